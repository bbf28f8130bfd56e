"""Trajectory buffer in the reference's format (v1/python/trajectory_buffer.py:11-211): five training tensors
``state_tensors f32[n,11,6,6] | legal_masks bool[n,220] | policy_targets f32[n,220] | value_targets f32[n] |
soft_value_targets f32[n]`` (2,692 B / position), rows in append order (ply-major within a wave), value targets NaN
until the owning game is finalised with ``player_sign * result_from_black``.

Differences from the reference that do not change the format: the arena is sized once from
``max_steps_hint * concurrent_games_hint`` and grows geometrically without re-filling, and ``build()`` returns
views (no clone) unless ``clone=True``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from . import v0_core


@dataclass
class TensorSelfPlayBatch:
    state_tensors: torch.Tensor
    legal_masks: torch.Tensor
    policy_targets: torch.Tensor
    value_targets: torch.Tensor
    soft_value_targets: torch.Tensor

    @property
    def num_samples(self) -> int:
        return int(self.state_tensors.shape[0])

    def to(self, device) -> "TensorSelfPlayBatch":
        dev = torch.device(device)
        return TensorSelfPlayBatch(self.state_tensors.to(dev), self.legal_masks.to(dev), self.policy_targets.to(dev),
                                   self.value_targets.to(dev), self.soft_value_targets.to(dev))

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.state_tensors, self.legal_masks, self.policy_targets,
                                                            self.value_targets, self.soft_value_targets))


class TensorTrajectoryBuffer:
    def __init__(self, device, action_dim: int, *, max_steps_hint: int = 512, concurrent_games_hint: int = 8,
                 initial_capacity: Optional[int] = None) -> None:
        self.device = torch.device(device)
        self.action_dim = int(action_dim)
        hint = max(1, int(max_steps_hint) * int(concurrent_games_hint))
        self._capacity = int(max(initial_capacity or 0, hint))
        self._size = 0
        self._state_shape = None
        self._state_tensors = self._legal_masks = self._policy_targets = None
        self._value_targets = self._soft_value_targets = self._player_signs = None

    def _allocate(self, capacity: int, shape) -> None:
        dev = self.device
        self._state_shape = shape
        self._state_tensors = torch.empty((capacity, *shape), dtype=torch.float32, device=dev)
        self._legal_masks = torch.empty((capacity, self.action_dim), dtype=torch.bool, device=dev)
        self._policy_targets = torch.empty((capacity, self.action_dim), dtype=torch.float32, device=dev)
        self._value_targets = torch.full((capacity,), float("nan"), dtype=torch.float32, device=dev)
        self._soft_value_targets = torch.full((capacity,), float("nan"), dtype=torch.float32, device=dev)
        self._player_signs = torch.empty((capacity,), dtype=torch.int8, device=dev)
        self._capacity = int(capacity)

    def _grow(self, required: int) -> None:
        old = (self._state_tensors, self._legal_masks, self._policy_targets, self._value_targets,
               self._soft_value_targets, self._player_signs)
        n = self._size
        self._allocate(max(int(required), 2 * max(1, self._capacity)), self._state_shape)
        for dst, src in zip((self._state_tensors, self._legal_masks, self._policy_targets, self._value_targets,
                             self._soft_value_targets, self._player_signs), old):
            dst[:n].copy_(src[:n])

    def append_steps(self, model_input: torch.Tensor, legal_mask: torch.Tensor, policy_dense: torch.Tensor,
                     player_sign: torch.Tensor) -> torch.Tensor:
        if model_input.dim() != 4:
            raise ValueError(f"model_input must be (N,C,H,W), got shape {tuple(model_input.shape)}")
        n = int(model_input.shape[0])
        if tuple(legal_mask.shape) != (n, self.action_dim) or tuple(policy_dense.shape) != (n, self.action_dim):
            raise ValueError(f"legal_mask / policy_dense must be (N,{self.action_dim})")
        sign_t = torch.as_tensor(player_sign, device=model_input.device).view(-1)
        if int(sign_t.numel()) != n:
            raise ValueError(f"player_sign must have {n} elements")
        shape = tuple(int(s) for s in model_input.shape[1:])
        if self._state_tensors is None:
            self._allocate(max(self._capacity, n), shape)
        elif self._state_shape != shape:
            raise ValueError(f"Inconsistent state shape: expected {self._state_shape}, got {shape}")
        end = self._size + n
        if end > self._capacity:
            self._grow(end)
        s = self._size
        self._state_tensors[s:end].copy_(model_input.detach())
        self._legal_masks[s:end].copy_(legal_mask.detach())
        self._policy_targets[s:end].copy_(policy_dense.detach())
        self._value_targets[s:end].fill_(float("nan"))
        self._soft_value_targets[s:end].fill_(float("nan"))
        self._player_signs[s:end].copy_(torch.where(sign_t >= 0, 1, -1))
        self._size = end
        return torch.arange(s, end, dtype=torch.int64, device=self.device)

    def append_step(self, model_input, legal_mask, policy_dense, player_sign: int) -> int:
        idx = self.append_steps(model_input.unsqueeze(0), legal_mask.unsqueeze(0), policy_dense.unsqueeze(0),
                                torch.tensor([int(player_sign)], dtype=torch.int64, device=model_input.device))
        return int(idx[0].item())

    def finalize_games_inplace(self, *, step_index_matrix, step_counts, slots, result_from_black,
                               soft_value_from_black):
        if self._size == 0:
            empty = torch.empty((0,), dtype=torch.int64, device=self.device)
            return empty, empty, torch.zeros((3,), dtype=torch.int64, device=self.device)
        return v0_core.finalize_trajectory_inplace(self._value_targets, self._soft_value_targets, self._player_signs,
                                                   step_index_matrix, step_counts, slots, result_from_black,
                                                   soft_value_from_black)

    def build(self, clone: bool = True) -> TensorSelfPlayBatch:
        if self._size == 0:
            shape = self._state_shape or (11, 6, 6)
            dev = self.device
            return TensorSelfPlayBatch(torch.empty((0, *shape), dtype=torch.float32, device=dev),
                                       torch.empty((0, self.action_dim), dtype=torch.bool, device=dev),
                                       torch.empty((0, self.action_dim), dtype=torch.float32, device=dev),
                                       torch.empty((0,), dtype=torch.float32, device=dev),
                                       torch.empty((0,), dtype=torch.float32, device=dev))
        e = self._size
        f = (lambda t: t[:e].clone()) if clone else (lambda t: t[:e])
        return TensorSelfPlayBatch(f(self._state_tensors), f(self._legal_masks), f(self._policy_targets),
                                   f(self._value_targets), f(self._soft_value_targets))
