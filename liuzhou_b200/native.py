"""Native (packed-bitboard) layer: game states live in HBM as 4 x uint64 per game (32 B), see
include/liuzhou_b200.h.  Packed batches are carried as ``torch.int64[B, 4]`` CUDA tensors.

These are the conversions at the boundary with the reference's tensor layout
(v0/src/game/tensor_state_batch.cpp) plus the rule-engine ops and the config-2 random-playout workload
directly on the packed layout.
"""
from __future__ import annotations

import ctypes

import torch

from ._lib import STATE_FIELDS, check, i64, lib, ptr, require_cuda, states_view, stream_ptr


def _dev(device) -> torch.device:
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("liuzhou_b200.native: CUDA device required (no CPU path)")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def init_states(batch_size: int, device="cuda") -> torch.Tensor:
    """Initial positions (GpuStateBatch.initial, v1/python/mcts_gpu.py:123-145) in packed form."""
    dev = _dev(device)
    with torch.cuda.device(dev):
        packed = torch.empty((int(batch_size), 4), dtype=torch.int64, device=dev)
        check(lib().lzb_init_states(ptr(packed), i64(batch_size), stream_ptr(dev)))
    return packed


def pack_states(tensors) -> torch.Tensor:
    """12 reference-layout tensors (GpuStateBatch field order) -> int64[B,4]."""
    tensors = list(tensors)
    require_cuda(tensors[0], "board")
    dev = tensors[0].device
    t = [tensors[0].to(torch.int8).contiguous(), tensors[1].to(torch.bool).contiguous(),
         tensors[2].to(torch.bool).contiguous()] + [x.to(torch.int64).contiguous() for x in tensors[3:12]]
    b = t[0].size(0)
    with torch.cuda.device(dev):
        packed = torch.empty((b, 4), dtype=torch.int64, device=dev)
        view = states_view(t)
        check(lib().lzb_pack_states(ctypes.byref(view), i64(b), ptr(packed), stream_ptr(dev)))
    return packed


def unpack_states(packed: torch.Tensor):
    """int64[B,4] -> 12 reference-layout tensors (board int8[B,6,6], marks bool[B,6,6], 9 x int64[B])."""
    require_cuda(packed, "packed")
    packed = packed.contiguous()
    dev = packed.device
    b = packed.size(0)
    with torch.cuda.device(dev):
        out = [torch.empty((b, 6, 6), dtype=torch.int8, device=dev),
               torch.empty((b, 6, 6), dtype=torch.bool, device=dev),
               torch.empty((b, 6, 6), dtype=torch.bool, device=dev)]
        out += [torch.empty((b,), dtype=torch.int64, device=dev) for _ in range(9)]
        view = states_view(out)
        check(lib().lzb_unpack_states(ptr(packed), i64(b), ctypes.byref(view), stream_ptr(dev)))
    return tuple(out)


def legal_masks(packed: torch.Tensor, scalar_semantics: bool = True):
    """-> (mask_words int64[B,4] (bit a of word a//64), counts int32[B])."""
    require_cuda(packed, "packed")
    packed = packed.contiguous()
    dev = packed.device
    b = packed.size(0)
    with torch.cuda.device(dev):
        words = torch.empty((b, 4), dtype=torch.int64, device=dev)
        counts = torch.empty((b,), dtype=torch.int32, device=dev)
        check(lib().lzb_legal_masks_packed(ptr(packed), i64(b), ctypes.c_int(1 if scalar_semantics else 0),
                                           ptr(words), ptr(counts), stream_ptr(dev)))
    return words, counts


def mask_words_to_bool(words: torch.Tensor) -> torch.Tensor:
    """int64[B,4] -> bool[B,220] (host-side convenience for tests)."""
    shifts = torch.arange(64, device=words.device, dtype=torch.int64)
    bits = (words.unsqueeze(-1) >> shifts) & 1
    return bits.reshape(words.size(0), 256)[:, :220].to(torch.bool)


def apply_actions(packed: torch.Tensor, actions: torch.Tensor, parent_indices: torch.Tensor | None = None,
                  out: torch.Tensor | None = None) -> torch.Tensor:
    """children[i] = apply(packed[parent_indices[i]], actions[i]) with 220-d action indices."""
    require_cuda(packed, "packed")
    dev = packed.device
    packed = packed.contiguous()
    actions = actions.to(device=dev, dtype=torch.int32).contiguous()
    n = actions.numel()
    if parent_indices is not None:
        parent_indices = parent_indices.to(device=dev, dtype=torch.int64).contiguous()
        if parent_indices.numel() != n:
            raise RuntimeError("parent_indices must align with actions")
    elif packed.size(0) != n:
        raise RuntimeError("actions must align with states when parent_indices is None")
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty((n, 4), dtype=torch.int64, device=dev)
        check(lib().lzb_apply_actions_packed(ptr(packed), ptr(parent_indices), ptr(actions), i64(n), ptr(out),
                                             stream_ptr(dev)))
    return out


class PlayoutBatch:
    """Config-2 workload state: B concurrent uniform-random games on the packed layout."""

    def __init__(self, batch_size: int, seed: int, device="cuda", game_offset: int = 0, track_hash: bool = False):
        self.device = _dev(device)
        self.batch_size = int(batch_size)
        self.seed = int(seed)
        self.game_offset = int(game_offset)
        self.packed = init_states(batch_size, self.device)
        self.plies = torch.zeros((batch_size,), dtype=torch.int32, device=self.device)
        self.result = torch.full((batch_size,), 2, dtype=torch.int8, device=self.device)
        self.hash = torch.zeros((batch_size,), dtype=torch.int64, device=self.device) if track_hash else None

    def reset(self) -> None:
        with torch.cuda.device(self.device):
            check(lib().lzb_init_states(ptr(self.packed), i64(self.batch_size), stream_ptr(self.device)))
        self.plies.zero_()
        self.result.fill_(2)
        if self.hash is not None:
            self.hash.zero_()

    def run(self, max_steps: int = 1 << 20, max_game_plies: int = 512) -> None:
        """Advance every unfinished game by up to max_steps plies in ONE kernel launch."""
        with torch.cuda.device(self.device):
            check(lib().lzb_playout_run(ptr(self.packed), ptr(self.plies), ptr(self.result), ptr(self.hash),
                                        i64(self.batch_size), ctypes.c_uint64(self.seed & (2**64 - 1)),
                                        ctypes.c_uint64(self.game_offset), ctypes.c_int32(int(min(max_steps, 2**31 - 1))),
                                        ctypes.c_int32(int(max_game_plies)), stream_ptr(self.device)))


__all__ = ["STATE_FIELDS", "init_states", "pack_states", "unpack_states", "legal_masks", "mask_words_to_bool",
           "apply_actions", "PlayoutBatch"]
