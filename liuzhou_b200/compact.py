"""Compact, lossless wire / disk form of a ``TensorSelfPlayBatch``.

The reference moves trajectories as five dense tensors, 2,692 B per position (v1/python/trajectory_buffer.py:11-33;
shard files of v1/python/self_play_worker.py:467-537).  Every byte of that is recoverable from far less:

    planes f32[11,6,6]   ->  4 x 36-bit boards (self, opponent, self marks, opponent marks) + the phase (1-7) packed
                             into 4 x int64 (32 B): planes 0-3 are 0/1 bitboards, planes 4-10 a one-hot phase plane
                             (v0/src/net/encoding.cpp:26-79)
    legal mask bool[220] ->  220 bits in 4 x int64 (32 B)
    policy f32[220]      ->  CSR over the non-zero entries: uint8 action index + f32 value (about 5 B x #legal)
    value / soft value   ->  2 x f32 kept as they are (NaN = not finalised yet survives)

About 170-200 B per position instead of 2,692 B (~15x less NCCL gather / shard-file traffic), and ``expand`` gives
back bit-identical tensors.  Pure tensor ops: works on whatever device the batch lives on (the gather path calls it on
the GPU; the tests run it on the CPU).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from .trajectory_buffer import TensorSelfPlayBatch

ACTION_DIM = 220


@dataclass
class CompactSelfPlayBatch:
    boards: torch.Tensor          # int64[n,4]: bits 0-35 of word k = plane k; bits 36-38 of word 0 = phase (1-7)
    legal_bits: torch.Tensor      # int64[n,4]: bit (a % 64) of word a // 64 = legal_masks[:, a]
    policy_offsets: torch.Tensor  # int64[n+1]: CSR row pointers
    policy_index: torch.Tensor    # uint8[nnz]
    policy_value: torch.Tensor    # f32[nnz]
    value_targets: torch.Tensor   # f32[n]
    soft_value_targets: torch.Tensor  # f32[n]

    @property
    def num_samples(self) -> int:
        return int(self.boards.shape[0])

    def tensors(self):
        return (self.boards, self.legal_bits, self.policy_offsets, self.policy_index, self.policy_value,
                self.value_targets, self.soft_value_targets)

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.tensors())

    def to(self, device) -> "CompactSelfPlayBatch":
        return CompactSelfPlayBatch(*(t.to(device) for t in self.tensors()))


def _pack_bits(mask: torch.Tensor, words: int) -> torch.Tensor:
    """bool[n, m] -> int64[n, words] (bit i % 64 of word i // 64)."""
    n, m = mask.shape
    pad = words * 64 - m
    if pad:
        mask = torch.cat([mask, torch.zeros((n, pad), dtype=torch.bool, device=mask.device)], 1)
    shifts = torch.arange(64, device=mask.device, dtype=torch.int64)
    return (mask.view(n, words, 64).to(torch.int64) << shifts).sum(dim=2)      # disjoint bits: the sum is an OR


def _unpack_bits(words: torch.Tensor, m: int) -> torch.Tensor:
    n, w = words.shape
    shifts = torch.arange(64, device=words.device, dtype=torch.int64)
    return (((words.unsqueeze(-1) >> shifts) & 1).reshape(n, w * 64)[:, :m]).to(torch.bool)


def compact(batch: TensorSelfPlayBatch) -> CompactSelfPlayBatch:
    """Lossless compaction; raises ValueError if the planes are not the 0/1 + one-hot-phase encoding."""
    x = batch.state_tensors
    n = int(x.shape[0])
    dev = x.device
    if tuple(x.shape[1:]) != (11, 6, 6):
        raise ValueError(f"state_tensors must be [n,11,6,6], got {tuple(x.shape)}")
    flat = x.reshape(n, 11, 36)
    if n and not bool(((flat == 0) | (flat == 1)).all()):
        raise ValueError("state planes must be exactly 0 / 1")
    phase_planes = flat[:, 4:, :]
    phase_on = phase_planes[:, :, 0] == 1                                   # [n,7]
    if n and not bool((phase_planes == phase_planes[:, :, :1]).all() & (phase_on.sum(1) == 1).all()):
        raise ValueError("planes 4-10 must be a one-hot, board-constant phase encoding")
    phase = phase_on.to(torch.int64).argmax(dim=1) + 1 if n else torch.zeros((0,), dtype=torch.int64, device=dev)
    boards = torch.stack([_pack_bits(flat[:, k, :] == 1, 1)[:, 0] for k in range(4)], dim=1) if n else \
        torch.zeros((0, 4), dtype=torch.int64, device=dev)
    if n:
        boards[:, 0] |= phase << 36
    legal_bits = _pack_bits(batch.legal_masks.to(torch.bool), 4)
    pol = batch.policy_targets
    nz = pol != 0
    counts = nz.sum(dim=1)
    offsets = torch.zeros((n + 1,), dtype=torch.int64, device=dev)
    offsets[1:] = torch.cumsum(counts, 0)
    rows, cols = torch.nonzero(nz, as_tuple=True)                            # row-major order == CSR order
    return CompactSelfPlayBatch(boards, legal_bits, offsets, cols.to(torch.uint8), pol[rows, cols].contiguous(),
                                batch.value_targets.clone(), batch.soft_value_targets.clone())


def expand(c: CompactSelfPlayBatch) -> TensorSelfPlayBatch:
    """Inverse of ``compact``: the reference's five dense tensors, bit-identical."""
    n = c.num_samples
    dev = c.boards.device
    planes = torch.zeros((n, 11, 36), dtype=torch.float32, device=dev)
    mask36 = (1 << 36) - 1
    for k in range(4):
        planes[:, k, :] = _unpack_bits((c.boards[:, k] & mask36).view(n, 1), 36).to(torch.float32)
    phase = (c.boards[:, 0] >> 36) & 7
    if n:
        planes[torch.arange(n, device=dev), 3 + phase, :] = 1.0
    legal = _unpack_bits(c.legal_bits, ACTION_DIM)
    policy = torch.zeros((n, ACTION_DIM), dtype=torch.float32, device=dev)
    counts = c.policy_offsets[1:] - c.policy_offsets[:-1]
    rows = torch.repeat_interleave(torch.arange(n, device=dev), counts)
    policy[rows, c.policy_index.to(torch.int64)] = c.policy_value
    return TensorSelfPlayBatch(planes.view(n, 11, 6, 6), legal, policy, c.value_targets.clone(),
                               c.soft_value_targets.clone())


def concat(parts) -> CompactSelfPlayBatch:
    """Concatenate compact batches (rank-major merge at the trainer)."""
    parts = list(parts)
    if not parts:
        raise ValueError("nothing to concatenate")
    offs, base = [parts[0].policy_offsets[:1]], 0
    for p in parts:
        offs.append(p.policy_offsets[1:] + base)
        base += int(p.policy_offsets[-1])
    return CompactSelfPlayBatch(torch.cat([p.boards for p in parts]), torch.cat([p.legal_bits for p in parts]),
                                torch.cat(offs), torch.cat([p.policy_index for p in parts]),
                                torch.cat([p.policy_value for p in parts]),
                                torch.cat([p.value_targets for p in parts]),
                                torch.cat([p.soft_value_targets for p in parts]))


# ----------------------------------------------------------------------------------------------------------------------
# Fixed-size rows for streaming (one ply of every game, every ply): no data-dependent shapes, hence no host
# synchronisation anywhere on the path -- compaction, the NCCL gather and the expansion can all be queued on a side
# stream while the next ply is searched.
#   words [0,4)   boards + phase (as above)
#   words [4,8)   legal bits
#   words [8,40)  64 x f32: the policy values of the legal actions in ascending action order (a position has at most
#                 64 legal actions: <= 36 placements; moves <= 4 * min(own pieces, empty cells) with own + empty <= 32
#                 because the opponent keeps >= 4 pieces; selections <= 18), zero beyond the legal count
#   word  40      value target | soft value target (2 x f32, NaN preserved)
#   word  41      flags: bit 0 = row not representable (policy mass on an illegal action or > 64 legal actions)
# 336 B per position instead of 2,692 B.
# ----------------------------------------------------------------------------------------------------------------------
FIXED_ROW_WORDS = 42
_FIXED_SLOTS = 64


def _legal_order(legal: torch.Tensor) -> torch.Tensor:
    """int64[n,64]: indices of the legal actions in ascending order first (then illegal ones, ascending)."""
    key = (~legal).to(torch.int16) * ACTION_DIM + torch.arange(ACTION_DIM, device=legal.device, dtype=torch.int16)
    return torch.argsort(key, dim=1)[:, :_FIXED_SLOTS]


def compact_rows_fixed(batch: TensorSelfPlayBatch) -> torch.Tensor:
    """-> int64[n, FIXED_ROW_WORDS]; never synchronises with the host (validity travels in the flags word)."""
    x = batch.state_tensors
    n = int(x.shape[0])
    dev = x.device
    flat = x.reshape(n, 11, 36)
    legal = batch.legal_masks.to(torch.bool)
    pol = batch.policy_targets.to(torch.float32)
    out = torch.empty((n, FIXED_ROW_WORDS), dtype=torch.int64, device=dev)
    for k in range(4):
        out[:, k] = _pack_bits(flat[:, k, :] == 1, 1)[:, 0]
    phase_on = flat[:, 4:, 0] == 1
    out[:, 0] |= (phase_on.to(torch.int64).argmax(dim=1) + 1) << 36
    out[:, 4:8] = _pack_bits(legal, 4)
    order = _legal_order(legal)
    out[:, 8:40] = pol.gather(1, order).contiguous().view(torch.int64)
    out[:, 40] = torch.stack([batch.value_targets.to(torch.float32), batch.soft_value_targets.to(torch.float32)],
                             dim=1).contiguous().view(torch.int64)[:, 0]
    planes_ok = ((flat[:, :4] == 0) | (flat[:, :4] == 1)).all(dim=2).all(dim=1) & (phase_on.sum(1) == 1) & \
        (flat[:, 4:, :] == flat[:, 4:, :1]).all(dim=2).all(dim=1)
    bad = ((pol != 0) & ~legal).any(dim=1) | (legal.sum(dim=1) > _FIXED_SLOTS) | ~planes_ok
    out[:, 41] = bad.to(torch.int64)
    return out


def expand_rows_fixed(rows: torch.Tensor, check: bool = True) -> TensorSelfPlayBatch:
    """Inverse of ``compact_rows_fixed`` (bit-identical tensors). ``check`` reads the flags (one host sync)."""
    n = int(rows.shape[0])
    dev = rows.device
    if tuple(rows.shape[1:]) != (FIXED_ROW_WORDS,) or rows.dtype != torch.int64:
        raise ValueError(f"rows must be int64[n,{FIXED_ROW_WORDS}]")
    if check and n and bool((rows[:, 41] != 0).any()):
        raise ValueError("fixed-row block holds rows that were not representable (flags word set)")
    planes = torch.zeros((n, 11, 36), dtype=torch.float32, device=dev)
    mask36 = (1 << 36) - 1
    for k in range(4):
        planes[:, k, :] = _unpack_bits((rows[:, k] & mask36).view(n, 1), 36).to(torch.float32)
    phase = (rows[:, 0] >> 36) & 7
    if n:
        planes[torch.arange(n, device=dev), 3 + phase, :] = 1.0
    legal = _unpack_bits(rows[:, 4:8], ACTION_DIM)
    vals = rows[:, 8:40].contiguous().view(torch.float32)                    # [n,64]
    policy = torch.zeros((n, ACTION_DIM), dtype=torch.float32, device=dev)
    policy.scatter_(1, _legal_order(legal), vals)
    vt = rows[:, 40:41].contiguous().view(torch.float32)                     # [n,2]
    return TensorSelfPlayBatch(planes.view(n, 11, 6, 6), legal, policy, vt[:, 0].clone(), vt[:, 1].clone())
