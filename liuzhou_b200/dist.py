"""Multi-GPU plumbing for self-play: one process per GPU (torch.distributed; NCCL over NVLink on the box, gloo in
the CPU tests).  Games are independent, so they shard across ranks with NO per-simulation communication; the only
collectives are the ones the reference does through files:

  * weights to every actor  -- reference: GPU -> CPU -> torch.save -> each worker torch.load
    (v1/train.py:966-979, self_play_worker.py:321-338)              -> here: one flat-buffer broadcast;
  * trajectories to the trainer -- reference: .pt shard files + manifests merged by the parent
    (self_play_worker.py:464-537, v1/train.py:1056-1153)             -> here: all_gather of row counts + gather of the
    five trajectory tensors (2,692 B / position) straight between device memories;
  * statistics -- reference: per-worker dicts summed by the parent   -> here: one all_reduce(SUM).
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

from .trajectory_buffer import TensorSelfPlayBatch


def split_games(total: int, n_ranks: int) -> List[int]:
    """v1/train.py:129-135 `_split_games`: base + 1 for the first `total % n` ranks."""
    if n_ranks <= 0:
        raise ValueError("n_ranks must be positive")
    base, rem = divmod(int(total), int(n_ranks))
    return [base + (1 if r < rem else 0) for r in range(n_ranks)]


def rank_seed(iteration_seed: int, rank: int) -> int:
    """Per-worker seed rule of the reference: iteration_seed * 10007 + (worker_idx + 1) * 9973 (v1/train.py:795,998)."""
    return int(iteration_seed) * 10007 + (int(rank) + 1) * 9973


def _world(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def _rank(group=None) -> int:
    return dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0


def broadcast_model(model: torch.nn.Module, src: int = 0, group=None) -> int:
    """Broadcast all parameters and buffers from `src` as ONE flat buffer per dtype. Returns bytes moved."""
    if _world(group) == 1:
        return 0
    tensors = [p.data for p in model.parameters()] + [b.data for b in model.buffers()]
    moved = 0
    by_dtype = {}
    for t in tensors:
        by_dtype.setdefault(t.dtype, []).append(t)
    for dtype, ts in by_dtype.items():
        flat = torch.cat([t.reshape(-1) for t in ts])
        dist.broadcast(flat, src=src, group=group)
        moved += flat.numel() * flat.element_size()
        off = 0
        for t in ts:
            n = t.numel()
            t.copy_(flat[off:off + n].view_as(t))
            off += n
    return moved


def gather_trajectories(batch: TensorSelfPlayBatch, dst: int = 0, group=None) -> Optional[TensorSelfPlayBatch]:
    """Variable-size gather of the five trajectory tensors to rank `dst`, rank-major order (== the order in which
    the reference's parent concatenates worker shards). Returns the merged batch on `dst`, None elsewhere."""
    world, rank = _world(group), _rank(group)
    if world == 1:
        return batch
    dev = batch.state_tensors.device
    n_local = torch.tensor([batch.num_samples], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c.item()) for c in counts]
    n_max = max(counts) if counts else 0
    merged = []
    for t in (batch.state_tensors, batch.legal_masks, batch.policy_targets, batch.value_targets,
              batch.soft_value_targets):
        carrier = t.to(torch.uint8) if t.dtype == torch.bool else t
        padded = torch.zeros((n_max,) + tuple(carrier.shape[1:]), dtype=carrier.dtype, device=dev)
        padded[: carrier.shape[0]].copy_(carrier)
        if rank == dst:
            bufs = [torch.empty_like(padded) for _ in range(world)]
            dist.gather(padded, bufs, dst=dst, group=group)
            out = torch.cat([b[:c] for b, c in zip(bufs, counts)], 0)
            merged.append(out.to(torch.bool) if t.dtype == torch.bool else out)
        else:
            dist.gather(padded, None, dst=dst, group=group)
    if rank != dst:
        return None
    return TensorSelfPlayBatch(*merged)


def _gather_rows(t: torch.Tensor, counts: List[int], dst: int, group) -> Optional[torch.Tensor]:
    """Variable-length gather along dim 0 (rows of rank r: counts[r]); returns the rank-major concatenation on dst."""
    world, rank = _world(group), _rank(group)
    n_max = max(counts) if counts else 0
    carrier = t.to(torch.uint8) if t.dtype == torch.bool else t
    padded = torch.zeros((n_max,) + tuple(carrier.shape[1:]), dtype=carrier.dtype, device=carrier.device)
    padded[: carrier.shape[0]].copy_(carrier)
    if rank == dst:
        bufs = [torch.empty_like(padded) for _ in range(world)]
        dist.gather(padded, bufs, dst=dst, group=group)
        out = torch.cat([b[:c] for b, c in zip(bufs, counts)], 0)
        return out.to(torch.bool) if t.dtype == torch.bool else out
    dist.gather(padded, None, dst=dst, group=group)
    return None


def gather_trajectories_compact(batch: TensorSelfPlayBatch, dst: int = 0, group=None) -> Optional[TensorSelfPlayBatch]:
    """Same result as ``gather_trajectories`` with ~15x fewer bytes on the wire: every rank compacts its batch
    (liuzhou_b200.compact: bitboards + legal bits + sparse policy, lossless), the compact pieces are gathered, and
    rank `dst` expands the merged batch back to the reference's five dense tensors."""
    from . import compact as cp

    world, rank = _world(group), _rank(group)
    if world == 1:
        return batch
    dev = batch.state_tensors.device
    c = cp.compact(batch)
    sizes = torch.tensor([c.num_samples, int(c.policy_index.numel())], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    n_rows = [int(s[0].item()) for s in all_sizes]
    n_nnz = [int(s[1].item()) for s in all_sizes]
    per_row_counts = c.policy_offsets[1:] - c.policy_offsets[:-1]           # offsets are rebuilt at the destination
    parts = [_gather_rows(t, n_rows, dst, group) for t in (c.boards, c.legal_bits, per_row_counts, c.value_targets,
                                                            c.soft_value_targets)]
    idx = _gather_rows(c.policy_index, n_nnz, dst, group)
    val = _gather_rows(c.policy_value, n_nnz, dst, group)
    if rank != dst:
        return None
    boards, legal_bits, counts, vt, svt = parts
    offsets = torch.zeros((counts.numel() + 1,), dtype=torch.int64, device=dev)
    offsets[1:] = torch.cumsum(counts, 0)
    return cp.expand(cp.CompactSelfPlayBatch(boards, legal_bits, offsets, idx, val, vt, svt))


def gather_rows_fixed(rows: torch.Tensor, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """Streaming gather of fixed-size compact rows (``compact.compact_rows_fixed``: int64[n,42], the same n on every
    rank -- one ply of every game): ONE collective with static shapes and no host synchronisation, so it can be queued
    on a side stream every ply.  Returns int64[world * n, 42] in rank-major order on ``dst``, None elsewhere."""
    world, rank = _world(group), _rank(group)
    if world == 1:
        return rows
    rows = rows.contiguous()
    if rank == dst:
        out = torch.empty((world,) + tuple(rows.shape), dtype=rows.dtype, device=rows.device)
        dist.gather(rows, list(out.unbind(0)), dst=dst, group=group)
        return out.view(world * rows.shape[0], rows.shape[1])
    dist.gather(rows, None, dst=dst, group=group)
    return None


def all_reduce_stats(values: List[float], device=None, group=None) -> List[float]:
    """Sum a small vector of counters (W/L/D, positions, lengths, seconds) over ranks."""
    if _world(group) == 1:
        return [float(v) for v in values]
    t = torch.tensor(values, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return [float(v) for v in t.tolist()]


def self_play_sharded(model, total_games: int, *, iteration_seed: int, device, group=None, compact_gather: bool = True,
                      timings: Optional[dict] = None, **self_play_kwargs):
    """Config-4 style entry: every rank plays its share of `total_games` (split_games) with the per-rank seed
    rule, weights come from rank 0, trajectories and statistics return to rank 0.
    Returns (merged TensorSelfPlayBatch or None, summed stats dict).  ``timings`` (optional dict) receives this rank's
    wall-clock seconds of the three stages (broadcast_s, self_play_s, handoff_s), each closed by a device synchronise."""
    import time

    from .self_play import self_play_v1_gpu

    def mark(key, t_prev):
        if timings is None:
            return t_prev
        if torch.device(device).type == "cuda":
            torch.cuda.synchronize(device)
        now = time.perf_counter()
        timings[key] = now - t_prev
        return now

    world, rank = _world(group), _rank(group)
    games = split_games(total_games, world)[rank]
    t_mark = time.perf_counter()
    model = model.to(device)
    broadcast_model(model, src=0, group=group)
    t_mark = mark("broadcast_s", t_mark)
    torch.manual_seed(rank_seed(iteration_seed, rank))
    if games > 0:
        batch, stats = self_play_v1_gpu(model, num_games=games, device=str(device), **self_play_kwargs)
        vec = [stats.num_games, stats.num_positions, stats.black_wins, stats.white_wins, stats.draws,
               stats.avg_game_length * stats.num_games, stats.elapsed_sec]
    else:
        dev = torch.device(device)
        batch = TensorSelfPlayBatch(torch.empty((0, 11, 6, 6), device=dev), torch.empty((0, 220), dtype=torch.bool, device=dev),
                                    torch.empty((0, 220), device=dev), torch.empty((0,), device=dev),
                                    torch.empty((0,), device=dev))
        vec = [0.0] * 7
    t_mark = mark("self_play_s", t_mark)
    merged = (gather_trajectories_compact if compact_gather else gather_trajectories)(batch, dst=0, group=group)
    tot = all_reduce_stats(vec, device=device, group=group)
    mark("handoff_s", t_mark)
    summary = {"num_games": tot[0], "num_positions": tot[1], "black_wins": tot[2], "white_wins": tot[3],
               "draws": tot[4], "avg_game_length": tot[5] / max(1.0, tot[0]), "sum_elapsed_sec": tot[6]}
    return merged, summary
