"""Multi-rank host logic on CPU: world_size-2 `gloo` process group exercising the sharding rules, the flat weight
broadcast, the variable-size trajectory gather and the statistics all-reduce (the N > 1 path of liuzhou_b200.dist;
NCCL replaces gloo on the GPU box, the code path is the same)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from liuzhou_b200.dist import (all_reduce_stats, broadcast_model, gather_trajectories, gather_trajectories_compact,
                               rank_seed, split_games)
from liuzhou_b200.trajectory_buffer import TensorSelfPlayBatch


def test_split_games_and_seeds_match_reference_rules():
    assert split_games(32768, 8) == [4096] * 8                 # BASELINE config 4
    assert split_games(10, 4) == [3, 3, 2, 2]                  # base + 1 for the first total % n ranks
    assert split_games(3, 4) == [1, 1, 1, 0]
    assert sum(split_games(522_488, 7)) == 522_488
    assert rank_seed(5, 0) == 5 * 10007 + 9973 and rank_seed(5, 3) == 5 * 10007 + 4 * 9973
    with pytest.raises(ValueError):
        split_games(4, 0)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_batch(n: int, rank: int, density: float = 0.3) -> TensorSelfPlayBatch:
    g = torch.Generator().manual_seed(100 + rank)
    planes = torch.zeros((n, 11, 6, 6))
    planes[:, :4] = (torch.rand((n, 4, 6, 6), generator=g) > 0.5).float()
    phase = torch.randint(1, 8, (n,), generator=g)
    planes[torch.arange(n), 3 + phase] = 1.0                  # one-hot phase plane (v0/src/net/encoding.cpp:26-79)
    legal = torch.rand((n, 220), generator=g) > 1.0 - density
    return TensorSelfPlayBatch(
        state_tensors=planes,
        legal_masks=legal,
        policy_targets=torch.rand((n, 220), generator=g) * legal,
        value_targets=torch.randint(-1, 2, (n,), generator=g).float(),
        soft_value_targets=torch.rand((n,), generator=g) * 2 - 1)


def _worker(rank: int, world: int, port: int, sizes, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from liuzhou_b200.net import ChessNet

        torch.manual_seed(1000 + rank)                          # different weights per rank before the broadcast
        model = ChessNet(trunk_channels=8, num_blocks=1, policy_channels=4, value_channels=4, value_mlp_channels=8)
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.fill_(float(rank))
        moved = broadcast_model(model, src=0)
        flat = torch.cat([p.detach().reshape(-1).double() for p in model.parameters()] +
                         [b.detach().reshape(-1).double() for b in model.buffers()])
        ref = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(ref, flat)
        same = all(torch.equal(ref[0], r) for r in ref)
        batch = _make_batch(sizes[rank], rank)
        merged = gather_trajectories(batch, dst=0)
        merged_c = gather_trajectories_compact(batch, dst=0)
        stats = all_reduce_stats([1.0, float(sizes[rank]), float(rank)])
        # streaming form: fixed-size rows, the same row count on every rank, one collective, no host sync
        from liuzhou_b200 import compact as cp
        from liuzhou_b200.dist import gather_rows_fixed

        fixed_in = _make_batch(6, 50 + rank, density=0.2)
        fixed_rows = gather_rows_fixed(cp.compact_rows_fixed(fixed_in), dst=0)
        if rank == 0:
            fb = cp.expand_rows_fixed(fixed_rows)
            fexp = [_make_batch(6, 50 + r, density=0.2) for r in range(world)]
            fixed_ok = fb.num_samples == 6 * world
            for name in ("state_tensors", "legal_masks", "policy_targets", "value_targets", "soft_value_targets"):
                cat = torch.cat([getattr(e, name) for e in fexp], 0)
                fixed_ok = fixed_ok and torch.equal(getattr(fb, name), cat) and getattr(fb, name).dtype == cat.dtype
        else:
            fixed_ok = fixed_rows is None
        if rank == 0:
            expect = [_make_batch(sizes[r], r) for r in range(world)]
            ok = merged is not None and merged.num_samples == sum(sizes)
            for name in ("state_tensors", "legal_masks", "policy_targets", "value_targets", "soft_value_targets"):
                cat = torch.cat([getattr(e, name) for e in expect], 0)
                ok = ok and torch.equal(getattr(merged, name), cat) and getattr(merged, name).dtype == cat.dtype
                got = getattr(merged_c, name)                       # compact wire format: same tensors, bit for bit
                ok = ok and got.dtype == cat.dtype and torch.equal(torch.nan_to_num(got.float(), nan=7.0),
                                                                   torch.nan_to_num(cat.float(), nan=7.0))
            ret["rank0"] = (same, ok and fixed_ok, stats, moved)
        else:
            ret[f"rank{rank}"] = (same, merged is None and merged_c is None and fixed_ok, stats, moved)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("sizes", [(5, 9), (0, 4), (7, 7)])
def test_broadcast_gather_allreduce_world2(sizes):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    procs = [mp.get_context("spawn").Process(target=_worker, args=(r, world, port, sizes, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    same0, ok0, stats0, moved0 = ret["rank0"]
    same1, none1, stats1, moved1 = ret["rank1"]
    assert same0 and same1, "weights differ after broadcast"
    assert ok0, "gathered trajectories differ from the rank-major concatenation"
    assert none1
    assert stats0 == stats1 == [2.0, float(sum(sizes)), 1.0]
    assert moved0 == moved1 and moved0 > 0
