#!/usr/bin/env python3
"""bench.py -- headline benchmark of the liuzhou_b200 hot path (contract: see README / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload selfplay|playout] [--impl ours|reference]

N > 1 is launched by the driver through torchrun (one rank per GPU, NCCL); games are independent, so ranks
shard the games with no data-path collective ("scaling": "weak").  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

SEED = 20260314  # reference stable-init seed (scripts/big_train_v1.sh:24)


# ----------------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------------
def measured_peaks() -> tuple[dict, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return json.loads(p.read_text()), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int = 0):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines: list[str] = []
        self._thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.lines.append(line.strip())
        self._thread = threading.Thread(target=pump, daemon=True)
        self._thread.start()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2])); power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def dist_setup(n_gpus: int):
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(0)
        local_rank = 0
    return world, rank, local_rank


def barrier_sync(world: int):
    import torch

    if world > 1:
        import torch.distributed as dist

        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(value: float, world: int) -> float:
    if world <= 1:
        return value
    import torch
    import torch.distributed as dist

    t = torch.tensor([value], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, world: int) -> float:
    if world <= 1:
        return value
    import torch
    import torch.distributed as dist

    t = torch.tensor([value], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


class L2Flusher:
    """Writes a buffer larger than the 126 MB L2 between timed iterations."""

    def __init__(self, device, mib: int = 256):
        import torch

        self.buf = torch.empty((mib << 20,), dtype=torch.uint8, device=device)

    def flush(self):
        self.buf.add_(1)


# ----------------------------------------------------------------------------------------------------
# workload: config 2 -- rule-engine-only random playouts, 65,536 concurrent games per GPU
# ----------------------------------------------------------------------------------------------------
PLAYOUT_GAMES = 65_536
PLAYOUT_BYTES_PER_PLY = 72   # 32 B state read + 32 B write + 8 B RNG / cursor


def cpu_playout_baseline(budget_s: float = 12.0, threads: int | None = None) -> dict:
    """Oracle port (plain C restatement of the v0 C++ scalar engine loop) on the host cores; bounded sample."""
    import oracle
    from concurrent.futures import ThreadPoolExecutor

    cores = threads or os.cpu_count() or 1
    oracle.lib()
    chunk = 256
    t0 = time.perf_counter()
    plies = games = 0
    next_game = 0
    with ThreadPoolExecutor(max_workers=cores) as pool:
        while time.perf_counter() - t0 < budget_s and next_game < PLAYOUT_GAMES:
            futs = [pool.submit(oracle.random_playouts_range, SEED, next_game + i * chunk, next_game + (i + 1) * chunk)
                    for i in range(cores)]
            next_game += cores * chunk
            for f in futs:
                r = f.result()
                plies += r["plies"]
                games += chunk
    dt = time.perf_counter() - t0
    return {"value": plies / dt, "unit": "positions/s", "cores": cores, "kind": "port",
            "sample": f"{games} of {PLAYOUT_GAMES} games ({plies} plies) of the same seeded workload in {dt:.1f}s, "
                      f"oracle/lz_oracle.c scalar-engine loop, {cores} threads"}


def run_playout(args, world, rank, local_rank):
    import torch

    from liuzhou_b200 import _lib, native

    dev = torch.device("cuda", local_rank)
    peaks, peak_kind = measured_peaks()
    games = PLAYOUT_GAMES
    pb = native.PlayoutBatch(games, seed=SEED, device=dev, game_offset=rank * games)
    flusher = L2Flusher(dev)
    stream = torch.cuda.current_stream(dev)

    def step():
        pb.reset()
        pb.run()

    for _ in range(args.warmup):
        step()
        flusher.flush()
    barrier_sync(world)
    launches0 = _lib.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    kernel_ms = []
    total_plies = 0
    t_step = []
    for _ in range(args.steps):
        flusher.flush()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(stream)
        pb.reset()
        e1.record(stream)
        pb.run()
        e2.record(stream)
        e2.synchronize()
        t_step.append(e0.elapsed_time(e2))
        kernel_ms.append(e1.elapsed_time(e2))
        total_plies += int(pb.plies.sum().item())
    barrier_sync(world)
    clocks = sampler.stop() if rank == 0 else {}
    launches = _lib.launch_count() - launches0
    elapsed_ms = max_over_ranks(sum(t_step), world)
    all_plies = sum_over_ranks(float(total_plies), world)
    value = all_plies / (elapsed_ms / 1e3)

    # e2e: the same pass through the public API with HOST buffers: initial states (reference byte layout) are
    # copied from pinned host memory, packed, played, unpacked and the final states + results copied back.
    st_host = [t.cpu().pin_memory() for t in native.unpack_states(native.init_states(games, dev))]
    out_host = [torch.empty_like(t).pin_memory() for t in st_host]
    res_host = torch.empty((games,), dtype=torch.int8).pin_memory()
    ply_host = torch.empty((games,), dtype=torch.int32).pin_memory()
    h2d = sum(t.numel() * t.element_size() for t in st_host)
    d2h = h2d + res_host.numel() + ply_host.numel() * 4
    e2e_ms, e2e_plies = [], 0
    for it in range(args.warmup + args.steps):
        flusher.flush()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dev_st = [t.to(dev, non_blocking=True) for t in st_host]
        pb.packed = native.pack_states(dev_st)
        pb.plies.zero_(); pb.result.fill_(2)
        pb.run()
        for dst, src in zip(out_host, native.unpack_states(pb.packed)):
            dst.copy_(src, non_blocking=True)
        res_host.copy_(pb.result, non_blocking=True)
        ply_host.copy_(pb.plies, non_blocking=True)
        torch.cuda.synchronize()
        if it >= args.warmup:
            e2e_ms.append((time.perf_counter() - t0) * 1e3)
            e2e_plies += int(ply_host.sum())
    e2e_elapsed = max_over_ranks(sum(e2e_ms), world)
    e2e_value = sum_over_ranks(float(e2e_plies), world) / (e2e_elapsed / 1e3)

    if rank != 0:
        return None
    k_ms = sum(kernel_ms) / len(kernel_ms)
    plies_per_launch = total_plies / args.steps
    achieved = plies_per_launch * PLAYOUT_BYTES_PER_PLY / (k_ms / 1e3) / 1e9
    cpu = cpu_playout_baseline()
    return {
        "metric": "selfplay_positions_per_sec", "value": value, "unit": "positions/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "rule_engine_random_playouts (BASELINE configs[1])", "games_per_gpu": games,
                   "policy": "uniform random (counter-based RNG), 0 MCTS sims", "max_game_plies": 512,
                   "seed": SEED, "l2": "256 MiB buffer rewritten between timed iterations"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "positions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": achieved / peaks["hbm_gbs"], "traffic": None, "peak_kind": peak_kind,
                     "kernel": "playout_kernel", "kernel_ms": k_ms,
                     "algorithmic_bytes_per_unit": PLAYOUT_BYTES_PER_PLY, "units_per_launch": plies_per_launch},
        "cpu_baseline": cpu,
    }


def run_reference_playout(args):
    """--impl reference for the playout workload: the CPU engine on all host threads, bounded sample per step."""
    per_step = max(2.0, min(10.0, 60.0 / max(1, args.steps + args.warmup)))
    vals = []
    last = None
    for it in range(args.warmup + args.steps):
        last = cpu_playout_baseline(budget_s=per_step)
        if it >= args.warmup:
            vals.append(last["value"])
    value = sum(vals) / len(vals)
    last["value"] = value
    return {
        "impl": "reference", "metric": "selfplay_positions_per_sec", "value": value, "unit": "positions/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i8", "data": "synthetic",
        "config": {"workload": "rule_engine_random_playouts (BASELINE configs[1])", "games_per_gpu": PLAYOUT_GAMES,
                   "seed": SEED},
        "cpu_baseline": last,
        "e2e": {"value": value, "unit": "positions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=["playout", "selfplay"], default="playout")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return 0
        print(json.dumps(run_reference_playout(args)), flush=True)
        return 0

    world, rank, local_rank = dist_setup(args.gpus)
    line = run_playout(args, world, rank, local_rank)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
