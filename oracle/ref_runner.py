#!/usr/bin/env python3
"""Run the reference's UNMODIFIED python host code (oracle/_ref/pysrc: v1/python/self_play_gpu_runner.py,
mcts_gpu.py, src/neural_network.py ...) in a process of its own.   TEST / BENCH INFRASTRUCTURE -- never imported by
liuzhou_b200/.

    python oracle/ref_runner.py selfplay --v0core {ref,shim} --games G --sims S [--device cuda:0] [--noise 0|1]
                                         [--sample 0|1] [--seed N] [--warmup-games W] [--dump out.pt]
    python oracle/ref_runner.py search   --v0core {ref,shim} --states in.pt --sims S [--dump out.pt]
    python oracle/ref_runner.py legacy   --games 1 --sims 64          (BASELINE configs[0]: legacy src/mcts.py self-play on CPU)
    python oracle/ref_runner.py portable --games 256 --sims 200 --max-plies 3   (the reference's CPU full-tree self-play entry)

`--v0core ref`  : `import v0_core` resolves to the reference's own extension (oracle/_ref/v0_core*.so, its three .cu kernels
                  compiled for sm_100 from the unmodified sources)      -> "the v1 reference on the same B200".
`--v0core shim` : sys.modules["v0_core"] = liuzhou_b200.v0_core BEFORE the reference modules are imported (the binding
                  INTEGRATION.md section 2a describes)                   -> the drop-in boundary proof.
The network is the reference's own `src.neural_network.ChessNet` (PyTorch eager, fp16 autocast as the reference runs it).
One JSON line on stdout; `--dump` saves the produced tensors for a bit-for-bit comparison between the two runs.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
REF_DIR = HERE / "_ref"
PYSRC = REF_DIR / "pysrc"


def _setup(v0core: str):
    if not PYSRC.is_dir():
        raise SystemExit(json.dumps({"unavailable": "oracle/_ref/pysrc missing (run oracle/build_ref.py where /root/reference is mounted)"}))
    sys.path.insert(0, str(PYSRC))
    import torch  # noqa: F401

    if v0core == "shim":
        sys.path.insert(0, str(ROOT))
        import liuzhou_b200.v0_core as shim

        sys.modules["v0_core"] = shim
    else:
        sys.path.insert(0, str(REF_DIR))
        import v0_core  # noqa: F401

        assert "_ref" in str(getattr(v0_core, "__file__", "")), "v0_core did not resolve to the reference build"


def _model(seed: int, device: str, small: bool = False):
    import torch
    from src.neural_network import ChessNet

    torch.manual_seed(seed)
    net = ChessNet(trunk_channels=8, num_blocks=1, policy_channels=4, value_channels=4, value_mlp_channels=8) if small \
        else ChessNet()
    return net.to(device).eval()


def cmd_selfplay(a) -> dict:
    _setup(a.v0core)
    import torch
    from v1.python.self_play_gpu_runner import self_play_v1_gpu

    model = _model(a.model_seed, a.device, a.small_net)
    kw = dict(mcts_simulations=a.sims, temperature_init=1.0, temperature_final=0.1, temperature_threshold=10,
              exploration_weight=1.0, device=a.device, add_dirichlet_noise=bool(a.noise), soft_value_k=2.0,
              max_game_plies=a.max_plies, sample_moves=bool(a.sample))
    if a.warmup_games > 0:                      # cuDNN autotune / allocator warm-up outside the timed iteration
        torch.manual_seed(a.seed + 1)
        self_play_v1_gpu(model, num_games=a.warmup_games, concurrent_games=a.warmup_games, **kw)
    torch.manual_seed(a.seed)
    cuda = a.device.startswith("cuda")
    if cuda:
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    batch, stats = self_play_v1_gpu(model, num_games=a.games, concurrent_games=a.concurrent or a.games, **kw)
    if cuda:
        torch.cuda.synchronize()
    t1 = time.perf_counter()
    host = batch.to("cpu")                       # e2e: the finished trajectory batch read back to the host
    t2 = time.perf_counter()
    if a.dump:
        torch.save({"state_tensors": host.state_tensors, "legal_masks": host.legal_masks,
                    "policy_targets": host.policy_targets, "value_targets": host.value_targets,
                    "soft_value_targets": host.soft_value_targets}, a.dump)
    n = int(host.num_samples)
    return {"mode": "selfplay", "v0core": a.v0core, "games": a.games, "sims": a.sims, "positions": n,
            "seconds": t1 - t0, "positions_per_sec": n / (t1 - t0), "e2e_seconds": t2 - t0,
            "e2e_positions_per_sec": n / (t2 - t0), "d2h_bytes": int(sum(t.numel() * t.element_size() for t in (
                host.state_tensors, host.legal_masks, host.policy_targets, host.value_targets, host.soft_value_targets))),
            "black_wins": int(stats.black_wins), "white_wins": int(stats.white_wins), "draws": int(stats.draws),
            "avg_game_length": float(stats.avg_game_length), "device": a.device,
            "net": "src.neural_network.ChessNet (reference module, PyTorch eager, fp16 autocast)"}


def cmd_search(a) -> dict:
    _setup(a.v0core)
    import torch
    from v1.python.mcts_gpu import GpuStateBatch, V1RootMCTS, V1RootMCTSConfig

    model = _model(a.model_seed, a.device, a.small_net)
    st = torch.load(a.states)
    fields = ("board", "marks_black", "marks_white", "phase", "current_player", "pending_marks_required",
              "pending_marks_remaining", "pending_captures_required", "pending_captures_remaining",
              "forced_removals_done", "move_count", "moves_since_capture")
    batch = GpuStateBatch(**{k: st[k].to(a.device) for k in fields})
    mcts = V1RootMCTS(model, V1RootMCTSConfig(num_simulations=a.sims, exploration_weight=1.0, temperature=1.0,
                                              add_dirichlet_noise=False, sample_moves=False), a.device)
    temps = torch.where(torch.arange(batch.batch_size, device=a.device) % 2 == 0, 1.0, 0.1).to(torch.float32)
    out = mcts.search_batch(batch, temperatures=temps, add_dirichlet_noise=False)
    if a.dump:
        torch.save({k: getattr(out, k).cpu() for k in ("model_input", "legal_mask", "policy_dense", "root_value",
                                                      "terminal_mask", "chosen_action_indices", "chosen_action_codes",
                                                      "chosen_valid_mask")}, a.dump)
    return {"mode": "search", "v0core": a.v0core, "roots": int(batch.batch_size), "sims": a.sims}


def cmd_legacy(a) -> dict:
    """BASELINE configs[0]: legacy `src/mcts.py::self_play` (python rule engine + python MCTS), 1 game, 64 sims/move, the
    reference tests' small random-init net, on the CPU (src/mcts.py:807-928)."""
    sys.path.insert(0, str(PYSRC))
    import torch
    from src.mcts import self_play
    from src.neural_network import ChessNet

    torch.set_num_threads(1)
    torch.manual_seed(7)
    model = ChessNet(trunk_channels=8, num_blocks=1, policy_channels=4, value_channels=4, value_mlp_channels=8).eval()
    t0 = time.perf_counter()
    data = self_play(model, num_games=a.games, mcts_simulations=a.sims, device="cpu", add_dirichlet_noise=True)
    dt = time.perf_counter() - t0
    positions = sum(len(g[0]) if isinstance(g, (tuple, list)) and g and hasattr(g[0], "__len__") else 0 for g in data) \
        if isinstance(data, list) else 0
    return {"mode": "legacy", "games": a.games, "sims": a.sims, "positions": int(positions), "seconds": dt,
            "positions_per_sec": positions / dt if positions else None,
            "sims_per_sec": positions * a.sims / dt if positions else None, "cores": 1,
            "what": "src.mcts.self_play, legacy python engine + python MCTS, tiny net, 1 CPU thread"}


def cmd_portable(a) -> dict:
    """The reference's own full-tree self-play on the CPU: `self_play_v1_portable_cpp` (v1/python/portable_cpp_self_play.py:26)
    over its compiled `_liuzhou_portable_cpp` tree (oracle/_ref) and its fp32 PyTorch ChessNet, `--threads` CPU threads,
    `--concurrent` games per network batch -- the reference's `--search_backend portable --portable_mcts_backend cpp` path,
    the only full-tree search it has.  Bounded by --max-plies (every game is cut after that many moves)."""
    sys.path.insert(0, str(PYSRC))
    sys.path.insert(0, str(REF_DIR))
    import torch
    from v1.python.portable_cpp_self_play import self_play_v1_portable_cpp

    threads = a.threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    model = _model(a.model_seed, "cpu", a.small_net)
    kw = dict(num_games=a.games, mcts_simulations=a.sims, temperature_init=1.0, temperature_final=0.1,
              temperature_threshold=10, exploration_weight=1.0, device="cpu", add_dirichlet_noise=bool(a.noise),
              soft_value_k=2.0, sample_moves=bool(a.sample), concurrent_games=a.concurrent or a.games, cpu_threads=threads)
    torch.manual_seed(a.seed)
    if a.warmup_games:
        self_play_v1_portable_cpp(model, **{**kw, "num_games": a.warmup_games, "concurrent_games": a.warmup_games,
                                            "max_game_plies": 1})
    t0 = time.perf_counter()
    batch, stats = self_play_v1_portable_cpp(model, max_game_plies=a.max_plies, **kw)
    dt = time.perf_counter() - t0
    positions = int(batch.num_samples)
    return {"mode": "portable", "games": a.games, "sims": a.sims, "max_plies": a.max_plies, "positions": positions,
            "seconds": dt, "positions_per_sec": positions / dt, "sims_per_sec": positions * a.sims / dt, "cores": threads,
            "what": "UNMODIFIED reference python self_play_v1_portable_cpp + its compiled C++ tree + fp32 ChessNet on the CPU"}


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["selfplay", "search", "legacy", "portable"])
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--v0core", choices=["ref", "shim"], default="ref")
    ap.add_argument("--games", type=int, default=64)
    ap.add_argument("--concurrent", type=int, default=0)
    ap.add_argument("--sims", type=int, default=200)
    ap.add_argument("--device", default="cuda:0")
    ap.add_argument("--noise", type=int, default=1)
    ap.add_argument("--sample", type=int, default=1)
    ap.add_argument("--seed", type=int, default=20260314)
    ap.add_argument("--model-seed", type=int, default=20260314)
    ap.add_argument("--small-net", type=int, default=0)
    ap.add_argument("--max-plies", type=int, default=512)
    ap.add_argument("--warmup-games", type=int, default=0)
    ap.add_argument("--states", default=None)
    ap.add_argument("--dump", default=None)
    a = ap.parse_args()
    os.environ.setdefault("CUBLAS_WORKSPACE_CONFIG", ":4096:8")
    out = {"selfplay": cmd_selfplay, "search": cmd_search, "legacy": cmd_legacy, "portable": cmd_portable}[a.mode](a)
    print(json.dumps(out), flush=True)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
