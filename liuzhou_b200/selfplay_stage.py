"""``--stage selfplay`` of the reference's ``v1/train.py`` (the command ``scripts/big_train_v1.sh:667-700`` issues once
per iteration) on the B200 engine:

    python -m liuzhou_b200.selfplay_stage --stage selfplay --devices cuda:0,cuda:1,... --self_play_games 32768 \
        --mcts_simulations 800 --self_play_concurrent_games 4096 --self_play_output RUN/selfplay_iter_001.pt \
        --self_play_iteration_seed 1 --self_play_stats_json RUN/selfplay_iter_001.json [--load_checkpoint CKPT]

Same flags, same outputs (``v1_sharded_manifest`` + chunk files next to ``--self_play_output``, the stats JSON of
``_build_self_play_report``, v1/train.py:439-463), so the next command of the script (``--stage train --self_play_input
…``) runs unchanged.  Flags of other stages are accepted and ignored.

Process model: one process per GPU.  Started plainly with several ``--devices`` it re-launches itself through
``torch.distributed.run`` (one rank per listed device, rendezvous on 127.0.0.1); started under torchrun it uses the
ranks it is given.  Rank r plays ``_split_games`` share r with seed ``iteration_seed*10007 + (r+1)*9973``
(v1/train.py:129-135,998); weights reach the ranks by NCCL broadcast, not through ``model_state_cpu.pt``.
"""
from __future__ import annotations

import argparse
import json
import os
import random
import subprocess
import sys
from typing import Any, Dict, List, Optional

import torch
from torch import nn

from .net import ChessNet

DEFAULT_MODEL_INIT_SEED = 20260314                      # v1/train.py:150-159, scripts/big_train_v1.sh:24


def init_model_stable_resnet(model: nn.Module, *, seed: int) -> None:
    """Seeded ResNet-style bootstrap init used when no checkpoint is given (v1/train.py:162-217): He-normal (fan_out)
    convs / linears, BatchNorm at identity, last BN of every block at 0, output layers ~N(0, 1e-3).  The ambient RNG
    state is restored afterwards.  Module traversal order and draw order match the reference, so the same seed gives
    the same weights."""
    if int(seed) <= 0:
        raise ValueError(f"seed must be positive for model init, got {int(seed)}")
    py_state, cpu_state = random.getstate(), torch.random.get_rng_state()
    cuda_states = torch.cuda.get_rng_state_all() if torch.cuda.is_available() else None
    try:
        random.seed(int(seed))
        torch.manual_seed(int(seed))
        if torch.cuda.is_available():
            torch.cuda.manual_seed_all(int(seed))
        for m in model.modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.BatchNorm2d):
                if m.weight is not None:
                    nn.init.ones_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
        for block in getattr(model, "blocks", []):
            bn2 = getattr(block, "bn2", None)
            if isinstance(bn2, nn.BatchNorm2d) and bn2.weight is not None:
                nn.init.zeros_(bn2.weight)
        head = getattr(model, "policy_head", None)
        for name in ("out_pos1", "out_pos2", "out_mark"):
            conv = getattr(head, name, None) if head is not None else None
            if isinstance(conv, nn.Conv2d):
                nn.init.normal_(conv.weight, mean=0.0, std=1e-3)
                if conv.bias is not None:
                    nn.init.zeros_(conv.bias)
        fc2 = getattr(getattr(model, "value_head", None), "fc2", None)
        if isinstance(fc2, nn.Linear):
            nn.init.normal_(fc2.weight, mean=0.0, std=1e-3)
            if fc2.bias is not None:
                nn.init.zeros_(fc2.bias)
    finally:
        random.setstate(py_state)
        torch.random.set_rng_state(cpu_state)
        if cuda_states is not None:
            torch.cuda.set_rng_state_all(cuda_states)


def load_checkpoint_into_model(model: nn.Module, path: Optional[str]) -> None:
    """``model_state_dict`` or a bare state dict, through CPU; strict first, then the shape-compatible subset
    (v1/train.py:1367-1411)."""
    if not path:
        return
    if not os.path.exists(path):
        raise FileNotFoundError(f"Checkpoint not found: {path}")
    ckpt = torch.load(path, map_location="cpu")
    state = ckpt.get("model_state_dict", ckpt) if isinstance(ckpt, dict) else ckpt
    try:
        model.load_state_dict(state, strict=True)
    except RuntimeError:
        own = model.state_dict()
        model.load_state_dict({k: v for k, v in state.items() if k in own and tuple(own[k].shape) == tuple(v.shape)},
                              strict=False)


def parse_device_list(primary: str, devices: Optional[str]) -> List[str]:
    """Comma list → canonical unique ``cuda:N`` names (v1/train.py:88-126); CUDA only -- there is no CPU path."""
    tokens = [t.strip() for t in str(devices or "").split(",") if t.strip()] or [str(primary).strip()]
    out: List[str] = []
    for t in tokens:
        d = torch.device(t)
        if d.type != "cuda":
            raise RuntimeError(f"liuzhou_b200 self-play runs on CUDA devices only, got {t!r}")
        name = f"cuda:{0 if d.index is None else int(d.index)}"
        if name not in out:
            out.append(name)
    return out


def build_self_play_report(stats, value_summary, soft_summary, mixed_summary) -> Dict[str, Any]:
    """The ``--self_play_stats_json`` document (v1/train.py:439-463)."""
    rep = stats.to_dict()
    games = max(1, int(stats.num_games))
    decisive = int(stats.black_wins + stats.white_wins)
    buckets = {str(d): int((stats.piece_delta_buckets or {}).get(str(d), 0) or 0) for d in range(-18, 19)}
    total = sum(buckets.values())
    rep.update({"decisive_games": decisive, "decisive_game_ratio": float(decisive / games),
                "draw_game_ratio": float(int(stats.draws) / games), "piece_delta_buckets": buckets,
                "piece_delta_bucket_total": int(total), "piece_delta_bucket_expected": int(stats.num_games),
                "piece_delta_bucket_coverage": float(total / games), "value_target_summary": dict(value_summary),
                "soft_value_target_summary": dict(soft_summary), "mixed_value_target_summary": dict(mixed_summary)})
    return rep


def build_parser() -> argparse.ArgumentParser:
    """The self-play-stage subset of v1/train.py:2817-3007 (same names, types and defaults)."""
    p = argparse.ArgumentParser(description="liuzhou_b200: v1 self-play stage on B200")
    p.add_argument("--stage", type=str, default="selfplay", choices=["selfplay"])
    p.add_argument("--pipeline", type=str, default="v1")                       # scripts/train_entry.py passes it
    p.add_argument("--self_play_games", type=int, default=4)
    p.add_argument("--mcts_simulations", type=int, default=32)
    p.add_argument("--soft_label_alpha", type=float, default=0.0)
    p.add_argument("--temperature_init", type=float, default=1.0)
    p.add_argument("--temperature_final", type=float, default=0.1)
    p.add_argument("--temperature_threshold", type=int, default=10)
    p.add_argument("--policy_target_temperature", type=float, default=None)
    p.add_argument("--policy_target_prior_pseudocount", type=float, default=0.0)
    p.add_argument("--self_play_sample_moves", action=argparse.BooleanOptionalAction, default=True)
    p.add_argument("--exploration_weight", type=float, default=1.0)
    p.add_argument("--dirichlet_alpha", type=float, default=0.3)
    p.add_argument("--dirichlet_epsilon", type=float, default=0.25)
    p.add_argument("--self_play_concurrent_games", type=int, default=8)
    p.add_argument("--self_play_opening_random_moves", type=int, default=0)
    p.add_argument("--self_play_backend", type=str, default=None, choices=["auto", "thread", "process"])
    p.add_argument("--search_backend", type=str, default="cuda_root", choices=["cuda_root", "portable"])
    p.add_argument("--portable_mcts_backend", type=str, default="python", choices=["python", "cpp"])
    p.add_argument("--portable_cpp_threads", type=int, default=1)
    p.add_argument("--self_play_shard_dir", type=str, default=None)
    p.add_argument("--self_play_target_samples_per_shard", type=int, default=0)
    p.add_argument("--self_play_chunk_target_bytes", type=int, default=0)
    p.add_argument("--soft_value_k", type=float, default=2.0)
    p.add_argument("--max_game_plies", type=int, default=512)
    p.add_argument("--sparse_ply", type=int, default=1)
    p.add_argument("--sparse_top_k", type=int, default=8)
    p.add_argument("--checkpoint_dir", type=str, default="./checkpoints_v1")
    p.add_argument("--device", type=str, default="cuda:0")
    p.add_argument("--devices", type=str, default=None)
    p.add_argument("--load_checkpoint", type=str, default=None)
    p.add_argument("--self_play_output", type=str, default=None)
    p.add_argument("--self_play_iteration_seed", type=int, default=None)
    p.add_argument("--self_play_stats_json", type=str, default=None)
    p.add_argument("--model_init_seed", type=int, default=None)
    p.add_argument("--leaves_per_wave", type=int, default=1, help="tree backend: leaves per tree per wave (ours)")
    return p


def _resolve_seeds(args) -> tuple:
    it_seed = 1 if args.self_play_iteration_seed is None else int(args.self_play_iteration_seed)
    if it_seed <= 0:
        raise ValueError(f"self_play_iteration_seed must be positive when provided, got {it_seed}")
    init_seed = args.model_init_seed
    if init_seed is None:
        env = str(os.environ.get("V1_MODEL_INIT_SEED", "")).strip()
        init_seed = int(env) if env else DEFAULT_MODEL_INIT_SEED
    return it_seed, int(init_seed)


def run_stage(args, devices: List[str]) -> Optional[Dict[str, Any]]:
    """One rank's part of the stage (collective when a process group exists). Rank 0 returns the report."""
    import torch.distributed as dist

    from .self_play_worker import run_self_play_iteration

    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    dev = torch.device(devices[int(os.environ.get("LOCAL_RANK", rank)) % len(devices)])
    torch.cuda.set_device(dev)
    it_seed, init_seed = _resolve_seeds(args)
    model = ChessNet()
    if rank == 0:                                       # the other ranks get the weights by broadcast
        if not args.load_checkpoint and init_seed > 0:
            init_model_stable_resnet(model, seed=init_seed)
        load_checkpoint_into_model(model, args.load_checkpoint)
    model.eval()
    out = str(args.self_play_output or os.path.join(args.checkpoint_dir, "selfplay_batch_v1.pt"))
    meta = {"stage": "selfplay", "source_checkpoint": str(args.load_checkpoint) if args.load_checkpoint else None,
            "self_play_devices": list(devices), "self_play_backend": args.self_play_backend or "auto",
            "search_backend": args.search_backend, "self_play_shard_dir": args.self_play_shard_dir,
            "mcts_simulations": int(args.mcts_simulations), "self_play_games": int(args.self_play_games),
            "self_play_concurrent_games": int(args.self_play_concurrent_games), "portable_self_play_workers": 1,
            "portable_mcts_backend": args.portable_mcts_backend, "portable_cpp_threads": int(args.portable_cpp_threads),
            "self_play_opening_random_moves": int(args.self_play_opening_random_moves),
            "self_play_iteration_seed": it_seed, "policy_target_temperature": args.policy_target_temperature,
            "policy_target_prior_pseudocount": float(args.policy_target_prior_pseudocount),
            "self_play_sample_moves": bool(args.self_play_sample_moves), "engine": "liuzhou_b200", "world_size": world}
    res = run_self_play_iteration(
        model, num_games=int(args.self_play_games), iteration_seed=it_seed, output_path=out, device=dev,
        metadata_base=meta, target_samples_per_shard=int(args.self_play_target_samples_per_shard),
        chunk_target_bytes=int(args.self_play_chunk_target_bytes), shard_dir=args.self_play_shard_dir,
        mcts_simulations=int(args.mcts_simulations), temperature_init=float(args.temperature_init),
        temperature_final=float(args.temperature_final), temperature_threshold=int(args.temperature_threshold),
        exploration_weight=float(args.exploration_weight), dirichlet_alpha=float(args.dirichlet_alpha),
        dirichlet_epsilon=float(args.dirichlet_epsilon), soft_value_k=float(args.soft_value_k),
        opening_random_moves=int(args.self_play_opening_random_moves), max_game_plies=int(args.max_game_plies),
        concurrent_games_per_device=int(args.self_play_concurrent_games), soft_label_alpha=float(args.soft_label_alpha),
        sample_moves=bool(args.self_play_sample_moves), sparse_ply=int(args.sparse_ply), sparse_top_k=int(args.sparse_top_k),
        search_backend=args.search_backend, portable_mcts_backend=args.portable_mcts_backend,
        portable_cpp_threads=int(args.portable_cpp_threads), policy_target_temperature=args.policy_target_temperature,
        policy_target_prior_pseudocount=float(args.policy_target_prior_pseudocount),
        leaves_per_wave=int(args.leaves_per_wave))
    if res is None:
        return None
    stats, v, s, m, n_shards = res
    report = build_self_play_report(stats, v, s, m)
    report["self_play_iteration_seed"] = it_seed
    if args.self_play_stats_json:
        os.makedirs(os.path.dirname(str(args.self_play_stats_json)) or ".", exist_ok=True)
        with open(str(args.self_play_stats_json), "w", encoding="utf-8") as f:
            json.dump(report, f, indent=2, ensure_ascii=False)
    print(f"[liuzhou_b200] selfplay saved: {out} (games={stats.num_games}, positions={stats.num_positions}, "
          f"W/L/D={stats.black_wins}/{stats.white_wins}/{stats.draws}, {stats.positions_per_sec:.0f} positions/s, "
          f"format=sharded_manifest num_shards={n_shards})", flush=True)
    return report


def main(argv: Optional[List[str]] = None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    args, _ignored = build_parser().parse_known_args(argv)           # flags of the train / infer stages are ignored
    devices = parse_device_list(args.device, args.devices)
    under_torchrun = "RANK" in os.environ and "WORLD_SIZE" in os.environ
    if len(devices) > 1 and not under_torchrun:
        port = 29500 + (os.getpid() % 2000)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={len(devices)}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), "-m", "liuzhou_b200.selfplay_stage", *argv]
        return subprocess.call(cmd)
    import torch.distributed as dist

    if under_torchrun and int(os.environ["WORLD_SIZE"]) > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(torch.device(devices[int(os.environ.get("LOCAL_RANK", "0")) % len(devices)]))
        dist.init_process_group("nccl")
    try:
        run_stage(args, devices)
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
