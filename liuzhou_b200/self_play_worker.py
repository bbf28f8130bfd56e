"""Self-play worker entry + iteration-level driver -- call-compatible with the reference's
``run_self_play_worker(**kw)`` (v1/python/self_play_worker.py:276-552) and producing the same three file formats
(``v1_sharded_shard`` chunks, a ``v1_worker_chunk_manifest`` per worker, the ``v1_sharded_manifest`` the trainer opens;
v1/train.py:932-1160), so ``v1/train.py --stage train`` and ``scripts/big_train_v1.sh`` consume our output unchanged.

Differences in *how*, not *what*:
  * one worker = one GPU process under ``torch.distributed`` (the reference spawns a ProcessPoolExecutor per iteration
    and hands weights over through ``model_state_cpu.pt``): ``run_self_play_iteration`` broadcasts the weights over
    NCCL, every rank runs ``run_self_play_worker`` on its own share of the games, rank 0 merges the worker manifests;
  * chunks leave the GPU through ``AsyncShardWriter`` (pinned staging on a copy stream, serialisation on a thread)
    while the next group of games is already being played;
  * ``search_backend``: ``cuda_root`` → our root-PUCT backend, ``portable`` → the device-resident tree search (there
    is no python/cpp distinction; ``portable_mcts_backend`` / ``portable_cpp_threads`` are recorded, not used;
    ``policy_target_temperature`` / ``policy_target_prior_pseudocount`` act on the tree backend as in the reference).
"""
from __future__ import annotations

import os
import time
import traceback
from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch

from . import dist as lzdist
from .net import ChessNet
from .self_play import SelfPlayV1Stats, self_play_v1_gpu
from .self_play_storage import (AsyncShardWriter, estimate_bytes_per_sample, merge_target_summaries,
                                mixed_value_targets, plan_sample_ranges, summarize_scalar_targets)

_DELTAS = [str(d) for d in range(-18, 19)]
_LEGACY_TIMING_KEYS = ("root_puct_ms", "pack_writeback_ms", "self_play_step_ms", "finalize_ms")


def normalize_search_backend(backend: str) -> str:
    """v1/train.py:699-707 (+ our own 'tree' spelling)."""
    raw = str(backend).strip().lower()
    if raw in {"cuda_root", "v1", "root", "root_only"}:
        return "cuda_root"
    if raw in {"portable", "reference", "full_mcts", "tree"}:
        return "portable"
    raise ValueError(f"Unsupported search_backend={backend!r}; expected cuda_root or portable.")


def merge_self_play_stats(stats_list: Sequence[SelfPlayV1Stats], elapsed_sec: float) -> SelfPlayV1Stats:
    """Sum counters, game-weighted mean length, union of devices (self_play_worker.py:165-273, v1/train.py:290-355)."""
    elapsed = max(1e-9, float(elapsed_sec))
    games = sum(int(s.num_games) for s in stats_list)
    positions = sum(int(s.num_positions) for s in stats_list)
    timing_ms: Dict[str, float] = {}
    timing_calls: Dict[str, int] = {}
    counters: Dict[str, int] = {}
    buckets = {k: 0 for k in _DELTAS}
    devices: List[str] = []
    reasons: List[str] = []
    fallbacks = 0
    for s in stats_list:
        for k, v in s.step_timing_ms.items():
            timing_ms[str(k)] = timing_ms.get(str(k), 0.0) + float(v)
        for k, v in s.step_timing_calls.items():
            timing_calls[str(k)] = timing_calls.get(str(k), 0) + int(v)
        for k, v in s.mcts_counters.items():
            counters[str(k)] = counters.get(str(k), 0) + int(v)
        raw = s.piece_delta_buckets if isinstance(s.piece_delta_buckets, dict) else {}
        for k in _DELTAS:
            buckets[k] += int(raw.get(k, 0) or 0)
        fallbacks += int(getattr(s, "fallback_count", 0) or 0)
        reasons.extend(str(x) for x in getattr(s, "fallback_reasons", ()))
        for d in str(getattr(s, "device", "")).split(","):
            if d and d not in devices:
                devices.append(d)
    for k in _LEGACY_TIMING_KEYS:
        timing_ms.setdefault(k, 0.0)
        timing_calls.setdefault(k, 0)
    worker_ms = 1000.0 * sum(max(0.0, float(s.elapsed_sec)) for s in stats_list)
    ratio = {k: (min(1.0, max(0.0, v / worker_ms)) if worker_ms > 0 else 0.0) for k, v in timing_ms.items()}
    return SelfPlayV1Stats(
        num_games=games, num_positions=positions, black_wins=sum(int(s.black_wins) for s in stats_list),
        white_wins=sum(int(s.white_wins) for s in stats_list), draws=sum(int(s.draws) for s in stats_list),
        avg_game_length=float(sum(float(s.avg_game_length) * float(s.num_games) for s in stats_list) / max(1.0, float(games))),
        elapsed_sec=elapsed, positions_per_sec=float(positions / elapsed), games_per_sec=float(games / elapsed),
        step_timing_ms=timing_ms, step_timing_ratio=ratio, step_timing_calls=timing_calls, mcts_counters=counters,
        piece_delta_buckets=buckets, policy_target_audit={}, device=",".join(devices), fallback_count=fallbacks,
        fallback_reasons=tuple(reasons))


def stats_from_payload(p: Dict[str, Any]) -> SelfPlayV1Stats:
    """Inverse of ``SelfPlayV1Stats.to_dict`` (v1/train.py:630-683)."""
    p = p if isinstance(p, dict) else {}

    def num(k, cast, default=0):
        try:
            return cast(p.get(k, default) or default)
        except (TypeError, ValueError):
            return cast(default)

    def dmap(k, cast):
        raw = p.get(k)
        return {str(a): cast(b) for a, b in raw.items()} if isinstance(raw, dict) else {}

    return SelfPlayV1Stats(
        num_games=int(num("num_games", float)), num_positions=int(num("num_positions", float)),
        black_wins=int(num("black_wins", float)), white_wins=int(num("white_wins", float)), draws=int(num("draws", float)),
        avg_game_length=num("avg_game_length", float, 0.0), elapsed_sec=num("elapsed_sec", float, 0.0),
        positions_per_sec=num("positions_per_sec", float, 0.0), games_per_sec=num("games_per_sec", float, 0.0),
        step_timing_ms=dmap("step_timing_ms", float), step_timing_ratio=dmap("step_timing_ratio", float),
        step_timing_calls=dmap("step_timing_calls", int), mcts_counters=dmap("mcts_counters", int),
        piece_delta_buckets=dmap("piece_delta_buckets", int),
        policy_target_audit=dict(p.get("policy_target_audit") or {}) if isinstance(p.get("policy_target_audit"), dict) else {},
        device=str(p.get("device", "")), fallback_count=int(num("fallback_count", float)),
        fallback_reasons=tuple(str(x) for x in (p.get("fallback_reasons") or ())))


def run_self_play_worker(
    *,
    worker_idx: int,
    shard_device: str,
    shard_games: int,
    seed: int,
    model_state_path: Optional[str],
    output_path: str,
    mcts_simulations: int,
    temperature_init: float,
    temperature_final: float,
    temperature_threshold: int,
    exploration_weight: float,
    dirichlet_alpha: float,
    dirichlet_epsilon: float,
    soft_value_k: float,
    opening_random_moves: int,
    max_game_plies: int,
    concurrent_games_per_device: int,
    soft_label_alpha: float = 0.0,
    sample_moves: bool = True,
    target_samples_per_shard: int = 0,
    chunk_target_bytes: int = 0,
    chunk_output_dir: Optional[str] = None,
    chunk_file_prefix: Optional[str] = None,
    chunk_file_ext: str = ".pt",
    sparse_ply: int = 1,
    sparse_top_k: int = 8,
    search_backend: str = "cuda_root",
    portable_mcts_backend: str = "python",
    portable_cpp_threads: int = 1,
    policy_target_temperature: Optional[float] = None,
    policy_target_prior_pseudocount: float = 0.0,
    model: Optional[torch.nn.Module] = None,
    leaves_per_wave: int = 1,
) -> Dict[str, Any]:
    """Play ``shard_games`` games on ``shard_device`` in groups of ``concurrent_games_per_device`` and persist them as
    chunk files + one worker manifest at ``output_path``.  ``model`` (already holding the weights, e.g. after the NCCL
    broadcast) may replace ``model_state_path``; everything else is the reference's keyword surface."""
    try:
        torch.manual_seed(int(seed))
        dev = torch.device(str(shard_device))
        if dev.type != "cuda":
            raise RuntimeError("liuzhou_b200.run_self_play_worker runs on CUDA devices only (no CPU fallback)")
        torch.cuda.set_device(dev)
        torch.cuda.manual_seed(int(seed))
        games_total = max(0, int(shard_games))
        if games_total <= 0:
            raise ValueError(f"shard_games must be positive in worker, got {games_total}")
        out_dir = str(chunk_output_dir or "").strip()
        prefix = str(chunk_file_prefix or "").strip()
        if not out_dir or not prefix:
            raise ValueError("run_self_play_worker requires chunk_output_dir and chunk_file_prefix "
                             "to emit worker manifest output.")
        backend = normalize_search_backend(search_backend)
        group = max(1, min(games_total, int(concurrent_games_per_device)))

        if model is None:
            state = torch.load(str(model_state_path), map_location="cpu")
            if not isinstance(state, dict):
                raise RuntimeError(f"Invalid model_state payload type: {type(state)!r} ({model_state_path})")
            model = ChessNet()
            model.load_state_dict(state, strict=True)
        model = model.to(dev).eval()
        # ONE inference wrapper (BatchNorm folding, packed weights) and ONE set of search engines (tree arenas, CUDA
        # graphs) per worker, reused by every group of games
        from .net import InferenceNet

        net = model if isinstance(model, InferenceNet) else InferenceNet(model, dev)
        engines: Dict[Any, Any] = {}

        common_meta = {
            "worker_idx": int(worker_idx), "device": str(dev), "games": games_total, "games_per_chunk": group,
            "graph_retry_off": False, "memory_anchor_mb": 0, "opening_random_moves": int(opening_random_moves),
            "search_backend": str(search_backend), "portable_mcts_backend": str(portable_mcts_backend),
            "portable_cpp_threads": int(portable_cpp_threads), "policy_target_temperature": policy_target_temperature,
            "policy_target_prior_pseudocount": float(policy_target_prior_pseudocount),
        }
        stats_parts: List[SelfPlayV1Stats] = []
        summaries: Tuple[List, List, List] = ([], [], [])
        files: List[str] = []
        sizes: List[int] = []
        bps_num = bps_den = 0
        started = time.perf_counter()
        with AsyncShardWriter(dev) as writer:
            left = games_total
            while left > 0:
                n_games = min(group, left)
                batch, stats = self_play_v1_gpu(
                    model=net, num_games=n_games, mcts_simulations=int(mcts_simulations),
                    temperature_init=float(temperature_init), temperature_final=float(temperature_final),
                    temperature_threshold=int(temperature_threshold), exploration_weight=float(exploration_weight),
                    device=str(dev), add_dirichlet_noise=True, dirichlet_alpha=float(dirichlet_alpha),
                    dirichlet_epsilon=float(dirichlet_epsilon), soft_value_k=float(soft_value_k),
                    opening_random_moves=int(opening_random_moves), max_game_plies=int(max_game_plies),
                    sample_moves=bool(sample_moves), concurrent_games=n_games, sparse_ply=int(sparse_ply),
                    sparse_top_k=int(sparse_top_k), verbose=False,
                    search_backend="root" if backend == "cuda_root" else "tree", leaves_per_wave=int(leaves_per_wave),
                    policy_target_temperature=policy_target_temperature,
                    policy_target_prior_pseudocount=float(policy_target_prior_pseudocount), engine_cache=engines)
                stats_parts.append(stats)
                summaries[0].append(summarize_scalar_targets(batch.value_targets))
                summaries[1].append(summarize_scalar_targets(batch.soft_value_targets))
                summaries[2].append(summarize_scalar_targets(mixed_value_targets(batch, soft_label_alpha)))
                bps = estimate_bytes_per_sample(batch)
                bps_num += bps * max(1, int(batch.num_samples))
                bps_den += max(1, int(batch.num_samples))
                for lo, hi in plan_sample_ranges(total_samples=int(batch.num_samples), num_shards=1,
                                                 target_samples_per_shard=int(target_samples_per_shard),
                                                 chunk_target_bytes=int(chunk_target_bytes), bytes_per_sample=int(bps)):
                    name = f"{prefix}.chunk{len(files):05d}{chunk_file_ext}"
                    meta = dict(common_meta)
                    meta.update({"payload_format": "v1_sharded_shard", "num_selfplay_batches": len(stats_parts),
                                 "saved_chunk_index": len(files), "source_worker_manifest": os.path.basename(str(output_path))})
                    writer.submit(os.path.join(out_dir, name), batch, start=lo, end=hi, stats_payload={}, metadata=meta)
                    files.append(name)
                    sizes.append(hi - lo)
                left -= n_games
        merged = merge_self_play_stats(stats_parts, elapsed_sec=time.perf_counter() - started)

        manifest_meta = dict(common_meta)
        manifest_meta.update({"num_selfplay_batches": len(stats_parts), "saved_chunks": len(files)})
        manifest = {
            "payload_format": "v1_worker_chunk_manifest", "version": 1, "num_samples": int(sum(sizes)),
            "num_shards": len(files), "shard_files": list(files), "shard_sizes": list(sizes),
            "chunk_target_bytes": int(chunk_target_bytes), "avg_bytes_per_sample": int(bps_num // max(1, bps_den)),
            "stats": merged.to_dict(), "value_target_summary": merge_target_summaries(summaries[0]),
            "soft_value_target_summary": merge_target_summaries(summaries[1]),
            "mixed_value_target_summary": merge_target_summaries(summaries[2]), "metadata": manifest_meta,
        }
        os.makedirs(os.path.dirname(str(output_path)) or ".", exist_ok=True)
        _atomic_save(manifest, str(output_path))      # after the writer has drained: every listed chunk file is on disk
        return {"worker_idx": int(worker_idx), "device": str(dev), "games": games_total, "output_path": str(output_path),
                "num_samples": int(sum(sizes)), "saved_chunks": len(files)}
    except Exception as exc:
        raise RuntimeError("v1 self-play process worker failed: "
                           f"worker={int(worker_idx)}, device={str(shard_device)}, games={int(shard_games)}\n"
                           f"{traceback.format_exc()}") from exc


def _atomic_save(obj, path: str) -> None:
    """torch.save through a temporary file + os.replace: a trainer polling for the manifest never sees a partial file."""
    tmp = f"{path}.tmp.{os.getpid()}"
    torch.save(obj, tmp)
    os.replace(tmp, path)


def merge_worker_manifests(manifest_paths: Sequence[str], *, output_path: str, metadata_base: Dict[str, Any],
                           target_samples_per_shard: int, chunk_target_bytes: int, elapsed_sec: float
                           ) -> Tuple[SelfPlayV1Stats, Dict[str, Any], Dict[str, Any], Dict[str, Any], int]:
    """Parent-side merge of ``v1_worker_chunk_manifest`` files into the ``v1_sharded_manifest`` at ``output_path``
    (v1/train.py:1053-1160).  Worker order = order of ``manifest_paths`` (ascending worker index)."""
    files: List[str] = []
    sizes: List[int] = []
    stats_list: List[SelfPlayV1Stats] = []
    sums: Tuple[List, List, List] = ([], [], [])
    bps_num = bps_den = 0
    for mp_ in manifest_paths:
        wm = torch.load(str(mp_), map_location="cpu")
        if not isinstance(wm, dict) or str(wm.get("payload_format", "")).strip().lower() != "v1_worker_chunk_manifest":
            raise RuntimeError(f"Invalid worker manifest payload_format in {mp_}")
        sf, ss = wm.get("shard_files"), wm.get("shard_sizes")
        if not isinstance(sf, list) or not isinstance(ss, list):
            raise RuntimeError(f"Worker manifest missing shard file lists: {mp_}")
        for i, entry in enumerate(sf):
            if str(entry).strip():
                files.append(str(entry).strip())
                sizes.append(int(ss[i]) if i < len(ss) else 0)
        stats_list.append(stats_from_payload(wm.get("stats", {})))
        for dst, key in zip(sums, ("value_target_summary", "soft_value_target_summary", "mixed_value_target_summary")):
            if isinstance(wm.get(key), dict):
                dst.append(wm[key])
        n, b = int(wm.get("num_samples", 0) or 0), int(wm.get("avg_bytes_per_sample", 0) or 0)
        if n > 0 and b > 0:
            bps_num += n * b
            bps_den += n
    if not files:
        raise RuntimeError("Process self-play direct-save produced no chunk files.")
    merged = merge_self_play_stats(stats_list, elapsed_sec=elapsed_sec)
    v, s, m = (merge_target_summaries(x) for x in sums)
    meta = dict(metadata_base)
    meta.update({"self_play_target_samples_per_shard": int(target_samples_per_shard),
                 "self_play_chunk_target_bytes": int(chunk_target_bytes), "value_target_summary": dict(v),
                 "soft_value_target_summary": dict(s), "mixed_value_target_summary": dict(m)})
    os.makedirs(os.path.dirname(str(output_path)) or ".", exist_ok=True)
    _atomic_save({"payload_format": "v1_sharded_manifest", "version": 1, "num_samples": int(sum(sizes)),
                  "num_shards": len(files), "shard_files": list(files), "shard_sizes": list(sizes),
                  "chunk_target_bytes": int(chunk_target_bytes), "avg_bytes_per_sample": int(bps_num // max(1, bps_den)),
                  "stats": merged.to_dict(), "metadata": meta}, str(output_path))
    return merged, v, s, m, len(files)


def run_self_play_iteration(model: torch.nn.Module, *, num_games: int, iteration_seed: int, output_path: str,
                            device=None, group=None, metadata_base: Optional[Dict[str, Any]] = None,
                            target_samples_per_shard: int = 0, chunk_target_bytes: int = 0,
                            shard_dir: Optional[str] = None, **worker_kwargs):
    """The self-play stage of one training iteration (``v1/train.py --stage selfplay``, :932-1160) as a collective
    call: every rank of the process group is one worker (= one GPU).  Weights travel from rank 0 by NCCL broadcast,
    games split as ``_split_games``, seeds as ``iteration_seed*10007 + (rank+1)*9973``, each rank writes its own chunk
    files next to ``output_path``, rank 0 merges the manifests.  Returns the reference's 5-tuple on rank 0, None elsewhere.
    ``worker_kwargs`` = the remaining ``run_self_play_worker`` keywords (mcts_simulations, temperatures, ...)."""
    import torch.distributed as dist

    world, rank = lzdist._world(group), lzdist._rank(group)
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    out_dir = os.path.dirname(str(output_path)) or "."
    stem, ext = os.path.splitext(os.path.basename(str(output_path)))
    ext = ext or ".pt"
    workspace = os.path.abspath(shard_dir) if shard_dir else os.path.join(out_dir, f".{stem}.workers")
    os.makedirs(workspace, exist_ok=True)
    os.makedirs(out_dir, exist_ok=True)
    games = lzdist.split_games(int(num_games), world)
    started = time.perf_counter()
    model = model.to(dev)
    lzdist.broadcast_model(model, src=0, group=group)
    my_manifest = os.path.join(workspace, f"worker_manifest_{int(iteration_seed):06d}_{rank:02d}.pt")
    failure: Optional[BaseException] = None
    if games[rank] > 0:
        try:
            run_self_play_worker(worker_idx=rank, shard_device=str(dev), shard_games=games[rank],
                                 seed=lzdist.rank_seed(iteration_seed, rank), model_state_path=None, model=model,
                                 output_path=my_manifest, target_samples_per_shard=int(target_samples_per_shard),
                                 chunk_target_bytes=int(chunk_target_bytes), chunk_output_dir=out_dir,
                                 chunk_file_prefix=f"{stem}.w{rank:02d}", chunk_file_ext=ext, **worker_kwargs)
        except Exception as exc:  # noqa: BLE001  reported to every rank below instead of leaving them in the barrier
            failure = exc
    if world > 1:
        # a failed worker must not leave the other ranks waiting: the failure count is all-reduced (this IS the barrier)
        failed = torch.tensor([1 if failure is not None else 0], dtype=torch.int32,
                              device=dev if dist.get_backend(group) == "nccl" else "cpu")
        dist.all_reduce(failed, op=dist.ReduceOp.SUM, group=group)
        if int(failed.item()) > 0 and failure is None:
            raise RuntimeError(f"self-play iteration aborted: {int(failed.item())} other rank(s) failed in their worker")
    if failure is not None:
        raise failure
    if rank != 0:
        return None
    manifests = [os.path.join(workspace, f"worker_manifest_{int(iteration_seed):06d}_{r:02d}.pt")
                 for r in range(world) if games[r] > 0]
    result = merge_worker_manifests(manifests, output_path=str(output_path), metadata_base=dict(metadata_base or {}),
                                    target_samples_per_shard=int(target_samples_per_shard),
                                    chunk_target_bytes=int(chunk_target_bytes), elapsed_sec=time.perf_counter() - started)
    if not shard_dir:                      # the auto-created workspace only held the per-rank manifests (v1/train.py:1165)
        import shutil

        shutil.rmtree(workspace, ignore_errors=True)
    return result
