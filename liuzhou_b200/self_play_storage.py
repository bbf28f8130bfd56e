"""Self-play payload files: the reference's on-disk formats, written from device-resident batches.

Formats (all ``torch.save``'d dicts of plain tensors / python scalars, loadable by the unmodified reference):
  * ``v1_sharded_shard``          one chunk of positions: the five trajectory tensors + ``stats`` + ``metadata``
                                  (v1/python/self_play_storage.py:92-110, self_play_worker.py:467-492);
  * ``v1_worker_chunk_manifest``  what one worker wrote (self_play_worker.py:504-537);
  * ``v1_sharded_manifest``       the iteration-level index the trainer opens (v1/train.py:1141-1153, loader :1627-1707).

Helper names / semantics follow v1/python/self_play_storage.py (``estimate_bytes_per_sample``, ``plan_sample_ranges``,
``slice_batch_cpu``, ``save_self_play_payload``).  What is new here is *how* the bytes leave the GPU: the reference does
``batch.to("cpu")`` (pageable, synchronous) and then ``torch.save`` on the critical path; ``AsyncShardWriter`` copies
each chunk into pinned staging on a side stream and serialises it on a writer thread while the GPU already plays the
next wave ("a streaming-compatible sharded writer that still emits v1_sharded_manifest asynchronously").
"""
from __future__ import annotations

import math
import os
import queue
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple

import torch

from .trajectory_buffer import TensorSelfPlayBatch

_FIELDS = ("state_tensors", "legal_masks", "policy_targets", "value_targets", "soft_value_targets")
_SUMMARY_COUNTS = ("total", "finite_count", "nonfinite_count", "nonzero_count", "zero_count", "positive_count",
                   "negative_count", "near_zero_count", "ge_abs_0p05_count", "ge_abs_0p10_count", "ge_abs_0p20_count")


# ----------------------------------------------------------------------------------------------------------------------
# planning helpers (v1/python/self_play_storage.py:14-72)
# ----------------------------------------------------------------------------------------------------------------------
def _even_parts(total: int, parts: int) -> List[int]:
    total, parts = max(0, int(total)), max(1, int(parts))
    if total == 0:
        return []
    parts = min(parts, total)
    q, r = divmod(total, parts)
    return [q + (i < r) for i in range(parts)]


def estimate_bytes_per_sample(samples: TensorSelfPlayBatch) -> int:
    """Bytes of one position over the five tensors, from the actual dtypes/shapes (2,692 for the default layout)."""
    n = max(1, int(samples.num_samples))
    per = 0
    for name in _FIELDS:
        t = getattr(samples, name)
        if t.numel() > 0:
            per += t.element_size() * (t.numel() // n)
    return max(1, per)


def plan_sample_ranges(*, total_samples: int, num_shards: int, target_samples_per_shard: int = 0,
                       chunk_target_bytes: int = 0, bytes_per_sample: int = 0) -> List[Tuple[int, int]]:
    """[start, end) row ranges: at least ``num_shards`` pieces, more if a sample / byte target asks for smaller ones;
    a byte target overrides the sample target; pieces differ by at most one row."""
    total = int(total_samples)
    if total <= 0:
        return []
    pieces = max(1, min(int(num_shards), total))
    per_piece = max(0, int(target_samples_per_shard))
    if int(chunk_target_bytes) > 0:
        per_piece = max(1, int(chunk_target_bytes) // max(1, int(bytes_per_sample)))
    if per_piece > 0:
        pieces = min(total, max(pieces, int(math.ceil(total / float(per_piece)))))
    out, lo = [], 0
    for size in _even_parts(total, pieces):
        out.append((lo, lo + size))
        lo += size
    return out


def slice_batch_cpu(samples: TensorSelfPlayBatch, *, start: int, end: int) -> TensorSelfPlayBatch:
    s, e = int(start), int(end)
    return TensorSelfPlayBatch(*(getattr(samples, f)[s:e].to("cpu") for f in _FIELDS))


def _payload_dict(samples: TensorSelfPlayBatch, stats_payload: Dict[str, Any], metadata: Dict[str, Any]) -> Dict[str, Any]:
    d: Dict[str, Any] = {f: getattr(samples, f).detach().cpu() for f in _FIELDS}
    d["stats"] = dict(stats_payload)
    d["metadata"] = dict(metadata)
    return d


def save_self_play_payload(*, path: str, samples: TensorSelfPlayBatch, stats_payload: Dict[str, Any],
                           metadata: Dict[str, Any]) -> None:
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    torch.save(_payload_dict(samples, stats_payload, metadata), path)


# ----------------------------------------------------------------------------------------------------------------------
# target summaries (self_play_worker.py:56-162, v1/train.py:358-416,1436-1483) -- computed where the batch lives
# ----------------------------------------------------------------------------------------------------------------------
def summarize_scalar_targets(values: torch.Tensor) -> Dict[str, Any]:
    total = int(values.numel())
    out: Dict[str, Any] = {k: 0 for k in _SUMMARY_COUNTS}
    out["total"] = total
    sum_abs = 0.0
    if total > 0:
        finite = torch.isfinite(values)
        a = torch.where(finite, values, torch.zeros_like(values)).abs().to(torch.float32)
        pos = (finite & (values > 0)).sum()
        neg = (finite & (values < 0)).sum()
        counts = torch.stack([finite.sum(), pos, neg, (finite & (a <= 1e-6)).sum(), (finite & (a >= 0.05)).sum(),
                              (finite & (a >= 0.10)).sum(), (finite & (a >= 0.20)).sum()]).tolist()   # one sync
        sum_abs = float(a.sum().item())
        fin, p, n, nz0, g05, g10, g20 = (int(c) for c in counts)
        out.update(finite_count=fin, nonfinite_count=total - fin, positive_count=p, negative_count=n,
                   nonzero_count=p + n, zero_count=fin - p - n, near_zero_count=nz0, ge_abs_0p05_count=g05,
                   ge_abs_0p10_count=g10, ge_abs_0p20_count=g20)
    out["sum_abs"] = sum_abs
    _fill_ratios(out)
    return _ordered_summary(out)


def _fill_ratios(s: Dict[str, Any]) -> None:
    fin = max(1, int(s.get("finite_count", 0)))
    s["nonzero_ratio"] = float(int(s.get("nonzero_count", 0)) / fin)
    s["abs_mean"] = float(float(s.get("sum_abs", 0.0)) / fin)
    s["near_zero_ratio"] = float(int(s.get("near_zero_count", 0)) / fin)
    for tag in ("0p05", "0p10", "0p20"):
        s[f"ge_abs_{tag}_ratio"] = float(int(s.get(f"ge_abs_{tag}_count", 0)) / fin)


def _ordered_summary(s: Dict[str, Any]) -> Dict[str, Any]:
    order = ("total", "finite_count", "nonfinite_count", "nonzero_count", "zero_count", "positive_count",
             "negative_count", "nonzero_ratio", "sum_abs", "abs_mean", "near_zero_count", "near_zero_ratio",
             "ge_abs_0p05_count", "ge_abs_0p05_ratio", "ge_abs_0p10_count", "ge_abs_0p10_ratio", "ge_abs_0p20_count",
             "ge_abs_0p20_ratio")
    return {k: s[k] for k in order}


def merge_target_summaries(summaries: Iterable[Dict[str, Any]]) -> Dict[str, Any]:
    acc: Dict[str, Any] = {k: 0 for k in _SUMMARY_COUNTS}
    acc["sum_abs"] = 0.0
    for s in summaries:
        if not isinstance(s, dict):
            continue
        for k in _SUMMARY_COUNTS:
            acc[k] += int(s.get(k, 0) or 0)
        acc["sum_abs"] += float(s.get("sum_abs", 0.0) or 0.0)
    _fill_ratios(acc)
    return _ordered_summary(acc)


def mixed_value_targets(samples: TensorSelfPlayBatch, soft_label_alpha: float) -> torch.Tensor:
    """clamp((1-a)*value + a*soft, -1, 1), a clipped to [0,1] (self_play_worker.py:430,447-451)."""
    a = float(max(0.0, min(1.0, soft_label_alpha)))
    return torch.clamp((1.0 - a) * samples.value_targets + a * samples.soft_value_targets, min=-1.0, max=1.0)


# ----------------------------------------------------------------------------------------------------------------------
# asynchronous writer: device rows -> pinned staging (side stream) -> torch.save on a writer thread
# ----------------------------------------------------------------------------------------------------------------------
class AsyncShardWriter:
    """Write ``v1_sharded_shard`` files without stalling the compute stream.

    ``submit`` enqueues the device→pinned copies of one row range on a private copy stream (after the producer's
    work, via an event) and returns at once; a writer thread waits for the copy event, serialises the file and recycles
    the staging buffers.  ``max_inflight`` bounds pinned memory.  ``close()`` drains and re-raises the first failure.
    For CPU batches (the CPU tests) the copy degenerates to a clone.
    """

    def __init__(self, device=None, max_inflight: int = 3):
        self._dev = torch.device(device) if device is not None else None
        self._cuda = self._dev is not None and self._dev.type == "cuda"
        self._stream = torch.cuda.Stream(self._dev) if self._cuda else None
        self._q: "queue.Queue" = queue.Queue()
        self._slots = threading.Semaphore(max(1, int(max_inflight)))
        self._pool: Dict[Tuple, List[torch.Tensor]] = {}
        self._pool_lock = threading.Lock()
        self._error: Optional[BaseException] = None
        self.bytes_written = 0
        self.files: List[str] = []
        self._thread = threading.Thread(target=self._run, name="lz-shard-writer", daemon=True)
        self._thread.start()

    # staging buffers are recycled by (shape-without-rows, dtype, capacity-bucket)
    def _stage(self, t: torch.Tensor) -> torch.Tensor:
        rows = int(t.shape[0])
        cap = -(-max(1, rows) // 16384) * 16384                      # equal-sized chunks share a bucket
        key = (tuple(t.shape[1:]), t.dtype, cap)
        with self._pool_lock:
            free = self._pool.get(key)
            buf = free.pop() if free else None
        if buf is None:
            buf = torch.empty((cap,) + tuple(t.shape[1:]), dtype=t.dtype, pin_memory=self._cuda)
        buf._lz_key = key                                             # type: ignore[attr-defined]
        return buf

    def _recycle(self, bufs: Sequence[torch.Tensor]) -> None:
        with self._pool_lock:
            for b in bufs:
                self._pool.setdefault(b._lz_key, []).append(b)        # type: ignore[attr-defined]

    def submit(self, path: str, samples: TensorSelfPlayBatch, *, start: int, end: int, stats_payload: Dict[str, Any],
               metadata: Dict[str, Any]) -> None:
        if self._error is not None:
            raise RuntimeError("AsyncShardWriter: an earlier write failed") from self._error
        self._slots.acquire()
        s, e = int(start), int(end)
        staged, event = [], None
        try:
            self._stage_rows(samples, s, e, staged)
            if self._cuda:
                event = torch.cuda.Event()
                event.record(self._stream)
        except BaseException:
            self._recycle(staged)             # staging failed (e.g. pinned allocation): give the slot back
            self._slots.release()
            raise
        self._q.put((str(path), staged, e - s, event, dict(stats_payload), dict(metadata)))

    def flush(self) -> None:
        """Block until every submitted shard file is on disk (re-raises the first failure); the writer stays usable."""
        done = threading.Event()
        self._q.put(done)
        done.wait()
        if self._error is not None:
            raise RuntimeError("AsyncShardWriter: shard write failed") from self._error

    def _stage_rows(self, samples: TensorSelfPlayBatch, s: int, e: int, staged: list) -> None:
        if self._cuda:
            self._stream.wait_stream(torch.cuda.current_stream(self._dev))
            with torch.cuda.stream(self._stream):
                for f in _FIELDS:
                    src = getattr(samples, f)[s:e]
                    buf = self._stage(src)
                    buf[: e - s].copy_(src, non_blocking=True)
                    src.record_stream(self._stream)
                    staged.append(buf)
        else:
            for f in _FIELDS:
                src = getattr(samples, f)[s:e]
                buf = self._stage(src)
                buf[: e - s].copy_(src)
                staged.append(buf)

    def _run(self) -> None:
        while True:
            item = self._q.get()
            if item is None:
                return
            if isinstance(item, threading.Event):       # flush(): everything queued before it has been written
                item.set()
                continue
            path, staged, rows, event, stats_payload, metadata = item
            try:
                if event is not None:
                    event.synchronize()
                # clone() drops the over-allocated pinned storage so the file holds exactly `rows` rows
                payload: Dict[str, Any] = {f: b[:rows].clone() for f, b in zip(_FIELDS, staged)}
                payload["stats"] = stats_payload
                payload["metadata"] = metadata
                os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
                tmp = path + ".tmp"
                torch.save(payload, tmp)
                os.replace(tmp, path)                                   # readers never see a half-written shard
                self.bytes_written += os.path.getsize(path)
                self.files.append(path)
            except BaseException as exc:                                # noqa: BLE001 - surfaced by close()/submit()
                if self._error is None:
                    self._error = exc
            finally:
                self._recycle(staged)
                self._slots.release()

    def close(self) -> None:
        self._q.put(None)
        self._thread.join()
        if self._error is not None:
            raise RuntimeError("AsyncShardWriter: shard write failed") from self._error

    def __enter__(self) -> "AsyncShardWriter":
        return self

    def __exit__(self, *exc) -> None:
        self.close()


# ----------------------------------------------------------------------------------------------------------------------
# whole-batch sharded save and the loader (v1/train.py:1486-1553 and :1588-1734)
# ----------------------------------------------------------------------------------------------------------------------
def save_self_play_payload_sharded(*, path: str, samples: TensorSelfPlayBatch, stats_payload: Dict[str, Any],
                                   metadata: Dict[str, Any], num_shards: int, target_samples_per_shard: int = 0,
                                   chunk_target_bytes: int = 0, writer: Optional[AsyncShardWriter] = None) -> int:
    """Write ``<stem>.shardNN<ext>`` files + the ``v1_sharded_manifest`` at ``path``; returns the shard count
    (0 = the batch was empty and a single plain payload was written, as in the reference)."""
    n = int(samples.num_samples)
    bps = estimate_bytes_per_sample(samples) if n > 0 else 0
    ranges = plan_sample_ranges(total_samples=n, num_shards=int(num_shards),
                                target_samples_per_shard=int(target_samples_per_shard),
                                chunk_target_bytes=int(chunk_target_bytes), bytes_per_sample=int(bps))
    if not ranges:
        save_self_play_payload(path=path, samples=samples, stats_payload=stats_payload, metadata=metadata)
        return 0
    out_dir = os.path.dirname(path) or "."
    stem, ext = os.path.splitext(os.path.basename(path))
    ext = ext or ".pt"
    os.makedirs(out_dir, exist_ok=True)
    width = max(2, len(str(len(ranges) - 1)))
    own = writer is None
    w = writer or AsyncShardWriter(samples.state_tensors.device)
    names, sizes = [], []
    try:
        for i, (lo, hi) in enumerate(ranges):
            name = f"{stem}.shard{i:0{width}d}{ext}"
            meta = dict(metadata)
            meta.update({"payload_format": "v1_sharded_shard", "shard_index": i, "shard_count": len(ranges),
                         "shard_num_samples": hi - lo, "source_manifest": str(path)})
            w.submit(os.path.join(out_dir, name), samples, start=lo, end=hi, stats_payload=stats_payload, metadata=meta)
            names.append(name)
            sizes.append(hi - lo)
    finally:
        if own:
            w.close()
    if not own:
        w.flush()           # the manifest must not appear before the shard files it lists
    tmp = f"{path}.tmp.{os.getpid()}"
    torch.save({"payload_format": "v1_sharded_manifest", "version": 1, "num_samples": n, "num_shards": len(names),
                "shard_files": names, "shard_sizes": sizes, "chunk_target_bytes": int(chunk_target_bytes),
                "avg_bytes_per_sample": int(bps), "stats": dict(stats_payload), "metadata": dict(metadata)}, tmp)
    os.replace(tmp, path)
    return len(names)


def _batch_from_mapping(obj: Any, where: str) -> TensorSelfPlayBatch:
    if isinstance(obj, TensorSelfPlayBatch):
        return obj.to("cpu")
    if not isinstance(obj, dict):
        raise RuntimeError(f"Unsupported shard format in {where}: {type(obj)!r}")
    missing = [f for f in _FIELDS if f not in obj]
    if missing:
        raise RuntimeError(f"Missing keys in shard {where}: {missing}")
    return TensorSelfPlayBatch(*(obj[f].to("cpu") for f in _FIELDS))


def concat_batches(batches: Sequence[TensorSelfPlayBatch]) -> TensorSelfPlayBatch:
    if not batches:
        raise ValueError("no batches to concatenate")
    if len(batches) == 1:
        return batches[0]
    return TensorSelfPlayBatch(*(torch.cat([getattr(b, f) for b in batches], 0) for f in _FIELDS))


def load_self_play_payload(path: str, *, ddp_rank: Optional[int] = None, ddp_world_size: Optional[int] = None,
                           device=None) -> Tuple[TensorSelfPlayBatch, Dict[str, Any], Dict[str, Any]]:
    """Open a plain payload or a ``v1_sharded_manifest`` (shards ``i % world == rank`` under DDP) →
    ``(batch, stats, metadata)``; ``device`` optionally moves the merged batch (pinned → device) for the trainer."""
    if not os.path.exists(path):
        raise FileNotFoundError(f"Self-play payload not found: {path}")
    payload = torch.load(path, map_location="cpu")
    if isinstance(payload, dict) and str(payload.get("payload_format", "")).strip().lower() == "v1_sharded_manifest":
        files = [str(x).strip() for x in (payload.get("shard_files") or []) if str(x).strip()]
        if not files:
            raise RuntimeError(f"Invalid sharded self-play manifest {path}: shard_files missing or empty.")
        base = os.path.dirname(path) or "."
        paths = [p if os.path.isabs(p) else os.path.join(base, p) for p in files]
        picked = list(range(len(paths)))
        if ddp_rank is not None and ddp_world_size is not None and int(ddp_world_size) > 1:
            r, w = int(ddp_rank), int(ddp_world_size)
            if not 0 <= r < w:
                raise RuntimeError(f"Invalid ddp rank/world for shard load: rank={r}, world={w}")
            picked = [i for i in picked if i % w == r]
            if not picked:
                raise RuntimeError(f"DDP rank={r} got no shard from manifest={path} (world={w}, num_shards={len(paths)}).")
        with ThreadPoolExecutor(max_workers=min(8, len(picked)), thread_name_prefix="shard-load") as pool:
            parts = list(pool.map(lambda i: _batch_from_mapping(torch.load(paths[i], map_location="cpu"), paths[i]), picked))
        merged = concat_batches(parts)
        stats = payload.get("stats") if isinstance(payload.get("stats"), dict) else {}
        meta = dict(payload.get("metadata")) if isinstance(payload.get("metadata"), dict) else {}
        meta.update({"payload_sharded_manifest": True, "payload_format": "v1_sharded_manifest", "manifest_path": str(path),
                     "manifest_num_shards": len(paths), "loaded_shard_indices": picked, "loaded_shard_count": len(picked),
                     "loaded_num_samples": int(merged.num_samples)})
    else:
        merged = _batch_from_mapping(payload, path)
        stats = payload.get("stats") if isinstance(payload, dict) and isinstance(payload.get("stats"), dict) else {}
        meta = payload.get("metadata") if isinstance(payload, dict) and isinstance(payload.get("metadata"), dict) else {}
    if device is not None:
        merged = merged.to(device)
    return merged, stats, meta
