"""Op-level timing of the drop-in rule-engine / root-PUCT kernels on one B200: our kernel vs the reference's own CUDA
kernel (oracle/_ref/v0_core, compiled from the unmodified reference sources for sm_100) on the same inputs, with the
algorithmic bytes of SURVEY 8(d) -> GB/s -> fraction of the measured HBM peak.  CUDA events, 20 warm launches, median of
30.  Prints one JSON document (kept under profiles/)."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from liuzhou_b200 import native, v0_core  # noqa: E402
from tests._util import load_ref, random_apply_batch, random_mask_states, to_torch  # noqa: E402

DEV = "cuda:0"
peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6540.2}
HBM = float(peaks["hbm_gbs"])


def timed(fn, reps=30, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def graph_ms(fn, reps=50):
    """Kernel-only time: the call captured into a CUDA graph (its output allocations come from the graph's pool) and
    replayed back to back -- no Python / ctypes / allocator time between launches."""
    try:
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            keep = fn()                                     # noqa: F841  (outputs stay alive in the graph pool)
        for _ in range(5):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1) / reps
    except RuntimeError:
        torch.cuda.synchronize()
        return None


def row(name, units, bytes_per_unit, ours_ms, ref_ms, note="", kernel_ms=None):
    gbs = units * bytes_per_unit / (ours_ms / 1e3) / 1e9
    kgbs = None if not kernel_ms else units * bytes_per_unit / (kernel_ms / 1e3) / 1e9
    return {"op": name, "units": units, "algorithmic_bytes_per_unit": bytes_per_unit, "ours_ms": round(ours_ms, 4),
            "ours_gbs": round(gbs, 1), "frac_of_hbm_peak": round(gbs / HBM, 3),
            "kernel_ms (graph replay)": None if not kernel_ms else round(kernel_ms, 4),
            "kernel_gbs": None if kgbs is None else round(kgbs, 1),
            "kernel_frac_of_hbm_peak": None if kgbs is None else round(kgbs / HBM, 3),
            "reference_cuda_ms": None if ref_ms is None else round(ref_ms, 4),
            "speedup_vs_reference_cuda": None if ref_ms is None else round(ref_ms / ours_ms, 1), "note": note}


def main():
    ref = load_ref()
    ref_core = ref[0] if ref else None
    out = []
    b = 65536
    st = random_mask_states(b, 0xF00DCAFE)
    t = to_torch(st, DEV)

    def ref_or_none(fn):
        if ref_core is None:
            return None
        try:
            return timed(fn)
        except RuntimeError:
            return None

    out.append(row("encode_actions_fast", b, 3888, timed(lambda: v0_core.encode_actions_fast(*t[:10], 36, 144, 36, 4)),
                   ref_or_none(lambda: ref_core.encode_actions_fast(*t[:10], 36, 144, 36, 4)),
                   "148 B in + 3,740 B out per state (mask + metadata), outputs allocated per call as in the reference",
                   kernel_ms=graph_ms(lambda: v0_core.encode_actions_fast(*t[:10], 36, 144, 36, 4))))
    n, n_base = 1 << 20, 1 << 14                       # 16,384 distinct (state, action) pairs, each applied 64 times
    st2, codes, parents = random_apply_batch(n_base, 0xA11CEB0B)
    t2 = to_torch(st2, DEV)
    c = torch.from_numpy(codes).to(DEV).repeat(n // n_base, 1).contiguous()
    p = torch.from_numpy(parents).to(DEV).repeat(n // n_base).contiguous()
    out.append(row("batch_apply_moves", n, 384, timed(lambda: v0_core.batch_apply_moves(*t2, c, p)),
                   ref_or_none(lambda: ref_core.batch_apply_moves(*t2, c, p)), "204 B in + 180 B out per action, 12 output tensors",
                   kernel_ms=graph_ms(lambda: v0_core.batch_apply_moves(*t2, c, p))))
    out.append(row("states_to_model_input", b, 1692, timed(lambda: v0_core.states_to_model_input(*t[:5])),
                   ref_or_none(lambda: ref_core.states_to_model_input(*t[:5])), "108 B in + 1,584 B out per state",
                   kernel_ms=graph_ms(lambda: v0_core.states_to_model_input(*t[:5]))))
    bp = 1 << 22                                            # 4 M packed states (128 MB): larger than L2
    packed = native.pack_states(t).repeat(bp // b, 1).contiguous()
    out.append(row("legal_masks (packed)", bp, 68, timed(lambda: native.legal_masks(packed)), None,
                   "native layout: 32 B in + 32 B mask words + 4 B count out per state; ALU-bound bit logic",
                   kernel_ms=graph_ms(lambda: native.legal_masks(packed))))
    acts = torch.zeros((bp,), dtype=torch.int32, device=DEV)
    out.append(row("apply_actions (packed)", bp, 68, timed(lambda: native.apply_actions(packed, acts)), None,
                   "native layout: 32 + 4 B in, 32 B out per action", kernel_ms=graph_ms(lambda: native.apply_actions(packed, acts))))
    for r, m, s in ((4096, 64, 200), (4096, 64, 800), (4096, 64, 65536)):
        g = torch.Generator(device=DEV).manual_seed(1)
        valid = torch.rand((r, m), device=DEV, generator=g) < 0.4
        valid[:, 0] = True
        pri = torch.rand((r, m), device=DEV, generator=g) * valid
        pri = pri / pri.sum(1, keepdim=True)
        leaf = (torch.rand((r, m), device=DEV, generator=g) * 2 - 1) * valid
        reps = 5 if s > 1000 else 30
        ours = timed(lambda: v0_core.root_puct_allocate_visits(pri, leaf, valid, s, 1.0), reps=reps, warm=3)
        refms = None
        if ref_core is not None:
            try:
                refms = timed(lambda: ref_core.root_puct_allocate_visits(pri, leaf, valid, s, 1.0), reps=reps, warm=3)
            except RuntimeError:
                refms = None
        out.append(row(f"root_puct_allocate_visits S={s}", r * m, 17 + 4.0 / m, ours, refms,
                       "R=4096 roots x M=64 slots; sequential in S: latency-bound, bytes independent of S"))
    print(json.dumps({"device": torch.cuda.get_device_name(0), "hbm_peak_gbs": HBM, "ops": out}, indent=1))


if __name__ == "__main__":
    main()
