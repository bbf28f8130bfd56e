"""bench.py contract checks that need no GPU: the reference arm prints ONE JSON line with the keys the driver parses, the
CPU arm of config 2 is the reference's own compiled engine where oracle/_ref is present, and our own arm refuses to run
without a GPU instead of falling back to anything."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def _run(*args, timeout=300):
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=timeout,
                          cwd=str(ROOT))


def test_reference_arm_playout_line():
    res = _run("--impl", "reference", "--workload", "playout", "--steps", "1", "--warmup", "1", "--ref-budget", "2")
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "selfplay_positions_per_sec" and d["unit"] == "positions/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1
    assert d["gpu_launches"] == 0
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cpu = d["cpu_baseline"]
    assert cpu["kind"] in ("reference", "port") and cpu["cores"] >= 1 and cpu["sample"]
    if (ROOT / "oracle" / "_ref" / "ref_playout").exists():
        assert cpu["kind"] == "reference"                       # the reference's own compiled scalar engine


def test_our_arm_needs_a_gpu():
    import torch

    if torch.cuda.is_available():
        return
    res = _run("--steps", "1", "--warmup", "3", timeout=120)
    assert "selfplay_positions_per_sec" not in res.stdout           # no number without the CUDA path
    assert "NVIDIA" in res.stderr or "CUDA" in res.stderr or "cuda" in res.stderr
