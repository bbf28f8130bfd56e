"""Timing experiments on the fused trunk kernel (debug bits: 4 no MMAs, 16 no copy stores, 64 two extra tcgen05.commit per tap,
128 half the weight bytes per tap)."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tools"))
from trunk_decompose import CHILD  # noqa: E402

if __name__ == "__main__":
    for bits in [int(a) for a in sys.argv[1:]] or (0, 64, 128, 20, 84, 148):
        env = dict(os.environ, LZB_TRUNK_DEBUG=str(bits))
        r = subprocess.run([sys.executable, "-c", CHILD % str(ROOT)], env=env, capture_output=True, text=True)
        out = r.stdout.strip() or r.stderr.strip()[-200:]
        print(f"debug={bits:3d}: {out}", flush=True)
