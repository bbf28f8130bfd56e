"""Is the fused trunk kernel's weight stream latency-bound?  Time at 4,096 boards with the weight ring limited to 2..5
stages (LZB_TRUNK_W_STAGES), for the full kernel and with MMAs + copy stores switched off (debug 20: pure streaming)."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tools"))
from trunk_decompose import CHILD  # noqa: E402

if __name__ == "__main__":
    for bits in (0, 20, 4):
        for stages in (2, 3, 4, 5):
            env = dict(os.environ, LZB_TRUNK_DEBUG=str(bits), LZB_TRUNK_W_STAGES=str(stages))
            r = subprocess.run([sys.executable, "-c", CHILD % str(ROOT)], env=env, capture_output=True, text=True)
            out = r.stdout.strip() or r.stderr.strip()[-200:]
            print(f"debug={bits:2d} w_stages={stages}: {out}", flush=True)
