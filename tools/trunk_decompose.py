"""Fused trunk kernel time at 4,096 boards with one engine switched off at a time (LZB_TRUNK_DEBUG bits: 4 = no MMAs,
16 = epilogue skips the shared-memory copy stores, 32 = no weight loads, 1 = no global output)."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
import bench
from liuzhou_b200.net import ChessNet, InferenceNet
torch.manual_seed(0)
n = 4096
net = InferenceNet(ChessNet(), "cuda:0")
x = net.new_input(n)
x[:, :11] = (torch.rand((n, 11, 6, 6), device="cuda") > 0.6).to(torch.bfloat16)
for _ in range(3): net._trunk_heads_conv(x)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    net._trunk_heads_conv(x)
ms = bench.sustained_replay_ms(g.replay, torch.cuda.current_stream(), seconds=0.5, warm_seconds=0.3)
print(f"{ms * 1e3:8.1f} us")
'''
names = {0: "full kernel", 16: "no copy stores (STS)", 32: "no weight loads", 48: "no STS, no weight loads", 4: "no MMAs",
         20: "no MMAs, no STS", 52: "no MMAs, no STS, no W loads (skeleton + TMEM loads)"}
for bits, name in (names.items() if __name__ == "__main__" else ()):
    env = dict(os.environ, LZB_TRUNK_DEBUG=str(bits))
    r = subprocess.run([sys.executable, "-c", CHILD % str(ROOT)], env=env, capture_output=True, text=True)
    print(f"debug={bits:2d} ({name:52s}): {r.stdout.strip() or r.stderr.strip()[-300:]}", flush=True)
