"""A/B of LZB_HEADS_OVERLAP: graph replays of the network forward incl. heads (forward_priors) at 4,096 boards."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
import bench
from liuzhou_b200 import native
from liuzhou_b200.net import ChessNet, InferenceNet
torch.manual_seed(0)
n = 4096
net = InferenceNet(ChessNet(), "cuda:0")
x = net.new_input(n)
x[:, :11] = (torch.rand((n, 11, 6, 6), device="cuda") > 0.6).to(torch.bfloat16)
pb = native.PlayoutBatch(n, seed=1, device=torch.device("cuda:0"))
pb.run(max_steps=30)
states = pb.packed
pri = torch.empty((n, 220), device="cuda"); val = torch.empty((n,), device="cuda")
for _ in range(3): net.forward_priors(x, states, priors_out=pri, values_out=val)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    net.forward_priors(x, states, priors_out=pri, values_out=val)
ms = bench.sustained_replay_ms(g.replay, torch.cuda.current_stream(), seconds=0.6, warm_seconds=0.3)
print(f"{ms * 1e3:8.1f} us  checksum {float(pri.sum()):.3f} {float(val.sum()):.4f}")
'''
if __name__ == "__main__":
    for o in ("0", "1", "0", "1"):
        env = dict(os.environ, LZB_HEADS_OVERLAP=o)
        r = subprocess.run([sys.executable, "-c", CHILD % str(ROOT)], env=env, capture_output=True, text=True)
        print(f"overlap={o}: {r.stdout.strip() or r.stderr.strip()[-300:]}", flush=True)
