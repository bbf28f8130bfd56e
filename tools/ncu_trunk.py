"""Small program for ncu: a few forwards of the default network at 4,096 boards (21 conv launches + heads each)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from liuzhou_b200.net import ChessNet, InferenceNet  # noqa: E402

torch.manual_seed(20260314)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
net = InferenceNet(ChessNet(), "cuda:0")
x = net.new_input(n)
x[:, :11] = (torch.rand((n, 11, 6, 6), device="cuda") > 0.6).to(torch.bfloat16)
for _ in range(3):
    out = net._forward_eager(x)
torch.cuda.synchronize()
print("ok", [tuple(o.shape) for o in out])
