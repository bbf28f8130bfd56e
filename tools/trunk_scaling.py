"""Trunk (21 conv launches) + heads timing vs batch size: does the working set (3 activation buffers of n x 9.2 KB) fit
the L2?  Sustained graph replays (>= 0.6 s each)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from liuzhou_b200.net import ChessNet, InferenceNet  # noqa: E402

torch.manual_seed(20260314)
net = InferenceNet(ChessNet(), "cuda:0")
stream = torch.cuda.current_stream()
for n in [int(a) for a in sys.argv[1:]] or [1024, 2048, 3072, 4096]:
    x = net.new_input(n)
    x[:, :11] = (torch.rand((n, 11, 6, 6), device="cuda") > 0.6).to(torch.bfloat16)
    for fn, name in ((lambda: net.trunk(x), "trunk"), (lambda: net._forward_eager(x), "forward")):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        ms = bench.sustained_replay_ms(g.replay, stream, seconds=0.6, warm_seconds=0.3)
        print(f"n={n:5d} {name:8s} {ms * 1e3:8.1f} us  = {ms * 1e6 / n:7.1f} ns/board  "
              f"({n * net.flops_per_state / ms / 1e9:6.0f} TFLOP/s whole-net flops)", flush=True)
