"""Experiment: cuDNN fused conv+bias+relu availability / speed for bf16 channels-last, and batch-size scaling."""
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def timeit(fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g.replay()
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for n in (1024, 2048, 4096, 8192):
    x = torch.randn(n, 128, 6, 6, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = torch.randn(128, 128, 3, 3, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last) * 0.05
    b = torch.randn(128, device="cuda", dtype=torch.bfloat16)
    z = torch.randn_like(x)
    t_conv = timeit(lambda: F.conv2d(x, w, None, 1, 1))
    print(f"n={n} conv2d: {t_conv:.1f} us ({n*36*128*1152*2/t_conv/1e6:.0f} TFLOP/s)")
    try:
        t = timeit(lambda: torch.cudnn_convolution_relu(x, w, b, [1, 1], [1, 1], [1, 1], 1))
        y = torch.cudnn_convolution_relu(x, w, b, [1, 1], [1, 1], [1, 1], 1)
        ref = torch.relu(F.conv2d(x, w, b, 1, 1))
        print(f"n={n} cudnn_convolution_relu: {t:.1f} us, max err {(y.float()-ref.float()).abs().max().item():.3f}, cl={y.is_contiguous(memory_format=torch.channels_last)}")
    except Exception as exc:
        print("cudnn_convolution_relu failed:", str(exc)[:200])
    try:
        t = timeit(lambda: torch.cudnn_convolution_add_relu(x, w, z, 1.0, b, [1, 1], [1, 1], [1, 1], 1))
        print(f"n={n} cudnn_convolution_add_relu: {t:.1f} us")
    except Exception as exc:
        print("cudnn_convolution_add_relu failed:", str(exc)[:200])
    t = timeit(lambda: F.conv2d(x, w, b, 1, 1))
    print(f"n={n} conv2d+bias: {t:.1f} us")
