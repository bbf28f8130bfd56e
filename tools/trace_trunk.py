"""Debug: clock64 timeline of the fused trunk kernel (cluster 0, leader CTA): MMA thread (before / after every operand-copy
wait) and epilogue warp 2 (waiting / accumulator ready / job done).  Run with LZB_TRUNK_DEBUG=8."""
import ctypes
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from liuzhou_b200._lib import lib  # noqa: E402
from liuzhou_b200.net import ChessNet, InferenceNet  # noqa: E402

torch.manual_seed(0)
n = 4096
net = InferenceNet(ChessNet(), "cuda:0")
x = net.new_input(n)
x[:, :11] = (torch.rand((n, 11, 6, 6), device="cuda") > 0.6).to(torch.bfloat16)
for _ in range(3):
    net._trunk_heads_conv(x)
torch.cuda.synchronize()
buf = np.zeros(8192, dtype=np.uint64)
lib().lzb_trunk_debug_trace(buf.ctypes.data_as(ctypes.POINTER(ctypes.c_ulonglong)))
m = buf[:4096][buf[:4096] > 0].astype(np.int64)
e = buf[4096:][buf[4096:] > 0].astype(np.int64)
t0 = m[0]
m -= t0
e -= t0
print("MMA stamps", len(m), "epilogue stamps", len(e), "span cycles", int(max(m[-1], e[-1])))
# MMA: pairs (before wait, after wait) per copy; 3 copies per 3x3 job
waits = m[1::2] - m[0::2]
print("MMA thread: total wait for operand copies", int(waits.sum()), "cycles of", int(m[-1]), "; mean per copy", float(waits.mean()))
starts = m[1::2]
print("copy-ready stamps, first 40:", starts[:40].tolist())
print("copy waits, first 40:", waits[:40].tolist())
# epilogue stamps per 3x3 job: waiting, ready, phase1 done, (buffer free, published) x 3, job done = 10 stamps
# (jobs whose next layer is the heads conv have 1 copy = 6 stamps, the heads job itself has 4) -> print raw deltas of the first jobs
d = np.diff(e)
print("epilogue stamp deltas (first 64):", d[:64].tolist())
print("epilogue stamps (first 12):", e[:12].tolist())
