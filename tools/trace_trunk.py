"""Debug: clock64 timeline of the fused trunk kernel (cluster 0, leader CTA): MMA thread (before the operand-copy wait /
copy ready / its taps issued) and epilogue warp 2 (waiting / accumulator ready / phase 1 done / (buffer free, copy
published) x copies / job done).  Run with LZB_TRUNK_DEBUG=8 (+4 no MMAs, +16 no copy stores, +32 no weight loads)."""
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
k = len(m) // 3 * 3
mm = m[:k].reshape(-1, 3)
wait_copy = mm[:, 1] - mm[:, 0]
issue = mm[:, 2] - mm[:, 1]
print(f"MMA thread per copy (3 taps): wait for the copy mean {wait_copy.mean():.0f}, wait-for-weights + issue mean {issue.mean():.0f} cycles;"
      f" totals {wait_copy.sum()} + {issue.sum()} of {m[-1]}")
# steady state: copies 60..78 = jobs 20..25 (layers 10..12 of both slots in round 0)
print("copy  t_before_wait  wait_copy  taps_issue")
for i in range(60, 78):
    print(f"{i:4d} {mm[i,0]:12d} {wait_copy[i]:9d} {issue[i]:9d}")
# epilogue: a 3x3 -> 3x3 job has 10 stamps; the stem job and the jobs around the heads conv differ, so walk by pattern:
# find jobs by matching against the MMA timeline is overkill -- print stamps 10 jobs from stamp 200 (deep inside round 0)
print("epilogue stamps from job ~20 (10 per job: waiting, acc ready, phase1 done, free0, pub0, free1, pub1, free2, pub2, done):")
for j in range(20, 26):
    st = e[j * 10:(j + 1) * 10]
    if len(st) < 10:
        break
    d = np.diff(st)
    print(f"job {j}: start {st[0]:9d}  wait_acc {d[0]:6d} phase1 {d[1]:6d} | free0 {d[2]:6d} write0 {d[3]:6d} | free1 {d[4]:6d} write1 {d[5]:6d} | free2 {d[6]:6d} write2 {d[7]:6d} | end {d[8]:4d}")
