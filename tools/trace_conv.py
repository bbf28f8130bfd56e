"""Debug: per-stage clock64 timeline of the conv kernel's MMA-issuer and producer threads (cluster 0, leader CTA).
Run with LZB_CONV_DEBUG=8 (|1 |2 |4 for the no-epilogue / no-TMA / no-MMA variants)."""
import ctypes
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from liuzhou_b200._lib import lib  # noqa: E402
from liuzhou_b200.net import conv_bf16, pack_conv_weight  # noqa: E402

n = 4096
x = torch.randn(n, 128, 6, 6, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
w = torch.randn(128, 128, 3, 3, device="cuda", dtype=torch.bfloat16) * 0.03
wp = pack_conv_weight(w)
o = torch.empty_like(x)
for _ in range(3):
    conv_bf16(x, wp, relu1=True, out1=o)
torch.cuda.synchronize()
buf = np.zeros(8192, dtype=np.uint64)
lib().lzb_conv_debug_trace(buf.ctypes.data_as(ctypes.POINTER(ctypes.c_ulonglong)))
e = buf[2048:4096][buf[2048:4096] > 0].astype(np.int64)
m = buf[:2048][buf[:2048] > 0].astype(np.int64)
p = buf[4096:][buf[4096:] > 0].astype(np.int64)
t0 = m[0]
e = e - t0
print("epilogue (warp 2) per tile [tfull ok | +tmem ld | +chunk0 | +chunk1]:",
      [(int(a), int(b - a), int(c - b), int(d - c)) for a, b, c, d in zip(e[0::4], e[1::4], e[2::4], e[3::4])])
print("MMA thread stamps:", len(m), "producer stamps:", len(p))
m = m - t0
p = p - t0
# layout of m: [start], then per tile: [tempty ok], then per stage: [full ok, committed]
idx = 0
tiles = []
for t in range(8):
    if idx >= len(m):
        break
    te = 0
    st = []
    for s in range(9):
        if idx + 1 >= len(m):
            break
        st.append((m[idx], m[idx + 1])); idx += 2
    tiles.append((te, st))
for ti, (te, st) in enumerate(tiles):
    fulls = np.array([a for a, b in st]); comm = np.array([b for a, b in st])
    print(f"tile {ti}: tmem_empty ok @{te:7d}  first full @{fulls[0]:7d}  last commit @{comm[-1]:7d}  "
          f"stage period mean {np.diff(fulls).mean():6.0f} cyc (min {np.diff(fulls).min()}, max {np.diff(fulls).max()})  "
          f"issue (full->commit) mean {(comm - fulls).mean():5.0f}")
print("producer: empty-ok stamps, first 24 deltas:", np.diff(p[:25]))
print("MMA full-ok stamps first tile:", [int(a) for a, b in tiles[0][1]])
print("total span cycles:", int(m[-1]))
