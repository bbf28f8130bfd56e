"""Decomposition of conv_tc_kernel<9,2> time at 4,096 states: the kernel's debug bits switch off one engine at a time
(LZB_CONV_DEBUG: 1 = epilogue drains TMEM but skips global loads / stores, 2 = no TMA loads, 4 = no MMAs), each variant
in its own process (the flag is read once), graph replays timed with CUDA events over >= 0.5 s."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
from liuzhou_b200.net import conv_bf16, pack_conv_weight
n = 4096
cl = torch.channels_last
x = torch.relu(torch.randn(n, 128, 6, 6, device="cuda", dtype=torch.bfloat16)).contiguous(memory_format=cl)
w = (torch.randn(128, 128, 3, 3, device="cuda", dtype=torch.bfloat16) * 0.03)
wp = pack_conv_weight(w)
bf = torch.randn(128, device="cuda")
res = torch.randn_like(x)
sc, sh = torch.rand(128, device="cuda") + 0.5, torch.randn(128, device="cuda")
o1, o2 = torch.empty_like(x), torch.empty_like(x)
def t(fn, chain=1):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(chain): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(200 // chain + 5): g.replay()
    reps = 3000 // chain
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps / chain * 1e3
a = t(lambda: conv_bf16(x, wp, bias=bf, relu1=True, out1=o1))
b = t(lambda: conv_bf16(x, wp, residual=res, scale=sc, shift=sh, want_out2=True, out1=o1, out2=o2))
c = t(lambda: conv_bf16(x, wp, bias=bf, relu1=True, out1=o1), chain=20)
print(f"conv1-epilogue {a:6.1f} us | conv2-epilogue {b:6.1f} us | conv1 x20 chained (PDL) {c:6.1f} us/launch")
'''
names = {0: "full kernel", 1: "no epilogue global IO", 2: "no TMA loads", 4: "no MMAs", 3: "no loads, no epilogue IO",
         5: "no MMAs, no epilogue IO", 6: "no loads, no MMAs (epilogue + skeleton)", 7: "skeleton only"}
for bits in (0, 1, 2, 4, 3, 5, 6, 7):
    env = dict(os.environ, LZB_CONV_DEBUG=str(bits))
    r = subprocess.run([sys.executable, "-c", CHILD % str(ROOT)], env=env, capture_output=True, text=True)
    print(f"debug={bits} ({names[bits]:40s}): {r.stdout.strip() or r.stderr.strip()[-300:]}", flush=True)
