"""Debug: does the overlapped heads kernel really start under the trunk kernel's tail?  globaltimer stamps of every trunk CTA's
exit and every heads block's start for one forward at 4,096 boards (LZB_HEADS_OVERLAP=1 LZB_OVERLAP_TRACE=1 LZB_TRUNK_DEBUG=1024)."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from liuzhou_b200 import native  # noqa: E402
from liuzhou_b200.net import ChessNet, InferenceNet  # noqa: E402

torch.manual_seed(0)
n = 4096
net = InferenceNet(ChessNet(), "cuda:0")
x = net.new_input(n)
x[:, :11] = (torch.rand((n, 11, 6, 6), device="cuda") > 0.6).to(torch.bfloat16)
pb = native.PlayoutBatch(n, seed=1, device=torch.device("cuda:0"))
pb.run(max_steps=30)
pri = torch.empty((n, 220), device="cuda")
val = torch.empty((n,), device="cuda")
for _ in range(5):
    net.forward_priors(x, pb.packed, priors_out=pri, values_out=val)
torch.cuda.synchronize()
f = net._flags[n].cpu().numpy()
heads = f[4096:4096 + 2048].view(np.uint64)[:512].astype(np.int64)
trunk = f[4096 + 2048:4096 + 2048 + 1024].view(np.uint64)[:148].astype(np.int64)
t0 = trunk.min()
print("trunk CTA exits (us after the first exit): min %.1f  median %.1f  max %.1f" % (0.0, np.median(trunk - t0) / 1e3, (trunk.max() - t0) / 1e3))
print("sorted trunk exits:", np.round(np.sort(trunk - t0) / 1e3, 1)[::8].tolist())
hs = np.sort(heads - t0) / 1e3
print("heads block starts (us after the first trunk exit): min %.1f  p25 %.1f  median %.1f  p75 %.1f  max %.1f" %
      (hs.min(), hs[128], hs[256], hs[384], hs.max()))
