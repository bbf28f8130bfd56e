"""Debug: clock64 timeline of tree_expand_select_kernel (8 sampled warps) in the last wave of a 60-wave search at 4,096
trees.  Run with LZB_TREE_TRACE=1.  Stamps: 0 start, 1 leaf slot read, 2 state + info read / legal set, 3 priors staged,
4 prior sum done, 5 arena allocation returned, 6 children written, 7 backup done, 8 root read, 9.. one per level, 20 descent
done, 21 leaf state + input rows written."""
import ctypes
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from liuzhou_b200 import native  # noqa: E402
from liuzhou_b200._lib import lib  # noqa: E402
from liuzhou_b200.net import ChessNet, InferenceNet  # noqa: E402
from liuzhou_b200.tree_search import TreeMCTS, TreeMCTSConfig  # noqa: E402

n = 4096
dev = torch.device("cuda:0")
torch.manual_seed(20260314)
net = InferenceNet(ChessNet(), dev)
pb = native.PlayoutBatch(n, seed=20260314, device=dev)
pb.run(max_steps=40)
sims = int(sys.argv[1]) if len(sys.argv) > 1 else 60
m = TreeMCTS(net, n, TreeMCTSConfig(num_simulations=sims, add_dirichlet_noise=False, use_cuda_graph=False), dev)
m.search(pb.packed)
torch.cuda.synchronize()
buf = np.zeros(256, dtype=np.uint64)
rc = lib().lzb_tree_debug_trace(buf.ctypes.data_as(ctypes.POINTER(ctypes.c_ulonglong)))
print("rc", rc)
tr = buf.reshape(8, 32).astype(np.int64)
names = {1: "slot", 2: "state/legal", 3: "priors", 4: "sum", 5: "alloc", 6: "children", 7: "backup", 8: "root"}
for w in range(8):
    t = tr[w]
    if t[0] == 0:
        continue
    depth = int(t[22])
    seq = [0, 1, 2, 3, 4, 5, 6, 7, 8] + [8 + d for d in range(1, min(depth, 8) + 1)] + [20, 21]
    parts = []
    prev = t[0]
    for i in seq[1:]:
        if t[i] == 0:
            continue
        label = names.get(i, f"L{i - 8}" if 9 <= i <= 16 else ("descent" if i == 20 else "leaf+encode"))
        parts.append(f"{label} {t[i] - prev}")
        prev = t[i]
    print(f"warp {w}: total {t[21] - t[0]} cycles, depth {depth}: " + ", ".join(parts))
