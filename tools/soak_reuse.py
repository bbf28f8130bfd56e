"""Soak test of the steady-state stepper with subtree reuse at full size: N plies, then report the sticky tree flags,
the arena high-water mark after compaction, the largest root visit count (inherited + new) and the outcome counts."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from liuzhou_b200.engine import SelfPlayStepper  # noqa: E402
from liuzhou_b200.net import ChessNet, InferenceNet  # noqa: E402

plies = int(sys.argv[1]) if len(sys.argv) > 1 else 300
games = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
sims = int(sys.argv[3]) if len(sys.argv) > 3 else 200
dev = torch.device("cuda:0")
torch.manual_seed(20260314)
net = InferenceNet(ChessNet(), dev)
sp = SelfPlayStepper(net, games, simulations=sims, seed=1, device=dev, reuse_subtree=True)
sp.diversify(seed=3)
hi_nodes, hi_visits, finished = 0, 0, 0
t0 = time.perf_counter()
for p in range(plies):
    sp.step()
    if p % 10 == 9:
        st = sp.mcts.tree.stats()
        hi_nodes = max(hi_nodes, st["nodes_used"])
        hi_visits = max(hi_visits, int(sp.mcts.tree.visit[:games].max()))
        finished += 0
        if st["flags"]:
            print("FLAGS", st["flags"], "at ply", p)
            break
torch.cuda.synchronize()
st = sp.mcts.tree.stats()
print({"plies": p + 1, "seconds": round(time.perf_counter() - t0, 1), "flags": st["flags"], "nodes_after_compaction_max": hi_nodes,
       "capacity": st["capacity"], "max_inherited_root_visits": hi_visits, "outcomes_bwd": sp.outcomes.tolist()})
