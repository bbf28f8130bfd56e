"""In-situ duration of the single tree kernel of a wave (expand + backup + next select + input encoding) and of the
heads tail, measured with CUDA events around eager launches inside a sustained stream of waves (4,096 x 200)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from liuzhou_b200.engine import SelfPlayStepper  # noqa: E402
from liuzhou_b200.net import ChessNet, InferenceNet  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(20260314)
net = InferenceNet(ChessNet(), dev)
sp = SelfPlayStepper(net, 4096, simulations=200, seed=1, device=dev, reuse_subtree=True)
sp.diversify(seed=3)
for _ in range(3):
    sp.step()
m, tree = sp.mcts, sp.mcts.tree
m._first_select()
ev = []
for i in range(400):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    m.net.forward_priors(m._wave_in, tree.pending_states, priors_out=m._wave_pri, values_out=m._wave_val)
    e[1].record()
    tree.complete_and_select(m._wave_pri, m._wave_val, m._wave_in)
    e[2].record()
    ev.append(e)
torch.cuda.synchronize()
late = ev[200:]
fwd = sum(a.elapsed_time(b) for a, b, _ in late) / len(late)
trk = sum(b.elapsed_time(c) for _, b, c in late) / len(late)
print(f"eager forward (22 convs + heads tail) {fwd * 1e3:8.1f} us   expand+select kernel {trk * 1e3:8.1f} us   sum {(fwd + trk) * 1e3:8.1f} us")
print(tree.stats())
