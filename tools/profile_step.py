"""Where does one self-play ply go?  CUDA-event breakdown of SelfPlayStepper.step() (4,096 games x 200 sims)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from liuzhou_b200.engine import SelfPlayStepper  # noqa: E402
from liuzhou_b200.net import ChessNet, InferenceNet  # noqa: E402

games = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 200
dev = torch.device("cuda:0")
torch.manual_seed(20260314)
net = InferenceNet(ChessNet(), dev)
st = SelfPlayStepper(net, games, simulations=sims, seed=20260314, device=dev)
st.diversify()
for _ in range(3):
    st.step()
torch.cuda.synchronize()


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


m = st.mcts
marks = []
orig_reset, orig_noise, orig_out = m.tree.reset, m._apply_root_noise, m.tree.root_outputs


def reset(*a, **k):
    marks.append(("search start", ev()))
    r = orig_reset(*a, **k)
    marks.append(("tree.reset", ev()))
    return r


def noise():
    marks.append(("root graph", ev()))
    r = orig_noise()
    marks.append(("dirichlet noise", ev()))
    return r


def outputs(with_priors=True):
    if not with_priors:
        marks.append(("waves", ev()))
    r = orig_out(with_priors)
    if not with_priors:
        marks.append(("root_outputs", ev()))
    return r


m.tree.reset, m._apply_root_noise, m.tree.root_outputs = reset, noise, outputs
for it in range(3):
    marks.clear()
    marks.append(("step start", ev()))
    st.step()
    marks.append(("policy + sample + trajectory + apply + refill", ev()))
    torch.cuda.synchronize()
    print(f"--- ply {it}")
    for (na, a), (nb, b) in zip(marks[:-1], marks[1:]):
        print(f"{a.elapsed_time(b):9.3f} ms  {nb}")
    print(f"{marks[0][1].elapsed_time(marks[-1][1]):9.3f} ms  total")
