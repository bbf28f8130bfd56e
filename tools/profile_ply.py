"""Where one ply of the steady-state stepper goes: CUDA-event timing of its segments (4,096 games x 200 sims)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from liuzhou_b200.engine import SelfPlayStepper  # noqa: E402
from liuzhou_b200.net import ChessNet, InferenceNet  # noqa: E402
import liuzhou_b200.tree_search as ts  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(20260314)
net = InferenceNet(ChessNet(), dev)
sp = SelfPlayStepper(net, 4096, simulations=200, seed=1, device=dev, reuse_subtree=True)
sp.diversify(seed=3)
for _ in range(3):
    sp.step()
torch.cuda.synchronize()

marks = []


def mark(name):
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    marks.append((name, e))


orig_search = ts.TreeMCTS.search


def timed_search(self, root_states, **kw):
    cfg, tree = self.cfg, self.tree
    mark("search:start")
    keep = bool(cfg.reuse_subtree) and self._advanced
    self._advanced = False
    if keep:
        pass
    else:
        tree.reset(root_states, kw.get("active"))
    self._root_graph.replay()
    mark("root graph")
    self._apply_root_noise()
    mark("root noise")
    self._first_graph.replay()
    for _ in range(self.waves - 1):
        self._wave_graph.replay()
    self._last_graph.replay()
    mark("200 waves")
    self._advanced_saved = keep
    # the rest of search(): outputs / policy / sampling -- call the original tail by re-running its code path
    self.evals += self.num_trees * (1 + self.waves * tree.k)
    out = tree.root_outputs(with_priors=False)
    visits, legal, terminal = out["visit_counts"], out["legal_masks"], out["terminal"]
    temps = torch.as_tensor(kw["temperatures"], dtype=torch.float32, device=self.device).view(-1)
    selection = ts.policy_from_visits(visits, temps)
    has_mass = selection.sum(dim=1) > 0
    safe = torch.where(has_mass.view(-1, 1), selection, torch.full_like(selection, 1.0 / 220))
    chosen = torch.multinomial(safe, num_samples=1).view(-1)
    chosen = torch.where(has_mass & ~terminal, chosen, torch.full_like(chosen, -1))
    mark("outputs + policy + sample")
    return ts.TreeSearchOutput(legal_mask=legal, visit_counts=visits, policy_dense=selection, selection_policy_dense=selection,
                               root_value=out["root_values"], root_action_values=out["root_action_values"],
                               terminal_mask=terminal | ~has_mass, chosen_action_indices=chosen)


ts.TreeMCTS.search = timed_search
orig_advance = ts.TreeMCTS.advance


def timed_advance(self, *a, **k):
    mark("stepper: rows / apply / status")
    orig_advance(self, *a, **k)
    mark("advance_roots")


ts.TreeMCTS.advance = timed_advance
import time
n = 4
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(n):
    sp.step()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / n * 1e3
acc = {}
for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
    if n1 == "search:start":
        n1 = "between plies"
    acc[n1] = acc.get(n1, 0.0) + e0.elapsed_time(e1)
for k, v in acc.items():
    print(f"{k:32s} {v / n:9.3f} ms / ply")
print(f"{'wall per ply':32s} {wall:9.3f} ms")
