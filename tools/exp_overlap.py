"""Experiment: does splitting the wave batch into G game groups on G free-running streams hide the HBM-bound
passes (bn_relu, heads, tree kernels) behind the other group's tensor-core convolutions?

usage: exp_overlap.py [total_games] [waves]
Prints forward-only and full-wave (select -> network -> expand/backup) times per 4,096-leaf wave."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from liuzhou_b200 import native  # noqa: E402
from liuzhou_b200.net import ChessNet, InferenceNet  # noqa: E402
from liuzhou_b200.tree_search import TreeMCTS, TreeMCTSConfig  # noqa: E402

total = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
waves = int(sys.argv[2]) if len(sys.argv) > 2 else 100
dev = torch.device("cuda:0")
torch.manual_seed(20260314)
model = ChessNet()
net = InferenceNet(model, dev)


def time_streams(fns, streams, reps):
    """fns[i]() is enqueued reps times on streams[i]; returns ms per rep (all streams together)."""
    cur = torch.cuda.current_stream(dev)
    for s in streams:
        s.wait_stream(cur)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(cur)
    for s in streams:
        s.wait_stream(cur)
    for _ in range(reps):
        for fn, s in zip(fns, streams):
            with torch.cuda.stream(s):
                fn()
    for s in streams:
        cur.wait_stream(s)
    e1.record(cur)
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


# ---- forward only -----------------------------------------------------------------------------------------
for groups in (1, 2, 4):
    n = total // groups
    nets = [InferenceNet(model, dev) for _ in range(groups)]
    graphs = []
    for nt in nets:
        x = nt.new_input(n)
        x.copy_((torch.rand_like(x.float()) > 0.5).to(x.dtype))
        nt.capture(n, x)
        graphs.append(nt._graphs[n][0])
    streams = [torch.cuda.Stream(device=dev) for _ in range(groups)]
    fns = [g.replay for g in graphs]
    time_streams(fns, streams, 5)
    ms = time_streams(fns, streams, 50)
    print(f"forward: {groups} group(s) x {n}: {ms:.3f} ms per {total} states "
          f"({total * net.flops_per_state / ms / 1e9:.0f} TFLOP/s)", flush=True)
    del nets, graphs

# ---- full wave: select -> encode -> network -> expand/backup ----------------------------------------------
for groups in (1, 2, 4):
    n = total // groups
    searchers = []
    for gi in range(groups):
        pb = native.PlayoutBatch(n, seed=20260314, device=dev, game_offset=gi * n)
        pb.run(max_steps=30)
        st = pb.packed.clone()
        m = TreeMCTS(InferenceNet(model, dev), n, TreeMCTSConfig(num_simulations=waves, add_dirichlet_noise=False), dev)
        m.tree.reset(st, None)
        m._capture()
        m.tree.reset(st, None)
        m._root_graph.replay()
        m._first_graph.replay()          # the wave graph is [network, expand + next select]: it needs a select first
        searchers.append(m)
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(device=dev) for _ in range(groups)]
    fns = [m._wave_graph.replay for m in searchers]
    time_streams(fns, streams, 10)
    ms = time_streams(fns, streams, waves - 10)
    print(f"wave: {groups} group(s) x {n}: {ms:.3f} ms per wave of {total} leaves "
          f"-> {total / ms / 200 * 1e3:.0f} positions/s at 200 sims", flush=True)
    del searchers
