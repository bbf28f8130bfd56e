#!/usr/bin/env python3
"""Golden forward pass of the REFERENCE's network module.

    python tests/golden/make_chessnet_golden.py      ->  tests/golden/chessnet_forward.npz

`/root/reference/src/neural_network.py::ChessNet` (default architecture, fp32, eval) is built with
`torch.manual_seed(20260314)` (the stable-init seed of scripts/big_train_v1.sh:24) and evaluated on 256 fixed positions
(every 5th state of 10 seeded uniform-random playouts: all phases); a second pass uses the same weights with
non-trivial BatchNorm statistics (drawn from a seeded CPU generator in module order, so any ChessNet with the same
parameter names reproduces them).  Stored: the 256 states (reference tensor layout), the three log-softmax policy heads
and the 101 value-bucket logits of both passes.  The product's bf16 tcgen05 path is compared against these on the GPU
(tests/test_gpu_conv.py) -- a check against the reference's module, not against our own re-declaration of it."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402  (drives the playouts and encodes the planes; the network outputs stored are the REFERENCE's)

sys.path.insert(0, "/root/reference")
from src.neural_network import ChessNet  # noqa: E402

SEED, N = 20260314, 256
OUT = Path(__file__).resolve().parent / "chessnet_forward.npz"


def positions():
    states = []
    for g in range(10):
        st = oracle.initial_states(1)
        for i, a in enumerate(oracle.random_playout(SEED, g, 512, want_trace=True)["trace"]):
            if i % 5 == g % 5:
                states.append(st)
            st = oracle.apply_move_scalar(st, int(a))
    st = {k: np.concatenate([s[k] for s in states])[:N] for k in oracle.STATE_FIELDS}
    assert st["board"].shape[0] == N
    return st


def randomize_bn(model, seed=7):
    """Deterministic non-trivial eval-mode BatchNorm statistics, module order, CPU generator."""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            c = m.num_features
            m.running_mean.copy_(torch.randn(c, generator=g) * 0.2)
            m.running_var.copy_(torch.rand(c, generator=g) * 0.8 + 0.6)
            m.weight.data.copy_(torch.rand(c, generator=g) * 0.6 + 0.7)
            m.bias.data.copy_(torch.randn(c, generator=g) * 0.1)


def main():
    st = positions()
    x = torch.from_numpy(oracle.states_to_model_input(st))
    torch.manual_seed(SEED)
    model = ChessNet().eval()
    out = {}
    with torch.no_grad():
        for tag in ("init", "bn"):
            if tag == "bn":
                randomize_bn(model)
            lp1, lp2, lpm, vl = model(x)
            out[f"{tag}_log_p1"], out[f"{tag}_log_p2"], out[f"{tag}_log_pmc"], out[f"{tag}_value_logits"] = (
                t.numpy().astype(np.float32) for t in (lp1, lp2, lpm, vl))
    packed = {k: (np.packbits(st[k].reshape(N, 36), axis=1) if st[k].dtype == np.bool_ else st[k].astype(np.int8))
              for k in oracle.STATE_FIELDS}
    np.savez_compressed(OUT, **packed, **out)
    print(OUT, OUT.stat().st_size, "bytes; phases:", np.unique(st["phase"], return_counts=True))


if __name__ == "__main__":
    main()
