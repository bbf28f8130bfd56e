"""`ChessNet` pinned to the REFERENCE's module (src/neural_network.py:213-259): golden fp32 forward of the reference's own
class on 256 fixed positions (tests/golden/make_chessnet_golden.py), default init (seed 20260314) and with non-trivial
BatchNorm statistics.

CPU : our re-declaration of the module (liuzhou_b200/net.py, fp32) reproduces the reference's outputs (same seed ->
      same initial weights -> same numbers): architecture, parameter order and initialisation are the reference's.
GPU : the product path -- bf16, every convolution on the tcgen05 kernel, fused heads -- against the same golden with a
      stated per-head tolerance; the measured max-abs errors are printed (pytest -s)."""
import numpy as np
import pytest
import torch

import oracle
from tests._util import GOLDEN

# tolerances of the bf16 product path against the fp32 reference (22 convolutions deep, bf16 operands, fp32 accumulation
# and fp32 residual stream in the fused trunk kernel).  Measured on B200: <= 4.6e-5 on the policy probabilities and
# <= 2.3e-5 on the value for both weight sets; the bounds below are ~20x that.
POLICY_PROB_ATOL = 1e-3        # max |softmax prob - reference prob| per head entry
VALUE_ATOL = 1e-3              # |bucket expectation - reference| (values live in [-1, 1])


def _golden():
    z = np.load(GOLDEN / "chessnet_forward.npz")
    n = z["board"].shape[0]
    st = {}
    for k in oracle.STATE_FIELDS:
        a = z[k]
        if k in ("marks_black", "marks_white"):
            st[k] = np.unpackbits(a, axis=1)[:, :36].reshape(n, 6, 6).astype(bool)
        elif k == "board":
            st[k] = a.reshape(n, 6, 6).astype(np.int8)
        else:
            st[k] = a.astype(np.int64)
    return z, st


def _randomize_bn(model, seed=7):
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            c = m.num_features
            m.running_mean.copy_(torch.randn(c, generator=g) * 0.2)
            m.running_var.copy_(torch.rand(c, generator=g) * 0.8 + 0.6)
            m.weight.data.copy_(torch.rand(c, generator=g) * 0.6 + 0.7)
            m.bias.data.copy_(torch.randn(c, generator=g) * 0.1)


def _value(logits):
    from liuzhou_b200.net import bucket_logits_to_scalar

    return bucket_logits_to_scalar(torch.as_tensor(logits).float())


def test_module_declaration_matches_reference_outputs_fp32():
    from liuzhou_b200.net import ChessNet

    z, st = _golden()
    x = torch.from_numpy(oracle.states_to_model_input(st))
    torch.manual_seed(20260314)
    model = ChessNet().eval()
    with torch.no_grad():
        for tag in ("init", "bn"):
            if tag == "bn":
                _randomize_bn(model)
            outs = model(x)
            for name, o in zip(("log_p1", "log_p2", "log_pmc", "value_logits"), outs):
                np.testing.assert_allclose(o.numpy(), z[f"{tag}_{name}"], rtol=0, atol=2e-5, err_msg=f"{tag}_{name}")


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["init", "bn"])
def test_bf16_tcgen05_forward_vs_reference_golden(tag):
    from liuzhou_b200 import _lib
    from liuzhou_b200.net import ChessNet, InferenceNet

    z, st = _golden()
    torch.manual_seed(20260314)
    model = ChessNet().eval()
    if tag == "bn":
        with torch.no_grad():
            _randomize_bn(model)
    net = InferenceNet(model, "cuda:0")
    assert not net.library_convs
    x = torch.from_numpy(oracle.states_to_model_input(st)).cuda()
    c0 = _lib.launch_count()
    outs = [o.float().cpu() for o in net.forward(x)]           # public path: pads 256 -> 256 rows, 11 -> 64 channels
    assert _lib.launch_count() - c0 >= 2                        # the fused trunk kernel + the heads tail, both ours
    worst = {}
    for name, o in zip(("log_p1", "log_p2", "log_pmc"), outs[:3]):
        ref = torch.from_numpy(z[f"{tag}_{name}"])
        worst[name] = float((o.exp() - ref.exp()).abs().max())
        assert worst[name] <= POLICY_PROB_ATOL, (name, worst[name])
        assert torch.allclose(o.exp().sum(1), torch.ones(o.size(0)), atol=1e-3)
    worst["value"] = float((_value(outs[3]) - _value(z[f"{tag}_value_logits"])).abs().max())
    assert worst["value"] <= VALUE_ATOL, worst
    print(f"[chessnet golden {tag}] max-abs error of the bf16 tcgen05 path vs the reference fp32 module: {worst}")
