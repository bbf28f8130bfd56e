"""GPU parity of the device tree MCTS (select / expand / backup kernels) vs the oracle's restatement of the
reference's PortableTreeBatch and vs the golden vectors produced by the reference binary.

Bar (north_star): given identical network outputs, IDENTICAL visit counts; Q / root values within 1e-5 rel
(they are in fact bit-identical here: fp64 tree statistics rounded to fp32 on output, like the reference)."""
import numpy as np
import pytest
import torch

import oracle
from tests._util import STATE_FIELDS, concat_states, fake_net, golden_states, load_golden, to_torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _np(t):
    return t.detach().cpu().numpy()


def _playout_states(n_games, seed, every):
    out = []
    for g in range(n_games):
        trace = oracle.random_playout(seed, g, 512, want_trace=True)["trace"]
        st = oracle.initial_states(1)
        for ply, a in enumerate(trace):
            if ply % every == 0:
                out.append(st)
            st = oracle.apply_move_scalar(st, int(a))
        out.append(st)   # terminal root as well
    return concat_states(out)


def _pending_from_device(tree, native):
    """(tree indices, model inputs, legal masks) of the LEAF_EVAL slots, in tree order."""
    status = _np(tree.pending_status)
    rows = np.nonzero(status == 0)[0]
    inputs = _np(tree.pending_inputs("f32_nchw"))[rows]
    words, _ = native.legal_masks(tree.pending_states, scalar_semantics=True)
    masks = _np(native.mask_words_to_bool(words))[rows].astype(np.uint8)
    return rows, inputs, masks


def _complete(tree, rows, pri, val):
    slots = tree.pending_status.numel()
    p = np.zeros((slots, 220), np.float32)
    v = np.zeros((slots,), np.float32)
    p[rows] = pri
    v[rows] = val
    tree.complete_pending(torch.from_numpy(p).to(DEV), torch.from_numpy(v).to(DEV))


def _run_both(st, sims, c_puct, check_every_wave=True, net=None):
    fake = net or fake_net
    from liuzhou_b200 import native
    from liuzhou_b200.tree import DeviceTreeBatch

    n = st["board"].shape[0]
    ref = oracle.TreeBatch(st, c_puct)
    tree = DeviceTreeBatch(n, DEV, exploration_weight=c_puct, nodes_per_tree_hint=(sims + 2) * 64)
    tree.reset(native.pack_states(to_torch(st, DEV)))
    pend = ref.prepare_roots()
    tree.prepare_roots()
    rows, inputs, masks = _pending_from_device(tree, native)
    assert np.array_equal(rows, pend["tree_indices"])
    assert np.array_equal(inputs, pend["model_inputs"]) and np.array_equal(masks, pend["legal_masks"])
    pri, val = fake(pend["model_inputs"], pend["legal_masks"], 0)
    ref.complete_pending(pri, val)
    _complete(tree, rows, pri, val)
    for s in range(sims):
        pend = ref.select_leaves()
        tree.select_leaves()
        rows, inputs, masks = _pending_from_device(tree, native)
        if check_every_wave:
            assert np.array_equal(rows, pend["tree_indices"]), s
            assert np.array_equal(inputs, pend["model_inputs"]), s
            assert np.array_equal(masks, pend["legal_masks"]), s
        pri, val = fake(inputs, masks, 1)
        ref.complete_pending(*fake(pend["model_inputs"], pend["legal_masks"], 1))
        _complete(tree, rows, pri, val)
    tree.check_capacity()
    return ref.root_outputs(), ref.root_priors(), tree.root_outputs()


@pytest.mark.parametrize("sims,c_puct", [(64, 1.0), (200, 1.5)])
def test_tree_visit_counts_identical_to_oracle(sims, c_puct):
    st = _playout_states(4, 77, every=11)
    ro, rp, mo = _run_both(st, sims, c_puct)
    assert np.array_equal(_np(mo["visit_counts"]), ro["visit_counts"])          # identical visit counts
    assert np.array_equal(_np(mo["legal_masks"]).astype(np.uint8), ro["legal_masks"])
    assert np.array_equal(_np(mo["root_action_values"]), ro["root_action_values"])
    assert np.array_equal(_np(mo["root_values"]), ro["root_values"])
    assert np.array_equal(_np(mo["terminal"]).astype(np.uint8), ro["terminal"])
    assert np.array_equal(_np(mo["root_priors"]), rp["priors"])
    total = _np(mo["visit_counts"]).sum(1)
    live = ro["terminal"] == 0
    assert (total[live] == sims).all()


def test_tree_deep_paths_beyond_recorded_depth():
    """A network that puts (almost) all prior mass on the lowest legal action and returns value 0 makes every
    simulation descend the same line one level deeper: depth reaches the number of simulations, far beyond the 32
    levels the select kernel records for the one-round-trip backup -- the parent-walk fallback and the hand-over
    between the two must leave visit counts / values identical to the oracle."""
    def line_net(inputs, masks, salt):
        n = masks.shape[0]
        pri = np.where(masks != 0, 1e-7, 0.0).astype(np.float32)
        first = (masks != 0).argmax(1)
        pri[np.arange(n), first] = 1.0
        return pri, np.zeros((n,), np.float32)

    st = _playout_states(3, 5, every=40)
    sims = 90
    ro, rp, mo = _run_both(st, sims, 1.0, net=line_net)
    v = _np(mo["visit_counts"])
    assert np.array_equal(v, ro["visit_counts"])
    assert np.array_equal(_np(mo["root_action_values"]), ro["root_action_values"])
    assert np.array_equal(_np(mo["root_values"]), ro["root_values"])
    live = ro["terminal"] == 0
    assert (v.max(1)[live] >= sims - 2).all()          # one child took (almost) every visit: the line is really deep


def test_tree_golden_reference_vectors():
    """Visit counts / Q / root values produced by the reference's own _liuzhou_portable_cpp binary."""
    from liuzhou_b200 import native
    from liuzhou_b200.tree import DeviceTreeBatch

    z = load_golden("tree_mcts")
    st = golden_states(z)
    n = st["board"].shape[0]
    tree = DeviceTreeBatch(n, DEV, exploration_weight=float(z["c"]), nodes_per_tree_hint=100 * 64)
    tree.reset(native.pack_states(to_torch(st, DEV)))
    tree.prepare_roots()
    rows, inputs, masks = _pending_from_device(tree, native)
    _complete(tree, rows, *fake_net(inputs, masks, 0))
    for _ in range(int(z["sims"])):
        tree.select_leaves()
        rows, inputs, masks = _pending_from_device(tree, native)
        _complete(tree, rows, *fake_net(inputs, masks, 1))
    out = tree.root_outputs()
    assert np.array_equal(_np(out["visit_counts"]), z["visit_counts"])
    assert np.array_equal(_np(out["root_action_values"]), z["root_action_values"])
    assert np.array_equal(_np(out["root_values"]), z["root_values"])
    assert np.array_equal(_np(out["terminal"]).astype(np.uint8), z["terminal"])
    assert np.array_equal(_np(out["root_priors"]), z["root_priors"])


def test_tree_set_root_priors_and_inactive():
    from liuzhou_b200 import native
    from liuzhou_b200.tree import DeviceTreeBatch

    st = _playout_states(2, 5, every=17)
    n = st["board"].shape[0]
    ref = oracle.TreeBatch(st, 1.0)
    ref.deactivate([1])
    active = np.ones((n,), bool)
    active[1] = False
    tree = DeviceTreeBatch(n, DEV, exploration_weight=1.0, nodes_per_tree_hint=40 * 64)
    tree.reset(native.pack_states(to_torch(st, DEV)), torch.from_numpy(active).to(DEV))
    pend = ref.prepare_roots()
    tree.prepare_roots()
    rows, inputs, masks = _pending_from_device(tree, native)
    assert np.array_equal(rows, pend["tree_indices"])
    pri, val = fake_net(inputs, masks, 0)
    ref.complete_pending(pri, val)
    _complete(tree, rows, pri, val)
    rng = np.random.default_rng(0)
    noisy = ref.root_priors()["priors"] * 0.75 + 0.25 * rng.random((n, 220)).astype(np.float32) * ref.root_priors()["legal_masks"]
    noisy = noisy.astype(np.float32)
    ref.set_root_priors(noisy)
    tree.set_root_priors(torch.from_numpy(noisy).to(DEV))
    for _ in range(32):
        pend = ref.select_leaves()
        tree.select_leaves()
        rows, inputs, masks = _pending_from_device(tree, native)
        assert np.array_equal(rows, pend["tree_indices"])
        pri, val = fake_net(inputs, masks, 1)
        ref.complete_pending(pri, val)
        _complete(tree, rows, pri, val)
    out = tree.root_outputs()
    assert np.array_equal(_np(out["visit_counts"]), ref.root_outputs()["visit_counts"])
    assert np.array_equal(_np(out["root_priors"]), ref.root_priors()["priors"])
    assert int(_np(out["visit_counts"])[1].sum()) == 0


def test_tree_multi_leaf_virtual_loss_properties():
    """K > 1 has no deterministic reference (SURVEY hard parts): check invariants instead -- every wave adds at
    most K visits per tree, total root visits == sum of child visits, virtual loss fully reverted, run-to-run
    determinism."""
    from liuzhou_b200 import native
    from liuzhou_b200.tree import DeviceTreeBatch

    st = _playout_states(3, 9, every=13)
    n = st["board"].shape[0]
    outs = []
    for _rep in range(2):
        tree = DeviceTreeBatch(n, DEV, exploration_weight=1.0, leaves_per_wave=4, virtual_loss=1.0,
                               nodes_per_tree_hint=200 * 64)
        tree.reset(native.pack_states(to_torch(st, DEV)))
        tree.prepare_roots()
        rows, inputs, masks = _pending_from_device(tree, native)
        _complete(tree, rows, *fake_net(inputs, masks, 0))
        for _ in range(25):
            tree.select_leaves()
            rows, inputs, masks = _pending_from_device(tree, native)
            _complete(tree, rows, *fake_net(inputs, masks, 1))
        out = tree.root_outputs()
        visits = _np(out["visit_counts"])
        root_n = _np(tree.visit[:n])
        assert (visits.sum(1) == root_n).all()                  # no virtual visits left behind
        assert (root_n <= 25 * 4).all()
        live = ~_np(out["terminal"])
        assert (root_n[live] >= 25).all()
        assert (_np(tree.info) & (1 << 20))[: tree.stats()["nodes_used"]].sum() == 0   # no pending flags left
        outs.append((visits, _np(out["root_action_values"])))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


def test_encode_inputs_and_heads_to_priors():
    from liuzhou_b200 import native
    from liuzhou_b200.tree import encode_inputs, heads_to_priors

    st = _playout_states(4, 3, every=1)
    n = st["board"].shape[0]
    packed = native.pack_states(to_torch(st, DEV))
    exp = oracle.states_to_model_input(st)
    assert np.array_equal(_np(encode_inputs(packed, "f32_nchw")), exp)
    bf = encode_inputs(packed, "bf16_nhwc")
    assert bf.dtype == torch.bfloat16 and bf.is_contiguous(memory_format=torch.channels_last)
    assert np.array_equal(_np(bf.float()), exp)
    rng = np.random.default_rng(1)
    heads = [torch.log_softmax(torch.from_numpy(rng.standard_normal((n, 36)).astype(np.float32)), 1) for _ in range(3)]
    logits = torch.from_numpy(rng.standard_normal((n, 101)).astype(np.float32))
    pri, val = heads_to_priors(packed, *(h.to(DEV) for h in heads), logits.to(DEV))
    legal = np.zeros((n, 220), bool)
    for i in range(n):
        legal[i, oracle.legal_actions(st, i)[0]] = True
    op, _ = oracle.project_policy_logits_fast(*(h.numpy() for h in heads), legal)
    np.testing.assert_allclose(_np(pri), op, rtol=1e-5, atol=1e-7)
    probs = torch.softmax(logits, 1)
    ev = (probs * torch.linspace(-1.0, 1.0, 101)).sum(1).numpy()   # neural_network.py:201-210
    np.testing.assert_allclose(_np(val), ev, rtol=1e-5, atol=1e-6)
