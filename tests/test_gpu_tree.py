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


def _search_both(ref, tree, native, sims, move):
    """prepare_roots + `sims` waves on both; every pending batch (tree indices, inputs, masks) must agree."""
    pend = ref.prepare_roots()
    tree.prepare_roots()
    rows, inputs, masks = _pending_from_device(tree, native)
    assert np.array_equal(rows, pend["tree_indices"]), move
    assert np.array_equal(inputs, pend["model_inputs"]) and np.array_equal(masks, pend["legal_masks"]), move
    pri, val = fake_net(pend["model_inputs"], pend["legal_masks"], 0)
    ref.complete_pending(pri, val)
    _complete(tree, rows, pri, val)
    for s in range(sims):
        pend = ref.select_leaves()
        tree.select_leaves()
        rows, inputs, masks = _pending_from_device(tree, native)
        assert np.array_equal(rows, pend["tree_indices"]), (move, s)
        assert np.array_equal(inputs, pend["model_inputs"]), (move, s)
        pri, val = fake_net(inputs, masks, 1)
        ref.complete_pending(pri, val)
        _complete(tree, rows, pri, val)


def _assert_root_outputs_equal(ref, tree):
    ro, mo = ref.root_outputs(), tree.root_outputs()
    assert np.array_equal(_np(mo["visit_counts"]), ro["visit_counts"])
    assert np.array_equal(_np(mo["legal_masks"]).astype(np.uint8), ro["legal_masks"])
    assert np.array_equal(_np(mo["root_action_values"]), ro["root_action_values"])
    assert np.array_equal(_np(mo["root_values"]), ro["root_values"])
    assert np.array_equal(_np(mo["terminal"]).astype(np.uint8), ro["terminal"])
    assert np.array_equal(_np(mo["root_priors"]), ref.root_priors()["priors"])
    return ro


@pytest.mark.parametrize("sims,c_puct,pick", [(64, 1.0, "max"), (120, 1.5, "second")])
def test_tree_advance_roots_subtree_reuse_identical_to_oracle(sims, c_puct, pick):
    """advance_roots (portable_mcts.cpp:739-768): after each searched move the played child becomes the root and
    keeps its subtree; the next search starts from those statistics.  Visit counts / Q / priors / root values and
    every pending batch stay identical to the oracle over 5 consecutive moves (the oracle itself is pinned against
    the reference binary on exactly this protocol, tests/test_oracle_vs_reference.py)."""
    from liuzhou_b200 import native
    from liuzhou_b200.tree import DeviceTreeBatch

    st = _playout_states(4, 77, every=11)
    n = st["board"].shape[0]
    ref = oracle.TreeBatch(st, c_puct)
    tree = DeviceTreeBatch(n, DEV, exploration_weight=c_puct, nodes_per_tree_hint=(sims + 2) * 64 * 3)
    tree.reset(native.pack_states(to_torch(st, DEV)))
    for move in range(5):
        _search_both(ref, tree, native, sims, move)
        ro = _assert_root_outputs_equal(ref, tree)
        v = ro["visit_counts"].astype(np.int64)
        if pick == "second":                      # a less-visited (possibly unvisited -> unexpanded) child
            order = np.argsort(-(v * 1000 + ro["legal_masks"]), axis=1, kind="stable")
            cand = order[:, 1]
            ok = ro["legal_masks"][np.arange(n), cand] != 0
            actions = np.where(ok, cand, order[:, 0])
        else:
            actions = v.argmax(1)
        actions = np.where((ro["terminal"] != 0) | (ro["legal_masks"].sum(1) == 0), -1, actions).astype(np.int32)
        ref.advance_roots(actions)
        tree.advance_roots(torch.from_numpy(actions).to(DEV))
        tree.check_capacity()
        # the new roots' states agree with the oracle's
        for i in (0, n // 2, n - 1):
            want = native.pack_states(to_torch(ref.root_state(i), DEV))
            assert torch.equal(tree.root_states()[i], want[0])
        status = tree.root_status()
        assert status["game_over"].shape == (n,)
    # inherited statistics really were reused: live roots carry more visits than one search adds
    total = _np(tree.root_outputs()["visit_counts"]).sum(1)
    assert total.max() > 0


def test_tree_advance_roots_reset_deactivate_and_errors():
    from liuzhou_b200 import native
    from liuzhou_b200.tree import DeviceTreeBatch

    st = _playout_states(2, 5, every=17)
    n = st["board"].shape[0]
    sims = 40
    ref = oracle.TreeBatch(st, 1.0)
    tree = DeviceTreeBatch(n, DEV, exploration_weight=1.0, nodes_per_tree_hint=(sims + 2) * 64 * 3)
    packed0 = native.pack_states(to_torch(st, DEV))
    tree.reset(packed0)
    _search_both(ref, tree, native, sims, 0)
    ro = _assert_root_outputs_equal(ref, tree)
    actions = np.where(ro["terminal"] != 0, -1, ro["visit_counts"].argmax(1)).astype(np.int32)
    # tree 0 restarts from the initial position, tree 1 is deactivated, the rest advance
    reset_mask = np.zeros((n,), bool)
    reset_mask[0] = True
    reset_states = packed0.clone()
    reset_states[0] = native.init_states(1, DEV)[0]
    tree.advance_roots(torch.from_numpy(actions).to(DEV), reset_states, torch.from_numpy(reset_mask).to(DEV))
    tree.deactivate([1])
    ref.advance_roots(actions)
    ref.deactivate([1])
    fresh = oracle.TreeBatch(oracle.initial_states(1), 1.0)
    # second search: compare all trees but 0 with `ref`, tree 0 with a fresh oracle tree
    pend_ref = ref.prepare_roots()
    pend_fresh = fresh.prepare_roots()
    tree.prepare_roots()
    rows, inputs, masks = _pending_from_device(tree, native)
    want_rows = sorted([0] + [int(r) for r in pend_ref["tree_indices"] if r != 0])
    assert list(rows) == want_rows and 1 not in rows
    pri, val = fake_net(inputs, masks, 0)
    _complete(tree, rows, pri, val)
    ref.complete_pending(*fake_net(pend_ref["model_inputs"], pend_ref["legal_masks"], 0))
    fresh.complete_pending(*fake_net(pend_fresh["model_inputs"], pend_fresh["legal_masks"], 0))
    for _ in range(sims):
        pr, pf = ref.select_leaves(), fresh.select_leaves()
        tree.select_leaves()
        rows, inputs, masks = _pending_from_device(tree, native)
        _complete(tree, rows, *fake_net(inputs, masks, 1))
        ref.complete_pending(*fake_net(pr["model_inputs"], pr["legal_masks"], 1))
        fresh.complete_pending(*fake_net(pf["model_inputs"], pf["legal_masks"], 1))
    mo = tree.root_outputs()
    rv, fv = ref.root_outputs()["visit_counts"], fresh.root_outputs()["visit_counts"]
    got = _np(mo["visit_counts"])
    assert np.array_equal(got[0], fv[0])
    assert np.array_equal(got[2:], rv[2:])
    assert got[1].sum() == rv[1].sum()          # deactivated: nothing added after the advance
    tree.check_capacity()
    # an action that is not a child of the root is reported like the reference does (it throws)
    bad = np.full((n,), -1, np.int32)
    live = np.nonzero((_np(mo["terminal"]) == 0) & (np.arange(n) != 1))[0]
    illegal = int(np.nonzero(_np(mo["legal_masks"])[live[0]] == 0)[0][0])
    bad[live[0]] = illegal
    tree.advance_roots(torch.from_numpy(bad).to(DEV))
    with pytest.raises(RuntimeError, match="not a child"):
        tree.check_capacity()
    with pytest.raises(RuntimeError):
        tree.advance_roots(torch.zeros((n + 1,), dtype=torch.int32, device=DEV))


def test_select_with_fused_input_encoding():
    """select_leaves / prepare_roots with encode_out: the rows of the pending (status 0) slots equal the separate
    lzb_encode_inputs_packed launch bit for bit (channels 0..15 written, the zero padding left alone), other rows are left
    untouched, tree statistics are unaffected."""
    from liuzhou_b200 import native
    from liuzhou_b200.tree import DeviceTreeBatch, encode_inputs

    st = _playout_states(3, 21, every=9)
    n = st["board"].shape[0]
    packed = native.pack_states(to_torch(st, DEV))
    trees = [DeviceTreeBatch(n, DEV, exploration_weight=1.0, nodes_per_tree_hint=40 * 64) for _ in range(2)]
    sentinel = 0.25
    # contract of the fused encoding: channels 16..63 are zero padding that the caller zero-fills once and the kernels
    # never write; the sentinel sits in the 16 channels they do write
    buf = torch.zeros((n, 64, 6, 6), dtype=torch.bfloat16, device=DEV).contiguous(memory_format=torch.channels_last)
    buf[:, :16] = sentinel
    for t in trees:
        t.reset(packed)
    trees[0].prepare_roots()
    trees[1].prepare_roots(encode_out=buf)
    for wave in range(12):
        status = _np(trees[1].pending_status)
        assert np.array_equal(status, _np(trees[0].pending_status))
        want = encode_inputs(trees[1].pending_states, "bf16_nhwc",
                             out=torch.empty_like(buf).contiguous(memory_format=torch.channels_last))
        live = torch.from_numpy(status == 0).to(DEV)
        assert torch.equal(buf[live], want[live])
        if wave == 0 and (~live).any():
            assert bool((buf[~live][:, :16] == sentinel).all())         # untouched rows
            assert bool((buf[:, 16:] == 0).all())
        rows, inputs, masks = _pending_from_device(trees[1], native)
        pri, val = fake_net(inputs, masks, 1)
        for t in trees:
            _complete(t, rows, pri, val)
        trees[0].select_leaves()
        trees[1].select_leaves(encode_out=buf)
    a, b = trees[0].root_outputs(), trees[1].root_outputs()
    assert torch.equal(a["visit_counts"], b["visit_counts"]) and torch.equal(a["root_values"], b["root_values"])
    with pytest.raises(RuntimeError):
        trees[1].select_leaves(encode_out=torch.zeros((n, 11, 6, 6), dtype=torch.bfloat16, device=DEV))
