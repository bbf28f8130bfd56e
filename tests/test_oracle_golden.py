"""CPU oracle vs the committed golden vectors (tests/golden/*.npz, produced from the reference's own
binaries by tests/golden/make_golden.py).  Needs neither /root/reference nor oracle/_ref nor a GPU."""
import numpy as np

import oracle
from tests._util import STATE_FIELDS, fake_net, golden_states, load_golden


def test_encode_actions_golden():
    z = load_golden("encode_actions")
    st = golden_states(z)
    for aux in (1, 4):
        mask, meta = oracle.encode_actions_fast(st, 36, 144, 36, aux)
        total = 216 + aux
        ref_mask = np.unpackbits(z[f"mask_aux{aux}"], axis=1)[:, :total].astype(bool)
        assert np.array_equal(mask, ref_mask)
        assert np.array_equal(meta, z[f"meta_aux{aux}"].astype(np.int32))


def test_apply_moves_golden():
    z = load_golden("apply_moves")
    out = oracle.batch_apply_moves(golden_states(z, "in_"), z["codes"], z["parents"])
    for k in STATE_FIELDS:
        assert np.array_equal(np.asarray(out[k]).reshape(z[f"out_{k}"].shape), z[f"out_{k}"]), k
    assert z["codes"].shape[0] > 5000


def test_scalar_playouts_golden():
    z = load_golden("scalar_playouts")
    st = golden_states(z)
    n = st["board"].shape[0]
    for i in range(n):
        legal = list(z["legal_idx"][z["legal_ptr"][i]:z["legal_ptr"][i + 1]])
        assert oracle.legal_actions(st, i)[0] == legal
        assert oracle.is_game_over(st, i) == bool(z["over"][i])
        assert oracle.winner(st, i) == int(z["winner"][i])
    # transitions: state i --chosen[i]--> state i+1 within a game
    gp = z["game_ptr"]
    for g in range(len(gp) - 1):
        for i in range(gp[g], gp[g + 1] - 1):
            nxt = oracle.apply_move_scalar(st, int(z["chosen"][i]), i)
            for k in STATE_FIELDS:
                assert np.array_equal(np.asarray(nxt[k])[0], st[k][i + 1]), (g, i, k)
        assert z["chosen"][gp[g + 1] - 1] == -1


def test_root_puct_golden():
    z = load_golden("root_puct")
    for tag in "abc":
        v, w, rv = oracle.root_puct_allocate_visits(z[f"{tag}_priors"], z[f"{tag}_leaf"], z[f"{tag}_valid"],
                                                    int(z[f"{tag}_sims"]), float(z[f"{tag}_c"]))
        assert np.array_equal(v, z[f"{tag}_visits"])
        assert np.array_equal(w, z[f"{tag}_value_sum"])
        np.testing.assert_allclose(rv, z[f"{tag}_root_values"], rtol=1e-5, atol=1e-6)


def test_tree_mcts_golden():
    z = load_golden("tree_mcts")
    tb = oracle.TreeBatch(golden_states(z), float(z["c"]))
    pend = tb.prepare_roots()
    tb.complete_pending(*fake_net(pend["model_inputs"], pend["legal_masks"], 0))
    for _ in range(int(z["sims"])):
        pend = tb.select_leaves()
        tb.complete_pending(*fake_net(pend["model_inputs"], pend["legal_masks"], 1))
    ro = tb.root_outputs()
    assert np.array_equal(ro["visit_counts"], z["visit_counts"])
    assert np.array_equal(ro["root_action_values"], z["root_action_values"])
    assert np.array_equal(ro["root_values"], z["root_values"])
    assert np.array_equal(ro["terminal"], z["terminal"])
    assert np.array_equal(tb.root_priors()["priors"], z["root_priors"])


def test_composites_golden():
    z = load_golden("composites")
    st = golden_states(z)
    mask, meta = z["mask"], z["meta"].astype(np.int32)
    got = oracle.root_pack_sparse_actions(mask, z["probs"], meta)
    for i, g in enumerate(got):
        if i == 5:
            np.testing.assert_allclose(g, z[f"pack{i}"], rtol=1e-6, atol=1e-7)
        else:
            assert np.array_equal(g, z[f"pack{i}"]), i
    fin = oracle.root_finalize_from_visits(z["pack4"], z["pack6"], z["pack3"], z["visits"], z["value_sum"], z["pack1"],
                                           mask.shape[0], 220, z["temps"])
    np.testing.assert_allclose(fin[0], z["fin0"], rtol=1e-5, atol=1e-7)
    for i in (1, 2, 3):
        assert np.array_equal(fin[i], z[f"fin{i}"])
    np.testing.assert_allclose(fin[4], z["fin4"], rtol=1e-5, atol=1e-6)
    mi = oracle.states_to_model_input(st)
    assert np.array_equal(np.packbits(mi.astype(np.uint8)), z["model_input"])
    p, _l = oracle.project_policy_logits_fast(z["head0"], z["head1"], z["head2"], mask)
    np.testing.assert_allclose(p, z["proj_probs"], rtol=1e-5, atol=1e-7)
    o_state = {k: np.array(v, copy=True) for k, v in st.items()}
    plies, done = z["step_plies_in"].copy(), z["step_done_in"].copy()
    out = oracle.self_play_step_inplace(o_state, plies, done, z["step_active"], z["step_codes"], z["step_terminal"],
                                        z["step_valid"], 130, 2.0)
    assert np.array_equal(out[0], z["step_out0"]) and np.array_equal(out[1], z["step_out1"])
    np.testing.assert_allclose(out[2], z["step_out2"], rtol=1e-6, atol=1e-7)
    for k in STATE_FIELDS:
        assert np.array_equal(np.asarray(o_state[k]).reshape(z[f"step_state_{k}"].shape), z[f"step_state_{k}"]), k
    assert np.array_equal(plies, z["step_plies_out"]) and np.array_equal(done, z["step_done_out"])


def test_playout_statistics():
    """Size-independent sanity of the config-2 workload on the oracle: mean length ~127 plies, ~93% draws
    (measured on the reference's engines)."""
    n, plies, draws = 400, 0, 0
    for g in range(n):
        r = oracle.random_playout(20260314, g)
        plies += r["plies"]
        draws += r["result"] == 0
        assert r["result"] in (-1, 0, 1)
    assert 120 < plies / n < 135
    assert draws / n > 0.85
