"""Rule-engine parity against the reference's LEGACY python engine (src/rule_engine.py + src/move_generator.py): 6,029
states / 5,981 transitions of 48 uniformly random games recorded by tests/golden/make_legacy_golden.py (sampling
convention of the reference's tests/v0/test_actions.py, seed 0x7777).  This is a pin that is independent of the v0
C++/CUDA engine behind the other golden files: legal sets, every atomic transition (placement -> marking -> removal
-> forced removal -> movement -> capture / no-move removal / counter removal), terminal flags and winners must be
bit-exact for the oracle (CPU) and for the CUDA kernels (drop-in byte layout and packed bitboards)."""
import numpy as np
import pytest

import oracle
from tests._util import STATE_FIELDS, load_golden, to_torch

_SCALARS = ("phase", "current_player", "pending_marks_required", "pending_marks_remaining", "pending_captures_required",
            "pending_captures_remaining", "forced_removals_done", "move_count", "moves_since_capture")


@pytest.fixture(scope="module")
def legacy():
    z = load_golden("legacy_playouts")
    n = z["board"].shape[0]
    st = {"board": z["board"].astype(np.int8),
          "marks_black": np.unpackbits(z["marks_black"], axis=1)[:, :36].reshape(n, 6, 6).astype(np.bool_),
          "marks_white": np.unpackbits(z["marks_white"], axis=1)[:, :36].reshape(n, 6, 6).astype(np.bool_)}
    for j, name in enumerate(_SCALARS):
        st[name] = z["scalars"][:, j].astype(np.int64)
    legal = np.zeros((n, 220), np.bool_)
    ptr = z["legal_ptr"]
    for i in range(n):
        legal[i, z["legal_idx"][ptr[i]:ptr[i + 1]]] = True
    chosen = z["chosen"].astype(np.int64)
    has_next = chosen >= 0                               # row i + 1 is the successor of row i
    return {"st": st, "legal": legal, "chosen": chosen, "has_next": has_next, "over": z["over"], "winner": z["winner"],
            "n": n}


def _rows(st, idx):
    return {k: np.ascontiguousarray(st[k][idx]) for k in STATE_FIELDS}


def test_golden_covers_every_phase_and_outcome(legacy):
    st, n = legacy["st"], legacy["n"]
    assert n >= 6000 and int(legacy["has_next"].sum()) >= 5900
    assert set(np.unique(st["phase"])) == {1, 2, 3, 4, 5, 6, 7}
    assert legacy["over"].sum() >= 40 and {int(w) for w in np.unique(legacy["winner"])} >= {0}
    assert (st["pending_captures_remaining"] > 0).any() and (st["pending_marks_remaining"] > 1).any()


def test_oracle_tensor_ops_match_legacy_engine(legacy):
    st, legal, chosen = legacy["st"], legacy["legal"], legacy["chosen"]
    live = ~legacy["over"]
    mask, meta = oracle.encode_actions_fast(st)
    assert np.array_equal(mask[live], legal[live])                     # the mask ignores game-over by design
    rows = np.nonzero(legacy["has_next"])[0]
    codes = meta[rows, chosen[rows]]
    nxt, applied = oracle.batch_apply_moves(st, codes, rows, return_applied=True)
    assert applied.all()
    want = _rows(st, rows + 1)
    for k in STATE_FIELDS:
        assert np.array_equal(np.asarray(nxt[k]).reshape(want[k].shape), want[k]), k


def test_oracle_scalar_engine_matches_legacy_engine(legacy):
    st, legal, chosen = legacy["st"], legacy["legal"], legacy["chosen"]
    for i in range(0, legacy["n"], 3):                                  # every third state keeps the CPU suite short
        assert oracle.is_game_over(st, i) == bool(legacy["over"][i]), i
        assert oracle.winner(st, i) == int(legacy["winner"][i]), i
        if legacy["over"][i]:
            continue
        idx, _ = oracle.legal_actions(st, i)
        assert sorted(idx) == list(np.nonzero(legal[i])[0]), i
        if chosen[i] >= 0:
            got = oracle.apply_move_scalar(st, int(chosen[i]), i)
            for k in STATE_FIELDS:
                assert np.array_equal(np.asarray(got[k]).reshape(-1), np.asarray(st[k][i + 1]).reshape(-1)), (i, k)


@pytest.mark.gpu
def test_cuda_kernels_match_legacy_engine(legacy):
    import torch

    from liuzhou_b200 import native, v0_core
    from liuzhou_b200.engine import packed_status

    dev = "cuda:0"
    st, legal, chosen = legacy["st"], legacy["legal"], legacy["chosen"]
    live = ~legacy["over"]
    rows = np.nonzero(legacy["has_next"])[0]
    t = to_torch(st, dev)
    # drop-in ops on the reference byte layout
    mask, meta = v0_core.encode_actions_fast(*t[:10], 36, 144, 36, 4)
    assert np.array_equal(mask.cpu().numpy()[live], legal[live])
    r = torch.from_numpy(rows).to(dev)
    codes = meta[r, torch.from_numpy(chosen[rows]).to(dev)]
    out = v0_core.batch_apply_moves(*t, codes, r)
    want = _rows(st, rows + 1)
    for k, o in zip(STATE_FIELDS, out):
        assert np.array_equal(o.cpu().numpy().reshape(want[k].shape), want[k]), k
    # native packed bitboards: legal words, apply by action index, terminal status
    packed = native.pack_states(t)
    words, counts = native.legal_masks(packed, scalar_semantics=True)
    got = native.mask_words_to_bool(words).cpu().numpy()
    assert np.array_equal(got[live], legal[live]) and np.array_equal(counts.cpu().numpy()[live], legal[live].sum(1))
    nxt = native.apply_actions(packed, torch.from_numpy(chosen[rows]).to(dev), parent_indices=r)
    assert torch.equal(nxt, native.pack_states(to_torch(want, dev)))
    over, winner = packed_status(packed)
    assert np.array_equal(over.cpu().numpy(), legacy["over"]) and np.array_equal(winner.cpu().numpy(), legacy["winner"])
