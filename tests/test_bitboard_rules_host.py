"""Host-side check of the product's bitboard rule engine (liuzhou_b200/csrc/lz_rules.cuh, the header the CUDA
kernels are built from) against the oracle, via a TEST-ONLY host build (tests/host_shim).  No GPU needed.

This is what lets us trust the bit logic before spending GPU time; the `-m gpu` tests then check the
kernels themselves through the C ABI."""
import ctypes
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle
from tests._util import (STATE_FIELDS, concat_states, random_apply_batch, random_mask_states,
                         sparse_random_states)

SHIM_DIR = Path(__file__).resolve().parent / "host_shim"
SCALARS = STATE_FIELDS[3:]


@pytest.fixture(scope="module")
def shim():
    src = SHIM_DIR / "lz_host_shim.cpp"
    lib = SHIM_DIR / "liblz_host_shim.so"
    hdr = SHIM_DIR.parents[1] / "liuzhou_b200" / "csrc" / "lz_rules.cuh"
    if not lib.exists() or lib.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w", str(src), "-o", str(lib)], check=True)
    return ctypes.CDLL(str(lib))


def _flat(st):
    board = np.ascontiguousarray(st["board"], np.int8).reshape(-1, 36).copy()
    mb = np.ascontiguousarray(st["marks_black"]).astype(np.uint8).reshape(-1, 36).copy()
    mw = np.ascontiguousarray(st["marks_white"]).astype(np.uint8).reshape(-1, 36).copy()
    sc = np.stack([np.asarray(st[k], np.int64) for k in SCALARS], 1).copy()
    return board, mb, mw, sc


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


def _legal(shim, st, scalar):
    board, mb, mw, sc = _flat(st)
    n = board.shape[0]
    mask = np.zeros((n, 220), np.uint8)
    meta = np.zeros((n, 220, 4), np.int32)
    counts = np.zeros((n,), np.int32)
    kth = np.zeros((n, 220), np.int32)
    rank = np.zeros((n, 220), np.int32)
    shim.hs_legal(ctypes.c_int64(n), _p(board), _p(mb), _p(mw), _p(sc), ctypes.c_int(scalar), _p(mask), _p(meta),
                  _p(counts), _p(kth), _p(rank))
    return mask.astype(bool), meta, counts, kth, rank


@pytest.mark.parametrize("sampler,seed", [(random_mask_states, 0xF00DCAFE), (sparse_random_states, 0x5EED),
                                          (sparse_random_states, 99)])
def test_legal_mask_tensor_semantics(shim, sampler, seed):
    st = sampler(10_000, seed)
    mask, meta, counts, kth, rank = _legal(shim, st, 0)
    o_mask, o_meta = oracle.encode_actions_fast(st)
    assert np.array_equal(mask, o_mask)
    assert np.array_equal(meta, o_meta)
    assert np.array_equal(counts, o_mask.sum(1))
    # legal_kth / legal_rank enumerate in ascending action index
    for i in range(0, st["board"].shape[0], 13):
        idx = np.nonzero(o_mask[i])[0]
        assert list(kth[i, :idx.size]) == list(idx)
        assert np.array_equal(rank[i], np.concatenate([[0], np.cumsum(o_mask[i])[:-1]]))


def _playout_states(n_games, seed):
    out = []
    for g in range(n_games):
        trace = oracle.random_playout(seed, g, 512, want_trace=True)["trace"]
        st = oracle.initial_states(1)
        for a in trace:
            out.append(st)
            st = oracle.apply_move_scalar(st, int(a))
        out.append(st)
    return concat_states(out)


def test_legal_scalar_semantics_and_status(shim):
    st = concat_states([_playout_states(40, 5), sparse_random_states(4000, 17)])
    n = st["board"].shape[0]
    mask, _meta, counts, _kth, _rank = _legal(shim, st, 1)
    board, mb, mw, sc = _flat(st)
    win = np.zeros((n,), np.int32)
    over = np.zeros((n,), np.uint8)
    hashes = np.zeros((n,), np.uint64)
    shim.hs_status(ctypes.c_int64(n), _p(board), _p(mb), _p(mw), _p(sc), _p(win), _p(over), _p(hashes))
    for i in range(n):
        legal = oracle.legal_actions(st, i)[0]
        assert list(np.nonzero(mask[i])[0]) == legal, i
        assert counts[i] == len(legal)
        assert win[i] == oracle.winner(st, i)
        assert bool(over[i]) == oracle.is_game_over(st, i)
        if i % 50 == 0:
            assert int(hashes[i]) == oracle.state_hash(st, i)


def test_apply_cuda_semantics(shim):
    # (1) the reference's synthetic sampler (many illegal -> silent no-op rows)
    st, codes, _parents = random_apply_batch(6000, 0xA11CEB0B)
    # (2) every legal child of reachable states
    ps = _playout_states(25, 8)
    m, meta = oracle.encode_actions_fast(ps)
    rows, cols = np.nonzero(m)
    st2 = {k: v[rows] for k, v in ps.items()}
    codes2 = meta[rows, cols]
    # (3) random garbage: random states x random codes
    st3 = sparse_random_states(6000, 4)
    rng = np.random.default_rng(3)
    codes3 = np.stack([rng.integers(0, 10, 6000), rng.integers(-2, 38, 6000), rng.integers(-1, 5, 6000),
                       rng.integers(-1, 36, 6000)], 1).astype(np.int32)
    for s, c in ((st, codes), (st2, codes2), (st3, codes3)):
        n = c.shape[0]
        expect, applied = oracle.batch_apply_moves(s, c, np.arange(n), return_applied=True)
        board, mb, mw, sc = _flat(s)
        got_applied = np.zeros((n,), np.uint8)
        c = np.ascontiguousarray(c)
        shim.hs_apply(ctypes.c_int64(n), _p(board), _p(mb), _p(mw), _p(sc), _p(c), _p(got_applied))
        assert np.array_equal(board.reshape(-1, 6, 6), expect["board"])
        assert np.array_equal(mb.reshape(-1, 6, 6).astype(bool), expect["marks_black"])
        assert np.array_equal(mw.reshape(-1, 6, 6).astype(bool), expect["marks_white"])
        for j, k in enumerate(SCALARS):
            assert np.array_equal(sc[:, j], expect[k]), k
        assert np.array_equal(got_applied.astype(bool), applied)


def test_apply_index_and_pack_roundtrip(shim):
    ps = _playout_states(25, 12)
    m, _ = oracle.encode_actions_fast(ps)
    rows, cols = np.nonzero(m)
    keep = [i for i in range(rows.size) if not oracle.is_game_over(ps, int(rows[i]))][::3]
    rows, cols = rows[keep], cols[keep]
    st = {k: v[rows] for k, v in ps.items()}
    board, mb, mw, sc = _flat(st)
    actions = np.ascontiguousarray(cols.astype(np.int32))
    shim.hs_apply_index(ctypes.c_int64(rows.size), _p(board), _p(mb), _p(mw), _p(sc), _p(actions))
    for i in range(rows.size):
        exp = oracle.apply_move_scalar(st, int(cols[i]), i)
        assert np.array_equal(board[i].reshape(6, 6), exp["board"][0])
        assert np.array_equal(mb[i].reshape(6, 6).astype(bool), exp["marks_black"][0])
        assert np.array_equal(mw[i].reshape(6, 6).astype(bool), exp["marks_white"][0])
        for j, k in enumerate(SCALARS):
            assert sc[i, j] == exp[k][0], (i, k)


@pytest.mark.parametrize("chunk", [1, 7, 1000])
def test_playout_matches_oracle(shim, chunk):
    shim.hs_playout.restype = ctypes.c_int
    for g in range(150):
        exp = oracle.random_playout(20260314, g, 512)
        res = ctypes.c_int(0)
        h = ctypes.c_uint64(0)
        board = np.zeros((36,), np.int8)
        mb = np.zeros((36,), np.uint8)
        mw = np.zeros((36,), np.uint8)
        sc = np.zeros((9,), np.int64)
        plies = shim.hs_playout(ctypes.c_uint64(20260314), ctypes.c_uint64(g), ctypes.c_int(512), ctypes.c_int(chunk),
                                ctypes.byref(res), ctypes.byref(h), _p(board), _p(mb), _p(mw), _p(sc))
        assert plies == exp["plies"] and res.value == exp["result"] and h.value == exp["hash"], g
        assert np.array_equal(board.reshape(6, 6), exp["final"]["board"][0])
