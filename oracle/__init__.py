"""CPU ORACLE -- test infrastructure, NOT product code.

numpy/ctypes front-end of ``oracle/liblz_oracle.so`` (plain-C restatement of the reference's rule engine,
legal-mask encoder, move application, root-PUCT loop and full-tree MCTS; see ``lz_oracle.h`` for the
reference file:line each function follows) plus small numpy restatements of the reference's ATen
composites (``root_pack_sparse_actions``, ``root_finalize_from_visits``, ``self_play_step_inplace``,
``finalize_trajectory_inplace``, ``project_policy_logits_fast``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
may import this package.  ``liuzhou_b200`` never does.

Parity status: pinned against the reference's own binaries (``oracle/_ref``) and committed golden vectors.
"""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_LIB_PATH = HERE / "liblz_oracle.so"

PLACEMENT_DIM, MOVEMENT_DIM, SELECTION_DIM, AUXILIARY_DIM = 36, 144, 36, 4
TOTAL_DIM = 220

STATE_FIELDS = (
    "board", "marks_black", "marks_white", "phase", "current_player",
    "pending_marks_required", "pending_marks_remaining",
    "pending_captures_required", "pending_captures_remaining",
    "forced_removals_done", "move_count", "moves_since_capture",
)


def build(force: bool = False) -> Path:
    srcs = [HERE / "lz_oracle.c", HERE / "lz_tree_oracle.c", HERE / "lz_oracle.h"]
    if (not force and _LIB_PATH.exists()
            and all(_LIB_PATH.stat().st_mtime >= s.stat().st_mtime for s in srcs)):
        return _LIB_PATH
    cmd = ["gcc", "-O2", "-fPIC", "-std=c11", "-ffp-contract=off", "-shared",
           str(srcs[0]), str(srcs[1]), "-lm", "-o", str(_LIB_PATH)]
    subprocess.run(cmd, check=True)
    return _LIB_PATH


class _State(ctypes.Structure):
    _fields_ = [
        ("board", ctypes.c_int8 * 36), ("marks_black", ctypes.c_uint8 * 36), ("marks_white", ctypes.c_uint8 * 36),
        ("phase", ctypes.c_int64), ("current_player", ctypes.c_int64),
        ("pending_marks_required", ctypes.c_int64), ("pending_marks_remaining", ctypes.c_int64),
        ("pending_captures_required", ctypes.c_int64), ("pending_captures_remaining", ctypes.c_int64),
        ("forced_removals_done", ctypes.c_int64), ("move_count", ctypes.c_int64),
        ("moves_since_capture", ctypes.c_int64),
    ]


class _Batch(ctypes.Structure):
    _fields_ = [(name, ctypes.c_void_p) for name in STATE_FIELDS]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(str(_LIB_PATH))
        L.or_mix64.restype = ctypes.c_uint64
        L.or_mix64.argtypes = [ctypes.c_uint64]
        L.or_playout_pick.restype = ctypes.c_uint32
        L.or_playout_pick.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32]
        L.or_state_hash.restype = ctypes.c_uint64
        L.or_random_playout.restype = ctypes.c_int
        L.or_random_playout.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.or_tree_create.restype = ctypes.c_void_p
        L.or_tree_create.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_double]
        for name in ("or_tree_free", "or_tree_prepare_roots", "or_tree_select_leaves", "or_tree_complete_pending",
                     "or_tree_root_priors", "or_tree_set_root_priors", "or_tree_root_outputs",
                     "or_tree_advance_roots", "or_tree_deactivate", "or_tree_root_state",
                     "or_tree_root_visit_count", "or_tree_pending_states"):
            getattr(L, name).argtypes = None
        L.or_tree_free.restype = None
        _lib = L
    return _lib


# ----------------------------------------------------------------------------------------------
# State batches in the reference tensor layout, as dicts of numpy arrays
# ----------------------------------------------------------------------------------------------
def empty_states(n: int) -> dict:
    st = {
        "board": np.zeros((n, 6, 6), np.int8),
        "marks_black": np.zeros((n, 6, 6), np.bool_),
        "marks_white": np.zeros((n, 6, 6), np.bool_),
    }
    for name in STATE_FIELDS[3:]:
        st[name] = np.zeros((n,), np.int64)
    return st


def initial_states(n: int) -> dict:
    st = empty_states(n)
    st["phase"][:] = 1
    st["current_player"][:] = 1
    return st


def _norm(st: dict) -> dict:
    out = {}
    out["board"] = np.ascontiguousarray(st["board"], dtype=np.int8)
    out["marks_black"] = np.ascontiguousarray(st["marks_black"]).astype(np.uint8, copy=False)
    out["marks_white"] = np.ascontiguousarray(st["marks_white"]).astype(np.uint8, copy=False)
    n = out["board"].shape[0]
    for name in STATE_FIELDS[3:]:
        if name in st and st[name] is not None:
            out[name] = np.ascontiguousarray(st[name], dtype=np.int64)
        else:
            out[name] = np.zeros((n,), np.int64)
    return out


def _view(st: dict) -> _Batch:
    b = _Batch()
    for name in STATE_FIELDS:
        setattr(b, name, st[name].ctypes.data)
    return b


def _to_public(st: dict) -> dict:
    out = dict(st)
    out["marks_black"] = st["marks_black"].astype(np.bool_).reshape(-1, 6, 6)
    out["marks_white"] = st["marks_white"].astype(np.bool_).reshape(-1, 6, 6)
    out["board"] = st["board"].reshape(-1, 6, 6)
    return out


def states_to_structs(st: dict):
    s = _norm(st)
    n = s["board"].shape[0]
    arr = (_State * n)()
    b = _view(s)
    for i in range(n):
        lib().or_load(ctypes.byref(b), ctypes.c_int64(i), ctypes.byref(arr[i]))
    return arr


def structs_to_states(arr, n: int) -> dict:
    st = _norm(empty_states(n))
    b = _view(st)
    for i in range(n):
        lib().or_store(ctypes.byref(arr[i]), ctypes.byref(b), ctypes.c_int64(i))
    return _to_public(st)


# ----------------------------------------------------------------------------------------------
# Tensor-op semantics
# ----------------------------------------------------------------------------------------------
def encode_actions_fast(st: dict, placement_dim=36, movement_dim=144, selection_dim=36, auxiliary_dim=4):
    s = _norm(st)
    n = s["board"].shape[0]
    total = placement_dim + movement_dim + selection_dim + auxiliary_dim
    mask = np.zeros((n, total), np.uint8)
    meta = np.zeros((n, total, 4), np.int32)
    b = _view(s)
    lib().or_encode_actions(ctypes.c_int64(n), ctypes.byref(b), ctypes.c_int64(placement_dim),
                            ctypes.c_int64(movement_dim), ctypes.c_int64(selection_dim),
                            ctypes.c_int64(auxiliary_dim), ctypes.c_void_p(mask.ctypes.data),
                            ctypes.c_void_p(meta.ctypes.data))
    return mask.astype(np.bool_), meta


def batch_apply_moves(st: dict, action_codes, parent_indices, return_applied: bool = False):
    s = _norm(st)
    n = s["board"].shape[0]
    codes = np.ascontiguousarray(action_codes, dtype=np.int32).reshape(-1, 4)
    parents = np.ascontiguousarray(parent_indices, dtype=np.int64).reshape(-1)
    m = codes.shape[0]
    out = _norm(empty_states(m))
    applied = np.zeros((m,), np.uint8)
    bi, bo = _view(s), _view(out)
    lib().or_batch_apply_moves(ctypes.c_int64(n), ctypes.byref(bi), ctypes.c_int64(m),
                               ctypes.c_void_p(codes.ctypes.data), ctypes.c_void_p(parents.ctypes.data),
                               ctypes.byref(bo), ctypes.c_void_p(applied.ctypes.data))
    res = _to_public(out)
    return (res, applied.astype(np.bool_)) if return_applied else res


def batch_apply_moves_inplace(st: dict, action_codes, slot_indices) -> dict:
    s = {k: np.array(v, copy=True) for k, v in _norm(st).items()}
    n = s["board"].shape[0]
    codes = np.ascontiguousarray(action_codes, dtype=np.int32).reshape(-1, 4)
    slots = np.ascontiguousarray(slot_indices, dtype=np.int64).reshape(-1)
    b = _view(s)
    lib().or_batch_apply_moves_inplace(ctypes.c_int64(n), ctypes.byref(b), ctypes.c_int64(codes.shape[0]),
                                       ctypes.c_void_p(codes.ctypes.data), ctypes.c_void_p(slots.ctypes.data))
    return _to_public(s)


def states_to_model_input(st: dict) -> np.ndarray:
    s = _norm(st)
    n = s["board"].shape[0]
    out = np.zeros((n, 11, 6, 6), np.float32)
    b = _view(s)
    lib().or_states_to_model_input(ctypes.c_int64(n), ctypes.byref(b), ctypes.c_void_p(out.ctypes.data))
    return out


def root_puct_allocate_visits(priors, leaf_values, valid_mask, num_simulations: int, exploration_weight: float):
    p = np.ascontiguousarray(priors, dtype=np.float32)
    lv = np.ascontiguousarray(leaf_values, dtype=np.float32)
    vm = np.ascontiguousarray(valid_mask).astype(np.uint8)
    r, m = p.shape
    visits = np.zeros((r, m), np.float32)
    value_sum = np.zeros((r, m), np.float32)
    root_values = np.zeros((r,), np.float32)
    if r and m:
        lib().or_root_puct_allocate_visits(
            ctypes.c_int64(r), ctypes.c_int64(m), ctypes.c_void_p(p.ctypes.data), ctypes.c_void_p(lv.ctypes.data),
            ctypes.c_void_p(vm.ctypes.data), ctypes.c_int64(int(num_simulations)),
            ctypes.c_float(float(exploration_weight)), ctypes.c_void_p(visits.ctypes.data),
            ctypes.c_void_p(value_sum.ctypes.data), ctypes.c_void_p(root_values.ctypes.data))
    return visits, value_sum, root_values


# ----------------------------------------------------------------------------------------------
# Scalar-engine semantics
# ----------------------------------------------------------------------------------------------
def legal_actions(st: dict, i: int = 0):
    arr = states_to_structs({k: v[i:i + 1] for k, v in st.items()})
    idx = (ctypes.c_int * 220)()
    codes = (ctypes.c_int32 * 880)()
    n = lib().or_legal_actions(ctypes.byref(arr[0]), idx, codes)
    return list(idx[:n]), np.array(codes[: n * 4], dtype=np.int32).reshape(n, 4)


def is_game_over(st: dict, i: int = 0) -> bool:
    arr = states_to_structs({k: v[i:i + 1] for k, v in st.items()})
    return bool(lib().or_is_game_over(ctypes.byref(arr[0])))


def winner(st: dict, i: int = 0) -> int:
    arr = states_to_structs({k: v[i:i + 1] for k, v in st.items()})
    return int(lib().or_winner(ctypes.byref(arr[0])))


def apply_move_scalar(st: dict, action_index: int, i: int = 0):
    arr = states_to_structs({k: v[i:i + 1] for k, v in st.items()})
    out = (_State * 1)()
    ok = lib().or_apply_move_scalar(ctypes.byref(arr[0]), ctypes.c_int(int(action_index)), ctypes.byref(out[0]))
    if not ok:
        raise RuntimeError("illegal action for the scalar engine")
    return structs_to_states(out, 1)


def random_playout(seed: int, game: int, max_plies: int = 512, want_trace: bool = False):
    final = (_State * 1)()
    result = ctypes.c_int(0)
    h = ctypes.c_uint64(0)
    trace = np.full((max_plies,), -1, np.int16) if want_trace else None
    plies = lib().or_random_playout(
        ctypes.c_uint64(seed), ctypes.c_uint64(game), ctypes.c_int(max_plies),
        ctypes.cast(final, ctypes.c_void_p), ctypes.cast(ctypes.byref(result), ctypes.c_void_p),
        ctypes.c_void_p(trace.ctypes.data) if want_trace else None,
        ctypes.cast(ctypes.byref(h), ctypes.c_void_p))
    out = {"plies": int(plies), "result": int(result.value), "hash": int(h.value),
           "final": structs_to_states(final, 1)}
    if want_trace:
        out["trace"] = trace[:plies].copy()
    return out


def random_playouts_range(seed: int, g0: int, g1: int, max_plies: int = 512):
    """Bench helper: plays games [g0, g1) in C (releases the GIL) -> dict(plies, black, white, draws, hash_xor)."""
    out = (ctypes.c_uint64 * 5)()
    lib().or_random_playouts_range(ctypes.c_uint64(seed), ctypes.c_uint64(g0), ctypes.c_uint64(g1),
                                   ctypes.c_int(max_plies), out)
    return {"plies": int(out[0]), "black": int(out[1]), "white": int(out[2]), "draws": int(out[3]),
            "hash_xor": int(out[4])}


def random_playouts_each(seed: int, g0: int, g1: int, max_plies: int = 512, threads: int = 0):
    """Per-game (plies int32, result int8, hash uint64) of games [g0, g1), played in C on `threads` host threads (the C
    call releases the GIL)."""
    import os
    from concurrent.futures import ThreadPoolExecutor

    n = int(g1) - int(g0)
    plies = np.zeros((n,), np.int32)
    result = np.zeros((n,), np.int8)
    hashes = np.zeros((n,), np.uint64)
    threads = int(threads) or (os.cpu_count() or 1)
    chunk = max(1, -(-n // (threads * 4)))
    L = lib()

    def work(lo):
        hi = min(n, lo + chunk)
        L.or_random_playouts_each(ctypes.c_uint64(seed), ctypes.c_uint64(g0 + lo), ctypes.c_uint64(g0 + hi),
                                  ctypes.c_int(max_plies), ctypes.c_void_p(plies[lo:].ctypes.data),
                                  ctypes.c_void_p(result[lo:].ctypes.data), ctypes.c_void_p(hashes[lo:].ctypes.data))

    with ThreadPoolExecutor(max_workers=threads) as pool:
        list(pool.map(work, range(0, n, chunk)))
    return plies, result, hashes


def state_hash(st: dict, i: int = 0) -> int:
    arr = states_to_structs({k: v[i:i + 1] for k, v in st.items()})
    return int(lib().or_state_hash(ctypes.byref(arr[0])))


# ----------------------------------------------------------------------------------------------
# Full-tree MCTS (PortableTreeBatch protocol)
# ----------------------------------------------------------------------------------------------
class TreeBatch:
    """Restatement of ``_liuzhou_portable_cpp.PortableTreeBatch`` (portable_mcts.cpp:437-977)."""

    def __init__(self, st: dict, exploration_weight: float = 1.0):
        self._structs = states_to_structs(st)
        self.num_trees = len(self._structs)
        self._h = ctypes.c_void_p(lib().or_tree_create(
            ctypes.c_int(self.num_trees), ctypes.cast(self._structs, ctypes.c_void_p),
            ctypes.c_double(float(exploration_weight))))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().or_tree_free(self._h)
            self._h = None

    def _pending(self, fn):
        t = self.num_trees
        tree_idx = np.zeros((t,), np.int32)
        inputs = np.zeros((t, 11, 6, 6), np.float32)
        masks = np.zeros((t, 220), np.uint8)
        n = fn(self._h, ctypes.c_void_p(tree_idx.ctypes.data), ctypes.c_void_p(inputs.ctypes.data),
               ctypes.c_void_p(masks.ctypes.data))
        if n < 0:
            raise RuntimeError("complete the current model evaluation before starting another operation")
        return {"tree_indices": tree_idx[:n].copy(), "model_inputs": inputs[:n].copy(),
                "legal_masks": masks[:n].copy()}

    def prepare_roots(self):
        return self._pending(lib().or_tree_prepare_roots)

    def select_leaves(self):
        return self._pending(lib().or_tree_select_leaves)

    def complete_pending(self, priors, values):
        p = np.ascontiguousarray(priors, dtype=np.float32)
        v = np.ascontiguousarray(values, dtype=np.float32)
        if lib().or_tree_complete_pending(self._h, ctypes.c_void_p(p.ctypes.data), ctypes.c_void_p(v.ctypes.data)) < 0:
            raise RuntimeError("there is no pending evaluation")

    def root_priors(self):
        t = self.num_trees
        priors = np.zeros((t, 220), np.float32)
        masks = np.zeros((t, 220), np.uint8)
        active = np.zeros((t,), np.uint8)
        lib().or_tree_root_priors(self._h, ctypes.c_void_p(priors.ctypes.data), ctypes.c_void_p(masks.ctypes.data),
                                  ctypes.c_void_p(active.ctypes.data))
        return {"priors": priors, "legal_masks": masks, "active": active}

    def set_root_priors(self, priors):
        p = np.ascontiguousarray(priors, dtype=np.float32)
        if lib().or_tree_set_root_priors(self._h, ctypes.c_void_p(p.ctypes.data)) < 0:
            raise RuntimeError("root priors have a non-positive sum")

    def root_outputs(self):
        t = self.num_trees
        masks = np.zeros((t, 220), np.uint8)
        visits = np.zeros((t, 220), np.int32)
        qv = np.zeros((t, 220), np.float32)
        rv = np.zeros((t,), np.float32)
        players = np.zeros((t,), np.int32)
        term = np.zeros((t,), np.uint8)
        inputs = np.zeros((t, 11, 6, 6), np.float32)
        lib().or_tree_root_outputs(self._h, *(ctypes.c_void_p(a.ctypes.data)
                                              for a in (masks, visits, qv, rv, players, term, inputs)))
        return {"legal_masks": masks, "visit_counts": visits, "root_action_values": qv, "root_values": rv,
                "current_players": players, "terminal": term, "model_inputs": inputs}

    def advance_roots(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.int32)
        rc = lib().or_tree_advance_roots(self._h, ctypes.c_void_p(a.ctypes.data))
        if rc < 0:
            raise RuntimeError("selected action is not a child of the current root")

    def deactivate(self, indices):
        for i in indices:
            lib().or_tree_deactivate(self._h, ctypes.c_int(int(i)))

    def pending_states(self) -> dict:
        arr = (_State * self.num_trees)()
        n = lib().or_tree_pending_states(self._h, ctypes.cast(arr, ctypes.c_void_p))
        return structs_to_states(arr, n)

    def root_state(self, tree: int) -> dict:
        out = (_State * 1)()
        lib().or_tree_root_state(self._h, ctypes.c_int(int(tree)), ctypes.byref(out[0]))
        return structs_to_states(out, 1)


from .composites import (  # noqa: E402,F401
    finalize_trajectory_inplace,
    project_policy_logits_fast,
    root_finalize_from_visits,
    root_pack_sparse_actions,
    root_sparse_writeback,
    self_play_step_inplace,
    soft_value_from_board,
    terminal_mask_from_next_state,
)
