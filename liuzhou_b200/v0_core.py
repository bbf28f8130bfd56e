"""Drop-in replacement for the tensor-op surface of the reference's ``v0_core`` extension module
(/root/reference/v0/src/bindings/module.cpp:1286-1420): same function names, positional argument order,
dtypes, shapes, ownership (fresh tensors returned, ``*_inplace`` ops mutate) and error type (RuntimeError).

Every op launches hand-written sm_100a kernels through the C ABI of libliuzhou_b200.so on the tensors'
device and the current CUDA stream.  CUDA tensors only: there is deliberately no CPU fallback.

    import sys, liuzhou_b200.v0_core as v0_core; sys.modules["v0_core"] = v0_core   # before importing v1.python.*
"""
from __future__ import annotations

import ctypes
from enum import IntEnum

import torch

from . import _lib
from ._lib import check, i64, lib, ptr, require_cuda, states_view, stream_ptr

__all__ = [
    "Phase", "Player", "ActionType", "encode_actions_fast", "batch_apply_moves", "batch_apply_moves_inplace",
    "states_to_model_input", "project_policy_logits_fast", "root_puct_allocate_visits",
    "root_pack_sparse_actions", "root_sparse_writeback", "root_finalize_from_visits", "self_play_step_inplace",
    "finalize_trajectory_inplace", "version",
]


class Phase(IntEnum):  # v0/include/v0/game_state.hpp:24-32
    PLACEMENT = 1
    MARK_SELECTION = 2
    REMOVAL = 3
    MOVEMENT = 4
    CAPTURE_SELECTION = 5
    FORCED_REMOVAL = 6
    COUNTER_REMOVAL = 7


class Player(IntEnum):  # game_state.hpp:34-37
    BLACK = 1
    WHITE = -1


class ActionType(IntEnum):  # v0/include/v0/move_generator.hpp:13-22
    PLACE = 1
    MOVE = 2
    MARK = 3
    CAPTURE = 4
    FORCED_REMOVAL = 5
    COUNTER_REMOVAL = 6
    NO_MOVES_REMOVAL = 7
    PROCESS_REMOVAL = 8


def version() -> str:
    return lib().lzb_version().decode()


def _prep_states(tensors, n_required):
    """`.contiguous()` + dtype checks like the reference launchers (fast_legal_mask_cuda.cu:425-437)."""
    board = tensors[0]
    require_cuda(board, "board")
    out = []
    for i, t in enumerate(tensors):
        if t.device != board.device:
            raise RuntimeError("all state tensors must be on the same device as board")
        if i == 0:
            if t.dtype != torch.int8:
                raise RuntimeError("board must be int8")
        elif i in (1, 2):
            if t.dtype != torch.bool:
                t = t.to(torch.bool)
        elif t.dtype != torch.int64:
            raise RuntimeError("state scalars must be int64")
        out.append(t.contiguous())
    if board.dim() != 3 or board.size(1) != 6 or board.size(2) != 6:
        raise RuntimeError("board must be (B, 6, 6)")
    assert len(out) == n_required
    return out


def encode_actions_fast(board, marks_black, marks_white, phase, current_player, pending_marks_required,
                        pending_marks_remaining, pending_captures_required, pending_captures_remaining,
                        forced_removals_done, placement_dim, movement_dim, selection_dim, auxiliary_dim):
    """-> (mask bool[B,T], metadata int32[B,T,4]); reference: fast_legal_mask_cuda.cu:406-485."""
    st = _prep_states([board, marks_black, marks_white, phase, current_player, pending_marks_required,
                       pending_marks_remaining, pending_captures_required, pending_captures_remaining,
                       forced_removals_done], 10)
    b = st[0].size(0)
    total = int(placement_dim) + int(movement_dim) + int(selection_dim) + int(auxiliary_dim)
    dev = st[0].device
    with torch.cuda.device(dev):
        mask = torch.empty((b, total), dtype=torch.bool, device=dev)
        meta = torch.empty((b, total, 4), dtype=torch.int32, device=dev)
        view = states_view(st)
        check(lib().lzb_encode_actions_fast(ctypes.byref(view), i64(b), i64(placement_dim), i64(movement_dim),
                                            i64(selection_dim), i64(auxiliary_dim), ptr(mask), ptr(meta),
                                            stream_ptr(dev)))
    return mask, meta


def _alloc_children(n, ref):
    dev = ref[0].device
    out = [torch.empty((n, 6, 6), dtype=torch.int8, device=dev),
           torch.empty((n, 6, 6), dtype=torch.bool, device=dev),
           torch.empty((n, 6, 6), dtype=torch.bool, device=dev)]
    out += [torch.empty((n,), dtype=torch.int64, device=dev) for _ in range(9)]
    return out


def batch_apply_moves(board, marks_black, marks_white, phase, current_player, pending_marks_required,
                      pending_marks_remaining, pending_captures_required, pending_captures_remaining,
                      forced_removals_done, move_count, moves_since_capture, action_codes, parent_indices):
    """-> 12-tuple of child tensors; reference: fast_apply_moves_cuda.cu:920-1046."""
    st = _prep_states([board, marks_black, marks_white, phase, current_player, pending_marks_required,
                       pending_marks_remaining, pending_captures_required, pending_captures_remaining,
                       forced_removals_done, move_count, moves_since_capture], 12)
    dev = st[0].device
    if action_codes.dim() != 2 or action_codes.size(1) != 4:
        raise RuntimeError("action_codes must be (N, 4).")
    codes = action_codes.to(device=dev, dtype=torch.int32).contiguous()
    parents = parent_indices.to(device=dev, dtype=torch.int64).contiguous()
    n = codes.size(0)
    if parents.numel() != n:
        raise RuntimeError("parent_indices must align with action_codes.")
    with torch.cuda.device(dev):
        out = _alloc_children(n, st)
        vin, vout = states_view(st), states_view(out)
        check(lib().lzb_batch_apply_moves(ctypes.byref(vin), i64(st[0].size(0)), ptr(codes), ptr(parents), i64(n),
                                          ctypes.byref(vout), stream_ptr(dev)))
    return tuple(out)


def batch_apply_moves_inplace(board, marks_black, marks_white, phase, current_player, pending_marks_required,
                              pending_marks_remaining, pending_captures_required, pending_captures_remaining,
                              forced_removals_done, move_count, moves_since_capture, action_codes, slot_indices):
    """Mutates the 12 state tensors; reference: fast_apply_moves_cuda.cu:1048-1134."""
    tensors = [board, marks_black, marks_white, phase, current_player, pending_marks_required,
               pending_marks_remaining, pending_captures_required, pending_captures_remaining,
               forced_removals_done, move_count, moves_since_capture]
    for t in tensors:
        require_cuda(t)
        if not t.is_contiguous():
            raise RuntimeError("in-place apply needs contiguous state tensors")
    dev = board.device
    codes = action_codes.to(device=dev, dtype=torch.int32).contiguous()
    slots = slot_indices.to(device=dev, dtype=torch.int64).contiguous()
    if codes.dim() != 2 or codes.size(1) != 4:
        raise RuntimeError("action_codes must be (N, 4).")
    if slots.numel() != codes.size(0):
        raise RuntimeError("slot_indices must align with action_codes.")
    with torch.cuda.device(dev):
        view = states_view(tensors)
        check(lib().lzb_batch_apply_moves_inplace(ctypes.byref(view), i64(board.size(0)), ptr(codes), ptr(slots),
                                                  i64(codes.size(0)), stream_ptr(dev)))


def states_to_model_input(board, marks_black, marks_white, phase, current_player):
    """-> float32[B,11,6,6]; reference: v0/src/net/encoding.cpp:26-79."""
    require_cuda(board, "board")
    if board.dim() != 3:
        raise RuntimeError("board must be (B, H, W)")
    dev = board.device
    b = board.size(0)
    tensors = [board.contiguous(), marks_black.to(torch.bool).contiguous(), marks_white.to(torch.bool).contiguous(),
               phase.to(torch.int64).contiguous(), current_player.to(torch.int64).contiguous()]
    if tensors[1].shape != board.shape or tensors[2].shape != board.shape:
        raise RuntimeError("marks shape mismatch")
    if tensors[3].numel() != b or tensors[4].numel() != b:
        raise RuntimeError("phase / current_player length mismatch")
    with torch.cuda.device(dev):
        out = torch.empty((b, 11, 6, 6), dtype=torch.float32, device=dev)
        view = states_view(tensors + [None] * 5)
        check(lib().lzb_states_to_model_input(ctypes.byref(view), i64(b), ptr(out), stream_ptr(dev)))
    return out


def project_policy_logits_fast(log_p1, log_p2, log_pmc, legal_mask, placement_dim, movement_dim, selection_dim,
                               auxiliary_dim):
    """-> (probs[B,220], masked_logits[B,220]) in the heads' dtype; reference: project_policy_logits_fast.cpp:16-164."""
    require_cuda(log_p1, "log_p1")
    if log_p1.dim() != 2 or log_p1.shape != log_p2.shape or log_p1.shape != log_pmc.shape:
        raise RuntimeError("All policy heads must share the same shape.")
    if legal_mask.dtype != torch.bool:
        raise RuntimeError("legal_mask must be of dtype bool.")
    if (int(placement_dim), int(movement_dim), int(selection_dim), int(auxiliary_dim)) != (36, 144, 36, 4):
        raise RuntimeError("liuzhou_b200 projects the 36/144/36/4 action layout only")
    b = log_p1.size(0)
    if log_p1.size(1) != 36 or tuple(legal_mask.shape) != (b, 220):
        raise RuntimeError("Policy head dimension mismatch")
    dev = log_p1.device
    dtype = log_p1.dtype
    h = [t.to(torch.float32).contiguous() for t in (log_p1, log_p2, log_pmc)]
    lm = legal_mask.contiguous()
    with torch.cuda.device(dev):
        probs = torch.empty((b, 220), dtype=torch.float32, device=dev)
        logits = torch.empty((b, 220), dtype=torch.float32, device=dev)
        check(lib().lzb_project_policy_logits_fast(ptr(h[0]), ptr(h[1]), ptr(h[2]), ptr(lm), i64(b), ptr(probs),
                                                   ptr(logits), stream_ptr(dev)))
    if dtype != torch.float32:
        probs, logits = probs.to(dtype), logits.to(dtype)
    return probs, logits


def root_puct_allocate_visits(priors, leaf_values, valid_mask, num_simulations, exploration_weight):
    """-> (visits f32[R,M], value_sum f32[R,M], root_values f32[R]); reference: root_puct_fused.cu:130-183."""
    if priors.dim() != 2 or leaf_values.dim() != 2 or valid_mask.dim() != 2:
        raise RuntimeError("priors / leaf_values / valid_mask must be 2D [R, A]")
    if priors.shape != leaf_values.shape or priors.shape != valid_mask.shape:
        raise RuntimeError("priors/leaf_values/valid_mask shape mismatch")
    if int(num_simulations) <= 0:
        raise RuntimeError("num_simulations must be positive")
    require_cuda(priors, "priors")
    dev = priors.device
    p = priors.to(torch.float32).contiguous()
    lv = leaf_values.to(device=dev, dtype=torch.float32).contiguous()
    vm = valid_mask.to(device=dev, dtype=torch.bool).contiguous()
    r, m = p.shape
    with torch.cuda.device(dev):
        visits = torch.zeros((r, m), dtype=torch.float32, device=dev)
        value_sum = torch.zeros((r, m), dtype=torch.float32, device=dev)
        root_values = torch.zeros((r,), dtype=torch.float32, device=dev)
        check(lib().lzb_root_puct_allocate_visits(ptr(p), ptr(lv), ptr(vm), i64(r), i64(m), i64(num_simulations),
                                                  ctypes.c_float(float(exploration_weight)), ptr(visits),
                                                  ptr(value_sum), ptr(root_values), stream_ptr(dev)))
    return visits, value_sum, root_values


def root_pack_sparse_actions(legal_mask, probs, metadata):
    raise NotImplementedError


def root_sparse_writeback(*args):
    raise NotImplementedError


def root_finalize_from_visits(*args):
    raise NotImplementedError


def self_play_step_inplace(*args):
    raise NotImplementedError


def finalize_trajectory_inplace(*args):
    raise NotImplementedError
