"""Drop-in replacement for the tensor-op surface of the reference's ``v0_core`` extension module
(/root/reference/v0/src/bindings/module.cpp:1286-1420): same function names, positional argument order,
dtypes, shapes, ownership (fresh tensors returned, ``*_inplace`` ops mutate) and error type (RuntimeError).

Every op launches hand-written sm_100a kernels through the C ABI of libliuzhou_b200.so on the tensors'
device and the current CUDA stream.  CUDA tensors only: there is deliberately no CPU fallback.

    import sys, liuzhou_b200.v0_core as v0_core; sys.modules["v0_core"] = v0_core   # before importing v1.python.*
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import check, i64, lib, ptr, require_cuda, states_view, stream_ptr

__all__ = [
    "Phase", "Player", "ActionType", "encode_actions_fast", "batch_apply_moves", "batch_apply_moves_inplace",
    "states_to_model_input", "project_policy_logits_fast", "root_puct_allocate_visits",
    "root_pack_sparse_actions", "root_sparse_writeback", "root_finalize_from_visits", "self_play_step_inplace",
    "finalize_trajectory_inplace", "version",
]


# enums, GameState / MoveRecord / ActionCode and the scalar rule functions (module.cpp:877-1156): liuzhou_b200/scalar_api.py
from .scalar_api import *  # noqa: E402,F401,F403
from .scalar_api import ActionType, Phase, Player  # noqa: E402,F401
from . import scalar_api as _scalar_api  # noqa: E402

__all__ += [n for n in _scalar_api.__all__ if n not in __all__]

# pybind11's export_values(): enum members are also module attributes (v0_core.PLACEMENT, v0_core.BLACK, v0_core.PLACE, ...)
for _e in (Phase, Player, ActionType):
    for _m in _e:
        globals().setdefault(_m.name, _m)


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
    """-> the reference's 10-tuple (terminal_mask, valid_root_indices, counts, valid_mask, legal_index_mat,
    priors_mat, action_code_mat, flat_indices, action_codes_all, parent_indices_all); module.cpp:258-363.
    One host sync for the data-dependent shapes (R, M, N), like the reference's `.item()` at :310."""
    if legal_mask.dim() != 2 or probs.dim() != 2:
        raise RuntimeError("legal_mask / probs must be 2D [B, A]")
    if metadata.dim() != 3 or metadata.size(2) != 4:
        raise RuntimeError("metadata must be 3D [B, A, 4]")
    if legal_mask.shape != probs.shape or tuple(metadata.shape[:2]) != tuple(legal_mask.shape):
        raise RuntimeError("legal/probs/metadata shape mismatch")
    require_cuda(legal_mask, "legal_mask")
    dev = legal_mask.device
    if probs.device != dev or metadata.device != dev:
        raise RuntimeError("legal_mask/probs/metadata must be on the same device")
    lm = legal_mask.to(torch.bool).contiguous()
    pr = probs.to(torch.float32).contiguous()
    md = metadata.to(torch.int32).contiguous()
    b, a = lm.shape
    i64t = dict(dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        row_counts = torch.empty((b,), **i64t)
        terminal = torch.empty((b,), dtype=torch.bool, device=dev)
        root_rank = torch.empty((b,), **i64t)
        flat_offset = torch.empty((b,), **i64t)
        summary = torch.empty((3,), **i64t)
        check(lib().lzb_root_pack_count(ptr(lm), i64(b), i64(a), ptr(row_counts), ptr(terminal), ptr(root_rank),
                                        ptr(flat_offset), ptr(summary), stream_ptr(dev)))
        r, m, n = (int(x) for x in summary.tolist())
        if r == 0:
            return (terminal, torch.empty((0,), **i64t), torch.empty((0,), **i64t),
                    torch.empty((0, 0), dtype=torch.bool, device=dev), torch.empty((0, 0), **i64t),
                    torch.empty((0, 0), dtype=torch.float32, device=dev),
                    torch.empty((0, 0, 4), dtype=torch.int32, device=dev), torch.empty((0,), **i64t),
                    torch.empty((0, 4), dtype=torch.int32, device=dev), torch.empty((0,), **i64t))
        roots = torch.empty((r,), **i64t)
        counts = torch.empty((r,), **i64t)
        valid_mask = torch.empty((r, m), dtype=torch.bool, device=dev)
        legal_index_mat = torch.empty((r, m), **i64t)
        priors_mat = torch.empty((r, m), dtype=torch.float32, device=dev)
        action_code_mat = torch.empty((r, m, 4), dtype=torch.int32, device=dev)
        flat_indices = torch.empty((n,), **i64t)
        action_codes_all = torch.empty((n, 4), dtype=torch.int32, device=dev)
        parent_indices_all = torch.empty((n,), **i64t)
        check(lib().lzb_root_pack_fill(ptr(lm), ptr(pr), ptr(md), i64(b), i64(a), ptr(row_counts), ptr(root_rank),
                                       ptr(flat_offset), i64(r), i64(m), ptr(roots), ptr(counts), ptr(valid_mask),
                                       ptr(legal_index_mat), ptr(priors_mat), ptr(action_code_mat), ptr(flat_indices),
                                       ptr(action_codes_all), ptr(parent_indices_all), stream_ptr(dev)))
    return (terminal, roots, counts, valid_mask, legal_index_mat, priors_mat, action_code_mat, flat_indices,
            action_codes_all, parent_indices_all)


def root_sparse_writeback(legal_index_mat, action_code_mat, valid_mask, legal_policy, local_picks,
                          valid_root_indices, batch_size, total_action_dim):
    """-> (policy_dense f32[B,A], chosen_action_indices i64[B], chosen_action_codes i32[B,4], chosen_valid_mask bool[B]);
    module.cpp:365-439: an externally supplied legal policy + per-root picks scattered back to dense rows."""
    if int(batch_size) < 0:
        raise RuntimeError("batch_size must be non-negative")
    if int(total_action_dim) <= 0:
        raise RuntimeError("total_action_dim must be positive")
    if legal_index_mat.dim() != 2 or valid_mask.dim() != 2 or legal_policy.dim() != 2:
        raise RuntimeError("legal_index_mat / valid_mask / legal_policy must be [R, M]")
    if action_code_mat.dim() != 3 or action_code_mat.size(2) != 4:
        raise RuntimeError("action_code_mat must be [R, M, 4]")
    r, m = legal_index_mat.shape
    if tuple(valid_mask.shape) != (r, m) or tuple(legal_policy.shape) != (r, m) or tuple(action_code_mat.shape[:2]) != (r, m):
        raise RuntimeError("legal_index_mat / valid_mask / legal_policy / action_code_mat shape mismatch")
    if local_picks.dim() != 1 or local_picks.numel() != r or valid_root_indices.dim() != 1 or valid_root_indices.numel() != r:
        raise RuntimeError("local_picks / valid_root_indices size mismatch")
    require_cuda(legal_index_mat, "legal_index_mat")
    dev = legal_index_mat.device
    for t in (action_code_mat, valid_mask, legal_policy, local_picks, valid_root_indices):
        if t.device != dev:
            raise RuntimeError("all tensors must be on the same device")
    li = legal_index_mat.to(torch.int64).contiguous()
    ac = action_code_mat.to(torch.int32).contiguous()
    vm = valid_mask.to(torch.bool).contiguous()
    lp = legal_policy.to(torch.float32).contiguous()
    pk = local_picks.to(torch.int64).contiguous()
    ro = valid_root_indices.to(torch.int64).contiguous()
    bsz, adim = int(batch_size), int(total_action_dim)
    with torch.cuda.device(dev):
        policy_dense = torch.empty((bsz, adim), dtype=torch.float32, device=dev)
        chosen_idx = torch.empty((bsz,), dtype=torch.int64, device=dev)
        chosen_codes = torch.empty((bsz, 4), dtype=torch.int32, device=dev)
        chosen_valid = torch.empty((bsz,), dtype=torch.bool, device=dev)
        check(lib().lzb_root_sparse_writeback(ptr(li), ptr(ac), ptr(vm), ptr(lp), ptr(pk), ptr(ro), i64(r), i64(m),
                                              i64(bsz), i64(adim), ptr(policy_dense), ptr(chosen_idx), ptr(chosen_codes),
                                              ptr(chosen_valid), stream_ptr(dev)))
    return policy_dense, chosen_idx, chosen_codes, chosen_valid


def root_finalize_from_visits(legal_index_mat, action_code_mat, valid_mask, visits, value_sum, valid_root_indices,
                              batch_size, total_action_dim, root_temperatures, sample_moves):
    """-> (policy_dense f32[B,A], chosen_action_indices i64[B], chosen_action_codes i32[B,4],
    chosen_valid_mask bool[B], root_value f32[R]); module.cpp:441-535.  sample_moves=True (never used by v1,
    mcts_gpu.py:1408) draws with torch.multinomial from the kernel's policy, then rewrites the picks."""
    if int(batch_size) < 0:
        raise RuntimeError("batch_size must be non-negative")
    if int(total_action_dim) <= 0:
        raise RuntimeError("total_action_dim must be positive")
    if legal_index_mat.dim() != 2 or valid_mask.dim() != 2 or visits.dim() != 2 or value_sum.dim() != 2:
        raise RuntimeError("legal_index_mat / valid_mask / visits / value_sum must be [R, M]")
    if action_code_mat.dim() != 3 or action_code_mat.size(2) != 4:
        raise RuntimeError("action_code_mat must be [R, M, 4]")
    r, m = legal_index_mat.shape
    for t in (valid_mask, visits, value_sum):
        if tuple(t.shape) != (r, m):
            raise RuntimeError("legal_index_mat / valid_mask / visits / value_sum shape mismatch")
    if valid_root_indices.numel() != r or root_temperatures.numel() != r:
        raise RuntimeError("valid_root_indices / root_temperatures size mismatch")
    require_cuda(legal_index_mat, "legal_index_mat")
    dev = legal_index_mat.device
    li = legal_index_mat.to(torch.int64).contiguous()
    ac = action_code_mat.to(torch.int32).contiguous()
    vm = valid_mask.to(torch.bool).contiguous()
    vi = visits.to(torch.float32).contiguous()
    vs = value_sum.to(torch.float32).contiguous()
    ro = valid_root_indices.to(torch.int64).contiguous()
    te = root_temperatures.to(torch.float32).contiguous()
    bsz, adim = int(batch_size), int(total_action_dim)
    with torch.cuda.device(dev):
        policy_dense = torch.empty((bsz, adim), dtype=torch.float32, device=dev)
        chosen_idx = torch.empty((bsz,), dtype=torch.int64, device=dev)
        chosen_codes = torch.empty((bsz, 4), dtype=torch.int32, device=dev)
        chosen_valid = torch.empty((bsz,), dtype=torch.bool, device=dev)
        root_value = torch.empty((r,), dtype=torch.float32, device=dev)
        check(lib().lzb_root_finalize_from_visits(ptr(li), ptr(ac), ptr(vm), ptr(vi), ptr(vs), ptr(ro), i64(r), i64(m),
                                                  i64(bsz), i64(adim), ptr(te), ptr(policy_dense), ptr(chosen_idx),
                                                  ptr(chosen_codes), ptr(chosen_valid), ptr(root_value),
                                                  stream_ptr(dev)))
    if sample_moves and m > 1 and r > 0:
        legal_policy = policy_dense.index_select(0, ro).gather(1, li) * vm.to(torch.float32)
        picks = torch.multinomial(legal_policy, 1, False).view(-1)
        chosen_idx.index_copy_(0, ro, li.gather(1, picks.view(-1, 1)).view(-1))
        chosen_codes.index_copy_(0, ro, ac.gather(1, picks.view(-1, 1, 1).expand(-1, 1, 4)).view(-1, 4))
    return policy_dense, chosen_idx, chosen_codes, chosen_valid, root_value


def self_play_step_inplace(board, marks_black, marks_white, phase, current_player, pending_marks_required,
                           pending_marks_remaining, pending_captures_required, pending_captures_remaining,
                           forced_removals_done, move_count, moves_since_capture, plies, done, active_idx,
                           chosen_action_codes, terminal_mask, chosen_valid_mask, max_game_plies, soft_value_k):
    """Mutates the 12 state tensors, plies, done; -> (finalize_slots i64[F], result_from_black f32[F],
    soft_value_from_black f32[F]); module.cpp:632-871.  One host sync to learn F (the reference does 3-4)."""
    if int(max_game_plies) <= 0:
        raise RuntimeError("max_game_plies must be positive")
    tensors = [board, marks_black, marks_white, phase, current_player, pending_marks_required,
               pending_marks_remaining, pending_captures_required, pending_captures_remaining,
               forced_removals_done, move_count, moves_since_capture]
    require_cuda(board, "board")
    dev = board.device
    bn = board.size(0)
    for t in tensors + [plies, done]:
        if t.device != dev:
            raise RuntimeError("all tensors must be on the same device as board")
        if t.size(0) != bn:
            raise RuntimeError("batch mismatch")
        if not t.is_contiguous():
            raise RuntimeError("self_play_step_inplace needs contiguous state tensors")
    if plies.dtype != torch.int64 or done.dtype != torch.bool:
        raise RuntimeError("plies must be int64 and done bool")
    active = active_idx.to(device=dev, dtype=torch.int64).reshape(-1).contiguous()
    codes = chosen_action_codes.to(device=dev, dtype=torch.int32).contiguous()
    term = terminal_mask.to(device=dev, dtype=torch.bool).reshape(-1).contiguous()
    cvalid = chosen_valid_mask.to(device=dev, dtype=torch.bool).reshape(-1).contiguous()
    k = active.numel()
    if codes.dim() != 2 or codes.size(0) != k or codes.size(1) != 4:
        raise RuntimeError("chosen_action_codes must be [A, 4]")
    if term.numel() != k or cvalid.numel() != k:
        raise RuntimeError("terminal_mask / chosen_valid_mask batch mismatch")
    with torch.cuda.device(dev):
        if k == 0:
            e = torch.empty((0,), dtype=torch.float32, device=dev)
            return torch.empty((0,), dtype=torch.int64, device=dev), e, e.clone()
        slots = torch.empty((k,), dtype=torch.int64, device=dev)
        result = torch.empty((k,), dtype=torch.float32, device=dev)
        soft = torch.empty((k,), dtype=torch.float32, device=dev)
        nfin = torch.empty((1,), dtype=torch.int64, device=dev)
        scratch = torch.empty((k, 4), dtype=torch.int32, device=dev)
        view = states_view(tensors)
        check(lib().lzb_self_play_step_inplace(ctypes.byref(view), i64(bn), ptr(plies), ptr(done), ptr(active),
                                               ptr(codes), ptr(term), ptr(cvalid), i64(k), i64(max_game_plies),
                                               ctypes.c_float(float(soft_value_k)), ptr(slots), ptr(result), ptr(soft),
                                               ptr(nfin), ptr(scratch), stream_ptr(dev)))
        f = int(nfin.item())
    return slots[:f], result[:f], soft[:f]


def finalize_trajectory_inplace(value_targets, soft_value_targets, player_signs, step_index_matrix, step_counts,
                                slots, result_from_black, soft_value_from_black):
    """Mutates value_targets / soft_value_targets; -> (final_slots, final_counts, counts_out i64[3]);
    module.cpp:547-630."""
    if value_targets.dim() != 1 or soft_value_targets.dim() != 1 or player_signs.dim() != 1:
        raise RuntimeError("value_targets / soft_value_targets / player_signs must be [S]")
    if step_index_matrix.dim() != 2 or step_counts.dim() != 1:
        raise RuntimeError("step_index_matrix must be [G, T] and step_counts [G]")
    require_cuda(value_targets, "value_targets")
    dev = value_targets.device
    for t in (soft_value_targets, player_signs, step_index_matrix, step_counts):
        if t.device != dev:
            raise RuntimeError("all tensors must be on the same device")
    if soft_value_targets.size(0) != value_targets.size(0) or player_signs.size(0) != value_targets.size(0):
        raise RuntimeError("target buffers shape mismatch")
    if step_counts.size(0) != step_index_matrix.size(0):
        raise RuntimeError("step_counts/step_index_matrix shape mismatch")
    if (value_targets.dtype != torch.float32 or soft_value_targets.dtype != torch.float32
            or not value_targets.is_contiguous() or not soft_value_targets.is_contiguous()):
        raise RuntimeError("target buffers must be contiguous float32")
    i64t = dict(dtype=torch.int64, device=dev)
    sl = slots.to(**i64t).reshape(-1).contiguous()
    res = result_from_black.to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
    soft = soft_value_from_black.to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
    f = sl.numel()
    counts_out = torch.zeros((3,), **i64t)
    if f == 0:
        return torch.empty((0,), **i64t), torch.empty((0,), **i64t), counts_out
    if res.numel() != f or soft.numel() != f:
        raise RuntimeError("result_from_black / soft_value_from_black must align with slots")
    signs = player_signs.to(torch.int8).contiguous()
    sim = step_index_matrix.to(torch.int64).contiguous()
    sc = step_counts.to(torch.int64).contiguous()
    with torch.cuda.device(dev):
        final_slots = torch.empty((f,), **i64t)
        final_counts = torch.empty((f,), **i64t)
        summary = torch.empty((4,), **i64t)
        check(lib().lzb_finalize_trajectory_inplace(ptr(value_targets), ptr(soft_value_targets), ptr(signs), ptr(sim),
                                                    i64(sim.size(0)), i64(sim.size(1)), ptr(sc), ptr(sl), ptr(res),
                                                    ptr(soft), i64(f), ptr(final_slots), ptr(final_counts),
                                                    ptr(summary), stream_ptr(dev)))
        kept = int(summary[0].item())
    if kept == 0:
        return torch.empty((0,), **i64t), torch.empty((0,), **i64t), counts_out
    return final_slots[:kept], final_counts[:kept], summary[1:4].clone()
