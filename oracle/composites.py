"""CPU ORACLE (test infrastructure) -- numpy restatements of the reference's ATen composite ops.

Each function follows /root/reference/v0/src/bindings/module.cpp (line ranges cited per function) and
returns numpy arrays with the reference's dtypes / shapes / ordering.  fp32 arithmetic is kept in
np.float32 with the reference's operation order; transcendental ops (pow, tanh, softmax) differ from
ATen's vectorised kernels by a few ulp, so tests compare those outputs with the tolerance stated there.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def root_pack_sparse_actions(legal_mask, probs, metadata):
    """module.cpp:258-363.  Returns the reference's 10-tuple."""
    legal = np.asarray(legal_mask).astype(bool)
    probs = np.asarray(probs, dtype=F32)
    meta = np.asarray(metadata, dtype=np.int32)
    b, a = legal.shape
    row_counts = legal.sum(1).astype(np.int64)
    terminal_mask = row_counts == 0
    valid_root_indices = np.nonzero(~terminal_mask)[0].astype(np.int64)
    counts = row_counts[valid_root_indices]
    r = int(valid_root_indices.size)
    if r == 0:
        return (terminal_mask, valid_root_indices, counts, np.zeros((0, 0), bool), np.zeros((0, 0), np.int64),
                np.zeros((0, 0), F32), np.zeros((0, 0, 4), np.int32), np.zeros((0,), np.int64),
                np.zeros((0, 4), np.int32), np.zeros((0,), np.int64))
    m = int(counts.max())
    valid_mask = np.zeros((r, m), bool)
    legal_index_mat = np.zeros((r, m), np.int64)
    priors_mat = np.zeros((r, m), F32)
    action_code_mat = np.zeros((r, m, 4), np.int32)
    flat, codes_all, parents = [], [], []
    for row, root in enumerate(valid_root_indices):
        idx = np.nonzero(legal[root])[0]
        k = idx.size
        valid_mask[row, :k] = True
        legal_index_mat[row, :k] = idx
        priors_mat[row, :k] = probs[root, idx]
        action_code_mat[row, :k] = meta[root, idx]
        flat.append(row * m + np.arange(k, dtype=np.int64))
        codes_all.append(meta[root, idx])
        parents.append(np.full((k,), root, np.int64))
    # priors_mat / priors_mat.sum(1, keepdim).clamp_min(1e-8)   (:333-335); ATen sums rows in fp32
    sums = priors_mat.sum(1, keepdims=True, dtype=F32)
    priors_mat = (priors_mat / np.maximum(sums, F32(1e-8))).astype(F32)
    return (terminal_mask, valid_root_indices, counts, valid_mask, legal_index_mat, priors_mat, action_code_mat,
            np.concatenate(flat), np.concatenate(codes_all).astype(np.int32), np.concatenate(parents))


def root_finalize_from_visits(legal_index_mat, action_code_mat, valid_mask, visits, value_sum,
                              valid_root_indices, batch_size, total_action_dim, root_temperatures):
    """module.cpp:441-535 with sample_moves=False (argmax pick; v1 always passes False, mcts_gpu.py:1408)."""
    legal_idx = np.asarray(legal_index_mat, dtype=np.int64)
    codes = np.asarray(action_code_mat, dtype=np.int32)
    mask = np.asarray(valid_mask).astype(bool)
    visits = np.asarray(visits, dtype=F32)
    value_sum = np.asarray(value_sum, dtype=F32)
    roots = np.asarray(valid_root_indices, dtype=np.int64)
    temps = np.maximum(np.asarray(root_temperatures, dtype=F32), F32(1e-6))
    r, m = legal_idx.shape
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        expo = (F32(1.0) / temps).reshape(-1, 1).astype(F32)
        legal_policy = np.power(np.maximum(visits, F32(1e-8)), expo, dtype=F32)
        legal_policy = (legal_policy * mask.astype(F32)).astype(F32)
        legal_policy = (legal_policy / np.maximum(legal_policy.sum(1, keepdims=True, dtype=F32), F32(1e-8))).astype(F32)
    # torch.max(dim) returns the first maximal index; NaN compares as maximal in ATen
    picks = np.zeros((r,), np.int64)
    for i in range(r):
        row = legal_policy[i]
        nan = np.isnan(row)
        picks[i] = int(np.argmax(nan)) if nan.any() else int(np.argmax(row))
    root_value = (value_sum.sum(1, dtype=F32) / np.maximum(visits.sum(1, dtype=F32), F32(1.0))).astype(F32)
    policy_dense = np.zeros((batch_size, total_action_dim), F32)
    chosen_idx = np.full((batch_size,), -1, np.int64)
    chosen_codes = np.full((batch_size, 4), -1, np.int32)
    chosen_valid = np.zeros((batch_size,), bool)
    for i in range(r):
        dense = np.zeros((total_action_dim,), F32)
        np.add.at(dense, legal_idx[i], (legal_policy[i] * mask[i].astype(F32)).astype(F32))
        policy_dense[roots[i]] = dense
        chosen_idx[roots[i]] = legal_idx[i, picks[i]]
        chosen_codes[roots[i]] = codes[i, picks[i]]
        chosen_valid[roots[i]] = True
    return policy_dense, chosen_idx, chosen_codes, chosen_valid, root_value


def root_sparse_writeback(legal_index_mat, action_code_mat, valid_mask, legal_policy, local_picks,
                          valid_root_indices, batch_size: int, total_action_dim: int):
    """module.cpp:365-439: dense scatter of a legal policy [R,M] (x valid_mask, repeated indices add up as in
    scatter_add_) and of the picked column per root; rows that are not valid roots stay 0 / -1 / False."""
    li = np.asarray(legal_index_mat, dtype=np.int64)
    ac = np.asarray(action_code_mat, dtype=np.int32)
    vm = np.asarray(valid_mask).astype(bool)
    lp = np.asarray(legal_policy, dtype=F32)
    picks = np.asarray(local_picks, dtype=np.int64).reshape(-1)
    roots = np.asarray(valid_root_indices, dtype=np.int64).reshape(-1)
    r = li.shape[0]
    policy_dense = np.zeros((batch_size, total_action_dim), F32)
    chosen_idx = np.full((batch_size,), -1, np.int64)
    chosen_codes = np.full((batch_size, 4), -1, np.int32)
    chosen_valid = np.zeros((batch_size,), bool)
    for i in range(r):
        row = np.zeros((total_action_dim,), F32)
        np.add.at(row, li[i], (lp[i] * vm[i].astype(F32)).astype(F32))
        policy_dense[roots[i]] = row
        chosen_idx[roots[i]] = li[i, picks[i]]
        chosen_codes[roots[i]] = ac[i, picks[i]]
        chosen_valid[roots[i]] = True
    return policy_dense, chosen_idx, chosen_codes, chosen_valid


def soft_value_from_board(boards, soft_value_k: float):
    """module.cpp:537-545 == mcts_gpu.py:677-686: tanh(k * (black - white) / 18)."""
    boards = np.asarray(boards).reshape(-1, 36)
    black = (boards == 1).sum(1).astype(F32)
    white = (boards == -1).sum(1).astype(F32)
    delta = ((black - white) / F32(18.0)).astype(F32)
    return np.tanh((delta * F32(soft_value_k)).astype(F32)).astype(F32)


def terminal_mask_from_next_state(st: dict):
    """mcts_gpu.py:658-675."""
    phase = np.asarray(st["phase"])
    post = (phase == 4) | (phase == 5) | (phase == 7)
    board = np.asarray(st["board"]).reshape(-1, 36)
    black = (board == 1).sum(1)
    white = (board == -1).sum(1)
    winner = post & ((black < 4) | (white < 4))
    draw = (np.asarray(st["move_count"]) >= 144) | (np.asarray(st["moves_since_capture"]) >= 36)
    return winner | draw


def self_play_step_inplace(st: dict, plies, done, active_idx, chosen_action_codes, terminal_mask,
                           chosen_valid_mask, max_game_plies: int, soft_value_k: float):
    """module.cpp:632-871.  Mutates st / plies / done (numpy, in place) and returns
    (finalize_slots i64[F], result_from_black f32[F], soft_value_from_black f32[F])."""
    from . import batch_apply_moves  # local import: package init order

    active_idx = np.asarray(active_idx, dtype=np.int64).reshape(-1)
    codes = np.asarray(chosen_action_codes, dtype=np.int32).reshape(-1, 4)
    terminal_mask = np.asarray(terminal_mask).astype(bool).reshape(-1)
    chosen_valid_mask = np.asarray(chosen_valid_mask).astype(bool).reshape(-1)
    slots_out, res_out, soft_out = [], [], []
    if active_idx.size == 0:
        return np.zeros((0,), np.int64), np.zeros((0,), F32), np.zeros((0,), F32)
    immediate = terminal_mask | ~chosen_valid_mask
    imm_idx = np.nonzero(immediate)[0]
    if imm_idx.size:
        imm_slots = active_idx[imm_idx]
        done[imm_slots] = True
        player = np.asarray(st["current_player"])[imm_slots].astype(F32)
        res = np.where(terminal_mask[imm_idx], -player, F32(0.0)).astype(F32)
        slots_out.append(imm_slots)
        res_out.append(res)
        soft_out.append(soft_value_from_board(np.asarray(st["board"])[imm_slots], soft_value_k))
    val_idx = np.nonzero(~immediate)[0]
    if val_idx.size:
        val_slots = active_idx[val_idx]
        nxt = batch_apply_moves(st, codes[val_idx], val_slots)
        for name, arr in nxt.items():
            st[name][val_slots] = arr.reshape((val_slots.size,) + st[name].shape[1:])
        plies[val_slots] += 1
        phase = nxt["phase"]
        post = (phase == 4) | (phase == 5) | (phase == 7)
        board = nxt["board"].reshape(-1, 36)
        black = (board == 1).sum(1)
        white = (board == -1).sum(1)
        winner_sign = np.zeros((val_slots.size,), np.int8)
        winner_sign = np.where(post & (black < 4), np.int8(-1), winner_sign)   # :817-820
        winner_sign = np.where(post & (white < 4), np.int8(1), winner_sign)    # :821-824 (overrides)
        draw = (nxt["move_count"] >= 144) | (nxt["moves_since_capture"] >= 36)
        hit = plies[val_slots] >= max_game_plies
        fin = (winner_sign != 0) | draw | hit
        fin_idx = np.nonzero(fin)[0]
        if fin_idx.size:
            fin_slots = val_slots[fin_idx]
            done[fin_slots] = True
            slots_out.append(fin_slots)
            res_out.append(winner_sign[fin_idx].astype(F32))
            soft_out.append(soft_value_from_board(board[fin_idx], soft_value_k))
    if not slots_out:
        return np.zeros((0,), np.int64), np.zeros((0,), F32), np.zeros((0,), F32)
    return np.concatenate(slots_out), np.concatenate(res_out).astype(F32), np.concatenate(soft_out).astype(F32)


def finalize_trajectory_inplace(value_targets, soft_value_targets, player_signs, step_index_matrix, step_counts,
                                slots, result_from_black, soft_value_from_black):
    """module.cpp:547-630.  Mutates value_targets / soft_value_targets; returns
    (final_slots, final_counts, counts_out[3] = black wins, white wins, draws)."""
    counts_out = np.zeros((3,), np.int64)
    slots = np.asarray(slots, dtype=np.int64).reshape(-1)
    res = np.asarray(result_from_black, dtype=F32).reshape(-1)
    soft = np.asarray(soft_value_from_black, dtype=F32).reshape(-1)
    empty = np.zeros((0,), np.int64)
    if slots.size == 0:
        return empty, empty, counts_out
    counts = np.asarray(step_counts, dtype=np.int64)[slots]
    keep = np.nonzero(counts > 0)[0]
    if keep.size == 0:
        return empty, empty, counts_out
    fs, fc, fr, fsoft = slots[keep], counts[keep], res[keep], soft[keep]
    counts_out[0] = int((fr > 0).sum())
    counts_out[1] = int((fr < 0).sum())
    counts_out[2] = int((fr == 0).sum())
    sim = np.asarray(step_index_matrix, dtype=np.int64)
    signs = np.asarray(player_signs).astype(np.int8)
    for g, n, r, s in zip(fs, fc, fr, fsoft):
        rows = sim[g, :n]
        sg = signs[rows].astype(F32)
        value_targets[rows] = sg * r
        soft_value_targets[rows] = sg * s
    return fs, fc, counts_out


def project_policy_logits_fast(log_p1, log_p2, log_pmc, legal_mask, placement_dim=36, movement_dim=144,
                               selection_dim=36, auxiliary_dim=4):
    """v0/src/net/project_policy_logits_fast.cpp:16-164 (fp32)."""
    log_p1 = np.asarray(log_p1, dtype=F32)
    log_p2 = np.asarray(log_p2, dtype=F32)
    log_pmc = np.asarray(log_pmc, dtype=F32)
    legal = np.asarray(legal_mask).astype(bool)
    b = log_p1.shape[0]
    size = int(round(placement_dim ** 0.5))
    total = placement_dim + movement_dim + selection_dim + auxiliary_dim
    combined = np.zeros((b, total), F32)
    combined[:, :placement_dim] = log_p1
    dirs = ((-1, 0), (1, 0), (0, -1), (0, 1))
    for cell in range(placement_dim):
        r, c = divmod(cell, size)
        for d, (dr, dc) in enumerate(dirs):
            nr, nc = r + dr, c + dc
            col = placement_dim + cell * 4 + d
            if 0 <= nr < size and 0 <= nc < size:
                combined[:, col] = log_p2[:, cell] + log_p1[:, nr * size + nc]
            else:
                combined[:, col] = -np.inf
    combined[:, placement_dim + movement_dim: placement_dim + movement_dim + selection_dim] = log_pmc
    masked = np.where(legal, combined, F32(-np.inf)).astype(F32)
    probs = np.zeros((b, total), F32)
    for i in range(b):
        if not legal[i].any():
            continue
        row = masked[i]
        if np.isfinite(row).any():
            mx = row.max()
            e = np.exp((row - mx).astype(F32), dtype=F32)
            probs[i] = (e / e.sum(dtype=F32)).astype(F32)
        else:
            masked[i] = np.where(legal[i], F32(0.0), row)
    return probs, masked
