"""CPU tests of the host-side (torch tensor) logic that surrounds the kernels: policy from visit counts,
deterministic move choice, packed-state status decoding, trajectory buffer bookkeeping."""
import numpy as np
import pytest
import torch

import oracle
from tests._util import concat_states


def _pack_np(st, i):
    """numpy restatement of lz::pack (include/liuzhou_b200.h layout) for the test."""
    def bits(a):
        v = 0
        for c, x in enumerate(np.asarray(a).reshape(36)):
            if x:
                v |= 1 << c
        return v
    board = np.asarray(st["board"][i]).reshape(36)
    meta = (int(st["phase"][i]) & 7) | ((1 if st["current_player"][i] == -1 else 0) << 3) | \
           ((int(st["forced_removals_done"][i]) & 3) << 4) | ((int(st["pending_marks_required"][i]) & 3) << 6) | \
           ((int(st["pending_marks_remaining"][i]) & 3) << 8) | ((int(st["pending_captures_required"][i]) & 3) << 10) | \
           ((int(st["pending_captures_remaining"][i]) & 3) << 12) | ((int(st["move_count"][i]) & 255) << 14) | \
           ((int(st["moves_since_capture"][i]) & 63) << 22)
    w = [bits(board == 1) | (meta << 36), bits(board == -1), bits(st["marks_black"][i]), bits(st["marks_white"][i])]
    return [x - (1 << 64) if x >= (1 << 63) else x for x in w]


def test_packed_status_matches_oracle():
    from liuzhou_b200.engine import packed_status

    states = []
    for g in range(30):
        trace = oracle.random_playout(3, g, 512, want_trace=True)["trace"]
        st = oracle.initial_states(1)
        for a in trace:
            st = oracle.apply_move_scalar(st, int(a))
            states.append(st)
    st = concat_states(states)
    n = st["board"].shape[0]
    packed = torch.tensor([_pack_np(st, i) for i in range(n)], dtype=torch.int64)
    over, winner = packed_status(packed)
    exp_over = np.array([oracle.is_game_over(st, i) for i in range(n)])
    exp_win = np.array([oracle.winner(st, i) for i in range(n)])
    assert np.array_equal(over.numpy(), exp_over)
    assert np.array_equal(winner.numpy(), exp_win)
    assert exp_over.sum() >= 30


def test_policy_from_visits_and_deterministic_choice():
    from liuzhou_b200.tree_search import deterministic_action, policy_from_visits

    visits = torch.tensor([[0, 3, 1, 0], [5, 5, 0, 0], [0, 0, 0, 0], [2, 0, 0, 8]], dtype=torch.int32)
    temps = torch.tensor([1.0, 0.5, 1.0, 0.0])
    p = policy_from_visits(visits, temps)
    assert torch.allclose(p[0], torch.tensor([0.0, 0.75, 0.25, 0.0]))
    assert torch.allclose(p[1], torch.tensor([0.5, 0.5, 0.0, 0.0]))           # N^(1/T) normalised
    assert float(p[2].sum()) == 0.0                                             # no visits -> no mass
    assert torch.equal(p[3], torch.tensor([0.0, 0.0, 0.0, 1.0]))              # T <= 1e-6 -> one-hot argmax
    # softmax(log N / T) == N^(1/T) / sum N^(1/T)
    v = torch.tensor([[7, 1, 2, 0]], dtype=torch.int32)
    q = policy_from_visits(v, torch.tensor([0.25]))
    ref = v.float() ** 4
    assert torch.allclose(q, ref / ref.sum(), atol=1e-6)
    legal = torch.tensor([[True, True, True, False], [True, True, True, True]])
    vis = torch.tensor([[4, 4, 1, 9], [2, 2, 2, 2]], dtype=torch.int32)
    qv = torch.tensor([[0.1, 0.3, 0.9, 0.9], [0.2, 0.2 + 5e-7, 0.1, 0.2]])
    a = deterministic_action(vis, qv, legal)
    assert a.tolist() == [1, 0]       # max N among legal, then max Q (atol 1e-6), then lowest index


def test_trajectory_buffer_growth_and_views_cpu():
    """append / grow bookkeeping of TensorTrajectoryBuffer (finalisation itself is a CUDA kernel)."""
    from liuzhou_b200.trajectory_buffer import TensorTrajectoryBuffer

    buf = TensorTrajectoryBuffer("cpu", 220, max_steps_hint=2, concurrent_games_hint=2)
    rows = []
    for step in range(5):
        n = 3
        x = torch.full((n, 11, 6, 6), float(step))
        legal = torch.zeros((n, 220), dtype=torch.bool)
        legal[:, step] = True
        pol = torch.zeros((n, 220))
        pol[:, step] = 1.0
        idx = buf.append_steps(x, legal, pol, torch.tensor([1, -1, 1]))
        rows.append(idx)
        assert idx.tolist() == list(range(step * n, step * n + n))
    b = buf.build()
    assert b.num_samples == 15 and b.nbytes() == 15 * 2692
    assert torch.isnan(b.value_targets).all()
    assert float(b.state_tensors[7].mean()) == 2.0 and bool(b.legal_masks[7, 2])
    assert buf._player_signs[:3].tolist() == [1, -1, 1]
    empty = TensorTrajectoryBuffer("cpu", 220).build()
    assert empty.num_samples == 0 and tuple(empty.state_tensors.shape) == (0, 11, 6, 6)


def test_eval_game_count_normalisation_and_stats():
    """Host logic of liuzhou_b200/evaluate.py (no GPU): even game counts (eval_checkpoint.py:48-55), outcome
    aggregation with the colour breakdown (:73-124)."""
    import torch

    from liuzhou_b200.evaluate import DRAW, LOSS, WIN, _stats_from_outcomes, normalize_eval_games

    assert [normalize_eval_games(n) for n in (-3, 0, 1, 2, 3, 2000, 2001)] == [2, 2, 2, 2, 4, 2000, 2002]
    outcomes = torch.tensor([WIN, WIN, LOSS, DRAW, WIN, DRAW, LOSS, LOSS])
    black = torch.tensor([True] * 4 + [False] * 4)
    st = _stats_from_outcomes(outcomes, black, seed=9)
    assert (st.wins, st.losses, st.draws, st.total_games, st.seed) == (3, 3, 2, 8, 9)
    assert st.color_breakdown["challenger_black"] == {"wins": 2, "losses": 1, "draws": 1, "games": 4}
    assert st.color_breakdown["challenger_white"] == {"wins": 1, "losses": 2, "draws": 1, "games": 4}
    assert abs(st.win_rate - 0.375) < 1e-12 and abs(st.draw_rate - 0.25) < 1e-12


# ---- the three helper checks of the reference's tests/v1/test_v1_tensor_pipeline_smoke.py:20-70, same inputs and expectations,
# ---- on our mirror of V1RootMCTS (plain torch host logic: runs on the CPU)
def test_v1_soft_tanh_range_sign_and_scale():
    import math

    from liuzhou_b200.mcts_gpu import V1RootMCTS

    board = torch.zeros((3, 6, 6), dtype=torch.int8)
    board[0, :2, :] = 1
    board[1, :2, :] = -1
    soft = V1RootMCTS._soft_tanh_from_board_black(board, soft_value_k=2.0)
    assert tuple(soft.shape) == (3,)
    assert torch.all(soft <= 1.0 + 1e-6) and torch.all(soft >= -1.0 - 1e-6)
    assert float(soft[0]) > float(soft[2]) > float(soft[1])
    assert float(soft[0]) == pytest.approx(math.tanh(4.0 / 3.0))


def test_v1_terminal_mask_next_state():
    from liuzhou_b200.mcts_gpu import GpuStateBatch, V1RootMCTS

    board = torch.zeros((3, 6, 6), dtype=torch.int8)
    board[0, 0, 0] = 1          # white has no pieces, but mark selection is not a terminal phase
    board[1, 0, 0] = 1
    board[1, 0, 1] = -1
    board[2, 0, 0] = 1
    board[2, 0, 1] = -1
    zeros = torch.zeros((3,), dtype=torch.int64)
    batch = GpuStateBatch(
        board=board, marks_black=torch.zeros((3, 6, 6), dtype=torch.bool), marks_white=torch.zeros((3, 6, 6), dtype=torch.bool),
        phase=torch.tensor([2, 4, 1], dtype=torch.int64), current_player=torch.ones((3,), dtype=torch.int64),
        pending_marks_required=zeros.clone(), pending_marks_remaining=zeros.clone(), pending_captures_required=zeros.clone(),
        pending_captures_remaining=zeros.clone(), forced_removals_done=zeros.clone(),
        move_count=torch.tensor([0, 144, 0], dtype=torch.int64),            # GameState.MAX_MOVE_COUNT (game_state.hpp:14)
        moves_since_capture=torch.tensor([0, 0, 36], dtype=torch.int64))    # NO_CAPTURE_DRAW_LIMIT (game_state.hpp:16)
    assert V1RootMCTS._terminal_mask_from_next_state(batch).tolist() == [False, True, True]


def test_v1_child_value_perspective_alignment():
    from liuzhou_b200.mcts_gpu import V1RootMCTS

    aligned = V1RootMCTS._child_values_to_parent_perspective(
        child_values=torch.tensor([0.2, -0.5, 0.8, -0.1], dtype=torch.float32),
        parent_players=torch.tensor([1, 1, -1, -1], dtype=torch.int64),
        child_players=torch.tensor([1, -1, -1, 1], dtype=torch.int64))
    assert torch.allclose(aligned, torch.tensor([0.2, 0.5, 0.8, 0.1]), atol=1e-6, rtol=0.0)
