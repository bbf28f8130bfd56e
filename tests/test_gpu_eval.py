"""Device-resident evaluation / tournament (liuzhou_b200/evaluate.py) -- bookkeeping checked by replaying the
recorded action traces of every game on the CPU oracle's scalar rule engine (oracle/lz_oracle.c), which restates
v0/src/rules/rule_engine.cpp: same terminal detection, same winner, same colour assignment."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _tiny_net(seed):
    from liuzhou_b200.net import ChessNet

    torch.manual_seed(seed)
    from liuzhou_b200.net import InferenceNet

    return InferenceNet(ChessNet(trunk_channels=8, num_blocks=1, policy_channels=4, value_channels=4,
                                 value_mlp_channels=8), DEV, allow_library_convs=True)


def _replay(trace, g):
    """-> list of (outcome for the challenger in {0 win, 1 loss, 2 draw}, plies) per game, from the oracle."""
    out = []
    half = g // 2
    for game in range(g):
        st = oracle.initial_states(1)
        chal_black = game < half
        res, n = None, 0
        for ply in range(trace.shape[0] + 1):
            if oracle.is_game_over(st):
                w = oracle.winner(st)
                res = 2 if w == 0 else (0 if (w > 0) == chal_black else 1)
                break
            legal = oracle.legal_actions(st, 0)[0]
            white_to_move = int(st["current_player"][0]) == -1
            chal_to_move = white_to_move != chal_black
            if len(legal) == 0:
                res = 1 if chal_to_move else 0
                break
            if ply == trace.shape[0]:
                break
            a = int(trace[ply, game])
            if a < 0:
                continue                          # the device loop did not move this game this ply (cannot happen while live)
            assert a in legal, (game, ply, a)
            st = oracle.apply_move_scalar(st, a)
            n += 1
        out.append((2 if res is None else res, n))
    return out


@pytest.mark.parametrize("vs_random", [True, False])
def test_match_outcomes_replay_on_oracle(vs_random):
    from liuzhou_b200.evaluate import play_match

    g = 30
    stats, trace = play_match(_tiny_net(1), None if vs_random else _tiny_net(2), num_games=g, mcts_simulations=8,
                              temperature=0.0, sample_moves=False, device=DEV, seed=5, record_actions=True,
                              sync_every=4)
    assert stats.total_games == g and stats.wins + stats.losses + stats.draws == g
    cb = stats.color_breakdown
    assert cb["challenger_black"]["games"] == cb["challenger_white"]["games"] == g // 2
    ref = _replay(trace.numpy(), g)
    wins = sum(1 for r, _ in ref if r == 0)
    losses = sum(1 for r, _ in ref if r == 1)
    draws = sum(1 for r, _ in ref if r == 2)
    assert (stats.wins, stats.losses, stats.draws) == (wins, losses, draws)
    assert stats.plies == sum(n for _, n in ref)
    bw = sum(1 for i, (r, _) in enumerate(ref) if r == 0 and i < g // 2)
    assert cb["challenger_black"]["wins"] == bw


def test_game_count_is_made_even_and_opening_moves_are_random():
    from liuzhou_b200.evaluate import normalize_eval_games, play_match

    assert [normalize_eval_games(n) for n in (0, 1, 2, 7, 8)] == [2, 2, 2, 8, 8]
    stats, trace = play_match(_tiny_net(3), None, num_games=7, mcts_simulations=4, device=DEV, seed=1,
                              opening_random_moves=6, record_actions=True)
    assert stats.total_games == 8
    # during the forced-uniform opening the 8 games must not all play the same first move
    assert len(set(trace[0].tolist())) > 1
    assert stats.searches < stats.plies


def test_round_robin_tournament_table():
    from liuzhou_b200.evaluate import round_robin_tournament

    res = round_robin_tournament([_tiny_net(s) for s in (1, 2, 3)], names=["a", "b", "c"], games_per_match=8,
                                 mcts_simulations=4, temperature=1.0, sample_moves=True, device=DEV, seed=0)
    assert len(res["matches"]) == 3
    assert sum(r["games"] for r in res["standings"]) == 2 * 3 * 8
    for m in res["matches"]:
        assert m["a_wins"] + m["b_wins"] + m["draws"] == m["games"] == 8
    pts = [r["match_points"] for r in res["standings"]]
    assert pts == sorted(pts, reverse=True)
    total_pts = sum(pts)
    decisive = sum(1 for m in res["matches"] if m["a_wins"] != m["b_wins"])
    assert total_pts == 3 * decisive + 2 * (3 - decisive)
