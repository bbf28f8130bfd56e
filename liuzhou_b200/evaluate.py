"""Device-resident checkpoint evaluation and round-robin tournament (BASELINE configs[4]).

Mirrors the reference's ``scripts/eval_checkpoint.py`` (worker loop :448-652, aggregation :73-124, entry
``evaluate_against_agent_parallel_v1`` :655-739) and the match scoring of ``scripts/tournament_v1_eval.py`` (3 / 1 / 0
points, :27-29, :233-262) -- but where the reference keeps one host ``GameState`` object per game and rebuilds tensors
every ply, here ALL games of a match live in HBM as packed bitboards and every ply is: terminal / no-move detection,
one batched search per agent (device tree, CUDA-graphed waves, ``active`` mask = games where that agent is to move),
a uniform-random pick for a ``RandomAgent`` opponent, one apply kernel.  No per-game host work, one host
synchronisation every ``sync_every`` plies (to see whether every game has finished).

Semantics kept from the reference: the game count is rounded up to an even number and the challenger plays black in
the first half (:48-55, :491-500); a game ends when ``get_winner()`` is set (win / loss), at the move limit
(``move_count >= 144`` or ``moves_since_capture >= 36``: draw), or when the side to move has no legal move (it
loses, :548-557); forced-uniform opening plies while ``move_count < opening_random_moves`` (:559-561).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch

from . import native
from .engine import packed_status
from .net import ChessNet, InferenceNet
from .tree_search import TreeMCTS, TreeMCTSConfig

WIN, LOSS, DRAW = 0, 1, 2


@dataclass
class EvaluationStats:
    """wins / losses / draws from the challenger's point of view (src/evaluate.py:228-248 + the colour breakdown the
    v1 evaluator attaches, eval_checkpoint.py:103-124)."""
    wins: int
    losses: int
    draws: int
    total_games: int
    seed: int = 0
    color_breakdown: Dict[str, Dict[str, int]] = field(default_factory=dict)
    plies: int = 0                     # positions played in the match (all games)
    searches: int = 0                  # MCTS-searched positions (network-guided moves)

    def _safe_rate(self, value: int) -> float:
        return 0.0 if self.total_games == 0 else value / self.total_games

    @property
    def win_rate(self) -> float:
        return self._safe_rate(self.wins)

    @property
    def loss_rate(self) -> float:
        return self._safe_rate(self.losses)

    @property
    def draw_rate(self) -> float:
        return self._safe_rate(self.draws)


def normalize_eval_games(num_games: int) -> int:
    """Even number of games, at least 2 (eval_checkpoint.py:48-55)."""
    n = max(2, int(num_games))
    return n + (n % 2)


class _SearchAgent:
    """One network + one device tree search over all games of the match (only `active` rows are searched)."""

    def __init__(self, model, num_slots: int, simulations: int, temperature: float, sample_moves: bool, device):
        self.net = model if isinstance(model, InferenceNet) else InferenceNet(model, device)
        self.temperature = float(temperature)
        self.sample_moves = bool(sample_moves)
        self.mcts = TreeMCTS(self.net, num_slots, TreeMCTSConfig(
            num_simulations=int(simulations), add_dirichlet_noise=False, sample_moves=bool(sample_moves),
            temperature=float(temperature)), device)
        self._temps = torch.full((num_slots,), float(temperature), dtype=torch.float32, device=self.net.device)

    def select(self, states: torch.Tensor, active: torch.Tensor) -> torch.Tensor:
        """Only the games where this agent moves are searched: their leaves are compacted into the first rows of the
        wave batch and the network runs on ceil(live) rows (one host read of the count per ply -- cheap next to the
        64+ network evaluations of a search)."""
        n_live = int(active.sum().item())
        if n_live == 0:
            return torch.full((active.numel(),), -1, dtype=torch.int64, device=active.device)
        self.mcts.set_live(active)
        out = self.mcts.search(states, active=active, temperatures=self._temps, add_dirichlet_noise=False,
                               sample_moves=self.sample_moves, live_rows=n_live)
        self.searched_rows = getattr(self, "searched_rows", 0) + self.mcts._bucket
        return out.chosen_action_indices


def _random_legal_actions(states: torch.Tensor) -> torch.Tensor:
    """Uniform pick among the legal actions of every state (RandomAgent: random.choice(legal)); -1 if none."""
    words, counts = native.legal_masks(states, scalar_semantics=True)
    mask = native.mask_words_to_bool(words).to(torch.float32)
    has = counts > 0
    safe = torch.where(has.view(-1, 1), mask, torch.ones_like(mask))
    pick = torch.multinomial(safe, num_samples=1).view(-1)
    return torch.where(has, pick, torch.full_like(pick, -1))


@torch.no_grad()
def play_match(challenger, opponent=None, *, num_games: int = 2000, mcts_simulations: int = 64,
               temperature: float = 0.0, sample_moves: bool = False, device="cuda:0", seed: int = 0,
               opening_random_moves: int = 0, max_plies: int = 1024, sync_every: int = 8,
               record_actions: bool = False):
    """All games of one challenger-vs-opponent match on one GPU.  `challenger` / `opponent` are ``ChessNet`` modules
    or ``InferenceNet`` wrappers; ``opponent=None`` is the uniform-random agent.  Returns ``EvaluationStats`` (and,
    with ``record_actions``, the int16[plies, G] action trace with -1 for games that did not move that ply)."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("liuzhou_b200.evaluate runs on CUDA devices only (no CPU fallback)")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    g = normalize_eval_games(num_games)
    slots = -(-g // 64) * 64                       # the tcgen05 conv path wants batches that are multiples of 64
    torch.manual_seed(int(seed))
    with torch.cuda.device(dev):
        states = native.init_states(slots, dev)
        chal_black = torch.arange(slots, device=dev) < (g // 2)
        done = torch.arange(slots, device=dev) >= g                # padding slots never play
        outcome = torch.full((slots,), -1, dtype=torch.int64, device=dev)
        a_chal = _SearchAgent(challenger, slots, mcts_simulations, temperature, sample_moves, dev)
        a_opp = None if opponent is None else _SearchAgent(opponent, slots, mcts_simulations, temperature,
                                                           sample_moves, dev)
        plies = torch.zeros((), dtype=torch.int64, device=dev)
        searched = torch.zeros((), dtype=torch.int64, device=dev)
        trace: List[torch.Tensor] = []
        win_t, loss_t, draw_t = (torch.full((slots,), v, dtype=torch.int64, device=dev) for v in (WIN, LOSS, DRAW))
        for it in range(int(max_plies)):
            over, winner = packed_status(states)
            meta = (states[:, 0] >> 36) & 0xFFFFFFF
            white_to_move = ((meta >> 3) & 1).to(torch.bool)
            move_count = (meta >> 14) & 255
            chal_to_move = white_to_move != chal_black
            # finished positions: winner (challenger's colour wins / loses) or move limit (draw)
            chal_won = torch.where(chal_black, winner > 0, winner < 0)
            res = torch.where(winner == 0, draw_t, torch.where(chal_won, win_t, loss_t))
            newly = ~done & over
            outcome = torch.where(newly, res, outcome)
            done = done | newly
            # side to move without a legal move loses
            _, counts = native.legal_masks(states, scalar_semantics=True)
            stuck = ~done & (counts == 0)
            outcome = torch.where(stuck, torch.where(chal_to_move, loss_t, win_t), outcome)
            done = done | stuck
            if it % max(1, int(sync_every)) == 0 and bool(done.all().item()):
                break
            live = ~done
            opening = live & (move_count < int(opening_random_moves))
            act_c = live & chal_to_move & ~opening
            act_o = live & ~chal_to_move & ~opening
            chosen = torch.full((slots,), -1, dtype=torch.int64, device=dev)
            need_random = opening if a_opp is not None else (opening | act_o)
            rnd = _random_legal_actions(states)
            chosen = torch.where(need_random, rnd, chosen)
            pick_c = a_chal.select(states, act_c)
            chosen = torch.where(act_c, pick_c, chosen)
            searched += act_c.sum()
            if a_opp is not None:
                pick_o = a_opp.select(states, act_o)
                chosen = torch.where(act_o, pick_o, chosen)
                searched += act_o.sum()
            # an agent that returns no move for a live game forfeits (eval_checkpoint.py:606-614)
            forfeit = live & (chosen < 0)
            outcome = torch.where(forfeit, torch.where(chal_to_move, loss_t, win_t), outcome)
            done = done | forfeit
            moving = live & ~forfeit
            nxt = native.apply_actions(states, chosen.clamp_min(0).to(torch.int32))
            states = torch.where(moving.view(-1, 1), nxt, states).contiguous()
            plies += moving.sum()
            if record_actions:
                trace.append(torch.where(moving, chosen, torch.full_like(chosen, -1)).to(torch.int16))
        # anything still running after max_plies is a draw (the reference's loop cannot get here: move limit)
        outcome = torch.where(done, outcome, draw_t)
        oc = outcome[:g].cpu()
        cb = chal_black[:g].cpu()
    stats = _stats_from_outcomes(oc, cb, seed)
    stats.plies = int(plies.item())
    stats.searches = int(searched.item())
    if record_actions:
        tr = torch.stack(trace)[:, :g].cpu() if trace else torch.zeros((0, g), dtype=torch.int16)
        return stats, tr
    return stats


def _stats_from_outcomes(outcomes: torch.Tensor, chal_black: torch.Tensor, seed: int) -> EvaluationStats:
    def count(mask, v):
        return int(((outcomes == v) & mask).sum())

    all_rows = torch.ones_like(chal_black)
    wins, losses, draws = (count(all_rows, v) for v in (WIN, LOSS, DRAW))
    total = int(outcomes.numel())
    if wins + losses + draws != total:
        raise ValueError(f"evaluation produced {wins + losses + draws} outcomes; expected {total} games")
    breakdown = {}
    for name, mask in (("challenger_black", chal_black), ("challenger_white", ~chal_black)):
        breakdown[name] = {"wins": count(mask, WIN), "losses": count(mask, LOSS), "draws": count(mask, DRAW),
                           "games": int(mask.sum())}
    return EvaluationStats(wins=wins, losses=losses, draws=draws, total_games=total, seed=int(seed),
                           color_breakdown=breakdown)


def evaluate_against_agent_parallel_v1(challenger_model, opponent_model=None, *, num_games: int = 2000,
                                       device: str = "cuda:0", mcts_simulations: int = 64, temperature: float = 0.0,
                                       seed: int = 0, opening_random_moves: int = 0, sample_moves: bool = False,
                                       **_ignored) -> EvaluationStats:
    """Name / keyword compatible front end of the reference's evaluator (eval_checkpoint.py:655-739): challenger vs a
    previous checkpoint (``opponent_model``) or vs the random agent (``None``)."""
    return play_match(challenger_model, opponent_model, num_games=num_games, mcts_simulations=mcts_simulations,
                      temperature=temperature, sample_moves=sample_moves, device=device, seed=seed,
                      opening_random_moves=opening_random_moves)


MATCH_POINTS_WIN, MATCH_POINTS_DRAW, MATCH_POINTS_LOSS = 3, 1, 0      # tournament_v1_eval.py:27-29


def round_robin_tournament(models: Sequence, *, names: Optional[Sequence[str]] = None, games_per_match: int = 1000,
                           mcts_simulations: int = 64, temperature: float = 1.0, sample_moves: bool = True,
                           device: str = "cuda:0", seed: int = 0) -> dict:
    """Every pair plays one colour-balanced match; standings by match points (3 / 1 / 0), then game win rate
    (tournament_v1_eval.py:132-165, :233-262).  Models are wrapped once and reused across matches."""
    dev = torch.device(device)
    nets = [m if isinstance(m, InferenceNet) else InferenceNet(m, dev) for m in models]
    names = list(names) if names is not None else [f"model_{i}" for i in range(len(nets))]
    table = [{"name": nm, "match_points": 0, "match_wins": 0, "match_draws": 0, "match_losses": 0, "game_wins": 0,
              "game_losses": 0, "game_draws": 0, "games": 0} for nm in names]
    matches = []
    k = 0
    for i in range(len(nets)):
        for j in range(i + 1, len(nets)):
            st = play_match(nets[i], nets[j], num_games=games_per_match, mcts_simulations=mcts_simulations,
                            temperature=temperature, sample_moves=sample_moves, device=dev, seed=seed + k)
            k += 1
            matches.append({"a": names[i], "b": names[j], "a_wins": st.wins, "b_wins": st.losses, "draws": st.draws,
                            "games": st.total_games})
            for row, w, l in ((table[i], st.wins, st.losses), (table[j], st.losses, st.wins)):
                row["game_wins"] += w
                row["game_losses"] += l
                row["game_draws"] += st.draws
                row["games"] += st.total_games
                if w > l:
                    row["match_points"] += MATCH_POINTS_WIN
                    row["match_wins"] += 1
                elif w == l:
                    row["match_points"] += MATCH_POINTS_DRAW
                    row["match_draws"] += 1
                else:
                    row["match_points"] += MATCH_POINTS_LOSS
                    row["match_losses"] += 1
    for row in table:
        row["game_win_rate"] = row["game_wins"] / max(1, row["games"])
    standings = sorted(table, key=lambda r: (-r["match_points"], -r["game_win_rate"], -r["game_wins"], r["name"]))
    return {"standings": standings, "matches": matches}


__all__ = ["EvaluationStats", "evaluate_against_agent_parallel_v1", "play_match", "round_robin_tournament",
           "normalize_eval_games", "ChessNet"]
