"""Wave-batched self-play entry -- call-compatible with the reference's ``self_play_v1_gpu``
(v1/python/self_play_gpu_runner.py:21-307): same arguments, same ``(TensorSelfPlayBatch, SelfPlayV1Stats)`` result,
same trajectory format and schedule (per-game temperature switch at ``temperature_threshold`` plies, forced-uniform
opening, Dirichlet noise every ply, no refill of finished games inside a wave).

``search_backend``:
  "root"  reference production semantics (root-PUCT over all children, V1RootMCTS.search_batch);
  "tree"  full MCTS on the device tree (select / expand / backup kernels), the search north_star asks for; mirrors
          the reference's portable self-play (v1/python/portable_cpp_self_play.py): subtree reuse after every move
          (``tree_reuse``), policy target temperature / prior pseudocount (``policy_target_*``, tree backend only).
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import Any, Dict, Optional, Tuple

import torch

from . import native, v0_core
from .mcts_gpu import TOTAL_ACTION_DIM, GpuStateBatch, V1RootMCTS, V1RootMCTSConfig
from .net import InferenceNet
from .trajectory_buffer import TensorSelfPlayBatch, TensorTrajectoryBuffer
from .tree import encode_inputs
from .tree_search import TreeMCTS, TreeMCTSConfig

_PIECE_DELTA_MIN, _PIECE_DELTA_MAX = -18, 18


@dataclass
class SelfPlayV1Stats:
    """v1/python/self_play_types.py:11-53."""

    num_games: int
    num_positions: int
    black_wins: int
    white_wins: int
    draws: int
    avg_game_length: float
    elapsed_sec: float
    positions_per_sec: float
    games_per_sec: float
    step_timing_ms: Dict[str, float] = field(default_factory=dict)
    step_timing_ratio: Dict[str, float] = field(default_factory=dict)
    step_timing_calls: Dict[str, int] = field(default_factory=dict)
    mcts_counters: Dict[str, int] = field(default_factory=dict)
    piece_delta_buckets: Dict[str, int] = field(default_factory=dict)
    policy_target_audit: Dict[str, Any] = field(default_factory=dict)
    device: str = ""
    fallback_count: int = 0
    fallback_reasons: Tuple[str, ...] = ()

    def to_dict(self) -> Dict[str, object]:
        d: Dict[str, object] = {k: float(getattr(self, k)) for k in (
            "num_games", "num_positions", "black_wins", "white_wins", "draws", "avg_game_length", "elapsed_sec",
            "positions_per_sec", "games_per_sec")}
        d["step_timing_ms"] = {k: float(v) for k, v in self.step_timing_ms.items()}
        d["step_timing_ratio"] = {k: float(v) for k, v in self.step_timing_ratio.items()}
        d["step_timing_calls"] = {k: int(v) for k, v in self.step_timing_calls.items()}
        d["mcts_counters"] = {k: int(v) for k, v in self.mcts_counters.items()}
        d["piece_delta_buckets"] = {k: int(v) for k, v in self.piece_delta_buckets.items()}
        d["policy_target_audit"] = dict(self.policy_target_audit or {})
        d["device"] = str(self.device)
        d["fallback_count"] = int(self.fallback_count)
        d["fallback_reasons"] = list(self.fallback_reasons)
        return d


def _soft_tanh_from_board_black(board: torch.Tensor, k: float) -> torch.Tensor:
    black = board.eq(1).sum(dim=(1, 2)).to(torch.float32)
    white = board.eq(-1).sum(dim=(1, 2)).to(torch.float32)
    return torch.tanh((black - white) / 18.0 * float(k))


def self_play_v1_gpu(
    model,
    num_games: int,
    mcts_simulations: int,
    temperature_init: float,
    temperature_final: float,
    temperature_threshold: int,
    exploration_weight: float,
    device: str,
    add_dirichlet_noise: bool = True,
    dirichlet_alpha: float = 0.3,
    dirichlet_epsilon: float = 0.25,
    soft_value_k: float = 2.0,
    opening_random_moves: int = 0,
    max_game_plies: int = 512,
    sample_moves: bool = True,
    concurrent_games: int = 8,
    child_eval_mode: str = "value_only",
    sparse_ply: int = 1,
    sparse_top_k: int = 8,
    inference_engine=None,
    collect_step_timing: bool = False,
    verbose: bool = False,
    search_backend: str = "root",
    leaves_per_wave: int = 1,
    tree_reuse: bool = True,
    policy_target_temperature: Optional[float] = None,
    policy_target_prior_pseudocount: float = 0.0,
) -> Tuple[TensorSelfPlayBatch, SelfPlayV1Stats]:
    if num_games <= 0:
        raise ValueError("num_games must be positive.")
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("liuzhou_b200.self_play_v1_gpu runs on CUDA devices only (no CPU fallback)")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    if search_backend not in ("root", "tree"):
        raise ValueError("search_backend must be 'root' or 'tree'")
    max_plies = max(1, int(max_game_plies))
    opening_random_n = max(0, int(opening_random_moves))
    wave_size = max(1, min(int(concurrent_games), int(num_games)))
    sims = max(1, int(mcts_simulations))
    net = model if isinstance(model, InferenceNet) else InferenceNet(model, dev)

    root_mcts = tree_mcts = None
    if search_backend == "root":
        root_mcts = V1RootMCTS(net, V1RootMCTSConfig(
            num_simulations=sims, exploration_weight=float(exploration_weight), temperature=float(temperature_init),
            add_dirichlet_noise=bool(add_dirichlet_noise), dirichlet_alpha=float(dirichlet_alpha),
            dirichlet_epsilon=float(dirichlet_epsilon), sample_moves=bool(sample_moves),
            child_eval_mode=str(child_eval_mode), soft_value_k=float(soft_value_k), sparse_ply=int(sparse_ply),
            sparse_top_k=int(sparse_top_k)), dev)

    buffer = TensorTrajectoryBuffer(dev, TOTAL_ACTION_DIM, max_steps_hint=min(max_plies, 160),
                                    concurrent_games_hint=min(wave_size, num_games))
    outcome_counts = torch.zeros((3,), dtype=torch.int64, device=dev)
    piece_delta_hist = torch.zeros((_PIECE_DELTA_MAX - _PIECE_DELTA_MIN + 1,), dtype=torch.int64, device=dev)
    game_lengths = torch.zeros((int(num_games),), dtype=torch.int64, device=dev)
    started = time.perf_counter()
    evals = 0

    for wave_base in range(0, int(num_games), wave_size):
        wave_games = min(wave_size, int(num_games) - wave_base)
        states = GpuStateBatch.initial(dev, batch_size=wave_games)
        step_index_matrix = torch.full((wave_games, max_plies), -1, dtype=torch.int64, device=dev)
        step_counts = torch.zeros((wave_games,), dtype=torch.int64, device=dev)
        plies = torch.zeros((wave_games,), dtype=torch.int64, device=dev)
        done = torch.zeros((wave_games,), dtype=torch.bool, device=dev)
        if search_backend == "tree" and (tree_mcts is None or tree_mcts.num_trees != wave_games):
            tree_mcts = TreeMCTS(net, wave_games, TreeMCTSConfig(
                num_simulations=sims, exploration_weight=float(exploration_weight),
                add_dirichlet_noise=bool(add_dirichlet_noise), dirichlet_alpha=float(dirichlet_alpha),
                dirichlet_epsilon=float(dirichlet_epsilon), sample_moves=bool(sample_moves),
                leaves_per_wave=int(leaves_per_wave), reuse_subtree=bool(tree_reuse),
                policy_target_temperature=policy_target_temperature,
                policy_target_prior_pseudocount=float(policy_target_prior_pseudocount)), dev)
        if tree_mcts is not None:
            tree_mcts._advanced = False          # a new wave of games starts from reset roots

        while True:
            active_idx = torch.where(~done)[0]
            n_active = int(active_idx.numel())
            if n_active == 0:
                break
            active_plies = plies.index_select(0, active_idx)
            if search_backend == "root":
                active_states = states.select(active_idx)
                temps = torch.where(active_plies < int(temperature_threshold), float(temperature_init),
                                    float(temperature_final)).to(torch.float32)
                force_uniform = active_plies < opening_random_n if opening_random_n > 0 else None
                search = root_mcts.search_batch(active_states, temperatures=temps,
                                                add_dirichlet_noise=add_dirichlet_noise,
                                                force_uniform_random_mask=force_uniform)
                model_input, legal_mask, policy_dense = search.model_input, search.legal_mask, search.policy_dense
                player_sign = active_states.current_player
                chosen_codes, terminal_mask, chosen_valid = (search.chosen_action_codes, search.terminal_mask,
                                                             search.chosen_valid_mask)
            else:
                # static shapes: all wave_games trees are searched, finished games are inactive trees
                packed = native.pack_states(states.tensors())
                temps_all = torch.where(plies < int(temperature_threshold), float(temperature_init),
                                        float(temperature_final)).to(torch.float32)
                out = tree_mcts.search(packed, active=~done, temperatures=temps_all,
                                       add_dirichlet_noise=add_dirichlet_noise, sample_moves=sample_moves)
                chosen_all = out.chosen_action_indices
                if opening_random_n > 0:
                    force = (plies < opening_random_n) & ~out.terminal_mask
                    lm = out.legal_mask.to(torch.float32)
                    uni = lm / lm.sum(dim=1, keepdim=True).clamp_min(1.0)
                    uni = torch.where(lm.sum(dim=1, keepdim=True) > 0, uni, torch.full_like(uni, 1.0 / 220))
                    chosen_all = torch.where(force, torch.multinomial(uni, 1).view(-1), chosen_all)
                mask_all, meta_all = v0_core.encode_actions_fast(*states.tensors()[:10], 36, 144, 36, 4)
                chosen_a = chosen_all.index_select(0, active_idx)
                model_input = encode_inputs(packed.index_select(0, active_idx), "f32_nchw")
                legal_mask = mask_all.index_select(0, active_idx)
                policy_dense = out.policy_dense.index_select(0, active_idx)
                player_sign = states.current_player.index_select(0, active_idx)
                terminal_mask = out.terminal_mask.index_select(0, active_idx)
                chosen_valid = chosen_a >= 0
                meta_a = meta_all.index_select(0, active_idx)
                chosen_codes = meta_a.gather(1, chosen_a.clamp_min(0).view(-1, 1, 1).expand(-1, 1, 4)).view(-1, 4)
                chosen_codes = torch.where(chosen_valid.view(-1, 1), chosen_codes, torch.full_like(chosen_codes, -1))
                # subtree reuse (portable_cpp_self_play.py:170): played child -> root; finished / inactive games keep -1
                tree_mcts.advance(torch.where(done | out.terminal_mask, torch.full_like(chosen_all, -1), chosen_all))

            step_indices = buffer.append_steps(model_input=model_input, legal_mask=legal_mask,
                                               policy_dense=policy_dense, player_sign=player_sign)
            step_positions = step_counts.index_select(0, active_idx)
            step_index_matrix[active_idx, step_positions] = step_indices
            step_counts.index_add_(0, active_idx, torch.ones((n_active,), dtype=torch.int64, device=dev))
            finalize_slots, result_local, soft_local = v0_core.self_play_step_inplace(
                *states.tensors(), plies, done, active_idx, chosen_codes, terminal_mask, chosen_valid, int(max_plies),
                float(soft_value_k))
            if int(finalize_slots.numel()) > 0:
                final_boards = states.board.index_select(0, finalize_slots)
                delta = (final_boards.eq(1).sum(dim=(1, 2)) - final_boards.eq(-1).sum(dim=(1, 2))).to(torch.int64)
                piece_delta_hist.add_(torch.bincount(torch.clamp(delta - _PIECE_DELTA_MIN, 0, piece_delta_hist.numel() - 1),
                                                     minlength=piece_delta_hist.numel()))
                soft_local = _soft_tanh_from_board_black(final_boards, float(soft_value_k))
                fin_slots, fin_lengths, outcome_delta = buffer.finalize_games_inplace(
                    step_index_matrix=step_index_matrix, step_counts=step_counts, slots=finalize_slots,
                    result_from_black=result_local, soft_value_from_black=soft_local)
                if int(fin_slots.numel()) > 0:
                    game_lengths.index_copy_(0, fin_slots + int(wave_base), fin_lengths)
                outcome_counts.add_(outcome_delta)
        if verbose:
            oc = outcome_counts.tolist()
            print(f"[v1.self_play] games={min(wave_base + wave_games, num_games)}/{num_games} W/L/D={oc[0]}/{oc[1]}/{oc[2]}")

    torch.cuda.synchronize(dev)
    elapsed = max(1e-9, time.perf_counter() - started)
    batch = buffer.build()
    oc = outcome_counts.tolist()
    hist = piece_delta_hist.cpu().tolist()
    if tree_mcts is not None:
        evals = tree_mcts.evals
    counters = {"network_evals": int(evals)}
    if root_mcts is not None:
        counters.update(root_mcts.get_timing()["counters"])
    keys = ("root_puct_ms", "pack_writeback_ms", "self_play_step_ms", "finalize_ms")
    stats = SelfPlayV1Stats(
        num_games=int(num_games), num_positions=batch.num_samples, black_wins=int(oc[0]), white_wins=int(oc[1]),
        draws=int(oc[2]), avg_game_length=float(game_lengths.to(torch.float32).mean().item()), elapsed_sec=elapsed,
        positions_per_sec=float(batch.num_samples / elapsed), games_per_sec=float(num_games / elapsed),
        step_timing_ms={k: 0.0 for k in keys}, step_timing_ratio={k: 0.0 for k in keys},
        step_timing_calls={k: 0 for k in keys}, mcts_counters=counters,
        piece_delta_buckets={str(d): int(hist[d - _PIECE_DELTA_MIN]) for d in range(_PIECE_DELTA_MIN, _PIECE_DELTA_MAX + 1)},
        device=str(dev))
    return batch, stats
