"""Test-only: the tree-backend self-play wave written the slow, literal way -- the reference's wave loop
(v1/python/self_play_gpu_runner.py:159-256) over the reference-layout drop-in ops, one host round trip per ply, every
tree searched at full batch size.  `liuzhou_b200.self_play._play_wave_tree` (packed layout, no per-ply sync, compacted
leaf batches) must produce exactly the same trajectory batch."""
import torch

from liuzhou_b200 import native, v0_core
from liuzhou_b200.mcts_gpu import TOTAL_ACTION_DIM, GpuStateBatch
from liuzhou_b200.trajectory_buffer import TensorTrajectoryBuffer
from liuzhou_b200.tree import encode_inputs
from liuzhou_b200.tree_search import TreeMCTS, TreeMCTSConfig


def slow_tree_self_play(net, num_games, sims, *, temperature_init=1.0, temperature_final=0.1, temperature_threshold=10,
                        add_dirichlet_noise=False, sample_moves=False, max_plies=512, soft_value_k=2.0, tree_reuse=True,
                        device="cuda:0"):
    dev = torch.device(device)
    g = int(num_games)
    mcts = TreeMCTS(net, g, TreeMCTSConfig(num_simulations=sims, add_dirichlet_noise=add_dirichlet_noise,
                                           sample_moves=sample_moves, reuse_subtree=tree_reuse), dev)
    buffer = TensorTrajectoryBuffer(dev, TOTAL_ACTION_DIM, max_steps_hint=160, concurrent_games_hint=g)
    states = GpuStateBatch.initial(dev, batch_size=g)
    step_index_matrix = torch.full((g, max_plies), -1, dtype=torch.int64, device=dev)
    step_counts = torch.zeros((g,), dtype=torch.int64, device=dev)
    plies = torch.zeros((g,), dtype=torch.int64, device=dev)
    done = torch.zeros((g,), dtype=torch.bool, device=dev)
    lengths = torch.zeros((g,), dtype=torch.int64, device=dev)
    outcomes = torch.zeros((3,), dtype=torch.int64, device=dev)
    while True:
        active_idx = torch.where(~done)[0]
        n_active = int(active_idx.numel())
        if n_active == 0:
            break
        packed = native.pack_states(states.tensors())
        temps = torch.where(plies < temperature_threshold, float(temperature_init), float(temperature_final)).float()
        out = mcts.search(packed, active=~done, temperatures=temps, add_dirichlet_noise=add_dirichlet_noise,
                          sample_moves=sample_moves)
        chosen_all = out.chosen_action_indices
        mask_all, meta_all = v0_core.encode_actions_fast(*states.tensors()[:10], 36, 144, 36, 4)
        chosen_a = chosen_all.index_select(0, active_idx)
        chosen_valid = chosen_a >= 0
        meta_a = meta_all.index_select(0, active_idx)
        codes = meta_a.gather(1, chosen_a.clamp_min(0).view(-1, 1, 1).expand(-1, 1, 4)).view(-1, 4)
        codes = torch.where(chosen_valid.view(-1, 1), codes, torch.full_like(codes, -1))
        terminal = out.terminal_mask.index_select(0, active_idx)
        mcts.advance(torch.where(done | out.terminal_mask, torch.full_like(chosen_all, -1), chosen_all))
        idx = buffer.append_steps(model_input=encode_inputs(packed.index_select(0, active_idx), "f32_nchw"),
                                  legal_mask=mask_all.index_select(0, active_idx),
                                  policy_dense=out.policy_dense.index_select(0, active_idx),
                                  player_sign=states.current_player.index_select(0, active_idx))
        step_index_matrix[active_idx, step_counts.index_select(0, active_idx)] = idx
        step_counts.index_add_(0, active_idx, torch.ones((n_active,), dtype=torch.int64, device=dev))
        slots, res, _soft = v0_core.self_play_step_inplace(*states.tensors(), plies, done, active_idx, codes, terminal,
                                                          chosen_valid, int(max_plies), float(soft_value_k))
        if int(slots.numel()) > 0:
            boards = states.board.index_select(0, slots)
            black = boards.eq(1).sum(dim=(1, 2)).float()
            white = boards.eq(-1).sum(dim=(1, 2)).float()
            soft = torch.tanh((black - white) / 18.0 * float(soft_value_k))
            fs, fl, od = buffer.finalize_games_inplace(step_index_matrix=step_index_matrix, step_counts=step_counts,
                                                      slots=slots, result_from_black=res, soft_value_from_black=soft)
            if int(fs.numel()) > 0:
                lengths.index_copy_(0, fs, fl)
            outcomes.add_(od)
    return buffer.build(), lengths, outcomes
