"""Steady-state self-play stepper on the native layout: G concurrent games as packed bitboards in HBM, one
``step()`` = one ply of every game (full tree MCTS with S simulations per move on the device tree, bf16 network
on the tensor cores, move selection, atomic move application, terminal detection, trajectory rows in the
reference format written into a device ring).  Finished games are refilled with fresh initial positions so the
batch stays full; this is the continuous variant of the reference's wave loop
(v1/python/self_play_gpu_runner.py:159-256) used for throughput measurement and long-running actors.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from . import native, v0_core
from .net import InferenceNet
from .tree import encode_inputs
from .tree_search import TreeMCTS, TreeMCTSConfig

ACTION_DIM = 220
BYTES_PER_POSITION = 2692      # 1,584 planes + 220 mask + 880 policy + 4 + 4


@dataclass
class StepResult:
    positions: int               # trajectory rows produced (= live games this ply)
    finished: int                # games that ended this ply (refilled)


class SelfPlayStepper:
    def __init__(self, net: InferenceNet, num_games: int, *, simulations: int = 200, exploration_weight: float = 1.0,
                 leaves_per_wave: int = 1, add_dirichlet_noise: bool = True, dirichlet_alpha: float = 0.3,
                 dirichlet_epsilon: float = 0.25, temperature_init: float = 1.0, temperature_final: float = 0.1,
                 temperature_threshold: int = 10, max_game_plies: int = 512, sample_moves: bool = True,
                 seed: int = 0, ring_steps: int = 4, device=None, reuse_subtree: bool = False):
        self.net = net
        self.device = torch.device(device) if device is not None else net.device
        self.g = int(num_games)
        self.t_init, self.t_final, self.t_thr = float(temperature_init), float(temperature_final), int(temperature_threshold)
        self.max_plies = int(max_game_plies)
        self.mcts = TreeMCTS(net, self.g, TreeMCTSConfig(
            num_simulations=int(simulations), exploration_weight=float(exploration_weight),
            add_dirichlet_noise=bool(add_dirichlet_noise), dirichlet_alpha=float(dirichlet_alpha),
            dirichlet_epsilon=float(dirichlet_epsilon), sample_moves=bool(sample_moves),
            leaves_per_wave=int(leaves_per_wave), reuse_subtree=bool(reuse_subtree)), self.device)
        dev = self.device
        self.states = native.init_states(self.g, dev)
        self.plies = torch.zeros((self.g,), dtype=torch.int32, device=dev)
        self._init_state = native.init_states(1, dev)
        # trajectory ring in the reference format (rows of one ply are contiguous; ply-major like the reference)
        self.ring_steps = int(ring_steps)
        r = self.ring_steps * self.g
        self.traj_planes = torch.empty((r, 11, 6, 6), dtype=torch.float32, device=dev)
        self.traj_legal = torch.empty((r, ACTION_DIM), dtype=torch.bool, device=dev)
        self.traj_policy = torch.empty((r, ACTION_DIM), dtype=torch.float32, device=dev)
        self.traj_sign = torch.empty((r,), dtype=torch.int8, device=dev)
        self.traj_game = torch.empty((r,), dtype=torch.int32, device=dev)
        self._ring_pos = 0
        self.games_finished = 0
        self.positions = 0
        self.outcomes = torch.zeros((3,), dtype=torch.int64, device=dev)   # black wins, white wins, draws

    def diversify(self, seed: int = 20260314, max_random_plies: int = 120, groups: int = 16) -> None:
        """Synthetic steady state: advance game groups by 0 .. max_random_plies uniform-random plies so that the
        batch mixes placement / marking / removal / movement positions like a long-running actor does."""
        g = self.g
        per = (g + groups - 1) // groups
        chunks = []
        for i in range(groups):
            n = min(per, g - i * per)
            if n <= 0:
                break
            pb = native.PlayoutBatch(n, seed=seed, device=self.device, game_offset=i * per)
            steps = (max_random_plies * i) // max(1, groups - 1)
            if steps > 0:
                pb.run(max_steps=steps)
            alive = pb.result == 2
            st = torch.where(alive.view(-1, 1), pb.packed, self._init_state.expand(n, 4))
            pl = torch.where(alive, pb.plies, torch.zeros_like(pb.plies))
            chunks.append((st, pl))
        self.states = torch.cat([c[0] for c in chunks]).contiguous()
        self.plies = torch.cat([c[1] for c in chunks]).contiguous()

    @torch.no_grad()
    def step(self) -> None:
        """One ply of all G games; no host synchronisation."""
        dev, g = self.device, self.g
        temps = torch.where(self.plies < self.t_thr, self.t_init, self.t_final).to(torch.float32)
        out = self.mcts.search(self.states, temperatures=temps)
        # trajectory rows (reference format) for this ply
        s = self._ring_pos * g
        encode_inputs(self.states, "f32_nchw", out=self.traj_planes[s:s + g])
        self.traj_legal[s:s + g] = out.legal_mask
        self.traj_policy[s:s + g] = out.policy_dense
        words0 = self.states[:, 0]
        white = (words0 >> 39) & 1                               # meta bit 3 of w0 >> 36: white to move
        self.traj_sign[s:s + g] = (1 - 2 * white).to(torch.int8)
        self._ring_pos = (self._ring_pos + 1) % self.ring_steps
        # apply the chosen moves; roots without a legal action lose (module.cpp:733-735)
        chosen = out.chosen_action_indices
        stuck = chosen < 0
        nxt = native.apply_actions(self.states, chosen.clamp_min(0).to(torch.int32))
        self.plies += 1
        # terminal detection on the packed layout (game_state.cpp:59-79 + max plies)
        over, winner = packed_status(nxt)
        done = over | stuck | (self.plies >= self.max_plies)
        res = torch.where(stuck, -(1 - 2 * white), winner)      # result from black's perspective
        self.outcomes += torch.stack([(done & (res > 0)).sum(), (done & (res < 0)).sum(), (done & (res == 0)).sum()])
        self.states = torch.where(done.view(-1, 1), self._init_state.expand(g, 4), nxt).contiguous()
        self.plies = torch.where(done, torch.zeros_like(self.plies), self.plies)
        self._last_done = done
        # subtree reuse: the played child becomes the root; finished games restart from a fresh root
        self.mcts.advance(torch.where(done, torch.full_like(chosen, -1), chosen), self.states, done)

    def trajectory_block(self, ring_index: Optional[int] = None):
        """(planes, legal, policy, sign) rows of one ply in the reference format."""
        i = (self._ring_pos - 1) % self.ring_steps if ring_index is None else ring_index
        s = i * self.g
        return (self.traj_planes[s:s + self.g], self.traj_legal[s:s + self.g], self.traj_policy[s:s + self.g],
                self.traj_sign[s:s + self.g])


def packed_status(packed: torch.Tensor):
    """(game_over bool[n], winner int64[n] in {+1 black, -1 white, 0}) of packed states -- the host-tensor form
    of lz::game_over / lz::winner for the few places that need it outside a kernel."""
    w0, w1 = packed[:, 0], packed[:, 1]
    mask36 = (1 << 36) - 1
    meta = (w0 >> 36) & 0xFFFFFFF
    phase = meta & 7
    move_count = (meta >> 14) & 255
    msc = (meta >> 22) & 63
    black = _popcount36(w0 & mask36)
    white = _popcount36(w1 & mask36)
    post = (phase == 4) | (phase == 5) | (phase == 7)
    winner = torch.where(post & (black < 4), -1, torch.where(post & (white < 4), 1, 0))
    over = (winner != 0) | (move_count >= 144) | (msc >= 36)
    return over, winner


def _popcount36(x: torch.Tensor) -> torch.Tensor:
    x = x - ((x >> 1) & 0x5555555555555555)
    x = (x & 0x3333333333333333) + ((x >> 2) & 0x3333333333333333)
    x = (x + (x >> 4)) & 0x0F0F0F0F0F0F0F0F
    return (x * 0x0101010101010101 >> 56) & 0xFF
