"""liuzhou_b200 -- B200-native (sm_100a) batched MCTS self-play engine for Liuzhou Chess (六洲棋).

Layers (host side mirrors the reference's operator surface, kernels live in csrc/):
  v0_core     drop-in for the reference's `v0_core` tensor ops (legal mask, apply-move, root-PUCT, ...)
  native      packed-bitboard state layout, rule ops and the random-playout workload
"""
__version__ = "0.1.0"
