#!/usr/bin/env python3
"""Golden transitions from the reference's LEGACY python rule engine (`src/rule_engine.py`, `src/move_generator.py`) --
an implementation of the rules that is independent of the v0 C++/CUDA engine the other golden files come from.

    python tests/golden/make_legacy_golden.py       ->  tests/golden/legacy_playouts.npz

Same sampling convention as the reference's tests/v0/test_actions.py (python `random.Random(0x7777)`, uniformly random
legal moves from the initial position), but every ply of every game is recorded instead of one end state per game:
the state in the reference tensor layout (v0/python/state_batch.py:79-139 + moves_since_capture), the legal action
indices (v0/python/move_encoder.py:250-285), the played index, `is_game_over` and the winner.  Row i+1 of a game is the
successor of row i under `chosen[i]`, so consumers check mask, transition and terminal status without the reference.
"""
import random
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, "/root/reference")

from src.game_state import GameState  # noqa: E402
from src.move_generator import apply_move, generate_all_legal_moves  # noqa: E402

_DIRS = {(-1, 0): 0, (1, 0): 1, (0, -1): 2, (0, 1): 3}


def action_to_index(move: dict, n: int = 6) -> int:
    """Index layout of v0/python/move_encoder.py:250-285 restated here so that this generator depends on the legacy
    `src/` engine only (importing v0.python would pull in the v0_core extension)."""
    kind = move["action_type"]
    if kind == "place":
        r, c = move["position"]
        return r * n + c
    if kind == "move":
        (r0, c0), (r1, c1) = move["from_position"], move["to_position"]
        return 36 + (r0 * n + c0) * 4 + _DIRS[(r1 - r0, c1 - c0)]
    if kind in ("mark", "capture", "remove", "counter_remove", "no_moves_remove"):
        r, c = move["position"]
        return 180 + r * n + c
    if kind == "process_removal":
        return 216
    raise ValueError(kind)


GAMES, MAX_PLIES, SEED = 48, 400, 0x7777
OUT = Path(__file__).resolve().parent / "legacy_playouts.npz"


def tensor_row(s: GameState):
    board = np.asarray(s.board, dtype=np.int8)
    mb = np.zeros((6, 6), np.bool_)
    mw = np.zeros((6, 6), np.bool_)
    for r, c in s.marked_black:
        mb[r, c] = True
    for r, c in s.marked_white:
        mw[r, c] = True
    scal = [int(s.phase.value), int(s.current_player.value), int(s.pending_marks_required), int(s.pending_marks_remaining),
            int(s.pending_captures_required), int(s.pending_captures_remaining), int(s.forced_removals_done),
            int(s.move_count), int(s.moves_since_capture)]
    return board, mb, mw, scal


def main():
    rng = random.Random(SEED)
    boards, mbs, mws, scalars, legal_ptr, legal_idx, chosen, over, winner, game_ptr = [], [], [], [], [0], [], [], [], [], [0]
    for _g in range(GAMES):
        s = GameState()
        for _ply in range(MAX_PLIES + 1):
            b, mb, mw, sc = tensor_row(s)
            boards.append(b), mbs.append(mb), mws.append(mw), scalars.append(sc)
            go = bool(s.is_game_over())
            w = s.get_winner()
            over.append(go)
            winner.append(0 if w is None else int(w.value))
            moves = [] if go else generate_all_legal_moves(s)
            idx = sorted(int(action_to_index(m, 6)) for m in moves)
            assert len(set(idx)) == len(idx)
            legal_idx.extend(idx)
            legal_ptr.append(len(legal_idx))
            if go or not moves or _ply == MAX_PLIES:
                chosen.append(-1)
                break
            mv = rng.choice(moves)
            chosen.append(int(action_to_index(mv, 6)))
            s = apply_move(s, mv, quiet=True)
        game_ptr.append(len(boards))
    np.savez_compressed(
        OUT, board=np.stack(boards), marks_black=np.packbits(np.stack(mbs).reshape(len(mbs), 36), axis=1),
        marks_white=np.packbits(np.stack(mws).reshape(len(mws), 36), axis=1), scalars=np.asarray(scalars, np.int16),
        legal_ptr=np.asarray(legal_ptr, np.int32), legal_idx=np.asarray(legal_idx, np.uint8),
        chosen=np.asarray(chosen, np.int16), over=np.asarray(over, np.bool_), winner=np.asarray(winner, np.int8),
        game_ptr=np.asarray(game_ptr, np.int32))
    print(f"wrote {OUT.name}: {len(boards)} states, {GAMES} games, {OUT.stat().st_size / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
