"""The object-level half of the ``v0_core`` surface (v0/src/bindings/module.cpp:877-1156): ``Phase`` / ``Player`` /
``ActionType``, ``GameState``, ``MoveRecord``, ``ActionCode``, the scalar rule functions (``generate_all_legal_moves_struct``,
``generate_moves_with_codes``, ``apply_move_struct``, the per-phase generators / appliers, ``encode_action_code(s)``) and the
``tensor_batch_from_game_states`` / ``tensor_batch_to_game_states`` converters.

The reference implements these on the host in C++ (v0/src/rules/rule_engine.cpp, v0/src/moves/move_generator.cpp); its
callers are tests, the evaluation scripts and the human-play UI -- one state at a time, never the self-play hot loop.
Here they are views over the SAME device kernels the batched ops use: a state is packed into the 32-byte bitboard form,
``lzb_legal_masks_packed`` / ``lzb_apply_actions_packed`` run on a batch of one on the current CUDA device, and the
result is unpacked into python objects.  So there is still exactly one rule engine in the product (and no CPU path):
these calls need a GPU and cost a kernel launch each -- use the batched ops for anything that matters.
"""
from __future__ import annotations

import enum
from typing import List, Optional, Sequence, Tuple

import torch

from . import native

Coord = Tuple[int, int]
BOARD_SIZE = 6
_DIRS = ((-1, 0), (1, 0), (0, -1), (0, 1))                      # up, down, left, right (fast_legal_mask_common.hpp:24-29)


class Phase(enum.IntEnum):                                      # v0/include/v0/game_state.hpp; module.cpp:877-886
    PLACEMENT = 1
    MARK_SELECTION = 2
    REMOVAL = 3
    MOVEMENT = 4
    CAPTURE_SELECTION = 5
    FORCED_REMOVAL = 6
    COUNTER_REMOVAL = 7


class Player(enum.IntEnum):                                     # module.cpp:888-891
    BLACK = 1
    WHITE = -1


class ActionType(enum.IntEnum):                                 # module.cpp:893-902; values = action-code kinds 1-8
    PLACE = 1
    MOVE = 2
    MARK = 3
    CAPTURE = 4
    FORCED_REMOVAL = 5
    COUNTER_REMOVAL = 6
    NO_MOVES_REMOVAL = 7
    PROCESS_REMOVAL = 8


_ACTION_NAMES = {ActionType.PLACE: "place", ActionType.MOVE: "move", ActionType.MARK: "mark", ActionType.CAPTURE: "capture",
                 ActionType.FORCED_REMOVAL: "remove", ActionType.COUNTER_REMOVAL: "counter_remove",
                 ActionType.NO_MOVES_REMOVAL: "no_moves_remove", ActionType.PROCESS_REMOVAL: "process_removal"}
_SELECT_KIND = {Phase.MARK_SELECTION: ActionType.MARK, Phase.CAPTURE_SELECTION: ActionType.CAPTURE,
                Phase.FORCED_REMOVAL: ActionType.FORCED_REMOVAL, Phase.COUNTER_REMOVAL: ActionType.COUNTER_REMOVAL,
                Phase.MOVEMENT: ActionType.NO_MOVES_REMOVAL}


class MoveRecord:
    """v0::MoveRecord (move_generator.hpp:24-45) with the property surface of module.cpp:904-957."""

    __slots__ = ("phase", "action_type", "primary", "secondary")

    def __init__(self, phase: Phase, action_type: ActionType, primary: Coord = (-1, -1), secondary: Coord = (-1, -1)):
        self.phase, self.action_type = Phase(phase), ActionType(action_type)
        self.primary, self.secondary = (int(primary[0]), int(primary[1])), (int(secondary[0]), int(secondary[1]))

    placement = staticmethod(lambda position: MoveRecord(Phase.PLACEMENT, ActionType.PLACE, position))
    mark = staticmethod(lambda position: MoveRecord(Phase.MARK_SELECTION, ActionType.MARK, position))
    capture = staticmethod(lambda position: MoveRecord(Phase.CAPTURE_SELECTION, ActionType.CAPTURE, position))
    forced_removal = staticmethod(lambda position: MoveRecord(Phase.FORCED_REMOVAL, ActionType.FORCED_REMOVAL, position))
    counter_removal = staticmethod(lambda position: MoveRecord(Phase.COUNTER_REMOVAL, ActionType.COUNTER_REMOVAL, position))
    no_moves_removal = staticmethod(lambda position: MoveRecord(Phase.MOVEMENT, ActionType.NO_MOVES_REMOVAL, position))
    process_removal = staticmethod(lambda: MoveRecord(Phase.REMOVAL, ActionType.PROCESS_REMOVAL))
    movement = staticmethod(lambda from_position, to_position: MoveRecord(Phase.MOVEMENT, ActionType.MOVE, from_position,
                                                                          to_position))

    @property
    def action_type_name(self) -> str:
        return _ACTION_NAMES[self.action_type]

    @property
    def position(self) -> Optional[Coord]:
        return None if self.action_type in (ActionType.MOVE, ActionType.PROCESS_REMOVAL) else self.primary

    @property
    def from_position(self) -> Optional[Coord]:
        return self.primary if self.action_type == ActionType.MOVE else None

    @property
    def to_position(self) -> Optional[Coord]:
        return self.secondary if self.action_type == ActionType.MOVE else None

    def to_dict(self) -> dict:
        d = {"phase": self.phase, "action_type": self.action_type_name}
        if self.action_type == ActionType.MOVE:
            d["from_position"], d["to_position"] = self.primary, self.secondary
        elif self.position is not None:
            d["position"] = self.primary
        return d

    def action_index(self) -> int:
        """Index in the 220-d action space (v0/python/move_encoder.py:250-285)."""
        if self.action_type == ActionType.PLACE:
            return self.primary[0] * BOARD_SIZE + self.primary[1]
        if self.action_type == ActionType.MOVE:
            d = (self.secondary[0] - self.primary[0], self.secondary[1] - self.primary[1])
            if d not in _DIRS:
                raise RuntimeError("只能水平或垂直移动一格")          # rule_engine.cpp:451
            return 36 + (self.primary[0] * BOARD_SIZE + self.primary[1]) * 4 + _DIRS.index(d)
        if self.action_type == ActionType.PROCESS_REMOVAL:
            return 216
        return 180 + self.primary[0] * BOARD_SIZE + self.primary[1]

    def __eq__(self, other) -> bool:
        return isinstance(other, MoveRecord) and (self.phase, self.action_type, self.primary, self.secondary) == (
            other.phase, other.action_type, other.primary, other.secondary)

    def __repr__(self) -> str:
        return f"MoveRecord({self.to_dict()})"


class ActionCode:
    """v0::ActionCode (move_generator.hpp:47-52)."""

    __slots__ = ("kind", "primary", "secondary", "extra")

    def __init__(self, kind: int = 0, primary: int = 0, secondary: int = 0, extra: int = 0):
        self.kind, self.primary, self.secondary, self.extra = int(kind), int(primary), int(secondary), int(extra)

    def to_tuple(self) -> Tuple[int, int, int, int]:
        return (self.kind, self.primary, self.secondary, self.extra)


class GameState:
    """v0::GameState (game_state.hpp:93-147) as bound at module.cpp:973-1006."""

    BOARD_SIZE = BOARD_SIZE
    MAX_MOVE_COUNT = 144
    NO_CAPTURE_DRAW_LIMIT = 36

    def __init__(self):
        self.board: List[List[int]] = [[0] * BOARD_SIZE for _ in range(BOARD_SIZE)]
        self.marked_black: List[Coord] = []
        self.marked_white: List[Coord] = []
        self.phase = Phase.PLACEMENT
        self.current_player = Player.BLACK
        self.forced_removals_done = 0
        self.move_count = 0
        self.moves_since_capture = 0
        self.pending_marks_required = self.pending_marks_remaining = 0
        self.pending_captures_required = self.pending_captures_remaining = 0

    def copy(self) -> "GameState":
        s = GameState()
        s.board = [list(r) for r in self.board]
        s.marked_black, s.marked_white = list(self.marked_black), list(self.marked_white)
        for k in _SCALARS:
            setattr(s, k, getattr(self, k))
        return s

    def switch_player(self) -> None:
        self.current_player = Player(-int(self.current_player))

    def is_board_full(self) -> bool:
        return all(v != 0 for row in self.board for v in row)

    def count_player_pieces(self, player) -> int:
        return sum(1 for row in self.board for v in row if v == int(player))

    def get_player_pieces(self, player) -> List[Coord]:
        return [(r, c) for r in range(BOARD_SIZE) for c in range(BOARD_SIZE) if self.board[r][c] == int(player)]

    def get_winner(self) -> Optional[Player]:                       # game_state.cpp:59-72
        if self.phase in (Phase.MOVEMENT, Phase.CAPTURE_SELECTION, Phase.COUNTER_REMOVAL):
            if self.count_player_pieces(Player.BLACK) < 4:
                return Player.WHITE
            if self.count_player_pieces(Player.WHITE) < 4:
                return Player.BLACK
        return None

    def is_game_over(self) -> bool:                                 # game_state.cpp:74-79
        return (self.get_winner() is not None or self.move_count >= self.MAX_MOVE_COUNT
                or self.moves_since_capture >= self.NO_CAPTURE_DRAW_LIMIT)

    def __eq__(self, other) -> bool:
        return isinstance(other, GameState) and self.board == other.board and \
            sorted(self.marked_black) == sorted(other.marked_black) and \
            sorted(self.marked_white) == sorted(other.marked_white) and \
            all(int(getattr(self, k)) == int(getattr(other, k)) for k in _SCALARS)


_SCALARS = ("phase", "current_player", "pending_marks_required", "pending_marks_remaining", "pending_captures_required",
            "pending_captures_remaining", "forced_removals_done", "move_count", "moves_since_capture")


class TensorStateBatch:
    """v0::TensorStateBatch as bound at module.cpp:1104-1141 (the reference's "SoA of bytes" layout)."""

    FIELDS = native.STATE_FIELDS

    def __init__(self, tensors: Optional[Sequence[torch.Tensor]] = None):
        t = list(tensors) if tensors is not None else [None] * 12
        for name, x in zip(self.FIELDS, t):
            setattr(self, name, x)
        n = 0 if t[0] is None else int(t[0].shape[0])
        self.mask_alive = None if t[0] is None else torch.ones((n,), dtype=torch.bool, device=t[0].device)
        self.board_size = BOARD_SIZE

    def tensors(self) -> List[torch.Tensor]:
        return [getattr(self, k) for k in self.FIELDS]

    def device(self):
        return self.board.device

    def to(self, device) -> "TensorStateBatch":
        return TensorStateBatch([x.to(device) for x in self.tensors()])

    def clone(self) -> "TensorStateBatch":
        return TensorStateBatch([x.clone() for x in self.tensors()])


# ----------------------------------------------------------------------------------------------------------------------
# converters
# ----------------------------------------------------------------------------------------------------------------------
def tensor_batch_from_game_states(states: Sequence[GameState], device: str = "cpu") -> TensorStateBatch:
    """tensor_state_batch.cpp FromGameStates (module.cpp:1143-1150)."""
    n = len(states)
    if n == 0:
        raise RuntimeError("tensor_batch_from_game_states requires at least one state")
    board = torch.tensor([s.board for s in states], dtype=torch.int8)
    mb = torch.zeros((n, 6, 6), dtype=torch.bool)
    mw = torch.zeros((n, 6, 6), dtype=torch.bool)
    for i, s in enumerate(states):
        for r, c in s.marked_black:
            mb[i, r, c] = True
        for r, c in s.marked_white:
            mw[i, r, c] = True
    scal = [torch.tensor([int(getattr(s, k)) for s in states], dtype=torch.int64) for k in _SCALARS]
    return TensorStateBatch([x.to(device) for x in [board, mb, mw] + scal])


def tensor_batch_to_game_states(batch) -> List[GameState]:
    """tensor_state_batch.cpp ToGameStates (module.cpp:1151-1153); accepts a TensorStateBatch or 12 tensors."""
    t = [x.detach().cpu() for x in (batch.tensors() if hasattr(batch, "tensors") else list(batch))]
    out = []
    for i in range(int(t[0].shape[0])):
        s = GameState()
        s.board = [[int(v) for v in row] for row in t[0][i].tolist()]
        s.marked_black = [(int(r), int(c)) for r, c in t[1][i].nonzero().tolist()]
        s.marked_white = [(int(r), int(c)) for r, c in t[2][i].nonzero().tolist()]
        for k, x in zip(_SCALARS, t[3:]):
            setattr(s, k, int(x[i]))
        s.phase, s.current_player = Phase(s.phase), Player(s.current_player)
        out.append(s)
    return out


def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("liuzhou_b200 scalar rule functions run on a CUDA device (no CPU rule engine in this package)")
    return torch.device("cuda", torch.cuda.current_device())


def _packed(state: GameState) -> torch.Tensor:
    return native.pack_states(tensor_batch_from_game_states([state], str(_device())).tensors())


def _legal_indices(state: GameState, ignore_game_over: bool = False) -> List[int]:
    """Legal action indices with the scalar engine's rules (move_generator.cpp:143-297).  ``GenerateAllLegalMoves``
    returns nothing once the game is over; the per-phase generators / appliers of rule_engine.cpp do not look at that,
    so for them (``ignore_game_over``) the draw counters are masked out (and a position decided by piece count falls
    back to the tensor-op mask, which never checks game-over)."""
    scalar = True
    if ignore_game_over:
        state = state.copy()
        state.move_count = state.moves_since_capture = 0
        scalar = state.get_winner() is None
    words, _ = native.legal_masks(_packed(state), scalar_semantics=scalar)
    return native.mask_words_to_bool(words)[0].nonzero().view(-1).tolist()


def _record(state: GameState, a: int) -> MoveRecord:
    if a < 36:
        return MoveRecord.placement(divmod(a, 6))
    if a < 180:
        cell, d = divmod(a - 36, 4)
        r, c = divmod(cell, 6)
        return MoveRecord.movement((r, c), (r + _DIRS[d][0], c + _DIRS[d][1]))
    if a < 216:
        return MoveRecord(state.phase, _SELECT_KIND[Phase(state.phase)], divmod(a - 180, 6))
    return MoveRecord.process_removal()


# ----------------------------------------------------------------------------------------------------------------------
# move generation / application (move_generator.cpp:242-432)
# ----------------------------------------------------------------------------------------------------------------------
def generate_all_legal_moves_struct(state: GameState) -> List[MoveRecord]:
    """Ascending action-index order == the reference's scan order; empty when the game is over (:243-245)."""
    return [_record(state, a) for a in _legal_indices(state)]


def encode_action_code(move: MoveRecord) -> ActionCode:
    """Scalar encoding (move_generator.cpp:299-347): moves carry ``secondary = to_cell`` (the tensor-op metadata uses
    ``(from, dir, to)`` instead), selections ``secondary = 0``, process-removal all zero."""
    if move.action_type == ActionType.MOVE:
        return ActionCode(2, move.primary[0] * 6 + move.primary[1], move.secondary[0] * 6 + move.secondary[1], 0)
    if move.action_type == ActionType.PROCESS_REMOVAL:
        return ActionCode(8, 0, 0, 0)
    return ActionCode(int(move.action_type), move.primary[0] * 6 + move.primary[1], 0, 0)


def encode_action_codes(moves: Sequence[MoveRecord]) -> List[ActionCode]:
    return [encode_action_code(m) for m in moves]


def generate_moves_with_codes(state: GameState):
    moves = generate_all_legal_moves_struct(state)
    return moves, encode_action_codes(moves)


_PHASE_KIND = {Phase.PLACEMENT: (ActionType.PLACE,), Phase.MARK_SELECTION: (ActionType.MARK,),
               Phase.REMOVAL: (ActionType.PROCESS_REMOVAL,), Phase.FORCED_REMOVAL: (ActionType.FORCED_REMOVAL,),
               Phase.MOVEMENT: (ActionType.MOVE, ActionType.NO_MOVES_REMOVAL), Phase.CAPTURE_SELECTION: (ActionType.CAPTURE,),
               Phase.COUNTER_REMOVAL: (ActionType.COUNTER_REMOVAL,)}


def _apply(state: GameState, move: MoveRecord, count_move: bool) -> GameState:
    if Phase(move.phase) != Phase(state.phase):
        raise RuntimeError("Move phase does not match state phase.")                   # move_generator.cpp:362
    if move.action_type not in _PHASE_KIND[Phase(state.phase)]:
        raise RuntimeError(f"{Phase(state.phase).name} phase does not allow '{move.action_type_name}'.")   # :368-408
    cells = (move.primary, move.secondary) if move.action_type == ActionType.MOVE else (
        () if move.action_type == ActionType.PROCESS_REMOVAL else (move.primary,))
    if any(not (0 <= r < BOARD_SIZE and 0 <= c < BOARD_SIZE) for r, c in cells):
        raise RuntimeError(f"position outside the board: {move.to_dict()}")             # rule_engine.cpp bounds checks
    a = move.action_index()
    if a not in _legal_indices(state, ignore_game_over=True):
        raise RuntimeError(f"illegal move for the current state: {move.to_dict()}")    # rule_engine.cpp:229-666
    dev = _device()
    nxt = native.apply_actions(_packed(state), torch.tensor([a], dtype=torch.int32, device=dev))
    out = tensor_batch_to_game_states(native.unpack_states(nxt))[0]
    if not count_move:               # the per-phase appliers leave the two counters alone; only ApplyMove advances them
        out.move_count, out.moves_since_capture = state.move_count, state.moves_since_capture
    return out


def apply_move_struct(state: GameState, move: MoveRecord, quiet: bool = False) -> GameState:
    """v0::ApplyMove (:360-432): one atomic action + ``move_count`` / ``moves_since_capture`` bookkeeping."""
    return _apply(state, move, True)


def _positions(state: GameState, phase: Phase, kind: ActionType) -> List[Coord]:
    if Phase(state.phase) != phase:
        return []
    return [m.primary for m in (_record(state, a) for a in _legal_indices(state, ignore_game_over=True))
            if m.action_type == kind]


def generate_placement_positions(state):
    return _positions(state, Phase.PLACEMENT, ActionType.PLACE)


def generate_mark_targets(state):
    return _positions(state, Phase.MARK_SELECTION, ActionType.MARK)


def generate_capture_targets(state):
    return _positions(state, Phase.CAPTURE_SELECTION, ActionType.CAPTURE)


def generate_movement_moves(state):
    if Phase(state.phase) != Phase.MOVEMENT:
        return []
    return [(m.primary, m.secondary) for m in (_record(state, a) for a in _legal_indices(state, ignore_game_over=True))
            if m.action_type == ActionType.MOVE]


def has_legal_movement_moves(state) -> bool:
    if Phase(state.phase) != Phase.MOVEMENT:
        raise RuntimeError("当前不是走子阶段")                                           # rule_engine.cpp:422-424
    return len(generate_movement_moves(state)) > 0


def _struct_moves(state, phase, kind):
    if Phase(state.phase) != phase:
        return []
    return [m for m in (_record(state, a) for a in _legal_indices(state, ignore_game_over=True)) if m.action_type == kind]


def generate_forced_removal_moves_struct(state):
    return _struct_moves(state, Phase.FORCED_REMOVAL, ActionType.FORCED_REMOVAL)


def generate_no_moves_options_struct(state):
    return _struct_moves(state, Phase.MOVEMENT, ActionType.NO_MOVES_REMOVAL)


def generate_counter_removal_moves_struct(state):
    return _struct_moves(state, Phase.COUNTER_REMOVAL, ActionType.COUNTER_REMOVAL)


def apply_placement_move(state, position):
    return _apply(state, MoveRecord.placement(position), False)


def apply_mark_selection(state, position):
    return _apply(state, MoveRecord.mark(position), False)


def process_phase2_removals(state):
    return _apply(state, MoveRecord.process_removal(), False)


def apply_movement_move(state, move, quiet: bool = False):
    return _apply(state, MoveRecord.movement(move[0], move[1]), False)


def apply_capture_selection(state, position, quiet: bool = False):
    return _apply(state, MoveRecord.capture(position), False)


def apply_forced_removal(state, piece_to_remove):
    return _apply(state, MoveRecord.forced_removal(piece_to_remove), False)


def handle_no_moves_phase3(state, stucked_player_removes, quiet: bool = False):
    return _apply(state, MoveRecord.no_moves_removal(stucked_player_removes), False)


def apply_counter_removal_phase3(state, opponent_removes, quiet: bool = False):
    return _apply(state, MoveRecord.counter_removal(opponent_removes), False)


# ----------------------------------------------------------------------------------------------------------------------
# composite phase-1 / phase-3 calls (rule_engine.cpp:682-728, bound at module.cpp:1050-1084): the API the reference's
# hand-built rule cases drive (tests/check_rule_engine_cases.py)
# ----------------------------------------------------------------------------------------------------------------------
def generate_legal_moves_phase1(state):
    return generate_placement_positions(state)


def apply_move_phase1(state, move, mark_positions=None):
    """Placement followed by the mark selections it earned; marks passed for a placement that formed no shape raise."""
    new_state = apply_placement_move(state, move)
    if mark_positions:
        if Phase(new_state.phase) != Phase.MARK_SELECTION:
            raise RuntimeError("当前状态不需要标记，但传入了 mark_positions")          # rule_engine.cpp:693
        for pos in mark_positions:
            new_state = apply_mark_selection(new_state, pos)
    return new_state


def generate_legal_moves_phase3(state):
    return generate_movement_moves(state)


def has_legal_moves_phase3(state) -> bool:
    return has_legal_movement_moves(state)


def apply_move_phase3(state, move, capture_positions=None, quiet: bool = False):
    """Movement followed by the captures it earned (without them the state stays in CAPTURE_SELECTION, as in the v0
    C++ engine -- the legacy python engine raises there instead, src/rule_engine.py:621)."""
    new_state = apply_movement_move(state, move, quiet)
    if capture_positions:
        if Phase(new_state.phase) != Phase.CAPTURE_SELECTION:
            raise RuntimeError("当前状态不需要提子，但传入了 capture_positions")        # rule_engine.cpp:719
        for pos in capture_positions:
            new_state = apply_capture_selection(new_state, pos, quiet)
    return new_state


__all__ = [n for n in dir() if not n.startswith("_") and n not in ("annotations", "enum", "torch", "native", "List",
                                                                   "Optional", "Sequence", "Tuple")]
