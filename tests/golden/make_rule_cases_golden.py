#!/usr/bin/env python3
"""Golden cases from the reference's HAND-BUILT rule regression suite.

    python tests/golden/make_rule_cases_golden.py      ->  tests/golden/rule_cases.json

The reference keeps its hand-written rule scenarios in /root/reference/tests/check_rule_engine_cases.py:75-1031 (every
`main_*`, `sample_case_*` and `test_phase*` function; positions built cell by cell, illegal inputs expected to raise)
and /root/reference/tests/test_game_state_phase_gate.py.  They drive the legacy python engine (`src/rule_engine.py`)
through its composite API (`apply_move_phase1(state, pos, mark_positions)`, `apply_move_phase3(state, move,
capture_positions)`, `process_phase2_removals`, `apply_forced_removal`, `handle_no_moves_phase3`,
`apply_counter_removal_phase3`, `has_legal_moves_phase3`, `generate_legal_moves_phase1`).

This script imports that file UNMODIFIED, wraps the engine functions it imported with a recorder, runs every scenario
function and dumps one record per call: (function, input state, arguments) -> (output state | returned value | raised).
Each record is also replayed on the reference's C++ scalar engine (`oracle/_ref/v0_core`, same function names,
module.cpp:1009-1069) and the agreement is stored (`cpp`: "same" or what differs), because the CUDA engine has to follow
the v0 C++ engine where the two reference implementations disagree.  Consumers need neither /root/reference nor oracle/_ref.
"""
import contextlib
import io
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
REF = Path("/root/reference")
sys.path.insert(0, str(REF))
sys.path.insert(0, str(ROOT / "oracle" / "_ref"))

OUT = Path(__file__).resolve().parent / "rule_cases.json"
FUNCS = ("generate_legal_moves_phase1", "apply_move_phase1", "process_phase2_removals", "apply_forced_removal",
         "has_legal_moves_phase3", "handle_no_moves_phase3", "apply_counter_removal_phase3", "apply_move_phase3")


def dump_state(s) -> dict:
    return {"board": [int(v) for row in s.board for v in row],
            "marked_black": sorted([int(r), int(c)] for r, c in s.marked_black),
            "marked_white": sorted([int(r), int(c)] for r, c in s.marked_white),
            "phase": int(getattr(s.phase, "value", s.phase)),
            "current_player": int(getattr(s.current_player, "value", s.current_player)),
            "pending_marks_required": int(s.pending_marks_required),
            "pending_marks_remaining": int(s.pending_marks_remaining),
            "pending_captures_required": int(s.pending_captures_required),
            "pending_captures_remaining": int(s.pending_captures_remaining),
            "forced_removals_done": int(s.forced_removals_done), "move_count": int(s.move_count),
            "moves_since_capture": int(getattr(s, "moves_since_capture", 0))}


def jsonable(x):
    if isinstance(x, (list, tuple, set)):
        return [jsonable(v) for v in (sorted(x) if isinstance(x, set) else x)]
    if x is None or isinstance(x, (bool, int, str)):
        return x
    raise TypeError(type(x))


def to_cpp_state(v0_core, d: dict):
    s = v0_core.GameState()
    s.board = [d["board"][r * 6:(r + 1) * 6] for r in range(6)]
    s.marked_black = [tuple(p) for p in d["marked_black"]]
    s.marked_white = [tuple(p) for p in d["marked_white"]]
    s.phase = v0_core.Phase(d["phase"])
    s.current_player = v0_core.Player(d["current_player"])
    for k in ("pending_marks_required", "pending_marks_remaining", "pending_captures_required",
              "pending_captures_remaining", "forced_removals_done", "move_count", "moves_since_capture"):
        if hasattr(s, k):                 # the pybind GameState does not expose moves_since_capture (module.cpp:973-1006)
            setattr(s, k, d[k])
    return s


def outcome(fn, state, args, kwargs, dump):
    try:
        r = fn(state, *args, **kwargs)
    except Exception as e:  # noqa: BLE001  (the scenarios expect ValueError; C++ raises RuntimeError)
        return {"raises": True, "exception": type(e).__name__}
    if hasattr(r, "board"):
        return {"state": dump(r)}
    if isinstance(r, bool):
        return {"value": r}
    return {"value": sorted(jsonable(r))}


def main() -> int:
    import importlib.util

    def load(name):                                       # the reference's files, unmodified, loaded by path
        spec = importlib.util.spec_from_file_location(f"ref_{name}", REF / "tests" / f"{name}.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    cases, gate = load("check_rule_engine_cases"), load("test_game_state_phase_gate")
    from src.game_state import GameState
    from src.move_generator import apply_move

    records = []
    scenario = ["?"]

    def traced(name, fn):
        def wrapper(state, *args, **kwargs):
            rec = {"scenario": scenario[0], "fn": name, "state": dump_state(state), "args": jsonable(list(args)),
                   "kwargs": {k: jsonable(v) for k, v in kwargs.items()}}
            res = outcome(fn, state, args, kwargs, dump_state)
            rec["result"] = {k: v for k, v in res.items() if k != "exception"}
            records.append(rec)
            if res.get("raises"):
                raise ValueError("recorded")          # what the scenarios' assert_raises expects
            return fn(state, *args, **kwargs)
        return wrapper

    for name in FUNCS:
        setattr(cases, name, traced(name, getattr(cases, name)))
    scenario_fns = [n for n in dir(cases) if n.startswith(("main_", "sample_case_", "test_phase"))]
    sink = io.StringIO()
    failed = {}
    for n in sorted(scenario_fns):
        scenario[0] = n
        with contextlib.redirect_stdout(sink):
            try:
                getattr(cases, n)()
            except Exception as e:  # noqa: BLE001  a scenario that dies half way still contributed its calls
                failed[n] = f"{type(e).__name__}: {e}"
    # tests/test_game_state_phase_gate.py: a 7-move opening through src.move_generator.apply_move + two is_game_over checks
    scenario[0] = "phase_gate"
    st = GameState()
    for mv in [(2, 4), (5, 2), (1, 4), (0, 4), (2, 3), (1, 0), (1, 3)]:
        before = dump_state(st)
        st = apply_move(st, {"phase": st.phase, "action_type": "place", "position": mv}, quiet=True)
        records.append({"scenario": "phase_gate", "fn": "apply_move_struct", "state": before, "args": [list(mv)],
                        "kwargs": {}, "result": {"state": dump_state(st)}})
    records.append({"scenario": "phase_gate", "fn": "is_game_over", "state": dump_state(st), "args": [], "kwargs": {},
                    "result": {"value": bool(st.is_game_over())}})
    gate.test_mark_selection_does_not_adjudicate_before_movement_starts()
    gate.test_winner_check_remains_active_in_movement_stage()
    from src.game_state import Phase
    s2 = GameState()
    s2.phase = Phase.MOVEMENT
    for c in range(4):
        s2.board[0][c] = 1
    for c in range(3):
        s2.board[5][c] = -1
    records.append({"scenario": "phase_gate", "fn": "is_game_over", "state": dump_state(s2), "args": [], "kwargs": {},
                    "result": {"value": bool(s2.is_game_over())}})

    # replay on the reference's C++ scalar engine
    import torch  # noqa: F401  (v0_core links libtorch)
    import v0_core

    agree = 0
    for rec in records:
        if rec["fn"] in ("is_game_over", "apply_move_struct"):
            s = to_cpp_state(v0_core, rec["state"])
            if rec["fn"] == "is_game_over":
                # the pybind GameState has no is_game_over(); the portable module exposes v0::GameState::IsGameOver
                import _liuzhou_portable_cpp as portable
                from types import SimpleNamespace

                d = rec["state"]
                obj = SimpleNamespace(**{**d, "board": [d["board"][r * 6:(r + 1) * 6] for r in range(6)],
                                         "marked_black": [tuple(p) for p in d["marked_black"]],
                                         "marked_white": [tuple(p) for p in d["marked_white"]]})
                got = {"value": bool(portable.inspect_state(obj)["game_over"])}
            else:
                got = {"state": dump_state(v0_core.apply_move_struct(s, v0_core.MoveRecord.placement(tuple(rec["args"][0]))))}
        else:
            args = [tuple(tuple(x) if isinstance(x, list) else x for x in a) if isinstance(a, list) else a
                    for a in rec["args"]]
            kwargs = {k: [tuple(p) for p in v] if isinstance(v, list) else v for k, v in rec["kwargs"].items()}
            got = outcome(getattr(v0_core, rec["fn"]), to_cpp_state(v0_core, rec["state"]), args, kwargs, dump_state)
            got = {k: v for k, v in got.items() if k != "exception"}
        if got == rec["result"]:
            rec["cpp"] = "same"
            agree += 1
        else:
            rec["cpp"] = got
    OUT.write_text(json.dumps({"source": "reference tests/check_rule_engine_cases.py + tests/test_game_state_phase_gate.py, "
                                         "legacy src/ engine outputs; cpp = v0 C++ scalar engine on the same call",
                               "scenarios_that_raised": failed, "cases": records}, separators=(",", ":")))
    print(f"{len(records)} calls recorded over {len(scenario_fns) + 1} scenarios; v0 C++ agrees on {agree}; "
          f"scenarios that raised: {failed}")
    for rec in records:
        if rec["cpp"] != "same":
            print("  differs:", rec["scenario"], rec["fn"], rec["args"], rec["kwargs"], "legacy",
                  {k: (v if k != "state" else "...") for k, v in rec["result"].items()}, "cpp",
                  {k: (v if k != "state" else "...") for k, v in rec["cpp"].items()})
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
