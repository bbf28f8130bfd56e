"""ctypes binding of libliuzhou_b200.so (the C ABI declared in include/liuzhou_b200.h).

There is no CPU execution path: if the shared library is missing, was not built for sm_100a, or no CUDA
device is visible, every op raises instead of falling back.
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

import torch

_CSRC = Path(__file__).resolve().parent / "csrc"
LIB_PATH = Path(os.environ["LZB_LIB_PATH"]) if os.environ.get("LZB_LIB_PATH") else _CSRC / "libliuzhou_b200.so"

STATE_FIELDS = (
    "board", "marks_black", "marks_white", "phase", "current_player",
    "pending_marks_required", "pending_marks_remaining",
    "pending_captures_required", "pending_captures_remaining",
    "forced_removals_done", "move_count", "moves_since_capture",
)


class StatesView(ctypes.Structure):
    """lzb_states_in / lzb_states_out (identical layout: 12 pointers)."""

    _fields_ = [(name, ctypes.c_void_p) for name in STATE_FIELDS]


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m liuzhou_b200.build` "
                "(there is no CPU fallback for the liuzhou_b200 ops)")
        L = ctypes.CDLL(str(LIB_PATH))
        L.lzb_last_error.restype = ctypes.c_char_p
        L.lzb_version.restype = ctypes.c_char_p
        L.lzb_launch_count.restype = ctypes.c_uint64
        _lib = L
    return _lib


def launch_count() -> int:
    return int(lib().lzb_launch_count())


def check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError(lib().lzb_last_error().decode("utf-8", "replace"))


def require_cuda(t: torch.Tensor, name: str = "tensor") -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"liuzhou_b200: {name} must be a CUDA tensor (got device {t.device}); this package has no CPU path")


def ptr(t: torch.Tensor | None) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def stream_ptr(device=None) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def i64(v) -> ctypes.c_int64:
    return ctypes.c_int64(int(v))


def states_view(tensors) -> StatesView:
    """tensors: sequence of 12 (or 10; trailing move_count / moves_since_capture may be None) CUDA tensors."""
    v = StatesView()
    for name, t in zip(STATE_FIELDS, list(tensors) + [None] * (12 - len(tensors))):
        setattr(v, name, 0 if t is None else t.data_ptr())
    return v
