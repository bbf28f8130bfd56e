#!/usr/bin/env python3
"""Build the UNMODIFIED reference implementation of the hot path into oracle/_ref/.

TEST INFRASTRUCTURE ONLY. Nothing under liuzhou_b200/ may import what this builds.

The reference (kuailehaha/liuzhou, mounted read-only at /root/reference) is compiled
from its own source files *where they lie* -- no sources are copied into this repo,
and the reference's CMake build is not used.  Outputs (binaries only):

  oracle/_ref/_liuzhou_portable_cpp<EXT>   v1/cpp/portable_mcts.cpp + v0 scalar rule engine
                                            (pybind11 only, no torch)          -> tree-MCTS reference
  oracle/_ref/v0_core<EXT>                 v0/src/** (pybind11 + libtorch, CPU ops;
                                            with --cuda also the three reference .cu kernels
                                            compiled for sm_100)               -> tensor-op reference

  oracle/_ref/pysrc/{src,v0/python,v1}     (--pysrc, default) a plain copy of the reference's PYTHON host code for
                                            this path (v1/python/mcts_gpu.py, self_play_gpu_runner.py, src/ ...), so that
                                            the GPU box -- which has no /root/reference -- can run the UNMODIFIED
                                            reference entry over its own v0_core CUDA (the "reference on the same
                                            B200" denominator) and over our v0_core shim (the boundary proof).
                                            Never imported by liuzhou_b200/, never committed.

oracle/_ref/ is git-ignored but NOT gpurun-ignored, so the binaries travel to the GPU box.
The GPU box has no /root/reference: this script is a no-op there (keeps prebuilt files).

Recipe mirrors what the reference's own build does (file lists from
/root/reference/v0/src/CMakeLists.txt:100-117 and
/root/reference/scripts/build_portable_cpp.py:32-48), as plain g++/nvcc command lines.
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys
import sysconfig
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE / "_ref"
REF = Path(os.environ.get("LIUZHOU_REFERENCE", "/root/reference"))
EXT = sysconfig.get_config_var("EXT_SUFFIX")

V0_CPP = [
    "v0/src/bindings/module.cpp",
    "v0/src/game/game_state.cpp",
    "v0/src/game/tensor_state_batch.cpp",
    "v0/src/rules/rule_engine.cpp",
    "v0/src/moves/move_generator.cpp",
    "v0/src/net/encoding.cpp",
    "v0/src/net/inference_engine.cpp",
    "v0/src/net/torchscript_runner.cpp",
    "v0/src/net/project_policy_logits_fast.cpp",
    "v0/src/game/fast_legal_mask.cpp",
    "v0/src/game/fast_apply_moves.cpp",
    "v0/src/mcts/mcts_core.cpp",
    "v0/src/mcts/eval_batcher.cpp",
]
V0_CU = [
    "v0/src/game/fast_legal_mask_cuda.cu",
    "v0/src/game/fast_apply_moves_cuda.cu",
    "v0/src/mcts/root_puct_fused.cu",
]
PORTABLE_CPP = [
    "v1/cpp/portable_mcts.cpp",
    "v0/src/game/game_state.cpp",
    "v0/src/rules/rule_engine.cpp",
    "v0/src/moves/move_generator.cpp",
]


def _run(cmd: list[str]) -> None:
    t0 = time.time()
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + res.stdout[-4000:] + "\n")
        raise RuntimeError(f"command failed ({res.returncode}): {cmd[0]} ... {cmd[-1]}")
    print(f"  [{time.time() - t0:5.1f}s] {cmd[0]} {Path(cmd[-1]).name}", flush=True)


def _py_includes() -> list[str]:
    import pybind11

    return [f"-I{sysconfig.get_paths()['include']}", f"-I{pybind11.get_include()}"]


def build_portable(force: bool) -> Path:
    target = OUT / f"_liuzhou_portable_cpp{EXT}"
    if target.exists() and not force:
        return target
    objs = []
    tmp = OUT / "obj_portable"
    tmp.mkdir(parents=True, exist_ok=True)
    cmds = []
    for src in PORTABLE_CPP:
        obj = tmp / (Path(src).stem + ".o")
        objs.append(str(obj))
        cmds.append(
            ["g++", "-O3", "-std=c++17", "-fPIC", "-fvisibility=hidden", f"-I{REF / 'v0/include'}",
             *_py_includes(), "-c", str(REF / src), "-o", str(obj)]
        )
    with ThreadPoolExecutor(max_workers=4) as pool:
        list(pool.map(_run, cmds))
    _run(["g++", "-shared", *objs, "-o", str(target)])
    return target


def build_ref_playout(force: bool) -> Path:
    """oracle/_ref/ref_playout: oracle/ref_playout_harness.cpp (ours: a random-playout driver) + the reference's scalar
    engine sources where they lie.  The CPU arm of BASELINE configs[1] with cpu_baseline.kind = "reference"."""
    target = OUT / "ref_playout"
    if target.exists() and not force:
        return target
    OUT.mkdir(parents=True, exist_ok=True)
    srcs = [str(HERE / "ref_playout_harness.cpp")] + [str(REF / s) for s in PORTABLE_CPP[1:]]
    _run(["g++", "-O3", "-std=c++17", "-pthread", f"-I{REF / 'v0/include'}", *srcs, "-o", str(target)])
    return target


def build_v0_core(force: bool, cuda: bool) -> Path:
    target = OUT / f"v0_core{EXT}"
    stamp = OUT / ("v0_core.cuda" if cuda else "v0_core.cpu")
    if target.exists() and stamp.exists() and not force:
        return target
    import torch
    from torch.utils import cpp_extension

    tmp = OUT / ("obj_v0_cuda" if cuda else "obj_v0_cpu")
    tmp.mkdir(parents=True, exist_ok=True)
    inc = [f"-I{REF / 'v0/include'}", f"-I{REF / 'v0/src/game'}"]
    inc += [f"-I{p}" for p in cpp_extension.include_paths()]
    inc += _py_includes()
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    defs = [
        "-DTORCH_API_INCLUDE_EXTENSION_H", "-DNOMINMAX", "-DPROJECT_POLICY_NO_MODULE",
        "-DFAST_LEGAL_MASK_NO_MODULE", "-DFAST_APPLY_MOVES_NO_MODULE",
        "-DTORCH_EXTENSION_NAME=v0_core", f"-D_GLIBCXX_USE_CXX11_ABI={abi}",
    ]
    if cuda:
        defs += ["-DV0_HAS_CUDA_LEGAL_MASK", "-DV0_HAS_CUDA_APPLY_MOVES", "-DV0_HAS_CUDA_ROOT_PUCT",
                 "-DTORCH_CUDA_AVAILABLE"]
        inc += ["-I/usr/local/cuda/include"]
    cmds, objs = [], []
    for src in V0_CPP:
        obj = tmp / (Path(src).stem + ".o")
        objs.append(str(obj))
        cmds.append(["g++", "-O2", "-std=c++17", "-fPIC", "-fvisibility=hidden", "-w", *defs, *inc,
                     "-c", str(REF / src), "-o", str(obj)])
    if cuda:
        for src in V0_CU:
            obj = tmp / (Path(src).stem + ".cu.o")
            objs.append(str(obj))
            cmds.append(["nvcc", "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-w",
                         "-gencode", "arch=compute_100,code=sm_100",
                         "--expt-relaxed-constexpr", "--expt-extended-lambda", *defs, *inc,
                         "-c", str(REF / src), "-o", str(obj)])
    with ThreadPoolExecutor(max_workers=6) as pool:
        list(pool.map(_run, cmds))
    libdir = cpp_extension.library_paths()[0]
    libs = ["-ltorch", "-ltorch_cpu", "-lc10", "-ltorch_python"]
    if cuda:
        libs += ["-ltorch_cuda", "-lc10_cuda", "-L/usr/local/cuda/lib64", "-lcudart"]
    _run(["g++", "-shared", *objs, f"-L{libdir}", f"-Wl,-rpath,{libdir}", *libs, "-lpthread",
          "-o", str(target)])
    for other in ("v0_core.cuda", "v0_core.cpu"):
        (OUT / other).unlink(missing_ok=True)
    stamp.write_text("built from /root/reference v0/src by oracle/build_ref.py\n")
    return target


def copy_pysrc() -> Path:
    """The reference's python packages for this path, byte for byte, into the git-ignored oracle/_ref/pysrc."""
    import shutil

    dst = OUT / "pysrc"
    if dst.exists():
        shutil.rmtree(dst)
    ignore = shutil.ignore_patterns("__pycache__", "*.pyc", "*.so", "*.md")
    shutil.copytree(REF / "src", dst / "src", ignore=ignore)
    shutil.copytree(REF / "v0" / "python", dst / "v0" / "python", ignore=ignore)
    shutil.copytree(REF / "v1", dst / "v1", ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.so", "*.md", "cpp"))
    return dst


def main() -> int:
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--cuda", action="store_true", help="also compile the reference .cu kernels (sm_100)")
    ap.add_argument("--no-pysrc", action="store_true", help="skip the copy of the reference's python packages")
    ap.add_argument("--only", choices=["portable", "v0_core", "playout"], default=None)
    args = ap.parse_args()
    if not REF.is_dir():
        print(f"[build_ref] {REF} not present (GPU box?) -- keeping prebuilt oracle/_ref as is")
        return 0
    OUT.mkdir(parents=True, exist_ok=True)
    if args.only in (None, "portable"):
        print("[build_ref] portable tree MCTS ->", build_portable(args.force))
    if args.only in (None, "playout"):
        print("[build_ref] scalar-engine playout driver ->", build_ref_playout(args.force))
    if args.only in (None, "v0_core"):
        print("[build_ref] v0_core ->", build_v0_core(args.force, args.cuda))
    if args.only is None and not args.no_pysrc:
        print("[build_ref] python host code ->", copy_pysrc())
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
