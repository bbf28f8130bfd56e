"""Boundary proof: the reference's UNMODIFIED python host code (`v1/python/mcts_gpu.py`, `self_play_gpu_runner.py`,
`trajectory_buffer.py`, copied byte for byte into the git-ignored oracle/_ref/pysrc) runs over `liuzhou_b200.v0_core`
installed as `sys.modules["v0_core"]` (/root/reference/v1/python/mcts_gpu.py:23, self_play_gpu_runner.py:10) and
produces EXACTLY what it produces over the reference's own `v0_core` CUDA extension (oracle/_ref/v0_core*.so) on the
same GPU: `V1RootMCTS.search_batch` outputs and a whole 64-game `self_play_v1_gpu` trajectory batch, `torch.equal`.

Each side runs in its own process (oracle/ref_runner.py), because a process can hold only one module named v0_core.
Noise off / sample_moves off: the remaining RNG-free pipeline is deterministic, so equality is bit for bit."""
import json
import subprocess
import sys

import numpy as np
import pytest
import torch

import oracle
from tests._util import REF_DIR, ROOT, STATE_FIELDS

pytestmark = pytest.mark.gpu

RUNNER = ROOT / "oracle" / "ref_runner.py"


def _have_reference() -> bool:
    return (REF_DIR / "pysrc" / "v1" / "python" / "mcts_gpu.py").exists() and bool(list(REF_DIR.glob("v0_core*.so")))


def _run(*args, timeout=1500):
    res = subprocess.run([sys.executable, str(RUNNER), *map(str, args)], cwd=str(ROOT), stdout=subprocess.PIPE,
                         stderr=subprocess.PIPE, text=True, timeout=timeout)
    assert res.returncode == 0, f"ref_runner {args} failed:\n{res.stdout[-2000:]}\n{res.stderr[-6000:]}"
    return json.loads(res.stdout.strip().splitlines()[-1])


def _mixed_states(n_games=12, stride=5):
    """Reachable positions of all phases: every `stride`-th state of seeded uniform-random playouts."""
    states = []
    for g in range(n_games):
        st = oracle.initial_states(1)
        for i, a in enumerate(oracle.random_playout(0xB0A7, g, 512, want_trace=True)["trace"]):
            if i % stride == g % stride:
                states.append(st)
            st = oracle.apply_move_scalar(st, int(a))
    return {k: torch.from_numpy(np.ascontiguousarray(np.concatenate([s[k] for s in states]))) for k in STATE_FIELDS}


@pytest.mark.skipif(not _have_reference(), reason="oracle/_ref (reference binaries + python copy) not present")
def test_reference_search_batch_over_shim_equals_reference_cuda(tmp_path):
    st = _mixed_states()
    n = st["board"].shape[0]
    assert n >= 150
    torch.save(st, tmp_path / "states.pt")
    outs = {}
    for side in ("ref", "shim"):
        info = _run("search", "--v0core", side, "--states", tmp_path / "states.pt", "--sims", 200, "--dump",
                    tmp_path / f"search_{side}.pt")
        assert info["roots"] == n
        outs[side] = torch.load(tmp_path / f"search_{side}.pt")
    for k, ref in outs["ref"].items():
        got = outs["shim"][k]
        assert got.dtype == ref.dtype and got.shape == ref.shape, k
        assert torch.equal(got, ref), (k, (got.float() - ref.float()).abs().max().item())
    assert int(outs["ref"]["chosen_valid_mask"].sum()) > 100          # the comparison was not vacuous


@pytest.mark.skipif(not _have_reference(), reason="oracle/_ref (reference binaries + python copy) not present")
def test_reference_self_play_over_shim_equals_reference_cuda(tmp_path):
    outs, infos = {}, {}
    for side in ("ref", "shim"):
        infos[side] = _run("selfplay", "--v0core", side, "--games", 64, "--sims", 64, "--noise", 0, "--sample", 0,
                           "--dump", tmp_path / f"sp_{side}.pt")
        outs[side] = torch.load(tmp_path / f"sp_{side}.pt")
    assert infos["ref"]["positions"] == infos["shim"]["positions"] > 64 * 20
    for k in ("black_wins", "white_wins", "draws", "avg_game_length"):
        assert infos["ref"][k] == infos["shim"][k], k
    for k, ref in outs["ref"].items():
        got = outs["shim"][k]
        assert got.dtype == ref.dtype and got.shape == ref.shape, k
        assert torch.equal(torch.nan_to_num(got.float(), nan=-7.0), torch.nan_to_num(ref.float(), nan=-7.0)), (
            k, (torch.nan_to_num(got.float()) - torch.nan_to_num(ref.float())).abs().max().item())
