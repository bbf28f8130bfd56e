#!/usr/bin/env python3
"""Golden vectors for the self-play payload formats, produced by the REFERENCE's own python (imported from
/root/reference in the build container; needs oracle/_ref for its `v0_core` import):

    python tests/golden/make_storage_golden.py      ->  tests/golden/storage_formats.json

Contents: `plan_sample_ranges` on a parameter grid, `_summarize_scalar_targets` / `_merge_scalar_target_summaries`
on seeded vectors, `_merge_self_play_stats` on two synthetic workers, and the manifest the reference's
`_save_self_play_payload_sharded` writes for a seeded synthetic batch (file names, sizes, keys).
The synthetic inputs are rebuilt from seeds by tests/test_storage_formats.py (`synthetic_batch`, `target_vector`).
"""
import json
import sys
import tempfile
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from tests.test_storage_formats import PLAN_GRID, STATS_A, STATS_B, synthetic_batch, target_vector  # noqa: E402

sys.path[:0] = ["/root/reference", str(ROOT / "oracle" / "_ref")]   # after our own `tests` package is bound

import v1.train as T  # noqa: E402
from v1.python.self_play_storage import estimate_bytes_per_sample, plan_sample_ranges  # noqa: E402
from v1.python.self_play_types import SelfPlayV1Stats  # noqa: E402
from v1.python.trajectory_buffer import TensorSelfPlayBatch  # noqa: E402


def main():
    out = {}
    out["plan"] = [[list(r) for r in plan_sample_ranges(**kw)] for kw in PLAN_GRID]
    out["summaries"] = [T._summarize_scalar_targets(target_vector(seed, n)) for seed, n in ((1, 0), (2, 257), (3, 4099))]
    out["summary_merged"] = T._merge_scalar_target_summaries(out["summaries"])
    merged = T._merge_self_play_stats([SelfPlayV1Stats(**STATS_A), SelfPlayV1Stats(**STATS_B)], elapsed_sec=12.5)
    out["stats_merged"] = merged.to_dict()
    b = synthetic_batch(1000, 7)
    ref_batch = TensorSelfPlayBatch(b.state_tensors, b.legal_masks, b.policy_targets, b.value_targets, b.soft_value_targets)
    out["bytes_per_sample"] = estimate_bytes_per_sample(ref_batch)
    with tempfile.TemporaryDirectory() as d:
        path = str(Path(d) / "selfplay_iter_001.pt")
        n = T._save_self_play_payload_sharded(path=path, samples=ref_batch, stats=SelfPlayV1Stats(**STATS_A),
                                              metadata={"iteration": 1}, num_shards=3, chunk_target_bytes=300 * 2692)
        man = torch.load(path)
        out["sharded"] = {"count": n, "shard_files": man["shard_files"], "shard_sizes": man["shard_sizes"],
                          "manifest_keys": sorted(man.keys()), "avg_bytes_per_sample": man["avg_bytes_per_sample"],
                          "shard_keys": sorted(torch.load(str(Path(d) / man["shard_files"][0])).keys()),
                          "shard_meta_keys": sorted(torch.load(str(Path(d) / man["shard_files"][0]))["metadata"].keys())}
    # stable bootstrap init (v1/train.py:162-217) of the reference's ChessNet under the production seed
    import hashlib

    from src.neural_network import ChessNet as RefNet

    torch.manual_seed(123)
    net = RefNet()
    T._init_model_stable_resnet(net, seed=20260314)
    h = hashlib.sha256()
    for k, v in net.state_dict().items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    out["stable_init_sha256"] = h.hexdigest()
    # policy targets / deterministic move choice of the portable search (v1/python/portable_mcts.py:149-261)
    from tests.test_storage_formats import policy_cases
    from v1.python.portable_mcts import deterministic_action_from_search, policy_from_visits_and_priors

    pol = []
    for visits, priors, legal, qv, temp, beta in policy_cases():
        idx = torch.where(legal)[0]
        dense = torch.zeros(220)
        dense[idx] = policy_from_visits_and_priors(visits[idx].float(), priors[idx], temperature=temp, prior_pseudocount=beta)
        pol.append({"policy": dense.tolist(), "choice": deterministic_action_from_search(visits, qv, priors, legal)})
    out["policy_targets"] = pol
    (Path(__file__).resolve().parent / "storage_formats.json").write_text(json.dumps(out, indent=1, sort_keys=True))
    print("wrote storage_formats.json")


if __name__ == "__main__":
    main()
