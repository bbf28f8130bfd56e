"""Compact trajectory wire format (liuzhou_b200/compact.py): lossless round trip on positions produced by the CPU
oracle (real planes / legal masks of reachable states), size accounting, rejection of non-canonical planes."""
import numpy as np
import pytest
import torch

import oracle
from liuzhou_b200 import compact as cp
from liuzhou_b200.trajectory_buffer import TensorSelfPlayBatch


def _oracle_batch(games=6, seed=11):
    planes, masks, pols = [], [], []
    rng = np.random.default_rng(seed)
    for g in range(games):
        st = oracle.initial_states(1)
        for a in oracle.random_playout(seed, g, 512, want_trace=True)["trace"]:
            m, _ = oracle.encode_actions_fast(st)
            planes.append(oracle.states_to_model_input(st)[0])
            masks.append(m[0])
            p = rng.random(220).astype(np.float32) * m[0]
            p[rng.random(220) < 0.3] = 0.0                       # legal actions with zero visits
            s = p.sum()
            pols.append(p / s if s > 0 else p)
            st = oracle.apply_move_scalar(st, int(a))
    n = len(planes)
    vt = torch.from_numpy(rng.choice([-1.0, 0.0, 1.0, np.nan], size=n).astype(np.float32))
    return TensorSelfPlayBatch(torch.from_numpy(np.stack(planes)), torch.from_numpy(np.stack(masks)),
                               torch.from_numpy(np.stack(pols)), vt, torch.from_numpy(rng.random(n).astype(np.float32)))


def _same(a: TensorSelfPlayBatch, b: TensorSelfPlayBatch):
    assert torch.equal(a.state_tensors, b.state_tensors) and a.state_tensors.dtype == b.state_tensors.dtype
    assert torch.equal(a.legal_masks, b.legal_masks) and b.legal_masks.dtype == torch.bool
    assert torch.equal(a.policy_targets, b.policy_targets)
    assert torch.equal(torch.isnan(a.value_targets), torch.isnan(b.value_targets))
    assert torch.equal(torch.nan_to_num(a.value_targets), torch.nan_to_num(b.value_targets))
    assert torch.equal(a.soft_value_targets, b.soft_value_targets)


def test_round_trip_is_lossless_and_15x_smaller():
    batch = _oracle_batch()
    c = cp.compact(batch)
    assert c.num_samples == batch.num_samples > 500
    _same(batch, cp.expand(c))
    ratio = batch.nbytes() / c.nbytes()
    assert ratio > 10.0, ratio                                  # 2,692 B -> < 270 B per position
    halves = [TensorSelfPlayBatch(*(t[a:b] for t in (batch.state_tensors, batch.legal_masks, batch.policy_targets,
                                                      batch.value_targets, batch.soft_value_targets)))
              for a, b in ((0, 100), (100, batch.num_samples))]
    _same(batch, cp.expand(cp.concat([cp.compact(h) for h in halves])))


def test_empty_batch_and_rejections():
    empty = TensorSelfPlayBatch(torch.zeros((0, 11, 6, 6)), torch.zeros((0, 220), dtype=torch.bool),
                                torch.zeros((0, 220)), torch.zeros((0,)), torch.zeros((0,)))
    e = cp.expand(cp.compact(empty))
    assert e.num_samples == 0 and tuple(e.state_tensors.shape) == (0, 11, 6, 6)
    batch = _oracle_batch(games=1)
    bad = TensorSelfPlayBatch(batch.state_tensors.clone(), batch.legal_masks, batch.policy_targets,
                              batch.value_targets, batch.soft_value_targets)
    bad.state_tensors[0, 0, 0, 0] = 0.5
    with pytest.raises(ValueError):
        cp.compact(bad)
    bad.state_tensors[0, 0, 0, 0] = 0.0
    bad.state_tensors[1, 4:] = 1.0                              # two phase planes on
    with pytest.raises(ValueError):
        cp.compact(bad)


def test_fixed_rows_round_trip_and_flags():
    """Streaming form (compact_rows_fixed / expand_rows_fixed): 336 B per position, static shapes, bit-identical
    expansion on positions of real games; rows that cannot be represented are flagged, not silently mangled."""
    b = _oracle_batch()
    rows = cp.compact_rows_fixed(b)
    n = b.num_samples
    assert rows.dtype == torch.int64 and tuple(rows.shape) == (n, cp.FIXED_ROW_WORDS) and cp.FIXED_ROW_WORDS * 8 == 336
    assert int(rows[:, 41].sum()) == 0
    assert int(b.legal_masks.sum(1).max()) <= 64                 # the bound the format relies on
    _same(cp.expand_rows_fixed(rows), b)
    _same(cp.expand_rows_fixed(rows), cp.expand(cp.compact(b)))   # and it agrees with the CSR form
    # policy mass on an illegal action / more than 64 legal actions / non-canonical planes -> flags
    bad = TensorSelfPlayBatch(b.state_tensors.clone(), b.legal_masks.clone(), b.policy_targets.clone(),
                              b.value_targets.clone(), b.soft_value_targets.clone())
    illegal = int((~bad.legal_masks[0]).nonzero()[0])
    bad.policy_targets[0, illegal] = 0.5
    bad.legal_masks[1, :80] = True
    bad.state_tensors[2, 0, 0, 0] = 0.5
    flags = cp.compact_rows_fixed(bad)[:, 41]
    assert flags[:3].tolist() == [1, 1, 1] and int(flags[3:].sum()) == 0
    with pytest.raises(ValueError):
        cp.expand_rows_fixed(cp.compact_rows_fixed(bad))
    empty = TensorSelfPlayBatch(b.state_tensors[:0], b.legal_masks[:0], b.policy_targets[:0], b.value_targets[:0],
                                b.soft_value_targets[:0])
    assert cp.expand_rows_fixed(cp.compact_rows_fixed(empty)).num_samples == 0
