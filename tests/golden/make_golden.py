#!/usr/bin/env python3
"""Generate the committed golden vectors from the REFERENCE's own binaries (oracle/_ref).

Run in the build container (needs /root/reference -> `python oracle/build_ref.py` first):
    python tests/golden/make_golden.py
Outputs (small, compressed) under tests/golden/:
    encode_actions.npz   2,048 random + 1,024 sparse states -> v0_core.encode_actions_fast (CPU) mask + metadata
    apply_moves.npz      all legal children of the states of 6 playouts -> v0_core.batch_apply_moves (CPU)
    scalar_playouts.npz  24 playouts: per-ply legal index lists + next states from the scalar engine
                         (_liuzhou_portable_cpp.inspect_state / apply_action)
    root_puct.npz        v0_core.root_puct_allocate_visits (CPU/ATen) on 3 shapes
    tree_mcts.npz        _liuzhou_portable_cpp.PortableTreeBatch visit counts / Q / root values, 96 sims
    composites.npz       root_pack_sparse_actions / root_finalize_from_visits / self_play_step_inplace /
                         states_to_model_input / project_policy_logits_fast outputs
Inputs are stored next to the outputs, so consumers never need the reference.
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import oracle  # noqa: E402  (only used to drive playouts; outputs stored are the REFERENCE's)
from tests._util import (STATE_FIELDS, concat_states, dict_to_state, fake_net, load_ref,  # noqa: E402
                         random_mask_states, sparse_random_states, state_obj, to_torch)

OUT = Path(__file__).resolve().parent


def pack_states(prefix, st):
    return {f"{prefix}{k}": np.asarray(st[k]) for k in STATE_FIELDS}


def main():
    ref = load_ref()
    assert ref is not None, "build oracle/_ref first"
    v0_core, portable = ref

    # 1. encode_actions_fast
    st = concat_states([random_mask_states(2048, 0xF00DCAFE), sparse_random_states(1024, 0x5EED)])
    out = {}
    for aux in (1, 4):
        mask, meta = v0_core.encode_actions_fast(*to_torch(st)[:10], 36, 144, 36, aux)
        out[f"mask_aux{aux}"] = np.packbits(mask.numpy(), axis=1)
        out[f"meta_aux{aux}"] = meta.numpy().astype(np.int8)  # values in [-1, 35]
    np.savez_compressed(OUT / "encode_actions.npz", **pack_states("", st), **out)

    # 2. scalar playouts (reference scalar engine decides everything)
    rec = {"legal_ptr": [0], "legal_idx": [], "chosen": [], "game_ptr": [0], "over": [], "winner": []}
    states = []
    for g in range(24):
        s = oracle.initial_states(1)
        ply = 0
        while True:
            info = portable.inspect_state(state_obj(s))
            legal = list(info["legal_action_indices"])
            states.append(s)
            rec["legal_idx"] += legal
            rec["legal_ptr"].append(len(rec["legal_idx"]))
            rec["over"].append(bool(info["game_over"]))
            rec["winner"].append(int(info["winner"]))
            if not legal:
                rec["chosen"].append(-1)
                break
            a = legal[oracle.lib().or_playout_pick(0x60D, g, ply, len(legal))]
            rec["chosen"].append(a)
            s = dict_to_state(portable.apply_action(state_obj(s), int(a)))
            ply += 1
        rec["game_ptr"].append(len(states))
    all_states = concat_states(states)
    np.savez_compressed(OUT / "scalar_playouts.npz", **pack_states("", all_states),
                        **{k: np.asarray(v, np.int32) for k, v in rec.items()})

    # 3. batch_apply_moves on all legal children of 6 of those games
    end = rec["game_ptr"][6]
    sub = {k: all_states[k][:end] for k in STATE_FIELDS}
    mask, meta = v0_core.encode_actions_fast(*to_torch(sub)[:10], 36, 144, 36, 4)
    rows, cols = np.nonzero(mask.numpy())
    codes = meta.numpy()[rows, cols]
    parents = rows.astype(np.int64)
    children = v0_core.batch_apply_moves(*to_torch(sub), torch.from_numpy(codes), torch.from_numpy(parents))
    np.savez_compressed(OUT / "apply_moves.npz", **pack_states("in_", sub), codes=codes, parents=parents,
                        **{f"out_{k}": t.numpy() for k, t in zip(STATE_FIELDS, children)})

    # 4. root PUCT
    rng = np.random.default_rng(5)
    out = {}
    for tag, (r, m, sims, c) in {"a": (48, 40, 200, 1.0), "b": (8, 220, 800, 1.0), "c": (16, 7, 64, 1.5)}.items():
        valid = rng.random((r, m)) < 0.8
        valid[:, 0] = True
        pri = rng.random((r, m)).astype(np.float32) * valid
        pri = (pri / pri.sum(1, keepdims=True)).astype(np.float32)
        leaf = ((rng.random((r, m)).astype(np.float32) * 2 - 1) * valid).astype(np.float32)
        leaf[: r // 2] = np.round(leaf[: r // 2] * 4) / 4
        pri[: r // 4] = (valid[: r // 4] / valid[: r // 4].sum(1, keepdims=True)).astype(np.float32)
        v, w, rv = v0_core.root_puct_allocate_visits(torch.from_numpy(pri), torch.from_numpy(leaf),
                                                     torch.from_numpy(valid), sims, c)
        out.update({f"{tag}_priors": pri, f"{tag}_leaf": leaf, f"{tag}_valid": valid,
                    f"{tag}_sims": np.int64(sims), f"{tag}_c": np.float32(c),
                    f"{tag}_visits": v.numpy(), f"{tag}_value_sum": w.numpy(), f"{tag}_root_values": rv.numpy()})
    np.savez_compressed(OUT / "root_puct.npz", **out)

    # 5. tree MCTS (fake net defined above is deterministic in the model input only)
    idx = np.arange(0, all_states["board"].shape[0], 97)[:24]
    roots = {k: all_states[k][idx] for k in STATE_FIELDS}
    n = idx.size
    tb = portable.PortableTreeBatch([state_obj(roots, i) for i in range(n)], exploration_weight=1.25, num_threads=2)
    pend = tb.prepare_roots()
    tb.complete_pending(*fake_net(pend["model_inputs"], pend["legal_masks"], 0))
    for _ in range(96):
        pend = tb.select_leaves()
        tb.complete_pending(*fake_net(pend["model_inputs"], pend["legal_masks"], 1))
    ro = tb.root_outputs()
    np.savez_compressed(OUT / "tree_mcts.npz", **pack_states("", roots), sims=np.int64(96), c=np.float64(1.25),
                        visit_counts=ro["visit_counts"], root_action_values=ro["root_action_values"],
                        root_values=ro["root_values"], terminal=ro["terminal"], legal_masks=ro["legal_masks"],
                        root_priors=tb.root_priors()["priors"])

    # 6. composites
    sub = {k: all_states[k][::5] for k in STATE_FIELDS}
    b = sub["board"].shape[0]
    t = to_torch(sub)
    mask, meta = v0_core.encode_actions_fast(*t[:10], 36, 144, 36, 4)
    mask = mask.clone()
    mask[::17] = False
    rng = np.random.default_rng(8)
    probs = (rng.random((b, 220)).astype(np.float32) + 0.01) * mask.numpy()
    probs = (probs / np.maximum(probs.sum(1, keepdims=True), 1e-8)).astype(np.float32)
    pack = v0_core.root_pack_sparse_actions(mask, torch.from_numpy(probs), meta)
    valid_mask, legal_idx, priors, code_mat = pack[3], pack[4], pack[5], pack[6]
    leaf = torch.from_numpy((rng.random(tuple(priors.shape)).astype(np.float32) * 2 - 1)) * valid_mask
    visits, value_sum, _ = v0_core.root_puct_allocate_visits(priors, leaf, valid_mask, 128, 1.0)
    temps = torch.from_numpy(np.where(np.arange(pack[1].numel()) % 2 == 0, 1.0, 0.5).astype(np.float32))
    fin = v0_core.root_finalize_from_visits(legal_idx, code_mat, valid_mask, visits, value_sum, pack[1], b, 220,
                                            temps, False)
    model_input = v0_core.states_to_model_input(*t[:5])
    heads = [torch.log_softmax(torch.from_numpy(rng.standard_normal((b, 36)).astype(np.float32)), 1) for _ in range(3)]
    proj = v0_core.project_policy_logits_fast(*heads, mask, 36, 144, 36, 4)
    # self_play_step_inplace
    plies = sub["move_count"].copy()
    done = np.zeros((b,), bool)
    done[::11] = True
    active = np.nonzero(~done)[0].astype(np.int64)
    m_np, meta_np = mask.numpy(), meta.numpy()
    terminal = ~m_np[active].any(1)
    codes = np.full((active.size, 4), -1, np.int32)
    cvalid = np.zeros((active.size,), bool)
    for j, g in enumerate(active):
        ii = np.nonzero(m_np[g])[0]
        if ii.size and rng.random() > 0.03:
            codes[j] = meta_np[g, rng.choice(ii)]
            cvalid[j] = True
    t_state = [x.clone() for x in t]
    t_plies, t_done = torch.from_numpy(plies.copy()), torch.from_numpy(done.copy())
    step = v0_core.self_play_step_inplace(*t_state, t_plies, t_done, torch.from_numpy(active), torch.from_numpy(codes),
                                          torch.from_numpy(terminal), torch.from_numpy(cvalid), 130, 2.0)
    np.savez_compressed(
        OUT / "composites.npz", **pack_states("", sub), mask=m_np, meta=meta_np.astype(np.int8), probs=probs,
        **{f"pack{i}": x.numpy() for i, x in enumerate(pack)}, leaf=leaf.numpy(), visits=visits.numpy(),
        value_sum=value_sum.numpy(), temps=temps.numpy(), **{f"fin{i}": x.numpy() for i, x in enumerate(fin)},
        model_input=np.packbits(model_input.numpy().astype(np.uint8)), head0=heads[0].numpy(), head1=heads[1].numpy(),
        head2=heads[2].numpy(), proj_probs=proj[0].numpy(), proj_logits=proj[1].numpy(),
        step_plies_in=plies, step_done_in=done, step_active=active, step_codes=codes, step_terminal=terminal,
        step_valid=cvalid, **{f"step_out{i}": x.numpy() for i, x in enumerate(step)},
        **{f"step_state_{k}": x.numpy() for k, x in zip(STATE_FIELDS, t_state)},
        step_plies_out=t_plies.numpy(), step_done_out=t_done.numpy())
    for f in sorted(OUT.glob("*.npz")):
        print(f.name, f.stat().st_size // 1024, "KiB")


if __name__ == "__main__":
    main()
