"""The HBM-bound rule kernels at the op-level bench sizes, three calls each (for an ncu capture: tools/gpu_ops_ncu.sh)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from liuzhou_b200 import native, v0_core  # noqa: E402
from tests._util import random_apply_batch, random_mask_states, to_torch  # noqa: E402

DEV = "cuda:0"
b = 65536 * int(sys.argv[1]) if len(sys.argv) > 1 else 65536
st = random_mask_states(65536, 0xF00DCAFE)
t = to_torch(st, DEV)
if b > 65536:
    t = [x.repeat((b // 65536,) + (1,) * (x.dim() - 1)).contiguous() for x in t]
n, n_base = 1 << 20, 1 << 14
st2, codes, parents = random_apply_batch(n_base, 0xA11CEB0B)
t2 = to_torch(st2, DEV)
c = torch.from_numpy(codes).to(DEV).repeat(n // n_base, 1).contiguous()
p = torch.from_numpy(parents).to(DEV).repeat(n // n_base).contiguous()
bp = 1 << 22
packed = native.pack_states(to_torch(st, DEV)).repeat(bp // 65536, 1).contiguous()
acts = torch.zeros((bp,), dtype=torch.int32, device=DEV)
for _ in range(3):
    v0_core.encode_actions_fast(*t[:10], 36, 144, 36, 4)
    v0_core.batch_apply_moves(*t2, c, p)
    v0_core.states_to_model_input(*t[:5])
    native.legal_masks(packed)
    native.apply_actions(packed, acts)
    torch.cuda.synchronize()
print("ok")
