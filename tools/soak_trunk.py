"""Race soak for the fused trunk kernel's relaxed copy hand-off: the same input through the kernel N times, every output
compared bit for bit with the first one (a lost or early operand copy would change at least one element), at several batch
sizes, while a second stream keeps the memory system busy."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from liuzhou_b200.net import ChessNet, InferenceNet  # noqa: E402

torch.manual_seed(1)
net = InferenceNet(ChessNet(), "cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
noise_stream = torch.cuda.Stream()
junk = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
bad_total = 0
for n in (64, 192, 1024, 4096):
    x = net.new_input(n)
    x[:, :11] = (torch.rand((n, 11, 6, 6), device="cuda") > 0.6).to(torch.bfloat16)
    ref = net._trunk_heads_conv(x).clone()
    bad = torch.zeros((), dtype=torch.int64, device="cuda")
    for i in range(reps):
        if i % 8 == 0:
            with torch.cuda.stream(noise_stream):
                junk.add_(1.0)
        out = net._trunk_heads_conv(x)
        bad += (out != ref).any().to(torch.int64)
    torch.cuda.synchronize()
    print(f"n={n}: {reps} launches, {int(bad)} differing outputs")
    bad_total += int(bad)
print("SOAK", "FAILED" if bad_total else "clean")
