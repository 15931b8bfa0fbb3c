"""One fb200_ess launch at the c5 size (S = 30, N = 109 483 778) for an ncu capture; MODE = white | ar1 | nocut."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fortuna_b200 import ops  # noqa: E402

mode = os.environ.get("MODE", "white")
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
S, N = 30, 109_483_778
x = torch.randn((S, N), device=dev, generator=g) * 0.05 + 1.0
if mode != "white":
    tmp = torch.empty(N, device=dev)
    for t in range(1, S):
        torch.randn(N, device=dev, generator=g, out=tmp)
        x[t] = 0.9 * x[t - 1] + 0.436 * tmp
for _ in range(3):
    out = ops.ess(x, None if mode == "nocut" else 0.0)
torch.cuda.synchronize()
print(mode, float(out[:1000].mean()))
