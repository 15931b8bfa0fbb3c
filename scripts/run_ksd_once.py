"""One fb200_ksd_imq call at the c5 size (S = 30, N = 109 483 778) for an ncu capture."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fortuna_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
S, N = 30, 109_483_778
x = torch.randn((S, N), device=dev, generator=g) * 0.05 + 1.0
gr = torch.randn((S, N), device=dev, generator=g)
for _ in range(2):
    out = ops.ksd_imq(x, gr)
torch.cuda.synchronize()
print(float(out[0]))
