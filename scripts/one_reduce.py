#!/usr/bin/env python
"""One K3 launch configuration at c4, for ncu captures: --mode tiled|rows|fused, --dtype f32|bf16."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fortuna_b200 import _lib, ops  # noqa: E402
from scripts.bench_reduce import WANT  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="fused")
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--S", type=int, default=64)
    ap.add_argument("--B", type=int, default=65536)
    ap.add_argument("--C", type=int, default=1000)
    ap.add_argument("--iters", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev)
    gen.manual_seed(1)
    lib = _lib.lib()
    S, B, Cn = a.S, a.B, a.C
    t = torch.randint(0, Cn, (B,), device=dev, generator=gen, dtype=torch.int32)
    n_val = 50000
    vl = torch.empty((n_val, Cn), device=dev).normal_(0.0, 4.0, generator=gen)
    vt = torch.randint(0, Cn, (n_val,), device=dev, generator=gen, dtype=torch.int32)
    vm = ops.bma_reduce_cls(vl.unsqueeze(0).contiguous(), ("mean",))["mean"]
    val_sorted = ops.sort_f32_(ops.aps_scores(vm, vt))
    thr = float((1 - 0.05) * (n_val + 1))
    z = torch.empty((S, B, Cn), dtype=torch.float32 if a.dtype == "f32" else torch.bfloat16, device=dev)
    for s in range(S):
        z[s].copy_(torch.empty((B, Cn), device=dev).normal_(0.0, 4.0, generator=gen))
    outs = ops._alloc_cls_outputs(WANT, S, B, Cn, dev)
    lib.fb200_debug_force_tiled(1 if a.mode == "tiled" else 0)
    for _ in range(a.iters):
        if a.mode == "fused":
            ops.bma_reduce_cls_sets(z, WANT, val_sorted, thr, targets=t, out=outs)
        else:
            ops.bma_reduce_cls(z, WANT, targets=t, out=outs)
    torch.cuda.synchronize()
    print("ok")


if __name__ == "__main__":
    main()
