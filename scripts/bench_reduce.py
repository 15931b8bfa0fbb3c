#!/usr/bin/env python
"""A/B timing of the K3 reduction designs at c4 (z[64,65536,1000]): tiled kernel vs warp-autonomous kernel, float32 and
bf16 logits, with and without the fused APS conformal sets.  CUDA events, inputs larger than L2."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fortuna_b200 import _lib, ops  # noqa: E402

WANT = ("mean", "aleatoric_variance", "epistemic_variance", "entropy", "aleatoric_entropy", "epistemic_entropy", "log_prob",
        "mode", "confidence", "brier_terms")


def timeit(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(steps):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--S", type=int, default=64)
    ap.add_argument("--B", type=int, default=65536)
    ap.add_argument("--C", type=int, default=1000)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--dtypes", default="f32,bf16")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    S, B, Cn = args.S, args.B, args.C
    gen = torch.Generator(device=dev)
    gen.manual_seed(1)
    lib = _lib.lib()
    peak = 6541.5
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    t = torch.randint(0, Cn, (B,), device=dev, generator=gen, dtype=torch.int32)
    n_val = 50000
    vl = torch.empty((n_val, Cn), device=dev).normal_(0.0, 4.0, generator=gen)
    vt = torch.randint(0, Cn, (n_val,), device=dev, generator=gen, dtype=torch.int32)
    vm = ops.bma_reduce_cls(vl.unsqueeze(0).contiguous(), ("mean",))["mean"]
    val_sorted = ops.sort_f32_(ops.aps_scores(vm, vt))
    thr = float((1 - 0.05) * (n_val + 1))
    del vl, vm
    for dt in args.dtypes.split(","):
        z = torch.empty((S, B, Cn), dtype=torch.float32 if dt == "f32" else torch.bfloat16, device=dev)
        for s in range(S):
            z[s].copy_(torch.empty((B, Cn), device=dev).normal_(0.0, 4.0, generator=gen))
        esz = z.element_size()
        alg = esz * S * B * Cn + 4.0 * 3 * B * Cn + 4.0 * B * 8
        outs = ops._alloc_cls_outputs(WANT, S, B, Cn, dev)
        mask = torch.empty((B, Cn), dtype=torch.uint8, device=dev)
        sizes = torch.empty(B, dtype=torch.int32, device=dev)
        ws = torch.empty(int(lib.fb200_bma_reduce_cls_sets_workspace_bytes(B)), dtype=torch.uint8, device=dev)
        res = {"dtype": dt, "S": S, "B": B, "C": Cn, "algorithmic_GB": alg / 1e9}
        for name, tiled in (("tiled", 1), ("rows", 0)):
            lib.fb200_debug_force_tiled(tiled)
            ms = timeit(lambda: ops.bma_reduce_cls(z, WANT, targets=t, out=outs), args.steps)
            res[name + "_ms"] = ms
            res[name + "_hbm_frac"] = alg / (ms * 1e-3) / 1e9 / peak
        lib.fb200_debug_force_tiled(0)  # warp-autonomous kernel: fused epilogue
        ms = timeit(lambda: ops.bma_reduce_cls_sets(z, WANT, val_sorted, thr, targets=t, out=outs, mask=mask, sizes=sizes,
                                                    workspace=ws), args.steps)
        res["rows_fused_sets_ms"] = ms
        res["fallback_rows"] = int(ws[:4].view(torch.int32)[0])
        lib.fb200_debug_force_tiled(1)
        ms2 = timeit(lambda: ops.aps_conformal_sets(outs["mean"], val_sorted, thr), args.steps)
        res["separate_sets_kernel_ms"] = ms2
        m2, s2 = ops.aps_conformal_sets(outs["mean"], val_sorted, thr)
        res["fused_equals_separate"] = bool(torch.equal(m2, mask) and torch.equal(s2, sizes))
        res["mean_set_size"] = float(sizes.float().mean())
        print(json.dumps(res), flush=True)
        del z


if __name__ == "__main__":
    main()
