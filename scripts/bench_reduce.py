#!/usr/bin/env python
"""K3 (single-pass reductions) and K5+K6 (APS conformal sets) alone at c4 (z[64,65536,1000]), float32 and bf16 logits.
CUDA events, inputs larger than L2."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fortuna_b200 import ops  # noqa: E402

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
    ap.add_argument("--val-samples", type=int, default=1,
                    help="posterior samples behind the validation probabilities (1: peaked rows, so the quantile sits deep "
                         "in the flatter S-sample test rows; S: the validation rows are BMA means like the test rows)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    S, B, Cn = args.S, args.B, args.C
    gen = torch.Generator(device=dev)
    gen.manual_seed(1)
    peak = 6541.5
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    t = torch.randint(0, Cn, (B,), device=dev, generator=gen, dtype=torch.int32)
    n_val = 50000
    vl = torch.empty((args.val_samples, n_val, Cn), device=dev).normal_(0.0, 4.0, generator=gen)
    vt = torch.randint(0, Cn, (n_val,), device=dev, generator=gen, dtype=torch.int32)
    vm = ops.bma_reduce_cls(vl, ("mean",))["mean"]
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
        res = {"dtype": dt, "S": S, "B": B, "C": Cn, "algorithmic_GB": alg / 1e9}
        ms = timeit(lambda: ops.bma_reduce_cls(z, WANT, targets=t, out=outs), args.steps)
        res["reduce_ms"] = ms
        res["reduce_hbm_frac"] = alg / (ms * 1e-3) / 1e9 / peak
        ms2 = timeit(lambda: ops.aps_conformal_sets(outs["mean"], val_sorted, thr), args.steps)
        res["sets_kernel_ms"] = ms2
        m2, s2 = ops.aps_conformal_sets(outs["mean"], val_sorted, thr)
        res["mean_set_size"] = float(s2.float().mean())
        res["val_samples"] = args.val_samples
        m3, s3, _ = ops.aps_conformal_sets(outs["mean"], val_sorted, thr, want_scores=True)  # sorts every row
        res["sets_equal_full_sort"] = bool(torch.equal(m2, m3) and torch.equal(s2, s3))
        print(json.dumps(res), flush=True)
        del z


if __name__ == "__main__":
    main()
