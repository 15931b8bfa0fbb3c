#!/usr/bin/env python
"""Where the absolute floors in tests/test_gpu_predictive.py come from: the float64 error of BOTH the float32 oracle
(the reference's op order) and the CUDA kernel against an exact (float64) evaluation of the same statistic, on the
test shapes.  Prints one JSON line per shape: for every output the largest absolute error of oracle and kernel, the floor
the test uses, and the largest |value| for scale.  (Round-1 verdict, "What's weak" 4.)"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fortuna_b200 import ops  # noqa: E402
from oracle import predictive as P  # noqa: E402


def exact(z, t):
    z = z.astype(np.float64)
    m = z.max(-1, keepdims=True)
    e = np.exp(z - m)
    p = e / e.sum(-1, keepdims=True)
    logp = (z - m) - np.log(e.sum(-1, keepdims=True))
    S = z.shape[0]
    mean = p.mean(0)
    alea = (p * (1 - p)).mean(0)
    epi = np.maximum(0.0, (p * p).mean(0) - mean * mean)
    ent = -(mean * np.log(np.where(mean > 0, mean, 1.0))).sum(-1)
    alea_ent = -(p * logp).sum(-1).mean(0)
    ll = np.take_along_axis(logp, np.broadcast_to(t[None, :, None], (S, len(t), 1)), 2)[..., 0]
    mx = ll.max(0)
    lp = mx + np.log(np.exp(ll - mx).sum(0)) - np.log(S)
    return {"mean": mean, "aleatoric_variance": alea, "epistemic_variance": epi, "variance": alea + epi, "entropy": ent,
            "aleatoric_entropy": alea_ent, "epistemic_entropy": ent - alea_ent, "log_prob": lp}


def main():
    dev = torch.device("cuda:0")
    names = ("mean", "aleatoric_variance", "epistemic_variance", "variance", "entropy", "aleatoric_entropy",
             "epistemic_entropy", "log_prob")
    for S, B, C in ((8, 512, 1000), (64, 1024, 10), (30, 500, 2), (64, 256, 1000)):
        z = (4.0 * np.random.default_rng(S * 1000 + C).standard_normal((S, B, C))).astype(np.float32)
        t = np.random.default_rng(1).integers(0, C, B).astype(np.int32)
        ex = exact(z, t)
        orc = {"mean": P.cls_mean(z), "aleatoric_variance": P.cls_aleatoric_variance(z), "epistemic_variance": P.cls_epistemic_variance(z),
               "variance": P.cls_variance(z), "entropy": P.cls_entropy(z), "aleatoric_entropy": P.cls_aleatoric_entropy(z),
               "epistemic_entropy": P.cls_epistemic_entropy(z), "log_prob": P.cls_log_prob(z, t)}
        got = ops.bma_reduce_cls(torch.from_numpy(z).to(dev), names, targets=torch.from_numpy(t).to(dev))
        mean_max = float(ex["mean"].max())
        floors = {"mean": 1e-9, "aleatoric_variance": 2e-7, "epistemic_variance": 4e-7 * mean_max ** 2 + 2e-7,
                  "variance": 4e-7 * mean_max ** 2 + 2e-7, "entropy": 2e-6, "aleatoric_entropy": 2e-6, "epistemic_entropy": 4e-6,
                  "log_prob": 2e-6}
        row = {"shape": [S, B, C]}
        for n in names:
            k = got[n].cpu().numpy().astype(np.float64)
            row[n] = {"oracle_f32_max_abs_err": float(np.abs(orc[n].astype(np.float64) - ex[n]).max()),
                      "kernel_max_abs_err": float(np.abs(k - ex[n]).max()), "test_atol": floors[n],
                      "max_abs_value": float(np.abs(ex[n]).max())}
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
