"""CUDA-event timings of the diagnostics kernels (fb200_ksd_imq, fb200_ess) at the BASELINE configs' sizes, next to the
numpy oracle on a bounded sample.  usage: python scripts/bench_diagnostic.py [--cpu]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fortuna_b200 import ops  # noqa: E402


def timed(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu", action="store_true", help="also time the numpy oracle on the small configs")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    rows = []
    for name, S, N in (("c1 two-moons MLP", 30, 1082), ("c2 LeNet-5, 100 samples", 100, 44426),
                       ("c3 ResNet-18 last layer, 256 samples", 256, 5130), ("c5 BERT-base tree, 30 samples", 30, 109_483_778)):
        x = torch.randn((S, N), device=dev, generator=g) * 0.05 + 1.0
        gr = torch.randn((S, N), device=dev, generator=g)
        ksd_ms = timed(lambda: ops.ksd_imq(x, gr))
        ess_ms = timed(lambda: ops.ess(x, 0.0))
        pairs = S * (S + 1) / 2
        row = {
            "config": name, "S": S, "N": N,
            "ksd_ms": ksd_ms,
            # per pair and column: 1 subtract + 4 FMA = 9 flop
            "ksd_fp32_tflops": 9 * (S * (S + 1) / 2) * N / ksd_ms / 1e9,  # useful flop: pairs i <= j only
            "ksd_pairs": pairs,
            "ksd_hbm_gbs": 2 * 4 * S * N / ksd_ms / 1e6,
            "ess_ms": ess_ms,
            "ess_fma_tflops": 2 * (S * (S + 1) / 2) * N / ess_ms / 1e9,
            "ess_hbm_gbs": (4 * S * N + 4 * N) / ess_ms / 1e6,
        }
        if S * N >= 1_000_000_000:
            # the same size as a correlated chain (AR(1), phi = 0.9): the cut-off comes tens of lags later than for white
            # noise, so the warp-level early exit of the short-chain kernel saves little there
            x[0].normal_(generator=g)
            for t in range(1, S):
                torch.randn(N, device=dev, generator=g, out=gr[0])
                x[t] = 0.9 * x[t - 1] + 0.436 * gr[0]
            row["ess_ms_ar1_phi0.9"] = timed(lambda: ops.ess(x, 0.0))
            row["ess_hbm_gbs_ar1_phi0.9"] = (4 * S * N + 4 * N) / row["ess_ms_ar1_phi0.9"] / 1e6
            row["ess_ms_no_cutoff"] = timed(lambda: ops.ess(x, None))
        if a.cpu and S * N <= 5_000_000:
            from oracle import diagnostic as D

            xn, gn = x.cpu().numpy(), gr.cpu().numpy()
            t0 = time.perf_counter()
            D.kernel_stein_discrepancy_imq(xn, gn)
            row["ksd_cpu_oracle_ms"] = (time.perf_counter() - t0) * 1e3
            t0 = time.perf_counter()
            D.effective_sample_size(xn)
            row["ess_cpu_oracle_fft_ms"] = (time.perf_counter() - t0) * 1e3
        rows.append(row)
        del x, gr
    print(json.dumps(rows))


if __name__ == "__main__":
    main()
