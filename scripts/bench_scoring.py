"""CUDA-event timings of the scoring kernels (K5-K7) at the c4 shape: mean[B=65536, C=1000] float32.
usage: python scripts/bench_scoring.py [--B 65536] [--C 1000] [--iters 20]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fortuna_b200 import ops  # noqa: E402
from fortuna_b200.prob_model.predictive.pipeline import conformal_threshold, ece_thresholds  # noqa: E402


def timed(fn, iters):
    for _ in range(3):
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
    ap.add_argument("--B", type=int, default=65536)
    ap.add_argument("--C", type=int, default=1000)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    z = 4 * torch.randn((8, a.B, a.C), device=dev, generator=g)
    mean = torch.softmax(z, -1).mean(0).contiguous()
    del z
    t = torch.randint(0, a.C, (a.B,), device=dev, generator=g, dtype=torch.int32)
    out = {"B": a.B, "C": a.C}
    out["aps_all_ms"] = timed(lambda: ops.aps_scores(mean, None), a.iters)
    out["aps_all_perm_ms"] = timed(lambda: ops.aps_scores(mean, None, want_perm=True), a.iters)
    out["aps_target_ms"] = timed(lambda: ops.aps_scores(mean, t), a.iters)
    uni = torch.full_like(mean, 1.0 / a.C)
    out["aps_all_uniform_rows_ms"] = timed(lambda: ops.aps_scores(uni, None), a.iters)
    scores = ops.aps_scores(mean, None)
    val = ops.sort_f32_(ops.aps_scores(mean[:50000], t[:50000]))
    thr = conformal_threshold(val.numel(), 0.05)
    out["conformal_sets_ms"] = timed(lambda: ops.conformal_sets(val, scores, thr), a.iters)
    out["aps_sets_fused_ms"] = timed(lambda: ops.aps_conformal_sets(mean, val, thr), a.iters)
    out["aps_sets_fused_with_scores_ms"] = timed(lambda: ops.aps_conformal_sets(mean, val, thr, want_scores=True), a.iters)
    ethr = torch.from_numpy(ece_thresholds(a.B)).to(dev)
    out["ece_bins_ms"] = timed(lambda: ops.ece_bins(mean, t, ethr), a.iters)
    # conformal Bayes at the c2 shape: 30 draws, 60 000 training points, 10 000 test points x 10 classes
    S, n_train, n_test, Cn = 30, 60000, 10000, 10
    ell = torch.log_softmax(2 * torch.randn((S, n_train, Cn), device=dev, generator=g), -1)[..., 0].contiguous()
    lsm = torch.log_softmax(2 * torch.randn((S, n_test, Cn), device=dev, generator=g), -1).contiguous()
    ms = timed(lambda: ops.conformal_bayes_sets(ell, lsm, 0.05), 5)
    out["conformal_bayes_c2_ms"] = ms
    out["conformal_bayes_c2_fp32_tflops"] = 2.0 * n_test * Cn * n_train * S / ms / 1e9
    print(json.dumps(out))


if __name__ == "__main__":
    main()
