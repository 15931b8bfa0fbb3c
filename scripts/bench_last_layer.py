#!/usr/bin/env python
"""K2 measurement: S-batched last-layer GEMM (tcgen05) at the c4 shape, TFLOP/s against the measured
cuBLAS bf16 peak of MEASURED_PEAKS.json, with torch.matmul (cuBLAS) timed beside it on the same operands."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fortuna_b200 import ops  # noqa: E402


def timeit(fn, iters, warmup=2):
    for _ in range(warmup):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--S", type=int, default=64)
    ap.add_argument("--B", type=int, default=8192)
    ap.add_argument("--D", type=int, default=2048)
    ap.add_argument("--C", type=int, default=1000)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--out", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--no-cublas", action="store_true")
    ap.add_argument("--path", type=int, default=0, help="0 auto, 2 one-CTA tcgen05 kernel, 3 CTA-pair (cta_group::2) kernel")
    ap.add_argument("--ex", action="store_true", help="also time the row-statistics / softmax epilogue (fb200_last_layer_logits_ex) "
                    "and log_prob from logits + lse against the single-pass reduction kernel")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    g = torch.Generator(device=dev)
    g.manual_seed(3)
    x = torch.empty((a.B, a.D), device=dev, dtype=torch.float32).normal_(generator=g).to(torch.bfloat16)
    wt = (torch.empty((a.S, a.C, a.D), device=dev, dtype=torch.float32).normal_(generator=g) / a.D ** 0.5).to(torch.bfloat16)
    bias = torch.empty((a.S, a.C), device=dev, dtype=torch.float32).normal_(generator=g)
    lt = torch.tensor([0.3], device=dev)
    odt = torch.float32 if a.out == "f32" else torch.bfloat16
    out = torch.empty((a.S, a.B, a.C), device=dev, dtype=odt)
    flops = 2.0 * a.S * a.B * a.D * a.C
    ms = timeit(lambda: ops.last_layer_logits(x, wt, bias, lt, w_layout=ops.W_SCD, path=a.path, out=out), a.iters)
    res = {"shape": {"S": a.S, "B": a.B, "D": a.D, "C": a.C, "out": a.out}, "path": a.path, "ms": ms, "tflops": flops / ms / 1e9,
           "frac_of_burst": flops / ms / 1e9 / peaks.get("bf16_tflops", 1656.1),
           "frac_of_sustained": flops / ms / 1e9 / peaks.get("bf16_tflops_sustained", 1382.0),
           "out_write_gbs": out.numel() * out.element_size() / ms / 1e6}
    if not a.no_cublas:
        o2 = torch.empty((a.S, a.B, a.C), device=dev, dtype=torch.bfloat16)
        wT = wt.transpose(1, 2)  # [S, D, C] view, K-major storage
        ms2 = timeit(lambda: torch.matmul(x.unsqueeze(0), wT, out=o2), a.iters)
        res["cublas_bf16_out_ms"] = ms2
        res["cublas_tflops"] = flops / ms2 / 1e9
    if a.ex:
        def ex(epi, lse):
            return timeit(lambda: ops.last_layer_logits_ex(x, wt, bias, lt, out_dtype=odt, w_layout=ops.W_SCD, epilogue=epi,
                                                           want_lse=lse), a.iters)

        ms1 = timeit(lambda: ops.last_layer_logits(x, wt, bias, lt, w_layout=ops.W_SCD, path=2, out=out), a.iters)
        res["one_cta_logits_ms"] = ms1
        res["one_cta_logits_plus_lse_ms"] = ex(ops.EPI_LOGITS, True)
        if a.C <= 256:
            res["softmax_epilogue_ms"] = ex(ops.EPI_SOFTMAX, False)
            res["log_softmax_epilogue_ms"] = ex(ops.EPI_LOG_SOFTMAX, True)
        z, lse = ops.last_layer_logits_ex(x, wt, bias, lt, out_dtype=odt, w_layout=ops.W_SCD, want_lse=True)
        t = torch.randint(0, a.C, (a.B,), device=dev, dtype=torch.int32, generator=g)
        res["log_prob_from_lse_ms"] = timeit(lambda: ops.log_prob_from_lse(z, lse, t), a.iters)
        res["log_prob_single_pass_reduce_ms"] = timeit(lambda: ops.bma_reduce_cls(z, ("log_prob",), targets=t), a.iters)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
