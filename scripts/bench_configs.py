#!/usr/bin/env python
"""Per-config kernel timings for the BASELINE.json configs (c1..c5): sampler step on each parameter tree,
last-layer contraction + single-pass reductions at each (S, B, D, C).  One JSON line per measurement."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fortuna_b200 import ops  # noqa: E402
from fortuna_b200.prob_model.posterior.sgmcmc import sgmcmc_preconditioner as pc  # noqa: E402
from fortuna_b200.prob_model.posterior.sgmcmc import sgmcmc_step_schedule as sched  # noqa: E402
from fortuna_b200.prob_model.posterior.sgmcmc.sghmc.sghmc_integrator import sghmc_integrator  # noqa: E402
from fortuna_b200.utils.flat_tree import FlatTree  # noqa: E402
from tests import trees  # noqa: E402

dev = torch.device("cuda:0")
HBM = 6541.5
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    HBM = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])


def timeit(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def sampler(name, tree):
    params = FlatTree.from_tree(tree, device=dev) if not isinstance(tree, FlatTree) else tree
    grad = params.empty_like()
    grad.flat.normal_(0, 0.01)
    tx = sghmc_integrator(0.01, None, ops.PRNGKey(0, dev), sched.constant_schedule(4e-6), pc.identity_preconditioner())
    st = [tx.init(params)]

    def step():
        _, st[0] = tx.apply(params, grad, st[0])

    us = timeit(step, 50)
    n = params.n_params
    print(json.dumps({"config": name, "kernel": "K1 sghmc step (identity)", "n_params": n, "n_leaves": len(params.lens),
                      "us": round(us, 2), "GB/s": round(20 * n / us / 1e3, 1), "hbm_frac": round(20 * n / us / 1e3 / HBM, 4)}))


def predictive(name, S, B, D, C, dtype=torch.float32):
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    x = torch.empty((B, D), device=dev).normal_(generator=g).to(dtype)
    w = (torch.empty((S, C, D), device=dev).normal_(generator=g) / D ** 0.5).to(dtype)
    bias = torch.empty((S, C), device=dev).normal_(generator=g)
    t = torch.randint(0, C, (B,), device=dev, dtype=torch.int32)
    z = torch.empty((S, B, C), device=dev)
    us2 = timeit(lambda: ops.last_layer_logits(x, w, bias, w_layout=ops.W_SCD, out=z), 10)
    want = ("mean", "aleatoric_variance", "epistemic_variance", "entropy", "aleatoric_entropy", "epistemic_entropy", "log_prob", "mode")
    outs = ops.bma_reduce_cls(z, want, targets=t)
    us3 = timeit(lambda: ops.bma_reduce_cls(z, want, targets=t, out=outs), 10)
    byt = 4.0 * S * B * C + 12.0 * B * C + 24.0 * B
    k2_bytes = x.numel() * x.element_size() + w.numel() * w.element_size() + 4.0 * S * B * C
    print(json.dumps({"config": name, "S": S, "B": B, "D": D, "C": C, "in": str(dtype).split(".")[-1],
                      "K2_us": round(us2, 1), "K2_TFLOPs": round(2.0 * S * B * D * C / us2 / 1e6, 1),
                      "K2_GB/s": round(k2_bytes / us2 / 1e3, 1),
                      "K3_us": round(us3, 1), "K3_GB/s": round(byt / us3 / 1e3, 1), "K3_hbm_frac": round(byt / us3 / 1e3 / HBM, 3),
                      "elems_per_s_K2+K3": round(S * B * C / (us2 + us3) * 1e6, 0)}))


if __name__ == "__main__":
    sampler("c1 two-moons MLP", trees.mlp_two_moons_tree(0))
    sampler("c2 LeNet-5", trees.lenet5_tree(0))
    p = FlatTree.from_shapes(trees.bert_base_shapes(), device=dev)
    p.flat.normal_(0, 0.01)
    sampler("c5 BERT-base", p)
    predictive("c1 two-moons grid", 30, 22500, 30, 2)
    predictive("c2 MNIST LeNet-5", 100, 10000, 84, 10)
    predictive("c3 ResNet-18 last layer (32 of 256 samples per GPU)", 32, 10000, 512, 10)
    predictive("c3 ResNet-18 last layer (all 256 samples)", 256, 10000, 512, 10)
    predictive("c5 BERT-base classifier", 64, 65536, 768, 2, torch.bfloat16)
    predictive("c4 slice", 64, 8192, 2048, 1000, torch.bfloat16)
