#!/usr/bin/env python
"""How much of an e2e step (bench_e2e_api's call) is HOST time: wall clock until the final stream synchronize is reached
(everything enqueued) against the wall clock of the whole call.  If the two are close the step is bound by the Python /
ctypes / torch host work per batch, not by the GPU or the copies.  Diagnostic, not a benchmark."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from fortuna_b200 import ops  # noqa: E402
from fortuna_b200.prob_model.predictive import pipeline  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    S, B, D, Cn = 64, 65536, 2048, 1000
    g = torch.Generator(device=dev)
    g.manual_seed(99)
    vl = torch.empty((bench.N_VAL, Cn), device=dev).normal_(0.0, 4.0, generator=g)
    vt = torch.randint(0, Cn, (bench.N_VAL,), device=dev, generator=g, dtype=torch.int32)
    vm = ops.bma_reduce_cls(vl.unsqueeze(0).contiguous(), ("mean",))["mean"]
    calib = pipeline.ConformalCalibration(ops.sort_f32_(ops.aps_scores(vm, vt)), bench.N_VAL, bench.ERROR)
    pm = bench.build_prob_classifier(dev, S, D, Cn, 0, torch.bfloat16)
    Bc = int(os.environ.get("E2E_BATCH", 8192))
    xs = [torch.randn((Bc, D)).to(torch.bfloat16).pin_memory() for _ in range(B // Bc)]
    ts = [torch.randint(0, Cn, (Bc,), dtype=torch.int32).pin_memory() for _ in range(B // Bc)]
    loader = bench.HostBatches(xs, ts)
    key = pm.rng.get()
    host = True
    reached = []
    orig = torch.cuda.Stream.synchronize

    def sync(self):
        reached.append(time.perf_counter())
        return orig(self)

    torch.cuda.Stream.synchronize = sync
    loop_done = []
    orig_ece = ops.ece_from_packed_bins

    def ece_from_packed(*a, **k):  # the first call after the batch loop: every batch has been enqueued by now
        loop_done.append(time.perf_counter())
        return orig_ece(*a, **k)

    ops.ece_from_packed_bins = ece_from_packed
    for it in range(8):
        torch.cuda.synchronize()
        reached.clear()
        t0 = time.perf_counter()
        host = pm.predictive.statistics(loader, bench.WANT, n_posterior_samples=S, rng=key, conformal=calib, ece=True, to_host=host)
        t1 = time.perf_counter()
        if it >= 3:
            print(f"step {it}: total {1e3 * (t1 - t0):6.2f} ms; batch loop enqueued after {1e3 * (loop_done[-1] - t0):6.2f} ms; "
                  f"final synchronize reached after {1e3 * (reached[-1] - t0):6.2f} ms")


if __name__ == "__main__":
    main()
