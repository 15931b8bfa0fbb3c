#!/usr/bin/env python
"""Kernel / copy timeline summary of ONE e2e step (bench_e2e_api's call) under torch.profiler, rank 0 - a diagnostic for
where an e2e step spends its time (not a benchmark: profiler overhead is inside).  Run alone or under torchrun."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from fortuna_b200 import ops  # noqa: E402
from fortuna_b200.prob_model.predictive import pipeline  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    S, B, D, Cn = 64, 65536, 2048, 1000
    g = torch.Generator(device=dev)
    g.manual_seed(99)
    vl = torch.empty((bench.N_VAL, Cn), device=dev).normal_(0.0, 4.0, generator=g)
    vt = torch.randint(0, Cn, (bench.N_VAL,), device=dev, generator=g, dtype=torch.int32)
    vm = ops.bma_reduce_cls(vl.unsqueeze(0).contiguous(), ("mean",))["mean"]
    calib = pipeline.ConformalCalibration(ops.sort_f32_(ops.aps_scores(vm, vt)), bench.N_VAL, bench.ERROR)
    pm = bench.build_prob_classifier(dev, S, D, Cn, 0, torch.bfloat16)
    Bl = B // world
    Bc = int(os.environ.get("E2E_BATCH", 8192)) // world
    xs = [torch.randn((Bc, D)).to(torch.bfloat16).pin_memory() for _ in range(Bl // Bc)]
    ts = [torch.randint(0, Cn, (Bc,), dtype=torch.int32).pin_memory() for _ in range(Bl // Bc)]
    loader = bench.HostBatches(xs, ts)
    key = pm.rng.get()
    host = True

    def step():
        return pm.predictive.statistics(loader, bench.WANT, n_posterior_samples=S, rng=key, conformal=calib, ece=True,
                                        to_host=host, distribute=world > 1, sharded_io=world > 1)

    for _ in range(3):
        host = step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        host = step()
        torch.cuda.synchronize()
    if rank == 0:
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        t0 = min(e.time_range.start for e in evs)
        t1 = max(e.time_range.end for e in evs)
        print(f"GPU span of the step: {(t1 - t0) / 1e3:.2f} ms; {len(evs)} device activities")
        agg = {}
        for e in evs:
            k = e.name[:70]
            a = agg.setdefault(k, [0, 0.0])
            a[0] += 1
            a[1] += (e.time_range.end - e.time_range.start) / 1e3
        for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
            print(f"{ms:9.3f} ms  x{n:4d}  {k}")
        # timeline of the first two batches
        evs.sort(key=lambda e: e.time_range.start)
        for e in evs[:60]:
            print(f"  +{(e.time_range.start - t0) / 1e3:8.3f} ms  {(e.time_range.end - e.time_range.start) / 1e3:8.3f} ms  {e.name[:60]}")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
