#!/usr/bin/env python
"""Parity of the S-sharded multi-GPU predictive against the single-GPU kernel on the same samples.
Run under torchrun with N ranks (one per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/check_sharded_predictive.py
Every rank builds the same z[S,B,C]; rank r reduces samples [r*S/N, (r+1)*S/N) and, after the exchange, must hold
the rows [r*B/N, (r+1)*B/N) of what one GPU computes from all S samples (float32 sums in a different order:
rtol 1e-5; mode, conformal masks/sizes and ECE counts equal)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fortuna_b200 import ops  # noqa: E402
from fortuna_b200.prob_model.predictive import pipeline  # noqa: E402


def run_check(dev, rank, world, transport="p2p", verbose=True):
    """Compare the S-sharded path with the single-GPU kernel on z[8 N, 1024 N, 1000]; returns the parity record every
    rank agrees on: {"outputs_mismatching", "mask_flips", "mask_elems", "ece_counts_equal", "ece_equal", "ok"}."""
    S, B, C = 8 * world, 1024 * world, 1000
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    z = torch.empty((S, B, C), device=dev).normal_(0, 4, generator=g)
    t = torch.randint(0, C, (B,), device=dev, generator=g, dtype=torch.int32)
    want = ("mean", "aleatoric_variance", "epistemic_variance", "entropy", "aleatoric_entropy", "epistemic_entropy",
            "log_prob", "mode")
    full = ops.bma_reduce_cls(z, want, targets=t)
    val = ops.sort_f32_(ops.aps_scores(full["mean"], t).clone())
    calib = pipeline.ConformalCalibration(val, B, 0.1)
    ref = pipeline.PredictivePipeline(S, B, C, dev, want)
    ref.outs = dict(full)
    mask_f, sizes_f, ece_f = ref.score(t, calib)

    per = S // world
    pipe = pipeline.PredictivePipeline(per, B, C, dev, want, world=world, rank=rank, transport=transport)
    zl = z[rank * per : (rank + 1) * per].contiguous()
    bad = 0
    flips = 0
    lo, hi = rank * (B // world), (rank + 1) * (B // world)
    for rep in range(3):  # three steps: both receive buffers of the double-buffered exchange and the wrap-around
        outs = pipe.reduce(zl, t)
        mask, sizes, ece = pipe.score(t, calib)
        torch.cuda.synchronize()
        for k in want:
            a, b = outs[k].cpu().numpy(), full[k][lo:hi].cpu().numpy()
            if k == "mode":
                ok = np.array_equal(a, b)
            else:
                ok = np.allclose(a, b, rtol=1e-5, atol=2e-6 if a.ndim == 1 else 2e-7)
            if not ok:
                bad += 1
                if verbose:
                    print(f"rank {rank}: MISMATCH in {k} (step {rep}): max abs diff {np.abs(a - b).max()}")
        # scores are prefix sums of means that differ in the last ulp: masks may differ only where a score sits within
        # 1e-6 of the threshold value; count such flips instead of demanding equality
        flips = max(flips, int((mask.cpu().numpy() != mask_f[lo:hi].cpu().numpy()).sum()))
    ok_counts = np.array_equal(ece["counts_global"].cpu().numpy(), ece_f["counts"].cpu().numpy())
    ece_ok = np.allclose(ece["result_global"].cpu().numpy(), ece_f["result"][:2].cpu().numpy(), rtol=1e-5, atol=1e-7)
    if verbose:
        print(f"rank {rank}: outputs mismatching: {bad}; conformal mask flips vs single GPU: {flips} of {mask.numel()}; "
              f"global ECE counts equal: {ok_counts}; global ECE/MCE {ece['result_global'].cpu().numpy()} vs single GPU "
              f"{ece_f['result'][:2].cpu().numpy()}: {ece_ok}")
    tot = torch.tensor([bad, flips, 0 if ok_counts else 1, 0 if ece_ok else 1], device=dev, dtype=torch.int64)
    dist.all_reduce(tot)
    rec = {"config": f"z[{S},{B},{C}] f32, S-sharded x{world} ({pipe.transport}) vs one GPU on all samples, 3 steps",
           "outputs_mismatching": int(tot[0]), "mask_flips": int(tot[1]), "mask_elems": B * C,
           "ece_counts_equal": int(tot[2]) == 0, "ece_equal": int(tot[3]) == 0}
    rec["ok"] = rec["outputs_mismatching"] == 0 and rec["ece_counts_equal"] and rec["ece_equal"] and rec["mask_flips"] <= B * C * 1e-4
    del pipe, z, zl
    return rec


def main():
    transport = sys.argv[1] if len(sys.argv) > 1 else "nccl"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rec = run_check(dev, rank, world, transport)
    dist.destroy_process_group()
    if not rec["ok"]:
        sys.exit(1)
    if rank == 0:
        print(f"sharded predictive parity ok (transport {transport}): {rec}")


if __name__ == "__main__":
    main()
