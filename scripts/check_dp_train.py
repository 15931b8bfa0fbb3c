#!/usr/bin/env python
"""The user API data-parallel: ``ProbClassifier.train`` under torchrun (one rank per GPU, FitProcessor(devices=-1)) against
the same fit in one process (FitProcessor(devices=0)) on the same batches.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 \
        scripts/check_dp_train.py
Checked: every rank ends with bit-identical parameters and sample bank; the data-parallel chain follows the one-process
chain (same noise stream, gradients summed in a different order: rtol 1e-4 after a short run); the logged loss agrees."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "examples"))
from two_moons_sghmc import MLP, make_moons  # noqa: E402

from fortuna_b200.data import DataLoader  # noqa: E402
from fortuna_b200.prob_model import (CyclicalSGLDPosteriorApproximator, FitConfig, FitOptimizer,  # noqa: E402
                                     FitProcessor, ProbClassifier, SGHMCPosteriorApproximator)


def fit(devices, seed=0, cyclical=False):
    torch.manual_seed(seed)  # same initial weights on every rank and in both fits
    train = make_moons(512, 0.07, 0)
    loader = DataLoader.from_array_data(train, batch_size=128, shuffle=False)
    approx = (CyclicalSGLDPosteriorApproximator(n_samples=4, n_thinning=1, cycle_length=8, init_step_size=1e-4,
                                                exploration_ratio=0.5) if cyclical else
              SGHMCPosteriorApproximator(n_samples=6, n_thinning=2, burnin_length=4, step_schedule=1e-4))
    pm = ProbClassifier(model=MLP(), posterior_approximator=approx, seed=seed)
    status = pm.train(loader, fit_config=FitConfig(optimizer=FitOptimizer(n_epochs=4), processor=FitProcessor(devices=devices)))
    torch.cuda.synchronize()
    return pm, status["fit_status"]["loss"]


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    bad = 0
    for cyclical in (False, True):
        bad += compare(rank, dev, cyclical)
    tot = torch.tensor([bad], device=dev)
    dist.all_reduce(tot)
    if rank == 0:
        print("dp_train:", "OK" if tot.item() == 0 else f"{tot.item()} MISMATCHES")
    dist.destroy_process_group()
    sys.exit(1 if tot.item() else 0)


def compare(rank, dev, cyclical):
    bad = 0
    pm_dp, loss_dp = fit(-1, cyclical=cyclical)
    assert pm_dp.posterior.bound.dp is not None, "the fit did not run data-parallel"
    flat = pm_dp.posterior.bound.params.flat
    bank = pm_dp.posterior.state.bank
    for name, t in (("params", flat), ("bank", bank)):
        r0 = t.clone()
        dist.broadcast(r0, 0)
        if not torch.equal(r0, t):
            bad += 1
            print(f"rank {rank}: {name} differ from rank 0's")
    pm_1, loss_1 = fit(0, cyclical=cyclical)
    assert pm_1.posterior.bound.dp is None
    a, b = flat.cpu().numpy(), pm_1.posterior.bound.params.flat.cpu().numpy()
    if not np.allclose(a, b, rtol=1e-4, atol=1e-5):
        bad += 1
        print(f"rank {rank}: data-parallel chain left the one-process chain: max abs {np.abs(a - b).max()}")
    if not np.allclose(bank.cpu().numpy(), pm_1.posterior.state.bank.cpu().numpy(), rtol=1e-4, atol=1e-5):
        bad += 1
        print(f"rank {rank}: sample banks differ")
    if not np.allclose(loss_dp, loss_1, rtol=1e-4):
        bad += 1
        print(f"rank {rank}: logged losses differ: {loss_dp} vs {loss_1}")
    if rank == 0:
        print("cyclical SGLD" if cyclical else "SGHMC", "| losses", np.round(loss_dp, 3).tolist(), "| multicast",
              pm_dp.posterior.bound.dp.multicast, "| mismatches", bad)
    return bad


if __name__ == "__main__":
    main()
