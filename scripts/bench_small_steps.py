#!/usr/bin/env python
"""Sampler step on the small trees (c1 two-moons MLP, c2 LeNet-5): launch-bound - microseconds per step through the
integrator (host overhead included) and through the bare ops call, CUDA events over many steps."""
import os
import sys
import time

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
for name, tree in (("c1 two-moons MLP", trees.mlp_two_moons_tree(0)), ("c2 LeNet-5", trees.lenet5_tree(0))):
    params = FlatTree.from_tree(tree, device=dev)
    grad = params.clone()
    tx = sghmc_integrator(0.01, None, ops.PRNGKey(0, dev), sched.constant_schedule(1e-5), pc.identity_preconditioner())
    state = tx.init(params)
    for _ in range(20):
        params, state = tx.apply(params, grad, state)
    torch.cuda.synchronize()
    n = 2000
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        params, state = tx.apply(params, grad, state)
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n * 1e6
    print(f"{name}: N={params.n_params}  {e0.elapsed_time(e1) / n * 1e3:.1f} us/step (CUDA events), {wall:.1f} us/step wall")
