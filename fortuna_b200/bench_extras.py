"""Secondary measurements printed next to the headline: SG-MCMC updates/s on the BERT-base-shaped
flat pytree (BASELINE config 5) with its HBM roofline fraction."""

import torch

from . import ops
from .prob_model.posterior.sgmcmc import sgmcmc_preconditioner as pc
from .prob_model.posterior.sgmcmc import sgmcmc_step_schedule as sched
from .prob_model.posterior.sgmcmc.sghmc.sghmc_integrator import sghmc_integrator
from .utils.flat_tree import FlatTree


def sampler_extras(device, hbm_peak_gbs, steps=20, warmup=3):
    import sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from tests import trees

    out = {}
    params = FlatTree.from_shapes(trees.bert_base_shapes(), device=device)
    params.flat.normal_(0.0, 0.01)
    grad = params.empty_like()
    grad.flat.normal_(0.0, 0.01)
    n = params.n_params
    for name, pre, bytes_per in (("identity", pc.identity_preconditioner(), 20), ("rmsprop", pc.rmsprop_preconditioner(), 28)):
        tx = sghmc_integrator(0.01, None, ops.PRNGKey(0, device), sched.constant_schedule(4e-6), pre)
        state = tx.init(params)
        for _ in range(warmup):
            params, state = tx.apply(params, grad, state)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            params, state = tx.apply(params, grad, state)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        gbs = bytes_per * n / (ms * 1e-3) / 1e9
        out[f"sghmc_{name}"] = {
            "n_params": n, "n_leaves": len(params.lens), "ms_per_step": ms,
            "param_updates_per_s": n / (ms * 1e-3), "bytes_per_param": bytes_per,
            "achieved_gbs": gbs, "hbm_frac": gbs / hbm_peak_gbs,
        }
        del state
    return out
