"""Secondary measurements printed next to the headline: SG-MCMC updates/s on the BERT-base-shaped
flat pytree (BASELINE config 5) with its HBM roofline fraction, and the S-batched last-layer GEMM (K2) at a
B = 8192 slice of config 4 with its tensor-pipe fraction."""

import torch

from . import ops
from .prob_model.posterior.sgmcmc import sgmcmc_preconditioner as pc
from .prob_model.posterior.sgmcmc import sgmcmc_step_schedule as sched
from .prob_model.posterior.sgmcmc.sghmc.sghmc_integrator import sghmc_integrator
from .utils.flat_tree import FlatTree


def sampler_extras(device, hbm_peak_gbs, steps=20, warmup=3, seed=0, only=None, joint=False):
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
        if only is not None and name not in only:
            continue
        tx = sghmc_integrator(0.01, None, ops.PRNGKey(seed, device), sched.constant_schedule(4e-6), pre)
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
        if joint:
            # fused log-joint mode: likelihood scaling + Gaussian-prior gradient in the prologue, sum(theta^2) returned
            sq = torch.empty(1, dtype=torch.float64, device=device)
            jt = (468.75, 1.0, sq)
            for _ in range(warmup):
                params, state = tx.apply(params, grad, state, joint=jt)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(steps):
                params, state = tx.apply(params, grad, state, joint=jt)
            e1.record()
            torch.cuda.synchronize()
            msj = e0.elapsed_time(e1) / steps
            out[f"sghmc_{name}"]["joint_ms_per_step"] = msj
            out[f"sghmc_{name}"]["joint_hbm_frac"] = bytes_per * n / (msj * 1e-3) / 1e9 / hbm_peak_gbs
        del state
    return out


def last_layer_extras(device, tf_burst, tf_sustained, S=64, B=8192, D=2048, Cn=1000, iters=3):
    """K2 at the c4 shape (B sliced to 8192 so the fp32 logits are 2.1 GB): TFLOP/s of fb200_last_layer_logits on
    the tcgen05 path, bf16 operands, fp32 accumulate, bias + temperature fused, fp32 and bf16 logits."""
    g = torch.Generator(device=device)
    g.manual_seed(3)
    x = torch.empty((B, D), device=device, dtype=torch.float32).normal_(generator=g).to(torch.bfloat16)
    wt = (torch.empty((S, Cn, D), device=device, dtype=torch.float32).normal_(generator=g) / D ** 0.5).to(torch.bfloat16)
    bias = torch.empty((S, Cn), device=device, dtype=torch.float32).normal_(generator=g)
    lt = torch.tensor([0.3], device=device)
    flops = 2.0 * S * B * D * Cn
    out = {}
    for name, dt in (("f32_logits", torch.float32), ("bf16_logits", torch.bfloat16)):
        o = torch.empty((S, B, Cn), device=device, dtype=dt)
        for _ in range(2):
            ops.last_layer_logits(x, wt, bias, lt, w_layout=ops.W_SCD, path=0, out=o)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            ops.last_layer_logits(x, wt, bias, lt, w_layout=ops.W_SCD, path=0, out=o)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        tf = flops / ms / 1e9
        out[name] = {"S": S, "B": B, "D": D, "C": Cn, "ms": ms, "tflops": tf, "frac_of_burst_peak": tf / tf_burst,
                     "frac_of_sustained_peak": tf / tf_sustained}
        del o
    return out
