"""Secondary measurements printed next to the headline: SG-MCMC updates/s on the BERT-base-shaped
flat pytree (BASELINE config 5) with its HBM roofline fraction, and the S-batched last-layer GEMM (K2) at a
B = 8192 slice of config 4 with its tensor-pipe fraction."""

import torch

from . import ops
from .prob_model.posterior.sgmcmc import sgmcmc_preconditioner as pc
from .prob_model.posterior.sgmcmc import sgmcmc_step_schedule as sched
from .prob_model.posterior.sgmcmc.sghmc.sghmc_integrator import sghmc_integrator
from .utils.flat_tree import FlatTree
from .utils.tree_shapes import bert_base_shapes


def sampler_extras(device, hbm_peak_gbs, steps=20, warmup=3, seed=0, only=None, joint=False):
    out = {}
    params = FlatTree.from_shapes(bert_base_shapes(), device=device)
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


def dp_step_extras(device, rank, world, steps=10, warmup=3):
    """Data-parallel sampler step (fb200_sgmcmc_step_dp_f32: gradient average + step + parameter broadcast in one kernel
    per rank over NVLink) on the BERT-base-sized flat tree, against NCCL all-reduce + the replicated step.  CUDA events,
    max over ranks; every rank must call this (it allocates symmetric memory collectively)."""
    import torch.distributed as dist

    from . import ops
    from .prob_model.posterior.sgmcmc.dp_step import DataParallelStep

    lens = [23440896, 393216, 1536] + [2359296] * 36 + [589824] + [768] * 100
    offs, n = [], 0
    for l in lens:
        offs.append(n)
        n += (l + 3) // 4 * 4
    dp = DataParallelStep(lens, offs, n, device, world, rank, seed=0)
    dp.params.normal_(0, 0.02)
    dp.grad.normal_(0, 1)
    hp = ops.make_hparams(1e-5, 0.05)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

    def timed(fn):
        for _ in range(warmup):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(steps):
            fn()
        ev[1].record()
        torch.cuda.synchronize()
        ms = torch.tensor([ev[0].elapsed_time(ev[1]) / steps], device=device)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0])

    fused = timed(lambda: dp.step(hp))
    p, m, g = dp.params.clone(), torch.zeros(n, device=device), dp.grad.clone()
    key, key2 = ops.PRNGKey(0, device), torch.empty(2, dtype=ops.KEY_DTYPE, device=device)
    lk = torch.empty((len(lens), 2), dtype=ops.KEY_DTYPE, device=device)

    def base():
        dist.all_reduce(g, op=dist.ReduceOp.AVG)
        ops.sgmcmc_step(p, g, m, None, None, dp.leaves, dp.tiles, key, key2, lk, hp)

    nccl = timed(base)
    return {"n_params": n, "ranks": world, "mode": "nvls multicast" if dp.multicast else "peer loads/stores",
            "fused_ms": fused, "nccl_allreduce_plus_step_ms": nccl, "speedup": nccl / fused,
            "param_updates_per_s": n / (fused * 1e-3),
            "link_bytes_in_per_rank": 4 * (n // world) * (world if dp.multicast else 2 * (world - 1))}
