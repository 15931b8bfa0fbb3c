#!/usr/bin/env python
"""Parity and timing of the data-parallel fused sampler step (fb200_sgmcmc_step_dp_f32) under torchrun, one rank per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
        scripts/check_dp_step.py [--bench]
Parity: every rank holds a different gradient; after K fused steps every replica must be BIT-identical to K single-GPU
steps (fb200_sgmcmc_step_f32) on the rank-ordered average of the gradients, for SGHMC and RMSprop-preconditioned
SGHMC, with and without the fused log-joint gradient, on a ragged tree.
--bench: the BERT-base-sized tree (109.5 M parameters): fused step vs NCCL all-reduce + replicated step, CUDA events,
max over ranks."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fortuna_b200 import ops  # noqa: E402
from fortuna_b200.prob_model.posterior.sgmcmc.dp_step import DataParallelStep  # noqa: E402


def _tree(lens):
    offs, o = [], 0
    for l in lens:
        offs.append(o)
        o += (l + 3) // 4 * 4  # leaves start 16-byte aligned like the flat tree
    return offs, o


def parity(dev, rank, world, multicast):
    """multicast: sum the gradients inside the NVSwitch (multimem.ld_reduce) - the sum order is then the switch's, so for
    more than two ranks the comparison with the rank-ordered average is within rounding, not bitwise; the replicas must
    still be bit-identical to each other."""
    lens = [4096 * 3, 7, 1025, 64 * 1024 + 3, 2048, 1, 333]
    offs, n = _tree(lens)
    bad = 0
    exact = (not multicast) or world <= 2
    for rms in (False, True):
        for joint in (None, (12.5, 0.75)):
            dp = DataParallelStep(lens, offs, n, dev, world, rank, rmsprop=rms, seed=3, multicast=multicast)
            if multicast and not dp.multicast:
                if rank == 0:
                    print("multicast (NVLS) addresses not available on this box")
                return -1
            g = torch.Generator(device=dev)
            g.manual_seed(11)  # same initial parameters everywhere
            p0 = torch.empty(n, device=dev).normal_(0, 1, generator=g)
            dp.params.copy_(p0)
            # single-GPU reference state on this rank
            ref_p, ref_m = p0.clone(), torch.zeros(n, device=dev)
            ref_v = torch.zeros(n, device=dev) if rms else None
            key, key2 = ops.PRNGKey(3, dev), torch.empty(2, dtype=ops.KEY_DTYPE, device=dev)
            lk = torch.empty((len(lens), 2), dtype=ops.KEY_DTYPE, device=dev)
            hp = ops.make_hparams(1e-3, 0.05, rmsprop=rms)
            for it in range(3):
                grads = []
                for r in range(world):  # every rank can rebuild every rank's gradient (seeded), for the reference
                    gr = torch.Generator(device=dev)
                    gr.manual_seed(1000 * it + r)
                    grads.append(torch.empty(n, device=dev).normal_(0, 1, generator=gr))
                dp.grad.copy_(grads[rank])
                dp.step(hp, joint=joint)
                acc = grads[0].clone()
                for r in range(1, world):
                    acc += grads[r]
                acc *= 1.0 / world
                ops.sgmcmc_step(ref_p, acc, ref_m, ref_v, None, dp.leaves, dp.tiles, key, key2, lk, hp,
                                joint=None if joint is None else (joint[0], joint[1], None))
                key, key2 = key2, key
                torch.cuda.synchronize()
                if exact:
                    same = torch.equal(dp.params, ref_p)
                else:
                    # RMSprop divides by eps + sqrt(v): where the averaged gradient is ~0 the update is ill-conditioned in
                    # the gradient's last ulp, so a handful of elements may move visibly under another sum order
                    off = ((dp.params - ref_p).abs() > 2e-6 + 2e-5 * ref_p.abs()).sum().item()
                    same = off <= (n // 2000 if rms else 0)
                if not same:
                    bad += 1
                    d = (dp.params - ref_p).abs().max().item()
                    print(f"rank {rank}: rms={rms} joint={joint} mc={multicast} step {it}: params differ, max abs {d}")
                r0 = dp.params.clone()
                dist.broadcast(r0, 0)
                if not torch.equal(r0, dp.params):
                    bad += 1
                    print(f"rank {rank}: replica differs from rank 0's (mc={multicast}, step {it})")
                if not exact:
                    ref_p.copy_(dp.params)  # follow the fused chain so that rounding does not accumulate in the check
            # exploration phase of cyclical SGLD (no noise): same comparison against fb200_sgd_explore_step_f32
            for it in range(2):
                grads = []
                for r in range(world):
                    gr = torch.Generator(device=dev)
                    gr.manual_seed(5000 + 10 * it + r)
                    grads.append(torch.empty(n, device=dev).normal_(0, 1, generator=gr))
                dp.grad.copy_(grads[rank])
                before = dp.params.clone()
                dp.explore_step(hp, joint=joint)
                acc = grads[0].clone()
                for r in range(1, world):
                    acc += grads[r]
                acc *= 1.0 / world
                ex_p = before.clone()
                ex_v = dp.precond_v.clone() if rms else None  # the owned slice already advanced; compare params only
                ops.sgd_explore_step(ex_p, acc, torch.zeros(n, device=dev) if rms else None, None, hp,
                                     joint=None if joint is None else (joint[0], joint[1], None))
                torch.cuda.synchronize()
                if not rms:  # with RMSprop the reference state above is not the chain's; the no-preconditioner case is exact
                    same = torch.equal(dp.params, ex_p) if exact else torch.allclose(dp.params, ex_p, rtol=2e-5, atol=2e-6)
                    if not same:
                        bad += 1
                        print(f"rank {rank}: explore joint={joint} mc={multicast} step {it}: params differ, "
                              f"max abs {(dp.params - ex_p).abs().max().item()}")
                r0 = dp.params.clone()
                dist.broadcast(r0, 0)
                if not torch.equal(r0, dp.params):
                    bad += 1
                    print(f"rank {rank}: explore: replica differs from rank 0's (mc={multicast})")
            t0, t1 = dp.tile_range
            # the owned slice of the optimizer state matches; the rest was never touched
            own = torch.zeros(n, dtype=torch.bool, device=dev)
            tl = dp.tiles.cpu().numpy()
            lv = dp.leaves.cpu().numpy()
            for t in range(t0, t1):
                leaf, ps = int(tl[t, 0]), int(tl[t, 1])
                off, ln = int(lv[leaf, 0]), int(lv[leaf, 1])
                h = (ln + 1) // 2
                e = min(ps + 1024, h)
                own[off + ps : off + e] = True
                own[off + h + ps : off + min(h + e, ln)] = True
            m_ok = torch.equal(dp.momentum[own], ref_m[own]) if exact else (
                ((dp.momentum[own] - ref_m[own]).abs() > 1e-5 + 1e-4 * ref_m[own].abs()).sum().item() <= (n // 2000 if rms else 0))
            if not m_ok or bool((dp.momentum[~own] != 0).any()):
                bad += 1
                print(f"rank {rank}: rms={rms} joint={joint}: momentum slice mismatch")
            if not torch.equal(dp.key, key):
                bad += 1
                print(f"rank {rank}: key chain differs")
            del dp
    return bad


def phase_switch(dev, rank, world):
    """Cyclical SGLD with the RMSprop preconditioner, data-parallel: sampling steps (tile split) and exploration steps
    (element split) interleaved - sampling x2, exploration x2, sampling x2, exploration x1 - against the single-GPU chain
    on the averaged gradients.  The preconditioner moments are rank-local, so every phase change has to hand them over
    (round-1 bug: the new owner saw moments that had missed the other phase's updates - zero in the first cycle)."""
    lens = [4096 * 3, 7, 1025, 64 * 1024 + 3, 2048, 1, 333]
    offs, n = _tree(lens)
    dp = DataParallelStep(lens, offs, n, dev, world, rank, rmsprop=True, seed=5, multicast=False)
    g = torch.Generator(device=dev)
    g.manual_seed(13)
    p0 = torch.empty(n, device=dev).normal_(0, 1, generator=g)
    dp.params.copy_(p0)
    ref_p, ref_m, ref_v = p0.clone(), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    key, key2 = ops.PRNGKey(5, dev), torch.empty(2, dtype=ops.KEY_DTYPE, device=dev)
    lk = torch.empty((len(lens), 2), dtype=ops.KEY_DTYPE, device=dev)
    hp = ops.make_hparams(1e-3, 0.0, rmsprop=True)
    bad = 0
    for it, phase in enumerate("ssEEssE"):
        grads = []
        for r in range(world):
            gr = torch.Generator(device=dev)
            gr.manual_seed(7000 + 10 * it + r)
            grads.append(torch.empty(n, device=dev).normal_(0, 1, generator=gr))
        dp.grad.copy_(grads[rank])
        acc = grads[0].clone()
        for r in range(1, world):
            acc += grads[r]
        acc *= 1.0 / world
        if phase == "s":
            dp.step(hp)
            ops.sgmcmc_step(ref_p, acc, ref_m, ref_v, None, dp.leaves, dp.tiles, key, key2, lk, hp)
            key, key2 = key2, key
        else:
            dp.explore_step(hp)
            ops.sgd_explore_step(ref_p, acc, ref_v, None, hp)
        torch.cuda.synchronize()
        # 1 / (eps + sqrt(v)) is ill-conditioned in the last ulp of a near-zero gradient: allow a handful of elements
        off = ((dp.params - ref_p).abs() > 2e-6 + 2e-5 * ref_p.abs()).sum().item()
        if off > n // 2000:
            bad += 1
            print(f"rank {rank}: phase-switch step {it} ({phase}): {off} elements differ, max abs "
                  f"{(dp.params - ref_p).abs().max().item()}")
        ref_p.copy_(dp.params)
    dp.gather_state()
    torch.cuda.synchronize()
    real = torch.zeros(n, dtype=torch.bool, device=dev)  # alignment padding between leaves belongs to no tile
    for ln, o in zip(lens, offs):
        real[o : o + ln] = True
    if not torch.allclose(dp.precond_v[real], ref_v[real], rtol=1e-5, atol=1e-9):
        bad += 1
        print(f"rank {rank}: gathered preconditioner moments differ from the single-GPU chain's")
    return bad


def bench(dev, rank, world, multicast, steps=20):
    lens = [23440896, 393216, 1536] + [2359296 // 4 * 4] * 36 + [589824] * 1 + [768] * 100  # ~109.5 M, BERT-base-like
    offs, n = _tree(lens)
    dp = DataParallelStep(lens, offs, n, dev, world, rank, seed=0, multicast=multicast)
    dp.params.normal_(0, 0.02)
    dp.grad.normal_(0, 1)
    hp = ops.make_hparams(1e-5, 0.05)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

    def timed(fn):
        for _ in range(3):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(steps):
            fn()
        ev[1].record()
        torch.cuda.synchronize()
        ms = torch.tensor([ev[0].elapsed_time(ev[1]) / steps], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    fused = timed(lambda: dp.step(hp))
    # baseline: NCCL all-reduce of the gradient, then the replicated single-GPU step on every rank
    p, m, gbuf = dp.params.clone(), torch.zeros(n, device=dev), dp.grad.clone()
    key, key2 = ops.PRNGKey(0, dev), torch.empty(2, dtype=ops.KEY_DTYPE, device=dev)
    lk = torch.empty((len(lens), 2), dtype=ops.KEY_DTYPE, device=dev)

    def base():
        dist.all_reduce(gbuf, op=dist.ReduceOp.AVG)
        ops.sgmcmc_step(p, gbuf, m, None, None, dp.leaves, dp.tiles, key, key2, lk, hp)

    nccl = timed(base)
    if rank == 0:
        print(json.dumps({"metric": "dp_sampler_step", "n_params": n, "n_gpus": world, "multicast": bool(dp.multicast),
                          "fused_ms": round(fused, 4),
                          "nccl_allreduce_plus_step_ms": round(nccl, 4), "speedup": round(nccl / fused, 3),
                          "param_updates_per_s": n / (fused * 1e-3),
                          "link_bytes_in_per_rank": 4 * (n // world) * (world if dp.multicast else 2 * (world - 1))}))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    fails = 0
    for multicast in (False, True):
        bad = parity(dev, rank, world, multicast)
        if bad < 0:
            continue
        tot = torch.tensor([bad], device=dev)
        dist.all_reduce(tot)
        fails += int(tot.item())
        if rank == 0:
            print(f"dp_step parity (multicast={multicast}):", "OK" if tot.item() == 0 else f"{tot.item()} MISMATCHES")
        if "--bench" in sys.argv and tot.item() == 0:
            bench(dev, rank, world, multicast)
    tot = torch.tensor([phase_switch(dev, rank, world)], device=dev)
    dist.all_reduce(tot)
    fails += int(tot.item())
    if rank == 0:
        print("dp_step RMSprop phase switch (cyclical SGLD):", "OK" if tot.item() == 0 else f"{tot.item()} MISMATCHES")
    dist.destroy_process_group()
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
