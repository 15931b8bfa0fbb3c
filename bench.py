#!/usr/bin/env python
"""Benchmark of the BMA-predictive hot path (BASELINE.json config 4) on B200.

One "step" = one pass of the predictive path over one batch of synthetic calibrated logits
z[S=64, B=65536, C=1000] (float32, 16.8 GB, larger than L2):
    fb200_bma_reduce_cls      -> mean, aleatoric/epistemic variance, entropy + split, log_prob, mode
                                 (+ per-row confidence and Brier term for the ECE)
    fb200_aps_conformal_sets  -> adaptive-prediction-set scores for every class of the mean and the rank-count sets
                                 against 50 000 pre-sorted validation scores, one kernel
    fb200_ece_bins_from_rows  -> ECE / MCE / accuracy / Brier of the mean from those per-row quantities
`value` times that with z resident in HBM; `e2e` times the user's call -
`ProbClassifier.predictive.statistics(data_loader, ..., conformal=..., ece=True, to_host=True)` - from pinned HOST
features [B, 2048] through the last-layer GEMM (K2, tcgen05), the reductions, the sets and the ECE back to pinned host
results (H2D of every batch and D2H of every result inside the timed region).  `--impl reference` times the CPU oracle (numpy restatement of the
reference's jnp path - JAX itself cannot be installed here) on a bounded sample with all host cores.

Contract: python bench.py --gpus N --steps K --warmup W  (N > 1 under torchrun, one rank per GPU).
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "posterior_predictive_SBC_elems_per_s"
UNIT = "elems/s"
S_FULL, B_FULL, C_FULL = 64, 65536, 1000
N_VAL = 50_000
ERROR = 0.05


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _tensor_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p.get("bf16_tflops", 1590.0)), float(p.get("bf16_tflops_sustained", 1400.0))
    return 1590.0, 1400.0


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: NVML polled every few milliseconds from a thread
    (nvidia_ml_py), nvidia-smi -lms 100 as the fallback (one sample per 100 ms - too coarse for a 70 ms region)."""

    FIELDS = (
        "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
        "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    )
    NVML_PERIOD_S = 0.004

    def __init__(self, index=0):
        self.lines = []  # (wall time, csv line)
        self.proc = None
        self.index = index
        self.t0 = self.t1 = None
        self.nvml = None
        self._stop = threading.Event()

    def _nvml_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._nvml_index())
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            pynvml.nvmlDeviceGetClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        n = self.nvml
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))
        while not self._stop.is_set():
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                r = int(get_reasons(self.handle))
                flags = ["Active" if r & b else "Not Active" for _, b in bits]
                self.lines.append((time.time(), ",".join([str(self.index), str(sm), str(self.max_sm), "0", hex(r)] + flags)))
            except Exception:
                pass
            self._stop.wait(self.NVML_PERIOD_S)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self.thread.join(timeout=2)
        elif self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        else:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        inside = [ln for (t, ln) in self.lines if self.t0 is not None and self.t0 <= t <= (self.t1 or t)]
        window = "timed region"
        if not inside:  # timed region shorter than the sampling period: fall back to warm-up + timed samples
            inside = [ln for (_, ln) in self.lines[1:]] or [ln for (_, ln) in self.lines]
            window = "warm-up + timed region"
        for ln in inside:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "window": window, "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# --------------------------------------------------------------------------------------------- CPU arm
def _cpu_worker(args):
    z, t, val_sorted, thr = args
    from oracle import predictive as P
    from oracle import scoring as Sc
    import numpy as np

    mean = P.cls_mean(z)
    P.cls_aleatoric_variance(z)
    P.cls_epistemic_variance(z)
    P.cls_entropy(z)
    P.cls_aleatoric_entropy(z)
    P.cls_epistemic_entropy(z)
    P.cls_log_prob(z, t)
    P.cls_mode(z)
    scores = Sc.aps_cumsums(mean)
    counts = (len(val_sorted) - np.searchsorted(val_sorted, scores, side="right")).astype(np.float32)
    (counts < thr).sum(1)
    Sc.compute_counts_confs_accs(mean.argmax(-1), mean, t)
    return z.size


_FAITHFUL_NOTE = None


def cpu_reference_throughput(rows_per_core, steps, warmup, cores=None):
    """Times the oracle (fair variant: no O(C^2) entropy loop) on [S, rows_per_core*cores, C] per step,
    the batch split over `cores` forked workers.  Returns (elems/s, ms_per_step, cores, sample text)."""
    import multiprocessing as mp

    import numpy as np

    from oracle import predictive as P
    from oracle import scoring as Sc

    if _FULL_AFFINITY is not None:
        try:
            os.sched_setaffinity(0, _FULL_AFFINITY)  # the GPU arm may have bound this process to one NUMA node
        except OSError:
            pass
    cores = cores or len(os.sched_getaffinity(0)) or os.cpu_count() or 1
    r = np.random.default_rng(0)
    shards = []
    # validation scores as in the GPU arm: APS scores of BMA mean rows against random targets (built once, outside the timed
    # steps; 4 samples per row instead of 64 to keep the set-up short - argsort + searchsorted cost the same whatever the values)
    vmean = np.zeros((N_VAL, C_FULL), dtype=np.float32)
    for _ in range(4):
        vmean += P.softmax((4 * r.standard_normal((N_VAL, C_FULL))).astype(np.float32))
    val = np.sort(Sc.aps_score(vmean / np.float32(4), r.integers(0, C_FULL, N_VAL)))
    del vmean
    thr = Sc.conformal_threshold(len(val), ERROR)
    for _ in range(cores):
        z = (4 * r.standard_normal((S_FULL, rows_per_core, C_FULL))).astype(np.float32)
        t = r.integers(0, C_FULL, rows_per_core).astype(np.int32)
        shards.append((z, t, val, thr))
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(cores) as pool:
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            done = pool.map(_cpu_worker, shards)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    # the reference's literal entropy evaluates the full softmax once per class (predictive/classification.py:300-310:
    # O(S B C^2)) and its conformal sets broadcast [n_val, B, C] (conformal/classification/base.py:51-53); the timed "fair"
    # variant above replaces both by their O(S B C) / searchsorted equivalents.  One small faithful sample for scale:
    faithful = None
    try:
        zf = shards[0][0][:8, :4]
        t0 = time.perf_counter()
        P.cls_entropy_faithful(zf, "entropy")
        tf = time.perf_counter() - t0
        t0 = time.perf_counter()
        P.cls_entropy(zf)
        tq = time.perf_counter() - t0
        faithful = {"entropy_faithful_s": tf, "entropy_fair_s": tq, "sample": "z[8,4,1000]", "slowdown": tf / max(tq, 1e-9)}
    except Exception:
        pass
    global _FAITHFUL_NOTE
    _FAITHFUL_NOTE = faithful
    elems = S_FULL * rows_per_core * cores * C_FULL
    total = sum(times)
    sample = f"oracle (numpy) on z[{S_FULL},{rows_per_core * cores},{C_FULL}] f32 per step, {cores} forked workers"
    return elems * len(times) / total, 1e3 * total / len(times), cores, sample


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    v, ms, cores, sample = cpu_reference_throughput(args.cpu_rows_per_core, args.steps, args.warmup)
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": v,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "c4 predictive z[S=64,B,C=1000] -> BMA reductions + APS conformal sets + ECE (bounded CPU sample)",
                   "S": S_FULL, "C": C_FULL},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "variant": "fair (O(S B C) entropies, searchsorted rank counts); n_val = %d" % N_VAL,
                         "faithful_variant": _FAITHFUL_NOTE},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU restatement of the reference's jnp path (oracle/), not JAX: jax/flax/optax are not installable here",
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- GPU arm
WANT = ("mean", "aleatoric_variance", "epistemic_variance", "entropy", "aleatoric_entropy", "epistemic_entropy",
        "log_prob", "mode")


try:
    _FULL_AFFINITY = set(os.sched_getaffinity(0))
except (AttributeError, OSError):
    _FULL_AFFINITY = None


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank's CPU threads to the cores NVML reports as closest to its GPU, before any pinned host buffer is
    allocated: first-touch then places the staging buffers on the GPU's NUMA node (at N = 8 every rank streams 17 GB per
    step from host memory; cross-socket buffers halve that).  Returns the number of cores, or None if unavailable."""
    try:
        import pynvml

        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = local_rank
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if local_rank < len(ids) and ids[local_rank].isdigit():
                idx = int(ids[local_rank])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return None


class HostBatches:
    """A loader of pinned HOST batches: (features [Bc, D] bf16, targets [Bc] int32) pairs - what a data loader hands
    ``predictive.statistics`` (fortuna/data/loader/base.py); iterating yields the same pinned tensors every pass."""

    def __init__(self, xs, ts):
        self.xs, self.ts = xs, ts
        self.size = sum(int(x.shape[0]) for x in xs)

    def __iter__(self):
        return iter(zip(self.xs, self.ts))

    def to_array_targets(self):  # marks this as a DATA loader (targets are read from the batches)
        import torch

        return torch.cat(self.ts)


def build_prob_classifier(dev, S, D, Cn, seed, logits_dtype):
    """ProbClassifier over a last-layer posterior of S random-init samples: ``dfe_subnet`` = identity (the features ARE the
    inputs: the backbone is outside the hot path), ``output_subnet`` = Linear(D, C); bf16 contraction on tcgen05."""
    import torch

    from fortuna_b200.prob_model import ProbClassifier, SGHMCPosteriorApproximator

    class Head(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.dfe_subnet = torch.nn.Identity()
            self.output_subnet = torch.nn.Linear(D, Cn)

        def forward(self, x):
            return self.output_subnet(self.dfe_subnet(x).float())

    pm = ProbClassifier(model=Head(), posterior_approximator=SGHMCPosteriorApproximator(n_samples=S), seed=seed, device=dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(11 + seed)
    n_tr = Cn * D + Cn  # flat layout (sorted leaf names): output_subnet/bias [C], output_subnet/weight [C, D]
    bank = torch.empty((S, n_tr), device=dev)
    bank[:, :Cn].normal_(generator=gen)
    bank[:, Cn:].normal_(generator=gen).mul_(D ** -0.5)
    pm.attach_posterior_samples(bank, freeze_fun=lambda path, v: "trainable" if "output_subnet" in path else "frozen",
                                last_layer_dtype=torch.bfloat16, logits_dtype=logits_dtype)
    pm.predictive.log_temp = torch.tensor([0.3], device=dev)
    return pm


def bench_e2e_api(dev, S, B, D, Cn, calib, args, world, rank):
    """`e2e`: the metric through the public call a user makes -
    ``ProbClassifier.predictive.statistics(data_loader, names, n_posterior_samples=S, conformal=..., ece=True,
    to_host=True)`` - with pinned HOST inputs: every step copies the step's features and targets to the device and all
    results back inside the timed region.  N > 1: ``sharded_io`` - every rank feeds its own B / N rows of each batch and
    receives the results of those rows; the S draws are sharded over the ranks (strong scaling: S, B fixed)."""
    import torch
    import torch.distributed as dist

    logits_dtype = torch.bfloat16 if args.e2e_logits == "bf16" else torch.float32
    pm = build_prob_classifier(dev, S, D, Cn, 0, logits_dtype)
    B_local = B // world
    # rows per batch per rank: the whole job's batch at N = 1; at N > 1 at least 4096 rows per rank (a smaller batch leaves
    # the GEMM launch-sized), so fewer batches per step as N grows
    Bc = min(args.e2e_batch if world == 1 else max(4096, args.e2e_batch // world), B_local)
    n_batches = (B_local + Bc - 1) // Bc
    gen = torch.Generator(device=dev)
    gen.manual_seed(21 + rank)
    xs, ts = [], []
    for i in range(n_batches):
        rows = min(Bc, B_local - i * Bc)
        d = torch.empty((rows, D), device=dev).normal_(generator=gen).to(torch.bfloat16)
        hx = torch.empty((rows, D), dtype=torch.bfloat16, pin_memory=True)
        hx.copy_(d)
        xs.append(hx)
        ts.append(torch.randint(0, Cn, (rows,), dtype=torch.int32).pin_memory())
        del d
    loader = HostBatches(xs, ts)
    key = pm.rng.get()
    host = True
    times = []
    for it in range(2 + args.e2e_steps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        host = pm.predictive.statistics(loader, WANT, n_posterior_samples=S, rng=key, conformal=calib, ece=True,
                                        to_host=host, distribute=world > 1, sharded_io=world > 1)
        torch.cuda.synchronize()
        if it >= 2:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    if world > 1:
        tt = torch.tensor([sec], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        sec = float(tt[0])
    h2d = sum(x.numel() * x.element_size() for x in xs) + sum(t.numel() * 4 for t in ts)
    d2h = sum(v.numel() * v.element_size() for v in host.values())
    return {
        "value": float(S) * B * Cn / sec,
        "unit": UNIT,
        "ms_per_step": sec * 1e3,
        "h2d_bytes_per_step": int(h2d) * world,
        "d2h_bytes_per_step": int(d2h) * world,
        "steps": len(times),
        "api": "fortuna_b200.prob_model.ProbClassifier.predictive.statistics(data_loader, names, n_posterior_samples=S, "
               "conformal=calib, ece=True, to_host=True" + (", distribute=True, sharded_io=True)" if world > 1 else ")"),
        "what": f"pinned host features [B={B},{D}] bf16 + targets -> last-layer GEMM over S={S} drawn samples of the resident bank "
                f"(tcgen05, {args.e2e_logits} logits) -> single-pass reductions -> APS conformal sets + ECE -> pinned host results; "
                f"{n_batches} batches of {Bc} rows per rank, copies overlapped with compute",
        "sharding": "single GPU" if world == 1 else
                    f"strong scaling: S={S} draws sharded x{world} ({S // world} per rank), every rank feeds and receives B/{world} rows; "
                    "inputs all-gathered over NVLink, partial sums exchanged by peer stores",
        "gemm_flops_per_step": 2.0 * S * B * D * Cn,
    }


def run_gpu_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from fortuna_b200 import ops
    from fortuna_b200.prob_model.predictive import pipeline

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: fortuna_b200 has no CPU path")
    numa_cores = bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    strong = args.scaling == "strong" and world > 1
    bshard = args.scaling == "strong-b" and world > 1  # inputs sharded, samples replicated: no exchange at all
    S_total, B, Cn = args.S, args.B, args.C
    if B % world or (strong and S_total % world):
        raise SystemExit("B (and, for strong scaling, S) must be divisible by the number of GPUs")
    S = S_total // world if strong else S_total  # samples THIS rank evaluates

    parity = None
    if world > 1 and not bshard:
        # the S-sharded path against the single-GPU kernel on the same samples, once, outside the timed region
        from scripts.check_sharded_predictive import run_check

        parity = run_check(dev, rank, world, args.transport, verbose=False)
        if not parity["ok"]:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "error": "S-sharded parity check failed", "parity": parity}), flush=True)
            dist.destroy_process_group()
            raise SystemExit(1)

    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    # synthetic calibrated logits 4*N(0,1): softmax neither flat nor one-hot; generated in slabs
    B_rows = B // world if bshard else B  # rows of the logits THIS rank holds
    z = torch.empty((S, B_rows, Cn), dtype=torch.float32, device=dev)
    for s in range(S):
        z[s].normal_(0.0, 4.0, generator=gen)
    gen.manual_seed(99)  # targets / validation scores identical on every rank
    targets = torch.randint(0, Cn, (B,), device=dev, generator=gen, dtype=torch.int32)
    if bshard:
        targets = targets[rank * B_rows : (rank + 1) * B_rows].contiguous()
    # validation probabilities = BMA means over S_FULL samples, like the test rows (the reference feeds both
    # predictive.mean(val) and predictive.mean(test) to conformal_set); round 1 / early round 2 used single-sample
    # validation rows, whose quantile sits ~670 ranks deep in the flatter test rows (--val-samples 1 reproduces that)
    val_logits = torch.empty((args.val_samples, N_VAL, Cn), dtype=torch.float32, device=dev)
    for s in range(args.val_samples):
        val_logits[s].normal_(0.0, 4.0, generator=gen)
    val_targets = torch.randint(0, Cn, (N_VAL,), device=dev, generator=gen, dtype=torch.int32)
    val_mean = ops.bma_reduce_cls(val_logits, ("mean",))["mean"]
    val_sorted = ops.sort_f32_(ops.aps_scores(val_mean, val_targets))
    del val_logits, val_mean
    calib = pipeline.ConformalCalibration(val_sorted, N_VAL, ERROR)

    if bshard:
        # every rank runs the single-GPU pipeline on ITS B / N rows with all S samples (a last-layer sample bank is a few
        # hundred MB: replicate it); the only traffic between the ranks is the packed ECE bin counters
        from fortuna_b200.prob_model.predictive.sharded import SampleShardExchange

        runner = pipeline.PredictivePipeline(S, B_rows, Cn, dev, WANT)
        runner.thresholds = torch.from_numpy(pipeline.ece_thresholds(B)).to(dev)  # bins of the WHOLE data set
        bins = SampleShardExchange(world, rank)
    else:
        runner = pipeline.PredictivePipeline(S, B, Cn, dev, WANT, world=world, rank=rank, transport=args.transport)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * (args.steps + 1))]

    def one_step(i=None):
        runner.reduce(z, targets, kernel_events=(ev[2 * i], ev[2 * i + 1]) if i is not None else None)
        runner.score(targets, calib)
        if bshard:
            bins.allreduce_bins(runner.ece)
            runner.ece["result_global"] = ops.ece_from_packed_bins(runner.ece["packed"], B)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)  # let nvidia-smi deliver its first sample before the GPU gets busy
    for _ in range(args.warmup):
        one_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    sampler.mark_begin()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for i in range(args.steps):
        one_step(i)
    t_end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = t_start.elapsed_time(t_end)
    reduce_ms = sum(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(args.steps)) / args.steps
    if world > 1:
        tt = torch.tensor([total_ms, reduce_ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms, reduce_ms = float(tt[0]), float(tt[1])
    ms_per_step = total_ms / args.steps
    # weak: every rank its own S samples; strong: S_total = S * world fixed; strong-b: S samples x B / N rows per rank
    elems_per_step = float(S) * B * Cn * (1 if bshard else world)
    value = elems_per_step / (ms_per_step * 1e-3)

    # per-phase breakdown of a step (CUDA events on the launching stream), measured in a second, untimed run of a few steps
    phases = None
    if world > 1 or args.phases:
        runner.timer = pipeline.PhaseTimer()
        for _ in range(min(10, args.steps)):
            one_step()
        torch.cuda.synchronize()
        ph = runner.timer.summary_ms()
        runner.timer = None
        if world > 1:
            keys = sorted(ph)
            tt = torch.tensor([ph[k] for k in keys], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ph = {k: float(v) for k, v in zip(keys, tt)}
        phases = {"ms_max_over_ranks": ph, "sum_ms": sum(ph.values()),
                  "note": "reduce = reduction kernel incl. its peer stores; exchange = the one symmetric-memory barrier (or the "
                          "all-to-all); finalize = sum the N source slots + finalise the owned rows; sets = APS sort + conformal "
                          "sets; ece = binning; ece_allreduce = pack + ONE all-reduce + finalise"}

    # roofline of the dominant kernel: algorithmic bytes per launch / measured launch time
    # logits read once + outputs written once: single GPU -> mean + two variance arrays [B,C] and six [B] vectors;
    # shard mode -> one slot row of 2C+4 floats per input (sum p, sum p^2, three scalars)
    # with the peer-store exchange only the rows this rank owns (1/N of them) land in its own HBM - the rest leaves over
    # NVLink - and it RECEIVES (N-1)/N of a slot buffer from its peers: HBM bytes per rank = logits + one slot buffer
    if world == 1 or bshard:
        alg_bytes = 4.0 * S * B_rows * Cn + 4.0 * 3 * B_rows * Cn + 4.0 * B_rows * 8
    else:
        alg_bytes = 4.0 * S * B * Cn + 4.0 * B * (2 * Cn + 4)
    peak, peak_src = _peaks()
    achieved = alg_bytes / (reduce_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_bytes_per_launch.json")
    if os.path.exists(tpath) and world == 1:
        with open(tpath) as f:
            traffic = json.load(f).get("bma_reduce_cls_kernel")

    line = {
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms_per_step,
        "higher_is_better": True,
        "scaling": "strong" if (strong or bshard) else "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {
            "workload": f"c4 predictive z[S={S * world if strong else S},B={B},C={Cn}] f32 -> BMA reductions + APS conformal sets + ECE",
            "S_per_gpu": S, "S_total": S if bshard else S * world, "B": B, "B_per_gpu": B_rows, "C": Cn, "n_val": N_VAL,
            "val_samples": args.val_samples, "error": ERROR,
            "parallelism": "single GPU" if world == 1 else (f"B-sharded x{world}: every rank holds all S samples and B/{world} inputs; no "
                                                            "exchange (one packed all-reduce of the ECE bin counters)") if bshard else (
                f"S-sharded x{world}: partial sums exchanged by " + ("ONE NCCL all-to-all" if runner.transport == "nccl" else
                "peer stores from the reduction kernel over NVLink (symmetric memory, double-buffered) + ONE barrier")),
            "l2": "inputs larger than L2; no flush needed" if 4.0 * S * B * Cn > 2.5e8 else "inputs re-read from HBM each step",
        },
        "roofline": {
            "bound": "hbm", "kernel": "bma_reduce_cls_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src, "kernel_ms": reduce_ms,
            "algorithmic_bytes_per_launch": alg_bytes,
        },
        "gpu_launches": runner.launches_per_step * args.steps,
        "clocks": clocks,
    }
    if parity is not None:
        line["parity"] = parity
    if phases is not None:
        line["phases"] = phases
    if world > 1 and not bshard:
        # the exchange: every rank ships the slot rows of the inputs it does not own and receives as many
        slot_bytes = 4.0 * B * (2 * Cn + 4)
        link = slot_bytes * (world - 1) / world
        ex_ms = (phases["ms_max_over_ranks"].get("exchange", 0.0) + reduce_ms) if phases else None
        line["nvlink"] = {"bytes_out_per_rank_per_step": link, "bytes_in_per_rank_per_step": link,
                          "note": "peer stores are issued by the reduction kernel's row epilogue and overlap it; the lower bound of "
                                  "reduce + exchange is max(HBM time of the logits, link bytes / 770 GB/s)",
                          "link_bound_ms_at_770GBs": link / 770e9 * 1e3,
                          "reduce_plus_exchange_ms": ex_ms,
                          "achieved_GBs_per_direction": (link / (ex_ms * 1e-3) / 1e9) if ex_ms else None}
    del z, runner
    torch.cuda.empty_cache()

    if not args.no_e2e:
        e2e = bench_e2e_api(dev, S_total, B, args.D, Cn, calib, args, world, rank)
        e2e["host_threads_bound_to_gpu_numa_node"] = numa_cores
        line["e2e"] = e2e
    if rank == 0 and world == 1 and not args.no_extras:
        try:
            from fortuna_b200.bench_extras import last_layer_extras, sampler_extras

            line["extras"] = sampler_extras(dev, peak)
            tfb, tfs = _tensor_peaks()
            line["extras"]["last_layer_gemm"] = last_layer_extras(dev, tfb, tfs)
        except Exception as exc:  # extras never invalidate the headline
            line["extras"] = {"error": repr(exc)}
    if world > 1 and not args.no_extras:
        # SG-MCMC chains: one per GPU, different seeds, NO collective - every rank steps its own
        # BERT-base-shaped chain, the slowest rank sets the time, the aggregate is world x updates/s
        try:
            from fortuna_b200.bench_extras import sampler_extras

            ex = sampler_extras(dev, peak, seed=rank, only=("identity",))["sghmc_identity"]
            tt = torch.tensor([ex["ms_per_step"]], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt[0])
            line["extras"] = {"sghmc_chains": {
                "chains": world, "n_params_per_chain": ex["n_params"], "ms_per_step_max_over_ranks": ms,
                "param_updates_per_s": world * ex["n_params"] / (ms * 1e-3), "collectives": 0,
                "hbm_frac_per_gpu": 20 * ex["n_params"] / (ms * 1e-3) / 1e9 / peak}}
        except Exception as exc:
            line["extras"] = {"error": repr(exc)}
        # the other way to use N GPUs for sampling: ONE data-parallel chain, gradient average + step + parameter broadcast
        # fused into one kernel per rank over NVLink peer memory
        try:
            from fortuna_b200.bench_extras import dp_step_extras

            line["extras"]["sghmc_data_parallel_step"] = dp_step_extras(dev, rank, world)
        except Exception as exc:
            line["extras"]["sghmc_data_parallel_step"] = {"error": repr(exc)}
    if rank == 0 and world == 1 and not args.no_cpu:
        v, ms, cores, sample = cpu_reference_throughput(args.cpu_rows_per_core, 3, 1)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                                "variant": "fair (O(S B C) entropies, searchsorted rank counts); n_val = %d" % N_VAL,
                                "faithful_variant": _FAITHFUL_NOTE}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30,
                    help="timed steps; beyond ~50 back-to-back steps the board reaches its software power cap and the SM clock "
                         "drops (reported in `clocks`): 20 steps 3.52 ms/step at 1965 MHz, 100 steps 3.62 at 1875, 300 steps 3.70 at 1792")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--S", type=int, default=S_FULL)
    ap.add_argument("--B", type=int, default=B_FULL)
    ap.add_argument("--C", type=int, default=C_FULL)
    ap.add_argument("--D", type=int, default=2048, help="last-layer feature width for e2e_from_features / GEMM extras")
    ap.add_argument("--cpu-rows-per-core", type=int, default=64)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-batch", type=int, default=8192, help="rows per batch of the e2e loader (whole job; / N per rank)")
    ap.add_argument("--e2e-logits", default="bf16", choices=["bf16", "f32"],
                    help="dtype of the logits between the last-layer GEMM and the reductions in the e2e path")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong", "strong-b"],
                    help="N > 1: weak = every rank evaluates S samples (S_total = N S); strong = S_total = S fixed, S / N per rank "
                         "(S-sharded, exchange of partial sums); strong-b = S and B fixed, B / N inputs per rank with the samples "
                         "replicated (no exchange)")
    ap.add_argument("--val-samples", type=int, default=S_FULL,
                    help="posterior samples behind the validation rows of the conformal calibration (default: as many as the test rows)")
    ap.add_argument("--phases", action="store_true", help="add the per-phase CUDA-event breakdown at N = 1 too")
    ap.add_argument("--transport", default="p2p", choices=["nccl", "p2p"],
                    help="N > 1 exchange: one NCCL all-to-all, or peer stores from the reduction kernel over NVLink")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
