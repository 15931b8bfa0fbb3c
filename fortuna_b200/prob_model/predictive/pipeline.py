"""Device-side orchestration of the predictive path over calibrated logits z[S,B,C]:
reduce (K3/K4) -> APS scores (K5) -> rank-count conformal sets (K6) -> ECE (K7).

Two callers:
  * ``PredictivePipeline``            - logits already resident in HBM (one GPU, or S-sharded over the
                                        ranks of a torch.distributed NCCL group);
  * ``predictive_from_logits_batches`` - the loader-shaped public call: an iterable of pinned HOST batches
    (like Fortuna's ``_loop_fun_through_inputs_loader``, fortuna/prob_model/predictive/base.py:952-975),
    double-buffered H2D copies overlapped with the kernels, results copied back to pinned host memory.
All arithmetic is in the library kernels; torch provides memory, streams and NCCL only.
"""

import time
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from ... import ops


def conformal_threshold(n_val: int, error: float) -> np.float32:
    """float32((1 - error) * (n_val + 1)) - right-hand side of
    fortuna/conformal/classification/base.py:51-53 under jax weak typing."""
    return np.float32((1 - error) * (n_val + 1))


def ece_thresholds(n_data: int, n_bins: int = 10) -> np.ndarray:
    """jnp.linspace(1 / n_data, 1, n_bins) in float32 (fortuna/metric/classification.py:76)."""
    start, stop = np.float32(1 / n_data), np.float32(1)
    div = n_bins - 1
    step = (np.arange(div, dtype=np.float32) / np.float32(div)).astype(np.float32)
    out = start * (np.float32(1) - step) + stop * step
    return np.concatenate([out, np.array([stop], dtype=np.float32)]).astype(np.float32)


class ConformalCalibration:
    """Validation scores, sorted once on the device, plus the float32 count threshold."""

    def __init__(self, val_scores_sorted: torch.Tensor, n_val: int, error: float):
        self.val_sorted = val_scores_sorted
        self.n_val = int(n_val)
        self.error = float(error)
        self.threshold = conformal_threshold(n_val, error)


def aps_sets(mean: torch.Tensor, calib) -> Tuple[torch.Tensor, torch.Tensor]:
    """APS conformal sets of the rows of `mean`: one fused kernel (scores never materialised) when a row fits the warp
    kernel, else scores + sets."""
    if mean.shape[1] <= 1024:
        return ops.aps_conformal_sets(mean, calib.val_sorted, calib.threshold)
    return ops.conformal_sets(calib.val_sorted, ops.aps_scores(mean, None), calib.threshold)


class PhaseTimer:
    """CUDA events around the phases of a step, on the current stream (optional: the events cost ~2 us each)."""

    def __init__(self):
        self.marks = []  # [(name, event)] of the step being recorded
        self.steps = []  # finished steps

    def mark(self, name):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self.marks.append((name, ev))

    def end_step(self):
        self.steps.append(self.marks)
        self.marks = []

    def summary_ms(self):
        """Mean milliseconds per phase over the recorded steps (call after a synchronize)."""
        acc = {}
        for marks in self.steps:
            for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
                acc.setdefault(n1, []).append(e0.elapsed_time(e1))
        return {k: sum(v) / len(v) for k, v in acc.items()}


class PredictivePipeline:
    def __init__(self, S, B, C, device, want: Sequence[str], world: int = 1, rank: int = 0, transport: Optional[str] = None,
                 fused_sets: bool = False):
        self.S, self.B, self.C = S, B, C
        self.device = device
        self.want = tuple(want)
        # single GPU: the reduction kernel also emits the three per-row quantities the ECE needs (K7 then reads 3 B floats
        # instead of the [B,C] mean)
        self._kernel_want = self.want + tuple(n for n in ("mode", "confidence", "brier_terms") if n not in want) if world == 1 else self.want
        self.world, self.rank = world, rank
        self.B_local = B // world
        self.outs: Dict[str, torch.Tensor] = {}
        self.scores = None
        self.thresholds = torch.from_numpy(ece_thresholds(B)).to(device)
        self.transport = "single"
        self.timer: Optional[PhaseTimer] = None
        self.fused_sets = bool(fused_sets) and world == 1 and C <= 1024
        self._sets_ws = None
        if world > 1:
            from .sharded import PeerSlots, SampleShardExchange

            self.exchange = SampleShardExchange(world, rank)
            rowf = ops.cls_slot_row_floats(C)
            self.transport = transport or "p2p"
            if self.transport == "p2p":
                try:
                    self.peer = PeerSlots(world, rank, self.B_local, rowf, device)
                except Exception as exc:  # no peer mapping on this system (no NVLink / symmetric memory): one all-to-all
                    import warnings

                    warnings.warn(f"peer-memory exchange unavailable ({exc!r}); using the NCCL all-to-all")
                    self.transport = "nccl"
            if self.transport != "p2p":
                self.send = ops.alloc_cls_slots(world, self.B_local, C, device)
                self.recv = torch.empty_like(self.send)
                self.slot_ptrs = ops.slot_pointer_table([self.send[o] for o in range(world)], device)
        # reduce, sort + sets (fused), ece_bins, ece_finalize; sharded: + finalize_slots, pack_bins, ece_from_packed
        fused_sort = C <= 1024
        self.launches_per_step = (4 if world == 1 else 7) + (0 if fused_sort else 1)

    def _mark(self, name):
        if self.timer is not None:
            self.timer.mark(name)

    # -------------------------------------------------------------------------------- K3 / K4
    def reduce(self, z: torch.Tensor, targets: Optional[torch.Tensor], kernel_events=None, calib=None):
        """kernel_events: optional (start, end) CUDA events recorded around the reduction kernel alone.
        calib: with ``fused_sets``, the conformal calibration whose APS sets the same launch emits."""
        self._mark("begin")
        if kernel_events:
            kernel_events[0].record()
        if self.world == 1:
            want = self._kernel_want if targets is not None else self.want
            if self.fused_sets and calib is not None:
                if self._sets_ws is None:
                    self.mask = torch.empty((self.B, self.C), dtype=torch.uint8, device=self.device)
                    self.sizes = torch.empty(self.B, dtype=torch.int32, device=self.device)
                    self._sets_ws = torch.empty(int(ops.lib().fb200_bma_reduce_cls_sets_workspace_bytes(self.B)),
                                                dtype=torch.uint8, device=self.device)
                self.outs, self.mask, self.sizes = ops.bma_reduce_cls_sets(
                    z, want, calib.val_sorted, calib.threshold, targets=targets, out=self.outs or None, mask=self.mask,
                    sizes=self.sizes, workspace=self._sets_ws)
            else:
                self.outs = ops.bma_reduce_cls(z, want, targets=targets, out=self.outs or None)
            if kernel_events:
                kernel_events[1].record()
            self._mark("reduce")
            return self.outs
        if self.transport == "p2p":
            par = self.peer.step & 1
            self.peer.step += 1
            ops.bma_reduce_cls_shard(z, self.peer.slot_table(parity=par), targets=targets, rank=self.rank)
        else:
            ops.bma_reduce_cls_shard(z, self.slot_ptrs, targets=targets, rank=self.rank)
        if kernel_events:
            kernel_events[1].record()
        self._mark("reduce")
        # the one exchange step of the S-sharded path: every rank ends up with all ranks' partials for its B/N rows
        if self.transport == "p2p":
            self.peer.barrier()  # the kernel's epilogue already stored into the owners' buffers over NVLink
            recv = self.peer.recv_view(parity=par)
        else:
            self.exchange.exchange_slots(self.send, self.recv)
            recv = self.recv
        self._mark("exchange")
        want = [w for w in self.want if w != "ensemble_log_prob"]
        self.outs = ops.bma_finalize_cls_slots(recv, self.S * self.world, self.C, want, out=self.outs or None)
        self._mark("finalize")
        return self.outs

    # -------------------------------------------------------------------------------- K5 / K6 / K7
    def score(self, targets: torch.Tensor, calib: ConformalCalibration):
        mean = self.outs["mean"]
        lo, hi = self.rank * self.B_local, (self.rank + 1) * self.B_local
        t_local = targets[lo:hi] if self.world > 1 else targets
        if not (self.fused_sets and self._sets_ws is not None):
            self.mask, self.sizes = aps_sets(mean, calib)
        self._mark("sets")
        if self.world == 1 and "confidence" in self.outs:
            # the reduction kernel's epilogue already holds max_c mean, its arg-max and the Brier term of every row
            self.ece = ops.ece_bins_from_rows(self.outs["confidence"], self.outs["mode"], t_local, self.thresholds,
                                              brier_terms=self.outs.get("brier_terms"))
        else:
            self.ece = ops.ece_bins(mean, t_local, self.thresholds)
        self._mark("ece")
        if self.world > 1:
            self.exchange.allreduce_bins(self.ece)
            self.ece["result_global"], self.ece["counts_global"] = ops.ece_from_packed_bins(self.ece["packed"], self.B, True)
            self._mark("ece_allreduce")
        if self.timer is not None:
            self.timer.end_step()
        return self.mask, self.sizes, self.ece


def predictive_from_logits_batches(
    batches: Iterable[torch.Tensor],
    target_batches: Iterable[torch.Tensor],
    calib: ConformalCalibration,
    want: Sequence[str],
    device,
    host_out: Optional[Dict[str, torch.Tensor]] = None,
):
    """Public loader-shaped call.  `batches`: pinned host tensors [S, Bc, C] (f32 or bf16);
    `target_batches`: pinned host int32 [Bc].  Returns a dict of pinned host tensors over all inputs
    (mean, ..., conformal mask/sizes, ECE result).  H2D copies run on their own stream, two device buffers
    deep, overlapped with the kernels of the previous batch."""
    batches = list(batches)
    target_batches = list(target_batches)
    S, Bc, Cn = batches[0].shape
    n_b = len(batches)
    B = sum(int(b.shape[1]) for b in batches)
    copy_s, comp_s, out_s = torch.cuda.Stream(device), torch.cuda.Stream(device), torch.cuda.Stream(device)
    zbuf = [torch.empty((S, Bc, Cn), dtype=batches[0].dtype, device=device) for _ in range(2)]
    tbuf = [torch.empty(Bc, dtype=torch.int32, device=device) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]
    bc_names = [w for w in want if w in ("mean", "aleatoric_variance", "epistemic_variance", "variance", "std")]
    full = {w: torch.empty((B, Cn), dtype=torch.float32, device=device) for w in bc_names}
    vec = {w: torch.empty(B, dtype=torch.int32 if w == "mode" else torch.float32, device=device)
           for w in want if w not in bc_names and w != "ensemble_log_prob"}
    tfull = torch.empty(B, dtype=torch.int32, device=device)
    mask = torch.empty((B, Cn), dtype=torch.uint8, device=device)
    sizes = torch.empty(B, dtype=torch.int32, device=device)
    if host_out is None:
        host_out = {}
        for k, v in list(full.items()) + list(vec.items()) + [("conformal_mask", mask), ("conformal_sizes", sizes)]:
            host_out[k] = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
        host_out["ece_result"] = torch.empty(4, dtype=torch.float32, pin_memory=True)
    cur = torch.cuda.current_stream(device)
    copy_s.wait_stream(cur)
    comp_s.wait_stream(cur)
    out_s.wait_stream(cur)
    b0 = 0
    for i, (hb, ht) in enumerate(zip(batches, target_batches)):
        k = i & 1
        rows = int(hb.shape[1])
        with torch.cuda.stream(copy_s):
            if i >= 2:
                copy_s.wait_event(free[k])
            zbuf[k][:, :rows].copy_(hb, non_blocking=True) if rows != Bc else zbuf[k].copy_(hb, non_blocking=True)
            tbuf[k][:rows].copy_(ht, non_blocking=True)
            ready[k].record(copy_s)
        with torch.cuda.stream(comp_s):
            comp_s.wait_event(ready[k])
            zin = zbuf[k] if rows == Bc else zbuf[k][:, :rows].contiguous()
            tin = tbuf[k][:rows]
            outs = {w: full[w][b0 : b0 + rows] for w in bc_names}
            outs.update({w: vec[w][b0 : b0 + rows] for w in vec})
            ops.bma_reduce_cls(zin, want, targets=tin, out=outs)
            free[k].record(comp_s)
            tfull[b0 : b0 + rows].copy_(tin, non_blocking=True)
            m, sz = aps_sets(full["mean"][b0 : b0 + rows], calib)
            mask[b0 : b0 + rows].copy_(m, non_blocking=True)
            sizes[b0 : b0 + rows].copy_(sz, non_blocking=True)
            done = torch.cuda.Event()
            done.record(comp_s)
        # results of this batch go back to the host on their own stream while the next batch computes
        with torch.cuda.stream(out_s):
            out_s.wait_event(done)
            for k2, v in list(full.items()) + list(vec.items()) + [("conformal_mask", mask), ("conformal_sizes", sizes)]:
                host_out[k2][b0 : b0 + rows].copy_(v[b0 : b0 + rows], non_blocking=True)
        b0 += rows
    with torch.cuda.stream(comp_s):
        thr = torch.from_numpy(ece_thresholds(B)).to(device, non_blocking=True)
        ece = ops.ece_bins(full["mean"], tfull, thr)
        host_out["ece_result"].copy_(ece["result"], non_blocking=True)
    cur.wait_stream(comp_s)
    cur.wait_stream(copy_s)
    cur.wait_stream(out_s)
    return host_out


def predictive_from_feature_batches(
    x_batches: Iterable[torch.Tensor],
    target_batches: Iterable[torch.Tensor],
    wt: torch.Tensor,
    bias: Optional[torch.Tensor],
    log_temp: Optional[torch.Tensor],
    calib: ConformalCalibration,
    want: Sequence[str],
    device,
    sample_idx: Optional[torch.Tensor] = None,
    logits_dtype=torch.float32,
    host_out: Optional[Dict[str, torch.Tensor]] = None,
):
    """The loader-shaped predictive call for a last-layer posterior (what ``prob_model.predictive.mean/...``
    amounts to with ``freeze_fun``): pinned HOST feature batches ``[Bc, D]`` (bf16) stream to the device, the
    S-batched last-layer contraction (K2, tcgen05) forms the calibrated logits of every posterior sample from the
    device-resident sample bank ``wt [n, C, D]`` (bf16, K-major) / ``bias [n, C]`` - ``sample_idx`` (int32 [S]) are
    the ``posterior.sample`` draws, None = every stored sample once - and the logits go straight into the
    single-pass reductions (K3/K4) and the conformal / ECE scoring (K5-K7).  The [S,B,C] logits never leave the
    GPU.  Returns pinned host tensors over all inputs."""
    x_batches = list(x_batches)
    target_batches = list(target_batches)
    Bc, D = x_batches[0].shape
    S = wt.shape[0] if sample_idx is None else sample_idx.numel()
    Cn = wt.shape[1]
    B = sum(int(b.shape[0]) for b in x_batches)
    copy_s, comp_s, out_s = torch.cuda.Stream(device), torch.cuda.Stream(device), torch.cuda.Stream(device)
    xbuf = [torch.empty((Bc, D), dtype=x_batches[0].dtype, device=device) for _ in range(2)]
    tbuf = [torch.empty(Bc, dtype=torch.int32, device=device) for _ in range(2)]
    zbuf = torch.empty((S, Bc, Cn), dtype=logits_dtype, device=device)
    ready = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]
    bc_names = [w for w in want if w in ("mean", "aleatoric_variance", "epistemic_variance", "variance", "std")]
    full = {w: torch.empty((B, Cn), dtype=torch.float32, device=device) for w in bc_names}
    vec = {w: torch.empty(B, dtype=torch.int32 if w == "mode" else torch.float32, device=device)
           for w in want if w not in bc_names and w != "ensemble_log_prob"}
    tfull = torch.empty(B, dtype=torch.int32, device=device)
    mask = torch.empty((B, Cn), dtype=torch.uint8, device=device)
    sizes = torch.empty(B, dtype=torch.int32, device=device)
    if host_out is None:
        host_out = {}
        for k, v in list(full.items()) + list(vec.items()) + [("conformal_mask", mask), ("conformal_sizes", sizes)]:
            host_out[k] = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
        host_out["ece_result"] = torch.empty(4, dtype=torch.float32, pin_memory=True)
    cur = torch.cuda.current_stream(device)
    copy_s.wait_stream(cur)
    comp_s.wait_stream(cur)
    out_s.wait_stream(cur)
    b0 = 0
    for i, (hx, ht) in enumerate(zip(x_batches, target_batches)):
        k = i & 1
        rows = int(hx.shape[0])
        with torch.cuda.stream(copy_s):
            if i >= 2:
                copy_s.wait_event(free[k])
            xbuf[k][:rows].copy_(hx, non_blocking=True)
            tbuf[k][:rows].copy_(ht, non_blocking=True)
            ready[k].record(copy_s)
        with torch.cuda.stream(comp_s):
            comp_s.wait_event(ready[k])
            xin = xbuf[k][:rows]
            tin = tbuf[k][:rows]
            z = zbuf if rows == Bc else torch.empty((S, rows, Cn), dtype=logits_dtype, device=device)
            ops.last_layer_logits(xin, wt, bias, log_temp, w_layout=ops.W_SCD, out=z, sample_idx=sample_idx)
            free[k].record(comp_s)
            outs = {w: full[w][b0 : b0 + rows] for w in bc_names}
            outs.update({w: vec[w][b0 : b0 + rows] for w in vec})
            ops.bma_reduce_cls(z, want, targets=tin, out=outs)
            tfull[b0 : b0 + rows].copy_(tin, non_blocking=True)
            m, sz = aps_sets(full["mean"][b0 : b0 + rows], calib)
            mask[b0 : b0 + rows].copy_(m, non_blocking=True)
            sizes[b0 : b0 + rows].copy_(sz, non_blocking=True)
            done = torch.cuda.Event()
            done.record(comp_s)
        # results of this batch go back to the host on their own stream while the next batch computes
        with torch.cuda.stream(out_s):
            out_s.wait_event(done)
            for k2, v in list(full.items()) + list(vec.items()) + [("conformal_mask", mask), ("conformal_sizes", sizes)]:
                host_out[k2][b0 : b0 + rows].copy_(v[b0 : b0 + rows], non_blocking=True)
        b0 += rows
    with torch.cuda.stream(comp_s):
        thr = torch.from_numpy(ece_thresholds(B)).to(device, non_blocking=True)
        ece = ops.ece_bins(full["mean"], tfull, thr)
        host_out["ece_result"].copy_(ece["result"], non_blocking=True)
    cur.wait_stream(comp_s)
    cur.wait_stream(copy_s)
    cur.wait_stream(out_s)
    return host_out


def bench_e2e_from_features(S, B, D, Cn, device, want, calib, args, unit, logits_dtype=torch.float32):
    """Times predictive_from_feature_batches: pinned host features [B, D] bf16 in, pinned host results out; the S
    last-layer samples are resident on the device (that is where the chains left them)."""
    Bc = min(8192, B)
    n_batches = (B + Bc - 1) // Bc
    gen = torch.Generator(device=device)
    gen.manual_seed(11)
    wt = (torch.empty((S, Cn, D), device=device, dtype=torch.float32).normal_(generator=gen) / D ** 0.5).to(torch.bfloat16)
    bias = torch.empty((S, Cn), device=device, dtype=torch.float32).normal_(generator=gen)
    log_temp = torch.tensor([0.3], device=device)
    host_x, host_t = [], []
    for _ in range(n_batches):
        d = torch.empty((Bc, D), device=device, dtype=torch.float32).normal_(generator=gen).to(torch.bfloat16)
        hx = torch.empty((Bc, D), dtype=torch.bfloat16, pin_memory=True)
        hx.copy_(d)
        host_x.append(hx)
        host_t.append(torch.randint(0, Cn, (Bc,), dtype=torch.int32).pin_memory())
        del d
    host_out = None
    times = []
    for it in range(1 + args.e2e_steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        host_out = predictive_from_feature_batches(host_x, host_t, wt, bias, log_temp, calib, want, device,
                                                   logits_dtype=logits_dtype, host_out=host_out)
        torch.cuda.synchronize()
        if it > 0:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    h2d = n_batches * (Bc * D * 2 + Bc * 4)
    d2h = sum(v.numel() * v.element_size() for v in host_out.values())
    return {
        "value": float(S) * B * Cn / sec,
        "unit": unit,
        "ms_per_step": sec * 1e3,
        "h2d_bytes_per_step": int(h2d),
        "d2h_bytes_per_step": int(d2h),
        "steps": len(times),
        "D": D,
        "gemm_flops_per_step": 2.0 * S * B * D * Cn,
        "logits": "float32, device-only" if logits_dtype == torch.float32 else "bfloat16, device-only",
        "api": "fortuna_b200.prob_model.predictive.pipeline.predictive_from_feature_batches",
        "what": f"host features [B,{D}] bf16 -> last-layer GEMM over {S} resident samples -> reductions -> APS sets + ECE -> host",
    }


def bench_e2e(S, B, Cn, device, want, calib, args, unit):
    """Times predictive_from_logits_batches with pinned host inputs; returns the bench `e2e` object."""
    Bc = min(args.e2e_batch, B)
    n_batches = (B + Bc - 1) // Bc
    n_host = min(4, n_batches)  # distinct pinned batches, cycled (bounds pinned memory to ~4 GB)
    gen = torch.Generator(device=device)
    gen.manual_seed(7)
    host_b, host_t = [], []
    for _ in range(n_host):
        d = torch.empty((S, Bc, Cn), dtype=torch.float32, device=device).normal_(0.0, 4.0, generator=gen)
        hb = torch.empty((S, Bc, Cn), dtype=torch.float32, pin_memory=True)
        hb.copy_(d)
        ht = torch.randint(0, Cn, (Bc,), dtype=torch.int32).pin_memory()
        host_b.append(hb)
        host_t.append(ht)
        del d
    batches = [host_b[i % n_host] for i in range(n_batches)]
    tbs = [host_t[i % n_host] for i in range(n_batches)]
    host_out = None
    times = []
    for it in range(1 + args.e2e_steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        host_out = predictive_from_logits_batches(batches, tbs, calib, want, device, host_out)
        torch.cuda.synchronize()
        if it > 0:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    h2d = n_batches * (S * Bc * Cn * 4 + Bc * 4)
    d2h = sum(v.numel() * v.element_size() for v in host_out.values())
    return {
        "value": float(S) * B * Cn / sec,
        "unit": unit,
        "h2d_bytes_per_step": int(h2d),
        "d2h_bytes_per_step": int(d2h),
        "ms_per_step": sec * 1e3,
        "steps": len(times),
        "api": "fortuna_b200.prob_model.predictive.pipeline.predictive_from_logits_batches",
        "host_batches": f"{n_batches} batches of [S,{Bc},C] per step from {n_host} distinct pinned buffers (cycled)",
    }
