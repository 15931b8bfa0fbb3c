"""Device-side orchestration of the predictive path over calibrated logits z[S,B,C] that are already resident in HBM
(one GPU, or S-sharded over the ranks of a torch.distributed NCCL group):
reduce (K3/K4) -> APS scores (K5) -> rank-count conformal sets (K6) -> ECE (K7) - ``PredictivePipeline``, the benchmark's
device-resident step.  The loader-shaped path users call (``prob_model.predictive.*``: host batches in, results out) is
``streaming.stream_cls_reduce``.  All arithmetic is in the library kernels; torch provides memory, streams and NCCL only.
"""

from typing import Dict, Optional, Sequence, Tuple

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
    def __init__(self, S, B, C, device, want: Sequence[str], world: int = 1, rank: int = 0, transport: Optional[str] = None):
        self.S, self.B, self.C = S, B, C
        self.device = device
        self.want = tuple(want)
        # single GPU: the reduction kernel also emits the three per-row quantities the ECE needs (K7 then reads 3 B floats
        # instead of the [B,C] mean)
        # (sharded: the finalisation kernel emits them for the owned rows)
        self._kernel_want = self.want + tuple(n for n in ("mode", "confidence", "brier_terms") if n not in want)
        self.world, self.rank = world, rank
        self.B_local = B // world
        self.outs: Dict[str, torch.Tensor] = {}
        self.scores = None
        self.thresholds = torch.from_numpy(ece_thresholds(B)).to(device)
        self.transport = "single"
        self.timer: Optional[PhaseTimer] = None
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
        # the sets are one launch for C <= 256 (sort + sets fused), two for wide rows (select kernel + the full sort over the
        # rows it left undecided, or - C > 1024 - scores then sets)
        self.launches_per_step = (4 if world == 1 else 7) + (0 if C <= 256 or (C <= 1024 and C % 4) else 1)

    def _mark(self, name):
        if self.timer is not None:
            self.timer.mark(name)

    # -------------------------------------------------------------------------------- K3 / K4
    def reduce(self, z: torch.Tensor, targets: Optional[torch.Tensor], kernel_events=None):
        """kernel_events: optional (start, end) CUDA events recorded around the reduction kernel alone."""
        self._mark("begin")
        if kernel_events:
            kernel_events[0].record()
        if self.world == 1:
            want = self._kernel_want if targets is not None else self.want
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
        want = [w for w in (self._kernel_want if targets is not None else self.want) if w != "ensemble_log_prob"]
        lo = self.rank * self.B_local
        self.outs = ops.bma_finalize_cls_slots(recv, self.S * self.world, self.C, want, out=self.outs or None,
                                               targets=targets[lo : lo + self.B_local] if targets is not None else None)
        self._mark("finalize")
        return self.outs

    # -------------------------------------------------------------------------------- K5 / K6 / K7
    def score(self, targets: torch.Tensor, calib: ConformalCalibration):
        mean = self.outs["mean"]
        lo, hi = self.rank * self.B_local, (self.rank + 1) * self.B_local
        t_local = targets[lo:hi] if self.world > 1 else targets
        self.mask, self.sizes = aps_sets(mean, calib)
        self._mark("sets")
        if "confidence" in self.outs:
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
