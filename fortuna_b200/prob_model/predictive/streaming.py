"""The loader-shaped engine behind ``prob_model.predictive.*``: what the reference's ``_loop_fun_through_inputs_loader``
/ ``_loop_fun_through_data_loader`` do around every ``_batched_*`` statistic (fortuna/prob_model/predictive/base.py:
946-1013: iterate the loader batch by batch, evaluate, concatenate; with ``distribute=True`` through ``pmap``,
fortuna/data/loader/base.py:535-553).

Here one pass over the loader computes ANY set of statistics:

  host batch -> pinned staging buffer -> H2D copy on its own stream (two buffers deep: the copy of batch i + 1 overlaps
  the kernels of batch i) -> ``logits_fn`` (the backbone once + the S-batched last-layer contraction K2, or one forward
  pass per stored sample) -> the single-pass reduction kernel K3 writing straight into this batch's rows of the result.

``distribute=True`` in a multi-rank NCCL job shards the POSTERIOR SAMPLES (SURVEY 8e / BASELINE c3: "S samples sharded
8-way"): rank r evaluates draws ``[r S/N, (r+1) S/N)`` of every batch, the reduction kernel's epilogue stores its partial
sums row by row into the owning rank's receive buffer over NVLink (or one all-to-all), ONE barrier, the owners finalise
their ``B/N`` rows, and the finalised rows are all-gathered so that every rank returns the whole result - the same arrays
a single GPU returns (sums associate per rank instead of strictly sample by sample: equal to float32 rounding).
The reference's ``pmap`` shards the batch instead; outputs are identical either way, the sample split is what lets
every GPU keep only its own chain's samples hot.

All arithmetic is in the library kernels (plus the user's backbone); torch provides memory, streams and NCCL."""
from typing import Callable, Dict, Iterable, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from ... import ops

_BC = ("mean", "aleatoric_variance", "epistemic_variance", "variance", "std")
_ROW_I32 = ("mode",)


def dist_world(distribute: bool):
    """(world, rank) to shard over, or (1, 0): an initialised NCCL group with more than one rank and distribute=True."""
    if distribute and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 \
            and dist.get_backend() == "nccl":
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


class ShardContext:
    """Per-predictive state of the S-sharded exchange: peer-mapped receive buffers (symmetric memory rendezvous is a
    collective and takes milliseconds, so the buffers are cached and grown only when a larger batch shows up) or the
    send / receive pair of the all-to-all fallback."""

    def __init__(self, world: int, rank: int, device, transport: Optional[str] = None):
        self.world, self.rank, self.device = world, rank, device
        self.transport = transport or "p2p"
        self.peer = None
        self.send = self.recv = None
        self._send_rows = 0
        self._C = None

    def _ensure(self, rows_per_rank: int, C: int):
        rowf = ops.cls_slot_row_floats(C)
        if self.transport == "p2p" and (self.peer is None or self.peer.rows < rows_per_rank or self._C != C):
            from .sharded import PeerSlots

            try:
                self.peer = PeerSlots(self.world, self.rank, rows_per_rank, rowf, self.device)
            except Exception as exc:  # no peer mapping on this system: one all-to-all per batch instead
                import warnings

                warnings.warn(f"peer-memory exchange unavailable ({exc!r}); using the NCCL all-to-all")
                self.transport = "nccl"
        if self.transport != "p2p" and (self.send is None or self._send_rows < rows_per_rank or self._C != C):
            self.send = ops.alloc_cls_slots(self.world, rows_per_rank, C, self.device)
            self.recv = torch.empty_like(self.send)
            self._send_rows = rows_per_rank
        self._C = C

    def reduce(self, z: torch.Tensor, targets: Optional[torch.Tensor], s_total: int, want: Sequence[str],
               log_temp=None, ens_local: Optional[torch.Tensor] = None, out=None) -> Dict[str, torch.Tensor]:
        """z [S_local, rows, C] (rows % world == 0) -> finalised outputs of THIS rank's rows / world rows.
        `targets`: int32 [rows] of the WHOLE batch (every rank scores its samples on all rows)."""
        _, rows, C = z.shape
        per = rows // self.world
        self._ensure(per, C)
        if self.transport == "p2p":
            par = self.peer.step & 1
            self.peer.step += 1
            ops.bma_reduce_cls_shard(z, self.peer.slot_table(per, par), targets=targets, log_temp=log_temp, rank=self.rank,
                                     ensemble_log_prob=ens_local)
            self.peer.barrier()
            recv = self.peer.recv_view(per, par)
        else:
            rowf = ops.cls_slot_row_floats(C)
            send = self.send.reshape(-1)[: self.world * per * rowf].view(self.world, per, rowf)
            recv = self.recv.reshape(-1)[: self.world * per * rowf].view(self.world, per, rowf)
            table = ops.slot_pointer_table([send[o] for o in range(self.world)], self.device)
            ops.bma_reduce_cls_shard(z, table, targets=targets, log_temp=log_temp, rank=self.rank,
                                     ensemble_log_prob=ens_local)
            dist.all_to_all_single(recv.reshape(-1), send.reshape(-1))
        t_own = targets[self.rank * per : (self.rank + 1) * per] if targets is not None else None
        return ops.bma_finalize_cls_slots(recv, s_total, C, want, out=out, targets=t_own)


def _alloc_result(names, S, B, C, device):
    out = {}
    for n in names:
        if n in _BC:
            out[n] = torch.empty((B, C), dtype=torch.float32, device=device)
        elif n == "ensemble_log_prob":
            out[n] = torch.empty((S, B), dtype=torch.float32, device=device)
        elif n in _ROW_I32:
            out[n] = torch.empty(B, dtype=torch.int32, device=device)
        else:
            out[n] = torch.empty(B, dtype=torch.float32, device=device)
    return out


class _Staging:
    """Two pinned host buffers + two device buffers for one stream of batches (inputs or targets)."""

    def __init__(self, device):
        self.device = device
        self.host = [None, None]
        self.dev = [None, None]
        self.done = [None, None]  # the H2D copy that last read host buffer k

    def put(self, k: int, arr, rows_padded: int, dtype) -> torch.Tensor:
        """Copy `arr` (numpy / tensor, [rows, ...]) into device buffer k (zero-padded to rows_padded) on the current
        stream; returns the [rows_padded, ...] device view.  ``dtype=None`` keeps a floating tensor's dtype (bf16 features
        stay bf16) and makes everything else float32.  A PINNED host tensor of the right dtype is copied from directly (no
        staging memcpy); a CUDA tensor is used as it is."""
        if dtype is None:
            dtype = arr.dtype if isinstance(arr, torch.Tensor) and arr.dtype in (torch.float32, torch.bfloat16) else torch.float32
        if isinstance(arr, torch.Tensor) and arr.is_cuda:
            src = arr.to(dtype)
            rows = src.shape[0]
            if rows == rows_padded:
                return src.contiguous()
            out = torch.zeros((rows_padded,) + tuple(src.shape[1:]), dtype=dtype, device=self.device)
            out[:rows].copy_(src)
            return out
        a = arr if isinstance(arr, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(arr))
        rows, tail = a.shape[0], tuple(a.shape[1:])
        need = (rows_padded,) + tail
        d = self.dev[k]
        if d is None or d.dtype != dtype or tuple(d.shape[1:]) != tail or d.shape[0] < rows_padded:
            self.dev[k] = d = torch.empty(need, dtype=dtype, device=self.device)
            self.host[k] = None
        d = d[:rows_padded]
        if a.is_pinned() and a.dtype == dtype and a.is_contiguous() and rows == rows_padded:
            d.copy_(a, non_blocking=True)  # the caller's pinned buffer is the DMA source
            return d
        h = self.host[k]
        if h is None or h.shape[0] < rows_padded:
            self.host[k] = h = torch.empty(need, dtype=dtype, pin_memory=True)
        elif self.done[k] is not None:
            self.done[k].synchronize()  # the DMA that read this pinned buffer two batches ago must have finished
        h[:rows].copy_(a)  # host memcpy (+ dtype conversion) into pinned memory
        if rows_padded > rows:
            h[rows:rows_padded].zero_()
        d.copy_(h[:rows_padded], non_blocking=True)
        self.done[k] = torch.cuda.Event()
        self.done[k].record()
        return d


def stream_cls_reduce(
    batches: Iterable,
    logits_fn: Callable[[torch.Tensor], torch.Tensor],
    names: Sequence[str],
    n_rows: int,
    n_samples: int,
    device,
    targets=None,
    log_temp: Optional[torch.Tensor] = None,
    shard: Optional[ShardContext] = None,
    sets=None,
    staging: Optional[dict] = None,
    to_host=False,
    sharded_io: bool = False,
    ece_total: Optional[int] = None,
) -> Dict[str, torch.Tensor]:
    """One pass over `batches` (arrays [Bc, ...], or (inputs, targets) pairs as a data loader yields them - the targets
    then come from the batches, in the loader's own order); ``logits_fn(x_dev) -> z [S_local, Bc, C]`` (calibrated logits
    of this rank's draws unless ``log_temp`` is given, which the reduction then applies).  `targets`: optional int array
    over all rows.  `sets`: optional ConformalCalibration - additionally returns the APS conformal ``conformal_mask``
    [B, C] uint8 and ``conformal_sizes`` [B] of the mean (fortuna/conformal/classification/adaptive_prediction.py on
    ``predictive.mean``).  `ece_total`: the TOTAL number of data points of the evaluation (all ranks) - additionally
    returns ``ece`` float32 [2] = (ECE, MCE) of the mean against the targets (fortuna/metric/classification.py:43-133):
    every batch's rows are binned with the thresholds of the whole data set, the additive bin counters accumulate over
    the batches (and, with sharded inputs / outputs, over the ranks in one all-reduce).

    Returns CUDA tensors over all `n_rows` rows, on every rank - or, with ``to_host`` (True, or a dict of pinned host
    tensors to reuse), PINNED HOST tensors: each batch's rows leave on a device-to-host stream of their own while the next
    batch computes, and the call returns when they have landed.

    ``sharded_io`` (with `shard`): every rank passes ITS OWN rows of each global batch (the way the reference's loaders hand
    each device its slice, fortuna/data/loader/base.py:535-553) and gets back the results of exactly those rows: the
    inputs are all-gathered over NVLink, every rank evaluates its S / N draws on the whole batch, and after the exchange
    rank r owns rows ``[r Bc/N, (r+1) Bc/N)`` - its own.  Host traffic and result storage are 1/N per rank."""
    names = tuple(names)
    world = shard.world if shard is not None else 1
    sharded_io = bool(sharded_io) and world > 1
    if targets is not None and not isinstance(targets, torch.Tensor):
        targets = np.asarray(targets)
    staging = staging if staging is not None else {}
    if "x" not in staging:  # streams and staging buffers live as long as the predictive object
        staging["x"], staging["t"] = _Staging(device), _Staging(device)
        staging["copy_stream"], staging["out_stream"] = torch.cuda.Stream(device), torch.cuda.Stream(device)
    st_x, st_t = staging["x"], staging["t"]
    cur = torch.cuda.current_stream(device)
    copy_s = staging["copy_stream"]
    copy_s.wait_stream(cur)
    out_s = None
    if to_host is not False and to_host is not None:
        out_s = staging["out_stream"]
        out_s.wait_stream(cur)
    host = to_host if isinstance(to_host, dict) else ({} if out_s is not None else None)
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    free = [torch.cuda.Event(), torch.cuda.Event()]
    result = None
    want_ens = "ensemble_log_prob" in names
    if want_ens and sharded_io:
        raise NotImplementedError("ensemble_log_prob lives with the samples: not available with sharded inputs / outputs")
    kernel_names = tuple(n for n in names if n != "ensemble_log_prob" or world == 1)
    want_targets = any(n in names for n in ("log_prob", "ensemble_log_prob", "brier_terms"))
    if sets is not None and "mean" not in kernel_names:
        kernel_names = kernel_names + ("mean",)
    ece_packed = ece_thr = None
    if ece_total is not None:
        from .pipeline import ece_thresholds

        ece_thr = torch.from_numpy(ece_thresholds(int(ece_total))).to(device)
        want_targets = True
        for n in ("confidence", "mode") + (("brier_terms",) if world == 1 or sharded_io else ()):  # per-row ECE inputs
            if n not in kernel_names:
                kernel_names = kernel_names + (n,)
        names = names + tuple(n for n in kernel_names if n not in names)
    b0 = 0
    for i, xb in enumerate(batches):
        k = i & 1
        tb = None
        if isinstance(xb, (tuple, list)):  # a data loader's (inputs, targets) batch: the targets travel with their inputs
            xb, tb = xb[0], xb[1]
        rows = int(xb.shape[0])
        rows_p = rows if sharded_io else (rows + world - 1) // world * world
        with torch.cuda.stream(copy_s):
            if i >= 2:
                copy_s.wait_event(free[k])
            x_dev = st_x.put(k, xb, rows_p, None)
            t_dev = None
            if tb is not None and want_targets:
                t_dev = st_t.put(k, tb, rows_p, torch.int32)
            elif targets is not None:
                t_dev = st_t.put(k, targets[b0 : b0 + rows], rows_p, torch.int32)
            ready[k].record(copy_s)
        cur.wait_event(ready[k])
        if sharded_io:  # the batch's inputs (and targets) of all ranks, rank order = row order
            x_all = torch.empty((world * rows,) + tuple(x_dev.shape[1:]), dtype=x_dev.dtype, device=device)
            dist.all_gather_into_tensor(x_all, x_dev)
            x_dev = x_all
            if t_dev is not None:
                t_all = torch.empty(world * rows, dtype=torch.int32, device=device)
                dist.all_gather_into_tensor(t_all, t_dev)
                t_dev = t_all
        z = logits_fn(x_dev)
        S_local, _, C = z.shape
        if result is None:
            result = _alloc_result(names, n_samples, n_rows, C, device)
            if sets is not None:
                result["conformal_mask"] = torch.empty((n_rows, C), dtype=torch.uint8, device=device)
                result["conformal_sizes"] = torch.empty(n_rows, dtype=torch.int32, device=device)
                if "mean" not in result:
                    result["mean"] = torch.empty((n_rows, C), dtype=torch.float32, device=device)
            if host is not None:
                for n, v in result.items():
                    if n not in host or host[n].shape != v.shape or host[n].dtype != v.dtype:
                        host[n] = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
        if world == 1:
            outs = {n: (result[n][b0 : b0 + rows] if n != "ensemble_log_prob" else
                        torch.empty((n_samples, rows), dtype=torch.float32, device=device)) for n in kernel_names}
            ops.bma_reduce_cls(z, kernel_names, targets=t_dev, log_temp=log_temp, out=outs)
            if want_ens:
                result["ensemble_log_prob"][:, b0 : b0 + rows].copy_(outs["ensemble_log_prob"])
        elif sharded_io:
            outs = {n: result[n][b0 : b0 + rows] for n in kernel_names}
            shard.reduce(z, t_dev, n_samples, kernel_names, log_temp=log_temp, out=outs)
        else:
            ens_local = torch.empty((S_local, rows_p), dtype=torch.float32, device=device) if want_ens else None
            local = shard.reduce(z, t_dev, n_samples, kernel_names, log_temp=log_temp, ens_local=ens_local)
            for n in kernel_names:  # every rank returns the whole result, like the reference's gathered pmap output
                full = torch.empty((rows_p,) + tuple(local[n].shape[1:]), dtype=local[n].dtype, device=device)
                dist.all_gather_into_tensor(full, local[n].contiguous())
                result[n][b0 : b0 + rows].copy_(full[:rows])
            if want_ens:
                allens = torch.empty((world * S_local, rows_p), dtype=torch.float32, device=device)
                dist.all_gather_into_tensor(allens, ens_local)
                result["ensemble_log_prob"][:, b0 : b0 + rows].copy_(allens[:, :rows])
        if sets is not None:
            from .pipeline import aps_sets

            m, sz = aps_sets(result["mean"][b0 : b0 + rows], sets)
            result["conformal_mask"][b0 : b0 + rows].copy_(m)
            result["conformal_sizes"][b0 : b0 + rows].copy_(sz)
        if ece_thr is not None:
            lo = shard.rank * rows if sharded_io else 0  # with sharded IO t_dev holds the whole batch's targets
            e = ops.ece_bins_from_rows(result["confidence"][b0 : b0 + rows], result["mode"][b0 : b0 + rows],
                                       t_dev[lo : lo + rows], ece_thr,
                                       brier_terms=result["brier_terms"][b0 : b0 + rows] if "brier_terms" in result else None)
            ece_packed = ops.ece_pack_bins(e, out=ece_packed, accumulate=ece_packed is not None)
        free[k].record(cur)  # staging buffers k may be overwritten once everything of this batch has run
        if out_s is not None:  # this batch's rows go home while the next batch computes
            out_s.wait_event(free[k])
            with torch.cuda.stream(out_s):
                for n, v in result.items():
                    if n == "ensemble_log_prob":
                        host[n][:, b0 : b0 + rows].copy_(v[:, b0 : b0 + rows], non_blocking=True)
                    else:
                        host[n][b0 : b0 + rows].copy_(v[b0 : b0 + rows], non_blocking=True)
        b0 += rows
    if result is None:
        raise ValueError("the loader yielded no batches")
    if b0 != n_rows:
        raise ValueError(f"the loader yielded {b0} rows, expected {n_rows}")
    cur.wait_stream(copy_s)
    if ece_packed is not None:
        if sharded_io:
            dist.all_reduce(ece_packed, op=dist.ReduceOp.SUM)
        result["ece"] = ops.ece_from_packed_bins(ece_packed, int(ece_total))
        if host is not None:
            host["ece"] = result["ece"].cpu()
    if out_s is not None:
        out_s.synchronize()
        return host
    return result
