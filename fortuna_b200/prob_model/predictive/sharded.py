"""The one exchange step of the S-sharded predictive: each rank holds S/N posterior
samples and all B inputs.  fb200_bma_reduce_cls_shard reduces the rank's own samples and writes, row by row, the
partial sums into the "slot" of the rank that owns the row (owner(b) = b // (B/N)); a slot row is
``[sum_s p (C) | sum_s p^2 (C) | sum_s sum_c p ln p, max_s ll, sum_s exp(ll - max), 0]``.  Two transports:

  * ``nccl`` / ``gloo``: slots are a local send buffer [N, B/N, row]; ONE all-to-all moves slot o to rank o;
  * ``p2p``: slots point into the owners' receive buffers mapped over NVLink (torch symmetric memory), so the
    reduction kernel's epilogue IS the exchange - its peer stores overlap the rest of the kernel - and only a
    barrier separates it from the finalisation.

Either way rank r ends up with ``recv[N_src, B/N, row]`` and fb200_bma_finalize_cls_slots sums the sources in rank
order and merges the (max, sumexp) pairs.  After scoring the local rows, the additive ECE bin counters are
all-reduced.  Only torch.distributed calls and index arithmetic live here - no arithmetic of the path.
"""

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist


def local_rows(B: int, world: int, rank: int) -> Tuple[int, int]:
    """Rows of the input axis owned by `rank` after the exchange; B must divide evenly."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} / world {world}")
    if B % world:
        raise ValueError(f"the number of inputs ({B}) must be divisible by the number of ranks ({world})")
    per = B // world
    return rank * per, (rank + 1) * per


class SampleShardExchange:
    def __init__(self, world: Optional[int] = None, rank: Optional[int] = None, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if world is None else world
        self.rank = dist.get_rank(group) if rank is None else rank

    def exchange_slots(self, send: torch.Tensor, recv: Optional[torch.Tensor] = None) -> torch.Tensor:
        """send [N, rows, row_floats]: slot o holds this rank's partials for the rows rank o owns.
        Returns recv [N, rows, row_floats]: slot s holds rank s's partials for the rows THIS rank owns."""
        if send.shape[0] != self.world:
            raise ValueError(f"send buffer has {send.shape[0]} slots for {self.world} ranks")
        if recv is None:
            recv = torch.empty_like(send)
        dist.all_to_all_single(recv.view(-1), send.view(-1), group=self.group)
        return recv

    def allreduce_bins(self, ece: Dict[str, torch.Tensor]):
        """Sum the additive ECE counters over the ranks.  On the GPU: ONE collective over the packed float64 vector
        (fb200_ece_pack_bins), which ``ece["packed"]`` then holds; CPU tensors (the gloo tests): one per counter."""
        if ece["counts"].is_cuda:
            from ... import ops

            ece["packed"] = ops.ece_pack_bins(ece, out=ece.get("packed"))
            dist.all_reduce(ece["packed"], op=dist.ReduceOp.SUM, group=self.group)
            return ece
        for n in ("counts", "correct", "conf_sums"):
            dist.all_reduce(ece[n], op=dist.ReduceOp.SUM, group=self.group)
        return ece


class PeerSlots:
    """Receive buffers mapped into every rank's address space (torch.distributed._symmetric_memory): rank r's
    reduction kernel stores its partials for rows owned by rank o straight into ``recv_o[r]`` over NVLink.

    TWO receive buffers, used alternately: step i stores into buffer i & 1, so ONE barrier per step is enough - it makes
    step i's peer stores visible before the finalisation reads them, and, because every rank reaches the barrier of step
    i + 1 only after its own finalisation of step i (stream order), nobody's step i + 2 stores can land in a buffer that
    is still being read.  (A single buffer needs a second barrier after the finalisation.)

    ``rows`` is the LARGEST number of rows per rank; a step with fewer rows (a loader's last batch) uses the front of the
    buffer with its own slot stride (``dest_ptrs(rows)`` / ``recv_view(rows)``)."""

    def __init__(self, world: int, rank: int, rows: int, row_floats: int, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem

        self.world, self.rank, self.rows, self.row_floats = world, rank, int(rows), int(row_floats)
        group = group if group is not None else dist.group.WORLD
        self._buf = symm_mem.empty((2, world, self.rows, self.row_floats), dtype=torch.float32, device=device)
        self.handle = symm_mem.rendezvous(self._buf, group.group_name if hasattr(group, "group_name") else group)
        self._half_bytes = world * self.rows * self.row_floats * 4
        self._tables = {}
        self.step = 0
        self.recv = self._buf[0]  # first buffer, full size (kept for callers that look at one step)

    def dest_ptrs(self, rows: int = None, parity: int = 0):
        """Where THIS rank writes inside every owner's receive buffer `parity`: slot index = this rank."""
        rows = self.rows if rows is None else int(rows)
        slot_bytes = rows * self.row_floats * 4
        return [int(self.handle.buffer_ptrs[o]) + parity * self._half_bytes + self.rank * slot_bytes
                for o in range(self.world)]

    def slot_table(self, rows: int = None, parity: int = 0) -> torch.Tensor:
        """Device pointer table for fb200_bma_reduce_cls_shard (cached per (rows, parity))."""
        from ... import ops

        rows = self.rows if rows is None else int(rows)
        key = (rows, parity)
        if key not in self._tables:
            self._tables[key] = ops.slot_pointer_table(self.dest_ptrs(rows, parity), self._buf.device)
        return self._tables[key]

    def recv_view(self, rows: int = None, parity: int = 0) -> torch.Tensor:
        """[world, rows, row_floats] view of receive buffer `parity` for a step with `rows` rows per rank."""
        rows = self.rows if rows is None else int(rows)
        flat = self._buf[parity].reshape(-1)
        return flat[: self.world * rows * self.row_floats].view(self.world, rows, self.row_floats)

    def barrier(self):
        """All ranks' peer stores issued before this point (on the current stream) are visible after it."""
        self.handle.barrier()
