"""The one exchange step of the S-sharded predictive: each rank holds S/N posterior
samples and all B inputs, reduces its own samples to partial sums with fb200_bma_reduce_cls (partial mode),
then

  * reduce-scatter (sum) over the input axis of  sum_s p, sum_s p^2, sum_s p(1-p)  [B,C]  and
    sum_s sum_c p ln p  [B]  -> rank r owns rows [r*B/N, (r+1)*B/N);
  * all-gather of the per-rank (max_s ll, sum_s exp(ll - max)) [B] pairs, sliced to the local rows, merged by
    fb200_merge_lse_pairs;
  * all-reduce of the ECE bin counters after scoring the local rows.

Only torch.distributed calls and index arithmetic live here - no arithmetic of the path.  NCCL runs the
reduce-scatter natively; gloo (CPU tests, tests/test_dist_exchange.py) has no reduce-scatter, so there the
same result is formed by all-reduce + slice.
"""

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

SUM_FIELDS = ("sum_p", "sum_p2", "sum_pq", "sum_plogp")


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
        self._native_rs = dist.get_backend(group) == "nccl"

    def scatter_sum(self, full: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """full [B, ...] on every rank -> sum over ranks of rows local_rows(B) -> [B/N, ...]."""
        lo, hi = local_rows(full.shape[0], self.world, self.rank)
        if out is None:
            out = torch.empty((hi - lo,) + tuple(full.shape[1:]), dtype=full.dtype, device=full.device)
        if self._native_rs:
            dist.reduce_scatter_tensor(out, full, op=dist.ReduceOp.SUM, group=self.group)
        else:
            tmp = full.clone()
            dist.all_reduce(tmp, op=dist.ReduceOp.SUM, group=self.group)
            out.copy_(tmp[lo:hi])
        return out

    def gather_rows(self, vec: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """vec [B] on every rank -> [N, B/N]: every rank's values for the rows this rank owns."""
        lo, hi = local_rows(vec.shape[0], self.world, self.rank)
        if out is None:
            out = torch.empty((self.world, vec.shape[0]), dtype=vec.dtype, device=vec.device)
        dist.all_gather_into_tensor(out.view(-1), vec, group=self.group)  # rank-major concatenation
        return out[:, lo:hi].contiguous()

    def exchange_partials(self, p: Dict[str, torch.Tensor], buffers: Optional[Dict[str, torch.Tensor]] = None):
        """Partial sums of this rank's samples -> (summed partials for the local rows, gathered (max, sumexp)
        pairs [N, B/N] for the local rows)."""
        buffers = buffers or {}
        summed = {n: self.scatter_sum(p[n], buffers.get(n)) for n in SUM_FIELDS}
        gmax = self.gather_rows(p["ll_max"], buffers.get("gather_max"))
        gsum = self.gather_rows(p["ll_sumexp"], buffers.get("gather_sum"))
        return summed, gmax, gsum

    def allreduce_bins(self, ece: Dict[str, torch.Tensor]):
        for n in ("counts", "correct", "conf_sums"):
            dist.all_reduce(ece[n], op=dist.ReduceOp.SUM, group=self.group)
        return ece
