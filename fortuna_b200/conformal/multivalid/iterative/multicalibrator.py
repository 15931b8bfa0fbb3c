"""``Multicalibrator`` - fortuna/conformal/multivalid/iterative/multicalibrator.py:14-151 with the loss of
mixins/multicalibrator.py:6-15."""
from typing import Tuple

import numpy as np
import torch

from .... import ops
from .base import F32, IterativeMultivalidMethod


def _rates(stats: np.ndarray, n: int):
    """(prob_b, mean(stat b), mean(x b)) in float32 from (count, sum stat, sum x): jnp.mean = float32 sum / n."""
    with np.errstate(all="ignore"):
        prob_b = F32(stats[..., 0]) / F32(n)
        ms = F32(stats[..., 1]) / F32(n)
        mx = F32(stats[..., 2]) / F32(n)
    return prob_b.astype(F32), ms.astype(F32), mx.astype(F32)


class Multicalibrator(IterativeMultivalidMethod):
    _mode = 0

    def __init__(self, seed: int = 0):
        super().__init__(seed=seed)
        self._patch_list = []

    def mean_squared_error(self, values, scores) -> torch.Tensor:
        from ..base import to_f32

        v, s = to_f32(values), to_f32(scores)
        v = v[:, None].contiguous() if v.ndim == 1 else v
        s = s[:, None].contiguous() if s.ndim == 1 else s
        return torch.tensor(self._loss(v, s), dtype=torch.float32)

    def _loss(self, values: torch.Tensor, scores: torch.Tensor) -> float:
        """mixins/multicalibrator.py:10-15.  With 1-D values the reference's ``mean(sum((values - scores) ** 2, -1))`` is
        the SUM over the data (the inner sum already runs over them); with [n, D] values it is the mean over n D."""
        if values.shape[0] == 0:
            return float("nan")
        total = ops.multivalid_loss(values, scores, 0).item()
        if values.shape[1] == 1:
            return float(F32(total))
        return float(F32(F32(total) / F32(values.numel())))

    def _cell_errors(self, stats, n, **kwargs) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """multicalibrator.py:21-79: patch = E[score | b] - E[value | b] (0 where b is empty), error = P(b) patch^2."""
        prob_b, ms, mx = _rates(stats, n)
        with np.errstate(all="ignore"):
            ey = (ms / prob_b).astype(F32)
            mv = (mx / prob_b).astype(F32)
            patch = np.where(prob_b > 0, ey - mv, F32(0)).astype(F32)
            err = (prob_b * patch ** 2).astype(F32)
        self._ratio = (ey, mv)
        return err, prob_b, patch

    def _get_patch(self, cell, add_patch, patch_type, **kwargs) -> float:
        if patch_type == "additive":
            return float(add_patch[cell])
        ey, mv = self._ratio
        with np.errstate(all="ignore"):
            return float(F32(ey[cell]) / F32(mv[cell]))  # multicalibrator.py:128-141
