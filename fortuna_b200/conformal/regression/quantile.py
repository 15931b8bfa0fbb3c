"""Conformal quantile of the scores - the part of fortuna/conformal/regression/quantile.py:49-84 (and of
onedim_uncertainty.py:54-90, identical formula) on the hot path: ``jnp.quantile(scores, ceil((n+1)(1-error))/n)``.
``score`` / ``conformal_interval`` are small-n host interval algebra and stay out of scope."""
from typing import Optional

import numpy as np
import torch

from ... import ops
from ...utils.arrays import as_device, device_of


def conformal_quantile_level(n: int, error: float) -> np.float32:
    return np.float32(np.ceil(np.float32((n + 1) * (1 - error))) / np.float32(n))


def conformal_quantile(scores, error: float) -> torch.Tensor:
    if error < 0 or error > 1:
        raise ValueError("""`error` must be a scalar between 0 and 1.""")
    s = as_device(scores, torch.float32, device_of(scores)).clone().view(-1)
    ops.sort_f32_(s)
    q = torch.tensor([conformal_quantile_level(s.numel(), error)], dtype=torch.float32, device=s.device)
    return ops.quantile_sorted(s, q)[0]


class QuantileConformalRegressor:
    def quantile(self, val_lower_bounds=None, val_upper_bounds=None, val_targets=None, error: float = 0.05,
                 scores: Optional[torch.Tensor] = None) -> torch.Tensor:
        if scores is None:
            raise NotImplementedError("pass `scores`: computing them from interval bounds is host algebra outside the hot path")
        return conformal_quantile(scores, error)


class OneDimensionalUncertaintyConformalRegressor:
    def quantile(self, val_preds=None, val_uncertainties=None, val_targets=None, error: float = 0.05,
                 scores: Optional[torch.Tensor] = None) -> torch.Tensor:
        if scores is None:
            raise NotImplementedError("pass `scores`: computing them from predictions is host algebra outside the hot path")
        return conformal_quantile(scores, error)
