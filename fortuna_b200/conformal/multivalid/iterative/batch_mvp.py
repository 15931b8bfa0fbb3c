"""``BatchMVPConformalMethod`` - fortuna/conformal/multivalid/iterative/batch_mvp.py:17-288 with the pinball loss of
mixins/batchmvp.py:8-27."""
from typing import Dict, Optional, Tuple, Union

import numpy as np
import torch

from .... import ops
from ..base import np_of, to_f32
from .base import F32, IterativeMultivalidMethod


class BatchMVPConformalMethod(IterativeMultivalidMethod):
    _mode = 1  # the statistic is [score <= threshold]

    def __init__(self, seed: int = 0):
        super().__init__(seed=seed)
        self._coverage = None

    def calibrate(self, scores, groups=None, thresholds=None, test_groups=None, test_thresholds=None, atol: float = 1e-4,
                  rtol: float = 1e-6, min_prob_b: Union[float, str] = "auto", n_buckets: int = 100, n_rounds: int = 1000,
                  eta: float = 0.1, split: float = 0.8, bucket_types: Tuple[str, ...] = (">=", "<="),
                  patch_type: str = "additive", coverage: float = 0.95) -> Union[Dict, Tuple[torch.Tensor, Dict]]:
        self._check_coverage(coverage)
        self._coverage = coverage
        return super().calibrate(scores=scores, groups=groups, values=thresholds, test_groups=test_groups,
                                 test_values=test_thresholds, atol=atol, rtol=rtol, min_prob_b=min_prob_b, n_buckets=n_buckets,
                                 n_rounds=n_rounds, eta=eta, split=split, bucket_types=bucket_types, patch_type=patch_type,
                                 coverage=coverage)

    def apply_patches(self, groups=None, thresholds=None) -> torch.Tensor:
        return super().apply_patches(groups=groups, values=thresholds)

    def calibration_error(self, scores, groups=None, thresholds=None, n_buckets: int = 10, **kwargs) -> torch.Tensor:
        return super().calibration_error(scores=scores, groups=groups, values=thresholds, n_buckets=n_buckets,
                                         coverage=self._coverage, **kwargs)

    def pinball_loss(self, values, scores, coverage: float) -> torch.Tensor:
        v, s = to_f32(values), to_f32(scores)
        v = v[:, None].contiguous() if v.ndim == 1 else v
        s = s[:, None].contiguous() if s.ndim == 1 else s
        return torch.tensor(self._loss(v, s, coverage), dtype=torch.float32)

    def _loss(self, values: torch.Tensor, scores: torch.Tensor, coverage: Optional[float] = None) -> float:
        if values.shape[0] == 0:
            return float("nan")
        coverage = self._coverage if coverage is None else coverage
        total = ops.multivalid_loss(values, scores, 1, coverage).item()
        return float(F32(F32(total) / F32(values.shape[0])))

    def _cell_errors(self, stats, n, coverage: float = None, **kwargs):
        """batch_mvp.py:104-192: error = P(b) (coverage - P(score <= threshold | b))^2."""
        with np.errstate(all="ignore"):
            prob_b = (F32(stats[..., 0]) / F32(n)).astype(F32)
            mc = (F32(stats[..., 1]) / F32(n)).astype(F32)
            prob = np.where(prob_b > 0, mc / prob_b, F32(0)).astype(F32)
            err = (prob_b * (F32(coverage) - prob) ** 2).astype(F32)
        return err, prob_b, None

    def _get_patch(self, values, scores, groups, buckets, patch_type, tau, g, v, coverage: float = None, **kwargs) -> float:
        """batch_mvp.py:194-230: the shift delta in {-1, ..., -v_1, 0, v_1, ..., 1} whose shifted thresholds bring the
        coverage inside the chosen set closest to the target; counts of all 2 n_buckets - 1 shifts in one pass."""
        if patch_type == "multiplicative":
            raise NotImplementedError("Please use `patch_type='additive'`.")
        n_buckets = len(buckets)
        deltas = np.concatenate((-buckets[::-1][:-1], buckets)).astype(F32)
        counts = ops.batchmvp_delta_counts(values[:, 0].contiguous(), scores[:, 0].contiguous(), groups, g, tau, v,
                                           n_buckets, torch.from_numpy(deltas).to(values.device)).cpu().numpy()
        n = values.shape[0]
        with np.errstate(all="ignore"):
            prob_b = F32(counts[-1]) / F32(n)
            mc = (F32(counts[:-1]) / F32(n)).astype(F32)
            prob = np.where(prob_b > 0, mc / prob_b, F32(0)).astype(F32)
            errors = ((F32(coverage) - prob) ** 2).astype(F32)
        idx = np.where(errors == errors.min())[0]
        if np.sum(idx >= n_buckets):
            return float(deltas[idx.min()])
        return float(deltas[idx.max()])

    @staticmethod
    def _maybe_check_values(values, test_values=None):
        if values is not None:
            v = np_of(values)
            if np.any(v < 0) or np.any(v > 1):
                raise ValueError("All elements in `thresholds` must be within [0, 1].")

    @staticmethod
    def _check_coverage(coverage: float):
        if coverage < 0 or coverage > 1:
            raise ValueError("`coverage` must be a float between 0 and 1.")
