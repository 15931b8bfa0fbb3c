"""``OneShotMulticalibrator`` - fortuna/conformal/multivalid/one_shot/multicalibrator.py:8-26."""
import numpy as np

from ..iterative.multicalibrator import Multicalibrator
from .base import F32, OneShotMultivalidMethod


class OneShotMulticalibrator(OneShotMultivalidMethod):
    def mean_squared_error(self, values, scores):
        return Multicalibrator.mean_squared_error(Multicalibrator(), values, scores)

    def _get_patches(self, stats, n, buckets, min_prob_b) -> np.ndarray:
        """patch[v, c] = E[score | |value - v| < h] where that set has probability above ``min_prob_b``, else v."""
        with np.errstate(all="ignore"):
            prob_b = (F32(stats[..., 0]) / F32(n)).astype(F32)
            ms = (F32(stats[..., 1]) / F32(n)).astype(F32)
            return np.where(prob_b > F32(min_prob_b), ms / prob_b, buckets[:, None].astype(F32)).astype(F32)
