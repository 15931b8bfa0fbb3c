"""``BinaryClassificationMulticalibrator`` - fortuna/conformal/multivalid/iterative/classification/
binary_multicalibrator.py:20-103 with the checks of mixins/classification/binary_multicalibrator.py:9-41."""
from typing import Dict, Tuple, Union

import numpy as np
import torch

from ...base import np_of
from ..multicalibrator import Multicalibrator


class BinaryClassificationMulticalibrator(Multicalibrator):
    def calibrate(self, targets, groups=None, probs=None, test_groups=None, test_probs=None, atol: float = 1e-4,
                  rtol: float = 1e-6, min_prob_b: Union[float, str] = "auto", n_buckets: int = 100, n_rounds: int = 1000,
                  eta: float = 0.1, split: float = 0.8, bucket_types: Tuple[str, ...] = ("<=", ">="),
                  patch_type: str = "additive", **kwargs) -> Union[Dict, Tuple[torch.Tensor, Dict]]:
        return super().calibrate(scores=targets, groups=groups, values=probs, test_groups=test_groups, test_values=test_probs,
                                 atol=atol, rtol=rtol, min_prob_b=min_prob_b, n_buckets=n_buckets, n_rounds=n_rounds, eta=eta,
                                 split=split, bucket_types=bucket_types, patch_type=patch_type, **kwargs)

    def apply_patches(self, groups=None, probs=None) -> torch.Tensor:
        return super().apply_patches(groups=groups, values=probs)

    def calibration_error(self, targets, groups=None, probs=None, n_buckets: int = 10, **kwargs) -> torch.Tensor:
        return super().calibration_error(scores=targets, groups=groups, values=probs, n_buckets=n_buckets, **kwargs)

    def mean_squared_error(self, probs, targets) -> torch.Tensor:
        return super().mean_squared_error(values=probs, scores=targets)

    @staticmethod
    def _maybe_check_values(values, test_values=None):
        if values is not None:
            if np.ndim(values) != 1:
                raise ValueError("`probs` must be a 1-dimensional array representing the probability that the target "
                                 "variable is 1.")
            v = np_of(values)
            if np.any(v < 0) or np.any(v > 1):
                raise ValueError("All elements in `values` must be within [0, 1].")
        if test_values is not None:
            if np.ndim(test_values) != 1:
                raise ValueError("`test_probs` must be a 1-dimensional array representing the probability that the target "
                                 "variable is 1.")
            t = np_of(test_values)
            if np.any(t < 0) or np.any(t > 1):
                raise ValueError("All elements in `test_values` must be within [0, 1].")

    @staticmethod
    def _check_scores(scores):
        s = np_of(scores)
        if s.ndim != 1:
            raise ValueError("`targets` must be a 1-dimensional array of integers.")
        if set(np.unique(s).tolist()) != {0, 1}:
            raise ValueError("All values in `targets` must be 0 or 1.")
        if s.dtype not in [np.int32, np.int64]:
            raise ValueError("All elements in `targets` must be integers")
