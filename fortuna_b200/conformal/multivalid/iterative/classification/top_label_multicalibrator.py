"""``TopLabelMulticalibrator`` - fortuna/conformal/multivalid/iterative/classification/top_label_multicalibrator.py:19-133
with mixins/classification/top_label_multicalibrator.py:9-74: one score column per class, a set also requires the class
to be the arg-max of the probabilities."""
from typing import Dict, Optional, Tuple, Union

import numpy as np
import torch

from ..... import ops
from ...base import np_of, to_f32
from ..multicalibrator import Multicalibrator


class TopLabelMulticalibratorMixin:
    def _get_scores(self, targets) -> torch.Tensor:
        """one-hot [n, n_classes] of integer targets (mixins/...:23-28)."""
        t = np_of(targets)
        if t.ndim != 1:
            raise ValueError("`targets` must be a 1-dimensional array of integers.")
        if t.dtype not in [np.int32, np.int64]:
            raise ValueError("All elements in `targets` must be integers")
        oh = (t[:, None] == np.arange(self.n_classes)[None]).astype(np.float32)
        return torch.from_numpy(oh).cuda()

    @staticmethod
    def _check_scores(scores):
        pass

    @staticmethod
    def _maybe_check_values(values, test_values=None):
        if values is not None:
            if np.ndim(values) != 2:
                raise ValueError("`probs` must be a 2-dimensional array representing the probabilities for each input and "
                                 "each class.")
            v = np_of(values)
            if np.any(v < 0) or np.any(v > 1):
                raise ValueError("All elements in `values` must be within [0, 1].")
        if test_values is not None:
            if np.ndim(test_values) != 2:
                raise ValueError("`test_probs` must be a 2-dimensional array representing the probabilities for each input "
                                 "and each class.")
            t = np_of(test_values)
            if np.any(t < 0) or np.any(t > 1):
                raise ValueError("All elements in `test_values` must be within [0, 1].")


class TopLabelMulticalibrator(TopLabelMulticalibratorMixin, Multicalibrator):
    _top_label = True

    def __init__(self, n_classes: int, seed: int = 0):
        super().__init__(seed=seed)
        self.n_classes = n_classes

    def calibrate(self, targets, groups=None, probs=None, test_groups=None, test_probs=None, atol: float = 1e-4,
                  rtol: float = 1e-6, min_prob_b: Union[float, str] = "auto", n_buckets: int = 100, n_rounds: int = 1000,
                  eta: float = 0.1, split: float = 0.8, bucket_types: Tuple[str, ...] = ("<=", ">="),
                  patch_type: str = "additive", **kwargs) -> Union[Dict, Tuple[torch.Tensor, Dict]]:
        return super().calibrate(scores=self._get_scores(targets), groups=groups, values=probs, test_groups=test_groups,
                                 test_values=test_probs, atol=atol, rtol=rtol, min_prob_b=min_prob_b, n_buckets=n_buckets,
                                 n_rounds=n_rounds, eta=eta, split=split, bucket_types=bucket_types, patch_type=patch_type,
                                 **kwargs)

    def apply_patches(self, groups=None, probs=None) -> torch.Tensor:
        return super().apply_patches(groups=groups, values=probs)

    def calibration_error(self, targets, groups=None, probs=None, n_buckets: int = 10, **kwargs) -> torch.Tensor:
        return super().calibration_error(scores=self._get_scores(targets), groups=groups, values=probs, n_buckets=n_buckets,
                                         **kwargs)

    def mean_squared_error(self, probs, targets) -> torch.Tensor:
        return super().mean_squared_error(values=probs, scores=self._get_scores(targets))

    def _maybe_init_values(self, values, size: Optional[int], device) -> torch.Tensor:
        """top_label_multicalibrator.py:112-133: uniform probabilities jittered by 0.01 N(0, 1) drawn with jax's own normal
        stream from PRNGKey(seed), reflected into [0, 1] and normalised."""
        if values is not None:
            return to_f32(values, device).clone()
        if size is None:
            raise ValueError("If `values` is not provided, `size` must be provided.")
        f = np.float32
        z = ops.random_normal(ops.PRNGKey(self._seed, device), size * self.n_classes).cpu().numpy().reshape(size, self.n_classes)
        v = (f(1 / self.n_classes) * np.ones((size, self.n_classes), f) + f(0.01) * z).astype(f)
        v = np.abs(v)
        v = (v - f(2) * np.maximum(f(0), v - f(1))).astype(f)
        v = (v / v.sum(1, keepdims=True, dtype=f)).astype(f)
        return torch.from_numpy(v).to(device)
