"""``MultivalidMethod`` - fortuna/conformal/multivalid/base.py:11-84: bucket grid, argument checks.  Arrays are device
tensors; the checks read them once."""
from typing import Optional, Union

import numpy as np
import torch

from ...utils.arrays import as_device


def linspace_f32(start, stop, num: int) -> np.ndarray:
    """jnp.linspace(start, stop, num) in float32 (jax/_src/numpy/lax_numpy.py: start (1 - step) + stop step, endpoint set)."""
    f = np.float32
    div = num - 1
    if div <= 0:
        return np.array([f(start)] * num, dtype=f)
    step = (np.arange(div, dtype=f) / f(div)).astype(f)
    out = f(start) * (f(1) - step) + f(stop) * step
    return np.concatenate([out, np.array([f(stop)], dtype=f)]).astype(f)


def to_f32(x, device=None) -> Optional[torch.Tensor]:
    return None if x is None else as_device(x, torch.float32, device)


def np_of(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


class MultivalidMethod:
    def __init__(self, seed: int = 0):
        self._seed = seed
        self._patches = None
        self._n_buckets = None

    @property
    def patches(self):
        return self._patches

    @property
    def n_buckets(self):
        return self._n_buckets

    @n_buckets.setter
    def n_buckets(self, n_buckets):
        self._n_buckets = n_buckets

    @staticmethod
    def _get_buckets(n_buckets: int) -> np.ndarray:
        return linspace_f32(0, 1, n_buckets)

    @staticmethod
    def _maybe_check_values(values, test_values=None):
        if values is not None:
            if np.ndim(values) != 1:
                raise ValueError("`values` must be a 1-dimensional array.")
            v = np_of(values)
            if np.any(v < 0) or np.any(v > 1):
                raise ValueError("All elements in `values` must be within [0, 1].")
        if test_values is not None:
            t = np_of(test_values)
            if np.any(t < 0) or np.any(t > 1):
                raise ValueError("All elements in `test_values` must be within [0, 1].")

    @staticmethod
    def _check_scores(scores):
        if np.ndim(scores) != 1:
            raise ValueError("`scores` must be a 1-dimensional array.")
        s = np_of(scores)
        if np.any(s < 0) or np.any(s > 1):
            raise ValueError("All elements in `scores` must be within [0, 1].")

    @staticmethod
    def _process_scores(scores) -> torch.Tensor:
        s = to_f32(scores).clone()
        return s[:, None].contiguous() if s.ndim == 1 else s

    @staticmethod
    def _maybe_init_min_prob_b(min_prob_b: Union[float, str], n_buckets: int, n_dims: int):
        if min_prob_b == "auto":
            return 1 / (n_buckets * n_dims)
        return min_prob_b
