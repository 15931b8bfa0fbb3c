"""``MaxCoverageFixedPrecisionBinaryClassificationCalibrator`` -
fortuna/conformal/classification/maxcovfixprec_binary_classfication.py:13-124 (the file name is the reference's).

Two scalings of a binary classifier's probabilities, tau_pos on p > 1/2 and tau_neg on p < 1/2, chosen on a grid to
maximise coverage subject to true-positive / true-negative precision thresholds.  The reference evaluates its objective
once per grid point over all data (``vmap`` over ``n_taus``); here one kernel pass per side (fb200_maxcov_counts)
returns the exact counts for the whole grid, and the per-grid-point float32 ratios and the arg-max are a 100-element
host computation."""
import logging
from typing import Optional, Union

import numpy as np
import torch

from ... import ops

F32 = np.float32


def _linspace(start, stop, num: int) -> np.ndarray:
    """jnp.linspace in float32: start * (1 - i/div) + stop * (i/div), endpoint appended."""
    start, stop = F32(start), F32(stop)
    if num == 1:
        return np.array([start], F32)
    div = num - 1
    step = (np.arange(div, dtype=F32) / F32(div)).astype(F32)
    return np.concatenate([(start * (F32(1) - step) + stop * step).astype(F32), np.array([stop], F32)])


def _device(x, dtype) -> torch.Tensor:
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(np.asarray(x))
    return x.to(device="cuda", dtype=dtype).contiguous().reshape(-1)


def _values(counts: np.ndarray, labels: np.ndarray, n: int, threshold: float) -> np.ndarray:
    """prob_b * [precision >= threshold] per grid point, float32 like the reference's means (:56-69)."""
    prob_b = counts.astype(F32) / F32(n)
    with np.errstate(invalid="ignore", divide="ignore"):
        prec = (labels.astype(F32) / F32(n)) / prob_b
    return (prob_b * (prec >= F32(threshold)).astype(F32)).astype(F32)


class MaxCoverageFixedPrecisionBinaryClassificationCalibrator:
    def __init__(self):
        self._patches = dict()

    def calibrate(self, targets, probs, true_positive_precision_threshold: float, true_negative_precision_threshold: float,
                  test_probs=None, n_taus: int = 100, margin: float = 0.0) -> Union[None, torch.Tensor]:
        if true_negative_precision_threshold <= 0.5 or true_positive_precision_threshold <= 0.5:
            raise ValueError("Both `true_negative_precision_threshold` and"
                             " `true_positive_precision_threshold` must be greater than 0.5.")
        p, y = _device(probs, torch.float32), _device(targets, torch.int32)
        n = p.numel()
        taus_pos = _linspace(1, 2 * true_positive_precision_threshold, n_taus)
        taus_neg = _linspace(2 * (1 - true_negative_precision_threshold), 1, n_taus)[::-1].copy()
        cp, lp = ops.maxcov_counts(p, y, torch.from_numpy(taus_pos).to(p.device), 0, true_positive_precision_threshold)
        cn, ln = ops.maxcov_counts(p, y, torch.from_numpy(taus_neg).to(p.device), 1, 1 - true_negative_precision_threshold)
        values_pos = _values(cp.cpu().numpy(), lp.cpu().numpy(), n, true_positive_precision_threshold + margin)
        values_neg = _values(cn.cpu().numpy(), ln.cpu().numpy(), n, true_negative_precision_threshold + margin)
        msg = "The {} could not be satisfied. Please consider improving the classifier or decreasing the threshold."
        if values_pos.max() == 0:
            logging.warning(msg.format("`true_positive_precision_threshold`"))
        if values_neg.max() == 0:
            logging.warning(msg.format("`true_negative_precision_threshold`"))
        self._patches["tau_pos"] = taus_pos[int(np.argmax(values_pos))]
        self._patches["tau_neg"] = taus_neg[int(np.argmax(values_neg))]
        self._values = (values_pos, values_neg)
        if test_probs is not None:
            return self.apply_patches(test_probs)

    def apply_patches(self, probs) -> torch.Tensor:
        if not len(self._patches):
            logging.warning("No patches available.")
            return probs
        shape = tuple(probs.shape)
        return ops.maxcov_apply_patches(_device(probs, torch.float32), self._patches["tau_pos"],
                                        self._patches["tau_neg"]).reshape(shape)

    # ------------------------------------------------------------------ precision / coverage of given probabilities
    @staticmethod
    def _count(probs, targets, side: int, threshold: float):
        p = _device(probs, torch.float32)
        y = _device(targets, torch.int32) if targets is not None else torch.zeros(p.numel(), dtype=torch.int32, device=p.device)
        one = torch.ones(1, dtype=torch.float32, device=p.device)  # tau = 1 leaves the probabilities unchanged
        c, l = ops.maxcov_counts(p, y, one, side, threshold)
        return int(c.item()), int(l.item()), p.numel()

    @classmethod
    def true_positive_precision(cls, probs, targets, threshold: float):
        c, l, n = cls._count(probs, targets, 0, threshold)
        with np.errstate(invalid="ignore", divide="ignore"):
            return (F32(l) / F32(n)) / (F32(c) / F32(n))

    @classmethod
    def true_negative_precision(cls, probs, targets, threshold: float):
        c, l, n = cls._count(probs, targets, 1, 1 - threshold)
        with np.errstate(invalid="ignore", divide="ignore"):
            return (F32(l) / F32(n)) / (F32(c) / F32(n))

    @classmethod
    def true_positive_coverage(cls, probs, threshold: float):
        c, _, n = cls._count(probs, None, 0, threshold)
        return F32(c) / F32(n)

    @classmethod
    def true_negative_coverage(cls, probs, threshold: float):
        c, _, n = cls._count(probs, None, 1, threshold)
        return F32(c) / F32(n)

    @property
    def patches(self):
        return self._patches
