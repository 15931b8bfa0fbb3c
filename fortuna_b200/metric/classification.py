"""Classification metrics - same functions and argument checks as fortuna/metric/classification.py:16-220.

Everything is one launch of fb200_ece_bins (max / arg-max per row, ten-bin assignment, per-bin counters, Brier
term) instead of ten host-driven ``where`` scans; results are 0-d / 1-d CUDA tensors."""
from typing import Optional, Tuple

import numpy as np
import torch

from .. import ops
from ..utils.arrays import as_device, device_of

N_BINS = 10


def linspace_thresholds(n_data: int, n_bins: int = N_BINS) -> np.ndarray:
    """jnp.linspace(1 / n_data, 1, n_bins) in float32 (metric/classification.py:76): start*(1-i/(n-1)) + stop*i/(n-1),
    last point exactly `stop`."""
    start, stop = np.float32(1 / n_data), np.float32(1)
    step = (np.arange(n_bins - 1, dtype=np.float32) / np.float32(n_bins - 1)).astype(np.float32)
    out = start * (np.float32(1) - step) + stop * step
    return np.concatenate([out, np.array([stop], dtype=np.float32)]).astype(np.float32)


def _bins(preds, probs, targets):
    dev = device_of(probs, preds, targets)
    probs = as_device(probs, torch.float32, dev)
    preds = as_device(preds, torch.int32, dev)
    targets = as_device(targets, torch.int32, dev)
    thr = torch.from_numpy(linspace_thresholds(probs.shape[0])).to(dev)
    return ops.ece_bins(probs, targets, thr, preds=preds), probs.shape[0]


def accuracy(preds, targets) -> torch.Tensor:
    if np.ndim(preds) > 1 if not isinstance(preds, torch.Tensor) else preds.dim() > 1:
        raise ValueError("""`preds` must be a one-dimensional array of predicted classes.""")
    if np.ndim(targets) > 1 if not isinstance(targets, torch.Tensor) else targets.dim() > 1:
        raise ValueError("""`targets` must be a one-dimensional array of target classes.""")
    dev = device_of(preds, targets)
    preds = as_device(preds, torch.int32, dev)
    conf = torch.ones(preds.shape[0], dtype=torch.float32, device=dev)  # confidences are irrelevant to accuracy
    thr = torch.from_numpy(linspace_thresholds(preds.shape[0])).to(dev)
    return ops.ece_bins(conf, as_device(targets, torch.int32, dev), thr, preds=preds)["result"][2]


def compute_counts_confs_accs(preds, probs, targets, plot: bool = False, plot_options: Optional[dict] = None
                              ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(counts int32[10], confs f32[10], accs f32[10]); empty bins give 0 like nan_to_num(mean([]))."""
    if plot:
        raise NotImplementedError("reliability-diagram plotting is outside the hot path (fortuna/plot.py)")
    e, _ = _bins(preds, probs, targets)
    return e["counts"], e["confs"], e["accs"]


def expected_calibration_error(preds, probs, targets, plot: bool = False, plot_options: Optional[dict] = None):
    if plot:
        raise NotImplementedError("reliability-diagram plotting is outside the hot path (fortuna/plot.py)")
    return _bins(preds, probs, targets)[0]["result"][0]


def ece(preds, probs, targets, plot: bool = False, plot_options: Optional[dict] = None):
    return expected_calibration_error(preds, probs, targets, plot, plot_options)


def maximum_calibration_error(preds, probs, targets, plot: bool = False, plot_options: Optional[dict] = None):
    if plot:
        raise NotImplementedError("reliability-diagram plotting is outside the hot path (fortuna/plot.py)")
    return _bins(preds, probs, targets)[0]["result"][1]


def mce(preds, probs, targets, plot: bool = False, plot_options: Optional[dict] = None):
    return maximum_calibration_error(preds, probs, targets, plot, plot_options)


def brier_score(probs, targets) -> torch.Tensor:
    nd = probs.dim() if isinstance(probs, torch.Tensor) else np.ndim(probs)
    if nd > 2:
        raise ValueError("`probs` can be at most 2 dimensional.")
    if nd < 2:
        raise NotImplementedError("binary (one-dimensional) Brier score is outside the hot path")
    dev = device_of(probs, targets)
    probs = as_device(probs, torch.float32, dev)
    thr = torch.from_numpy(linspace_thresholds(probs.shape[0])).to(dev)
    return ops.ece_bins(probs, as_device(targets, torch.int32, dev), thr)["result"][3]
