"""Adaptive-prediction-set scores - fortuna/conformal/classification/adaptive_prediction.py:11-39."""
import torch

from ... import ops
from ...utils.arrays import as_device, device_of
from .base import CVPlusConformalClassifier, SplitConformalClassifier


def score_fn(probs, targets):
    """cumsum(probs[perm])[inv_perm[target]] per row: the probability mass ranked at or before the target."""
    dev = device_of(probs, targets)
    return ops.aps_scores(as_device(probs, torch.float32, dev), as_device(targets, torch.int32, dev))


def all_class_scores(probs):
    """The [B, C] matrix the reference builds with one score_fn call per class (base.py:76-81)."""
    return ops.aps_scores(as_device(probs, torch.float32, device_of(probs)), None)


class AdaptivePredictionConformalClassifier(SplitConformalClassifier):
    def score_fn(self, probs, targets):
        return score_fn(probs, targets)

    def _all_class_scores(self, probs):
        return all_class_scores(probs)


class CVPlusAdaptivePredictionConformalClassifier(CVPlusConformalClassifier):
    def score_fn(self, probs, targets):
        return score_fn(probs, targets)

    def _all_class_scores(self, probs):
        return all_class_scores(probs)
