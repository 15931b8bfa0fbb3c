"""Simple conformal scores 1 - p[target] - fortuna/conformal/classification/simple_prediction.py:10-35."""
import torch

from ... import ops
from ...utils.arrays import as_device, device_of
from .base import CVPlusConformalClassifier, SplitConformalClassifier


def score_fn(probs, targets):
    dev = device_of(probs, targets)
    return ops.simple_scores(as_device(probs, torch.float32, dev), as_device(targets, torch.int32, dev))


def all_class_scores(probs):
    return ops.simple_scores(as_device(probs, torch.float32, device_of(probs)), None)


class SimplePredictionConformalClassifier(SplitConformalClassifier):
    def score_fn(self, probs, targets):
        return score_fn(probs, targets)

    def _all_class_scores(self, probs):
        return all_class_scores(probs)


class CVPlusSimplePredictionConformalClassifier(CVPlusConformalClassifier):
    def score_fn(self, probs, targets):
        return score_fn(probs, targets)

    def _all_class_scores(self, probs):
        return all_class_scores(probs)
