"""``OneShotTopLabelMulticalibrator`` - fortuna/conformal/multivalid/one_shot/classification/
top_label_multicalibrator.py:15-50."""
from typing import Union

from ...iterative.classification.top_label_multicalibrator import TopLabelMulticalibratorMixin
from ..multicalibrator import OneShotMulticalibrator


class OneShotTopLabelMulticalibrator(TopLabelMulticalibratorMixin, OneShotMulticalibrator):
    _top_label = True

    def __init__(self, n_classes: int, seed: int = 0):
        super().__init__(seed=seed)
        self.n_classes = n_classes

    def calibrate(self, targets, probs=None, test_probs=None, n_buckets: int = 100, min_prob_b: Union[float, str] = "auto"):
        return super().calibrate(scores=self._get_scores(targets), values=probs, test_values=test_probs, n_buckets=n_buckets,
                                 min_prob_b=min_prob_b)

    def apply_patches(self, probs):
        return super().apply_patches(values=probs)

    def mean_squared_error(self, probs, targets):
        return super().mean_squared_error(values=probs, scores=self._get_scores(targets))
