"""``OneShotBinaryClassificationMulticalibrator`` - fortuna/conformal/multivalid/one_shot/classification/
binary_multicalibrator.py:13-39."""
from typing import Union

from ...iterative.classification.binary_multicalibrator import BinaryClassificationMulticalibrator
from ..multicalibrator import OneShotMulticalibrator


class OneShotBinaryClassificationMulticalibrator(OneShotMulticalibrator):
    _check_scores = staticmethod(BinaryClassificationMulticalibrator._check_scores)
    _maybe_check_values = staticmethod(BinaryClassificationMulticalibrator._maybe_check_values)

    def calibrate(self, targets, probs=None, test_probs=None, n_buckets: int = 100, min_prob_b: Union[float, str] = "auto"):
        return super().calibrate(scores=targets, values=probs, test_values=test_probs, n_buckets=n_buckets,
                                 min_prob_b=min_prob_b)

    def apply_patches(self, probs):
        return super().apply_patches(values=probs)

    def mean_squared_error(self, probs, targets):
        return super().mean_squared_error(values=probs, scores=targets)
