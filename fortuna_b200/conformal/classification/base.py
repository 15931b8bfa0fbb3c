"""Conformal classifiers - same classes and methods as fortuna/conformal/classification/base.py:14-132.

The reference materialises an ``[n_val, B, C]`` boolean tensor for the rank count and re-sorts the test
probabilities once per class for the all-class scores.  Here ``get_scores`` is one sort per row
(fb200_aps_scores / fb200_simple_scores) and the sets come from one compare per element against the validation
scores sorted once on the device (fb200_sort_f32 + fb200_conformal_sets); only the ragged python lists the public
API returns are built on the host, from the uint8 mask."""
import abc
from typing import List, Tuple

import numpy as np
import torch

from ... import ops
from ...utils.arrays import as_device, device_of


def conformal_threshold(n_val: int, error: float) -> np.float32:
    """float32((1 - error) * (n_val + 1)): right-hand side of base.py:51-53 under jax weak typing."""
    return np.float32((1 - error) * (n_val + 1))


def conformal_mask_from_scores(val_scores, test_scores, error: float):
    """Device-side core of ``_get_conformal_sets_from_scores``: (mask uint8 [B,C], sizes int32 [B])."""
    dev = device_of(val_scores, test_scores)
    val = as_device(val_scores, torch.float32, dev).clone().view(-1)
    test = as_device(test_scores, torch.float32, dev)
    if test.dim() != 2:
        raise ValueError("`test_scores` must be a two-dimensional array [n_test, n_classes].")
    ops.sort_f32_(val)
    return ops.conformal_sets(val, test, conformal_threshold(val.numel(), error))


def sets_from_mask(mask: torch.Tensor) -> List[List[int]]:
    m = mask.detach().cpu().numpy().astype(bool)
    return [np.nonzero(row)[0].tolist() for row in m]


class ConformalClassifier:
    def is_in(self, values, conformal_sets: List) -> np.ndarray:
        vals = values.tolist() if hasattr(values, "tolist") else list(values)
        return np.array([v in s for v, s in zip(vals, conformal_sets)])

    @abc.abstractmethod
    def score_fn(self, probs, targets):
        pass

    def _all_class_scores(self, probs):
        """[B, C] scores of every class of every row (``targets=None`` in the kernels)."""
        raise NotImplementedError

    @staticmethod
    def _get_conformal_sets_from_scores(val_scores, test_scores, error: float) -> List[List[int]]:
        mask, _ = conformal_mask_from_scores(val_scores, test_scores, error)
        return sets_from_mask(mask)

    @abc.abstractmethod
    def get_scores(self, *args, **kwargs) -> Tuple[torch.Tensor, torch.Tensor]:
        pass


class SplitConformalClassifier(ConformalClassifier, abc.ABC):
    def get_scores(self, val_probs, val_targets, test_probs) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.score_fn(val_probs, val_targets), self._all_class_scores(test_probs)

    def conformal_set(self, val_probs, val_targets, test_probs, error: float) -> List[List[int]]:
        val_scores, test_scores = self.get_scores(val_probs=val_probs, val_targets=val_targets, test_probs=test_probs)
        return self._get_conformal_sets_from_scores(val_scores=val_scores, test_scores=test_scores, error=error)


class CVPlusConformalClassifier(ConformalClassifier):
    def get_scores(self, cross_val_probs: List, cross_val_targets: List, cross_test_probs: List):
        vs = [self.score_fn(p, t) for p, t in zip(cross_val_probs, cross_val_targets)]
        ts = [self._all_class_scores(p) for p in cross_test_probs]
        return torch.cat(vs), torch.cat(ts, dim=0)

    def conformal_set(self, cross_val_probs: List, cross_val_targets: List, cross_test_probs: List, error: float):
        val_scores, test_scores = self.get_scores(cross_val_probs=cross_val_probs, cross_val_targets=cross_val_targets,
                                                  cross_test_probs=cross_test_probs)
        return self._get_conformal_sets_from_scores(val_scores=val_scores, test_scores=test_scores, error=error)
