"""``BatchMVPConformalClassifier`` - fortuna/conformal/multivalid/iterative/regression/batch_mvp.py:13-69 (the reference
keeps the classifier under ``regression/``)."""
from typing import List

import numpy as np
import torch

from ..batch_mvp import BatchMVPConformalMethod


class BatchMVPConformalClassifier(BatchMVPConformalMethod):
    def __init__(self, seed: int = 0):
        super().__init__(seed=seed)

    @staticmethod
    def conformal_set(class_scores, thresholds) -> List[List[int]]:
        cs = class_scores.detach().cpu().numpy() if isinstance(class_scores, torch.Tensor) else np.asarray(class_scores)
        th = thresholds.detach().cpu().numpy() if isinstance(thresholds, torch.Tensor) else np.asarray(thresholds)
        if cs.ndim != 2:
            raise ValueError("`class_scores` must bse a 2-dimensional array. The first dimension is over the different "
                             "inputs. The second dimension is over all the possible classes.")
        if th.ndim != 1:
            raise ValueError("`thresholds` must be a 1-dimensional array.")
        if cs.shape[0] != th.shape[0]:
            raise ValueError("The first dimension of `class_scores` and `thresholds` must be over the same input data "
                             "points.")
        bools = cs <= th[:, None]
        return [np.nonzero(row)[0].tolist() for row in bools]
