"""fortuna/calib_model/predictive/regression.py:16-166."""
from typing import List, Optional, Union

import numpy as np
import torch

from ...output_calib_model.predictive.regression import RegressionPredictive as _OutputLevel
from .base import Predictive


class RegressionPredictive(Predictive):
    def __init__(self, device="cuda"):
        super().__init__(_OutputLevel(), device)

    def entropy(self, inputs_loader, n_samples: int = 30, rng=None, distribute: bool = True) -> torch.Tensor:
        self._out.rng = self.rng
        return self._out.entropy(self._outputs(inputs_loader), calibrated=False, n_target_samples=n_samples, rng=rng)

    def quantile(self, q: Union[float, List, np.ndarray], inputs_loader, n_samples: Optional[int] = 30, rng=None,
                 distribute: bool = True) -> torch.Tensor:
        self._out.rng = self.rng
        return self._out.quantile(q, self._outputs(inputs_loader), n_samples, rng, calibrated=False)

    def credible_interval(self, inputs_loader, n_samples: int = 30, error: float = 0.05, interval_type: str = "two-tailed",
                          rng=None, distribute: bool = True) -> torch.Tensor:
        self._out.rng = self.rng
        return self._out.credible_interval(self._outputs(inputs_loader), n_samples, error, interval_type, rng,
                                           calibrated=False)
