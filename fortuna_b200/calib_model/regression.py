"""``CalibRegressor`` - fortuna/calib_model/regression.py:25-139.  ``model`` predicts the mean and
``likelihood_log_variance_model`` the log-variance (outputs concatenated to ``[mu | log var]`` as in
fortuna/model/model_manager/regression.py); a single module returning ``[n, 2d]`` is accepted too."""
from typing import Callable, Optional

import numpy as np
import torch

from .. import ops
from ..loss.regression.scaled_mse import scaled_mse_fn
from ..prob_model.regression import _MeanAndLogVar
from .base import CalibModel
from .config.base import Config
from .predictive.regression import RegressionPredictive


def _stat(outputs: torch.Tensor, name: str) -> torch.Tensor:
    return ops.bma_reduce_reg(outputs.unsqueeze(0).contiguous(), (name,))[name]


class CalibRegressor(CalibModel):
    _regression = True

    def __init__(self, model: torch.nn.Module, likelihood_log_variance_model: Optional[torch.nn.Module] = None,
                 model_editor=None, seed: int = 0, device="cuda"):
        if model_editor is not None:
            raise NotImplementedError("model editors are outside the hot path")
        self.predictive = RegressionPredictive(device=device)
        full = model if likelihood_log_variance_model is None else _MeanAndLogVar(model, likelihood_log_variance_model)
        super().__init__(full, seed=seed, device=device)

    def _predict_fn(self, outputs: torch.Tensor) -> torch.Tensor:  # RegressionProbOutputLayer.predict: the mean
        return _stat(outputs, "mean")

    def _check_output_dim(self, data_loader):
        """regression.py:88-107: the model outputs [mu | log var] must be twice as wide as the targets."""
        if data_loader.size == 0:
            raise ValueError("`data_loader` is either empty or incorrectly constructed.")
        for x, y in data_loader:
            self.model.to(self.device)
            with torch.no_grad():
                out_dim = self.model(torch.as_tensor(np.asarray(x[:1]), dtype=torch.float32, device=self.device)).shape[-1]
            y = np.asarray(y)
            d = 1 if y.ndim == 1 else y.shape[-1]
            break
        if out_dim != 2 * d:
            raise ValueError(
                f"""The outputs dimension of both `model` and `likelihood_log_variance_model` must be the same as
                the dimension of the target variables in `data_loader`. However, {out_dim // 2} and {d} were found,
                respectively.""")

    def calibrate(self, calib_data_loader, val_data_loader=None, loss_fn: Callable = scaled_mse_fn,
                  config: Optional[Config] = None):
        """regression.py:109-139; the default uncertainty estimates are the predictive variances."""
        config = config or Config()
        self._check_output_dim(calib_data_loader)
        if val_data_loader is not None:
            self._check_output_dim(val_data_loader)
        user_fn = config.monitor.uncertainty_fn
        unc = user_fn if user_fn is not None else (lambda outputs: _stat(outputs, "aleatoric_variance"))
        return super()._calibrate(calib_data_loader=calib_data_loader, uncertainty_fn=unc, loss_fn=loss_fn,
                                  val_data_loader=val_data_loader, config=config)
