"""``OutputCalibRegressor`` - fortuna/output_calib_model/regression.py:24-122: temperature scaling of the log variances
of given regression outputs [n, 2d] = [mu | log var]."""
from typing import Callable, Optional

import numpy as np
import torch

from ..loss.regression.scaled_mse import scaled_mse_fn
from ..prob_model.calibration import RegressionTemperatureScaler
from .base import OutputCalibManager, OutputCalibModel
from .config.base import Config
from .predictive.regression import RegressionPredictive

_DEFAULT = object()


class OutputCalibRegressor(OutputCalibModel):
    _regression = True

    def __init__(self, output_calibrator=_DEFAULT, seed: int = 0) -> None:
        if output_calibrator is _DEFAULT:
            output_calibrator = RegressionTemperatureScaler()
        if output_calibrator is not None and not isinstance(output_calibrator, RegressionTemperatureScaler):
            raise NotImplementedError("output_calibrator must be RegressionTemperatureScaler() or None")
        self.output_calibrator = output_calibrator
        self.output_calib_manager = OutputCalibManager(output_calibrator=output_calibrator)
        self.predictive = RegressionPredictive(output_calib_manager=self.output_calib_manager)
        self.prob_output_layer = self.predictive
        super().__init__(seed=seed)

    def _stat(self, z: torch.Tensor, name: str, log_temp) -> torch.Tensor:
        from .. import ops

        lt = None if self.output_calibrator is None else torch.as_tensor(
            np.array(log_temp, dtype=np.float32).reshape(1), device=z.device)
        return ops.bma_reduce_reg(z.unsqueeze(0), (name,), log_temp=lt)[name]

    def _predict_fn(self, outputs: torch.Tensor, log_temp) -> torch.Tensor:
        return self._stat(outputs, "mean", log_temp)

    def calibrate(self, calib_outputs, calib_targets, val_outputs=None, val_targets=None,
                  loss_fn: Callable = scaled_mse_fn, config: Optional[Config] = None):
        """regression.py:66-114; the default uncertainty estimates are the calibrated variances."""
        config = config or Config()
        self._check_output_dim(calib_outputs, calib_targets)
        if val_outputs is not None:
            self._check_output_dim(val_outputs, val_targets)
        user_fn = config.monitor.uncertainty_fn
        if user_fn is not None:
            raise NotImplementedError("a custom uncertainty_fn is not supported for regression outputs")
        unc = lambda z, lt: self._stat(z, "aleatoric_variance", lt)
        return super()._calibrate(uncertainty_fn=unc, calib_outputs=calib_outputs, calib_targets=calib_targets,
                                  val_outputs=val_outputs, val_targets=val_targets, loss_fn=loss_fn, config=config)

    @staticmethod
    def _check_output_dim(outputs, targets):
        if outputs.shape[1] != 2 * targets.shape[1]:
            raise ValueError(
                f"""`outputs.shape[1]` must be twice the dimension of the target variables in `targets`, with
                first and second halves corresponding to the mean and log-variance of the likelihood, respectively.
                However, `outputs.shape[1]={outputs.shape[1]}` and `targets.shape[1]={targets.shape[1]}`.""")
