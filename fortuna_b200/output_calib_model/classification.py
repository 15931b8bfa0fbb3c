"""``OutputCalibClassifier`` - fortuna/output_calib_model/classification.py:26-125: temperature scaling of given logits."""
from typing import Callable, Optional

import numpy as np
import torch

from ..loss.classification.focal_loss import focal_loss_fn
from ..prob_model.calibration import ClassificationTemperatureScaler
from .base import OutputCalibManager, OutputCalibModel
from .config.base import Config
from .predictive.classification import ClassificationPredictive

_DEFAULT = object()


class OutputCalibClassifier(OutputCalibModel):
    def __init__(self, output_calibrator=_DEFAULT, seed: int = 0) -> None:
        if output_calibrator is _DEFAULT:
            output_calibrator = ClassificationTemperatureScaler()
        if output_calibrator is not None and not isinstance(output_calibrator, ClassificationTemperatureScaler):
            raise NotImplementedError("output_calibrator must be ClassificationTemperatureScaler() or None")
        self.output_calibrator = output_calibrator
        self.output_calib_manager = OutputCalibManager(output_calibrator=output_calibrator)
        self.predictive = ClassificationPredictive(output_calib_manager=self.output_calib_manager)
        self.prob_output_layer = self.predictive  # mean / mode / ... of the output layer are the S = 1 kernels
        super().__init__(seed=seed)

    def _predict_fn(self, outputs: torch.Tensor, log_temp) -> torch.Tensor:
        return _cls_stat(outputs, "mode", log_temp, self.output_calibrator)

    def calibrate(self, calib_outputs, calib_targets, val_outputs=None, val_targets=None,
                  loss_fn: Callable = focal_loss_fn, config: Optional[Config] = None):
        """classification.py:71-116; the default uncertainty estimates are the calibrated probabilities."""
        config = config or Config()
        self._check_output_dim(calib_outputs, calib_targets)
        if val_outputs is not None:
            self._check_output_dim(val_outputs, val_targets)
        user_fn = config.monitor.uncertainty_fn
        unc = (lambda z, lt: user_fn(_calibrated(z, lt, self.output_calibrator))) if user_fn is not None else (
            lambda z, lt: _cls_stat(z, "mean", lt, self.output_calibrator))
        return super()._calibrate(uncertainty_fn=unc, calib_outputs=calib_outputs, calib_targets=calib_targets,
                                  val_outputs=val_outputs, val_targets=val_targets, loss_fn=loss_fn, config=config)

    @staticmethod
    def _check_output_dim(outputs, targets):
        t = targets.cpu().numpy() if isinstance(targets, torch.Tensor) else np.asarray(targets)
        n_classes = len(np.unique(t))
        if outputs.shape[1] != n_classes:
            raise ValueError(
                f"""`outputs.shape[1]` must be the same as the dimension of the number of classes in `targets`.
                However, `outputs.shape[1]={outputs.shape[1]}` and `len(np.unique(targets))={n_classes}`.""")


def _lt_tensor(log_temp, device, calibrator):
    if calibrator is None:
        return None
    return torch.as_tensor(np.array(log_temp, dtype=np.float32).reshape(1), device=device)


def _cls_stat(z: torch.Tensor, name: str, log_temp, calibrator) -> torch.Tensor:
    from .. import ops

    return ops.bma_reduce_cls(z.unsqueeze(0), (name,), log_temp=_lt_tensor(log_temp, z.device, calibrator))[name]


def _calibrated(z: torch.Tensor, log_temp, calibrator) -> torch.Tensor:
    """Calibrated logits as a new tensor (for a user-supplied ``uncertainty_fn``)."""
    from .. import ops

    out = z.clone()
    lt = _lt_tensor(log_temp, z.device, calibrator)
    if lt is not None:
        ops.softmax_rows_(out, ops.EPI_LOGITS, log_temp=lt)
    return out
