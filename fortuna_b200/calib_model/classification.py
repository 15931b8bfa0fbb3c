"""``CalibClassifier`` - fortuna/calib_model/classification.py:34-193.  ``model`` is a torch ``nn.Module`` returning logits
(model definitions are outside the hot path)."""
from typing import Callable, Optional

import numpy as np
import torch

from .. import ops
from ..loss.classification.focal_loss import focal_loss_fn
from .base import CalibModel
from .config.base import Config
from .predictive.classification import ClassificationPredictive


def _stat(outputs: torch.Tensor, name: str) -> torch.Tensor:
    return ops.bma_reduce_cls(outputs.unsqueeze(0).contiguous(), (name,))[name]


class CalibClassifier(CalibModel):
    def __init__(self, model: torch.nn.Module, model_editor=None, seed: int = 0, device="cuda"):
        if model_editor is not None:
            raise NotImplementedError("model editors are outside the hot path")
        self.predictive = ClassificationPredictive(device=device)
        super().__init__(model, seed=seed, device=device)

    def _predict_fn(self, outputs: torch.Tensor) -> torch.Tensor:  # ClassificationProbOutputLayer.predict: the mode
        return _stat(outputs, "mode")

    def _check_output_dim(self, data_loader):
        """classification.py:120-144: the model's output width must equal the number of classes in the targets."""
        if data_loader.size == 0:
            raise ValueError("`data_loader` is either empty or incorrectly constructed.")
        for x, y in data_loader:
            self.model.to(self.device)
            with torch.no_grad():
                out_dim = self.model(torch.as_tensor(np.asarray(x[:1]), dtype=torch.float32, device=self.device)).shape[-1]
            break
        n_classes = len(np.unique(data_loader.to_array_targets()))
        if out_dim != n_classes:
            raise ValueError(
                f"""The outputs dimension of `model` must be the same as the number of classes in the target variables
                of `data_loader`. However, {out_dim} and {n_classes} were found, respectively.""")

    def calibrate(self, calib_data_loader, val_data_loader=None, loss_fn: Callable = focal_loss_fn,
                  config: Optional[Config] = None):
        """classification.py:146-193; the default uncertainty estimates are the predictive probabilities."""
        config = config or Config()
        self._check_output_dim(calib_data_loader)
        if val_data_loader is not None:
            self._check_output_dim(val_data_loader)
        user_fn = config.monitor.uncertainty_fn
        unc = user_fn if user_fn is not None else (lambda outputs: _stat(outputs, "mean"))
        return super()._calibrate(calib_data_loader=calib_data_loader, uncertainty_fn=unc, loss_fn=loss_fn,
                                  val_data_loader=val_data_loader, config=config)
