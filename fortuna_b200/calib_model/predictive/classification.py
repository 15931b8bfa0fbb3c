"""fortuna/calib_model/predictive/classification.py:8-47."""
import torch

from ...output_calib_model.predictive.classification import ClassificationPredictive as _OutputLevel
from .base import Predictive


class ClassificationPredictive(Predictive):
    def __init__(self, device="cuda"):
        super().__init__(_OutputLevel(), device)

    def entropy(self, inputs_loader, distribute: bool = True) -> torch.Tensor:
        """-sum_c p_c log p_c of softmax(model(x))."""
        return self._out.entropy(self._outputs(inputs_loader), calibrated=False)
