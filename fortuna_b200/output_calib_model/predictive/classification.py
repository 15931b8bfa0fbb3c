"""``ClassificationPredictive`` of the output calibration models -
fortuna/output_calib_model/predictive/classification.py:10-109 over predictive/base.py:38-254 and
``ClassificationProbOutputLayer`` (fortuna/prob_output_layer/classification.py:15-148).  ``outputs`` are logits [n, C]."""
from typing import Optional

import numpy as np
import torch

from ... import ops
from .base import Predictive


class ClassificationPredictive(Predictive):
    def _reduce(self, outputs, name: str, calibrated: bool, targets=None) -> torch.Tensor:
        if calibrated:
            self._check_calibrated()
        z = self._device_f32(outputs)
        t = None
        if targets is not None:
            t = torch.as_tensor(np.asarray(targets) if not isinstance(targets, torch.Tensor) else targets).to(
                device=z.device, dtype=torch.int32).contiguous()
        return ops.bma_reduce_cls(z.unsqueeze(0), (name,), targets=t, log_temp=self._log_temp(calibrated, z.device))[name]

    def log_prob(self, outputs, targets, calibrated: bool = True, **kwargs) -> torch.Tensor:
        return self._reduce(outputs, "log_prob", calibrated, targets)

    def mean(self, outputs, calibrated: bool = True, **kwargs) -> torch.Tensor:
        """softmax of the (calibrated) logits."""
        return self._reduce(outputs, "mean", calibrated)

    def mode(self, outputs, calibrated: bool = True, **kwargs) -> torch.Tensor:
        return self._reduce(outputs, "mode", calibrated)

    def variance(self, outputs, calibrated: bool = True, **kwargs) -> torch.Tensor:
        """p (1 - p) of the one-hot encoded target (prob_output_layer/classification.py:87-103)."""
        return self._reduce(outputs, "aleatoric_variance", calibrated)

    def std(self, outputs, variances: Optional[torch.Tensor] = None, calibrated: bool = True) -> torch.Tensor:
        if variances is not None:
            return torch.sqrt(torch.as_tensor(variances))
        return self._reduce(outputs, "std", calibrated)

    def entropy(self, outputs, calibrated: bool = True, **kwargs) -> torch.Tensor:
        """-sum_c p_c log p_c (prob_output_layer/classification.py:124-148)."""
        return self._reduce(outputs, "entropy", calibrated)

    def sample(self, n_samples: int, outputs, rng=None, calibrated: bool = True, **kwargs) -> torch.Tensor:
        """int32 [n_samples, n]: ``keys = split(rng, n); choice(keys[b], C, p=softmax(outputs[b]), shape=(n_samples,))``
        with jax's own uniforms and prefix-sum order (fb200_cls_output_sample_targets)."""
        if calibrated:
            self._check_calibrated()
        z = self._device_f32(outputs)
        return ops.cls_output_sample_targets(z, self._key(rng), n_samples, self._log_temp(calibrated, z.device))
