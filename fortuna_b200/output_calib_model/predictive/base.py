"""Predictive distribution p(y | calibrated output) of the output calibration models -
fortuna/output_calib_model/predictive/base.py:23-309.  Every method takes the model OUTPUTS (no network, no posterior
samples): the S = 1 case of the maps the BMA reduction kernels (K3 / K4) compute, with the temperature of the
calibration state applied inside those kernels."""
from typing import Optional

import numpy as np
import torch

from ...utils.random import WithRNG


class Predictive(WithRNG):
    def __init__(self, output_calib_manager=None, prob_output_layer=None):
        self.output_calib_manager = output_calib_manager
        self.prob_output_layer = prob_output_layer
        self.state = None

    # ------------------------------------------------------------------ plumbing
    @staticmethod
    def _device_f32(x) -> torch.Tensor:
        if not isinstance(x, torch.Tensor):
            x = torch.as_tensor(np.asarray(x))
        return x.to(device="cuda", dtype=torch.float32).contiguous()

    def _check_calibrated(self) -> None:
        if self.state is None:
            raise ValueError("No calibration state was found. The model must be calibrated beforehand.")

    def _log_temp(self, calibrated: bool, device) -> Optional[torch.Tensor]:
        """float32[1] on the device, or None for uncalibrated outputs / an identity calibrator."""
        if not calibrated:
            return None
        self._check_calibrated()
        if self.output_calib_manager is None or self.output_calib_manager.output_calibrator is None:
            return None
        lt = self.state.get().log_temp
        return torch.as_tensor(np.array(lt, dtype=np.float32).reshape(1), device=device)

    def _key(self, rng):
        return self.rng.get() if rng is None else rng
