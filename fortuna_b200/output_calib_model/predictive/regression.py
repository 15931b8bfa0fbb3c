"""``RegressionPredictive`` of the output calibration models -
fortuna/output_calib_model/predictive/regression.py:16-112 over predictive/base.py:38-254 and
``RegressionProbOutputLayer`` (fortuna/prob_output_layer/regression.py:18-251).  ``outputs`` are [n, 2d] = [mu | log var];
the calibrated log variance is log var + log_temp."""
from typing import List, Optional, Union

import numpy as np
import torch

from ... import ops
from .base import Predictive


class RegressionPredictive(Predictive):
    def _reduce(self, outputs, name: str, calibrated: bool, targets=None) -> torch.Tensor:
        if calibrated:
            self._check_calibrated()
        z = self._device_f32(outputs)
        t = None
        if targets is not None:
            t = self._device_f32(targets).reshape(z.shape[0], -1).contiguous()
        return ops.bma_reduce_reg(z.unsqueeze(0), (name,), targets=t, log_temp=self._log_temp(calibrated, z.device))[name]

    def log_prob(self, outputs, targets, calibrated: bool = True, **kwargs) -> torch.Tensor:
        return self._reduce(outputs, "log_prob", calibrated, targets)

    def mean(self, outputs, calibrated: bool = True, **kwargs) -> torch.Tensor:
        return self._reduce(outputs, "mean", calibrated)

    def mode(self, outputs, calibrated: bool = True, **kwargs) -> torch.Tensor:
        return self._reduce(outputs, "mean", calibrated)

    def variance(self, outputs, calibrated: bool = True, **kwargs) -> torch.Tensor:
        return self._reduce(outputs, "aleatoric_variance", calibrated)

    def std(self, outputs, variances: Optional[torch.Tensor] = None, calibrated: bool = True) -> torch.Tensor:
        if variances is not None:
            return torch.sqrt(torch.as_tensor(variances))
        return self._reduce(outputs, "std", calibrated)

    def sample(self, n_samples: int, outputs, rng=None, calibrated: bool = True, **kwargs) -> torch.Tensor:
        """float32 [n_samples, n, d] = mu + exp(0.5 log var) * normal(rng, (n_samples, n, d)), jax's own normals."""
        if calibrated:
            self._check_calibrated()
        z = self._device_f32(outputs)
        return ops.reg_output_sample_targets(z, self._key(rng), n_samples, self._log_temp(calibrated, z.device))

    def entropy(self, outputs, calibrated: bool = True, n_target_samples: int = 30, rng=None, **kwargs) -> torch.Tensor:
        """Monte-Carlo -mean_t log N(y_t; outputs) over the target samples above (prob_output_layer/regression.py:
        174-203): the S = 1 case of fb200_reg_mc_entropies, samples drawn in registers."""
        if calibrated:
            self._check_calibrated()
        z = self._device_f32(outputs)
        key = self._key(rng)
        idx = torch.zeros(1, dtype=torch.int32, device=z.device)
        return ops.reg_mc_entropies(z.unsqueeze(0), idx, key.reshape(1, 2).contiguous(), n_target_samples,
                                    names=("aleatoric_entropy",), log_temp=self._log_temp(calibrated, z.device))["aleatoric_entropy"]

    def quantile(self, q: Union[float, List, np.ndarray], outputs, n_samples: Optional[int] = 30, rng=None,
                 calibrated: bool = True) -> torch.Tensor:
        samples = self.sample(n_samples, outputs, rng=rng, calibrated=calibrated)
        scalar = np.ndim(q) == 0
        qq = torch.as_tensor(np.atleast_1d(np.asarray(q, dtype=np.float32)), device=samples.device)
        out = ops.quantile_axis0(samples, qq)
        return out[0] if scalar else out

    def credible_interval(self, outputs, n_samples: int = 30, error: float = 0.05, interval_type: str = "two-tailed",
                          rng=None, calibrated: bool = True) -> torch.Tensor:
        supported_types = ["two-tailed", "right-tailed", "left-tailed"]
        if interval_type not in supported_types:
            raise ValueError("`type={}` not recognised. Please choose among the following supported types: {}.".format(
                interval_type, supported_types))
        q = [0.5 * error, 1 - 0.5 * error] if interval_type == "two-tailed" else (
            error if interval_type == "left-tailed" else 1 - error)
        qq = self.quantile(q, outputs, n_samples, rng, calibrated)
        if qq.shape[-1] != 1:
            raise ValueError("""Credibility intervals are only supported for scalar target variables.""")
        if interval_type == "two-tailed":
            return qq.squeeze(2).transpose(0, 1).contiguous()  # [n, 2]: (lower, upper)
        return qq
