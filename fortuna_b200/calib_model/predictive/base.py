"""Predictive distribution of a calibration model - fortuna/calib_model/predictive/base.py:14-250.  Every method walks the
loader through the model (the backbone, outside the hot path: torch forward under no_grad, batch by batch with the state
held by the calibrated module), concatenates the outputs like ``Likelihood.get_calibrated_outputs``
(fortuna/likelihood/base.py:399-458, no output calibrator here: ``output_calib_manager=None``,
calib_model/classification.py:52-56) and applies the S = 1 case of the predictive kernels to them."""
from typing import Optional

import numpy as np
import torch

from ...utils.random import WithRNG


class Predictive(WithRNG):
    def __init__(self, output_predictive, device):
        self._out = output_predictive  # the output-level maps (S = 1 kernels), uncalibrated outputs
        self.device = torch.device(device)
        self.state = None              # CalibStateRepository once calibrated / loaded

    def _check_state(self):
        if self.state is None:
            raise ValueError("No state available. You must first either calibrate the model, or load a saved checkpoint.")

    def _outputs(self, inputs_loader) -> torch.Tensor:
        """float32 [n, out_dim] on the device: the model applied to every batch of the loader."""
        self._check_state()
        model = self.state.get().model
        was_training = model.training
        model.eval()
        outs = []
        with torch.no_grad():
            for x in inputs_loader:
                xt = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))
                outs.append(model(xt.to(device=self.device, dtype=torch.float32)).float())
        model.train(was_training)
        if not outs:
            raise ValueError("`inputs_loader` is empty.")
        return torch.cat(outs, 0).contiguous()

    def log_prob(self, data_loader, distribute: bool = True) -> torch.Tensor:
        """log p(y | x) of every data point (predictive/base.py:27-57)."""
        return self._out.log_prob(self._outputs(data_loader.to_inputs_loader()), data_loader.to_array_targets(),
                                  calibrated=False)

    def sample(self, inputs_loader, n_samples: int = 1, rng=None, distribute: bool = True) -> torch.Tensor:
        """Target samples [n_samples, n(, d)] drawn with jax's own streams from ``rng`` (predictive/base.py:59-101)."""
        self._out.rng = self.rng
        return self._out.sample(n_samples, self._outputs(inputs_loader), rng=rng, calibrated=False)

    def mean(self, inputs_loader, distribute: bool = True) -> torch.Tensor:
        return self._out.mean(self._outputs(inputs_loader), calibrated=False)

    def mode(self, inputs_loader, distribute: bool = True) -> torch.Tensor:
        return self._out.mode(self._outputs(inputs_loader), calibrated=False)

    def variance(self, inputs_loader, distribute: bool = True) -> torch.Tensor:
        return self._out.variance(self._outputs(inputs_loader), calibrated=False)

    def std(self, inputs_loader, variances: Optional[torch.Tensor] = None, distribute: bool = True) -> torch.Tensor:
        if variances is not None:
            return torch.sqrt(torch.as_tensor(variances))
        return self._out.std(self._outputs(inputs_loader), calibrated=False)
