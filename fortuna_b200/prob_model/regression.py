"""``ProbRegressor`` - fortuna/prob_model/regression.py with an SG-MCMC posterior: ``model`` predicts the mean,
``likelihood_log_variance_model`` the log-variance (outputs concatenated to ``[mu, log sigma^2]`` as in
fortuna/model/model_manager/regression.py); a single module returning ``[B, 2d]`` is accepted too."""
import math
from typing import Optional

import torch

from ..utils.random import RandomNumberGenerator
from .calib_config import CalibConfig
from .calibration import RegressionTemperatureScaler, calibrate_temperature
from .fit_config import FitConfig
from .posterior.sgmcmc.cyclical_sgld.cyclical_sgld_approximator import CyclicalSGLDPosteriorApproximator
from .posterior.sgmcmc.cyclical_sgld.cyclical_sgld_posterior import CyclicalSGLDPosterior
from .posterior.sgmcmc.sghmc.sghmc_approximator import SGHMCPosteriorApproximator
from .posterior.sgmcmc.sghmc.sghmc_posterior import SGHMCPosterior
from .posterior.sgmcmc.torch_model import BoundModel, bind_for_fit
from .predictive.regression import RegressionPredictive
from .prior import IsotropicGaussianPrior


class _MeanAndLogVar(torch.nn.Module):
    def __init__(self, model, lik_log_var):
        super().__init__()
        self.model, self.lik_log_var = model, lik_log_var

    def forward(self, x):
        return torch.cat([self.model(x), self.lik_log_var(x)], dim=-1)


def gaussian_log_likelihood_sum(outputs: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
    """sum_b log N(y_b; mu_b, exp(logvar_b)) - fortuna/prob_output_layer/regression.py:30-34."""
    d = outputs.shape[-1] // 2
    mu, lv = outputs[..., :d], outputs[..., d:]
    t = targets.reshape(mu.shape).to(mu.dtype)
    return -0.5 * (torch.exp(-lv) * (t - mu) ** 2 + lv + math.log(2 * math.pi)).sum()


class ProbRegressor:
    def __init__(self, model: torch.nn.Module, likelihood_log_variance_model: Optional[torch.nn.Module] = None, prior=None,
                 posterior_approximator=None, output_calibrator="default", model_editor=None, seed: int = 0, device="cuda"):
        if model_editor is not None:
            raise NotImplementedError("model editors are outside the hot path")
        if output_calibrator == "default":  # the reference's default (regression.py:64)
            output_calibrator = RegressionTemperatureScaler()
        if output_calibrator is not None and not isinstance(output_calibrator, RegressionTemperatureScaler):
            raise NotImplementedError("only the default RegressionTemperatureScaler (or None) is supported")
        self.output_calibrator = output_calibrator
        self.model = model if likelihood_log_variance_model is None else _MeanAndLogVar(model, likelihood_log_variance_model)
        self.prior = prior if prior is not None else IsotropicGaussianPrior()
        self.posterior_approximator = posterior_approximator if posterior_approximator is not None else SGHMCPosteriorApproximator()
        if not isinstance(self.posterior_approximator, (SGHMCPosteriorApproximator, CyclicalSGLDPosteriorApproximator)):
            raise NotImplementedError("only the SG-MCMC posterior approximators (sghmc, cyclical_sgld) are on the hot path")
        self.device = torch.device(device)
        self.rng = RandomNumberGenerator(seed=seed, device=self.device)
        self.posterior = None
        self.predictive = None
        self._predictive_cls = RegressionPredictive

    def train(self, train_data_loader, val_data_loader=None, calib_data_loader=None, fit_config: Optional[FitConfig] = None,
              calib_config: Optional[CalibConfig] = None, **_ignored):
        fit_config = fit_config or FitConfig()
        bound = bind_for_fit(self.model, self.device, fit_config.optimizer.freeze_fun, fit_config.processor)
        cls = SGHMCPosterior if isinstance(self.posterior_approximator, SGHMCPosteriorApproximator) else CyclicalSGLDPosterior
        self.posterior = cls(bound, self.prior, self.posterior_approximator, self.rng)
        self.posterior.log_likelihood_sum = gaussian_log_likelihood_sum
        status = {"fit_status": self.posterior.fit(train_data_loader, val_data_loader, fit_config)}
        self.predictive = RegressionPredictive(self.posterior)
        if calib_data_loader is not None:
            status["calib_status"] = self.calibrate(calib_data_loader, val_data_loader, calib_config)
        return status

    def calibrate(self, calib_data_loader, val_data_loader=None, calib_config: Optional[CalibConfig] = None):
        """fortuna/prob_model/base.py:109-235 for the regression temperature scaler (log var + log_temp)."""
        if self.output_calibrator is None:
            return None
        if self.posterior is None:
            raise ValueError("Before calibration, you must either train the probabilistic model or load a state from an "
                             "existing checkpoint.")
        cfg = calib_config or CalibConfig()
        S = cfg.processor.n_posterior_samples
        pred = self.predictive
        pred.log_temp = None
        idx = self.posterior.sample_indices(self.rng.get(), S)

        def outputs(loader):
            return pred._stored_outputs(pred._x(loader.to_inputs_loader())).index_select(0, idx.long()).contiguous()

        z = outputs(calib_data_loader)
        zv = outputs(val_data_loader) if val_data_loader is not None else None
        tv = val_data_loader.to_array_targets() if val_data_loader is not None else None
        bs = getattr(calib_data_loader, "_batch_size", None) or z.shape[1]
        vbs = (getattr(val_data_loader, "_batch_size", None) or zv.shape[1]) if zv is not None else None
        status, log_temp = calibrate_temperature(z, calib_data_loader.to_array_targets(), bs, cfg, zv, tv, vbs, regression=True)
        pred.log_temp = torch.from_numpy(log_temp).to(self.device)
        self.posterior.log_temp = pred.log_temp
        if self.posterior.state.checkpoint_dir:
            self.posterior.save_state(self.posterior.state.checkpoint_dir)
        return status

    def attach_posterior_samples(self, samples: torch.Tensor, freeze_fun=None, steps=None) -> None:
        """Make ``samples`` - float32 [n_samples, N_trainable], rows in the flat parameter layout of ``utils.flat_tree`` over
        the leaves ``freeze_fun`` keeps trainable - the posterior of this model without running a fit here (chains that ran
        elsewhere).  The in-memory sibling of ``load_state``; same contract as ``ProbClassifier.attach_posterior_samples``."""
        from .posterior.sgmcmc.sgmcmc_posterior_state_repository import SGMCMCPosteriorStateRepository

        self.model.to(self.device)
        bound = BoundModel(self.model, self.device, freeze_fun=freeze_fun)
        n, width = samples.shape
        if width != bound.params.flat.numel():
            raise ValueError(f"samples have {width} parameters per row, the trainable tree has {bound.params.flat.numel()}")
        cls = SGHMCPosterior if isinstance(self.posterior_approximator, SGHMCPosteriorApproximator) else CyclicalSGLDPosterior
        self.posterior = cls(bound, self.prior, self.posterior_approximator, self.rng)
        self.posterior.log_likelihood_sum = gaussian_log_likelihood_sum
        repo = SGMCMCPosteriorStateRepository(n, template=bound.params)
        repo.bank.copy_(samples.to(device=self.device, dtype=torch.float32))
        repo._filled = [True] * n
        repo.steps = list(steps) if steps is not None else list(range(n))
        self.posterior.state = repo
        self.predictive = self._predictive_cls(self.posterior)

    # ------------------------------------------------------------------ state on disk (fortuna/prob_model/base.py:236-261)
    def load_state(self, checkpoint_path) -> None:
        """Load posterior samples from a checkpoint directory in the reference's layout (``<dir>/c`` chain state,
        ``<dir>/<i>`` samples) into the device sample bank; which leaves are sampled comes from the chain file."""
        import os

        from ..training import checkpoint as ckpt

        d = ckpt.restore_checkpoint(os.path.join(str(checkpoint_path), "c")) if os.path.isdir(os.path.join(str(checkpoint_path), "c")) else None
        if d is None:
            raise ValueError(f"No checkpoint was found in `checkpoint_dir={checkpoint_path}`.")
        enc = d.get("_encoded_which_params")
        freeze_fun = None
        if enc is not None:
            which = {tuple("".join(chr(int(o)) for o in v) for v in path.values()) for path in enc.values()}
            freeze_fun = lambda path, v: "trainable" if tuple(path) in which else "frozen"  # noqa: E731
        self.model.to(self.device)
        bound = BoundModel(self.model, self.device, freeze_fun=freeze_fun)
        cls = SGHMCPosterior if isinstance(self.posterior_approximator, SGHMCPosteriorApproximator) else CyclicalSGLDPosterior
        self.posterior = cls(bound, self.prior, self.posterior_approximator, self.rng)
        self.posterior.load_state(checkpoint_path)
        self.predictive = self._predictive_cls(self.posterior, log_temp=self.posterior.log_temp)

    def save_state(self, checkpoint_path, keep_top_n_checkpoints: int = 1) -> None:
        if self.posterior is None:
            raise ValueError("No state available.")
        return self.posterior.save_state(checkpoint_path, keep_top_n_checkpoints)
