"""``ProbClassifier`` - the user-facing object of fortuna/prob_model/classification.py:48-261 and
fortuna/prob_model/base.py:50-107 for the SG-MCMC posteriors on the hot path:

    prob_model = ProbClassifier(model=..., posterior_approximator=SGHMCPosteriorApproximator(...))
    status = prob_model.train(train_data_loader, fit_config=FitConfig(optimizer=FitOptimizer(n_epochs=30)))
    means = prob_model.predictive.mean(inputs_loader=test_inputs_loader)

``model`` is a torch ``nn.Module`` returning logits (model definitions are outside the hot path; modules that follow
the reference's ``dfe_subnet`` / ``output_subnet`` split get the S-batched last-layer contraction when only the
output layer is sampled, via ``FitOptimizer(freeze_fun=...)``).  Other posterior approximators of the reference are
out of scope and rejected."""
from typing import Optional

import torch

from ..utils.random import RandomNumberGenerator
from .calib_config import CalibConfig
from .calibration import ClassificationTemperatureScaler, calibrate_temperature
from .fit_config import FitConfig
from .posterior.sgmcmc.cyclical_sgld.cyclical_sgld_approximator import CyclicalSGLDPosteriorApproximator
from .posterior.sgmcmc.cyclical_sgld.cyclical_sgld_posterior import CyclicalSGLDPosterior
from .posterior.sgmcmc.sghmc.sghmc_approximator import SGHMCPosteriorApproximator
from .posterior.sgmcmc.sghmc.sghmc_posterior import SGHMCPosterior
from .posterior.sgmcmc.torch_model import BoundModel, bind_for_fit
from .predictive.classification import ClassificationPredictive
from .prior import IsotropicGaussianPrior


class ProbClassifier:
    def __init__(self, model: torch.nn.Module, prior=None, posterior_approximator=None,
                 output_calibrator="default", model_editor=None, seed: int = 0, device="cuda"):
        if model_editor is not None:
            raise NotImplementedError("model editors are outside the hot path")
        if output_calibrator == "default":  # the reference's default (classification.py:60)
            output_calibrator = ClassificationTemperatureScaler()
        if output_calibrator is not None and not isinstance(output_calibrator, ClassificationTemperatureScaler):
            raise NotImplementedError("only the default ClassificationTemperatureScaler (or None) is supported")
        self.output_calibrator = output_calibrator
        self.model = model
        self.prior = prior if prior is not None else IsotropicGaussianPrior()
        self.posterior_approximator = posterior_approximator if posterior_approximator is not None else SGHMCPosteriorApproximator()
        if not isinstance(self.posterior_approximator, (SGHMCPosteriorApproximator, CyclicalSGLDPosteriorApproximator)):
            raise NotImplementedError(f"posterior approximator {self.posterior_approximator!s} is outside the hot path "
                                      "(SG-MCMC only: sghmc, cyclical_sgld)")
        self.device = torch.device(device)
        self.rng = RandomNumberGenerator(seed=seed, device=self.device)  # one key chain shared by all parts (base.py:36-48)
        self.posterior = None
        self.predictive = None
        self._predictive_cls = ClassificationPredictive

    def _check_output_dim(self, data_loader):
        """fortuna/prob_model/classification.py:218-242: targets must be integer classes below the output width."""
        for x, y in data_loader:
            out_dim = self.model(torch.as_tensor(x[:1], dtype=torch.float32, device=next(self.model.parameters()).device)).shape[-1]
            if int(y.max()) >= out_dim or int(y.min()) < 0:
                raise ValueError(f"The outputs dimension of `model` must be at least {int(y.max()) + 1}, but {out_dim} was found.")
            break

    def train(self, train_data_loader, val_data_loader=None, calib_data_loader=None, fit_config: Optional[FitConfig] = None,
              calib_config=None, map_fit_config=None, **kwargs):
        if map_fit_config is not None:
            raise NotImplementedError("a preliminary MAP fit is outside the hot path")
        fit_config = fit_config or FitConfig()
        self.model.to(self.device)
        self._check_output_dim(train_data_loader)
        bound = bind_for_fit(self.model, self.device, fit_config.optimizer.freeze_fun, fit_config.processor)
        cls = SGHMCPosterior if isinstance(self.posterior_approximator, SGHMCPosteriorApproximator) else CyclicalSGLDPosterior
        self.posterior = cls(bound, self.prior, self.posterior_approximator, self.rng)
        status = {"fit_status": self.posterior.fit(train_data_loader, val_data_loader, fit_config)}
        self.predictive = ClassificationPredictive(self.posterior)
        if calib_data_loader is not None:  # fortuna/prob_model/base.py:94-105
            status["calib_status"] = self.calibrate(calib_data_loader, val_data_loader, calib_config)
        return status

    def calibrate(self, calib_data_loader, val_data_loader=None, calib_config: Optional[CalibConfig] = None):
        """fortuna/prob_model/base.py:109-235 for the temperature scaler: the ensemble of outputs on the calibration data
        is computed once, log_temp is fitted on it (fb200_calib_cls_loss_terms per step), and every later predictive call
        uses the fitted temperature."""
        if self.output_calibrator is None:
            return None  # "Nothing to calibrate. No calibrator was passed to the probabilistic model."
        if self.posterior is None:
            raise ValueError("Before calibration, you must either train the probabilistic model or load a state from an "
                             "existing checkpoint.")
        cfg = calib_config or CalibConfig()
        S = cfg.processor.n_posterior_samples
        self.predictive.log_temp = None  # the ensemble is UNcalibrated outputs
        key = self.rng.get()
        z = self.predictive.sample_calibrated_outputs(calib_data_loader.to_inputs_loader(), S, rng=key).float().contiguous()
        zv = tv = None
        if val_data_loader is not None:
            zv = self.predictive.sample_calibrated_outputs(val_data_loader.to_inputs_loader(), S, rng=key).float().contiguous()
            tv = val_data_loader.to_array_targets()
        bs = getattr(calib_data_loader, "_batch_size", None) or z.shape[1]
        vbs = (getattr(val_data_loader, "_batch_size", None) or zv.shape[1]) if zv is not None else None
        status, log_temp = calibrate_temperature(z, calib_data_loader.to_array_targets(), bs, cfg, zv, tv, vbs)
        self.predictive.log_temp = torch.from_numpy(log_temp).to(self.device)
        self.posterior.log_temp = self.predictive.log_temp
        if self.posterior.state.checkpoint_dir:  # posterior.state.update(calib_params=...): base.py:213-215
            self.posterior.save_state(self.posterior.state.checkpoint_dir)
        return status

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

    def attach_posterior_samples(self, samples: torch.Tensor, freeze_fun=None, steps=None, last_layer_dtype=torch.float32,
                                 logits_dtype=torch.float32) -> None:
        """Make ``samples`` - float32 [n_samples, N_trainable], rows in the flat parameter layout of
        ``utils.flat_tree`` over the leaves ``freeze_fun`` keeps trainable - the posterior of this model without running a
        fit here: chains that ran elsewhere (one per GPU, another job), or the benchmark's random-init weights.  The
        in-memory sibling of ``load_state`` (which reads the same rows from ``<dir>/<i>`` checkpoints).
        ``last_layer_dtype`` / ``logits_dtype``: see ``ClassificationPredictive``."""
        from .posterior.sgmcmc.sgmcmc_posterior_state_repository import SGMCMCPosteriorStateRepository

        self.model.to(self.device)
        bound = BoundModel(self.model, self.device, freeze_fun=freeze_fun)
        n, width = samples.shape
        if width != bound.params.flat.numel():
            raise ValueError(f"samples have {width} parameters per row, the trainable tree has {bound.params.flat.numel()}")
        cls = SGHMCPosterior if isinstance(self.posterior_approximator, SGHMCPosteriorApproximator) else CyclicalSGLDPosterior
        self.posterior = cls(bound, self.prior, self.posterior_approximator, self.rng)
        repo = SGMCMCPosteriorStateRepository(n, template=bound.params)
        repo.bank.copy_(samples.to(device=self.device, dtype=torch.float32))
        repo._filled = [True] * n
        repo.steps = list(steps) if steps is not None else list(range(n))
        self.posterior.state = repo
        self.predictive = self._predictive_cls(self.posterior, last_layer_dtype=last_layer_dtype, logits_dtype=logits_dtype)

    def save_state(self, checkpoint_path, keep_top_n_checkpoints: int = 1) -> None:
        if self.posterior is None:
            raise ValueError("No state available.")
        return self.posterior.save_state(checkpoint_path, keep_top_n_checkpoints)
