"""Configuration objects with the reference's names (fortuna/prob_model/fit_config/*.py): only the fields the
SG-MCMC fit reads.  NB the reference's SG-MCMC ``fit`` ignores ``optimizer.method`` (the sampler IS the optimizer,
sghmc_posterior.py:97-103) - same here."""
from typing import Callable, Optional, Tuple


class FitOptimizer:
    def __init__(self, method=None, n_epochs: int = 100, freeze_fun: Optional[Callable[[Tuple[str, ...], object], str]] = None):
        self.method, self.n_epochs, self.freeze_fun = method, n_epochs, freeze_fun


class FitMonitor:
    def __init__(self, metrics: Optional[Tuple[Callable, ...]] = None, **_ignored):
        self.metrics = metrics or ()


class FitCheckpointer:
    """fortuna/prob_model/fit_config/checkpointer.py:6-44, same fields and defaults."""

    def __init__(self, save_checkpoint_dir=None, restore_checkpoint_path=None, start_from_current_state: bool = False,
                 save_every_n_steps: Optional[int] = None, keep_top_n_checkpoints: Optional[int] = 2,
                 dump_state: bool = False):
        self.save_checkpoint_dir = save_checkpoint_dir
        self.save_every_n_steps = save_every_n_steps
        self.restore_checkpoint_path = restore_checkpoint_path
        self.start_from_current_state = start_from_current_state
        self.keep_top_n_checkpoints = keep_top_n_checkpoints
        self.dump_state = dump_state


class FitProcessor:
    """fortuna/prob_model/fit_config/processor.py:1-23.  ``devices=-1`` (the reference's default: use every device):
    when the process runs under ``torchrun`` with more than one rank, the fit is data-parallel over the ranks (one GPU
    each); ``devices=0``: this process alone.  ``disable_jit`` is accepted and unused (nothing is traced here)."""

    def __init__(self, devices: int = -1, disable_jit: bool = False):
        self.devices = devices
        self.disable_jit = disable_jit


class FitHyperparameters:
    """fortuna/prob_model/fit_config/hyperparameters.py:4-21.  As in the reference, the SG-MCMC fits do NOT read it
    (only the MAP / ADVI trainers pass ``max_grad_norm`` on: posterior/map/map_posterior.py:109; sghmc_posterior.py:97-111
    does not) - so setting it changes nothing for SGHMC / cyclical SGLD, there or here.  The clipping itself
    (fortuna/training/trainer.py:158-169, utils/training.py:14-20) is available on the sampler step:
    ``sghmc_integrator(...).apply(..., max_grad_norm=...)`` / ``run_sgmcmc(..., max_grad_norm=...)`` - one extra read of
    the gradient for the norm, the scaling rides inside the sampler kernel.  Gradient accumulation is not built."""

    def __init__(self, max_grad_norm: Optional[float] = None, gradient_accumulation_steps: Optional[int] = None):
        if gradient_accumulation_steps is not None and gradient_accumulation_steps > 1:
            raise NotImplementedError("gradient accumulation is outside the hot path")
        self.max_grad_norm = max_grad_norm
        self.gradient_accumulation_steps = gradient_accumulation_steps


class FitConfig:
    def __init__(self, optimizer: FitOptimizer = None, checkpointer: FitCheckpointer = None, monitor: FitMonitor = None,
                 processor: FitProcessor = None, hyperparameters: FitHyperparameters = None, **_ignored):
        self.optimizer = optimizer or FitOptimizer()
        self.checkpointer = checkpointer or FitCheckpointer()
        self.monitor = monitor or FitMonitor()
        self.processor = processor or FitProcessor()
        self.hyperparameters = hyperparameters or FitHyperparameters()
