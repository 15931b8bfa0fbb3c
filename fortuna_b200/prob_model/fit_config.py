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


class FitConfig:
    def __init__(self, optimizer: FitOptimizer = None, checkpointer: FitCheckpointer = None, monitor: FitMonitor = None,
                 processor: FitProcessor = None, **_ignored):
        self.optimizer = optimizer or FitOptimizer()
        self.checkpointer = checkpointer or FitCheckpointer()
        self.monitor = monitor or FitMonitor()
        self.processor = processor or FitProcessor()
