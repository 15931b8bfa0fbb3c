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


class FitConfig:
    def __init__(self, optimizer: FitOptimizer = None, monitor: FitMonitor = None, **_ignored):
        self.optimizer = optimizer or FitOptimizer()
        self.monitor = monitor or FitMonitor()
