"""fortuna/prob_model/posterior/sgmcmc/sghmc/sghmc_approximator.py:16-63 - same defaults and checks."""

from typing import Union

from ..sgmcmc_preconditioner import Preconditioner, identity_preconditioner
from ..sgmcmc_step_schedule import StepSchedule, constant_schedule


class SGHMCPosteriorApproximator:
    def __init__(self, n_samples: int = 10, n_thinning: int = 1, burnin_length: int = 1000,
                 momentum_decay: float = 0.01, step_schedule: Union[StepSchedule, float] = 1e-5,
                 preconditioner: Preconditioner = identity_preconditioner()):
        self.n_samples = n_samples
        self.n_thinning = n_thinning
        self.burnin_length = burnin_length
        self.momentum_decay = momentum_decay
        if isinstance(step_schedule, float):
            step_schedule = constant_schedule(step_schedule)
        elif not callable(step_schedule):
            raise ValueError("`step_schedule` must be a a callable function.")
        self.step_schedule = step_schedule
        self.preconditioner = preconditioner

    def __str__(self):
        return "sghmc"
