"""fortuna/prob_model/posterior/sgmcmc/cyclical_sgld/cyclical_sgld_approximator.py:10-55."""

from ..sgmcmc_preconditioner import Preconditioner, identity_preconditioner


class CyclicalSGLDPosteriorApproximator:
    def __init__(self, n_samples: int = 10, n_thinning: int = 1, cycle_length: int = 1000,
                 init_step_size: float = 1e-5, exploration_ratio: float = 0.25,
                 preconditioner: Preconditioner = identity_preconditioner()):
        self.n_samples = n_samples
        self.n_thinning = n_thinning
        self.cycle_length = cycle_length
        self.init_step_size = init_step_size
        self.exploration_ratio = exploration_ratio
        self.preconditioner = preconditioner

    def __str__(self):
        return "cyclical_sgld"
