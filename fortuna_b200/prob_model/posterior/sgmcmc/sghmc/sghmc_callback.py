"""SGHMC sampling callback; constructor signature of
fortuna/prob_model/posterior/sgmcmc/sghmc/sghmc_callback.py:8-64 (same ValueError when too few samples)."""

from functools import partial

from ..sgmcmc_sampling_callback import SGMCMCSamplingCallback, sghmc_stores_step


class SGHMCSamplingCallback(SGMCMCSamplingCallback):
    def __init__(self, n_epochs, n_training_steps, n_samples, n_thinning, burnin_length, trainer, state_repository,
                 keep_top_n_checkpoints):
        rule = partial(sghmc_stores_step, burnin_length=burnin_length, n_thinning=n_thinning)
        super().__init__(trainer, state_repository, keep_top_n_checkpoints, rule=rule, n_samples=n_samples)
        self._require(n_epochs * n_training_steps, "the burnin length, number of epochs, or the thinning parameter")
