"""Cyclical-SGLD sampling callback; constructor signature of
fortuna/prob_model/posterior/sgmcmc/cyclical_sgld/cyclical_sgld_callback.py:8-70."""

from functools import partial

from ..sgmcmc_sampling_callback import SGMCMCSamplingCallback, cyclical_sgld_stores_step


class CyclicalSGLDSamplingCallback(SGMCMCSamplingCallback):
    def __init__(self, n_epochs, n_training_steps, n_samples, n_thinning, cycle_length, exploration_ratio, trainer,
                 state_repository, keep_top_n_checkpoints):
        rule = partial(cyclical_sgld_stores_step, cycle_length=cycle_length, exploration_ratio=exploration_ratio,
                       n_thinning=n_thinning)
        super().__init__(trainer, state_repository, keep_top_n_checkpoints, rule=rule, n_samples=n_samples)
        self._require(n_epochs * n_training_steps,
                      "the cycle length, number of epochs, exploration ratio, or the thinning parameter")
