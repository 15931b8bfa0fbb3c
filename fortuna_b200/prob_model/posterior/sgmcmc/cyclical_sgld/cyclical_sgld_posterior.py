"""Cyclical SGLD posterior - fortuna/prob_model/posterior/sgmcmc/cyclical_sgld/cyclical_sgld_posterior.py."""
from ..sgmcmc_fit import run_sgmcmc
from ..sgmcmc_posterior import SGMCMCPosterior
from ..sgmcmc_posterior_state_repository import SGMCMCPosteriorStateRepository
from .cyclical_sgld_callback import CyclicalSGLDSamplingCallback
from .cyclical_sgld_integrator import cyclical_sgld_integrator


class CyclicalSGLDPosterior(SGMCMCPosterior):
    STATE_NAME = "CyclicalSGLDState"

    def __init__(self, bound_model, prior, posterior_approximator, rng):
        self.bound, self.prior, self.posterior_approximator = bound_model, prior, posterior_approximator
        super().__init__(SGMCMCPosteriorStateRepository(posterior_approximator.n_samples, template=bound_model.params), rng)

    def fit(self, train_data_loader, val_data_loader=None, fit_config=None, **_):
        a = self.posterior_approximator
        n_epochs = fit_config.optimizer.n_epochs
        c = fit_config.checkpointer
        self._restore_chain_params(fit_config)
        self._fit_repository(fit_config)
        tx = cyclical_sgld_integrator(self.rng.get(), a.init_step_size, a.cycle_length, a.exploration_ratio, a.preconditioner)
        cb = CyclicalSGLDSamplingCallback(n_epochs=n_epochs, n_training_steps=len(train_data_loader),
                                          n_samples=a.n_samples, n_thinning=a.n_thinning, cycle_length=a.cycle_length,
                                          exploration_ratio=a.exploration_ratio, trainer=None,
                                          state_repository=self.state, keep_top_n_checkpoints=c.keep_top_n_checkpoints)
        kw = {} if getattr(self, "log_likelihood_sum", None) is None else {"log_likelihood_sum": self.log_likelihood_sum}
        status, self.chain_state = run_sgmcmc(self.bound, self.prior, tx, [cb], train_data_loader, n_epochs, val_data_loader,
                                              save_chain=self._chain_saver(fit_config), save_every_n_steps=c.save_every_n_steps, **kw)
        return status
