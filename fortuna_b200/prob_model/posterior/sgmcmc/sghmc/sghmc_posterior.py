"""SGHMC posterior - ``fit`` / ``sample`` of fortuna/prob_model/posterior/sgmcmc/sghmc/sghmc_posterior.py:41-225."""
from ..sgmcmc_fit import run_sgmcmc
from ..sgmcmc_posterior import SGMCMCPosterior
from ..sgmcmc_posterior_state_repository import SGMCMCPosteriorStateRepository
from .sghmc_callback import SGHMCSamplingCallback
from .sghmc_integrator import sghmc_integrator


class SGHMCPosterior(SGMCMCPosterior):
    def __init__(self, bound_model, prior, posterior_approximator, rng):
        self.bound, self.prior, self.posterior_approximator = bound_model, prior, posterior_approximator
        super().__init__(SGMCMCPosteriorStateRepository(posterior_approximator.n_samples, template=bound_model.params), rng)

    def fit(self, train_data_loader, val_data_loader=None, fit_config=None, **_):
        a = self.posterior_approximator
        n_epochs = fit_config.optimizer.n_epochs
        c = fit_config.checkpointer
        self._restore_chain_params(fit_config)
        self._fit_repository(fit_config)
        tx = sghmc_integrator(a.momentum_decay, None, self.rng.get(), a.step_schedule, a.preconditioner)
        cb = SGHMCSamplingCallback(n_epochs=n_epochs, n_training_steps=len(train_data_loader), n_samples=a.n_samples,
                                   n_thinning=a.n_thinning, burnin_length=a.burnin_length, trainer=None,
                                   state_repository=self.state, keep_top_n_checkpoints=c.keep_top_n_checkpoints)
        kw = {} if getattr(self, "log_likelihood_sum", None) is None else {"log_likelihood_sum": self.log_likelihood_sum}
        status, self.chain_state = run_sgmcmc(self.bound, self.prior, tx, [cb], train_data_loader, n_epochs, val_data_loader,
                                              save_chain=self._chain_saver(fit_config), save_every_n_steps=c.save_every_n_steps, **kw)
        return status
