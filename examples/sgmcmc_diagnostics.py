#!/usr/bin/env python
"""Quality of SG-MCMC samples (the reference's examples/sgmcmc_diagnostics.pct.py): kernel Stein discrepancy with the IMQ
kernel and effective sample size, computed over the on-device sample bank; plus calibration, conformal Bayes sets and a
checkpoint round trip in the reference's on-disk layout.  Needs a B200 (there is no CPU path)."""
import os
import sys
import tempfile

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from two_moons_sghmc import MLP, make_moons  # noqa: E402

from fortuna_b200.data import DataLoader  # noqa: E402
from fortuna_b200.prob_model import (  # noqa: E402
    CalibConfig, CalibOptimizer, FitCheckpointer, FitConfig, FitOptimizer, ProbClassifier, SGHMCPosteriorApproximator,
)
from fortuna_b200.prob_model.posterior.sgmcmc.sgmcmc_diagnostic import (  # noqa: E402
    effective_sample_size, kernel_stein_discrepancy_imq,
)


def main():
    torch.manual_seed(0)
    train, test = make_moons(500, 0.07, 0), make_moons(400, 0.07, 2)
    train_loader = DataLoader.from_array_data(train, batch_size=128, shuffle=True)
    test_loader = DataLoader.from_array_data(test, batch_size=128)
    with tempfile.TemporaryDirectory() as ckpt_dir:
        pm = ProbClassifier(model=MLP(), posterior_approximator=SGHMCPosteriorApproximator(
            n_samples=30, n_thinning=10, burnin_length=600, step_schedule=1e-3))
        pm.train(train_loader, calib_data_loader=test_loader,
                 fit_config=FitConfig(optimizer=FitOptimizer(n_epochs=250), checkpointer=FitCheckpointer(save_checkpoint_dir=ckpt_dir)),
                 calib_config=CalibConfig(optimizer=CalibOptimizer(n_epochs=30)))
        print("fitted log-temperature:", float(pm.predictive.log_temp[0]))
        bank = pm.posterior.state.bank  # [n_samples, n_params] on the device: the chain, in storage order
        ess = effective_sample_size(bank[:, : pm.posterior.bound.params.flat.numel()])
        print("ESS over parameters: median %.1f, min %.1f of %d samples" % (float(ess.nanmedian()), float(ess[~ess.isnan()].min()), bank.shape[0]))
        # gradients of the log-joint at every stored sample (one backward pass each) for the Stein discrepancy
        bound, x, y = pm.posterior.bound, torch.as_tensor(train[0], device="cuda"), torch.as_tensor(train[1], device="cuda")
        keep, grads = bound.params.flat.clone(), []
        for i in range(bank.shape[0]):
            bound.load_flat(bank[i])
            bound.zero_grad()
            lj = -torch.nn.functional.cross_entropy(bound.model(x), y, reduction="sum") - 0.5 * bound.params.flat.square().sum()
            lj.backward()
            grads.append(bound.grads.flat.clone())
        bound.load_flat(keep)
        print("KSD (IMQ):", float(kernel_stein_discrepancy_imq(bank, torch.stack(grads))))
        sets = pm.predictive.conformal_set(train_loader, test_loader.to_inputs_loader(), n_posterior_samples=30, error=0.05)
        print("conformal Bayes: mean set size %.3f, coverage %.3f" % (
            np.mean([len(s) for s in sets]), np.mean([int(t) in s for s, t in zip(sets, test[1])])))
        pm2 = ProbClassifier(model=MLP(), posterior_approximator=SGHMCPosteriorApproximator(n_samples=30))
        pm2.load_state(ckpt_dir)  # flax-msgpack files <dir>/c and <dir>/<i>, as Fortuna writes them
        key = pm.rng.get()
        inputs = test_loader.to_inputs_loader()
        same = torch.equal(pm2.predictive.mean(inputs, 30, rng=key), pm.predictive.mean(inputs, 30, rng=key))
        print("reloaded from checkpoints:", pm2.posterior.state.n_filled, "samples;", "predictive reproducible" if same else "mismatch")


if __name__ == "__main__":
    main()
