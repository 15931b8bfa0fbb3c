"""The user API on BASELINE config 1 (two-moons MLP ProbClassifier, SGHMC / cyclical-SGLD posterior, predictive
mean / entropy), exercised like tests/fortuna/prob_model/test_train.py:49-60,206-389 does (train -> sample ->
predictive), plus value-level checks the reference's smoke tests do not make."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples"))


def _setup():
    from two_moons_sghmc import MLP, make_moons

    from fortuna_b200.data import DataLoader

    train, test = make_moons(500, 0.07, 0), make_moons(400, 0.07, 2)
    return MLP, train, test, DataLoader.from_array_data(train, batch_size=128, shuffle=True), DataLoader.from_array_data(test, batch_size=128)


def test_sghmc_two_moons_train_and_predict(cuda):
    from fortuna_b200.data import InputsLoader
    from fortuna_b200.metric.classification import accuracy
    from fortuna_b200.prob_model import FitConfig, FitOptimizer, ProbClassifier, SGHMCPosteriorApproximator
    from oracle import predictive as P

    MLP, train, test, train_loader, test_loader = _setup()
    torch.manual_seed(0)
    pm = ProbClassifier(model=MLP(), posterior_approximator=SGHMCPosteriorApproximator(
        n_samples=20, n_thinning=10, burnin_length=600, step_schedule=1e-3), seed=0)
    status = pm.train(train_loader, val_data_loader=test_loader, fit_config=FitConfig(optimizer=FitOptimizer(n_epochs=250)))
    assert status["fit_status"]["loss"].shape == (250,) and status["fit_status"]["val_loss"].shape == (250,)
    assert pm.posterior.state.n_filled == 20
    names = ("mean", "mode", "entropy", "aleatoric_entropy", "epistemic_entropy", "variance", "aleatoric_variance",
             "epistemic_variance", "log_prob", "ensemble_log_prob")
    key = pm.rng.get()
    st = pm.predictive.statistics(test_loader, names, n_posterior_samples=30, rng=key)
    assert st["mean"].shape == (400, 2) and st["ensemble_log_prob"].shape == (30, 400)
    assert float(accuracy(st["mode"], test[1])) > 0.97
    # the same key gives the same draws: each single-statistic method equals the joint pass
    inputs = test_loader.to_inputs_loader()
    assert torch.equal(pm.predictive.mean(inputs, 30, rng=key), st["mean"])
    assert torch.equal(pm.predictive.entropy(inputs, 30, rng=key), st["entropy"])
    assert torch.equal(pm.predictive.log_prob(test_loader, 30, rng=key), st["log_prob"])
    # everything against the oracle on the very logits the predictive used
    z = pm.predictive.sample_calibrated_outputs(inputs, 30, rng=key).cpu().numpy()
    np.testing.assert_allclose(st["mean"].cpu().numpy(), P.cls_mean(z), rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(st["entropy"].cpu().numpy(), P.cls_entropy(z), rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(st["epistemic_entropy"].cpu().numpy(), P.cls_epistemic_entropy(z), rtol=1e-5, atol=4e-6)
    np.testing.assert_allclose(st["log_prob"].cpu().numpy(), P.cls_log_prob(z, test[1]), rtol=1e-5, atol=2e-6)
    # the S draws are the reference's: idx_s = choice(split(key, S)[s], n_samples)
    from oracle import prng as oprng
    from fortuna_b200 import ops

    keys = oprng.split(ops.key_to_host(key), 30)
    want_idx = np.array([oprng.choice(k, 20) for k in keys])
    assert np.array_equal(pm.posterior.sample_indices(key, 30).cpu().numpy(), want_idx)
    # epistemic uncertainty is larger far from the data than on it
    grid = np.stack(np.meshgrid(np.linspace(-4, 5, 20), np.linspace(-4, 4, 20)), -1).reshape(-1, 2).astype(np.float32)
    far = grid[(np.abs(grid[:, 0] - 0.5) > 3) | (np.abs(grid[:, 1]) > 3)]
    ent_far = float(pm.predictive.epistemic_entropy(InputsLoader.from_array_inputs(far), 30).mean())
    assert ent_far > float(st["epistemic_entropy"].mean())
    sets = pm.predictive.credible_set(inputs, 30, error=0.05)
    assert len(sets) == 400 and all(1 <= len(s) <= 2 for s in sets)


def test_last_layer_posterior_uses_batched_contraction(cuda):
    """freeze_fun keeps only output_subnet trainable (examples/scaling_up_bayesian_inference.pct.py:90-104): the
    predictive then runs features once + K2 over the sample bank, and must equal one forward pass per drawn sample."""
    from fortuna_b200.prob_model import CyclicalSGLDPosteriorApproximator, FitConfig, FitOptimizer, ProbClassifier

    MLP, train, test, train_loader, test_loader = _setup()
    torch.manual_seed(1)
    pm = ProbClassifier(model=MLP(), posterior_approximator=CyclicalSGLDPosteriorApproximator(
        n_samples=8, n_thinning=5, cycle_length=40, init_step_size=1e-3, exploration_ratio=0.25), seed=3)
    freeze = lambda path, v: "trainable" if "output_subnet" in path else "frozen"
    pm.train(train_loader, fit_config=FitConfig(optimizer=FitOptimizer(n_epochs=40, freeze_fun=freeze)))
    bound = pm.posterior.bound
    assert bound.is_last_layer_only() and bound.params.n_params == 30 * 2 + 2
    assert pm.posterior.state.n_filled == 8
    key = pm.rng.get()
    inputs = test_loader.to_inputs_loader()
    z = pm.predictive.sample_calibrated_outputs(inputs, 12, rng=key)
    idx = pm.posterior.sample_indices(key, 12).cpu().numpy()
    x = torch.as_tensor(test[0], device=cuda)
    keep = bound.params.flat.clone()
    for s, i in enumerate(idx):
        bound.load_flat(pm.posterior.state.bank[int(i)])
        with torch.no_grad():
            ref = bound.model(x)
        np.testing.assert_allclose(z[s].cpu().numpy(), ref.cpu().numpy(), rtol=1e-4, atol=1e-5)
    bound.load_flat(keep)


def test_rejections(cuda):
    from fortuna_b200.prob_model import ProbClassifier, SGHMCPosteriorApproximator

    MLP = _setup()[0]
    with pytest.raises(NotImplementedError):
        ProbClassifier(model=MLP(), posterior_approximator="advi")
    with pytest.raises(ValueError):  # too few steps to ever collect n_samples (sghmc_callback.py:55-64)
        from fortuna_b200.data import DataLoader
        from fortuna_b200.prob_model import FitConfig, FitOptimizer

        _, train, _, train_loader, _ = _setup()
        ProbClassifier(model=MLP(), posterior_approximator=SGHMCPosteriorApproximator(n_samples=10, burnin_length=1000)).train(
            train_loader, fit_config=FitConfig(optimizer=FitOptimizer(n_epochs=2)))
