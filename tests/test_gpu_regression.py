"""Regression side of the path: Monte-Carlo target samples with jax's exact draws, axis-0 quantiles, and the
ProbRegressor / RegressionPredictive user API on a 1-D sine regression (fortuna/prob_model/predictive/regression.py,
fortuna/prob_output_layer/regression.py, tests/fortuna/test_predictive.py:61-230 check shapes and variance >= 0)."""
import numpy as np
import pytest
import torch

from oracle import predictive as P
from oracle import prng as oprng

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,B,d,T", [(5, 7, 2, 6), (10, 129, 1, 30), (3, 33, 3, 1)])
def test_reg_sample_targets_matches_oracle(cuda, n, B, d, T):
    from fortuna_b200 import ops

    r = np.random.default_rng(n * 10 + d)
    z = r.standard_normal((n, B, 2 * d)).astype(np.float32)
    lt = np.array([0.25], np.float32)
    want, want_idx = P.reg_sample_targets(z, oprng.PRNGKey(9), T, lt)
    y, idx = ops.reg_sample_targets(torch.from_numpy(z).to(cuda), ops.PRNGKey(9, cuda), T,
                                    log_temp=torch.from_numpy(lt).to(cuda), want_idx=True)
    assert np.array_equal(idx.cpu().numpy(), want_idx)  # posterior-sample draws: bit-exact
    np.testing.assert_allclose(y.cpu().numpy(), want, rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("T,shape", [(30, (50, 1)), (7, (13, 3)), (1, (4, 2)), (256, (9, 1))])
def test_quantile_axis0_matches_oracle(cuda, T, shape):
    from fortuna_b200 import ops

    x = np.random.default_rng(T).standard_normal((T,) + shape).astype(np.float32)
    q = np.array([0.0, 0.025, 0.5, 0.975, 1.0], np.float32)
    got = ops.quantile_axis0(torch.from_numpy(x).to(cuda), torch.from_numpy(q).to(cuda)).cpu().numpy()
    np.testing.assert_allclose(got, P.quantile_axis0(x, q), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(got, np.quantile(x.astype(np.float64), q.astype(np.float64), axis=0), rtol=1e-5, atol=1e-6)


def test_prob_regressor_sine(cuda):
    from fortuna_b200.data import DataLoader
    from fortuna_b200.prob_model import FitConfig, FitOptimizer, ProbRegressor, SGHMCPosteriorApproximator

    r = np.random.default_rng(0)
    x = r.uniform(-3, 3, (400, 1)).astype(np.float32)
    y = (np.sin(x) + 0.2 * r.standard_normal(x.shape)).astype(np.float32)
    xt = np.linspace(-3, 3, 64, dtype=np.float32)[:, None]
    yt = np.sin(xt)
    mlp = lambda out: torch.nn.Sequential(torch.nn.Linear(1, 50), torch.nn.Tanh(), torch.nn.Linear(50, out))
    torch.manual_seed(0)
    pr = ProbRegressor(model=mlp(1), likelihood_log_variance_model=mlp(1),
                       posterior_approximator=SGHMCPosteriorApproximator(n_samples=10, n_thinning=10, burnin_length=1500,
                                                                         step_schedule=1e-5), seed=1)
    pr.train(DataLoader.from_array_data((x, y), batch_size=100, shuffle=True),
             fit_config=FitConfig(optimizer=FitOptimizer(n_epochs=650)))
    test = DataLoader.from_array_data((xt, yt), batch_size=32)
    inputs = test.to_inputs_loader()
    key = pr.rng.get()
    mean = pr.predictive.mean(inputs, 20, rng=key)
    var = pr.predictive.variance(inputs, 20, rng=key)
    assert mean.shape == (64, 1) and var.shape == (64, 1) and bool((var >= 0).all())
    assert float((mean.cpu() - torch.from_numpy(yt)).abs().mean()) < 0.15
    lp = pr.predictive.log_prob(test, 20, rng=key)
    assert lp.shape == (64,) and bool(torch.isfinite(lp).all())
    # batch by batch == all at once (to float32 rounding: the torch backbone picks its GEMM by batch size)
    whole = DataLoader.from_array_data((xt, yt), batch_size=None)
    np.testing.assert_allclose(lp.cpu().numpy(), pr.predictive.log_prob(whole, 20, rng=key).cpu().numpy(), rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(mean.cpu().numpy(), pr.predictive.mean(whole.to_inputs_loader(), 20, rng=key).cpu().numpy(),
                               rtol=2e-5, atol=2e-6)
    elp = pr.predictive.ensemble_log_prob(test, 20, rng=key)
    assert elp.shape == (20, 64)
    # sample / quantile / credible interval: reference semantics (per-batch noise with the same key) vs the oracle
    smp = pr.predictive.sample(inputs, n_target_samples=30, rng=key)
    assert smp.shape == (30, 64, 1)
    stored = pr.predictive._stored_outputs(torch.from_numpy(xt).to(cuda)).cpu().numpy()
    from fortuna_b200 import ops

    want = np.concatenate([P.reg_sample_targets(stored[:, i : i + 32], ops.key_to_host(key), 30)[0] for i in (0, 32)], axis=1)
    np.testing.assert_allclose(smp.cpu().numpy(), want, rtol=1e-5, atol=2e-6)
    ci = pr.predictive.credible_interval(inputs, n_target_samples=30, error=0.05, rng=key)
    assert ci.shape == (64, 2) and bool((ci[:, 0] < ci[:, 1]).all())
    np.testing.assert_allclose(ci.cpu().numpy()[:, 0], P.quantile_axis0(want, [0.025])[0, :, 0], rtol=1e-5, atol=1e-6)
    q = pr.predictive.quantile(0.5, inputs, n_target_samples=30, rng=key)
    assert q.shape == (64, 1)
    with pytest.raises(ValueError):
        pr.predictive.credible_interval(inputs, interval_type="both")
    # Monte-Carlo entropies (regression.py:51-267) through the user API, against the oracle on the stored outputs
    # like the reference's batched loop (predictive/base.py:946-1013) every BATCH of the loader is evaluated with the same
    # keys - the target-sample noise has the batch's shape - so the oracle runs per batch of 32 too
    parts = [P.reg_mc_entropies(stored[:, i : i + 32], ops.key_to_host(key), 12, 9) for i in (0, 32)]
    want_e = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
    for name in ("entropy", "aleatoric_entropy", "epistemic_entropy"):
        got = getattr(pr.predictive, name)(inputs, n_posterior_samples=12, n_target_samples=9, rng=key)
        assert got.shape == (64,)
        np.testing.assert_allclose(got.cpu().numpy(), want_e[name], rtol=2e-5, atol=2e-5, err_msg=name)



@pytest.mark.parametrize("n,B,d,S,T", [(5, 7, 2, 6, 4), (10, 129, 1, 30, 30), (3, 33, 3, 2, 1), (4, 40, 8, 5, 3)])
def test_reg_mc_entropies_match_oracle(cuda, n, B, d, S, T):
    """fb200_reg_mc_entropies: bit-exact target draws (threefry element (t, b, k) of keys[i]), Monte-Carlo means of
    log N within rtol 1e-5; entropy = aleatoric + epistemic by construction."""
    from fortuna_b200 import ops

    r = np.random.default_rng(n + B + d)
    z = np.concatenate([r.standard_normal((n, B, d)), 0.5 * r.standard_normal((n, B, d)) - 1.0], axis=-1).astype(np.float32)
    lt = np.array([0.1], np.float32)
    want = P.reg_mc_entropies(z, oprng.PRNGKey(21), S, T, lt)
    key = ops.PRNGKey(21, cuda)
    ks = ops.split(key, 1 + S)
    idx = ops.sample_indices(ks[0].contiguous(), S, n)
    got = ops.reg_mc_entropies(torch.from_numpy(z).to(cuda), idx, ks[1:].contiguous(), T,
                               ("entropy", "aleatoric_entropy", "epistemic_entropy"), log_temp=torch.from_numpy(lt).to(cuda))
    for name in want:
        np.testing.assert_allclose(got[name].cpu().numpy(), want[name], rtol=2e-5, atol=2e-5, err_msg=name)
    np.testing.assert_allclose((got["aleatoric_entropy"] + got["epistemic_entropy"]).cpu().numpy(),
                               got["entropy"].cpu().numpy(), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("S,B,d", [(5, 40, 1), (12, 1000, 3), (2, 257, 8)])
@pytest.mark.parametrize("lt", [0.0, -0.6])
def test_calib_reg_loss_terms_match_oracle(cuda, S, B, d, lt):
    from fortuna_b200 import ops
    from oracle import calibration as Cal

    r = np.random.default_rng(S + B)
    z = np.concatenate([r.standard_normal((S, B, d)), 0.5 * r.standard_normal((S, B, d)) - 0.5], axis=-1).astype(np.float32)
    y = r.standard_normal((B, d)).astype(np.float32)
    got = ops.calib_reg_loss_terms(torch.from_numpy(z).to(cuda), torch.from_numpy(y).to(cuda),
                                   torch.tensor([lt], dtype=torch.float32, device=cuda)).cpu().numpy()
    L, dL = Cal.calib_terms_reg(z, y, np.float32(lt))
    np.testing.assert_allclose(got[0], L, rtol=2e-6, atol=1e-3)
    np.testing.assert_allclose(got[1], dL, rtol=2e-6, atol=1e-3)
    h = 1e-5  # the derivative is the derivative
    Lp, _ = Cal.calib_terms_reg(z, y, lt + h)
    Lm, _ = Cal.calib_terms_reg(z, y, lt - h)
    np.testing.assert_allclose(dL, (Lp - Lm) / (2 * h), rtol=1e-5, atol=1e-4)


def test_regression_temperature_calibration_recovers_variance_scale(cuda):
    """Predicted variances too small by e^1.2: Adam on the kernel's loss finds log_temp ~ 1.2."""
    from fortuna_b200.prob_model import Adam, CalibConfig, CalibOptimizer
    from fortuna_b200.prob_model.calibration import calibrate_temperature

    r = np.random.default_rng(4)
    B = 5000
    mu = r.standard_normal((1, B, 1))
    true_lv = -0.5 + 0.3 * r.standard_normal((1, B, 1))
    y = (mu + np.exp(0.5 * true_lv) * r.standard_normal((1, B, 1)))[0]
    z = torch.from_numpy(np.repeat(np.concatenate([mu, true_lv - 1.2], -1), 3, axis=0).astype(np.float32)).to(cuda)
    status, log_temp = calibrate_temperature(z, y.astype(np.float32), 1000, CalibConfig(
        optimizer=CalibOptimizer(Adam(5e-2), n_epochs=40)), regression=True)
    assert status["loss"][-1] < status["loss"][0] and abs(float(log_temp[0]) - 1.2) < 0.1


def test_regression_reductions_special_values(cuda):
    """inf / NaN means, +-inf and very large log-variances: NaN and inf land where the reference's float32 expressions
    put them (elementwise maps + max(0, .) that keeps a NaN)."""
    from fortuna_b200 import ops
    from oracle import predictive as P

    r = np.random.default_rng(0)
    S, B, d = 5, 40, 3
    z = r.standard_normal((S, B, 2 * d)).astype(np.float32)
    y = r.standard_normal((B, d)).astype(np.float32)
    z[0, 1, 0] = np.inf
    z[1, 2, d] = np.inf
    z[2, 3, d] = -np.inf
    z[3, 4, 1] = np.nan
    z[4, 5, d + 1] = 200.0
    z[0, 6, d + 2] = -200.0
    z[1, 7, :d] = 1e30
    with np.errstate(all="ignore"):
        want = {"mean": P.reg_mean(z), "aleatoric_variance": P.reg_aleatoric_variance(z),
                "epistemic_variance": P.reg_epistemic_variance(z), "log_prob": P.reg_log_prob(z, y),
                "ensemble_log_prob": P.reg_ensemble_log_prob(z, y)}
    got = ops.bma_reduce_reg(torch.from_numpy(z).to(cuda), tuple(want), targets=torch.from_numpy(y).to(cuda))
    for n, w in want.items():
        g, w = got[n].cpu().numpy(), np.asarray(w, np.float32)
        assert np.array_equal(np.isnan(g), np.isnan(w)) and np.array_equal(np.isinf(g), np.isinf(w)), n
        ok = np.isfinite(w)
        np.testing.assert_allclose(g[ok], w[ok], rtol=1e-5, atol=1e-6, err_msg=n)
