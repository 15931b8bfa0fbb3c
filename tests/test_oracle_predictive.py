"""Pins oracle/predictive.py: the reference's output-layer KATs
(tests/fortuna/test_prob_output_layer.py:38-43, 76-81) and analytic identities."""
import numpy as np
from scipy.special import logsumexp as sp_lse, softmax as sp_softmax

from oracle import predictive as P


def _z(S, B, C, seed=0, scale=4.0):
    return (scale * np.random.default_rng(seed).standard_normal((S, B, C))).astype(np.float32)


def test_reference_log_prob_kats():
    out = np.ones((1, 1, 2), np.float32)
    want = 1.0 - sp_lse([1.0, 1.0])
    np.testing.assert_allclose(P.cls_ensemble_log_prob(out, [0])[0, 0], want, rtol=1e-6)
    reg = np.ones((1, 1, 2), np.float32)
    np.testing.assert_allclose(
        P.reg_ensemble_log_prob(reg, np.zeros((1, 1), np.float32))[0, 0],
        -0.5 * (np.log(2 * np.pi) + 1 + np.exp(-1)), rtol=1e-6)


def test_against_float64():
    z = _z(7, 40, 33)
    t = np.random.default_rng(1).integers(0, 33, 40)
    p = sp_softmax(z.astype(np.float64), axis=-1)
    np.testing.assert_allclose(P.cls_mean(z), p.mean(0), rtol=1e-5)
    # 1 - p is rounded to float32 by the reference's formula: absolute error ~ulp(p), not relative
    np.testing.assert_allclose(P.cls_aleatoric_variance(z), (p * (1 - p)).mean(0), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(P.cls_epistemic_variance(z), (p**2).mean(0) - p.mean(0) ** 2, rtol=1e-3, atol=1e-7)
    m = p.mean(0)
    np.testing.assert_allclose(P.cls_entropy(z), -(m * np.log(m)).sum(-1), rtol=1e-5)
    np.testing.assert_allclose(P.cls_aleatoric_entropy(z), -(p * np.log(p)).sum(-1).mean(0), rtol=1e-5)
    ll = np.log(p)[:, np.arange(40), t]
    np.testing.assert_allclose(P.cls_log_prob(z, t), sp_lse(ll, axis=0) - np.log(7), rtol=1e-5, atol=1e-6)


def test_identities():
    z = _z(6, 50, 10, seed=2)
    np.testing.assert_allclose(P.cls_mean(z).sum(-1), 1.0, rtol=1e-6)
    np.testing.assert_allclose(P.cls_entropy(z), P.cls_aleatoric_entropy(z) + P.cls_epistemic_entropy(z), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(P.cls_variance(z), P.cls_aleatoric_variance(z) + P.cls_epistemic_variance(z))
    assert (P.cls_epistemic_variance(z[:1]) == 0).all()  # S = 1
    np.testing.assert_allclose(P.cls_entropy(np.zeros((3, 4, 10), np.float32)), np.log(10.0), rtol=1e-6)
    assert (P.cls_epistemic_entropy(z) > -1e-6).all()


def test_faithful_class_loop_matches_vectorised():
    z = _z(5, 20, 12, seed=3)
    for kind, fn in (("entropy", P.cls_entropy), ("aleatoric", P.cls_aleatoric_entropy), ("epistemic", P.cls_epistemic_entropy)):
        np.testing.assert_allclose(P.cls_entropy_faithful(z, kind), fn(z), rtol=1e-5, atol=2e-6)


def test_underflow_and_minus_inf_semantics():
    z = np.zeros((2, 3, 1000), np.float32)
    z[..., 1] = 200.0
    assert np.isfinite(P.cls_aleatoric_entropy(z)).all() and np.isfinite(P.cls_entropy(z)).all()
    z = _z(2, 3, 5, seed=4)
    z[0, :, 1] = -np.inf
    assert np.isnan(P.cls_aleatoric_entropy(z)).all()


def test_credible_set_and_mode():
    mean = np.array([[0.1, 0.6, 0.3], [0.5, 0.25, 0.25]], np.float32)
    assert P.cls_credible_set(mean, 0.05) == [[1, 2, 0], [0, 2, 1]]
    assert P.cls_credible_set(mean, 0.5) == [[1], [0, 2]]
    z = np.log(np.maximum(mean, 1e-9))[None]
    assert P.cls_mode(z).tolist() == [1, 0]


def test_last_layer_and_bf16_rounding():
    r = np.random.default_rng(0)
    x, w, b = r.standard_normal((9, 16)).astype(np.float32), r.standard_normal((3, 16, 5)).astype(np.float32), r.standard_normal((3, 5)).astype(np.float32)
    lt = np.array([0.3], np.float32)
    want = (np.einsum("bd,sdc->sbc", x.astype(np.float64), w.astype(np.float64)) + b[:, None]) * np.exp(-0.3)
    np.testing.assert_allclose(P.last_layer_logits(x, w, b, lt), want, rtol=1e-5, atol=1e-5)
    v = np.array([1.0, 1.00390625, 1.005859375, -3.1415927], np.float32)
    assert P.round_to_bf16(v).tolist() == [1.0, 1.0, 1.0078125, -3.140625]


def test_regression_maps():
    r = np.random.default_rng(1)
    z = r.standard_normal((6, 10, 4)).astype(np.float32)
    y = r.standard_normal((10, 2)).astype(np.float32)
    mu, lv = z[..., :2].astype(np.float64), z[..., 2:].astype(np.float64)
    np.testing.assert_allclose(P.reg_mean(z), mu.mean(0), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(P.reg_aleatoric_variance(z), np.exp(lv).mean(0), rtol=1e-5)
    ll = -0.5 * (np.exp(-lv) * (y - mu) ** 2 + lv + np.log(2 * np.pi)).sum(-1)
    np.testing.assert_allclose(P.reg_log_prob(z, y), sp_lse(ll, axis=0) - np.log(6), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(P.reg_calibrate(z, np.array([0.5], np.float32))[..., 2:], z[..., 2:] + 0.5, rtol=1e-6)


def test_cls_sample_targets_oracle_properties():
    """Class frequencies of the restated Predictive.sample approach softmax; C = 1 always draws class 0."""
    from oracle import prng

    z = np.tile(np.array([[[0.0, 1.0, 2.0, -1.0]]], np.float32), (1, 2000, 1))
    y, idx = P.cls_sample_targets(z, prng.PRNGKey(3), 3)
    assert y.shape == (3, 2000) and idx.tolist() == [0, 0, 0]
    freq = np.bincount(y.reshape(-1), minlength=4) / y.size
    np.testing.assert_allclose(freq, P.softmax(z[0, :1])[0], atol=0.02)
    assert not P.cls_sample_targets(np.zeros((2, 5, 1), np.float32), prng.PRNGKey(0), 4)[0].any()


def test_reg_mc_entropies_oracle_properties():
    """entropy = aleatoric + epistemic; one posterior sample has no epistemic part; the aleatoric entropy of a Gaussian
    approaches 0.5 log(2 pi e sigma^2) as the number of target samples grows."""
    from oracle import prng

    r = np.random.default_rng(0)
    z = np.concatenate([r.standard_normal((4, 6, 1)), np.full((4, 6, 1), -1.0)], axis=-1).astype(np.float32)
    e = P.reg_mc_entropies(z, prng.PRNGKey(1), 5, 7)
    np.testing.assert_allclose(e["aleatoric_entropy"] + e["epistemic_entropy"], e["entropy"], rtol=1e-5, atol=1e-5)
    one = P.reg_mc_entropies(z[:1], prng.PRNGKey(1), 1, 50)
    np.testing.assert_allclose(one["epistemic_entropy"], 0.0, atol=1e-6)
    big = P.reg_mc_entropies(z[:1], prng.PRNGKey(2), 1, 4000)
    np.testing.assert_allclose(big["aleatoric_entropy"], 0.5 * (np.log(2 * np.pi * np.e) - 1.0), atol=0.05)
