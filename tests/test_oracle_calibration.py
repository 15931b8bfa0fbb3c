"""Oracle of the temperature-scaling calibration loss: derivative vs finite differences, Adam restatement, and the
product-side Adam (fortuna_b200/prob_model/calib_config.py) against it."""
import numpy as np

from oracle import calibration as Cal


def _data(S=5, B=40, C=7, seed=0):
    r = np.random.default_rng(seed)
    z = (3 * r.standard_normal((S, B, C))).astype(np.float32)
    y = r.integers(0, C, B).astype(np.int32)
    return z, y


def test_terms_derivative_matches_finite_differences():
    z, y = _data()
    for lt in (-0.5, 0.0, 0.8):
        _, dL = Cal.calib_terms(z, y, lt)
        h = 1e-6
        tau = np.exp(-lt)
        Lp, _ = Cal.calib_terms(z, y, -np.log(tau + h))
        Lm, _ = Cal.calib_terms(z, y, -np.log(tau - h))
        np.testing.assert_allclose(dL, (Lp - Lm) / (2 * h), rtol=1e-5, atol=1e-5)


def test_loss_is_minimised_at_the_generating_temperature():
    r = np.random.default_rng(1)
    z_true = 2 * r.standard_normal((1, 4000, 5))
    p = np.exp(z_true[0] - z_true[0].max(-1, keepdims=True))
    p /= p.sum(-1, keepdims=True)
    y = np.array([r.choice(5, p=pi) for pi in p])
    z = 3.0 * z_true  # over-confident by a factor 3: optimum at log_temp = log 3
    lts = np.linspace(0.0, 2.0, 41)
    best = lts[np.argmin([Cal.calib_loss(z, y, lt, 4000) for lt in lts])]
    assert abs(best - np.log(3.0)) < 0.15


def test_adam_restatements_agree():
    from fortuna_b200.prob_model.calib_config import Adam

    opt = Adam(1e-2)
    st = opt.init(np.zeros(1, np.float32))
    so = {"count": 0, "mu": np.float32(0), "nu": np.float32(0)}
    x = xo = np.float32(0)
    for g in (0.3, -1.2, 0.05, 2.0, -0.7):
        u, st = opt.update(np.array([g], np.float32), st)
        uo, so = Cal.adam_step(g, so)
        x, xo = x + u[0], xo + uo
        assert np.float32(u[0]) == np.float32(uo)
    # first step of Adam moves by -lr * sign(g)
    assert abs(Cal.adam_step(5.0, {"count": 0, "mu": np.float32(0), "nu": np.float32(0)})[0] + 1e-2) < 1e-6
