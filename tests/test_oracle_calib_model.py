"""The calibration-model oracle against finite differences (float64) and Adam's closed-form first step."""
import numpy as np

from oracle import calib_model as O


def _fd(f, z, eps=1e-6):
    g = np.zeros_like(z)
    for i in np.ndindex(*z.shape):
        zp, zm = z.copy(), z.copy()
        zp[i] += eps
        zm[i] -= eps
        g[i] = (f(zp) - f(zm)) / (2 * eps)
    return g


def test_focal_grad_matches_finite_differences():
    r = np.random.default_rng(0)
    z = r.standard_normal((5, 7))
    y = r.integers(0, 7, 5)
    for gamma in (0.0, 1.0, 2.0, 3.5):
        g = O.focal_loss_grad(z, y, gamma)
        fd = _fd(lambda v: float(O.focal_loss(v, y, gamma, dtype=np.float64)), z)
        np.testing.assert_allclose(g, fd, rtol=1e-6, atol=1e-9)


def test_scaled_mse_grad_matches_finite_differences():
    r = np.random.default_rng(1)
    o = r.standard_normal((6, 4))
    t = r.standard_normal((6, 2))
    g = O.scaled_mse_grad(o, t)
    fd = _fd(lambda v: float(O.scaled_mse(v, t, dtype=np.float64)), o)
    np.testing.assert_allclose(g, fd, rtol=1e-6, atol=1e-9)


def test_adam_first_step_is_lr_times_sign():
    opt = O.Adam(1e-2)
    p = np.array([1.0, -2.0, 3.0], np.float32)
    g = np.array([0.5, -4.0, 1e-3], np.float32)
    q, st = opt.step(p, g, opt.init(p))
    np.testing.assert_allclose(q, p - 1e-2 * np.sign(g), rtol=0, atol=2e-7)  # mu/c1 = g, sqrt(nu/c2) = |g|
    assert st["count"] == 1


def test_linear_calibration_lowers_the_loss():
    r = np.random.default_rng(2)
    x = r.standard_normal((200, 5)).astype(np.float32)
    y = (x[:, 0] > 0).astype(np.int32) + (x[:, 1] > 0).astype(np.int32)
    W = 0.1 * r.standard_normal((3, 5)).astype(np.float32)
    losses, _, _ = O.linear_calibration(W, np.zeros(3, np.float32), [(x[:100], y[:100]), (x[100:], y[100:])], n_epochs=20)
    assert losses[-1] < losses[0]
