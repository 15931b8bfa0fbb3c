"""Temperature-scaling calibration loss restated in numpy.  TEST INFRASTRUCTURE ONLY.

Follows fortuna/prob_model/predictive/base.py:205-287 (``_batched_log_joint_prob`` with ``ensemble_outputs``) through
fortuna/likelihood/base.py:124-204 (batch log-likelihood sum scaled by n_data / batch) and
fortuna/output_calibrator/classification.py:7-17 (logits * exp(-log_temp)); optax 0.1.9 ``adam`` for the update.
Parity unpinned: the reference's tests run the calibration for smoke only (tests/fortuna/prob_model/test_train.py)."""
import numpy as np


def calib_terms(z, targets, log_temp, dtype=np.float64):
    """(L_s, dL_s/dtau) per posterior sample: L_s = sum_b log softmax(tau z[s,b])[y_b], tau = exp(-log_temp)."""
    z = np.asarray(z, dtype=dtype)
    y = np.asarray(targets).astype(np.int64)
    tau = np.exp(-dtype(log_temp))
    zc = tau * z
    m = zc.max(-1, keepdims=True)
    lse = (m + np.log(np.exp(zc - m).sum(-1, keepdims=True)))[..., 0]
    zy = np.take_along_axis(z, np.broadcast_to(y[None, :, None], z.shape[:2] + (1,)), axis=2)[..., 0]
    p = np.exp(zc - lse[..., None])
    return (tau * zy - lse).sum(axis=1), (zy - (p * z).sum(-1)).sum(axis=1)


def calib_loss(z, targets, log_temp, n_data, dtype=np.float64):
    """-(logsumexp_s((n_data / B) L_s) - log S)."""
    L, _ = calib_terms(z, targets, log_temp, dtype)
    a = (n_data / np.asarray(z).shape[1]) * L
    m = a.max()
    return -(m + np.log(np.exp(a - m).sum()) - np.log(len(a)))


def adam_step(grad, state, lr=1e-2, b1=0.9, b2=0.999, eps=1e-8):
    """optax.adam: scale_by_adam (bias-corrected moments, eps outside the square root) then scale(-lr)."""
    f = np.float32
    count = state["count"] + 1
    mu = f(b1) * state["mu"] + f(1 - b1) * f(grad)
    nu = f(b2) * state["nu"] + f(1 - b2) * f(grad) * f(grad)
    mu_hat = mu / f(1 - f(b1) ** count)
    nu_hat = nu / f(1 - f(b2) ** count)
    return f(-lr) * mu_hat / (np.sqrt(nu_hat) + f(eps)), {"count": count, "mu": mu, "nu": nu}


def calib_terms_reg(z, targets, log_temp, dtype=np.float64):
    """Regression temperature scaler (output_calibrator/regression.py:7-19: log var + log_temp):
    (L_s, dL_s/dlog_temp), L_s = sum_b log N(y_b; mu, exp(lv + lt))  (prob_output_layer/regression.py:30-34)."""
    z = np.asarray(z, dtype=dtype)
    d = z.shape[-1] // 2
    mu, lv = z[..., :d], z[..., d:] + dtype(log_temp)
    y = np.asarray(targets, dtype=dtype).reshape(1, z.shape[1], d)
    q = np.exp(-lv) * (y - mu) ** 2
    L = (-0.5 * (q + lv + np.log(2 * np.pi)).sum(-1)).sum(axis=1)
    dL = (-0.5 * (1.0 - q).sum(-1)).sum(axis=1)
    return L, dL
