"""Output calibration models restated in numpy.  TEST INFRASTRUCTURE ONLY.

Follows fortuna/output_calib_model/loss.py:26-79 (calibrate the outputs, apply ``loss_fn``) with the default loss
functions fortuna/loss/classification/focal_loss.py:10-36 and fortuna/loss/regression/scaled_mse.py:6-26, the
temperature scalers fortuna/output_calibrator/classification.py:7-17 (logits * exp(-log_temp)) and regression.py:7-19
(log var + log_temp), the one-step-per-epoch loop of
fortuna/output_calib_model/output_calib_model_calibrator.py:170-257 with optax 0.1.9 ``adam``, and the output-layer
target samples fortuna/prob_output_layer/classification.py:33-51 / regression.py:38-52.
Parity unpinned: the reference's own tests of this path are dry runs (tests/fortuna/calib_model/
test_output_calib_model.py); the derivatives below are checked against central differences in float64 and the random
draws ride on oracle/prng.py, which is pinned to jax's threefry known answers."""
import numpy as np

from . import prng
from .calibration import adam_step

F32 = np.float32


def focal_loss(z, targets, log_temp=0.0, gamma=2.0, dtype=np.float64):
    """(loss, dloss/dlog_temp): loss = -mean_b (1 - p_b)^gamma log p_b, p_b = softmax(tau z_b)[y_b], tau = exp(-log_temp)."""
    z = np.asarray(z, dtype=dtype)
    y = np.asarray(targets).astype(np.int64)
    tau = np.exp(-dtype(log_temp))
    zc = tau * z
    m = zc.max(-1, keepdims=True)
    e = np.exp(zc - m)
    probs = e / e.sum(-1, keepdims=True)
    rows = np.arange(z.shape[0])
    p = probs[rows, y]
    logp = np.log(p)
    q = 1.0 - p
    loss = -np.mean(q**gamma * logp)
    # d log p / d tau = z_y - sum_c p_c z_c;  d/dlog p [(1-p)^g log p] = (1-p)^g - g (1-p)^(g-1) p log p
    dlogp = z[rows, y] - (probs * z).sum(-1)
    dpow = np.zeros_like(p) if gamma == 0 else gamma * q ** (gamma - 1.0)
    dl = q**gamma - dpow * p * logp
    dloss_dtau = -np.mean(dl * dlogp)
    return dtype(loss), dtype(-tau * dloss_dtau)


def scaled_mse(z, targets, log_temp=0.0, dtype=np.float64):
    """(loss, dloss/dlog_temp): mean_b sum_k ( exp(-lv') (y - mu)^2 + lv' ), lv' = log var + log_temp."""
    z = np.asarray(z, dtype=dtype)
    d = z.shape[-1] // 2
    mu, lv = z[:, :d], z[:, d:] + dtype(log_temp)
    y = np.asarray(targets, dtype=dtype).reshape(z.shape[0], d)
    q = np.exp(-lv) * (y - mu) ** 2
    return dtype(np.mean((q + lv).sum(-1))), dtype(np.mean((1.0 - q).sum(-1)))


def calibrate(z, targets, n_epochs, regression=False, gamma=2.0, lr=1e-2):
    """The reference's loop: every epoch is one Adam step on the whole calibration set.  Returns (losses, log_temp
    trajectory) with log_temp carried in float32 as the reference's parameter is."""
    lt = F32(0.0)
    state = {"count": 0, "mu": F32(0.0), "nu": F32(0.0)}
    losses, traj = [], []
    for _ in range(n_epochs):
        loss, g = scaled_mse(z, targets, lt) if regression else focal_loss(z, targets, lt, gamma)
        upd, state = adam_step(F32(g), state, lr=lr)
        lt = F32(lt + upd)
        losses.append(F32(loss))
        traj.append(lt)
    return np.array(losses, F32), np.array(traj, F32)


def cls_output_sample(logits, key, n_target_samples, log_temp=None):
    """ClassificationProbOutputLayer.sample: keys = split(rng, n); y[:, b] = choice(keys[b], C, p=probs[b], shape=(T,)) =
    searchsorted(cumsum(p), cumsum(p)[-1] * (1 - uniform(keys[b], (T,)))).  Returns int32 [T, n]."""
    from .predictive import softmax
    from .scoring import cumsum_f32

    z = np.asarray(logits, dtype=F32)
    if log_temp is not None:
        z = (z * np.exp(-F32(np.asarray(log_temp).reshape(())), dtype=F32)).astype(F32)
    n = z.shape[0]
    p_cuml = cumsum_f32(softmax(z))
    keys = prng.split(np.asarray(key, dtype=np.uint32), n)
    out = np.empty((n_target_samples, n), np.int32)
    for b in range(n):
        u = prng.uniform(keys[b], n_target_samples).astype(F32)
        r = (p_cuml[b, -1] * (F32(1) - u)).astype(F32)
        out[:, b] = (p_cuml[b][None, :] < r[:, None]).sum(axis=1)
    return out


def reg_output_sample(outputs, key, n_target_samples, log_temp=None):
    """RegressionProbOutputLayer.sample: mu + exp(0.5 lv') * normal(rng, (T, n, d)).  Returns float32 [T, n, d]."""
    z = np.asarray(outputs, dtype=F32)
    n, d = z.shape[0], z.shape[1] // 2
    mu, lv = z[:, :d], z[:, d:]
    if log_temp is not None:
        lv = (lv + F32(np.asarray(log_temp).reshape(()))).astype(F32)
    eps = prng.normal(np.asarray(key, dtype=np.uint32), n_target_samples * n * d).reshape(n_target_samples, n, d)
    return (mu[None] + np.exp(F32(0.5) * lv, dtype=F32)[None] * eps).astype(F32)


def reg_output_entropy(outputs, key, n_target_samples, log_temp=None):
    """RegressionProbOutputLayer.entropy: -mean_t log N(y_t; outputs) over the samples above.  float32 [n]."""
    z = np.asarray(outputs, dtype=F32)
    n, d = z.shape[0], z.shape[1] // 2
    mu, lv = z[:, :d], z[:, d:]
    if log_temp is not None:
        lv = (lv + F32(np.asarray(log_temp).reshape(()))).astype(F32)
    y = reg_output_sample(outputs, key, n_target_samples, log_temp)
    lp = F32(-0.5) * (np.exp(-lv, dtype=F32)[None] * (y - mu[None]) ** 2 + lv[None] + F32(np.log(2 * np.pi))).sum(-1)
    return (-lp.mean(axis=0)).astype(F32)
