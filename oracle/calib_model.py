"""ORACLE (test infrastructure only - never imported by the product): CPU restatement of the calibration-model loss heads
and of optax.adam.  parity unpinned: the reference's tests hold no numbers for these; fixed by finite differences
(tests/test_oracle_calib_model.py) and by the closed form of Adam's first step.

    focal_loss_fn   fortuna/loss/classification/focal_loss.py:10-36
    scaled_mse_fn   fortuna/loss/regression/scaled_mse.py:7-30
    the step        fortuna/calib_model/calib_model_calibrator.py:29-66 + optax.adam (optax/_src/transform.py
                    scale_by_adam, bias_correction, scale_by_learning_rate), fortuna/calib_model/config/optimizer.py:20
"""
import numpy as np

F32 = np.float32


def softmax_rows(z, dtype=F32):
    z = np.asarray(z, dtype=dtype)
    e = np.exp(z - z.max(-1, keepdims=True))
    return (e / e.sum(-1, keepdims=True)).astype(dtype)


def focal_loss(outputs, targets, gamma=2.0, dtype=F32):
    """-mean((1 - p)^gamma log p), p = sum(softmax(outputs) * one_hot(targets), -1)."""
    p = softmax_rows(outputs, dtype)[np.arange(len(targets)), np.asarray(targets)]
    return (-np.mean((dtype(1) - p) ** dtype(gamma) * np.log(p))).astype(dtype)


def focal_loss_rows(outputs, targets, gamma=2.0, dtype=F32):
    p = softmax_rows(outputs, dtype)[np.arange(len(targets)), np.asarray(targets)]
    return (-((dtype(1) - p) ** dtype(gamma)) * np.log(p)).astype(dtype)


def focal_loss_grad(outputs, targets, gamma=2.0, dtype=np.float64):
    """d focal_loss / d outputs, analytic: (1/n) dl/dp p (onehot - softmax)."""
    z = np.asarray(outputs, dtype=dtype)
    n, C = z.shape
    sm = softmax_rows(z, dtype)
    y = np.asarray(targets)
    p = sm[np.arange(n), y]
    q = 1 - p
    dldp = -(q ** gamma) / p
    if gamma != 0:
        dldp = dldp + gamma * q ** (gamma - 1) * np.log(p)
    oh = np.zeros((n, C), dtype=dtype)
    oh[np.arange(n), y] = 1
    return ((dldp * p)[:, None] * (oh - sm) / n).astype(dtype)


def scaled_mse(outputs, targets, dtype=F32):
    o = np.asarray(outputs, dtype=dtype)
    d = o.shape[-1] // 2
    mu, lv = o[:, :d], o[:, d:]
    t = np.asarray(targets, dtype=dtype).reshape(mu.shape)
    return np.mean(np.sum(np.exp(-lv) * (t - mu) ** 2 + lv, -1)).astype(dtype)


def scaled_mse_grad(outputs, targets, dtype=np.float64):
    o = np.asarray(outputs, dtype=dtype)
    n, d = o.shape[0], o.shape[-1] // 2
    mu, lv = o[:, :d], o[:, d:]
    t = np.asarray(targets, dtype=dtype).reshape(mu.shape)
    e = np.exp(-lv)
    return (np.concatenate([-2 * e * (t - mu), 1 - e * (t - mu) ** 2], -1) / n).astype(dtype)


class Adam:
    """optax.adam(lr, b1, b2, eps): scale_by_adam then scale(-lr), float32 throughout."""

    def __init__(self, lr=1e-2, b1=0.9, b2=0.999, eps=1e-8):
        self.lr, self.b1, self.b2, self.eps = F32(lr), F32(b1), F32(b2), F32(eps)

    def init(self, params):
        p = np.asarray(params, dtype=F32)
        return {"count": 0, "mu": np.zeros_like(p), "nu": np.zeros_like(p)}

    def step(self, params, grads, state):
        g = np.asarray(grads, dtype=F32)
        count = state["count"] + 1
        mu = ((F32(1) - self.b1) * g + self.b1 * state["mu"]).astype(F32)
        nu = ((F32(1) - self.b2) * (g * g) + self.b2 * state["nu"]).astype(F32)
        c1 = F32(1) - self.b1 ** F32(count)
        c2 = F32(1) - self.b2 ** F32(count)
        u = ((mu / c1) / (np.sqrt(nu / c2) + self.eps)).astype(F32)
        return (np.asarray(params, dtype=F32) + (-self.lr) * u).astype(F32), {"count": count, "mu": mu, "nu": nu}


def linear_calibration(W, b, loader, gamma=2.0, n_epochs=3, lr=1e-2):
    """CalibClassifier.calibrate on a single Dense layer logits = x W^T + b (torch layout W [C, D]) with the focal loss
    and Adam: per-epoch mean losses and the final (W, b).  ``loader``: a list of (x, y) batches, visited every epoch."""
    W, b = np.asarray(W, dtype=F32).copy(), np.asarray(b, dtype=F32).copy()
    C, D = W.shape
    opt = Adam(lr)
    flat = np.concatenate([b.ravel(), W.ravel()])  # leaf order of the flat tree: sorted names ("bias" < "weight")
    st = opt.init(flat)
    losses = []
    for _ in range(n_epochs):
        ep = []
        for x, y in loader:
            x = np.asarray(x, dtype=F32)
            z = (x @ W.T + b).astype(F32)
            ep.append(float(focal_loss(z, y, gamma)))
            g = focal_loss_grad(z, y, gamma).astype(F32)
            gW, gb = (g.T @ x).astype(F32), g.sum(0).astype(F32)
            flat, st = opt.step(np.concatenate([b.ravel(), W.ravel()]), np.concatenate([gb.ravel(), gW.ravel()]), st)
            b, W = flat[:C].copy(), flat[C:].reshape(C, D).copy()
        losses.append(np.mean(ep))
    return np.array(losses, dtype=F32), W, b
