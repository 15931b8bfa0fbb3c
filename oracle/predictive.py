"""Bayesian-model-averaging predictive arithmetic restated in numpy float32.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows:
  fortuna/likelihood/base.py:429-458 restricted to the last ``nn.Dense`` (fortuna/model/mlp.py:227,
      lenet.py:129, resnet.py:248) + fortuna/output_calibrator/classification.py:14-17 (x * exp(-log_temp)),
      fortuna/output_calibrator/regression.py:16-19 (log_var += log_temp)
  fortuna/prob_output_layer/classification.py:25-28 (log_prob), :54-69 (mean = softmax), :88-104 (p(1-p))
  fortuna/prob_output_layer/regression.py:30-34 (Gaussian log_prob), :67-74 (mean / variance)
  fortuna/likelihood/classification.py:102-114,181-199, fortuna/likelihood/regression.py:55-67,87-99
  fortuna/prob_model/predictive/base.py:104-128 (ensemble_log_prob), :179-203 (log_prob),
      :624-647 (mean), :735-758 (aleatoric variance), :806-836 (epistemic variance), :883-898, :944
  fortuna/prob_model/predictive/classification.py:86 (mode), :300-310, :359-374, :419-432 (entropies),
      :465-483 (credible set)
with ``jax.nn.softmax`` = exp(z - max) / sum and ``jsp.special.logsumexp`` = log(sum(exp(z - m))) + m,
m = max replaced by 0 when non-finite (jax 0.4.30).

All functions take the *calibrated* logits ``z`` of shape [S, B, C] (float32) for the S posterior
samples actually evaluated (the index draw over the sample bank is oracle.prng.choice).
Sums over samples are sequential in s like the reference's ``lax.fori_loop``; sums over classes use
numpy's pairwise sum (XLA's reduction order is unspecified, so these agree to float32 rounding only).
Parity unpinned: the reference tests check shapes and ``variance >= 0`` only
(tests/fortuna/test_predictive.py:61-230); guarded by identities in tests/test_oracle_predictive.py.
"""

import numpy as np

F32 = np.float32


def _f32(x):
    return np.asarray(x, dtype=F32)


# ----------------------------------------------------------------------------- last layer
def last_layer_logits(x, w, bias, log_temp=None, bf16_inputs=False):
    """logits[s,b,c] = x[b,:] @ w[s,:,c] + bias[s,c], then * exp(-log_temp).

    x [B,D], w [S,D,C], bias [S,C] or None.  float32 path: float32 products accumulated by numpy's
    BLAS (order unspecified, like XLA's dot).  bf16_inputs=True: x and w are first rounded to
    bfloat16 (round-to-nearest-even), products accumulated in float64 and rounded once to float32 -
    the exact value a float32-accumulating tensor-core GEMM approximates.
    """
    x = _f32(x)
    w = _f32(w)
    if bf16_inputs:
        x = round_to_bf16(x).astype(np.float64)
        w = round_to_bf16(w).astype(np.float64)
        out = np.einsum("bd,sdc->sbc", x, w).astype(F32)
    else:
        out = np.stack([x @ w[s] for s in range(w.shape[0])]).astype(F32)
    if bias is not None:
        out = out + _f32(bias)[:, None, :]
    if log_temp is not None:
        out = out * np.exp(-_f32(log_temp).reshape(()), dtype=F32)
    return out.astype(F32)


def round_to_bf16(a):
    """float32 -> nearest-even bfloat16, returned as float32."""
    a = _f32(a)
    u = a.view(np.uint32).astype(np.uint64)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    out = rounded.astype(np.uint32).view(F32)
    return np.where(np.isnan(a), a, out).astype(F32)


# ----------------------------------------------------------------------------- per-sample maps
def softmax(z):
    z = _f32(z)
    m = np.max(z, axis=-1, keepdims=True)
    with np.errstate(invalid="ignore"):
        e = np.exp(z - m, dtype=F32)
    return (e / np.sum(e, axis=-1, keepdims=True, dtype=F32)).astype(F32)


def logsumexp(z, axis=-1):
    z = _f32(z)
    m = np.max(z, axis=axis, keepdims=True)
    m = np.where(np.isfinite(m), m, F32(0)).astype(F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        out = np.log(np.sum(np.exp(z - m, dtype=F32), axis=axis, keepdims=True, dtype=F32), dtype=F32) + m
    return np.squeeze(out, axis=axis).astype(F32)


def log_softmax_rows(z):
    """l[.., c] = z[.., c] - logsumexp_c(z)  (prob_output_layer/classification.py:25-28 for every class)."""
    z = _f32(z)
    with np.errstate(invalid="ignore"):
        return (z - logsumexp(z, -1)[..., None]).astype(F32)


# ----------------------------------------------------------------------------- classification reductions
def cls_mean(z):
    """predictive/base.py:645-647: sequential sum over s of softmax, then / S."""
    z = _f32(z)
    acc = softmax(z[0])
    for s in range(1, z.shape[0]):
        acc = acc + softmax(z[s])
    return (acc / F32(z.shape[0])).astype(F32)


def cls_mode(z):
    """predictive/classification.py:86: argmax of the mean (first maximum)."""
    return np.argmax(cls_mean(z), axis=-1).astype(np.int32)


def cls_aleatoric_variance(z):
    z = _f32(z)
    acc = None
    for s in range(z.shape[0]):
        p = softmax(z[s])
        t = p * (F32(1) - p)
        acc = t if acc is None else acc + t
    return (acc / F32(z.shape[0])).astype(F32)


def cls_epistemic_variance(z):
    z = _f32(z)
    S = F32(z.shape[0])
    a = None
    b = None
    for s in range(z.shape[0]):
        p = softmax(z[s])
        a = p if a is None else a + p
        b = p * p if b is None else b + p * p
    m = a / S
    return np.maximum(F32(0), b / S - m * m).astype(F32)


def cls_variance(z):
    return (cls_aleatoric_variance(z) + cls_epistemic_variance(z)).astype(F32)


def cls_std(z):
    return np.sqrt(cls_variance(z)).astype(F32)


def _mean_p_logp(l):
    """mean over s of exp(l) * l, [S,B,C] -> [B,C]."""
    with np.errstate(invalid="ignore"):
        t = np.exp(l, dtype=F32) * l
    return (np.sum(t, axis=0, dtype=F32) / F32(l.shape[0])).astype(F32)


def cls_aleatoric_entropy(z):
    """predictive/classification.py:300-310."""
    l = log_softmax_rows(z)
    return (-np.sum(_mean_p_logp(l), axis=-1, dtype=F32)).astype(F32)


def _log_preds(l):
    """lp[b,c] = logsumexp_s l[s,b,c] - log(S)."""
    return (logsumexp(l, axis=0) - np.log(F32(l.shape[0]), dtype=F32)).astype(F32)


def cls_entropy(z):
    """predictive/classification.py:419-432."""
    lp = _log_preds(log_softmax_rows(z))
    with np.errstate(invalid="ignore"):
        return (-np.sum(np.exp(lp, dtype=F32) * lp, axis=-1, dtype=F32)).astype(F32)


def cls_epistemic_entropy(z):
    """predictive/classification.py:359-374."""
    l = log_softmax_rows(z)
    lp = _log_preds(l)
    with np.errstate(invalid="ignore"):
        term = np.exp(lp, dtype=F32) * lp - _mean_p_logp(l)
    return (-np.sum(term, axis=-1, dtype=F32)).astype(F32)


def cls_entropy_faithful(z, kind="entropy"):
    """The reference's literal structure: a loop over classes, each one re-evaluating
    log_prob = onehot . z - logsumexp_c z over the whole [S,B,C] tensor (O(S*B*C^2)).
    Used to cross-check the vectorised forms above and as the 'faithful' CPU timing variant."""
    z = _f32(z)
    S, B, C = z.shape
    out = np.zeros(B, dtype=F32)
    lse = logsumexp(z, -1)
    for c in range(C):
        onehot = np.zeros(C, dtype=F32)
        onehot[c] = 1
        ll = (np.sum(z * onehot, axis=-1, dtype=F32) - lse).astype(F32)  # [S,B]
        mpl = np.sum(np.exp(ll, dtype=F32) * ll, axis=0, dtype=F32) / F32(S)
        if kind == "aleatoric":
            term = mpl
        else:
            lp = logsumexp(ll, axis=0) - np.log(F32(S), dtype=F32)
            term = np.exp(lp, dtype=F32) * lp
            if kind == "epistemic":
                term = term - mpl
        out = out + term
    return (-out).astype(F32)


def cls_ensemble_log_prob(z, targets):
    """[S,B]: z[s,b,y_b] - logsumexp_c z[s,b,:]  (the matrix ensemble_log_prob returns transposed)."""
    z = _f32(z)
    t = np.asarray(targets).astype(np.int64)
    picked = np.take_along_axis(z, np.broadcast_to(t[None, :, None], z.shape[:2] + (1,)), axis=2)[..., 0]
    return (picked - logsumexp(z, -1)).astype(F32)


def cls_log_prob(z, targets):
    """predictive/base.py:201-203: logsumexp_s(ll) - log(S)."""
    ll = cls_ensemble_log_prob(z, targets)
    return (logsumexp(ll, axis=0) - np.log(F32(z.shape[0]), dtype=F32)).astype(F32)


def cls_credible_set(mean, error):
    """predictive/classification.py:471-483 -> list of lists of class labels."""
    mean = _f32(mean)
    labels = np.argsort(mean, axis=-1, kind="stable")[:, ::-1]
    srt = np.sort(mean, axis=-1, kind="stable")[:, ::-1]
    from .scoring import cumsum_f32  # jnp.cumsum's association order (associative_scan), see oracle/scoring.py

    region = cumsum_f32(srt) > F32(1 - error)
    first = np.argmax(region, axis=1)
    return [labels[i, : first[i] + 1].tolist() for i in range(mean.shape[0])]


# ----------------------------------------------------------------------------- regression maps
def reg_split(z):
    z = _f32(z)
    d = z.shape[-1] // 2
    return z[..., :d], z[..., d:]


def reg_calibrate(z, log_temp):
    """output_calibrator/regression.py:16-19: log_var += log_temp."""
    mu, lv = reg_split(z)
    return np.concatenate([mu, lv + _f32(log_temp).reshape(())], axis=-1).astype(F32)


def reg_mean(z):
    mu, _ = reg_split(z)
    acc = mu[0].copy()
    for s in range(1, mu.shape[0]):
        acc = acc + mu[s]
    return (acc / F32(mu.shape[0])).astype(F32)


def reg_aleatoric_variance(z):
    _, lv = reg_split(z)
    acc = np.exp(lv[0], dtype=F32)
    for s in range(1, lv.shape[0]):
        acc = acc + np.exp(lv[s], dtype=F32)
    return (acc / F32(lv.shape[0])).astype(F32)


def reg_epistemic_variance(z):
    mu, _ = reg_split(z)
    S = F32(mu.shape[0])
    a = mu[0].copy()
    b = mu[0] * mu[0]
    for s in range(1, mu.shape[0]):
        a = a + mu[s]
        b = b + mu[s] * mu[s]
    m = a / S
    return np.maximum(F32(0), b / S - m * m).astype(F32)


_LOG2PI = np.log(F32(2) * F32(np.pi), dtype=F32)


def reg_ensemble_log_prob(z, targets):
    """[S,B]: -0.5 * sum_d( exp(-lv) (y - mu)^2 + lv + log 2 pi )."""
    mu, lv = reg_split(z)
    y = _f32(targets)[None]
    d = y - mu
    t = np.exp(-lv, dtype=F32) * (d * d) + lv + _LOG2PI
    return (F32(-0.5) * np.sum(t, axis=-1, dtype=F32)).astype(F32)


def reg_log_prob(z, targets):
    ll = reg_ensemble_log_prob(z, targets)
    return (logsumexp(ll, axis=0) - np.log(F32(ll.shape[0]), dtype=F32)).astype(F32)


def reg_sample_targets(outputs_stored, key, n_target_samples, log_temp=None):
    """Predictive.sample for regression (predictive/base.py:407-441, likelihood/base.py:363-369,
    prob_output_layer/regression.py:38-52) on ONE input batch: keys = split(rng, T); per key (k1, k2) = split(key);
    posterior sample i = choice(k1, n_stored); y = mu_i + exp(0.5 logvar_i) * normal(k2, (1,) + mu.shape).
    outputs_stored: [n_stored, B, 2d].  Returns (y [T, B, d], idx [T])."""
    from . import prng

    z = _f32(outputs_stored)
    n, B, two_d = z.shape
    d = two_d // 2
    keys = prng.split(np.asarray(key, dtype=np.uint32), n_target_samples)
    ys, idx = [], []
    for k in keys:
        k1, k2 = prng.split(k, 2)
        i = prng.choice(k1, n)
        zi = z[i] if log_temp is None else reg_calibrate(z[i : i + 1], log_temp)[0]
        mu, lv = zi[:, :d], zi[:, d:]
        eps = prng.normal(k2, B * d).reshape(B, d)
        ys.append((mu + np.exp(F32(0.5) * lv, dtype=F32) * eps).astype(F32))
        idx.append(i)
    return np.stack(ys), np.array(idx, dtype=np.int32)


def cls_sample_targets(logits_stored, key, n_target_samples, log_temp=None):
    """Predictive.sample for classification (predictive/base.py:406-440, likelihood/base.py:363-369,
    prob_output_layer/classification.py:33-51) on ONE input batch.  jax.random.choice(key, C, p=p, shape=(1,)) is
    ``p_cuml = cumsum(p); r = p_cuml[-1] * (1 - uniform(key, (1,))); searchsorted(p_cuml, r)`` (jax/_src/random.py,
    ``choice`` with replacement), cumsum in associative_scan order (oracle/scoring.py).
    logits_stored: [n_stored, B, C].  Returns (y int32 [T, B], idx int32 [T])."""
    from . import prng
    from .scoring import cumsum_f32

    z = _f32(logits_stored)
    n, B, Cn = z.shape
    keys = prng.split(np.asarray(key, dtype=np.uint32), n_target_samples)
    ys, idx = [], []
    for k in keys:
        k1, k2 = prng.split(k, 2)
        i = prng.choice(k1, n)
        zi = z[i] if log_temp is None else (z[i] * np.exp(-_f32(log_temp).reshape(()), dtype=F32)).astype(F32)
        p_cuml = cumsum_f32(softmax(zi))
        kb = prng.split(k2, B)
        u = np.array([prng.uniform(kb[b], 1)[0] for b in range(B)], dtype=F32)
        r = (p_cuml[:, -1] * (F32(1) - u)).astype(F32)
        ys.append((p_cuml < r[:, None]).sum(axis=1).astype(np.int32))
        idx.append(i)
    return np.stack(ys), np.array(idx, dtype=np.int32)


def reg_mc_entropies(outputs_stored, key, n_posterior_samples, n_target_samples, log_temp=None):
    """RegressionPredictive.aleatoric_entropy / epistemic_entropy / entropy (predictive/regression.py:51-267) for ONE rng:
    key1, *keys = split(rng, 1 + S); outputs_i drawn with key1 (idx_s = choice(split(key1, S)[s], n_stored),
    predictive/base.py:487-499); y_i = sample(T, outputs_i, keys[i]); the three Monte-Carlo means.
    Returns dict(entropy, aleatoric_entropy, epistemic_entropy) of float32 [B]."""
    from . import prng

    z = _f32(outputs_stored)
    n, B, two_d = z.shape
    d = two_d // 2
    S, T = int(n_posterior_samples), int(n_target_samples)
    ks = prng.split(np.asarray(key, dtype=np.uint32), 1 + S)
    key1, keys = ks[0], ks[1:]
    idx = [prng.choice(k, n) for k in prng.split(key1, S)]
    ens = np.stack([z[i] for i in idx])
    if log_temp is not None:
        ens = reg_calibrate(ens, log_temp)
    mu, lv = reg_split(ens)
    ys = []
    for i in range(S):
        eps = prng.normal(keys[i], T * B * d).reshape(T, B, d)
        ys.append((mu[i][None] + np.exp(F32(0.5) * lv[i], dtype=F32)[None] * eps).astype(F32))
    log_S = np.log(F32(S), dtype=F32)
    alea = np.zeros(B, F32)
    ent = np.zeros(B, F32)
    epi = np.zeros(B, F32)
    for i in range(S):
        log_liks = np.stack([reg_ensemble_log_prob(ens[i : i + 1], ys[i][t])[0] for t in range(T)])       # [T, B]
        log_preds = np.stack([logsumexp(reg_ensemble_log_prob(ens, ys[i][t]), axis=0) - log_S for t in range(T)])
        alea = alea - log_liks.mean(axis=0, dtype=F32)
        ent = ent - log_preds.mean(axis=0, dtype=F32)
        epi = epi - (log_preds - log_liks).mean(axis=0, dtype=F32)
    return {"entropy": (ent / F32(S)).astype(F32), "aleatoric_entropy": (alea / F32(S)).astype(F32),
            "epistemic_entropy": (epi / F32(S)).astype(F32)}


def quantile_axis0(x, q):
    """jnp.quantile(x, q, axis=0), method="linear"."""
    from .scoring import quantile_linear

    x = _f32(x)
    T = x.shape[0]
    flat = x.reshape(T, -1)
    q = np.atleast_1d(_f32(q))
    out = np.stack([quantile_linear(flat[:, m], q) for m in range(flat.shape[1])], axis=1)
    return out.reshape((len(q),) + x.shape[1:])
