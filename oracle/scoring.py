"""Conformal / calibration scoring arithmetic restated in numpy.  TEST INFRASTRUCTURE ONLY.

Follows:
  fortuna/conformal/classification/adaptive_prediction.py:11-22  (APS score)
  fortuna/conformal/classification/simple_prediction.py:10-18    (simple score)
  fortuna/conformal/classification/base.py:45-64, 72-82          (rank-count sets, all-class test scores)
  fortuna/metric/classification.py:16-40, 43-95, 98-133, 147-181, 195-220
  fortuna/conformal/regression/quantile.py:89-90, onedim_uncertainty.py:89-90 (jnp.quantile)
with jax 0.4.30 semantics for ``jnp.argsort`` (stable ascending), ``jnp.linspace`` (start*(1-i/(n-1)) +
stop*(i/(n-1)), last point = stop) and ``jnp.quantile`` (method="linear").

Order of float additions in ``jnp.cumsum``: at jax 0.4.30 the cumulative-reduction primitives lower, on every
platform but TPU, through ``jax.lax.associative_scan`` (jax/_src/lax/control_flow/loops.py,
``_cumulative_reduction_primitive``: default lowering ``partial(associative_scan, reduce_fn)``; reduce_window is
registered for 'tpu' only, a size-dependent choice for 'cuda'/'rocm').  The reference's JAX-CPU path therefore
sums in associative_scan's recursive pairwise order, restated in ``cumsum_f32`` below; the CUDA kernels follow the
same tree, so scores, sets and bins are compared bit-for-bit.  (JAX is not installable here, so this order is
restated from the pinned source, not observed.)
Parity pinned only by the simple-score case of tests/fortuna/test_conformal_methods.py:37-53;
everything else here is "parity unpinned" (reference tests check shapes only; ECE has no test).
"""

import numpy as np

F32 = np.float32


def _f32(x):
    return np.asarray(x, dtype=F32)


# ----------------------------------------------------------------------------- conformal scores
def cumsum_f32(x):
    """jnp.cumsum(x, axis=-1) in float32 == jax.lax.associative_scan(add, x): combine adjacent pairs, scan the
    half-length array recursively (those are the odd positions), then even position 2i = odd[i-1] + x[2i]."""
    x = _f32(x)
    n = x.shape[-1]
    if n < 2:
        return x.copy()
    reduced = (x[..., 0:-1:2] + x[..., 1::2]).astype(F32)
    odd = cumsum_f32(reduced)
    if n % 2 == 0:
        even = (odd[..., :-1] + x[..., 2::2]).astype(F32)
    else:
        even = (odd + x[..., 2::2]).astype(F32)
    out = np.empty_like(x)
    out[..., 0] = x[..., 0]
    out[..., 2::2] = even
    out[..., 1::2] = odd
    return out


def descending_perm(probs):
    """argsort(probs, axis=1)[:, ::-1] with a stable ascending sort: descending values, ties broken
    by LARGER index first."""
    return np.argsort(_f32(probs), axis=1, kind="stable")[:, ::-1]


def aps_cumsums(probs):
    """[B,C] inclusive float32 prefix sums of each row sorted descending, un-permuted back so that
    out[b, c] = sum of all probs ranked at or before class c.  This is the [B,C] test-score matrix of
    base.py:76-81 (out_axes=1)."""
    probs = _f32(probs)
    perm = descending_perm(probs)
    srt = np.take_along_axis(probs, perm, axis=1)
    cs = cumsum_f32(srt)
    out = np.empty_like(cs)
    np.put_along_axis(out, perm, cs, axis=1)
    return out


def gather_index(targets, n):
    """jnp indexing with a traced index: negative indices count from the end, then the gather clamps to the array
    (jax "sharp bits": ``jnp.arange(10)[11] == 9``)."""
    t = np.asarray(targets).astype(np.int64)
    return np.clip(np.where(t < 0, t + n, t), 0, n - 1)


def aps_score(probs, targets):
    """adaptive_prediction.score_fn: cumsum(probs[perm])[inv_perm[target]]."""
    allc = aps_cumsums(probs)
    t = gather_index(targets, allc.shape[1])
    return allc[np.arange(allc.shape[0]), t]


def simple_score(probs, targets):
    """simple_prediction.score_fn: 1 - probs[target]."""
    probs = _f32(probs)
    t = gather_index(targets, probs.shape[1])
    return (F32(1) - probs[np.arange(probs.shape[0]), t]).astype(F32)


def simple_scores_all(probs):
    return (F32(1) - _f32(probs)).astype(F32)


def get_scores(val_probs, val_targets, test_probs, kind="aps"):
    """SplitConformalClassifier.get_scores (base.py:72-82) -> (val_scores [n], test_scores [B,C])."""
    if kind == "aps":
        return aps_score(val_probs, val_targets), aps_cumsums(test_probs)
    return simple_score(val_probs, val_targets), simple_scores_all(test_probs)


def conformal_threshold(n_val, error):
    """float32 right-hand side of base.py:51-53: (1 - error) * (len(val_scores) + 1)."""
    return F32((1 - error) * (n_val + 1))


def conformal_conds_literal(val_scores, test_scores, error):
    """base.py:51-53 verbatim: broadcast compare [n,B,C], integer count, float32 compare."""
    v = _f32(val_scores)
    t = _f32(test_scores)
    counts = np.sum(v[:, None, None] > t[None], axis=0).astype(np.int32)
    return counts.astype(F32) < conformal_threshold(len(v), error)


def conformal_conds(val_scores, test_scores, error):
    """Same booleans via sort + searchsorted: count(val > t) = n - upper_bound(sorted, t)."""
    v = np.sort(_f32(val_scores))  # NaNs last
    t = _f32(test_scores)
    n = len(v)
    v = v[: n - int(np.isnan(v).sum())]  # a NaN validation score is above nothing; it still counts in the threshold
    counts = (len(v) - np.searchsorted(v, t, side="right")).astype(np.int32)
    return counts.astype(F32) < conformal_threshold(n, error)


def conformal_sets(val_scores, test_scores, error):
    """_get_conformal_sets_from_scores -> (mask [B,C] bool, sizes [B] int32, list of lists)."""
    conds = conformal_conds(val_scores, test_scores, error)
    sizes = conds.sum(1).astype(np.int32)
    sets = [np.nonzero(row)[0].tolist() for row in conds]
    return conds, sizes, sets


# ----------------------------------------------------------------------------- calibration metrics
def linspace_f32(start, stop, num):
    """jnp.linspace(start, stop, num) float32, endpoint=True (jax/_src/numpy/lax_numpy.py)."""
    start = F32(start)
    stop = F32(stop)
    div = num - 1
    step = (np.arange(div, dtype=F32) / F32(div)).astype(F32)
    out = start * (F32(1) - step) + stop * step
    return np.concatenate([out, np.array([stop], dtype=F32)]).astype(F32)


def ece_thresholds(n_data):
    """metric/classification.py:76: linspace(1 / len(probs), 1, 10)."""
    return linspace_f32(1 / n_data, 1, 10)


def bin_indices(conf, thresholds):
    """Bin i holds thr[i-1] < conf <= thr[i] (bin 0: conf <= thr[0]); -1 when conf > thr[-1] or NaN."""
    conf = _f32(conf)
    idx = np.full(conf.shape, -1, dtype=np.int32)
    idx[conf <= thresholds[0]] = 0
    for i in range(1, len(thresholds)):
        idx[(conf <= thresholds[i]) & (conf > thresholds[i - 1])] = i
    return idx


def compute_counts_confs_accs(preds, probs, targets):
    """metric/classification.py:74-87 -> (counts int32[10], confs f32[10], accs f32[10], bin_idx)."""
    probs = _f32(probs)
    conf = probs.max(-1) if probs.ndim == 2 else probs
    thr = ece_thresholds(len(conf))
    idx = bin_indices(conf, thr)
    diff = np.asarray(targets).astype(np.int64) - np.asarray(preds).astype(np.int64)
    counts = np.zeros(10, dtype=np.int32)
    confs = np.zeros(10, dtype=F32)
    accs = np.zeros(10, dtype=F32)
    for i in range(10):
        sel = idx == i
        counts[i] = sel.sum()
        if counts[i] > 0:
            accs[i] = F32((diff[sel] == 0).sum()) / F32(counts[i])
            confs[i] = np.sum(conf[sel], dtype=F32) / F32(counts[i])
        # empty bin: mean of empty = nan -> nan_to_num -> 0
    return counts, confs, accs, idx


def expected_calibration_error(preds, probs, targets):
    counts, confs, accs, _ = compute_counts_confs_accs(preds, probs, targets)
    return F32(np.sum(counts.astype(F32) * np.abs(accs - confs), dtype=F32) / F32(len(np.asarray(preds))))


def maximum_calibration_error(preds, probs, targets):
    counts, confs, accs, _ = compute_counts_confs_accs(preds, probs, targets)
    return F32(np.max(counts.astype(F32) * np.abs(accs - confs)))


def accuracy(preds, targets):
    preds = np.asarray(preds)
    targets = np.asarray(targets)
    return F32(F32((preds == targets).sum()) / F32(preds.shape[0]))


def brier_score(probs, targets):
    probs = _f32(probs)
    t = np.asarray(targets)
    if probs.ndim > 1:
        onehot = np.zeros_like(probs)
        onehot[np.arange(len(t)), t.astype(np.int64)] = 1
        d = probs - onehot
        return F32(np.mean(np.sum(d * d, axis=-1, dtype=F32), dtype=F32))
    d = probs - _f32(t)
    return F32(np.mean(d * d, dtype=F32))


# ----------------------------------------------------------------------------- quantile
def quantile_linear(a, q):
    """jnp.quantile(a, q) for 1-D ``a``, method="linear": sort; pos = q*(n-1); lo/hi = floor/ceil
    clamped to [0, n-1]; a[lo]*(1-w) + a[hi]*w with w = pos - floor(pos).  NaN in -> NaN out."""
    a = np.sort(_f32(a))
    n = a.shape[0]
    q = _f32(q)
    if np.isnan(a).any():
        return np.full(q.shape, np.nan, dtype=F32)
    pos = q * F32(n - 1)
    low = np.floor(pos)
    high = np.ceil(pos)
    hw = (pos - low).astype(F32)
    lw = (F32(1) - hw).astype(F32)
    lo = np.clip(low, 0, n - 1).astype(np.int64)
    hi = np.clip(high, 0, n - 1).astype(np.int64)
    return (a[lo] * lw + a[hi] * hw).astype(F32)


def conformal_quantile_level(n, error):
    """quantile.py:89-90: jnp.ceil((n + 1) * (1 - error)) / n in float32."""
    return F32(np.ceil(F32((n + 1) * (1 - error))) / F32(n))


def conformal_quantile(scores, error):
    scores = _f32(scores)
    return quantile_linear(scores, conformal_quantile_level(scores.shape[0], error))


# ----------------------------------------------------------------------------- conformal Bayes (importance sampling)
def conformal_bayes_counts(ell_train, lsm_test, dtype=np.float32):
    """predictive/classification.py:557-579 for every test point: ell_train [S, n_train] (ensemble log-probs of the
    training data), lsm_test [S, n_test, C] (log-softmax of the test logits, same S samples).
    Returns (counts int [n_test, C] = #{i : p_train[i] <= p_test}, p_train [n_test, C, n_train], p_test [n_test, C])."""
    ell = np.asarray(ell_train, dtype=dtype)
    lsm = np.asarray(lsm_test, dtype=dtype)
    S, n_test, Cn = lsm.shape
    e_train = np.exp(ell).astype(dtype)                          # jnp.exp(ensemble_train_log_probs)
    w = np.exp(np.transpose(lsm, (1, 2, 0))).astype(dtype)       # [n_test, C, S] importance_weights
    z = w.sum(axis=2, keepdims=True, dtype=dtype)                # normalizing_constant
    p_train = np.einsum("tcs,si->tci", (w / z).astype(dtype), e_train).astype(dtype)
    p_test = ((w**2).sum(axis=2, dtype=dtype) / z[..., 0]).astype(dtype)
    counts = (p_train <= p_test[..., None]).sum(axis=2)
    return counts, p_train, p_test


def conformal_bayes_region(ell_train, lsm_test, error, dtype=np.float32):
    """rank = count + 1 (the test point itself satisfies <=); region = float32(rank) > float32(error * (n_train + 1))
    (:576-582)."""
    counts, _, _ = conformal_bayes_counts(ell_train, lsm_test, dtype)
    n_train = np.asarray(ell_train).shape[1]
    return (counts + 1).astype(F32) > F32(error * (n_train + 1))


def conformal_bayes_ess(lsm_test, dtype=np.float32):
    """:601-617: 1 / sum_s softmax_s(log w)^2 per (test point, class)."""
    lw = np.transpose(np.asarray(lsm_test, dtype=dtype), (1, 2, 0))
    m = lw.max(axis=2, keepdims=True)
    lse = m + np.log(np.exp(lw - m).sum(axis=2, keepdims=True, dtype=dtype))
    iw = np.exp(lw - lse).astype(dtype)
    return (dtype(1) / (iw**2).sum(axis=2, dtype=dtype)).astype(dtype)
