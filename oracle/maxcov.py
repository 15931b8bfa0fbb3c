"""MaxCoverageFixedPrecisionBinaryClassificationCalibrator restated in numpy float32.  TEST INFRASTRUCTURE ONLY.

Follows fortuna/conformal/classification/maxcovfixprec_binary_classfication.py:33-124 literally: one objective
evaluation over ALL data points per candidate tau (the reference vmaps it), float32 op by op; ``jnp.linspace`` as jax
0.4 computes it (start * (1 - i/div) + stop * (i/div), endpoint appended; jax/_src/numpy/lax_numpy.py).
Pinned by the reference's own test case (tests/fortuna/test_conformal_methods.py:828-846: all-zero probabilities give
test values of zero) and by hand-worked cases in tests/test_oracle_maxcov.py; otherwise parity unpinned (jax is not
installed here)."""
import numpy as np

F32 = np.float32


def linspace_f32(start, stop, num):
    start, stop = F32(start), F32(stop)
    if num == 1:
        return np.array([start], F32)
    div = num - 1
    step = (np.arange(div, dtype=F32) / F32(div)).astype(F32)
    out = (start * (F32(1) - step) + stop * step).astype(F32)
    return np.concatenate([out, np.array([stop], F32)])


def tau_grids(tp_thr, tn_thr, n_taus=100):
    """(taus_pos ascending, taus_neg descending)  (:77-80)."""
    return linspace_f32(1, 2 * tp_thr, n_taus), linspace_f32(2 * (1 - tn_thr), 1, n_taus)[::-1].copy()


def _mean(x):
    return F32(np.asarray(x, dtype=F32).sum(dtype=np.float64)) / F32(len(x))  # exact count, then one float32 division


def objective_pos(tau, probs, targets, tp_thr, margin=0.0):
    p = np.asarray(probs, F32)
    calib = np.minimum((F32(1) + (F32(tau) - F32(1)) * (p > F32(0.5)).astype(F32)) * p, F32(1)).astype(F32)
    b = calib >= F32(tp_thr)
    prob_b = _mean(b)
    with np.errstate(invalid="ignore", divide="ignore"):
        prec = _mean(np.asarray(targets) * b) / prob_b
    return F32(prob_b * F32(prec >= F32(tp_thr + margin)))


def objective_neg(tau, probs, targets, tn_thr, margin=0.0):
    p = np.asarray(probs, F32)
    calib = ((F32(1) + (F32(tau) - F32(1)) * (p < F32(0.5)).astype(F32)) * p).astype(F32)
    b = calib <= F32(1 - tn_thr)
    prob_b = _mean(b)
    with np.errstate(invalid="ignore", divide="ignore"):
        prec = _mean((1 - np.asarray(targets)) * b) / prob_b
    return F32(prob_b * F32(prec >= F32(tn_thr + margin)))


def calibrate(targets, probs, tp_thr, tn_thr, n_taus=100, margin=0.0):
    """Returns (tau_pos, tau_neg, values_pos, values_neg)."""
    taus_pos, taus_neg = tau_grids(tp_thr, tn_thr, n_taus)
    vp = np.array([objective_pos(t, probs, targets, tp_thr, margin) for t in taus_pos], F32)
    vn = np.array([objective_neg(t, probs, targets, tn_thr, margin) for t in taus_neg], F32)
    return taus_pos[int(np.argmax(vp))], taus_neg[int(np.argmax(vn))], vp, vn


def apply_patches(probs, tau_pos, tau_neg):
    p = np.array(probs, dtype=F32, copy=True)
    hi = p > F32(0.5)
    p[hi] = np.minimum(F32(tau_pos) * p[hi], F32(1))
    lo = p < F32(0.5)
    p[lo] = F32(tau_neg) * p[lo]
    return p


def true_positive_precision(probs, targets, threshold):
    b = np.asarray(probs, F32) >= F32(threshold)
    with np.errstate(invalid="ignore", divide="ignore"):
        return _mean(np.asarray(targets) * b) / _mean(b)


def true_negative_precision(probs, targets, threshold):
    b = np.asarray(probs, F32) <= F32(1 - threshold)
    with np.errstate(invalid="ignore", divide="ignore"):
        return _mean((1 - np.asarray(targets)) * b) / _mean(b)
