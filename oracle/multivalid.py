"""ORACLE (test infrastructure only - never imported by the product): numpy restatement of the multivalid scorers, written
the way the reference writes them (explicit membership arrays per (tau, g, v, c), float32 means).  parity unpinned: the
reference's tests for these classes (tests/fortuna/test_conformal_methods.py:346-826) assert no numbers.

    fortuna/conformal/multivalid/base.py:11-84
    fortuna/conformal/multivalid/iterative/base.py:22-593          calibrate / apply_patches / calibration_error / _get_b
    fortuna/conformal/multivalid/iterative/multicalibrator.py      _calibration_error, _get_patch
    fortuna/conformal/multivalid/iterative/batch_mvp.py            _calibration_error, _get_patch (delta search)
    fortuna/conformal/multivalid/iterative/classification/top_label_multicalibrator.py   _get_b, _maybe_init_values
    fortuna/conformal/multivalid/mixins/{multicalibrator,batchmvp}.py                    losses
    fortuna/conformal/multivalid/one_shot/{base,multicalibrator}.py, one_shot/classification/top_label_multicalibrator.py
"""
import numpy as np

from . import prng
from .scoring import linspace_f32

F32 = np.float32


def permutation_from_key(key, n: int) -> np.ndarray:
    """jax.random.choice(key, n, shape=(n,), replace=False) -> jax/_src/random.py _shuffle."""
    key = np.asarray(key, dtype=np.uint32)
    rounds = int(np.ceil(3 * np.log(max(1, n)) / np.log(np.iinfo(np.uint32).max)))
    x = np.arange(n)
    for _ in range(rounds):
        key, sub = prng.split(key, 2)
        x = x[np.argsort(prng.random_bits(sub, n), kind="stable")]
    return x


def permutation(seed: int, n: int) -> np.ndarray:
    return permutation_from_key(prng.PRNGKey(seed), n)


def get_b(groups, values, v, g, c, tau, n_buckets, top_label=False):
    x = values if values.ndim == 1 else values[:, c]
    x = x.astype(F32)
    h = F32(0.5 / n_buckets)
    if tau == 0:
        b = np.abs(x - F32(v)) < h
    elif tau == 1:
        b = (x + h) >= F32(v)
    else:
        b = (x - h) <= F32(v)
    b = b & groups[:, g].astype(bool)
    if top_label:
        b = b & (values.argmax(1) == c)
    return b


def _mean(a):
    return np.mean(np.asarray(a, dtype=F32), dtype=F32)


class Iterative:
    """kind: "multicalibrator" | "top_label" | "batchmvp"."""

    def __init__(self, kind="multicalibrator", seed=0, n_classes=None):
        self.kind, self.seed, self.n_classes = kind, seed, n_classes
        self.patches, self.coverage = [], None

    # losses
    def loss(self, values, scores):
        with np.errstate(all="ignore"):
            if self.kind == "batchmvp":
                s = scores[:, 0] if scores.ndim == 2 and values.ndim == 1 else scores
                diff = (s - values).astype(F32)
                cov = F32(self.coverage)
                return float(_mean(diff * cov * (s > values) - diff * (F32(1) - cov) * (s <= values)))
            if scores.ndim == 2 and values.ndim == 1:
                return float(_mean(np.sum(((values - scores[:, 0]).astype(F32)) ** 2, axis=-1, dtype=F32)))
            return float(_mean(((values - scores).astype(F32)) ** 2))

    def cell(self, v, g, c, tau, scores_c, groups, values, n_buckets, b=None):
        """(error, b, patch, ey, mean_values): _calibration_error of the subclass."""
        if b is None:
            b = get_b(groups, values, v, g, c, tau, n_buckets, self.kind == "top_label")
        prob_b = _mean(b)
        with np.errstate(all="ignore"):
            if self.kind == "batchmvp":
                conds = (scores_c <= values) & b
                prob = np.where(prob_b > 0, _mean(conds) / prob_b, F32(0)).astype(F32)
                return prob_b * (F32(self.coverage) - prob) ** 2, b, None, None, None
            ey = _mean(scores_c * b) / prob_b
            x = values if values.ndim == 1 else values[:, c]
            mv = _mean(x * b) / _mean(b)
            patch = np.where(prob_b > 0, ey - mv, F32(0)).astype(F32)
            return prob_b * patch ** 2, b, patch, ey, mv

    def batchmvp_patch(self, b, scores, values, buckets):
        n_buckets = len(buckets)
        deltas = np.concatenate((-buckets[::-1][:-1], buckets)).astype(F32)
        prob_b = _mean(b)
        errs = []
        with np.errstate(all="ignore"):
            for d in deltas:
                conds = (scores <= (values + d).astype(F32)) & b
                prob = np.where(prob_b > 0, _mean(conds) / prob_b, F32(0)).astype(F32)
                errs.append((F32(self.coverage) - prob) ** 2)
        errs = np.array(errs, dtype=F32)
        idx = np.where(errs == errs.min())[0]
        if np.sum(idx >= n_buckets):
            return deltas[idx.min()]
        return deltas[idx.max()]

    @staticmethod
    def patch_values(values, patch, b, c, eta, patch_type):
        out = values.copy()
        sel = (b,) if values.ndim == 1 else (b, c)
        if patch_type == "additive":
            out[sel] = np.clip((values[sel] + F32(eta) * F32(patch)).astype(F32), 0, 1)
        else:
            out[sel] = np.clip((values[sel] * (F32(1) - F32(eta) * (F32(1) - F32(patch)))).astype(F32), 0, 1)
        return out

    def init_values(self, values, size):
        if values is not None:
            return np.asarray(values, dtype=F32).copy()
        if self.kind != "top_label":
            return np.zeros(size, dtype=F32)
        z = prng.normal(prng.PRNGKey(self.seed), size * self.n_classes).reshape(size, self.n_classes)
        v = (F32(1 / self.n_classes) * np.ones((size, self.n_classes), F32) + F32(0.01) * z).astype(F32)
        v = np.abs(v)
        v = (v - F32(2) * np.maximum(F32(0), v - F32(1))).astype(F32)
        return (v / v.sum(1, keepdims=True, dtype=F32)).astype(F32)

    def calibrate(self, scores, groups=None, values=None, atol=1e-4, rtol=1e-6, min_prob_b="auto", n_buckets=100,
                  n_rounds=1000, eta=0.1, split=0.8, bucket_types=("<=", ">="), patch_type="additive", coverage=0.95):
        self.coverage = coverage
        scores = np.asarray(scores, dtype=F32)
        if scores.ndim == 1:
            scores = scores[:, None]
        n_dims = scores.shape[1]
        if min_prob_b == "auto":
            min_prob_b = 1 / (n_buckets * n_dims)
        groups = np.ones((len(scores), 1), bool) if groups is None else np.asarray(groups, bool).copy()
        values = self.init_values(values, len(groups))
        self.n_buckets, self.eta, self.patch_type = n_buckets, eta, patch_type
        buckets = linspace_f32(0, 1, n_buckets)
        taus = [t for t, s in ((0, "="), (1, ">="), (2, "<=")) if s in bucket_types]
        size = len(scores)
        calib_size = int(np.ceil(F32(split * size)))
        perm = permutation(self.seed, size)
        tr, va = perm[:calib_size], perm[calib_size:]
        scores, val_scores = scores[tr], scores[va]
        values, val_values = values[tr], values[va]
        groups, val_groups = groups[tr], groups[va]
        top = self.kind == "top_label"
        val_losses = [self.loss(val_values, val_scores)]
        old = self.loss(values, scores)
        self.patches, converged = [], False
        for t in range(n_rounds):
            shape = (len(taus), groups.shape[1], n_buckets, n_dims)
            err, probs, bs = np.zeros(shape, F32), np.zeros(shape, F32), {}
            for it, tau in enumerate(taus):
                for g in range(groups.shape[1]):
                    for iv, v in enumerate(buckets):
                        for c in range(n_dims):
                            e, b, _, _, _ = self.cell(v, g, c, tau, scores[:, c], groups, values, n_buckets)
                            err[it, g, iv, c], probs[it, g, iv, c] = e, _mean(b)
                            bs[it, g, iv, c] = b
            cond_b = probs >= F32(min_prob_b)
            if not np.count_nonzero(cond_b):
                break
            err = np.where(np.isnan(err), err, err) * cond_b
            it, gt, iv, ct = np.unravel_index(int(np.argmax(err)), shape)
            taut, vt, bt = taus[it], buckets[iv], bs[it, gt, iv, ct]
            if self.kind == "batchmvp":
                patch = self.batchmvp_patch(bt, scores[:, ct], values, buckets)
            else:
                _, _, add, ey, mv = self.cell(vt, gt, ct, taut, scores[:, ct], groups, values, n_buckets, b=bt)
                with np.errstate(all="ignore"):
                    patch = add if patch_type == "additive" else F32(ey / mv)
            values = self.patch_values(values, patch, bt, ct, eta, patch_type)
            val_bt = get_b(val_groups, val_values, vt, gt, ct, taut, n_buckets, top)
            val_values = self.patch_values(val_values, patch, val_bt, ct, eta, patch_type)
            losses = self.loss(values, scores)
            val_losses.append(self.loss(val_values, val_scores))
            if losses > old:
                break
            if val_losses[-1] <= atol:
                converged = True
                break
            if val_losses[-1] > val_losses[-2]:
                break
            if np.any(val_bt) and (val_losses[-2] - val_losses[-1]) / val_losses[-2] <= rtol:
                converged = True
                break
            old = losses
            self.patches.append((int(taut), int(gt), float(vt), int(ct), float(patch)))
        self.final_values, self.final_val_values = values, val_values
        return dict(n_rounds=len(self.patches), losses=val_losses, converged=converged)

    def apply_patches(self, groups=None, values=None):
        values = self.init_values(values, len(groups) if groups is not None else None)
        groups = np.ones((len(values), 1), bool) if groups is None else np.asarray(groups, bool)
        for taut, gt, vt, ct, patch in self.patches:
            b = get_b(groups, values, F32(vt), gt, ct, taut, self.n_buckets, self.kind == "top_label")
            values = self.patch_values(values, patch, b, ct, self.eta, self.patch_type)
        return values

    def calibration_error(self, scores, groups=None, values=None, n_buckets=10):
        scores = np.asarray(scores, dtype=F32)
        if scores.ndim == 1:
            scores = scores[:, None]
        groups = np.ones((len(scores), 1), bool) if groups is None else np.asarray(groups, bool)
        values = self.init_values(values, len(scores))
        buckets = linspace_f32(0, 1, n_buckets)
        err = np.zeros((groups.shape[1], n_buckets, scores.shape[1]), F32)
        for g in range(groups.shape[1]):
            for iv, v in enumerate(buckets):
                for c in range(scores.shape[1]):
                    err[g, iv, c] = self.cell(v, g, c, 0, scores[:, c], groups, values, n_buckets)[0]
        out = err.sum(1, dtype=F32) / np.mean(groups.astype(F32), 0, dtype=F32)[:, None]
        return out[:, 0] if scores.shape[1] == 1 else out


def one_shot_patches(scores, values, n_buckets=100, min_prob_b="auto", top_label=False):
    """OneShotMultivalidMethod.calibrate -> patches [n_buckets, n_dims]."""
    scores = np.asarray(scores, dtype=F32)
    if scores.ndim == 1:
        scores = scores[:, None]
    values = np.asarray(values, dtype=F32)
    D = scores.shape[1]
    if min_prob_b == "auto":
        min_prob_b = 1 / (n_buckets * D)
    buckets = linspace_f32(0, 1, n_buckets)
    h = F32(0.5 / n_buckets)
    out = np.zeros((n_buckets, D), F32)
    for iv, v in enumerate(buckets):
        for c in range(D):
            x = values if values.ndim == 1 else values[:, c]
            b = np.abs(x - v) < h
            if top_label:
                b = b & (values.argmax(1) == c)
            prob_b = _mean(b)
            with np.errstate(all="ignore"):
                out[iv, c] = np.where(prob_b > F32(min_prob_b), _mean(scores[:, c] * b) / prob_b, v)
    return out


def one_shot_apply(patches, values, n_buckets, top_label=False):
    """OneShotMultivalidMethod.apply_patches, loop for loop (one_shot/base.py:68-126)."""
    values = np.asarray(values, dtype=F32)
    buckets = linspace_f32(0, 1, n_buckets)
    h = F32(0.5 / n_buckets)
    out = values.copy()
    D = 1 if values.ndim == 1 else values.shape[1]
    for u in np.unique(values):
        idx_v = int(np.argmin(np.abs(buckets - u)))
        for c in range(D):
            x = values if values.ndim == 1 else values[:, c]
            b = np.abs(x - u) < h
            if top_label:
                b = b & (values.argmax(1) == c)
            if values.ndim == 1:
                out[b] = patches[idx_v, c]
            else:
                out[b, c] = patches[idx_v, c]
    return out
