"""Pins oracle/scoring.py: the one value-level conformal test of the reference
(tests/fortuna/test_conformal_methods.py:37-53) and cross-checks of every restated formula."""
import numpy as np

from oracle import predictive as P
from oracle import scoring as Sc


def test_reference_simple_conformal_case():
    vp = np.array([[0.1, 0.9], [0.8, 0.2], [0.3, 0.7]])
    vt = np.array([1, 1, 0])
    tp = np.array([[0.5, 0.5], [0.8, 0.2]])
    vs, ts = Sc.get_scores(vp, vt, tp, "simple")
    _, _, sets = Sc.conformal_sets(vs, ts, 0.05)
    assert 0 in sets[0] and 1 in sets[0] and 0 in sets[1]


def test_reference_shape_cases():
    vp = np.array([[0.1, 0.7, 0.2], [0.5, 0.4, 0.1], [0.15, 0.6, 0.35]])
    vt = np.array([1, 1, 2])
    tp = np.array([[0.2, 0.5, 0.3], [0.6, 0.01, 0.39]])
    for kind in ("simple", "aps"):
        vs, ts = Sc.get_scores(vp, vt, tp, kind)
        assert vs.shape == (3,) and ts.shape == (2, 3)
        assert len(Sc.conformal_sets(vs, ts, 0.05)[2]) == 2


def _tree_prefix(x):
    """associative_scan(add) on a python list of float32, straight from its definition."""
    n = len(x)
    if n < 2:
        return list(x)
    red = [np.float32(x[2 * i] + x[2 * i + 1]) for i in range(n // 2)]
    odd = _tree_prefix(red)
    out = [None] * n
    out[0] = x[0]
    for i in range(n // 2):
        out[2 * i + 1] = odd[i]
    for i in range(1, (n + 1) // 2):
        out[2 * i] = np.float32(odd[i - 1] + x[2 * i])
    return out


def test_cumsum_f32_is_associative_scan_and_padding_invariant():
    r = np.random.default_rng(3)
    for n in (1, 2, 3, 5, 8, 9, 31, 100, 1000, 1001):
        x = r.random((3, n)).astype(np.float32)
        got = Sc.cumsum_f32(x)
        for row in range(3):
            assert got[row].tolist() == [float(v) for v in _tree_prefix([np.float32(v) for v in x[row]])]
        n2 = 1
        while n2 < n:
            n2 *= 2
        xp = np.zeros((3, 2 * n2), np.float32)
        xp[:, :n] = x
        assert np.array_equal(Sc.cumsum_f32(xp)[:, :n], got)  # zero padding to a power of two changes no prefix
        np.testing.assert_allclose(got, np.cumsum(x.astype(np.float64), axis=1), rtol=2e-6)


def test_aps_definition_by_brute_force():
    r = np.random.default_rng(0)
    p = P.softmax(np.round(2 * r.standard_normal((40, 9))).astype(np.float32))  # many ties
    allc = Sc.aps_cumsums(p)
    for b in range(p.shape[0]):
        order = sorted(range(9), key=lambda c: (p[b, c], c), reverse=True)  # desc value, ties: larger index first
        assert order == Sc.descending_perm(p)[b].tolist()
        # the prefix sum at each rank, written out as the explicit pairwise tree of jax.lax.associative_scan
        srt = [np.float32(p[b, c]) for c in order]
        for c, want in zip(order, _tree_prefix(srt)):
            assert allc[b, c] == want
        run = np.float64(0)
        for c in order:  # and, independently of the association order, the float64 sequential sum to 1e-6
            run += np.float64(p[b, c])
            assert abs(float(allc[b, c]) - run) < 1e-6
    t = r.integers(0, 9, 40)
    assert np.array_equal(Sc.aps_score(p, t), allc[np.arange(40), t])
    top = p.argmax(-1)
    same = p[np.arange(40), top]
    # score of a unique arg-max class is its own probability
    uniq = (p == same[:, None]).sum(-1) == 1
    assert np.array_equal(Sc.aps_score(p, top)[uniq], same[uniq])


def test_literal_vs_searchsorted_sets_and_extremes():
    r = np.random.default_rng(1)
    vs = np.round(r.random(200), 2).astype(np.float32)
    ts = np.round(r.random((30, 7)), 2).astype(np.float32)
    for err in (0.0, 0.01, 0.05, 0.3, 0.9, 1.0):
        assert np.array_equal(Sc.conformal_conds_literal(vs, ts, err), Sc.conformal_conds(vs, ts, err))
    assert Sc.conformal_conds(vs, ts, 0.0).all()  # error -> 0: full sets
    assert not Sc.conformal_conds(vs, np.full((2, 2), -1.0, np.float32), 1.0).any()


def test_linspace_and_bins():
    thr = Sc.ece_thresholds(100)
    assert thr[0] == np.float32(0.01) and thr[-1] == np.float32(1.0) and len(thr) == 10
    assert np.all(np.diff(thr) > 0)
    conf = np.array([0.0, 0.01, 0.010001, 0.5, 1.0, 1.5, np.nan], np.float32)
    assert Sc.bin_indices(conf, thr).tolist() == [0, 0, 1, 5, 9, -1, -1]


def test_ece_against_direct_formula():
    r = np.random.default_rng(2)
    p = P.softmax((3 * r.standard_normal((500, 4))).astype(np.float32))
    t = r.integers(0, 4, 500)
    preds = p.argmax(-1)
    counts, confs, accs, idx = Sc.compute_counts_confs_accs(preds, p, t)
    assert counts.sum() == 500
    conf = p.max(-1)
    ece = sum(abs((preds[idx == i] == t[idx == i]).mean() - conf[idx == i].mean()) * (idx == i).sum() for i in range(10) if (idx == i).any()) / 500
    np.testing.assert_allclose(Sc.expected_calibration_error(preds, p, t), ece, rtol=1e-5)
    assert Sc.maximum_calibration_error(preds, p, t) >= Sc.expected_calibration_error(preds, p, t)
    np.testing.assert_allclose(Sc.accuracy(preds, t), (preds == t).mean(), rtol=1e-6)
    onehot = np.eye(4)[t]
    np.testing.assert_allclose(Sc.brier_score(p, t), ((p - onehot) ** 2).sum(-1).mean(), rtol=1e-5)


def test_quantile_matches_numpy_linear_and_clamps():
    r = np.random.default_rng(3)
    x = r.standard_normal(101).astype(np.float32)
    for q in (0.0, 0.25, 0.5, 0.9, 1.0):
        np.testing.assert_allclose(Sc.quantile_linear(x, q), np.quantile(x.astype(np.float64), q), rtol=1e-5, atol=1e-6)
    assert Sc.quantile_linear(x, np.float32(1.2)) == np.float32(x.max())  # q > 1 clamps to the maximum
    lvl = Sc.conformal_quantile_level(100, 0.05)
    assert lvl == np.float32(np.ceil(101 * 0.95) / 100)
    s = r.standard_normal(100).astype(np.float32)
    assert s.min() <= Sc.conformal_quantile(s, 0.05) <= s.max()
    assert np.isnan(Sc.quantile_linear(np.array([1.0, np.nan], np.float32), 0.5))
