"""CPU checks of the MaxCovFixPrec restatement (oracle/maxcov.py): the reference's own test case
(tests/fortuna/test_conformal_methods.py:828-846), hand-worked cases, and the structural property the CUDA kernel relies
on - per data point the selection indicator switches on at ONE grid index."""
import numpy as np
import pytest

from oracle import maxcov as M


def test_linspace_matches_definition():
    g = M.linspace_f32(1, 1.98, 100)
    assert g.dtype == np.float32 and len(g) == 100 and g[0] == np.float32(1) and g[-1] == np.float32(1.98)
    assert np.all(np.diff(g) > 0)
    np.testing.assert_allclose(g, np.linspace(1, 1.98, 100), rtol=3e-7)
    assert np.array_equal(M.linspace_f32(0.25, 3, 1), np.array([0.25], np.float32))


def test_reference_test_case_all_zero_probs():
    y = np.array([0, 1] * 15)
    tp, tn, vp, vn = M.calibrate(y, np.zeros(30, np.float32), 0.99, 0.99)
    assert tp == np.float32(1) and np.all(vp == 0)  # nothing is ever predicted positive
    assert tn == np.float32(1) and np.all(vn == 0)  # everything is "negative" but only half truly is
    assert np.all(M.apply_patches(np.zeros(20, np.float32), tp, tn) == 0)


def test_hand_worked_case():
    # positives: p = 0.6 (y=1), 0.7 (y=1), 0.8 (y=0); T_tp = 0.9: tau*p >= 0.9 needs tau >= 1.5, 1.2857, 1.125.
    # precision must be >= 0.9: selecting {0.8} gives 0, {0.8, 0.7} gives 1/2, all three 2/3 -> never satisfied
    p = np.array([0.6, 0.7, 0.8, 0.1], np.float32)
    y = np.array([1, 1, 0, 0])
    tp, tn, vp, vn = M.calibrate(y, p, 0.9, 0.9, n_taus=50)
    assert np.all(vp == 0) and tp == np.float32(1)
    # negatives: only p = 0.1 < 0.5; tau*0.1 <= 0.1 for every tau in [0.2, 1] -> coverage 1/4 with precision 1
    assert np.all(vn == np.float32(0.25)) and tn == np.float32(1)  # arg-max takes the first grid point (tau = 1)
    # flip the label of p = 0.8: {0.8} alone has precision 1 from tau >= 1.125 on, {0.8, 0.7} as well from 1.2857
    y2 = np.array([1, 1, 1, 0])
    tp2, _, vp2, _ = M.calibrate(y2, p, 0.9, 0.9, n_taus=50)
    taus = M.tau_grids(0.9, 0.9, 50)[0]
    assert vp2.max() == np.float32(0.75) and tp2 == taus[np.argmax(taus * np.float32(0.6) >= np.float32(0.9))]
    out = M.apply_patches(p, tp2, 0.5)
    np.testing.assert_allclose(out, [min(tp2 * 0.6, 1), min(tp2 * 0.7, 1), 1.0, 0.05], rtol=1e-6)
    assert M.true_positive_precision(p, y, 0.7) == np.float32(0.5)
    assert M.true_negative_precision(p, y, 0.9) == np.float32(1.0)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_indicator_switches_once_along_the_grid(seed):
    r = np.random.default_rng(seed)
    p = np.concatenate([r.uniform(0, 1, 500), [0.5, 0.0, 1.0, 0.4999999, 0.5000001]]).astype(np.float32)
    tp_thr, tn_thr = 0.8, 0.7
    taus_pos, taus_neg = M.tau_grids(tp_thr, tn_thr, 100)
    bp = np.stack([np.minimum((np.float32(1) + (t - np.float32(1)) * (p > 0.5)) * p, 1) >= np.float32(tp_thr) for t in taus_pos])
    bn = np.stack([((np.float32(1) + (t - np.float32(1)) * (p < 0.5)) * p) <= np.float32(1 - tn_thr) for t in taus_neg])
    for b in (bp, bn):
        assert np.all(np.diff(b.astype(np.int8), axis=0) >= 0)  # false ... false true ... true


def test_host_side_values_from_counts_match_oracle():
    """The product's host logic (float32 ratios, margin test, arg-max, jnp.linspace restatement) fed with brute-force
    counts reproduces the oracle's objective values and chosen taus - the GPU kernel only has to deliver the counts."""
    from fortuna_b200.conformal.classification import maxcovfixprec_binary_classfication as H

    r = np.random.default_rng(7)
    p = r.uniform(0, 1, 3000).astype(np.float32)
    y = (r.uniform(0, 1, 3000) < p).astype(np.int32)
    for tp_thr, tn_thr, margin in ((0.8, 0.7, 0.0), (0.9, 0.9, 0.03)):
        taus_pos, taus_neg = M.tau_grids(tp_thr, tn_thr, 100)
        assert np.array_equal(H._linspace(1, 2 * tp_thr, 100), taus_pos)
        assert np.array_equal(H._linspace(2 * (1 - tn_thr), 1, 100)[::-1], taus_neg)
        cp, lp, cn, ln = [], [], [], []
        for t in taus_pos:
            b = np.minimum((np.float32(1) + (t - np.float32(1)) * (p > 0.5)) * p, np.float32(1)) >= np.float32(tp_thr)
            cp.append(b.sum()), lp.append((y * b).sum())
        for t in taus_neg:
            b = ((np.float32(1) + (t - np.float32(1)) * (p < 0.5)) * p).astype(np.float32) <= np.float32(1 - tn_thr)
            cn.append(b.sum()), ln.append(((1 - y) * b).sum())
        vp = H._values(np.array(cp), np.array(lp), len(p), tp_thr + margin)
        vn = H._values(np.array(cn), np.array(ln), len(p), tn_thr + margin)
        tp, tn, want_vp, want_vn = M.calibrate(y, p, tp_thr, tn_thr, margin=margin)
        assert np.array_equal(vp, want_vp) and np.array_equal(vn, want_vn)
        assert taus_pos[int(np.argmax(vp))] == tp and taus_neg[int(np.argmax(vn))] == tn
