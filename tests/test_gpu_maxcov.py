"""MaxCoverageFixedPrecisionBinaryClassificationCalibrator on the GPU against the brute-force oracle (one objective
evaluation per grid point): exact counts, identical chosen taus and patched probabilities."""
import numpy as np
import pytest
import torch

from oracle import maxcov as M

pytestmark = pytest.mark.gpu


def _data(n, seed):
    r = np.random.default_rng(seed)
    y = r.integers(0, 2, n)
    # informative but imperfect scores; some exactly at the 1/2 boundary and at 0 / 1
    p = np.clip(0.5 + (y - 0.5) * r.uniform(0.0, 0.9, n) + 0.25 * r.standard_normal(n), 0, 1).astype(np.float32)
    p[:: max(1, n // 50)] = 0.5
    return p, y.astype(np.int32)


@pytest.mark.parametrize("n,n_taus", [(1, 100), (30, 100), (5000, 100), (200000, 257), (4097, 1)])
@pytest.mark.parametrize("thr", [(0.8, 0.7), (0.99, 0.95), (0.55, 0.6)])
def test_counts_match_bruteforce(cuda, n, n_taus, thr):
    from fortuna_b200 import ops

    p, y = _data(n, n + n_taus)
    taus_pos, taus_neg = M.tau_grids(thr[0], thr[1], n_taus)
    pd, yd = torch.from_numpy(p).to(cuda), torch.from_numpy(y).to(cuda)
    for side, taus, t in ((0, taus_pos, thr[0]), (1, taus_neg, 1 - thr[1])):
        c, l = ops.maxcov_counts(pd, yd, torch.from_numpy(taus).to(cuda), side, t)
        ind = (p > 0.5) if side == 0 else (p < 0.5)
        want_c, want_l = [], []
        for tau in taus:
            calib = ((np.float32(1) + (tau - np.float32(1)) * ind.astype(np.float32)) * p).astype(np.float32)
            b = (np.minimum(calib, np.float32(1)) >= np.float32(t)) if side == 0 else (calib <= np.float32(t))
            want_c.append(int(b.sum()))
            want_l.append(int((y * b).sum() if side == 0 else ((1 - y) * b).sum()))
        assert c.cpu().tolist() == want_c and l.cpu().tolist() == want_l


@pytest.mark.parametrize("n", [30, 4000, 100000])
@pytest.mark.parametrize("thr,margin", [((0.8, 0.7), 0.0), ((0.9, 0.9), 0.02), ((0.99, 0.99), 0.0)])
def test_calibrator_matches_oracle(cuda, n, thr, margin):
    from fortuna_b200.conformal import MaxCoverageFixedPrecisionBinaryClassificationCalibrator as Cal

    p, y = _data(n, n)
    pt, _ = _data(777, n + 1)
    cal = Cal()
    assert cal.apply_patches(pt) is pt  # no patches yet: returned unchanged (:99-101)
    out = cal.calibrate(targets=y, probs=p, true_positive_precision_threshold=thr[0],
                        true_negative_precision_threshold=thr[1], test_probs=pt, margin=margin)
    tp, tn, vp, vn = M.calibrate(y, p, thr[0], thr[1], margin=margin)
    assert cal.patches["tau_pos"] == tp and cal.patches["tau_neg"] == tn
    assert np.array_equal(cal._values[0], vp) and np.array_equal(cal._values[1], vn)
    assert np.array_equal(out.cpu().numpy(), M.apply_patches(pt, tp, tn))
    got = Cal.true_positive_precision(p, y, thr[0])
    want = M.true_positive_precision(p, y, thr[0])
    assert (np.isnan(got) and np.isnan(want)) or got == want
    got, want = Cal.true_negative_precision(p, y, thr[1]), M.true_negative_precision(p, y, thr[1])
    assert (np.isnan(got) and np.isnan(want)) or got == want
    assert Cal.true_positive_coverage(p, 0.8) == np.float32((p >= np.float32(0.8)).sum()) / np.float32(n)
    assert Cal.true_negative_coverage(p, 0.2) == np.float32((p <= np.float32(0.2)).sum()) / np.float32(n)


def test_reference_case_and_errors(cuda):
    from fortuna_b200.conformal import MaxCoverageFixedPrecisionBinaryClassificationCalibrator as Cal

    cal = Cal()
    y = np.array([0, 1] * 15)
    assert cal.calibrate(targets=y, probs=np.zeros(30, np.float32), true_positive_precision_threshold=0.99,
                         true_negative_precision_threshold=0.99) is None
    out = cal.calibrate(targets=y, probs=np.zeros(30, np.float32), test_probs=np.zeros(20, np.float32),
                        true_positive_precision_threshold=0.99, true_negative_precision_threshold=0.99)
    assert out.shape == (20,) and bool((out == 0).all())
    assert cal.patches["tau_pos"] == np.float32(1) and cal.patches["tau_neg"] == np.float32(1)
    for bad in ((0.5, 0.9), (0.9, 0.3)):
        with pytest.raises(ValueError):
            cal.calibrate(targets=y, probs=np.zeros(30, np.float32), true_positive_precision_threshold=bad[0],
                          true_negative_precision_threshold=bad[1])


def test_counts_with_nan_probabilities(cuda):
    """A NaN probability satisfies neither p > 0.5 nor the precision comparison (jnp.minimum keeps the NaN, NaN >= t is
    False): it is in no count on either side."""
    from fortuna_b200 import ops

    p, y = _data(5000, 11)
    p[::7] = np.nan
    taus_pos, taus_neg = M.tau_grids(0.8, 0.7, 33)
    pd, yd = torch.from_numpy(p).to(cuda), torch.from_numpy(y).to(cuda)
    for side, taus, t in ((0, taus_pos, 0.8), (1, taus_neg, 1 - 0.7)):
        c, l = ops.maxcov_counts(pd, yd, torch.from_numpy(taus).to(cuda), side, t)
        ind = (p > 0.5) if side == 0 else (p < 0.5)
        want_c, want_l = [], []
        for tau in taus:
            calib = ((np.float32(1) + (tau - np.float32(1)) * ind.astype(np.float32)) * p).astype(np.float32)
            with np.errstate(invalid="ignore"):
                b = (np.minimum(calib, np.float32(1)) >= np.float32(t)) if side == 0 else (calib <= np.float32(t))
            want_c.append(int(b.sum()))
            want_l.append(int((y * b).sum() if side == 0 else ((1 - y) * b).sum()))
        assert c.cpu().tolist() == want_c and l.cpu().tolist() == want_l
        assert max(want_c) <= int((~np.isnan(p)).sum())
