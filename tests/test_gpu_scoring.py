"""K5-K9 parity: integer / index outputs bit-exact, float32 prefix sums bit-exact (same sequential order)."""
import numpy as np
import pytest
import torch

from oracle import predictive as P
from oracle import scoring as Sc

pytestmark = pytest.mark.gpu


def _probs(B, C, seed, ties=False):
    z = (4 * np.random.default_rng(seed).standard_normal((B, C))).astype(np.float32)
    if ties:
        z = np.round(z)
    return P.softmax(z)


@pytest.mark.parametrize("B,C", [(64, 1000), (200, 10), (50, 2), (33, 1), (17, 37), (9, 1024), (5, 3000)])
@pytest.mark.parametrize("ties", [False, True])
def test_aps_scores_bit_exact(cuda, B, C, ties):
    from fortuna_b200 import ops

    p = _probs(B, C, seed=B + C, ties=ties)
    t = np.random.default_rng(0).integers(0, C, B).astype(np.int32)
    pt = torch.from_numpy(p).to(cuda)
    allc, perm = ops.aps_scores(pt, None, want_perm=True)
    assert np.array_equal(perm.cpu().numpy(), Sc.descending_perm(p))
    assert np.array_equal(allc.cpu().numpy(), Sc.aps_cumsums(p))
    # without the permutation output the key-only kernel runs (32-bit sort + rank lookup): same bits
    assert np.array_equal(ops.aps_scores(pt, None).cpu().numpy(), Sc.aps_cumsums(p))
    tgt = ops.aps_scores(pt, torch.from_numpy(t).to(cuda))
    assert np.array_equal(tgt.cpu().numpy(), Sc.aps_score(p, t))
    s = ops.simple_scores(pt, torch.from_numpy(t).to(cuda))
    assert np.array_equal(s.cpu().numpy(), Sc.simple_score(p, t))
    assert np.array_equal(ops.simple_scores(pt).cpu().numpy(), Sc.simple_scores_all(p))


@pytest.mark.parametrize("C", [1, 2, 31, 32, 33, 100, 257, 512, 513, 1000, 1024])
def test_aps_scores_degenerate_rows(cuda, C):
    """Rows where the rank of a class is decided by the tie rule (larger class first): all-equal rows, rows that
    underflowed to exact zeros, runs of duplicates crossing lane boundaries, a single dominant class."""
    from fortuna_b200 import ops

    r = np.random.default_rng(C)
    rows = [np.full(C, 1.0 / C, dtype=np.float32)]
    z = np.zeros(C, dtype=np.float32)
    z[r.integers(0, C)] = 1.0
    rows.append(z)
    q = P.softmax((4 * r.standard_normal((1, C))).astype(np.float32))[0]
    q[r.random(C) < 0.6] = 0.0
    rows.append(q)
    d = P.softmax((4 * r.standard_normal((1, C))).astype(np.float32))[0]
    d = d[r.integers(0, max(C // 7, 1), C)]  # few distinct values, each repeated many times
    rows.append(d)
    e = P.softmax((4 * r.standard_normal((1, C))).astype(np.float32))[0]
    e[::2] = e[0]
    rows.append(e)
    rows.append(np.where(np.arange(C) % 3 == 0, np.float32(-0.0), np.float32(0.0)).astype(np.float32))
    rows.append(P.softmax((4 * r.standard_normal((1, C))).astype(np.float32))[0])
    p = np.stack(rows).astype(np.float32)
    p = np.concatenate([p] * 3)  # more rows than one warp sees: tie rows and clean rows alternate on a warp
    t = r.integers(0, C, p.shape[0]).astype(np.int32)
    pt = torch.from_numpy(p).to(cuda)
    want = Sc.aps_cumsums(p)
    assert np.array_equal(ops.aps_scores(pt, None).cpu().numpy(), want)
    allc, perm = ops.aps_scores(pt, None, want_perm=True)
    assert np.array_equal(allc.cpu().numpy(), want)
    assert np.array_equal(perm.cpu().numpy(), Sc.descending_perm(p))
    assert np.array_equal(ops.aps_scores(pt, torch.from_numpy(t).to(cuda)).cpu().numpy(), Sc.aps_score(p, t))


@pytest.mark.parametrize("n", [1, 2, 100, 4096, 4097, 50_000, 300_001])
def test_sort_bit_exact(cuda, n):
    from fortuna_b200 import ops

    x = np.random.default_rng(n).standard_normal(n).astype(np.float32)
    if n > 10:
        x[3] = x[7]  # duplicates
        x[5] = np.inf
    got = ops.sort_f32_(torch.from_numpy(x.copy()).to(cuda)).cpu().numpy()
    assert np.array_equal(got, np.sort(x))


@pytest.mark.parametrize("error", [0.05, 0.1, 0.5, 0.0, 1.0])
@pytest.mark.parametrize("B,C,n_val", [(64, 1000, 5000), (100, 10, 37), (7, 3, 3)])
def test_conformal_sets_bit_exact(cuda, B, C, n_val, error):
    from fortuna_b200 import ops

    vp = _probs(n_val, C, seed=2, ties=True)
    vt = np.random.default_rng(3).integers(0, C, n_val).astype(np.int32)
    tp = _probs(B, C, seed=4, ties=True)
    val_scores, test_scores = Sc.get_scores(vp, vt, tp, "aps")
    conds, sizes, _ = Sc.conformal_sets(val_scores, test_scores, error)
    if n_val * B * C <= 4_000_000:
        assert np.array_equal(conds, Sc.conformal_conds_literal(val_scores, test_scores, error))
    vs = ops.sort_f32_(torch.from_numpy(val_scores.copy()).to(cuda))
    mask, sz = ops.conformal_sets(vs, torch.from_numpy(test_scores).to(cuda), Sc.conformal_threshold(n_val, error))
    assert np.array_equal(mask.cpu().numpy().astype(bool), conds)
    assert np.array_equal(sz.cpu().numpy(), sizes)


def test_reference_conformal_known_case(cuda):
    """tests/fortuna/test_conformal_methods.py:37-53 (simple score): {0,1} in set 0, 0 in set 1."""
    from fortuna_b200 import ops

    vp = np.array([[0.1, 0.9], [0.8, 0.2], [0.3, 0.7]], dtype=np.float32)
    vt = np.array([1, 1, 0], dtype=np.int32)
    tp = np.array([[0.5, 0.5], [0.8, 0.2]], dtype=np.float32)
    vs = ops.simple_scores(torch.from_numpy(vp).to(cuda), torch.from_numpy(vt).to(cuda))
    ts = ops.simple_scores(torch.from_numpy(tp).to(cuda))
    mask, _ = ops.conformal_sets(ops.sort_f32_(vs), ts, Sc.conformal_threshold(3, 0.05))
    m = mask.cpu().numpy().astype(bool)
    assert m[0, 0] and m[0, 1] and m[1, 0]


@pytest.mark.parametrize("B,C", [(1000, 10), (4096, 1000), (77, 2), (513, 1)])
def test_ece_bins(cuda, B, C):
    from fortuna_b200 import ops

    r = np.random.default_rng(B)
    t = r.integers(0, max(C, 2), B).astype(np.int32)
    if C == 1:
        p = r.random(B).astype(np.float32)
        preds = r.integers(0, 2, B).astype(np.int32)
        probs_t = torch.from_numpy(p).to(cuda)
        preds_t = torch.from_numpy(preds).to(cuda)
    else:
        p = _probs(B, C, seed=5, ties=(C == 2))
        preds = p.argmax(-1).astype(np.int32)
        probs_t = torch.from_numpy(p).to(cuda)
        preds_t = None
    thr = Sc.ece_thresholds(B)
    o = ops.ece_bins(probs_t, torch.from_numpy(t).to(cuda), torch.from_numpy(thr).to(cuda), preds=preds_t, want_bin_idx=True)
    counts, confs, accs, idx = Sc.compute_counts_confs_accs(preds, p, t)
    assert np.array_equal(o["bin_idx"].cpu().numpy(), idx)
    assert np.array_equal(o["counts"].cpu().numpy(), counts)
    got_counts = o["counts"].cpu().numpy().astype(np.float32)
    got_confs = np.where(counts > 0, o["conf_sums"].cpu().numpy() / np.maximum(got_counts, 1), 0)
    got_accs = np.where(counts > 0, o["correct"].cpu().numpy() / np.maximum(got_counts, 1), 0)
    np.testing.assert_allclose(got_confs, confs, rtol=1e-5)
    np.testing.assert_allclose(got_accs, accs, rtol=1e-6)
    res = o["result"].cpu().numpy()
    np.testing.assert_allclose(res[0], Sc.expected_calibration_error(preds, p, t), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(res[1], Sc.maximum_calibration_error(preds, p, t), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(res[2], Sc.accuracy(preds, t), rtol=1e-6)
    if C > 1:
        np.testing.assert_allclose(res[3], Sc.brier_score(p, t), rtol=1e-5)


@pytest.mark.parametrize("n", [1, 2, 100, 5001])
@pytest.mark.parametrize("error", [0.05, 0.1, 0.5, 0.001])
def test_quantile(cuda, n, error):
    from fortuna_b200 import ops

    x = np.random.default_rng(n).standard_normal(n).astype(np.float32)
    q = np.array([Sc.conformal_quantile_level(n, error), 0.0, 0.5, 1.0, 0.3333], dtype=np.float32)
    xs = ops.sort_f32_(torch.from_numpy(x.copy()).to(cuda))
    got = ops.quantile_sorted(xs, torch.from_numpy(q).to(cuda)).cpu().numpy()
    assert np.array_equal(got, Sc.quantile_linear(x, q))


@pytest.mark.parametrize("error", [0.05, 0.3])
def test_credible_sets(cuda, error):
    from fortuna_b200 import ops

    p = _probs(200, 50, seed=8, ties=True)
    labels, size = ops.credible_sets(torch.from_numpy(p).to(cuda), error)
    want = P.cls_credible_set(p, error)
    lab = labels.cpu().numpy()
    sz = size.cpu().numpy()
    for i in range(p.shape[0]):
        assert lab[i, : sz[i]].tolist() == want[i]


def _cb_inputs(S, n_train, n_test, C, seed):
    r = np.random.default_rng(seed)
    zt = (2.0 * r.standard_normal((S, n_train, C))).astype(np.float32)
    yt = r.integers(0, C, n_train)
    ell = np.take_along_axis(P.log_softmax_rows(zt), np.broadcast_to(yt[None, :, None], (S, n_train, 1)), axis=2)[..., 0]
    lsm = P.log_softmax_rows((2.0 * r.standard_normal((S, n_test, C))).astype(np.float32))
    return np.ascontiguousarray(ell, dtype=np.float32), np.ascontiguousarray(lsm, dtype=np.float32)


@pytest.mark.parametrize("S,n_train,n_test,C", [(30, 500, 40, 2), (7, 63, 5, 3), (100, 3000, 130, 10), (33, 1000, 7, 101), (256, 257, 65, 4)])
@pytest.mark.parametrize("error", [0.05, 0.3])
def test_conformal_bayes_sets(cuda, S, n_train, n_test, C, error):
    """fb200_conformal_bayes_sets against the float64 evaluation of predictive/classification.py:557-582.  The rank is an
    integer count of float comparisons p_train[i] <= p_test: entries within a few float32 ulps of p_test may fall
    either way, so each row's count may differ by at most its number of such entries, and the region must agree
    wherever the rank is further than that from the threshold."""
    from fortuna_b200 import ops

    ell, lsm = _cb_inputs(S, n_train, n_test, C, seed=S + C)
    out = ops.conformal_bayes_sets(torch.from_numpy(ell).to(cuda), torch.from_numpy(lsm).to(cuda), error,
                                   want_counts=True, want_ess=True)
    cnt64, p_train, p_test = Sc.conformal_bayes_counts(ell, lsm, dtype=np.float64)
    amb = (np.abs(p_train - p_test[..., None]) <= 4e-6 * p_test[..., None]).sum(axis=2)
    got = out["rank_counts"].cpu().numpy()
    assert np.all(np.abs(got - cnt64) <= amb)
    assert (amb == 0).mean() > 0.8 and np.array_equal(got[amb == 0], cnt64[amb == 0])
    thr = np.float32(error * (n_train + 1))
    clear = np.abs((cnt64 + 1).astype(np.float64) - float(thr)) > amb + 0.5
    want_region = Sc.conformal_bayes_region(ell, lsm, error, dtype=np.float64)
    region = out["region"].cpu().numpy().astype(bool)
    assert np.array_equal(region[clear], want_region[clear])
    assert np.array_equal(region, (got + 1).astype(np.float32) > thr)
    assert np.array_equal(out["sizes"].cpu().numpy(), region.sum(axis=1))
    np.testing.assert_allclose(out["ess"].cpu().numpy(), Sc.conformal_bayes_ess(lsm, dtype=np.float64), rtol=2e-5)
    # the float32 restatement agrees with the float64 one to the same standard
    cnt32, _, _ = Sc.conformal_bayes_counts(ell, lsm)
    assert np.all(np.abs(cnt32 - cnt64) <= amb)


def test_softmax_rows_in_place(cuda):
    from fortuna_b200 import ops

    z = (3 * np.random.default_rng(0).standard_normal((5, 37, 1000))).astype(np.float32)
    zt = torch.from_numpy(z.copy()).to(cuda)
    _, lse = ops.softmax_rows_(zt, ops.EPI_LOGITS, want_lse=True)
    assert torch.equal(zt.cpu(), torch.from_numpy(z))
    m = z.max(-1, keepdims=True)
    want_lse = (m + np.log(np.exp(z - m).sum(-1, keepdims=True)))[..., 0]
    np.testing.assert_allclose(lse.cpu().numpy(), want_lse, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(ops.softmax_rows_(zt.clone(), ops.EPI_LOG_SOFTMAX).cpu().numpy(), P.log_softmax_rows(z), rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(ops.softmax_rows_(zt.clone(), ops.EPI_SOFTMAX).cpu().numpy(), P.softmax(z), rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("B,C,n_val", [(64, 1000, 5000), (100, 10, 37), (7, 3, 3), (33, 1024, 200), (50, 2, 11)])
@pytest.mark.parametrize("error", [0.05, 0.3, 0.0, 1.0])
def test_fused_aps_conformal_sets_equal_the_two_kernels(cuda, B, C, n_val, error):
    """fb200_aps_conformal_sets (sets decided in the sort kernel's epilogue) == fb200_aps_scores + fb200_conformal_sets,
    and == the oracle."""
    from fortuna_b200 import ops

    vp = _probs(n_val, C, seed=2, ties=True)
    vt = np.random.default_rng(3).integers(0, C, n_val).astype(np.int32)
    tp = _probs(B, C, seed=4, ties=(B % 2 == 0))
    val_scores, test_scores = Sc.get_scores(vp, vt, tp, "aps")
    conds, sizes, _ = Sc.conformal_sets(val_scores, test_scores, error)
    vs = ops.sort_f32_(torch.from_numpy(val_scores.copy()).to(cuda))
    thr = Sc.conformal_threshold(n_val, error)
    tpt = torch.from_numpy(tp).to(cuda)
    mask, sz, scores = ops.aps_conformal_sets(tpt, vs, thr, want_scores=True)
    assert np.array_equal(scores.cpu().numpy(), test_scores)
    assert np.array_equal(mask.cpu().numpy().astype(bool), conds) and np.array_equal(sz.cpu().numpy(), sizes)
    m2, s2 = ops.aps_conformal_sets(tpt, vs, thr)
    assert torch.equal(m2, mask) and torch.equal(s2, sz)
    m3, s3 = ops.conformal_sets(vs, ops.aps_scores(tpt, None), thr)
    assert torch.equal(m3, mask) and torch.equal(s3, sz)


@pytest.mark.parametrize("S,B,C", [(8, 4096, 1000), (16, 1000, 10), (5, 77, 2), (3, 40, 1500)])
def test_ece_from_reduction_rows_equals_ece_on_the_mean(cuda, S, B, C):
    """confidence / mode / brier_terms out of the reduction kernel's epilogue + fb200_ece_bins_from_rows == fb200_ece_bins
    on the [B,C] mean: identical bin indices and counts, ECE / MCE / accuracy / Brier within float rounding."""
    from fortuna_b200 import ops

    r = np.random.default_rng(S + C)
    z = torch.from_numpy((3 * r.standard_normal((S, B, C))).astype(np.float32)).to(cuda)
    t = torch.from_numpy(r.integers(0, C, B).astype(np.int32)).to(cuda)
    o = ops.bma_reduce_cls(z, ("mean", "mode", "confidence", "brier_terms"), targets=t)
    mean = o["mean"]
    assert torch.equal(o["confidence"], mean.max(-1).values)
    oh = torch.zeros_like(mean)
    oh[torch.arange(B, device=cuda), t.long()] = 1
    np.testing.assert_allclose(o["brier_terms"].cpu().numpy(), ((mean - oh) ** 2).sum(-1).cpu().numpy(), rtol=2e-5, atol=1e-7)
    thr = torch.from_numpy(Sc.ece_thresholds(B)).to(cuda)
    a = ops.ece_bins(mean, t, thr, want_bin_idx=True)
    b = ops.ece_bins_from_rows(o["confidence"], o["mode"], t, thr, brier_terms=o["brier_terms"], want_bin_idx=True)
    assert torch.equal(a["bin_idx"], b["bin_idx"]) and torch.equal(a["counts"], b["counts"]) and torch.equal(a["correct"], b["correct"])
    assert torch.equal(a["conf_sums"], b["conf_sums"])  # fixed-point sums of the same confidences: identical
    np.testing.assert_allclose(b["result"].cpu().numpy(), a["result"].cpu().numpy(), rtol=1e-5, atol=1e-7)


def _select_rows(C, seed):
    """Rows that exercise every exit of the select kernel: peaked / very peaked / flat softmax rows, tie-heavy rows,
    exact zeros, uniform, one-hot, -0.0 and a slightly negative value."""
    r = np.random.default_rng(seed)
    blocks = [
        P.softmax((4 * r.standard_normal((1500, C))).astype(np.float32)),
        P.softmax((8 * r.standard_normal((300, C))).astype(np.float32)),
        P.softmax((1 * r.standard_normal((300, C))).astype(np.float32)),   # boundary deeper than the selection holds
        P.softmax((2 * r.standard_normal((300, C))).astype(np.float32)),
        P.softmax(np.round(4 * r.standard_normal((300, C))).astype(np.float32)),  # ties, some across the boundary
        P.softmax(np.round(2 * r.standard_normal((200, C))).astype(np.float32)),
    ]
    z = P.softmax((4 * r.standard_normal((60, C))).astype(np.float32))
    z[r.random(z.shape) < 0.6] = 0.0
    blocks.append(z)
    blocks.append(np.full((3, C), 1.0 / C, dtype=np.float32))
    oh = np.zeros((3, C), dtype=np.float32)
    oh[np.arange(3), r.integers(0, C, 3)] = 1.0
    blocks.append(oh)
    nz = P.softmax((4 * r.standard_normal((4, C))).astype(np.float32))
    nz[:2, ::5] = np.float32(-0.0)
    nz[2:, 7] = np.float32(-1e-9)
    blocks.append(nz)
    p = np.concatenate(blocks).astype(np.float32)
    return p[r.permutation(len(p))]


@pytest.mark.parametrize("C", [1000, 1024, 512, 260])
@pytest.mark.parametrize("error", [0.02, 0.05, 0.3, 0.7])
def test_aps_sets_select_kernel(cuda, C, error):
    """Sets without scores on wide rows run the select kernel (threshold selection + a sort of the selected values only,
    undecided rows re-done by the full sort): same mask and sizes as the oracle and as the path that sorts every row."""
    from fortuna_b200 import ops

    n_val = 5000
    vp = _probs(n_val, C, seed=12)
    vt = np.random.default_rng(13).integers(0, C, n_val).astype(np.int32)
    tp = _select_rows(C, seed=C + int(error * 100))
    val_scores, test_scores = Sc.get_scores(vp, vt, tp, "aps")
    conds, sizes, _ = Sc.conformal_sets(val_scores, test_scores, error)
    vs = ops.sort_f32_(torch.from_numpy(val_scores.copy()).to(cuda))
    thr = Sc.conformal_threshold(n_val, error)
    tpt = torch.from_numpy(tp).to(cuda)
    mask, sz = ops.aps_conformal_sets(tpt, vs, thr)
    assert np.array_equal(sz.cpu().numpy(), sizes)
    assert np.array_equal(mask.cpu().numpy().astype(bool), conds)
    m1, s1, _ = ops.aps_conformal_sets(tpt, vs, thr, want_scores=True)
    assert torch.equal(m1, mask) and torch.equal(s1, sz)


@pytest.mark.parametrize("C", [1000, 512])
def test_aps_sets_select_kernel_quantile_equal_to_a_prefix_sum(cuda, C):
    """The validation scores ARE prefix sums of the test rows, so the quantile equals a test score exactly: the
    boundary rank is decided by equality, and the select kernel's margin test has to send near-misses to the full sort."""
    from fortuna_b200 import ops

    r = np.random.default_rng(C)
    tp = P.softmax((4 * r.standard_normal((2000, C))).astype(np.float32))
    n_val = 400
    vp = tp[:n_val]
    cums = Sc.aps_cumsums(vp)
    order = Sc.descending_perm(vp)
    vt = order[np.arange(n_val), r.integers(5, 80, n_val)].astype(np.int32)  # class at a random shallow rank
    for error in (0.1, 0.5, 0.9):
        val_scores, test_scores = Sc.get_scores(vp, vt, tp, "aps")
        assert np.array_equal(val_scores, cums[np.arange(n_val), vt])
        conds, sizes, _ = Sc.conformal_sets(val_scores, test_scores, error)
        vs = ops.sort_f32_(torch.from_numpy(val_scores.copy()).to(cuda))
        mask, sz = ops.aps_conformal_sets(torch.from_numpy(tp).to(cuda), vs, Sc.conformal_threshold(n_val, error))
        assert np.array_equal(sz.cpu().numpy(), sizes)
        assert np.array_equal(mask.cpu().numpy().astype(bool), conds)


def test_ece_bins_rows_with_nan_fall_into_no_bin(cuda):
    """metric/classification.py:75-83: probs.max(-1) of a row holding a NaN is NaN, and NaN <= threshold is False for
    every bin - the row is not counted (the oracle's bin index -1).  Also rows sitting exactly on a threshold."""
    from fortuna_b200 import ops
    from oracle import predictive as P
    from oracle import scoring as Sc

    r = np.random.default_rng(1)
    B, C = 500, 10
    p = P.softmax((3 * r.standard_normal((B, C))).astype(np.float32))
    t = r.integers(0, C, B).astype(np.int32)
    thr = Sc.ece_thresholds(B)
    p[3] = np.nan
    p[5, 4] = np.nan
    p[7] = 0.1
    p[9] = 0.0
    p[9, 2] = 1.0
    p[11] = 0.0
    p[11, 0], p[11, 1] = thr[4], 1 - thr[4]  # a confidence exactly on a threshold
    preds = np.where(np.isnan(p).any(1), 0, np.nan_to_num(p).argmax(1)).astype(np.int32)
    with np.errstate(all="ignore"):
        counts, confs, accs, idx = Sc.compute_counts_confs_accs(preds, p, t)
    e = ops.ece_bins(torch.from_numpy(p).to(cuda), torch.from_numpy(t).to(cuda), torch.from_numpy(thr).to(cuda),
                     preds=torch.from_numpy(preds).to(cuda), want_bin_idx=True)
    assert np.array_equal(e["counts"].cpu().numpy(), counts) and int(counts.sum()) == B - 2
    assert np.array_equal(e["bin_idx"].cpu().numpy(), idx) and idx[3] == -1 and idx[5] == -1
    np.testing.assert_allclose(e["confs"].cpu().numpy(), confs, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(e["accs"].cpu().numpy(), accs, rtol=1e-6)


def test_sort_and_quantile_with_nan_scores(cuda):
    """lax.sort puts NaNs (of either sign) last and treats -0 as 0; jnp.quantile of scores holding a NaN is NaN
    (conformal/regression/quantile.py:89-90 on scores computed from NaN predictions)."""
    from fortuna_b200 import ops
    from fortuna_b200.conformal.regression.quantile import conformal_quantile
    from oracle import scoring as Sc

    s = np.array([0.5, np.nan, 0.25, -1.0, np.inf, -np.inf, 0.0, -0.0, -np.nan], np.float32)
    x = ops.sort_f32_(torch.from_numpy(s.copy()).to(cuda)).cpu().numpy()
    assert np.array_equal(x[:7], np.array([-np.inf, -1.0, 0.0, 0.0, 0.25, 0.5, np.inf], np.float32))
    assert np.isnan(x[7:]).all() and not np.signbit(x[2:4]).any()
    r = np.random.default_rng(2)
    for n in (3, 101, 5000):
        sc = r.random(n).astype(np.float32)
        want_clean = Sc.conformal_quantile(sc, 0.1)
        assert float(conformal_quantile(sc, 0.1)) == float(want_clean)
        sc[n // 2] = np.nan
        assert np.isnan(float(conformal_quantile(sc, 0.1)))
        got = ops.sort_f32_(torch.from_numpy(sc.copy()).to(cuda)).cpu().numpy()
        assert np.array_equal(got[:-1], np.sort(sc)[:-1]) and np.isnan(got[-1])


@pytest.mark.parametrize("C", [10, 33, 130, 1000, 1024, 3000])
def test_rows_with_some_nan_classes_sort_them_first(cuda, C):
    """lax.sort puts NaNs last, so the descending order of adaptive_prediction.py:13,20 and
    classification.py:471-483 starts with the NaN classes (larger class first) and every prefix sum after them is
    NaN: the credible set is the single highest NaN class, all APS scores of the row are NaN.  Positive and
    negative NaNs alike."""
    from fortuna_b200 import ops

    r = np.random.default_rng(C)
    p = P.softmax((3 * r.standard_normal((12, C))).astype(np.float32))
    neg_nan = np.array([0xFFC00000], dtype=np.uint32).view(np.float32)[0]
    p[4, 3] = np.nan
    p[5, 2] = np.nan
    p[5, 7] = neg_nan
    p[6, C - 1] = neg_nan
    p[7, :] = np.nan
    p[8, 0] = np.nan
    p[8, C // 2] = np.inf
    t = r.integers(0, C, p.shape[0]).astype(np.int32)
    pt = torch.from_numpy(p).to(cuda)
    with np.errstate(all="ignore"):
        want_sets = P.cls_credible_set(p, 0.1)
        want_perm = Sc.descending_perm(p)
        want_all = Sc.aps_cumsums(p)
        want_tgt = Sc.aps_score(p, t)
    labels, size = ops.credible_sets(pt, 0.1)
    lab, sz = labels.cpu().numpy(), size.cpu().numpy()
    for i in range(p.shape[0]):
        assert lab[i, : sz[i]].tolist() == want_sets[i], i
    assert want_sets[4] == [3] and want_sets[5] == [7] and want_sets[6] == [C - 1]
    allc, perm = ops.aps_scores(pt, None, want_perm=True)
    assert np.array_equal(perm.cpu().numpy(), want_perm)
    assert np.array_equal(allc.cpu().numpy(), want_all, equal_nan=True)
    assert np.array_equal(ops.aps_scores(pt, None).cpu().numpy(), want_all, equal_nan=True)
    assert np.array_equal(ops.aps_scores(pt, torch.from_numpy(t).to(cuda)).cpu().numpy(), want_tgt, equal_nan=True)


@pytest.mark.parametrize("C", [10, 130, 1000, 1024])
@pytest.mark.parametrize("error", [0.05, 0.3])
@pytest.mark.parametrize("n_nan_val", [0, 7, 120, 300])
def test_conformal_sets_of_rows_with_nan_scores_hold_every_class(cuda, C, error, n_nan_val):
    """base.py:51-53 counts the validation scores ABOVE a test score; no number is above a NaN, so the count is 0 and
    every class of a row whose APS scores are NaN is in the set.  Three routes: scores then sets, the fused sort
    kernel, the select kernel (which hands such rows to the sort).  NaN VALIDATION scores are above nothing either, yet
    count in the threshold (1 - error) (n + 1): with n_nan_val of them the boundary moves down the sorted vector, and
    past its start every class is in."""
    from fortuna_b200 import ops

    n_val = 300
    vp = _probs(n_val, C, seed=2)
    vt = np.random.default_rng(3).integers(0, C, n_val).astype(np.int32)
    tp = _probs(40, C, seed=4)
    tp[3, 1] = np.nan
    tp[17, C - 1] = np.array([0xFFC00000], dtype=np.uint32).view(np.float32)[0]
    tp[18, :] = np.nan
    tp[30, 0] = np.inf
    with np.errstate(all="ignore"):
        val_scores, test_scores = Sc.get_scores(vp, vt, tp, "aps")
        val_scores[np.random.default_rng(5).permutation(n_val)[:n_nan_val]] = np.nan
        conds = Sc.conformal_conds_literal(val_scores, test_scores, error)
        assert np.array_equal(conds, Sc.conformal_conds(val_scores, test_scores, error))
    assert conds[3].all() and conds[17].all() and conds[18].all()
    assert n_nan_val < 120 or conds.all()
    sizes = conds.sum(1).astype(np.int32)
    vs = ops.sort_f32_(torch.from_numpy(val_scores.copy()).to(cuda))
    thr = Sc.conformal_threshold(n_val, error)
    tpt = torch.from_numpy(tp).to(cuda)
    m1, s1 = ops.conformal_sets(vs, ops.aps_scores(tpt, None), thr)
    m2, s2 = ops.aps_conformal_sets(tpt, vs, thr)
    m3, s3, sc = ops.aps_conformal_sets(tpt, vs, thr, want_scores=True)
    assert np.array_equal(sc.cpu().numpy(), test_scores, equal_nan=True)
    for route, (m, s) in enumerate(((m1, s1), (m2, s2), (m3, s3))):
        bad = np.argwhere(m.cpu().numpy().astype(bool) != conds)
        assert bad.size == 0, (route, bad[:8].tolist())
        assert np.array_equal(s.cpu().numpy(), sizes), route


@pytest.mark.parametrize("error", [0.05, 0.3])
def test_conformal_bayes_special_values(cuda, error):
    """Weights of exactly zero (log-probability -inf) in some or in all posterior samples (0 / 0 = NaN: nothing compares
    <= a NaN, rank 1), a NaN weight, training points impossible under one or under every sample, a NaN training
    log-probability: counts, region and the NaN pattern of the ESS as predictive/classification.py:557-617 gives them."""
    from fortuna_b200 import ops

    S, n_train, n_test, C = 30, 500, 12, 5
    ell, lsm = _cb_inputs(S, n_train, n_test, C, seed=1)
    lsm[:7, 1, 2] = -np.inf
    lsm[:, 2, 3] = -np.inf
    lsm[4, 3, 1] = np.nan
    ell[5, 17] = -np.inf
    ell[:, 18] = -np.inf
    ell[6, 19] = np.nan
    out = ops.conformal_bayes_sets(torch.from_numpy(ell).to(cuda), torch.from_numpy(lsm).to(cuda), error,
                                   want_counts=True, want_ess=True)
    with np.errstate(all="ignore"):
        cnt, p_train, p_test = Sc.conformal_bayes_counts(ell, lsm, dtype=np.float64)
        ess = Sc.conformal_bayes_ess(lsm, dtype=np.float64)
        amb = (np.abs(p_train - p_test[..., None]) <= 4e-6 * p_test[..., None]).sum(axis=2)
    got = out["rank_counts"].cpu().numpy()
    assert np.all(np.abs(got - cnt) <= amb)
    assert got[2, 3] == 0 and got[3, 1] == 0 and cnt[2, 3] == 0 and cnt[3, 1] == 0
    assert got[0, 0] == cnt[0, 0] == n_train - 1  # the NaN training point is never counted
    region = out["region"].cpu().numpy().astype(bool)
    assert np.array_equal(region, (got + 1).astype(np.float32) > np.float32(error * (n_train + 1)))
    e = out["ess"].cpu().numpy()
    assert np.array_equal(np.isnan(e), np.isnan(ess)) and np.isnan(e[2, 3]) and np.isnan(e[3, 1])
    ok = ~np.isnan(ess)
    np.testing.assert_allclose(e[ok], ess[ok], rtol=2e-5)


@pytest.mark.parametrize("n_data", [100, 1000, 5000])
def test_ece_bin_edges(cuda, n_data):
    """Confidences exactly ON the bin thresholds, one ulp either side of them, 0, 1, just above 1 (a rounded sum) and
    below 0: the bin is the one metric/classification.py's comparisons pick (oracle: bit-exact indices and counts)."""
    from fortuna_b200 import ops

    thr = Sc.ece_thresholds(n_data)
    edge = np.concatenate([thr, np.nextafter(thr, np.float32(2)), np.nextafter(thr, np.float32(-2)),
                           np.array([0.0, 1.0, 1.0000001, 1.5, -0.1, -0.0], dtype=np.float32)]).astype(np.float32)
    r = np.random.default_rng(n_data)
    p = np.concatenate([edge, r.random(n_data - edge.size).astype(np.float32)]) if n_data > edge.size else edge[:n_data]
    p = p[r.permutation(p.size)]
    preds = r.integers(0, 2, p.size).astype(np.int32)
    t = r.integers(0, 2, p.size).astype(np.int32)
    thr = Sc.ece_thresholds(p.size)
    o = ops.ece_bins(torch.from_numpy(p).to(cuda), torch.from_numpy(t).to(cuda), torch.from_numpy(thr).to(cuda),
                     preds=torch.from_numpy(preds).to(cuda), want_bin_idx=True)
    counts, confs, accs, idx = Sc.compute_counts_confs_accs(preds, p, t)
    assert np.array_equal(o["bin_idx"].cpu().numpy(), idx)
    assert np.array_equal(o["counts"].cpu().numpy(), counts)


@pytest.mark.parametrize("C", [10, 130, 1000, 3000])
def test_scores_of_labels_outside_the_row_follow_jax_indexing(cuda, C):
    """probs[target] / inv_perm[target] with a traced index: a negative label counts from the end and the gather clamps
    to the row (jnp.arange(10)[11] == 9) - the kernels used to answer NaN (APS) or read outside the row (simple)."""
    from fortuna_b200 import ops

    p = _probs(64, C, seed=C)
    t = np.random.default_rng(1).integers(0, C, 64).astype(np.int32)
    t[:8] = [-1, -C, -C - 5, C, C + 7, 2**31 - 1, -(2**31), C - 1]
    pt, tt = torch.from_numpy(p).to(cuda), torch.from_numpy(t).to(cuda)
    assert np.array_equal(ops.aps_scores(pt, tt).cpu().numpy(), Sc.aps_score(p, t))
    assert np.array_equal(ops.simple_scores(pt, tt).cpu().numpy(), Sc.simple_score(p, t))
    assert Sc.gather_index(t, C)[:8].tolist() == [C - 1, 0, 0, C - 1, C - 1, C - 1, 0, C - 1]
