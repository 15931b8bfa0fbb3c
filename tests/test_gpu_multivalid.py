"""Multivalid scorers: the statistics kernel against brute force, the iterative and one-shot classes against
the numpy restatement round by round (the chosen (tau, g, v, c) exactly, patches / losses / patched values within float32
rounding), and the reference tests' call patterns (tests/fortuna/test_conformal_methods.py:346-826)."""
import numpy as np
import pytest
import torch

from oracle import multivalid as M

pytestmark = pytest.mark.gpu


def _data(n, G=3, seed=0):
    r = np.random.default_rng(seed)
    x = r.random(n).astype(np.float32)
    scores = np.clip(x + 0.15 * r.standard_normal(n) + 0.1 * (x > 0.5), 0, 1).astype(np.float32)
    groups = r.random((n, G)) < 0.5
    return x, scores, groups


@pytest.mark.parametrize("n,G,nb,D,top", [(1000, 3, 10, 1, False), (2500, 1, 100, 1, False), (700, 2, 7, 4, True)])
@pytest.mark.parametrize("mode", [0, 1])
def test_stats_kernel_against_brute_force(cuda, n, G, nb, D, top, mode):
    from fortuna_b200 import ops

    r = np.random.default_rng(n)
    values = r.random((n, D)).astype(np.float32)
    values[::7] = np.float32(0.5)  # values on bucket boundaries
    scores = r.random((n, D)).astype(np.float32)
    groups = r.random((n, G)) < 0.6
    buckets = M.linspace_f32(0, 1, nb)
    out = ops.multivalid_stats(torch.from_numpy(values).to(cuda), torch.from_numpy(scores).to(cuda),
                               torch.from_numpy(groups.astype(np.uint8)).to(cuda), torch.from_numpy(buckets).to(cuda), top, mode
                               ).cpu().numpy()
    for tau in range(3):
        for g in range(G):
            for j in (0, 1, nb // 2, nb - 1):
                for c in range(D):
                    b = M.get_b(groups, values if D > 1 else values[:, 0], buckets[j], g, c, tau, nb, top)
                    x = values[:, c]
                    stat = (scores[:, c] <= x).astype(np.float64) if mode == 1 else scores[:, c].astype(np.float64)
                    assert out[tau, g, j, c, 0] == b.sum()
                    np.testing.assert_allclose(out[tau, g, j, c, 1], stat[b].sum(), rtol=1e-12, atol=1e-12)
                    np.testing.assert_allclose(out[tau, g, j, c, 2], x[b].astype(np.float64).sum(), rtol=1e-12, atol=1e-12)


def test_permutation_is_the_oracles(cuda):
    from fortuna_b200.conformal.multivalid.iterative.base import jax_permutation

    for seed, n in ((0, 10), (3, 1000), (7, 5000)):
        assert np.array_equal(jax_permutation(seed, n, cuda), M.permutation(seed, n))


def _same_run(got_status, got_patches, want_status, want_patches):
    assert got_status["n_rounds"] == want_status["n_rounds"] and got_status["converged"] == want_status["converged"]
    np.testing.assert_allclose(got_status["losses"], want_status["losses"], rtol=2e-5)
    assert len(got_patches) == len(want_patches)
    for a, b in zip(got_patches, want_patches):
        assert a[:2] == b[:2] and a[3] == b[3] and abs(a[2] - b[2]) < 1e-7, (a, b)
        np.testing.assert_allclose(a[4], b[4], rtol=2e-4, atol=1e-7)


@pytest.mark.parametrize("patch_type", ["additive", "multiplicative"])
@pytest.mark.parametrize("bucket_types", [("<=", ">="), ("=",), ("=", ">=", "<=")])
def test_multicalibrator_matches_oracle(cuda, patch_type, bucket_types):
    from fortuna_b200.conformal import Multicalibrator

    x, scores, groups = _data(1200)
    tx, ts, tg = _data(300, seed=5)
    v0 = np.clip(x + 0.2, 0, 1).astype(np.float32)
    kw = dict(n_rounds=8, n_buckets=10, eta=0.5, bucket_types=bucket_types, patch_type=patch_type)
    mc = Multicalibrator(seed=1)
    tv, status = mc.calibrate(scores=scores, groups=groups, values=v0, test_groups=tg, test_values=tx, **kw)
    o = M.Iterative("multicalibrator", seed=1)
    want = o.calibrate(scores, groups, values=v0, **kw)
    _same_run(status, mc.patches, want, o.patches)
    np.testing.assert_allclose(tv.cpu().numpy(), o.apply_patches(tg, tx), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(mc.calibration_error(ts, tg, tv).cpu().numpy(), o.calibration_error(ts, tg, tv.cpu().numpy()),
                               rtol=1e-4, atol=1e-7)


def test_batchmvp_matches_oracle(cuda):
    from fortuna_b200.conformal import BatchMVPConformalClassifier, BatchMVPConformalRegressor

    x, scores, groups = _data(1500, seed=2)
    tx, ts, tg = _data(200, seed=6)
    kw = dict(n_rounds=6, n_buckets=20, eta=0.5, coverage=0.9)
    bm = BatchMVPConformalRegressor(seed=2)
    status = bm.calibrate(scores=scores, groups=groups, **kw)
    o = M.Iterative("batchmvp", seed=2)
    want = o.calibrate(scores, groups, bucket_types=(">=", "<="), **kw)
    _same_run(status, bm.patches, want, o.patches)
    th = bm.apply_patches(tg)
    np.testing.assert_allclose(th.cpu().numpy(), o.apply_patches(tg), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(bm.calibration_error(ts, tg, th).cpu().numpy(), o.calibration_error(ts, tg, th.cpu().numpy()),
                               rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(float(bm.pinball_loss(th, ts, 0.9)), o.loss(th.cpu().numpy(), ts[:, None]), rtol=1e-5)
    sets = BatchMVPConformalClassifier.conformal_set(np.array([[0.1, 0.7], [0.5, 0.2]]), np.array([0.3, 0.6]))
    assert sets == [[0], [0, 1]]
    with pytest.raises(NotImplementedError):
        bm.calibrate(scores=scores, n_rounds=2, n_buckets=4, patch_type="multiplicative")


def test_top_label_and_binary_multicalibrators_match_oracle(cuda):
    from fortuna_b200.conformal import BinaryClassificationMulticalibrator, TopLabelMulticalibrator

    r = np.random.default_rng(4)
    n, C = 900, 4
    logits = r.standard_normal((n, C)).astype(np.float32)
    probs = (np.exp(logits) / np.exp(logits).sum(1, keepdims=True)).astype(np.float32)
    targets = np.array([r.choice(C, p=p / p.sum()) for p in probs.astype(np.float64)], dtype=np.int32)
    groups = r.random((n, 2)) < 0.5
    kw = dict(n_rounds=5, n_buckets=8, eta=0.5)
    tl = TopLabelMulticalibrator(n_classes=C, seed=3)
    status = tl.calibrate(targets=targets, groups=groups, probs=probs, **kw)
    o = M.Iterative("top_label", seed=3, n_classes=C)
    want = o.calibrate((targets[:, None] == np.arange(C)[None]).astype(np.float32), groups, values=probs, **kw)
    _same_run(status, tl.patches, want, o.patches)
    np.testing.assert_allclose(tl.apply_patches(groups, probs).cpu().numpy(), o.apply_patches(groups, probs), rtol=1e-5, atol=1e-6)
    # probabilities initialised by the class itself: jax's normal stream from PRNGKey(seed)
    st2 = tl.calibrate(targets=targets, groups=groups, n_rounds=2, n_buckets=8)
    want2 = o.calibrate((targets[:, None] == np.arange(C)[None]).astype(np.float32), groups, n_rounds=2, n_buckets=8)
    _same_run(st2, tl.patches, want2, o.patches)
    # binary
    y = (r.random(n) < probs[:, 0]).astype(np.int32)
    bc = BinaryClassificationMulticalibrator(seed=0)
    sb = bc.calibrate(targets=y, groups=groups, probs=probs[:, 0].copy(), **kw)
    ob = M.Iterative("multicalibrator", seed=0)
    wb = ob.calibrate(y.astype(np.float32), groups, values=probs[:, 0].copy(), **kw)
    _same_run(sb, bc.patches, wb, ob.patches)
    with pytest.raises(ValueError):
        bc.calibrate(targets=y.astype(np.float32), probs=probs[:, 0].copy())  # targets must be integers


def test_one_shot_multicalibrators_match_oracle(cuda):
    from fortuna_b200.conformal import (OneShotBinaryClassificationMulticalibrator, OneShotMulticalibrator,
                                        OneShotTopLabelMulticalibrator)

    x, scores, _ = _data(2000, seed=8)
    x = np.round(x, 2).astype(np.float32)  # repeated values: several data points per unique value
    tx = np.round(np.random.default_rng(9).random(500), 2).astype(np.float32)
    os_ = OneShotMulticalibrator()
    tv = os_.calibrate(scores=scores, values=x, test_values=tx, n_buckets=20)
    want = M.one_shot_patches(scores, x, n_buckets=20)
    np.testing.assert_allclose(os_.patches, want, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(tv.cpu().numpy(), M.one_shot_apply(want, tx, 20), rtol=1e-5, atol=1e-7)
    y = (np.random.default_rng(1).random(2000) < x).astype(np.int32)
    ob = OneShotBinaryClassificationMulticalibrator()
    ob.calibrate(targets=y, probs=x, n_buckets=10)
    np.testing.assert_allclose(ob.patches, M.one_shot_patches(y.astype(np.float32), x, n_buckets=10), rtol=1e-5, atol=1e-7)
    r = np.random.default_rng(2)
    C = 3
    probs = np.round(r.dirichlet(np.ones(C), 800), 2).astype(np.float32)
    t = r.integers(0, C, 800).astype(np.int32)
    ot = OneShotTopLabelMulticalibrator(n_classes=C)
    out = ot.calibrate(targets=t, probs=probs, test_probs=probs, n_buckets=10)
    wp = M.one_shot_patches((t[:, None] == np.arange(C)[None]).astype(np.float32), probs, n_buckets=10, top_label=True)
    np.testing.assert_allclose(ot.patches, wp, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(out.cpu().numpy(), M.one_shot_apply(wp, probs, 10, top_label=True), rtol=1e-5, atol=1e-7)


def test_reference_call_patterns(cuda):
    """tests/fortuna/test_conformal_methods.py:486-549 (Multicalibrator) and :346-411 (BatchMVP): same calls, same errors."""
    from fortuna_b200.conformal import BatchMVPConformalRegressor, Multicalibrator

    r = np.random.default_rng(0)
    size, test_size = 10, 20
    scores = r.random(size).astype(np.float32)
    groups = r.random((size, 3)) < 0.5
    values = np.zeros(size, np.float32)
    test_scores = r.random(test_size).astype(np.float32)
    test_groups = r.random((test_size, 3)) < 0.5
    for cls, vname, tname in ((Multicalibrator, "values", "test_values"), (BatchMVPConformalRegressor, "thresholds", "test_thresholds")):
        mc = cls()
        mc.calibrate(scores=scores, n_rounds=3, n_buckets=4)
        mc.calibrate(scores=scores, groups=groups, n_rounds=3, n_buckets=4, **{vname: values})
        test_values, status = mc.calibrate(scores=scores, groups=groups, test_groups=test_groups, n_rounds=3, n_buckets=4)
        assert test_values.shape == (test_size,) and set(status) == {"n_rounds", "losses", "converged"}
        with pytest.raises(ValueError):
            mc.calibrate(scores=scores, groups=groups, test_groups=test_groups, n_rounds=3, n_buckets=4, **{vname: values})
        with pytest.raises(ValueError):
            mc.calibrate(scores=scores, groups=groups, n_rounds=3, n_buckets=4, **{vname: values, tname: test_values})
        with pytest.raises(ValueError):
            mc.calibrate(scores=scores, groups=groups, test_groups=test_groups, n_rounds=3, n_buckets=4, **{tname: test_values})
        mc.calibrate(scores=scores, groups=groups, n_rounds=3, n_buckets=4)
        tv = mc.apply_patches(test_groups)
        tv = mc.apply_patches(test_groups, tv)
        err = mc.calibration_error(test_scores, test_groups, tv)
        assert err.shape == (3,)
        mc.calibrate(scores=scores, groups=groups, test_groups=test_groups, n_rounds=3, n_buckets=4,
                     **{vname: values, tname: tv.cpu().numpy()})
