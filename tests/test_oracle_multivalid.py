"""Sanity of the multivalid oracle (CPU): the permutation is one, calibration lowers the loss, one-shot patches are the
conditional means."""
import numpy as np

from oracle import multivalid as M


def test_permutation():
    for n in (1, 10, 2000):
        p = M.permutation(3, n)
        assert sorted(p.tolist()) == list(range(n))
    assert not np.array_equal(M.permutation(0, 50), M.permutation(1, 50))


def test_multicalibrator_and_batchmvp_lower_their_losses():
    r = np.random.default_rng(0)
    n = 600
    x = r.random(n).astype(np.float32)
    scores = np.clip(x + 0.2, 0, 1).astype(np.float32)
    groups = r.random((n, 2)) < 0.5
    it = M.Iterative("multicalibrator")
    st = it.calibrate(scores, groups, values=x, n_rounds=10, n_buckets=10, eta=1.0)
    assert st["losses"][-1] < st["losses"][0]
    bm = M.Iterative("batchmvp")
    sb = bm.calibrate(scores, groups, n_rounds=10, n_buckets=20, eta=1.0, coverage=0.8, bucket_types=(">=", "<="))
    assert sb["losses"][-1] < sb["losses"][0]  # the pinball loss at the target coverage
    th = bm.apply_patches(groups)
    assert float(np.mean(scores <= th)) > float(np.mean(scores <= 0.0))


def test_one_shot_patches_are_conditional_means():
    r = np.random.default_rng(1)
    x = np.round(r.random(500), 1).astype(np.float32)
    s = r.random(500).astype(np.float32)
    p = M.one_shot_patches(s, x, n_buckets=11, min_prob_b=0.0)
    for j, v in enumerate(M.linspace_f32(0, 1, 11)):
        sel = np.abs(x - v) < np.float32(0.5 / 11)
        if sel.any():
            np.testing.assert_allclose(p[j, 0], s[sel].mean(), rtol=1e-5)
    out = M.one_shot_apply(p, x, 11)
    assert out.shape == x.shape
