"""Pins oracle/sampler.py: the reference's own schedule tests
(tests/fortuna/prob_model/test_step_schedule.py:21-50) plus structural checks of the step."""
import numpy as np
import pytest

from oracle import prng, sampler
from tests import trees


def test_reference_schedule_values():
    s = sampler.constant_schedule(1e-1)
    assert np.allclose(s(0), 1e-1) and np.allclose(s(1), 1e-1)
    s = sampler.cosine_schedule(1e-1, 10)
    assert np.allclose(s(0), 1e-1) and not np.allclose(s(1), s(0))
    assert np.allclose(s(10), 0, atol=1e-8) and np.allclose(s(20), 1e-1)
    s = sampler.polynomial_schedule()
    assert s(1) < s(0)
    s = sampler.constant_schedule_with_cosine_burnin(1e-1, 1e-2, 10)
    assert np.allclose(s(0), 1e-1) and not np.allclose(s(1), s(0))
    assert np.allclose(s(10), 1e-2) and np.allclose(s(11), 1e-2)
    s = sampler.cyclical_cosine_schedule_with_const_burnin(1e-1, 10, 10)
    assert np.allclose(s(0), 1e-1) and np.allclose(s(1), 1e-1) and not np.allclose(s(12), 1e-1)


def test_schedule_argument_errors():
    with pytest.raises(ValueError):
        sampler.constant_schedule(-1.0)
    with pytest.raises(ValueError):
        sampler.cosine_schedule(1e-3, 0)
    with pytest.raises(ValueError):
        sampler.polynomial_schedule(gamma=0.5)
    with pytest.raises(ValueError):
        sampler.constant_schedule_with_cosine_burnin(1e-3, -1e-3, 10)
    with pytest.raises(ValueError):
        sampler.cyclical_cosine_schedule_with_const_burnin(1e-3, -1, 10)


def test_flatten_order_is_sorted_keys():
    paths = ["/".join(p) for p, _ in sampler.tree_flatten(trees.mlp_two_moons_tree())]
    assert paths == sorted(paths)
    assert paths[0].endswith("hidden1/bias") and paths[1].endswith("hidden1/kernel")


def test_sghmc_step_structure():
    params = trees.ragged_tree(0)
    pre = sampler.Preconditioner("identity")
    st = sampler.sghmc_init(params, prng.PRNGKey(0), pre)
    zero = sampler.tree_map(np.zeros_like, params)
    upd, st1 = sampler.sghmc_update(zero, st, 0.01, sampler.constant_schedule(1e-4), pre)
    # key chain: new key = split(key)[1]; noise uses split(key)[0] then one key per leaf in flatten order
    k = prng.split(prng.PRNGKey(0))
    assert st1.rng_key.tolist() == k[1].tolist() and st1.count == 1
    flat = sampler.tree_flatten(params)
    lk = prng.split(k[0], len(flat))
    scale = np.sqrt(np.float32(2 * (1 - 0.01)))
    for (path, leaf), key, (_, m) in zip(flat, lk, sampler.tree_flatten(st1.momentum)):
        np.testing.assert_array_equal(m, (prng.normal(key, leaf.size).reshape(leaf.shape) * scale).astype(np.float32))
    # zero gradient: update = momentum * sqrt(eps)
    for (_, u), (_, m) in zip(sampler.tree_flatten(upd), sampler.tree_flatten(st1.momentum)):
        np.testing.assert_array_equal(u, m * np.sqrt(np.float32(1e-4)))


def test_rmsprop_moments_and_float64_agreement():
    params = trees.mlp_two_moons_tree(0)
    pre = sampler.Preconditioner("rmsprop")
    st = sampler.sghmc_init(params, prng.PRNGKey(1), pre)
    g = sampler.tree_map(lambda p: (p + 0.5).astype(np.float32), params)
    upd, st1 = sampler.sghmc_update(g, st, 0.01, sampler.constant_schedule(1e-3), pre)
    noise = sampler.normal_like_tree(prng.split(prng.PRNGKey(1))[0], g)
    for (_, gl), (_, v), (_, n), (_, u) in zip(
        sampler.tree_flatten(g), sampler.tree_flatten(st1.precond), sampler.tree_flatten(noise), sampler.tree_flatten(upd)
    ):
        g64, n64 = gl.astype(np.float64), n.astype(np.float64)
        v64 = 0.01 * g64**2
        np.testing.assert_allclose(v, v64, rtol=1e-5)
        m64 = g64 * np.sqrt(1e-3) + n64 * np.sqrt(1e-7 + np.sqrt(v64)) * np.sqrt(2 * 0.99)
        np.testing.assert_allclose(u, m64 / (1e-7 + np.sqrt(v64)) * np.sqrt(1e-3), rtol=2e-5, atol=1e-7)


def test_cyclical_phases_and_exploration_step():
    params = trees.mlp_two_moons_tree(0)
    pre = sampler.Preconditioner("identity")
    st = sampler.sghmc_init(params, prng.PRNGKey(2), pre)
    g = sampler.tree_map(lambda p: np.ones_like(p), params)
    phases = [sampler.cyclical_sgld_is_sampling_phase(c, 4, 0.25) for c in range(8)]
    assert phases == [False, True, True, True, False, True, True, True]
    upd, st1 = sampler.cyclical_sgld_update(g, st, 1e-2, 4, 0.25, pre)  # count 0: exploration
    assert st1.rng_key.tolist() == st.rng_key.tolist() and st1.count == 1
    eps0 = sampler.cyclical_cosine_schedule_with_const_burnin(1e-2, 0, 4)(0)
    for _, u in sampler.tree_flatten(upd):
        np.testing.assert_array_equal(u, np.full_like(u, eps0))
    upd, st2 = sampler.cyclical_sgld_update(g, st1, 1e-2, 4, 0.25, pre)  # count 1: SGLD
    assert st2.rng_key.tolist() != st1.rng_key.tolist()


def test_sampling_callbacks_logic():
    do = lambda step, cnt: sampler.sghmc_do_sample(step, cnt, 3, 2, 4)
    assert [s for s in range(1, 12) if do(s, 0)] == [6, 8, 10]
    assert sampler.count_collectable(do, 1, 11) == 3
    dc = lambda step, cnt: sampler.cyclical_sgld_do_sample(step, cnt, 3, 1, 4, 0.25)
    assert [s for s in range(1, 9) if dc(s, 0)] == [1, 2, 3, 5, 6, 7]


def test_selected_leaves_step_equals_full_step():
    """sghmc_update_selected (used by the 110 M-parameter GPU parity test) is sghmc_update on the chosen leaves."""
    params = trees.ragged_tree(3)
    flat = sampler.tree_flatten(params)
    r = np.random.default_rng(0)
    for kind in ("identity", "rmsprop"):
        pre = sampler.Preconditioner(kind)
        st = sampler.sghmc_init(params, prng.PRNGKey(7), pre)
        sched = sampler.constant_schedule(1e-3)
        for _ in range(2):
            grad = sampler.tree_map(lambda p: (0.2 * r.standard_normal(p.shape)).astype(np.float32), params)
            gflat = sampler.tree_flatten(grad)
            sel = [0, 3, len(flat) - 1]
            mflat = sampler.tree_flatten(st.momentum)
            vflat = sampler.tree_flatten(st.precond) if st.precond is not None else None
            upd_s, mom_s, v_s, new_key = sampler.sghmc_update_selected(
                {i: gflat[i][1] for i in sel}, {i: mflat[i][1] for i in sel},
                {i: vflat[i][1] for i in sel} if vflat else None, st.rng_key, st.count, len(flat), 0.01, sched, pre)
            upd, st = sampler.sghmc_update(grad, st, 0.01, sched, pre)
            uflat, m2 = sampler.tree_flatten(upd), sampler.tree_flatten(st.momentum)
            assert np.array_equal(new_key, st.rng_key)
            for i in sel:
                assert np.array_equal(upd_s[i], uflat[i][1]) and np.array_equal(mom_s[i], m2[i][1])
                if kind == "rmsprop":
                    assert np.array_equal(v_s[i], sampler.tree_flatten(st.precond)[i][1])


def test_clip_by_norm_restatement():
    """clip_gradients_by_norm (fortuna/utils/training.py:14-20): identity below the threshold, norm == max above it."""
    r = np.random.default_rng(1)
    g = {"a": r.standard_normal((7, 5)).astype(np.float32), "b": {"c": r.standard_normal(11).astype(np.float32)}}
    norm = float(np.sqrt(sum(float((v.astype(np.float64) ** 2).sum()) for _, v in sampler.tree_flatten(g))))
    out, mult = sampler.clip_gradients_by_norm(g, 10 * norm)
    assert mult == np.float32(1) and all(np.array_equal(a, b) for (_, a), (_, b) in zip(sampler.tree_flatten(out), sampler.tree_flatten(g)))
    out, mult = sampler.clip_gradients_by_norm(g, 0.25 * norm)
    new_norm = float(np.sqrt(sum(float((v.astype(np.float64) ** 2).sum()) for _, v in sampler.tree_flatten(out))))
    np.testing.assert_allclose(new_norm, 0.25 * norm, rtol=1e-5)
    x = np.array([1.0, 1.00390625, 1.01171875, -3.3895314e38, 1e-40], np.float32)
    np.testing.assert_array_equal(sampler.round_to_bf16(x)[:3], np.array([1.0, 1.0, 1.015625], np.float32))
