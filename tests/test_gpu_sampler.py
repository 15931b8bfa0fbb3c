"""K1 parity: fused SG-MCMC step vs oracle/sampler.py over consecutive steps.

Bit-exact: rng_key chain and per-leaf keys (uint32).  Floats: rtol 1e-5 (erf_inv / FMA contraction
differ from numpy by rounding only), with an absolute floor of 1e-6 * max|x| for cancellation.
"""
import numpy as np
import pytest
import torch

from oracle import prng as oprng
from oracle import sampler as osamp
from tests import trees

pytestmark = pytest.mark.gpu


def _grad_tree(params, seed):
    r = np.random.default_rng(seed)
    return osamp.tree_map(lambda p: (0.3 * r.standard_normal(p.shape) + p).astype(np.float32), params)


def _assert_tree_close(got, want, what, floor=0.0):
    """floor: absolute floor for trees whose steps are much larger than some (cancelling) leaf values."""
    for (path, g), (_, w) in zip(osamp.tree_flatten(got), osamp.tree_flatten(want)):
        atol = max(floor, 1e-6 * max(1e-30, float(np.max(np.abs(w)))))
        np.testing.assert_allclose(g, w, rtol=1e-5, atol=atol, err_msg=f"{what} {'/'.join(path)}")


@pytest.mark.parametrize("precond", ["identity", "rmsprop"])
@pytest.mark.parametrize("tree_fn", [trees.ragged_tree, trees.mlp_two_moons_tree, trees.lenet5_tree])
def test_sghmc_steps_match_oracle(cuda, precond, tree_fn):
    from fortuna_b200 import ops
    from fortuna_b200.prob_model.posterior.sgmcmc.sghmc.sghmc_integrator import sghmc_integrator
    from fortuna_b200.prob_model.posterior.sgmcmc import sgmcmc_preconditioner as pc
    from fortuna_b200.prob_model.posterior.sgmcmc import sgmcmc_step_schedule as sched
    from fortuna_b200.utils.flat_tree import FlatTree

    params_np = tree_fn(1)
    sched_args = dict(init_step_size=1e-3, final_step_size=1e-4, burnin_steps=3)
    o_pre = osamp.Preconditioner(precond)
    o_sched = osamp.constant_schedule_with_cosine_burnin(**sched_args)
    o_state = osamp.sghmc_init(params_np, oprng.PRNGKey(5), o_pre)
    o_params = params_np

    pre = pc.rmsprop_preconditioner() if precond == "rmsprop" else pc.identity_preconditioner()
    tx = sghmc_integrator(0.01, None, ops.PRNGKey(5, cuda), sched.constant_schedule_with_cosine_burnin(**sched_args), pre)
    params = FlatTree.from_tree(params_np, device=cuda)
    state = tx.init(params)

    for step in range(4):
        g_np = _grad_tree(o_params, 100 + step)
        upd, o_state = osamp.sghmc_update(g_np, o_state, 0.01, o_sched, o_pre)
        o_params = osamp.apply_updates(o_params, upd)
        if step % 2 == 0:  # fused apply_gradients path
            params, state = tx.apply(params, FlatTree.from_tree(g_np, device=cuda), state)
        else:  # contract-exact optax path: updates returned, caller applies
            updates, state = tx.update(g_np, state)
            _assert_tree_close(updates.to_numpy_tree(), upd, f"updates step {step}")
            params.flat.add_(updates.flat)  # test-side optax.apply_updates
        assert np.array_equal(ops.key_to_host(state.rng_key), o_state.rng_key), "rng_key chain must be bit-exact"
        assert state.count == o_state.count
        _assert_tree_close(params.to_numpy_tree(), o_params, f"params step {step}")
        _assert_tree_close(state.momentum.to_numpy_tree(), o_state.momentum, f"momentum step {step}")
        if precond == "rmsprop":
            _assert_tree_close(
                state.preconditioner_state.grad_moment_estimates.to_numpy_tree(), o_state.precond, f"v step {step}"
            )


def test_leaf_keys_bit_exact(cuda):
    from fortuna_b200 import ops
    from fortuna_b200.utils.flat_tree import FlatTree

    params = FlatTree.from_tree(trees.ragged_tree(0), device=cuda)
    leaves, tiles = params.tables()
    L = leaves.shape[0]
    key_in = ops.PRNGKey(9, cuda)
    key_out = torch.empty_like(key_in)
    ws = torch.empty((L, 2), dtype=ops.KEY_DTYPE, device=cuda)
    g = params.zeros_like()
    m = params.zeros_like()
    ops.sgmcmc_step(params.flat, g.flat, m.flat, None, None, leaves, tiles, key_in, key_out, ws, ops.make_hparams(1e-3, 0.01))
    k = oprng.split(oprng.PRNGKey(9))
    assert np.array_equal(ops.key_to_host(key_out), k[1])
    assert np.array_equal(ws.cpu().numpy().view(np.uint32), oprng.split(k[0], L))
    # zero gradient, zero momentum: momentum == sqrt(2(1-beta)) * N(0,1) leaf by leaf -> checks the noise layout
    want = osamp.normal_like_tree(k[0], trees.ragged_tree(0))
    scale = np.sqrt(np.float32(2 * (1 - 0.01)), dtype=np.float32)
    got = m.to_numpy_tree()
    for (path, gl), (_, wl) in zip(osamp.tree_flatten(got), osamp.tree_flatten(want)):
        np.testing.assert_allclose(gl, wl * scale, rtol=1e-5, atol=1e-6, err_msg="/".join(path))


@pytest.mark.parametrize("precond", ["identity", "rmsprop"])
def test_cyclical_sgld_both_phases(cuda, precond):
    from fortuna_b200 import ops
    from fortuna_b200.prob_model.posterior.sgmcmc.cyclical_sgld.cyclical_sgld_integrator import (
        cyclical_sgld_integrator,
    )
    from fortuna_b200.prob_model.posterior.sgmcmc import sgmcmc_preconditioner as pc
    from fortuna_b200.utils.flat_tree import FlatTree

    L, ratio, eps0 = 4, 0.5, 1e-3  # steps 0,1 explore; 2,3 sample; 4,5 explore ...
    params_np = trees.ragged_tree(2)
    o_pre = osamp.Preconditioner(precond)
    o_state = osamp.sghmc_init(params_np, oprng.PRNGKey(1), o_pre)
    o_params = params_np
    pre = pc.rmsprop_preconditioner() if precond == "rmsprop" else pc.identity_preconditioner()
    tx = cyclical_sgld_integrator(ops.PRNGKey(1, cuda), eps0, L, ratio, pre)
    params = FlatTree.from_tree(params_np, device=cuda)
    state = tx.init(params)
    phases = []
    for step in range(7):
        g_np = _grad_tree(o_params, 300 + step)
        phases.append(osamp.cyclical_sgld_is_sampling_phase(o_state.count, L, ratio))
        upd, o_state = osamp.cyclical_sgld_update(g_np, o_state, eps0, L, ratio, o_pre)
        o_params = osamp.apply_updates(o_params, upd)
        params, state = tx.apply(params, g_np, state)
        assert np.array_equal(ops.key_to_host(state.sgld_state.rng_key), o_state.rng_key)
        _assert_tree_close(params.to_numpy_tree(), o_params, f"params step {step}")
        _assert_tree_close(state.sgld_state.momentum.to_numpy_tree(), o_state.momentum, f"momentum step {step}")
    assert phases == [False, False, True, True, False, False, True]


def test_argument_errors(cuda):
    from fortuna_b200 import _lib, ops
    from fortuna_b200.utils.flat_tree import FlatTree

    params = FlatTree.from_tree(trees.mlp_two_moons_tree(0), device=cuda)
    leaves, tiles = params.tables()
    key = ops.PRNGKey(0, cuda)
    ws = torch.empty((leaves.shape[0], 2), dtype=ops.KEY_DTYPE, device=cuda)
    with pytest.raises(_lib.FortunaB200Error):  # key_in aliases key_out
        ops.sgmcmc_step(params.flat, params.flat, params.flat, None, None, leaves, tiles, key, key, ws, ops.make_hparams(1e-3))
    with pytest.raises(_lib.FortunaB200Error):  # rmsprop without its moment buffer
        ops.sgmcmc_step(
            params.flat, params.flat, params.flat, None, None, leaves, tiles, key, torch.empty_like(key), ws,
            ops.make_hparams(1e-3, rmsprop=True),
        )
    with pytest.raises(ValueError):  # CPU tensors are rejected: no CPU path
        ops.random_bits(torch.zeros(2, dtype=torch.int32), 4)


@pytest.mark.parametrize("precond", ["identity", "rmsprop"])
@pytest.mark.parametrize("kind", ["sghmc", "cyclical"])
def test_fused_log_joint_gradient_matches_oracle(cuda, precond, kind):
    """fb200_sgmcmc_step_joint_f32 / fb200_sgd_explore_step_joint_f32: the kernel forms
    g = (n_data / batch) * grad_loglik - prec * theta itself (joint/base.py:54-122, prior/gaussian.py:29-32) and returns
    sum(theta^2) of the pre-update parameters; the oracle is stepped with that gradient formed in float32 numpy."""
    from fortuna_b200 import ops
    from fortuna_b200.prob_model.posterior.sgmcmc import sgmcmc_preconditioner as pc
    from fortuna_b200.prob_model.posterior.sgmcmc import sgmcmc_step_schedule as sched
    from fortuna_b200.prob_model.posterior.sgmcmc.cyclical_sgld.cyclical_sgld_integrator import cyclical_sgld_integrator
    from fortuna_b200.prob_model.posterior.sgmcmc.sghmc.sghmc_integrator import sghmc_integrator
    from fortuna_b200.utils.flat_tree import FlatTree

    params_np = trees.lenet5_tree(2)
    lik_scale, prec = np.float32(60000 / 128), np.float32(np.exp(-0.7))
    o_pre = osamp.Preconditioner(precond)
    pre = pc.rmsprop_preconditioner() if precond == "rmsprop" else pc.identity_preconditioner()
    if kind == "sghmc":
        o_state = osamp.sghmc_init(params_np, oprng.PRNGKey(9), o_pre)
        o_sched = osamp.constant_schedule(1e-5)
        tx = sghmc_integrator(0.05, None, ops.PRNGKey(9, cuda), sched.constant_schedule(1e-5), pre)
    else:
        o_state = osamp.sghmc_init(params_np, oprng.PRNGKey(9), o_pre)  # the oracle steps the inner SGLD state
        tx = cyclical_sgld_integrator(ops.PRNGKey(9, cuda), 1e-5, 4, 0.5, pre)  # steps 0,1 explore; 2,3 sample; ...
    o_params = params_np
    params = FlatTree.from_tree(params_np, device=cuda)
    state = tx.init(params)
    sq = torch.empty(1, dtype=torch.float64, device=cuda)
    for step in range(6):
        gl = _grad_tree(o_params, 300 + step)  # stands for the gradient of the batch log-likelihood sum
        g_joint = osamp.tree_map(lambda a, p: (lik_scale * a - prec * p).astype(np.float32), gl, o_params)
        want_sq = sum(float(np.sum(p.astype(np.float64) ** 2)) for _, p in osamp.tree_flatten(o_params))
        if kind == "sghmc":
            upd, o_state = osamp.sghmc_update(g_joint, o_state, 0.05, o_sched, o_pre)
        else:
            upd, o_state = osamp.cyclical_sgld_update(g_joint, o_state, 1e-5, 4, 0.5, o_pre)
        o_params = osamp.apply_updates(o_params, upd)
        params, state = tx.apply(params, FlatTree.from_tree(gl, device=cuda), state, joint=(lik_scale, prec, sq))
        np.testing.assert_allclose(float(sq.item()), want_sq, rtol=1e-6)
        _assert_tree_close(params.to_numpy_tree(), o_params, f"params step {step}")
    inner = state.sgld_state if kind == "cyclical" else state
    o_inner = o_state
    assert np.array_equal(ops.key_to_host(inner.rng_key), o_inner.rng_key)
    _assert_tree_close(inner.momentum.to_numpy_tree(), o_inner.momentum, "momentum")


@pytest.mark.parametrize("precond", ["identity", "rmsprop"])
def test_bert_base_tree_steps_match_oracle(cuda, precond):
    """BASELINE config 5: the 110 M-parameter BERT-base-shaped tree the sampler benchmark is quoted on (201 leaves,
    53 K tiles, flat offsets beyond 2^24, a 23.4 M-element embedding leaf whose half-offset is 11.7 M).  Two SGHMC steps
    on the GPU; the key chain is compared bit for bit and four leaves - the embedding, one 768 x 3072 kernel, one bias and
    the 2-element classifier bias - against oracle/sampler.py's arithmetic under the FULL tree's key schedule
    (sghmc_integrator.py:59-92, utils/random.py:11-23), every element, rtol 1e-5."""
    from fortuna_b200 import ops
    from fortuna_b200.prob_model.posterior.sgmcmc import sgmcmc_preconditioner as pc
    from fortuna_b200.prob_model.posterior.sgmcmc import sgmcmc_step_schedule as sched
    from fortuna_b200.prob_model.posterior.sgmcmc.sghmc.sghmc_integrator import sghmc_integrator
    from fortuna_b200.utils.flat_tree import FlatTree

    params = FlatTree.from_shapes(trees.bert_base_shapes(), device=cuda)
    gen = torch.Generator(device=cuda)
    gen.manual_seed(3)
    params.flat.normal_(0.0, 0.02, generator=gen)
    n_leaves = len(params.lens)
    assert params.n_params > 109_000_000 and max(params.lens) == 30522 * 768
    want = {"embedding": params.lens.index(30522 * 768), "classifier_bias": params.lens.index(2),
            "kernel": params.lens.index(768 * 3072), "bias": params.lens.index(3072)}
    sel = sorted(set(want.values()))
    assert any(params.offsets[i] > (1 << 24) for i in sel)

    def leaf(ft, i):
        o, n = ft.offsets[i], ft.lens[i]
        return ft.flat[o : o + n].cpu().numpy().reshape(ft.shapes[i])

    o_pre = osamp.Preconditioner(precond)
    o_sched = osamp.constant_schedule(2e-5)
    pre = pc.rmsprop_preconditioner() if precond == "rmsprop" else pc.identity_preconditioner()
    tx = sghmc_integrator(0.01, None, ops.PRNGKey(11, cuda), sched.constant_schedule(2e-5), pre)
    state = tx.init(params)
    o_p = {i: leaf(params, i) for i in sel}
    o_m = {i: np.zeros_like(o_p[i]) for i in sel}
    o_v = {i: np.zeros_like(o_p[i]) for i in sel} if precond == "rmsprop" else None
    o_key, o_count = oprng.PRNGKey(11), 0
    grad = params.empty_like()
    for step in range(2):
        grad.flat.normal_(0.0, 0.5, generator=gen)
        g_sel = {i: leaf(grad, i) for i in sel}
        upd, o_m, o_v, o_key = osamp.sghmc_update_selected(g_sel, o_m, o_v, o_key, o_count, n_leaves, 0.01, o_sched, o_pre)
        o_count += 1
        o_p = {i: (o_p[i] + upd[i]).astype(np.float32) for i in sel}
        params, state = tx.apply(params, grad, state)
        assert np.array_equal(ops.key_to_host(state.rng_key), o_key), "rng_key chain must be bit-exact"
        for name, i in want.items():
            w = o_p[i]
            np.testing.assert_allclose(leaf(params, i), w, rtol=1e-5, atol=1e-6 * float(np.max(np.abs(w))),
                                       err_msg=f"params {name} step {step}")
            wm = o_m[i]
            np.testing.assert_allclose(leaf(state.momentum, i), wm, rtol=1e-5, atol=1e-6 * float(np.max(np.abs(wm))),
                                       err_msg=f"momentum {name} step {step}")
            if precond == "rmsprop":
                np.testing.assert_allclose(leaf(state.preconditioner_state.grad_moment_estimates, i), o_v[i], rtol=1e-5,
                                           err_msg=f"v {name} step {step}")


@pytest.mark.parametrize("precond", ["identity", "rmsprop"])
@pytest.mark.parametrize("joint", [False, True])
def test_gradient_clipping_fused_into_the_step(cuda, precond, joint):
    """clip_grandients_by_norm (fortuna/utils/training.py:14-20) as a norm pre-pass whose multiplier the
    step kernel applies while loading the gradient - vs the oracle: clip the (log-joint) gradient, then the plain step.
    Both regimes: a threshold above the norm (multiplier exactly 1) and one far below it."""
    from fortuna_b200 import ops
    from fortuna_b200.prob_model.posterior.sgmcmc import sgmcmc_preconditioner as pc
    from fortuna_b200.prob_model.posterior.sgmcmc import sgmcmc_step_schedule as sched
    from fortuna_b200.prob_model.posterior.sgmcmc.sghmc.sghmc_integrator import sghmc_integrator
    from fortuna_b200.utils.flat_tree import FlatTree

    params_np = trees.ragged_tree(4, scale=0.3)
    lik_scale, prec = np.float32(7.5), np.float32(0.6)
    for max_norm in (1e6, 0.05):
        o_pre = osamp.Preconditioner(precond)
        o_state = osamp.sghmc_init(params_np, oprng.PRNGKey(2), o_pre)
        o_params = params_np
        pre = pc.rmsprop_preconditioner() if precond == "rmsprop" else pc.identity_preconditioner()
        tx = sghmc_integrator(0.02, None, ops.PRNGKey(2, cuda), sched.constant_schedule(1e-3), pre)
        params = FlatTree.from_tree(params_np, device=cuda)
        state = tx.init(params)
        for step in range(3):
            g_np = _grad_tree(o_params, 40 + step)
            g_used = osamp.tree_map(lambda g, p: (lik_scale * g + (-prec) * p).astype(np.float32), g_np, o_params) if joint else g_np
            g_clip, mult = osamp.clip_gradients_by_norm(g_used, max_norm)
            assert (mult == 1) == (max_norm > 1)
            upd, o_state = osamp.sghmc_update(g_clip, o_state, 0.02, osamp.constant_schedule(1e-3), o_pre)
            o_params = osamp.apply_updates(o_params, upd)
            params, state = tx.apply(params, FlatTree.from_tree(g_np, device=cuda), state,
                                     joint=(float(lik_scale), float(prec), None) if joint else None, max_grad_norm=max_norm)
            # parameters of scale 0.3 take steps of scale 0.05: a 1-element leaf that lands near zero keeps the absolute
            # rounding of the larger operands (one ulp of 0.5 = 6e-8)
            _assert_tree_close(params.to_numpy_tree(), o_params, f"params step {step} max_norm {max_norm}", floor=2e-7)
            _assert_tree_close(state.momentum.to_numpy_tree(), o_state.momentum, f"momentum step {step}", floor=2e-7)


def test_a_nan_gradient_makes_the_clipped_step_nan_everywhere(cuda):
    """utils/training.py:18: jnp.minimum(1, max_norm / (1e-7 + norm)) keeps a NaN norm, so ONE NaN gradient entry turns
    every gradient - and with it every parameter and momentum - into NaN (fminf would have answered 1 and confined the
    NaN to its own entry)."""
    from fortuna_b200 import ops
    from fortuna_b200.prob_model.posterior.sgmcmc import sgmcmc_preconditioner as pc
    from fortuna_b200.prob_model.posterior.sgmcmc import sgmcmc_step_schedule as sched
    from fortuna_b200.prob_model.posterior.sgmcmc.sghmc.sghmc_integrator import sghmc_integrator
    from fortuna_b200.utils.flat_tree import FlatTree

    params_np = trees.ragged_tree(4, scale=0.3)
    g_np = _grad_tree(params_np, 40)
    first = osamp.tree_flatten(g_np)[0][1]
    first.reshape(-1)[0] = np.nan
    g_clip, mult = osamp.clip_gradients_by_norm(g_np, 0.05)
    assert np.isnan(mult) and all(np.isnan(leaf).all() for _, leaf in osamp.tree_flatten(g_clip))
    tx = sghmc_integrator(0.02, None, ops.PRNGKey(2, cuda), sched.constant_schedule(1e-3), pc.identity_preconditioner())
    params = FlatTree.from_tree(params_np, device=cuda)
    state = tx.init(params)
    params, state = tx.apply(params, FlatTree.from_tree(g_np, device=cuda), state, max_grad_norm=0.05)
    assert all(np.isnan(leaf).all() for _, leaf in osamp.tree_flatten(params.to_numpy_tree()))
    assert all(np.isnan(leaf).all() for _, leaf in osamp.tree_flatten(state.momentum.to_numpy_tree()))


def test_bf16_gradient_input(cuda):
    """the step reads a bfloat16 gradient vector (18 instead of 20 B/param) - identical to the float32 step
    on the same values (bf16 -> f32 is exact), for the vector path and the ragged leaves, with clipping on top."""
    from fortuna_b200 import ops
    from fortuna_b200.prob_model.posterior.sgmcmc import sgmcmc_preconditioner as pc
    from fortuna_b200.prob_model.posterior.sgmcmc import sgmcmc_step_schedule as sched
    from fortuna_b200.prob_model.posterior.sgmcmc.sghmc.sghmc_integrator import sghmc_integrator
    from fortuna_b200.utils.flat_tree import FlatTree

    params_np = trees.ragged_tree(6)
    g_np = osamp.tree_map(osamp.round_to_bf16, _grad_tree(params_np, 77))
    outs = []
    for as_bf16 in (False, True):
        tx = sghmc_integrator(0.05, None, ops.PRNGKey(8, cuda), sched.constant_schedule(2e-3), pc.rmsprop_preconditioner())
        params = FlatTree.from_tree(params_np, device=cuda)
        state = tx.init(params)
        g = FlatTree.from_tree(g_np, device=cuda)
        if as_bf16:
            g = FlatTree(g.flat.to(torch.bfloat16), g.paths, g.shapes, g.offsets)
            assert torch.equal(g.flat.float(), FlatTree.from_tree(g_np, device=cuda).flat)
        for _ in range(2):
            params, state = tx.apply(params, g, state, max_grad_norm=0.5)
        outs.append((params.flat.clone(), state.momentum.flat.clone(), ops.key_to_host(state.rng_key)))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][2], outs[1][2])
    # and against the oracle, one step, no clipping
    o_pre = osamp.Preconditioner("rmsprop")
    o_state = osamp.sghmc_init(params_np, oprng.PRNGKey(8), o_pre)
    upd, o_state = osamp.sghmc_update(g_np, o_state, 0.05, osamp.constant_schedule(2e-3), o_pre)
    tx = sghmc_integrator(0.05, None, ops.PRNGKey(8, cuda), sched.constant_schedule(2e-3), pc.rmsprop_preconditioner())
    params = FlatTree.from_tree(params_np, device=cuda)
    state = tx.init(params)
    g = FlatTree.from_tree(g_np, device=cuda)
    params, state = tx.apply(params, FlatTree(g.flat.to(torch.bfloat16), g.paths, g.shapes, g.offsets), state)
    _assert_tree_close(params.to_numpy_tree(), osamp.apply_updates(params_np, upd), "params (bf16 gradient)")
