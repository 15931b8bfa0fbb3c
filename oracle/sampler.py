"""SG-MCMC sampler arithmetic restated in numpy float32.  TEST INFRASTRUCTURE ONLY.

Follows, line by line:
  fortuna/prob_model/posterior/sgmcmc/sghmc/sghmc_integrator.py:51-92        (SGHMC step)
  fortuna/prob_model/posterior/sgmcmc/cyclical_sgld/cyclical_sgld_integrator.py:43-100
  fortuna/prob_model/posterior/sgmcmc/sgmcmc_preconditioner.py:60-93,117-130 (RMSprop / identity)
  fortuna/prob_model/posterior/sgmcmc/sgmcmc_step_schedule.py:26-27,51-53,80-81,111-114,144-147
  fortuna/utils/random.py:11-23                                              (per-leaf normal tree)
  fortuna/prob_model/posterior/sgmcmc/sghmc/sghmc_callback.py:49-53,
  .../cyclical_sgld/cyclical_sgld_callback.py:55-59                           (which steps are stored)
plus flax 0.8.5 ``TrainState.apply_gradients`` / optax 0.1.9 ``apply_updates`` (params + updates)
and optax ``sgd(1.0)`` = scale(-1.0).

A "tree" here is a nested dict of float32 numpy arrays; leaves are visited in
jax.tree_util order = dict keys sorted at every level.  Parity unpinned: the
reference's tests only smoke-run the sampler (tests/fortuna/prob_model/test_train.py:57-60).
"""

import numpy as np

from . import prng

F32 = np.float32


# ----------------------------------------------------------------------------- pytree helpers
def tree_flatten(tree, prefix=()):
    """[(path, leaf)] in jax flatten order (sorted dict keys, depth first)."""
    if isinstance(tree, dict):
        out = []
        for k in sorted(tree.keys()):
            out.extend(tree_flatten(tree[k], prefix + (k,)))
        return out
    return [(prefix, tree)]


def tree_map(fn, tree, *rest):
    if isinstance(tree, dict):
        return {k: tree_map(fn, tree[k], *[r[k] for r in rest]) for k in tree}
    return fn(tree, *rest)


def tree_unflatten_like(tree, leaves):
    it = iter(leaves)

    def rec(t):
        if isinstance(t, dict):
            return {k: rec(t[k]) for k in sorted(t.keys())}
        return next(it)

    return rec(tree)


# ----------------------------------------------------------------------------- step schedules
def constant_schedule(init_step_size):
    if not init_step_size >= 0:
        raise ValueError("`init_step_size` should be >= 0.")
    return lambda step: F32(init_step_size)


def cosine_schedule(init_step_size, total_steps):
    if not init_step_size >= 0:
        raise ValueError("`init_step_size` should be >= 0.")
    if not total_steps > 0:
        raise ValueError("`total_steps` should be > 0.")

    def schedule(step):
        t = F32(np.int32(step)) / F32(total_steps)
        return F32(0.5 * init_step_size) * (F32(1) + np.cos(t * F32(np.pi), dtype=F32))

    return schedule


def polynomial_schedule(a=1.0, b=1.0, gamma=0.55):
    if not 0.5 < gamma <= 1.0:
        raise ValueError("`gamma` should be in (0.5, 1.0] range.")

    def schedule(step):
        return F32(a) * np.power(F32(b) + F32(np.int32(step)), F32(-gamma), dtype=F32)

    return schedule


def constant_schedule_with_cosine_burnin(init_step_size, final_step_size, burnin_steps):
    if not init_step_size >= 0:
        raise ValueError("`init_step_size` should be >= 0.")
    if not final_step_size >= 0:
        raise ValueError("`final_step_size` should be >= 0.")
    if not burnin_steps >= 0:
        raise ValueError("`burnin_steps` should be >= 0.")

    def schedule(step):
        with np.errstate(divide="ignore", invalid="ignore"):
            t = np.minimum(F32(np.int32(step)) / F32(burnin_steps), F32(1.0))
        coef = (F32(1) + np.cos(t * F32(np.pi), dtype=F32)) * F32(0.5)
        return coef * F32(init_step_size) + (F32(1) - coef) * F32(final_step_size)

    return schedule


def cyclical_cosine_schedule_with_const_burnin(init_step_size, burnin_steps, cycle_length):
    if not init_step_size >= 0:
        raise ValueError("`init_step_size` should be >= 0.")
    if not burnin_steps >= 0:
        raise ValueError("`burnin_steps` should be >= 0.")
    if not cycle_length >= 0:
        raise ValueError("`cycle_length` should be >= 0.")

    def schedule(step):
        t = np.maximum(F32(np.int32(step) - np.int32(burnin_steps) - np.int32(1)), F32(0.0))
        t = np.fmod(t, F32(cycle_length)) / F32(cycle_length)  # t >= 0, so fmod == jnp.remainder
        return F32(0.5 * init_step_size) * (F32(1) + np.cos(t * F32(np.pi), dtype=F32))

    return schedule


# ----------------------------------------------------------------------------- noise tree
def leaf_keys(key, n_leaves):
    """generate_rng_like_tree: split(key, num_leaves) (fortuna/utils/random.py:11-14)."""
    return prng.split(key, n_leaves)


def normal_like_tree(key, tree):
    """generate_random_normal_like_tree (fortuna/utils/random.py:17-23)."""
    flat = tree_flatten(tree)
    keys = leaf_keys(key, len(flat))
    leaves = [prng.normal(k, leaf.size).reshape(leaf.shape) for k, (_, leaf) in zip(keys, flat)]
    return tree_unflatten_like(tree, leaves)


# ----------------------------------------------------------------------------- SGHMC
class Preconditioner:
    """identity (sgmcmc_preconditioner.py:108-138) or RMSprop (:44-105)."""

    def __init__(self, kind="identity", running_average_factor=0.99, eps=1.0e-7):
        assert kind in ("identity", "rmsprop")
        self.kind = kind
        self.alpha = running_average_factor
        self.eps = eps

    def init(self, params):
        if self.kind == "identity":
            return None
        return tree_map(np.zeros_like, params)

    def update(self, grad, v):
        if self.kind == "identity":
            return None
        a = F32(self.alpha)
        b = F32(1 - self.alpha)  # python double (1 - alpha), then weak-typed to f32
        return tree_map(lambda e, g: e * a + (g * g) * b, v, grad)

    def m_inv(self, vec, v):
        if self.kind == "identity":
            return vec
        eps = F32(self.eps)
        return tree_map(lambda e, x: x / (eps + np.sqrt(e)), v, vec)

    def m_sqrt(self, vec, v):
        if self.kind == "identity":
            return vec
        eps = F32(self.eps)
        return tree_map(lambda e, x: x * np.sqrt(eps + np.sqrt(e)), v, vec)


class SGHMCState:
    def __init__(self, count, rng_key, momentum, precond):
        self.count = int(count)
        self.rng_key = np.asarray(rng_key, dtype=np.uint32)
        self.momentum = momentum
        self.precond = precond


def sghmc_init(params, rng_key, preconditioner):
    return SGHMCState(0, rng_key, tree_map(np.zeros_like, params), preconditioner.init(params))


def sghmc_update(grad, state, momentum_decay, step_schedule, preconditioner, momentum_resample_steps=None):
    """sghmc_integrator.update_fn (:59-92).  Returns (updates, new_state)."""
    step_size = F32(step_schedule(state.count))
    v = preconditioner.update(grad, state.precond)
    keys = prng.split(state.rng_key)
    key, new_key = keys[0], keys[1]
    noise = normal_like_tree(key, grad)
    noise = preconditioner.m_sqrt(noise, v)
    if momentum_resample_steps is not None and state.count % momentum_resample_steps == 0:
        momentum = tree_map(np.zeros_like, grad)
    else:
        momentum = state.momentum
    beta = F32(momentum_decay)
    sqrt_eps = np.sqrt(step_size, dtype=F32)
    noise_scale = np.sqrt(F32(2 * (1 - momentum_decay)), dtype=F32)
    momentum = tree_map(lambda m, g, n: beta * m + g * sqrt_eps + n * noise_scale, momentum, grad, noise)
    updates = preconditioner.m_inv(momentum, v)
    updates = tree_map(lambda m: m * sqrt_eps, updates)
    return updates, SGHMCState(state.count + 1, new_key, momentum, v)


def sghmc_update_selected(grad_sel, mom_sel, v_sel, rng_key, count, n_leaves, momentum_decay, step_schedule,
                          preconditioner):
    """``sghmc_update`` restricted to some leaves of a large tree: ``grad_sel`` / ``mom_sel`` / ``v_sel`` are dicts
    {leaf position in flatten order: array}; the key schedule is that of the WHOLE tree (``split(key, n_leaves)``,
    fortuna/utils/random.py:11-14), the arithmetic per leaf is sghmc_integrator.py:59-92 unchanged.  Lets a parity
    test step a 110 M-parameter tree on the GPU and check chosen leaves without drawing 110 M normals in numpy.
    Returns (updates_sel, momentum_sel, v_sel, new_key)."""
    step_size = F32(step_schedule(count))
    v = preconditioner.update(grad_sel, v_sel)
    keys = prng.split(np.asarray(rng_key, dtype=np.uint32))
    key, new_key = keys[0], keys[1]
    lk = leaf_keys(key, n_leaves)
    noise = {i: prng.normal(lk[i], g.size).reshape(g.shape) for i, g in grad_sel.items()}
    noise = preconditioner.m_sqrt(noise, v)
    beta = F32(momentum_decay)
    sqrt_eps = np.sqrt(step_size, dtype=F32)
    noise_scale = np.sqrt(F32(2 * (1 - momentum_decay)), dtype=F32)
    momentum = tree_map(lambda m, g, n: beta * m + g * sqrt_eps + n * noise_scale, mom_sel, grad_sel, noise)
    updates = preconditioner.m_inv(momentum, v)
    updates = tree_map(lambda m: m * sqrt_eps, updates)
    return updates, momentum, v, new_key


def clip_gradients_by_norm(grad, max_grad_norm):
    """fortuna/utils/training.py:14-20 (``clip_grandients_by_norm``): norm = sqrt(tree_reduce(x + sum(y**2))) over the
    leaves in flatten order, mult = minimum(1, max_grad_norm / (1e-7 + norm)), grad <- mult * grad, all float32.
    (XLA's and numpy's float32 ``sum`` associate differently - pairwise blocks either way - so the norm agrees to
    rounding, not bits; the GPU kernel sums in float64.)"""
    acc = F32(0)
    for _, leaf in tree_flatten(grad):
        acc = F32(acc + np.sum(leaf.astype(F32) ** 2, dtype=F32))
    norm = np.sqrt(acc, dtype=F32)
    mult = np.minimum(F32(1), F32(max_grad_norm) / (F32(1e-7) + norm)).astype(F32)
    return tree_map(lambda z: (mult * z).astype(F32), grad), mult


def round_to_bf16(x):
    """float32 -> nearest bfloat16 (ties to even) -> float32: what a bf16 gradient buffer holds."""
    u = np.ascontiguousarray(x, dtype=F32).view(np.uint32)
    r = ((u >> 16) & 1) + np.uint32(0x7FFF)
    return ((u + r) & np.uint32(0xFFFF0000)).view(F32)


def apply_updates(params, updates):
    """optax.apply_updates as used by flax TrainState.apply_gradients: p + u."""
    return tree_map(lambda p, u: (p + u).astype(p.dtype), params, updates)


# ----------------------------------------------------------------------------- cyclical SGLD
def cyclical_sgld_is_sampling_phase(count, cycle_length, exploration_ratio):
    """cyclical_sgld_integrator.py:94-96: ((count % L) / L) >= ratio with int32 mod and a
    float32 true divide (jax default dtype), compared with the weak-typed python ratio."""
    return bool(F32(np.int32(count) % np.int32(cycle_length)) / F32(cycle_length) >= F32(exploration_ratio))


def cyclical_sgld_update(grad, state, init_step_size, cycle_length, exploration_ratio, preconditioner):
    """cyclical_sgld_integrator.update_fn (:64-100); ``state`` is the inner sgld (SGHMC) state.
    Returns (updates, new_state)."""
    schedule = cyclical_cosine_schedule_with_const_burnin(init_step_size, 0, cycle_length)
    if cyclical_sgld_is_sampling_phase(state.count, cycle_length, exploration_ratio):
        return sghmc_update(grad, state, 0.0, schedule, preconditioner, None)
    # exploration: preconditioned SGD ascent step, no noise, key and momentum untouched (:65-84)
    step_size = F32(schedule(state.count))
    v = preconditioner.update(grad, state.precond)
    rescaled = tree_map(lambda g: F32(-1.0) * step_size * g, grad)
    updates = tree_map(lambda g: F32(-1.0) * g, rescaled)  # optax.sgd(1.0) == scale(-1.0)
    updates = preconditioner.m_inv(updates, v)
    return updates, SGHMCState(state.count + 1, state.rng_key, state.momentum, v)


# ----------------------------------------------------------------------------- which steps are stored
def sghmc_do_sample(current_step, samples_count, n_samples, n_thinning, burnin_length):
    """sghmc_callback.py:49-53 (current_step is 1-based)."""
    return (
        samples_count < n_samples
        and current_step > burnin_length
        and (current_step - burnin_length) % n_thinning == 0
    )


def cyclical_sgld_do_sample(current_step, samples_count, n_samples, n_thinning, cycle_length, exploration_ratio):
    """cyclical_sgld_callback.py:55-59."""
    return (
        samples_count < n_samples
        and ((current_step % cycle_length) / cycle_length) >= exploration_ratio
        and (current_step % cycle_length) % n_thinning == 0
    )


def count_collectable(do_sample, n_epochs, n_training_steps):
    return sum(bool(do_sample(step, 0)) for step in range(1, n_epochs * n_training_steps + 1))
