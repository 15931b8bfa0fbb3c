"""Cyclical SGLD integrator - drop-in for
fortuna/prob_model/posterior/sgmcmc/cyclical_sgld/cyclical_sgld_integrator.py:14-102.

Phase test ``((count % cycle_length) / cycle_length) >= exploration_ratio`` (:94-96) is evaluated on the
host from the mirrored count in float32, as jax does (int32 mod, float32 true divide, weak-typed ratio):
  * sampling phase   -> the SGHMC kernel with momentum_decay = 0 (SGLD);
  * exploration phase -> fb200_sgd_explore_step_f32 (preconditioned SGD ascent, no noise, key and
    momentum untouched).
"""

from typing import NamedTuple

import numpy as np

from ..... import ops
from .....utils.flat_tree import FlatTree, as_flat_tree
from ..sghmc.sghmc_integrator import GradientTransformation, OptaxSGHMCState, clip_multiplier, sghmc_integrator
from ..sgmcmc_preconditioner import Preconditioner, precond_buffer
from ..sgmcmc_step_schedule import cyclical_cosine_schedule_with_const_burnin


class OptaxCyclicalSGLDState(NamedTuple):
    """Same fields as the reference (:14-18); optax ``sgd(1.0)`` carries no arrays."""

    sgd_state: tuple
    sgld_state: OptaxSGHMCState


def is_sampling_phase(count: int, cycle_length: int, exploration_ratio: float) -> bool:
    return bool(
        np.float32(np.int32(count) % np.int32(cycle_length)) / np.float32(cycle_length)
        >= np.float32(exploration_ratio)
    )


def cyclical_sgld_integrator(
    rng_key,
    init_step_size: float,
    cycle_length: int,
    exploration_ratio: float,
    preconditioner: Preconditioner,
) -> GradientTransformation:
    step_schedule = cyclical_cosine_schedule_with_const_burnin(
        init_step_size=init_step_size, burnin_steps=0, cycle_length=cycle_length
    )
    sgld = sghmc_integrator(
        momentum_decay=0.0,
        momentum_resample_steps=None,
        rng_key=rng_key,
        step_schedule=step_schedule,
        preconditioner=preconditioner,
    )

    def init_fn(params):
        return OptaxCyclicalSGLDState(sgd_state=((), ()), sgld_state=sgld.init(params))

    ws = {}

    def _explore(gradient, state, params, joint=None, dp=None, max_grad_norm=None):
        inner = state.sgld_state
        grad = as_flat_tree(gradient, like=inner.momentum)
        hp = ops.make_hparams(
            step_schedule(inner.count), 0.0, preconditioner.is_rmsprop, preconditioner.running_average_factor,
            preconditioner.eps, grad_mult=clip_multiplier(grad, params, joint, max_grad_norm, dp, ws),
        )
        v = precond_buffer(inner.preconditioner_state)
        updates = grad.empty_like() if params is None else None
        if dp is not None:
            if params is None or params.flat.data_ptr() != dp.params.data_ptr() or grad.flat.data_ptr() != dp.grad.data_ptr():
                raise ValueError("the data-parallel step works in place on the peer-mapped parameter / gradient vectors")
            dp.explore_step(v.flat if v is not None else None, hp, joint=None if joint is None else joint[:2])
            return None, OptaxCyclicalSGLDState(sgd_state=state.sgd_state, sgld_state=inner._replace(count=inner.count + 1))
        ops.sgd_explore_step(
            params.flat if params is not None else None,
            grad.flat,
            v.flat if v is not None else None,
            updates.flat if updates is not None else None,
            hp,
            joint=joint,
            ws_cache=ws,
        )
        new_inner = inner._replace(count=inner.count + 1)
        return updates, OptaxCyclicalSGLDState(sgd_state=state.sgd_state, sgld_state=new_inner)

    def update_fn(gradient, state, *_, max_grad_norm=None):
        if is_sampling_phase(state.sgld_state.count, cycle_length, exploration_ratio):
            updates, inner = sgld.update(gradient, state.sgld_state, max_grad_norm=max_grad_norm)
            return updates, OptaxCyclicalSGLDState(sgd_state=state.sgd_state, sgld_state=inner)
        return _explore(gradient, state, None, max_grad_norm=max_grad_norm)

    def _sync_precond(state, dp, scheme):
        """Data-parallel RMSprop: the two phases split the elements differently over the ranks and the moment estimates
        are rank-local, so on a phase change the pieces advanced under the previous split are gathered first - the new
        owner of an element must see every update its estimate has had (the reference's replicas all hold the whole
        state, fortuna/training/trainer.py:807)."""
        v = precond_buffer(state.sgld_state.preconditioner_state)
        last = ws.get("v_scheme")
        if dp is not None and v is not None and last is not None and last != scheme:
            dp.gather_owned(v.flat, last, v.lens, v.offsets)
        ws["v_scheme"] = scheme if dp is not None else None

    def apply_fn(params: FlatTree, gradient, state, joint=None, dp=None, max_grad_norm=None):
        sampling = is_sampling_phase(state.sgld_state.count, cycle_length, exploration_ratio)
        _sync_precond(state, dp, "tiles" if sampling else "elems")
        if sampling:
            _, inner = sgld.apply(params, gradient, state.sgld_state, joint=joint, dp=dp, max_grad_norm=max_grad_norm)
            return params, OptaxCyclicalSGLDState(sgd_state=state.sgd_state, sgld_state=inner)
        _, new_state = _explore(gradient, state, params, joint=joint, dp=dp, max_grad_norm=max_grad_norm)
        return params, new_state

    def dp_gather(state, dp):
        inner = state.sgld_state
        m = inner.momentum
        dp.gather_owned(m.flat, "tiles", m.lens, m.offsets)  # only the sampling phase advances the momentum
        v = precond_buffer(inner.preconditioner_state)
        if v is not None and ws.get("v_scheme") is not None:
            dp.gather_owned(v.flat, ws["v_scheme"], v.lens, v.offsets)
            ws["v_scheme"] = None  # whole on every rank: the next phase may split it either way

    return GradientTransformation(init_fn, update_fn, apply_fn, dp_gather)
