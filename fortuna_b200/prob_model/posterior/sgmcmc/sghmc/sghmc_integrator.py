"""SGHMC integrator - drop-in for
fortuna/prob_model/posterior/sgmcmc/sghmc/sghmc_integrator.py:20-94 (an optax ``GradientTransformation``).

``init_fn(params) -> OptaxSGHMCState(count, rng_key, momentum, preconditioner_state)`` and
``update_fn(gradient, state, *_) -> (updates, new_state)`` keep the reference's names and meaning; the
arithmetic of one step (key split, per-leaf threefry noise, preconditioner, momentum, update) is ONE
launch of fb200_sgmcmc_step_f32.  Two entry points:

  * ``update``  - contract-exact optax mode: returns the ``updates`` tree, caller adds them;
  * ``apply``   - fused ``TrainState.apply_gradients`` (the reference's call site,
                  fortuna/training/trainer.py:170): params += updates inside the same launch, no
                  ``updates`` round trip through HBM (20 instead of 36 B/param).

Buffers are donated like XLA donates under jit: momentum / preconditioner moments are updated in place
and the returned state aliases them; ``rng_key`` is a fresh 2-word tensor per step so old states keep a
valid key.  ``count`` is mirrored on the host (it is deterministic), so no step ever synchronises.
"""

from typing import Callable, NamedTuple, Optional

import numpy as np
import torch

from ..... import ops
from .....utils.flat_tree import FlatTree, as_flat_tree
from ..sgmcmc_preconditioner import Preconditioner, precond_buffer
from ..sgmcmc_step_schedule import StepSchedule


class OptaxSGHMCState(NamedTuple):
    """Optax state for the SGHMC integrator (same fields as the reference, :20-26)."""

    count: int
    rng_key: torch.Tensor  # uint32[2] bit pattern on the device
    momentum: FlatTree
    preconditioner_state: NamedTuple


class GradientTransformation(NamedTuple):
    init: Callable
    update: Callable
    apply: Callable  # fused update + apply_updates
    # data-parallel only: dp_gather(opt_state, dp) makes the rank-local momentum / preconditioner moments whole on every
    # rank (each rank advances only the elements it owns), e.g. before the chain state is written to a checkpoint
    dp_gather: Optional[Callable] = None


def clip_multiplier(grad: FlatTree, params, joint, max_grad_norm, dp, ws):
    """Device float32[1] multiplier of ``clip_grandients_by_norm`` (fortuna/utils/training.py:14-20, call site
    fortuna/training/trainer.py:158-169) for the gradient this step uses, or None when clipping is off."""
    if not max_grad_norm or max_grad_norm <= 0:
        return None
    if dp is not None:
        raise NotImplementedError("gradient clipping needs the norm of the AVERAGED gradient: not built for the data-parallel step")
    return ops.grad_clip_mult(grad.flat, max_grad_norm, params=params.flat if joint is not None else None,
                              joint=None if joint is None else joint[:2], ws_cache=ws)


def _step(gradient, state, params, momentum_decay, momentum_resample_steps, step_schedule, preconditioner, ws,
          joint=None, dp=None, max_grad_norm=None):
    grad = as_flat_tree(gradient, like=state.momentum)
    step_size = np.float32(step_schedule(state.count))
    zero_m = momentum_resample_steps is not None and state.count % momentum_resample_steps == 0
    gm = clip_multiplier(grad, params, joint, max_grad_norm, dp, ws)
    hp = ops.make_hparams(
        step_size,
        momentum_decay,
        preconditioner.is_rmsprop,
        preconditioner.running_average_factor,
        preconditioner.eps,
        zero_m,
        grad_mult=gm,
        grad_dtype=grad.flat.dtype,
    )
    leaves, tiles = state.momentum.tables()
    key = ws.get("leaf_keys")
    if key is None or key.shape[0] != leaves.shape[0] or key.device != grad.flat.device:
        key = ws["leaf_keys"] = torch.empty((leaves.shape[0], 2), dtype=ops.KEY_DTYPE, device=grad.flat.device)
    new_key = torch.empty(2, dtype=ops.KEY_DTYPE, device=grad.flat.device)
    v = precond_buffer(state.preconditioner_state)
    updates = None
    if params is None:
        updates = grad.empty_like()
    if dp is not None:
        # data-parallel: gradient average over the ranks + step on the owned tiles + parameter broadcast, one kernel
        if params is None or params.flat.data_ptr() != dp.params.data_ptr() or grad.flat.data_ptr() != dp.grad.data_ptr():
            raise ValueError("the data-parallel step works in place on the peer-mapped parameter / gradient vectors")
        dp.step(state.momentum.flat, v.flat if v is not None else None, leaves, tiles, state.rng_key, new_key, key, hp,
                joint=None if joint is None else joint[:2])
        return None, OptaxSGHMCState(count=state.count + 1, rng_key=new_key, momentum=state.momentum,
                                     preconditioner_state=state.preconditioner_state)
    ops.sgmcmc_step(
        params.flat if params is not None else None,
        grad.flat,
        state.momentum.flat,
        v.flat if v is not None else None,
        updates.flat if updates is not None else None,
        leaves,
        tiles,
        state.rng_key,
        new_key,
        key,
        hp,
        joint=joint,
        ws_cache=ws,
    )
    new_state = OptaxSGHMCState(
        count=state.count + 1,
        rng_key=new_key,
        momentum=state.momentum,
        preconditioner_state=state.preconditioner_state,
    )
    return updates, new_state


def sghmc_integrator(
    momentum_decay: float,
    momentum_resample_steps: Optional[int],
    rng_key: torch.Tensor,
    step_schedule: StepSchedule,
    preconditioner: Preconditioner,
) -> GradientTransformation:
    """Same signature as the reference factory (:29-49)."""
    ws = {}

    def init_fn(params):
        params = as_flat_tree(params)
        return OptaxSGHMCState(
            count=0,
            rng_key=rng_key,
            momentum=params.zeros_like(),
            preconditioner_state=preconditioner.init(params),
        )

    def update_fn(gradient, state, *_, max_grad_norm=None):
        return _step(
            gradient, state, None, momentum_decay, momentum_resample_steps, step_schedule, preconditioner, ws,
            max_grad_norm=max_grad_norm,
        )

    def apply_fn(params: FlatTree, gradient, state, joint=None, dp=None, max_grad_norm=None):
        """``joint=(n_data / batch_size, prior_precision, float64[1] tensor | None)``: ``gradient`` is the gradient of
        the batch log-likelihood sum; scaling and the Gaussian-prior gradient happen inside the kernel.
        ``max_grad_norm`` > 0: the gradient is clipped by its global norm (one extra read; scaled inside the kernel).
        ``gradient`` may be a FlatTree whose flat vector is bfloat16 (18 instead of 20 B/param)."""
        if not isinstance(params, FlatTree):
            raise TypeError("the fused apply path updates a FlatTree in place")
        _, new_state = _step(
            gradient, state, params, momentum_decay, momentum_resample_steps, step_schedule, preconditioner, ws,
            joint=joint, dp=dp, max_grad_norm=max_grad_norm,
        )
        return params, new_state

    def dp_gather(state, dp):
        m = state.momentum
        dp.gather_owned(m.flat, "tiles", m.lens, m.offsets)
        v = precond_buffer(state.preconditioner_state)
        if v is not None:
            dp.gather_owned(v.flat, "tiles", v.lens, v.offsets)

    return GradientTransformation(init_fn, update_fn, apply_fn, dp_gather)
