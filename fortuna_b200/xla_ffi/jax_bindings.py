"""Reference-side binding: registers the XLA-FFI handlers of fb200_xla_ffi.cc with JAX and exposes the
drop-ins a Fortuna maintainer would import (one worked example, the SGHMC integrator; the other handlers are bound the
same way - INTEGRATION.md lists all of them with the reference call each replaces).  Requires jax/jaxlib (absent from the build container and the
GPU box, so this module is NOT imported by the package, the tests or the benchmark - the ctypes driver in
fortuna_b200/ops.py is what runs here).  Tolerates both spellings of the FFI module: ``jax.extend.ffi`` at the
pinned jax 0.4.30, ``jax.ffi`` from 0.4.38.

    from fortuna_b200.xla_ffi.jax_bindings import sghmc_integrator   # instead of fortuna's
"""
import ctypes
import os


def _ffi():
    import jax

    try:
        from jax import ffi  # jax >= 0.4.38

        return ffi
    except ImportError:
        from jax.extend import ffi  # jax 0.4.30 .. 0.4.37

        return ffi


_TARGETS = (
    "fb200_xla_split",
    "fb200_xla_fold_in",
    "fb200_xla_random_bits",
    "fb200_xla_random_normal",
    "fb200_xla_sample_indices",
    "fb200_xla_choice_from_keys",
    "fb200_xla_sgmcmc_step",
    "fb200_xla_sgmcmc_step_joint",
    "fb200_xla_sgd_explore_step",
    "fb200_xla_sgd_explore_step_joint",
    "fb200_xla_bank_put",
    "fb200_xla_last_layer",
    "fb200_xla_last_layer_ex",
    "fb200_xla_log_prob_from_lse",
    "fb200_xla_transpose_w",
    "fb200_xla_softmax_rows",
    "fb200_xla_bma_reduce_cls",
    "fb200_xla_bma_reduce_cls_shard",
    "fb200_xla_bma_finalize_cls_slots",
    "fb200_xla_variance_combine",
    "fb200_xla_bma_reduce_reg",
    "fb200_xla_aps_scores",
    "fb200_xla_simple_scores",
    "fb200_xla_conformal_sets",
    "fb200_xla_aps_conformal_sets",
    "fb200_xla_credible_sets",
    "fb200_xla_sort_f32",
    "fb200_xla_quantile_sorted",
    "fb200_xla_quantile_axis0",
    "fb200_xla_ece_bins",
    "fb200_xla_ece_bins_from_rows",
    "fb200_xla_ece_pack_bins",
    "fb200_xla_ece_from_packed_bins",
    "fb200_xla_cls_sample_targets",
    "fb200_xla_ksd_imq",
    "fb200_xla_ess",
    "fb200_xla_grad_clip_mult",
    "fb200_xla_reg_sample_targets",
    "fb200_xla_reg_mc_entropies",
    "fb200_xla_cls_output_sample_targets",
    "fb200_xla_reg_output_sample_targets",
    "fb200_xla_conformal_bayes_sets",
    "fb200_xla_calib_cls_loss_terms",
    "fb200_xla_calib_reg_loss_terms",
    "fb200_xla_focal_loss_grad",
    "fb200_xla_scaled_mse_loss_grad",
    "fb200_xla_adam_step",
    "fb200_xla_maxcov_counts",
    "fb200_xla_maxcov_apply_patches",
)
_registered = False


def register(lib_path=None):
    global _registered
    if _registered:
        return
    ffi = _ffi()
    here = os.path.dirname(os.path.abspath(__file__))
    lib = ctypes.CDLL(lib_path or os.path.join(here, "libfb200_xla_ffi.so"))
    for name in _TARGETS:
        ffi.register_ffi_target(name, ffi.pycapsule(getattr(lib, name)), platform="CUDA")
    _registered = True


def sghmc_integrator(momentum_decay, momentum_resample_steps, rng_key, step_schedule, preconditioner_kind="identity",
                     running_average_factor=0.99, eps=1e-7):
    """optax.GradientTransformation with Fortuna's (init_fn, update_fn) contract whose update is ONE custom call.
    Works on the flattened pytree: `flatten(params) -> (flat f32[N], leaves i64[L,2], tiles i32[T,2])` is built
    once with fb200_sgmcmc_build_tiles_host; updates = new_params - params keeps optax's return contract while
    `apply_gradients` users can take new_params directly (input_output_aliases donates the buffers)."""
    import jax
    import jax.numpy as jnp
    import numpy as np
    import optax
    from jax.flatten_util import ravel_pytree

    register()
    ffi = _ffi()
    rms = int(preconditioner_kind == "rmsprop")

    def init_fn(params):
        flat, _ = ravel_pytree(params)
        return dict(count=jnp.zeros([], jnp.int32), rng_key=rng_key, momentum=jnp.zeros_like(flat),
                    v=jnp.zeros_like(flat))

    def update_fn(gradient, state, params):
        gflat, unravel = ravel_pytree(gradient)
        pflat, _ = ravel_pytree(params)
        lens = [int(np.prod(l.shape)) for l in jax.tree_util.tree_leaves(gradient)]
        offs = np.concatenate([[0], np.cumsum(lens)[:-1]])
        leaves = np.stack([offs, lens], 1).astype(np.int64)
        tiles = np.array([(i, p) for i, n in enumerate(lens) for p in range(0, (n + 1) // 2, 1024)], np.int32)
        n = gflat.shape[0]
        out_types = (jax.ShapeDtypeStruct((n,), jnp.float32),) * 3 + (
            jax.ShapeDtypeStruct((2,), jnp.uint32), jax.ShapeDtypeStruct((len(lens), 2), jnp.uint32))
        new_p, new_m, new_v, new_key, _ = ffi.ffi_call(
            "fb200_xla_sgmcmc_step", out_types, pflat, gflat, state["momentum"], state["v"], jnp.asarray(leaves),
            jnp.asarray(tiles), state["rng_key"], step_size=np.float32(step_schedule(state["count"])),
            momentum_decay=float(momentum_decay), zero_momentum=np.int32(
                momentum_resample_steps is not None and int(state["count"]) % momentum_resample_steps == 0),
            rmsprop=np.int32(rms), rms_alpha=float(running_average_factor),
            rms_eps=float(eps), input_output_aliases={0: 0, 2: 1, 3: 2})
        new_state = dict(count=state["count"] + 1, rng_key=new_key, momentum=new_m, v=new_v)
        return unravel(new_p - pflat), new_state

    return optax.GradientTransformation(init_fn, update_fn)
