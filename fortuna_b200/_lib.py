"""ctypes binding of libfortuna_b200.so (the C ABI in include/fortuna_b200.h).

The library is the product: there is no CPU or PyTorch fallback.  Importing this module never builds
or loads anything; the first call to :func:`lib` loads the shared object and raises
``FortunaB200LibraryError`` if it is missing (run ``python -m fortuna_b200.build`` or
``__graft_entry__.build()``).
"""

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FB200_LIB: another build of the same library (A/B comparisons of a kernel variant, scripts/ab_lib.sh)
LIB_PATH = os.environ.get("FB200_LIB") or os.path.join(_HERE, "lib", "libfortuna_b200.so")

F32, BF16 = 0, 1
W_SDC, W_SCD = 0, 1
TILE_PAIRS = 1024


class FortunaB200LibraryError(RuntimeError):
    pass


class FortunaB200Error(RuntimeError):
    def __init__(self, code, where):
        self.code = code
        super().__init__(f"{where} failed with code {code}: {error_string(code)}")


class Leaf(C.Structure):
    _fields_ = [("offset", C.c_int64), ("len", C.c_int64)]


class Tile(C.Structure):
    _fields_ = [("leaf", C.c_int32), ("pair_start", C.c_uint32)]


class SgmcmcHParams(C.Structure):
    _fields_ = [
        ("step_size", C.c_float),
        ("momentum_decay", C.c_double),
        ("zero_momentum", C.c_int32),
        ("rmsprop", C.c_int32),
        ("rms_alpha", C.c_double),
        ("rms_eps", C.c_double),
        ("grad_mult", C.c_void_p),
        ("grad_dtype", C.c_int32),
    ]


class JointGrad(C.Structure):
    _fields_ = [("lik_scale", C.c_float), ("prior_prec", C.c_float)]


_P = C.c_void_p

MAX_PEERS = 8


class DpPeers(C.Structure):
    _fields_ = [("n_ranks", C.c_int), ("rank", C.c_int), ("grad", C.c_void_p * MAX_PEERS), ("params", C.c_void_p * MAX_PEERS),
                ("mc_grad", C.c_void_p), ("mc_params", C.c_void_p)]


class ClsOutputs(C.Structure):
    _fields_ = [
        (n, _P)
        for n in (
            "mean",
            "aleatoric_variance",
            "epistemic_variance",
            "variance",
            "std",
            "mode",
            "entropy",
            "aleatoric_entropy",
            "epistemic_entropy",
            "log_prob",
            "ensemble_log_prob",
            "confidence",
            "brier_terms",
        )
    ]


class RegOutputs(C.Structure):
    _fields_ = [
        (n, _P)
        for n in (
            "mean",
            "aleatoric_variance",
            "epistemic_variance",
            "variance",
            "std",
            "log_prob",
            "ensemble_log_prob",
        )
    ]


_i, _i64, _u32, _f, _sz = C.c_int, C.c_int64, C.c_uint32, C.c_float, C.c_size_t

# name -> (restype, argtypes); mirrors include/fortuna_b200.h one to one
SIGNATURES = {
    "fb200_abi_version": (_i, []),
    "fb200_error_string": (C.c_char_p, [_i]),
    "fb200_threefry_split": (_i, [_P, _i, _P, _P]),
    "fb200_threefry_fold_in": (_i, [_P, _u32, _P, _P]),
    "fb200_random_bits": (_i, [_P, _i64, _P, _P]),
    "fb200_random_normal": (_i, [_P, _i64, _P, _P]),
    "fb200_sample_indices": (_i, [_P, _i, _i, _P, _P]),
    "fb200_choice_from_keys": (_i, [_P, _i, _i, _P, _P]),
    "fb200_sgmcmc_num_tiles_host": (_i64, [C.POINTER(Leaf), _i]),
    "fb200_sgmcmc_build_tiles_host": (_i, [C.POINTER(Leaf), _i, C.POINTER(Tile), _i64]),
    "fb200_sgmcmc_step_f32": (
        _i,
        [_P, _P, _P, _P, _P, _i64, _P, _i, _P, _i64, _P, _P, _P, C.POINTER(SgmcmcHParams), _P],
    ),
    "fb200_grad_clip_workspace_doubles": (_i64, [_i64]),
    "fb200_grad_clip_mult": (_i, [_P, _i, _P, C.POINTER(JointGrad), _i64, _f, _P, _P, _P]),
    "fb200_sgd_explore_step_f32": (_i, [_P, _P, _P, _P, _i64, C.POINTER(SgmcmcHParams), _P]),
    "fb200_sgmcmc_step_dp_f32": (_i, [_P, _P, _i64, _P, _i, _P, _i64, _i64, _P, _P, _P, C.POINTER(SgmcmcHParams),
                                 C.POINTER(JointGrad), C.POINTER(DpPeers), _P]),
    "fb200_sgd_explore_step_dp_f32": (_i, [_P, _i64, _i64, _i64, C.POINTER(SgmcmcHParams), C.POINTER(JointGrad),
                                      C.POINTER(DpPeers), _P]),
    "fb200_sgmcmc_step_joint_f32": (_i, [_P, _P, _P, _P, _i64, _P, _i, _P, _i64, _P, _P, _P, C.POINTER(SgmcmcHParams),
                                         C.POINTER(JointGrad), _P, _P, _P]),
    "fb200_sgd_explore_joint_workspace_doubles": (_i64, [_i64]),
    "fb200_sgd_explore_step_joint_f32": (_i, [_P, _P, _P, _i64, C.POINTER(SgmcmcHParams), C.POINTER(JointGrad), _P, _P, _P]),
    "fb200_bank_put_f32": (_i, [_P, _i64, _i, _P, _P]),
    "fb200_last_layer_logits": (_i, [_P, _P, _P, _P, _P, _i, _P, _i, _i, _i, _i, _i, _i, _i, _i, _P]),
    "fb200_last_layer_ex_workspace_bytes": (_sz, [_i, _i, _i]),
    "fb200_last_layer_logits_ex": (_i, [_P, _P, _P, _P, _P, _i, _P, _i, _i, _i, _i, _i, _i, _i, _i, _P, _P, _sz, _P]),
    "fb200_softmax_rows": (_i, [_P, _i, _i64, _i, _i, _P, _P, _P]),
    "fb200_log_prob_from_lse": (_i, [_P, _i, _P, _P, _i, _i, _i, _P, _P, _P]),
    "fb200_transpose_w_bf16": (_i, [_P, _i, _P, _i, _i, _i, _P]),
    "fb200_bma_reduce_cls": (_i, [_P, _i, _P, _P, _i, _i, _i, C.POINTER(ClsOutputs), _P]),
    "fb200_bma_reduce_cls_wide_workspace_bytes": (_sz, [_i, _i]),
    "fb200_bma_reduce_cls_wide": (_i, [_P, _i, _P, _P, _i, _i, _i, C.POINTER(ClsOutputs), _P, _sz, _P]),
    "fb200_cls_slot_row_floats": (_sz, [_i]),
    "fb200_bma_reduce_cls_shard": (_i, [_P, _i, _P, _P, _i, _i, _i, _i, _i, _P, _P, _P]),
    "fb200_bma_finalize_cls_slots": (_i, [_P, _i, _i, _i, _i, C.POINTER(ClsOutputs), _P, _P]),
    "fb200_variance_combine": (_i, [_P, _P, _i64, _P, _P, _P]),
    "fb200_bma_reduce_reg": (_i, [_P, _P, _P, _i, _i, _i, C.POINTER(RegOutputs), _P]),
    "fb200_aps_scores": (_i, [_P, _P, _i, _i, _P, _P, _P]),
    "fb200_simple_scores": (_i, [_P, _P, _i, _i, _P, _P]),
    "fb200_credible_sets": (_i, [_P, _i, _i, _f, _P, _P, _P]),
    "fb200_sort_workspace_bytes": (_sz, [_i64]),
    "fb200_sort_f32": (_i, [_P, _i64, _P, _sz, _P]),
    "fb200_aps_conformal_sets": (_i, [_P, _i, _i, _P, _i, _f, _P, _P, _P, _P]),
    "fb200_conformal_sets": (_i, [_P, _i, _P, _i, _i, _f, _P, _P, _P]),
    "fb200_ece_bins": (_i, [_P, _i, _i, _P, _P, _P, _i, _P, _P, _P, _P, _P, _P]),
    "fb200_ece_bins_from_rows": (_i, [_P, _P, _P, _P, _i, _P, _i, _P, _P, _P, _P, _P, _P]),
    "fb200_ece_from_bins": (_i, [_P, _P, _P, _i, _i64, _P, _P]),
    "fb200_ece_pack_bins": (_i, [_P, _P, _P, _i, _P, _i, _P]),
    "fb200_ece_from_packed_bins": (_i, [_P, _i, _i64, _P, _P, _P]),
    "fb200_quantile_axis0": (_i, [_P, _i, _i64, _P, _i, _P, _P]),
    "fb200_calib_cls_workspace_bytes": (_sz, [_i, _i]),
    "fb200_calib_cls_loss_terms": (_i, [_P, _i, _P, _P, _i, _i, _i, _P, _sz, _P, _P]),
    "fb200_calib_reg_workspace_bytes": (_sz, [_i, _i]),
    "fb200_calib_reg_loss_terms": (_i, [_P, _P, _P, _i, _i, _i, _P, _sz, _P, _P]),
    "fb200_cls_sample_targets": (_i, [_P, _P, _P, _i, _i, _i, _i, _P, _P, _P]),
    "fb200_maxcov_workspace_bytes": (_sz, [_i]),
    "fb200_maxcov_counts": (_i, [_P, _P, _i64, _P, _i, _i, _f, _P, _sz, _P, _P, _P]),
    "fb200_maxcov_apply_patches": (_i, [_P, _i64, _f, _f, _P, _P]),
    "fb200_cls_output_sample_targets": (_i, [_P, _P, _P, _i, _i, _i, _P, _P]),
    "fb200_reg_output_sample_targets": (_i, [_P, _P, _P, _i, _i, _i, _P, _P]),
    "fb200_output_calib_focal_loss_terms": (_i, [_P, _i, _P, _P, _i, _i, _f, _P, _sz, _P, _P]),
    "fb200_output_calib_scaled_mse_terms": (_i, [_P, _P, _P, _i, _i, _P, _sz, _P, _P]),
    "fb200_focal_loss_grad": (_i, [_P, _P, _i, _i, _f, _P, _P, _P, _P]),
    "fb200_scaled_mse_loss_grad": (_i, [_P, _P, _i, _i, _P, _P, _P, _P]),
    "fb200_dot_f32": (_i, [_P, _P, _i64, _P, _P]),
    "fb200_adam_step_f32": (_i, [_P, _P, _P, _P, _i64, _f, _f, _f, _f, _f, _f, _P]),
    "fb200_multivalid_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "fb200_multivalid_stats": (_i, [_P, _P, _P, _P, _i, _i, _i, _i, _i, _i, _i, _i, _P, _sz, _P, _P]),
    "fb200_multivalid_patch": (_i, [_P, _P, _i, _i, _i, _i, _i, _i, _f, _i, _i, _f, _f, _i, _P, _P]),
    "fb200_multivalid_loss": (_i, [_P, _P, _i, _i, _i, _i, _f, _P, _P, _P]),
    "fb200_batchmvp_delta_counts": (_i, [_P, _P, _P, _i, _i, _i, _i, _f, _i, _P, _i, _P, _P]),
    "fb200_reg_mc_entropies": (_i, [_P, _P, _P, _P, _i, _i, _i, _i, _i, _P, _P, _P, _P]),
    "fb200_conformal_bayes_workspace_bytes": (_sz, [_i, _i, _i]),
    "fb200_conformal_bayes_sets": (_i, [_P, _P, _i, _i, _i, _i, _f, _P, _sz, _P, _P, _P, _P, _P]),
    "fb200_ksd_imq_workspace_bytes": (_sz, [_i, _i64]),
    "fb200_ksd_imq": (_i, [_P, _P, _i, _i64, _i64, _f, _f, _P, _sz, _P, _P]),
    "fb200_ess": (_i, [_P, _i, _i64, _i64, _i, _f, _P, _P]),
    "fb200_reg_sample_targets": (_i, [_P, _P, _P, _i, _i, _i, _i, _P, _P, _P]),
    "fb200_quantile_sorted": (_i, [_P, _i64, _P, _i, _P, _P]),
}

_LIB = None


def lib():
    """Load (once) and return the ctypes handle; raises if the CUDA library has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise FortunaB200LibraryError(
                f"{LIB_PATH} is missing. fortuna_b200 has no CPU fallback: build the CUDA library with "
                "`python -m fortuna_b200.build` (needs nvcc) before using it."
            )
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _LIB = handle
    return _LIB


def error_string(code):
    return lib().fb200_error_string(int(code)).decode()


def check(code, where):
    if code != 0:
        raise FortunaB200Error(code, where)
