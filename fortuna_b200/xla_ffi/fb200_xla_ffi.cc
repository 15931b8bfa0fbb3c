// XLA-FFI shim over the C ABI of include/fortuna_b200.h - the binding north_star names ("thin jax.ffi custom-call
// bindings").  One handler per hot-path entry point the Python mirror (fortuna_b200/ops.py) calls: PRNG, the sampler
// steps, the sample bank, the last-layer contraction, the single-pass reductions (plain, wide, shard, finalise), the scoring kernels and the ECE.  Every handler only unpacks buffers / attributes and forwards to ONE
// fb200_* call on XLA's stream; no arithmetic lives here.  Conventions:
//   * an OPTIONAL input (targets, log_temp, bias, sample_idx, ...) is passed as a zero-element buffer;
//   * an unwanted OUTPUT of the reduction kernels is a zero-element result buffer (XLA still needs a result slot);
//   * logits may be f32 or bf16 (AnyBuffer); scalars the reference passes as Python numbers are attributes;
//   * in-place state (params, momentum, preconditioner moments) uses input_output_aliases on the Python side, and a
//     copy first when XLA could not alias.
//
// NOT built by fortuna_b200/build.py: it needs the XLA FFI headers that ship inside jaxlib
// (jaxlib/include/xla/ffi/api/ffi.h), and jax / jaxlib cannot be installed in the build container or on the GPU box
// (SURVEY.md section 0).  What IS checked here: scripts/check_xla_ffi_syntax.sh compiles this file with
// `g++ -fsyntax-only` against mock/xla/ffi/api/ffi.h, whose Binding::To() static_asserts every handler against its
// binding - a syntax / signature check, not an integration test.  Where jaxlib is present:
//
//   JAXLIB_INC=$(python -c "import jaxlib,os;print(os.path.join(os.path.dirname(jaxlib.__file__),'include'))")
//   g++ -O2 -fPIC -shared -std=c++17 fb200_xla_ffi.cc -I$JAXLIB_INC -I../../include -I/usr/local/cuda/include
//       -L../lib -lfortuna_b200 -Wl,-rpath,'$ORIGIN/../lib' -o libfb200_xla_ffi.so
//
// and register the handlers from Python with jax_bindings.py (same directory).
#if __has_include("xla/ffi/api/ffi.h")
#include <cuda_runtime_api.h>

#include <cstdint>
#include <string>

#include "fortuna_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

ffi::Error Check(int rc, const char* what) {
  if (rc == 0) return ffi::Error::Success();
  return ffi::Error(rc < 0 ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal,
                    std::string(what) + ": " + fb200_error_string(rc));
}

template <typename B>
int Dim(const B& b, int i) {
  return static_cast<int>(b.dimensions()[i]);
}
// optional input / unwanted output: zero elements -> NULL
template <ffi::DataType dt>
ffi::NativeType<dt>* Opt(const ffi::Buffer<dt>& b) {
  return b.element_count() ? b.typed_data() : nullptr;
}
template <ffi::DataType dt>
ffi::NativeType<dt>* Opt(ffi::ResultBuffer<dt>& b) {
  return b->element_count() ? b->typed_data() : nullptr;
}
// f32 / bf16 tensors arrive as AnyBuffer
int DTypeOf(const ffi::AnyBuffer& b) {
  if (b.element_type() == ffi::F32) return FB200_F32;
  if (b.element_type() == ffi::BF16) return FB200_BF16;
  return -1;
}
ffi::Error BadDType(const char* what) {
  return ffi::Error(ffi::ErrorCode::kInvalidArgument, std::string(what) + ": expected a float32 or bfloat16 tensor");
}
void CopyIfNotAliased(const void* src, void* dst, size_t bytes, cudaStream_t s) {
  if (src != dst && bytes) cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s);
}

using F32 = ffi::Buffer<ffi::F32>;
using F64 = ffi::Buffer<ffi::F64>;
using S32 = ffi::Buffer<ffi::S32>;
using S64 = ffi::Buffer<ffi::S64>;
using U32 = ffi::Buffer<ffi::U32>;
using U64 = ffi::Buffer<ffi::U64>;
using U8 = ffi::Buffer<ffi::U8>;
using BF16 = ffi::Buffer<ffi::BF16>;
using Any = ffi::AnyBuffer;
using RF32 = ffi::ResultBuffer<ffi::F32>;
using RF64 = ffi::ResultBuffer<ffi::F64>;
using RS32 = ffi::ResultBuffer<ffi::S32>;
using RU32 = ffi::ResultBuffer<ffi::U32>;
using RU8 = ffi::ResultBuffer<ffi::U8>;
using RBF16 = ffi::ResultBuffer<ffi::BF16>;
using RAny = ffi::Result<ffi::AnyBuffer>;
using Stream = ffi::PlatformStream<cudaStream_t>;

// the reduction kernels' output struct from eleven result slots (zero elements = not wanted) + the three ECE row inputs
struct ClsOuts {
  RF32 &mean, &alea_var, &epi_var, &variance, &std_, &entropy, &alea_ent, &epi_ent, &log_prob, &ens_log_prob, &confidence,
      &brier;
  RS32& mode;
  fb200_cls_outputs_t get() {
    fb200_cls_outputs_t o{};
    o.mean = Opt(mean);
    o.aleatoric_variance = Opt(alea_var);
    o.epistemic_variance = Opt(epi_var);
    o.variance = Opt(variance);
    o.std = Opt(std_);
    o.mode = Opt(mode);
    o.entropy = Opt(entropy);
    o.aleatoric_entropy = Opt(alea_ent);
    o.epistemic_entropy = Opt(epi_ent);
    o.log_prob = Opt(log_prob);
    o.ensemble_log_prob = Opt(ens_log_prob);
    o.confidence = Opt(confidence);
    o.brier_terms = Opt(brier);
    return o;
  }
};
#define FB200_CLS_OUT_PARAMS                                                                                          \
  RF32 mean, RF32 alea_var, RF32 epi_var, RF32 variance, RF32 std_, RS32 mode, RF32 entropy, RF32 alea_ent, RF32 epi_ent, \
      RF32 log_prob, RF32 ens_log_prob, RF32 confidence, RF32 brier
#define FB200_CLS_OUT_STRUCT \
  ClsOuts{mean, alea_var, epi_var, variance, std_, entropy, alea_ent, epi_ent, log_prob, ens_log_prob, confidence, brier, mode}.get()
#define FB200_CLS_OUT_BINDING                                                                                       \
  .Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<S32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>() \
      .Ret<F32>().Ret<F32>()

}  // namespace

// ================================================================================================ PRNG (jax.random)
// jax.random.split(key, num) - fortuna/utils/random.py:11-14, 38-49
static ffi::Error Split(cudaStream_t s, U32 key, RU32 out) {
  return Check(fb200_threefry_split(key.typed_data(), Dim(*out, 0), out->typed_data(), s), "fb200_threefry_split");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_split, Split, ffi::Ffi::Bind().Ctx<Stream>().Arg<U32>().Ret<U32>());

// jax.random.fold_in(key, data) - fortuna/training/trainer.py:676-680
static ffi::Error FoldIn(cudaStream_t s, U32 key, RU32 out, int64_t data) {
  return Check(fb200_threefry_fold_in(key.typed_data(), static_cast<uint32_t>(data), out->typed_data(), s),
               "fb200_threefry_fold_in");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_fold_in, FoldIn, ffi::Ffi::Bind().Ctx<Stream>().Arg<U32>().Ret<U32>().Attr<int64_t>("data"));

static ffi::Error RandomBits(cudaStream_t s, U32 key, RU32 out) {
  return Check(fb200_random_bits(key.typed_data(), static_cast<int64_t>(out->element_count()), out->typed_data(), s),
               "fb200_random_bits");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_random_bits, RandomBits, ffi::Ffi::Bind().Ctx<Stream>().Arg<U32>().Ret<U32>());

// jax.random.normal(key, shape, float32) - fortuna/utils/random.py:17-23
static ffi::Error RandomNormal(cudaStream_t s, U32 key, RF32 out) {
  return Check(fb200_random_normal(key.typed_data(), static_cast<int64_t>(out->element_count()), out->typed_data(), s),
               "fb200_random_normal");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_random_normal, RandomNormal, ffi::Ffi::Bind().Ctx<Stream>().Arg<U32>().Ret<F32>());

// idx[s] = choice(split(key, S)[s], n_samples) - predictive/base.py:632 + sgmcmc_posterior.py:48-52
static ffi::Error SampleIndices(cudaStream_t s, U32 key, RS32 idx, int32_t n_samples) {
  return Check(fb200_sample_indices(key.typed_data(), Dim(*idx, 0), n_samples, idx->typed_data(), s), "fb200_sample_indices");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_sample_indices, SampleIndices,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<U32>().Ret<S32>().Attr<int32_t>("n_samples"));

// idx[i] = choice(keys[i], n_samples)
static ffi::Error ChoiceFromKeys(cudaStream_t s, U32 keys, RS32 idx, int32_t n_samples) {
  return Check(fb200_choice_from_keys(keys.typed_data(), Dim(keys, 0), n_samples, idx->typed_data(), s), "fb200_choice_from_keys");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_choice_from_keys, ChoiceFromKeys,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<U32>().Ret<S32>().Attr<int32_t>("n_samples"));

// ================================================================================================ K1 sampler
// One SGHMC / SGLD step with TrainState.apply_gradients fused (sghmc_integrator.py:59-92 + optax.apply_updates).
// params, momentum, precond_v are donated (input_output_aliases {0:0, 2:1, 3:2}); key -> new key.
static ffi::Error SgmcmcStep(cudaStream_t s, F32 params, F32 grad, F32 momentum, F32 precond_v, S64 leaves, S32 tiles, U32 key,
                             RF32 params_out, RF32 momentum_out, RF32 precond_out, RU32 key_out, RU32 leaf_keys_ws,
                             float step_size, double momentum_decay, int32_t zero_momentum, int32_t rmsprop, double rms_alpha,
                             double rms_eps) {
  const size_t n = grad.element_count();
  CopyIfNotAliased(params.typed_data(), params_out->typed_data(), n * 4, s);
  CopyIfNotAliased(momentum.typed_data(), momentum_out->typed_data(), n * 4, s);
  if (rmsprop) CopyIfNotAliased(precond_v.typed_data(), precond_out->typed_data(), n * 4, s);
  fb200_sgmcmc_hparams_t hp{step_size, momentum_decay, zero_momentum, rmsprop, rms_alpha, rms_eps, nullptr, FB200_F32};
  return Check(fb200_sgmcmc_step_f32(params_out->typed_data(), grad.typed_data(), momentum_out->typed_data(),
                                     rmsprop ? precond_out->typed_data() : nullptr, nullptr, static_cast<int64_t>(n),
                                     reinterpret_cast<const fb200_leaf_t*>(leaves.typed_data()), Dim(leaves, 0),
                                     reinterpret_cast<const fb200_tile_t*>(tiles.typed_data()), tiles.dimensions()[0],
                                     key.typed_data(), key_out->typed_data(), leaf_keys_ws->typed_data(), &hp, s),
               "fb200_sgmcmc_step_f32");
}
#define FB200_HP_ATTRS                                                                                                  \
  .Attr<float>("step_size").Attr<double>("momentum_decay").Attr<int32_t>("zero_momentum").Attr<int32_t>("rmsprop")      \
      .Attr<double>("rms_alpha").Attr<double>("rms_eps")
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_sgmcmc_step, SgmcmcStep,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<S64>().Arg<S32>()
                                  .Arg<U32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<U32>().Ret<U32>() FB200_HP_ATTRS);

// The same with the log-joint gradient assembled in the kernel (joint/base.py:54-122, prior/gaussian.py:29-32):
// grad = gradient of the batch log-likelihood SUM; also returns sum(theta^2) of the pre-update parameters.
static ffi::Error SgmcmcStepJoint(cudaStream_t s, F32 params, F32 grad, F32 momentum, F32 precond_v, S64 leaves, S32 tiles,
                                  U32 key, RF32 params_out, RF32 momentum_out, RF32 precond_out, RU32 key_out,
                                  RU32 leaf_keys_ws, RF64 sq_ws, RF64 theta_sq, float step_size, double momentum_decay,
                                  int32_t zero_momentum, int32_t rmsprop, double rms_alpha, double rms_eps, float lik_scale,
                                  float prior_prec) {
  const size_t n = grad.element_count();
  CopyIfNotAliased(params.typed_data(), params_out->typed_data(), n * 4, s);
  CopyIfNotAliased(momentum.typed_data(), momentum_out->typed_data(), n * 4, s);
  if (rmsprop) CopyIfNotAliased(precond_v.typed_data(), precond_out->typed_data(), n * 4, s);
  fb200_sgmcmc_hparams_t hp{step_size, momentum_decay, zero_momentum, rmsprop, rms_alpha, rms_eps, nullptr, FB200_F32};
  fb200_joint_grad_t jg{lik_scale, prior_prec};
  return Check(fb200_sgmcmc_step_joint_f32(params_out->typed_data(), grad.typed_data(), momentum_out->typed_data(),
                                           rmsprop ? precond_out->typed_data() : nullptr, static_cast<int64_t>(n),
                                           reinterpret_cast<const fb200_leaf_t*>(leaves.typed_data()), Dim(leaves, 0),
                                           reinterpret_cast<const fb200_tile_t*>(tiles.typed_data()), tiles.dimensions()[0],
                                           key.typed_data(), key_out->typed_data(), leaf_keys_ws->typed_data(), &hp, &jg,
                                           Opt(sq_ws), Opt(theta_sq), s),
               "fb200_sgmcmc_step_joint_f32");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_sgmcmc_step_joint, SgmcmcStepJoint,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<S64>().Arg<S32>()
                                  .Arg<U32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<U32>().Ret<U32>().Ret<F64>().Ret<F64>()
                                      FB200_HP_ATTRS.Attr<float>("lik_scale").Attr<float>("prior_prec"));

// Exploration phase of cyclical SGLD (cyclical_sgld_integrator.py:65-84): preconditioned ascent step, no noise.
static ffi::Error SgdExploreStep(cudaStream_t s, F32 params, F32 grad, F32 precond_v, RF32 params_out, RF32 precond_out,
                                 float step_size, double momentum_decay, int32_t zero_momentum, int32_t rmsprop,
                                 double rms_alpha, double rms_eps) {
  const size_t n = grad.element_count();
  CopyIfNotAliased(params.typed_data(), params_out->typed_data(), n * 4, s);
  if (rmsprop) CopyIfNotAliased(precond_v.typed_data(), precond_out->typed_data(), n * 4, s);
  fb200_sgmcmc_hparams_t hp{step_size, momentum_decay, zero_momentum, rmsprop, rms_alpha, rms_eps, nullptr, FB200_F32};
  return Check(fb200_sgd_explore_step_f32(params_out->typed_data(), grad.typed_data(),
                                          rmsprop ? precond_out->typed_data() : nullptr, nullptr, static_cast<int64_t>(n), &hp, s),
               "fb200_sgd_explore_step_f32");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_sgd_explore_step, SgdExploreStep,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Arg<F32>().Ret<F32>().Ret<F32>() FB200_HP_ATTRS);

static ffi::Error SgdExploreStepJoint(cudaStream_t s, F32 params, F32 grad, F32 precond_v, RF32 params_out, RF32 precond_out,
                                      RF64 sq_ws, RF64 theta_sq, float step_size, double momentum_decay, int32_t zero_momentum,
                                      int32_t rmsprop, double rms_alpha, double rms_eps, float lik_scale, float prior_prec) {
  const size_t n = grad.element_count();
  CopyIfNotAliased(params.typed_data(), params_out->typed_data(), n * 4, s);
  if (rmsprop) CopyIfNotAliased(precond_v.typed_data(), precond_out->typed_data(), n * 4, s);
  fb200_sgmcmc_hparams_t hp{step_size, momentum_decay, zero_momentum, rmsprop, rms_alpha, rms_eps, nullptr, FB200_F32};
  fb200_joint_grad_t jg{lik_scale, prior_prec};
  return Check(fb200_sgd_explore_step_joint_f32(params_out->typed_data(), grad.typed_data(),
                                                rmsprop ? precond_out->typed_data() : nullptr, static_cast<int64_t>(n), &hp, &jg,
                                                Opt(sq_ws), Opt(theta_sq), s),
               "fb200_sgd_explore_step_joint_f32");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_sgd_explore_step_joint, SgdExploreStepJoint,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Arg<F32>().Ret<F32>().Ret<F32>().Ret<F64>()
                                  .Ret<F64>() FB200_HP_ATTRS.Attr<float>("lik_scale").Attr<float>("prior_prec"));

// sample bank row write (sgmcmc_posterior_state_repository.py:39-54): bank is donated
static ffi::Error BankPut(cudaStream_t s, F32 bank, F32 src, RF32 bank_out, int32_t row) {
  CopyIfNotAliased(bank.typed_data(), bank_out->typed_data(), bank.size_bytes(), s);
  return Check(fb200_bank_put_f32(bank_out->typed_data(), bank.dimensions()[1], row, src.typed_data(), s), "fb200_bank_put_f32");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_bank_put, BankPut,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Ret<F32>().Attr<int32_t>("row"));

// ================================================================================================ K2 last layer
// calibrated logits of S drawn samples: x [B,D], w [n,C,D] (K-major) or [n,D,C], bias [n,C], log_temp [1], idx [S]
static ffi::Error LastLayer(cudaStream_t s, Any x, Any w, F32 bias, F32 log_temp, S32 sample_idx, RAny out, int32_t w_layout,
                            int32_t path) {
  const int in_dt = DTypeOf(x), out_dt = DTypeOf(*out);
  if (in_dt < 0 || out_dt < 0 || DTypeOf(w) != in_dt) return BadDType("fb200_xla_last_layer");
  const int S = Dim(*out, 0), B = Dim(*out, 1), C = Dim(*out, 2), D = Dim(x, 1);
  return Check(fb200_last_layer_logits(x.untyped_data(), w.untyped_data(), Opt(bias), Opt(log_temp), Opt(sample_idx), Dim(w, 0),
                                       out->untyped_data(), S, B, D, C, in_dt, out_dt, w_layout, path, s),
               "fb200_last_layer_logits");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_last_layer, LastLayer,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<Any>().Arg<Any>().Arg<F32>().Arg<F32>().Arg<S32>().Ret<Any>()
                                  .Attr<int32_t>("w_layout").Attr<int32_t>("path"));

// the same with the softmax / log-softmax epilogue and / or the row logsumexp (prob_output_layer/classification.py:25-69)
static ffi::Error LastLayerEx(cudaStream_t s, Any x, Any w, F32 bias, F32 log_temp, S32 sample_idx, RAny out, RF32 lse, RU8 ws,
                              int32_t w_layout, int32_t epilogue) {
  const int in_dt = DTypeOf(x), out_dt = DTypeOf(*out);
  if (in_dt < 0 || out_dt < 0 || DTypeOf(w) != in_dt) return BadDType("fb200_xla_last_layer_ex");
  const int S = Dim(*out, 0), B = Dim(*out, 1), C = Dim(*out, 2), D = Dim(x, 1);
  return Check(fb200_last_layer_logits_ex(x.untyped_data(), w.untyped_data(), Opt(bias), Opt(log_temp), Opt(sample_idx),
                                          Dim(w, 0), out->untyped_data(), S, B, D, C, in_dt, out_dt, w_layout, epilogue,
                                          Opt(lse), Opt(ws), ws->size_bytes(), s),
               "fb200_last_layer_logits_ex");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_last_layer_ex, LastLayerEx,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<Any>().Arg<Any>().Arg<F32>().Arg<F32>().Arg<S32>().Ret<Any>()
                                  .Ret<F32>().Ret<U8>().Attr<int32_t>("w_layout").Attr<int32_t>("epilogue"));

// log_prob / ensemble_log_prob from logits + their row logsumexp (predictive/base.py:104-128, 179-203)
static ffi::Error LogProbFromLse(cudaStream_t s, Any logits, F32 lse, S32 targets, RF32 ens, RF32 log_prob) {
  const int dt = DTypeOf(logits);
  if (dt < 0) return BadDType("fb200_xla_log_prob_from_lse");
  return Check(fb200_log_prob_from_lse(logits.untyped_data(), dt, lse.typed_data(), targets.typed_data(), Dim(logits, 0),
                                       Dim(logits, 1), Dim(logits, 2), Opt(ens), log_prob->typed_data(), s),
               "fb200_log_prob_from_lse");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_log_prob_from_lse, LogProbFromLse,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<Any>().Arg<F32>().Arg<S32>().Ret<F32>().Ret<F32>());

// W [S,D,C] (flax nn.Dense layout) -> K-major bf16 [S,C,D] for the tcgen05 path, once per sample bank
static ffi::Error TransposeW(cudaStream_t s, Any w, RBF16 wt) {
  const int dt = DTypeOf(w);
  if (dt < 0) return BadDType("fb200_xla_transpose_w");
  return Check(fb200_transpose_w_bf16(w.untyped_data(), dt, wt->untyped_data(), Dim(w, 0), Dim(w, 1), Dim(w, 2), s),
               "fb200_transpose_w_bf16");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_transpose_w, TransposeW, ffi::Ffi::Bind().Ctx<Stream>().Arg<Any>().Ret<BF16>());

// in-place row softmax / log_softmax / temperature scaling (+ row logsumexp); z is donated
static ffi::Error SoftmaxRows(cudaStream_t s, Any z, F32 log_temp, RAny z_out, RF32 lse, int32_t epilogue) {
  const int dt = DTypeOf(z);
  if (dt < 0) return BadDType("fb200_xla_softmax_rows");
  CopyIfNotAliased(z.untyped_data(), z_out->untyped_data(), z.size_bytes(), s);
  const int nd = static_cast<int>(z.dimensions().size());
  const int C = Dim(z, nd - 1);
  return Check(fb200_softmax_rows(z_out->untyped_data(), dt, static_cast<int64_t>(z.element_count() / (C ? C : 1)), C, epilogue,
                                  Opt(log_temp), Opt(lse), s),
               "fb200_softmax_rows");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_softmax_rows, SoftmaxRows,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<Any>().Arg<F32>().Ret<Any>().Ret<F32>().Attr<int32_t>("epilogue"));

// ================================================================================================ K3 / K4 reductions
// every classification statistic of one [S,B,C] logits tensor in one pass (predictive/base.py:624-836,
// predictive/classification.py:71-432); C > 1024 takes the two-pass kernel with `ws` = S*B floats
static ffi::Error BmaReduceCls(cudaStream_t s, Any logits, F32 log_temp, S32 targets, FB200_CLS_OUT_PARAMS, RU8 ws) {
  const int dt = DTypeOf(logits);
  if (dt < 0) return BadDType("fb200_xla_bma_reduce_cls");
  const int S = Dim(logits, 0), B = Dim(logits, 1), C = Dim(logits, 2);
  const fb200_cls_outputs_t o = FB200_CLS_OUT_STRUCT;
  if (C > 1024)
    return Check(fb200_bma_reduce_cls_wide(logits.untyped_data(), dt, Opt(log_temp), Opt(targets), S, B, C, &o, Opt(ws),
                                           ws->size_bytes(), s),
                 "fb200_bma_reduce_cls_wide");
  return Check(fb200_bma_reduce_cls(logits.untyped_data(), dt, Opt(log_temp), Opt(targets), S, B, C, &o, s), "fb200_bma_reduce_cls");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_bma_reduce_cls, BmaReduceCls,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<Any>().Arg<F32>().Arg<S32>() FB200_CLS_OUT_BINDING.Ret<U8>());

// S-sharded mode: this rank's partial sums, row by row, into the owners' slots (peer or local memory)
static ffi::Error BmaReduceClsShard(cudaStream_t s, Any logits, F32 log_temp, S32 targets, U64 slot_ptrs, RF32 ens_log_prob,
                                    int32_t rank) {
  const int dt = DTypeOf(logits);
  if (dt < 0) return BadDType("fb200_xla_bma_reduce_cls_shard");
  return Check(fb200_bma_reduce_cls_shard(logits.untyped_data(), dt, Opt(log_temp), Opt(targets), Dim(logits, 0), Dim(logits, 1),
                                          Dim(logits, 2), Dim(slot_ptrs, 0), rank, slot_ptrs.typed_data(), Opt(ens_log_prob), s),
               "fb200_bma_reduce_cls_shard");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_bma_reduce_cls_shard, BmaReduceClsShard,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<Any>().Arg<F32>().Arg<S32>().Arg<U64>().Ret<F32>()
                                  .Attr<int32_t>("rank"));

// ... and the owners' finalisation of the received slots [n_src, rows, 2C + 4]
static ffi::Error BmaFinalizeSlots(cudaStream_t s, F32 slots, S32 targets, FB200_CLS_OUT_PARAMS, int32_t n_classes,
                                   int32_t s_total) {
  const fb200_cls_outputs_t o = FB200_CLS_OUT_STRUCT;
  return Check(fb200_bma_finalize_cls_slots(slots.typed_data(), Dim(slots, 0), Dim(slots, 1), n_classes, s_total, &o,
                                            Opt(targets), s),
               "fb200_bma_finalize_cls_slots");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_bma_finalize_cls_slots, BmaFinalizeSlots,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<S32>() FB200_CLS_OUT_BINDING
                                  .Attr<int32_t>("n_classes").Attr<int32_t>("s_total"));

// Predictive.variance / .std: aleatoric + epistemic estimates that may come from different draws (predictive/base.py:838-944)
static ffi::Error VarianceCombine(cudaStream_t s, F32 aleatoric, F32 epistemic, RF32 variance, RF32 std_) {
  return Check(fb200_variance_combine(aleatoric.typed_data(), Opt(epistemic), static_cast<int64_t>(aleatoric.element_count()),
                                      Opt(variance), Opt(std_), s),
               "fb200_variance_combine");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_variance_combine, VarianceCombine,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Ret<F32>().Ret<F32>());

// regression maps over [S,B,2d] = [mu | log var] (prob_output_layer/regression.py:30-74)
static ffi::Error BmaReduceReg(cudaStream_t s, F32 outputs, F32 log_temp, F32 targets, RF32 mean, RF32 alea_var, RF32 epi_var,
                               RF32 variance, RF32 std_, RF32 log_prob, RF32 ens_log_prob) {
  fb200_reg_outputs_t o{Opt(mean), Opt(alea_var), Opt(epi_var), Opt(variance), Opt(std_), Opt(log_prob), Opt(ens_log_prob)};
  return Check(fb200_bma_reduce_reg(outputs.typed_data(), Opt(log_temp), Opt(targets), Dim(outputs, 0), Dim(outputs, 1),
                                    Dim(outputs, 2) / 2, &o, s),
               "fb200_bma_reduce_reg");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_bma_reduce_reg, BmaReduceReg,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Arg<F32>().Ret<F32>().Ret<F32>().Ret<F32>()
                                  .Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>());

// ================================================================================================ K5 - K9 scoring
// adaptive_prediction.py:11-22 (targets given: [B]) / conformal/classification/base.py:72-82 (all classes: [B,C])
static ffi::Error ApsScores(cudaStream_t s, F32 probs, S32 targets, RF32 scores, RS32 perm) {
  return Check(fb200_aps_scores(probs.typed_data(), Opt(targets), Dim(probs, 0), Dim(probs, 1), scores->typed_data(), Opt(perm), s),
               "fb200_aps_scores");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_aps_scores, ApsScores,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<S32>().Ret<F32>().Ret<S32>());

// simple_prediction.py:10-18
static ffi::Error SimpleScores(cudaStream_t s, F32 probs, S32 targets, RF32 scores) {
  return Check(fb200_simple_scores(probs.typed_data(), Opt(targets), Dim(probs, 0), Dim(probs, 1), scores->typed_data(), s),
               "fb200_simple_scores");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_simple_scores, SimpleScores, ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<S32>().Ret<F32>());

// conformal/classification/base.py:45-64: mask[b,c] = count(val > score) < threshold, sizes = row sums
static ffi::Error ConformalSets(cudaStream_t s, F32 val_sorted, F32 scores, RU8 mask, RS32 sizes, float threshold) {
  return Check(fb200_conformal_sets(val_sorted.typed_data(), Dim(val_sorted, 0), scores.typed_data(), Dim(scores, 0),
                                    Dim(scores, 1), threshold, mask->typed_data(), Opt(sizes), s),
               "fb200_conformal_sets");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_conformal_sets, ConformalSets,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Ret<U8>().Ret<S32>().Attr<float>("threshold"));

// APS scores of every class + the sets, one kernel (scores optional)
static ffi::Error ApsConformalSets(cudaStream_t s, F32 probs, F32 val_sorted, RF32 scores, RU8 mask, RS32 sizes, float threshold) {
  return Check(fb200_aps_conformal_sets(probs.typed_data(), Dim(probs, 0), Dim(probs, 1), val_sorted.typed_data(),
                                        Dim(val_sorted, 0), threshold, Opt(scores), mask->typed_data(), Opt(sizes), s),
               "fb200_aps_conformal_sets");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_aps_conformal_sets, ApsConformalSets,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Ret<F32>().Ret<U8>().Ret<S32>()
                                  .Attr<float>("threshold"));

// predictive/classification.py:465-483
static ffi::Error CredibleSets(cudaStream_t s, F32 mean, RS32 labels, RS32 sizes, float threshold) {
  return Check(fb200_credible_sets(mean.typed_data(), Dim(mean, 0), Dim(mean, 1), threshold, labels->typed_data(),
                                   sizes->typed_data(), s),
               "fb200_credible_sets");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_credible_sets, CredibleSets,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Ret<S32>().Ret<S32>().Attr<float>("threshold"));

// ascending sort of a float32 vector (validation scores, quantiles); x is donated
static ffi::Error SortF32(cudaStream_t s, F32 x, RF32 x_out, RU8 ws) {
  CopyIfNotAliased(x.typed_data(), x_out->typed_data(), x.size_bytes(), s);
  return Check(fb200_sort_f32(x_out->typed_data(), static_cast<int64_t>(x.element_count()), ws->typed_data(), ws->size_bytes(), s),
               "fb200_sort_f32");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_sort_f32, SortF32, ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Ret<F32>().Ret<U8>());

// jnp.quantile (conformal/regression/quantile.py:89-90)
static ffi::Error QuantileSorted(cudaStream_t s, F32 x_sorted, F32 q, RF32 out) {
  return Check(fb200_quantile_sorted(x_sorted.typed_data(), static_cast<int64_t>(x_sorted.element_count()), q.typed_data(),
                                     Dim(q, 0), out->typed_data(), s),
               "fb200_quantile_sorted");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_quantile_sorted, QuantileSorted, ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Ret<F32>());

static ffi::Error QuantileAxis0(cudaStream_t s, F32 x, F32 q, RF32 out) {
  const int T = Dim(x, 0);
  return Check(fb200_quantile_axis0(x.typed_data(), T, static_cast<int64_t>(x.element_count() / (T ? T : 1)), q.typed_data(),
                                    Dim(q, 0), out->typed_data(), s),
               "fb200_quantile_axis0");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_quantile_axis0, QuantileAxis0, ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Ret<F32>());

// metric/classification.py:43-220: bins, ECE, MCE, accuracy, Brier.  result = [ece, mce, accuracy, brier | confs | accs]
static ffi::Error EceBins(cudaStream_t s, F32 probs, S32 preds, S32 targets, F32 thresholds, RS32 bin_idx, RS32 counts,
                          RF32 conf_sums, RS32 correct, RF32 result) {
  const int nd = static_cast<int>(probs.dimensions().size());
  return Check(fb200_ece_bins(probs.typed_data(), Dim(probs, 0), nd > 1 ? Dim(probs, 1) : 1, Opt(preds), Opt(targets),
                              thresholds.typed_data(), Dim(thresholds, 0), Opt(bin_idx), counts->typed_data(),
                              conf_sums->typed_data(), correct->typed_data(), result->typed_data(), s),
               "fb200_ece_bins");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_ece_bins, EceBins,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<S32>().Arg<S32>().Arg<F32>().Ret<S32>().Ret<S32>()
                                  .Ret<F32>().Ret<S32>().Ret<F32>());

// the same from the per-row quantities the reduction kernel emits (confidence, mode, brier_terms)
static ffi::Error EceBinsFromRows(cudaStream_t s, F32 confidence, S32 preds, F32 brier, S32 targets, F32 thresholds, RS32 bin_idx,
                                  RS32 counts, RF32 conf_sums, RS32 correct, RF32 result) {
  return Check(fb200_ece_bins_from_rows(confidence.typed_data(), preds.typed_data(), Opt(brier), Opt(targets), Dim(confidence, 0),
                                        thresholds.typed_data(), Dim(thresholds, 0), Opt(bin_idx), counts->typed_data(),
                                        conf_sums->typed_data(), correct->typed_data(), result->typed_data(), s),
               "fb200_ece_bins_from_rows");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_ece_bins_from_rows, EceBinsFromRows,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<S32>().Arg<F32>().Arg<S32>().Arg<F32>().Ret<S32>()
                                  .Ret<S32>().Ret<F32>().Ret<S32>().Ret<F32>());

// sharded evaluation: pack the additive counters (-> one psum), finalise from the packed sums
static ffi::Error EcePackBins(cudaStream_t s, S32 counts, F32 conf_sums, S32 correct, F64 packed_in, RF64 packed,
                              int32_t accumulate) {
  if (accumulate) CopyIfNotAliased(packed_in.typed_data(), packed->typed_data(), packed_in.size_bytes(), s);
  return Check(fb200_ece_pack_bins(counts.typed_data(), conf_sums.typed_data(), correct.typed_data(), Dim(counts, 0),
                                   packed->typed_data(), accumulate, s),
               "fb200_ece_pack_bins");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_ece_pack_bins, EcePackBins,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<S32>().Arg<F32>().Arg<S32>().Arg<F64>().Ret<F64>()
                                  .Attr<int32_t>("accumulate"));

static ffi::Error EceFromPackedBins(cudaStream_t s, F64 packed, RS32 counts, RF32 result, int64_t n_total) {
  return Check(fb200_ece_from_packed_bins(packed.typed_data(), Dim(packed, 0) / 3, n_total, Opt(counts), result->typed_data(), s),
               "fb200_ece_from_packed_bins");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_ece_from_packed_bins, EceFromPackedBins,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F64>().Ret<S32>().Ret<F32>().Attr<int64_t>("n_total"));

// Predictive.sample for classification (predictive/base.py:318-440, prob_output_layer/classification.py:33-51)
static ffi::Error ClsSampleTargets(cudaStream_t s, F32 logits, F32 log_temp, U32 key, RS32 y, RS32 idx) {
  return Check(fb200_cls_sample_targets(logits.typed_data(), Opt(log_temp), key.typed_data(), Dim(*y, 0), Dim(logits, 0),
                                        Dim(logits, 1), Dim(logits, 2), y->typed_data(), Opt(idx), s),
               "fb200_cls_sample_targets");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_cls_sample_targets, ClsSampleTargets,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Arg<U32>().Ret<S32>().Ret<S32>());

// SG-MCMC diagnostics over the sample bank (sgmcmc_diagnostic.py:16-140)
static ffi::Error KsdImq(cudaStream_t s, F32 samples, F32 grads, RU8 ws, RF32 out, float c, float beta) {
  return Check(fb200_ksd_imq(samples.typed_data(), grads.typed_data(), Dim(samples, 0), samples.dimensions()[1],
                             samples.dimensions()[1], c, beta, ws->typed_data(), ws->size_bytes(), out->typed_data(), s),
               "fb200_ksd_imq");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_ksd_imq, KsdImq,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Ret<U8>().Ret<F32>().Attr<float>("c")
                                  .Attr<float>("beta"));

static ffi::Error Ess(cudaStream_t s, F32 samples, RF32 out, int32_t use_threshold, float threshold) {
  return Check(fb200_ess(samples.typed_data(), Dim(samples, 0), samples.dimensions()[1], samples.dimensions()[1], use_threshold,
                         threshold, out->typed_data(), s),
               "fb200_ess");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_ess, Ess,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Ret<F32>().Attr<int32_t>("use_threshold")
                                  .Attr<float>("threshold"));

// ================================================================================================ the rows of SURVEY 8(f)
// clip_grandients_by_norm's multiplier (utils/training.py:14-20): min(1, max_norm / (1e-7 + ||g||)); ws: float64 scratch
static ffi::Error GradClipMult(cudaStream_t s, Any grad, RF64 ws, RF32 mult, float max_grad_norm) {
  const int dt = DTypeOf(grad);
  if (dt < 0) return BadDType("grad");
  return Check(fb200_grad_clip_mult(grad.untyped_data(), dt, nullptr, nullptr, static_cast<int64_t>(grad.element_count()),
                                    max_grad_norm, ws->typed_data(), mult->typed_data(), s),
               "fb200_grad_clip_mult");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_grad_clip_mult, GradClipMult,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<Any>().Ret<F64>().Ret<F32>().Attr<float>("max_grad_norm"));

// Predictive.sample for regression (predictive/base.py:407-441, prob_output_layer/regression.py:38-52)
static ffi::Error RegSampleTargets(cudaStream_t s, F32 outputs, F32 log_temp, U32 key, RF32 y, RS32 idx) {
  return Check(fb200_reg_sample_targets(outputs.typed_data(), Opt(log_temp), key.typed_data(), Dim(*y, 0), Dim(outputs, 0),
                                        Dim(outputs, 1), Dim(outputs, 2) / 2, y->typed_data(), Opt(idx), s),
               "fb200_reg_sample_targets");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_reg_sample_targets, RegSampleTargets,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Arg<U32>().Ret<F32>().Ret<S32>());

// RegressionPredictive.entropy / aleatoric_entropy / epistemic_entropy (predictive/regression.py:51-267)
static ffi::Error RegMcEntropies(cudaStream_t s, F32 outputs, F32 log_temp, S32 sample_idx, U32 target_keys, RF32 ent,
                                 RF32 alea, RF32 epi, int32_t n_target_samples) {
  return Check(fb200_reg_mc_entropies(outputs.typed_data(), Opt(log_temp), sample_idx.typed_data(), target_keys.typed_data(),
                                      Dim(outputs, 0), Dim(sample_idx, 0), n_target_samples, Dim(outputs, 1),
                                      Dim(outputs, 2) / 2, Opt(ent), Opt(alea), Opt(epi), s),
               "fb200_reg_mc_entropies");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_reg_mc_entropies, RegMcEntropies,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Arg<S32>().Arg<U32>().Ret<F32>().Ret<F32>()
                                  .Ret<F32>().Attr<int32_t>("n_target_samples"));

// output_calib_model Predictive.sample (output_calib_model/predictive/base.py:69-105) for both output layers
static ffi::Error ClsOutputSampleTargets(cudaStream_t s, F32 logits, F32 log_temp, U32 key, RS32 y) {
  return Check(fb200_cls_output_sample_targets(logits.typed_data(), Opt(log_temp), key.typed_data(), Dim(*y, 0),
                                               Dim(logits, 0), Dim(logits, 1), y->typed_data(), s),
               "fb200_cls_output_sample_targets");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_cls_output_sample_targets, ClsOutputSampleTargets,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Arg<U32>().Ret<S32>());
static ffi::Error RegOutputSampleTargets(cudaStream_t s, F32 outputs, F32 log_temp, U32 key, RF32 y) {
  return Check(fb200_reg_output_sample_targets(outputs.typed_data(), Opt(log_temp), key.typed_data(), Dim(*y, 0),
                                               Dim(outputs, 0), Dim(outputs, 1) / 2, y->typed_data(), s),
               "fb200_reg_output_sample_targets");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_reg_output_sample_targets, RegOutputSampleTargets,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Arg<U32>().Ret<F32>());

// ClassificationPredictive.conformal_set by importance sampling (predictive/classification.py:485-626)
static ffi::Error ConformalBayesSets(cudaStream_t s, F32 ell_train, F32 lsm_test, RU8 ws, RU8 region, RS32 sizes,
                                     RS32 rank_counts, RF32 ess, float threshold) {
  return Check(fb200_conformal_bayes_sets(ell_train.typed_data(), lsm_test.typed_data(), Dim(ell_train, 0), Dim(ell_train, 1),
                                          Dim(lsm_test, 1), Dim(lsm_test, 2), threshold, ws->typed_data(), ws->size_bytes(),
                                          region->typed_data(), Opt(sizes), Opt(rank_counts), Opt(ess), s),
               "fb200_conformal_bayes_sets");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_conformal_bayes_sets, ConformalBayesSets,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Ret<U8>().Ret<U8>().Ret<S32>().Ret<S32>()
                                  .Ret<F32>().Attr<float>("threshold"));

// calibration loss with ensemble outputs (predictive/base.py:205-287): per-sample log-likelihood sums and derivatives
static ffi::Error CalibClsLossTerms(cudaStream_t s, Any logits, S32 targets, F32 log_temp, RU8 ws, RF64 out) {
  const int dt = DTypeOf(logits);
  if (dt < 0) return BadDType("logits");
  return Check(fb200_calib_cls_loss_terms(logits.untyped_data(), dt, targets.typed_data(), Opt(log_temp), Dim(logits, 0),
                                          Dim(logits, 1), Dim(logits, 2), ws->typed_data(), ws->size_bytes(),
                                          out->typed_data(), s),
               "fb200_calib_cls_loss_terms");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_calib_cls_loss_terms, CalibClsLossTerms,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<Any>().Arg<S32>().Arg<F32>().Ret<U8>().Ret<F64>());
static ffi::Error CalibRegLossTerms(cudaStream_t s, F32 outputs, F32 targets, F32 log_temp, RU8 ws, RF64 out) {
  return Check(fb200_calib_reg_loss_terms(outputs.typed_data(), targets.typed_data(), Opt(log_temp), Dim(outputs, 0),
                                          Dim(outputs, 1), Dim(outputs, 2) / 2, ws->typed_data(), ws->size_bytes(),
                                          out->typed_data(), s),
               "fb200_calib_reg_loss_terms");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_calib_reg_loss_terms, CalibRegLossTerms,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Arg<F32>().Ret<U8>().Ret<F64>());

// calibration-model loss heads (loss/classification/focal_loss.py, loss/regression/scaled_mse.py) with their gradients,
// and optax.adam over a flat parameter vector (calib_model/config/optimizer.py:20)
static ffi::Error FocalLossGrad(cudaStream_t s, F32 logits, S32 targets, RF32 grad, RF32 loss_rows, RF64 loss_mean,
                                float gamma) {
  return Check(fb200_focal_loss_grad(logits.typed_data(), targets.typed_data(), Dim(logits, 0), Dim(logits, 1), gamma,
                                     Opt(grad), loss_rows->typed_data(), loss_mean->typed_data(), s),
               "fb200_focal_loss_grad");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_focal_loss_grad, FocalLossGrad,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<S32>().Ret<F32>().Ret<F32>().Ret<F64>()
                                  .Attr<float>("gamma"));
static ffi::Error ScaledMseLossGrad(cudaStream_t s, F32 outputs, F32 targets, RF32 grad, RF32 loss_terms, RF64 loss_mean) {
  return Check(fb200_scaled_mse_loss_grad(outputs.typed_data(), targets.typed_data(), Dim(outputs, 0), Dim(outputs, 1) / 2,
                                          Opt(grad), loss_terms->typed_data(), loss_mean->typed_data(), s),
               "fb200_scaled_mse_loss_grad");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_scaled_mse_loss_grad, ScaledMseLossGrad,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Ret<F32>().Ret<F32>().Ret<F64>());
// params / mu / nu are donated operands aliased to the results (input_output_aliases on the jax side)
static ffi::Error AdamStep(cudaStream_t s, F32 params, F32 grads, F32 mu, F32 nu, RF32 params_out, RF32 mu_out, RF32 nu_out,
                           float lr, float b1, float b2, float eps, float c1, float c2) {
  const size_t bytes = params.size_bytes();
  CopyIfNotAliased(params.typed_data(), params_out->typed_data(), bytes, s);
  CopyIfNotAliased(mu.typed_data(), mu_out->typed_data(), bytes, s);
  CopyIfNotAliased(nu.typed_data(), nu_out->typed_data(), bytes, s);
  return Check(fb200_adam_step_f32(params_out->typed_data(), grads.typed_data(), mu_out->typed_data(), nu_out->typed_data(),
                                   static_cast<int64_t>(params.element_count()), lr, b1, b2, eps, c1, c2, s),
               "fb200_adam_step_f32");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_adam_step, AdamStep,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Ret<F32>().Ret<F32>()
                                  .Ret<F32>().Attr<float>("lr").Attr<float>("b1").Attr<float>("b2").Attr<float>("eps")
                                  .Attr<float>("c1").Attr<float>("c2"));

// MaxCoverageFixedPrecisionBinaryClassificationCalibrator (conformal/classification/maxcovfixprec_binary_classfication.py)
static ffi::Error MaxcovCounts(cudaStream_t s, F32 probs, S32 targets, F32 taus, RU8 ws, RS32 counts, RS32 labels,
                               int32_t side, float threshold) {
  return Check(fb200_maxcov_counts(probs.typed_data(), targets.typed_data(), static_cast<int64_t>(probs.element_count()),
                                   taus.typed_data(), Dim(taus, 0), side, threshold, ws->typed_data(), ws->size_bytes(),
                                   counts->typed_data(), labels->typed_data(), s),
               "fb200_maxcov_counts");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_maxcov_counts, MaxcovCounts,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<S32>().Arg<F32>().Ret<U8>().Ret<S32>().Ret<S32>()
                                  .Attr<int32_t>("side").Attr<float>("threshold"));
static ffi::Error MaxcovApplyPatches(cudaStream_t s, F32 probs, RF32 out, float tau_pos, float tau_neg) {
  return Check(fb200_maxcov_apply_patches(probs.typed_data(), static_cast<int64_t>(probs.element_count()), tau_pos, tau_neg,
                                          out->typed_data(), s),
               "fb200_maxcov_apply_patches");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_maxcov_apply_patches, MaxcovApplyPatches,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Ret<F32>().Attr<float>("tau_pos").Attr<float>("tau_neg"));

#else
#error "xla/ffi/api/ffi.h not found: add jaxlib's include directory (or, for the syntax check only, fortuna_b200/xla_ffi/mock)"
#endif
