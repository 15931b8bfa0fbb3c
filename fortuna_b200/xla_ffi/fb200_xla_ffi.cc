// XLA-FFI shim over the C ABI of include/fortuna_b200.h - the binding north_star names ("thin jax.ffi
// custom-call bindings").  NOT built by fortuna_b200/build.py: it needs the XLA FFI headers that ship inside
// jaxlib (jaxlib/include/xla/ffi/api/ffi.h), and jax/jaxlib cannot be installed in the build container or on
// the GPU box.  Where jaxlib is present:
//
//   g++ -O2 -fPIC -shared -std=c++17 fb200_xla_ffi.cc -I$(python -c "import jaxlib,os;print(os.path.join(os.path.dirname(jaxlib.__file__),'include'))") \
//       -I../../include -I/usr/local/cuda/include -L../lib -lfortuna_b200 -Wl,-rpath,'$ORIGIN/../lib' -o libfb200_xla_ffi.so
//
// and register the handlers from Python with jax_bindings.py (same directory).  Every handler only unpacks
// buffers/attributes and forwards to ONE fb200_* call on XLA's stream; no arithmetic lives here.
#if __has_include("xla/ffi/api/ffi.h")
#include <cuda_runtime_api.h>

#include <cstdint>

#include "fortuna_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error Check(int rc, const char* what) {
  if (rc == 0) return ffi::Error::Success();
  return ffi::Error(rc < 0 ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal,
                    std::string(what) + ": " + fb200_error_string(rc));
}

// ---- K1: one SGHMC / SGLD step.  params, momentum, precond_v are donated (input_output_aliases) so the
// update is in place; key -> new key.  Replaces sghmc_integrator.update_fn + optax.apply_updates.
static ffi::Error SgmcmcStep(cudaStream_t stream, ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::F32> grad,
                             ffi::Buffer<ffi::F32> momentum, ffi::Buffer<ffi::F32> precond_v,
                             ffi::Buffer<ffi::S64> leaves, ffi::Buffer<ffi::S32> tiles, ffi::Buffer<ffi::U32> key,
                             ffi::ResultBuffer<ffi::F32> params_out, ffi::ResultBuffer<ffi::F32> momentum_out,
                             ffi::ResultBuffer<ffi::F32> precond_out, ffi::ResultBuffer<ffi::U32> key_out,
                             ffi::ResultBuffer<ffi::U32> leaf_keys_ws, float step_size, double momentum_decay,
                             int32_t rmsprop, double rms_alpha, double rms_eps) {
  // aliasing makes *_out the same buffers as the inputs; when XLA could not alias, copy first
  const size_t n = grad.element_count();
  auto same_or_copy = [&](const float* src, float* dst) {
    if (src != dst) cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, stream);
  };
  same_or_copy(params.typed_data(), params_out->typed_data());
  same_or_copy(momentum.typed_data(), momentum_out->typed_data());
  if (rmsprop) same_or_copy(precond_v.typed_data(), precond_out->typed_data());
  fb200_sgmcmc_hparams_t hp{step_size, momentum_decay, 0, rmsprop, rms_alpha, rms_eps};
  const int n_leaves = static_cast<int>(leaves.dimensions()[0]);
  const int64_t n_tiles = tiles.dimensions()[0];
  return Check(fb200_sgmcmc_step_f32(params_out->typed_data(), grad.typed_data(), momentum_out->typed_data(),
                                     rmsprop ? precond_out->typed_data() : nullptr, nullptr, (int64_t)n,
                                     reinterpret_cast<const fb200_leaf_t*>(leaves.typed_data()), n_leaves,
                                     reinterpret_cast<const fb200_tile_t*>(tiles.typed_data()), n_tiles,
                                     key.typed_data(), key_out->typed_data(), leaf_keys_ws->typed_data(), &hp, stream),
               "fb200_sgmcmc_step_f32");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_sgmcmc_step, SgmcmcStep,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S64>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::U32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::U32>>()
                                  .Ret<ffi::Buffer<ffi::U32>>()
                                  .Attr<float>("step_size")
                                  .Attr<double>("momentum_decay")
                                  .Attr<int32_t>("rmsprop")
                                  .Attr<double>("rms_alpha")
                                  .Attr<double>("rms_eps"));

// ---- K2: last-layer logits for S samples (bf16 operands, K-major weights) -> [S,B,C] float32
static ffi::Error LastLayer(cudaStream_t stream, ffi::Buffer<ffi::BF16> x, ffi::Buffer<ffi::BF16> wt,
                            ffi::Buffer<ffi::F32> bias, ffi::Buffer<ffi::F32> log_temp,
                            ffi::Buffer<ffi::S32> sample_idx, ffi::ResultBuffer<ffi::F32> out) {
  const int B = (int)x.dimensions()[0], D = (int)x.dimensions()[1];
  const int n = (int)wt.dimensions()[0], C = (int)wt.dimensions()[1];
  const int S = (int)sample_idx.dimensions()[0];
  return Check(fb200_last_layer_logits(x.untyped_data(), wt.untyped_data(), bias.typed_data(), log_temp.typed_data(),
                                       sample_idx.typed_data(), n, out->untyped_data(), S, B, D, C, FB200_BF16,
                                       FB200_F32, FB200_W_SCD, /*path=*/0, stream),
               "fb200_last_layer_logits");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_last_layer, LastLayer,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());

// ---- K3/K4: all classification reductions of one [S,B,C] logits tensor in one pass
static ffi::Error BmaReduceCls(cudaStream_t stream, ffi::Buffer<ffi::F32> logits, ffi::Buffer<ffi::S32> targets,
                               ffi::ResultBuffer<ffi::F32> mean, ffi::ResultBuffer<ffi::F32> alea_var,
                               ffi::ResultBuffer<ffi::F32> epi_var, ffi::ResultBuffer<ffi::F32> entropy,
                               ffi::ResultBuffer<ffi::F32> alea_ent, ffi::ResultBuffer<ffi::F32> epi_ent,
                               ffi::ResultBuffer<ffi::F32> log_prob, ffi::ResultBuffer<ffi::S32> mode) {
  const int S = (int)logits.dimensions()[0], B = (int)logits.dimensions()[1], C = (int)logits.dimensions()[2];
  fb200_cls_outputs_t o{};
  o.mean = mean->typed_data();
  o.aleatoric_variance = alea_var->typed_data();
  o.epistemic_variance = epi_var->typed_data();
  o.entropy = entropy->typed_data();
  o.aleatoric_entropy = alea_ent->typed_data();
  o.epistemic_entropy = epi_ent->typed_data();
  o.log_prob = log_prob->typed_data();
  o.mode = mode->typed_data();
  return Check(fb200_bma_reduce_cls(logits.untyped_data(), FB200_F32, nullptr, targets.typed_data(), S, B, C, &o, stream),
               "fb200_bma_reduce_cls");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_bma_reduce_cls, BmaReduceCls,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>());

// ---- K5: APS scores (targets rank-0 buffer of size B, or size 0 for "every class")
static ffi::Error ApsScores(cudaStream_t stream, ffi::Buffer<ffi::F32> probs, ffi::Buffer<ffi::S32> targets,
                            ffi::ResultBuffer<ffi::F32> scores) {
  const int B = (int)probs.dimensions()[0], C = (int)probs.dimensions()[1];
  const int32_t* t = targets.element_count() ? targets.typed_data() : nullptr;
  return Check(fb200_aps_scores(probs.typed_data(), t, B, C, scores->typed_data(), nullptr, stream), "fb200_aps_scores");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_aps_scores, ApsScores,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());

// ---- K6: rank-count conformal sets against sorted validation scores
static ffi::Error ConformalSets(cudaStream_t stream, ffi::Buffer<ffi::F32> val_sorted, ffi::Buffer<ffi::F32> test,
                                ffi::ResultBuffer<ffi::U8> mask, ffi::ResultBuffer<ffi::S32> sizes, float threshold) {
  const int B = (int)test.dimensions()[0], C = (int)test.dimensions()[1];
  return Check(fb200_conformal_sets(val_sorted.typed_data(), (int)val_sorted.element_count(), test.typed_data(), B, C,
                                    threshold, mask->typed_data(), sizes->typed_data(), stream),
               "fb200_conformal_sets");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_conformal_sets, ConformalSets,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Attr<float>("threshold"));

// ---- K7: ECE / MCE binning
static ffi::Error EceBins(cudaStream_t stream, ffi::Buffer<ffi::F32> probs, ffi::Buffer<ffi::S32> targets,
                          ffi::Buffer<ffi::F32> thresholds, ffi::ResultBuffer<ffi::S32> counts,
                          ffi::ResultBuffer<ffi::F32> conf_sums, ffi::ResultBuffer<ffi::S32> correct,
                          ffi::ResultBuffer<ffi::F32> result) {
  const int B = (int)probs.dimensions()[0], C = (int)probs.dimensions()[1];
  return Check(fb200_ece_bins(probs.typed_data(), B, C, nullptr, targets.typed_data(), thresholds.typed_data(),
                              (int)thresholds.element_count(), nullptr, counts->typed_data(),
                              conf_sums->typed_data(), correct->typed_data(), result->typed_data(), stream),
               "fb200_ece_bins");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fb200_xla_ece_bins, EceBins,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());
#else
#error "XLA FFI headers not found: this shim only builds where jaxlib is installed (see the header comment)"
#endif
