/*
 * fortuna_b200.h - C ABI of libfortuna_b200.so: the B200 (sm_100a) implementation of
 * awslabs/fortuna's posterior-sampling + Bayesian-model-averaging hot path.
 *
 * The reference has no FFI: the path is ordinary Python on JAX.  Each entry point below replaces
 * the jnp expression group at the cited reference file:line; INTEGRATION.md shows the XLA-FFI /
 * ctypes stub a Fortuna maintainer would bind to each one.
 *
 * Conventions (all entry points):
 *   - every pointer is a DEVICE pointer owned by the caller unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued asynchronously on it and the
 *     call returns without synchronising; no global state, no allocation, reentrant across streams;
 *   - return value: 0 = enqueued; negative = argument error (FB200_E_*); positive = cudaError_t
 *     reported by the launch;
 *   - arrays are dense row-major; `dtype` is FB200_F32 or FB200_BF16.
 */
#ifndef FORTUNA_B200_H_
#define FORTUNA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FB200_ABI_VERSION 1

#define FB200_F32 0
#define FB200_BF16 1

#define FB200_E_NULL (-1)        /* a required pointer is NULL */
#define FB200_E_SHAPE (-2)       /* a size is <= 0 or inconsistent */
#define FB200_E_ALIGN (-3)       /* a base pointer is not 16-byte aligned */
#define FB200_E_UNSUPPORTED (-4) /* shape/dtype outside what the kernels cover (stated per function) */
#define FB200_E_WORKSPACE (-5)   /* workspace too small */

int fb200_abi_version(void);
/* Human-readable text for a return code of this library (static storage). */
const char* fb200_error_string(int code);

/* ------------------------------------------------------------------------------------------------
 * jax.random primitives (threefry2x32, jax 0.4.30 non-partitionable layout), bit-exact uint32.
 * Replace: jax.random.split / fold_in / bits / normal as called from fortuna/utils/random.py:11-23,
 * 38-49 and fortuna/prob_model/posterior/sgmcmc/sghmc/sghmc_integrator.py:66.
 * `key` is uint32[2] on the device.
 * ---------------------------------------------------------------------------------------------- */
/* out_keys[num][2] = jax.random.split(key, num). */
int fb200_threefry_split(const uint32_t* key, int num, uint32_t* out_keys, void* stream);
/* out_key[2] = jax.random.fold_in(key, data)  (fortuna/training/trainer.py:676-680). */
int fb200_threefry_fold_in(const uint32_t* key, uint32_t data, uint32_t* out_key, void* stream);
/* out[n] = the 32-bit stream jax.random.bits(key, (n,), uint32); n < 2^32-1. */
int fb200_random_bits(const uint32_t* key, int64_t n, uint32_t* out, void* stream);
/* out[n] = jax.random.normal(key, (n,), float32): uniform in (-1,1) then sqrt(2)*erf_inv (Giles). */
int fb200_random_normal(const uint32_t* key, int64_t n, float* out, void* stream);
/* out_idx[s] = jax.random.choice(split(key, n_draws)[s], n_samples) - the S posterior-sample index
 * draws of fortuna/prob_model/predictive/base.py:632 + .../sgmcmc/sgmcmc_posterior.py:51. */
int fb200_sample_indices(const uint32_t* key, int n_draws, int n_samples, int32_t* out_idx, void* stream);
/* out_idx[i] = jax.random.choice(keys[i], n_samples) for n_keys caller-supplied keys (uint32[n_keys][2]):
 * the single draw of SGMCMCPosterior.sample(rng) (.../sgmcmc/sgmcmc_posterior.py:51) without the split. */
int fb200_choice_from_keys(const uint32_t* keys, int n_keys, int n_samples, int32_t* out_idx, void* stream);

/* ------------------------------------------------------------------------------------------------
 * SG-MCMC sampler step (K1).  Replaces sghmc_integrator.update_fn
 * (fortuna/prob_model/posterior/sgmcmc/sghmc/sghmc_integrator.py:59-92), the preconditioner
 * (.../sgmcmc_preconditioner.py:65-93), the per-leaf noise tree (fortuna/utils/random.py:11-23),
 * the cyclical-SGLD phase switch (.../cyclical_sgld/cyclical_sgld_integrator.py:64-100) and
 * flax TrainState.apply_gradients / optax.apply_updates (call site fortuna/training/trainer.py:170).
 *
 * The parameter pytree is one flat float32 vector: leaves concatenated in jax tree-flatten order.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int64_t offset; /* first element of the leaf in the flat vectors                      */
  int64_t len;    /* number of elements (leaf.size); 0 < len < 2^32-1                   */
} fb200_leaf_t;

typedef struct {
  int32_t leaf;        /* index into the leaf table                                      */
  uint32_t pair_start; /* first threefry counter pair of this tile; tile = 1024 pairs    */
} fb200_tile_t;

#define FB200_TILE_PAIRS 1024

/* Host helper: number of tiles / fill the tile table for a leaf table (plain host memory). */
int64_t fb200_sgmcmc_num_tiles_host(const fb200_leaf_t* leaves_host, int n_leaves);
int fb200_sgmcmc_build_tiles_host(const fb200_leaf_t* leaves_host, int n_leaves, fb200_tile_t* tiles_host,
                                  int64_t n_tiles);

typedef struct {
  float lik_scale;
  float prior_prec;
} fb200_joint_grad_t;

typedef struct {
  float step_size;        /* epsilon_t = step_schedule(count), evaluated by the caller in float32       */
  double momentum_decay;  /* beta as the python float Fortuna holds; 0 for SGLD.  The library forms      */
                          /* float32(beta) and sqrt(float32(2*(1-beta))) exactly as jax weak-typing does */
  int32_t zero_momentum;  /* 1: use m=0 this step (momentum_resample_steps hit); Fortuna never sets it   */
  int32_t rmsprop;        /* 0 identity preconditioner, 1 RMSprop                                        */
  double rms_alpha;       /* running_average_factor (python float); float32(alpha), float32(1-alpha)     */
  double rms_eps;         /* eps (python float)                                                          */
  /* gradient-side options; zero-initialised = off, so older callers are unaffected */
  const float* grad_mult; /* optional DEVICE float[1]: the gradient is scaled by it inside the step - the multiplier of
                             clip_grandients_by_norm (fortuna/utils/training.py:14-20), see fb200_grad_clip_mult      */
  int32_t grad_dtype;     /* dtype of `grad`: FB200_F32 (0) or FB200_BF16 (1: 18 instead of 20 B/param; single-GPU
                             fb200_sgmcmc_step_f32 / _joint_f32 only)                                                */
} fb200_sgmcmc_hparams_t;

/* One sampling step (SGHMC, or the SGLD phase of cyclical SGLD with momentum_decay = 0):
 *   (key, new_key) = split(key_in); key_out <- new_key; leaf keys = split(key, n_leaves);
 *   v' = alpha v + (1-alpha) g^2; n = N(0,1) * sqrt(eps + sqrt(v')); m' = beta m + g sqrt(eps_t) + n sqrt(2(1-beta));
 *   u = m' / (eps + sqrt(v')) * sqrt(eps_t);  params += u;  updates_out = u.
 * params / updates_out may each be NULL (contract-exact optax mode returns only `updates`);
 * precond_v must be non-NULL iff hp->rmsprop.  key_in and key_out must not alias.
 * leaf_keys_ws: device scratch uint32[n_leaves][2] (receives the per-leaf keys; inspectable).
 * leaves / tiles: device copies of the tables above.  All float vectors 16-byte aligned. */
int fb200_sgmcmc_step_f32(float* params, const float* grad, float* momentum, float* precond_v,
                          float* updates_out, int64_t n, const fb200_leaf_t* leaves, int n_leaves,
                          const fb200_tile_t* tiles, int64_t n_tiles, const uint32_t* key_in,
                          uint32_t* key_out, uint32_t* leaf_keys_ws, const fb200_sgmcmc_hparams_t* hp,
                          void* stream);

/* Gradient clipping by global norm (fortuna/utils/training.py:14-20, call site fortuna/training/trainer.py:168-169):
 *   mult = min(1, max_grad_norm / (1e-7 + sqrt(sum g^2)))   (float32; the sum of squares in float64, deterministic)
 * written to mult_out (device float[1]); hand that pointer to the step through hp->grad_mult and the scaling rides inside
 * the step kernel - the clipped gradient is never written.  joint (optional): g is the log-joint gradient the fused
 * mode forms, lik_scale * grad - prior_prec * params (then `params` is required), else g = grad.  One extra read of the
 * gradient (4 B/param; 8 with the prior).  workspace: fb200_grad_clip_workspace_doubles(n) doubles. */
int64_t fb200_grad_clip_workspace_doubles(int64_t n);
int fb200_grad_clip_mult(const void* grad, int grad_dtype, const float* params, const fb200_joint_grad_t* joint, int64_t n,
                         float max_grad_norm, double* workspace, float* mult_out, void* stream);

/* Exploration step of cyclical SGLD (cyclical_sgld_integrator.py:65-84): preconditioner update and
 * u = (eps_t g) / (eps + sqrt(v')); params += u; no noise; key and momentum untouched. */
int fb200_sgd_explore_step_f32(float* params, const float* grad, float* precond_v, float* updates_out,
                               int64_t n, const fb200_sgmcmc_hparams_t* hp, void* stream);

/* Fused log-joint mode of the two step kernels: `grad_loglik` is the gradient of the batch
 * log-likelihood SUM only, and the kernel forms the gradient of Joint._batched_log_joint_prob
 * (fortuna/prob_model/joint/base.py:54-122) with an isotropic Gaussian prior (fortuna/prob_model/prior/gaussian.py:29-32)
 * in its prologue:  g = lik_scale * grad_loglik - prior_prec * theta,  lik_scale = n_data / batch_size,
 * prior_prec = exp(-log_var).  theta is read by the update anyway, so the prior gradient costs no HBM traffic (autograd
 * through the prior re-reads every parameter several times per step).  theta_sq_out (optional, double[1]) receives
 * sum(theta^2) of the PRE-update parameters - the log-prior value -0.5 (prec * sum + n (log 2 pi + log_var)) of the
 * logged loss; sq_workspace: double[n_tiles] (step) / double[fb200_sgd_explore_joint_workspace_doubles(n)] (explore),
 * required iff theta_sq_out is given.  params is required (in-place apply mode only). */
int fb200_sgmcmc_step_joint_f32(float* params, const float* grad_loglik, float* momentum, float* precond_v, int64_t n,
                                const fb200_leaf_t* leaves, int n_leaves, const fb200_tile_t* tiles, int64_t n_tiles,
                                const uint32_t* key_in, uint32_t* key_out, uint32_t* leaf_keys_ws,
                                const fb200_sgmcmc_hparams_t* hp, const fb200_joint_grad_t* joint, double* sq_workspace,
                                double* theta_sq_out, void* stream);

/* Data-parallel sampler step: the multi-device trainer's gradient all-reduce (lax.pmean over the "batch" axis,
 * fortuna/training/trainer.py:793-795, MultiDeviceMixin) + the replicated optimizer step
 * (fortuna/training/trainer.py:170 -> the integrators above) + nothing else, as ONE kernel per rank over NVLink peer
 * memory.  Every rank holds a full parameter replica and a full gradient buffer, both mapped into every other rank's
 * address space (symmetric memory); rank r owns tiles [tile_begin, tile_end) of the common tile table and, for their
 * elements only: loads the gradient from all ranks (summed in rank order, times 1 / n_ranks), updates ITS slice of
 * momentum / precond_v (full-size arrays, only the owned slice is touched), and stores the new parameters into every
 * replica.  The noise of an element depends only on (key, leaf, element), so the result is bit-identical to the
 * single-GPU step on the averaged gradient whatever the split.  The caller orders the launch between two barriers
 * (all gradients written / all parameters stored).  key_in/key_out advance identically on every rank.
 * joint: optional (NULL = grad buffers hold the full log-joint gradient). */
#define FB200_MAX_PEERS 8
typedef struct {
  int n_ranks;
  int rank;
  const float* grad[FB200_MAX_PEERS];   /* [n] per rank, peer-mapped */
  float* params[FB200_MAX_PEERS];       /* [n] per rank, peer-mapped */
  const float* mc_grad;                 /* optional: NVLS multicast addresses of the same buffers (both or neither). */
  float* mc_params;                     /* When set the gradient is summed inside the NVSwitch (multimem.ld_reduce - the
                                           sum order is the switch's, not rank order) and the parameters are replicated by
                                           it (multimem.st): every rank then moves 4 B/param in and out whatever N is. */
} fb200_dp_peers_t;
int fb200_sgmcmc_step_dp_f32(float* momentum, float* precond_v, int64_t n, const fb200_leaf_t* leaves, int n_leaves,
                             const fb200_tile_t* tiles, int64_t tile_begin, int64_t tile_end, const uint32_t* key_in,
                             uint32_t* key_out, uint32_t* leaf_keys_ws, const fb200_sgmcmc_hparams_t* hp,
                             const fb200_joint_grad_t* joint, const fb200_dp_peers_t* peers, void* stream);
/* The exploration phase of cyclical SGLD (fb200_sgd_explore_step_f32) in the same data-parallel form: this rank updates
 * elements [elem_begin, elem_end) (elem_begin a multiple of 4) of every replica from the rank-averaged gradient. */
int fb200_sgd_explore_step_dp_f32(float* precond_v, int64_t n, int64_t elem_begin, int64_t elem_end,
                                  const fb200_sgmcmc_hparams_t* hp, const fb200_joint_grad_t* joint,
                                  const fb200_dp_peers_t* peers, void* stream);
int64_t fb200_sgd_explore_joint_workspace_doubles(int64_t n);
int fb200_sgd_explore_step_joint_f32(float* params, const float* grad_loglik, float* precond_v, int64_t n,
                                     const fb200_sgmcmc_hparams_t* hp, const fb200_joint_grad_t* joint,
                                     double* sq_workspace, double* theta_sq_out, void* stream);

/* dst[row, :] = src[:] : snapshot of the chain's (trainable) parameters into row `row` of the
 * on-device sample bank [n_rows, n] (replaces the device_get + msgpack put of
 * .../sgmcmc/sgmcmc_sampling_callback.py:36-47). */
int fb200_bank_put_f32(float* bank, int64_t n, int row, const float* src, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Last-layer logits for S posterior samples (K2).  Replaces Likelihood._get_batched_calibrated_outputs
 * (fortuna/likelihood/base.py:429-458) restricted to the final nn.Dense (fortuna/model/mlp.py:227,
 * lenet.py:129, resnet.py:248) + ClassificationTemperatureScaler (fortuna/output_calibrator/classification.py:14-17).
 *   out[s,b,c] = (sum_d x[b,d] * w[s,d,c] + bias[s,c]) * exp(-log_temp)
 * x [B,D]; w: see layout; bias [S,C] float32 or NULL; log_temp: device float[1] or NULL; out [S,B,C].
 * w_layout: FB200_W_SDC = [S,D,C] (flax kernel layout), FB200_W_SCD = [S,C,D] (K-major copy).
 * The tcgen05 path needs in_dtype = BF16, w_layout = SCD, D % 64 == 0 (use fb200_transpose_w_bf16 once
 * per sample bank); every other shape runs the CUDA-core path.  `path`: 0 auto, 1 force CUDA-core,
 * 2 force the one-CTA tcgen05 kernel, 3 force the CTA-pair (cta_group::2) tcgen05 kernel (FB200_E_UNSUPPORTED if the
 * shape does not qualify).  Auto: CTA-pair kernel for D >= 1024, one-CTA kernel for D >= 256 and C >= 64, bf16 layers
 * narrower than 64 classes with the samples folded into the column axis, CUDA cores otherwise.
 * sample_idx: optional device int32[S]; row s of the output uses w[sample_idx[s]] / bias[sample_idx[s]] - the
 * posterior.sample() index draw (fortuna/prob_model/posterior/sgmcmc/sgmcmc_posterior.py:48-52) fused into
 * the contraction, so drawn samples are never gathered or copied.  NULL = identity.
 * n_stored: number of sample rows held by w / bias (the sample bank size); must equal S when sample_idx is NULL.
 * ---------------------------------------------------------------------------------------------- */
#define FB200_W_SDC 0
#define FB200_W_SCD 1
int fb200_last_layer_logits(const void* x, const void* w, const float* bias, const float* log_temp,
                            const int32_t* sample_idx, int n_stored, void* out, int S, int B, int D, int C,
                            int in_dtype, int out_dtype, int w_layout, int path, void* stream);
/* The same contraction with the softmax / log-softmax of ClassificationProbOutputLayer fused behind it
 * (fortuna/prob_output_layer/classification.py:25-28 log_prob = logits[y] - logsumexp, :54-69 mean = softmax):
 *   epilogue = FB200_EPI_LOGITS       out = calibrated logits (as above)
 *            = FB200_EPI_SOFTMAX      out = softmax_c(logits)
 *            = FB200_EPI_LOG_SOFTMAX  out = logits - logsumexp_c(logits)
 *   lse_out (optional, float32 [S,B]) = logsumexp_c of the calibrated logits of every row.
 * On the tcgen05 path (bf16, K-major W, D >= 256, C >= 64, B >= 128) the row statistics are taken in the kernel's
 * epilogue straight from the TMEM accumulators: for C <= 256 a row is complete inside one 128 x 256 tile (two passes
 * over TMEM: statistics, then the normalised values are what gets stored); for wider layers (epilogue must be
 * FB200_EPI_LOGITS) each column tile leaves a (max, sum-exp) pair per row in `workspace`
 * (fb200_last_layer_ex_workspace_bytes) and a small kernel merges them into lse_out.  With lse_out,
 * fb200_log_prob_from_lse turns predictive.log_prob into S*B gathers instead of a second pass over [S,B,C].
 * Shapes outside the tcgen05 path run the regular dispatch followed by one in-place row pass on CUDA cores. */
#define FB200_EPI_LOGITS 0
#define FB200_EPI_SOFTMAX 1
#define FB200_EPI_LOG_SOFTMAX 2
size_t fb200_last_layer_ex_workspace_bytes(int S, int B, int C);
int fb200_last_layer_logits_ex(const void* x, const void* w, const float* bias, const float* log_temp,
                               const int32_t* sample_idx, int n_stored, void* out, int S, int B, int D, int C,
                               int in_dtype, int out_dtype, int w_layout, int epilogue, float* lse_out,
                               void* workspace, size_t workspace_bytes, void* stream);
/* The row pass on its own, in place on z[rows, C]: softmax / log_softmax per epilogue of z * exp(-log_temp) (log_temp
 * optional device float[1], the temperature scaler); FB200_EPI_LOGITS leaves the logits - multiplied by exp(-log_temp)
 * when log_temp is given - and fills lse_out; lse_out optional float32 [rows]. */
int fb200_softmax_rows(void* z, int dtype, int64_t rows, int C, int epilogue, const float* log_temp, float* lse_out,
                       void* stream);
/* ll[s,b] = logits[s,b,targets[b]] - lse[s,b];  ensemble_log_prob[S,B] = ll (optional);
 * log_prob[B] = logsumexp_s ll - log S (optional)  (fortuna/prob_model/predictive/base.py:104-128, 179-203).
 * logits: [S,B,C] calibrated logits (dtype FB200_F32 / FB200_BF16), lse: float32 [S,B] from the call above. */
int fb200_log_prob_from_lse(const void* logits, int dtype, const float* lse, const int32_t* targets, int S, int B,
                            int C, float* ensemble_log_prob, float* log_prob, void* stream);
/* wt[s,c,d] (bf16) = w[s,d,c] (float32 or bf16 per in_dtype). */
int fb200_transpose_w_bf16(const void* w, int in_dtype, void* wt, int S, int D, int C, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Classification BMA reductions over calibrated logits z[S,B,C] (K3 + K4), one pass over HBM.
 * Replaces Predictive._batched_mean / _batched_aleatoric_variance / _batched_epistemic_variance /
 * _batched_log_prob / _batched_ensemble_log_prob (fortuna/prob_model/predictive/base.py:624-647,
 * 735-758, 806-836, 179-203, 104-128), ClassificationPredictive.mode / aleatoric_entropy /
 * epistemic_entropy / entropy (fortuna/prob_model/predictive/classification.py:71-86, 257-432) and the
 * per-sample maps of fortuna/prob_output_layer/classification.py:25-28, 54-69, 88-104.
 * Every output pointer is optional (NULL = not wanted).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  float* mean;               /* [B,C]  sum_s softmax / S                                   */
  float* aleatoric_variance; /* [B,C]  sum_s p(1-p) / S                                    */
  float* epistemic_variance; /* [B,C]  max(0, sum p^2 / S - mean^2)                        */
  float* variance;           /* [B,C]  aleatoric + epistemic (same sample set)             */
  float* std;                /* [B,C]  sqrt(variance)                                      */
  int32_t* mode;             /* [B]    argmax_c mean (first maximum)                       */
  float* entropy;            /* [B]    -sum_c mean log mean                                */
  float* aleatoric_entropy;  /* [B]    -sum_c mean_s(p log p)                              */
  float* epistemic_entropy;  /* [B]    entropy - aleatoric_entropy                         */
  float* log_prob;           /* [B]    logsumexp_s(ll) - log S          (needs targets)    */
  float* ensemble_log_prob;  /* [S,B]  ll[s,b] = z[s,b,y_b] - lse_c     (needs targets)    */
  float* confidence;         /* [B]    max_c mean - the confidence the ECE bins (metric/classification.py:74-87) */
  float* brier_terms;        /* [B]    sum_c (mean_c - onehot(y)_c)^2 (metric/classification.py:213-220; needs targets) */
} fb200_cls_outputs_t;

/* logits: [S,B,C] in `dtype`; log_temp: optional device float[1] (z = logits * exp(-log_temp));
 * targets: optional int32[B].  Supports C <= 1024 (FB200_E_UNSUPPORTED above that). */
int fb200_bma_reduce_cls(const void* logits, int dtype, const float* log_temp, const int32_t* targets, int S,
                         int B, int C, const fb200_cls_outputs_t* out, void* stream);

/* The same outputs for heads of ANY width (the single-pass kernel above keeps a row in one warp's registers: C <= 1024).
 * Two passes over [S,B,C] - row logsumexp into `workspace` (fb200_bma_reduce_cls_wide_workspace_bytes = S*B floats),
 * then one CTA per input - i.e. twice the HBM traffic. */
size_t fb200_bma_reduce_cls_wide_workspace_bytes(int S, int B);
int fb200_bma_reduce_cls_wide(const void* logits, int dtype, const float* log_temp, const int32_t* targets, int S, int B,
                              int C, const fb200_cls_outputs_t* out, void* workspace, size_t workspace_bytes,
                              void* stream);

/* S-sharded multi-GPU mode (each rank holds S/N of the posterior samples and all B inputs).
 * The same single pass, but instead of finalising it exports this rank's partial sums, row by row, into the
 * "slot" of the rank that will own the row: owner(b) = b / (B / n_ranks).  slot_ptrs is a DEVICE array of n_ranks
 * float* (stored as uint64); slot_ptrs[o] receives rows_per_rank = B / n_ranks rows of fb200_cls_slot_row_floats(C)
 * floats each:  [ sum_s p (C) | sum_s p^2 (C) | sum_s sum_c p ln p, max_s ll, sum_s exp(ll - max), 0 ].
 * A slot pointer may be local memory (a send buffer, then ONE all-to-all moves slot o to rank o) or PEER memory
 * mapped over NVLink (the kernel's epilogue stores straight into rank o's receive buffer: no collective at all,
 * only a barrier before the finalisation).
 */
size_t fb200_cls_slot_row_floats(int C);
/*
 * `rank` (this rank's index) only staggers the order in which input tiles are visited - rank r starts r/N of the way
 * through - so that at any moment the N ranks store into N different owners instead of all into the same one. */
int fb200_bma_reduce_cls_shard(const void* logits, int dtype, const float* log_temp, const int32_t* targets, int S,
                               int B, int C, int n_ranks, int rank, const uint64_t* slot_ptrs,
                               float* ensemble_log_prob /* optional [S,B]: ll of THIS rank's samples */, void* stream);
/* Finalise `rows` owned rows from the n_src slots received (slots: [n_src][rows][fb200_cls_slot_row_floats(C)],
 * one per source rank, summed in source order; the (max, sumexp) pairs are merged), over s_total samples. */
int fb200_bma_finalize_cls_slots(const float* slots, int n_src, int rows, int C, int s_total,
                                 const fb200_cls_outputs_t* out,
                                 const int32_t* targets /* optional int32[rows]: the owned rows' targets, for brier_terms */,
                                 void* stream);

/* Predictive.variance / Predictive.std (fortuna/prob_model/predictive/base.py:838-944): variance = aleatoric + epistemic
 * (the two estimates may come from different posterior draws, as in the reference), std = sqrt(variance).  epistemic may
 * be NULL (then the first array IS the variance: std(variances=...)); either output may be NULL. */
int fb200_variance_combine(const float* aleatoric, const float* epistemic, int64_t n, float* variance_out, float* std_out,
                           void* stream);

/* Regression BMA reductions over calibrated outputs z[S,B,2d] = [mu, log var]
 * (fortuna/prob_output_layer/regression.py:30-34, 67-74; fortuna/likelihood/regression.py:55-99;
 * same predictive/base.py loops).  log_temp (optional) is added to log var
 * (fortuna/output_calibrator/regression.py:16-19).  targets: optional float32 [B,d]. */
typedef struct {
  float* mean;               /* [B,d] */
  float* aleatoric_variance; /* [B,d] */
  float* epistemic_variance; /* [B,d] */
  float* variance;           /* [B,d] */
  float* std;                /* [B,d] */
  float* log_prob;           /* [B]   */
  float* ensemble_log_prob;  /* [S,B] */
} fb200_reg_outputs_t;
int fb200_bma_reduce_reg(const float* outputs, const float* log_temp, const float* targets, int S, int B, int d,
                         const fb200_reg_outputs_t* out, void* stream);

/* Monte-Carlo target samples of the regression predictive (Predictive.sample, fortuna/prob_model/predictive/base.py:
 * 318-441 -> fortuna/likelihood/base.py:339-372 -> fortuna/prob_output_layer/regression.py:38-52), same draws as jax:
 *   keys = split(key, T); (k1, k2) = split(keys[t]); i_t = choice(k1, n_stored);
 *   y[t, b, j] = mu[i_t, b, j] + exp(0.5 * logvar[i_t, b, j]) * normal(k2, (1, B, d))[0, b, j]
 * outputs: [n_stored, B, 2d] calibrated outputs of every STORED posterior sample; y: [T, B, d];
 * idx_out: optional int32[T] (the drawn sample indices).  The reference calls this once per input batch with the
 * same key, so B here is the batch the noise shape refers to. */
int fb200_reg_sample_targets(const float* outputs, const float* log_temp, const uint32_t* key, int n_target_samples,
                             int n_stored, int B, int d, float* y, int32_t* idx_out, void* stream);
/* Target samples given ONE set of outputs - the output_calib_model predictive (Predictive.sample,
 * fortuna/output_calib_model/predictive/base.py:69-105 -> RegressionProbOutputLayer.sample,
 * fortuna/prob_output_layer/regression.py:38-52):  y = mu + exp(0.5 (log var + log_temp)) * normal(key, (T, B, d)),
 * the same normals as jax.  outputs: float32 [B, 2d] UNcalibrated; y: float32 [T, B, d]; T * B * d < 2^32 - 1. */
int fb200_reg_output_sample_targets(const float* outputs, const float* log_temp, const uint32_t* key,
                                    int n_target_samples, int B, int d, float* y, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Conformal / calibration scoring (K5-K9), bit-exact integer / index outputs.
 * ---------------------------------------------------------------------------------------------- */
/* Adaptive-prediction-set scores (fortuna/conformal/classification/adaptive_prediction.py:11-22,
 * base.py:72-82).  Per row: stable descending sort (ties: larger class index first), sequential
 * float32 inclusive prefix sum.  targets != NULL: scores[B] = prefix sum at the target's rank;
 * targets == NULL: scores[B,C] for every class.  perm_out (optional) int32[B,C] = the descending
 * permutation (argsort(probs)[:, ::-1]).  C <= 4096. */
int fb200_aps_scores(const float* probs, const int32_t* targets, int B, int C, float* scores, int32_t* perm_out,
                     void* stream);
/* 1 - probs[b, y_b] (targets != NULL -> [B]) or 1 - probs ([B,C])
 * (fortuna/conformal/classification/simple_prediction.py:10-18). */
int fb200_simple_scores(const float* probs, const int32_t* targets, int B, int C, float* scores, void* stream);
/* Credible sets (fortuna/prob_model/predictive/classification.py:465-483): labels[B,C] = descending
 * permutation of mean, set_size[b] = first k with prefix sum > threshold, plus one;
 * threshold = float32(1 - error) evaluated by the caller (jax weak typing). */
int fb200_credible_sets(const float* mean, int B, int C, float threshold, int32_t* labels, int32_t* set_size,
                        void* stream);
/* In-place ascending sort of a float32 vector (used for the validation scores; n <= 2^26).
 * workspace: fb200_sort_workspace_bytes(n) bytes. */
size_t fb200_sort_workspace_bytes(int64_t n);
int fb200_sort_f32(float* x, int64_t n, void* workspace, size_t workspace_bytes, void* stream);
/* Rank-count conformal sets (fortuna/conformal/classification/base.py:45-64):
 *   mask[b,c] = ( #{v : val_scores[v] > test_scores[b,c]} as float32 ) < threshold,
 * threshold = float32((1-error)*(n_val+1)) evaluated by the caller; sizes[b] = sum_c mask.
 * val_scores_sorted: ascending, n_val values.  mask uint8 [B,C]; sizes int32 [B] (optional). */
int fb200_conformal_sets(const float* val_scores_sorted, int n_val, const float* test_scores, int B, int C,
                         float threshold, uint8_t* mask, int32_t* sizes, void* stream);
/* fb200_aps_scores (all classes) and fb200_conformal_sets in ONE kernel: the sets are decided in the epilogue of the sort
 * kernel, so the [B,C] score matrix is neither written nor read back (scores: optional output; C <= 1024). */
int fb200_aps_conformal_sets(const float* probs, int B, int C, const float* val_scores_sorted, int n_val, float threshold,
                             float* scores, uint8_t* mask, int32_t* sizes, void* stream);
/* ECE / MCE binning (fortuna/metric/classification.py:43-95, 98-133, 147-181): conf = max_c probs (C > 1)
 * or probs itself (C == 1); thresholds = the caller's float32 linspace(1/B, 1, n_bins) (device, n_bins <= 32).
 * Outputs: bin_idx int32[B] (optional; -1 = in no bin), counts int32[n_bins], conf_sums float32[n_bins],
 * correct int32[n_bins] (#(targets == preds)), and result float32[4 + 2*n_bins] = {ECE, MCE, accuracy,
 * brier (C>1 only), confs[n_bins] (mean confidence per bin, 0 if empty), accs[n_bins] (accuracy per bin, 0 if empty)}.
 * preds: optional int32[B]; NULL = argmax_c probs (first maximum).  workspace: none (outputs are zeroed here). */
int fb200_ece_bins(const float* probs, int B, int C, const int32_t* preds, const int32_t* targets,
                   const float* thresholds, int n_bins, int32_t* bin_idx, int32_t* counts, float* conf_sums,
                   int32_t* correct, float* result, void* stream);
/* The same outputs from per-row quantities the reduction kernel already has in its epilogue (fb200_cls_outputs_t:
 * confidence = max_c mean, mode = argmax, brier_terms): the [B,C] mean is not read a second time.
 * brier_terms optional (result[3] = NaN without it). */
int fb200_ece_bins_from_rows(const float* confidence, const int32_t* preds, const float* brier_terms,
                             const int32_t* targets, int B, const float* thresholds, int n_bins, int32_t* bin_idx,
                             int32_t* counts, float* conf_sums, int32_t* correct, float* result, void* stream);
/* result[2] = {ECE, MCE} over n_total data points from bin counters that were summed over shards (the counters
 * fb200_ece_bins returns are additive: all-reduce them across the ranks of a B-sharded evaluation, then call this). */
int fb200_ece_from_bins(const int32_t* counts, const float* conf_sums, const int32_t* correct, int n_bins,
                        int64_t n_total, float* result, void* stream);
/* The same in one collective: pack the three counters into packed[3 n_bins] float64 = [counts | correct | conf_sums]
 * (all-reduce THAT with a sum), then result[2] = {ECE, MCE} from the packed sums (counts_out optional int32[n_bins]).
 * accumulate != 0 adds to what `packed` holds - the bins of a loader evaluated batch by batch (thresholds from the TOTAL
 * number of data points, metric/classification.py:76). */
int fb200_ece_pack_bins(const int32_t* counts, const float* conf_sums, const int32_t* correct, int n_bins, double* packed,
                        int accumulate, void* stream);
int fb200_ece_from_packed_bins(const double* packed, int n_bins, int64_t n_total, int32_t* counts_out, float* result,
                               void* stream);
/* jnp.quantile(x, q) for sorted ascending x[n], method="linear"
 * (fortuna/conformal/regression/quantile.py:89-90, onedim_uncertainty.py:89-90): out[nq]. */
int fb200_quantile_sorted(const float* x_sorted, int64_t n, const float* q, int nq, float* out, void* stream);

/* jnp.quantile(x, q, axis=0) for x[T, M] (T <= 256), method="linear": out[nq, M]
 * (RegressionPredictive.quantile over the target-sample axis, fortuna/prob_model/predictive/regression.py:331-369). */
int fb200_quantile_axis0(const float* x, int T, int64_t M, const float* q, int nq, float* out, void* stream);

/* Calibration of ClassificationTemperatureScaler (fortuna/output_calibrator/classification.py:7-17) over a fixed ensemble
 * of outputs - the loss ProbModelOutputCalibrator minimises (fortuna/prob_model/prob_model_calibrator.py:30-61,
 * fortuna/prob_model/predictive/base.py:205-287): with tau = exp(-log_temp) and ll[s,b] = log softmax(tau z[s,b])[y_b],
 * out[s] = L_s = sum_b ll[s,b] and out[S + s] = dL_s/dtau = sum_b (z_y - sum_c p_c z_c), float64, one pass over
 * [S,B,C].  The caller forms loss = -(logsumexp_s((n_data / B) L_s) - log S) and its derivative by the scalar chain
 * rule.  logits: [S,B,C] UNcalibrated (dtype FB200_F32 / FB200_BF16); out: double[2S]. */
size_t fb200_calib_cls_workspace_bytes(int S, int B);
int fb200_calib_cls_loss_terms(const void* logits, int dtype, const int32_t* targets, const float* log_temp, int S,
                               int B, int C, void* workspace, size_t workspace_bytes, double* out, void* stream);
/* The same for RegressionTemperatureScaler (fortuna/output_calibrator/regression.py:7-19, log var + log_temp): outputs
 * float32 [S,B,2d] = [mu | log var] UNcalibrated, targets float32 [B,d];  out[s] = L_s = sum_b log N(y_b; mu, exp(lv + lt)),
 * out[S + s] = dL_s/dlog_temp. */
size_t fb200_calib_reg_workspace_bytes(int S, int B);
int fb200_calib_reg_loss_terms(const float* outputs, const float* targets, const float* log_temp, int S, int B, int d,
                               void* workspace, size_t workspace_bytes, double* out, void* stream);
/* Losses of the output calibration models (OutputCalibClassifier / OutputCalibRegressor .calibrate,
 * fortuna/output_calib_model/classification.py:71-116 and regression.py:66-114 -> Loss.__call__,
 * fortuna/output_calib_model/loss.py:26-79, differentiated by OutputCalibModelCalibrator.training_step,
 * fortuna/output_calib_model/output_calib_model_calibrator.py:224-257) with their DEFAULT loss functions, over ONE set of
 * outputs (the whole calibration set is one step); one pass returns the loss sum and its derivative:
 *   focal (fortuna/loss/classification/focal_loss.py:10-36): p_b = softmax(tau z_b)[y_b], tau = exp(-log_temp);
 *     out[0] = sum_b -(1 - p_b)^gamma log p_b,  out[1] = d out[0] / d tau   (gamma = 0 is the cross entropy)
 *   scaled MSE (fortuna/loss/regression/scaled_mse.py:6-26): lv' = log var + log_temp;
 *     out[0] = sum_b sum_k ( exp(-lv') (y - mu)^2 + lv' ),  out[1] = d out[0] / d log_temp
 * The reference's loss is the mean over b (the caller divides).  Workspaces: fb200_calib_cls_workspace_bytes(1, B) /
 * fb200_calib_reg_workspace_bytes(1, B).  out: double[2]. */
int fb200_output_calib_focal_loss_terms(const void* logits, int dtype, const int32_t* targets, const float* log_temp,
                                        int B, int C, float gamma, void* workspace, size_t workspace_bytes, double* out,
                                        void* stream);
int fb200_output_calib_scaled_mse_terms(const float* outputs, const float* targets, const float* log_temp, int B, int d,
                                        void* workspace, size_t workspace_bytes, double* out, void* stream);

/* Loss heads of the calibration models (CalibClassifier / CalibRegressor .calibrate, fortuna/calib_model/base.py:62-154
 * -> CalibModelCalibrator.training_loss_step, fortuna/calib_model/calib_model_calibrator.py:29-77 -> Loss.__call__,
 * fortuna/calib_model/loss.py:26-128, which the reference differentiates through the model by autodiff) for their DEFAULT
 * loss functions: one pass over the model outputs returns the per-input loss terms and the gradient in the outputs,
 * scaled by 1/n (the losses are means), for the backbone's backward pass:
 *   focal (fortuna/loss/classification/focal_loss.py:10-36): p = softmax(logits_i)[y_i]; loss_rows[i] = -(1-p)^gamma log p;
 *     grad[i,c] = (1/n) (gamma (1-p)^(gamma-1) log p - (1-p)^gamma / p) p (delta(c, y_i) - softmax(logits_i)[c])
 *   scaled MSE (fortuna/loss/regression/scaled_mse.py:6-26): outputs_i = [mu | log var];
 *     loss_terms[i,j] = exp(-lv) (y - mu)^2 + lv;  grad = (1/n) [-2 exp(-lv)(y - mu) | 1 - exp(-lv)(y - mu)^2]
 * grad / loss_rows may be NULL; loss_mean (double[1], optional, needs the loss terms) = their sum / n in a fixed order. */
int fb200_focal_loss_grad(const float* logits, const int32_t* targets, int n, int C, float gamma, float* grad,
                          float* loss_rows, double* loss_mean, void* stream);
int fb200_scaled_mse_loss_grad(const float* outputs, const float* targets, int n, int d, float* grad, float* loss_terms,
                               double* loss_mean, void* stream);
/* optax.adam(lr, b1, b2, eps) - the default `Optimizer.method` of the calibration configs
 * (fortuna/calib_model/config/optimizer.py:20, applied by TrainState.apply_gradients at fortuna/training/trainer.py:170) -
 * over a flat float32 parameter vector, in place: mu = (1-b1) g + b1 mu; nu = (1-b2) g^2 + b2 nu;
 * params += -lr (mu / c1) / (sqrt(nu / c2) + eps), c1 = 1 - b1^t, c2 = 1 - b2^t (float32, from the caller). */
int fb200_adam_step_f32(float* params, const float* grads, float* mu, float* nu, int64_t n, float lr, float b1, float b2,
                        float eps, float c1, float c2, void* stream);
/* out[0] = sum_i a[i] b[i] in double, fixed order: the chain rule of a USER-SUPPLIED output-calibration loss through the
 * temperature, d loss / d log_temp = -sum dloss/dz' . z' with z' the calibrated outputs
 * (fortuna/output_calib_model/loss.py:26-79, output_calibrator/classification.py:7-17). */
int fb200_dot_f32(const float* a, const float* b, int64_t n, double* out, void* stream);

/* Multivalid scorers: BatchMVP and the multicalibrators (fortuna/conformal/multivalid/iterative/base.py:34-333,
 * multicalibrator.py:14-151, batch_mvp.py:17-288, classification/top_label_multicalibrator.py:19-133).
 * fb200_multivalid_stats - the per-round statistics the reference computes as a vmap over (bucket type tau, group g, bucket
 * v_j, score dimension c) of _calibration_error: for b = cond_tau(x, v_j) * groups[:, g] (* [argmax_k values[:, k] == c]
 * when top_label), cond_0: |x - v| < h, cond_1: x + h >= v, cond_2: x - h <= v, h = float32(0.5 / n_buckets)
 * (iterative/base.py:346-364), x = values[:, 0] (Dv = 1) or values[:, c] (Dv = D):
 *   out[((tau * G + g) * n_buckets + j) * D + c][0..2] = ( |b|,  sum stat b,  sum x b )   (double; tau = 0..2)
 *   mode 0: stat = scores[:, c] (mean(scores b), multicalibrator.py:81-100); mode 1: stat = [scores <= x]
 *   (mean(conds), batch_mvp.py:166-192).  tau_mask: bit tau = compute that bucket type (others stay zero).
 * values float32 [n, Dv]; scores float32 [n, D]; groups uint8 [n, G] or NULL (one all-true group, G = 1); buckets float32
 * [n_buckets] = linspace(0, 1, n_buckets).  Deterministic (integer counts, float sums added in a fixed order).
 * fb200_multivalid_patch - iterative/base.py:381-397 on the members of ONE (tau, g, v, c), in place:
 *   values[b(, c)] = clip(values + eta patch, 0, 1)  |  clip(values (1 - eta (1 - patch)), 0, 1);  n_members = |b| (int32[1],
 *   optional).
 * fb200_multivalid_loss - mode 0: sum (values - scores)^2 over all n * Dv elements (mixins/multicalibrator.py:10-15; the
 *   caller divides where the reference takes a mean); mode 1: sum of the pinball terms at `coverage` (mixins/batchmvp.py:
 *   12-27).  terms: float32 [n * Dv] scratch.  out: double[1].
 * fb200_batchmvp_delta_counts - the patch search of BatchMVP (batch_mvp.py:194-230): counts[k] = #{i in b : scores_i <=
 *   values_i + deltas[k]}, counts[m] = |b| (uint64 [m + 1]). */
size_t fb200_multivalid_workspace_bytes(int n, int G, int n_buckets, int D);
int fb200_multivalid_stats(const float* values, const float* scores, const uint8_t* groups, const float* buckets, int n, int D,
                           int Dv, int G, int n_buckets, int top_label, int mode, int tau_mask, void* workspace,
                           size_t workspace_bytes, double* out, void* stream);
int fb200_multivalid_patch(float* values, const uint8_t* groups, int n, int Dv, int G, int g, int c, int tau, float v,
                           int n_buckets, int top_label, float eta, float patch, int multiplicative, int32_t* n_members,
                           void* stream);
int fb200_multivalid_loss(const float* values, const float* scores, int n, int D, int Dv, int mode, float coverage,
                          float* terms, double* out, void* stream);
int fb200_batchmvp_delta_counts(const float* values, const float* scores, const uint8_t* groups, int n, int G, int g, int tau,
                                float v, int n_buckets, const float* deltas, int m, uint64_t* counts, void* stream);

/* Maximum-coverage fixed-precision calibration of a binary classifier
 * (MaxCoverageFixedPrecisionBinaryClassificationCalibrator, fortuna/conformal/classification/
 * maxcovfixprec_binary_classfication.py:33-124).  For a SORTED grid of candidate taus (ascending for side 0, descending
 * for side 1 - the reference's own grids, :77-80) one pass returns, for every tau_j,
 *   side 0 (true positives): counts[j] = #{ clip((1 + (tau_j - 1) [p > 0.5]) p, max 1) >= threshold }, labels[j] = sum of
 *                            y over those points                                           (:56-61, :109-113)
 *   side 1 (true negatives): counts[j] = #{ (1 + (tau_j - 1) [p < 0.5]) p <= threshold },  labels[j] = sum of (1 - y)
 *                            (threshold = 1 - true_negative_precision_threshold)           (:63-69, :115-119)
 * as exact integers (the reference takes float32 means of the same indicators; mean = count / n).  The predicate is
 * evaluated with the reference's float32 expression, op by op.  probs float32 [n], targets int32 [n] in {0, 1},
 * taus float32 [n_taus] (n_taus <= 4096) on the device. */
size_t fb200_maxcov_workspace_bytes(int n_taus);
int fb200_maxcov_counts(const float* probs, const int32_t* targets, int64_t n, const float* taus, int n_taus, int side,
                        float threshold, void* workspace, size_t workspace_bytes, int32_t* counts, int32_t* labels,
                        void* stream);
/* apply_patches (:98-107): out = clip(tau_pos p, max 1) where p > 0.5, tau_neg p where p < 0.5, p otherwise. */
int fb200_maxcov_apply_patches(const float* probs, int64_t n, float tau_pos, float tau_neg, float* out, void* stream);

/* Target samples of the classification predictive - Predictive.sample (fortuna/prob_model/predictive/base.py:318-440)
 * -> ClassificationProbOutputLayer.sample (fortuna/prob_output_layer/classification.py:33-51), the reference's draws:
 * keys = split(key, T); (k1, k2) = split(keys[t]); i = choice(k1, n_stored); y[t, b] = choice(split(k2, B)[b], C,
 * p = softmax(logits[i, b] * exp(-log_temp)), shape = (1,)) = searchsorted(cumsum(p), cumsum(p)[-1] * (1 - uniform)).
 * logits: float32 [n_stored, B, C] (C <= 1024) of every stored posterior sample for ONE input batch; y: int32 [T, B];
 * idx_out: optional int32 [T] (the drawn posterior samples). */
int fb200_cls_sample_targets(const float* logits, const float* log_temp, const uint32_t* key, int n_target_samples,
                             int n_stored, int B, int C, int32_t* y, int32_t* idx_out, void* stream);
/* Target samples given ONE set of logits - the output_calib_model predictive (Predictive.sample,
 * fortuna/output_calib_model/predictive/base.py:69-105 -> ClassificationProbOutputLayer.sample,
 * fortuna/prob_output_layer/classification.py:33-51): y[:, b] = choice(split(key, B)[b], C, p = softmax(logits[b] *
 * exp(-log_temp)), shape = (T,)) - T uniforms under one key per input.  logits: float32 [B, C] (C <= 1024); y: int32 [T, B]. */
int fb200_cls_output_sample_targets(const float* logits, const float* log_temp, const uint32_t* key,
                                    int n_target_samples, int B, int C, int32_t* y, void* stream);

/* Monte-Carlo entropies of the regression predictive (fortuna/prob_model/predictive/regression.py:51-267), bit-exact
 * target draws.  outputs: float32 [n_stored, B, 2d] = [mu, log sigma^2] of every stored posterior sample (d <= 8);
 * sample_idx: int32 [S] drawn posterior samples (fb200_sample_indices with key1 of `key1, *keys = split(rng, 1 + S)`);
 * target_keys: uint32 [S, 2] = those `keys` (rows 1..S of fb200_threefry_split(rng, 1 + S)); T = n_target_samples.
 * Each output is optional float32 [B]:  aleatoric = -(1/S) sum_i mean_t log N(y_it; out_i),
 * entropy = -(1/S) sum_i mean_t [logsumexp_j log N(y_it; out_j) - log S],  epistemic = entropy - aleatoric
 * with y_it = mu_i + sigma_i * normal(keys[i], (T, B, d))[t]. */
int fb200_reg_mc_entropies(const float* outputs, const float* log_temp, const int32_t* sample_idx,
                           const uint32_t* target_keys, int n_stored, int S, int T, int B, int d, float* entropy,
                           float* aleatoric_entropy, float* epistemic_entropy, void* stream);

/* Conformal Bayes sets by importance sampling over the posterior samples - ClassificationPredictive.conformal_set,
 * fortuna/prob_model/predictive/classification.py:485-626.  ell_train: float32 [S, n_train] = ensemble_log_prob of the
 * training data; lsm_test: float32 [S, n_test, C] = log-softmax of the test logits under the SAME S drawn samples
 * (fb200_last_layer_logits_ex with FB200_EPI_LOG_SOFTMAX); threshold = float32(error * (n_train + 1)) evaluated by the
 * caller.  Outputs: region uint8 [n_test, C] (1 = class in the set), sizes int32 [n_test] (optional), rank_counts int32
 * [n_test, C] (optional; #{i : p_train[i] <= p_test}, rank = count + 1), ess float32 [n_test, C] (optional; effective
 * sample size of the importance weights, :601-617).  S <= 768. */
size_t fb200_conformal_bayes_workspace_bytes(int S, int n_test, int C);
int fb200_conformal_bayes_sets(const float* ell_train, const float* lsm_test, int S, int n_train, int n_test, int C,
                               float threshold, void* workspace, size_t workspace_bytes, uint8_t* region, int32_t* sizes,
                               int32_t* rank_counts, float* ess, void* stream);

/* ------------------------------------------------------------------------------------------------
 * SG-MCMC diagnostics over stored samples (fortuna/prob_model/posterior/sgmcmc/sgmcmc_diagnostic.py).
 * `samples` / `grads`: float32 [S, N] with `row_stride` floats between rows (rows of the sample bank, or any dense
 * matrix with row_stride == N); the flat order is the caller's ravel_pytree order.
 * ---------------------------------------------------------------------------------------------- */
/* kernel_stein_discrepancy_imq(samples, grads, c, beta) (:16-70): out[1] = sqrt(sum_ij k0(x_i, x_j)) / S with the IMQ
 * Stein kernel k0; c > 0, beta < 0 (else FB200_E_SHAPE, the reference raises ValueError), S <= 4096.
 * The four length-N sums per pair are accumulated in float32, k0 and the S^2 sum in float64. */
size_t fb200_ksd_imq_workspace_bytes(int S, int64_t N);
int fb200_ksd_imq(const float* samples, const float* grads, int S, int64_t N, int64_t row_stride, float c, float beta,
                  void* workspace, size_t workspace_bytes, float* out, void* stream);
/* effective_sample_size(samples, filter_threshold) (:73-140), per parameter: ess[N].  The reference's FFT
 * auto-correlation of the zero-padded chain equals the direct lagged products computed here; use_threshold = 0 is
 * filter_threshold=None.  S up to ~3000 (the chain of one 8-column slab must fit shared memory). */
int fb200_ess(const float* samples, int S, int64_t N, int64_t row_stride, int use_threshold, float threshold,
              float* ess, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FORTUNA_B200_H_ */
