// Monte-Carlo entropies of the regression predictive.
//
// Replaces RegressionPredictive.aleatoric_entropy / epistemic_entropy / entropy,
// fortuna/prob_model/predictive/regression.py:51-267:
//   key1, *keys = split(rng, 1 + S);  outputs_i = calibrated outputs of the i-th drawn posterior sample (drawn with key1)
//   y_i[t]      = mu_i + exp(0.5 logvar_i) * normal(keys[i], (T, B, d))[t]          prob_output_layer/regression.py:38-52
//   log_lik     = log N(y_i[t]; outputs_i)                                            :30-34
//   log_pred    = logsumexp_j log N(y_i[t]; outputs_j) - log S
//   aleatoric   = -(1/S) sum_i mean_t log_lik          (:104-113)
//   entropy     = -(1/S) sum_i mean_t log_pred         (:250-265)
//   epistemic   = -(1/S) sum_i mean_t (log_pred - log_lik)   (:176-192)
// The reference materialises y [S, T, B, d] and, for the two predictive entropies, evaluates S x S x T Gaussian
// log-densities per input through vmap/fori_loop.  Here one warp owns an input: lanes take (i, t) pairs, draw their
// target sample in registers (threefry block of element (t, b, k) of keys[i] - bit-identical to jax.random.normal) and
// run the j-loop with an online logsumexp; nothing but the three [B] results is written.
#include <math.h>
#include <math_constants.h>

#include "common.cuh"

namespace fb200 {

struct RegEntArgs {
  const float* z;          // [n_stored, B, 2d]
  const float* log_temp;   // [1] or null
  const int32_t* idx;      // [S] drawn sample indices
  const uint32_t* keys;    // [S, 2] target-sample keys
  int S, T, B, d;
  float* entropy;
  float* aleatoric;
  float* epistemic;
};

constexpr int REG_MAX_D = 8;

__global__ void __launch_bounds__(256) reg_mc_entropy_kernel(const RegEntArgs a) {
  const int lane = threadIdx.x & 31;
  const int b = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (b >= a.B) return;  // whole warps leave together
  const int d = a.d;
  const float lt = a.log_temp ? a.log_temp[0] : 0.0f;
  const float log2pi = 1.8378770664093453f;
  const uint32_t n = (uint32_t)a.T * (uint32_t)a.B * (uint32_t)d;
  const size_t sample_stride = (size_t)a.B * 2 * d;
  const float* zb = a.z + (size_t)b * 2 * d;
  float sum_lik = 0.0f, sum_pred = 0.0f;
  const bool need_pred = a.entropy || a.epistemic;
  for (int p = lane; p < a.S * a.T; p += 32) {
    const int i = p / a.T, t = p - i * a.T;
    const float* zi = zb + (size_t)a.idx[i] * sample_stride;
    const uint32_t k0 = a.keys[2 * i], k1 = a.keys[2 * i + 1];
    float y[REG_MAX_D];
    float ll_own = 0.0f;
#pragma unroll
    for (int k = 0; k < REG_MAX_D; ++k) {
      if (k < d) {
        const float mu = zi[k];
        const float lv = zi[d + k] + lt;
        const uint32_t e = ((uint32_t)t * (uint32_t)a.B + (uint32_t)b) * (uint32_t)d + (uint32_t)k;
        const float eps = bits_to_normal(random_bits_elem(k0, k1, n, e));
        y[k] = mu + expf(0.5f * lv) * eps;
        const float diff = y[k] - mu;
        ll_own += expf(-lv) * (diff * diff) + lv + log2pi;
      }
    }
    ll_own *= -0.5f;
    sum_lik += ll_own;
    if (need_pred) {
      float m = -CUDART_INF_F, s = 0.0f;
      for (int j = 0; j < a.S; ++j) {
        const float* zj = zb + (size_t)a.idx[j] * sample_stride;
        float acc = 0.0f;
#pragma unroll
        for (int k = 0; k < REG_MAX_D; ++k) {
          if (k < d) {
            const float mu = zj[k];
            const float lv = zj[d + k] + lt;
            const float diff = y[k] - mu;
            acc += expf(-lv) * (diff * diff) + lv + log2pi;
          }
        }
        const float ll = -0.5f * acc;
        if (ll > m) {
          s = s * expf(m - ll) + 1.0f;
          m = ll;
        } else {
          s += expf(ll - m);
        }
      }
      sum_pred += m + logf(s);
    }
  }
  sum_lik = group_sum<32>(sum_lik);
  sum_pred = group_sum<32>(sum_pred);
  if (lane == 0) {
    const float inv = 1.0f / ((float)a.S * (float)a.T);
    const float log_S = logf((float)a.S);
    const float mean_lik = sum_lik * inv;
    const float mean_pred = sum_pred * inv - log_S;
    if (a.aleatoric) a.aleatoric[b] = -mean_lik;
    if (a.entropy) a.entropy[b] = -mean_pred;
    if (a.epistemic) a.epistemic[b] = -(mean_pred - mean_lik);
  }
}

}  // namespace fb200

using namespace fb200;

extern "C" int fb200_reg_mc_entropies(const float* outputs, const float* log_temp, const int32_t* sample_idx,
                                      const uint32_t* target_keys, int n_stored, int S, int T, int B, int d,
                                      float* entropy, float* aleatoric_entropy, float* epistemic_entropy,
                                      void* stream) {
  if (!outputs || !sample_idx || !target_keys) return FB200_E_NULL;
  if (!entropy && !aleatoric_entropy && !epistemic_entropy) return FB200_E_NULL;
  if (n_stored <= 0 || S <= 0 || T <= 0 || B <= 0 || d <= 0) return FB200_E_SHAPE;
  if (d > REG_MAX_D) return FB200_E_UNSUPPORTED;
  if ((long long)T * B * d >= 0xFFFFFFFFll) return FB200_E_UNSUPPORTED;
  RegEntArgs a{};
  a.z = outputs;
  a.log_temp = log_temp;
  a.idx = sample_idx;
  a.keys = target_keys;
  a.S = S;
  a.T = T;
  a.B = B;
  a.d = d;
  a.entropy = entropy;
  a.aleatoric = aleatoric_entropy;
  a.epistemic = epistemic_entropy;
  const int64_t threads = (int64_t)B * 32;
  reg_mc_entropy_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
  FB200_LAUNCH_CHECK();
  return 0;
}
