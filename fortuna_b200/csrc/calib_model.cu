// Loss-and-gradient kernels of the calibration models (`calib_model`).
//
// CalibClassifier / CalibRegressor train a backbone against a calibration loss
// (fortuna/calib_model/base.py:62-154 -> calib_model_calibrator.py:29-77 -> calib_model/loss.py:26-128), by default
//   focal_loss_fn   (fortuna/loss/classification/focal_loss.py:10-36):  -mean_i (1 - p_i)^gamma log p_i,
//                                                                        p_i = softmax(outputs_i)[target_i]
//   scaled_mse_fn   (fortuna/loss/regression/scaled_mse.py:7-30):       mean_i sum_j exp(-lv_ij) (t_ij - m_ij)^2 + lv_ij,
//                                                                        outputs_i = [m_i | lv_i]
// The reference differentiates them with autodiff through the model; the backbone (forward and backward) is outside the
// hot path, so what the path owns is the loss head: ONE pass over the outputs [n, C] returns the per-input loss terms
// and the gradient in the outputs, already scaled by 1/n, which the host hands to the backbone's backward.
#include <math.h>
#include <math_constants.h>

#include "common.cuh"

namespace fb200 {

constexpr int CM_WARPS = 8;

// one warp per row; dL/dz_c = (1/n) dl/dp p (delta_ct - p_c),
// dl/dp = gamma (1 - p)^(gamma - 1) log p - (1 - p)^gamma / p      (gamma = 0: -1 / p, the cross entropy)
__global__ void __launch_bounds__(32 * CM_WARPS) focal_loss_grad_kernel(const float* __restrict__ z,
                                                                        const int32_t* __restrict__ targets, int n, int C,
                                                                        float gamma, float inv_n, float* __restrict__ grad,
                                                                        float* __restrict__ loss_rows) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int b = blockIdx.x * CM_WARPS + w; b < n; b += gridDim.x * CM_WARPS) {
    const float* row = z + (size_t)b * C;
    const int y = targets[b];
    float m = -CUDART_INF_F;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, row[c]);
    m = group_max<32>(m);
    float se = 0.0f;
    for (int c = lane; c < C; c += 32) se += expf(row[c] - m);
    se = group_sum<32>(se);
    const bool valid = (unsigned)y < (unsigned)C;
    const float p = valid ? expf(row[y] - m) / se : 0.0f;  // one_hot of an out-of-range target is all zero: p = 0
    const float q = 1.0f - p;
    const float lp = logf(p);
    const float qg = gamma == 0.0f ? 1.0f : powf(q, gamma);
    if (lane == 0 && loss_rows) loss_rows[b] = -(qg * lp);
    if (grad) {
      float dldp = -qg / p;
      if (gamma != 0.0f) dldp += gamma * powf(q, gamma - 1.0f) * lp;
      const float k = valid ? inv_n * dldp * p : 0.0f;
      float* grow = grad + (size_t)b * C;
      for (int c = lane; c < C; c += 32) {
        const float pc = expf(row[c] - m) / se;
        grow[c] = k * ((c == y ? 1.0f : 0.0f) - pc);
      }
    }
  }
}

// one thread per (input, output dimension)
__global__ void __launch_bounds__(256) scaled_mse_grad_kernel(const float* __restrict__ outputs,
                                                              const float* __restrict__ targets, int n, int d, float inv_n,
                                                              float* __restrict__ grad, float* __restrict__ loss_terms) {
  const int64_t total = (int64_t)n * d;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / d, j = i - b * d;
    const float mu = outputs[b * 2 * d + j], lv = outputs[b * 2 * d + d + j];
    const float r = targets[i] - mu;
    const float e = expf(-lv);
    if (loss_terms) loss_terms[i] = e * (r * r) + lv;
    if (grad) {
      grad[b * 2 * d + j] = inv_n * (-2.0f * e * r);
      grad[b * 2 * d + d + j] = inv_n * (1.0f - e * (r * r));
    }
  }
}

// sum of n floats in double, fixed order (one CTA: thread t adds elements t, t + 1024, ...; then a tree): the loss value
// is the same bits run to run
__global__ void __launch_bounds__(1024) ordered_sum_kernel(const float* __restrict__ x, int64_t n, double scale,
                                                           double* __restrict__ out) {
  __shared__ double sh[1024];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) acc += (double)x[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0] * scale;
}

}  // namespace fb200

using namespace fb200;

extern "C" int fb200_focal_loss_grad(const float* logits, const int32_t* targets, int n, int C, float gamma, float* grad,
                                     float* loss_rows, double* loss_mean, void* stream) {
  if (!logits || !targets) return FB200_E_NULL;
  if (loss_mean && !loss_rows) return FB200_E_NULL;
  if (n <= 0 || C <= 0) return FB200_E_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  int grid = (n + CM_WARPS - 1) / CM_WARPS;
  if (grid > 148 * 8) grid = 148 * 8;
  focal_loss_grad_kernel<<<grid, 32 * CM_WARPS, 0, st>>>(logits, targets, n, C, gamma, 1.0f / (float)n, grad, loss_rows);
  FB200_LAUNCH_CHECK();
  if (loss_mean) {
    ordered_sum_kernel<<<1, 1024, 0, st>>>(loss_rows, n, 1.0 / (double)n, loss_mean);
    FB200_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int fb200_scaled_mse_loss_grad(const float* outputs, const float* targets, int n, int d, float* grad,
                                          float* loss_terms, double* loss_mean, void* stream) {
  if (!outputs || !targets) return FB200_E_NULL;
  if (loss_mean && !loss_terms) return FB200_E_NULL;
  if (n <= 0 || d <= 0) return FB200_E_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  int64_t blocks = ((int64_t)n * d + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  scaled_mse_grad_kernel<<<(unsigned)blocks, 256, 0, st>>>(outputs, targets, n, d, 1.0f / (float)n, grad, loss_terms);
  FB200_LAUNCH_CHECK();
  if (loss_mean) {
    ordered_sum_kernel<<<1, 1024, 0, st>>>(loss_terms, (int64_t)n * d, 1.0 / (double)n, loss_mean);
    FB200_LAUNCH_CHECK();
  }
  return 0;
}

// optax.adam over a flat float32 parameter vector (optax/_src/transform.py scale_by_adam + scale(-lr); the reference's
// default calibration optimizer, fortuna/calib_model/config/optimizer.py:20): one elementwise pass, every operation
// rounded on its own like the jnp expressions (no contraction):
//   mu = (1-b1) g + b1 mu;  nu = (1-b2) g^2 + b2 nu;  u = (mu / c1) / (sqrt(nu / c2) + eps);  p += (-lr) u
// with c1 = 1 - b1^t, c2 = 1 - b2^t computed by the caller in float32.
namespace fb200 {
__global__ void __launch_bounds__(256) adam_step_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ mu, float* __restrict__ nu, int64_t n, float lr,
                                                        float b1, float b2, float eps, float c1, float c2) {
  const float omb1 = __fsub_rn(1.0f, b1), omb2 = __fsub_rn(1.0f, b2);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float m = __fadd_rn(__fmul_rn(omb1, gi), __fmul_rn(b1, mu[i]));
    const float v = __fadd_rn(__fmul_rn(omb2, __fmul_rn(gi, gi)), __fmul_rn(b2, nu[i]));
    mu[i] = m;
    nu[i] = v;
    const float mh = __fdiv_rn(m, c1), vh = __fdiv_rn(v, c2);
    const float u = __fdiv_rn(mh, __fadd_rn(__fsqrt_rn(vh), eps));
    p[i] = __fadd_rn(p[i], __fmul_rn(-lr, u));
  }
}
}  // namespace fb200

extern "C" int fb200_adam_step_f32(float* params, const float* grads, float* mu, float* nu, int64_t n, float lr, float b1,
                                   float b2, float eps, float c1, float c2, void* stream) {
  if (!params || !grads || !mu || !nu) return FB200_E_NULL;
  if (n <= 0) return FB200_E_SHAPE;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  adam_step_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(params, grads, mu, nu, n, lr, b1, b2, eps, c1, c2);
  FB200_LAUNCH_CHECK();
  return 0;
}

// dot(a, b) in double, fixed order - the chain rule of a user-supplied loss through the temperature
// (d loss / d log_temp = -sum dloss/dz' * z', fortuna/output_calib_model/loss.py:26-79 differentiated by hand)
namespace fb200 {
__global__ void __launch_bounds__(1024) ordered_dot_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                           int64_t n, double* __restrict__ out) {
  __shared__ double sh[1024];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) acc += (double)a[i] * (double)b[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0];
}
}  // namespace fb200

extern "C" int fb200_dot_f32(const float* a, const float* b, int64_t n, double* out, void* stream) {
  if (!a || !b || !out) return FB200_E_NULL;
  if (n <= 0) return FB200_E_SHAPE;
  ordered_dot_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(a, b, n, out);
  FB200_LAUNCH_CHECK();
  return 0;
}
