// Calibration loss of the temperature scaler over a fixed ensemble of outputs.
//
// Replaces, for ClassificationTemperatureScaler (fortuna/output_calibrator/classification.py:7-17), the loss and
// gradient that ProbModelOutputCalibrator differentiates (fortuna/prob_model/prob_model_calibrator.py:30-61 ->
// Predictive._batched_log_joint_prob with ensemble_outputs, fortuna/prob_model/predictive/base.py:205-287 ->
// Likelihood._batched_log_joint_prob, fortuna/likelihood/base.py:124-204):
//   tau = exp(-log_temp);  ll[s,b] = tau z[s,b,y_b] - logsumexp_c(tau z[s,b,:])
//   L_s = sum_b ll[s,b];   log_pred = logsumexp_s( (n_data / B) L_s ) - log S;   loss = -log_pred
// The kernel returns, per posterior sample, L_s and dL_s/dtau = sum_b ( z_y - sum_c p_c z_c ) in ONE pass over [S,B,C]
// (autodiff would hold the [S,B,C] softmax for the backward pass); the scalar chain rule and the optimizer step stay
// on the host.  Sums are double, partials per CTA added in index order (deterministic).
#include <math.h>
#include <math_constants.h>

#include <cuda_bf16.h>

#include "common.cuh"

namespace fb200 {

constexpr int CAL_WARPS = 8;

template <typename T>
__device__ __forceinline__ float cal_load(const T* p);
template <>
__device__ __forceinline__ float cal_load<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float cal_load<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// grid (row blocks, S).  part[(s * gridDim.x + blockIdx.x) * 2 + {0, 1}] = (sum ll, sum dll/dtau) of the CTA's rows.
// FOCAL: the per-row term is the focal loss of OutputCalibClassifier.calibrate's default loss_fn
// (fortuna/loss/classification/focal_loss.py:10-36, called from fortuna/output_calib_model/loss.py:26-79):
//   p = softmax(tau z)[y];  loss = -(1 - p)^gamma log p;
//   dloss/dtau = -( (1 - p)^gamma - gamma (1 - p)^(gamma - 1) p log p ) (z_y - sum_c p_c z_c)
template <typename T, bool FOCAL>
__global__ void __launch_bounds__(32 * CAL_WARPS) calib_cls_kernel(const T* __restrict__ z,
                                                                   const int32_t* __restrict__ targets,
                                                                   const float* __restrict__ log_temp, int B, int C,
                                                                   float gamma, double* __restrict__ part) {
  __shared__ double red[CAL_WARPS][2];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int s = blockIdx.y;
  const float tau = log_temp ? expf(-log_temp[0]) : 1.0f;
  double acc_ll = 0.0, acc_d = 0.0;
  for (int b = blockIdx.x * CAL_WARPS + w; b < B; b += gridDim.x * CAL_WARPS) {
    const T* row = z + ((size_t)s * B + b) * C;
    const int y = targets[b];
    float m = -CUDART_INF_F;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, cal_load<T>(row + c));
    m = group_max<32>(m);
    float se = 0.0f, sez = 0.0f;
    for (int c = lane; c < C; c += 32) {
      const float v = cal_load<T>(row + c);
      const float e = __expf(tau * (v - m));  // tau > 0: max_c (tau z) = tau max_c z
      se += e;
      sez = fmaf(e, v, sez);
    }
    se = group_sum<32>(se);
    sez = group_sum<32>(sez);
    if (lane == 0) {
      const bool ok = (unsigned)y < (unsigned)C;
      const float zy = ok ? cal_load<T>(row + y) : 0.0f;
      if constexpr (FOCAL) {
        const float p = ok ? expf(tau * (zy - m)) / se : 0.0f;
        const float logp = logf(p), q = 1.0f - p;
        const float powq = gamma == 2.0f ? q * q : (gamma == 0.0f ? 1.0f : powf(q, gamma));
        const float dpow = gamma == 0.0f ? 0.0f : gamma * (gamma == 2.0f ? q : powf(q, gamma - 1.0f));
        float dl = powq;
        if (dpow != 0.0f && p > 0.0f) dl -= dpow * p * logp;
        acc_ll += (double)(-(powq * logp));
        acc_d += (double)(-dl * (zy - sez / se));
      } else {
        acc_ll += (double)(tau * (zy - m) - __logf(se));
        acc_d += (double)(zy - sez / se);
      }
    }
  }
  if (lane == 0) {
    red[w][0] = acc_ll;
    red[w][1] = acc_d;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a0 = 0.0, a1 = 0.0;
    for (int q = 0; q < CAL_WARPS; ++q) {
      a0 += red[q][0];
      a1 += red[q][1];
    }
    double* o = part + ((size_t)s * gridDim.x + blockIdx.x) * 2;
    o[0] = a0;
    o[1] = a1;
  }
}

__global__ void calib_sum_kernel(const double* __restrict__ part, int S, int nblk, double* __restrict__ out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  double a0 = 0.0, a1 = 0.0;
  for (int q = 0; q < nblk; ++q) {
    a0 += part[((size_t)s * nblk + q) * 2];
    a1 += part[((size_t)s * nblk + q) * 2 + 1];
  }
  out[s] = a0;
  out[S + s] = a1;
}

// Regression temperature scaler (fortuna/output_calibrator/regression.py:7-19: log_var + log_temp): per posterior sample
//   L_s = sum_b -0.5 sum_k ( exp(-(lv + lt)) (y - mu)^2 + lv + lt + log 2 pi ),   dL_s/dlt = sum_b -0.5 sum_k ( 1 - exp(-(lv + lt)) (y - mu)^2 )
// z: [S, B, 2d] = [mu | log var], targets [B, d].  One thread per input, grid (blocks, S).
// (c0, scale) = (log 2 pi, -0.5) gives the log-likelihood above; (0, 1) gives OutputCalibRegressor.calibrate's default
// loss_fn, the variance-scaled squared error sum_k ( exp(-lv') (y - mu)^2 + lv' ) (fortuna/loss/regression/scaled_mse.py:6-26).
__global__ void __launch_bounds__(256) calib_reg_kernel(const float* __restrict__ z, const float* __restrict__ targets,
                                                        const float* __restrict__ log_temp, int B, int d, float c0,
                                                        float scale, double* __restrict__ part) {
  __shared__ double red[8][2];
  const int s = blockIdx.y;
  const float lt = log_temp ? log_temp[0] : 0.0f;
  double acc_ll = 0.0, acc_d = 0.0;
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
    const float* r = z + ((size_t)s * B + b) * 2 * d;
    float a = 0.0f, g = 0.0f;
    for (int k = 0; k < d; ++k) {
      const float lv = r[d + k] + lt;
      const float diff = targets[(size_t)b * d + k] - r[k];
      const float q = expf(-lv) * (diff * diff);
      a += q + lv + c0;
      g += 1.0f - q;
    }
    acc_ll += (double)(scale * a);
    acc_d += (double)(scale * g);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc_ll += __shfl_xor_sync(0xffffffffu, acc_ll, o);
    acc_d += __shfl_xor_sync(0xffffffffu, acc_d, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[threadIdx.x >> 5][0] = acc_ll;
    red[threadIdx.x >> 5][1] = acc_d;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a0 = 0.0, a1 = 0.0;
    for (int q = 0; q < 8; ++q) {
      a0 += red[q][0];
      a1 += red[q][1];
    }
    double* o = part + ((size_t)s * gridDim.x + blockIdx.x) * 2;
    o[0] = a0;
    o[1] = a1;
  }
}

static int calib_reg_blocks(int B) {
  int nb = (B + 255) / 256;
  if (nb > 148 * 4) nb = 148 * 4;
  return nb < 1 ? 1 : nb;
}

static int calib_blocks(int B) {
  int nb = (B + CAL_WARPS - 1) / CAL_WARPS;
  if (nb > 148 * 4) nb = 148 * 4;
  return nb < 1 ? 1 : nb;
}

}  // namespace fb200

using namespace fb200;

extern "C" size_t fb200_calib_cls_workspace_bytes(int S, int B) {
  if (S <= 0 || B <= 0) return 0;
  return (size_t)S * (size_t)calib_blocks(B) * 2 * sizeof(double);
}

template <bool FOCAL>
static int calib_cls_launch(const void* logits, int dtype, const int32_t* targets, const float* log_temp, int S, int B,
                            int C, float gamma, void* workspace, size_t workspace_bytes, double* out, void* stream) {
  if (!logits || !targets || !workspace || !out) return FB200_E_NULL;
  if (S <= 0 || B <= 0 || C <= 0 || S > 65535) return FB200_E_SHAPE;
  if (workspace_bytes < fb200_calib_cls_workspace_bytes(S, B)) return FB200_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = calib_blocks(B);
  double* part = reinterpret_cast<double*>(workspace);
  const dim3 grid((unsigned)nb, (unsigned)S);
  if (dtype == FB200_F32)
    calib_cls_kernel<float, FOCAL><<<grid, 32 * CAL_WARPS, 0, st>>>(reinterpret_cast<const float*>(logits), targets,
                                                                    log_temp, B, C, gamma, part);
  else if (dtype == FB200_BF16)
    calib_cls_kernel<__nv_bfloat16, FOCAL><<<grid, 32 * CAL_WARPS, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(logits), targets, log_temp, B, C, gamma, part);
  else
    return FB200_E_UNSUPPORTED;
  FB200_LAUNCH_CHECK();
  calib_sum_kernel<<<(S + 127) / 128, 128, 0, st>>>(part, S, nb, out);
  FB200_LAUNCH_CHECK();
  return 0;
}

extern "C" int fb200_calib_cls_loss_terms(const void* logits, int dtype, const int32_t* targets, const float* log_temp,
                                          int S, int B, int C, void* workspace, size_t workspace_bytes, double* out,
                                          void* stream) {
  return calib_cls_launch<false>(logits, dtype, targets, log_temp, S, B, C, 0.0f, workspace, workspace_bytes, out,
                                 stream);
}

extern "C" int fb200_output_calib_focal_loss_terms(const void* logits, int dtype, const int32_t* targets,
                                                   const float* log_temp, int B, int C, float gamma, void* workspace,
                                                   size_t workspace_bytes, double* out, void* stream) {
  if (!(gamma >= 0.0f)) return FB200_E_SHAPE;
  return calib_cls_launch<true>(logits, dtype, targets, log_temp, 1, B, C, gamma, workspace, workspace_bytes, out,
                                stream);
}

extern "C" size_t fb200_calib_reg_workspace_bytes(int S, int B) {
  if (S <= 0 || B <= 0) return 0;
  return (size_t)S * (size_t)calib_reg_blocks(B) * 2 * sizeof(double);
}

static int calib_reg_launch(const float* outputs, const float* targets, const float* log_temp, int S, int B, int d,
                            float c0, float scale, void* workspace, size_t workspace_bytes, double* out, void* stream) {
  if (!outputs || !targets || !workspace || !out) return FB200_E_NULL;
  if (S <= 0 || B <= 0 || d <= 0 || S > 65535) return FB200_E_SHAPE;
  if (workspace_bytes < fb200_calib_reg_workspace_bytes(S, B)) return FB200_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = calib_reg_blocks(B);
  double* part = reinterpret_cast<double*>(workspace);
  calib_reg_kernel<<<dim3((unsigned)nb, (unsigned)S), 256, 0, st>>>(outputs, targets, log_temp, B, d, c0, scale, part);
  FB200_LAUNCH_CHECK();
  calib_sum_kernel<<<(S + 127) / 128, 128, 0, st>>>(part, S, nb, out);
  FB200_LAUNCH_CHECK();
  return 0;
}

extern "C" int fb200_calib_reg_loss_terms(const float* outputs, const float* targets, const float* log_temp, int S, int B,
                                          int d, void* workspace, size_t workspace_bytes, double* out, void* stream) {
  return calib_reg_launch(outputs, targets, log_temp, S, B, d, 1.8378770664093453f, -0.5f, workspace, workspace_bytes,
                          out, stream);
}

extern "C" int fb200_output_calib_scaled_mse_terms(const float* outputs, const float* targets, const float* log_temp,
                                                   int B, int d, void* workspace, size_t workspace_bytes, double* out,
                                                   void* stream) {
  return calib_reg_launch(outputs, targets, log_temp, 1, B, d, 0.0f, 1.0f, workspace, workspace_bytes, out, stream);
}
