// K2 (CUDA-core path) + dispatch: last-layer logits for S posterior samples.
//
// Replaces Likelihood._get_batched_calibrated_outputs (fortuna/likelihood/base.py:429-458) restricted to
// the final nn.Dense (fortuna/model/mlp.py:227, lenet.py:129, resnet.py:248) and the temperature scaler
// (fortuna/output_calibrator/classification.py:14-17):
//     out[s,b,c] = (sum_d x[b,d] w[s,d,c] + bias[s,c]) * exp(-log_temp)
// The reference runs S full forward passes; here the features x are computed once by the caller and only
// this contraction is repeated per sample.  Shapes whose D*C is tiny (two-moons 30x2, LeNet 84x10, BERT
// 768x2) are bandwidth/latency bound and run on this shared-memory-tiled FFMA kernel; large bf16 problems
// go to the tcgen05 kernel in last_layer_tc.cu.
#include <cuda_bf16.h>

#include "common.cuh"

namespace fb200 {

int last_layer_tc_supported(int S, int B, int D, int C, int in_dtype, int out_dtype, int w_layout);
int last_layer_tc_launch(const void* x, const void* w, const float* bias, const float* log_temp,
                         const int32_t* sample_idx, int n_stored, void* out, int S, int B, int D, int C,
                         int out_dtype, cudaStream_t st);

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ void store_out(float* p, float v) { *p = v; }
__device__ __forceinline__ void store_out(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

constexpr int kBM = 64, kBN = 64, kBK = 16;

template <typename TI, typename TO, int WL>
__global__ void __launch_bounds__(256) last_layer_simt_kernel(const TI* __restrict__ x, const TI* __restrict__ w,
                                                              const float* __restrict__ bias,
                                                              const float* __restrict__ log_temp,
                                                              const int32_t* __restrict__ sample_idx,
                                                              TO* __restrict__ out, int S, int B, int D, int C) {
  __shared__ float As[kBK][kBM + 4];
  __shared__ float Bs[kBK][kBN + 4];
  const int s = blockIdx.z;
  const int b0 = blockIdx.y * kBM;
  const int c0 = blockIdx.x * kBN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int sw = sample_idx ? sample_idx[s] : s;  // which stored posterior sample this output row uses
  const TI* ws = w + (size_t)sw * D * C;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  for (int k0 = 0; k0 < D; k0 += kBK) {
    {  // A tile: 64 rows x 16 k
      const int r = threadIdx.x >> 2, kk = (threadIdx.x & 3) * 4;
      const int b = b0 + r;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int k = k0 + kk + q;
        As[kk + q][r] = (b < B && k < D) ? to_f32(x[(size_t)b * D + k]) : 0.0f;
      }
    }
    if (WL == FB200_W_SDC) {  // w[s][d][c]: 16 k x 64 c, c contiguous
      const int kk = threadIdx.x >> 4, cc = (threadIdx.x & 15) * 4;
      const int k = k0 + kk;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int c = c0 + cc + q;
        Bs[kk][cc + q] = (k < D && c < C) ? to_f32(ws[(size_t)k * C + c]) : 0.0f;
      }
    } else {  // w[s][c][d]: d contiguous
      const int cc = threadIdx.x >> 2, kk = (threadIdx.x & 3) * 4;
      const int c = c0 + cc;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int k = k0 + kk + q;
        Bs[kk + q][cc] = (k < D && c < C) ? to_f32(ws[(size_t)c * D + k]) : 0.0f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const float scale = log_temp ? expf(-log_temp[0]) : 1.0f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int b = b0 + ty * 4 + i;
    if (b >= B) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + tx * 4 + j;
      if (c >= C) continue;
      float v = acc[i][j];
      if (bias) v += bias[(size_t)sw * C + c];
      if (log_temp) v *= scale;
      store_out(out + ((size_t)s * B + b) * C + c, v);
    }
  }
}

// wt[s][c][d] (bf16) = w[s][d][c]
template <typename TI>
__global__ void __launch_bounds__(256) transpose_w_kernel(const TI* __restrict__ w, __nv_bfloat16* __restrict__ wt,
                                                          int D, int C) {
  __shared__ float tile[32][33];
  const int s = blockIdx.z;
  const int d0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const TI* ws = w + (size_t)s * D * C;
  __nv_bfloat16* wts = wt + (size_t)s * D * C;
  for (int i = ty; i < 32; i += 8) {
    const int d = d0 + i, c = c0 + tx;
    tile[i][tx] = (d < D && c < C) ? to_f32(ws[(size_t)d * C + c]) : 0.0f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, d = d0 + tx;
    if (c < C && d < D) wts[(size_t)c * D + d] = __float2bfloat16_rn(tile[tx][i]);
  }
}

template <typename TI, typename TO>
static int launch_simt(const void* x, const void* w, const float* bias, const float* log_temp,
                       const int32_t* sample_idx, void* out, int S, int B, int D, int C, int w_layout,
                       cudaStream_t st) {
  dim3 grid((C + kBN - 1) / kBN, (B + kBM - 1) / kBM, S);
  if (grid.y > 65535 || grid.z > 65535) return FB200_E_UNSUPPORTED;
  if (w_layout == FB200_W_SDC)
    last_layer_simt_kernel<TI, TO, FB200_W_SDC><<<grid, 256, 0, st>>>(
        reinterpret_cast<const TI*>(x), reinterpret_cast<const TI*>(w), bias, log_temp, sample_idx,
        reinterpret_cast<TO*>(out), S, B, D, C);
  else
    last_layer_simt_kernel<TI, TO, FB200_W_SCD><<<grid, 256, 0, st>>>(
        reinterpret_cast<const TI*>(x), reinterpret_cast<const TI*>(w), bias, log_temp, sample_idx,
        reinterpret_cast<TO*>(out), S, B, D, C);
  FB200_LAUNCH_CHECK();
  return 0;
}

}  // namespace fb200

using namespace fb200;

extern "C" int fb200_last_layer_logits(const void* x, const void* w, const float* bias, const float* log_temp,
                                       const int32_t* sample_idx, int n_stored, void* out, int S, int B, int D,
                                       int C, int in_dtype, int out_dtype, int w_layout, int path, void* stream) {
  if (!x || !w || !out) return FB200_E_NULL;
  if (S <= 0 || B <= 0 || D <= 0 || C <= 0 || n_stored <= 0) return FB200_E_SHAPE;
  if (!sample_idx && n_stored != S) return FB200_E_SHAPE;
  if ((in_dtype != FB200_F32 && in_dtype != FB200_BF16) || (out_dtype != FB200_F32 && out_dtype != FB200_BF16))
    return FB200_E_UNSUPPORTED;
  if (w_layout != FB200_W_SDC && w_layout != FB200_W_SCD) return FB200_E_UNSUPPORTED;
  if (path < 0 || path > 2) return FB200_E_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int tc_ok = last_layer_tc_supported(S, B, D, C, in_dtype, out_dtype, w_layout);
  if (path == 2 && !tc_ok) return FB200_E_UNSUPPORTED;
  // auto: tensor cores only when the contraction is large enough to be compute-bound (D >= 256, C >= 64)
  const bool use_tc = path == 2 || (path == 0 && tc_ok && D >= 256 && C >= 64 && B >= 128);
  if (use_tc) return last_layer_tc_launch(x, w, bias, log_temp, sample_idx, n_stored, out, S, B, D, C, out_dtype, st);
  if (in_dtype == FB200_F32) {
    if (out_dtype == FB200_F32) return launch_simt<float, float>(x, w, bias, log_temp, sample_idx, out, S, B, D, C, w_layout, st);
    return launch_simt<float, __nv_bfloat16>(x, w, bias, log_temp, sample_idx, out, S, B, D, C, w_layout, st);
  }
  if (out_dtype == FB200_F32)
    return launch_simt<__nv_bfloat16, float>(x, w, bias, log_temp, sample_idx, out, S, B, D, C, w_layout, st);
  return launch_simt<__nv_bfloat16, __nv_bfloat16>(x, w, bias, log_temp, sample_idx, out, S, B, D, C, w_layout, st);
}

extern "C" int fb200_transpose_w_bf16(const void* w, int in_dtype, void* wt, int S, int D, int C, void* stream) {
  if (!w || !wt) return FB200_E_NULL;
  if (S <= 0 || D <= 0 || C <= 0) return FB200_E_SHAPE;
  dim3 grid((C + 31) / 32, (D + 31) / 32, S);
  if (grid.y > 65535 || grid.z > 65535) return FB200_E_UNSUPPORTED;
  if (in_dtype == FB200_F32)
    transpose_w_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float*>(w),
                                                                      reinterpret_cast<__nv_bfloat16*>(wt), D, C);
  else if (in_dtype == FB200_BF16)
    transpose_w_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(w), reinterpret_cast<__nv_bfloat16*>(wt), D, C);
  else
    return FB200_E_UNSUPPORTED;
  FB200_LAUNCH_CHECK();
  return 0;
}
