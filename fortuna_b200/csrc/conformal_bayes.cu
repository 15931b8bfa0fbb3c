// Conformal Bayes sets by importance sampling over posterior samples.
//
// Replaces ClassificationPredictive.conformal_set, fortuna/prob_model/predictive/classification.py:485-626
// (`_compute_cb_region_importancesampling` :557-587 and `_diagnose_importancesampling_weights` :601-617).
//
// For every test input t and candidate class c (a "row" r = t * C + c):
//   w[s]         = exp(log p(y = c | x_t, theta_s))                       importance weights          (:562)
//   Z            = sum_s w[s]                                                                          (:563)
//   p_train[i]   = sum_s (w[s] / Z) * exp(log p(y_i | x_i, theta_s))       [R, S] x [S, n_train]       (:566-569)
//   p_test       = sum_s w[s]^2 / Z                                                                    (:570-573)
//   rank         = #{i : p_train[i] <= p_test} + 1                          (the test point itself)    (:576-579)
//   region[t,c]  = float32(rank) > float32(error * (n_train + 1))                                      (:582)
// The reference materialises p_train ([C, n_train] per test point, lax.map over test points); here the contraction is
// tiled like a GEMM on the CUDA cores (K = S is 30-256: far too short for tensor cores to matter, and the comparison
// against p_test needs float32 products) and its epilogue only COUNTS - p_train never exists in memory.
#include <math.h>
#include <math_constants.h>

#include "common.cuh"

namespace fb200 {

constexpr int CB_T = 64;   // rows / columns per tile
constexpr int CB_KC = 32;  // posterior samples per staged chunk of the training matrix
constexpr int CB_THREADS = 256;

// Row preparation: a[s][r] = w[s] / Z (K-major for the tiles), p_test[r], optional ESS[r] = 1 / sum_s (w[s]/Z)^2 via the
// reference's logsumexp route.  One thread per row; lsm is [S, n_test, C] so consecutive rows are consecutive floats.
__global__ void __launch_bounds__(256) cb_rows_kernel(const float* __restrict__ lsm, int S, int64_t R,
                                                      float* __restrict__ a, float* __restrict__ p_test,
                                                      float* __restrict__ ess) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  float z = 0.0f, z2 = 0.0f, mx = -INFINITY;
  for (int s = 0; s < S; ++s) {
    const float lw = lsm[(int64_t)s * R + r];
    const float w = expf(lw);
    z += w;
    z2 += w * w;
    mx = fmaxf(mx, lw);
  }
  for (int s = 0; s < S; ++s) a[(int64_t)s * R + r] = expf(lsm[(int64_t)s * R + r]) / z;
  p_test[r] = z2 / z;
  if (ess) {
    float se = 0.0f;
    for (int s = 0; s < S; ++s) se += expf(lsm[(int64_t)s * R + r] - mx);
    const float lse = mx + logf(se);
    float q = 0.0f;
    for (int s = 0; s < S; ++s) {
      const float iw = expf(lsm[(int64_t)s * R + r] - lse);
      q += iw * iw;
    }
    ess[r] = 1.0f / q;
  }
}

struct CbArgs {
  const float* a;          // [S, R]
  const float* ell_train;  // [S, n_train] log-likelihoods of the training points under each drawn sample
  const float* p_test;     // [R]
  int S;
  int64_t R;
  int n_train;
  int col_tiles_per_split;
  int32_t* counts;  // [R], zero-initialised; receives #{i : p_train[i] <= p_test}
};

// grid (row tiles, column splits).  The CTA keeps its 64 rows of `a` (all S samples) in shared memory and walks its
// column tiles; per tile the training log-likelihoods are staged 32 samples at a time with the exp applied on the way in.
__global__ void __launch_bounds__(CB_THREADS) cb_count_kernel(const CbArgs p) {
  extern __shared__ __align__(16) float cb_smem[];
  float* sa = cb_smem;                              // [S][CB_T + 4]
  float* se = cb_smem + (size_t)p.S * (CB_T + 4);   // [CB_KC][CB_T + 4]
  const int tid = threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.x * CB_T;
  const int pi = (tid / 16) * 4, pj = (tid % 16) * 4;
  for (int e = tid; e < p.S * CB_T; e += CB_THREADS) {
    const int s = e / CB_T, j = e % CB_T;
    sa[(size_t)s * (CB_T + 4) + j] = (r0 + j < p.R) ? p.a[(int64_t)s * p.R + r0 + j] : 0.0f;
  }
  float pt[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) pt[i] = (r0 + pi + i < p.R) ? p.p_test[r0 + pi + i] : -INFINITY;
  int cnt[4] = {0, 0, 0, 0};
  const int n_col_tiles = (p.n_train + CB_T - 1) / CB_T;
  const int ct0 = blockIdx.y * p.col_tiles_per_split;
  const int ct1 = min(n_col_tiles, ct0 + p.col_tiles_per_split);
  for (int ct = ct0; ct < ct1; ++ct) {
    const int c0 = ct * CB_T;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    for (int k0 = 0; k0 < p.S; k0 += CB_KC) {
      __syncthreads();  // previous chunk consumed (and, first time round, `sa` complete)
      for (int e = tid; e < CB_KC * CB_T; e += CB_THREADS) {
        const int k = e / CB_T, j = e % CB_T;
        const int s = k0 + k, i = c0 + j;
        se[(size_t)k * (CB_T + 4) + j] = (s < p.S && i < p.n_train) ? expf(p.ell_train[(int64_t)s * p.n_train + i]) : 0.0f;
      }
      __syncthreads();
      const int kn = min(CB_KC, p.S - k0);
      for (int k = 0; k < kn; ++k) {
        const float4 a4 = *reinterpret_cast<const float4*>(&sa[(size_t)(k0 + k) * (CB_T + 4) + pi]);
        const float4 e4 = *reinterpret_cast<const float4*>(&se[(size_t)k * (CB_T + 4) + pj]);
        const float av[4] = {a4.x, a4.y, a4.z, a4.w}, ev[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], ev[j], acc[i][j]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c0 + pj + j < p.n_train && acc[i][j] <= pt[i]) ++cnt[i];
  }
  // the 16 threads of a half-warp share their 4 rows
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int v = cnt[i];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 15) == 0 && r0 + pi + i < p.R && v) atomicAdd(&p.counts[r0 + pi + i], v);
  }
}

// region[t, c] = float32(count + 1) > threshold; sizes[t] = number of classes in the set.  One thread per test point.
__global__ void __launch_bounds__(256) cb_region_kernel(const int32_t* __restrict__ counts, int n_test, int C,
                                                        float threshold, uint8_t* __restrict__ region,
                                                        int32_t* __restrict__ sizes) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_test) return;
  int sz = 0;
  for (int c = 0; c < C; ++c) {
    const bool in = (float)(counts[(int64_t)t * C + c] + 1) > threshold;
    region[(int64_t)t * C + c] = in ? 1 : 0;
    sz += in ? 1 : 0;
  }
  if (sizes) sizes[t] = sz;
}

}  // namespace fb200

using namespace fb200;

extern "C" size_t fb200_conformal_bayes_workspace_bytes(int S, int n_test, int C) {
  if (S <= 0 || n_test <= 0 || C <= 0) return 0;
  const size_t R = (size_t)n_test * C;
  return (size_t)S * R * sizeof(float) + R * sizeof(float) + R * sizeof(int32_t);
}

extern "C" int fb200_conformal_bayes_sets(const float* ell_train, const float* lsm_test, int S, int n_train, int n_test,
                                          int C, float threshold, void* workspace, size_t workspace_bytes,
                                          uint8_t* region, int32_t* sizes, int32_t* rank_counts, float* ess,
                                          void* stream) {
  if (!ell_train || !lsm_test || !workspace || !region) return FB200_E_NULL;
  if (S <= 0 || n_train <= 0 || n_test <= 0 || C <= 0) return FB200_E_SHAPE;
  if (S > 768) return FB200_E_UNSUPPORTED;  // the row tile of importance weights must fit shared memory
  if (workspace_bytes < fb200_conformal_bayes_workspace_bytes(S, n_test, C)) return FB200_E_WORKSPACE;
  if (!fb200_aligned16(workspace)) return FB200_E_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t R = (int64_t)n_test * C;
  float* a = reinterpret_cast<float*>(workspace);
  float* p_test = a + (size_t)S * R;
  int32_t* counts = rank_counts ? rank_counts : reinterpret_cast<int32_t*>(p_test + R);
  cudaError_t e = cudaMemsetAsync(counts, 0, (size_t)R * sizeof(int32_t), st);
  if (e != cudaSuccess) return (int)e;
  cb_rows_kernel<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(lsm_test, S, R, a, p_test, ess);
  FB200_LAUNCH_CHECK();
  CbArgs p{};
  p.a = a;
  p.ell_train = ell_train;
  p.p_test = p_test;
  p.S = S;
  p.R = R;
  p.n_train = n_train;
  p.counts = counts;
  const int64_t row_tiles = (R + CB_T - 1) / CB_T;
  const int n_col_tiles = (n_train + CB_T - 1) / CB_T;
  // enough CTAs to fill the GPU: split the columns when there are few row tiles
  int splits = (int)((148 * 4 + row_tiles - 1) / row_tiles);
  if (splits > n_col_tiles) splits = n_col_tiles;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  p.col_tiles_per_split = (n_col_tiles + splits - 1) / splits;
  splits = (n_col_tiles + p.col_tiles_per_split - 1) / p.col_tiles_per_split;
  if (row_tiles > 0x7FFFFFFFll) return FB200_E_UNSUPPORTED;
  const size_t smem = ((size_t)S + CB_KC) * (CB_T + 4) * sizeof(float);
  e = cudaFuncSetAttribute(cb_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  cb_count_kernel<<<dim3((unsigned)row_tiles, (unsigned)splits), CB_THREADS, smem, st>>>(p);
  FB200_LAUNCH_CHECK();
  cb_region_kernel<<<(unsigned)((n_test + 255) / 256), 256, 0, st>>>(counts, n_test, C, threshold, region, sizes);
  FB200_LAUNCH_CHECK();
  return 0;
}
