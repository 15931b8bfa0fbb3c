// Maximum-coverage fixed-precision calibration of a binary classifier ("MaxCovFixPrec").
//
// Replaces the two vmapped objective sweeps of MaxCoverageFixedPrecisionBinaryClassificationCalibrator.calibrate
// (fortuna/conformal/classification/maxcovfixprec_binary_classfication.py:33-96) and its precision / coverage helpers
// (:109-124).  For every candidate tau the reference recomputes, over all n data points,
//   positive side: c = clip((1 + (tau - 1) [p > 0.5]) p, max 1);  b = c >= T_tp;       mean(b), mean(y b)
//   negative side: c =      (1 + (tau - 1) [p < 0.5]) p;          b = c <= 1 - T_tn;   mean(b), mean((1 - y) b)
// i.e. n_taus passes over the data.  Both float32 expressions are monotone in tau, and the candidate grids are sorted
// (ascending on the positive side, descending on the negative side), so b switches from false to true at ONE grid index
// per data point: a binary search with the reference's own float32 expression finds it, a shared-memory histogram
// collects (count, sum of labels) per index, and an inclusive prefix sum over the grid gives every tau's counts - one
// pass over the data, exact integers.
#include <math.h>

#include "common.cuh"

namespace fb200 {

constexpr int MAXCOV_MAX_TAUS = 4096;

__device__ __forceinline__ bool maxcov_pred(float p, float tau, int side, float thr) {
  if (side == 0) {
    const float ind = p > 0.5f ? 1.0f : 0.0f;
    const float u = __fmul_rn(__fadd_rn(1.0f, __fmul_rn(__fadd_rn(tau, -1.0f), ind)), p);
    const float c = (u != u) ? u : fminf(u, 1.0f);  // jnp.minimum keeps a NaN: a NaN probability predicts nothing
    return c >= thr;
  }
  const float ind = p < 0.5f ? 1.0f : 0.0f;
  const float c = __fmul_rn(__fadd_rn(1.0f, __fmul_rn(__fadd_rn(tau, -1.0f), ind)), p);
  return c <= thr;
}

// hist[j] / hist[n_taus + j]: number of points (and of counted labels) whose predicate first holds at grid index j.
__global__ void __launch_bounds__(256) maxcov_hist_kernel(const float* __restrict__ probs,
                                                          const int32_t* __restrict__ targets, int64_t n,
                                                          const float* __restrict__ taus, int n_taus, int side,
                                                          float thr, unsigned int* __restrict__ hist) {
  extern __shared__ unsigned int sh[];  // [2 * n_taus] counters, then the grid
  float* st = reinterpret_cast<float*>(sh + 2 * n_taus);
  for (int j = threadIdx.x; j < 2 * n_taus; j += blockDim.x) sh[j] = 0u;
  for (int j = threadIdx.x; j < n_taus; j += blockDim.x) st[j] = taus[j];
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float p = probs[i];
    if (!maxcov_pred(p, st[n_taus - 1], side, thr)) continue;  // never true on this grid (NaN lands here too)
    int lo = 0, hi = n_taus - 1;                               // invariant: pred(hi) is true
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (maxcov_pred(p, st[mid], side, thr)) hi = mid;
      else lo = mid + 1;
    }
    const int y = targets[i];
    atomicAdd(&sh[lo], 1u);
    const unsigned int lab = side == 0 ? (unsigned int)y : (unsigned int)(1 - y);
    if (lab) atomicAdd(&sh[n_taus + lo], lab);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < 2 * n_taus; j += blockDim.x)
    if (sh[j]) atomicAdd(&hist[j], sh[j]);
}

// Inclusive prefix sums over the grid: counts[j] = #{b true at tau_j}, labels[j] = sum of the counted labels.
__global__ void maxcov_scan_kernel(const unsigned int* __restrict__ hist, int n_taus, int32_t* __restrict__ counts,
                                   int32_t* __restrict__ labels) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  unsigned int c = 0u, l = 0u;
  for (int j = 0; j < n_taus; ++j) {
    c += hist[j];
    l += hist[n_taus + j];
    counts[j] = (int32_t)c;
    labels[j] = (int32_t)l;
  }
}

__global__ void __launch_bounds__(256) maxcov_apply_kernel(const float* __restrict__ probs, int64_t n, float tau_pos,
                                                           float tau_neg, float* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float p = probs[i];
    float q = p;
    if (p > 0.5f) q = fminf(__fmul_rn(tau_pos, p), 1.0f);
    else if (p < 0.5f) q = __fmul_rn(tau_neg, p);
    out[i] = q;
  }
}

}  // namespace fb200

using namespace fb200;

extern "C" size_t fb200_maxcov_workspace_bytes(int n_taus) {
  return n_taus > 0 ? (size_t)2 * n_taus * sizeof(unsigned int) : 0;
}

extern "C" int fb200_maxcov_counts(const float* probs, const int32_t* targets, int64_t n, const float* taus, int n_taus,
                                   int side, float threshold, void* workspace, size_t workspace_bytes, int32_t* counts,
                                   int32_t* labels, void* stream) {
  if (!probs || !targets || !taus || !workspace || !counts || !labels) return FB200_E_NULL;
  if (n <= 0 || n >= 0x7FFFFFFFll || n_taus <= 0 || (side != 0 && side != 1)) return FB200_E_SHAPE;
  if (n_taus > MAXCOV_MAX_TAUS) return FB200_E_UNSUPPORTED;
  if (workspace_bytes < fb200_maxcov_workspace_bytes(n_taus)) return FB200_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned int* hist = reinterpret_cast<unsigned int*>(workspace);
  cudaError_t e = cudaMemsetAsync(hist, 0, fb200_maxcov_workspace_bytes(n_taus), st);
  if (e != cudaSuccess) return (int)e;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  const size_t smem = (size_t)3 * n_taus * sizeof(unsigned int);
  maxcov_hist_kernel<<<(unsigned)blocks, 256, smem, st>>>(probs, targets, n, taus, n_taus, side, threshold, hist);
  FB200_LAUNCH_CHECK();
  maxcov_scan_kernel<<<1, 32, 0, st>>>(hist, n_taus, counts, labels);
  FB200_LAUNCH_CHECK();
  return 0;
}

extern "C" int fb200_maxcov_apply_patches(const float* probs, int64_t n, float tau_pos, float tau_neg, float* out,
                                          void* stream) {
  if (!probs || !out) return FB200_E_NULL;
  if (n <= 0) return FB200_E_SHAPE;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  maxcov_apply_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(probs, n, tau_pos, tau_neg, out);
  FB200_LAUNCH_CHECK();
  return 0;
}
