// K5-K9: conformal / calibration scoring kernels.  Integer outputs (permutations, set masks, sizes, bin
// indices, counts) are bit-exact against oracle/scoring.py; float32 prefix sums follow its sequential order.
//
// Replaces (reference file:line)
//   fortuna/conformal/classification/adaptive_prediction.py:11-22, simple_prediction.py:10-18
//   fortuna/conformal/classification/base.py:45-64, 72-82
//   fortuna/prob_model/predictive/classification.py:465-483   (credible sets)
//   fortuna/metric/classification.py:16-40, 43-95, 98-133, 147-181, 195-220
//   fortuna/conformal/regression/quantile.py:89-90, onedim_uncertainty.py:89-90
#include <math.h>
#include <math_constants.h>

#include "common.cuh"

namespace fb200 {

// Monotone float32 -> uint32 map (total order; -0 is canonicalised to +0 by the caller).
__device__ __forceinline__ uint32_t f32_ordered(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float f32_unordered(uint32_t o) {
  const uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
  return __uint_as_float(u);
}

// ------------------------------------------------------------------------------------------------
// Per-row descending sort + sequential inclusive prefix sum (APS scores, credible sets).
// One CTA per row; keys (ordered(prob) << 32 | class) sorted descending by a shared-memory bitonic
// network: descending prob, ties by larger class index first == argsort(stable ascending)[::-1].
// ------------------------------------------------------------------------------------------------
struct RowSortArgs {
  const float* probs;      // [B,C]
  const int32_t* targets;  // optional [B]
  int B, C, N2;            // N2 = next power of two >= C
  float* scores;           // [B] if targets else [B,C]; may be NULL
  int32_t* perm;           // optional [B,C]
  int32_t* set_size;       // optional [B] (credible sets)
  float set_threshold;     // float32(1 - error)
};

__global__ void __launch_bounds__(256) row_sort_cumsum_kernel(const RowSortArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);
  float* cum = reinterpret_cast<float*>(keys + a.N2);
  const int C = a.C, N2 = a.N2;
  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    const float* row = a.probs + (size_t)b * C;
    for (int i = threadIdx.x; i < N2; i += blockDim.x) {
      unsigned long long k = 0ull;  // padding sorts last
      if (i < C) {
        const float p = row[i] + 0.0f;  // -0 -> +0 (lax.sort treats them as equal)
        k = ((unsigned long long)f32_ordered(p) << 32) | (unsigned)i;
      }
      keys[i] = k;
    }
    __syncthreads();
    for (int k = 2; k <= N2; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int t = threadIdx.x; t < (N2 >> 1); t += blockDim.x) {
          const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
          const int hi = lo | j;
          const bool desc = ((lo & k) == 0);  // final pass (k == N2): all descending
          const unsigned long long x = keys[lo], y = keys[hi];
          if ((x < y) == desc) {
            keys[lo] = y;
            keys[hi] = x;
          }
        }
        __syncthreads();
      }
    }
    // sequential float32 inclusive prefix sum in rank order (oracle/scoring.py defines this order)
    if (threadIdx.x == 0) {
      float run = 0.0f;
      int first = -1;
      const int y = a.targets ? a.targets[b] : -1;
      float ys = CUDART_NAN_F;
      for (int r = 0; r < C; ++r) {
        const unsigned long long k = keys[r];
        const int idx = (int)(k & 0xFFFFFFFFull);
        run += f32_unordered((uint32_t)(k >> 32));
        cum[r] = run;
        if (idx == y) ys = run;
        if (first < 0 && run > a.set_threshold) first = r;
      }
      if (a.scores && a.targets) a.scores[b] = ys;
      if (a.set_size) a.set_size[b] = (first < 0 ? 0 : first) + 1;  // argmax of an all-False row is 0
    }
    __syncthreads();
    if ((a.scores && !a.targets) || a.perm) {
      for (int r = threadIdx.x; r < C; r += blockDim.x) {
        const int idx = (int)(keys[r] & 0xFFFFFFFFull);
        if (a.scores && !a.targets) a.scores[(size_t)b * C + idx] = cum[r];
        if (a.perm) a.perm[(size_t)b * C + r] = idx;
      }
    }
    __syncthreads();
  }
}

static int launch_row_sort(RowSortArgs a, cudaStream_t st) {
  int n2 = 1;
  while (n2 < a.C) n2 <<= 1;
  a.N2 = n2;
  const size_t smem = (size_t)n2 * 8 + (size_t)a.C * 4;
  int threads = n2 / 2;
  if (threads < 32) threads = 32;
  if (threads > 256) threads = 256;
  cudaError_t e =
      cudaFuncSetAttribute(row_sort_cumsum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  if (e != cudaSuccess) return (int)e;
  int grid = a.B;
  if (grid > 148 * 64) grid = 148 * 64;
  row_sort_cumsum_kernel<<<grid, threads, smem, st>>>(a);
  FB200_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) simple_scores_kernel(const float* __restrict__ probs,
                                                            const int32_t* __restrict__ targets, int B, int C,
                                                            float* __restrict__ scores) {
  const int64_t n = targets ? (int64_t)B : (int64_t)B * C;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (targets) {
      const int y = targets[i];
      scores[i] = 1.0f - probs[i * C + y];
    } else {
      scores[i] = 1.0f - probs[i];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Global-memory bitonic sort on the ordered-uint image of the floats (NaNs last, like lax.sort).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sort_load_kernel(const float* __restrict__ x, int64_t n, int64_t n2,
                                                        uint32_t* __restrict__ w) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride)
    w[i] = i < n ? f32_ordered(x[i] + 0.0f) : 0xFFFFFFFFu;
}
__global__ void __launch_bounds__(256) sort_store_kernel(const uint32_t* __restrict__ w, int64_t n,
                                                         float* __restrict__ x) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] = f32_unordered(w[i]);
}
__global__ void __launch_bounds__(256) bitonic_global_step_kernel(uint32_t* __restrict__ w, int64_t n2, int64_t k,
                                                                  int64_t j) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < (n2 >> 1); t += stride) {
    const int64_t lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
    const int64_t hi = lo | j;
    const bool asc = ((lo & k) == 0);
    const uint32_t x = w[lo], y = w[hi];
    if ((x > y) == asc) {
      w[lo] = y;
      w[hi] = x;
    }
  }
}
// All (k, j) steps with j < kSortTile/2... done inside shared memory for one tile of kSortTile keys.
constexpr int kSortTile = 4096;
__global__ void __launch_bounds__(512) bitonic_smem_kernel(uint32_t* __restrict__ w, int64_t n2, int64_t k_first,
                                                           int64_t k_last) {
  // Runs, for each k in [k_first, k_last] (doubling), the steps j = min(k/2, kSortTile/2) .. 1.
  __shared__ uint32_t s[kSortTile];
  const int64_t base = (int64_t)blockIdx.x * kSortTile;
  for (int i = threadIdx.x; i < kSortTile; i += blockDim.x) s[i] = w[base + i];
  __syncthreads();
  for (int64_t k = k_first; k <= k_last; k <<= 1) {
    int64_t j0 = k >> 1;
    if (j0 > kSortTile / 2) j0 = kSortTile / 2;
    for (int64_t j = j0; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < kSortTile / 2; t += blockDim.x) {
        const int lo = (int)(((t & ~(j - 1)) << 1) | (t & (j - 1)));
        const int hi = lo | (int)j;
        const bool asc = (((base + lo) & k) == 0);
        const uint32_t x = s[lo], y = s[hi];
        if ((x > y) == asc) {
          s[lo] = y;
          s[hi] = x;
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < kSortTile; i += blockDim.x) w[base + i] = s[i];
}

// ------------------------------------------------------------------------------------------------
// Rank-count conformal sets: mask = (count(val > t) <= K)  <=>  K >= n_val  or  !(t < val_sorted[n_val-1-K]).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conformal_sets_kernel(const float* __restrict__ val_sorted, int n_val,
                                                             const float* __restrict__ test, int B, int C, int K,
                                                             uint8_t* __restrict__ mask, int32_t* __restrict__ sizes) {
  const int warp = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  const int nwarps = (int)(((int64_t)gridDim.x * blockDim.x) >> 5);
  const bool all_true = K >= n_val;
  const bool all_false = K < 0;
  const float q = (all_true || all_false) ? 0.0f : val_sorted[n_val - 1 - K];
  for (int b = warp; b < B; b += nwarps) {
    const float* row = test + (size_t)b * C;
    uint8_t* mrow = mask + (size_t)b * C;
    int cnt = 0;
    if ((C & 3) == 0) {
      for (int c = lane * 4; c < C; c += 128) {
        const float4 t = *reinterpret_cast<const float4*>(row + c);
        uchar4 m;
        m.x = all_true || (!all_false && !(t.x < q));
        m.y = all_true || (!all_false && !(t.y < q));
        m.z = all_true || (!all_false && !(t.z < q));
        m.w = all_true || (!all_false && !(t.w < q));
        *reinterpret_cast<uchar4*>(mrow + c) = m;
        cnt += m.x + m.y + m.z + m.w;
      }
    } else {
      for (int c = lane; c < C; c += 32) {
        const uint8_t m = all_true || (!all_false && !(row[c] < q));
        mrow[c] = m;
        cnt += m;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (sizes && lane == 0) sizes[b] = cnt;
  }
}

// ------------------------------------------------------------------------------------------------
// ECE / MCE binning
// ------------------------------------------------------------------------------------------------
constexpr int kMaxBins = 32;

__global__ void __launch_bounds__(256) ece_bins_kernel(const float* __restrict__ probs, int B, int C,
                                                       const int32_t* __restrict__ preds,
                                                       const int32_t* __restrict__ targets,
                                                       const float* __restrict__ thresholds, int n_bins,
                                                       int32_t* __restrict__ bin_idx, int32_t* __restrict__ counts,
                                                       float* __restrict__ conf_sums, int32_t* __restrict__ correct,
                                                       float* __restrict__ brier_sum) {
  __shared__ float thr[kMaxBins];
  __shared__ int s_counts[kMaxBins];
  __shared__ int s_correct[kMaxBins];
  __shared__ float s_conf[kMaxBins];
  __shared__ float s_brier;
  __shared__ int s_ncorrect;
  if (threadIdx.x < kMaxBins) {
    thr[threadIdx.x] = threadIdx.x < n_bins ? thresholds[threadIdx.x] : CUDART_INF_F;
    s_counts[threadIdx.x] = 0;
    s_correct[threadIdx.x] = 0;
    s_conf[threadIdx.x] = 0.0f;
  }
  if (threadIdx.x == 0) {
    s_brier = 0.0f;
    s_ncorrect = 0;
  }
  __syncthreads();
  const int warp = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  const int nwarps = (int)(((int64_t)gridDim.x * blockDim.x) >> 5);
  for (int b0 = warp * 32; b0 < B; b0 += nwarps * 32) {
    // each warp handles 32 consecutive rows; lane `r` ends up owning row b0 + r
    float my_conf = 0.0f;
    int my_pred = 0;
    float my_brier = 0.0f;
    if (C == 1) {
      const int b = b0 + lane;
      if (b < B) my_conf = probs[b];
    } else {
      for (int r = 0; r < 32 && b0 + r < B; ++r) {
        const float* row = probs + (size_t)(b0 + r) * C;
        const int y = targets ? targets[b0 + r] : -1;
        float best = -CUDART_INF_F;
        int best_c = 0x7FFFFFFF;
        float sq = 0.0f;
        for (int c = lane; c < C; c += 32) {
          const float p = row[c];
          if (p > best) {
            best = p;
            best_c = c;
          }
          const float d = p - (c == y ? 1.0f : 0.0f);
          sq = fmaf(d, d, sq);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, best, o);
          const int oc = __shfl_xor_sync(0xffffffffu, best_c, o);
          sq += __shfl_xor_sync(0xffffffffu, sq, o);
          if (ov > best || (ov == best && oc < best_c)) {
            best = ov;
            best_c = oc;
          }
        }
        if (lane == r) {
          my_conf = best;
          my_pred = best_c == 0x7FFFFFFF ? 0 : best_c;
          my_brier = sq;
        }
      }
    }
    const int b = b0 + lane;
    if (b < B) {
      const int pred = preds ? preds[b] : my_pred;
      int bin = -1;
      for (int i = 0; i < n_bins; ++i) {
        if (my_conf <= thr[i]) {
          bin = i;
          break;
        }
      }
      if (bin_idx) bin_idx[b] = bin;
      if (bin >= 0) {
        atomicAdd(&s_counts[bin], 1);
        atomicAdd(&s_conf[bin], my_conf);
        if (targets && targets[b] == pred) atomicAdd(&s_correct[bin], 1);
      }
      if (targets && targets[b] == pred) atomicAdd(&s_ncorrect, 1);  // accuracy counts every row
      if (C > 1 && targets) atomicAdd(&s_brier, my_brier);
    }
  }
  __syncthreads();
  if (threadIdx.x < n_bins) {
    if (s_counts[threadIdx.x]) atomicAdd(&counts[threadIdx.x], s_counts[threadIdx.x]);
    if (s_correct[threadIdx.x]) atomicAdd(&correct[threadIdx.x], s_correct[threadIdx.x]);
    if (s_counts[threadIdx.x]) atomicAdd(&conf_sums[threadIdx.x], s_conf[threadIdx.x]);
  }
  if (threadIdx.x == 0) {
    // scratch layout of `result` until ece_finalize_kernel runs: [0] = float brier sum, [1] = int #correct
    if (s_brier != 0.0f) atomicAdd(&brier_sum[0], s_brier);
    if (s_ncorrect) atomicAdd(reinterpret_cast<int*>(&brier_sum[1]), s_ncorrect);
  }
}

__global__ void ece_finalize_kernel(const int32_t* __restrict__ counts, const float* __restrict__ conf_sums,
                                    const int32_t* __restrict__ correct, int n_bins, int B, int C,
                                    float* __restrict__ result) {
  // result[0..3] = ECE, MCE, accuracy, brier; on entry result[0] holds the brier sum and result[1] the
  // (int) number of correct predictions
  const float brier_sum = result[0];
  const int n_correct = reinterpret_cast<const int*>(result)[1];
  float ece = 0.0f, mce = -CUDART_INF_F;
  for (int i = 0; i < n_bins; ++i) {
    const float cnt = (float)counts[i];
    const float acc = counts[i] ? (float)correct[i] / cnt : 0.0f;    // nan_to_num(mean(empty)) = 0
    const float conf = counts[i] ? conf_sums[i] / cnt : 0.0f;
    const float w = cnt * fabsf(acc - conf);
    ece += w;
    mce = fmaxf(mce, w);
  }
  result[0] = ece / (float)B;
  result[1] = mce;
  result[2] = (float)n_correct / (float)B;
  result[3] = C > 1 ? brier_sum / (float)B : CUDART_NAN_F;
}

__global__ void quantile_sorted_kernel(const float* __restrict__ x, int64_t n, const float* __restrict__ q, int nq,
                                       float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  // jnp.quantile, method="linear": pos = q (n-1); weights from the unclamped floor; indices clamped
  const float pos = __fmul_rn(q[i], (float)(n - 1));
  const float low = floorf(pos), high = ceilf(pos);
  const float hw = __fsub_rn(pos, low);
  const float lw = __fsub_rn(1.0f, hw);
  const float nm1 = (float)(n - 1);
  const int64_t lo = (int64_t)fmaxf(0.0f, fminf(low, nm1));
  const int64_t hi = (int64_t)fmaxf(0.0f, fminf(high, nm1));
  const float last = x[n - 1];
  if (last != last || pos != pos) {  // NaNs sort last; any NaN poisons the quantile
    out[i] = CUDART_NAN_F;
    return;
  }
  out[i] = __fadd_rn(__fmul_rn(x[lo], lw), __fmul_rn(x[hi], hw));
}

}  // namespace fb200

using namespace fb200;

extern "C" int fb200_aps_scores(const float* probs, const int32_t* targets, int B, int C, float* scores,
                                int32_t* perm_out, void* stream) {
  if (!probs || (!scores && !perm_out)) return FB200_E_NULL;
  if (B <= 0 || C <= 0) return FB200_E_SHAPE;
  if (C > 4096) return FB200_E_UNSUPPORTED;
  RowSortArgs a{};
  a.probs = probs;
  a.targets = targets;
  a.B = B;
  a.C = C;
  a.scores = scores;
  a.perm = perm_out;
  a.set_size = nullptr;
  a.set_threshold = INFINITY;
  return launch_row_sort(a, (cudaStream_t)stream);
}

extern "C" int fb200_credible_sets(const float* mean, int B, int C, float threshold, int32_t* labels,
                                   int32_t* set_size, void* stream) {
  if (!mean || !labels || !set_size) return FB200_E_NULL;
  if (B <= 0 || C <= 0) return FB200_E_SHAPE;
  if (C > 4096) return FB200_E_UNSUPPORTED;
  RowSortArgs a{};
  a.probs = mean;
  a.B = B;
  a.C = C;
  a.perm = labels;
  a.set_size = set_size;
  a.set_threshold = threshold;
  return launch_row_sort(a, (cudaStream_t)stream);
}

extern "C" int fb200_simple_scores(const float* probs, const int32_t* targets, int B, int C, float* scores,
                                   void* stream) {
  if (!probs || !scores) return FB200_E_NULL;
  if (B <= 0 || C <= 0) return FB200_E_SHAPE;
  const int64_t n = targets ? (int64_t)B : (int64_t)B * C;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  simple_scores_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(probs, targets, B, C, scores);
  FB200_LAUNCH_CHECK();
  return 0;
}

static int64_t sort_pow2(int64_t n) {
  int64_t n2 = kSortTile;
  while (n2 < n) n2 <<= 1;
  return n2;
}

extern "C" size_t fb200_sort_workspace_bytes(int64_t n) {
  if (n <= 0) return 0;
  return (size_t)sort_pow2(n) * sizeof(uint32_t);
}

extern "C" int fb200_sort_f32(float* x, int64_t n, void* workspace, size_t workspace_bytes, void* stream) {
  if (!x || !workspace) return FB200_E_NULL;
  if (n <= 0 || n > (1ll << 26)) return FB200_E_SHAPE;
  const int64_t n2 = sort_pow2(n);
  if (workspace_bytes < (size_t)n2 * sizeof(uint32_t)) return FB200_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  uint32_t* w = reinterpret_cast<uint32_t*>(workspace);
  int64_t blocks = (n2 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  sort_load_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, n, n2, w);
  FB200_LAUNCH_CHECK();
  const unsigned tiles = (unsigned)(n2 / kSortTile);
  // every k up to the tile size entirely in shared memory
  bitonic_smem_kernel<<<tiles, 512, 0, st>>>(w, n2, 2, kSortTile);
  FB200_LAUNCH_CHECK();
  for (int64_t k = (int64_t)kSortTile * 2; k <= n2; k <<= 1) {
    for (int64_t j = k >> 1; j >= kSortTile; j >>= 1) {
      int64_t gb = ((n2 >> 1) + 255) / 256;
      if (gb > 148 * 16) gb = 148 * 16;
      bitonic_global_step_kernel<<<(unsigned)gb, 256, 0, st>>>(w, n2, k, j);
      FB200_LAUNCH_CHECK();
    }
    bitonic_smem_kernel<<<tiles, 512, 0, st>>>(w, n2, k, k);
    FB200_LAUNCH_CHECK();
  }
  sort_store_kernel<<<(unsigned)blocks, 256, 0, st>>>(w, n, x);
  FB200_LAUNCH_CHECK();
  return 0;
}

extern "C" int fb200_conformal_sets(const float* val_scores_sorted, int n_val, const float* test_scores, int B,
                                    int C, float threshold, uint8_t* mask, int32_t* sizes, void* stream) {
  if (!val_scores_sorted || !test_scores || !mask) return FB200_E_NULL;
  if (n_val <= 0 || B <= 0 || C <= 0) return FB200_E_SHAPE;
  if ((C & 3) == 0 && (!fb200_aligned16(test_scores) || (reinterpret_cast<uintptr_t>(mask) & 3u))) return FB200_E_ALIGN;
  // K = largest integer count with float32(count) < threshold (count is compared as float32, base.py:51-53)
  long long K;
  if (!(threshold > 0.0f)) {
    K = -1;
  } else if (threshold > 4.0e9f) {
    K = n_val;
  } else {
    K = (long long)ceil((double)threshold) - 1;
    while ((float)(K + 1) < threshold) ++K;
    while (K >= 0 && !((float)K < threshold)) --K;
    if (K > n_val) K = n_val;
  }
  int64_t blocks = ((int64_t)B * 32 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  conformal_sets_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(val_scores_sorted, n_val, test_scores,
                                                                            B, C, (int)K, mask, sizes);
  FB200_LAUNCH_CHECK();
  return 0;
}

extern "C" int fb200_ece_bins(const float* probs, int B, int C, const int32_t* preds, const int32_t* targets,
                              const float* thresholds, int n_bins, int32_t* bin_idx, int32_t* counts,
                              float* conf_sums, int32_t* correct, float* result, void* stream) {
  if (!probs || !thresholds || !counts || !conf_sums || !correct || !result) return FB200_E_NULL;
  if (B <= 0 || C <= 0 || n_bins <= 0) return FB200_E_SHAPE;
  if (n_bins > kMaxBins) return FB200_E_UNSUPPORTED;
  if (C == 1 && !preds) return FB200_E_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  if ((e = cudaMemsetAsync(counts, 0, sizeof(int32_t) * n_bins, st)) != cudaSuccess) return (int)e;
  if ((e = cudaMemsetAsync(correct, 0, sizeof(int32_t) * n_bins, st)) != cudaSuccess) return (int)e;
  if ((e = cudaMemsetAsync(conf_sums, 0, sizeof(float) * n_bins, st)) != cudaSuccess) return (int)e;
  if ((e = cudaMemsetAsync(result, 0, sizeof(float) * 4, st)) != cudaSuccess) return (int)e;
  int64_t blocks = ((int64_t)B + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  ece_bins_kernel<<<(unsigned)blocks, 256, 0, st>>>(probs, B, C, preds, targets, thresholds, n_bins, bin_idx,
                                                    counts, conf_sums, correct, result);
  FB200_LAUNCH_CHECK();
  ece_finalize_kernel<<<1, 1, 0, st>>>(counts, conf_sums, correct, n_bins, B, C, result);
  FB200_LAUNCH_CHECK();
  return 0;
}

extern "C" int fb200_quantile_sorted(const float* x_sorted, int64_t n, const float* q, int nq, float* out,
                                     void* stream) {
  if (!x_sorted || !q || !out) return FB200_E_NULL;
  if (n <= 0 || nq <= 0) return FB200_E_SHAPE;
  quantile_sorted_kernel<<<(nq + 127) / 128, 128, 0, (cudaStream_t)stream>>>(x_sorted, n, q, nq, out);
  FB200_LAUNCH_CHECK();
  return 0;
}
