// Multivalid scorers: the per-round statistics of the iterative multivalid methods
// (fortuna/conformal/multivalid/iterative/base.py:34-333 - BatchMVP and the multicalibrators) in ONE pass over the data.
//
// A round of the reference evaluates, for every bucket type tau, group g, bucket value v_j and score dimension c, the set
//   b_i = cond_tau(x_i, v_j) * groups[i, g] (* [argmax_k values[i, k] == c] for the top-label multicalibrator)
//     tau 0 "=":  |x_i - v_j| < h          tau 1 ">=":  x_i + h >= v_j          tau 2 "<=":  x_i - h <= v_j
//   (iterative/base.py:346-364; x_i = values[i] or values[i, c], h = float32(0.5 / n_buckets), float32 arithmetic)
// and from it  prob_b = mean(b),  mean(scores b),  mean(values b)  (multicalibrator.py:54-100) or
// mean((scores <= values) b)  (batch_mvp.py:166-192) - as a vmap over (tau, g, v, c) that materialises
// [n_taus, G, n_buckets, D, n] booleans every round.  Here a CTA stages a chunk of points in shared memory once and every
// thread owns a few (tau, j) bins whose sums it keeps in registers: no atomics in the inner loop, the predicates are the
// reference's own float32 expressions evaluated per (point, bucket) - no boundary-index shortcuts to get wrong.
// Counts are integers (order-free atomics); the float sums go through per-chunk partials added in chunk order, so the
// statistics - and with them the argmax that picks each round's patch - are the same bits run to run.
#include <math.h>
#include <math_constants.h>

#include "common.cuh"

namespace fb200 {

constexpr int MV_CHUNK = 1024;
constexpr int MV_THREADS = 128;

struct MvArgs {
  const float* values;   // [n, Dv], Dv = 1 or D
  const float* scores;   // [n, D]
  const uint8_t* groups; // [n, G] (0 / 1)
  const float* buckets;  // [n_buckets]
  int n, D, Dv, G, n_buckets;
  int top_label;         // b also requires argmax_k values[i, k] == c  (Dv == D)
  int mode;              // 0: stat = score (multicalibrators); 1: stat = [score <= value] (BatchMVP)
  float h;               // float32(0.5 / n_buckets)
  int tau_mask;          // bit tau set: that bucket type is wanted
};

__device__ __forceinline__ bool mv_cond(int tau, float x, float v, float h) {
  if (tau == 0) return fabsf(__fsub_rn(x, v)) < h;
  if (tau == 1) return __fadd_rn(x, h) >= v;
  return __fsub_rn(x, h) <= v;
}

__device__ __forceinline__ int mv_argmax(const float* row, int D) {
  int k = 0;
  float m = row[0];
  for (int j = 1; j < D; ++j)
    if (row[j] > m) m = row[j], k = j;  // first maximum, like jnp.argmax
  return k;
}

// grid (chunks, G, D).  counts[((tau * G + g) * nb + j) * D + c] += |b|;  part[chunk][same index][2] = (sum stat, sum x)
__global__ void __launch_bounds__(MV_THREADS) mv_stats_kernel(const MvArgs a, unsigned long long* __restrict__ counts,
                                                              double* __restrict__ part) {
  __shared__ float sx[MV_CHUNK], ss[MV_CHUNK];
  const int g = blockIdx.y, c = blockIdx.z;
  const int i0 = blockIdx.x * MV_CHUNK;
  const int m = min(MV_CHUNK, a.n - i0);
  for (int t = threadIdx.x; t < m; t += MV_THREADS) {
    const int i = i0 + t;
    const float* vrow = a.values + (size_t)i * a.Dv;
    const float x = a.Dv == 1 ? vrow[0] : vrow[c];
    bool in = a.groups ? a.groups[(size_t)i * a.G + g] != 0 : true;
    if (a.top_label) in = in && mv_argmax(vrow, a.Dv) == c;
    const float s = a.scores[(size_t)i * a.D + c];
    sx[t] = in ? x : CUDART_NAN_F;  // NaN fails every predicate: a point outside the group is in no bucket
    ss[t] = a.mode == 1 ? (s <= x ? 1.0f : 0.0f) : s;
  }
  __syncthreads();
  const int nb = a.n_buckets;
  const size_t cells = (size_t)3 * a.G * nb * a.D;
  for (int bin = threadIdx.x; bin < 3 * nb; bin += MV_THREADS) {
    const int tau = bin / nb, j = bin - tau * nb;
    if (!((a.tau_mask >> tau) & 1)) continue;
    const float v = a.buckets[j];
    unsigned cnt = 0;
    double sum_s = 0.0, sum_x = 0.0;
    for (int t = 0; t < m; ++t) {
      const float x = sx[t];
      if (mv_cond(tau, x, v, a.h)) {
        ++cnt;
        sum_s += (double)ss[t];
        sum_x += (double)x;
      }
    }
    const size_t cell = (((size_t)tau * a.G + g) * nb + j) * a.D + c;
    if (cnt) atomicAdd(counts + cell, (unsigned long long)cnt);
    part[((size_t)blockIdx.x * cells + cell) * 2] = sum_s;
    part[((size_t)blockIdx.x * cells + cell) * 2 + 1] = sum_x;
  }
}

// out[cell] = (count, sum stat, sum x): the partials of the chunks added in chunk order
__global__ void __launch_bounds__(256) mv_reduce_kernel(const unsigned long long* __restrict__ counts,
                                                        const double* __restrict__ part, int n_chunks, size_t cells,
                                                        double* __restrict__ out) {
  for (size_t cell = blockIdx.x * (size_t)blockDim.x + threadIdx.x; cell < cells; cell += (size_t)gridDim.x * blockDim.x) {
    double s = 0.0, x = 0.0;
    for (int k = 0; k < n_chunks; ++k) {
      s += part[((size_t)k * cells + cell) * 2];
      x += part[((size_t)k * cells + cell) * 2 + 1];
    }
    out[cell * 3] = (double)counts[cell];
    out[cell * 3 + 1] = s;
    out[cell * 3 + 2] = x;
  }
}

// values[b(, c)] = clip(values + eta patch, 0, 1)  or  clip(values (1 - eta (1 - patch)), 0, 1)   (iterative/base.py:
// 381-397), b of (tau, g, v, c) taken on the values BEFORE the update; n_members += |b|
__global__ void __launch_bounds__(256) mv_patch_kernel(float* __restrict__ values, const uint8_t* __restrict__ groups, int n,
                                                       int Dv, int G, int g, int c, int tau, float v, float h, int top_label,
                                                       float eta, float patch, int multiplicative,
                                                       int* __restrict__ n_members) {
  int mine = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float* vrow = values + (size_t)i * Dv;
    const int col = Dv == 1 ? 0 : c;
    const float x = vrow[col];
    bool in = groups ? groups[(size_t)i * G + g] != 0 : true;
    if (top_label) in = in && mv_argmax(vrow, Dv) == c;
    in = in && mv_cond(tau, x, v, h);
    if (in) {
      const float y = multiplicative ? __fmul_rn(x, __fsub_rn(1.0f, __fmul_rn(eta, __fsub_rn(1.0f, patch))))
                                     : __fadd_rn(x, __fmul_rn(eta, patch));
      vrow[col] = (y != y) ? y : fminf(fmaxf(y, 0.0f), 1.0f);  // jnp.clip keeps a NaN
      ++mine;
    }
  }
  mine = __reduce_add_sync(0xffffffffu, mine);
  if ((threadIdx.x & 31) == 0 && mine && n_members) atomicAdd(n_members, mine);
}

// per-point loss terms: mode 0 (values - scores)^2 (mixins/multicalibrator.py:10-15), mode 1 the pinball loss at
// `coverage` (mixins/batchmvp.py:12-27)
__global__ void __launch_bounds__(256) mv_loss_terms_kernel(const float* __restrict__ values, const float* __restrict__ scores,
                                                            int64_t total, int D, int Dv, int mode, float coverage,
                                                            float* __restrict__ terms) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    // Dv == D: element k of both; Dv == 1 < D never happens for a loss (scores[:, 0] is used): k indexes points
    const float x = values[k];
    const float s = Dv == 1 ? scores[k * D] : scores[k];
    const float diff = __fsub_rn(s, x);
    float t;
    if (mode == 0) {
      t = __fmul_rn(diff, diff);
    } else {
      t = s > x ? __fmul_rn(diff, coverage) : -__fmul_rn(diff, __fsub_rn(1.0f, coverage));
    }
    terms[k] = t;
  }
}

__global__ void __launch_bounds__(1024) mv_ordered_sum_kernel(const float* __restrict__ x, int64_t n, double* __restrict__ out) {
  __shared__ double sh[1024];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) acc += (double)x[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0];
}

// BatchMVP patch search (iterative/batch_mvp.py:194-230): counts[k] = #{i in b : scores_i <= values_i + delta_k},
// counts[m] = |b|, b of (tau, g, v) on the current values.  One thread per delta over chunks staged in shared memory.
__global__ void __launch_bounds__(256) mv_delta_counts_kernel(const float* __restrict__ values,
                                                              const float* __restrict__ scores,
                                                              const uint8_t* __restrict__ groups, int n, int G, int g, int tau,
                                                              float v, float h, const float* __restrict__ deltas, int m,
                                                              unsigned long long* __restrict__ counts) {
  __shared__ float sx[MV_CHUNK], ss[MV_CHUNK];
  const int i0 = blockIdx.x * MV_CHUNK;
  const int len = min(MV_CHUNK, n - i0);
  int members = 0;
  for (int t = threadIdx.x; t < len; t += blockDim.x) {
    const int i = i0 + t;
    const float x = values[i];
    bool in = groups ? groups[(size_t)i * G + g] != 0 : true;
    in = in && mv_cond(tau, x, v, h);
    sx[t] = in ? x : CUDART_NAN_F;  // scores <= NaN is false
    ss[t] = scores[i];
    members += in ? 1 : 0;
  }
  members = __reduce_add_sync(0xffffffffu, members);
  if ((threadIdx.x & 31) == 0 && members) atomicAdd(counts + m, (unsigned long long)members);
  __syncthreads();
  for (int k = threadIdx.x; k < m; k += blockDim.x) {
    const float d = deltas[k];
    unsigned cnt = 0;
    for (int t = 0; t < len; ++t) cnt += ss[t] <= __fadd_rn(sx[t], d) ? 1u : 0u;
    if (cnt) atomicAdd(counts + k, (unsigned long long)cnt);
  }
}

}  // namespace fb200

using namespace fb200;

static size_t mv_cells(int G, int n_buckets, int D) { return (size_t)3 * G * n_buckets * D; }

extern "C" size_t fb200_multivalid_workspace_bytes(int n, int G, int n_buckets, int D) {
  if (n <= 0 || G <= 0 || n_buckets <= 0 || D <= 0) return 0;
  const size_t chunks = ((size_t)n + MV_CHUNK - 1) / MV_CHUNK;
  const size_t cells = mv_cells(G, n_buckets, D);
  return cells * sizeof(unsigned long long) + chunks * cells * 2 * sizeof(double);
}

extern "C" int fb200_multivalid_stats(const float* values, const float* scores, const uint8_t* groups, const float* buckets,
                                      int n, int D, int Dv, int G, int n_buckets, int top_label, int mode, int tau_mask,
                                      void* workspace, size_t workspace_bytes, double* out, void* stream) {
  if (!values || !scores || !buckets || !workspace || !out) return FB200_E_NULL;
  if (n <= 0 || D <= 0 || G <= 0 || n_buckets <= 0 || (Dv != 1 && Dv != D) || (top_label && Dv != D)) return FB200_E_SHAPE;
  if (G > 65535 || D > 65535) return FB200_E_UNSUPPORTED;
  if (workspace_bytes < fb200_multivalid_workspace_bytes(n, G, n_buckets, D)) return FB200_E_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(workspace) & 7u) return FB200_E_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t cells = mv_cells(G, n_buckets, D);
  const int chunks = (n + MV_CHUNK - 1) / MV_CHUNK;
  unsigned long long* counts = static_cast<unsigned long long*>(workspace);
  double* part = reinterpret_cast<double*>(counts + cells);
  cudaError_t e = cudaMemsetAsync(workspace, 0, fb200_multivalid_workspace_bytes(n, G, n_buckets, D), st);
  if (e != cudaSuccess) return (int)e;
  MvArgs a{values, scores, groups, buckets, n, D, Dv, G, n_buckets, top_label, mode, (float)(0.5 / (double)n_buckets), tau_mask};
  a.h = (float)(0.5 / (double)n_buckets);
  mv_stats_kernel<<<dim3(chunks, G, D), MV_THREADS, 0, st>>>(a, counts, part);
  FB200_LAUNCH_CHECK();
  int blocks = (int)((cells + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  mv_reduce_kernel<<<blocks, 256, 0, st>>>(counts, part, chunks, cells, out);
  FB200_LAUNCH_CHECK();
  return 0;
}

extern "C" int fb200_multivalid_patch(float* values, const uint8_t* groups, int n, int Dv, int G, int g, int c, int tau,
                                      float v, int n_buckets, int top_label, float eta, float patch, int multiplicative,
                                      int32_t* n_members, void* stream) {
  if (!values) return FB200_E_NULL;
  if (n <= 0 || Dv <= 0 || n_buckets <= 0 || tau < 0 || tau > 2 || c < 0 || (Dv > 1 && c >= Dv) || (groups && (g < 0 || g >= G)))
    return FB200_E_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  if (n_members) {
    cudaError_t e = cudaMemsetAsync(n_members, 0, sizeof(int32_t), st);
    if (e != cudaSuccess) return (int)e;
  }
  int blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  mv_patch_kernel<<<blocks, 256, 0, st>>>(values, groups, n, Dv, G, g, c, tau, v, (float)(0.5 / (double)n_buckets), top_label,
                                          eta, patch, multiplicative, n_members);
  FB200_LAUNCH_CHECK();
  return 0;
}

extern "C" int fb200_multivalid_loss(const float* values, const float* scores, int n, int D, int Dv, int mode, float coverage,
                                     float* terms, double* out, void* stream) {
  if (!values || !scores || !terms || !out) return FB200_E_NULL;
  if (n <= 0 || D <= 0 || (Dv != 1 && Dv != D)) return FB200_E_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = (int64_t)n * Dv;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  mv_loss_terms_kernel<<<(unsigned)blocks, 256, 0, st>>>(values, scores, total, D, Dv, mode, coverage, terms);
  FB200_LAUNCH_CHECK();
  mv_ordered_sum_kernel<<<1, 1024, 0, st>>>(terms, total, out);
  FB200_LAUNCH_CHECK();
  return 0;
}

extern "C" int fb200_batchmvp_delta_counts(const float* values, const float* scores, const uint8_t* groups, int n, int G, int g,
                                           int tau, float v, int n_buckets, const float* deltas, int m, uint64_t* counts,
                                           void* stream) {
  if (!values || !scores || !deltas || !counts) return FB200_E_NULL;
  if (n <= 0 || m <= 0 || n_buckets <= 0 || tau < 0 || tau > 2 || (groups && (g < 0 || g >= G))) return FB200_E_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(uint64_t) * (m + 1), st);
  if (e != cudaSuccess) return (int)e;
  const int chunks = (n + MV_CHUNK - 1) / MV_CHUNK;
  mv_delta_counts_kernel<<<chunks, 256, 0, st>>>(values, scores, groups, n, G, g, tau, v, (float)(0.5 / (double)n_buckets), deltas,
                                                 m, reinterpret_cast<unsigned long long*>(counts));
  FB200_LAUNCH_CHECK();
  return 0;
}
