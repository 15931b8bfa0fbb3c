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
#include "row_sort.cuh"

namespace fb200 {

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
  // fused rank-count conformal sets (key-only kernel, all-class mode): mask[b,c] = !(score < val_sorted[n_val-1-K])
  uint8_t* cs_mask;        // optional [B,C]
  int32_t* cs_sizes;       // optional [B]
  const float* cs_val_sorted;
  int cs_q_index;          // n_val - 1 - K
  int cs_n_val;
  int cs_all_true, cs_all_false;
  int only_flagged;        // all-class key-only kernel: visit only the rows whose cs_sizes entry is -1 (left by the
                           // select kernel below for the rows it could not decide)
};

// The validation score the rank count compares against: count(val > t) <= K  <=>  !(t < val_sorted[n' - 1 - K]), n' the
// number of validation scores that are numbers - a NaN is above nothing (base.py:51-53) and fb200_sort_f32 leaves the
// NaNs at the end like lax.sort.  Without NaNs (the last element says so; its load is independent of the boundary's)
// this is val_sorted[q_index].  K >= n' puts every class in: -inf does that through the same comparison.
__device__ __forceinline__ float conformal_boundary_score(const float* __restrict__ v, int n, int q_index) {
  const float last = __ldg(v + n - 1);
  const float q = __ldg(v + q_index);
  if (last == last) return q;
  int lo = 0, hi = n - 1;  // first NaN
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const float x = __ldg(v + mid);
    if (x != x) hi = mid;
    else lo = mid + 1;
  }
  const int idx = q_index - (n - lo);
  return idx < 0 ? -CUDART_INF_F : __ldg(v + idx);
}

__global__ void __launch_bounds__(256) row_sort_cumsum_kernel(const RowSortArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);
  float* cum = reinterpret_cast<float*>(keys + a.N2);
  float* tree = cum + a.N2;  // 2 * N2 floats
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
    // inclusive prefix sum in rank order, association order of jax.lax.associative_scan (see the warp kernel):
    // up-sweep levels live in tree[] (level l at offset 2*N2 - (2*N2 >> l)), then the down-sweep overwrites them
    for (int i = threadIdx.x; i < N2; i += blockDim.x)
      tree[i] = i < C ? f32_unordered((uint32_t)(keys[i] >> 32)) : 0.0f;
    __syncthreads();
    {
      int off = 0;
      for (int m = N2; m > 1; m >>= 1) {
        for (int i = threadIdx.x; i < m / 2; i += blockDim.x) tree[off + m + i] = tree[off + 2 * i] + tree[off + 2 * i + 1];
        off += m;
        __syncthreads();
      }
      // tree[off] is the total == out of the top level; walk down: out_l in place of u_l
      for (int m = 2; m <= N2; m <<= 1) {
        const int hi_off = off;
        off -= m;
        for (int i = threadIdx.x; i < m / 2; i += blockDim.x) {
          const float odd = tree[hi_off + i];
          if (i > 0) tree[off + 2 * i] = tree[hi_off + i - 1] + tree[off + 2 * i];
          tree[off + 2 * i + 1] = odd;
        }
        __syncthreads();
      }
    }
    for (int i = threadIdx.x; i < C; i += blockDim.x) cum[i] = tree[i];
    __syncthreads();
    if (threadIdx.x == 0) {
      const int y = a.targets ? jax_gather_index(a.targets[b], C) : -1;
      float ys = CUDART_NAN_F;
      int first = -1;
      for (int r = 0; r < C; ++r) {
        const int idx = (int)(keys[r] & 0xFFFFFFFFull);
        if (idx == y) ys = cum[r];
        if (first < 0 && cum[r] > a.set_threshold) first = r;
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

// ------------------------------------------------------------------------------------------------
// Warp-per-row variant for C <= 1024: the 64-bit keys live in registers (NR per lane, logical position
// e = lane * NR + r), so 2/3 of the bitonic network is register-to-register compare-exchange and only the
// strides >= NR cross lanes (shfl.xor).  Keys are stored complemented so an ASCENDING network yields the
// descending (prob, class) order; padding keys are all-ones and sort to the end.  Blocks that the network
// wants descending are handled by complementing the lane's keys for that merge level, so every
// compare-exchange is a plain "min to the low slot" with no run-time direction logic.
//
// Prefix sum: jnp.cumsum on CPU lowers to jax.lax.associative_scan (jax/_src/lax/control_flow/loops.py,
// _cumulative_reduction_primitive: reduce_window only on TPU), i.e. the recursive pairwise tree
//   r[i] = x[2i] + x[2i+1];  odd = scan(r);  out[2i+1] = odd[i];  out[0] = x[0];  out[2i] = odd[i-1] + x[2i].
// Padding the row with zeros to a power of two does not change any prefix (each out[k] only sees x[0..k]), so
// the tree runs on 32*NR slots: levels inside a lane's block in registers, the 32 lane totals by shuffles.
// ------------------------------------------------------------------------------------------------
typedef unsigned long long u64;

__device__ __forceinline__ void ce_min_first(u64& a, u64& b) {
  const bool sw = a > b;
  const u64 lo = sw ? b : a;
  const u64 hi = sw ? a : b;
  a = lo;
  b = hi;
}

template <int NR>
__device__ __forceinline__ void warp_bitonic_sort(u64 (&key)[NR], int lane) {
  constexpr int N = 32 * NR;
  // levels whose blocks fit inside a lane: direction is a compile-time function of the register index
#pragma unroll
  for (int k = 2; k < NR; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        if ((r & j) == 0) {
          if ((r & k) == 0) ce_min_first(key[r], key[r | j]);
          else ce_min_first(key[r | j], key[r]);
        }
      }
    }
  }
  // levels k >= NR: the direction of the lane's block is a lane bit; complement instead of reversing.  Every such level
  // is [complement if the direction changes] + [cross-lane stages, the lane mask a loop variable] + [the same in-lane
  // merge]: ONE rolled loop body (fully unrolled the kernel was ~150 KB of straight-line code and stalled on
  // instruction fetch, see the key-only kernel below).
  u64 flipped = 0ull;
#pragma unroll 1
  for (int kl = 1; kl <= 32; kl <<= 1) {  // k = kl * NR
    const u64 want = (kl < 32 && (lane & kl) != 0) ? ~0ull : 0ull;
    const u64 x = want ^ flipped;
    flipped = want;
#pragma unroll
    for (int r = 0; r < NR; ++r) key[r] ^= x;
#pragma unroll 1
    for (int lmask = kl >> 1; lmask > 0; lmask >>= 1) {
      const bool upper = (lane & lmask) != 0;  // the upper lane of a pair keeps the larger key
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const u64 other = __shfl_xor_sync(0xffffffffu, key[r], lmask);
        const bool lt = key[r] < other;
        key[r] = (lt == upper) ? other : key[r];
      }
    }
#pragma unroll
    for (int j = NR >> 1; j > 0; j >>= 1) {
#pragma unroll
      for (int r = 0; r < NR; ++r)
        if ((r & j) == 0) ce_min_first(key[r], key[r | j]);
    }
  }
}

constexpr int kSortWarps = 8;

template <int NR, int MODE>  // MODE 0: all-class scores (+perm); 1: target score; 2: credible sets (+perm)
__global__ void __launch_bounds__(kSortWarps * 32) row_sort_warp_kernel(const RowSortArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // warp index through a shuffle: provably warp-uniform, so the row loop is convergent for the compiler and the
  // 64-bit shuffles of the network need no per-call WARPSYNC
  const int warp_in_cta = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int C = a.C;
  float* unperm = reinterpret_cast<float*>(smem_raw) + (size_t)warp_in_cta * (32 * NR);
  const int nwarps = gridDim.x * kSortWarps;
  for (int b = blockIdx.x * kSortWarps + warp_in_cta; b < a.B; b += nwarps) {
    const float* row = a.probs + (size_t)b * C;
    u64 key[NR];
    // coalesced load; the initial arrangement is irrelevant to the sort (the class index rides in the key)
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const int c = r * 32 + lane;
      u64 k = ~0ull;
      if (c < C) {
        const float p = row[c] + 0.0f;  // -0 -> +0
        k = ~(((u64)f32_ordered(p) << 32) | (unsigned)c);
      }
      key[r] = k;
    }
    warp_bitonic_sort<NR>(key, lane);
    // rank of key[r] is lane * NR + r; split into (prob, class)
    float x[NR];
    int cls[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const u64 kk = ~key[r];
      cls[r] = (int)(kk & 0xFFFFFFFFull);
      x[r] = (lane * NR + r < C) ? f32_unordered((uint32_t)(kk >> 32)) : 0.0f;
    }
    warp_row_tree_prefix<NR>(x, lane);
    if (MODE == 1) {
      const int y = a.targets ? jax_gather_index(a.targets[b], C) : -1;
      float ys = CUDART_NAN_F;
      bool mine = false;
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        if (lane * NR + r < C && cls[r] == y) {
          ys = x[r];
          mine = true;
        }
      }
      const unsigned m = __ballot_sync(0xffffffffu, mine);  // exactly one lane saw the target (none if out of range)
      if (m) ys = __shfl_sync(0xffffffffu, ys, __ffs(m) - 1);
      if (lane == 0 && a.scores) a.scores[b] = ys;
    }
    if (MODE == 2) {
      int first = 0x7FFFFFFF;
#pragma unroll
      for (int r = NR - 1; r >= 0; --r)
        if (lane * NR + r < C && x[r] > a.set_threshold) first = lane * NR + r;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
      if (lane == 0 && a.set_size) a.set_size[b] = (first == 0x7FFFFFFF ? 0 : first) + 1;
    }
    if (MODE == 0 && a.scores) {
#pragma unroll
      for (int r = 0; r < NR; ++r)
        if (lane * NR + r < C) unperm[cls[r]] = x[r];
      __syncwarp();
      float* orow = a.scores + (size_t)b * C;
      for (int c = lane; c < C; c += 32) orow[c] = unperm[c];
      __syncwarp();
    }
    if (a.perm) {
      int32_t* prow = a.perm + (size_t)b * C;
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const int rank = lane * NR + r;
        if (rank < C) prow[rank] = cls[r];
      }
    }
  }
}

template <int NR>
static int launch_row_sort_warp(const RowSortArgs& a, cudaStream_t st) {
  const int mode = a.set_size ? 2 : (a.targets ? 1 : 0);
  const size_t smem = mode == 0 ? (size_t)kSortWarps * 32 * NR * sizeof(float) : 0;
  int grid = (a.B + kSortWarps - 1) / kSortWarps;
  if (grid > 148 * 16) grid = 148 * 16;
  if (mode == 0) row_sort_warp_kernel<NR, 0><<<grid, kSortWarps * 32, smem, st>>>(a);
  else if (mode == 1) row_sort_warp_kernel<NR, 1><<<grid, kSortWarps * 32, smem, st>>>(a);
  else row_sort_warp_kernel<NR, 2><<<grid, kSortWarps * 32, smem, st>>>(a);
  FB200_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Key-only variant (no permutation requested): the network sorts bare 32-bit keys ~ordered(prob), so a
// compare-exchange is one min and one max (the 64-bit (prob, class) keys cost a 2-instruction compare and four
// selects), and the class a rank belongs to is never carried around.  Instead every class finds its own rank
// afterwards: rank(c) = #{keys strictly before key(c)} + #{c' > c with key(c') == key(c)} - the stable
// descending order of argsort(...)[::-1] (adaptive_prediction.py:13,20).  The first term is a branch-free lower
// bound over the sorted keys in shared memory (log2(32*NR) probes; slot i lives at i + i/32 so that probes of
// different 32-slot blocks fall into different banks); the second term is zero unless the row holds equal
// probabilities, which the sorted registers reveal with one compare per slot - only such rows run the tie pass
// (match_any per register slot, classes visited from high to low, one running counter per group of equal keys).
// Prefix sums are taken over the sorted values exactly as in the 64-bit kernel, so the outputs are the same bits.
// ------------------------------------------------------------------------------------------------
template <int NR>
__host__ __device__ constexpr int sort32_pad() { return 32 * NR + NR; }  // slot i at i + (i >> 5)
// shared-memory words per warp: raw row [32 NR] (+ sorted keys and prefix sums, padded), 16-byte granular so that
// every warp's staging buffer is a legal cp.async 16 destination
template <int NR, int MODE>
__host__ __device__ constexpr int sort32_warp_words() {
  return ((MODE == 0 ? 32 * NR + 2 * sort32_pad<NR>() : 32 * NR) + 3) & ~3;
}

// One level of the branch-free lower bound for all NR searches of the lane, then the next level (the step is a
// template parameter because the LDS offset has to be an immediate).  Slot pos + STEP - 1 is probed: for STEP >= 32
// both pos and STEP are multiples of 32, so its padded address is addr(pos) + STEP + STEP/32 - 2.
template <int NR, int STEP>
__device__ __forceinline__ void lower_bound_levels(uint32_t (&ap)[NR], const uint32_t (&own)[NR]) {
  if constexpr (STEP >= 1) {
    constexpr int inc = STEP + (STEP >> 5);
    constexpr int probe = STEP >= 32 ? inc - 2 : STEP - 1;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      uint32_t v;
      asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(ap[r]), "n"(probe * 4));
      if (v < own[r]) ap[r] += inc * 4;
    }
    lower_bound_levels<NR, STEP / 2>(ap, own);
  }
}

// Asynchronous copy of one row (class order) into the warp's staging buffer: LDGSTS, 16 bytes per lane where the
// row is 16-byte aligned, else 4.  The next row is requested as soon as the current one has been consumed, so its
// HBM latency hides behind the rank lookup instead of stalling the 4 warps a scheduler has.
template <int NR>
__device__ __forceinline__ void row_prefetch(float* srow, const float* row, int C, int lane, bool vec16) {
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(srow);
  if (vec16) {
    for (int q = lane; q * 4 < C; q += 32)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + q * 16), "l"(row + q * 4) : "memory");
  } else {
    for (int c = lane; c < C; c += 32)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + c * 4), "l"(row + c) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

template <int NR, int MODE>  // MODE 0: all-class scores; 1: target score
__global__ void __launch_bounds__(kSortWarps * 32, 2) row_sort32_warp_kernel(const RowSortArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int N = 32 * NR;
  constexpr int PAD = sort32_pad<NR>();
  constexpr int kWarpWords = sort32_warp_words<NR, MODE>();
  const int warp_in_cta = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int C = a.C;
  float* srow = reinterpret_cast<float*>(smem_raw) + (size_t)warp_in_cta * kWarpWords;  // the raw row, class order
  uint32_t* skey = reinterpret_cast<uint32_t*>(srow + N);                                // sorted keys, padded
  int* scnt = reinterpret_cast<int*>(skey);  // tie pass only; the sorted keys are dead by then
  const bool vec16 = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.probs) & 15) == 0);
  const int nwarps = gridDim.x * kSortWarps;
  // the warp's next row at or after b: every row, or (second pass of the select kernel) only the flagged ones
  auto row_from = [&](int r) {
    if (MODE == 0 && a.only_flagged)
      while (r < a.B && a.cs_sizes[r] != -1) r += nwarps;
    return r;
  };
  int b = row_from(blockIdx.x * kSortWarps + warp_in_cta);
  if (b < a.B) row_prefetch<NR>(srow, a.probs + (size_t)b * C, C, lane, vec16);
  for (int nb; b < a.B; b = nb) {
    nb = row_from(b + nwarps);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    uint32_t own[NR], key[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const int c = r * 32 + lane;
      key[r] = (c < C) ? ~f32_ordered(srow[c] + 0.0f) : 0xFFFFFFFFu;  // -0 -> +0; padding sorts last
    }
    int tpos = -1;
    if (MODE == 1) {
      // rank of the target class, counted on the unsorted row
      const int y = a.targets ? jax_gather_index(a.targets[b], C) : -1;
      if ((unsigned)y < (unsigned)C) {
        const uint32_t ky = ~f32_ordered(srow[y] + 0.0f);
        int cnt = 0;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          const int c = r * 32 + lane;
          cnt += (key[r] < ky || (key[r] == ky && c > y && c < C)) ? 1 : 0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        tpos = cnt;
      }
      __syncwarp();
      if (nb < a.B) row_prefetch<NR>(srow, a.probs + (size_t)nb * C, C, lane, vec16);
    }
    warp_bitonic_sort32<NR>(key, lane);
    float x[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) x[r] = (lane * NR + r < C) ? f32_unordered(~key[r]) : 0.0f;
    warp_row_tree_prefix<NR>(x, lane);
    if (MODE == 1) {
      float ys = CUDART_NAN_F;
      if (tpos >= 0) {
        float v = 0.0f;
#pragma unroll
        for (int r = 0; r < NR; ++r) v = (lane * NR + r == tpos) ? x[r] : v;
        ys = __shfl_sync(0xffffffffu, v, tpos / NR);
      }
      if (lane == 0 && a.scores) a.scores[b] = ys;
    }
    if (MODE == 0) {
      // Sets without scores: no class needs its exact rank.  score(rank) >= q is a threshold function of the rank
      // whenever the prefix sums are monotone around q (checked on the actual values: the tree order can break
      // monotonicity by an ulp); the set is then the ranks >= first, and  rank(c) >= first  <=>  prob(c) <= prob at rank
      // `first`  as long as the key just BEFORE the boundary differs from the boundary key (equal probabilities anywhere
      // else leave the sorted values, their prefix sums and this test untouched) - one compare per class instead of the
      // ten-probe rank lookup.  Rows with a tie across the boundary or a non-monotone boundary take the general path.
      bool done = false;
      if (a.cs_mask && !a.scores && !a.cs_all_true && !a.cs_all_false) {
        const float q = conformal_boundary_score(a.cs_val_sorted, a.cs_n_val, a.cs_q_index);
        int n_true = 0, first = 0x7FFFFFFF;
#pragma unroll
        for (int r = NR - 1; r >= 0; --r) {
          const int rank = lane * NR + r;
          const bool t = (rank < C) && !(x[r] < q);
          n_true += t ? 1 : 0;
          first = t ? rank : first;  // descending r: ends at the lane's smallest rank that is in
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          n_true += __shfl_xor_sync(0xffffffffu, n_true, o);
          first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
        }
        const bool none = first == 0x7FFFFFFF;
        // boundary key and the key before it (previous register slot, or the previous lane's last slot)
        const uint32_t prev_last = __shfl_up_sync(0xffffffffu, key[NR - 1], 1);
        uint32_t kb = 0u, kp = 0u;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          const bool hit = lane * NR + r == first;
          kb = hit ? key[r] : kb;
          kp = hit ? (r > 0 ? key[r > 0 ? r - 1 : 0] : prev_last) : kp;
        }
        const int src = none ? 0 : first / NR;
        kb = __shfl_sync(0xffffffffu, kb, src);
        kp = __shfl_sync(0xffffffffu, kp, src);
        const bool boundary_tie = !none && first > 0 && kp == kb;
        // key 0 is a NaN probability: it ranks first, every score of the row is NaN and "<= pb" decides nothing -
        // the general path below puts every class in (no validation score is above a NaN)
        const bool nan_row = !none && kb == 0u;
        if ((none ? (n_true == 0) : (n_true == C - first)) && !boundary_tie && !nan_row) {  // warp-uniform
          const float pb = f32_unordered(~kb);  // probability at the boundary rank
          uint8_t* mrow = a.cs_mask + (size_t)b * C;
          if (vec16 && (reinterpret_cast<uintptr_t>(mrow) & 3u) == 0) {
            // four classes per lane and store: class c = (lane + 32 j) * 4 + k
#pragma unroll
            for (int j = 0; j < (NR + 3) / 4; ++j) {
              const int c0 = (lane + 32 * j) * 4;
              if (c0 < C) {
                const float4 v = *reinterpret_cast<const float4*>(srow + c0);
                const uint32_t w = ((!none && v.x <= pb) ? 1u : 0u) | ((!none && v.y <= pb) ? 0x100u : 0u) |
                                   ((!none && v.z <= pb) ? 0x10000u : 0u) | ((!none && v.w <= pb) ? 0x1000000u : 0u);
                *reinterpret_cast<uint32_t*>(mrow + c0) = w;
              }
            }
          } else {
#pragma unroll
            for (int r = 0; r < NR; ++r) {
              const int c = r * 32 + lane;
              if (c < C) mrow[c] = (!none && srow[c] <= pb) ? 1 : 0;  // -0 == +0: no canonicalisation needed to compare
            }
          }
          if (lane == 0 && a.cs_sizes) a.cs_sizes[b] = n_true;
          __syncwarp();
          if (nb < a.B) row_prefetch<NR>(srow, a.probs + (size_t)nb * C, C, lane, vec16);
          done = true;
        }
      }
      if (!done) {
      // equal neighbours among the real (rank < C) keys?  (only such rows run the exact tie pass of the rank lookup)
      bool tie = false;
#pragma unroll
      for (int r = 0; r + 1 < NR; ++r) tie |= (key[r] == key[r + 1]) && (lane * NR + r + 1 < C);
      {
        const uint32_t nxt = __shfl_down_sync(0xffffffffu, key[0], 1);
        tie |= (lane < 31) && (key[NR - 1] == nxt) && ((lane + 1) * NR < C);
      }
      const bool any_tie = __any_sync(0xffffffffu, tie);
      // the row's own keys again, class order, then the staging buffer is free for the next row
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const int c = r * 32 + lane;
        own[r] = (c < C) ? ~f32_ordered(srow[c] + 0.0f) : 0xFFFFFFFFu;
      }
      const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(skey);
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const int i = lane * NR + r;
        const uint32_t ad = sbase + (uint32_t)(i + (i >> 5)) * 4u;
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(ad), "r"(key[r]) : "memory");
        asm volatile("st.shared.f32 [%0+%2], %1;" ::"r"(ad), "f"(x[r]), "n"(PAD * 4) : "memory");
      }
      __syncwarp();
      if (nb < a.B) row_prefetch<NR>(srow, a.probs + (size_t)nb * C, C, lane, vec16);
      // lower bound, kept as the shared-space BYTE address of the slot so that a probe is one LDS with an
      // immediate offset, one compare and one predicated add
      uint32_t ap[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r) ap[r] = sbase;
      lower_bound_levels<NR, N / 2>(ap, own);
      if (any_tie) {
        __syncwarp();
        for (int i = lane; i < PAD; i += 32) scnt[i] = 0;
        __syncwarp();
#pragma unroll
        for (int r = NR - 1; r >= 0; --r) {
          const unsigned m = __match_any_sync(0xffffffffu, own[r]);
          const int above = __popc(m & ~((2u << lane) - 1u));  // equal keys in higher lanes == larger classes
          const int slot = (int)((ap[r] - sbase) >> 2);
          const int base = scnt[slot];
          __syncwarp();
          if (lane == __ffs(m) - 1) scnt[slot] = base + __popc(m);
          __syncwarp();
          const int pos = slot - slot / 33 + base + above;
          ap[r] = sbase + (uint32_t)(pos + (pos >> 5)) * 4u;
        }
      }
      if (a.cs_mask) {
        // conformal sets straight from the scores (conformal/classification/base.py:45-64, see conformal_sets_kernel):
        // the [B,C] score matrix need not be written and read back
        const float q = (a.cs_all_true || a.cs_all_false)
                            ? 0.0f
                            : conformal_boundary_score(a.cs_val_sorted, a.cs_n_val, a.cs_q_index);
        float* orow = a.scores ? a.scores + (size_t)b * C : nullptr;
        uint8_t* mrow = a.cs_mask + (size_t)b * C;
        int cnt = 0;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          const int c = r * 32 + lane;
          float v;
          asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(ap[r]), "n"(PAD * 4));
          const bool in = a.cs_all_true || (!a.cs_all_false && !(v < q));
          if (c < C) {
            if (orow) orow[c] = v;
            mrow[c] = in ? 1 : 0;
            cnt += in ? 1 : 0;
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0 && a.cs_sizes) a.cs_sizes[b] = cnt;
      } else if (a.scores) {
        float* orow = a.scores + (size_t)b * C;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          const int c = r * 32 + lane;
          float v;
          asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(ap[r]), "n"(PAD * 4));
          if (c < C) orow[c] = v;
        }
      }
      }  // !done
      __syncwarp();
    }
  }
}

template <int NR>
static int launch_row_sort32_warp(const RowSortArgs& a, cudaStream_t st) {
  const int mode = a.targets ? 1 : 0;
  const size_t words = mode == 0 ? sort32_warp_words<NR, 0>() : sort32_warp_words<NR, 1>();
  const size_t smem = (size_t)kSortWarps * words * sizeof(uint32_t);
  int grid = (a.B + kSortWarps - 1) / kSortWarps;
  if (grid > 148 * 16) grid = 148 * 16;
  if (mode == 0) {
    auto kern = row_sort32_warp_kernel<NR, 0>;
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
    }
    kern<<<grid, kSortWarps * 32, smem, st>>>(a);
  } else {
    row_sort32_warp_kernel<NR, 1><<<grid, kSortWarps * 32, smem, st>>>(a);
  }
  FB200_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// K5+K6 "select": APS conformal sets without the scores and without sorting the row.
// The set of a row is {classes whose score >= q} = the ranks >= first, first = the smallest rank whose prefix sum
// reaches q (base.py:51-53 counts the validation scores ABOVE a test score, so the set is the low-probability tail).
// Only the exact prefix sums of the ranks up to `first` decide it, and in jax.lax.associative_scan's order the prefix
// at rank i depends on the values at ranks <= i alone.  So: pick a probability threshold t, take A = {p >= t} - the
// top-|A| of the row as a set, whatever the ties - and if A holds at most kSelMax values whose mass reaches q, sort
// just A (1..8 keys per lane instead of 32), take its prefix sums in the same tree order and find the boundary there.
// What A cannot show is checked instead of assumed:
//   * ranks inside A: every prefix is compared with q (the tree order may break monotonicity by an ulp);
//   * ranks beyond A: prefix[i] >= exact[i] (1 - g) >= exact[|A|-1] (1 - g) >= prefix[|A|-1] (1 - g) / (1 + g) for
//     non-negative values, g = 20 * 2^-24 (no value of a 1024-long scan goes through more than 20 additions), so
//     prefix[|A|-1] * (1 - 4e-6) >= q proves them all in;  rows with a negative value or -0 never get here;
//   * a tie across the boundary rank needs the stable order of the classes: not decided here.
// Rows this kernel cannot decide get size -1 and are done by row_sort32_warp_kernel<NR, 0> (only_flagged) right after.
// The threshold is t = max * 2^x; x is the warp's running estimate (found by bisection on its first row; afterwards
// loosened by 0.7 when the mass falls short, tightened by 1.2 after every accepted row - a warp only sees a few dozen
// rows, so it has to adapt quickly), so a typical row costs 1.2-1.5 counting passes and sorts 64-128 values.
// Rows are double-buffered in shared memory (cp.async): the next row's HBM latency hides behind the current one.
// A warp that fails kSelGiveUp rows in a row flags its next kSelSkip rows unread, so data whose boundaries lie deep
// (flat rows against a high quantile) costs the full sort plus about a tenth of a counting pass per row.
// ------------------------------------------------------------------------------------------------
constexpr int kSelWarps = 4;
constexpr int kSelMax = 256;  // most values a row may need sorted (8 per lane)
constexpr int kSelGiveUp = 3;  // undecided rows in a row after which a warp stops reading rows ...
constexpr int kSelSkip = 12;   // ... for this many of its rows

template <int NRS>
__device__ __forceinline__ bool aps_select_decide(const uint32_t* scratch, int cnt, int lane, float q, float& pb,
                                                  int& first_out) {
  uint32_t key[NRS];
#pragma unroll
  for (int r = 0; r < NRS; ++r) {
    const int i = r * 32 + lane;
    key[r] = i < cnt ? ~scratch[i] : 0xFFFFFFFFu;  // descending values = ascending complemented bits; padding last
  }
  warp_bitonic_sort32<NRS>(key, lane);
  float x[NRS];
#pragma unroll
  for (int r = 0; r < NRS; ++r) x[r] = (lane * NRS + r < cnt) ? __uint_as_float(~key[r]) : 0.0f;
  warp_row_tree_prefix<NRS>(x, lane);
  int n_true = 0, first = 0x7FFFFFFF;
  float xl = 0.0f;
#pragma unroll
  for (int r = NRS - 1; r >= 0; --r) {
    const int rank = lane * NRS + r;
    const bool t = (rank < cnt) && !(x[r] < q);
    n_true += t ? 1 : 0;
    first = t ? rank : first;
    xl = (rank == cnt - 1) ? x[r] : xl;
  }
  n_true = __reduce_add_sync(0xffffffffu, n_true);
  first = __reduce_min_sync(0xffffffffu, first);
  xl = __shfl_sync(0xffffffffu, xl, (cnt - 1) / NRS);
  if (first == 0x7FFFFFFF) return false;
  const uint32_t prev_last = __shfl_up_sync(0xffffffffu, key[NRS - 1], 1);
  uint32_t kb = 0u, kp = 0u;
#pragma unroll
  for (int r = 0; r < NRS; ++r) {
    const bool hit = lane * NRS + r == first;
    kb = hit ? key[r] : kb;
    kp = hit ? (r > 0 ? key[r > 0 ? r - 1 : 0] : prev_last) : kp;
  }
  kb = __shfl_sync(0xffffffffu, kb, first / NRS);
  kp = __shfl_sync(0xffffffffu, kp, first / NRS);
  pb = __uint_as_float(~kb);
  first_out = first;
  return n_true == cnt - first && !(first > 0 && kp == kb) && (xl * 0.999996f >= q);
}

template <int NR>  // 32 * NR >= C, C % 4 == 0, rows 16-byte aligned, mask rows 4-byte aligned
__global__ void __launch_bounds__(kSelWarps * 32, 5) aps_sets_select_kernel(const RowSortArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int N = 32 * NR;
  constexpr int kWarpWords = 2 * N + kSelMax;
  const int warp_in_cta = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int C = a.C;
  float* rows = reinterpret_cast<float*>(smem_raw) + (size_t)warp_in_cta * kWarpWords;
  uint32_t* scratch = reinterpret_cast<uint32_t*>(rows + 2 * N);
  const int nwarps = gridDim.x * kSelWarps;
  const float q = conformal_boundary_score(a.cs_val_sorted, a.cs_n_val, a.cs_q_index);
  const float qm = q * 1.00002f;  // the counting pass sums in another order than the scan: aim a little above q
  float xs = 0.0f;      // log2 of the threshold / row maximum ratio the warp tries first on its next row
  bool primed = false;  // ... once a row has told it where to look
  int buf = 0, fails = 0;
  // the rows are copied as C floats; the padding classes [C, 32 NR) of both buffers read as zero from here on
  for (int i = C + lane; i < N; i += 32) rows[i] = 0.0f, rows[N + i] = 0.0f;
  __syncwarp();
  int b = blockIdx.x * kSelWarps + warp_in_cta;
  if (b < a.B) row_prefetch<NR>(rows, a.probs + (size_t)b * C, C, lane, true);
  for (int nb; b < a.B; b = nb, buf ^= 1) {
    const float* srow = rows + buf * N;
    nb = b + nwarps;
    if (fails >= kSelGiveUp) {
      // the data is not of the kind this kernel decides (boundary ranks deeper than kSelMax): leave the next
      // kSelSkip rows to the full sort unread, then probe again (one more miss and the next burst is skipped)
      const int r = nb + lane * nwarps;
      if (lane < kSelSkip && r < a.B) a.cs_sizes[r] = -1;
      nb += kSelSkip * nwarps;
      fails = kSelGiveUp - 1;
    }
    if (nb < a.B) {
      row_prefetch<NR>(rows + (buf ^ 1) * N, a.probs + (size_t)nb * C, C, lane, true);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_all;" ::: "memory");
    }
    __syncwarp();
    // the row, four consecutive classes per lane and slot: class (lane + 32 j) * 4 + k in v[4 j + k]
    float v[NR];
    uint32_t orb = 0u, mxb = 0u;
#pragma unroll
    for (int j = 0; j < NR / 4; ++j) {
      const float4 f = *reinterpret_cast<const float4*>(srow + (lane + 32 * j) * 4);
      v[4 * j] = f.x, v[4 * j + 1] = f.y, v[4 * j + 2] = f.z, v[4 * j + 3] = f.w;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t u = __float_as_uint(v[4 * j + k]);
        orb |= u;
        mxb = max(mxb, u);
      }
    }
    __syncwarp();  // all lanes hold their part of the row: this buffer may be refilled by the NEXT iteration's prefetch
    orb = __reduce_or_sync(0xffffffffu, orb);
    mxb = __reduce_max_sync(0xffffffffu, mxb);
    // non-negative finite rows only (bit patterns then order like the values; -0, negatives, inf, NaN: general path)
    bool decided = false;
    float pb = 0.0f;
    int first = 0;
    if (!(orb & 0x80000000u) && mxb < 0x7F800000u && mxb != 0u) {
      const float mx = __uint_as_float(mxb);
      // threshold search in log2(t / mx).  First row of the warp: bisection over [-20, -1] from -6; afterwards one
      // step of -log2(0.7) when the mass falls short, +2 when more than kSelMax values pass, bisection once bracketed.
      float lo = primed ? -CUDART_INF_F : -20.0f, hi = primed ? CUDART_INF_F : -1.0f;
      float x = primed ? xs : -6.0f, t = 0.0f;
      const int max_tries = primed ? 5 : 7;
      int mycnt = 0, cnt = 0;
      bool accept = false, hopeless = false;
#pragma unroll 1
      for (int tr = 0; tr < max_tries && !accept && !hopeless; ++tr) {
        t = mx * exp2f(x);
        float mass = 0.0f, cntf = 0.0f;  // three FP-pipe instructions per value: FSET.BF, FFMA, FADD
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          const float in = v[r] >= t ? 1.0f : 0.0f;
          mass = fmaf(in, v[r], mass);
          cntf += in;
        }
        mycnt = (int)cntf;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mass += __shfl_xor_sync(0xffffffffu, mass, o);
        cnt = __reduce_add_sync(0xffffffffu, mycnt);
        if (!(mass >= qm)) {  // A too small (or q is NaN): lower the threshold ...
          hopeless = cnt >= kSelMax;  // ... unless that could only add values: boundary deeper than kSelMax ranks
          hi = x;
          x = lo > -CUDART_INF_F ? 0.5f * (lo + hi) : x - 0.515f;
        } else if (cnt > kSelMax) {  // enough mass but too many values: raise it
          lo = x;
          x = hi < CUDART_INF_F ? 0.5f * (lo + hi) : x + 2.0f;
        } else {
          accept = true;
        }
      }
      if (!hopeless) {  // (a hopeless row says nothing about where the decidable ones have their boundary)
        xs = fminf(fmaxf(accept ? x + 0.263f : x, -30.0f), -1.0f);  // accepted: try 1.2 x tighter on the next row
        primed = true;
      }
      if (accept) {
        // compact A into the scratch buffer (any order: it is sorted next)
        int pos = mycnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int up = __shfl_up_sync(0xffffffffu, pos, o);
          if (lane >= o) pos += up;
        }
        uint32_t wa = (uint32_t)__cvta_generic_to_shared(scratch) + (uint32_t)(pos - mycnt) * 4u;
#pragma unroll
        for (int r = 0; r < NR; ++r) {  // compare, predicated store, predicated address bump
          asm volatile(
              "{ .reg .pred p; setp.ge.f32 p, %1, %2; @p st.shared.f32 [%0], %1; @p add.u32 %0, %0, 4; }"
              : "+r"(wa)
              : "f"(v[r]), "f"(t)
              : "memory");
        }
        __syncwarp();
        if (cnt <= 32) decided = aps_select_decide<1>(scratch, cnt, lane, q, pb, first);
        else if (cnt <= 64) decided = aps_select_decide<2>(scratch, cnt, lane, q, pb, first);
        else if (cnt <= 128) decided = aps_select_decide<4>(scratch, cnt, lane, q, pb, first);
        else decided = aps_select_decide<8>(scratch, cnt, lane, q, pb, first);
        __syncwarp();  // scratch is free again
      }
    }
    if (decided) {
      uint8_t* mrow = a.cs_mask + (size_t)b * C;
#pragma unroll
      for (int j = 0; j < NR / 4; ++j) {
        const int c0 = (lane + 32 * j) * 4;
        if (c0 < C) {
          // FSET.BF per class (1.0f = 0x3F800000 or 0), the four top bytes gathered by three PRMTs, one AND
          const uint32_t m0 = __float_as_uint(v[4 * j] <= pb ? 1.0f : 0.0f);
          const uint32_t m1 = __float_as_uint(v[4 * j + 1] <= pb ? 1.0f : 0.0f);
          const uint32_t m2 = __float_as_uint(v[4 * j + 2] <= pb ? 1.0f : 0.0f);
          const uint32_t m3 = __float_as_uint(v[4 * j + 3] <= pb ? 1.0f : 0.0f);
          const uint32_t w = __byte_perm(__byte_perm(m0, m1, 0x0073), __byte_perm(m2, m3, 0x0073), 0x5410);
          *reinterpret_cast<uint32_t*>(mrow + c0) = w & 0x01010101u;
        }
      }
      if (lane == 0) a.cs_sizes[b] = C - first;
      fails = 0;
    } else {
      if (lane == 0) a.cs_sizes[b] = -1;
      ++fails;
    }
  }
}

template <int NR>
static int launch_aps_sets_select(RowSortArgs a, cudaStream_t st) {
  const size_t smem = (size_t)kSelWarps * (2 * 32 * NR + kSelMax) * sizeof(uint32_t);
  auto kern = aps_sets_select_kernel<NR>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  int grid = (a.B + kSelWarps - 1) / kSelWarps;
  if (grid > 148 * 5) grid = 148 * 5;
  kern<<<grid, kSelWarps * 32, smem, st>>>(a);
  FB200_LAUNCH_CHECK();
  a.only_flagged = 1;  // whatever it left undecided: full sort + rank lookup
  return launch_row_sort32_warp<NR>(a, st);
}

// ------------------------------------------------------------------------------------------------
// Target samples of the classification predictive (fortuna/prob_model/predictive/base.py:406-440 ->
// fortuna/likelihood/base.py:339-372 -> fortuna/prob_output_layer/classification.py:33-51), the reference's draws:
//   keys = split(rng, T); (key1, key2) = split(keys[t]); i = choice(key1, n_stored); probs = softmax(z_i);
//   keys_b = split(key2, B); y[t, b] = choice(keys_b[b], C, p=probs[b], shape=(1,))
// and jax.random.choice with p is  p_cuml = cumsum(p); r = p_cuml[-1] * (1 - uniform(key, (1,)));
// searchsorted(p_cuml, r) = #{c : p_cuml[c] < r}.  One warp per (t, b): the row's softmax in registers (rank-major,
// NR per lane), the inclusive prefix sums in jnp.cumsum's association order (warp_row_tree_prefix, as for the APS
// scores), one count.  The uniform is the threefry block of the single element of keys_b[b] - bit-identical to jax.
// ------------------------------------------------------------------------------------------------
struct ClsSampleArgs {
  const float* z;         // [n_stored, B, C] logits
  const float* log_temp;  // [1] or null: outputs are logits * exp(-log_temp)
  const uint32_t* key;    // [2]
  uint32_t T, n_stored;
  int B, C;
  int32_t* y;        // [T, B]
  int32_t* idx_out;  // [T] or null
  // direct: ClassificationProbOutputLayer.sample on ONE set of outputs (the output_calib_model predictive,
  // fortuna/output_calib_model/predictive/base.py:69-105 -> fortuna/prob_output_layer/classification.py:33-51):
  //   keys_b = split(rng, B); y[:, b] = choice(keys_b[b], C, p=probs[b], shape=(T,))  - T uniforms under one key per input
  int direct;
};

template <int NR>
__global__ void __launch_bounds__(256) cls_sample_targets_kernel(const ClsSampleArgs a) {
  const int lane = threadIdx.x & 31;
  const int b = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const uint32_t t = blockIdx.y;
  if (b >= a.B) return;
  uint32_t i = 0u, bits;
  if (a.direct) {
    const uint32_t kb0 = random_bits_elem(a.key[0], a.key[1], 2u * (uint32_t)a.B, 2u * (uint32_t)b);
    const uint32_t kb1 = random_bits_elem(a.key[0], a.key[1], 2u * (uint32_t)a.B, 2u * (uint32_t)b + 1u);
    bits = random_bits_elem(kb0, kb1, a.T, t);
  } else {
    // key schedule of draw t (same in every warp of the row)
    const uint32_t kt0 = random_bits_elem(a.key[0], a.key[1], 2u * a.T, 2u * t);
    const uint32_t kt1 = random_bits_elem(a.key[0], a.key[1], 2u * a.T, 2u * t + 1u);
    uint32_t a0 = 0u, a1 = 2u, b0 = 1u, b1 = 3u;  // split(keys[t]): key1 = (a0, b0), key2 = (a1, b1)
    threefry2x32(kt0, kt1, a0, a1);
    threefry2x32(kt0, kt1, b0, b1);
    uint32_t c0 = 0u, c1 = 2u, d0 = 1u, d1 = 3u;  // choice(key1, n_stored) = randint: split(key1) -> two single draws
    threefry2x32(a0, b0, c0, c1);
    threefry2x32(a0, b0, d0, d1);
    uint32_t hi = 0u, hi1 = 0u, lo = 0u, lo1 = 0u;
    threefry2x32(c0, d0, hi, hi1);
    threefry2x32(c1, d1, lo, lo1);
    uint32_t mult = 65536u % a.n_stored;
    mult = (mult * mult) % a.n_stored;
    i = ((hi % a.n_stored) * mult + (lo % a.n_stored)) % a.n_stored;
    if (a.idx_out && b == 0 && lane == 0) a.idx_out[t] = (int32_t)i;
    // keys_b[b] = split(key2, B)[b], then its single uniform in [0, 1)
    const uint32_t kb0 = random_bits_elem(a1, b1, 2u * (uint32_t)a.B, 2u * (uint32_t)b);
    const uint32_t kb1 = random_bits_elem(a1, b1, 2u * (uint32_t)a.B, 2u * (uint32_t)b + 1u);
    bits = random_bits_elem(kb0, kb1, 1u, 0u);
  }
  const float u = __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
  // softmax of the row, rank-major: element lane * NR + r
  const float scale = a.log_temp ? expf(-a.log_temp[0]) : 1.0f;
  const float* row = a.z + ((size_t)i * a.B + b) * a.C;
  float x[NR];
  float m = -CUDART_INF_F;
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const int c = lane * NR + r;
    x[r] = c < a.C ? row[c] * scale : -CUDART_INF_F;
    m = fmaxf(m, x[r]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float ssum = 0.0f;
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    x[r] = lane * NR + r < a.C ? expf(x[r] - m) : 0.0f;
    ssum += x[r];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
#pragma unroll
  for (int r = 0; r < NR; ++r) x[r] = x[r] / ssum;
  warp_row_tree_prefix<NR>(x, lane);
  // p_cuml[-1]: the prefix at position C - 1
  const int last = a.C - 1;
  float tot = 0.0f;
#pragma unroll
  for (int r = 0; r < NR; ++r) tot = (lane * NR + r == last) ? x[r] : tot;
  tot = __shfl_sync(0xffffffffu, tot, last / NR);
  const float rr = tot * (1.0f - u);
  int cnt = 0;
#pragma unroll
  for (int r = 0; r < NR; ++r) cnt += (lane * NR + r < a.C && x[r] < rr) ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) a.y[(size_t)t * a.B + b] = cnt;
}

template <int NR>
static int launch_cls_sample(const ClsSampleArgs& a, cudaStream_t st) {
  const int64_t threads = (int64_t)a.B * 32;
  cls_sample_targets_kernel<NR><<<dim3((unsigned)((threads + 255) / 256), a.T), 256, 0, st>>>(a);
  FB200_LAUNCH_CHECK();
  return 0;
}

static int launch_row_sort(RowSortArgs a, cudaStream_t st) {
  if (!a.perm && !a.set_size && !a.targets && !a.scores && a.cs_mask && a.cs_sizes && !a.cs_all_true && !a.cs_all_false &&
      a.C > 256 && a.C <= 1024 && a.C % 4 == 0 && fb200_aligned16(a.probs) &&
      (reinterpret_cast<uintptr_t>(a.cs_mask) & 3u) == 0) {
    // sets without scores on wide rows: threshold selection instead of the full sort (rows it cannot decide: below)
    return a.C <= 512 ? launch_aps_sets_select<16>(a, st) : launch_aps_sets_select<32>(a, st);
  }
  if (!a.perm && !a.set_size && a.C <= 1024) {
    if (a.C <= 32) return launch_row_sort32_warp<1>(a, st);
    if (a.C <= 64) return launch_row_sort32_warp<2>(a, st);
    if (a.C <= 128) return launch_row_sort32_warp<4>(a, st);
    if (a.C <= 256) return launch_row_sort32_warp<8>(a, st);
    if (a.C <= 512) return launch_row_sort32_warp<16>(a, st);
    return launch_row_sort32_warp<32>(a, st);
  }
  if (a.C <= 32) return launch_row_sort_warp<1>(a, st);
  if (a.C <= 64) return launch_row_sort_warp<2>(a, st);
  if (a.C <= 128) return launch_row_sort_warp<4>(a, st);
  if (a.C <= 256) return launch_row_sort_warp<8>(a, st);
  if (a.C <= 512) return launch_row_sort_warp<16>(a, st);
  if (a.C <= 1024) return launch_row_sort_warp<32>(a, st);
  int n2 = 1;
  while (n2 < a.C) n2 <<= 1;
  a.N2 = n2;
  const size_t smem = (size_t)n2 * 8 + (size_t)n2 * 4 + (size_t)n2 * 8;  // keys, cum, prefix tree
  int threads = n2 / 2;
  if (threads < 32) threads = 32;
  if (threads > 256) threads = 256;
  cudaError_t e =
      cudaFuncSetAttribute(row_sort_cumsum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
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
      const int y = jax_gather_index(targets[i], C);
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
  {
    // NaNs of either sign sort last (lax.sort), in front of the padding only: their key is set explicitly
    uint32_t key = 0xFFFFFFFFu;
    if (i < n) {
      const float v = x[i];
      key = (v != v) ? 0xFFFFFFFEu : f32_ordered(v + 0.0f);
    }
    w[i] = key;
  }
}
__global__ void __launch_bounds__(256) sort_store_kernel(const uint32_t* __restrict__ w, int64_t n,
                                                         float* __restrict__ x) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    x[i] = w[i] >= 0xFFFFFFFEu ? CUDART_NAN_F : f32_unordered(w[i]);
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
  const float q = (all_true || all_false) ? 0.0f : conformal_boundary_score(val_sorted, n_val, n_val - 1 - K);
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

__device__ __forceinline__ unsigned long long conf_fixed(float conf) {
  // confidences are probabilities; anything negative or NaN contributes 0 (it is in no bin anyway)
  return conf > 0.0f ? (unsigned long long)((double)conf * 1099511627776.0) : 0ull;  // 2^40
}

__global__ void __launch_bounds__(256) ece_bins_kernel(const float* __restrict__ probs, int B, int C,
                                                       const int32_t* __restrict__ preds,
                                                       const int32_t* __restrict__ targets,
                                                       const float* __restrict__ thresholds, int n_bins,
                                                       int32_t* __restrict__ bin_idx, int32_t* __restrict__ counts,
                                                       int32_t* __restrict__ correct, float* __restrict__ scratch,
                                                       const float* __restrict__ brier_rows) {
  // Float sums are accumulated in 64-bit fixed point (confidences in units of 2^-40, Brier terms of 2^-38):
  // integer atomics commute, so the result does not depend on the order CTAs and warps arrive in.
  __shared__ float thr[kMaxBins];
  __shared__ int s_counts[kMaxBins];
  __shared__ int s_correct[kMaxBins];
  __shared__ unsigned long long s_conf[kMaxBins];
  __shared__ unsigned long long s_brier;
  __shared__ int s_ncorrect;
  if (threadIdx.x < kMaxBins) {
    thr[threadIdx.x] = threadIdx.x < n_bins ? thresholds[threadIdx.x] : CUDART_INF_F;
    s_counts[threadIdx.x] = 0;
    s_correct[threadIdx.x] = 0;
    s_conf[threadIdx.x] = 0ull;
  }
  if (threadIdx.x == 0) {
    s_brier = 0ull;
    s_ncorrect = 0;
  }
  __syncthreads();
  const int warp = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  const int nwarps = (int)(((int64_t)gridDim.x * blockDim.x) >> 5);
  if (C == 1) {
    // confidences given directly: one thread per data point
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += stride) {
      const float conf = probs[b];
      int bin = -1;
      for (int i = 0; i < n_bins; ++i) {
        if (conf <= thr[i]) {
          bin = i;
          break;
        }
      }
      if (bin_idx) bin_idx[b] = bin;
      const bool ok = targets && targets[b] == preds[b];
      if (bin >= 0) {
        atomicAdd(&s_counts[bin], 1);
        atomicAdd(&s_conf[bin], conf_fixed(conf));
        if (ok) atomicAdd(&s_correct[bin], 1);
      }
      if (ok) atomicAdd(&s_ncorrect, 1);
      // per-row Brier terms computed upstream (the reduction kernel's epilogue): same fixed-point accumulation
      if (brier_rows && targets) atomicAdd(&s_brier, (unsigned long long)((double)brier_rows[b] * 274877906944.0));
    }
  } else {
    // one warp per row: max / first-argmax / Brier term by shuffle reduction, lane 0 bins the row
    for (int b = warp; b < B; b += nwarps) {
      const float* row = probs + (size_t)b * C;
      const int y = targets ? targets[b] : -1;
      float best = -CUDART_INF_F;
      int best_c = 0x7FFFFFFF;
      float sq = 0.0f;
      bool has_nan = false;  // probs.max(-1) of a row holding a NaN is NaN: the row falls into no bin (classification.py:75-83)
      if ((C & 3) == 0 && ((reinterpret_cast<uintptr_t>(row) & 15u) == 0)) {
        for (int c = lane * 4; c < C; c += 128) {
          const float4 v = *reinterpret_cast<const float4*>(row + c);
          const float pv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            has_nan |= pv[k] != pv[k];
            if (pv[k] > best) {
              best = pv[k];
              best_c = c + k;
            }
            const float d = pv[k] - ((c + k) == y ? 1.0f : 0.0f);
            sq = fmaf(d, d, sq);
          }
        }
      } else {
        for (int c = lane; c < C; c += 32) {
          const float p = row[c];
          has_nan |= p != p;
          if (p > best) {
            best = p;
            best_c = c;
          }
          const float d = p - (c == y ? 1.0f : 0.0f);
          sq = fmaf(d, d, sq);
        }
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
      if (__any_sync(0xffffffffu, has_nan)) best = CUDART_NAN_F;
      if (lane == 0) {
        const int pred = preds ? preds[b] : (best_c == 0x7FFFFFFF ? 0 : best_c);
        int bin = -1;
        for (int i = 0; i < n_bins; ++i) {
          if (best <= thr[i]) {
            bin = i;
            break;
          }
        }
        if (bin_idx) bin_idx[b] = bin;
        const bool ok = targets && y == pred;
        if (bin >= 0) {
          atomicAdd(&s_counts[bin], 1);
          atomicAdd(&s_conf[bin], conf_fixed(best));
          if (ok) atomicAdd(&s_correct[bin], 1);
        }
        if (ok) atomicAdd(&s_ncorrect, 1);
        if (targets && sq == sq) atomicAdd(&s_brier, (unsigned long long)((double)sq * 274877906944.0));  // 2^38 (a NaN has no fixed-point image)
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < n_bins) {
    if (s_counts[threadIdx.x]) atomicAdd(&counts[threadIdx.x], s_counts[threadIdx.x]);
    if (s_correct[threadIdx.x]) atomicAdd(&correct[threadIdx.x], s_correct[threadIdx.x]);
    if (s_counts[threadIdx.x])
      atomicAdd(reinterpret_cast<unsigned long long*>(scratch + 4) + threadIdx.x, s_conf[threadIdx.x]);
  }
  if (threadIdx.x == 0) {
    // scratch layout of `result` until ece_finalize_kernel runs: [0..1] = u64 Brier sum (2^-38 units),
    // [2] = int #correct, [4 + 2i .. 4 + 2i + 1] = u64 confidence sum of bin i (2^-40 units)
    if (s_brier) atomicAdd(reinterpret_cast<unsigned long long*>(scratch), s_brier);
    if (s_ncorrect) atomicAdd(reinterpret_cast<int*>(scratch + 2), s_ncorrect);
  }
}

__global__ void ece_finalize_kernel(const int32_t* __restrict__ counts, float* __restrict__ conf_sums,
                                    const int32_t* __restrict__ correct, int n_bins, int B, int C,
                                    float* __restrict__ result) {
  // result[0..3] = ECE, MCE, accuracy, brier, then confs[n_bins], accs[n_bins]; on entry result[0] holds the
  // brier sum and result[1] the (int) number of correct predictions
  const float brier_sum = (float)((double)reinterpret_cast<const unsigned long long*>(result)[0] * (1.0 / 274877906944.0));
  const int n_correct = reinterpret_cast<const int*>(result)[2];
  float csum[kMaxBins];
  for (int i = 0; i < n_bins; ++i) {  // read every fixed-point accumulator before result[4..] is overwritten
    csum[i] = (float)((double)reinterpret_cast<const unsigned long long*>(result + 4)[i] * (1.0 / 1099511627776.0));
    conf_sums[i] = csum[i];
  }
  float ece = 0.0f, mce = -CUDART_INF_F;
  for (int i = 0; i < n_bins; ++i) {
    const float cnt = (float)counts[i];
    const float acc = counts[i] ? (float)correct[i] / cnt : 0.0f;    // nan_to_num(mean(empty)) = 0
    const float conf = counts[i] ? csum[i] / cnt : 0.0f;
    const float w = cnt * fabsf(acc - conf);
    ece += w;
    mce = fmaxf(mce, w);
    result[4 + i] = conf;           // per-bin mean confidence  (compute_counts_confs_accs: confs)
    result[4 + n_bins + i] = acc;   // per-bin accuracy                                    (accs)
  }
  result[0] = ece / (float)B;
  result[1] = mce;
  result[2] = (float)n_correct / (float)B;
  result[3] = C > 1 ? brier_sum / (float)B : CUDART_NAN_F;
}

// jnp.quantile(x, q, axis=0) for x[T, M]: one thread per column sorts its T values (insertion sort in local
// memory; T is the number of Monte-Carlo target samples, 30 by default) and interpolates like quantile_sorted.
constexpr int kMaxAxis0 = 256;
__global__ void __launch_bounds__(128) quantile_axis0_kernel(const float* __restrict__ x, int T, int64_t M,
                                                             const float* __restrict__ q, int nq,
                                                             float* __restrict__ out) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float v[kMaxAxis0];
  bool has_nan = false;
  for (int t = 0; t < T; ++t) {
    const float val = x[(size_t)t * M + m];
    has_nan |= (val != val);
    int i = t;
    while (i > 0 && v[i - 1] > val) {
      v[i] = v[i - 1];
      --i;
    }
    v[i] = val;
  }
  const float nm1 = (float)(T - 1);
  for (int k = 0; k < nq; ++k) {
    const float pos = __fmul_rn(q[k], nm1);
    const float low = floorf(pos), high = ceilf(pos);
    const float hw = __fsub_rn(pos, low), lw = __fsub_rn(1.0f, hw);
    const int lo = (int)fmaxf(0.0f, fminf(low, nm1)), hi = (int)fmaxf(0.0f, fminf(high, nm1));
    out[(size_t)k * M + m] = (has_nan || pos != pos) ? CUDART_NAN_F : __fadd_rn(__fmul_rn(v[lo], lw), __fmul_rn(v[hi], hw));
  }
}

// ECE / MCE from (all-reduced) bin counters: the finalisation step of the B-sharded multi-GPU scoring.
__global__ void ece_from_bins_kernel(const int32_t* __restrict__ counts, const float* __restrict__ conf_sums,
                                     const int32_t* __restrict__ correct, int n_bins, float n_total,
                                     float* __restrict__ result) {
  float ece = 0.0f, mce = -CUDART_INF_F;
  for (int i = 0; i < n_bins; ++i) {
    const float cnt = (float)counts[i];
    const float acc = counts[i] ? (float)correct[i] / cnt : 0.0f;
    const float conf = counts[i] ? conf_sums[i] / cnt : 0.0f;
    const float w = cnt * fabsf(acc - conf);
    ece += w;
    mce = fmaxf(mce, w);
  }
  result[0] = ece / n_total;
  result[1] = mce;
}

// The three additive bin counters as ONE float64 vector [counts | correct | conf_sums] (3 n_bins): a B-sharded evaluation
// all-reduces it in a single collective (int32 counts are exact in float64) and finalises from the packed sums.
__global__ void ece_pack_bins_kernel(const int32_t* __restrict__ counts, const float* __restrict__ conf_sums,
                                     const int32_t* __restrict__ correct, int n_bins, double* __restrict__ packed,
                                     int accumulate) {
  const int i = threadIdx.x;
  if (i < n_bins) {
    const double base0 = accumulate ? packed[i] : 0.0, base1 = accumulate ? packed[n_bins + i] : 0.0,
                 base2 = accumulate ? packed[2 * n_bins + i] : 0.0;
    packed[i] = base0 + (double)counts[i];
    packed[n_bins + i] = base1 + (double)correct[i];
    packed[2 * n_bins + i] = base2 + (double)conf_sums[i];
  }
}

__global__ void ece_from_packed_kernel(const double* __restrict__ packed, int n_bins, float n_total,
                                       int32_t* __restrict__ counts_out, float* __restrict__ result) {
  float ece = 0.0f, mce = -CUDART_INF_F;
  for (int i = 0; i < n_bins; ++i) {
    const float cnt = (float)packed[i];
    const bool any = packed[i] > 0.0;
    const float acc = any ? (float)packed[n_bins + i] / cnt : 0.0f;
    const float conf = any ? (float)packed[2 * n_bins + i] / cnt : 0.0f;
    const float w = cnt * fabsf(acc - conf);
    ece += w;
    mce = fmaxf(mce, w);
    if (counts_out) counts_out[i] = (int32_t)packed[i];
  }
  result[0] = ece / n_total;
  result[1] = mce;
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

static int cls_sample_common(const float* logits, const float* log_temp, const uint32_t* key, int n_target_samples,
                             int n_stored, int B, int C, int32_t* y, int32_t* idx_out, int direct, void* stream) {
  if (!logits || !key || !y) return FB200_E_NULL;
  if (n_target_samples <= 0 || n_stored <= 0 || B <= 0 || C <= 0) return FB200_E_SHAPE;
  if (C > 1024 || n_target_samples > 65535 || B >= (1 << 30)) return FB200_E_UNSUPPORTED;
  ClsSampleArgs a{};
  a.z = logits;
  a.log_temp = log_temp;
  a.key = key;
  a.T = (uint32_t)n_target_samples;
  a.n_stored = (uint32_t)n_stored;
  a.B = B;
  a.C = C;
  a.y = y;
  a.idx_out = idx_out;
  a.direct = direct;
  cudaStream_t st = (cudaStream_t)stream;
  if (C <= 32) return launch_cls_sample<1>(a, st);
  if (C <= 64) return launch_cls_sample<2>(a, st);
  if (C <= 128) return launch_cls_sample<4>(a, st);
  if (C <= 256) return launch_cls_sample<8>(a, st);
  if (C <= 512) return launch_cls_sample<16>(a, st);
  return launch_cls_sample<32>(a, st);
}

extern "C" int fb200_cls_sample_targets(const float* logits, const float* log_temp, const uint32_t* key,
                                        int n_target_samples, int n_stored, int B, int C, int32_t* y, int32_t* idx_out,
                                        void* stream) {
  return cls_sample_common(logits, log_temp, key, n_target_samples, n_stored, B, C, y, idx_out, 0, stream);
}

extern "C" int fb200_cls_output_sample_targets(const float* logits, const float* log_temp, const uint32_t* key,
                                               int n_target_samples, int B, int C, int32_t* y, void* stream) {
  return cls_sample_common(logits, log_temp, key, n_target_samples, 1, B, C, y, nullptr, 1, stream);
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

// K = largest integer count with float32(count) < threshold (count is compared as float32, base.py:51-53)
static long long conformal_rank_bound(float threshold, int n_val) {
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
  return K;
}

extern "C" int fb200_aps_conformal_sets(const float* probs, int B, int C, const float* val_scores_sorted, int n_val,
                                        float threshold, float* scores, uint8_t* mask, int32_t* sizes, void* stream) {
  if (!probs || !val_scores_sorted || !mask) return FB200_E_NULL;
  if (B <= 0 || C <= 0 || n_val <= 0) return FB200_E_SHAPE;
  if (C > 1024) return FB200_E_UNSUPPORTED;  // the fused epilogue lives in the key-only warp kernel
  const long long K = conformal_rank_bound(threshold, n_val);
  RowSortArgs a{};
  a.probs = probs;
  a.B = B;
  a.C = C;
  a.scores = scores;
  a.set_threshold = INFINITY;
  a.cs_mask = mask;
  a.cs_sizes = sizes;
  a.cs_val_sorted = val_scores_sorted;
  a.cs_n_val = n_val;
  a.cs_all_true = K >= n_val;
  a.cs_all_false = K < 0;
  a.cs_q_index = (a.cs_all_true || a.cs_all_false) ? 0 : (int)(n_val - 1 - K);
  return launch_row_sort(a, (cudaStream_t)stream);
}

extern "C" int fb200_conformal_sets(const float* val_scores_sorted, int n_val, const float* test_scores, int B,
                                    int C, float threshold, uint8_t* mask, int32_t* sizes, void* stream) {
  if (!val_scores_sorted || !test_scores || !mask) return FB200_E_NULL;
  if (n_val <= 0 || B <= 0 || C <= 0) return FB200_E_SHAPE;
  if ((C & 3) == 0 && (!fb200_aligned16(test_scores) || (reinterpret_cast<uintptr_t>(mask) & 3u))) return FB200_E_ALIGN;
  const long long K = conformal_rank_bound(threshold, n_val);
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
  if ((reinterpret_cast<uintptr_t>(result) & 7u) != 0) return FB200_E_ALIGN;  // holds 64-bit accumulators meanwhile
  if ((e = cudaMemsetAsync(result, 0, sizeof(float) * (4 + 2 * n_bins), st)) != cudaSuccess) return (int)e;
  int64_t blocks = C == 1 ? ((int64_t)B + 255) / 256 : ((int64_t)B + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  ece_bins_kernel<<<(unsigned)blocks, 256, 0, st>>>(probs, B, C, preds, targets, thresholds, n_bins, bin_idx,
                                                    counts, correct, result, nullptr);
  FB200_LAUNCH_CHECK();
  ece_finalize_kernel<<<1, 1, 0, st>>>(counts, conf_sums, correct, n_bins, B, C, result);
  FB200_LAUNCH_CHECK();
  return 0;
}

extern "C" int fb200_ece_bins_from_rows(const float* confidence, const int32_t* preds, const float* brier_terms,
                                        const int32_t* targets, int B, const float* thresholds, int n_bins,
                                        int32_t* bin_idx, int32_t* counts, float* conf_sums, int32_t* correct,
                                        float* result, void* stream) {
  if (!confidence || !preds || !thresholds || !counts || !conf_sums || !correct || !result) return FB200_E_NULL;
  if (B <= 0 || n_bins <= 0) return FB200_E_SHAPE;
  if (n_bins > kMaxBins) return FB200_E_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  if ((e = cudaMemsetAsync(counts, 0, sizeof(int32_t) * n_bins, st)) != cudaSuccess) return (int)e;
  if ((e = cudaMemsetAsync(correct, 0, sizeof(int32_t) * n_bins, st)) != cudaSuccess) return (int)e;
  if ((reinterpret_cast<uintptr_t>(result) & 7u) != 0) return FB200_E_ALIGN;
  if ((e = cudaMemsetAsync(result, 0, sizeof(float) * (4 + 2 * n_bins), st)) != cudaSuccess) return (int)e;
  int64_t blocks = ((int64_t)B + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  ece_bins_kernel<<<(unsigned)blocks, 256, 0, st>>>(confidence, B, 1, preds, targets, thresholds, n_bins, bin_idx, counts,
                                                    correct, result, brier_terms);
  FB200_LAUNCH_CHECK();
  // C = 2 makes the finaliser report the Brier score (it is undefined for bare confidences, C = 1)
  ece_finalize_kernel<<<1, 1, 0, st>>>(counts, conf_sums, correct, n_bins, B, brier_terms ? 2 : 1, result);
  FB200_LAUNCH_CHECK();
  return 0;
}

extern "C" int fb200_quantile_axis0(const float* x, int T, int64_t M, const float* q, int nq, float* out, void* stream) {
  if (!x || !q || !out) return FB200_E_NULL;
  if (T <= 0 || M <= 0 || nq <= 0) return FB200_E_SHAPE;
  if (T > kMaxAxis0) return FB200_E_UNSUPPORTED;
  quantile_axis0_kernel<<<(unsigned)((M + 127) / 128), 128, 0, (cudaStream_t)stream>>>(x, T, M, q, nq, out);
  FB200_LAUNCH_CHECK();
  return 0;
}

extern "C" int fb200_ece_from_bins(const int32_t* counts, const float* conf_sums, const int32_t* correct, int n_bins,
                                   int64_t n_total, float* result, void* stream) {
  if (!counts || !conf_sums || !correct || !result) return FB200_E_NULL;
  if (n_bins <= 0 || n_total <= 0) return FB200_E_SHAPE;
  ece_from_bins_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counts, conf_sums, correct, n_bins, (float)n_total, result);
  FB200_LAUNCH_CHECK();
  return 0;
}

extern "C" int fb200_ece_pack_bins(const int32_t* counts, const float* conf_sums, const int32_t* correct, int n_bins,
                                   double* packed, int accumulate, void* stream) {
  if (!counts || !conf_sums || !correct || !packed) return FB200_E_NULL;
  if (n_bins <= 0 || n_bins > kMaxBins) return FB200_E_SHAPE;
  ece_pack_bins_kernel<<<1, kMaxBins, 0, (cudaStream_t)stream>>>(counts, conf_sums, correct, n_bins, packed, accumulate);
  FB200_LAUNCH_CHECK();
  return 0;
}

extern "C" int fb200_ece_from_packed_bins(const double* packed, int n_bins, int64_t n_total, int32_t* counts_out,
                                          float* result, void* stream) {
  if (!packed || !result) return FB200_E_NULL;
  if (n_bins <= 0 || n_total <= 0) return FB200_E_SHAPE;
  ece_from_packed_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(packed, n_bins, (float)n_total, counts_out, result);
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
