// Warp-level row sort building blocks of the scoring kernels (scoring.cu): order-preserving float <-> uint maps, the
// key-only 32-bit bitonic network over 32*NR register-resident keys, and the inclusive prefix sum in jnp.cumsum's
// association order (jax.lax.associative_scan).  See scoring.cu for the derivations.
#pragma once
#include <stdint.h>

namespace fb200 {

// Monotone float32 -> uint32 map (total order; -0 is canonicalised to +0 by the caller).  Every NaN, whatever its
// sign or payload, maps to the top value: lax.sort puts NaNs last, so a descending argsort has them first.  The
// NaN case is explicit because ptxas rewrites the sign-bit arithmetic below as FADD -|x|, -0 - exact for numbers,
// but a NaN leaves an FADD as 0x7FFFFFFF and would land between the negative and the positive values.
__device__ __forceinline__ uint32_t f32_ordered(float f) {
  const uint32_t u = __float_as_uint(f);
  const uint32_t o = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return (f != f) ? 0xFFFFFFFFu : o;
}
// jnp indexing with a traced index (probs[target], inv_perm[target]): a negative index counts from the end, then the
// gather clamps to the array (jax "sharp bits": jnp.arange(10)[11] == 9).  No label reads outside its row.
__device__ __forceinline__ int jax_gather_index(int y, int n) {
  y = y < 0 ? y + n : y;
  return min(max(y, 0), n - 1);
}
__device__ __forceinline__ float f32_unordered(uint32_t o) {
  const uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
  return __uint_as_float(u);
}


// Inclusive scan of one value per lane in jax.lax.associative_scan's association order (see above).
__device__ __forceinline__ float warp_tree_scan(float t, int lane) {
  float w = t;
#pragma unroll
  for (int s = 1; s < 32; s <<= 1) {  // up-sweep: the right child of each pair becomes the pair's sum
    const float o = __shfl_up_sync(0xffffffffu, w, s);
    if ((lane & (2 * s - 1)) == 2 * s - 1) w = o + w;
  }
  float out = w;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {  // down-sweep: even elements of a level add the prefix ending just before them
    const float o = __shfl_up_sync(0xffffffffu, out, s);
    if ((lane & (2 * s - 1)) == s - 1 && lane >= 3 * s - 1) out = o + out;
  }
  return out;
}

// x[NR] (rank order inside the lane's block) -> inclusive prefix sums over the whole warp-row, tree order.
__host__ __device__ constexpr int ilog2_c(int v) { return v <= 1 ? 0 : 1 + ilog2_c(v >> 1); }
// level l of the in-lane tree has NR >> l entries and starts at offset 2*NR - (2*NR >> l)
template <int NR>
__host__ __device__ constexpr int lvl_off(int l) { return 2 * NR - ((2 * NR) >> l); }

template <int NR>
__device__ __forceinline__ void warp_row_tree_prefix(float (&x)[NR], int lane) {
  constexpr int LOG = ilog2_c(NR);
  float u[2 * NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) u[r] = x[r];
#pragma unroll
  for (int l = 0; l < LOG; ++l) {
#pragma unroll
    for (int i = 0; i < (NR >> (l + 1)); ++i)
      u[lvl_off<NR>(l + 1) + i] = u[lvl_off<NR>(l) + 2 * i] + u[lvl_off<NR>(l) + 2 * i + 1];
  }
  const float incl = warp_tree_scan(u[lvl_off<NR>(LOG)], lane);
  const float carry = __shfl_up_sync(0xffffffffu, incl, 1);  // prefix through the previous lane's block
  const bool has_carry = lane > 0;
  float o[2 * NR];
  o[lvl_off<NR>(LOG)] = incl;
#pragma unroll
  for (int l = LOG - 1; l >= 0; --l) {
#pragma unroll
    for (int i = 0; i < (NR >> (l + 1)); ++i) {
      o[lvl_off<NR>(l) + 2 * i + 1] = o[lvl_off<NR>(l + 1) + i];
      if (i == 0) o[lvl_off<NR>(l)] = has_carry ? carry + u[lvl_off<NR>(l)] : u[lvl_off<NR>(l)];
      else o[lvl_off<NR>(l) + 2 * i] = o[lvl_off<NR>(l + 1) + i - 1] + u[lvl_off<NR>(l) + 2 * i];
    }
  }
#pragma unroll
  for (int r = 0; r < NR; ++r) x[r] = o[r];
}

__device__ __forceinline__ void ce_min_first(uint32_t& a, uint32_t& b) {
  const uint32_t lo = min(a, b);
  const uint32_t hi = max(a, b);
  a = lo;
  b = hi;
}

// The network in two parts.  Merge levels that stay inside a lane (k < NR) are straight-line code on static
// register indices.  From k = NR on, every level is: complement the lane's keys if its block changes direction, the
// cross-lane stages (one shuffle + one predicated min/max per key; the lane mask is a loop variable), then the same
// in-lane merge j = NR/2..1 for every k - so these levels are ONE rolled loop body.  Fully unrolled, the kernel was
// 90 KB of straight-line code and stalled on instruction fetch more than on anything else (icache hit rate 61 %).
template <int NR>
__device__ __forceinline__ void warp_bitonic_sort32(uint32_t (&key)[NR], int lane) {
  constexpr int N = 32 * NR;
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
  uint32_t flipped = 0u;  // all-ones while the lane's keys are stored complemented
#pragma unroll 1
  for (int kl = 1; kl <= 32; kl <<= 1) {  // k = kl * NR
    const uint32_t want = (kl < 32 && (lane & kl) != 0) ? 0xFFFFFFFFu : 0u;
    const uint32_t x = want ^ flipped;
    flipped = want;
#pragma unroll
    for (int r = 0; r < NR; ++r) key[r] ^= x;
#pragma unroll 1
    for (int lmask = kl >> 1; lmask > 0; lmask >>= 1) {
      const bool upper = (lane & lmask) != 0;  // the upper lane of a pair keeps the larger key
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const uint32_t other = __shfl_xor_sync(0xffffffffu, key[r], lmask);
        key[r] = upper ? max(key[r], other) : min(key[r], other);
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


}  // namespace fb200
