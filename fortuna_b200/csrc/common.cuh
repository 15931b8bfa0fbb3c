// Shared device helpers for libfortuna_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fortuna_b200.h"

#define FB200_LAUNCH_CHECK()                      \
  do {                                            \
    cudaError_t e__ = cudaPeekAtLastError();      \
    if (e__ != cudaSuccess) return (int)e__;      \
  } while (0)

static inline bool fb200_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

namespace fb200 {

// ------------------------------------------------------------------------------------------------
// threefry2x32, 20 rounds, jax/_src/prng.py layout.  rotl is one SHF.L.W; xors are LOP3.
// (Round 2 tried taking the rotation from the 64-bit product x1 * 2^r - one IMAD.WIDE on the FMA pipe, the OR of its two
// words folded into the XOR's LOP3 - to relieve the half-rate ALU pipe K1 is co-limited by: 0.381 ms instead of 0.355 ms
// per c5 step on the same box, scripts/ab_lib.sh + scripts/bench_sampler.py.  The plain form stays.)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return __funnelshift_l(x, x, r); }

#define FB200_TF_ROUND(r) \
  x0 += x1;               \
  x1 = rotl32(x1, r);     \
  x1 ^= x0;

__device__ __forceinline__ void threefry2x32(uint32_t k0, uint32_t k1, uint32_t& x0, uint32_t& x1) {
  const uint32_t k2 = k0 ^ k1 ^ 0x1BD11BDAu;
  x0 += k0;
  x1 += k1;
  FB200_TF_ROUND(13) FB200_TF_ROUND(15) FB200_TF_ROUND(26) FB200_TF_ROUND(6)
  x0 += k1;
  x1 += k2 + 1u;
  FB200_TF_ROUND(17) FB200_TF_ROUND(29) FB200_TF_ROUND(16) FB200_TF_ROUND(24)
  x0 += k2;
  x1 += k0 + 2u;
  FB200_TF_ROUND(13) FB200_TF_ROUND(15) FB200_TF_ROUND(26) FB200_TF_ROUND(6)
  x0 += k0;
  x1 += k1 + 3u;
  FB200_TF_ROUND(17) FB200_TF_ROUND(29) FB200_TF_ROUND(16) FB200_TF_ROUND(24)
  x0 += k1;
  x1 += k2 + 4u;
  FB200_TF_ROUND(13) FB200_TF_ROUND(15) FB200_TF_ROUND(26) FB200_TF_ROUND(6)
  x0 += k2;
  x1 += k0 + 5u;
}

// Element e of jax's random_bits(key, n): counters iota(n) (padded with one 0 when n is odd), first
// half in lane 0, second half in lane 1.  h = ceil(n/2).
__device__ __forceinline__ uint32_t random_bits_elem(uint32_t k0, uint32_t k1, uint32_t n, uint32_t e) {
  const uint32_t h = (n >> 1) + (n & 1u);
  uint32_t x0, x1;
  if (e < h) {
    x0 = e;
    x1 = (e + h < n) ? e + h : 0u;
    threefry2x32(k0, k1, x0, x1);
    return x0;
  }
  x0 = e - h;
  x1 = e;
  threefry2x32(k0, k1, x0, x1);
  return x1;
}

// ------------------------------------------------------------------------------------------------
// bits -> N(0,1), jax/_src/random.py _uniform + _normal_real + XLA ErfInv32 (Giles).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float bits_to_open_unit(uint32_t bits) {
  // f in [0,1): bitcast((bits >> 9) | 0x3F800000) - 1; u = max(lo, f*(hi-lo)+lo), lo = nextafter(-1,0), hi-lo == 2.0f.
  const float f = __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
  const float lo = -0.99999994f;
  return fmaxf(lo, fmaf(f, 2.0f, lo));
}

__device__ __forceinline__ float erfinv_giles(float x) {
  // w = -log1p(-(x*x)) with x*x rounded first (XLA does not fuse it); 1 - t is exact for t >= 0.5 and
  // within half an ulp below, so -log(1-t) via MUFU.LG2 is within ~2e-7 absolute of log1p.
  const float t = __fmul_rn(x, x);
  float w = -0.6931471805599453f * lg2_approx(1.0f - t);
  float p;
  if (w < 5.0f) {
    w = w - 2.5f;
    p = 2.81022636e-08f;
    p = fmaf(p, w, 3.43273939e-07f);
    p = fmaf(p, w, -3.5233877e-06f);
    p = fmaf(p, w, -4.39150654e-06f);
    p = fmaf(p, w, 0.00021858087f);
    p = fmaf(p, w, -0.00125372503f);
    p = fmaf(p, w, -0.00417768164f);
    p = fmaf(p, w, 0.246640727f);
    p = fmaf(p, w, 1.50140941f);
  } else {
    w = sqrtf(w) - 3.0f;
    p = -0.000200214257f;
    p = fmaf(p, w, 0.000100950558f);
    p = fmaf(p, w, 0.00134934322f);
    p = fmaf(p, w, -0.00367342844f);
    p = fmaf(p, w, 0.00573950773f);
    p = fmaf(p, w, -0.0076224613f);
    p = fmaf(p, w, 0.00943887047f);
    p = fmaf(p, w, 1.00167406f);
    p = fmaf(p, w, 2.83297682f);
  }
  return p * x;
}

__device__ __forceinline__ float bits_to_normal(uint32_t bits) {
  return 1.41421356237309505f * erfinv_giles(bits_to_open_unit(bits));
}

// ------------------------------------------------------------------------------------------------
// warp helpers
// ------------------------------------------------------------------------------------------------
template <int W>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int W>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace fb200
