// Shared pieces of the BMA reduction kernels (bma_reduce.cu: tiled / narrow / wide / finalise): launch arguments, row
// loads, the packed per-logit arithmetic (process_rows), the compact row-constant division.
#pragma once
#include <cuda_bf16.h>
#include <math_constants.h>

#include "async_copy.cuh"
#include "common.cuh"

namespace fb200 {

constexpr int kMaxStages = 8;
// NW consumer warps (whole warpgroups) + one producer warpgroup (one lane of it works).  The launch bound caps
// ptxas at 65536 / threads registers; the producer warpgroup gives most of its share back (setmaxnreg.dec) and the
// consumer warpgroups take it (setmaxnreg.inc):  NW = 8: 8*32*232 + 4*32*40;  12: 12*32*160 + 4*32*32;
// 16: 16*32*112 + 4*32*32.  The sum must not exceed what the CTA owns at launch (threads * the ptxas register
// count, which is 65536 / threads rounded DOWN to a multiple of 8: 168, 128, 96) or setmaxnreg.inc never returns.
template <int NW>
struct WarpCfg {
  static constexpr int kThreads = (NW + 4) * 32;
  static constexpr int kConsumerRegs = NW == 8 ? 232 : (NW == 12 ? 160 : 112);
  static constexpr int kProducerRegs = NW == 8 ? 40 : 32;
  static constexpr int kLaunchRegs = (65536 / kThreads) / 8 * 8;
  static_assert(NW * 32 * kConsumerRegs + 4 * 32 * kProducerRegs <= kThreads * kLaunchRegs,
                "setmaxnreg budget exceeds the registers the CTA owns at launch");
};

struct ClsArgs {
  const void* logits;
  const float* log_temp;
  const int32_t* targets;
  int S, B, C;
  fb200_cls_outputs_t out;
  const unsigned long long* slot_ptrs;  // shard mode: device array [n_ranks] of float* (local or peer memory)
  int n_ranks, rows_per_rank;
  int tile_rot;  // shard mode: tiles are visited starting here (staggered per rank: no owner is everyone's target at once)
  int partial_mode;
  int n_tiles;
  int n_stages;
  int ss;                      // samples per pipeline stage
  uint32_t tile_stride_bytes;  // bytes between consecutive samples inside a stage (multiple of 16)
  uint32_t stage_bytes;        // multiple of 128
  int bulk_ok;                 // base pointer and sample stride are 16-byte aligned
  float log_S;                 // float32 log(S)
};

template <typename T, int VEC>
struct RowLoad;
template <>
struct RowLoad<float, 4> {
  static __device__ __forceinline__ void load(const float* p, float (&z)[4]) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    z[0] = v.x; z[1] = v.y; z[2] = v.z; z[3] = v.w;
  }
};
template <>
struct RowLoad<float, 1> {
  static __device__ __forceinline__ void load(const float* p, float (&z)[1]) { z[0] = *p; }
};
template <>
struct RowLoad<__nv_bfloat16, 4> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&z)[4]) {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    z[0] = __uint_as_float(v.x << 16);
    z[1] = __uint_as_float(v.x & 0xFFFF0000u);
    z[2] = __uint_as_float(v.y << 16);
    z[3] = __uint_as_float(v.y & 0xFFFF0000u);
  }
};
template <>
struct RowLoad<__nv_bfloat16, 1> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&z)[1]) { z[0] = __bfloat162float(*p); }
};

__device__ __forceinline__ void st_f4(float* p, float x, float y, float z, float w) {
  asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ float ex2_approx_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx_ftz(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}


// Packed float32x2 arithmetic (Blackwell FFMA2 / FADD2 / FMUL2: one issue slot for two independent IEEE round-to-nearest
// operations on an aligned register pair).  The element-wise part of the reduction is issue-bound, not pipe-bound
// (r01i profile: 11 issued instructions per logit, 7 of them FFMA / FADD / FMUL), so pairing them removes a third of
// the instruction stream; every lane result is bit-identical to the scalar fmaf / + / * it replaces.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) { return __fmul2_rn(a, b); }

// Process RI rows (consecutive samples s of the same input b) held in shared memory: softmax of each row,
// accumulate sum_s p, sum_s p^2 and the entropy partials (sum_s p(1-p) is formed as sum p - sum p^2 at the end:
// two accumulators per (b,c) instead of three is what lets 12-16 warps share the register file).  The RI rows are independent until the final
// accumulation, so their shuffle reductions and MUFU chains overlap (ILP); accumulation keeps the s order.
// Accumulators and row values are float2 pairs (VEC = 4: two pairs per 16-byte chunk; VEC = 1: the .y half idles).
template <int VEC>
struct Pairs {
  static constexpr int kN = VEC == 4 ? 2 : 1;
};

template <typename T, int G, int NV, int VEC, bool TAILONLY, int RI>
__device__ __forceinline__ void process_rows(const T* const (&rowp)[RI], int g, int C, float k2,
                                             float2 (&sp)[NV][Pairs<VEC>::kN], float2 (&sp2)[NV][Pairs<VEC>::kN],
                                             float& spl, float& sum_l, float (&cshift)[RI], float (&l2sum)[RI]) {
  constexpr int NP = Pairs<VEC>::kN;
  float2 z[RI][NV][NP];
  float m[RI];
  // pass A: rows -> registers (padding lanes get -1e30: finite even after the log2(e)*scale multiply, so
  // 0 * u is never NaN for them, while a genuine -inf logit still yields the reference's 0 * -inf = NaN)
#pragma unroll
  for (int ri = 0; ri < RI; ++ri) {
    float pm = -CUDART_INF_F;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c0 = (g + G * j) * VEC;
      float t[VEC];
      // the maximum is taken over the row's own entries only: a padding value must not become the shift of a row whose
      // logits are all -inf (the reference's softmax of such a row is NaN, not 0)
      if ((TAILONLY && j < NV - 1) || c0 < C) {  // TAILONLY: every lane's chunk but the last is inside the row
        RowLoad<T, VEC>::load(rowp[ri] + c0, t);
        if constexpr (VEC == 4) {
          pm = fmaxf(fmaxf(pm, t[0]), t[1]);
          pm = fmaxf(fmaxf(pm, t[2]), t[3]);
        } else {
          pm = fmaxf(pm, t[0]);
        }
      } else {
#pragma unroll
        for (int k = 0; k < VEC; ++k) t[k] = -1.0e30f;
      }
      if constexpr (VEC == 4) {
        z[ri][j][0] = make_float2(t[0], t[1]);
        z[ri][j][1] = make_float2(t[2], t[3]);
      } else {
        z[ri][j][0] = make_float2(t[0], -1.0e30f);
      }
    }
    m[ri] = pm;
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1)
#pragma unroll
    for (int ri = 0; ri < RI; ++ri) m[ri] = fmaxf(m[ri], __shfl_xor_sync(0xffffffffu, m[ri], o));
  // pass B, in log2 units: u = z*k2 - c with c = fl(max*k2) (one FFMA; the rounding of c is a row constant
  // and cancels in p = e/sum), e = 2^u (MUFU.EX2), row sum, and sum_c e*u for the entropy:
  //   sum_c p ln p = ln2 * ( (sum_c e u) / sum - log2(sum) )
  float sum[RI], set[RI];
  const float2 k22 = make_float2(k2, k2);
#pragma unroll
  for (int ri = 0; ri < RI; ++ri) {
    // jsp.special.logsumexp replaces a non-finite max by 0 (m is the max of the raw logits here)
    const float ms = (fabsf(m[ri]) <= 3.4028234e38f) ? m[ri] : 0.0f;
    cshift[ri] = ms * k2;
    const float2 nc2 = make_float2(-cshift[ri], -cshift[ri]);
    float2 psum[NP], pset[NP];
#pragma unroll
    for (int k = 0; k < NP; ++k) psum[k] = pset[k] = make_float2(0.0f, 0.0f);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        const float2 u = ffma2(z[ri][j][k], k22, nc2);
        float2 ev;
        ev.x = ex2_approx_ftz(u.x);
        ev.y = (VEC == 4) ? ex2_approx_ftz(u.y) : 0.0f;
        z[ri][j][k] = ev;
        psum[k] = fadd2(psum[k], ev);
        pset[k] = ffma2(ev, u, pset[k]);
      }
    }
    if constexpr (VEC == 4) {
      sum[ri] = (psum[0].x + psum[0].y) + (psum[1].x + psum[1].y);
      set[ri] = (pset[0].x + pset[0].y) + (pset[1].x + pset[1].y);
    } else {
      sum[ri] = psum[0].x;
      set[ri] = pset[0].x;
    }
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1)
#pragma unroll
    for (int ri = 0; ri < RI; ++ri) sum[ri] += __shfl_xor_sync(0xffffffffu, sum[ri], o);
  // pass C: p = e / sum, accumulate in sample order: sum p += e * inv, sum p^2 += (e*e) * inv^2 (two FFMA2 + one FMUL2 per pair)
#pragma unroll
  for (int ri = 0; ri < RI; ++ri) {
    // jax.nn.softmax subtracts the row maximum as it is: a row whose maximum is +inf or -inf (an infinite logit, or every
    // logit -inf) is NaN in every class there - unlike logsumexp, which replaces a non-finite maximum by 0 (the shift
    // used above, and what log_prob keeps).  The normaliser carries the NaN into sum p, sum p^2 and the entropy terms.
    const float inv = (fabsf(m[ri]) <= 3.4028234e38f) ? rcp_approx_ftz(sum[ri]) : CUDART_NAN_F;
    const float inv2 = inv * inv;
    const float2 inv_2 = make_float2(inv, inv), inv2_2 = make_float2(inv2, inv2);
    l2sum[ri] = lg2_approx(sum[ri]);
    spl = fmaf(inv, set[ri], spl);  // lane-partial of sum_c p u (log2 units); -log2(sum) is row-uniform
    sum_l += l2sum[ri];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        const float2 e = z[ri][j][k];
        sp[j][k] = ffma2(e, inv_2, sp[j][k]);
        sp2[j][k] = ffma2(fmul2(e, e), inv2_2, sp2[j][k]);
      }
    }
  }
}

// x / S for a row-constant divisor S, correctly rounded like the reference's float32 division (jnp.mean = sum / S):
// r = RN(1 / S) once per row, then q = RN(x r), rem = x - S q (exact in one FMA), q' = RN(q + rem r) - Markstein's
// correction.  Three FMA-pipe instructions and no slow-path subroutine: 96 inlined IEEE divisions per row (each with its
// FCHK + CALL fallback) had made the row epilogue ~9 K instructions / 145 KB of code, more than the instruction cache.
struct DivS {
  float S, r;
};
__device__ __forceinline__ DivS make_divs(float S) { return DivS{S, 1.0f / S}; }
__device__ __forceinline__ float div_s(float x, const DivS d) {
  const float q = x * d.r;
  return fmaf(fmaf(-d.S, q, x), d.r, q);
}

// max(0, v) that keeps a NaN (jnp.maximum propagates it; fmaxf(0, NaN) is 0): the epistemic variance of an input whose
// mean is NaN is NaN in the reference.  The entropy terms likewise skip mean == 0 (an underflowed probability: 0 in the
// reference's log-domain expression too) but not a NaN mean: `!(mean <= 0)`.
__device__ __forceinline__ float clamp0_keep_nan(float v) { return v < 0.0f ? 0.0f : v; }

// element k (0..VEC-1) of chunk j of a pair-packed accumulator
template <int NV, int NP>
__device__ __forceinline__ float pk_get(const float2 (&a)[NV][NP], int j, int k) {
  return (k & 1) ? a[j][k >> 1].y : a[j][k >> 1].x;
}

static inline int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace fb200
