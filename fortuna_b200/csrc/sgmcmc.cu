// K1: fused SG-MCMC sampler step over the flattened parameter pytree.
//
// Replaces (reference file:line)
//   fortuna/prob_model/posterior/sgmcmc/sghmc/sghmc_integrator.py:59-92     update_fn
//   fortuna/prob_model/posterior/sgmcmc/sgmcmc_preconditioner.py:65-93      RMSprop / identity
//   fortuna/utils/random.py:11-23                                           per-leaf N(0,1) tree
//   fortuna/prob_model/posterior/sgmcmc/cyclical_sgld/cyclical_sgld_integrator.py:64-100
//   flax TrainState.apply_gradients -> optax.apply_updates (fortuna/training/trainer.py:170)
//
// One launch reads theta, g, m (and v) once and writes theta, m (and v) once: 20 B/param (identity),
// 28 B/param (RMSprop).  The noise never touches HBM: each thread evaluates the threefry2x32 block
// of counter pair (i, i + h) of its leaf, whose two outputs are jax's normal() draws for elements i
// and i + h of that leaf (h = ceil(len/2)), so a thread updates two 16-byte vectors, one in each half.
#include <stdlib.h>

#include "common.cuh"

namespace fb200 {

// ------------------------------------------------------------------------------------------------
// key schedule: (key, new_key) = split(key_in); leaf_keys = split(key, n_leaves); key_out = new_key
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) sgmcmc_keys_kernel(const uint32_t* __restrict__ key_in,
                                                          uint32_t* __restrict__ key_out,
                                                          uint32_t* __restrict__ leaf_keys, uint32_t n_leaves) {
  const uint32_t k0 = key_in[0], k1 = key_in[1];
  // split(key_in) = random_bits(key_in, 4).reshape(2,2): blocks (0,2) and (1,3)
  uint32_t a0 = 0u, a1 = 2u, b0 = 1u, b1 = 3u;
  threefry2x32(k0, k1, a0, a1);
  threefry2x32(k0, k1, b0, b1);
  // key = (a0, b0) drives the noise; new_key = (a1, b1) is stored
  const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l == 0) {
    key_out[0] = a1;
    key_out[1] = b1;
  }
  if (l < n_leaves) {
    leaf_keys[2 * l] = random_bits_elem(a0, b0, 2u * n_leaves, 2u * l);
    leaf_keys[2 * l + 1] = random_bits_elem(a0, b0, 2u * n_leaves, 2u * l + 1u);
  }
}

constexpr int64_t kFoldKeysMaxTiles = 1024;  // ~2 M parameters: below this the step is launch-bound, not bandwidth-bound

struct StepCoef {
  float sqrt_eps;     // sqrt(float32(step_size))
  float beta;         // float32(momentum_decay)
  float noise_scale;  // sqrt(float32(2 * (1 - momentum_decay)))
  float alpha;        // float32(running_average_factor)
  float one_m_alpha;  // float32(1 - running_average_factor)
  float rms_eps;      // float32(eps)
  float step_size;    // float32 epsilon_t (exploration phase)
  int zero_momentum;
  const float* grad_mult;  // device float[1] or null: clip_grandients_by_norm's multiplier (utils/training.py:14-20)
  int grad_bf16;           // `grad` holds bfloat16 (single-GPU kernels only)
};

// four consecutive bf16 gradients (8-byte aligned) -> float4
__device__ __forceinline__ float4 ld_bf16x4(const float* grad_as_bf16, int64_t e) {
  const uint2 w = *reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned short*>(grad_as_bf16) + e);
  return make_float4(__uint_as_float(w.x << 16), __uint_as_float(w.x & 0xFFFF0000u), __uint_as_float(w.y << 16),
                     __uint_as_float(w.y & 0xFFFF0000u));
}
__device__ __forceinline__ float ld_bf16(const float* grad_as_bf16, int64_t e) {
  return __uint_as_float((uint32_t)reinterpret_cast<const unsigned short*>(grad_as_bf16)[e] << 16);
}

__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// One parameter: returns the update u; m and v are updated in place (registers).  Every product and sum is rounded
// on its own (__fmul_rn / __fadd_rn: no FMA contraction), as the reference's op-by-op float32 graph does - and so that
// every instantiation of the step kernel (single-GPU, log-joint, data-parallel) computes bit-identical values.
template <bool RMS>
__device__ __forceinline__ float sghmc_elem(float g, float& m, float& v, uint32_t bits, const StepCoef& c) {
  float n = bits_to_normal(bits);
  float denom = 1.0f;
  if (RMS) {
    v = __fadd_rn(__fmul_rn(v, c.alpha), __fmul_rn(__fmul_rn(g, g), c.one_m_alpha));  // sgmcmc_preconditioner.py:67-68
    const float s = sqrt_approx(v);
    denom = __fadd_rn(c.rms_eps, s);            // :76
    n = __fmul_rn(n, sqrt_approx(denom));       // :83  multiply_by_m_sqrt
  }
  const float m_old = c.zero_momentum ? 0.0f : m;
  m = __fadd_rn(__fadd_rn(__fmul_rn(c.beta, m_old), __fmul_rn(g, c.sqrt_eps)),
                __fmul_rn(n, c.noise_scale));               // sghmc_integrator.py:77-84
  float u = m;
  if (RMS) u = __fdividef(u, denom);                        // :85  multiply_by_m_inv
  return __fmul_rn(u, c.sqrt_eps);                          // :86
}

// Gradient prologue of the fused log-joint mode (JOINT): `grad` holds the gradient of the batch log-likelihood SUM and
// the kernel forms the log-joint gradient of Joint._batched_log_joint_prob (fortuna/prob_model/joint/base.py:54-122,
// prior/gaussian.py:29-32) itself:  g = (n_data / batch) * grad - prec * theta  - theta is loaded anyway, so the prior
// costs no HBM traffic and no autograd passes.  The CTA also leaves sum(theta^2) of its tile in sq_part[blockIdx.x]
// (double), from which the caller gets the log-prior VALUE of the pre-update parameters for the logged loss.
struct JointCoef {
  float lik_scale;   // n_data / batch_size
  float prior_coef;  // -exp(-log_var)
  double* sq_part;   // [n_tiles] or NULL
};

// Data-parallel mode (DP): the gradient all-reduce and the parameter broadcast of the reference's multi-device trainer
// (lax.pmean of the per-device gradients, fortuna/training/trainer.py:793-795, then the replicated optimizer step) are
// folded into the step.  Every rank owns a contiguous range of tiles; for its elements it LOADS the gradient from every
// rank's buffer over NVLink (peer pointers into symmetric memory), averages in rank order, updates its slice of the
// momentum / preconditioner state, and STORES the new parameters into every rank's replica.  Per rank that is the
// traffic of a reduce-scatter plus an all-gather, but with the sampler arithmetic riding inside it and the optimizer
// state touched once per element across the whole job instead of once per rank.
struct DpPeers {
  const float* grad[FB200_MAX_PEERS];
  float* params[FB200_MAX_PEERS];
  const float* mc_grad;  // multicast (NVLS) addresses of the same buffers, or null: one multimem.ld_reduce returns the
  float* mc_params;      // switch-side sum over all ranks, one multimem.st lands in every replica
  int n;
  float inv_n;
};

// NVLink SHARP (NVLS) accesses through a multicast address: the NVSwitch performs the reduction / the replication, so
// an owner receives n/N gradients instead of (N-1) n/N and sends its parameters once instead of N-1 times.
__device__ __forceinline__ float4 mc_ld_reduce4(const float* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ float mc_ld_reduce1(const float* p) {
  float v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void mc_st4(float* p, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void mc_st1(float* p, float v) {
  asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// NP = compile-time bound on the ranks (2, 4 or 8): the loads of all ranks are issued before the first add, and the
// registers that holds them in flight should not be paid for ranks that do not exist.
template <int NP>
__device__ __forceinline__ float4 dp_grad4(const DpPeers& dp, int64_t e) {
  float4 v[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p)
    if (p < dp.n) v[p] = *reinterpret_cast<const float4*>(dp.grad[p] + e);
  float4 a = v[0];
#pragma unroll
  for (int p = 1; p < NP; ++p)
    if (p < dp.n) {
      a.x += v[p].x; a.y += v[p].y; a.z += v[p].z; a.w += v[p].w;
    }
  a.x *= dp.inv_n; a.y *= dp.inv_n; a.z *= dp.inv_n; a.w *= dp.inv_n;
  return a;
}
__device__ __forceinline__ float dp_grad1(const DpPeers& dp, int64_t e) {
  float a = dp.grad[0][e];
  for (int p = 1; p < dp.n; ++p) a += dp.grad[p][e];
  return a * dp.inv_n;
}

template <bool RMS, bool JOINT, int NP>  // NP = 0: single GPU; 2 / 4 / 8: data-parallel over up to NP ranks (peer
                                         // loads / stores); 1: data-parallel through the multicast addresses (NVLS)
__global__ void __launch_bounds__(256)
    sgmcmc_step_kernel(float* __restrict__ params, const float* __restrict__ grad, float* __restrict__ mom,
                       float* __restrict__ pv, float* __restrict__ upd, const fb200_leaf_t* __restrict__ leaves,
                       const uint32_t* __restrict__ key_in, uint32_t* __restrict__ key_out,
                       uint32_t* __restrict__ leaf_keys, uint32_t n_leaves, const fb200_tile_t* __restrict__ tiles,
                       const StepCoef c, const JointCoef jc, const DpPeers dp) {
  constexpr bool DP = NP > 0;
  constexpr bool MC = NP == 1;
  const fb200_tile_t tile = tiles[blockIdx.x];
  const fb200_leaf_t leaf = leaves[tile.leaf];
  // Key schedule.  Large trees: a tiny kernel of its own has already written leaf_keys (key_in == null here).  Small
  // trees are launch-bound and that second launch was a third of the step, so there the schedule is derived right here,
  // once per WARP and without a barrier: (key, new_key) = split(key_in) - blocks (0,2) and (1,3) of random_bits(key_in, 4),
  // one per lane - then this leaf's key = elements 2l, 2l+1 of random_bits(key, 2 n_leaves), again one per lane,
  // broadcast by shuffles (two dependent threefry blocks on two lanes: ~0.4 us per CTA, which matters at 53 K tiles and
  // not at 30).
  uint2 key;
  if (key_in == nullptr) {
    key = reinterpret_cast<const uint2*>(leaf_keys)[tile.leaf];
  } else {
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t kw = 0u;
    if (lane < 2u) {
      const uint32_t k0 = key_in[0], k1 = key_in[1];
      uint32_t x0 = lane, x1 = lane + 2u;
      threefry2x32(k0, k1, x0, x1);                          // lane 0: (a0, a1); lane 1: (b0, b1)
      const uint32_t o0 = __shfl_xor_sync(0x3u, x0, 1), o1 = __shfl_xor_sync(0x3u, x1, 1);
      const uint32_t a0 = lane ? o0 : x0, b0 = lane ? x0 : o0;  // key = (a0, b0) drives the noise
      if (blockIdx.x == 0 && threadIdx.x == 0) {
        key_out[0] = x1;                                     // new_key = (a1, b1) is stored
        key_out[1] = o1;
      }
      kw = random_bits_elem(a0, b0, 2u * n_leaves, 2u * (uint32_t)tile.leaf + lane);
      if (tile.pair_start == 0 && threadIdx.x < 2) leaf_keys[2 * tile.leaf + lane] = kw;  // split(key, L), inspectable
    }
    key.x = __shfl_sync(0xffffffffu, kw, 0);
    key.y = __shfl_sync(0xffffffffu, kw, 1);
  }
  const uint32_t n = (uint32_t)leaf.len;
  const uint32_t h = (n >> 1) + (n & 1u);
  const uint32_t i0 = tile.pair_start + threadIdx.x * 4u;
  float sq = 0.0f;  // JOINT: this thread's sum of theta^2
  const bool active = i0 < h;
  if (!JOINT && !active) return;
  if (active) {
  const int64_t e0 = leaf.offset + i0;      // first-half elements e0 .. e0+3
  const int64_t e1 = leaf.offset + h + i0;  // second-half elements e1 .. e1+3

  uint32_t r0[4], r1[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t i = i0 + k;
    r0[k] = i;
    r1[k] = (i + h < n) ? i + h : 0u;  // odd-size pad counter is 0
    threefry2x32(key.x, key.y, r0[k], r1[k]);
  }

  const bool vec = ((leaf.offset & 3) == 0) && ((h & 3u) == 0) && ((uint64_t)i0 + h + 3u < (uint64_t)n);
  if (vec) {
    float4 g0, g1;
    if constexpr (MC) {
      g0 = mc_ld_reduce4(dp.mc_grad + e0);
      g1 = mc_ld_reduce4(dp.mc_grad + e1);
      g0.x *= dp.inv_n; g0.y *= dp.inv_n; g0.z *= dp.inv_n; g0.w *= dp.inv_n;
      g1.x *= dp.inv_n; g1.y *= dp.inv_n; g1.z *= dp.inv_n; g1.w *= dp.inv_n;
    } else if constexpr (DP) {
      g0 = dp_grad4<NP>(dp, e0);
      g1 = dp_grad4<NP>(dp, e1);
    } else if (c.grad_bf16) {
      g0 = ld_bf16x4(grad, e0);
      g1 = ld_bf16x4(grad, e1);
    } else {
      g0 = *reinterpret_cast<const float4*>(grad + e0);
      g1 = *reinterpret_cast<const float4*>(grad + e1);
    }
    float4 m0 = *reinterpret_cast<const float4*>(mom + e0);
    float4 m1 = *reinterpret_cast<const float4*>(mom + e1);
    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
    if (RMS) {
      v0 = *reinterpret_cast<const float4*>(pv + e0);
      v1 = *reinterpret_cast<const float4*>(pv + e1);
    }
    float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = p0;
    if (params) {
      p0 = *reinterpret_cast<const float4*>(params + e0);
      p1 = *reinterpret_cast<const float4*>(params + e1);
    }
    if (JOINT) {
      g0.x = fmaf(jc.lik_scale, g0.x, jc.prior_coef * p0.x); g0.y = fmaf(jc.lik_scale, g0.y, jc.prior_coef * p0.y);
      g0.z = fmaf(jc.lik_scale, g0.z, jc.prior_coef * p0.z); g0.w = fmaf(jc.lik_scale, g0.w, jc.prior_coef * p0.w);
      g1.x = fmaf(jc.lik_scale, g1.x, jc.prior_coef * p1.x); g1.y = fmaf(jc.lik_scale, g1.y, jc.prior_coef * p1.y);
      g1.z = fmaf(jc.lik_scale, g1.z, jc.prior_coef * p1.z); g1.w = fmaf(jc.lik_scale, g1.w, jc.prior_coef * p1.w);
      sq = ((p0.x * p0.x + p0.y * p0.y) + (p0.z * p0.z + p0.w * p0.w)) +
           ((p1.x * p1.x + p1.y * p1.y) + (p1.z * p1.z + p1.w * p1.w));
    }
    if (c.grad_mult) {  // clip_grandients_by_norm: g <- mult * g (the multiplier comes from fb200_grad_clip_mult)
      const float gm = __ldg(c.grad_mult);
      g0.x = __fmul_rn(gm, g0.x); g0.y = __fmul_rn(gm, g0.y); g0.z = __fmul_rn(gm, g0.z); g0.w = __fmul_rn(gm, g0.w);
      g1.x = __fmul_rn(gm, g1.x); g1.y = __fmul_rn(gm, g1.y); g1.z = __fmul_rn(gm, g1.z); g1.w = __fmul_rn(gm, g1.w);
    }
    float4 u0, u1;
    u0.x = sghmc_elem<RMS>(g0.x, m0.x, v0.x, r0[0], c);
    u0.y = sghmc_elem<RMS>(g0.y, m0.y, v0.y, r0[1], c);
    u0.z = sghmc_elem<RMS>(g0.z, m0.z, v0.z, r0[2], c);
    u0.w = sghmc_elem<RMS>(g0.w, m0.w, v0.w, r0[3], c);
    u1.x = sghmc_elem<RMS>(g1.x, m1.x, v1.x, r1[0], c);
    u1.y = sghmc_elem<RMS>(g1.y, m1.y, v1.y, r1[1], c);
    u1.z = sghmc_elem<RMS>(g1.z, m1.z, v1.z, r1[2], c);
    u1.w = sghmc_elem<RMS>(g1.w, m1.w, v1.w, r1[3], c);
    *reinterpret_cast<float4*>(mom + e0) = m0;
    *reinterpret_cast<float4*>(mom + e1) = m1;
    if (RMS) {
      *reinterpret_cast<float4*>(pv + e0) = v0;
      *reinterpret_cast<float4*>(pv + e1) = v1;
    }
    if (params) {
      p0.x = __fadd_rn(p0.x, u0.x); p0.y = __fadd_rn(p0.y, u0.y); p0.z = __fadd_rn(p0.z, u0.z); p0.w = __fadd_rn(p0.w, u0.w);
      p1.x = __fadd_rn(p1.x, u1.x); p1.y = __fadd_rn(p1.y, u1.y); p1.z = __fadd_rn(p1.z, u1.z); p1.w = __fadd_rn(p1.w, u1.w);
      if constexpr (MC) {
        mc_st4(dp.mc_params + e0, p0);
        mc_st4(dp.mc_params + e1, p1);
      } else if constexpr (DP) {
#pragma unroll
        for (int p = 0; p < NP; ++p)
          if (p < dp.n) {
            *reinterpret_cast<float4*>(dp.params[p] + e0) = p0;
            *reinterpret_cast<float4*>(dp.params[p] + e1) = p1;
          }
      } else {
        *reinterpret_cast<float4*>(params + e0) = p0;
        *reinterpret_cast<float4*>(params + e1) = p1;
      }
    }
    if (upd) {
      *reinterpret_cast<float4*>(upd + e0) = u0;
      *reinterpret_cast<float4*>(upd + e1) = u1;
    }
  } else {
    // ragged leaf (odd size, unaligned offset or half): scalar path, same arithmetic
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t i = i0 + k;
      if (i >= h) break;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        if (half == 1 && !(i + h < n)) continue;
        const int64_t e = half ? e1 + k : e0 + k;
        float m = mom[e];
        float v = RMS ? pv[e] : 0.f;
        float g;
        if constexpr (MC) g = mc_ld_reduce1(dp.mc_grad + e) * dp.inv_n;
        else if constexpr (DP) g = dp_grad1(dp, e);
        else g = c.grad_bf16 ? ld_bf16(grad, e) : grad[e];
        if (JOINT) {
          const float p = params[e];
          g = fmaf(jc.lik_scale, g, jc.prior_coef * p);
          sq += p * p;
        }
        if (c.grad_mult) g = __fmul_rn(__ldg(c.grad_mult), g);
        const float u = sghmc_elem<RMS>(g, m, v, half ? r1[k] : r0[k], c);
        mom[e] = m;
        if (RMS) pv[e] = v;
        if constexpr (MC) {
          mc_st1(dp.mc_params + e, __fadd_rn(params[e], u));
        } else if constexpr (DP) {
          const float pn = __fadd_rn(params[e], u);
          for (int p = 0; p < dp.n; ++p) dp.params[p][e] = pn;
        } else if (params) params[e] = __fadd_rn(params[e], u);
        if (upd) upd[e] = u;
      }
    }
  }
  }  // active
  if (JOINT && jc.sq_part) {
    __shared__ float wsum[8];
    sq = group_sum<32>(sq);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += (double)wsum[w];
      jc.sq_part[blockIdx.x] = t;
    }
  }
}

// out[0] = sum of the per-CTA partials: 1024 threads take strided slices (independent loads, 4-way unrolled), then a
// fixed-shape tree in shared memory - the same association every run (deterministic), a few microseconds for the 53 K
// tiles of a 110 M-parameter tree.
__global__ void __launch_bounds__(1024) sum_partials_kernel(const double* __restrict__ part, int64_t n,
                                                            double* __restrict__ out) {
  __shared__ double sh[1024];
  double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
  int64_t i = threadIdx.x;
  for (; i + 3 * 1024 < n; i += 4 * 1024) {
    t0 += part[i];
    t1 += part[i + 1024];
    t2 += part[i + 2 * 1024];
    t3 += part[i + 3 * 1024];
  }
  for (; i < n; i += 1024) t0 += part[i];
  sh[threadIdx.x] = (t0 + t1) + (t2 + t3);
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0];
}

// Exploration phase of cyclical SGLD: no noise, no momentum (cyclical_sgld_integrator.py:65-84).  Elements
// [begin, end); NP as in sgmcmc_step_kernel (0: single GPU; 2 / 4 / 8: gradients averaged over the ranks by peer loads
// and the new parameters stored into every replica; 1: the same through the NVLS multicast addresses).  Products and
// sums are rounded one by one so that all instantiations agree bit for bit.
template <bool RMS, bool JOINT, int NP>
__global__ void __launch_bounds__(256)
    sgd_explore_kernel(float* __restrict__ params, const float* __restrict__ grad, float* __restrict__ pv,
                       float* __restrict__ upd, int64_t begin, int64_t n, const StepCoef c, const JointCoef jc,
                       const DpPeers dp) {
  constexpr bool DP = NP > 0;
  constexpr bool MC = NP == 1;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  float sq = 0.0f;
  for (int64_t e = begin + ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; e < n; e += stride) {
    if (e + 3 < n) {
      float4 g;
      if constexpr (MC) {
        g = mc_ld_reduce4(dp.mc_grad + e);
        g.x *= dp.inv_n; g.y *= dp.inv_n; g.z *= dp.inv_n; g.w *= dp.inv_n;
      } else if constexpr (DP) {
        g = dp_grad4<NP>(dp, e);
      } else {
        g = *reinterpret_cast<const float4*>(grad + e);
      }
      float gg[4] = {g.x, g.y, g.z, g.w};
      float pp[4] = {0.f, 0.f, 0.f, 0.f};
      if (params) {
        const float4 p = *reinterpret_cast<const float4*>(params + e);
        pp[0] = p.x; pp[1] = p.y; pp[2] = p.z; pp[3] = p.w;
      }
      if (JOINT) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          gg[k] = fmaf(jc.lik_scale, gg[k], __fmul_rn(jc.prior_coef, pp[k]));
          sq += pp[k] * pp[k];
        }
      }
      if (c.grad_mult) {
        const float gm = __ldg(c.grad_mult);
#pragma unroll
        for (int k = 0; k < 4; ++k) gg[k] = __fmul_rn(gm, gg[k]);
      }
      float uu[4];
      float vv[4] = {0.f, 0.f, 0.f, 0.f};
      if (RMS) {
        const float4 v = *reinterpret_cast<const float4*>(pv + e);
        vv[0] = v.x; vv[1] = v.y; vv[2] = v.z; vv[3] = v.w;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        // -1.0 * step * g, then optax.sgd(1.0) = scale(-1.0): exact sign flips
        float u = __fmul_rn(c.step_size, gg[k]);
        if (RMS) {
          vv[k] = __fadd_rn(__fmul_rn(vv[k], c.alpha), __fmul_rn(__fmul_rn(gg[k], gg[k]), c.one_m_alpha));
          u = __fdividef(u, __fadd_rn(c.rms_eps, sqrt_approx(vv[k])));
        }
        uu[k] = u;
      }
      if (RMS) *reinterpret_cast<float4*>(pv + e) = make_float4(vv[0], vv[1], vv[2], vv[3]);
      if (params) {
        const float4 p = make_float4(__fadd_rn(pp[0], uu[0]), __fadd_rn(pp[1], uu[1]), __fadd_rn(pp[2], uu[2]),
                                     __fadd_rn(pp[3], uu[3]));
        if constexpr (MC) {
          mc_st4(dp.mc_params + e, p);
        } else if constexpr (DP) {
#pragma unroll
          for (int q = 0; q < NP; ++q)
            if (q < dp.n) *reinterpret_cast<float4*>(dp.params[q] + e) = p;
        } else {
          *reinterpret_cast<float4*>(params + e) = p;
        }
      }
      if (upd) *reinterpret_cast<float4*>(upd + e) = make_float4(uu[0], uu[1], uu[2], uu[3]);
    } else {
      for (int64_t j = e; j < n; ++j) {
        float g;
        if constexpr (MC) g = mc_ld_reduce1(dp.mc_grad + j) * dp.inv_n;
        else if constexpr (DP) g = dp_grad1(dp, j);
        else g = grad[j];
        const float p = params ? params[j] : 0.0f;
        if (JOINT) {
          g = fmaf(jc.lik_scale, g, __fmul_rn(jc.prior_coef, p));
          sq += p * p;
        }
        if (c.grad_mult) g = __fmul_rn(__ldg(c.grad_mult), g);
        float u = __fmul_rn(c.step_size, g);
        if (RMS) {
          const float v = __fadd_rn(__fmul_rn(pv[j], c.alpha), __fmul_rn(__fmul_rn(g, g), c.one_m_alpha));
          pv[j] = v;
          u = __fdividef(u, __fadd_rn(c.rms_eps, sqrt_approx(v)));
        }
        if (params) {
          const float pn = __fadd_rn(p, u);
          if constexpr (MC) {
            mc_st1(dp.mc_params + j, pn);
          } else if constexpr (DP) {
            for (int q = 0; q < dp.n; ++q) dp.params[q][j] = pn;
          } else {
            params[j] = pn;
          }
        }
        if (upd) upd[j] = u;
      }
    }
  }
  if (JOINT && jc.sq_part) {
    __shared__ float wsum[8];
    sq = group_sum<32>(sq);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += (double)wsum[w];
      jc.sq_part[blockIdx.x] = t;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// clip_grandients_by_norm (fortuna/utils/training.py:14-20): the global norm of the gradient the step is about to use,
// as float64 partials per CTA (summed in order by sum_partials_kernel: deterministic), then
// mult = min(1, max_norm / (1e-7 + sqrt(sum))) in float32.  The step kernels scale by *mult as they load the gradient.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) grad_sq_partials_kernel(const float* __restrict__ grad, int grad_bf16,
                                                               const float* __restrict__ params, const JointCoef jc,
                                                               int joint, int64_t n, double* __restrict__ part) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  double acc = 0.0;
  for (int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; e < n; e += stride) {
    float g[4] = {0.f, 0.f, 0.f, 0.f}, p[4] = {0.f, 0.f, 0.f, 0.f};
    if (e + 3 < n) {
      const float4 gv = grad_bf16 ? ld_bf16x4(grad, e) : *reinterpret_cast<const float4*>(grad + e);
      g[0] = gv.x; g[1] = gv.y; g[2] = gv.z; g[3] = gv.w;
      if (joint) {
        const float4 pv = *reinterpret_cast<const float4*>(params + e);
        p[0] = pv.x; p[1] = pv.y; p[2] = pv.z; p[3] = pv.w;
      }
    } else {
      for (int k = 0; k < 4 && e + k < n; ++k) {
        g[k] = grad_bf16 ? ld_bf16(grad, e + k) : grad[e + k];
        if (joint) p[k] = params[e + k];
      }
    }
    float s4 = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gj = joint ? fmaf(jc.lik_scale, g[k], jc.prior_coef * p[k]) : g[k];  // as the step kernel forms it
      s4 = fmaf(gj, gj, s4);
    }
    acc += (double)s4;
  }
  __shared__ double sh[256];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}

__global__ void grad_clip_finish_kernel(const double* __restrict__ sum_sq, float max_norm, float* __restrict__ mult) {
  const float norm = sqrtf((float)sum_sq[0]);
  const float r = max_norm / (1e-7f + norm);
  mult[0] = (r != r) ? r : fminf(1.0f, r);  // jnp.minimum keeps a NaN (a NaN gradient makes the whole step NaN)
}

__global__ void __launch_bounds__(256) bank_put_kernel(float* __restrict__ dst, const float* __restrict__ src,
                                                       int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n4 = n >> 2;
  const bool al = ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15u) == 0;
  if (al) {
    for (int64_t i = i0; i < n4; i += stride)
      reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(src)[i];
    for (int64_t i = n4 * 4 + i0; i < n; i += stride) dst[i] = src[i];
  } else {
    for (int64_t i = i0; i < n; i += stride) dst[i] = src[i];
  }
}

static StepCoef make_coef(const fb200_sgmcmc_hparams_t* hp) {
  StepCoef c;
  c.step_size = hp->step_size;
  c.sqrt_eps = sqrtf(hp->step_size);
  c.beta = (float)hp->momentum_decay;
  c.noise_scale = sqrtf((float)(2.0 * (1.0 - hp->momentum_decay)));
  c.alpha = (float)hp->rms_alpha;
  c.one_m_alpha = (float)(1.0 - hp->rms_alpha);
  c.rms_eps = (float)hp->rms_eps;
  c.zero_momentum = hp->zero_momentum;
  c.grad_mult = hp->grad_mult;
  c.grad_bf16 = hp->grad_dtype == FB200_BF16;
  return c;
}

}  // namespace fb200

using namespace fb200;

extern "C" int64_t fb200_sgmcmc_num_tiles_host(const fb200_leaf_t* leaves, int n_leaves) {
  if (!leaves || n_leaves <= 0) return FB200_E_NULL;
  int64_t t = 0;
  for (int l = 0; l < n_leaves; ++l) {
    if (leaves[l].len <= 0 || leaves[l].len >= 0xFFFFFFFFll || leaves[l].offset < 0) return FB200_E_SHAPE;
    const int64_t h = (leaves[l].len + 1) / 2;
    t += (h + FB200_TILE_PAIRS - 1) / FB200_TILE_PAIRS;
  }
  return t;
}

extern "C" int fb200_sgmcmc_build_tiles_host(const fb200_leaf_t* leaves, int n_leaves, fb200_tile_t* tiles,
                                             int64_t n_tiles) {
  if (!leaves || !tiles) return FB200_E_NULL;
  const int64_t need = fb200_sgmcmc_num_tiles_host(leaves, n_leaves);
  if (need < 0) return (int)need;
  if (need != n_tiles) return FB200_E_SHAPE;
  int64_t t = 0;
  for (int l = 0; l < n_leaves; ++l) {
    const int64_t h = (leaves[l].len + 1) / 2;
    for (int64_t p = 0; p < h; p += FB200_TILE_PAIRS) {
      tiles[t].leaf = l;
      tiles[t].pair_start = (uint32_t)p;
      ++t;
    }
  }
  return 0;
}

static int sgmcmc_step_impl(float* params, const float* grad, float* momentum, float* precond_v, float* updates_out,
                            int64_t n, const fb200_leaf_t* leaves, int n_leaves, const fb200_tile_t* tiles,
                            int64_t n_tiles, const uint32_t* key_in, uint32_t* key_out, uint32_t* leaf_keys_ws,
                            const fb200_sgmcmc_hparams_t* hp, const fb200_joint_grad_t* joint, double* sq_ws,
                            double* sq_out, void* stream, const DpPeers* dpp = nullptr) {
  if (!grad || !momentum || !leaves || !tiles || !key_in || !key_out || !leaf_keys_ws || !hp) return FB200_E_NULL;
  if (!params && !updates_out) return FB200_E_NULL;
  if (joint && !params) return FB200_E_NULL;
  if (sq_out && !sq_ws) return FB200_E_NULL;
  if (hp->rmsprop && !precond_v) return FB200_E_NULL;
  if (n <= 0 || n_leaves <= 0 || n_tiles <= 0 || n_tiles > 0x7FFFFFFFll || n_leaves >= (1 << 30))
    return FB200_E_SHAPE;
  if (key_in == key_out) return FB200_E_SHAPE;
  if (hp->grad_dtype != FB200_F32 && hp->grad_dtype != FB200_BF16) return FB200_E_UNSUPPORTED;
  if (hp->grad_dtype == FB200_BF16 && dpp) return FB200_E_UNSUPPORTED;  // peer gradients are float32
  if (!fb200_aligned16(grad) || !fb200_aligned16(momentum) || (params && !fb200_aligned16(params)) ||
      (updates_out && !fb200_aligned16(updates_out)) || (hp->rmsprop && !fb200_aligned16(precond_v)))
    return FB200_E_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  const StepCoef c = make_coef(hp);
  JointCoef jc{1.0f, 0.0f, nullptr};
  if (joint) {
    jc.lik_scale = joint->lik_scale;
    jc.prior_coef = -joint->prior_prec;
    jc.sq_part = sq_out ? sq_ws : nullptr;
  }
  const unsigned g = (unsigned)n_tiles;
  const DpPeers nodp{};
  const DpPeers& dpa = dpp ? *dpp : nodp;
  // key schedule: folded into the step kernel for launch-bound (small) trees, a kernel of its own otherwise
  const bool fold_keys = n_tiles <= kFoldKeysMaxTiles;
  if (!fold_keys) {
    sgmcmc_keys_kernel<<<(n_leaves + 127) / 128, 128, 0, st>>>(key_in, key_out, leaf_keys_ws, (uint32_t)n_leaves);
    FB200_LAUNCH_CHECK();
    key_in = nullptr;
  }
  float* pv = hp->rmsprop ? precond_v : nullptr;
  float* up = dpp ? nullptr : updates_out;
#define FB200_STEP_LAUNCH(RMS_, JOINT_, NP_)                                                                          \
  sgmcmc_step_kernel<RMS_, JOINT_, NP_><<<g, 256, 0, st>>>(params, grad, momentum, pv, up, leaves, key_in, key_out,    \
                                                           leaf_keys_ws, (uint32_t)n_leaves, tiles, c, jc, dpa)
#define FB200_STEP_MODES(NP_)                         \
  do {                                                \
    if (joint) {                                      \
      if (hp->rmsprop) FB200_STEP_LAUNCH(true, true, NP_);   \
      else FB200_STEP_LAUNCH(false, true, NP_);       \
    } else {                                          \
      if (hp->rmsprop) FB200_STEP_LAUNCH(true, false, NP_);  \
      else FB200_STEP_LAUNCH(false, false, NP_);      \
    }                                                 \
  } while (0)
  if (!dpp) FB200_STEP_MODES(0);
  else if (dpp->mc_grad) FB200_STEP_MODES(1);
  else if (dpp->n <= 2) FB200_STEP_MODES(2);
  else if (dpp->n <= 4) FB200_STEP_MODES(4);
  else FB200_STEP_MODES(8);
#undef FB200_STEP_MODES
#undef FB200_STEP_LAUNCH
  FB200_LAUNCH_CHECK();
  if (joint && sq_out) {
    sum_partials_kernel<<<1, 1024, 0, st>>>(sq_ws, n_tiles, sq_out);
    FB200_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int fb200_sgmcmc_step_f32(float* params, const float* grad, float* momentum, float* precond_v,
                                     float* updates_out, int64_t n, const fb200_leaf_t* leaves, int n_leaves,
                                     const fb200_tile_t* tiles, int64_t n_tiles, const uint32_t* key_in,
                                     uint32_t* key_out, uint32_t* leaf_keys_ws, const fb200_sgmcmc_hparams_t* hp,
                                     void* stream) {
  return sgmcmc_step_impl(params, grad, momentum, precond_v, updates_out, n, leaves, n_leaves, tiles, n_tiles, key_in,
                          key_out, leaf_keys_ws, hp, nullptr, nullptr, nullptr, stream);
}

extern "C" int fb200_sgmcmc_step_joint_f32(float* params, const float* grad_loglik, float* momentum, float* precond_v,
                                           int64_t n, const fb200_leaf_t* leaves, int n_leaves,
                                           const fb200_tile_t* tiles, int64_t n_tiles, const uint32_t* key_in,
                                           uint32_t* key_out, uint32_t* leaf_keys_ws,
                                           const fb200_sgmcmc_hparams_t* hp, const fb200_joint_grad_t* joint,
                                           double* sq_workspace, double* theta_sq_out, void* stream) {
  if (!joint) return FB200_E_NULL;
  return sgmcmc_step_impl(params, grad_loglik, momentum, precond_v, nullptr, n, leaves, n_leaves, tiles, n_tiles,
                          key_in, key_out, leaf_keys_ws, hp, joint, sq_workspace, theta_sq_out, stream);
}

static int dp_peers_from_abi(const fb200_dp_peers_t* peers, DpPeers* dp) {
  if (!peers) return FB200_E_NULL;
  if (peers->n_ranks < 1 || peers->n_ranks > FB200_MAX_PEERS || peers->rank < 0 || peers->rank >= peers->n_ranks)
    return FB200_E_SHAPE;
  *dp = DpPeers{};
  dp->n = peers->n_ranks;
  dp->inv_n = 1.0f / (float)peers->n_ranks;
  for (int p = 0; p < peers->n_ranks; ++p) {
    if (!peers->grad[p] || !peers->params[p]) return FB200_E_NULL;
    if (!fb200_aligned16(peers->grad[p]) || !fb200_aligned16(peers->params[p])) return FB200_E_ALIGN;
    dp->grad[p] = peers->grad[p];
    dp->params[p] = peers->params[p];
  }
  if ((peers->mc_grad == nullptr) != (peers->mc_params == nullptr)) return FB200_E_NULL;
  if (peers->mc_grad && (!fb200_aligned16(peers->mc_grad) || !fb200_aligned16(peers->mc_params))) return FB200_E_ALIGN;
  dp->mc_grad = peers->mc_grad;
  dp->mc_params = peers->mc_params;
  return 0;
}

extern "C" int fb200_sgmcmc_step_dp_f32(float* momentum, float* precond_v, int64_t n, const fb200_leaf_t* leaves,
                                        int n_leaves, const fb200_tile_t* tiles, int64_t tile_begin, int64_t tile_end,
                                        const uint32_t* key_in, uint32_t* key_out, uint32_t* leaf_keys_ws,
                                        const fb200_sgmcmc_hparams_t* hp, const fb200_joint_grad_t* joint,
                                        const fb200_dp_peers_t* peers, void* stream) {
  if (!peers || !tiles) return FB200_E_NULL;
  if (tile_begin < 0 || tile_end < tile_begin) return FB200_E_SHAPE;
  DpPeers dp;
  const int rc = dp_peers_from_abi(peers, &dp);
  if (rc) return rc;
  if (tile_end == tile_begin) {  // a rank without tiles still advances its key like everyone else
    if (!key_in || !key_out || !leaf_keys_ws || n_leaves <= 0) return FB200_E_NULL;
    sgmcmc_keys_kernel<<<(n_leaves + 127) / 128, 128, 0, (cudaStream_t)stream>>>(key_in, key_out, leaf_keys_ws,
                                                                                 (uint32_t)n_leaves);
    FB200_LAUNCH_CHECK();
    return 0;
  }
  return sgmcmc_step_impl(peers->params[peers->rank], peers->grad[peers->rank], momentum, precond_v, nullptr, n, leaves,
                          n_leaves, tiles + tile_begin, tile_end - tile_begin, key_in, key_out, leaf_keys_ws, hp, joint,
                          nullptr, nullptr, stream, &dp);
}

extern "C" int64_t fb200_grad_clip_workspace_doubles(int64_t n) {
  if (n <= 0) return 0;
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 16) blocks = 148 * 16;
  return blocks + 1;  // per-CTA partials + their sum
}

extern "C" int fb200_grad_clip_mult(const void* grad, int grad_dtype, const float* params, const fb200_joint_grad_t* joint,
                                    int64_t n, float max_grad_norm, double* workspace, float* mult_out, void* stream) {
  if (!grad || !workspace || !mult_out) return FB200_E_NULL;
  if (joint && !params) return FB200_E_NULL;
  if (n <= 0 || !(max_grad_norm > 0.0f)) return FB200_E_SHAPE;
  if (grad_dtype != FB200_F32 && grad_dtype != FB200_BF16) return FB200_E_UNSUPPORTED;
  if ((grad_dtype == FB200_F32 && !fb200_aligned16(grad)) || (reinterpret_cast<uintptr_t>(grad) & 7u) ||
      (joint && !fb200_aligned16(params)))
    return FB200_E_ALIGN;
  const int64_t blocks = fb200_grad_clip_workspace_doubles(n) - 1;
  JointCoef jc{1.0f, 0.0f, nullptr};
  if (joint) {
    jc.lik_scale = joint->lik_scale;
    jc.prior_coef = -joint->prior_prec;
  }
  cudaStream_t st = (cudaStream_t)stream;
  grad_sq_partials_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const float*>(grad), grad_dtype == FB200_BF16,
                                                            params, jc, joint != nullptr, n, workspace);
  FB200_LAUNCH_CHECK();
  sum_partials_kernel<<<1, 1024, 0, st>>>(workspace, blocks, workspace + blocks);
  FB200_LAUNCH_CHECK();
  grad_clip_finish_kernel<<<1, 1, 0, st>>>(workspace + blocks, max_grad_norm, mult_out);
  FB200_LAUNCH_CHECK();
  return 0;
}

static int explore_blocks(int64_t n) {
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 32) blocks = 148 * 32;
  return (int)blocks;
}

extern "C" int64_t fb200_sgd_explore_joint_workspace_doubles(int64_t n) { return n > 0 ? explore_blocks(n) : 0; }

template <int NP>
static void explore_launch(float* params, const float* grad, float* precond_v, float* updates_out, int64_t begin,
                           int64_t end, bool rms, bool joint, unsigned blocks, const StepCoef& c, const JointCoef& jc,
                           const DpPeers& dp, cudaStream_t st) {
  if (joint) {
    if (rms) sgd_explore_kernel<true, true, NP><<<blocks, 256, 0, st>>>(params, grad, precond_v, updates_out, begin, end, c, jc, dp);
    else sgd_explore_kernel<false, true, NP><<<blocks, 256, 0, st>>>(params, grad, nullptr, updates_out, begin, end, c, jc, dp);
  } else {
    if (rms) sgd_explore_kernel<true, false, NP><<<blocks, 256, 0, st>>>(params, grad, precond_v, updates_out, begin, end, c, jc, dp);
    else sgd_explore_kernel<false, false, NP><<<blocks, 256, 0, st>>>(params, grad, nullptr, updates_out, begin, end, c, jc, dp);
  }
}

static int sgd_explore_impl(float* params, const float* grad, float* precond_v, float* updates_out, int64_t n,
                            const fb200_sgmcmc_hparams_t* hp, const fb200_joint_grad_t* joint, double* sq_ws,
                            double* sq_out, void* stream, const DpPeers* dpp = nullptr, int64_t begin = 0,
                            int64_t end = -1) {
  if (!grad || !hp) return FB200_E_NULL;
  if (!params && !updates_out) return FB200_E_NULL;
  if (joint && !params) return FB200_E_NULL;
  if (sq_out && !sq_ws) return FB200_E_NULL;
  if (hp->rmsprop && !precond_v) return FB200_E_NULL;
  if (n <= 0) return FB200_E_SHAPE;
  if (end < 0) end = n;
  if (begin < 0 || end > n || begin > end) return FB200_E_SHAPE;
  if (begin == end) return 0;
  if (begin & 3) return FB200_E_SHAPE;
  if (hp->grad_dtype != FB200_F32) return FB200_E_UNSUPPORTED;  // the exploration step takes float32 gradients
  if (!fb200_aligned16(grad) || (params && !fb200_aligned16(params)) ||
      (updates_out && !fb200_aligned16(updates_out)) || (hp->rmsprop && !fb200_aligned16(precond_v)))
    return FB200_E_ALIGN;
  const StepCoef c = make_coef(hp);
  const unsigned blocks = (unsigned)explore_blocks(end - begin);
  cudaStream_t st = (cudaStream_t)stream;
  JointCoef jc{1.0f, 0.0f, nullptr};
  if (joint) {
    jc.lik_scale = joint->lik_scale;
    jc.prior_coef = -joint->prior_prec;
    jc.sq_part = sq_out ? sq_ws : nullptr;
  }
  const DpPeers nodp{};
  const bool rms = hp->rmsprop != 0;
  if (!dpp) explore_launch<0>(params, grad, precond_v, updates_out, begin, end, rms, joint != nullptr, blocks, c, jc, nodp, st);
  else if (dpp->mc_grad) explore_launch<1>(params, grad, precond_v, nullptr, begin, end, rms, joint != nullptr, blocks, c, jc, *dpp, st);
  else if (dpp->n <= 2) explore_launch<2>(params, grad, precond_v, nullptr, begin, end, rms, joint != nullptr, blocks, c, jc, *dpp, st);
  else if (dpp->n <= 4) explore_launch<4>(params, grad, precond_v, nullptr, begin, end, rms, joint != nullptr, blocks, c, jc, *dpp, st);
  else explore_launch<8>(params, grad, precond_v, nullptr, begin, end, rms, joint != nullptr, blocks, c, jc, *dpp, st);
  FB200_LAUNCH_CHECK();
  if (joint && sq_out) {
    sum_partials_kernel<<<1, 1024, 0, st>>>(sq_ws, blocks, sq_out);
    FB200_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int fb200_sgd_explore_step_dp_f32(float* precond_v, int64_t n, int64_t elem_begin, int64_t elem_end,
                                             const fb200_sgmcmc_hparams_t* hp, const fb200_joint_grad_t* joint,
                                             const fb200_dp_peers_t* peers, void* stream) {
  DpPeers dp;
  const int rc = dp_peers_from_abi(peers, &dp);
  if (rc) return rc;
  return sgd_explore_impl(peers->params[peers->rank], peers->grad[peers->rank], precond_v, nullptr, n, hp, joint,
                          nullptr, nullptr, stream, &dp, elem_begin, elem_end);
}

extern "C" int fb200_sgd_explore_step_f32(float* params, const float* grad, float* precond_v, float* updates_out,
                                          int64_t n, const fb200_sgmcmc_hparams_t* hp, void* stream) {
  return sgd_explore_impl(params, grad, precond_v, updates_out, n, hp, nullptr, nullptr, nullptr, stream);
}

extern "C" int fb200_sgd_explore_step_joint_f32(float* params, const float* grad_loglik, float* precond_v, int64_t n,
                                                const fb200_sgmcmc_hparams_t* hp, const fb200_joint_grad_t* joint,
                                                double* sq_workspace, double* theta_sq_out, void* stream) {
  if (!joint) return FB200_E_NULL;
  return sgd_explore_impl(params, grad_loglik, precond_v, nullptr, n, hp, joint, sq_workspace, theta_sq_out, stream);
}

extern "C" int fb200_bank_put_f32(float* bank, int64_t n, int row, const float* src, void* stream) {
  if (!bank || !src) return FB200_E_NULL;
  if (n <= 0 || row < 0) return FB200_E_SHAPE;
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 16) blocks = 148 * 16;
  bank_put_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(bank + (int64_t)row * n, src, n);
  FB200_LAUNCH_CHECK();
  return 0;
}
