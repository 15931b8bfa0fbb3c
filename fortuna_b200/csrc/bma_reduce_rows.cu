// K3 + K4 (+ K5 + K6 fused): the warp-autonomous single-pass BMA reduction for wide rows, see bma_reduce.cu for the
// reference lines it replaces; the fused epilogue additionally replaces
//   fortuna/conformal/classification/adaptive_prediction.py:11-22 (APS scores of every class of the mean row)
//   fortuna/conformal/classification/base.py:45-64, 72-82       (rank-count conformal sets)
#include <stdlib.h>

#include "bma_reduce.cuh"
#include "row_sort.cuh"

namespace fb200 {

// ------------------------------------------------------------------------------------------------
// Warp-autonomous variant for wide rows (C > 128, C % 4 == 0: one warp per row, G = 32).  Every warp is its own
// stream processor: it owns whole input rows b (b = warp id, + number of warps, ...), walks their S samples and
// fetches them itself - lane 0 issues one cp.async.bulk per (sample, row) into the WARP's private shared-memory ring
// (mbarrier per slot) and re-arms a slot as soon as the warp has read it.  There is no producer warp, no CTA-wide
// barrier and no tile: a warp that is busy in its row epilogue stalls nothing but its own ring (a few KB), the other
// warps keep the HBM queue full.  That is what makes a LONG epilogue affordable - the fused conformal sets below sort
// the mean row in registers (~3 K instructions) - where the tiled kernel above loses every consumer warp's epilogue
// time on the shared ring.  Arithmetic per logit is process_rows, unchanged.
//
// Fused APS sets (SETS): the mean row is already in the warp's registers, 32 values per lane - what the key-only
// sort kernel (scoring.cu) loads from HBM.  Same network, same tree-order prefix sums, same threshold test; rows with
// equal probabilities or a non-monotone boundary (the general rank-lookup path) are appended to a fallback list and
// handled by that kernel afterwards, so the result is the same bits by construction.  The 262 MB mean re-read, the
// separate 0.47 ms launch and its HBM-latency exposure disappear.
// ------------------------------------------------------------------------------------------------
constexpr int kRowMaxStages = 8;
__host__ __device__ constexpr int next_pow2_c(int v) { return v <= 1 ? 1 : 2 * next_pow2_c((v + 1) / 2); }

template <typename T, int NV, int NW, int RI, bool SETS>
__global__ void __launch_bounds__(NW * 32, 1) bma_reduce_cls_rows_kernel(const ClsArgs a) {
  constexpr int G = 32, VEC = 4, NP = 2;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int g = lane;
  const int S = a.S, B = a.B, C = a.C;
  const int nst = a.n_stages;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw) + warp * kRowMaxStages;
  unsigned char* ring = smem_raw + ((NW * kRowMaxStages * 8 + 127) & ~127) + (size_t)warp * nst * a.row_stride_bytes;
  if (lane == 0) {
    for (int i = 0; i < nst; ++i) mbar_init(&bar[i], 1);
    fence_mbar_init();
  }
  __syncwarp();

  const int total_warps = gridDim.x * NW;
  const int gw = blockIdx.x * NW + warp;
  const int n_my = gw < B ? (B - gw + total_warps - 1) / total_warps : 0;
  auto row_of = [&](int i) {
    int b = gw + i * total_warps + a.row_rot;
    return b >= B ? b - B : b;
  };
  const T* base = reinterpret_cast<const T*>(a.logits);
  // prefetch cursor (warp-uniform; only lane 0 issues)
  int pf_i = 0, pf_s = 0, pf_stage = 0;
  auto issue_next = [&]() {
    if (pf_i < n_my) {
      if (lane == 0) {
        const T* src = base + ((size_t)pf_s * B + row_of(pf_i)) * C;
        mbar_arrive_expect_tx(&bar[pf_stage], a.row_bytes);
        bulk_g2s(ring + (size_t)pf_stage * a.row_stride_bytes, src, a.row_bytes, &bar[pf_stage]);
      }
      if (++pf_stage == nst) pf_stage = 0;
      if (++pf_s == S) {
        pf_s = 0;
        ++pf_i;
      }
    }
  };
  for (int k = 0; k < nst; ++k) issue_next();
  if (a.stagger_ns > 0) {
    for (int w = 0; w < warp; ++w) __nanosleep((unsigned)a.stagger_ns);
  }

  const float scale = a.log_temp ? expf(-a.log_temp[0]) : 1.0f;
  const float k2 = scale * 1.4426950408889634f;
  int stage = 0;
  uint32_t parity = 0;

  for (int i = 0; i < n_my; ++i) {
    const int b = row_of(i);
    float2 sp[NV][NP], sp2[NV][NP];
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
      for (int k = 0; k < NP; ++k) sp[j][k] = sp2[j][k] = make_float2(0.0f, 0.0f);
    float spl = 0.0f, sum_l = 0.0f;
    float lmax = -CUDART_INF_F, lsum = 0.0f;
    int y = -1;
    if (a.targets) y = a.targets[b];
    const bool y_ok = (y >= 0) && (y < C);
    auto log_prob_update = [&](const T* rp, float csh, float l2s, int s_idx) {
      float zt = 0.0f;
      if (y_ok) {
        float tmp[1];
        RowLoad<T, 1>::load(rp + y, tmp);
        zt = tmp[0];
      }
      // ll = z_y - logsumexp = ln2 * (u_y - log2(sum))
      const float ll = 0.6931471805599453f * (fmaf(zt, k2, -csh) - l2s);
      if (ll > lmax) {
        lsum = lsum * __expf(lmax - ll) + 1.0f;
        lmax = ll;
      } else if (ll > -CUDART_INF_F) {
        lsum += __expf(ll - lmax);
      } else if (ll != ll) {
        lsum = ll;  // NaN propagates like the reference's logsumexp
      }
      if (a.out.ensemble_log_prob && lane == 0) a.out.ensemble_log_prob[(size_t)s_idx * B + b] = ll;
    };
    int s = 0;
    if (RI == 2) {
      for (; s + 1 < S; s += 2) {
        const int st1 = (stage + 1 == nst) ? 0 : stage + 1;
        const uint32_t par1 = (stage + 1 == nst) ? parity ^ 1u : parity;
        mbar_wait(&bar[stage], parity);
        mbar_wait(&bar[st1], par1);
        const T* rp[2];
        rp[0] = reinterpret_cast<const T*>(ring + (size_t)stage * a.row_stride_bytes);
        rp[1] = reinterpret_cast<const T*>(ring + (size_t)st1 * a.row_stride_bytes);
        float csh[2], l2s[2];
        process_rows<T, G, NV, VEC, true, 2>(rp, g, C, k2, sp, sp2, spl, sum_l, csh, l2s);
        if (a.targets) {
          log_prob_update(rp[0], csh[0], l2s[0], s);
          log_prob_update(rp[1], csh[1], l2s[1], s + 1);
        }
        __syncwarp();  // every lane has read both slots: re-arm them
        issue_next();
        issue_next();
        stage = (st1 + 1 == nst) ? 0 : st1 + 1;
        parity = (st1 + 1 == nst) ? par1 ^ 1u : par1;
      }
    }
    for (; s < S; ++s) {
      mbar_wait(&bar[stage], parity);
      const T* rp[1];
      rp[0] = reinterpret_cast<const T*>(ring + (size_t)stage * a.row_stride_bytes);
      float csh[1], l2s[1];
      process_rows<T, G, NV, VEC, true, 1>(rp, g, C, k2, sp, sp2, spl, sum_l, csh, l2s);
      if (a.targets) log_prob_update(rp[0], csh[0], l2s[0], s);
      __syncwarp();
      issue_next();
      if (++stage == nst) {
        stage = 0;
        parity ^= 1u;
      }
    }

    // ------------------------------------------------------------------ row epilogue
    const float Sf = (float)S;
    const DivS dS = make_divs(Sf);
    const float spl_row = 0.6931471805599453f * (group_sum<G>(spl) - sum_l);  // sum_s sum_c p ln p of this row
    if (a.partial_mode) {
      const int owner = b / a.rows_per_rank;
      const int lr = b - owner * a.rows_per_rank;
      float* dst = reinterpret_cast<float*>(a.slot_ptrs[owner]) + (size_t)lr * (size_t)(2 * C + 4);
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int c0 = (g + G * j) * VEC;
        if (c0 < C) {
          *reinterpret_cast<float4*>(dst + c0) = make_float4(sp[j][0].x, sp[j][0].y, sp[j][1].x, sp[j][1].y);
          *reinterpret_cast<float4*>(dst + C + c0) = make_float4(sp2[j][0].x, sp2[j][0].y, sp2[j][1].x, sp2[j][1].y);
        }
      }
      if (g == 0) *reinterpret_cast<float4*>(dst + 2 * C) = make_float4(spl_row, lmax, lsum, 0.0f);
      continue;
    }

    float ent_terms = 0.0f;
    float best_v = -CUDART_INF_F;
    int best_c = 0x7FFFFFFF;
    float msq = 0.0f, mean_y = 0.0f;  // Brier term of the row: sum_c mean^2 - 2 mean_y + 1
    const bool want_brier = a.out.brier_terms != nullptr;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c0 = (g + G * j) * VEC;
      if (c0 < C) {
        float mean[VEC], av[VEC], ev[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          const float spv = pk_get(sp, j, k), sp2v = pk_get(sp2, j, k);
          mean[k] = div_s(spv, dS);
          if (want_brier) {
            msq = fmaf(mean[k], mean[k], msq);
            mean_y = (c0 + k == y) ? mean[k] : mean_y;
          }
          const float spq = spv - sp2v;  // sum_s p(1-p)
          av[k] = div_s(spq, dS);
          const float m2 = div_s(sp2v, dS);
          ev[k] = fmaxf(0.0f, __fsub_rn(m2, __fmul_rn(mean[k], mean[k])));
          if (mean[k] > 0.0f) ent_terms = fmaf(mean[k], __logf(mean[k]), ent_terms);
          if (mean[k] > best_v) {
            best_v = mean[k];
            best_c = c0 + k;
          }
        }
        const size_t o = (size_t)b * C + c0;
        if (a.out.mean) st_f4(a.out.mean + o, mean[0], mean[1], mean[2], mean[3]);
        if (a.out.aleatoric_variance) st_f4(a.out.aleatoric_variance + o, av[0], av[1], av[2], av[3]);
        if (a.out.epistemic_variance) st_f4(a.out.epistemic_variance + o, ev[0], ev[1], ev[2], ev[3]);
        if (a.out.variance) st_f4(a.out.variance + o, av[0] + ev[0], av[1] + ev[1], av[2] + ev[2], av[3] + ev[3]);
        if (a.out.std)
          st_f4(a.out.std + o, sqrtf(av[0] + ev[0]), sqrtf(av[1] + ev[1]), sqrtf(av[2] + ev[2]), sqrtf(av[3] + ev[3]));
        if constexpr (SETS) {  // the accumulator registers now hold the mean row
          sp[j][0] = make_float2(mean[0], mean[1]);
          sp[j][1] = make_float2(mean[2], mean[3]);
        }
      }
    }
    const float ent = -group_sum<G>(ent_terms);
    const float alea = -spl_row / Sf;
    // argmax with first-maximum tie-break across the lanes of the row
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best_v, o);
      const int oc = __shfl_xor_sync(0xffffffffu, best_c, o);
      if (ov > best_v || (ov == best_v && oc < best_c)) {
        best_v = ov;
        best_c = oc;
      }
    }
    if (want_brier) {
      msq = group_sum<G>(msq);
      mean_y = group_sum<G>(mean_y);  // one lane of the row held the target class
    }
    if (g == 0) {
      if (a.out.entropy) a.out.entropy[b] = ent;
      if (a.out.aleatoric_entropy) a.out.aleatoric_entropy[b] = alea;
      if (a.out.epistemic_entropy) a.out.epistemic_entropy[b] = ent - alea;
      if (a.out.mode) a.out.mode[b] = best_c == 0x7FFFFFFF ? 0 : best_c;
      if (a.out.log_prob) a.out.log_prob[b] = (lmax + logf(lsum)) - a.log_S;
      if (a.out.confidence) a.out.confidence[b] = best_v;
      if (want_brier) a.out.brier_terms[b] = y_ok ? (msq - 2.0f * mean_y) + 1.0f : msq;
    }

    if constexpr (SETS) {
      // -------------------------------------------------------------- fused APS conformal set of the mean row
      uint8_t* mrow = a.cs_mask + (size_t)b * C;
      auto store_mask = [&](bool any, float pb) {  // class c is in  <=>  any && mean_c <= pb   (4 classes per store)
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          const int c0 = (g + G * j) * VEC;
          if (c0 < C) {
            uint32_t w = 0u;
#pragma unroll
            for (int k = 0; k < VEC; ++k)
              w |= (any && (pk_get(sp, j, k) + 0.0f) <= pb) ? (1u << (8 * k)) : 0u;
            *reinterpret_cast<uint32_t*>(mrow + c0) = w;
          }
        }
      };
      if (a.cs_all_true || a.cs_all_false) {
        store_mask(a.cs_all_true != 0, CUDART_INF_F);
        if (lane == 0 && a.cs_sizes) a.cs_sizes[b] = a.cs_all_true ? C : 0;
      } else {
        constexpr int NR = next_pow2_c(NV * VEC);  // keys per lane of the network (power of two)
        uint32_t key[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          const int j = r / VEC, k = r % VEC;
          key[r] = 0xFFFFFFFFu;  // padding sorts last
          if (j < NV) {
            const int c = (g + G * j) * VEC + k;
            if (c < C) key[r] = ~f32_ordered(pk_get(sp, j < NV ? j : 0, k) + 0.0f);  // -0 -> +0
          }
        }
        warp_bitonic_sort32<NR>(key, lane);
        float x[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) x[r] = (lane * NR + r < C) ? f32_unordered(~key[r]) : 0.0f;
        warp_row_tree_prefix<NR>(x, lane);
        // equal neighbours among the real (rank < C) keys?
        bool tie = false;
#pragma unroll
        for (int r = 0; r + 1 < NR; ++r) tie |= (key[r] == key[r + 1]) && (lane * NR + r + 1 < C);
        {
          const uint32_t nxt = __shfl_down_sync(0xffffffffu, key[0], 1);
          tie |= (lane < 31) && (key[NR - 1] == nxt) && ((lane + 1) * NR < C);
        }
        bool ok = !__any_sync(0xffffffffu, tie);
        const float q = __ldg(a.cs_val_sorted + a.cs_q_index);
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
        ok = ok && (none ? (n_true == 0) : (n_true == C - first));  // the set is exactly the ranks >= first
        if (ok) {
          float pb = 0.0f;  // probability at the boundary rank
#pragma unroll
          for (int r = 0; r < NR; ++r) pb = (lane * NR + r == first) ? f32_unordered(~key[r]) : pb;
          pb = __shfl_sync(0xffffffffu, pb, none ? 0 : first / NR);
          store_mask(!none, pb);
          if (lane == 0 && a.cs_sizes) a.cs_sizes[b] = n_true;
        } else if (lane == 0) {
          a.fb_rows[atomicAdd(a.fb_count, 1)] = b;  // general path: scoring.cu's kernel on the stored mean row
        }
      }
    }
  }
}

// NW consumer warps x RI interleaved samples; registers: 65536 / (32 NW) per thread.
template <typename T, int NV, int NW, int RI, bool SETS>
static int launch_cls_rows_cfg(ClsArgs a, cudaStream_t st) {
  a.row_bytes = (uint32_t)a.C * (uint32_t)sizeof(T);
  a.row_stride_bytes = (a.row_bytes + 15u) & ~15u;
  const size_t bar_bytes = (NW * kRowMaxStages * 8 + 127) & ~127;
  const size_t budget = 224u * 1024u - bar_bytes;
  int nst = (int)(budget / NW / a.row_stride_bytes);
  if (nst > kRowMaxStages) nst = kRowMaxStages;
  if (nst < 2) return FB200_E_UNSUPPORTED;
  a.n_stages = nst;
  {
    const char* e = getenv("FB200_K3_STAGGER_NS");  // experiment knob
    a.stagger_ns = e ? atoi(e) : 0;
  }
  a.row_rot = a.partial_mode ? (int)(((long long)a.B * a.tile_rot) / (a.n_ranks > 0 ? a.n_ranks : 1)) : 0;
  const size_t smem = bar_bytes + (size_t)NW * nst * a.row_stride_bytes;
  auto kern = bma_reduce_cls_rows_kernel<T, NV, NW, RI, SETS>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  int grid = sm_count();
  const int need = (a.B + NW - 1) / NW;
  if (grid > need) grid = need;
  kern<<<grid, NW * 32, smem, st>>>(a);
  FB200_LAUNCH_CHECK();
  return 0;
}


// Warp / interleave configuration: rows of more than 512 classes keep 64 accumulator + 64 row registers per lane
// (two interleaved samples) -> 12 warps x 168 registers; narrower rows 16 warps x 128.
template <typename T, int NV, bool SETS>
static int launch_cls_rows_nv(const ClsArgs& a, cudaStream_t st) {
  if constexpr (NV >= 5) return launch_cls_rows_cfg<T, NV, 12, 2, SETS>(a, st);
  else return launch_cls_rows_cfg<T, NV, 16, 2, SETS>(a, st);
}

template <typename T, bool SETS>
static int dispatch_cls_rows(const ClsArgs& a, cudaStream_t st) {
  const int chunks = a.C / 4;
  if (chunks <= 64) return launch_cls_rows_nv<T, 2, SETS>(a, st);
  if (chunks <= 96) return launch_cls_rows_nv<T, 3, SETS>(a, st);
  if (chunks <= 128) return launch_cls_rows_nv<T, 4, SETS>(a, st);
  if (chunks <= 160) return launch_cls_rows_nv<T, 5, SETS>(a, st);
  if (chunks <= 192) return launch_cls_rows_nv<T, 6, SETS>(a, st);
  if (chunks <= 224) return launch_cls_rows_nv<T, 7, SETS>(a, st);
  return launch_cls_rows_nv<T, 8, SETS>(a, st);
}

// one warp per row, 16-byte chunks per lane, every row a legal bulk-copy source
bool cls_rows_supported(const ClsArgs& a, int dtype) {
  const size_t esz = dtype == FB200_F32 ? 4 : 2;
  if (a.C <= 128 || a.C > 1024 || (a.C % 4) != 0) return false;
  if (((size_t)a.C * esz) % 16 != 0 || !fb200_aligned16(a.logits)) return false;
  if (a.cs_mask && (reinterpret_cast<uintptr_t>(a.cs_mask) & 3u)) return false;  // 4 mask bytes per store
  return true;
}

int launch_cls_rows(const ClsArgs& a, int dtype, cudaStream_t st) {
  if (!cls_rows_supported(a, dtype)) return FB200_E_UNSUPPORTED;
  if (a.cs_mask) {
    if (dtype == FB200_F32) return dispatch_cls_rows<float, true>(a, st);
    return dispatch_cls_rows<__nv_bfloat16, true>(a, st);
  }
  if (dtype == FB200_F32) return dispatch_cls_rows<float, false>(a, st);
  return dispatch_cls_rows<__nv_bfloat16, false>(a, st);
}

}  // namespace fb200
