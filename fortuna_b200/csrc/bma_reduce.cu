// K3 + K4: single-pass Bayesian-model-averaging reductions over calibrated logits z[S,B,C].
//
// Replaces (reference file:line)
//   fortuna/prob_model/predictive/base.py:624-647   _batched_mean            sum_s softmax / S
//   fortuna/prob_model/predictive/base.py:735-758   _batched_aleatoric_variance   sum_s p(1-p) / S
//   fortuna/prob_model/predictive/base.py:806-836   _batched_epistemic_variance   max(0, E p^2 - (E p)^2)
//   fortuna/prob_model/predictive/base.py:179-203, 104-128   log_prob / ensemble_log_prob
//   fortuna/prob_model/predictive/classification.py:71-86, 257-432   mode, (aleatoric|epistemic) entropy
//   fortuna/prob_output_layer/classification.py:25-28, 54-69, 88-104  per-sample maps
//   fortuna/output_calibrator/classification.py:14-17                  z = logits * exp(-log_temp)
//
// Data movement: logits are read from HBM exactly once.  A CTA owns a tile of TB consecutive inputs b and
// walks the S samples; for each s the TB rows z[s, b0:b0+TB, :] are one contiguous block, fetched by a
// producer warp with cp.async.bulk (TMA engine) into an mbarrier-signalled shared-memory ring.  Eight
// consumer warps each own whole rows (G lanes per row), so the per-row max / sum-exp are warp-shuffle
// reductions and the per-(b,c) sums over s live in registers until the tile's single store.
#include "bma_reduce.cuh"


namespace fb200 {

template <typename T, int G, int NV, int VEC, bool TAILONLY, int NW, int RI>
__global__ void __launch_bounds__(WarpCfg<NW>::kThreads, 1) bma_reduce_cls_kernel(const ClsArgs a) {
  constexpr int kConsumerWarps = NW;
  constexpr int R = 32 / G;                 // rows per warp
  constexpr int TB = kConsumerWarps * R;    // rows per tile
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty_bar = full_bar + kMaxStages;
  unsigned char* stages = smem_raw + 128;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = a.S, B = a.B, C = a.C;

  if (threadIdx.x == 0) {
    for (int i = 0; i < a.n_stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], kConsumerWarps);
    }
    fence_mbar_init();
  }
  __syncthreads();

  const int n_chunks = (S + a.ss - 1) / a.ss;  // pipeline steps per tile

  if (warp >= kConsumerWarps) {
    // ------------------------------------------------------------------ producer warpgroup (warp 8 works)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(WarpCfg<NW>::kProducerRegs));
    if (warp != kConsumerWarps) return;
    const T* base = reinterpret_cast<const T*>(a.logits);
    uint32_t it = 0;
    for (int tv = blockIdx.x; tv < a.n_tiles; tv += gridDim.x) {
      int tile = tv + a.tile_rot;
      if (tile >= a.n_tiles) tile -= a.n_tiles;
      const int b0 = tile * TB;
      const int rows = min(TB, B - b0);
      const uint32_t bytes = (uint32_t)rows * (uint32_t)C * (uint32_t)sizeof(T);
      const bool bulk = a.bulk_ok && ((bytes & 15u) == 0);
      for (int ch = 0; ch < n_chunks; ++ch, ++it) {
        const int stage = it % a.n_stages;
        const uint32_t round = it / a.n_stages;
        if (round > 0) mbar_wait(&empty_bar[stage], (round - 1) & 1u);
        const int s0 = ch * a.ss;
        const int ns = min(a.ss, S - s0);
        unsigned char* dst = stages + (size_t)stage * a.stage_bytes;
        if (bulk) {
          if (lane == 0) {
            mbar_arrive_expect_tx(&full_bar[stage], bytes * (uint32_t)ns);
            for (int q = 0; q < ns; ++q) {
              const T* src = base + ((size_t)(s0 + q) * B + b0) * C;
              bulk_g2s(dst + (size_t)q * a.tile_stride_bytes, src, bytes, &full_bar[stage]);
            }
          }
        } else {
          // ragged tail tile or unaligned tensor: the producer warp copies element-wise
          const int n_el = rows * C;
          for (int q = 0; q < ns; ++q) {
            const T* src = base + ((size_t)(s0 + q) * B + b0) * C;
            T* d = reinterpret_cast<T*>(dst + (size_t)q * a.tile_stride_bytes);
            for (int i = lane; i < n_el; i += 32) d[i] = src[i];
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&full_bar[stage]);
        }
      }
    }
    return;
  }

  // -------------------------------------------------------------------- consumer warps
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(WarpCfg<NW>::kConsumerRegs));
  const int gr = lane / G;
  const int g = lane % G;
  const int row = warp * R + gr;
  // temperature: z = logits * exp(-log_temp) with exp(-log_temp) > 0, so max_c z = scale * max_c logits and the
  // scale folds into the single FFMA of pass B (k2 = scale * log2 e); no per-element multiply
  const float scale = a.log_temp ? expf(-a.log_temp[0]) : 1.0f;
  const float k2 = scale * 1.4426950408889634f;

  uint32_t it = 0;
  for (int tv = blockIdx.x; tv < a.n_tiles; tv += gridDim.x) {
    int tile = tv + a.tile_rot;
    if (tile >= a.n_tiles) tile -= a.n_tiles;
    const int b = tile * TB + row;
    const bool valid_row = b < B;
    constexpr int NP = Pairs<VEC>::kN;
    float2 sp[NV][NP], sp2[NV][NP];
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
      for (int k = 0; k < NP; ++k) sp[j][k] = sp2[j][k] = make_float2(0.0f, 0.0f);
    float spl = 0.0f, sum_l = 0.0f;
    float lmax = -CUDART_INF_F, lsum = 0.0f;
    int y = -1;
    if (a.targets && valid_row) y = a.targets[b];
    const bool y_ok = (y >= 0) && (y < C);

    for (int ch = 0; ch < n_chunks; ++ch, ++it) {
      const int stage = it % a.n_stages;
      const uint32_t round = it / a.n_stages;
      mbar_wait(&full_bar[stage], round & 1u);
      const int s0 = ch * a.ss;
      const int ns = min(a.ss, S - s0);
      const unsigned char* sbase = stages + (size_t)stage * a.stage_bytes;
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
        if (a.out.ensemble_log_prob && valid_row && g == 0) a.out.ensemble_log_prob[(size_t)s_idx * B + b] = ll;
      };
      int q = 0;
      if (RI == 2) {
        for (; q + 1 < ns; q += 2) {
          const T* rp[2];
          rp[0] = reinterpret_cast<const T*>(sbase + (size_t)q * a.tile_stride_bytes) + (size_t)row * C;
          rp[1] = reinterpret_cast<const T*>(sbase + (size_t)(q + 1) * a.tile_stride_bytes) + (size_t)row * C;
          float csh[2], l2s[2];
          process_rows<T, G, NV, VEC, TAILONLY, 2>(rp, g, C, k2, sp, sp2, spl, sum_l, csh, l2s);
          if (a.targets) {
            log_prob_update(rp[0], csh[0], l2s[0], s0 + q);
            log_prob_update(rp[1], csh[1], l2s[1], s0 + q + 1);
          }
        }
      }
      for (; q < ns; ++q) {
        const T* rp[1];
        rp[0] = reinterpret_cast<const T*>(sbase + (size_t)q * a.tile_stride_bytes) + (size_t)row * C;
        float csh[1], l2s[1];
        process_rows<T, G, NV, VEC, TAILONLY, 1>(rp, g, C, k2, sp, sp2, spl, sum_l, csh, l2s);
        if (a.targets) log_prob_update(rp[0], csh[0], l2s[0], s0 + q);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[stage]);
    }

    // ------------------------------------------------------------------ tile epilogue
    const float Sf = (float)S;
    const DivS dS = make_divs(Sf);
    const float spl_row = 0.6931471805599453f * (group_sum<G>(spl) - sum_l);  // sum_s sum_c p ln p of this row
    if (a.partial_mode) {
      // shard mode: this rank's partial sums for row b go to the slot of the rank that owns the row after the
      // exchange (owner = b / rows_per_rank).  A slot row is [sum_p C | sum_p2 C | sum plogp, ll_max, ll_sumexp, pad];
      // the slot pointer may be peer memory (NVLink stores straight from this epilogue) or a local send buffer.
      if (valid_row) {
        const int owner = b / a.rows_per_rank;
        const int lr = b - owner * a.rows_per_rank;
        float* base = reinterpret_cast<float*>(a.slot_ptrs[owner]) + (size_t)lr * (size_t)(2 * C + 4);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          const int c0 = (g + G * j) * VEC;
          if (c0 < C) {
            if constexpr (VEC == 4) {
              *reinterpret_cast<float4*>(base + c0) = make_float4(sp[j][0].x, sp[j][0].y, sp[j][1].x, sp[j][1].y);
              *reinterpret_cast<float4*>(base + C + c0) = make_float4(sp2[j][0].x, sp2[j][0].y, sp2[j][1].x, sp2[j][1].y);
            } else {
              base[c0] = sp[j][0].x;
              base[C + c0] = sp2[j][0].x;
            }
          }
        }
        if (g == 0) {
          base[2 * C] = spl_row;
          base[2 * C + 1] = lmax;
          base[2 * C + 2] = lsum;
          base[2 * C + 3] = 0.0f;
        }
      }
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
          ev[k] = clamp0_keep_nan(__fsub_rn(m2, __fmul_rn(mean[k], mean[k])));
          if (!(mean[k] <= 0.0f)) ent_terms = fmaf(mean[k], __logf(mean[k]), ent_terms);
          if (mean[k] > best_v) {
            best_v = mean[k];
            best_c = c0 + k;
          }
        }
        if (valid_row) {
          const size_t o = (size_t)b * C + c0;
          if constexpr (VEC == 4) {
            // C % 4 == 0 and 16-byte aligned outputs (checked on the host): one 128-bit store per output
            if (a.out.mean) st_f4(a.out.mean + o, mean[0], mean[1], mean[2], mean[3]);
            if (a.out.aleatoric_variance) st_f4(a.out.aleatoric_variance + o, av[0], av[1], av[2], av[3]);
            if (a.out.epistemic_variance) st_f4(a.out.epistemic_variance + o, ev[0], ev[1], ev[2], ev[3]);
            if (a.out.variance) st_f4(a.out.variance + o, av[0] + ev[0], av[1] + ev[1], av[2] + ev[2], av[3] + ev[3]);
            if (a.out.std)
              st_f4(a.out.std + o, sqrtf(av[0] + ev[0]), sqrtf(av[1] + ev[1]), sqrtf(av[2] + ev[2]), sqrtf(av[3] + ev[3]));
          } else {
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
              if (a.out.mean) a.out.mean[o + k] = mean[k];
              if (a.out.aleatoric_variance) a.out.aleatoric_variance[o + k] = av[k];
              if (a.out.epistemic_variance) a.out.epistemic_variance[o + k] = ev[k];
              if (a.out.variance) a.out.variance[o + k] = av[k] + ev[k];
              if (a.out.std) a.out.std[o + k] = sqrtf(av[k] + ev[k]);
            }
          }
        }
      }
    }
    const float ent = -group_sum<G>(ent_terms);
    const float alea = -spl_row / Sf;
    // argmax with first-maximum tie-break across the G lanes of the row
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
    if (valid_row && g == 0) {
      if (a.out.entropy) a.out.entropy[b] = ent;
      if (a.out.aleatoric_entropy) a.out.aleatoric_entropy[b] = alea;
      if (a.out.epistemic_entropy) a.out.epistemic_entropy[b] = ent - alea;
      if (a.out.mode) a.out.mode[b] = best_c == 0x7FFFFFFF ? 0 : best_c;
      if (a.out.log_prob) a.out.log_prob[b] = (lmax + logf(lsum)) - a.log_S;
      if (a.out.confidence) a.out.confidence[b] = best_c == 0x7FFFFFFF ? CUDART_NAN_F : best_v;  // a NaN row: max = NaN
      if (want_brier) a.out.brier_terms[b] = y_ok ? (msq - 2.0f * mean_y) + 1.0f : msq;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Narrow heads (C <= 16: two moons, MNIST / CIFAR-10 heads, BERT sentiment): a row is a handful of floats, so the
// warp-per-row machinery above spends its time in 16-lane shuffle reductions and the [S, B, C] tensor is a few tens of
// MB - latency-sized, not bandwidth-sized.  Here a whole row lives in ONE thread's registers (no shuffles inside a row)
// and 8 warps share a block of 32 inputs, each taking samples w, w + 8, ...; their partial sums - and the (max,
// sum-exp) pairs of log_prob - meet in shared memory and are added in stream order.  Same arithmetic per sample as process_rows (log2 units, MUFU.EX2),
// samples accumulated in 8 interleaved streams instead of one (inside the rtol 1e-5 of the parity tests).
// ------------------------------------------------------------------------------------------------
constexpr int kNarrowStreams = 8;  // warps per CTA = interleaved sample streams per input

template <typename T, int CMAX>
__global__ void __launch_bounds__(32 * kNarrowStreams) bma_reduce_cls_narrow_kernel(const ClsArgs a) {
  // CTA = 32 consecutive inputs (one per lane) x 8 sample streams (one per warp): a warp's loads of sample s are 32
  // adjacent rows = one contiguous 32*C*sizeof(T) span.  Partials meet in shared memory and warp 0 adds the streams in
  // order.
  __shared__ float part[kNarrowStreams][2 * CMAX + 4][32];
  const int S = a.S, B = a.B, C = a.C;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int b = blockIdx.x * 32 + lane;
  const bool valid = b < B;
  const float scale = a.log_temp ? expf(-a.log_temp[0]) : 1.0f;
  const float k2 = scale * 1.4426950408889634f;
  float sp[CMAX], sp2[CMAX];
#pragma unroll
  for (int c = 0; c < CMAX; ++c) sp[c] = sp2[c] = 0.0f;
  float spl = 0.0f, sum_l = 0.0f, lmax = -CUDART_INF_F, lsum = 0.0f;
  int y = -1;
  if (a.targets && valid) y = a.targets[b];
  const bool y_ok = (y >= 0) && (y < C);
  const T* base = reinterpret_cast<const T*>(a.logits);
  if (valid) {
    for (int s = w; s < S; s += kNarrowStreams) {
      const T* rp = base + ((size_t)s * B + b) * C;
      float z[CMAX];
      float m = -CUDART_INF_F;
#pragma unroll
      for (int c = 0; c < CMAX; ++c) {
        z[c] = -1.0e30f;
        if (c < C) {
          float tmp[1];
          RowLoad<T, 1>::load(rp + c, tmp);
          z[c] = tmp[0];
          m = fmaxf(m, z[c]);  // the row's own entries only: padding must not lend an all -inf row a finite maximum
        }
      }
      const float ms = (fabsf(m) <= 3.4028234e38f) ? m : 0.0f;
      const float csh = ms * k2;
      float sum = 0.0f, set = 0.0f;
      float zy = 0.0f;  // the target logit, selected without dynamic register indexing
#pragma unroll
      for (int c = 0; c < CMAX; ++c) {
        if (c == y) zy = z[c];
        const float u = fmaf(z[c], k2, -csh);
        const float ev = ex2_approx_ftz(u);
        z[c] = ev;
        sum += ev;
        set = fmaf(ev, u, set);
      }
      // softmax of a row whose maximum is not finite is NaN in every class (jax.nn.softmax; see process_rows)
      const float inv = (fabsf(m) <= 3.4028234e38f) ? rcp_approx_ftz(sum) : CUDART_NAN_F;
      const float inv2 = inv * inv;
      const float l2s = lg2_approx(sum);
      spl = fmaf(inv, set, spl);
      sum_l += l2s;
#pragma unroll
      for (int c = 0; c < CMAX; ++c) {
        sp[c] = fmaf(z[c], inv, sp[c]);
        sp2[c] = fmaf(z[c] * z[c], inv2, sp2[c]);
      }
      if (a.targets) {
        const float ll = 0.6931471805599453f * (fmaf(y_ok ? zy : 0.0f, k2, -csh) - l2s);
        if (ll > lmax) {
          lsum = lsum * __expf(lmax - ll) + 1.0f;
          lmax = ll;
        } else if (ll > -CUDART_INF_F) {
          lsum += __expf(ll - lmax);
        } else if (ll != ll) {
          lsum = ll;
        }
        if (a.out.ensemble_log_prob) a.out.ensemble_log_prob[(size_t)s * B + b] = ll;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < CMAX; ++c) {
    part[w][c][lane] = sp[c];
    part[w][CMAX + c][lane] = sp2[c];
  }
  part[w][2 * CMAX][lane] = spl;
  part[w][2 * CMAX + 1][lane] = sum_l;
  part[w][2 * CMAX + 2][lane] = lmax;
  part[w][2 * CMAX + 3][lane] = lsum;
  __syncthreads();
  if (w != 0 || !valid) return;
  for (int q = 1; q < kNarrowStreams; ++q) {
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      sp[c] += part[q][c][lane];
      sp2[c] += part[q][CMAX + c][lane];
    }
    spl += part[q][2 * CMAX][lane];
    sum_l += part[q][2 * CMAX + 1][lane];
    const float om = part[q][2 * CMAX + 2][lane], os = part[q][2 * CMAX + 3][lane];
    const float M = fmaxf(lmax, om);
    float acc = 0.0f;
    if (lsum != lsum || os != os) {
      acc = CUDART_NAN_F;  // a NaN log-likelihood in either stream poisons the sum, as in the sequential update
    } else {
      if (lmax > -CUDART_INF_F) acc += lsum * __expf(lmax - M);
      if (om > -CUDART_INF_F) acc += os * __expf(om - M);
    }
    lmax = M;
    lsum = acc;
  }
  const float Sf = (float)S;
  const float spl_row = 0.6931471805599453f * (spl - sum_l);
  float ent_terms = 0.0f, best_v = -CUDART_INF_F, msq = 0.0f, mean_y = 0.0f;
  int best_c = 0x7FFFFFFF;
#pragma unroll
  for (int c = 0; c < CMAX; ++c) {
    if (c < C) {
      const float mean = sp[c] / Sf;
      const float av = (sp[c] - sp2[c]) / Sf;
      const float ev = clamp0_keep_nan(__fsub_rn(sp2[c] / Sf, __fmul_rn(mean, mean)));
      msq = fmaf(mean, mean, msq);
      if (c == y) mean_y = mean;
      if (!(mean <= 0.0f)) ent_terms = fmaf(mean, logf(mean), ent_terms);
      if (mean > best_v) {
        best_v = mean;
        best_c = c;
      }
      const size_t o = (size_t)b * C + c;
      if (a.out.mean) a.out.mean[o] = mean;
      if (a.out.aleatoric_variance) a.out.aleatoric_variance[o] = av;
      if (a.out.epistemic_variance) a.out.epistemic_variance[o] = ev;
      if (a.out.variance) a.out.variance[o] = av + ev;
      if (a.out.std) a.out.std[o] = sqrtf(av + ev);
    }
  }
  const float ent = -ent_terms;
  const float alea = -spl_row / Sf;
  if (a.out.entropy) a.out.entropy[b] = ent;
  if (a.out.aleatoric_entropy) a.out.aleatoric_entropy[b] = alea;
  if (a.out.epistemic_entropy) a.out.epistemic_entropy[b] = ent - alea;
  if (a.out.mode) a.out.mode[b] = best_c == 0x7FFFFFFF ? 0 : best_c;
  if (a.out.log_prob) a.out.log_prob[b] = (lmax + logf(lsum)) - a.log_S;
  if (a.out.confidence) a.out.confidence[b] = best_c == 0x7FFFFFFF ? CUDART_NAN_F : best_v;  // a NaN row: max = NaN
  if (a.out.brier_terms) a.out.brier_terms[b] = y_ok ? (msq - 2.0f * mean_y) + 1.0f : msq;
}

template <typename T>
static int launch_cls_narrow(const ClsArgs& a, cudaStream_t st) {
  const unsigned grid = (unsigned)((a.B + 31) / 32);
  constexpr int th = 32 * kNarrowStreams;
  if (a.C <= 2) bma_reduce_cls_narrow_kernel<T, 2><<<grid, th, 0, st>>>(a);
  else if (a.C <= 4) bma_reduce_cls_narrow_kernel<T, 4><<<grid, th, 0, st>>>(a);
  else if (a.C <= 8) bma_reduce_cls_narrow_kernel<T, 8><<<grid, th, 0, st>>>(a);
  else bma_reduce_cls_narrow_kernel<T, 16><<<grid, th, 0, st>>>(a);
  FB200_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Wide heads (C > 1024: ImageNet-21k, vocabulary-sized heads): a row no longer fits a warp's registers, so the
// single-pass design does not apply.  Two passes over [S,B,C]: (1) the per-row logsumexp lse[s,b] (caller's workspace;
// row_lse pass), (2) one CTA per input walks its classes, each thread accumulating sum_s p, sum_s p^2 and sum_s p ln p
// for one class at a time with p = exp(z * scale - lse[s,b]); row-level results (entropies, mode) by a block reduction.
// log_prob / ensemble_log_prob come from the lse by gathers (log_prob_from_lse).  2x the HBM traffic of the fused kernel -
// the price of rows that do not fit on chip.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) wide_lse_kernel(const T* __restrict__ z, int64_t rows, int C,
                                                       const float* __restrict__ log_temp, float* __restrict__ lse) {
  const float scale = log_temp ? expf(-log_temp[0]) : 1.0f;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp0; r < rows; r += nwarps) {
    const T* row = z + r * C;
    float m = -CUDART_INF_F;
    for (int c = lane; c < C; c += 32) {
      float t[1];
      RowLoad<T, 1>::load(row + c, t);
      m = fmaxf(m, t[0] * scale);
    }
    m = group_max<32>(m);
    const float ms = (fabsf(m) <= 3.4028234e38f) ? m : 0.0f;
    float su = 0.0f;
    for (int c = lane; c < C; c += 32) {
      float t[1];
      RowLoad<T, 1>::load(row + c, t);
      su += __expf(t[0] * scale - ms);
    }
    su = group_sum<32>(su);
    if (lane == 0) lse[r] = ms + __logf(su);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) wide_reduce_kernel(const ClsArgs a, const float* __restrict__ lse) {
  const float scale = a.log_temp ? expf(-a.log_temp[0]) : 1.0f;
  __shared__ float s_lse[1024];  // lse[s, b] of this input for up to 1024 samples (more: read from global)
  __shared__ float red_f[8][4];
  __shared__ float red_v[8];
  __shared__ int red_c[8];
  const int S = a.S, B = a.B, C = a.C;
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  // jax.nn.softmax subtracts the row maximum as it is: a row holding +inf is NaN in EVERY class.  Only such a row
  // has lse = +inf (the log-probabilities keep logsumexp's convention and read the workspace as it is).
  auto softmax_lse = [](float v) { return v == CUDART_INF_F ? CUDART_NAN_F : v; };
  for (int s = tid; s < S && s < 1024; s += blockDim.x) s_lse[s] = softmax_lse(lse[(size_t)s * B + b]);
  __syncthreads();
  const T* base = reinterpret_cast<const T*>(a.logits) + (size_t)b * C;
  const size_t sstride = (size_t)B * C;
  const float Sf = (float)S;
  float ent_terms = 0.0f, plogp = 0.0f, best_v = -CUDART_INF_F, msq = 0.0f, mean_y = 0.0f;
  int best_c = 0x7FFFFFFF;
  const int y = a.targets ? a.targets[b] : -1;
  for (int c = tid; c < C; c += blockDim.x) {
    float sp = 0.0f, sp2 = 0.0f;
    for (int s = 0; s < S; ++s) {
      float t[1];
      RowLoad<T, 1>::load(base + (size_t)s * sstride + c, t);
      const float l = t[0] * scale - (s < 1024 ? s_lse[s] : softmax_lse(lse[(size_t)s * B + b]));  // log p
      const float p = __expf(l);
      sp += p;
      sp2 = fmaf(p, p, sp2);
      plogp = fmaf(p, l, plogp);  // 0 * -inf = NaN like the reference
    }
    const float mean = sp / Sf;
    const float av = (sp - sp2) / Sf;
    const float ev = clamp0_keep_nan(__fsub_rn(sp2 / Sf, __fmul_rn(mean, mean)));
    msq = fmaf(mean, mean, msq);
    if (c == y) mean_y = mean;
    if (!(mean <= 0.0f)) ent_terms = fmaf(mean, logf(mean), ent_terms);
    if (mean > best_v) {
      best_v = mean;
      best_c = c;
    }
    const size_t o = (size_t)b * C + c;
    if (a.out.mean) a.out.mean[o] = mean;
    if (a.out.aleatoric_variance) a.out.aleatoric_variance[o] = av;
    if (a.out.epistemic_variance) a.out.epistemic_variance[o] = ev;
    if (a.out.variance) a.out.variance[o] = av + ev;
    if (a.out.std) a.out.std[o] = sqrtf(av + ev);
  }
  ent_terms = group_sum<32>(ent_terms);
  plogp = group_sum<32>(plogp);
  msq = group_sum<32>(msq);
  mean_y = group_sum<32>(mean_y);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best_v, o);
    const int oc = __shfl_xor_sync(0xffffffffu, best_c, o);
    if (ov > best_v || (ov == best_v && oc < best_c)) {
      best_v = ov;
      best_c = oc;
    }
  }
  if (lane == 0) {
    red_f[w][0] = ent_terms;
    red_f[w][1] = plogp;
    red_f[w][2] = msq;
    red_f[w][3] = mean_y;
    red_v[w] = best_v;
    red_c[w] = best_c;
  }
  __syncthreads();
  if (tid == 0) {
    float e = 0.0f, pl = 0.0f, bv = -CUDART_INF_F, sq = 0.0f, my = 0.0f;
    int bc = 0x7FFFFFFF;
    for (int q = 0; q < (int)(blockDim.x >> 5); ++q) {
      e += red_f[q][0];
      pl += red_f[q][1];
      sq += red_f[q][2];
      my += red_f[q][3];
      if (red_v[q] > bv || (red_v[q] == bv && red_c[q] < bc)) {
        bv = red_v[q];
        bc = red_c[q];
      }
    }
    const float ent = -e, alea = -pl / Sf;
    if (a.out.entropy) a.out.entropy[b] = ent;
    if (a.out.aleatoric_entropy) a.out.aleatoric_entropy[b] = alea;
    if (a.out.epistemic_entropy) a.out.epistemic_entropy[b] = ent - alea;
    if (a.out.mode) a.out.mode[b] = bc == 0x7FFFFFFF ? 0 : bc;
    if (a.out.confidence) a.out.confidence[b] = bc == 0x7FFFFFFF ? CUDART_NAN_F : bv;  // jnp.max of a NaN row
    if (a.out.brier_terms) a.out.brier_terms[b] = (y >= 0 && y < C) ? (sq - 2.0f * my) + 1.0f : sq;
  }
}

// log_prob / ensemble_log_prob of the wide path: gathers against the row logsumexp (same update as the fused kernel)
template <typename T>
__global__ void __launch_bounds__(256) wide_log_prob_kernel(const ClsArgs a, const float* __restrict__ lse) {
  const float scale = a.log_temp ? expf(-a.log_temp[0]) : 1.0f;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  const int y = a.targets[b];
  const bool y_ok = (y >= 0) && (y < a.C);
  const T* base = reinterpret_cast<const T*>(a.logits);
  float lmax = -CUDART_INF_F, lsum = 0.0f;
  for (int s = 0; s < a.S; ++s) {
    const size_t row = (size_t)s * a.B + b;
    float zt = 0.0f;
    if (y_ok) {
      float t[1];
      RowLoad<T, 1>::load(base + row * a.C + y, t);
      zt = t[0];
    }
    const float ll = zt * scale - lse[row];
    if (ll > lmax) {
      lsum = lsum * __expf(lmax - ll) + 1.0f;
      lmax = ll;
    } else if (ll > -CUDART_INF_F) {
      lsum += __expf(ll - lmax);
    } else if (ll != ll) {
      lsum = ll;
    }
    if (a.out.ensemble_log_prob) a.out.ensemble_log_prob[row] = ll;
  }
  if (a.out.log_prob) a.out.log_prob[b] = (lmax + logf(lsum)) - a.log_S;
}

template <typename T>
static int launch_cls_wide(const ClsArgs& a, float* lse_ws, cudaStream_t st) {
  const int64_t rows = (int64_t)a.S * a.B;
  int64_t blocks = (rows + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  wide_lse_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const T*>(a.logits), rows, a.C, a.log_temp, lse_ws);
  FB200_LAUNCH_CHECK();
  wide_reduce_kernel<T><<<(unsigned)a.B, 256, 0, st>>>(a, lse_ws);
  FB200_LAUNCH_CHECK();
  if (a.targets && (a.out.log_prob || a.out.ensemble_log_prob)) {
    wide_log_prob_kernel<T><<<(unsigned)((a.B + 255) / 256), 256, 0, st>>>(a, lse_ws);
    FB200_LAUNCH_CHECK();
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Finalise the rows this rank owns from the n_src slots it received (one per source rank, same row layout as
// above): sums are taken in source-rank order (deterministic), the (max, sumexp) pairs are merged, then the same
// finalisation as the single-GPU kernel.  One warp per row.
// ------------------------------------------------------------------------------------------------
template <bool V4, int NSRC>  // NSRC > 0: the source count at compile time (the loads of all sources are in flight at once)
__global__ void __launch_bounds__(256) bma_finalize_slots_kernel(const float* __restrict__ slots, int n_src, int rows,
                                                                 int C, int s_total, const fb200_cls_outputs_t out,
                                                                 const int32_t* __restrict__ targets, float log_S) {
  // One warp per owned row: the N source rows are streamed with 128-bit loads (V4: C % 4 == 0; a slot row is 2C + 4
  // floats, so every row and its second half start 16-byte aligned), summed in source-rank order, finalised like the
  // single-GPU epilogue (div_s: correctly rounded x / S without the IEEE-division slow path).  HBM-bound:
  // n_src (2C + 4) floats read + the requested outputs written per row.
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int b = warp;
  const size_t RS = (size_t)(2 * C + 4);
  const size_t src_stride = (size_t)rows * RS;
  const float* row0 = slots + (size_t)b * RS;
  const DivS dS = make_divs((float)s_total);
  const int y = targets ? targets[b] : -1;
  float ent_terms = 0.0f, msq = 0.0f, mean_y = 0.0f;
  float best_v = -CUDART_INF_F;
  int best_c = 0x7FFFFFFF;
  constexpr int W = V4 ? 4 : 1;
  for (int c0 = lane * W; c0 < C; c0 += 32 * W) {
    float sp[W], sp2[W];
#pragma unroll
    for (int k = 0; k < W; ++k) sp[k] = sp2[k] = 0.0f;
    if constexpr (V4 && NSRC > 0) {
      float4 a[NSRC], q[NSRC];
#pragma unroll
      for (int r = 0; r < NSRC; ++r) {
        const float* pr = row0 + (size_t)r * src_stride;
        a[r] = *reinterpret_cast<const float4*>(pr + c0);
        q[r] = *reinterpret_cast<const float4*>(pr + C + c0);
      }
#pragma unroll
      for (int r = 0; r < NSRC; ++r) {  // same source-rank order as the rolled loop
        sp[0] += a[r].x; sp[1] += a[r].y; sp[2] += a[r].z; sp[3] += a[r].w;
        sp2[0] += q[r].x; sp2[1] += q[r].y; sp2[2] += q[r].z; sp2[3] += q[r].w;
      }
    } else {
      for (int r = 0; r < n_src; ++r) {
        const float* pr = row0 + (size_t)r * src_stride;
        if constexpr (V4) {
          const float4 a = *reinterpret_cast<const float4*>(pr + c0);
          const float4 q = *reinterpret_cast<const float4*>(pr + C + c0);
          sp[0] += a.x; sp[1] += a.y; sp[2] += a.z; sp[3] += a.w;
          sp2[0] += q.x; sp2[1] += q.y; sp2[2] += q.z; sp2[3] += q.w;
        } else {
          sp[0] += pr[c0];
          sp2[0] += pr[C + c0];
        }
      }
    }
    float mean[W], av[W], ev[W];
#pragma unroll
    for (int k = 0; k < W; ++k) {
      mean[k] = div_s(sp[k], dS);
      av[k] = div_s(sp[k] - sp2[k], dS);
      ev[k] = clamp0_keep_nan(__fsub_rn(div_s(sp2[k], dS), __fmul_rn(mean[k], mean[k])));
      if (!(mean[k] <= 0.0f)) ent_terms = fmaf(mean[k], __logf(mean[k]), ent_terms);
      msq = fmaf(mean[k], mean[k], msq);
      mean_y = (c0 + k == y) ? mean[k] : mean_y;
      if (mean[k] > best_v) {
        best_v = mean[k];
        best_c = c0 + k;
      }
    }
    const size_t o = (size_t)b * C + c0;
    if constexpr (V4) {
      if (out.mean) st_f4(out.mean + o, mean[0], mean[1], mean[2], mean[3]);
      if (out.aleatoric_variance) st_f4(out.aleatoric_variance + o, av[0], av[1], av[2], av[3]);
      if (out.epistemic_variance) st_f4(out.epistemic_variance + o, ev[0], ev[1], ev[2], ev[3]);
      if (out.variance) st_f4(out.variance + o, av[0] + ev[0], av[1] + ev[1], av[2] + ev[2], av[3] + ev[3]);
      if (out.std)
        st_f4(out.std + o, sqrtf(av[0] + ev[0]), sqrtf(av[1] + ev[1]), sqrtf(av[2] + ev[2]), sqrtf(av[3] + ev[3]));
    } else {
      if (out.mean) out.mean[o] = mean[0];
      if (out.aleatoric_variance) out.aleatoric_variance[o] = av[0];
      if (out.epistemic_variance) out.epistemic_variance[o] = ev[0];
      if (out.variance) out.variance[o] = av[0] + ev[0];
      if (out.std) out.std[o] = sqrtf(av[0] + ev[0]);
    }
  }
  const float ent = -group_sum<32>(ent_terms);
  msq = group_sum<32>(msq);
  mean_y = group_sum<32>(mean_y);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best_v, o);
    const int oc = __shfl_xor_sync(0xffffffffu, best_c, o);
    if (ov > best_v || (ov == best_v && oc < best_c)) {
      best_v = ov;
      best_c = oc;
    }
  }
  if (lane == 0) {
    float plogp = 0.0f, M = -CUDART_INF_F;
    for (int r = 0; r < n_src; ++r) {
      plogp += row0[(size_t)r * src_stride + 2 * C];
      M = fmaxf(M, row0[(size_t)r * src_stride + 2 * C + 1]);
    }
    float ssum = 0.0f;
    for (int r = 0; r < n_src; ++r) {
      const float mr = row0[(size_t)r * src_stride + 2 * C + 1];
      const float sr = row0[(size_t)r * src_stride + 2 * C + 2];
      if (mr > -CUDART_INF_F) ssum += sr * __expf(mr - M);
      else if (sr != sr) ssum = sr;  // NaN propagates like the reference's logsumexp
    }
    const float alea = -plogp / (float)s_total;
    if (out.entropy) out.entropy[b] = ent;
    if (out.aleatoric_entropy) out.aleatoric_entropy[b] = alea;
    if (out.epistemic_entropy) out.epistemic_entropy[b] = ent - alea;
    if (out.mode) out.mode[b] = best_c == 0x7FFFFFFF ? 0 : best_c;
    if (out.confidence) out.confidence[b] = best_c == 0x7FFFFFFF ? CUDART_NAN_F : best_v;  // a NaN row: max = NaN
    if (out.log_prob) out.log_prob[b] = (M + logf(ssum)) - log_S;
    if (out.brier_terms) out.brier_terms[b] = (y >= 0 && y < C) ? (msq - 2.0f * mean_y) + 1.0f : msq;
  }
}

// ------------------------------------------------------------------------------------------------
// Regression maps: z[S,B,2d] = [mu, log var]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bma_reduce_reg_moments_kernel(const float* __restrict__ z,
                                                                     const float* __restrict__ log_temp, int S,
                                                                     int B, int d, const fb200_reg_outputs_t out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // (b, j)
  if (i >= (int64_t)B * d) return;
  const int b = (int)(i / d), j = (int)(i % d);
  const float lt = log_temp ? log_temp[0] : 0.0f;
  float smu = 0.f, smu2 = 0.f, sv = 0.f;
  for (int s = 0; s < S; ++s) {
    const float* r = z + ((size_t)s * B + b) * (2 * d);
    const float mu = r[j];
    float lv = r[d + j];
    if (log_temp) lv += lt;
    smu += mu;
    smu2 = fmaf(mu, mu, smu2);
    sv += expf(lv);
  }
  const float Sf = (float)S;
  const float mean = smu / Sf;
  const float av = sv / Sf;
  const float ev = clamp0_keep_nan(__fsub_rn(smu2 / Sf, __fmul_rn(mean, mean)));
  if (out.mean) out.mean[i] = mean;
  if (out.aleatoric_variance) out.aleatoric_variance[i] = av;
  if (out.epistemic_variance) out.epistemic_variance[i] = ev;
  if (out.variance) out.variance[i] = av + ev;
  if (out.std) out.std[i] = sqrtf(av + ev);
}

__global__ void __launch_bounds__(256) bma_reduce_reg_logprob_kernel(const float* __restrict__ z,
                                                                     const float* __restrict__ log_temp,
                                                                     const float* __restrict__ targets, int S,
                                                                     int B, int d, const fb200_reg_outputs_t out,
                                                                     float log_S) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float lt = log_temp ? log_temp[0] : 0.0f;
  const float log2pi = 1.8378770664093453f;
  float lmax = -CUDART_INF_F, lsum = 0.0f;
  for (int s = 0; s < S; ++s) {
    const float* r = z + ((size_t)s * B + b) * (2 * d);
    float acc = 0.0f;
    for (int j = 0; j < d; ++j) {
      const float mu = r[j];
      float lv = r[d + j];
      if (log_temp) lv += lt;
      const float diff = targets[(size_t)b * d + j] - mu;
      acc += expf(-lv) * (diff * diff) + lv + log2pi;
    }
    const float ll = -0.5f * acc;
    if (out.ensemble_log_prob) out.ensemble_log_prob[(size_t)s * B + b] = ll;
    if (ll > lmax) {
      lsum = lsum * __expf(lmax - ll) + 1.0f;
      lmax = ll;
    } else if (ll > -CUDART_INF_F) {
      lsum += __expf(ll - lmax);
    } else if (ll != ll) {
      lsum = ll;
    }
  }
  if (out.log_prob) out.log_prob[b] = (lmax + logf(lsum)) - log_S;
}

// Target samples of the regression predictive (fortuna/prob_model/predictive/base.py:407-441 ->
// fortuna/likelihood/base.py:339-372 -> fortuna/prob_output_layer/regression.py:38-52), bit-exact draws:
//   keys = split(rng, T); (k1, k2) = split(keys[t]); i = choice(k1, n_stored);
//   y[t] = mu_i + exp(0.5 * logvar_i) * normal(k2, (1, B, d))
__global__ void __launch_bounds__(256) reg_sample_targets_kernel(const float* __restrict__ z,
                                                                 const float* __restrict__ log_temp,
                                                                 const uint32_t* __restrict__ key, uint32_t T,
                                                                 uint32_t n_stored, int B, int d,
                                                                 float* __restrict__ y, int32_t* __restrict__ idx_out,
                                                                 int direct) {
  const uint32_t t = blockIdx.y;
  const uint32_t n = (uint32_t)B * (uint32_t)d;
  if (direct) {
    // RegressionProbOutputLayer.sample on ONE set of outputs (fortuna/prob_output_layer/regression.py:38-52, the
    // output_calib_model predictive): y = mu + exp(0.5 logvar) * normal(rng, (T, B, d)) - one stream of T*B*d normals
    const float lt = log_temp ? log_temp[0] : 0.0f;
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
      const uint32_t b = e / (uint32_t)d, j = e - b * (uint32_t)d;
      const float mu = z[(size_t)b * (2 * d) + j];
      float lv = z[(size_t)b * (2 * d) + d + j];
      if (log_temp) lv += lt;
      const float eps = bits_to_normal(random_bits_elem(key[0], key[1], T * n, t * n + e));
      y[(size_t)t * n + e] = mu + expf(0.5f * lv) * eps;
    }
    return;
  }
  // per-t key schedule (identical in every thread of the row; a handful of threefry blocks)
  const uint32_t kt0 = random_bits_elem(key[0], key[1], 2u * T, 2u * t);
  const uint32_t kt1 = random_bits_elem(key[0], key[1], 2u * T, 2u * t + 1u);
  uint32_t a0 = 0u, a1 = 2u, b0 = 1u, b1 = 3u;  // split(keys[t]) = bits(4).reshape(2,2): key1 = (a0, b0), key2 = (a1, b1)
  threefry2x32(kt0, kt1, a0, a1);
  threefry2x32(kt0, kt1, b0, b1);
  // choice(key1, n_stored) = randint: split(key1) -> two single draws
  uint32_t c0 = 0u, c1 = 2u, d0 = 1u, d1 = 3u;
  threefry2x32(a0, b0, c0, c1);
  threefry2x32(a0, b0, d0, d1);
  uint32_t hi = 0u, hi1 = 0u, lo = 0u, lo1 = 0u;
  threefry2x32(c0, d0, hi, hi1);
  threefry2x32(c1, d1, lo, lo1);
  uint32_t mult = 65536u % n_stored;
  mult = (mult * mult) % n_stored;
  const uint32_t i = ((hi % n_stored) * mult + (lo % n_stored)) % n_stored;
  if (idx_out && blockIdx.x == 0 && threadIdx.x == 0) idx_out[t] = (int32_t)i;
  const float lt = log_temp ? log_temp[0] : 0.0f;
  const float* zi = z + (size_t)i * B * (2 * d);
  for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const uint32_t b = e / (uint32_t)d, j = e - b * (uint32_t)d;
    const float mu = zi[(size_t)b * (2 * d) + j];
    float lv = zi[(size_t)b * (2 * d) + d + j];
    if (log_temp) lv += lt;
    const float eps = bits_to_normal(random_bits_elem(a1, b1, n, e));
    y[(size_t)t * n + e] = mu + expf(0.5f * lv) * eps;
  }
}


template <typename T, int G, int NV, int VEC, bool TAILONLY, int NW, int RI>
static int launch_cls_cfg(ClsArgs a, cudaStream_t st) {
  constexpr int R = 32 / G;
  constexpr int TB = NW * R;
  constexpr int kReduceThreads = WarpCfg<NW>::kThreads;
  const uint32_t tile_bytes = (uint32_t)TB * (uint32_t)a.C * (uint32_t)sizeof(T);
  a.tile_stride_bytes = (tile_bytes + 15u) & ~15u;
  // aim for >= 16 KB per stage so a handful of stages covers the HBM latency-bandwidth product
  int ss = (int)(16384u / a.tile_stride_bytes);
  if (ss < RI) ss = RI;  // RI samples per stage at least: the consumers interleave them for ILP
  if (ss > 16) ss = 16;
  if (ss > a.S) ss = a.S;
  a.ss = ss;
  a.stage_bytes = ((uint32_t)ss * a.tile_stride_bytes + 127u) & ~127u;
  const uint32_t budget = 200u * 1024u;
  int nst = (int)(budget / a.stage_bytes);
  if (nst > kMaxStages) nst = kMaxStages;
  if (nst < 2) nst = 2;
  a.n_stages = nst;
  a.n_tiles = (a.B + TB - 1) / TB;
  a.tile_rot = a.partial_mode ? (int)(((long long)a.n_tiles * a.tile_rot) / (a.n_ranks > 0 ? a.n_ranks : 1)) : 0;  // rank -> first tile
  const size_t smem = 128 + (size_t)nst * a.stage_bytes;
  auto kern = bma_reduce_cls_kernel<T, G, NV, VEC, TAILONLY, NW, RI>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  int occ = 1;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kReduceThreads, smem);
  if (e != cudaSuccess) return (int)e;
  if (occ < 1) occ = 1;
  int grid = sm_count() * occ;
  if (grid > a.n_tiles) grid = a.n_tiles;
  kern<<<grid, kReduceThreads, smem, st>>>(a);
  FB200_LAUNCH_CHECK();
  return 0;
}

// Warp / interleave configuration per row width.  Wide rows (C > 512: 64 accumulator + 64 row registers per lane)
// run 12 consumer warps x 2 interleaved samples (160 registers); measured at c4 on B200: 12x2 2.93 ms,
// 16x1 3.00 ms, 12x1 3.36 ms, 8x2 3.34 ms.  Narrow rows have registers to spare: 16 warps x 2.
template <typename T, int G, int NV, int VEC, bool TAILONLY>
static int launch_cls(const ClsArgs& a, cudaStream_t st) {
  if constexpr (NV * VEC >= 20) {
    return launch_cls_cfg<T, G, NV, VEC, TAILONLY, 12, 2>(a, st);
  } else {
    return launch_cls_cfg<T, G, NV, VEC, TAILONLY, 16, 2>(a, st);
  }
}

template <typename T>
static int dispatch_cls(const ClsArgs& a, cudaStream_t st) {
  const int C = a.C;
  const int vec = (C % 4 == 0) ? 4 : 1;
  const int chunks = C / vec;
#define FB200_CASE(G_, NV_, VEC_, TO_) return launch_cls<T, G_, NV_, VEC_, TO_>(a, st)
  if (vec == 4) {
    if (chunks <= 1) FB200_CASE(1, 1, 4, true);
    if (chunks <= 2) FB200_CASE(2, 1, 4, true);
    if (chunks <= 4) FB200_CASE(4, 1, 4, true);
    if (chunks <= 8) FB200_CASE(8, 1, 4, true);
    if (chunks <= 16) FB200_CASE(16, 1, 4, true);
    // G = 32: NV = ceil(chunks / 32) exactly, so only the last chunk slot can be ragged
    if (chunks <= 32) FB200_CASE(32, 1, 4, true);
    if (chunks <= 64) FB200_CASE(32, 2, 4, true);
    if (chunks <= 96) FB200_CASE(32, 3, 4, true);
    if (chunks <= 128) FB200_CASE(32, 4, 4, true);
    if (chunks <= 160) FB200_CASE(32, 5, 4, true);
    if (chunks <= 192) FB200_CASE(32, 6, 4, true);
    if (chunks <= 224) FB200_CASE(32, 7, 4, true);
    if (chunks <= 256) FB200_CASE(32, 8, 4, true);
  } else {
    if (chunks <= 1) FB200_CASE(1, 1, 1, true);
    if (chunks <= 2) FB200_CASE(2, 1, 1, true);
    if (chunks <= 4) FB200_CASE(4, 1, 1, true);
    if (chunks <= 8) FB200_CASE(8, 1, 1, true);
    if (chunks <= 16) FB200_CASE(16, 1, 1, true);
    if (chunks <= 32) FB200_CASE(32, 1, 1, true);
    if (chunks <= 64) FB200_CASE(32, 2, 1, true);
    if (chunks <= 128) FB200_CASE(32, 4, 1, false);
    if (chunks <= 256) FB200_CASE(32, 8, 1, false);
    if (chunks <= 512) FB200_CASE(32, 16, 1, false);
    if (chunks <= 1024) FB200_CASE(32, 32, 1, false);
  }
#undef FB200_CASE
  return FB200_E_UNSUPPORTED;
}

}  // namespace fb200

using namespace fb200;

static int reduce_cls_common(ClsArgs& a, const void* logits, int dtype, const float* log_temp,
                             const int32_t* targets, int S, int B, int C, void* stream) {
  if (!logits) return FB200_E_NULL;
  if (S <= 0 || B <= 0 || C <= 0) return FB200_E_SHAPE;
  if (C > 1024) return FB200_E_UNSUPPORTED;
  if (dtype != FB200_F32 && dtype != FB200_BF16) return FB200_E_UNSUPPORTED;
  a.logits = logits;
  a.log_temp = log_temp;
  a.targets = targets;
  a.S = S;
  a.B = B;
  a.C = C;
  a.log_S = logf((float)S);
  const size_t esz = dtype == FB200_F32 ? 4 : 2;
  a.bulk_ok = fb200_aligned16(logits) && (((size_t)B * (size_t)C * esz) % 16 == 0);
  if (C % 4 == 0 && !(fb200_aligned16(a.out.mean) && fb200_aligned16(a.out.aleatoric_variance) &&
                      fb200_aligned16(a.out.epistemic_variance) && fb200_aligned16(a.out.variance) &&
                      fb200_aligned16(a.out.std)))
    return FB200_E_ALIGN;  // [B,C] outputs are written with 128-bit stores when C % 4 == 0
  if (C <= 16 && !a.partial_mode) {
    if (dtype == FB200_F32) return launch_cls_narrow<float>(a, (cudaStream_t)stream);
    return launch_cls_narrow<__nv_bfloat16>(a, (cudaStream_t)stream);
  }
  if (dtype == FB200_F32) return dispatch_cls<float>(a, (cudaStream_t)stream);
  return dispatch_cls<__nv_bfloat16>(a, (cudaStream_t)stream);
}

extern "C" int fb200_bma_reduce_cls(const void* logits, int dtype, const float* log_temp, const int32_t* targets,
                                    int S, int B, int C, const fb200_cls_outputs_t* out, void* stream) {
  if (!out) return FB200_E_NULL;
  ClsArgs a{};
  a.out = *out;
  if ((a.out.log_prob || a.out.ensemble_log_prob || a.out.brier_terms) && !targets) return FB200_E_NULL;
  return reduce_cls_common(a, logits, dtype, log_temp, targets, S, B, C, stream);
}

extern "C" size_t fb200_bma_reduce_cls_wide_workspace_bytes(int S, int B) {
  return (S > 0 && B > 0) ? (size_t)S * (size_t)B * sizeof(float) : 0;
}

extern "C" int fb200_bma_reduce_cls_wide(const void* logits, int dtype, const float* log_temp, const int32_t* targets,
                                         int S, int B, int C, const fb200_cls_outputs_t* out, void* workspace,
                                         size_t workspace_bytes, void* stream) {
  if (!logits || !out || !workspace) return FB200_E_NULL;
  if (S <= 0 || B <= 0 || C <= 0) return FB200_E_SHAPE;
  if (dtype != FB200_F32 && dtype != FB200_BF16) return FB200_E_UNSUPPORTED;
  if ((out->log_prob || out->ensemble_log_prob || out->brier_terms) && !targets) return FB200_E_NULL;
  if (workspace_bytes < fb200_bma_reduce_cls_wide_workspace_bytes(S, B)) return FB200_E_WORKSPACE;
  ClsArgs a{};
  a.out = *out;
  a.logits = logits;
  a.log_temp = log_temp;
  a.targets = targets;
  a.S = S;
  a.B = B;
  a.C = C;
  a.log_S = logf((float)S);
  float* lse_ws = reinterpret_cast<float*>(workspace);
  if (dtype == FB200_F32) return launch_cls_wide<float>(a, lse_ws, (cudaStream_t)stream);
  return launch_cls_wide<__nv_bfloat16>(a, lse_ws, (cudaStream_t)stream);
}

extern "C" size_t fb200_cls_slot_row_floats(int C) { return C > 0 ? (size_t)(2 * C + 4) : 0; }

extern "C" int fb200_bma_reduce_cls_shard(const void* logits, int dtype, const float* log_temp,
                                          const int32_t* targets, int S, int B, int C, int n_ranks, int rank,
                                          const uint64_t* slot_ptrs, float* ensemble_log_prob, void* stream) {
  if (!slot_ptrs) return FB200_E_NULL;
  if (ensemble_log_prob && !targets) return FB200_E_NULL;
  if (n_ranks <= 0 || B <= 0 || B % n_ranks != 0 || rank < 0 || rank >= n_ranks) return FB200_E_SHAPE;
  ClsArgs a{};
  a.partial_mode = 1;
  a.out.ensemble_log_prob = ensemble_log_prob;
  a.slot_ptrs = reinterpret_cast<const unsigned long long*>(slot_ptrs);
  a.n_ranks = n_ranks;
  a.tile_rot = rank;  // turned into a tile index once the tile size is known (launch_cls_cfg)
  a.rows_per_rank = B / n_ranks;
  return reduce_cls_common(a, logits, dtype, log_temp, targets, S, B, C, stream);
}

extern "C" int fb200_bma_finalize_cls_slots(const float* slots, int n_src, int rows, int C, int s_total,
                                            const fb200_cls_outputs_t* out, const int32_t* targets, void* stream) {
  if (!slots || !out) return FB200_E_NULL;
  if (out->ensemble_log_prob) return FB200_E_UNSUPPORTED;  // per-sample values live on the rank that holds the sample
  if (out->brier_terms && !targets) return FB200_E_NULL;   // the Brier term needs the owned rows' targets
  if (n_src <= 0 || rows <= 0 || C <= 0 || s_total <= 0) return FB200_E_SHAPE;
  const int64_t threads = (int64_t)rows * 32;
  const unsigned grid = (unsigned)((threads + 255) / 256);
  const bool v4 = (C % 4 == 0) && fb200_aligned16(slots) && fb200_aligned16(out->mean) &&
                  fb200_aligned16(out->aleatoric_variance) && fb200_aligned16(out->epistemic_variance) &&
                  fb200_aligned16(out->variance) && fb200_aligned16(out->std);
  const float log_S = logf((float)s_total);
  cudaStream_t st = (cudaStream_t)stream;
  if (v4 && n_src == 8)
    bma_finalize_slots_kernel<true, 8><<<grid, 256, 0, st>>>(slots, n_src, rows, C, s_total, *out, targets, log_S);
  else if (v4 && n_src == 4)
    bma_finalize_slots_kernel<true, 4><<<grid, 256, 0, st>>>(slots, n_src, rows, C, s_total, *out, targets, log_S);
  else if (v4 && n_src == 2)
    bma_finalize_slots_kernel<true, 2><<<grid, 256, 0, st>>>(slots, n_src, rows, C, s_total, *out, targets, log_S);
  else if (v4)
    bma_finalize_slots_kernel<true, 0><<<grid, 256, 0, st>>>(slots, n_src, rows, C, s_total, *out, targets, log_S);
  else
    bma_finalize_slots_kernel<false, 0><<<grid, 256, 0, st>>>(slots, n_src, rows, C, s_total, *out, targets, log_S);
  FB200_LAUNCH_CHECK();
  return 0;
}

// predictive.variance / predictive.std of the reference (fortuna/prob_model/predictive/base.py:838-944) combine an
// aleatoric and an epistemic estimate that may come from DIFFERENT posterior draws: variance = a + e, std = sqrt(variance).
__global__ void __launch_bounds__(256) variance_combine_kernel(const float* __restrict__ a, const float* __restrict__ e,
                                                               int64_t n, float* __restrict__ var_out,
                                                               float* __restrict__ std_out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = e ? __fadd_rn(a[i], e[i]) : a[i];
    if (var_out) var_out[i] = v;
    if (std_out) std_out[i] = sqrtf(v);
  }
}

extern "C" int fb200_variance_combine(const float* aleatoric, const float* epistemic, int64_t n, float* variance_out,
                                      float* std_out, void* stream) {
  if (!aleatoric || (!variance_out && !std_out)) return FB200_E_NULL;
  if (n <= 0) return FB200_E_SHAPE;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  variance_combine_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(aleatoric, epistemic, n, variance_out, std_out);
  FB200_LAUNCH_CHECK();
  return 0;
}

static int reg_sample_common(const float* outputs, const float* log_temp, const uint32_t* key, int n_target_samples,
                             int n_stored, int B, int d, float* y, int32_t* idx_out, int direct, void* stream) {
  if (!outputs || !key || !y) return FB200_E_NULL;
  if (n_target_samples <= 0 || n_stored <= 0 || B <= 0 || d <= 0) return FB200_E_SHAPE;
  if ((long long)B * d >= 0xFFFFFFFFll || n_target_samples > 65535) return FB200_E_UNSUPPORTED;
  const long long n = (long long)B * d;
  if (direct && n * n_target_samples >= 0xFFFFFFFFll) return FB200_E_UNSUPPORTED;  // one 32-bit counter stream
  unsigned gx = (unsigned)((n + 255) / 256);
  if (gx > 148u * 8u) gx = 148u * 8u;
  reg_sample_targets_kernel<<<dim3(gx, (unsigned)n_target_samples), 256, 0, (cudaStream_t)stream>>>(
      outputs, log_temp, key, (uint32_t)n_target_samples, (uint32_t)n_stored, B, d, y, idx_out, direct);
  FB200_LAUNCH_CHECK();
  return 0;
}

extern "C" int fb200_reg_sample_targets(const float* outputs, const float* log_temp, const uint32_t* key,
                                        int n_target_samples, int n_stored, int B, int d, float* y, int32_t* idx_out,
                                        void* stream) {
  return reg_sample_common(outputs, log_temp, key, n_target_samples, n_stored, B, d, y, idx_out, 0, stream);
}

extern "C" int fb200_reg_output_sample_targets(const float* outputs, const float* log_temp, const uint32_t* key,
                                               int n_target_samples, int B, int d, float* y, void* stream) {
  return reg_sample_common(outputs, log_temp, key, n_target_samples, 1, B, d, y, nullptr, 1, stream);
}

extern "C" int fb200_bma_reduce_reg(const float* outputs, const float* log_temp, const float* targets, int S,
                                    int B, int d, const fb200_reg_outputs_t* out, void* stream) {
  if (!outputs || !out) return FB200_E_NULL;
  if (S <= 0 || B <= 0 || d <= 0) return FB200_E_SHAPE;
  if ((out->log_prob || out->ensemble_log_prob) && !targets) return FB200_E_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  if (out->mean || out->aleatoric_variance || out->epistemic_variance || out->variance || out->std) {
    const int64_t n = (int64_t)B * d;
    bma_reduce_reg_moments_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(outputs, log_temp, S, B, d, *out);
    FB200_LAUNCH_CHECK();
  }
  if (out->log_prob || out->ensemble_log_prob) {
    bma_reduce_reg_logprob_kernel<<<(B + 255) / 256, 256, 0, st>>>(outputs, log_temp, targets, S, B, d, *out,
                                                                   logf((float)S));
    FB200_LAUNCH_CHECK();
  }
  return 0;
}
