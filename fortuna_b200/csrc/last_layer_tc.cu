// K2 (tensor-core path): S-batched last-layer GEMM on tcgen05 / TMEM / TMA for sm_100a.
//
//   out[s,b,c] = (sum_d x[b,d] * wt[s,c,d] + bias[s,c]) * exp(-log_temp)
//
// Replaces Likelihood._get_batched_calibrated_outputs restricted to the final nn.Dense
// (fortuna/likelihood/base.py:429-458; fortuna/model/mlp.py:227, lenet.py:129, resnet.py:248) + the
// temperature scaler (fortuna/output_calibrator/classification.py:14-17) for all S samples in one launch.
//
// Operands are bf16, K-major: A = x [B, D] (M = B), B = wt [S, C, D] (N = C of sample s).  A CTA owns a
// 128 x 256 output tile of one sample and walks K in 64-element (128-byte, SWIZZLE_128B) slabs:
//   warp 0   : TMA producer  (cp.async.bulk.tensor into a 4-stage shared-memory ring, mbarrier tx-count)
//   warp 1   : TMEM allocator + single-thread tcgen05.mma issuer (M128 N256 K16, fp32 accumulators in TMEM),
//              tcgen05.commit releases smem stages and publishes finished accumulators
//   warps 2-5: epilogue - tcgen05.ld the accumulator quadrant of the warp, + bias, * scale, store
// Two TMEM accumulator stages (2 x 256 columns) let the epilogue of tile i overlap the MMAs of tile i+1.
// The grid is persistent (one CTA per SM); tiles are ordered so that concurrently running CTAs share the
// same W tile and a group of M-blocks keeps its x rows L2-resident while it sweeps all (s, n) tiles.
#include "tc_common.cuh"

namespace fb200 {

constexpr int TC_BLOCK_M = 128;
constexpr int TC_BLOCK_N = 256;
constexpr int TC_BLOCK_K = 64;  // 64 bf16 = 128 bytes = one SWIZZLE_128B row
constexpr int TC_UMMA_K = 16;
constexpr int TC_STAGES = 4;
constexpr int TC_ACC_STAGES = 2;
constexpr int TC_THREADS = 192;
constexpr uint32_t TC_A_BYTES = TC_BLOCK_M * TC_BLOCK_K * 2;  // 16 KB
constexpr uint32_t TC_B_BYTES = TC_BLOCK_N * TC_BLOCK_K * 2;  // 32 KB
constexpr uint32_t TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;
constexpr uint32_t TC_TMEM_COLS = TC_ACC_STAGES * TC_BLOCK_N;  // 512
constexpr int TC_EPI_WARPS = 4;
constexpr uint32_t TC_OUT_BUF_BYTES = 32 * 128;  // 32 rows x 128 B: one warp's slice of a store, SWIZZLE_128B
constexpr uint32_t TC_OUT_BYTES = TC_EPI_WARPS * 2 * TC_OUT_BUF_BYTES;  // double-buffered per epilogue warp
constexpr size_t TC_SMEM_BYTES =
    1024 /*align slack*/ + (size_t)TC_STAGES * TC_STAGE_BYTES + TC_OUT_BYTES + 256 /*barriers*/;

// MODE 0: calibrated logits.  MODE 1: logits + per-row (max, sum-exp) of the tile's columns - the logsumexp over classes
// that log_prob needs, so that it can be formed from S*B gathers instead of a second pass over [S,B,C].  MODE 2: the
// row is complete inside the tile (C <= 256): two passes over the TMEM accumulator - statistics, then softmax /
// log_softmax written instead of the logits (prob_output_layer/classification.py:25-28, 54-69).  The exponentials
// (one MUFU.EX2 per element) hide behind the MMAs of the next tile, which run on the other accumulator stage.
template <typename TO, bool TMA_OUT, int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1)
    last_layer_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ CUtensorMap tmap_out, const TcParams p) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* stage_base = smem;
  unsigned char* out_stage = smem + (size_t)TC_STAGES * TC_STAGE_BYTES;  // 1024-aligned (stages are 48 KB)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(out_stage + TC_OUT_BYTES);
  uint64_t* empty_bar = full_bar + TC_STAGES;
  uint64_t* tmem_full_bar = empty_bar + TC_STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + TC_ACC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + TC_ACC_STAGES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tmap_prefetch(&tmap_a);
    tmap_prefetch(&tmap_b);
    if (TMA_OUT) tmap_prefetch(&tmap_out);
    for (int i = 0; i < TC_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < TC_ACC_STAGES; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 4);  // one arrival per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TC_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int group = gridDim.x;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const TileCoord tc = tile_coord(t, group, p);
        const int sw = p.sample_idx ? __ldg(p.sample_idx + tc.s) : tc.s;  // posterior.sample() gather, fused
        for (int kb = 0; kb < p.k_blocks; ++kb, ++it) {
          const int stage = it % TC_STAGES;
          const uint32_t round = it / TC_STAGES;
          if (round > 0) mbar_wait(&empty_bar[stage], (round - 1) & 1u);
          unsigned char* sa = stage_base + (size_t)stage * TC_STAGE_BYTES;
          unsigned char* sb = sa + TC_A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], TC_STAGE_BYTES);
          tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * TC_BLOCK_K, tc.m_blk * TC_BLOCK_M);
          tma_load_3d(sb, &tmap_b, &full_bar[stage], kb * TC_BLOCK_K, tc.n_blk * TC_BLOCK_N, sw);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(TC_BLOCK_M, TC_BLOCK_N);
      uint32_t it = 0, tile_it = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++tile_it) {
        const int acc = tile_it % TC_ACC_STAGES;
        const uint32_t acc_round = tile_it / TC_ACC_STAGES;
        if (acc_round > 0) mbar_wait(&tmem_empty_bar[acc], (acc_round - 1) & 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * TC_BLOCK_N;
        for (int kb = 0; kb < p.k_blocks; ++kb, ++it) {
          const int stage = it % TC_STAGES;
          const uint32_t round = it / TC_STAGES;
          mbar_wait(&full_bar[stage], round & 1u);
          tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + (size_t)stage * TC_STAGE_BYTES);
          const uint32_t sb = sa + TC_A_BYTES;
#pragma unroll
          for (int k = 0; k < TC_BLOCK_K / TC_UMMA_K; ++k) {
            const uint64_t da = umma_smem_desc(sa + k * TC_UMMA_K * 2);
            const uint64_t db = umma_smem_desc(sb + k * TC_UMMA_K * 2);
            umma_bf16(d_tmem, da, db, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);                          // smem slot free once these MMAs retire
          if (kb == p.k_blocks - 1) umma_commit(&tmem_full_bar[acc]);  // accumulator complete
        }
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------- epilogue warps (2..5)
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const float scale = p.log_temp ? expf(-p.log_temp[0]) : 1.0f;
    TO* out = reinterpret_cast<TO*>(p.out);
    constexpr int CPS = 128 / (int)sizeof(TO);  // columns per TMA store: one 128-byte swizzle row per output row
    unsigned char* my_stage = out_stage + (size_t)(warp - 2) * 2 * TC_OUT_BUF_BYTES;
    uint32_t store_it = 0;
    uint32_t tile_it = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++tile_it) {
      const TileCoord tc = tile_coord(t, group, p);
      const int acc = tile_it % TC_ACC_STAGES;
      const uint32_t acc_round = tile_it / TC_ACC_STAGES;
      mbar_wait(&tmem_full_bar[acc], acc_round & 1u);
      tc_fence_after();
      const int row0 = tc.m_blk * TC_BLOCK_M + quad * 32;
      const int b = row0 + lane;
      const int c_base = tc.n_blk * TC_BLOCK_N;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)acc * TC_BLOCK_N;
      const int sw = p.sample_idx ? __ldg(p.sample_idx + tc.s) : tc.s;
      const float* bias_row = p.bias ? p.bias + (size_t)sw * p.C : nullptr;
      TO* out_row = out + ((size_t)tc.s * p.B + (b < p.B ? b : 0)) * p.C;
      constexpr float kLog2e = 1.4426950408889634f;
      float run_m = -INFINITY, run_s = 0.0f;  // MODE != 0: running row max / sum exp(v - max) over this tile's columns
      // accumulator chunk -> calibrated logits v[32] of this thread's row (columns c0 .. c0+31)
      auto load_chunk = [&](int chunk, int c0, float (&v)[32]) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + (uint32_t)chunk * 32, r);
        // bias for these 32 columns (same addresses in every lane: broadcast loads), overlapped with the TMEM load
        float bv[32];
        if (bias_row) {
          if (c0 + 32 <= p.C && ((reinterpret_cast<uintptr_t>(bias_row + c0) & 15u) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 q = __ldg(reinterpret_cast<const float4*>(bias_row + c0 + j));
              bv[j] = q.x; bv[j + 1] = q.y; bv[j + 2] = q.z; bv[j + 3] = q.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) bv[j] = (c0 + j < p.C) ? __ldg(bias_row + c0 + j) : 0.0f;
          }
        }
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] = __uint_as_float(r[j]);
          if (bias_row) v[j] += bv[j];
          if (p.log_temp) v[j] *= scale;
          if (MODE != 0 && c0 + 32 > p.C && c0 + j >= p.C) v[j] = -INFINITY;  // columns past C: exp -> 0
        }
      };
      auto add_stats = [&](const float (&v)[32]) {
        float cm = v[0];
#pragma unroll
        for (int j = 1; j < 32; ++j) cm = fmaxf(cm, v[j]);
        const float m_new = fmaxf(run_m, cm);
        const float mb = m_new * kLog2e;
        float acc = 0.0f;
#pragma unroll
        for (int j = 0; j < 32; ++j) acc += exp2f(fmaf(v[j], kLog2e, -mb));
        run_s = run_s * exp2f(fmaf(run_m, kLog2e, -mb)) + acc;
        run_m = m_new;
      };
      auto store_chunk = [&](int chunk, int c0, float (&v)[32]) {
        if (TMA_OUT) {
            // stage this warp's 32 rows x 32 columns in shared memory (128-byte rows, SWIZZLE_128B: 16-byte unit q of
            // row r lives at unit q ^ (r & 7), so the 8 lanes of a store phase hit 8 different bank groups), then one
            // lane hands the block to the TMA engine: full-line coalesced global writes, off the LSU path.
            constexpr int per_store = CPS / 32;              // TMEM chunks per store: 1 (f32) or 2 (bf16)
            const int sub = chunk % per_store;
            unsigned char* buf = my_stage + (size_t)(store_it & 1u) * TC_OUT_BUF_BYTES;
            if (sub == 0) {
              if (lane == 0) bulk_wait_group_read<1>();      // the store that last read `buf` has drained it
              __syncwarp();
            }
            unsigned char* rowp = buf + lane * 128;
            if (sizeof(TO) == 4) {
#pragma unroll
              for (int q = 0; q < 8; ++q)
                *reinterpret_cast<float4*>(rowp + ((q ^ (lane & 7)) << 4)) =
                    make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            } else {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                uint4 pk;
                __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * q], v[8 * q + 1]);
                __nv_bfloat162 h1 = __floats2bfloat162_rn(v[8 * q + 2], v[8 * q + 3]);
                __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * q + 4], v[8 * q + 5]);
                __nv_bfloat162 h3 = __floats2bfloat162_rn(v[8 * q + 6], v[8 * q + 7]);
                pk.x = *reinterpret_cast<uint32_t*>(&h0);
                pk.y = *reinterpret_cast<uint32_t*>(&h1);
                pk.z = *reinterpret_cast<uint32_t*>(&h2);
                pk.w = *reinterpret_cast<uint32_t*>(&h3);
                *reinterpret_cast<uint4*>(rowp + (((sub * 4 + q) ^ (lane & 7)) << 4)) = pk;
              }
            }
            const bool last_sub = (sub == per_store - 1) || (c0 + 32 >= p.C) || (chunk == TC_BLOCK_N / 32 - 1);
            if (last_sub) {
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_3d(&tmap_out, buf, c0 - sub * 32, row0, tc.s);
                bulk_commit_group();
              }
              ++store_it;
            }
          } else if (b < p.B) {
            if (p.c_split) {
              // flattened N: scatter column n = (sample, class) to out[sample, b, class]; for a fixed column the 32 lanes
              // (consecutive b) write addresses c_split elements apart, neighbouring columns fill the gaps
              int sn = c0 / p.c_split, cn = c0 - sn * p.c_split;
              TO* o = out + ((size_t)sn * p.B + b) * p.c_split;
              const size_t sample_stride = (size_t)p.B * p.c_split;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (c0 + j < p.C) {
                  if (sizeof(TO) == 4) reinterpret_cast<float*>(o)[cn] = v[j];
                  else reinterpret_cast<__nv_bfloat16*>(o)[cn] = __float2bfloat16_rn(v[j]);
                }
                if (++cn == p.c_split) {
                  cn = 0;
                  o += sample_stride;
                }
              }
            } else if (sizeof(TO) == 4) {
              float* o = reinterpret_cast<float*>(out_row) + c0;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (c0 + j < p.C) o[j] = v[j];
            } else {
              __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out_row) + c0;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (c0 + j < p.C) o[j] = __float2bfloat16_rn(v[j]);
            }
          }
            };
      float lse = 0.0f;
      if (MODE == 2) {
#pragma unroll 1
        for (int chunk = 0; chunk < TC_BLOCK_N / 32; ++chunk) {
          const int c0 = c_base + chunk * 32;
          if (c0 >= p.C) break;  // warp-uniform
          float v[32];
          load_chunk(chunk, c0, v);
          add_stats(v);
        }
        lse = run_m + __logf(run_s);
      }
#pragma unroll 1
      for (int chunk = 0; chunk < TC_BLOCK_N / 32; ++chunk) {
        const int c0 = c_base + chunk * 32;
        if (c0 >= p.C) break;  // warp-uniform
        float v[32];
        load_chunk(chunk, c0, v);
        if (MODE == 1) add_stats(v);
        if (MODE == 2) {
          if (p.epi == 1) {
            const float lb = lse * kLog2e;
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = exp2f(fmaf(v[j], kLog2e, -lb));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = v[j] - lse;
          }
        }
        store_chunk(chunk, c0, v);
      }
      if (MODE != 0 && b < p.B) {
        const size_t row = (size_t)tc.s * p.B + b;
        if (p.lse_part) p.lse_part[row * p.n_blocks + tc.n_blk] = make_float2(run_m, run_s);
        else if (p.lse_out) p.lse_out[row] = run_m + __logf(run_s);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
    }
    if (TMA_OUT && lane == 0) bulk_wait_group_all();  // shared memory must outlive the last store's read
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TC_TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------- host
int last_layer_tc_supported(int S, int B, int D, int C, int in_dtype, int out_dtype, int w_layout) {
  (void)S;
  (void)B;
  (void)out_dtype;
  if (in_dtype != FB200_BF16 || w_layout != FB200_W_SCD) return 0;
  if (D % 8 != 0 || D < TC_BLOCK_K) return 0;  // TMA needs 16-byte global strides
  if (C < 8) return 0;
  return 1;
}

// c_split > 0: the caller folded the sample axis into the column axis (S = 1, C = samples * c_split, w viewed as
// [1, C, D], bias as [1, C]); the epilogue scatters column n to out[n / c_split, b, n % c_split].
// [rows, n_blocks] (max, sum exp) partials of the column tiles -> lse[rows] = logsumexp over the whole row
__global__ void __launch_bounds__(256) lse_combine_kernel(const float2* __restrict__ part, int64_t rows, int n_blocks,
                                                          float* __restrict__ lse) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += stride) {
    float m = -INFINITY;
    for (int n = 0; n < n_blocks; ++n) m = fmaxf(m, part[r * n_blocks + n].x);
    float s = 0.0f;
    for (int n = 0; n < n_blocks; ++n) {
      const float2 q = part[r * n_blocks + n];
      s += q.y * __expf(q.x - m);
    }
    lse[r] = m + __logf(s);
  }
}

size_t last_layer_tc_lse_workspace_bytes(int S, int B, int C) {
  const int n_blocks = (C + TC_BLOCK_N - 1) / TC_BLOCK_N;
  return n_blocks > 1 ? (size_t)S * (size_t)B * (size_t)n_blocks * sizeof(float2) : 0;
}

// epi: 0 logits / 1 softmax / 2 log_softmax (1, 2: C <= 256); lse_out: optional [S, B]; lse_ws: workspace of
// last_layer_tc_lse_workspace_bytes when C > 256 and lse_out is given.
int last_layer_tc_launch(const void* x, const void* w, const float* bias, const float* log_temp,
                         const int32_t* sample_idx, int n_stored, void* out, int S, int B, int D, int C,
                         int out_dtype, int c_split, cudaStream_t st, int epi, float* lse_out, void* lse_ws) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return FB200_E_UNSUPPORTED;
  if (!fb200_aligned16(x) || !fb200_aligned16(w)) return FB200_E_ALIGN;
  CUtensorMap ta, tb;
  {
    cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)B};
    cuuint64_t strides[1] = {(cuuint64_t)D * 2};
    cuuint32_t box[2] = {TC_BLOCK_K, TC_BLOCK_M};
    cuuint32_t el[2] = {1, 1};
    CUresult r = enc(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(x), dims, strides, box, el,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return FB200_E_UNSUPPORTED;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)C, (cuuint64_t)n_stored};
    cuuint64_t strides[2] = {(cuuint64_t)D * 2, (cuuint64_t)D * 2 * (cuuint64_t)C};
    cuuint32_t box[3] = {TC_BLOCK_K, TC_BLOCK_N, 1};
    cuuint32_t el[3] = {1, 1, 1};
    CUresult r = enc(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w), dims, strides, box, el,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return FB200_E_UNSUPPORTED;
  }
  // output map {C, B, S}: needs 16-byte row stride and base; otherwise the epilogue stores directly
  const size_t osz = out_dtype == FB200_F32 ? 4 : 2;
  const bool tma_out = c_split == 0 && fb200_aligned16(out) && (((size_t)C * osz) % 16 == 0);
  CUtensorMap to = ta;
  if (tma_out) {
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)B, (cuuint64_t)S};
    cuuint64_t strides[2] = {(cuuint64_t)C * osz, (cuuint64_t)C * osz * (cuuint64_t)B};
    cuuint32_t box[3] = {(cuuint32_t)(128 / osz), 32, 1};
    cuuint32_t el[3] = {1, 1, 1};
    CUresult r = enc(&to, out_dtype == FB200_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                     out, dims, strides, box, el, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return FB200_E_UNSUPPORTED;
  }
  TcParams p{};
  p.bias = bias;
  p.log_temp = log_temp;
  p.sample_idx = sample_idx;
  p.c_split = c_split;
  p.out = out;
  p.S = S;
  p.B = B;
  p.D = D;
  p.C = C;
  p.m_blocks = (B + TC_BLOCK_M - 1) / TC_BLOCK_M;
  p.n_blocks = (C + TC_BLOCK_N - 1) / TC_BLOCK_N;
  p.k_blocks = (D + TC_BLOCK_K - 1) / TC_BLOCK_K;
  const long long total = (long long)p.m_blocks * p.n_blocks * S;
  if (total > 0x7FFFFFFFll) return FB200_E_UNSUPPORTED;
  p.total_tiles = (int)total;
  p.epi = epi;
  int mode = 0;
  if (epi != 0 || lse_out) {
    if (c_split != 0) return FB200_E_UNSUPPORTED;
    if (epi != 0 && p.n_blocks != 1) return FB200_E_UNSUPPORTED;
    mode = epi != 0 ? 2 : 1;
    if (p.n_blocks == 1) {
      p.lse_out = lse_out;
    } else {
      if (!lse_ws) return FB200_E_WORKSPACE;
      p.lse_part = reinterpret_cast<float2*>(lse_ws);
    }
  }
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int grid = sms;
  if (grid > p.total_tiles) grid = p.total_tiles;
  auto launch = [&](auto kern) -> int {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    kern<<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(ta, tb, to, p);
    return 0;
  };
  int rc;
#define FB200_TC_DISPATCH(M)                                                                                   \
  if (out_dtype == FB200_F32)                                                                                  \
    rc = tma_out ? launch(last_layer_tc_kernel<float, true, M>) : launch(last_layer_tc_kernel<float, false, M>); \
  else                                                                                                         \
    rc = tma_out ? launch(last_layer_tc_kernel<__nv_bfloat16, true, M>)                                        \
                 : launch(last_layer_tc_kernel<__nv_bfloat16, false, M>);
  if (mode == 0) {
    FB200_TC_DISPATCH(0)
  } else if (mode == 1) {
    FB200_TC_DISPATCH(1)
  } else {
    FB200_TC_DISPATCH(2)
  }
#undef FB200_TC_DISPATCH
  if (rc != 0) return rc;
  FB200_LAUNCH_CHECK();
  if (p.lse_part) {
    const int64_t rows = (int64_t)S * B;
    int64_t blocks = (rows + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    lse_combine_kernel<<<(unsigned)blocks, 256, 0, st>>>(p.lse_part, rows, p.n_blocks, lse_out);
    FB200_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace fb200
