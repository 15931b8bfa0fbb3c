// K2, CTA-pair variant: the S-batched last-layer GEMM with tcgen05.mma.cta_group::2 (sm_100a).
//
// Same contract as last_layer_tc.cu, different tiling: a CLUSTER of two CTAs (one SM pair / TPC) owns a 256 x 256
// output tile.  CTA r of the pair loads its own 128 rows of x and its own 128 rows of W (half of the N tile) per
// 64-wide K slab - 32 KB per stage instead of 48 KB - and the pair's leader issues ONE M256 N256 K16 UMMA that reads
// both CTAs' shared memory and writes each CTA's 128 x 256 accumulator into that CTA's own TMEM.  Per output flop
// this moves 2/3 of the L2->SM bytes and 2/3 of the shared-memory operand reads of the 1-CTA kernel, which is what
// bounded it (profiles/r01d_tc_ncu_full.txt: tensor pipe 74 %, l1tex/smem 72 %).
//
// Synchronisation (all mbarriers live in each CTA's own shared memory at identical offsets):
//   full[stage]   (leader's copy is used)  : both CTAs' TMA loads complete_tx on the LEADER's barrier
//                                            (.cta_group::2, barrier address with the peer bit cleared); the leader's
//                                            producer arms it with the bytes of BOTH CTAs
//   empty[stage]  (one per CTA)            : tcgen05.commit ... multicast to both CTAs when the MMAs that read the
//                                            stage have retired -> each CTA's producer may refill its half
//   tmem_full[acc] (one per CTA)           : tcgen05.commit multicast when a tile's accumulators are complete
//   tmem_empty[acc] (leader's copy)        : 8 arrivals = 4 epilogue warps x 2 CTAs (the peer arrives remotely)
#include <cooperative_groups.h>

#include "tc_common.cuh"

namespace fb200 {

constexpr int T2_BLOCK_M = 128;  // rows per CTA (256 per pair)
constexpr int T2_BLOCK_N = 256;  // columns per pair; each CTA stages 128 of them
constexpr int T2_BLOCK_K = 64;
constexpr int T2_UMMA_K = 16;
constexpr int T2_STAGES = 6;
constexpr int T2_ACC_STAGES = 2;
constexpr int T2_THREADS = 192;
constexpr uint32_t T2_A_BYTES = T2_BLOCK_M * T2_BLOCK_K * 2;        // 16 KB
constexpr uint32_t T2_B_BYTES = (T2_BLOCK_N / 2) * T2_BLOCK_K * 2;  // 16 KB
constexpr uint32_t T2_STAGE_BYTES = T2_A_BYTES + T2_B_BYTES;        // 32 KB
constexpr uint32_t T2_TMEM_COLS = T2_ACC_STAGES * T2_BLOCK_N;       // 512
constexpr uint32_t T2_OUT_BUF_BYTES = 32 * 128;
constexpr uint32_t T2_OUT_BYTES = 4 * 2 * T2_OUT_BUF_BYTES;
constexpr size_t T2_SMEM_BYTES = 1024 + (size_t)T2_STAGES * T2_STAGE_BYTES + T2_OUT_BYTES + 256;

// ---------------------------------------------------------------------------------------------- PTX (pair)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Loads into this CTA's shared memory, signals the LEADER CTA's mbarrier (peer bit of the barrier address cleared).
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma2_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem2_alloc(uint32_t* slot_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem2_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs of the pair once the issued MMAs have completed
__device__ __forceinline__ void umma2_commit_both(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}
// wait on a local mbarrier whose arrivals come from the peer CTA too: acquire at cluster scope
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
// arrive on the mbarrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

template <typename TO, bool TMA_OUT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T2_THREADS, 1)
    last_layer_tc2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                          const __grid_constant__ CUtensorMap tmap_out, const TcParams p) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* stage_base = smem;
  unsigned char* out_stage = smem + (size_t)T2_STAGES * T2_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(out_stage + T2_OUT_BYTES);
  uint64_t* empty_bar = full_bar + T2_STAGES;
  uint64_t* tmem_full_bar = empty_bar + T2_STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + T2_ACC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + T2_ACC_STAGES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader of the pair
  const int pair = blockIdx.x >> 1;
  const int n_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tmap_prefetch(&tmap_a);
    tmap_prefetch(&tmap_b);
    if (TMA_OUT) tmap_prefetch(&tmap_out);
    for (int i = 0; i < T2_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);   // leader's producer arms it; TMA of both CTAs completes the bytes
      mbar_init(&empty_bar[i], 1);  // one multicast commit
    }
    for (int i = 0; i < T2_ACC_STAGES; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 8);  // 4 epilogue warps of each CTA
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem2_alloc(tmem_slot, T2_TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();  // both CTAs' barriers are initialised and TMEM is allocated before anyone signals across
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer (both CTAs, own halves)
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = pair; t < p.total_tiles; t += n_pairs) {
        const TileCoord tc = tile_coord(t, n_pairs, p);
        const int sw = p.sample_idx ? __ldg(p.sample_idx + tc.s) : tc.s;
        const int row_a = tc.m_blk * (2 * T2_BLOCK_M) + (int)rank * T2_BLOCK_M;
        const int row_b = tc.n_blk * T2_BLOCK_N + (int)rank * (T2_BLOCK_N / 2);
        for (int kb = 0; kb < p.k_blocks; ++kb, ++it) {
          const int stage = it % T2_STAGES;
          const uint32_t round = it / T2_STAGES;
          if (round > 0) mbar_wait(&empty_bar[stage], (round - 1) & 1u);
          unsigned char* sa = stage_base + (size_t)stage * T2_STAGE_BYTES;
          unsigned char* sb = sa + T2_A_BYTES;
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * T2_STAGE_BYTES);
          tma2_load_2d(sa, &tmap_a, &full_bar[stage], kb * T2_BLOCK_K, row_a);
          tma2_load_3d(sb, &tmap_b, &full_bar[stage], kb * T2_BLOCK_K, row_b, sw);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (leader CTA only)
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * T2_BLOCK_M, T2_BLOCK_N);
      uint32_t it = 0, tile_it = 0;
      for (int t = pair; t < p.total_tiles; t += n_pairs, ++tile_it) {
        const int acc = tile_it % T2_ACC_STAGES;
        const uint32_t acc_round = tile_it / T2_ACC_STAGES;
        if (acc_round > 0) mbar_wait_cluster(&tmem_empty_bar[acc], (acc_round - 1) & 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * T2_BLOCK_N;
        for (int kb = 0; kb < p.k_blocks; ++kb, ++it) {
          const int stage = it % T2_STAGES;
          const uint32_t round = it / T2_STAGES;
          mbar_wait(&full_bar[stage], round & 1u);
          tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + (size_t)stage * T2_STAGE_BYTES);
          const uint32_t sb = sa + T2_A_BYTES;
#pragma unroll
          for (int k = 0; k < T2_BLOCK_K / T2_UMMA_K; ++k) {
            const uint64_t da = umma_smem_desc(sa + k * T2_UMMA_K * 2);
            const uint64_t db = umma_smem_desc(sb + k * T2_UMMA_K * 2);
            umma2_bf16(d_tmem, da, db, idesc, (kb | k) != 0);
          }
          umma2_commit_both(&empty_bar[stage]);
          if (kb == p.k_blocks - 1) umma2_commit_both(&tmem_full_bar[acc]);
        }
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------- epilogue warps (2..5), own 128 rows
    const int quad = warp & 3;
    const float scale = p.log_temp ? expf(-p.log_temp[0]) : 1.0f;
    TO* out = reinterpret_cast<TO*>(p.out);
    constexpr int CPS = 128 / (int)sizeof(TO);
    unsigned char* my_stage = out_stage + (size_t)(warp - 2) * 2 * T2_OUT_BUF_BYTES;
    uint32_t store_it = 0;
    uint32_t tile_it = 0;
    for (int t = pair; t < p.total_tiles; t += n_pairs, ++tile_it) {
      const TileCoord tc = tile_coord(t, n_pairs, p);
      const int acc = tile_it % T2_ACC_STAGES;
      const uint32_t acc_round = tile_it / T2_ACC_STAGES;
      mbar_wait(&tmem_full_bar[acc], acc_round & 1u);
      tc_fence_after();
      const int row0 = tc.m_blk * (2 * T2_BLOCK_M) + (int)rank * T2_BLOCK_M + quad * 32;
      const int b = row0 + lane;
      const int c_base = tc.n_blk * T2_BLOCK_N;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)acc * T2_BLOCK_N;
      const int sw = p.sample_idx ? __ldg(p.sample_idx + tc.s) : tc.s;
      const float* bias_row = p.bias ? p.bias + (size_t)sw * p.C : nullptr;
      TO* out_row = out + ((size_t)tc.s * p.B + (b < p.B ? b : 0)) * p.C;
#pragma unroll 1
      for (int chunk = 0; chunk < T2_BLOCK_N / 32; ++chunk) {
        const int c0 = c_base + chunk * 32;
        if (c0 >= p.C) break;
        uint32_t r[32];
        tmem_ld_32x32(taddr + (uint32_t)chunk * 32, r);
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
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] = __uint_as_float(r[j]);
          if (bias_row) v[j] += bv[j];
          if (p.log_temp) v[j] *= scale;
        }
        if (TMA_OUT) {
          constexpr int per_store = CPS / 32;
          const int sub = chunk % per_store;
          unsigned char* buf = my_stage + (size_t)(store_it & 1u) * T2_OUT_BUF_BYTES;
          if (sub == 0) {
            if (lane == 0) bulk_wait_group_read<1>();
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
          const bool last_sub = (sub == per_store - 1) || (c0 + 32 >= p.C) || (chunk == T2_BLOCK_N / 32 - 1);
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
          if (sizeof(TO) == 4) {
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
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tmem_empty_bar[acc], 0);  // the leader's MMA warp waits for all 8 warps
    }
    if (TMA_OUT && lane == 0) bulk_wait_group_all();
  }

  // neither CTA may release TMEM (or exit, invalidating barriers the peer still signals) before both are done
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem2_dealloc(tmem_base, T2_TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------- host
int last_layer_tc2_launch(const void* x, const void* w, const float* bias, const float* log_temp,
                          const int32_t* sample_idx, int n_stored, void* out, int S, int B, int D, int C,
                          int out_dtype, cudaStream_t st) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return FB200_E_UNSUPPORTED;
  if (!fb200_aligned16(x) || !fb200_aligned16(w)) return FB200_E_ALIGN;
  CUtensorMap ta, tb;
  {
    cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)B};
    cuuint64_t strides[1] = {(cuuint64_t)D * 2};
    cuuint32_t box[2] = {T2_BLOCK_K, T2_BLOCK_M};
    cuuint32_t el[2] = {1, 1};
    if (enc(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(x), dims, strides, box, el,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return FB200_E_UNSUPPORTED;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)C, (cuuint64_t)n_stored};
    cuuint64_t strides[2] = {(cuuint64_t)D * 2, (cuuint64_t)D * 2 * (cuuint64_t)C};
    cuuint32_t box[3] = {T2_BLOCK_K, T2_BLOCK_N / 2, 1};
    cuuint32_t el[3] = {1, 1, 1};
    if (enc(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w), dims, strides, box, el,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return FB200_E_UNSUPPORTED;
  }
  const size_t osz = out_dtype == FB200_F32 ? 4 : 2;
  const bool tma_out = fb200_aligned16(out) && (((size_t)C * osz) % 16 == 0);
  CUtensorMap to = ta;
  if (tma_out) {
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)B, (cuuint64_t)S};
    cuuint64_t strides[2] = {(cuuint64_t)C * osz, (cuuint64_t)C * osz * (cuuint64_t)B};
    cuuint32_t box[3] = {(cuuint32_t)(128 / osz), 32, 1};
    cuuint32_t el[3] = {1, 1, 1};
    if (enc(&to, out_dtype == FB200_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, out,
            dims, strides, box, el, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return FB200_E_UNSUPPORTED;
  }
  TcParams p{};
  p.bias = bias;
  p.log_temp = log_temp;
  p.sample_idx = sample_idx;
  p.out = out;
  p.S = S;
  p.B = B;
  p.D = D;
  p.C = C;
  p.m_blocks = (B + 2 * T2_BLOCK_M - 1) / (2 * T2_BLOCK_M);  // pair tiles along M
  p.n_blocks = (C + T2_BLOCK_N - 1) / T2_BLOCK_N;
  p.k_blocks = (D + T2_BLOCK_K - 1) / T2_BLOCK_K;
  const long long total = (long long)p.m_blocks * p.n_blocks * S;
  if (total > 0x7FFFFFFFll) return FB200_E_UNSUPPORTED;
  p.total_tiles = (int)total;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int pairs = sms / 2;
  if (pairs > p.total_tiles) pairs = p.total_tiles;
  const int grid = 2 * pairs;
  auto launch = [&](auto kern) -> int {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T2_SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    kern<<<grid, T2_THREADS, T2_SMEM_BYTES, st>>>(ta, tb, to, p);
    return 0;
  };
  int rc;
  if (out_dtype == FB200_F32)
    rc = tma_out ? launch(last_layer_tc2_kernel<float, true>) : launch(last_layer_tc2_kernel<float, false>);
  else
    rc = tma_out ? launch(last_layer_tc2_kernel<__nv_bfloat16, true>)
                 : launch(last_layer_tc2_kernel<__nv_bfloat16, false>);
  if (rc != 0) return rc;
  FB200_LAUNCH_CHECK();
  return 0;
}

}  // namespace fb200
