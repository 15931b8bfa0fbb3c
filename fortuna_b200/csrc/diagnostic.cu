// SG-MCMC diagnostics over the sample bank.
//
// Replaces (reference file:line)
//   fortuna/prob_model/posterior/sgmcmc/sgmcmc_diagnostic.py:16-70    kernel_stein_discrepancy_imq
//   fortuna/prob_model/posterior/sgmcmc/sgmcmc_diagnostic.py:73-140   effective_sample_size
//
// KSD.  The reference evaluates k0(i, j) for all S^2 pairs, each from four length-N dot products
//   dd = <d, d>, gg = <g_i, g_j>, gid = <g_i, d>, gjd = <g_j, d>,  d = x_i - x_j
// (vmap over j inside a lax.scan over i).  dd / gg are symmetric and gid(j, i) = -gjd(i, j), so only sample tiles
// ti <= tj are accumulated: a 64 x 64 pair tile per CTA (32 x 32 for chains of <= 32 samples), 4 x 4 (2 x 2) pairs x 4 sums per thread in registers, the samples'
// columns staged through shared memory 32 at a time (LDGSTS, double-buffered); N is split over CTAs (each split writes its own partial, summed
// in split order - no float atomics, run-to-run identical).  Differences are formed explicitly: the Gram-matrix
// shortcut (|x_i|^2 + |x_j|^2 - 2 <x_i, x_j> on tensor cores) cancels catastrophically for SG-MCMC samples, which sit
// close together relative to their norms.  A second kernel turns the sums into k0 in double precision and adds them
// in the reference's order (row by row).
//
// ESS.  jnp.fft-based autocorrelation of the zero-padded, centred chain == the direct lagged products
//   prod[k] = sum_t x[t] x[t+k]   (padding to >= 2S makes the circular correlation linear),
// so each column's S lags are computed directly: a CTA holds a [S, COLS] slab of the chain in shared memory, a thread
// owns one column and 8 consecutive lags (a sliding register window: 2 shared loads per 8 FMAs), then one thread per
// column walks the lags in order for the unbiased normalisation, the cut at the first auto-correlation below the
// threshold and the final sum.
#include <math.h>

#include "common.cuh"

namespace fb200 {

constexpr int KSD_KC = 32;  // columns per shared-memory stage
constexpr int KSD_THREADS = 256;

struct KsdArgs {
  const float* x;
  const float* g;
  int S;
  int64_t N, stride;
  int tile, n_tiles_side, n_splits;
  int64_t cols_per_split;
  float* partial;  // [n_splits][S][S][4]  (dd, gg, gid, gjd) for pairs in tiles ti <= tj
};

// 4-byte asynchronous copy global -> shared (LDGSTS); src_bytes = 0 writes zeros (rows / columns past the end)
__device__ __forceinline__ void cp_async4_zfill(float* dst_smem, const float* src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
  const int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}

template <int KSD_T>  // samples per tile side: 64 (4 x 4 pairs per thread) or 32 (2 x 2; chains of <= 32 samples)
__global__ void __launch_bounds__(KSD_THREADS) ksd_pair_sums_kernel(const KsdArgs a) {
  constexpr int TP = KSD_T / 16;
  constexpr int ROW = KSD_T + 4;                 // floats per staged column; rows 16-byte aligned
  constexpr int ARR = KSD_KC * ROW;              // one [column][sample] array
  extern __shared__ __align__(16) float ksd_smem[];  // [2 buffers][x_I | g_I | x_J | g_J][KSD_KC][ROW]
  // tile index -> (ti, tj), ti <= tj
  int t = blockIdx.x, ti = 0;
  while (t >= a.n_tiles_side - ti) {
    t -= a.n_tiles_side - ti;
    ++ti;
  }
  const int tj = ti + t;
  const int split = blockIdx.y;
  const int64_t n0 = (int64_t)split * a.cols_per_split;
  const int64_t n1 = min(a.N, n0 + a.cols_per_split);
  const int tid = threadIdx.x;
  const int pi = (tid / 16) * TP, pj = (tid % 16) * TP;  // this thread's TP x TP pairs inside the tile
  float dd[TP][TP], gg[TP][TP], gid[TP][TP], gjd[TP][TP];
#pragma unroll
  for (int i = 0; i < TP; ++i)
#pragma unroll
    for (int j = 0; j < TP; ++j) dd[i][j] = gg[i][j] = gid[i][j] = gjd[i][j] = 0.0f;

  // stage [KSD_T samples][32 columns] of both tiles, transposed to [column][sample]: a warp reads 32 consecutive
  // columns of one sample (one 128-byte line); the copies are asynchronous so stage k + 1 lands while k is consumed
  auto stage = [&](int buf, int64_t nb) {
    float* base = ksd_smem + (size_t)buf * 4 * ARR;
    for (int e = tid; e < KSD_T * KSD_KC; e += KSD_THREADS) {
      const int s = e / KSD_KC, c = e % KSD_KC;
      const int64_t col = nb + c;
      const int si = ti * KSD_T + s, sj = tj * KSD_T + s;
      const bool okc = col < n1;
      const bool oki = okc && si < a.S, okj = okc && sj < a.S;
      const int64_t oi = oki ? (int64_t)si * a.stride + col : 0, oj = okj ? (int64_t)sj * a.stride + col : 0;
      float* d = base + (size_t)c * ROW + s;
      cp_async4_zfill(d, a.x + oi, oki);
      cp_async4_zfill(d + ARR, a.g + oi, oki);
      cp_async4_zfill(d + 2 * ARR, a.x + oj, okj);
      cp_async4_zfill(d + 3 * ARR, a.g + oj, okj);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int buf = 0;
  if (n0 < n1) stage(0, n0);
  for (int64_t nb = n0; nb < n1; nb += KSD_KC, buf ^= 1) {
    if (nb + KSD_KC < n1) {
      stage(buf ^ 1, nb + KSD_KC);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const float* sxI = ksd_smem + (size_t)buf * 4 * ARR;
    const float* sgI = sxI + ARR;
    const float* sxJ = sxI + 2 * ARR;
    const float* sgJ = sxI + 3 * ARR;
#pragma unroll 4
    for (int c = 0; c < KSD_KC; ++c) {
      float xi[TP], gi[TP], xj[TP], gj[TP];
      if constexpr (TP == 4) {
        const float4 a4 = *reinterpret_cast<const float4*>(sxI + c * ROW + pi);
        const float4 b4 = *reinterpret_cast<const float4*>(sgI + c * ROW + pi);
        const float4 c4 = *reinterpret_cast<const float4*>(sxJ + c * ROW + pj);
        const float4 d4 = *reinterpret_cast<const float4*>(sgJ + c * ROW + pj);
        xi[0] = a4.x; xi[1] = a4.y; xi[2] = a4.z; xi[3] = a4.w;
        gi[0] = b4.x; gi[1] = b4.y; gi[2] = b4.z; gi[3] = b4.w;
        xj[0] = c4.x; xj[1] = c4.y; xj[2] = c4.z; xj[3] = c4.w;
        gj[0] = d4.x; gj[1] = d4.y; gj[2] = d4.z; gj[3] = d4.w;
      } else {
        const float2 a2 = *reinterpret_cast<const float2*>(sxI + c * ROW + pi);
        const float2 b2 = *reinterpret_cast<const float2*>(sgI + c * ROW + pi);
        const float2 c2 = *reinterpret_cast<const float2*>(sxJ + c * ROW + pj);
        const float2 d2 = *reinterpret_cast<const float2*>(sgJ + c * ROW + pj);
        xi[0] = a2.x; xi[1] = a2.y; gi[0] = b2.x; gi[1] = b2.y;
        xj[0] = c2.x; xj[1] = c2.y; gj[0] = d2.x; gj[1] = d2.y;
      }
#pragma unroll
      for (int i = 0; i < TP; ++i)
#pragma unroll
        for (int j = 0; j < TP; ++j) {
          const float d = xi[i] - xj[j];
          dd[i][j] = fmaf(d, d, dd[i][j]);
          gg[i][j] = fmaf(gi[i], gj[j], gg[i][j]);
          gid[i][j] = fmaf(gi[i], d, gid[i][j]);
          gjd[i][j] = fmaf(gj[j], d, gjd[i][j]);
        }
    }
    __syncthreads();  // everyone is done with `buf` before the next iteration's copies overwrite it
  }
  float4* out = reinterpret_cast<float4*>(a.partial) + (int64_t)split * a.S * a.S;
#pragma unroll
  for (int i = 0; i < TP; ++i)
#pragma unroll
    for (int j = 0; j < TP; ++j) {
      const int si = ti * KSD_T + pi + i, sj = tj * KSD_T + pj + j;
      if (si < a.S && sj < a.S) out[(int64_t)si * a.S + sj] = make_float4(dd[i][j], gg[i][j], gid[i][j], gjd[i][j]);
    }
}

// Chains of <= 32 samples (every BASELINE config's default: 10 - 30 stored samples) are ONE diagonal tile: the square
// kernel above computes all 32 x 32 pairs of it although only i <= j are needed (2.2 x the work at S = 30) with 2 x 2
// pairs per thread (8 shared loads per 20 FMA-pipe instructions).  Here the tile is cut into 8 x 8 blocks of 4 x 4 pairs
// and only the 36 blocks bi <= bj exist: a thread owns one block (four 128-bit shared loads per 80 FMA-pipe instructions
// per column), 36 threads cover one column, and the CTA's 7 groups of 36 threads take every seventh column of a stage.
constexpr int KSD_TRI_GROUPS = 7;    // 7 x 36 = 252 computing threads in a 256-thread CTA: 8 warps, so that two CTAs per SM put
constexpr int KSD_TRI_THREADS = 256; // 4 warps on every scheduler and each thread may hold 128 registers (9 warps: 96)
constexpr int KSD_TRI_KC = 16 * KSD_TRI_GROUPS;       // columns per stage (16 per thread between two CTA barriers)
constexpr int KSD_TRI_ROW = 36;                       // floats per staged column: 32 samples + 4 (16-byte rows, bank shift)

__global__ void __launch_bounds__(KSD_TRI_THREADS, 2) ksd_tri32_kernel(const KsdArgs a) {
  constexpr int ARR = KSD_TRI_KC * KSD_TRI_ROW;
  extern __shared__ __align__(16) float ksd_smem[];  // [2 buffers][x | g][KSD_TRI_KC][KSD_TRI_ROW]
  const int split = blockIdx.y;
  const int64_t n0 = (int64_t)split * a.cols_per_split;
  const int64_t n1 = min(a.N, n0 + a.cols_per_split);
  const int tid = threadIdx.x;
  const bool computes = tid < 36 * KSD_TRI_GROUPS;  // the last 4 threads only help with the copies
  const int grp = computes ? tid / 36 : 0;
  int b = tid % 36, bi = 0;  // block index -> (bi, bj), bi <= bj
  while (b >= 8 - bi) {
    b -= 8 - bi;
    ++bi;
  }
  const int bj = bi + b;
  float dd[4][4], gg[4][4], gid[4][4], gjd[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) dd[i][j] = gg[i][j] = gid[i][j] = gjd[i][j] = 0.0f;

  // staging: warp w copies samples w, w + 8, w + 16, w + 24; a lane takes columns lane, lane + 32, ... of the stage (one
  // 128-byte line per warp instruction, no index division)
  const int lane = tid & 31, warp = tid >> 5;
  auto stage = [&](int buf, int64_t nb) {
    float* base = ksd_smem + (size_t)buf * 2 * ARR;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int s = warp + 8 * q;
      const bool oks = s < a.S;
      const float* rx = a.x + (int64_t)(oks ? s : 0) * a.stride;
      const float* rg = a.g + (int64_t)(oks ? s : 0) * a.stride;
#pragma unroll
      for (int c = lane; c < KSD_TRI_KC; c += 32) {
        const int64_t col = nb + c;
        const bool ok = oks && col < n1;
        float* d = base + (size_t)c * KSD_TRI_ROW + s;
        cp_async4_zfill(d, rx + (ok ? col : 0), ok);
        cp_async4_zfill(d + ARR, rg + (ok ? col : 0), ok);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int buf = 0;
  if (n0 < n1) stage(0, n0);
  for (int64_t nb = n0; nb < n1; nb += KSD_TRI_KC, buf ^= 1) {
    if (nb + KSD_TRI_KC < n1) {
      stage(buf ^ 1, nb + KSD_TRI_KC);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const float* sx = ksd_smem + (size_t)buf * 2 * ARR;
    const float* sg = sx + ARR;
    // operands of the next column are loaded before the current one is consumed: with 4 - 5 warps per scheduler the
    // 128-bit shared loads at the top of each column were the largest stall (ncu: 16 % of all samples)
    const float* px = sx + grp * KSD_TRI_ROW;
    const float* pg = sg + grp * KSD_TRI_ROW;
    float4 a4 = *reinterpret_cast<const float4*>(px + 4 * bi), b4 = *reinterpret_cast<const float4*>(pg + 4 * bi);
    float4 c4 = *reinterpret_cast<const float4*>(px + 4 * bj), d4 = *reinterpret_cast<const float4*>(pg + 4 * bj);
#pragma unroll 1
    for (int c = computes ? grp : KSD_TRI_KC; c < KSD_TRI_KC; c += KSD_TRI_GROUPS) {
      const float xi[4] = {a4.x, a4.y, a4.z, a4.w}, gi[4] = {b4.x, b4.y, b4.z, b4.w};
      const float xj[4] = {c4.x, c4.y, c4.z, c4.w}, gj[4] = {d4.x, d4.y, d4.z, d4.w};
      if (c + KSD_TRI_GROUPS < KSD_TRI_KC) {
        px += KSD_TRI_GROUPS * KSD_TRI_ROW;
        pg += KSD_TRI_GROUPS * KSD_TRI_ROW;
        a4 = *reinterpret_cast<const float4*>(px + 4 * bi);
        b4 = *reinterpret_cast<const float4*>(pg + 4 * bi);
        c4 = *reinterpret_cast<const float4*>(px + 4 * bj);
        d4 = *reinterpret_cast<const float4*>(pg + 4 * bj);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float d = xi[i] - xj[j];
          dd[i][j] = fmaf(d, d, dd[i][j]);
          gg[i][j] = fmaf(gi[i], gj[j], gg[i][j]);
          gid[i][j] = fmaf(gi[i], d, gid[i][j]);
          gjd[i][j] = fmaf(gj[j], d, gjd[i][j]);
        }
    }
    __syncthreads();  // everyone is done with `buf` before the next iteration's copies overwrite it
  }
  // the groups hold partial sums over disjoint columns of the same pairs: add them in group order through shared
  // memory (deterministic)
  __syncthreads();
  float4* red = reinterpret_cast<float4*>(ksd_smem);  // [groups][36 blocks][16 pairs]
  if (computes) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        red[((size_t)grp * 36 + (tid % 36)) * 16 + i * 4 + j] = make_float4(dd[i][j], gg[i][j], gid[i][j], gjd[i][j]);
  }
  __syncthreads();
  float4* out = reinterpret_cast<float4*>(a.partial) + (int64_t)split * a.S * a.S;
  for (int e = tid; e < 36 * 16; e += KSD_TRI_THREADS) {
    const int blk = e / 16, ij = e % 16;
    float4 acc = red[(size_t)blk * 16 + ij];
    for (int g2 = 1; g2 < KSD_TRI_GROUPS; ++g2) {
      const float4 v = red[((size_t)g2 * 36 + blk) * 16 + ij];
      acc.x += v.x;
      acc.y += v.y;
      acc.z += v.z;
      acc.w += v.w;
    }
    int bb = blk, ri = 0;
    while (bb >= 8 - ri) {
      bb -= 8 - ri;
      ++ri;
    }
    const int si = 4 * ri + ij / 4, sj = 4 * (ri + bb) + ij % 4;
    if (si < a.S && sj < a.S) out[(int64_t)si * a.S + sj] = acc;
  }
}

// k0 of one pair from its four sums, the reference's expression term by term (sgmcmc_diagnostic.py:52-61).
__device__ __forceinline__ double ksd_k0(double dd, double gg, double gid, double gjd, double c2, double beta,
                                         double dim) {
  const double base = c2 + dd;
  const double b0 = pow(base, beta);  // base^beta; the two lower powers by division
  const double b1 = b0 / base;
  const double b2 = b1 / base;
  double kern = gg * b0;
  kern += -2.0 * beta * gid * b1;
  kern += 2.0 * beta * gjd * b1;
  kern += -2.0 * dim * beta * b1;
  kern += -4.0 * beta * (beta - 1.0) * b2 * dd;
  return kern;
}

// Row i of the pair matrix: sum_j k0(i, j) in double (one CTA per row), then the rows are added in order - the
// reference's lax.scan over i with jnp.sum over j inside (:65-70) - and ksd = sqrt(total) / S.
__global__ void __launch_bounds__(256) ksd_row_kernel(const float* __restrict__ partial, int S, int64_t N, int n_splits,
                                                      int KSD_T, float c, float beta,
                                                      double* __restrict__ row_sums) {
  __shared__ double red[8];
  const int i = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float4* p = reinterpret_cast<const float4*>(partial);
  const double c2 = (double)c * (double)c;
  double acc = 0.0;
  for (int j = tid; j < S; j += blockDim.x) {
    // pairs of tiles ti > tj live at the mirrored entry: gid(i,j) = -gjd(j,i), gjd(i,j) = -gid(j,i)
    const bool mirror = (i / KSD_T) > (j / KSD_T);
    const int64_t idx = mirror ? (int64_t)j * S + i : (int64_t)i * S + j;
    double dd = 0., gg = 0., gid = 0., gjd = 0.;
    for (int s = 0; s < n_splits; ++s) {
      const float4 v = p[(int64_t)s * S * S + idx];
      dd += (double)v.x;
      gg += (double)v.y;
      gid += (double)v.z;
      gjd += (double)v.w;
    }
    const double a_gid = mirror ? -gjd : gid, a_gjd = mirror ? -gid : gjd;
    acc += ksd_k0(dd, gg, a_gid, a_gjd, c2, beta, (double)N);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (tid == 0) {
    double r = 0.0;
    for (int w = 0; w < 8; ++w) r += red[w];
    row_sums[i] = r;
  }
}

__global__ void ksd_total_kernel(const double* __restrict__ row_sums, int S, float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double t = 0.0;
  for (int i = 0; i < S; ++i) t += row_sums[i];
  out[0] = (float)(sqrt(t) / (double)S);
}

// ------------------------------------------------------------------------------------------------ ESS
constexpr int ESS_R = 8;  // consecutive lags per thread (sliding register window)

struct EssArgs {
  const float* x;
  int S;
  int64_t N, stride;
  int use_threshold;
  float threshold;
  float* ess;
};

// Short-chain kernel: per-lag / per-row constants live in the kernel parameters (constant bank, warp-uniform operands):
// live[t] = 1 for t < S else 0 (centres real rows only, one FFMA per element), inv_len[k] = 1 / (S - k), wgt[k] = (S - k) / S.
struct EssShortArgs {
  EssArgs a;
  float inv_S_rn;  // RN(1 / S) for the correctly rounded mean (Markstein)
  float live[64], inv_len[64], wgt[64];
};

template <int COLS>
__global__ void __launch_bounds__(256) ess_kernel(const EssArgs a) {
  extern __shared__ __align__(16) float ess_smem[];
  const int S = a.S;
  float* xs = ess_smem;                 // [S + ESS_R][COLS] centred chain, zero rows past the end
  float* ac = xs + (size_t)(S + ESS_R) * COLS;  // [max(S, GROUPS)][COLS] lagged products (first: column-sum scratch)
  constexpr int GROUPS = 256 / COLS;
  float* inv_den = ac + (size_t)(S > GROUPS ? S : GROUPS) * COLS;  // [S] 1 / (S - k)
  for (int k = threadIdx.x; k < S; k += 256) inv_den[k] = 1.0f / ((float)S - (float)k);
  const int col = threadIdx.x % COLS, grp = threadIdx.x / COLS;
  const int64_t n_slabs = (a.N + COLS - 1) / COLS;
  for (int64_t slab = blockIdx.x; slab < n_slabs; slab += gridDim.x) {
    const int64_t n = slab * COLS + col;
    const bool ok = n < a.N;
    // load + mean (sgmcmc_diagnostic.py:107-108): each column is summed by its GROUPS threads, then centred
    float part = 0.0f;
    for (int t = grp; t < S; t += GROUPS) {
      const float v = ok ? a.x[(int64_t)t * a.stride + n] : 0.0f;
      xs[(size_t)t * COLS + col] = v;
      part += v;
    }
    for (int t = S + grp; t < S + ESS_R; t += GROUPS) xs[(size_t)t * COLS + col] = 0.0f;
    ac[(size_t)grp * COLS + col] = part;
    __syncthreads();
    float mean = 0.0f;
    for (int q = 0; q < GROUPS; ++q) mean += ac[(size_t)q * COLS + col];
    mean = mean / (float)S;
    __syncthreads();
    for (int t = grp; t < S; t += GROUPS) xs[(size_t)t * COLS + col] -= mean;
    __syncthreads();
    // lagged products, ESS_R lags at a time: acc[r] = sum_t x[t] * x[t + k0 + r]
    for (int k0 = grp * ESS_R; k0 < S; k0 += GROUPS * ESS_R) {
      float acc[ESS_R], win[ESS_R];
#pragma unroll
      for (int r = 0; r < ESS_R; ++r) {
        acc[r] = 0.0f;
        win[r] = xs[(size_t)(k0 + r) * COLS + col];  // x[0 + k0 + r]; rows past S are zero
      }
      const int t_end = S - k0;  // x[t + k0] is zero for t >= S - k0
      int t = 0;
      // eight time steps per iteration with the window kept as a ring: at step u the lag-r partner of x[t + u] is
      // win[(u + r) & 7], and slot u is then refilled with x[t + u + k0 + 8] - no register moves, one load per step
      for (; t + ESS_R <= t_end; t += ESS_R) {
#pragma unroll
        for (int u = 0; u < ESS_R; ++u) {
          const float xt = xs[(size_t)(t + u) * COLS + col];
#pragma unroll
          for (int r = 0; r < ESS_R; ++r) acc[r] = fmaf(xt, win[(u + r) & (ESS_R - 1)], acc[r]);
          win[u] = xs[(size_t)(t + u + k0 + ESS_R) * COLS + col];  // rows S .. S+7 are zero
        }
      }
      for (; t < t_end; ++t) {
        const float xt = xs[(size_t)t * COLS + col];
#pragma unroll
        for (int r = 0; r < ESS_R; ++r) acc[r] = fmaf(xt, win[r], acc[r]);
#pragma unroll
        for (int r = 0; r + 1 < ESS_R; ++r) win[r] = win[r + 1];
        const int nx = t + 1 + k0 + ESS_R - 1;
        win[ESS_R - 1] = nx < S ? xs[(size_t)nx * COLS + col] : 0.0f;
      }
#pragma unroll
      for (int r = 0; r < ESS_R; ++r)
        if (k0 + r < S) ac[(size_t)(k0 + r) * COLS + col] = acc[r];
    }
    __syncthreads();
    // one thread per column walks the lags in order (:120-139)
    if (grp == 0 && ok) {
      const float fS = (float)S;
      const float cov0 = ac[col] / fS;  // denominator S - 0
      const float inv_cov0 = 1.0f / cov0;
      const float inv_S = 1.0f / fS;
      float sum = 0.0f;
      bool open = true;
      for (int k = 0; k < S; ++k) {
        const float cov = ac[(size_t)k * COLS + col] * inv_den[k];  // reciprocals: 1-2 ulp from the true quotients
        const float rho = cov * inv_cov0;
        if (a.use_threshold && rho < a.threshold) open = false;  // cumulative mask: this lag and every later one drop out
        const float w = ((fS - (float)k) * inv_S) * rho;
        if (open) sum += w;
        else if (rho != rho) sum += rho;  // NaN auto-correlations propagate like `weighted * mask` does
      }
      a.ess[n] = fS / (-1.0f + 2.0f * sum);
    }
    __syncthreads();
  }
}

// Short chains (S <= SMAX = 32 / 64 stored samples - every BASELINE config): ONE THREAD PER PARAMETER, the whole chain
// in its registers.  A warp's load of sample t is 32 consecutive floats (one 128-byte line), all S loads of a thread are
// in flight together, the lagged products are S^2/2 register FMAs (entries past S are zero, so the unrolled triangle
// needs no bounds), and the lag walk follows in the same thread: no shared memory, no barrier, no per-slab latency chain
// (the slab kernel below reached 19 % of the HBM peak at c5: load -> centre -> lags -> walk, one after the other).
// With a cut-off (the default filter_threshold = 0) the lags stop as soon as all 32 chains of the warp are masked.
// Prefetch ring: a thread's S loads for its NEXT parameters are issued as 4-byte cp.async (LDGSTS) into a private column
// of a DEPTH-stage shared-memory buffer before it computes the current one, so the loads of DEPTH - 1 iterations are in
// flight under the lag arithmetic (registers alone gave 2 blocks per SM that alternate between waiting for 30 loads and
// computing: 44 % of the HBM peak).  Every thread reads back only what it copied itself: cp.async.wait_group is all the
// synchronisation there is.  4-byte copies: rows of a [S, N] matrix are only 4-byte aligned to each other in general.
template <int SMAX, int DEPTH>
__global__ void __launch_bounds__(256) ess_short_kernel(const __grid_constant__ EssShortArgs args) {
  static_assert(SMAX % 2 == 0 && SMAX <= 64, "pairs");
  constexpr int HP = SMAX / 2;
  extern __shared__ __align__(16) float ess_ring[];  // [DEPTH][SMAX][256]
  const EssArgs& a = args.a;
  const int S = a.S;
  const float fS = (float)S;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  const int64_t n0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // rows S .. SMAX-1 of this thread's ring columns are never copied into: zero them once, the loads below need no guard
  for (int d = 0; d < DEPTH; ++d)
    for (int t = S; t < SMAX; ++t) ess_ring[((size_t)d * SMAX + t) * 256 + threadIdx.x] = 0.0f;
  auto prefetch = [&](int64_t n, int stage) {
    if (n < a.N) {
      float* dst = ess_ring + ((size_t)stage * SMAX) * 256 + threadIdx.x;
      const float* src = a.x + n;
#pragma unroll 6
      for (int t = 0; t < S; ++t, src += a.stride, dst += 256) cp_async4_zfill(dst, src, true);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");  // one group per iteration, empty past the end
  };
#pragma unroll
  for (int d = 0; d < DEPTH - 1; ++d) prefetch(n0 + d * step, d);
  int stage = 0;
  for (int64_t n = n0; n < a.N; n += step) {
    const unsigned lanes = __activemask();
    prefetch(n + (DEPTH - 1) * step, (stage + DEPTH - 1) % DEPTH);
    asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
    const float* src = ess_ring + ((size_t)stage * SMAX) * 256 + threadIdx.x;
    stage = (stage + 1) % DEPTH;
    float x[SMAX + 2];
#pragma unroll
    for (int t = 0; t < SMAX; ++t) x[t] = src[t * 256];
    x[SMAX] = x[SMAX + 1] = 0.0f;
    float sum = 0.0f;
#pragma unroll
    for (int t = 0; t < SMAX; ++t) sum += x[t];
    // mean = sum / S (sgmcmc_diagnostic.py:107-108), correctly rounded without the division's slow path: q = sum * r,
    // q' = q + (sum - S q) r with r = RN(1 / S) (Markstein)
    const float q = sum * args.inv_S_rn;
    const float mean = fmaf(fmaf(-fS, q, sum), args.inv_S_rn, q);
#pragma unroll
    for (int t = 0; t < SMAX; ++t) x[t] = fmaf(-mean, args.live[t], x[t]);  // real rows centred, padding stays zero
    // The lagged products as packed float32x2 FMAs (FFMA2: one issue slot for two products) on the chain held as aligned
    // pairs xe[i] = (x[2i], x[2i+1]); entries past the end are zero.
    //   even lag 2m:  E_m = sum_i xe[i] * xe[i+m]                 ->  prod = E_m.lo + E_m.hi
    //   odd lags:     O_m = sum_i xe[i] * swap(xe[i+m])           (the operand swizzle of FFMA2, no register moves)
    //                 O_m.lo = sum_i x[2i] x[2(i+m)+1]   = the even-t half of lag 2m+1
    //                 O_m.hi = sum_i x[2i+1] x[2(i+m)]   = the odd-t  half of lag 2m-1
    //                 ->  prod(2m+1) = O_m.lo + O_{m+1}.hi
    // (a shifted copy of the chain for the odd lags cost two MOVs per FFMA2: the compiler rebuilt the pairs every time.)
    // Even and odd t accumulate separately and meet at the end; the reference's FFT route has no defined order either.
    float2 xe[HP];
#pragma unroll
    for (int i = 0; i < HP; ++i) xe[i] = make_float2(x[2 * i], x[2 * i + 1]);
    auto lag_pairs = [&](int m, bool swapped) {  // two accumulator pairs per chain: independent FFMA2s in flight
      float2 pa = make_float2(0.0f, 0.0f), pb = make_float2(0.0f, 0.0f);
#pragma unroll
      for (int i = 0; i + m < HP; i += 2) {
        const float2 b0 = xe[i + m];
        pa = __ffma2_rn(xe[i], swapped ? make_float2(b0.y, b0.x) : b0, pa);
        if (i + 1 + m < HP) {
          const float2 b1 = xe[i + 1 + m];
          pb = __ffma2_rn(xe[i + 1], swapped ? make_float2(b1.y, b1.x) : b1, pb);
        }
      }
      return make_float2(pa.x + pb.x, pa.y + pb.y);
    };
    float inv_cov0 = 0.0f, acc_w = 0.0f;
    bool open = true;
    float2 odd_prev = lag_pairs(0, true);  // O_0
    // Two lags per round (2m and 2m+1); the warp votes on the cut-off once per round.
#pragma unroll
    for (int k0 = 0; k0 < SMAX; k0 += 2) {
      if (k0 >= S) break;
      const int m = k0 / 2;
      const float2 ev = lag_pairs(m, false);
      const float2 odd_next = m + 1 < HP ? lag_pairs(m + 1, true) : make_float2(0.0f, 0.0f);
      float prod[2];
      prod[0] = ev.x + ev.y;
      prod[1] = odd_prev.x + odd_next.y;
      odd_prev = odd_next;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int k = k0 + q;
        if (k < S) {
          const float cov = prod[q] * args.inv_len[k];  // unbiased normalisation (:120-128)
          if (k == 0) inv_cov0 = 1.0f / cov;
          const float rho = cov * inv_cov0;
          if (a.use_threshold && rho < a.threshold) open = false;  // cumulative mask: this lag and every later one drop out
          const float w = args.wgt[k] * rho;
          if (open) acc_w += w;
          else if (rho != rho) acc_w += rho;  // NaN auto-correlations propagate like `weighted * mask` does
        }
      }
      // Once every chain of the warp is past its cut-off the later lags cannot change any result: a closed chain only
      // takes NaN terms, and |sum_t x[t] x[t+k]| <= sum_t x[t]^2 (Cauchy-Schwarz), so a lag can only be NaN / inf if lag
      // 0 already was - and then acc_w has been NaN since k = 0.
      if (a.use_threshold && !__any_sync(lanes, open)) break;
    }
    a.ess[n] = fS / (-1.0f + 2.0f * acc_w);
  }
}

template <int SMAX, int DEPTH>
static int launch_ess_short(const EssArgs& a, cudaStream_t st) {
  constexpr size_t smem = (size_t)DEPTH * SMAX * 256 * sizeof(float);
  constexpr int per_sm = (int)((220 * 1024) / smem) > 3 ? 3 : (int)((220 * 1024) / smem);  // resident blocks per SM the ring leaves room for
  auto kern = ess_short_kernel<SMAX, DEPTH>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  EssShortArgs args;
  args.a = a;
  const float fS = (float)a.S;
  args.inv_S_rn = 1.0f / fS;
  for (int t = 0; t < 64; ++t) {
    args.live[t] = t < a.S ? 1.0f : 0.0f;
    args.inv_len[t] = t < a.S ? 1.0f / (fS - (float)t) : 0.0f;
    args.wgt[t] = t < a.S ? (fS - (float)t) * (1.0f / fS) : 0.0f;
  }
  int64_t blocks = (a.N + 255) / 256;
  if (blocks > 148 * per_sm) blocks = 148 * per_sm;  // persistent: every block walks its parameters through the ring
  kern<<<(unsigned)blocks, 256, smem, st>>>(args);
  FB200_LAUNCH_CHECK();
  return 0;
}

template <int COLS>
static int launch_ess(const EssArgs& a, cudaStream_t st) {
  const int groups = 256 / COLS;
  const size_t smem =
      (((size_t)(a.S + ESS_R) + (size_t)(a.S > groups ? a.S : groups)) * COLS + (size_t)a.S) * sizeof(float);
  auto kern = ess_kernel<COLS>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  int64_t slabs = (a.N + COLS - 1) / COLS;
  int ctas_per_sm = (int)((220 * 1024) / (smem + 1024));
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  if (ctas_per_sm > 8) ctas_per_sm = 8;
  int64_t grid = slabs < (int64_t)148 * ctas_per_sm ? slabs : (int64_t)148 * ctas_per_sm;
  kern<<<(unsigned)grid, 256, smem, st>>>(a);
  FB200_LAUNCH_CHECK();
  return 0;
}

}  // namespace fb200

using namespace fb200;

static int ksd_tile(int S) { return S <= 32 ? 32 : 64; }

static void ksd_plan(int S, int64_t N, int* tiles_side, int* n_pair_tiles, int* n_splits, int64_t* cols_per_split) {
  const int KSD_T = ksd_tile(S);
  const int ts = (S + KSD_T - 1) / KSD_T;
  const int npt = ts * (ts + 1) / 2;
  // enough CTAs to fill the GPU a few times over, but at least 256 columns per split
  int64_t want = (148 * 4 + npt - 1) / npt;
  int64_t max_splits = (N + 255) / 256;
  if (want > max_splits) want = max_splits;
  if (want < 1) want = 1;
  int64_t cps = (N + want - 1) / want;
  cps = (cps + KSD_KC - 1) / KSD_KC * KSD_KC;
  const int64_t splits = (N + cps - 1) / cps;
  *tiles_side = ts;
  *n_pair_tiles = npt;
  *n_splits = (int)splits;
  *cols_per_split = cps;
}

extern "C" size_t fb200_ksd_imq_workspace_bytes(int S, int64_t N) {
  if (S <= 0 || N <= 0) return 0;
  int ts, npt, ns;
  int64_t cps;
  ksd_plan(S, N, &ts, &npt, &ns, &cps);
  return (size_t)ns * (size_t)S * (size_t)S * 4 * sizeof(float) + (size_t)S * sizeof(double);
}

extern "C" int fb200_ksd_imq(const float* samples, const float* grads, int S, int64_t N, int64_t row_stride, float c,
                             float beta, void* workspace, size_t workspace_bytes, float* out, void* stream) {
  if (!samples || !grads || !workspace || !out) return FB200_E_NULL;
  if (S <= 0 || N <= 0 || row_stride < N) return FB200_E_SHAPE;
  if (!(c > 0.0f) || !(beta < 0.0f)) return FB200_E_SHAPE;
  if (S > 4096) return FB200_E_UNSUPPORTED;
  if (!fb200_aligned16(workspace)) return FB200_E_ALIGN;
  KsdArgs a{};
  a.x = samples;
  a.g = grads;
  a.S = S;
  a.N = N;
  a.stride = row_stride;
  int npt;
  ksd_plan(S, N, &a.n_tiles_side, &npt, &a.n_splits, &a.cols_per_split);
  const size_t partial_bytes = (size_t)a.n_splits * (size_t)S * (size_t)S * 4 * sizeof(float);
  if (workspace_bytes < partial_bytes + (size_t)S * sizeof(double)) return FB200_E_WORKSPACE;
  a.partial = reinterpret_cast<float*>(workspace);
  double* row_sums = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(workspace) + partial_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  a.tile = ksd_tile(S);
  const dim3 grid((unsigned)npt, (unsigned)a.n_splits);
  auto launch = [&](auto kern, int T) -> int {
    const size_t smem = (size_t)2 * 4 * KSD_KC * (T + 4) * sizeof(float);  // two stages of four [32][T + 4] arrays
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<grid, KSD_THREADS, smem, st>>>(a);
    return 0;
  };
  int mirror_granule = a.tile;  // pairs (i, j) with i / granule > j / granule are read at the mirrored entry
  if (S <= 32) {
    // one diagonal tile: the triangle kernel (4 x 4 blocks bi <= bj only)
    // two stages of [x | g][112][36] floats = 64 512 B, re-used by the final reduction over the 7 groups (also 64 512 B)
    static_assert(2 * 2 * KSD_TRI_KC * KSD_TRI_ROW * sizeof(float) >= KSD_TRI_GROUPS * 36 * 16 * sizeof(float4), "red fits");
    const size_t smem = (size_t)2 * 2 * KSD_TRI_KC * KSD_TRI_ROW * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(ksd_tri32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    ksd_tri32_kernel<<<dim3(1, (unsigned)a.n_splits), KSD_TRI_THREADS, smem, st>>>(a);
    mirror_granule = 4;
  } else {
    const int lrc = launch(ksd_pair_sums_kernel<64>, 64);
    if (lrc != 0) return lrc;
  }
  FB200_LAUNCH_CHECK();
  ksd_row_kernel<<<(unsigned)S, 256, 0, st>>>(a.partial, S, N, a.n_splits, mirror_granule, c, beta, row_sums);
  FB200_LAUNCH_CHECK();
  ksd_total_kernel<<<1, 32, 0, st>>>(row_sums, S, out);
  FB200_LAUNCH_CHECK();
  return 0;
}

extern "C" int fb200_ess(const float* samples, int S, int64_t N, int64_t row_stride, int use_threshold,
                         float threshold, float* ess, void* stream) {
  if (!samples || !ess) return FB200_E_NULL;
  if (S <= 0 || N <= 0 || row_stride < N) return FB200_E_SHAPE;
  EssArgs a{};
  a.x = samples;
  a.S = S;
  a.N = N;
  a.stride = row_stride;
  a.use_threshold = use_threshold;
  a.threshold = threshold;
  a.ess = ess;
  cudaStream_t st = (cudaStream_t)stream;
  if (S <= 32) return launch_ess_short<32, 2>(a, st);  // the chain of a parameter fits one thread's registers
  if (S <= 64) return launch_ess_short<64, 2>(a, st);
  const size_t per_col = ((size_t)(2 * S + ESS_R + 32)) * sizeof(float);
  const size_t budget = 200 * 1024 - (size_t)S * sizeof(float);
  if (per_col * 128 <= 110 * 1024 && N >= 128) return launch_ess<128>(a, st);  // short chains: wide slabs
  if (per_col * 32 <= budget) return launch_ess<32>(a, st);
  if (per_col * 16 <= budget) return launch_ess<16>(a, st);
  if (per_col * 8 <= budget) return launch_ess<8>(a, st);
  return FB200_E_UNSUPPORTED;  // chains longer than ~3000 stored samples
}
