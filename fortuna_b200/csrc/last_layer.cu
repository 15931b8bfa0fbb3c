// K2 (CUDA-core path) + dispatch: last-layer logits for S posterior samples.
//
// Replaces Likelihood._get_batched_calibrated_outputs (fortuna/likelihood/base.py:429-458) restricted to
// the final nn.Dense (fortuna/model/mlp.py:227, lenet.py:129, resnet.py:248) and the temperature scaler
// (fortuna/output_calibrator/classification.py:14-17):
//     out[s,b,c] = (sum_d x[b,d] w[s,d,c] + bias[s,c]) * exp(-log_temp)
// The reference runs S full forward passes; here the features x are computed once by the caller and only
// this contraction is repeated per sample.  Shapes whose D*C is tiny (two-moons 30x2, LeNet 84x10, BERT
// 768x2) are bandwidth/latency bound and run on this shared-memory-tiled FFMA kernel; large bf16 problems
// go to the tcgen05 kernel in last_layer_tc.cu.
#include <cuda_bf16.h>

#include <math_constants.h>

#include "common.cuh"

namespace fb200 {

int last_layer_tc_supported(int S, int B, int D, int C, int in_dtype, int out_dtype, int w_layout);
size_t last_layer_tc_lse_workspace_bytes(int S, int B, int C);
int last_layer_tc_launch(const void* x, const void* w, const float* bias, const float* log_temp,
                         const int32_t* sample_idx, int n_stored, void* out, int S, int B, int D, int C,
                         int out_dtype, int c_split, cudaStream_t st, int epi = 0, float* lse_out = nullptr,
                         void* lse_ws = nullptr);

int last_layer_tc2_launch(const void* x, const void* w, const float* bias, const float* log_temp,
                          const int32_t* sample_idx, int n_stored, void* out, int S, int B, int D, int C,
                          int out_dtype, cudaStream_t st);

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ void store_out(float* p, float v) { *p = v; }
__device__ __forceinline__ void store_out(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// The contraction is tiled over the FLATTENED output column n = s * C + c (N = S * C columns): a narrow last layer
// (C = 2 .. 10) still fills 64-column tiles, and the features x are re-read N / 64 times instead of S times.
constexpr int kBM = 128, kBN = 64, kBK = 16;

template <typename TI, typename TO, int WL>
__global__ void __launch_bounds__(256) last_layer_simt_kernel(const TI* __restrict__ x, const TI* __restrict__ w,
                                                              const float* __restrict__ bias,
                                                              const float* __restrict__ log_temp,
                                                              const int32_t* __restrict__ sample_idx,
                                                              TO* __restrict__ out, int S, int B, int D, int C) {
  __shared__ __align__(16) float As[kBK][kBM + 4];
  __shared__ __align__(16) float Bs[kBK][kBN + 4];
  const int N = S * C;
  const int n0 = blockIdx.x * kBN;
  const int b0 = blockIdx.y * kBM;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // thread tile: rows ty*8 .. +8, columns tx*4 .. +4

  // loader roles.  A: row r = tid / 2, 8 consecutive k.  B: one column, 4 consecutive k (SCD) or
  // one k, 4 consecutive columns (SDC); the weight row of a column, w[idx[s]][.][c], is resolved once.
  const int a_r = threadIdx.x >> 1, a_k = (threadIdx.x & 1) * 8;
  const int a_b = b0 + a_r;
  const TI* a_ptr = x + (size_t)(a_b < B ? a_b : 0) * D;
  int b_col, b_k;
  if (WL == FB200_W_SCD) {
    b_col = threadIdx.x >> 2;
    b_k = (threadIdx.x & 3) * 4;
  } else {
    b_col = (threadIdx.x & 15) * 4;
    b_k = threadIdx.x >> 4;
  }
  // (sample, class) of the 4 columns this thread loads (SDC) / of its single column (SCD, q = 0)
  size_t wbase[4];
  bool wok[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int n = n0 + b_col + (WL == FB200_W_SCD ? 0 : q);
    wok[q] = n < N;
    const int sn = wok[q] ? n / C : 0, cn = wok[q] ? n - sn * C : 0;
    const int sw = sample_idx ? sample_idx[sn] : sn;
    wbase[q] = WL == FB200_W_SCD ? ((size_t)sw * C + cn) * (size_t)D : (size_t)sw * D * C + cn;
  }

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  for (int k0 = 0; k0 < D; k0 += kBK) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int k = k0 + a_k + q;
      As[a_k + q][a_r] = (a_b < B && k < D) ? to_f32(a_ptr[k]) : 0.0f;
    }
    if (WL == FB200_W_SCD) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int k = k0 + b_k + q;
        Bs[b_k + q][b_col] = (wok[0] && k < D) ? to_f32(w[wbase[0] + k]) : 0.0f;
      }
    } else {
      const int k = k0 + b_k;
#pragma unroll
      for (int q = 0; q < 4; ++q) Bs[b_k][b_col + q] = (wok[q] && k < D) ? to_f32(w[wbase[q] + (size_t)k * C]) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      const float4 bq = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const float scale = log_temp ? expf(-log_temp[0]) : 1.0f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = n0 + tx * 4 + j;
    if (n >= N) continue;
    const int sn = n / C, cn = n - sn * C;
    const int sw = sample_idx ? sample_idx[sn] : sn;
    const float bj = bias ? bias[(size_t)sw * C + cn] : 0.0f;
    TO* ocol = out + (size_t)sn * B * C + cn;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int b = b0 + ty * 8 + i;
      if (b >= B) continue;
      float v = acc[i][j] + bj;
      if (log_temp) v *= scale;
      store_out(ocol + (size_t)b * C, v);
    }
  }
}

// float32, K-major weights (w[s][c][d]: what a torch nn.Linear holds and what the predictive passes), D % 4 == 0, 16-byte
// aligned operands - the c2 / c3 shapes (LeNet 84 x 10, ResNet-18 512 x 10).  Same 128 x 64 x 16 tiling over the flattened
// column n = s * C + c and the same k-ascending FMA order per output as the kernel above (bit-identical results), but:
// both operands stay K-major in shared memory ([row][k], 20-float rows) and arrive by 16-byte cp.async into a double
// buffer (one barrier per k-step, copies of step k + 1 under the arithmetic of step k); fragments are read as float4
// along k (12 LDS.128 per 128 FMAs; a thread's four columns are 16 apart, which makes the weight reads conflict-free);
// the finished tile goes through shared memory so that each sample's [rows, classes] block - contiguous in the output -
// is written with coalesced stores instead of 4-byte stores C floats apart.  Measured at c3 (S=256, B=10 000, D=512,
// C=10): 1.00 -> 0.68 ms = 38 TFLOP/s, which is this part's float32 CUDA-core rate (a 3-register FFMA issues every
// other cycle per scheduler: 148 x 64 lanes x 2 x 1.96 GHz = 37 TFLOP/s; the same loop on packed FFMA2 - even / odd k in
// the two halves - measured 0.68 ms too and was not kept: it is the pipe, not the issue slots).
constexpr int kKM_ROW = kBK + 4;  // floats per staged row: 16-byte aligned, 20 * r mod 32 distinct for 8 consecutive r

__device__ __forceinline__ void cp_async16_zfill(float* dst_smem, const float* src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}

template <typename TO>
__global__ void __launch_bounds__(256, 2) last_layer_simt_kmajor_kernel(const float* __restrict__ x,
                                                                        const float* __restrict__ w,
                                                                        const float* __restrict__ bias,
                                                                        const float* __restrict__ log_temp,
                                                                        const int32_t* __restrict__ sample_idx,
                                                                        TO* __restrict__ out, int S, int B, int D, int C) {
  constexpr int A_STAGE = kBM * kKM_ROW, B_STAGE = kBN * kKM_ROW;
  constexpr int C_ROW = kBN + 1;
  constexpr int SMEM_FLOATS = (2 * (A_STAGE + B_STAGE) > kBM * C_ROW) ? 2 * (A_STAGE + B_STAGE) : kBM * C_ROW;
  __shared__ __align__(16) float smem[SMEM_FLOATS];
  float* As = smem;                // [2][kBM][kKM_ROW]
  float* Bs = smem + 2 * A_STAGE;  // [2][kBN][kKM_ROW]
  const int N = S * C;
  const int n0 = blockIdx.x * kBN;
  const int b0 = blockIdx.y * kBM;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // thread tile: rows ty*8 .. +8, columns tx + 16 j

  // copy roles: A - two 16-byte chunks per thread (row = id / 4, k chunk = id % 4); B - one chunk (column tid / 4)
  const int kc = tid & 3;
  const int a_row0 = tid >> 2, a_row1 = a_row0 + 64;
  const bool a_ok0 = b0 + a_row0 < B, a_ok1 = b0 + a_row1 < B;
  const float* a_src0 = x + (size_t)(a_ok0 ? b0 + a_row0 : 0) * D + kc * 4;
  const float* a_src1 = x + (size_t)(a_ok1 ? b0 + a_row1 : 0) * D + kc * 4;
  const int b_col = tid >> 2;
  const bool b_ok = n0 + b_col < N;
  const float* b_src;
  {
    const int n = b_ok ? n0 + b_col : 0;
    const int sn = n / C, cn = n - sn * C;
    const int sw = sample_idx ? sample_idx[sn] : sn;
    b_src = w + ((size_t)sw * C + cn) * (size_t)D + kc * 4;
  }
  auto stage = [&](int buf, int k0) {
    const bool kin = k0 + kc * 4 < D;  // D % 4 == 0: a chunk is inside or outside as a whole
    cp_async16_zfill(As + buf * A_STAGE + a_row0 * kKM_ROW + kc * 4, a_src0 + (kin ? k0 : 0), a_ok0 && kin);
    cp_async16_zfill(As + buf * A_STAGE + a_row1 * kKM_ROW + kc * 4, a_src1 + (kin ? k0 : 0), a_ok1 && kin);
    cp_async16_zfill(Bs + buf * B_STAGE + b_col * kKM_ROW + kc * 4, b_src + (kin ? k0 : 0), b_ok && kin);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  stage(0, 0);
  int buf = 0;
  for (int k0 = 0; k0 < D; k0 += kBK, buf ^= 1) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();  // step k0 has landed for everyone, and everyone is done reading the other buffer
    if (k0 + kBK < D) stage(buf ^ 1, k0 + kBK);
    const float* Ab = As + buf * A_STAGE + (ty * 8) * kKM_ROW;
    const float* Bb = Bs + buf * B_STAGE + tx * kKM_ROW;
#pragma unroll
    for (int kq = 0; kq < kBK / 4; ++kq) {
      float4 a4[8], b4[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) a4[i] = *reinterpret_cast<const float4*>(Ab + i * kKM_ROW + kq * 4);
#pragma unroll
      for (int j = 0; j < 4; ++j) b4[j] = *reinterpret_cast<const float4*>(Bb + (16 * j) * kKM_ROW + kq * 4);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[i][j] = fmaf(a4[i].x, b4[j].x, acc[i][j]);
          acc[i][j] = fmaf(a4[i].y, b4[j].y, acc[i][j]);
          acc[i][j] = fmaf(a4[i].z, b4[j].z, acc[i][j]);
          acc[i][j] = fmaf(a4[i].w, b4[j].w, acc[i][j]);
        }
    }
  }
  __syncthreads();  // the operand buffers become the output tile

  const float scale = log_temp ? expf(-log_temp[0]) : 1.0f;
  float* Cs = smem;  // [kBM][C_ROW]
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int nl = tx + 16 * j, n = n0 + nl;
    float bj = 0.0f;
    if (bias && n < N) {
      const int sn = n / C, cn = n - sn * C;
      bj = bias[(size_t)(sample_idx ? sample_idx[sn] : sn) * C + cn];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = acc[i][j] + bj;
      if (log_temp) v *= scale;
      Cs[(ty * 8 + i) * C_ROW + nl] = v;
    }
  }
  __syncthreads();
  // per sample of the tile: out[s][b0 .. b0 + rows)[c_lo .. c_hi) - one contiguous block when the sample's classes are all
  // inside the tile, rows of (c_hi - c_lo) floats C apart at the tile's edges
  const int rows = min(kBM, B - b0);
  const int n_end = min(n0 + kBN, N);
  for (int ns = n0; ns < n_end;) {
    const int sn = ns / C, c_lo = ns - sn * C;
    const int c_hi = min(C, c_lo + (n_end - ns));
    const int wseg = c_hi - c_lo;
    TO* obase = out + ((size_t)sn * B + b0) * C + c_lo;
    const float* cbase = Cs + (ns - n0);
    for (int e = tid; e < rows * wseg; e += 256) {
      const int r = e / wseg, c = e - r * wseg;
      store_out(obase + (size_t)r * C + c, cbase[r * C_ROW + c]);
    }
    ns += wseg;
  }
}

// wt[s][c][d] (bf16) = w[s][d][c]
template <typename TI>
__global__ void __launch_bounds__(256) transpose_w_kernel(const TI* __restrict__ w, __nv_bfloat16* __restrict__ wt,
                                                          int D, int C) {
  __shared__ float tile[32][33];
  const int s = blockIdx.z;
  const int d0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const TI* ws = w + (size_t)s * D * C;
  __nv_bfloat16* wts = wt + (size_t)s * D * C;
  for (int i = ty; i < 32; i += 8) {
    const int d = d0 + i, c = c0 + tx;
    tile[i][tx] = (d < D && c < C) ? to_f32(ws[(size_t)d * C + c]) : 0.0f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, d = d0 + tx;
    if (c < C && d < D) wts[(size_t)c * D + d] = __float2bfloat16_rn(tile[tx][i]);
  }
}

template <typename TI, typename TO>
static int launch_simt(const void* x, const void* w, const float* bias, const float* log_temp,
                       const int32_t* sample_idx, void* out, int S, int B, int D, int C, int w_layout,
                       cudaStream_t st) {
  const long long N = (long long)S * C;
  if (N > 0x7FFFFFFFll) return FB200_E_UNSUPPORTED;
  dim3 grid((unsigned)((N + kBN - 1) / kBN), (B + kBM - 1) / kBM, 1);
  if (grid.y > 65535) return FB200_E_UNSUPPORTED;
  if constexpr (sizeof(TI) == 4) {
    // float32 operands, K-major weights, whole 16-byte chunks along k: the double-buffered kernel
    // (from D = 128 up: at LeNet's D = 84 - six k-steps - the tile epilogue weighs as much as the copies save, 89 vs 84 us)
    if (w_layout == FB200_W_SCD && D % 4 == 0 && D >= 128 && fb200_aligned16(x) && fb200_aligned16(w)) {
      last_layer_simt_kmajor_kernel<TO><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(x),
                                                                reinterpret_cast<const float*>(w), bias, log_temp,
                                                                sample_idx, reinterpret_cast<TO*>(out), S, B, D, C);
      FB200_LAUNCH_CHECK();
      return 0;
    }
  }
  if (w_layout == FB200_W_SDC)
    last_layer_simt_kernel<TI, TO, FB200_W_SDC><<<grid, 256, 0, st>>>(
        reinterpret_cast<const TI*>(x), reinterpret_cast<const TI*>(w), bias, log_temp, sample_idx,
        reinterpret_cast<TO*>(out), S, B, D, C);
  else
    last_layer_simt_kernel<TI, TO, FB200_W_SCD><<<grid, 256, 0, st>>>(
        reinterpret_cast<const TI*>(x), reinterpret_cast<const TI*>(w), bias, log_temp, sample_idx,
        reinterpret_cast<TO*>(out), S, B, D, C);
  FB200_LAUNCH_CHECK();
  return 0;
}


// Generic (CUDA-core) row pass for shapes outside the tcgen05 epilogue: lse[row] = logsumexp_c z[row, :], and
// optionally z <- softmax / log_softmax in place.  One warp per row.
template <typename TO>
__global__ void __launch_bounds__(256) row_lse_kernel(TO* __restrict__ z, int64_t rows, int C, int epi,
                                                      float* __restrict__ lse_out, const float* __restrict__ log_temp) {
  const int lane = threadIdx.x & 31;
  const float scale = log_temp ? expf(-log_temp[0]) : 1.0f;  // temperature scaler: statistics of z * exp(-log_temp)
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp0; r < rows; r += nwarps) {
    TO* row = z + r * C;
    float m = -INFINITY;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, (float)row[c] * scale);
    m = group_max<32>(m);
    // jsp.special.logsumexp replaces a non-finite maximum by 0 (lse, and the log-softmax z - lse the reference's
    // per-class log_prob is); jax.nn.softmax subtracts the maximum as it is: a row holding +inf, or only -inf, is NaN
    // in every class there
    const bool m_finite = fabsf(m) <= 3.4028234e38f;
    const float ms = m_finite ? m : 0.0f;
    float s = 0.0f;
    for (int c = lane; c < C; c += 32) s += __expf((float)row[c] * scale - ms);
    s = group_sum<32>(s);
    const float lse = ms + __logf(s);
    if (lane == 0 && lse_out) lse_out[r] = lse;
    if (epi == 1) {
      for (int c = lane; c < C; c += 32)
        row[c] = (TO)(m_finite ? __expf((float)row[c] * scale - lse) : CUDART_NAN_F);
    } else if (epi == 2) {
      for (int c = lane; c < C; c += 32) row[c] = (TO)((float)row[c] * scale - lse);
    } else if (log_temp) {
      for (int c = lane; c < C; c += 32) row[c] = (TO)((float)row[c] * scale);  // calibrated logits in place
    }
  }
}

// log_prob from logits + their per-row logsumexp: ll[s,b] = z[s,b,y_b] - lse[s,b] (S*B gathers, one 32-byte sector
// each, instead of a pass over [S,B,C]); out[b] = logsumexp_s ll - log S (predictive/base.py:179-203, 104-128).
template <typename T>
__global__ void __launch_bounds__(256) log_prob_from_lse_kernel(const T* __restrict__ z, const float* __restrict__ lse,
                                                                const int32_t* __restrict__ targets, int S, int B,
                                                                int C, float* __restrict__ ens,
                                                                float* __restrict__ log_prob) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int y = targets[b];
  const bool ok = (unsigned)y < (unsigned)C;
  float m = -INFINITY;
  for (int s = 0; s < S; ++s) {
    const size_t row = (size_t)s * B + b;
    const float ll = ok ? (float)z[row * C + y] - lse[row] : CUDART_NAN_F;
    if (ens) ens[row] = ll;
    m = fmaxf(m, ll);
  }
  if (!log_prob) return;
  float acc = 0.0f;
  for (int s = 0; s < S; ++s) {
    const size_t row = (size_t)s * B + b;
    const float ll = ok ? (float)z[row * C + y] - lse[row] : CUDART_NAN_F;
    acc += expf(ll - m);
  }
  log_prob[b] = ok ? (m + logf(acc)) - logf((float)S) : CUDART_NAN_F;
}

}  // namespace fb200

using namespace fb200;

extern "C" int fb200_last_layer_logits(const void* x, const void* w, const float* bias, const float* log_temp,
                                       const int32_t* sample_idx, int n_stored, void* out, int S, int B, int D,
                                       int C, int in_dtype, int out_dtype, int w_layout, int path, void* stream) {
  if (!x || !w || !out) return FB200_E_NULL;
  if (S <= 0 || B <= 0 || D <= 0 || C <= 0 || n_stored <= 0) return FB200_E_SHAPE;
  if (!sample_idx && n_stored != S) return FB200_E_SHAPE;
  if ((in_dtype != FB200_F32 && in_dtype != FB200_BF16) || (out_dtype != FB200_F32 && out_dtype != FB200_BF16))
    return FB200_E_UNSUPPORTED;
  if (w_layout != FB200_W_SDC && w_layout != FB200_W_SCD) return FB200_E_UNSUPPORTED;
  if (path < 0 || path > 3) return FB200_E_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int tc_ok = last_layer_tc_supported(S, B, D, C, in_dtype, out_dtype, w_layout);
  if ((path == 2 || path == 3) && !tc_ok) return FB200_E_UNSUPPORTED;
  if (path == 3) return last_layer_tc2_launch(x, w, bias, log_temp, sample_idx, n_stored, out, S, B, D, C, out_dtype, st);
  // auto: tensor cores only when the contraction is large enough to be compute-bound (D >= 256, C >= 64)
  const bool use_tc = path == 2 || (path == 0 && tc_ok && D >= 256 && C >= 64 && B >= 128);
  // long contractions are operand-traffic bound on one CTA: the CTA-pair kernel (cta_group::2) moves 2/3 of the bytes
  // per flop (measured at c4, D = 2048: 1306 vs 1212 TFLOP/s f32 logits; no difference at D = 512)
  if (use_tc && path == 0 && D >= 1024 && B >= 512 && C >= 256)
    return last_layer_tc2_launch(x, w, bias, log_temp, sample_idx, n_stored, out, S, B, D, C, out_dtype, st);
  if (use_tc) return last_layer_tc_launch(x, w, bias, log_temp, sample_idx, n_stored, out, S, B, D, C, out_dtype, 0, st);
  // narrow last layer in bf16 (e.g. BERT classifier 768 x 2): fold the sample axis into the column axis so the
  // tensor cores see N = S * C columns; needs the identity sample map (stored rows are then one [S*C, D] matrix)
  const long long flatN = (long long)S * C;
  if (path == 0 && !sample_idx && C < 64 && flatN >= 64 && flatN < (1ll << 30) && D >= 256 && B >= 128 &&
      last_layer_tc_supported(1, B, D, (int)flatN, in_dtype, out_dtype, w_layout))
    return last_layer_tc_launch(x, w, bias, log_temp, nullptr, 1, out, 1, B, D, (int)flatN, out_dtype, C, st);
  if (in_dtype == FB200_F32) {
    if (out_dtype == FB200_F32) return launch_simt<float, float>(x, w, bias, log_temp, sample_idx, out, S, B, D, C, w_layout, st);
    return launch_simt<float, __nv_bfloat16>(x, w, bias, log_temp, sample_idx, out, S, B, D, C, w_layout, st);
  }
  if (out_dtype == FB200_F32)
    return launch_simt<__nv_bfloat16, float>(x, w, bias, log_temp, sample_idx, out, S, B, D, C, w_layout, st);
  return launch_simt<__nv_bfloat16, __nv_bfloat16>(x, w, bias, log_temp, sample_idx, out, S, B, D, C, w_layout, st);
}

extern "C" size_t fb200_last_layer_ex_workspace_bytes(int S, int B, int C) {
  if (S <= 0 || B <= 0 || C <= 0) return 0;
  return last_layer_tc_lse_workspace_bytes(S, B, C);
}

extern "C" int fb200_last_layer_logits_ex(const void* x, const void* w, const float* bias, const float* log_temp,
                                          const int32_t* sample_idx, int n_stored, void* out, int S, int B, int D,
                                          int C, int in_dtype, int out_dtype, int w_layout, int epilogue,
                                          float* lse_out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!x || !w || !out) return FB200_E_NULL;
  if (S <= 0 || B <= 0 || D <= 0 || C <= 0 || n_stored <= 0) return FB200_E_SHAPE;
  if (!sample_idx && n_stored != S) return FB200_E_SHAPE;
  if (epilogue < FB200_EPI_LOGITS || epilogue > FB200_EPI_LOG_SOFTMAX) return FB200_E_UNSUPPORTED;
  if ((in_dtype != FB200_F32 && in_dtype != FB200_BF16) || (out_dtype != FB200_F32 && out_dtype != FB200_BF16))
    return FB200_E_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  if (epilogue == FB200_EPI_LOGITS && !lse_out)
    return fb200_last_layer_logits(x, w, bias, log_temp, sample_idx, n_stored, out, S, B, D, C, in_dtype, out_dtype,
                                   w_layout, 0, stream);
  // fused: the one-CTA tcgen05 kernel's epilogue owns whole rows of a 128 x 256 tile
  const bool tc = last_layer_tc_supported(S, B, D, C, in_dtype, out_dtype, w_layout) && D >= 256 && C >= 64 && B >= 128;
  if (tc && (epilogue == FB200_EPI_LOGITS || C <= 256)) {
    const size_t need = last_layer_tc_lse_workspace_bytes(S, B, C);
    if (need && (!workspace || workspace_bytes < need)) return FB200_E_WORKSPACE;
    return last_layer_tc_launch(x, w, bias, log_temp, sample_idx, n_stored, out, S, B, D, C, out_dtype, 0, st, epilogue,
                                lse_out, need ? workspace : nullptr);
  }
  // everything else: logits by the regular dispatch, then one row pass on CUDA cores (in place)
  const int rc = fb200_last_layer_logits(x, w, bias, log_temp, sample_idx, n_stored, out, S, B, D, C, in_dtype,
                                         out_dtype, w_layout, 0, stream);
  if (rc != 0) return rc;
  const int64_t rows = (int64_t)S * B;
  int64_t blocks = (rows + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (out_dtype == FB200_F32)
    row_lse_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<float*>(out), rows, C, epilogue, lse_out, nullptr);
  else
    row_lse_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<__nv_bfloat16*>(out), rows, C,
                                                                     epilogue, lse_out, nullptr);
  FB200_LAUNCH_CHECK();
  return 0;
}

extern "C" int fb200_softmax_rows(void* z, int dtype, int64_t rows, int C, int epilogue, const float* log_temp,
                                  float* lse_out, void* stream) {
  if (!z) return FB200_E_NULL;
  if (rows <= 0 || C <= 0) return FB200_E_SHAPE;
  if (epilogue < FB200_EPI_LOGITS || epilogue > FB200_EPI_LOG_SOFTMAX) return FB200_E_UNSUPPORTED;
  if (epilogue == FB200_EPI_LOGITS && !lse_out && !log_temp) return FB200_E_NULL;
  int64_t blocks = (rows + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == FB200_F32)
    row_lse_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<float*>(z), rows, C, epilogue, lse_out,
                                                            log_temp);
  else if (dtype == FB200_BF16)
    row_lse_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<__nv_bfloat16*>(z), rows, C,
                                                                     epilogue, lse_out, log_temp);
  else
    return FB200_E_UNSUPPORTED;
  FB200_LAUNCH_CHECK();
  return 0;
}

extern "C" int fb200_log_prob_from_lse(const void* logits, int dtype, const float* lse, const int32_t* targets, int S,
                                       int B, int C, float* ensemble_log_prob, float* log_prob, void* stream) {
  if (!logits || !lse || !targets) return FB200_E_NULL;
  if (!ensemble_log_prob && !log_prob) return FB200_E_NULL;
  if (S <= 0 || B <= 0 || C <= 0) return FB200_E_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)((B + 255) / 256);
  if (dtype == FB200_F32)
    log_prob_from_lse_kernel<float><<<blocks, 256, 0, st>>>(reinterpret_cast<const float*>(logits), lse, targets, S, B,
                                                            C, ensemble_log_prob, log_prob);
  else if (dtype == FB200_BF16)
    log_prob_from_lse_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(logits),
                                                                    lse, targets, S, B, C, ensemble_log_prob, log_prob);
  else
    return FB200_E_UNSUPPORTED;
  FB200_LAUNCH_CHECK();
  return 0;
}

extern "C" int fb200_transpose_w_bf16(const void* w, int in_dtype, void* wt, int S, int D, int C, void* stream) {
  if (!w || !wt) return FB200_E_NULL;
  if (S <= 0 || D <= 0 || C <= 0) return FB200_E_SHAPE;
  dim3 grid((C + 31) / 32, (D + 31) / 32, S);
  if (grid.y > 65535 || grid.z > 65535) return FB200_E_UNSUPPORTED;
  if (in_dtype == FB200_F32)
    transpose_w_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float*>(w),
                                                                      reinterpret_cast<__nv_bfloat16*>(wt), D, C);
  else if (in_dtype == FB200_BF16)
    transpose_w_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(w), reinterpret_cast<__nv_bfloat16*>(wt), D, C);
  else
    return FB200_E_UNSUPPORTED;
  FB200_LAUNCH_CHECK();
  return 0;
}
