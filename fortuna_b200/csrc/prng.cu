// jax.random primitives on the device: bit-exact threefry2x32 stream (jax 0.4.30, non-partitionable).
// Reference call sites: fortuna/utils/random.py:11-23,38-49; sghmc_integrator.py:66;
// sgmcmc_posterior.py:51; predictive/base.py:113,188,632; training/trainer.py:676-680.
#include "common.cuh"

namespace fb200 {

// One thread per counter pair i < h: out[i] = lane 0, out[h+i] = lane 1 (dropped for the odd-size pad).
template <bool NORMAL>
__global__ void __launch_bounds__(256) random_stream_kernel(const uint32_t* __restrict__ key, uint32_t n,
                                                            void* __restrict__ out) {
  const uint32_t k0 = key[0], k1 = key[1];
  const uint32_t h = (n >> 1) + (n & 1u);
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < h; i += stride) {
    uint32_t x0 = (uint32_t)i;
    const bool has1 = (uint32_t)i + h < n;
    uint32_t x1 = has1 ? (uint32_t)i + h : 0u;
    threefry2x32(k0, k1, x0, x1);
    if (NORMAL) {
      float* o = reinterpret_cast<float*>(out);
      o[i] = bits_to_normal(x0);
      if (has1) o[i + h] = bits_to_normal(x1);
    } else {
      uint32_t* o = reinterpret_cast<uint32_t*>(out);
      o[i] = x0;
      if (has1) o[i + h] = x1;
    }
  }
}

__global__ void fold_in_kernel(const uint32_t* __restrict__ key, uint32_t data, uint32_t* __restrict__ out) {
  // threefry_2x32(key, [0, data]): one block, counters (0, data).
  uint32_t x0 = 0u, x1 = data;
  threefry2x32(key[0], key[1], x0, x1);
  out[0] = x0;
  out[1] = x1;
}

// jax.random.choice(keys[s], n_samples) == randint(key, (), 0, n): two 32-bit draws, multiplier trick.
// SPLIT: key is one key and draw s uses split(key, n_draws)[s]; otherwise key is n_draws keys, one per draw.
template <bool SPLIT>
__global__ void __launch_bounds__(128) sample_indices_kernel(const uint32_t* __restrict__ key, uint32_t n_draws,
                                                             uint32_t span, int32_t* __restrict__ out) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_draws) return;
  uint32_t ks0, ks1;
  if (SPLIT) {
    // keys = split(key, n_draws): row s of random_bits(key, 2 * n_draws).reshape(n_draws, 2)
    const uint32_t k0 = key[0], k1 = key[1];
    ks0 = random_bits_elem(k0, k1, 2u * n_draws, 2u * s);
    ks1 = random_bits_elem(k0, k1, 2u * n_draws, 2u * s + 1u);
  } else {
    ks0 = key[2u * s];
    ks1 = key[2u * s + 1u];
  }
  // k1_, k2_ = split(keys[s]) : random_bits(keys[s], 4) = [y0(0,2), y0(1,3), y1(0,2), y1(1,3)]
  uint32_t a0 = 0u, a1 = 2u, b0 = 1u, b1 = 3u;
  threefry2x32(ks0, ks1, a0, a1);
  threefry2x32(ks0, ks1, b0, b1);
  // higher_bits = random_bits(k1_, 1)[0], lower_bits = random_bits(k2_, 1)[0]: counters (0, pad 0), lane 0
  uint32_t hi = 0u, hi1 = 0u, lo = 0u, lo1 = 0u;
  threefry2x32(a0, b0, hi, hi1);
  threefry2x32(a1, b1, lo, lo1);
  uint32_t mult = 65536u % span;
  mult = (mult * mult) % span;
  uint32_t off = (hi % span) * mult + (lo % span);
  out[s] = (int32_t)(off % span);
}

}  // namespace fb200

using namespace fb200;

static inline int grid_for(uint64_t work, int block) {
  uint64_t g = (work + block - 1) / block;
  const uint64_t cap = 148ull * 16ull;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

extern "C" int fb200_random_bits(const uint32_t* key, int64_t n, uint32_t* out, void* stream) {
  if (!key || !out) return FB200_E_NULL;
  if (n <= 0 || n >= 0xFFFFFFFFll) return FB200_E_SHAPE;
  const uint64_t h = ((uint64_t)n + 1) / 2;
  random_stream_kernel<false><<<grid_for(h, 256), 256, 0, (cudaStream_t)stream>>>(key, (uint32_t)n, out);
  FB200_LAUNCH_CHECK();
  return 0;
}

extern "C" int fb200_random_normal(const uint32_t* key, int64_t n, float* out, void* stream) {
  if (!key || !out) return FB200_E_NULL;
  if (n <= 0 || n >= 0xFFFFFFFFll) return FB200_E_SHAPE;
  const uint64_t h = ((uint64_t)n + 1) / 2;
  random_stream_kernel<true><<<grid_for(h, 256), 256, 0, (cudaStream_t)stream>>>(key, (uint32_t)n, out);
  FB200_LAUNCH_CHECK();
  return 0;
}

extern "C" int fb200_threefry_split(const uint32_t* key, int num, uint32_t* out_keys, void* stream) {
  if (num <= 0 || num >= (1 << 30)) return FB200_E_SHAPE;
  return fb200_random_bits(key, 2ll * num, out_keys, stream);
}

extern "C" int fb200_threefry_fold_in(const uint32_t* key, uint32_t data, uint32_t* out_key, void* stream) {
  if (!key || !out_key) return FB200_E_NULL;
  fold_in_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(key, data, out_key);
  FB200_LAUNCH_CHECK();
  return 0;
}

extern "C" int fb200_sample_indices(const uint32_t* key, int n_draws, int n_samples, int32_t* out_idx,
                                    void* stream) {
  if (!key || !out_idx) return FB200_E_NULL;
  if (n_draws <= 0 || n_samples <= 0 || n_draws >= (1 << 30)) return FB200_E_SHAPE;
  sample_indices_kernel<true><<<(n_draws + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      key, (uint32_t)n_draws, (uint32_t)n_samples, out_idx);
  FB200_LAUNCH_CHECK();
  return 0;
}

extern "C" int fb200_choice_from_keys(const uint32_t* keys, int n_keys, int n_samples, int32_t* out_idx,
                                      void* stream) {
  if (!keys || !out_idx) return FB200_E_NULL;
  if (n_keys <= 0 || n_samples <= 0 || n_keys >= (1 << 30)) return FB200_E_SHAPE;
  sample_indices_kernel<false><<<(n_keys + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      keys, (uint32_t)n_keys, (uint32_t)n_samples, out_idx);
  FB200_LAUNCH_CHECK();
  return 0;
}
