"""jax.random (jax 0.4.30, default ``threefry2x32``, ``jax_threefry_partitionable=False``)
restated in numpy.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference calls these through
  fortuna/utils/random.py:11-23            split(key, num_leaves) + normal per leaf
  fortuna/utils/random.py:38-49            RandomNumberGenerator.get = split(rng)[0]
  .../sghmc/sghmc_integrator.py:66         key, new_key = split(rng_key)
  .../sgmcmc/sgmcmc_posterior.py:51        random.choice(rng, n_samples)
  fortuna/prob_model/predictive/base.py:113,188,632   keys = split(rng, S)
  fortuna/training/trainer.py:680          fold_in(rng, step)
The algorithm is jax/_src/prng.py (threefry2x32, _threefry_split_original,
_threefry_random_bits_original, threefry_fold_in) and jax/_src/random.py
(_uniform, _normal_real, _randint, choice) at the pinned version; XLA's f32
``ErfInv`` (Giles' single-precision polynomial).  The uint32 stream is bit-exact
(pinned by the Random123 / JAX known answers in tests/test_oracle_prng.py); float
outputs are float32 op-by-op.
"""

import numpy as np

U32 = np.uint32
F32 = np.float32

_ROT_A = (13, 15, 26, 6)
_ROT_B = (17, 29, 16, 24)


def PRNGKey(seed: int) -> np.ndarray:
    """jax.random.PRNGKey: [hi32(seed), lo32(seed)] as uint32[2]."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=U32)


def _rotl(x, r):
    return (x << U32(r)) | (x >> U32(32 - r))


def threefry2x32(k0, k1, x0, x1):
    """Threefry-2x32, 20 rounds (jax/_src/prng.py ``_threefry2x32_lowering``).

    k0, k1: uint32 scalars; x0, x1: uint32 arrays of the same shape.
    Returns (y0, y1).
    """
    with np.errstate(over="ignore"):
        k0 = U32(k0)
        k1 = U32(k1)
        ks = (k0, k1, U32(k0 ^ k1 ^ U32(0x1BD11BDA)))
        x0 = np.asarray(x0, dtype=U32).copy()
        x1 = np.asarray(x1, dtype=U32).copy()
        x0 += ks[0]
        x1 += ks[1]
        for g in range(5):
            for r in (_ROT_A if g % 2 == 0 else _ROT_B):
                x0 += x1
                x1 = _rotl(x1, r)
                x1 ^= x0
            x0 += ks[(g + 1) % 3]
            x1 += U32(ks[(g + 2) % 3] + U32(g + 1))
    return x0, x1


def _threefry_2x32_counts(key, count):
    """jax ``threefry_2x32(keypair, count)``: counters split first-half / second-half."""
    count = np.asarray(count, dtype=U32).ravel()
    n = count.size
    if n % 2:
        count = np.concatenate([count, np.zeros(1, U32)])
    h = count.size // 2
    y0, y1 = threefry2x32(key[0], key[1], count[:h], count[h:])
    return np.concatenate([y0, y1])[:n]


def random_bits(key, n: int) -> np.ndarray:
    """32-bit ``_threefry_random_bits_original`` for ``n`` < 2**32 - 1 values."""
    n = int(n)
    if n >= 0xFFFFFFFF:
        raise ValueError("oracle random_bits restates the single-block case only (n < 2**32-1)")
    return _threefry_2x32_counts(key, np.arange(n, dtype=U32))


def split(key, num: int = 2) -> np.ndarray:
    """``_threefry_split_original``: random_bits(key, 2*num).reshape(num, 2)."""
    return random_bits(key, 2 * int(num)).reshape(int(num), 2)


def fold_in(key, data: int) -> np.ndarray:
    """``threefry_fold_in``: threefry_2x32(key, threefry_seed(data)) with seed = [0, data]."""
    return _threefry_2x32_counts(key, np.array([0, int(data) & 0xFFFFFFFF], dtype=U32))


def bits_to_unit_float(bits: np.ndarray) -> np.ndarray:
    """bitcast((bits >> 9) | 0x3F800000) - 1.0  in [0, 1)  (jax/_src/random.py ``_uniform``)."""
    fb = (np.asarray(bits, dtype=U32) >> U32(9)) | U32(0x3F800000)
    return fb.view(F32) - F32(1.0)


def uniform(key, n: int, minval=0.0, maxval=1.0) -> np.ndarray:
    """``_uniform`` float32: max(minval, f * (maxval - minval) + minval)."""
    lo = F32(minval)
    hi = F32(maxval)
    f = bits_to_unit_float(random_bits(key, n))
    return np.maximum(lo, f * F32(hi - lo) + lo)


_ERFINV_LT5 = tuple(
    F32(c)
    for c in (
        2.81022636e-08,
        3.43273939e-07,
        -3.5233877e-06,
        -4.39150654e-06,
        0.00021858087,
        -0.00125372503,
        -0.00417768164,
        0.246640727,
        1.50140941,
    )
)
_ERFINV_GE5 = tuple(
    F32(c)
    for c in (
        -0.000200214257,
        0.000100950558,
        0.00134934322,
        -0.00367342844,
        0.00573950773,
        -0.0076224613,
        0.00943887047,
        1.00167406,
        2.83297682,
    )
)


def erf_inv_f32(x: np.ndarray) -> np.ndarray:
    """XLA ``ErfInv32`` (xla/client/lib/math.cc; M. Giles' polynomial), float32 op order:
    w = -log1p(-(x*x)); two Horner branches selected on w < 5; result p*x; +-inf at |x| = 1.
    Note x*x is rounded to float32 BEFORE the log1p (no fused 1 - x*x)."""
    x = np.asarray(x, dtype=F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = x * x
        w = -np.log1p(-t)
        lt = w < F32(5.0)
        wv = np.where(lt, w - F32(2.5), np.sqrt(w) - F32(3.0)).astype(F32)
        p = np.where(lt, _ERFINV_LT5[0], _ERFINV_GE5[0]).astype(F32)
        for i in range(1, 9):
            p = (np.where(lt, _ERFINV_LT5[i], _ERFINV_GE5[i]).astype(F32) + p * wv).astype(F32)
        res = (p * x).astype(F32)
        res = np.where(np.abs(x) == F32(1.0), x * F32(np.inf), res).astype(F32)
    return res


_NORMAL_LO = np.nextafter(F32(-1.0), F32(0.0), dtype=F32)
_SQRT2 = F32(np.sqrt(2))


def normal_from_bits(bits: np.ndarray) -> np.ndarray:
    """bits -> N(0,1) float32 exactly as ``_normal_real``: u = uniform(lo=nextafter(-1,0), hi=1);
    sqrt(2) * erf_inv(u)."""
    f = bits_to_unit_float(bits)
    u = np.maximum(_NORMAL_LO, f * F32(F32(1.0) - _NORMAL_LO) + _NORMAL_LO)
    return (_SQRT2 * erf_inv_f32(u)).astype(F32)


def normal(key, n: int) -> np.ndarray:
    """jax.random.normal(key, (n,), float32) (flattened; shape is a reshape of this)."""
    return normal_from_bits(random_bits(key, n))


def randint(key, minval: int, maxval: int) -> int:
    """Scalar ``_randint`` (int32): two 32-bit draws reduce the modulo bias.
    NOT pinned by any known answer (restated from jax/_src/random.py at 0.4.30)."""
    k1, k2 = split(key, 2)
    hi = int(random_bits(k1, 1)[0])
    lo = int(random_bits(k2, 1)[0])
    span = (int(maxval) - int(minval)) & 0xFFFFFFFF
    if maxval <= minval:
        span = 1
    mult = (1 << 16) % span
    mult = ((mult * mult) & 0xFFFFFFFF) % span  # lax.mul on uint32 wraps (spans > 2**16 give multiplier 0)
    off = (((hi % span) * mult) & 0xFFFFFFFF) + (lo % span)
    off = (off & 0xFFFFFFFF) % span
    return int(minval) + off


def choice(key, n: int) -> int:
    """jax.random.choice(key, n) with shape=(), replace=True, p=None == randint(key, (), 0, n)
    (call site: .../sgmcmc/sgmcmc_posterior.py:51)."""
    return randint(key, 0, int(n))


class RandomNumberGenerator:
    """fortuna/utils/random.py:26-49 - stateful key chain: get() = split(rng)[0]."""

    def __init__(self, seed: int):
        self._rng = PRNGKey(seed)

    def get(self) -> np.ndarray:
        self._rng = split(self._rng)[0]
        return self._rng
