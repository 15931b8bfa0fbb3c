"""SG-MCMC diagnostics restated in numpy.  TEST INFRASTRUCTURE ONLY.

Follows fortuna/prob_model/posterior/sgmcmc/sgmcmc_diagnostic.py:
  kernel_stein_discrepancy_imq :16-70   (``_k_0`` term by term, vmap over the second sample, scan over the first)
  effective_sample_size        :73-140  (centre, zero-pad to the next power of two >= 2n, FFT power spectrum -> inverse
                                         FFT, unbiased normalisation, cumulative threshold mask, n / (-1 + 2 sum))
in float32 like the reference's default dtype (``dtype=np.float64`` gives the high-precision evaluation the tests use as
the yardstick).  ``effective_sample_size_direct`` is the same quantity from explicit lagged products - what the FFT
computes up to rounding, and what the CUDA kernel does.

Parity unpinned: the reference's tests hold orderings and shapes only (tests/fortuna/prob_model/test_diagnostic.py:42-69:
KSD of exact samples < KSD of over-dispersed samples, 0 <= ESS <= n, tree input == flat input); those are re-stated in
tests/test_oracle_diagnostic.py.  numpy's pocketfft in complex64 stands in for XLA-CPU's FFT.
"""
import numpy as np


def ravel_samples(samples):
    """ravel_pytree(samples)[0].reshape(len(samples), -1) for a list of nested dicts (sorted keys) or an array."""
    def flat(t):
        if isinstance(t, dict):
            return np.concatenate([flat(t[k]) for k in sorted(t)])
        return np.asarray(t).reshape(-1)

    if isinstance(samples, (list, tuple)):
        return np.stack([flat(s) for s in samples])
    a = np.asarray(samples)
    return a.reshape(a.shape[0], -1)


def ksd_pair_matrix(samples, grads, c=1.0, beta=-0.5, dtype=np.float32):
    """k0[i, j] (:52-61)."""
    x = ravel_samples(samples).astype(dtype)
    g = ravel_samples(grads).astype(dtype)
    S, dim = x.shape
    c, beta = dtype(c), dtype(beta)
    out = np.empty((S, S), dtype=dtype)
    for i in range(S):
        diff = (x[i][None, :] - x).astype(dtype)                      # param1 - param2, param2 = every sample
        dd = np.einsum("jn,jn->j", diff, diff).astype(dtype)
        base = (c**2 + dd).astype(dtype)
        kern = (g @ g[i]).astype(dtype) * base**beta
        kern = kern + dtype(-2) * beta * (diff @ g[i]).astype(dtype) * base ** (beta - dtype(1))
        kern = kern + dtype(2) * beta * np.einsum("jn,jn->j", g, diff).astype(dtype) * base ** (beta - dtype(1))
        kern = kern + dtype(-2 * dim) * beta * base ** (beta - dtype(1))
        kern = kern + dtype(-4) * beta * (beta - dtype(1)) * base ** (beta - dtype(2)) * dd
        out[i] = kern.astype(dtype)
    return out


def kernel_stein_discrepancy_imq(samples, grads, c=1.0, beta=-0.5, dtype=np.float32):
    if not c > 0:
        raise ValueError("`c` should be > 0.")
    if not beta < 0:
        raise ValueError("`beta` should be < 0.")
    k = ksd_pair_matrix(samples, grads, c, beta, dtype)
    total = dtype(0)
    for i in range(k.shape[0]):  # lax.scan over the first sample, jnp.sum over the second
        total = dtype(total + k[i].sum(dtype=dtype))
    return dtype(np.sqrt(total) / dtype(k.shape[0]))


def _finish_ess(auto_cov, n, filter_threshold, dtype):
    auto_corr = (auto_cov / auto_cov[:1]).astype(dtype)
    nk = ((n - np.arange(0.0, n)) / n).astype(dtype)
    w = (nk[:, None] * auto_corr).astype(dtype)
    if filter_threshold is not None:
        mask = (auto_corr < dtype(filter_threshold)).astype(dtype)
        mask = np.cumsum(mask, axis=0, dtype=dtype)
        mask = np.maximum(dtype(1) - mask, dtype(0))
        w = (w * mask).astype(dtype)
    return (dtype(n) / (dtype(-1) + dtype(2) * w.sum(axis=0, dtype=dtype))).astype(dtype)


def effective_sample_size(samples, filter_threshold=0.0, dtype=np.float32):
    """FFT route, as the reference (:97-139).  Returns the flat [n_params] vector."""
    x = ravel_samples(samples).astype(dtype).T.copy()  # [N, S]: the reference moves the chain axis last
    n = x.shape[-1]
    x = x - x.mean(axis=-1, keepdims=True, dtype=dtype)
    pad_len = int(2.0 ** np.ceil(np.log2(n * 2)) - n)
    x = np.pad(x, ((0, 0), (0, pad_len)))
    cdt = np.complex64 if dtype == np.float32 else np.complex128
    f = np.fft.fft(x.astype(cdt), axis=-1)
    prod = np.fft.ifft(f * np.conj(f), axis=-1)
    prod = np.real(prod[..., :n]).astype(dtype)
    auto_cov = (prod / (n - np.arange(0.0, n)).astype(dtype)).astype(dtype).T  # back to [S, N]
    return _finish_ess(auto_cov, n, filter_threshold, dtype)


def autocovariance_direct(samples, dtype=np.float64):
    x = ravel_samples(samples).astype(dtype)
    n = x.shape[0]
    x = x - x.mean(axis=0, keepdims=True, dtype=dtype)
    prod = np.stack([(x[: n - k] * x[k:]).sum(axis=0, dtype=dtype) for k in range(n)])
    return (prod / (n - np.arange(0.0, n))[:, None].astype(dtype)).astype(dtype)


def effective_sample_size_direct(samples, filter_threshold=0.0, dtype=np.float64):
    """Same estimator from explicit lagged products prod[k] = sum_t x[t] x[t+k] (what the zero-padded FFT computes)."""
    x = ravel_samples(samples)
    return _finish_ess(autocovariance_direct(x, dtype), x.shape[0], filter_threshold, dtype)


def ess_threshold_margin(samples, filter_threshold=0.0):
    """Per column: how far the auto-correlations that decide the cut are from the threshold (min |rho_k - thr| up to and
    including the first lag below it) - columns with a tiny margin may legitimately cut one lag apart in float32."""
    ac = autocovariance_direct(samples)
    rho = ac / ac[:1]
    below = rho < filter_threshold
    first = np.where(below.any(axis=0), below.argmax(axis=0), rho.shape[0] - 1)
    idx = np.arange(rho.shape[0])[:, None] <= first[None, :]
    return np.where(idx, np.abs(rho - filter_threshold), np.inf).min(axis=0)
