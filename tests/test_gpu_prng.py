"""Bit-exact jax.random stream on the device vs the oracle."""
import numpy as np
import pytest
import torch

from oracle import prng as oprng

pytestmark = pytest.mark.gpu


def _u32(t):
    return t.detach().cpu().numpy().view(np.uint32)


@pytest.mark.parametrize("seed", [0, 42, 2**40 + 7])
@pytest.mark.parametrize("n", [1, 2, 3, 4, 7, 1000, 1001, 65537])
def test_random_bits_bit_exact(cuda, seed, n):
    from fortuna_b200 import ops

    key = ops.PRNGKey(seed, cuda)
    got = _u32(ops.random_bits(key, n))
    want = oprng.random_bits(oprng.PRNGKey(seed), n)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("num", [1, 2, 3, 8, 203])
def test_split_and_fold_in_bit_exact(cuda, num):
    from fortuna_b200 import ops

    key = ops.PRNGKey(7, cuda)
    assert np.array_equal(_u32(ops.split(key, num)), oprng.split(oprng.PRNGKey(7), num))
    assert np.array_equal(_u32(ops.fold_in(key, num)), oprng.fold_in(oprng.PRNGKey(7), num))


def test_known_answers(cuda):
    from fortuna_b200 import ops

    k0 = ops.PRNGKey(0, cuda)
    assert _u32(ops.split(k0)).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]
    assert _u32(ops.fold_in(k0, 1)).tolist() == [928981903, 3453687069]
    # values printed in the JAX documentation
    assert abs(float(ops.random_normal(k0, 1)[0]) - (-0.20584226)) < 2e-7
    np.testing.assert_allclose(
        ops.random_normal(k0, 3).cpu().numpy(), [1.8160863, -0.48262316, 0.33988908], rtol=2e-6
    )
    assert abs(float(ops.random_normal(ops.PRNGKey(42, cuda), 1)[0]) - (-0.18471177)) < 2e-7


@pytest.mark.parametrize("n", [1, 5, 4096, 1_000_003])
def test_normal_matches_oracle(cuda, n):
    from fortuna_b200 import ops

    got = ops.random_normal(ops.PRNGKey(3, cuda), n).cpu().numpy()
    want = oprng.normal(oprng.PRNGKey(3), n)
    # uint32 stream is bit-exact (test above); erf_inv differs from XLA-CPU / numpy libm by rounding only
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("n_draws,n_samples", [(1, 1), (30, 10), (64, 100), (257, 3), (5, 2**31 - 1)])
def test_sample_indices_bit_exact(cuda, n_draws, n_samples):
    from fortuna_b200 import ops

    key = ops.PRNGKey(11, cuda)
    got = ops.sample_indices(key, n_draws, n_samples).cpu().numpy()
    keys = oprng.split(oprng.PRNGKey(11), n_draws)
    want = np.array([oprng.choice(k, n_samples) for k in keys], dtype=np.int64)
    assert np.array_equal(got.astype(np.int64), want)


@pytest.mark.parametrize("n_keys,n_samples", [(1, 10), (17, 100), (3, 1)])
def test_choice_from_keys_bit_exact(cuda, n_keys, n_samples):
    """posterior.sample(rng): idx = choice(rng, n_samples) with the caller's key (no split over draws)."""
    from fortuna_b200 import ops

    keys_o = oprng.split(oprng.PRNGKey(5), n_keys)
    keys = ops.split(ops.PRNGKey(5, cuda), n_keys)
    got = ops.sample_indices_from_keys(keys, n_samples).cpu().numpy()
    want = np.array([oprng.choice(k, n_samples) for k in keys_o], dtype=np.int64)
    assert np.array_equal(got.astype(np.int64), want)
