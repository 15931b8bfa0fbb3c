"""Pins oracle/prng.py against every known answer available for the path:
the Random123 threefry2x32 KAT and the values printed in JAX's own documentation."""
import numpy as np

from oracle import prng


def test_random123_kat():
    y0, y1 = prng.threefry2x32(0x13198A2E, 0x03707344, np.array([0x243F6A88], np.uint32), np.array([0x85A308D3], np.uint32))
    assert (int(y0[0]), int(y1[0])) == (0xC4923A9C, 0x483DF7A0)
    # all-zero and all-one vectors of the Random123 kat_vectors file (threefry2x32, 20 rounds)
    y0, y1 = prng.threefry2x32(0, 0, np.array([0], np.uint32), np.array([0], np.uint32))
    assert (int(y0[0]), int(y1[0])) == (0x6B200159, 0x99BA4EFE)
    y0, y1 = prng.threefry2x32(0xFFFFFFFF, 0xFFFFFFFF, np.array([0xFFFFFFFF], np.uint32), np.array([0xFFFFFFFF], np.uint32))
    assert (int(y0[0]), int(y1[0])) == (0x1CB996FC, 0xBB002BE7)


def test_jax_documented_values():
    k0 = prng.PRNGKey(0)
    assert k0.tolist() == [0, 0]
    assert prng.split(k0).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]
    assert prng.uniform(k0, 1)[0] == np.float32(0.41845703)
    assert prng.normal(k0, 1)[0] == np.float32(-0.20584226)
    np.testing.assert_array_equal(prng.normal(k0, 3), np.array([1.8160863, -0.48262316, 0.33988908], np.float32))
    k42 = prng.PRNGKey(42)
    assert prng.normal(k42, 1)[0] == np.float32(-0.18471177)
    new_key, subkey = prng.split(k42)
    assert new_key.tolist() == [2465931498, 3679230171] and subkey.tolist() == [255383827, 267815257]
    assert prng.normal(subkey, 1)[0] == np.float32(1.3694694)
    assert prng.split(k42, 3).tolist() == [[3134548294, 3733159049], [3746501087, 894150801], [801545058, 2363201431]]
    assert prng.fold_in(k0, 1).tolist() == [928981903, 3453687069]


def test_bits_layout_odd_even():
    key = prng.PRNGKey(3)
    for n in (1, 2, 3, 8, 9):
        bits = prng.random_bits(key, n)
        h = (n + 1) // 2
        for e in range(n):
            x0 = e if e < h else e - h
            x1 = (x0 + h) if (x0 + h) < n else 0
            y = prng.threefry2x32(key[0], key[1], np.array([x0], np.uint32), np.array([x1], np.uint32))
            assert int(bits[e]) == int(y[0][0] if e < h else y[1][0])


def test_normal_is_standard_and_erfinv_accurate():
    from scipy.special import erfinv

    x = prng.normal(prng.PRNGKey(1), 200_000)
    assert abs(float(x.mean())) < 0.01 and abs(float(x.std()) - 1.0) < 0.01
    u = np.linspace(-0.999, 0.999, 10001).astype(np.float32)
    np.testing.assert_allclose(prng.erf_inv_f32(u), erfinv(u.astype(np.float64)), rtol=2e-6, atol=1e-7)
    assert np.isinf(prng.erf_inv_f32(np.array([1.0, -1.0], np.float32))).all()


def test_choice_range_and_determinism():
    idx = [prng.choice(k, 10) for k in prng.split(prng.PRNGKey(5), 500)]
    assert min(idx) == 0 and max(idx) == 9
    assert idx == [prng.choice(k, 10) for k in prng.split(prng.PRNGKey(5), 500)]
    counts = np.bincount(idx, minlength=10)
    assert counts.min() > 20  # roughly uniform
    assert prng.choice(prng.PRNGKey(0), 1) == 0


def test_rng_holder_chain():
    r = prng.RandomNumberGenerator(0)
    assert r.get().tolist() == [4146024105, 967050713]
    assert r.get().tolist() == prng.split(np.array([4146024105, 967050713], np.uint32))[0].tolist()
