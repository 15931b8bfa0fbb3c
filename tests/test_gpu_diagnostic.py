"""fb200_ksd_imq / fb200_ess against the oracle (oracle/diagnostic.py) and its float64 evaluation, with the host
mirrors called the way tests/fortuna/prob_model/test_diagnostic.py:42-69 calls the reference."""
import numpy as np
import pytest
import torch

from oracle import diagnostic as D

pytestmark = pytest.mark.gpu

MU = np.array([0.0, 0.0])
SIGMA = np.array([[1.5, 0.5], [0.5, 1.5]])


def _grad_log_density(x):
    return -(x - MU) @ np.linalg.inv(SIGMA).T


def _ksd_tol(samples, grads, c, beta):
    """Float64 yardstick + the absolute slack float32 pair sums allow: the S^2 terms of mixed sign are each good to
    ~1e-6 relative, so sum k0 carries an error of ~1e-6 * sum |k0| (the total itself can be much smaller)."""
    k = D.ksd_pair_matrix(samples, grads, c, beta, dtype=np.float64)
    total, S = k.sum(), k.shape[0]
    return np.sqrt(total) / S, total, 4e-6 * np.abs(k).sum()


@pytest.mark.parametrize("S,N", [(1000, 2), (3, 5), (64, 1000), (65, 33), (130, 4099), (30, 44426), (1, 7)])
@pytest.mark.parametrize("c,beta", [(1.0, -0.5), (0.3, -1.5)])
def test_ksd_imq_matches_oracle(cuda, S, N, c, beta):
    from fortuna_b200 import ops

    r = np.random.default_rng(S * 31 + N)
    centre = r.standard_normal(N).astype(np.float32)
    x = (centre + 0.05 * r.standard_normal((S, N))).astype(np.float32)  # samples close together, far from the origin
    g = (-(x - centre) / 0.05**2 * 0.01 + 0.1 * r.standard_normal((S, N))).astype(np.float32)
    got = float(ops.ksd_imq(torch.from_numpy(x).to(cuda), torch.from_numpy(g).to(cuda), c, beta).item())
    want, total, slack = _ksd_tol(x, g, c, beta)
    assert abs(got**2 * S * S - total) <= slack + 1e-5 * abs(total)
    if S * N <= 200_000:  # the float32 restatement (slow python loop over i) agrees within the same slack
        o32 = float(D.kernel_stein_discrepancy_imq(x, g, c, beta))
        assert abs(o32**2 * S * S - total) <= 4 * slack + 1e-4 * abs(total)


def test_ksd_on_strided_bank_rows_and_reference_test_case(cuda):
    """tests/fortuna/prob_model/test_diagnostic.py:42-57: exact samples score lower than over-dispersed ones, pytree
    input equals flat input; plus rows taken straight out of a wider matrix (the sample bank with padding)."""
    from fortuna_b200.prob_model.posterior.sgmcmc.sgmcmc_diagnostic import kernel_stein_discrepancy_imq

    rng = np.random.default_rng(0)
    s1 = rng.multivariate_normal(MU, SIGMA, size=1000)
    s2 = rng.multivariate_normal(MU, SIGMA**3, size=1000)
    g1, g2 = _grad_log_density(s1), _grad_log_density(s2)
    k1 = kernel_stein_discrepancy_imq(s1, g1)
    k2 = kernel_stein_discrepancy_imq(s2, g2)
    assert float(k1) < float(k2)
    tree = lambda a: [{"x": v[0], "y": v[1]} for v in a]  # noqa: E731
    assert torch.allclose(k1, kernel_stein_discrepancy_imq(tree(s1), tree(g1)))
    np.testing.assert_allclose(float(k1), D.kernel_stein_discrepancy_imq(s1, g1, dtype=np.float64), rtol=1e-5)
    wide_x = torch.zeros((1000, 8), device=cuda)
    wide_g = torch.zeros((1000, 8), device=cuda)
    wide_x[:, :2] = torch.from_numpy(s1.astype(np.float32)).to(cuda)
    wide_g[:, :2] = torch.from_numpy(g1.astype(np.float32)).to(cuda)
    assert torch.equal(kernel_stein_discrepancy_imq(wide_x[:, :2], wide_g[:, :2]), k1)
    with pytest.raises(ValueError):
        kernel_stein_discrepancy_imq(s1, g1, c=0.0)
    with pytest.raises(ValueError):
        kernel_stein_discrepancy_imq(s1, g1, beta=0.5)


@pytest.mark.parametrize("S,N", [(200, 1), (1000, 2), (257, 40), (30, 1082), (100, 5000), (7, 33), (900, 70), (2500, 9)])
@pytest.mark.parametrize("thr", [0.0, 0.05])
def test_ess_matches_oracle(cuda, S, N, thr):
    from fortuna_b200 import ops

    r = np.random.default_rng(S + N)
    ar = np.zeros((S, N))
    e = r.standard_normal((S, N))
    rho = r.uniform(0.0, 0.95, N)
    for t in range(1, S):
        ar[t] = rho * ar[t - 1] + e[t]  # AR(1) chains with different mixing speeds
    x = (3.0 + ar).astype(np.float32)
    got = ops.ess(torch.from_numpy(x).to(cuda), thr).cpu().numpy()
    want = D.effective_sample_size_direct(x, thr)
    ok = D.ess_threshold_margin(x, thr) > 1e-4  # columns whose cut-off lag does not hinge on float32 rounding
    assert ok.mean() > 0.9
    np.testing.assert_allclose(got[ok], want[ok], rtol=2e-4)
    # the reference's FFT route in float32 lands on the same numbers
    np.testing.assert_allclose(D.effective_sample_size(x, thr)[ok], got[ok], rtol=5e-4)


def test_ess_reference_test_case_and_notebook(cuda):
    """test_diagnostic.py:59-69 and examples/sgmcmc_diagnostics.pct.py:122-126."""
    from fortuna_b200.prob_model.posterior.sgmcmc.sgmcmc_diagnostic import effective_sample_size

    rng = np.random.default_rng(0)
    s = rng.multivariate_normal(MU, SIGMA, size=1000)
    ess = effective_sample_size(s)
    assert ess.shape == (2,) and bool((ess >= 0).all()) and bool((ess <= len(s)).all())
    tree = effective_sample_size([{"x": v[0], "y": v[1]} for v in s])
    assert torch.allclose(torch.stack([tree["x"].reshape(()), tree["y"].reshape(())]), ess)
    white = effective_sample_size(rng.normal(size=200))
    trend = effective_sample_size(np.arange(200) + rng.normal(size=200))
    assert 100 < float(white[0]) <= 200 and float(trend[0]) < 10
    const = effective_sample_size(np.ones((50, 3), np.float32))
    assert bool(torch.isnan(const).all())  # 0 / 0 auto-correlation, as in the reference


def test_ess_of_the_sample_bank(cuda):
    """FlatTree samples (what the sampling callback stores): padding between leaves is not a parameter."""
    from fortuna_b200.prob_model.posterior.sgmcmc.sgmcmc_diagnostic import effective_sample_size
    from fortuna_b200.utils.flat_tree import FlatTree

    r = np.random.default_rng(3)
    trees = [{"a": {"kernel": r.standard_normal((3, 5)).astype(np.float32), "bias": r.standard_normal(5).astype(np.float32)},
              "b": r.standard_normal(7).astype(np.float32)} for _ in range(40)]
    fts = [FlatTree.from_tree(t, device=cuda) for t in trees]
    got = effective_sample_size(fts)
    want = D.effective_sample_size_direct(trees, 0.0)
    flat = torch.cat([got["a"]["bias"].reshape(-1), got["a"]["kernel"].reshape(-1), got["b"].reshape(-1)]).cpu().numpy()
    ok = D.ess_threshold_margin(trees, 0.0) > 1e-4
    np.testing.assert_allclose(flat[ok], want[ok], rtol=2e-4)
    assert got["a"]["kernel"].shape == (3, 5)


@pytest.mark.parametrize("S", [1, 2, 3, 31, 32, 33, 63, 64])
@pytest.mark.parametrize("N", [1, 255, 300_001])
def test_ess_short_chain_kernel_edges(cuda, S, N):
    """The register-resident kernel (S <= 64: prefetch ring, packed lag products, warp-level early exit) at the ends of its
    ranges: chains that fill / do not fill the 32- and 64-sample instantiations, one parameter, a ragged last warp, more
    parameters than one pass of the persistent grid, rows strided inside a wider matrix, with and without the cut-off."""
    from fortuna_b200 import ops

    r = np.random.default_rng(S * 7 + N % 1000)
    ar = np.zeros((S, N), np.float32)
    e = r.standard_normal((S, N)).astype(np.float32)
    rho = r.uniform(0.0, 0.9, N).astype(np.float32)
    for t in range(1, S):
        ar[t] = rho * ar[t - 1] + e[t]
    x = (1.0 + ar).astype(np.float32)
    wide = torch.zeros((S, N + 5), device=cuda)
    wide[:, 2 : N + 2] = torch.from_numpy(x).to(cuda)
    for thr in (0.0, None):
        got = ops.ess(wide[:, 2 : N + 2], thr).cpu().numpy()
        assert got.shape == (N,)
        if S < 3:
            continue  # S = 1: 0 / 0; S = 2: the two lags cancel exactly - NaN / inf / rounding noise in the reference too
        if thr is None:
            continue  # every lag kept: the denominator is rounding noise (see test_oracle_diagnostic), shapes only
        want = D.effective_sample_size_direct(x, thr)
        ok = D.ess_threshold_margin(x, thr) > 1e-4
        np.testing.assert_allclose(got[ok], want[ok], rtol=3e-4)
    if S == 1:
        assert np.isnan(ops.ess(wide[:, 2 : N + 2], 0.0).cpu().numpy()).all()


@pytest.mark.parametrize("S", [2, 5, 17, 31, 32])
@pytest.mark.parametrize("N", [1, 111, 113, 100_003])
def test_ksd_triangle_kernel_edges(cuda, S, N):
    """Chains of <= 32 samples take the triangle kernel (4 x 4 pair blocks bi <= bj): block edges inside and at the end of
    the tile, fewer columns than one 112-column stage, a ragged last stage, several splits, strided rows."""
    from fortuna_b200 import ops

    r = np.random.default_rng(S * 13 + N % 1000)
    centre = r.standard_normal(N).astype(np.float32)
    x = (centre + 0.05 * r.standard_normal((S, N))).astype(np.float32)
    g = (-(x - centre) / 0.05**2 * 0.01 + 0.1 * r.standard_normal((S, N))).astype(np.float32)
    wx = torch.zeros((S, N + 3), device=cuda)
    wg = torch.zeros((S, N + 3), device=cuda)
    wx[:, 1 : N + 1] = torch.from_numpy(x).to(cuda)
    wg[:, 1 : N + 1] = torch.from_numpy(g).to(cuda)
    got = float(ops.ksd_imq(wx[:, 1 : N + 1], wg[:, 1 : N + 1]).item())
    want, total, slack = _ksd_tol(x, g, 1.0, -0.5)
    assert abs(got**2 * S * S - total) <= slack + 1e-5 * abs(total)
