"""BASELINE.json's full size, c4 = z[64, 65536, 1000] float32 (16.8 GB: byte offsets beyond 2^32, row tiles the small
parity cases never reach): size-independent properties over ALL rows, and the oracle on row windows taken from the
start, the 2^31- and 2^32-byte boundaries of a sample's slab, the middle and the end."""
import numpy as np
import pytest
import torch

from oracle import predictive as P
from oracle import scoring as Sc

pytestmark = pytest.mark.gpu

S, B, C = 64, 65536, 1000
NAMES = ("mean", "aleatoric_variance", "epistemic_variance", "variance", "entropy", "aleatoric_entropy", "epistemic_entropy",
         "log_prob", "mode", "confidence", "brier_terms")


@pytest.fixture(scope="module")
def c4(cuda):
    free, _ = torch.cuda.mem_get_info(cuda)
    if free < 40e9:
        pytest.skip("needs 40 GB of free device memory")
    gen = torch.Generator(device=cuda)
    gen.manual_seed(7)
    z = torch.empty((S, B, C), dtype=torch.float32, device=cuda)
    for s in range(S):
        z[s].normal_(0.0, 4.0, generator=gen)
    t = torch.randint(0, C, (B,), device=cuda, generator=gen, dtype=torch.int32)
    from fortuna_b200 import ops

    outs = ops.bma_reduce_cls(z, NAMES, targets=t)
    torch.cuda.synchronize()
    yield z, t, outs
    del z, outs
    torch.cuda.empty_cache()


def _windows():
    # rows whose first byte inside a sample's [B, C] slab sits at 0, just below / above 2^31 bytes is beyond B*C*4 = 2.6e8, so
    # the interesting boundaries are those of the WHOLE tensor: sample s, row b starts at (s B + b) C 4 bytes
    w = [0, 17, B // 2 - 8, B - 24]
    for target in (2 ** 31, 2 ** 32):  # rows of sample 8 / 16 that straddle these byte offsets also exist in every window
        b = (target // (C * 4)) % B
        w.append(max(0, min(B - 24, b - 8)))
    return sorted(set(w))


def test_reductions_match_the_oracle_on_row_windows(c4):
    z, t, outs = c4
    for b0 in _windows():
        zs = z[:, b0 : b0 + 24].cpu().numpy()
        ts = t[b0 : b0 + 24].cpu().numpy()
        g = {k: v[b0 : b0 + 24].cpu().numpy() for k, v in outs.items()}
        mean = P.cls_mean(zs)
        np.testing.assert_allclose(g["mean"], mean, rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(g["aleatoric_variance"], P.cls_aleatoric_variance(zs), rtol=1e-5, atol=2e-7)
        ev = 4e-7 * float(np.max(mean) ** 2) + 2e-7
        np.testing.assert_allclose(g["epistemic_variance"], P.cls_epistemic_variance(zs), rtol=1e-5, atol=ev)
        np.testing.assert_allclose(g["entropy"], P.cls_entropy(zs), rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(g["aleatoric_entropy"], P.cls_aleatoric_entropy(zs), rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(g["log_prob"], P.cls_log_prob(zs, ts), rtol=1e-5, atol=2e-6)
        assert np.array_equal(g["mode"], P.cls_mode(zs))


def test_identities_over_all_rows(c4):
    _, t, outs = c4
    mean = outs["mean"]
    assert bool(torch.isfinite(mean).all()) and float(mean.min()) >= 0.0
    np.testing.assert_allclose(mean.sum(1, dtype=torch.float64).cpu().numpy(), 1.0, rtol=0, atol=2e-6)
    tot = (outs["aleatoric_variance"].double() + outs["epistemic_variance"].double() - outs["variance"].double()).abs().max()
    assert float(tot) < 1e-6
    ent = (outs["aleatoric_entropy"].double() + outs["epistemic_entropy"].double() - outs["entropy"].double()).abs().max()
    assert float(ent) < 1e-5
    assert float(outs["epistemic_entropy"].min()) > -1e-5  # Jensen: entropy of the mean >= mean entropy
    # mode / confidence are the arg-max and max of the mean; log_prob <= 0
    assert torch.equal(outs["mode"].long(), mean.argmax(1)) and torch.equal(outs["confidence"], mean.max(1).values)
    assert float(outs["log_prob"].max()) <= 0.0
    # the Brier terms against the definition, in float64
    oh = torch.nn.functional.one_hot(t.long(), C).double()
    want = ((mean.double() - oh) ** 2).sum(1)
    np.testing.assert_allclose(outs["brier_terms"].cpu().numpy(), want.cpu().numpy(), rtol=1e-5, atol=1e-7)


def test_conformal_sets_and_ece_at_full_size(c4, cuda):
    """select kernel == the path that sorts every row, sizes == mask row sums, sets grow as the error shrinks; the oracle on
    row windows; ECE bin counts add up to B."""
    from fortuna_b200 import ops

    _, t, outs = c4
    mean = outs["mean"]
    n_val = 50_000
    val_scores = ops.aps_scores(mean[:n_val].contiguous(), t[:n_val].contiguous())
    vs = ops.sort_f32_(val_scores.clone())
    prev = None
    for error in (0.2, 0.05):
        thr = Sc.conformal_threshold(n_val, error)
        mask, sizes = ops.aps_conformal_sets(mean, vs, thr)
        m2, s2, scores = ops.aps_conformal_sets(mean, vs, thr, want_scores=True)
        assert torch.equal(mask, m2) and torch.equal(sizes, s2)
        assert torch.equal(mask.sum(1, dtype=torch.int32), sizes)
        if prev is not None:
            assert bool((mask >= prev).all())  # a smaller error never removes a class
        prev = mask
        for b0 in _windows():
            p = mean[b0 : b0 + 24].cpu().numpy()
            conds, sz, _ = Sc.conformal_sets(val_scores.cpu().numpy(), Sc.aps_cumsums(p), error)
            assert np.array_equal(mask[b0 : b0 + 24].cpu().numpy().astype(bool), conds)
            assert np.array_equal(sizes[b0 : b0 + 24].cpu().numpy(), sz)
    thr_e = torch.from_numpy(Sc.ece_thresholds(B)).to(cuda)
    e = ops.ece_bins_from_rows(outs["confidence"], outs["mode"], t, thr_e, brier_terms=outs["brier_terms"])
    assert int(e["counts"].sum()) == B
