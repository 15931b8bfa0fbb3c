"""K3/K4 parity: single-pass BMA reductions vs oracle/predictive.py.

Tolerances (float32): rtol 1e-5 plus an absolute floor that reflects cancellation in the quantity
itself: epistemic variance = E[p^2] - E[p]^2 is compared with atol 4e-7 * max(E[p]^2) (both sides carry
half-ulp errors of the two subtracted terms); entropies with atol 2e-6 (sums of C terms of size <= 0.37).
"""
import numpy as np
import pytest
import torch

from oracle import predictive as P

pytestmark = pytest.mark.gpu

ALL = (
    "mean", "aleatoric_variance", "epistemic_variance", "variance", "std", "mode", "entropy",
    "aleatoric_entropy", "epistemic_entropy", "log_prob", "ensemble_log_prob",
)


def _logits(S, B, C, seed, ties=False, scale=4.0):
    z = (scale * np.random.default_rng(seed).standard_normal((S, B, C))).astype(np.float32)
    if ties:
        z = (np.round(z * 2) / 2).astype(np.float32)
    return z


def _check(outs, z, t, S):
    g = {k: v.cpu().numpy() for k, v in outs.items()}
    mean = P.cls_mean(z)
    np.testing.assert_allclose(g["mean"], mean, rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(g["aleatoric_variance"], P.cls_aleatoric_variance(z), rtol=1e-5, atol=2e-7)  # 1-p rounds at ulp(p)
    ev_atol = 4e-7 * float(np.max(mean) ** 2) + 2e-7
    np.testing.assert_allclose(g["epistemic_variance"], P.cls_epistemic_variance(z), rtol=1e-5, atol=ev_atol)
    np.testing.assert_allclose(g["variance"], P.cls_variance(z), rtol=1e-5, atol=ev_atol)
    np.testing.assert_allclose(g["std"] ** 2, P.cls_variance(z), rtol=2e-5, atol=ev_atol)
    np.testing.assert_allclose(g["entropy"], P.cls_entropy(z), rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(g["aleatoric_entropy"], P.cls_aleatoric_entropy(z), rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(g["epistemic_entropy"], P.cls_epistemic_entropy(z), rtol=1e-5, atol=4e-6)
    np.testing.assert_allclose(g["log_prob"], P.cls_log_prob(z, t), rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(g["ensemble_log_prob"], P.cls_ensemble_log_prob(z, t), rtol=1e-5, atol=2e-6)
    # mode: index-exact wherever the top-2 means are separated by more than float noise
    srt = np.sort(mean, axis=-1)
    clear = (srt[:, -1] - srt[:, -2]) > 1e-6 if mean.shape[1] > 1 else np.ones(mean.shape[0], bool)
    assert np.array_equal(g["mode"][clear], P.cls_mode(z)[clear])


@pytest.mark.parametrize(
    "S,B,C",
    [(8, 512, 1000), (64, 1024, 10), (30, 500, 2), (5, 517, 1000), (3, 37, 7), (4, 100, 33), (2, 19, 257),
     (6, 64, 128), (3, 41, 1024), (7, 9, 1), (4, 70, 4), (3, 25, 1001), (9, 130, 512), (2, 300, 64)],
)
def test_cls_reductions_f32(cuda, S, B, C):
    from fortuna_b200 import ops

    z = _logits(S, B, C, seed=S * 1000 + C)
    t = np.random.default_rng(1).integers(0, C, B).astype(np.int32)
    outs = ops.bma_reduce_cls(torch.from_numpy(z).to(cuda), ALL, targets=torch.from_numpy(t).to(cuda))
    _check(outs, z, t, S)


@pytest.mark.parametrize("S,B,C,temp", [(4, 33, 1025, False), (6, 40, 3000, True), (3, 17, 21841, False), (1030, 3, 1100, False)])
def test_cls_reductions_wide_heads(cuda, S, B, C, temp):
    """C > 1024 (rows wider than a warp's registers): the two-pass kernel fb200_bma_reduce_cls_wide, same outputs."""
    from fortuna_b200 import ops

    z = _logits(S, B, C, seed=S + C, scale=3.0)
    t = np.random.default_rng(1).integers(0, C, B).astype(np.int32)
    lt = np.array([0.25], dtype=np.float32)
    outs = ops.bma_reduce_cls(torch.from_numpy(z).to(cuda), ALL, targets=torch.from_numpy(t).to(cuda),
                              log_temp=torch.from_numpy(lt).to(cuda) if temp else None)
    zc = (z * np.exp(-lt[0], dtype=np.float32)).astype(np.float32) if temp else z
    _check(outs, zc, t, S)
    zb = torch.from_numpy(z).to(cuda).to(torch.bfloat16)
    ob = ops.bma_reduce_cls(zb, ("mean", "entropy", "log_prob"), targets=torch.from_numpy(t).to(cuda))
    zbn = zb.float().cpu().numpy()
    np.testing.assert_allclose(ob["mean"].cpu().numpy(), P.cls_mean(zbn), rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(ob["log_prob"].cpu().numpy(), P.cls_log_prob(zbn, t), rtol=1e-5, atol=2e-6)


def test_cls_reductions_ties_and_temperature(cuda):
    from fortuna_b200 import ops

    S, B, C = 8, 256, 1000
    z = _logits(S, B, C, seed=3, ties=True)
    t = np.random.default_rng(2).integers(0, C, B).astype(np.int32)
    lt = np.array([0.3], dtype=np.float32)
    outs = ops.bma_reduce_cls(
        torch.from_numpy(z).to(cuda), ALL, targets=torch.from_numpy(t).to(cuda), log_temp=torch.from_numpy(lt).to(cuda)
    )
    zc = (z * np.exp(-lt[0], dtype=np.float32)).astype(np.float32)
    _check(outs, zc, t, S)


@pytest.mark.parametrize("S,B,C", [(8, 256, 1000), (16, 100, 10), (4, 33, 6)])
def test_cls_reductions_bf16_input(cuda, S, B, C):
    from fortuna_b200 import ops

    z = P.round_to_bf16(_logits(S, B, C, seed=11))
    t = np.random.default_rng(4).integers(0, C, B).astype(np.int32)
    zt = torch.from_numpy(z).to(cuda).to(torch.bfloat16)
    outs = ops.bma_reduce_cls(zt, ALL, targets=torch.from_numpy(t).to(cuda))
    _check(outs, z, t, S)


def test_identities_and_edge_cases(cuda):
    from fortuna_b200 import ops

    # uniform logits -> entropy = log C, epistemic = 0; S = 1 -> epistemic variance = 0
    z = np.zeros((1, 16, 10), dtype=np.float32)
    o = ops.bma_reduce_cls(torch.from_numpy(z).to(cuda), ("entropy", "epistemic_entropy", "epistemic_variance", "mean"))
    np.testing.assert_allclose(o["entropy"].cpu().numpy(), np.log(10.0), rtol=1e-6)
    np.testing.assert_allclose(o["epistemic_entropy"].cpu().numpy(), 0.0, atol=1e-6)
    assert float(o["epistemic_variance"].abs().max()) == 0.0
    # very confident rows: p log p underflows to 0 * finite = 0 (no NaN), like the reference
    z = np.zeros((4, 8, 1000), dtype=np.float32)
    z[..., 3] = 200.0
    o = ops.bma_reduce_cls(torch.from_numpy(z).to(cuda), ("entropy", "aleatoric_entropy", "mean", "mode"))
    assert np.isfinite(o["entropy"].cpu().numpy()).all() and np.isfinite(o["aleatoric_entropy"].cpu().numpy()).all()
    assert (o["mode"].cpu().numpy() == 3).all()
    # classes masked with -inf in some samples: aleatoric entropy is NaN (0 * -inf), as in the reference
    z = _logits(3, 5, 8, seed=0)
    z[1, :, 2] = -np.inf
    o = ops.bma_reduce_cls(torch.from_numpy(z).to(cuda), ("aleatoric_entropy", "mean"))
    want = P.cls_aleatoric_entropy(z)
    assert np.isnan(want).all() and np.isnan(o["aleatoric_entropy"].cpu().numpy()).all()
    np.testing.assert_allclose(o["mean"].cpu().numpy(), P.cls_mean(z), rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("S,B,C,n_ranks", [(12, 300, 1000, 2), (9, 64, 10, 4), (6, 120, 33, 3)])
def test_shard_slots_finalize_equals_fused(cuda, S, B, C, n_ranks):
    """S-sharded path emulated on one device: every 'rank' reduces its samples into slots (owner = b // (B/N)), the
    all-to-all is emulated by indexing send[src][owner], each owner finalises its rows -> equals the oracle on all S."""
    from fortuna_b200 import ops

    z = _logits(S, B, C, seed=21 + C)
    t = np.random.default_rng(5).integers(0, C, B).astype(np.int32)
    zt = torch.from_numpy(z).to(cuda)
    tt = torch.from_numpy(t).to(cuda)
    rows = B // n_ranks
    bounds = np.linspace(0, S, n_ranks + 1).astype(int)  # uneven sample shards are fine
    sends = []
    for r in range(n_ranks):
        send = ops.alloc_cls_slots(n_ranks, rows, C, cuda)
        send.fill_(float("nan"))
        ptrs = ops.slot_pointer_table([send[o] for o in range(n_ranks)], cuda)
        ops.bma_reduce_cls_shard(zt[bounds[r] : bounds[r + 1]].contiguous(), ptrs, targets=tt, rank=r)
        sends.append(send)
    want = [n for n in ALL if n != "ensemble_log_prob"]
    parts = {n: [] for n in want}
    for o in range(n_ranks):
        recv = torch.stack([sends[src][o] for src in range(n_ranks)]).contiguous()  # what the all-to-all delivers to o
        outs = ops.bma_finalize_cls_slots(recv, S, C, want)
        for n in want:
            parts[n].append(outs[n])
    outs = {n: torch.cat(parts[n], dim=0) for n in want}
    outs["ensemble_log_prob"] = torch.from_numpy(P.cls_ensemble_log_prob(z, t)).to(cuda)
    _check(outs, z, t, S)


@pytest.mark.parametrize("S,B,d", [(8, 100, 1), (5, 33, 3)])
def test_regression_reductions(cuda, S, B, d):
    from fortuna_b200 import ops

    r = np.random.default_rng(7)
    z = r.standard_normal((S, B, 2 * d)).astype(np.float32)
    y = r.standard_normal((B, d)).astype(np.float32)
    lt = np.array([0.2], dtype=np.float32)
    outs = ops.bma_reduce_reg(
        torch.from_numpy(z).to(cuda),
        ("mean", "aleatoric_variance", "epistemic_variance", "variance", "std", "log_prob", "ensemble_log_prob"),
        targets=torch.from_numpy(y).to(cuda),
        log_temp=torch.from_numpy(lt).to(cuda),
    )
    zc = P.reg_calibrate(z, lt)
    g = {k: v.cpu().numpy() for k, v in outs.items()}
    np.testing.assert_allclose(g["mean"], P.reg_mean(zc), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(g["aleatoric_variance"], P.reg_aleatoric_variance(zc), rtol=1e-5)
    np.testing.assert_allclose(g["epistemic_variance"], P.reg_epistemic_variance(zc), rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(g["log_prob"], P.reg_log_prob(zc, y), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(g["ensemble_log_prob"], P.reg_ensemble_log_prob(zc, y), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize(
    "S,B,D,C,dtype", [(3, 70, 30, 2, "f32"), (4, 129, 84, 10, "f32"), (2, 100, 768, 2, "bf16"), (2, 65, 96, 130, "f32")]
)
@pytest.mark.parametrize("layout", ["sdc", "scd"])
def test_last_layer_cuda_core_path(cuda, S, B, D, C, dtype, layout):
    from fortuna_b200 import ops

    r = np.random.default_rng(9)
    x = r.standard_normal((B, D)).astype(np.float32)
    w = (r.standard_normal((S, D, C)) / np.sqrt(D)).astype(np.float32)
    bias = r.standard_normal((S, C)).astype(np.float32)
    lt = np.array([0.3], dtype=np.float32)
    bf = dtype == "bf16"
    want = P.last_layer_logits(x, w, bias, lt, bf16_inputs=bf)
    td = torch.bfloat16 if bf else torch.float32
    xt = torch.from_numpy(x).to(cuda).to(td)
    wt = torch.from_numpy(w).to(cuda).to(td)
    if layout == "scd":
        wt = wt.transpose(1, 2).contiguous()
    out = ops.last_layer_logits(
        xt, wt, torch.from_numpy(bias).to(cuda), torch.from_numpy(lt).to(cuda),
        w_layout=ops.W_SCD if layout == "scd" else ops.W_SDC, path=1,
    )
    np.testing.assert_allclose(out.cpu().numpy(), want, rtol=1e-5, atol=1e-5 * np.sqrt(D))


def test_pipeline_from_host_features_matches_oracle(cuda):
    """The loader-shaped last-layer predictive (streaming.stream_cls_reduce, the engine behind prob_model.predictive.*):
    pinned host features -> K2 -> K3 -> K5-K7 -> pinned host results, vs the oracle chain."""
    from fortuna_b200 import ops
    from fortuna_b200.prob_model.predictive import pipeline, streaming
    from oracle import scoring as Sc

    r = np.random.default_rng(21)
    n, S, B, D, C = 6, 9, 300, 128, 40
    x = r.standard_normal((B, D)).astype(np.float32)
    w = (r.standard_normal((n, D, C)) / np.sqrt(D)).astype(np.float32)
    bias = r.standard_normal((n, C)).astype(np.float32)
    lt = np.array([0.2], np.float32)
    t = r.integers(0, C, B).astype(np.int32)
    key = ops.PRNGKey(4, cuda)
    idx = ops.sample_indices(key, S, n)
    idx_np = idx.cpu().numpy()
    z = P.last_layer_logits(x, w[idx_np], bias[idx_np], lt, bf16_inputs=True)  # [S,B,C] oracle logits
    xs = P.round_to_bf16(x)
    batches = [(torch.from_numpy(xs[i : i + 128]).to(torch.bfloat16).pin_memory(), torch.from_numpy(t[i : i + 128]).pin_memory())
               for i in range(0, B, 128)]
    wt = ops.transpose_w_bf16(torch.from_numpy(w).to(cuda))
    bias_d, lt_d = torch.from_numpy(bias).to(cuda), torch.from_numpy(lt).to(cuda)
    val_probs = P.softmax((2 * r.standard_normal((500, C))).astype(np.float32))
    val_t = r.integers(0, C, 500)
    val_sorted = torch.from_numpy(np.sort(Sc.aps_score(val_probs, val_t))).to(cuda)
    calib = pipeline.ConformalCalibration(val_sorted, 500, 0.1)
    want = ("mean", "entropy", "log_prob", "mode")

    def logits_fn(xd):
        assert xd.dtype == torch.bfloat16  # pinned bf16 batches are the DMA source as they are
        return ops.last_layer_logits(xd, wt, bias_d, lt_d, w_layout=ops.W_SCD, sample_idx=idx)

    out = streaming.stream_cls_reduce(batches, logits_fn, want, B, S, cuda, sets=calib, to_host=True, ece_total=B)
    # bf16 GEMM with float32 accumulation in a different order than the oracle's: logits agree to ~1e-3 relative
    np.testing.assert_allclose(out["mean"].numpy(), P.cls_mean(z), rtol=5e-3, atol=1e-5)
    np.testing.assert_allclose(out["entropy"].numpy(), P.cls_entropy(z), rtol=5e-3, atol=1e-4)
    np.testing.assert_allclose(out["log_prob"].numpy(), P.cls_log_prob(z, t), rtol=5e-3, atol=1e-3)
    # scoring is integer-exact GIVEN the device mean: recompute the oracle's sets from the returned mean
    mean = out["mean"].numpy()
    conds, sizes, _ = Sc.conformal_sets(val_sorted.cpu().numpy(), Sc.aps_cumsums(mean), 0.1)
    assert np.array_equal(out["conformal_mask"].numpy().astype(bool), conds)
    assert np.array_equal(out["conformal_sizes"].numpy(), sizes)
    np.testing.assert_allclose(out["ece"].numpy()[0], Sc.expected_calibration_error(mean.argmax(-1), mean, t), rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("n,B,C,T", [(5, 7, 2, 6), (10, 129, 10, 30), (3, 33, 1000, 4), (4, 50, 37, 3), (2, 9, 1, 2), (6, 300, 1024, 2)])
def test_cls_sample_targets_are_the_reference_draws(cuda, n, B, C, T):
    """fb200_cls_sample_targets vs the oracle's restatement of Predictive.sample -> ClassificationProbOutputLayer.sample
    (jax.random.choice with p = cumsum + searchsorted): posterior-sample draws and class draws are integers, compared
    exactly."""
    from fortuna_b200 import ops
    from oracle import prng as oprng

    r = np.random.default_rng(n + B + C)
    z = (2.5 * r.standard_normal((n, B, C))).astype(np.float32)
    lt = np.array([0.2], np.float32)
    want, want_idx = P.cls_sample_targets(z, oprng.PRNGKey(13), T, lt)
    y, idx = ops.cls_sample_targets(torch.from_numpy(z).to(cuda), ops.PRNGKey(13, cuda), T,
                                    log_temp=torch.from_numpy(lt).to(cuda), want_idx=True)
    assert np.array_equal(idx.cpu().numpy(), want_idx)
    got = y.cpu().numpy()
    assert got.shape == (T, B) and got.min() >= 0 and got.max() < C
    assert np.array_equal(got, want)


def test_cls_sample_targets_follow_the_predictive_distribution(cuda):
    """Many draws from one stored sample: class frequencies approach softmax (a statistical sanity check of the draw)."""
    from fortuna_b200 import ops

    z = torch.tensor([[[0.0, 1.0, 2.0, -1.0]]], device=cuda).repeat(1, 4096, 1).contiguous()
    y = ops.cls_sample_targets(z, ops.PRNGKey(5, cuda), 16)
    freq = torch.bincount(y.reshape(-1).long(), minlength=4).float() / y.numel()
    np.testing.assert_allclose(freq.cpu().numpy(), P.softmax(np.array([[0.0, 1.0, 2.0, -1.0]], np.float32))[0], atol=0.01)


@pytest.mark.parametrize("S,B,C,dtype", [(5, 40, 7, torch.float32), (30, 1000, 10, torch.float32), (3, 257, 1000, torch.float32),
                                         (4, 64, 2, torch.bfloat16)])
@pytest.mark.parametrize("lt", [0.0, 0.7])
def test_calib_loss_terms_match_oracle(cuda, S, B, C, dtype, lt):
    """fb200_calib_cls_loss_terms: per-sample log-likelihood sums under the temperature scaler and their derivative in
    tau = exp(-log_temp) (float64 outputs) against the float64 oracle; then the host chain rule against finite
    differences of the oracle loss."""
    from fortuna_b200 import ops
    from fortuna_b200.prob_model.calibration import loss_and_grad
    from oracle import calibration as Cal

    r = np.random.default_rng(S + C)
    z = (3 * r.standard_normal((S, B, C))).astype(np.float32)
    y = r.integers(0, C, B).astype(np.int32)
    zt = torch.from_numpy(z).to(cuda).to(dtype)
    z_used = zt.float().cpu().numpy()
    tt = torch.from_numpy(y).to(cuda)
    ltt = torch.tensor([lt], dtype=torch.float32, device=cuda)
    got = ops.calib_cls_loss_terms(zt, tt, ltt).cpu().numpy()
    L, dL = Cal.calib_terms(z_used, y, np.float32(lt))
    np.testing.assert_allclose(got[0], L, rtol=2e-6, atol=1e-3)
    np.testing.assert_allclose(got[1], dL, rtol=2e-5, atol=2e-3)
    loss, g = loss_and_grad(zt, tt, np.float32(lt), n_data=3 * B)
    np.testing.assert_allclose(loss, Cal.calib_loss(z_used, y, np.float32(lt), 3 * B), rtol=1e-5)
    h = 1e-4
    fd = (Cal.calib_loss(z_used, y, lt + h, 3 * B) - Cal.calib_loss(z_used, y, lt - h, 3 * B)) / (2 * h)
    np.testing.assert_allclose(g, fd, rtol=2e-3, atol=1e-3)


def test_calibrate_temperature_recovers_the_generating_temperature(cuda):
    """Over-confident logits (true logits x 3): Adam on the kernel's loss drives log_temp to log 3 and the NLL down."""
    from fortuna_b200.prob_model import Adam, CalibConfig, CalibOptimizer
    from fortuna_b200.prob_model.calibration import calibrate_temperature

    r = np.random.default_rng(2)
    z_true = 2 * r.standard_normal((1, 6000, 5))
    p = P.softmax(z_true[0].astype(np.float32)).astype(np.float64)
    y = np.array([r.choice(5, p=pi / pi.sum()) for pi in p])
    z = torch.from_numpy(np.repeat(3.0 * z_true, 4, axis=0).astype(np.float32)).to(cuda)  # 4 identical "samples"
    status, log_temp = calibrate_temperature(z, y, batch_size=2000, calib_config=CalibConfig(
        optimizer=CalibOptimizer(method=Adam(5e-2), n_epochs=60)), z_val=z[:, :1000].contiguous(), targets_val=y[:1000])
    assert status["loss"].shape == (60,) and status["val_loss"].shape == (60,)
    assert status["loss"][-1] < status["loss"][0] and abs(float(log_temp[0]) - np.log(3.0)) < 0.12


@pytest.mark.parametrize("S,B,D,C", [(100, 1000, 84, 10), (32, 513, 512, 10), (7, 130, 128, 3), (3, 127, 132, 17), (5, 129, 256, 64),
                                     (64, 300, 768, 2), (1, 1, 128, 1)])
@pytest.mark.parametrize("gather", [False, True])
@pytest.mark.parametrize("out_bf16", [False, True])
def test_last_layer_kmajor_f32_kernel(cuda, S, B, D, C, gather, out_bf16):
    """float32 operands with K-major weights, D % 4 == 0 and D >= 128 take the double-buffered kernel (the c3 shape among them):
    same FMA order per output as the generic kernel - bit-identical to it (reached through the [S, D, C] layout) - and
    within rtol 1e-5 of the oracle; tiles that cut samples, ragged rows, one-class and one-row edges, gathered samples."""
    from fortuna_b200 import ops

    r = np.random.default_rng(S * 1000 + D)
    n = S + 3 if gather else S
    x = r.standard_normal((B, D)).astype(np.float32)
    w = (r.standard_normal((n, D, C)) / np.sqrt(D)).astype(np.float32)
    bias = r.standard_normal((n, C)).astype(np.float32)
    lt = np.array([-0.2], dtype=np.float32)
    idx = r.integers(0, n, S).astype(np.int32) if gather else None
    xt, w_sdc = torch.from_numpy(x).to(cuda), torch.from_numpy(w).to(cuda)
    w_scd = w_sdc.transpose(1, 2).contiguous()
    bt, ltt = torch.from_numpy(bias).to(cuda), torch.from_numpy(lt).to(cuda)
    it = torch.from_numpy(idx).to(cuda) if gather else None
    od = torch.bfloat16 if out_bf16 else torch.float32
    new = ops.last_layer_logits(xt, w_scd, bt, ltt, out_dtype=od, w_layout=ops.W_SCD, sample_idx=it, path=1)
    old = ops.last_layer_logits(xt, w_sdc, bt, ltt, out_dtype=od, w_layout=ops.W_SDC, sample_idx=it, path=1)
    assert new.shape == (S, B, C) and torch.equal(new, old)
    sel = idx if gather else np.arange(S)
    want = P.last_layer_logits(x, w[sel], bias[sel], lt)
    tol = dict(rtol=1e-2, atol=1e-2) if out_bf16 else dict(rtol=1e-5, atol=1e-5 * np.sqrt(D))
    np.testing.assert_allclose(new.float().cpu().numpy(), want, **tol)


@pytest.mark.parametrize("C", [10, 130, 1000, 1024, 1500, 4096])
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_reduction_special_values_follow_the_reference(cuda, C, dtype):
    """Rows that hold +inf, -inf or NaN, rows far below zero, rows whose lanes differ by hundreds of units, constant rows
    and a sample whose logits are ALL -inf: NaN / inf patterns and values as the reference's float32 expressions give
    them (oracle).  The bf16 instantiation shifts every lane's exponentials by the lane's own maximum and rescales
    afterwards; the float32 one keeps the row shift - both must land on the same answers.  (Padding lanes - C not a
    multiple of the lane tiling - must not lend an all -inf row a finite maximum: the reference's softmax of it is NaN.)"""
    from fortuna_b200 import ops

    r = np.random.default_rng(C)
    S, B = 6, 96
    z = P.round_to_bf16((4 * r.standard_normal((S, B, C))).astype(np.float32))
    z[0, 1, :] = -300.0                      # far below zero: exp underflows without a shift
    z[1, 2, :] = 0.0                         # constant row
    z[2, 3, : C // 2] = 250.0                # entries that differ by hundreds of units
    z[2, 3, C // 2 :] = -250.0
    z[3, 4, C // 3] = np.inf                 # +inf: jax.nn.softmax subtracts the maximum as it is - the whole row is NaN
    z[4, 5, C // 4] = -np.inf                # -inf: p = 0, p ln p = NaN in the aleatoric entropy
    z[5, 6, C - 1] = np.nan                  # NaN poisons its row
    z[0, 7, :] = -np.inf                     # a sample of -inf only: softmax = NaN
    names = ("mean", "aleatoric_variance", "epistemic_variance", "entropy", "aleatoric_entropy", "epistemic_entropy")
    zt = torch.from_numpy(z).to(cuda)
    got = ops.bma_reduce_cls(zt.to(torch.bfloat16) if dtype == "bf16" else zt, names)
    with np.errstate(all="ignore"):
        want = {"mean": P.cls_mean(z), "aleatoric_variance": P.cls_aleatoric_variance(z),
                "epistemic_variance": P.cls_epistemic_variance(z), "entropy": P.cls_entropy(z),
                "aleatoric_entropy": P.cls_aleatoric_entropy(z), "epistemic_entropy": P.cls_epistemic_entropy(z)}
    for n in names:
        x, y = got[n].float().cpu().numpy(), np.asarray(want[n], np.float32)
        assert np.array_equal(np.isnan(x), np.isnan(y)), n
        ok = np.isfinite(y)
        np.testing.assert_allclose(x[ok], y[ok], rtol=2e-5, atol=4e-6, err_msg=n)
    # arg-max and max of the mean (jnp.argmax / jnp.max: a NaN row gives class 0 and NaN)
    extra = ops.bma_reduce_cls(zt.to(torch.bfloat16) if dtype == "bf16" else zt, ("mean", "mode", "confidence"))
    with np.errstate(all="ignore"):
        nan_rows = np.isnan(want["mean"]).any(axis=1)
    mode, conf = extra["mode"].cpu().numpy(), extra["confidence"].cpu().numpy()
    assert (mode[nan_rows] == 0).all() and np.isnan(conf[nan_rows]).all() and np.isfinite(conf[~nan_rows]).all()
    assert np.array_equal(mode[~nan_rows], extra["mean"].cpu().numpy()[~nan_rows].argmax(1))


@pytest.mark.parametrize("C", [10, 130, 1000])
def test_row_softmax_special_values_follow_the_reference(cuda, C):
    """fb200_softmax_rows on rows holding +inf / -inf / NaN / only -inf: the softmax epilogue follows jax.nn.softmax
    (x - max as it is: such rows are NaN), the log-softmax epilogue and the row logsumexp follow jsp.special.logsumexp
    (a non-finite maximum is replaced by 0: z - lse is -inf for the finite classes of a row with a +inf logit)."""
    from fortuna_b200 import ops

    r = np.random.default_rng(C)
    z = (4 * r.standard_normal((64, C))).astype(np.float32)
    z[1, :] = -300.0
    z[3, C // 3] = np.inf
    z[4, C // 4] = -np.inf
    z[5, C - 1] = np.nan
    z[6, :] = -np.inf
    with np.errstate(all="ignore"):
        want_sm, want_lsm, want_lse = P.softmax(z), P.log_softmax_rows(z[None])[0], P.logsumexp(z)
    for epi, want in ((ops.EPI_SOFTMAX, want_sm), (ops.EPI_LOG_SOFTMAX, want_lsm)):
        got, lse = ops.softmax_rows_(torch.from_numpy(z.copy()).to(cuda)[None].contiguous(), epi, want_lse=True)
        got, lse = got[0].cpu().numpy(), lse.reshape(-1).cpu().numpy()
        assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(np.isinf(got), np.isinf(want))
        ok = np.isfinite(want)
        np.testing.assert_allclose(got[ok], want[ok], rtol=2e-5, atol=1e-6)
        assert np.array_equal(np.isnan(lse), np.isnan(want_lse)) and np.array_equal(np.isinf(lse), np.isinf(want_lse))


@pytest.mark.parametrize("C", [10, 1000])
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_log_prob_special_values_follow_the_reference(cuda, C, dtype):
    """log_prob / ensemble_log_prob keep jsp.special.logsumexp's convention (a non-finite maximum is replaced by 0): a +inf
    logit makes the other classes' log-likelihood -inf and its own NaN, a -inf target logit -inf, an all -inf sample NaN."""
    from fortuna_b200 import ops

    r = np.random.default_rng(C)
    S, B = 6, 96
    z = P.round_to_bf16((4 * r.standard_normal((S, B, C))).astype(np.float32))
    t = r.integers(0, C, B).astype(np.int32)
    z[0, 1, :] = -300.0
    z[3, 4, C // 3], t[4] = np.inf, C // 3            # the target is the +inf class
    z[3, 8, C // 3], t[8] = np.inf, (C // 3 + 1) % C  # the target is another class
    z[4, 5, C // 4], t[5] = -np.inf, C // 4
    z[4, 9, C // 4], t[9] = -np.inf, (C // 4 + 1) % C
    z[5, 6, C - 1] = np.nan
    z[0, 7, :] = -np.inf
    with np.errstate(all="ignore"):
        want = {"log_prob": P.cls_log_prob(z, t), "ensemble_log_prob": P.cls_ensemble_log_prob(z, t)}
    zt = torch.from_numpy(z).to(cuda)
    got = ops.bma_reduce_cls(zt.to(torch.bfloat16) if dtype == "bf16" else zt, tuple(want), targets=torch.from_numpy(t).to(cuda))
    for n, w in want.items():
        g, w = got[n].cpu().numpy(), np.asarray(w, np.float32)
        assert np.array_equal(np.isnan(g), np.isnan(w)) and np.array_equal(np.isinf(g), np.isinf(w)), n
        ok = np.isfinite(w)
        np.testing.assert_allclose(g[ok], w[ok], rtol=1e-5, atol=4e-6, err_msg=n)
