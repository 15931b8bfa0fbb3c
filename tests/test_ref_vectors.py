"""The oracle against tests/golden/ref_vectors.npz - outputs of the REFERENCE'S OWN SOURCE FILES, executed from
/root/reference over the numpy stand-in for jax (tests/golden/make_ref_vectors.py, tests/golden/jax_standin.py; the
header of the latter says what such a run pins and what it does not).

Bars: integers, keys, sets, masks, modes bit-exact; floats bit-exact wherever the reference's expression has one
evaluation order in float32 (elementwise maps, running sums over posterior samples, sequential scans), and to a few
float32 ulps where a jnp.sum / jnp.mean over many terms is involved (numpy's pairwise order on the stand-in's side,
the oracle's restatement on the other: XLA has a third order, so nothing tighter is meaningful).
"""
import os

import numpy as np
import pytest

from oracle import diagnostic as Dg
from oracle import predictive as P
from oracle import prng as OP
from oracle import sampler as S
from oracle import scoring as Sc
from tests import trees

HERE = os.path.dirname(os.path.abspath(__file__))
REF = np.load(os.path.join(HERE, "golden", "ref_vectors.npz"))


def case(prefix):
    return {k[len(prefix):]: REF[k] for k in REF.files if k.startswith(prefix)}


def drawn(key, S_, n):
    """SGMCMCPosterior.sample per key of split(rng, S): choice(key, n_samples) (sgmcmc_posterior.py:36-40)."""
    return np.array([OP.choice(q, n) for q in OP.split(key, int(S_))])


def log_temp_of(c):
    return None if np.isnan(c["log_temp"]) else np.float32(c["log_temp"])


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_classification_predictive(tag):
    c = case(f"cls_{tag}/")
    n, lt = c["W"].shape[0], log_temp_of(c)
    idx = drawn(c["key"], c["S"], n)
    z = P.last_layer_logits(c["x"], c["W"][idx], c["b"][idx], log_temp=lt)
    if "calibrated_outputs" in c:
        assert np.array_equal(z, c["calibrated_outputs"])
    # one evaluation order: exact
    assert np.array_equal(P.cls_mean(z), c["mean"])
    assert np.array_equal(P.cls_mode(z), c["mode"])
    assert np.array_equal(P.cls_aleatoric_variance(z), c["aleatoric_variance"])
    assert np.array_equal(P.cls_epistemic_variance(z), c["epistemic_variance"])
    assert np.array_equal(P.cls_ensemble_log_prob(z, c["t"]), c["ensemble_log_prob"])
    assert np.array_equal(P.cls_log_prob(z, c["t"]), c["log_prob"])
    # variance / std: rng, key1 = split(rng); rng, key2 = split(rng) - two different draws (base.py:878-897)
    r1, k1 = OP.split(c["key"])
    _, k2 = OP.split(r1)
    z1 = P.last_layer_logits(c["x"], c["W"][drawn(k1, c["S"], n)], c["b"][drawn(k1, c["S"], n)], log_temp=lt)
    z2 = P.last_layer_logits(c["x"], c["W"][drawn(k2, c["S"], n)], c["b"][drawn(k2, c["S"], n)], log_temp=lt)
    var = P.cls_aleatoric_variance(z1) + P.cls_epistemic_variance(z2)
    assert np.array_equal(var, c["variance"])
    assert np.array_equal(np.sqrt(var), c["std"])
    # entropies: a sum over the classes (the reference's vmap over i + jnp.sum): a few ulps
    for name, fn in (("entropy", P.cls_entropy), ("aleatoric_entropy", P.cls_aleatoric_entropy),
                     ("epistemic_entropy", P.cls_epistemic_entropy)):
        np.testing.assert_allclose(fn(z), c[name], rtol=2e-6, atol=5e-7, err_msg=name)
    # the literal O(C^2) form of the reference's loop agrees with the closed form the kernels use
    np.testing.assert_allclose(P.cls_entropy_faithful(z, "entropy"), c["entropy"], rtol=2e-6, atol=5e-7)
    # target samples (Predictive.sample -> ClassificationProbOutputLayer.sample): one key per target sample, split into
    # the posterior draw and the per-row keys of jax.random.choice(p=)
    stored = P.last_layer_logits(c["x"], c["W"], c["b"])
    y, _ = P.cls_sample_targets(stored, c["sample_key"], 6, log_temp=lt)
    assert np.array_equal(y, c["sample"])
    sets = P.cls_credible_set(P.cls_mean(z), 0.1)
    assert np.array_equal(np.array([len(s) for s in sets]), c["credible_set_sizes"])
    for i, s in enumerate(sets):
        assert s == c["credible_set"][i, : len(s)].tolist()


@pytest.mark.parametrize("tag", ["a", "b"])
def test_regression_predictive(tag):
    c = case(f"reg_{tag}/")
    n, lt = c["W"].shape[0], log_temp_of(c)
    stored = P.last_layer_logits(c["x"], c["W"], c["b"])  # [n, B, 2d] uncalibrated outputs of every stored sample
    idx = drawn(c["key"], c["S"], n)
    z = stored[idx] if lt is None else P.reg_calibrate(stored[idx], lt)
    assert np.array_equal(P.reg_mean(z), c["mean"])
    assert np.array_equal(P.reg_mean(z), c["mode"])
    assert np.array_equal(P.reg_aleatoric_variance(z), c["aleatoric_variance"])
    np.testing.assert_allclose(P.reg_epistemic_variance(z), c["epistemic_variance"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(P.reg_ensemble_log_prob(z, c["y"]), c["ensemble_log_prob"], rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(P.reg_log_prob(z, c["y"]), c["log_prob"], rtol=2e-6, atol=1e-6)
    r1, k1 = OP.split(c["key"])
    _, k2 = OP.split(r1)
    z1, z2 = stored[drawn(k1, c["S"], n)], stored[drawn(k2, c["S"], n)]
    if lt is not None:
        z1, z2 = P.reg_calibrate(z1, lt), P.reg_calibrate(z2, lt)
    var = P.reg_aleatoric_variance(z1) + P.reg_epistemic_variance(z2)
    np.testing.assert_allclose(var, c["variance"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(np.sqrt(var), c["std"], rtol=1e-6, atol=1e-9)
    # target samples: the key the reference's RandomNumberGenerator produced; exact (elementwise after the draws)
    y, _ = P.reg_sample_targets(stored, c["sample_key"], int(c["T"]), log_temp=lt)
    assert np.array_equal(y, c["sample"])
    np.testing.assert_allclose(P.quantile_axis0(y, [0.05, 0.5, 0.95]), c["quantile"], rtol=1e-6, atol=1e-7)
    ent = P.reg_mc_entropies(stored, c["key"], int(c["S"]), int(c["T"]), log_temp=lt)
    for name in ("entropy", "aleatoric_entropy", "epistemic_entropy"):
        np.testing.assert_allclose(ent[name], c[name], rtol=5e-6, atol=2e-6, err_msg=name)


def grad_of(p):
    return S.tree_map(lambda a: (a * 0.5 + 0.01).astype(np.float32), p)


def flat(tree):
    return np.concatenate([v.ravel() for _, v in S.tree_flatten(tree)])


TREES = {"mlp": trees.mlp_two_moons_tree, "ragged": trees.ragged_tree}


@pytest.mark.parametrize("tree_name", ["mlp", "ragged"])
@pytest.mark.parametrize("precond", ["identity", "rmsprop"])
@pytest.mark.parametrize("resample", [None, 3])
def test_sghmc_trajectory_bit_exact(tree_name, precond, resample):
    """sghmc_integrator.py:59-92 + sgmcmc_preconditioner.py + utils/random.py + cosine_schedule, six steps: parameters,
    momentum, second moments and the key chain equal the reference source's, bit for bit."""
    pre = f"sampler/sghmc_{precond}_{resample}/{tree_name}/"
    p = TREES[tree_name](0)
    pc = S.Preconditioner(precond)
    st = S.sghmc_init(p, OP.PRNGKey(7), pc)
    for step in range(int(REF["sampler/n_steps"])):
        upd, st = S.sghmc_update(grad_of(p), st, 0.05, S.cosine_schedule(1e-3, 10), pc, resample)
        p = S.apply_updates(p, upd)
        assert np.array_equal(st.rng_key, REF[pre + f"step{step}/key"])
        assert int(st.count) == int(REF[pre + f"step{step}/count"])
        if pre + f"step{step}/params" in REF.files:
            assert np.array_equal(flat(p), REF[pre + f"step{step}/params"])
            assert np.array_equal(flat(st.momentum), REF[pre + f"step{step}/momentum"])
            if precond == "rmsprop":
                assert np.array_equal(flat(st.precond), REF[pre + f"step{step}/v"])


@pytest.mark.parametrize("tree_name", ["mlp", "ragged"])
@pytest.mark.parametrize("precond", ["identity", "rmsprop"])
def test_cyclical_sgld_trajectory_bit_exact(tree_name, precond):
    """cyclical_sgld_integrator.py:64-100, cycle of 4 at exploration ratio 0.5: exploration (SGD) and sampling (SGLD)
    steps alternate in pairs; the key only advances in the sampling phase."""
    pre = f"sampler/csgld_{precond}/{tree_name}/"
    p = TREES[tree_name](0)
    pc = S.Preconditioner(precond)
    st = S.sghmc_init(p, OP.PRNGKey(9), pc)
    keys = []
    for step in range(int(REF["sampler/n_steps"])):
        upd, st = S.cyclical_sgld_update(grad_of(p), st, 1e-3, 4, 0.5, pc)
        p = S.apply_updates(p, upd)
        keys.append(st.rng_key.copy())
        assert np.array_equal(st.rng_key, REF[pre + f"step{step}/key"])
        assert int(st.count) == int(REF[pre + f"step{step}/count"])
        if pre + f"step{step}/params" in REF.files:
            assert np.array_equal(flat(p), REF[pre + f"step{step}/params"])
            assert np.array_equal(flat(st.momentum), REF[pre + f"step{step}/momentum"])
            if precond == "rmsprop":
                assert np.array_equal(flat(st.precond), REF[pre + f"step{step}/v"])
    assert np.array_equal(keys[0], keys[1]) and not np.array_equal(keys[1], keys[2])


def test_step_schedules():
    counts = REF["schedule/counts"]
    scheds = {
        "constant": S.constant_schedule(2e-3),
        "cosine": S.cosine_schedule(1e-2, 10),
        "polynomial": S.polynomial_schedule(a=0.5, b=2.0, gamma=0.55),
        "cosine_burnin": S.constant_schedule_with_cosine_burnin(1e-2, 1e-3, 8),
        "cyclical": S.cyclical_cosine_schedule_with_const_burnin(1e-2, 5, 6),
    }
    for name, fn in scheds.items():
        got = np.array([fn(np.int32(c)) for c in counts], dtype=np.float32)
        if name == "polynomial":  # powf: libm here and there, but not the same call (np.power vs **)
            np.testing.assert_allclose(got, REF[f"schedule/{name}"], rtol=3e-7)
        else:
            assert np.array_equal(got, REF[f"schedule/{name}"]), name


def test_clip_by_global_norm():
    g = grad_of(trees.ragged_tree(0))
    for i in range(2):
        clipped, mult = S.clip_gradients_by_norm(g, float(REF[f"clip/{i}/max_norm"]))
        np.testing.assert_allclose(flat(clipped), REF[f"clip/{i}/grad"], rtol=3e-7, atol=0)
    assert np.array_equal(flat(clipped), flat(g))  # max_norm = 1e3: untouched


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_scores_sets_and_calibration_metrics(tag):
    c = case(f"score_{tag}/")
    vp, tp, vt = c["val_probs"], c["test_probs"], c["val_targets"]
    assert np.array_equal(Sc.aps_score(vp, vt), c["aps_val_scores"])
    assert np.array_equal(Sc.aps_cumsums(tp), c["aps_test_scores"])
    assert np.array_equal(Sc.simple_score(vp, vt), c["simple_val_scores"])
    assert np.array_equal(Sc.simple_scores_all(tp), c["simple_test_scores"])
    for e in (0.05, 0.2) + ((0.0, 1.0, 0.999999) if tag == "a" else ()):
        conds, sizes, _ = Sc.conformal_sets(Sc.aps_score(vp, vt), Sc.aps_cumsums(tp), e)
        assert np.array_equal(conds, c[f"aps_sets_{e}"].astype(bool))
        conds, sizes, _ = Sc.conformal_sets(Sc.simple_score(vp, vt), Sc.simple_scores_all(tp), e)
        assert np.array_equal(conds, c[f"simple_sets_{e}"].astype(bool))
        # the literal [n_val, n_test, C] broadcast and the rank-count form give the same sets
        assert np.array_equal(Sc.conformal_conds_literal(Sc.aps_score(vp, vt), Sc.aps_cumsums(tp), e),
                              c[f"aps_sets_{e}"].astype(bool))
    if tag == "a":  # CV+: scores of the folds concatenated (conformal/classification/base.py:111-132)
        h, nt = vp.shape[0] // 2, tp.shape[0]
        for kind, sfn, afn in (("aps", Sc.aps_score, Sc.aps_cumsums), ("simple", Sc.simple_score, Sc.simple_scores_all)):
            val = np.concatenate([sfn(vp[:h], vt[:h]), sfn(vp[h:], vt[h:])])
            test = np.concatenate([afn(tp[: nt // 2]), afn(tp[nt // 2 :])], axis=0)
            conds, _, _ = Sc.conformal_sets(val, test, 0.1)
            assert np.array_equal(conds, c[f"cvplus_{kind}_sets_0.1"].astype(bool)), kind
    preds = vp.argmax(-1).astype(np.int32)
    counts, confs, accs, _ = Sc.compute_counts_confs_accs(preds, vp, vt)
    assert np.array_equal(counts, c["ece_counts"])
    np.testing.assert_allclose(confs, c["ece_confs"], rtol=3e-7)
    np.testing.assert_allclose(accs, c["ece_accs"], rtol=3e-7)
    np.testing.assert_allclose(Sc.expected_calibration_error(preds, vp, vt), c["ece"], rtol=1e-6)
    np.testing.assert_allclose(Sc.maximum_calibration_error(preds, vp, vt), c["mce"], rtol=1e-6)
    assert np.float32(Sc.accuracy(preds, vt)) == c["accuracy"]
    np.testing.assert_allclose(Sc.brier_score(vp, vt), c["brier"], rtol=1e-6)


def test_diagnostics():
    samples, grads = REF["diag/samples"], REF["diag/grads"]
    np.testing.assert_allclose(Dg.kernel_stein_discrepancy_imq(samples, grads), REF["diag/ksd"], rtol=2e-5)
    np.testing.assert_allclose(Dg.kernel_stein_discrepancy_imq(samples, grads, c=2.0, beta=-0.75), REF["diag/ksd_c2_b075"],
                               rtol=2e-5)
    # (filter_threshold=None is not compared: with every lag kept the weighted autocorrelations of a centred chain sum
    # to 1/2 identically, so the reference's denominator -1 + 2 * sum is pure rounding noise - inf or +-1e8 on both sides.)
    # complex64 FFTs on both sides: the autocovariances agree to float32 rounding of O(1) terms
    ess, ref = Dg.effective_sample_size(samples), REF["diag/ess"]
    # the cut-off index is a threshold test on the autocorrelations: compare where it is decided by more than rounding
    margin = Dg.ess_threshold_margin(samples)
    ok = margin > 1e-4
    assert ok.mean() > 0.8
    np.testing.assert_allclose(ess[ok], ref[ok], rtol=2e-4)


# ------------------------------------------------------------------------------------------------ multivalid scorers
from oracle import multivalid as M  # noqa: E402


def _same_mv_run(pre, status, patches):
    """Round for round: the chosen (tau, g, v, c) exactly, patches and validation losses to float32 rounding."""
    want = REF[pre + "patches"]
    assert status["n_rounds"] == int(REF[pre + "n_rounds"]) and bool(status["converged"]) == bool(REF[pre + "converged"])
    np.testing.assert_allclose(np.array(status["losses"], np.float32), REF[pre + "losses"], rtol=2e-6)
    assert len(patches) == len(want)
    for got, w in zip(patches, want):
        assert (got[0], got[1], got[3]) == (int(w[0]), int(w[1]), int(w[3])) and np.float32(got[2]) == np.float32(w[2])
        np.testing.assert_allclose(got[4], w[4], rtol=2e-6, atol=1e-9)


@pytest.mark.parametrize("tag,kw", [("add", dict(bucket_types=("<=", ">="), patch_type="additive")),
                                    ("mul", dict(bucket_types=("=", ">=", "<="), patch_type="multiplicative"))])
def test_multicalibrator(tag, kw):
    d = case("mv/data/")
    pre = f"mv/multicalibrator_{tag}/"
    o = M.Iterative("multicalibrator", seed=1)
    st = o.calibrate(d["scores"], d["groups"], values=d["v0"], n_rounds=6, n_buckets=8, eta=0.5, **kw)
    _same_mv_run(pre, st, o.patches)
    tv = o.apply_patches(d["tg"], d["tx"])
    np.testing.assert_allclose(tv, REF[pre + "test_values"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(o.calibration_error(d["ts"], d["tg"], tv), REF[pre + "calibration_error"], rtol=2e-5, atol=1e-8)


def test_batchmvp():
    d = case("mv/data/")
    o = M.Iterative("batchmvp", seed=2)
    st = o.calibrate(d["scores"], d["groups"], n_rounds=5, n_buckets=10, eta=0.5, coverage=0.9, bucket_types=(">=", "<="))
    _same_mv_run("mv/batchmvp/", st, o.patches)
    th = o.apply_patches(d["tg"])
    np.testing.assert_allclose(th, REF["mv/batchmvp/test_thresholds"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(o.calibration_error(d["ts"], d["tg"], th), REF["mv/batchmvp/calibration_error"], rtol=2e-5,
                               atol=1e-8)


def test_top_label_and_binary_multicalibrators():
    d = case("mv/cls/")
    C = d["probs"].shape[1]
    o = M.Iterative("top_label", seed=0, n_classes=C)
    onehot = (d["targets"][:, None] == np.arange(C)[None]).astype(np.float32)
    st = o.calibrate(onehot, d["groups"], values=d["probs"], n_rounds=4, n_buckets=5, eta=1.0)
    _same_mv_run("mv/top_label/", st, o.patches)
    np.testing.assert_allclose(o.apply_patches(d["groups"], d["probs"]), REF["mv/top_label/applied"], rtol=1e-6, atol=1e-7)
    ob = M.Iterative("multicalibrator", seed=0)
    sb = ob.calibrate(d["binary_targets"].astype(np.float32), d["groups"], values=d["probs"][:, 0].copy(), n_rounds=4,
                      n_buckets=6, eta=0.5)
    _same_mv_run("mv/binary/", sb, ob.patches)


def test_one_shot_multicalibrators():
    d, o = case("mv/data/"), case("mv/one_shot/")
    p = M.one_shot_patches(d["scores"], o["values"], n_buckets=10)
    np.testing.assert_allclose(p, o["patches"], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(M.one_shot_apply(p, o["test_values"], 10), o["applied"], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(M.one_shot_patches(o["binary_targets"].astype(np.float32), o["values"], n_buckets=10),
                               o["binary_patches"], rtol=1e-6, atol=1e-8)
    onehot = (o["tl_targets"][:, None] == np.arange(3)[None]).astype(np.float32)
    wp = M.one_shot_patches(onehot, o["tl_probs"], n_buckets=10, top_label=True)
    np.testing.assert_allclose(wp, o["tl_patches"], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(M.one_shot_apply(wp, o["tl_probs"], 10, top_label=True), o["tl_applied"], rtol=1e-6, atol=1e-8)


def test_maxcov_fixed_precision_calibrator():
    from oracle import maxcov as Mc

    d = case("maxcov/")
    for i in range(2):
        tp, tn, margin = d[f"{i}/thresholds"]
        tau_pos, tau_neg, _, _ = Mc.calibrate(d["targets"], d["probs"], tp, tn, n_taus=60, margin=margin)
        assert np.float32(tau_pos) == d[f"{i}/tau_pos"] and np.float32(tau_neg) == d[f"{i}/tau_neg"]
        out = Mc.apply_patches(d["test_probs"], tau_pos, tau_neg)
        assert np.array_equal(out, d[f"{i}/applied"])
        np.testing.assert_allclose(Mc.true_positive_precision(out, d["targets"][:200], tp), d[f"{i}/tp_precision"], rtol=1e-6)
        np.testing.assert_allclose(Mc.true_negative_precision(out, d["targets"][:200], tn), d[f"{i}/tn_precision"], rtol=1e-6)


def test_calibration_losses():
    from oracle import calib_model as Cm

    d = case("loss/")
    for g in (2.0, 0.5):
        np.testing.assert_allclose(Cm.focal_loss(d["logits"], d["targets"], gamma=g), d[f"focal_gamma{g}"], rtol=2e-6)
    np.testing.assert_allclose(Cm.scaled_mse(d["outputs"], d["y"]), d["scaled_mse"], rtol=2e-6)


def test_conformal_bayes_sets():
    """predictive/classification.py:485-626 through the reference's own code (the test grid, the shared draws, the
    importance weights, the rank count).  The rank is a count of float32 comparisons between dot products over the S
    draws: entries the reference decides by less than 1e-6 relative could flip with the summation order."""
    c = case("cb/")
    n = c["W"].shape[0]
    idx = drawn(c["key"], c["S"], n)
    ell = P.cls_ensemble_log_prob(P.last_layer_logits(c["x_train"], c["W"][idx], c["b"][idx]), c["y_train"])
    lsm = P.log_softmax_rows(P.last_layer_logits(c["x_test"], c["W"][idx], c["b"][idx]))
    for e in (0.1, 0.4):
        region = Sc.conformal_bayes_region(ell, lsm, e)
        assert (region != c[f"region_{e}"].astype(bool)).sum() <= 1
    np.testing.assert_allclose(Sc.conformal_bayes_ess(lsm), c["ess"], rtol=1e-5)


def test_conformal_regression_quantiles():
    """quantile.py / onedim_uncertainty.py: scores and the conformal quantile level + linear interpolation."""
    c = case("creg/")
    y = c["val_targets"][:, 0]
    qs = np.maximum(c["val_lower"] - y, y - c["val_upper"]).astype(np.float32)
    assert np.array_equal(qs, c["q_scores"])
    os_ = (np.abs(c["val_targets"] - c["val_preds"]) / c["val_unc"])[:, 0].astype(np.float32)
    assert np.array_equal(os_, c["o_scores"])
    for e in (0.05, 0.2):
        assert np.float32(Sc.conformal_quantile(qs, e)) == c[f"q_quantile_{e}"]
        assert np.float32(Sc.conformal_quantile(os_, e)) == c[f"o_quantile_{e}"]


def test_calibration_loss_with_ensemble_outputs():
    """predictive/base.py:205-316 + likelihood/base.py:124-262: -(logsumexp_s((n_data / B) L_s) - log S) under the
    temperature scalers (the oracle evaluates in float64, the reference source in float32: rtol 1e-5)."""
    from oracle import calibration as Cal

    c = case("calib/")
    n_data = int(c["n_data"])
    for lt in (0.0, 0.4, -0.3):
        np.testing.assert_allclose(Cal.calib_loss(c["logits"], c["targets"], lt, n_data), c[f"cls_loss_{lt}"], rtol=1e-5)
        L, _ = Cal.calib_terms_reg(c["outputs"], c["y"], lt)
        a = (n_data / c["y"].shape[0]) * L
        want = -(a.max() + np.log(np.exp(a - a.max()).sum()) - np.log(len(a)))
        np.testing.assert_allclose(want, c[f"reg_loss_{lt}"], rtol=1e-5)


@pytest.mark.skipif(not os.path.isdir("/root/reference/fortuna"), reason="the reference tree only exists in the build container")
def test_committed_vectors_are_what_the_reference_source_produces(tmp_path):
    """Provenance of ref_vectors.npz: re-run the generator (which imports and executes the reference's files from
    /root/reference) and compare every array with the committed fixture, bit for bit."""
    import subprocess
    import sys

    out = tmp_path / "regen.npz"
    root = os.path.dirname(HERE)
    r = subprocess.run([sys.executable, os.path.join(HERE, "golden", "make_ref_vectors.py"), "--out", str(out)], cwd=root,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    new = np.load(out)
    assert sorted(new.files) == sorted(REF.files)
    for k in REF.files:
        assert new[k].dtype == REF[k].dtype and np.array_equal(new[k], REF[k], equal_nan=True), k


def test_output_calibration_predictive_maps():
    """output_calib_model/predictive/* over prob_output_layer/*: the S = 1 maps on given outputs, uncalibrated and under
    the temperature scalers, including the per-row target draws of ClassificationProbOutputLayer.sample."""
    from oracle import output_calib as Oc

    c = case("ocal/")
    lt = np.float32(c["log_temp"])
    for cal in (0, 1):
        ltc = lt if cal else None
        z = c["logits"] if not cal else (c["logits"] * np.exp(-lt, dtype=np.float32)).astype(np.float32)
        p = P.softmax(z)
        assert np.array_equal(p, c[f"cls_{cal}/mean"]) and np.array_equal(p.argmax(-1), c[f"cls_{cal}/mode"])
        assert np.array_equal((p * (1 - p)).astype(np.float32), c[f"cls_{cal}/variance"])
        assert np.array_equal(np.sqrt(p * (1 - p), dtype=np.float32), c[f"cls_{cal}/std"])
        np.testing.assert_allclose(P.cls_entropy(z[None]), c[f"cls_{cal}/entropy"], rtol=2e-6, atol=5e-7)
        assert np.array_equal(P.cls_ensemble_log_prob(z[None], c["targets"])[0], c[f"cls_{cal}/log_prob"])
        assert np.array_equal(Oc.cls_output_sample(c["logits"], c["key"], 9, ltc), c[f"cls_{cal}/sample"])
        o = c["outputs"] if not cal else P.reg_calibrate(c["outputs"][None], lt)[0]
        mu, lv = P.reg_split(o)
        assert np.array_equal(mu, c[f"reg_{cal}/mean"]) and np.array_equal(mu, c[f"reg_{cal}/mode"])
        assert np.array_equal(np.exp(lv, dtype=np.float32), c[f"reg_{cal}/variance"])
        np.testing.assert_allclose(np.sqrt(np.exp(lv, dtype=np.float32)), c[f"reg_{cal}/std"], rtol=2e-7)
        np.testing.assert_allclose(P.reg_ensemble_log_prob(o[None], c["y"])[0], c[f"reg_{cal}/log_prob"], rtol=2e-6, atol=1e-6)
        smp = Oc.reg_output_sample(c["outputs"], c["key"], 6, ltc)
        assert np.array_equal(smp, c[f"reg_{cal}/sample"])
        np.testing.assert_allclose(Oc.reg_output_entropy(c["outputs"], c["key"], 6, ltc), c[f"reg_{cal}/entropy"], rtol=5e-6, atol=2e-6)
        np.testing.assert_allclose(P.quantile_axis0(smp, [0.1, 0.9]), c[f"reg_{cal}/quantile"], rtol=1e-6, atol=1e-7)


def test_sampling_callbacks_store_the_reference_steps():
    """sghmc_callback.py / cyclical_sgld_callback.py / sgmcmc_sampling_callback.py: which 1-based training steps store
    posterior sample i - the oracle's rules and the product's callback classes (host logic) against the reference's."""
    from fortuna_b200.prob_model.posterior.sgmcmc.cyclical_sgld.cyclical_sgld_callback import CyclicalSGLDSamplingCallback
    from fortuna_b200.prob_model.posterior.sgmcmc.sghmc.sghmc_callback import SGHMCSamplingCallback

    class Repo:
        def __init__(self):
            self.puts = []

        def put(self, state, i, keep):
            self.puts.append((i, state))

    names = {"sghmc": ("n_epochs", "n_training_steps", "n_samples", "n_thinning", "burnin_length"),
             "csgld": ("n_epochs", "n_training_steps", "n_samples", "n_thinning", "cycle_length", "exploration_ratio")}
    for tag in ("sghmc_0", "sghmc_1", "csgld_0", "csgld_1"):
        kind = tag.split("_")[0]
        kw = dict(zip(names[kind], REF[f"callback/{tag}/config"].tolist()))
        kw = {k: (v if k == "exploration_ratio" else int(v)) for k, v in kw.items()}
        want = REF[f"callback/{tag}/stored"]
        n_steps = kw["n_epochs"] * kw["n_training_steps"]
        # oracle
        stored, count = [], 0
        for step in range(1, n_steps + 1):
            rule_kw = {k: v for k, v in kw.items() if k not in ("n_epochs", "n_training_steps")}
            hit = S.sghmc_do_sample(step, count, **rule_kw) if kind == "sghmc" else S.cyclical_sgld_do_sample(step, count, **rule_kw)
            if hit:
                stored.append((count, step))
                count += 1
        assert np.array_equal(np.array(stored, np.int32).reshape(-1, 2), want), tag
        # product callback classes
        repo = Repo()
        cls = SGHMCSamplingCallback if kind == "sghmc" else CyclicalSGLDSamplingCallback
        cb = cls(trainer=None, state_repository=repo, keep_top_n_checkpoints=1, **kw)
        for step in range(1, n_steps + 1):
            cb.training_step_end(step)
        assert np.array_equal(np.array(repo.puts, np.int32).reshape(-1, 2), want), tag
    assert int(REF["callback/sghmc_refuses"]) == 1 and int(REF["callback/csgld_refuses"]) == 1
    with pytest.raises(ValueError):
        SGHMCSamplingCallback(n_epochs=1, n_training_steps=10, n_samples=5, n_thinning=3, burnin_length=4, trainer=None,
                              state_repository=Repo(), keep_top_n_checkpoints=1)
    with pytest.raises(ValueError):
        CyclicalSGLDSamplingCallback(n_epochs=1, n_training_steps=12, n_samples=4, n_thinning=2, cycle_length=6,
                                     exploration_ratio=0.5, trainer=None, state_repository=Repo(), keep_top_n_checkpoints=1)
