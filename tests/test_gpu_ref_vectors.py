"""The CUDA path against tests/golden/ref_vectors.npz - outputs of the REFERENCE'S OWN SOURCE FILES (executed from
/root/reference over a numpy stand-in for jax: tests/golden/make_ref_vectors.py / jax_standin.py).  The classification
predictive goes through the public API (``ProbClassifier.predictive.*`` with the fixture's key, loader and stored
last-layer samples); regression, samplers, scoring and diagnostics through the ops layer over the C ABI.

Bars (north_star): keys, indices, modes, sets, masks, sizes, bin counts bit-exact; float32 outputs rtol 1e-5 with the
absolute floors of DESIGN section 4 (profiles/r02_tolerance_floors.txt)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
REF = np.load(os.path.join(HERE, "golden", "ref_vectors.npz"))


def case(prefix):
    return {k[len(prefix):]: REF[k] for k in REF.files if k.startswith(prefix)}


def dev_key(words, cuda):
    from fortuna_b200 import ops

    return ops.key_from_host([int(w) for w in words], cuda)


def make_classifier(cuda, c):
    from fortuna_b200.prob_model import ProbClassifier, SGHMCPosteriorApproximator

    n, D, Cn = c["W"].shape

    class Head(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.dfe_subnet = torch.nn.Identity()  # the features are the inputs: the backbone is outside the path
            self.output_subnet = torch.nn.Linear(D, Cn)

        def forward(self, x):
            return self.output_subnet(self.dfe_subnet(x).float())

    pm = ProbClassifier(model=Head(), posterior_approximator=SGHMCPosteriorApproximator(n_samples=n), seed=0, device=cuda)
    # flat layout (sorted leaf names, leaves aligned as utils.flat_tree does): output_subnet/bias [C], then
    # output_subnet/weight [C, D] = kernel^T
    from fortuna_b200.utils.flat_tree import FlatTree

    tmpl = FlatTree.from_shapes({"model": {"params": {"output_subnet": {"bias": (Cn,), "weight": (Cn, D)}}}}, device="cpu")
    ob, ow = tmpl.offsets
    bank = np.zeros((n, tmpl.flat.numel()), np.float32)
    bank[:, ob : ob + Cn] = c["b"]
    bank[:, ow : ow + Cn * D] = c["W"].transpose(0, 2, 1).reshape(n, -1)
    pm.attach_posterior_samples(torch.from_numpy(bank).to(cuda),
                                freeze_fun=lambda path, v: "trainable" if "output_subnet" in path else "frozen")
    if not np.isnan(c["log_temp"]):
        pm.predictive.log_temp = torch.tensor([float(c["log_temp"])], device=cuda)
    return pm


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_classification_predictive_api(cuda, tag):
    from fortuna_b200.data import DataLoader, InputsLoader

    c = case(f"cls_{tag}/")
    pm = make_classifier(cuda, c)
    S, bs = int(c["S"]), int(c["sizes"][0])
    inputs = InputsLoader.from_array_inputs(c["x"], batch_size=bs)
    data = DataLoader.from_array_data((c["x"], c["t"]), batch_size=bs)
    key = dev_key(c["key"], cuda)
    p = pm.predictive
    got = lambda t: t.cpu().numpy()
    if "calibrated_outputs" in c:
        z = got(p.sample_calibrated_outputs(inputs, S, rng=key))
        np.testing.assert_allclose(z, c["calibrated_outputs"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(got(p.mean(inputs, S, rng=key)), c["mean"], rtol=1e-5, atol=1e-8)
    assert np.array_equal(got(p.mode(inputs, S, rng=key)), c["mode"])
    np.testing.assert_allclose(got(p.aleatoric_variance(inputs, S, rng=key)), c["aleatoric_variance"], rtol=1e-5, atol=2e-7)
    floor = 4e-7 * float(c["mean"].max()) ** 2
    np.testing.assert_allclose(got(p.epistemic_variance(inputs, S, rng=key)), c["epistemic_variance"], rtol=1e-5, atol=floor)
    np.testing.assert_allclose(got(p.variance(inputs, S, rng=key)), c["variance"], rtol=1e-5, atol=2e-7 + floor)
    np.testing.assert_allclose(got(p.std(inputs, S, rng=key)), c["std"], rtol=1e-5, atol=1e-3 * np.sqrt(2e-7 + floor) + 1e-6)
    for name in ("entropy", "aleatoric_entropy", "epistemic_entropy"):
        np.testing.assert_allclose(got(getattr(p, name)(inputs, S, rng=key)), c[name], rtol=1e-5, atol=4e-6, err_msg=name)
    np.testing.assert_allclose(got(p.log_prob(data, S, rng=key)), c["log_prob"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(got(p.ensemble_log_prob(data, S, rng=key)), c["ensemble_log_prob"], rtol=1e-5, atol=2e-6)
    # target samples: integer class draws; a draw whose uniform lands within float32 rounding of a cumulative probability
    # may fall on the neighbouring class
    ys = got(p.sample(InputsLoader.from_array_inputs(c["x"]), 6, rng=dev_key(c["sample_key"], cuda)))
    assert ys.shape == c["sample"].shape and (ys != c["sample"]).mean() <= 0.01
    sets = p.credible_set(inputs, S, error=0.1, rng=key)
    # a set boundary is a threshold test on a float32 prefix sum: rows whose boundary the reference decides by less
    # than the mean's tolerance are skipped (none in these fixtures would be a surprise; the count is asserted)
    same = sum(s == c["credible_set"][i, : c["credible_set_sizes"][i]].tolist() for i, s in enumerate(sets))
    assert same >= len(sets) - 1, (same, len(sets))


@pytest.mark.parametrize("tag", ["a", "b"])
def test_regression_predictive_ops(cuda, tag):
    from fortuna_b200 import ops

    c = case(f"reg_{tag}/")
    n, S, T = c["W"].shape[0], int(c["S"]), int(c["T"])
    lt = None if np.isnan(c["log_temp"]) else torch.tensor([float(c["log_temp"])], device=cuda)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    stored = ops.last_layer_logits(t(c["x"]), t(c["W"]), t(c["b"]))  # [n, B, 2d], uncalibrated
    key = dev_key(c["key"], cuda)
    idx = ops.sample_indices(key, S, n).long()
    outs = ops.bma_reduce_reg(stored[idx].contiguous(), ("mean", "aleatoric_variance", "epistemic_variance", "log_prob",
                                                          "ensemble_log_prob"), targets=t(c["y"]), log_temp=lt)
    got = lambda x: x.cpu().numpy()
    np.testing.assert_allclose(got(outs["mean"]), c["mean"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(got(outs["aleatoric_variance"]), c["aleatoric_variance"], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(got(outs["epistemic_variance"]), c["epistemic_variance"], rtol=1e-5,
                               atol=4e-7 * float(np.abs(c["mean"]).max()) ** 2)
    np.testing.assert_allclose(got(outs["log_prob"]), c["log_prob"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(got(outs["ensemble_log_prob"]), c["ensemble_log_prob"], rtol=1e-5, atol=2e-6)
    # variance: the reference's two splits of rng (base.py:878-897)
    r1k1 = ops.split(key)
    k2 = ops.split(r1k1[0].contiguous())[1].contiguous()
    i1 = ops.sample_indices(r1k1[1].contiguous(), S, n).long()
    i2 = ops.sample_indices(k2, S, n).long()
    a = ops.bma_reduce_reg(stored[i1].contiguous(), ("aleatoric_variance",), log_temp=lt)["aleatoric_variance"]
    e = ops.bma_reduce_reg(stored[i2].contiguous(), ("epistemic_variance",), log_temp=lt)["epistemic_variance"]
    np.testing.assert_allclose(got(a + e), c["variance"], rtol=1e-5, atol=1e-7)
    # target samples with the key the reference's RandomNumberGenerator drew; quantiles of them
    y = ops.reg_sample_targets(stored, dev_key(c["sample_key"], cuda), T, log_temp=lt)
    np.testing.assert_allclose(got(y), c["sample"], rtol=1e-5, atol=2e-6)
    q = ops.quantile_axis0(y.contiguous(), t(np.array([0.05, 0.5, 0.95], np.float32)))
    np.testing.assert_allclose(got(q), c["quantile"], rtol=1e-5, atol=2e-6)
    # Monte-Carlo entropies (regression.py:51-267): key1, *keys = split(rng, 1 + S)
    ks = ops.split(key, 1 + S)
    midx = ops.sample_indices(ks[0].contiguous(), S, n)
    ent = ops.reg_mc_entropies(stored, midx, ks[1:].contiguous(), T, ("entropy", "aleatoric_entropy", "epistemic_entropy"),
                               log_temp=lt)
    for name in ("entropy", "aleatoric_entropy", "epistemic_entropy"):
        np.testing.assert_allclose(got(ent[name]), c[name], rtol=2e-5, atol=5e-6, err_msg=name)


TREES = {"mlp": "mlp_two_moons_tree", "ragged": "ragged_tree"}


def _run(cuda, tx, tree_name, pre, inner_of, rms):
    from fortuna_b200 import ops
    from fortuna_b200.utils.flat_tree import FlatTree
    from tests import trees

    params = FlatTree.from_tree(getattr(trees, TREES[tree_name])(0), device=cuda)
    state = tx.init(params)
    for step in range(int(REF["sampler/n_steps"])):
        g = params.empty_like()
        g.flat.copy_(params.flat * 0.5 + 0.01)  # input construction only (the fixture script's rule)
        params, state = tx.apply(params, g, state)
        inner = inner_of(state)
        assert np.array_equal(ops.key_to_host(inner.rng_key), REF[pre + f"step{step}/key"]), step
        if pre + f"step{step}/params" not in REF.files:
            continue
        flat = np.concatenate([v.cpu().numpy().ravel() for v in params.leaf_views()])
        mom = np.concatenate([v.cpu().numpy().ravel() for v in inner.momentum.leaf_views()])
        # after k steps a parameter is p0 + a sum of k updates: rtol 1e-5 of the value, plus an absolute floor of 1e-6 of
        # the array's scale for the elements where those terms cancel (with RMSprop the updates are O(0.1) against
        # parameters of O(0.01); measured: 4 of 24 018 elements off by 1.5e-7 after six steps)
        want_p, want_m = REF[pre + f"step{step}/params"], REF[pre + f"step{step}/momentum"]
        np.testing.assert_allclose(flat, want_p, rtol=1e-5, atol=1e-6 * float(np.abs(want_p).max()))
        np.testing.assert_allclose(mom, want_m, rtol=1e-5, atol=1e-6 * float(np.abs(want_m).max()))
        if rms:
            v = np.concatenate([x.cpu().numpy().ravel() for x in inner.preconditioner_state.grad_moment_estimates.leaf_views()])
            np.testing.assert_allclose(v, REF[pre + f"step{step}/v"], rtol=1e-5, atol=1e-12)


@pytest.mark.parametrize("tree_name", ["mlp", "ragged"])
@pytest.mark.parametrize("precond", ["identity", "rmsprop"])
@pytest.mark.parametrize("resample", [None, 3])
def test_sghmc_trajectories(cuda, tree_name, precond, resample):
    from fortuna_b200 import ops
    from fortuna_b200.prob_model.posterior.sgmcmc import sgmcmc_preconditioner as pc
    from fortuna_b200.prob_model.posterior.sgmcmc import sgmcmc_step_schedule as sched
    from fortuna_b200.prob_model.posterior.sgmcmc.sghmc.sghmc_integrator import sghmc_integrator

    pre = pc.rmsprop_preconditioner() if precond == "rmsprop" else pc.identity_preconditioner()
    tx = sghmc_integrator(0.05, resample, ops.PRNGKey(7, cuda), sched.cosine_schedule(1e-3, 10), pre)
    _run(cuda, tx, tree_name, f"sampler/sghmc_{precond}_{resample}/{tree_name}/", lambda s: s, precond == "rmsprop")


@pytest.mark.parametrize("tree_name", ["mlp", "ragged"])
@pytest.mark.parametrize("precond", ["identity", "rmsprop"])
def test_cyclical_sgld_trajectories(cuda, tree_name, precond):
    from fortuna_b200 import ops
    from fortuna_b200.prob_model.posterior.sgmcmc import sgmcmc_preconditioner as pc
    from fortuna_b200.prob_model.posterior.sgmcmc.cyclical_sgld.cyclical_sgld_integrator import cyclical_sgld_integrator

    pre = pc.rmsprop_preconditioner() if precond == "rmsprop" else pc.identity_preconditioner()
    tx = cyclical_sgld_integrator(ops.PRNGKey(9, cuda), 1e-3, 4, 0.5, pre)
    _run(cuda, tx, tree_name, f"sampler/csgld_{precond}/{tree_name}/", lambda s: s.sgld_state, precond == "rmsprop")


def test_step_schedules_and_clip(cuda):
    from fortuna_b200 import ops
    from fortuna_b200.prob_model.posterior.sgmcmc import sgmcmc_step_schedule as sched
    from fortuna_b200.utils.flat_tree import FlatTree
    from tests import trees

    scheds = {
        "constant": sched.constant_schedule(2e-3),
        "cosine": sched.cosine_schedule(1e-2, 10),
        "polynomial": sched.polynomial_schedule(a=0.5, b=2.0, gamma=0.55),
        "cosine_burnin": sched.constant_schedule_with_cosine_burnin(1e-2, 1e-3, 8),
        "cyclical": sched.cyclical_cosine_schedule_with_const_burnin(1e-2, 5, 6),
    }
    for name, fn in scheds.items():
        got = np.array([fn(int(c)) for c in REF["schedule/counts"]], dtype=np.float32)
        np.testing.assert_allclose(got, REF[f"schedule/{name}"], rtol=3e-7, err_msg=name)
    params = FlatTree.from_tree(trees.ragged_tree(0), device=cuda)
    g = params.zeros_like()  # alignment gaps between the leaves stay zero: they must not enter the norm
    for pv, gv in zip(params.leaf_views(), g.leaf_views()):
        gv.copy_(pv * 0.5 + 0.01)
    for i in range(2):
        mult = ops.grad_clip_mult(g.flat, float(REF[f"clip/{i}/max_norm"]))
        got = np.concatenate([(mult * v).cpu().numpy().ravel() for v in g.leaf_views()])
        np.testing.assert_allclose(got, REF[f"clip/{i}/grad"], rtol=1e-6)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_scores_sets_and_calibration_metrics(cuda, tag):
    from fortuna_b200 import ops
    from oracle import scoring as Sc  # thresholds only (host constants)

    c = case(f"score_{tag}/")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    vp, tp, vt = t(c["val_probs"]), t(c["test_probs"]), t(c["val_targets"])
    got = lambda x: x.cpu().numpy()
    assert np.array_equal(got(ops.aps_scores(vp, vt)), c["aps_val_scores"])
    assert np.array_equal(got(ops.aps_scores(tp, None)), c["aps_test_scores"])
    assert np.array_equal(got(ops.simple_scores(vp, vt)), c["simple_val_scores"])
    assert np.array_equal(got(ops.simple_scores(tp, None)), c["simple_test_scores"])
    n_val = c["val_probs"].shape[0]
    for e in (0.05, 0.2) + ((0.0, 1.0, 0.999999) if tag == "a" else ()):
        thr = Sc.conformal_threshold(n_val, e)
        val = ops.sort_f32_(ops.aps_scores(vp, vt))
        mask, sizes = ops.conformal_sets(val, ops.aps_scores(tp, None), thr)
        assert np.array_equal(got(mask).astype(bool), c[f"aps_sets_{e}"].astype(bool))
        assert np.array_equal(got(sizes), c[f"aps_sets_{e}"].sum(1))
        # the fused one-pass kernel (select + partial sort) gives the same sets
        mask2, sizes2 = ops.aps_conformal_sets(tp, val, thr)
        assert np.array_equal(got(mask2).astype(bool), c[f"aps_sets_{e}"].astype(bool))
        assert np.array_equal(got(sizes2), c[f"aps_sets_{e}"].sum(1))
        vals = ops.sort_f32_(ops.simple_scores(vp, vt))
        mask, _ = ops.conformal_sets(vals, ops.simple_scores(tp, None), thr)
        assert np.array_equal(got(mask).astype(bool), c[f"simple_sets_{e}"].astype(bool))
    e = ops.ece_bins(vp, vt, t(Sc.ece_thresholds(n_val)))
    assert np.array_equal(got(e["counts"]).astype(np.int64), c["ece_counts"].astype(np.int64))
    res = got(e["result"])  # [ece, mce, accuracy, brier]
    np.testing.assert_allclose(res[0], c["ece"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(res[1], c["mce"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(res[2], c["accuracy"], rtol=1e-6)
    np.testing.assert_allclose(res[3], c["brier"], rtol=1e-5)


def test_diagnostics(cuda):
    from fortuna_b200 import ops
    from oracle import diagnostic as Dg  # only the margin of the cut-off test (which rows are decided by rounding)

    x, g = torch.from_numpy(REF["diag/samples"]).to(cuda), torch.from_numpy(REF["diag/grads"]).to(cuda)
    np.testing.assert_allclose(ops.ksd_imq(x, g).cpu().numpy()[0], REF["diag/ksd"], rtol=2e-5)
    np.testing.assert_allclose(ops.ksd_imq(x, g, c=2.0, beta=-0.75).cpu().numpy()[0], REF["diag/ksd_c2_b075"], rtol=2e-5)
    ok = Dg.ess_threshold_margin(REF["diag/samples"]) > 1e-4
    np.testing.assert_allclose(ops.ess(x).cpu().numpy()[ok], REF["diag/ess"][ok], rtol=2e-4)


# ------------------------------------------------------------------------------------------------ multivalid scorers
def _same_mv_run(pre, status, patches):
    want = REF[pre + "patches"]
    assert status["n_rounds"] == int(REF[pre + "n_rounds"]) and bool(status["converged"]) == bool(REF[pre + "converged"])
    np.testing.assert_allclose(np.array([float(v) for v in status["losses"]], np.float32), REF[pre + "losses"], rtol=2e-5)
    assert len(patches) == len(want)
    for got, w in zip(patches, want):
        assert (int(got[0]), int(got[1]), int(got[3])) == (int(w[0]), int(w[1]), int(w[3])), (got, w)
        assert abs(float(got[2]) - w[2]) < 1e-7
        np.testing.assert_allclose(float(got[4]), w[4], rtol=2e-4, atol=1e-7)


@pytest.mark.parametrize("tag,kw", [("add", dict(bucket_types=("<=", ">="), patch_type="additive")),
                                    ("mul", dict(bucket_types=("=", ">=", "<="), patch_type="multiplicative"))])
def test_multicalibrator(cuda, tag, kw):
    from fortuna_b200.conformal import Multicalibrator

    d = case("mv/data/")
    pre = f"mv/multicalibrator_{tag}/"
    mc = Multicalibrator(seed=1)
    tv, status = mc.calibrate(scores=d["scores"], groups=d["groups"], values=d["v0"], test_groups=d["tg"],
                              test_values=d["tx"], n_rounds=6, n_buckets=8, eta=0.5, **kw)
    _same_mv_run(pre, status, mc.patches)
    np.testing.assert_allclose(tv.cpu().numpy(), REF[pre + "test_values"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(mc.calibration_error(d["ts"], d["tg"], tv).cpu().numpy(), REF[pre + "calibration_error"],
                               rtol=1e-4, atol=1e-7)


def test_batchmvp(cuda):
    from fortuna_b200.conformal import BatchMVPConformalRegressor

    d = case("mv/data/")
    bm = BatchMVPConformalRegressor(seed=2)
    status = bm.calibrate(scores=d["scores"], groups=d["groups"], n_rounds=5, n_buckets=10, eta=0.5, coverage=0.9)
    _same_mv_run("mv/batchmvp/", status, bm.patches)
    th = bm.apply_patches(d["tg"])
    np.testing.assert_allclose(th.cpu().numpy(), REF["mv/batchmvp/test_thresholds"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(bm.calibration_error(d["ts"], d["tg"], th).cpu().numpy(), REF["mv/batchmvp/calibration_error"],
                               rtol=1e-4, atol=1e-7)


def test_top_label_binary_and_one_shot_multicalibrators(cuda):
    from fortuna_b200.conformal import (BinaryClassificationMulticalibrator, OneShotBinaryClassificationMulticalibrator,
                                        OneShotMulticalibrator, OneShotTopLabelMulticalibrator, TopLabelMulticalibrator)

    d = case("mv/cls/")
    C = d["probs"].shape[1]
    tl = TopLabelMulticalibrator(n_classes=C, seed=0)
    status = tl.calibrate(targets=d["targets"], groups=d["groups"], probs=d["probs"], n_rounds=4, n_buckets=5, eta=1.0)
    _same_mv_run("mv/top_label/", status, tl.patches)
    np.testing.assert_allclose(tl.apply_patches(d["groups"], d["probs"]).cpu().numpy(), REF["mv/top_label/applied"],
                               rtol=1e-5, atol=1e-6)
    bc = BinaryClassificationMulticalibrator(seed=0)
    sb = bc.calibrate(targets=d["binary_targets"], groups=d["groups"], probs=d["probs"][:, 0].copy(), n_rounds=4,
                      n_buckets=6, eta=0.5)
    _same_mv_run("mv/binary/", sb, bc.patches)

    dd, o = case("mv/data/"), case("mv/one_shot/")
    os_ = OneShotMulticalibrator()
    tv = os_.calibrate(scores=dd["scores"], values=o["values"], test_values=o["test_values"], n_buckets=10)
    np.testing.assert_allclose(os_.patches, o["patches"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(tv.cpu().numpy(), o["applied"], rtol=1e-5, atol=1e-7)
    ob = OneShotBinaryClassificationMulticalibrator()
    ob.calibrate(targets=o["binary_targets"], probs=o["values"], n_buckets=10)
    np.testing.assert_allclose(ob.patches, o["binary_patches"], rtol=1e-5, atol=1e-7)
    ot = OneShotTopLabelMulticalibrator(n_classes=3)
    out = ot.calibrate(targets=o["tl_targets"], probs=o["tl_probs"], test_probs=o["tl_probs"], n_buckets=10)
    np.testing.assert_allclose(ot.patches, o["tl_patches"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(out.cpu().numpy(), o["tl_applied"], rtol=1e-5, atol=1e-7)


def test_maxcov_fixed_precision_calibrator(cuda):
    from fortuna_b200.conformal import MaxCoverageFixedPrecisionBinaryClassificationCalibrator as Cal

    d = case("maxcov/")
    for i in range(2):
        tp, tn, margin = (float(v) for v in d[f"{i}/thresholds"])
        cal = Cal()
        out = cal.calibrate(targets=d["targets"], probs=d["probs"], true_positive_precision_threshold=tp,
                            true_negative_precision_threshold=tn, test_probs=d["test_probs"], n_taus=60, margin=margin)
        assert np.float32(cal.patches["tau_pos"]) == d[f"{i}/tau_pos"] and np.float32(cal.patches["tau_neg"]) == d[f"{i}/tau_neg"]
        assert np.array_equal(out.cpu().numpy(), d[f"{i}/applied"])


def test_calibration_losses(cuda):
    from fortuna_b200 import ops

    d = case("loss/")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    for g in (2.0, 0.5):
        loss, _, _ = ops.focal_loss_grad(t(d["logits"]), t(d["targets"]), gamma=g, want_grad=False)
        np.testing.assert_allclose(float(loss[0]), d[f"focal_gamma{g}"], rtol=1e-5)
    loss, _, _ = ops.scaled_mse_loss_grad(t(d["outputs"]), t(d["y"]), want_grad=False)
    np.testing.assert_allclose(float(loss[0]), d["scaled_mse"], rtol=1e-5)


def test_conformal_bayes_sets(cuda):
    from fortuna_b200.data import DataLoader, InputsLoader

    c = case("cb/")
    pm = make_classifier(cuda, dict(W=c["W"], b=c["b"], log_temp=np.float32(np.nan)))
    train = DataLoader.from_array_data((c["x_train"], c["y_train"]), batch_size=32)
    test = InputsLoader.from_array_inputs(c["x_test"], batch_size=16)
    key = dev_key(c["key"], cuda)
    for e in (0.1, 0.4):
        sets, ess = pm.predictive.conformal_set(train, test, n_posterior_samples=int(c["S"]), error=e, rng=key, return_ess=True)
        mask = np.zeros(c[f"region_{e}"].shape, bool)
        for i, s in enumerate(sets):
            mask[i, s] = True
        assert (mask != c[f"region_{e}"].astype(bool)).sum() <= 1
    np.testing.assert_allclose(ess.cpu().numpy().reshape(c["ess"].shape), c["ess"], rtol=1e-4)


def test_conformal_regressors(cuda):
    from fortuna_b200.conformal.regression.quantile import (OneDimensionalUncertaintyConformalRegressor,
                                                            QuantileConformalRegressor)

    c = case("creg/")
    q = QuantileConformalRegressor()
    assert np.array_equal(q.score(c["val_lower"], c["val_upper"], c["val_targets"]), c["q_scores"])
    o = OneDimensionalUncertaintyConformalRegressor()
    assert np.array_equal(o.score(c["val_preds"], c["val_unc"], c["val_targets"]), c["o_scores"])
    for e in (0.05, 0.2):
        assert np.float32(float(q.quantile(c["val_lower"], c["val_upper"], c["val_targets"], e))) == c[f"q_quantile_{e}"]
        ci = q.conformal_interval(c["val_lower"], c["val_upper"], c["test_lower"], c["test_upper"], c["val_targets"], e)
        assert np.array_equal(ci, c[f"q_interval_{e}"])
        assert np.array_equal(q.is_in(np.zeros(len(ci), np.float32), ci), c[f"q_is_in_{e}"].astype(bool))
        assert np.float32(float(o.quantile(c["val_preds"], c["val_unc"], c["val_targets"], e))) == c[f"o_quantile_{e}"]
        ci = o.conformal_interval(c["val_preds"], c["val_unc"], c["test_preds"], c["test_unc"], c["val_targets"], e)
        assert np.array_equal(ci, c[f"o_interval_{e}"])


def test_calibration_loss_with_ensemble_outputs(cuda):
    from fortuna_b200.prob_model.calibration import loss_and_grad

    c = case("calib/")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    n_data = int(c["n_data"])
    for lt in (0.0, 0.4, -0.3):
        loss, _ = loss_and_grad(t(c["logits"]), t(c["targets"]), np.float32(lt), n_data=n_data)
        np.testing.assert_allclose(loss, c[f"cls_loss_{lt}"], rtol=1e-5)
        loss, _ = loss_and_grad(t(c["outputs"]), t(c["y"]), np.float32(lt), n_data=n_data, regression=True)
        np.testing.assert_allclose(loss, c[f"reg_loss_{lt}"], rtol=1e-5)


def test_output_calibration_predictive_maps(cuda):
    """OutputCalibClassifier / OutputCalibRegressor predictives on given outputs against the reference source's
    (output_calib_model/predictive/* over prob_output_layer/*), uncalibrated and with a temperature put into the state."""
    from fortuna_b200.output_calib_model import OutputCalibClassifier, OutputCalibRegressor
    from fortuna_b200.output_calib_model.output_calib_state_repository import OutputCalibStateRepository
    from fortuna_b200.output_calib_model.state import OutputCalibState

    c = case("ocal/")
    key = dev_key(c["key"], cuda)

    def with_temperature(model):
        repo = OutputCalibStateRepository()
        repo.put(OutputCalibState.init({"output_calibrator": {"params": {"log_temp": np.array([c["log_temp"]], np.float32)}}}))
        model.predictive.state = repo
        return model

    mc, mr = with_temperature(OutputCalibClassifier()), with_temperature(OutputCalibRegressor())
    got = lambda t: t.cpu().numpy()
    for cal in (0, 1):
        p, pre = mc.predictive, f"cls_{cal}/"
        np.testing.assert_allclose(got(p.mean(c["logits"], calibrated=bool(cal))), c[pre + "mean"], rtol=2e-5, atol=1e-8)
        assert np.array_equal(got(p.mode(c["logits"], calibrated=bool(cal))), c[pre + "mode"])
        np.testing.assert_allclose(got(p.variance(c["logits"], calibrated=bool(cal))), c[pre + "variance"], rtol=2e-5, atol=1e-7)
        np.testing.assert_allclose(got(p.std(c["logits"], calibrated=bool(cal))), c[pre + "std"], rtol=2e-5, atol=2e-4)
        np.testing.assert_allclose(got(p.entropy(c["logits"], calibrated=bool(cal))), c[pre + "entropy"], rtol=1e-5, atol=4e-6)
        np.testing.assert_allclose(got(p.log_prob(c["logits"], c["targets"], calibrated=bool(cal))), c[pre + "log_prob"],
                                   rtol=1e-5, atol=2e-6)
        smp = got(p.sample(9, c["logits"], rng=key, calibrated=bool(cal)))
        assert smp.shape == c[pre + "sample"].shape and (smp != c[pre + "sample"]).mean() <= 0.01
        p, pre = mr.predictive, f"reg_{cal}/"
        np.testing.assert_allclose(got(p.mean(c["outputs"], calibrated=bool(cal))), c[pre + "mean"], rtol=1e-6)
        np.testing.assert_allclose(got(p.mode(c["outputs"], calibrated=bool(cal))), c[pre + "mode"], rtol=1e-6)
        np.testing.assert_allclose(got(p.variance(c["outputs"], calibrated=bool(cal))), c[pre + "variance"], rtol=2e-5)
        np.testing.assert_allclose(got(p.std(c["outputs"], calibrated=bool(cal))), c[pre + "std"], rtol=2e-5)
        np.testing.assert_allclose(got(p.log_prob(c["outputs"], c["y"], calibrated=bool(cal))), c[pre + "log_prob"], rtol=2e-5,
                                   atol=2e-6)
        np.testing.assert_allclose(got(p.sample(6, c["outputs"], rng=key, calibrated=bool(cal))), c[pre + "sample"], rtol=1e-5,
                                   atol=2e-6)
        np.testing.assert_allclose(got(p.entropy(c["outputs"], n_target_samples=6, rng=key, calibrated=bool(cal))),
                                   c[pre + "entropy"], rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(got(p.quantile([0.1, 0.9], c["outputs"], n_samples=6, rng=key, calibrated=bool(cal))),
                                   c[pre + "quantile"], rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_regression_predictive_api(cuda, tag):
    """The regression predictive through ``ProbRegressor.predictive.*`` with the fixture's stored samples attached (the
    whole network is the Dense layer of the fixture), key and loader batches - against the reference source's outputs."""
    from fortuna_b200.data import DataLoader, InputsLoader
    from fortuna_b200.prob_model import ProbRegressor, SGHMCPosteriorApproximator
    from fortuna_b200.utils.flat_tree import FlatTree

    c = case(f"reg_{tag}/")
    n, D, two_d = c["W"].shape
    S, T, bs = int(c["S"]), int(c["T"]), int(c["sizes"][0])
    pm = ProbRegressor(model=torch.nn.Linear(D, two_d), posterior_approximator=SGHMCPosteriorApproximator(n_samples=n),
                       output_calibrator=None, seed=0, device=cuda)
    tmpl = FlatTree.from_shapes({"model": {"params": {"bias": (two_d,), "weight": (two_d, D)}}}, device="cpu")
    ob, ow = tmpl.offsets
    bank = np.zeros((n, tmpl.flat.numel()), np.float32)
    bank[:, ob : ob + two_d] = c["b"]
    bank[:, ow : ow + two_d * D] = c["W"].transpose(0, 2, 1).reshape(n, -1)
    pm.attach_posterior_samples(torch.from_numpy(bank).to(cuda))
    if not np.isnan(c["log_temp"]):
        pm.predictive.log_temp = torch.tensor([float(c["log_temp"])], device=cuda)
    inputs = InputsLoader.from_array_inputs(c["x"], batch_size=bs)
    data = DataLoader.from_array_data((c["x"], c["y"]), batch_size=bs)
    one = InputsLoader.from_array_inputs(c["x"])
    key = dev_key(c["key"], cuda)
    p, got = pm.predictive, (lambda t: t.cpu().numpy())
    np.testing.assert_allclose(got(p.mean(inputs, S, rng=key)), c["mean"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(got(p.mode(inputs, S, rng=key)), c["mode"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(got(p.aleatoric_variance(inputs, S, rng=key)), c["aleatoric_variance"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(got(p.epistemic_variance(inputs, S, rng=key)), c["epistemic_variance"], rtol=1e-5,
                               atol=4e-7 * float(np.abs(c["mean"]).max()) ** 2 + 1e-7)
    np.testing.assert_allclose(got(p.log_prob(data, S, rng=key)), c["log_prob"], rtol=1e-5, atol=4e-6)
    np.testing.assert_allclose(got(p.ensemble_log_prob(data, S, rng=key)), c["ensemble_log_prob"], rtol=1e-5, atol=4e-6)
    np.testing.assert_allclose(got(p.variance(inputs, S, rng=key)), c["variance"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(got(p.std(inputs, S, rng=key)), c["std"], rtol=1e-5, atol=1e-6)
    skey = dev_key(c["sample_key"], cuda)
    np.testing.assert_allclose(got(p.sample(one, T, rng=skey)), c["sample"], rtol=1e-5, atol=4e-6)
    np.testing.assert_allclose(got(p.quantile([0.05, 0.5, 0.95], one, T, rng=skey)), c["quantile"], rtol=1e-5, atol=4e-6)
    for name in ("entropy", "aleatoric_entropy", "epistemic_entropy"):
        np.testing.assert_allclose(got(getattr(p, name)(one, S, T, rng=key)), c[name], rtol=2e-5, atol=8e-6, err_msg=name)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_conformal_classifier_and_metric_apis(cuda, tag):
    """The mirror classes / functions a user of the reference calls - ``AdaptivePredictionConformalClassifier`` /
    ``SimplePredictionConformalClassifier`` (and their CV+ variants) ``.conformal_set``, ``fortuna.metric.classification.*`` -
    against the reference source's results on the same arrays."""
    from fortuna_b200.conformal import (AdaptivePredictionConformalClassifier, CVPlusAdaptivePredictionConformalClassifier,
                                        CVPlusSimplePredictionConformalClassifier, SimplePredictionConformalClassifier)
    from fortuna_b200.metric import classification as M

    c = case(f"score_{tag}/")
    vp, tp, vt = c["val_probs"], c["test_probs"], c["val_targets"]
    nt, C = tp.shape

    def as_mask(sets):
        m = np.zeros((nt, C), bool)
        for i, s_ in enumerate(sets):
            m[i, list(s_)] = True
        return m

    for kind, cls in (("aps", AdaptivePredictionConformalClassifier), ("simple", SimplePredictionConformalClassifier)):
        for e in (0.05, 0.2):
            assert np.array_equal(as_mask(cls().conformal_set(vp, vt, tp, error=e)), c[f"{kind}_sets_{e}"].astype(bool)), (kind, e)
    if tag == "a":
        h, nv = vp.shape[0] // 2, vp.shape[0]
        for kind, cls in (("aps", CVPlusAdaptivePredictionConformalClassifier), ("simple", CVPlusSimplePredictionConformalClassifier)):
            sets = cls().conformal_set([vp[:h], vp[h:]], [vt[:h], vt[h:]], [tp[: nt // 2], tp[nt // 2 :]], error=0.1)
            assert np.array_equal(as_mask(sets), c[f"cvplus_{kind}_sets_0.1"].astype(bool)), kind
    preds = vp.argmax(-1).astype(np.int32)
    f = lambda t: np.asarray(t.cpu().numpy() if hasattr(t, "cpu") else t)
    counts, confs, accs = M.compute_counts_confs_accs(preds, vp, vt)
    assert np.array_equal(f(counts).astype(np.int64), c["ece_counts"].astype(np.int64))
    np.testing.assert_allclose(f(confs), c["ece_confs"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(f(accs), c["ece_accs"], rtol=1e-6)
    np.testing.assert_allclose(float(f(M.expected_calibration_error(preds, vp, vt))), c["ece"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(float(f(M.maximum_calibration_error(preds, vp, vt))), c["mce"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(float(f(M.accuracy(preds, vt))), c["accuracy"], rtol=1e-6)
    np.testing.assert_allclose(float(f(M.brier_score(vp, vt))), c["brier"], rtol=1e-5)
