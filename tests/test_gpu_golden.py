"""CUDA library (through the C ABI) vs the committed golden fixtures of tests/golden/: the external known
answers of kat.json and the frozen seeded vectors of oracle_vectors.npz.  Bit-exact for keys, indices,
permutations, masks, sizes and bins; rtol 1e-5 for float32 outputs (north_star tolerance)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "kat.json")))
VEC = np.load(os.path.join(HERE, "golden", "oracle_vectors.npz"))


def _u32(t):
    return t.detach().cpu().numpy().view(np.uint32)


def test_jax_documented_values(cuda):
    from fortuna_b200 import ops

    d = KAT["jax_docs"]
    k0, k42 = ops.PRNGKey(0, cuda), ops.PRNGKey(42, cuda)
    assert _u32(ops.split(k0)).tolist() == d["split_key0"]
    assert _u32(ops.split(k42)).tolist() == d["split_key42"]
    assert _u32(ops.split(k42, 3)).tolist() == d["split_key42_n3"]
    assert _u32(ops.fold_in(k0, 1)).tolist() == d["fold_in_key0_1"]
    np.testing.assert_allclose(ops.random_normal(k0, 1).cpu().numpy(), [d["normal_key0"]], rtol=2e-6)
    np.testing.assert_allclose(ops.random_normal(k0, 3).cpu().numpy(), d["normal_key0_n3"], rtol=2e-6)
    np.testing.assert_allclose(ops.random_normal(k42, 1).cpu().numpy(), [d["normal_key42"]], rtol=2e-6)
    sub = ops.key_from_host(d["split_key42"][1], cuda)
    np.testing.assert_allclose(ops.random_normal(sub, 1).cpu().numpy(), [d["normal_subkey42"]], rtol=2e-6)


def test_threefry_kat_through_fold_in(cuda):
    """fold_in(key, d) = threefry(key, (0, d)): reaches the raw block function through the ABI; the Random123
    vectors with counter word 0 == 0 are checked directly, the rest through the oracle-pinned stream."""
    from fortuna_b200 import ops

    for v in KAT["threefry2x32"]:
        if v["ctr"][0] != 0:
            continue
        key = ops.key_from_host(v["key"], cuda)
        assert _u32(ops.fold_in(key, v["ctr"][1])).tolist() == v["out"]


def test_reference_test_values(cuda):
    from fortuna_b200 import ops
    from fortuna_b200.prob_model.posterior.sgmcmc import sgmcmc_step_schedule as sched

    r = KAT["reference_tests"]
    c = r["cosine_schedule"]
    f = sched.cosine_schedule(c["init_step_size"], c["total_steps"])
    for step, val in c["step_value"]:
        np.testing.assert_allclose(f(step), val, rtol=1e-5, atol=1e-8)
    c = r["constant_schedule_with_cosine_burnin"]
    f = sched.constant_schedule_with_cosine_burnin(c["init_step_size"], c["final_step_size"], c["burnin_steps"])
    for step, val in c["step_value"]:
        np.testing.assert_allclose(f(step), val, rtol=1e-5, atol=1e-8)
    c = r["cyclical_cosine_schedule_with_const_burnin"]
    f = sched.cyclical_cosine_schedule_with_const_burnin(c["init_step_size"], c["burnin_steps"], c["cycle_length"])
    for step, val in c["step_value"]:
        np.testing.assert_allclose(f(step), val, rtol=1e-5, atol=1e-8)
    for step in c["step_differs_from_init"]:
        assert not np.allclose(f(step), c["init_step_size"])

    c = r["classification_log_prob"]
    z = torch.tensor([c["outputs"]], dtype=torch.float32, device=cuda)
    t = torch.tensor(c["target"], dtype=torch.int32, device=cuda)
    out = ops.bma_reduce_cls(z, ("log_prob", "ensemble_log_prob"), targets=t)
    np.testing.assert_allclose(out["ensemble_log_prob"].cpu().numpy()[0], [c["value"]], rtol=1e-5)
    np.testing.assert_allclose(out["log_prob"].cpu().numpy(), [c["value"]], rtol=1e-5)
    c = r["regression_log_prob"]
    z = torch.tensor([c["outputs"]], dtype=torch.float32, device=cuda)
    t = torch.tensor(c["target"], dtype=torch.float32, device=cuda)
    out = ops.bma_reduce_reg(z, ("log_prob",), targets=t)
    np.testing.assert_allclose(out["log_prob"].cpu().numpy(), [c["value"]], rtol=1e-5)

    c = r["simple_conformal_sets"]
    vp = torch.tensor(c["val_probs"], dtype=torch.float32, device=cuda)
    vt = torch.tensor(c["val_targets"], dtype=torch.int32, device=cuda)
    tp = torch.tensor(c["test_probs"], dtype=torch.float32, device=cuda)
    val = ops.sort_f32_(ops.simple_scores(vp, vt))
    thr = np.float32((1 - c["error"]) * (len(c["val_probs"]) + 1))
    mask, sizes = ops.conformal_sets(val, ops.simple_scores(tp, None), thr)
    mask = mask.cpu().numpy().astype(bool)
    for b, must in enumerate(c["must_contain"]):
        assert all(mask[b, k] for k in must)


def _run_sampler(cuda, kind, precond):
    from fortuna_b200 import ops
    from fortuna_b200.prob_model.posterior.sgmcmc import sgmcmc_preconditioner as pc
    from fortuna_b200.prob_model.posterior.sgmcmc import sgmcmc_step_schedule as sched
    from fortuna_b200.prob_model.posterior.sgmcmc.cyclical_sgld.cyclical_sgld_integrator import cyclical_sgld_integrator
    from fortuna_b200.prob_model.posterior.sgmcmc.sghmc.sghmc_integrator import sghmc_integrator
    from fortuna_b200.utils.flat_tree import FlatTree
    from tests import trees

    pre = pc.rmsprop_preconditioner() if precond == "rmsprop" else pc.identity_preconditioner()
    if kind == "sghmc":
        tx = sghmc_integrator(0.01, None, ops.PRNGKey(0, cuda), sched.constant_schedule(4e-6), pre)
    else:
        tx = cyclical_sgld_integrator(ops.PRNGKey(0, cuda), 1e-3, 4, 0.25, pre)
    params = FlatTree.from_tree(trees.mlp_two_moons_tree(0), device=cuda)
    state = tx.init(params)
    for _ in range(3):
        g = params.empty_like()
        g.flat.copy_(params.flat * 0.5 + 0.01)  # input construction only (same rule as the fixture script)
        params, state = tx.apply(params, g, state)
    inner = state if kind == "sghmc" else state.sgld_state
    flat = np.concatenate([v.cpu().numpy().ravel() for v in params.leaf_views()])
    mom = np.concatenate([v.cpu().numpy().ravel() for v in inner.momentum.leaf_views()])
    return flat, mom, ops.key_to_host(inner.rng_key)


@pytest.mark.parametrize("kind", ["sghmc", "csgld"])
@pytest.mark.parametrize("precond", ["identity", "rmsprop"])
def test_sampler_trajectories(cuda, kind, precond):
    flat, mom, key = _run_sampler(cuda, kind, precond)
    assert np.array_equal(key, VEC[f"{kind}_{precond}_key"])
    np.testing.assert_allclose(flat, VEC[f"{kind}_{precond}_params"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(mom, VEC[f"{kind}_{precond}_momentum"], rtol=1e-5, atol=1e-7)


def test_prng_vectors(cuda):
    from fortuna_b200 import ops

    assert np.array_equal(_u32(ops.random_bits(ops.PRNGKey(7, cuda), 9)), VEC["bits_seed7_n9"])
    np.testing.assert_allclose(ops.random_normal(ops.PRNGKey(7, cuda), 9).cpu().numpy(), VEC["normal_seed7_n9"], rtol=1e-5, atol=1e-6)
    assert np.array_equal(ops.sample_indices(ops.PRNGKey(11, cuda), 30, 10).cpu().numpy(), VEC["choice_seed11_30of10"])


def test_predictive_and_scoring_vectors(cuda):
    from fortuna_b200 import ops

    z = torch.from_numpy(VEC["z"]).to(cuda)
    t = torch.from_numpy(VEC["z_targets"]).to(cuda)
    names = ("mean", "mode", "aleatoric_variance", "epistemic_variance", "entropy", "aleatoric_entropy",
             "epistemic_entropy", "log_prob", "ensemble_log_prob")
    out = ops.bma_reduce_cls(z, names, targets=t)
    assert np.array_equal(out["mode"].cpu().numpy(), VEC["mode"])
    for n in names:
        if n == "mode":
            continue
        atol = 2e-6 if "entropy" in n or "log_prob" in n else 2e-7
        np.testing.assert_allclose(out[n].cpu().numpy(), VEC[n], rtol=1e-5, atol=atol, err_msg=n)

    # scoring runs on the FIXTURE's mean so that integer outputs can be compared bit for bit
    mean = torch.from_numpy(VEC["mean"]).to(cuda)
    for name, probs in (("aps", mean), ("aps_ties", torch.from_numpy(VEC["tie_probs"]).to(cuda))):
        allc, perm = ops.aps_scores(probs, None, want_perm=True)
        assert np.array_equal(perm.cpu().numpy(), VEC[f"{name}_perm"]), name
        assert np.array_equal(allc.cpu().numpy(), VEC[f"{name}_all"]), name
        assert np.array_equal(ops.aps_scores(probs, t).cpu().numpy(), VEC[f"{name}_target"]), name
    val = ops.sort_f32_(ops.aps_scores(mean, t))
    B = mean.shape[0]
    mask, sizes = ops.conformal_sets(val, ops.aps_scores(mean, None), np.float32((1 - 0.2) * (B + 1)))
    assert np.array_equal(mask.cpu().numpy(), VEC["conformal_mask"])
    assert np.array_equal(sizes.cpu().numpy(), VEC["conformal_sizes"])
    from fortuna_b200.prob_model.predictive.pipeline import ece_thresholds

    e = ops.ece_bins(mean, t, torch.from_numpy(ece_thresholds(B)).to(cuda), want_bin_idx=True)
    assert np.array_equal(e["bin_idx"].cpu().numpy(), VEC["ece_bin_idx"])
    assert np.array_equal(e["counts"].cpu().numpy(), VEC["ece_counts"])
    res = e["result"].cpu().numpy()
    np.testing.assert_allclose(res[0], VEC["ece"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(res[1], VEC["mce"], rtol=1e-5, atol=1e-7)
    qlevel = np.float32(np.ceil(np.float32((B + 1) * (1 - 0.1))) / np.float32(B))
    q = ops.quantile_sorted(val, torch.tensor([qlevel], dtype=torch.float32, device=cuda))
    np.testing.assert_allclose(q.cpu().numpy()[0], VEC["quantile_err0.1"], rtol=1e-6)


def test_widened_rows_vectors(cuda):
    """The frozen vectors of the widened scope rows (diagnostics, conformal Bayes, target samples, regression
    Monte-Carlo entropies, calibration terms) through the C ABI."""
    from fortuna_b200 import ops

    dev = cuda
    x, g = torch.from_numpy(VEC["diag_samples"]).to(dev), torch.from_numpy(VEC["diag_grads"]).to(dev)
    np.testing.assert_allclose(float(ops.ksd_imq(x, g, 1.0, -0.5).item()), VEC["ksd_imq"], rtol=2e-5)
    np.testing.assert_allclose(ops.ess(torch.from_numpy(VEC["ess_chain"]).to(dev), 0.0).cpu().numpy(), VEC["ess_direct"], rtol=5e-4)
    z = torch.from_numpy(VEC["z"]).to(dev)
    t = torch.from_numpy(VEC["z_targets"]).to(dev)
    ell = ops.bma_reduce_cls(z, ("ensemble_log_prob",), targets=t)["ensemble_log_prob"]
    cb = ops.conformal_bayes_sets(ell, torch.from_numpy(VEC["cb_lsm_test"]).to(dev), 0.3, want_counts=True)
    assert np.array_equal(cb["rank_counts"].cpu().numpy(), VEC["cb_counts"])
    assert np.array_equal(cb["region"].cpu().numpy(), VEC["cb_region_err0.3"])
    y, idx = ops.cls_sample_targets(z, ops.PRNGKey(5, dev), 6, want_idx=True)
    assert np.array_equal(idx.cpu().numpy(), VEC["cls_sample_idx"]) and np.array_equal(y.cpu().numpy(), VEC["cls_sample_y"])
    zr = torch.from_numpy(VEC["reg_outputs"]).to(dev)
    ks = ops.split(ops.PRNGKey(9, dev), 6)
    e = ops.reg_mc_entropies(zr, ops.sample_indices(ks[0].contiguous(), 5, 4), ks[1:].contiguous(), 4,
                             ("entropy", "aleatoric_entropy", "epistemic_entropy"))
    for k in ("entropy", "aleatoric_entropy", "epistemic_entropy"):
        np.testing.assert_allclose(e[k].cpu().numpy(), VEC[f"reg_{k}"], rtol=2e-5, atol=2e-5, err_msg=k)
    terms = ops.calib_cls_loss_terms(z, t, torch.tensor([0.4], dtype=torch.float32, device=dev)).cpu().numpy()
    np.testing.assert_allclose(terms[0], VEC["calib_L_lt0.4"], rtol=2e-6, atol=1e-4)
    np.testing.assert_allclose(terms[1], VEC["calib_dL_lt0.4"], rtol=2e-5, atol=1e-4)


def test_output_calib_and_maxcov_vectors(cuda):
    """Frozen vectors of the output calibration models (loss terms, Adam trajectories, output-layer target samples) and
    of the MaxCovFixPrec calibrator, through the C ABI and the user classes."""
    from fortuna_b200 import ops
    from fortuna_b200.conformal import MaxCoverageFixedPrecisionBinaryClassificationCalibrator as Cal
    from fortuna_b200.output_calib_model import Config, Monitor, Optimizer, OutputCalibClassifier, OutputCalibRegressor

    dev = cuda
    z0 = torch.from_numpy(VEC["z"][0].copy()).to(dev)
    t = torch.from_numpy(VEC["z_targets"]).to(dev)
    n = z0.shape[0]
    for key, lt, gamma in (("oc_focal_g2_lt0.3", 0.3, 2.0), ("oc_focal_g0_lt0", 0.0, 0.0)):
        terms = ops.output_calib_focal_loss_terms(z0, t, torch.tensor([lt], dtype=torch.float32, device=dev), gamma).cpu().numpy()
        np.testing.assert_allclose(terms[0] / n, VEC[key][0], rtol=2e-5)
        np.testing.assert_allclose(-np.exp(-lt) * terms[1] / n, VEC[key][1], rtol=2e-4, atol=1e-6)
    zr = torch.from_numpy(VEC["reg_outputs"][0].copy()).to(dev)
    yr = torch.from_numpy(VEC["oc_reg_targets"]).to(dev)
    terms = ops.output_calib_scaled_mse_terms(zr, yr, torch.tensor([0.2], dtype=torch.float32, device=dev)).cpu().numpy()
    np.testing.assert_allclose(terms / zr.shape[0], VEC["oc_scaled_mse_lt0.2"], rtol=2e-5)
    got = ops.cls_output_sample_targets(z0, ops.PRNGKey(13, dev), 5, torch.tensor([0.3], dtype=torch.float32, device=dev))
    assert np.array_equal(got.cpu().numpy(), VEC["oc_cls_sample"])
    got = ops.reg_output_sample_targets(zr, ops.PRNGKey(13, dev), 3, torch.tensor([0.2], dtype=torch.float32, device=dev))
    np.testing.assert_allclose(got.cpu().numpy(), VEC["oc_reg_sample"], rtol=1e-5, atol=2e-6)
    cfg = Config(optimizer=Optimizer(n_epochs=10), monitor=Monitor(verbose=False))
    clf = OutputCalibClassifier()
    clf._check_output_dim = lambda outputs, targets: None  # the 12 fixture targets do not cover all 7 classes
    clf.calibrate(VEC["z"][0], VEC["z_targets"], config=cfg)
    assert abs(float(clf.predictive.state.get().log_temp[0]) - float(VEC["oc_traj_cls"][-1])) < 1e-5
    reg = OutputCalibRegressor()
    reg.calibrate(VEC["reg_outputs"][0], VEC["oc_reg_targets"], config=cfg)
    assert abs(float(reg.predictive.state.get().log_temp[0]) - float(VEC["oc_traj_reg"][-1])) < 1e-5
    cal = Cal()
    patched = cal.calibrate(targets=VEC["mc_targets"], probs=VEC["mc_probs"], true_positive_precision_threshold=0.8,
                            true_negative_precision_threshold=0.8, test_probs=VEC["mc_probs"], n_taus=40)
    assert cal.patches["tau_pos"] == VEC["mc_taus"][0] and cal.patches["tau_neg"] == VEC["mc_taus"][1]
    assert np.array_equal(cal._values[0], VEC["mc_values_pos"]) and np.array_equal(cal._values[1], VEC["mc_values_neg"])
    assert np.array_equal(patched.cpu().numpy(), VEC["mc_patched"])


def test_calib_model_and_multivalid_against_frozen_vectors(cuda):
    """The loss head, the fused Adam and the multivalid classes on the frozen inputs of oracle_vectors.npz."""
    from fortuna_b200 import ops
    from fortuna_b200.conformal import BatchMVPConformalRegressor, Multicalibrator, OneShotMulticalibrator
    from fortuna_b200.conformal.multivalid.iterative.base import jax_permutation

    z = torch.from_numpy(VEC["cm_logits"]).to(cuda)
    y = torch.from_numpy(VEC["cm_targets"]).to(cuda)
    loss, grad, _ = ops.focal_loss_grad(z, y, 2.0)
    np.testing.assert_allclose(loss.item(), float(VEC["cm_focal_g2"]), rtol=1e-5)
    np.testing.assert_allclose(grad.cpu().numpy(), VEC["cm_focal_grad_g2"], rtol=2e-4, atol=1e-8)
    p = z.reshape(-1).clone()
    g = torch.from_numpy(VEC["cm_focal_grad_g2"].ravel().copy()).to(cuda)
    mu, nu = torch.zeros_like(p), torch.zeros_like(p)
    for step in (1, 2, 3):
        ops.adam_step_(p, g, mu, nu, step, 1e-2)
    assert np.array_equal(p.cpu().numpy(), VEC["cm_adam3"])
    assert np.array_equal(jax_permutation(2, 300, cuda), VEC["mv_perm_seed2_n300"])
    x, s, gr = VEC["mv_values"], VEC["mv_scores"], VEC["mv_groups"]
    mc = Multicalibrator(seed=2)
    st = mc.calibrate(scores=s, groups=gr, values=x, n_rounds=4, n_buckets=8, eta=0.5)
    np.testing.assert_allclose(st["losses"], VEC["mv_mc_losses"], rtol=2e-5)
    np.testing.assert_allclose(np.array(mc.patches, np.float64), VEC["mv_mc_patches"], rtol=2e-4, atol=1e-7)
    np.testing.assert_allclose(mc.apply_patches(gr, x).cpu().numpy(), VEC["mv_mc_applied"], rtol=1e-5, atol=1e-6)
    bm = BatchMVPConformalRegressor(seed=2)
    sb = bm.calibrate(scores=s, groups=gr, n_rounds=4, n_buckets=8, eta=0.5, coverage=0.9)
    np.testing.assert_allclose(sb["losses"], VEC["mv_bm_losses"], rtol=2e-5)
    np.testing.assert_allclose(np.array(bm.patches, np.float64), VEC["mv_bm_patches"], rtol=1e-6, atol=1e-7)
    os_ = OneShotMulticalibrator()
    os_.calibrate(scores=s, values=np.round(x, 1).astype(np.float32), n_buckets=11)
    np.testing.assert_allclose(os_.patches, VEC["mv_oneshot_patches"], rtol=1e-5, atol=1e-7)
