#!/usr/bin/env python
"""Generates the committed golden fixtures of tests/golden/.

Two kinds of vectors, kept apart because they carry different authority:

* ``kat.json`` - KNOWN ANSWERS that do not come from this repository: the Random123 threefry2x32 test
  vectors, the values printed in JAX's public documentation for ``PRNGKey(0)`` / ``PRNGKey(42)``, and the
  value-level assertions of the reference's own tests for this path
  (tests/fortuna/prob_model/test_step_schedule.py:21-50, tests/fortuna/test_prob_output_layer.py:38-43,76-81,
  tests/fortuna/test_conformal_methods.py:37-53).  They are written out literally below; nothing is computed.
  The oracle AND the CUDA library are both checked against them.

* ``oracle_vectors.npz`` - seeded inputs and the ORACLE's outputs for them (sampler trajectories, predictive
  reductions, APS / conformal / ECE / quantile).  The reference is pure Python on jax/flax/optax, none of
  which can be installed in the build container (no wheels, no network), so these could not be produced by
  importing /root/reference; they freeze the oracle's behaviour (so a later edit of the oracle or of a kernel
  that changes a result is caught on both sides) and are "parity unpinned" in the sense of oracle/__init__.py.

Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

KAT = {
    "_source": {
        "threefry2x32": "Random123 kat_vectors (threefry2x32, 20 rounds): key, counter -> output",
        "jax_docs": "values printed in the JAX documentation (jax.random module docs, 'Pseudorandom numbers' tutorial)",
        "reference_tests": "value assertions of the reference's own tests; file:line per entry",
    },
    "threefry2x32": [
        {"key": [0x13198A2E, 0x03707344], "ctr": [0x243F6A88, 0x85A308D3], "out": [0xC4923A9C, 0x483DF7A0]},
        {"key": [0, 0], "ctr": [0, 0], "out": [0x6B200159, 0x99BA4EFE]},
        {"key": [0xFFFFFFFF, 0xFFFFFFFF], "ctr": [0xFFFFFFFF, 0xFFFFFFFF], "out": [0x1CB996FC, 0xBB002BE7]},
    ],
    "jax_docs": {
        "split_key0": [[4146024105, 967050713], [2718843009, 1272950319]],
        "uniform_key0": 0.41845703,
        "normal_key0": -0.20584226,
        "normal_key0_n3": [1.8160863, -0.48262316, 0.33988908],
        "normal_key42": -0.18471177,
        "split_key42": [[2465931498, 3679230171], [255383827, 267815257]],
        "normal_subkey42": 1.3694694,
        "split_key42_n3": [[3134548294, 3733159049], [3746501087, 894150801], [801545058, 2363201431]],
        "fold_in_key0_1": [928981903, 3453687069],
    },
    "reference_tests": {
        "cosine_schedule": {
            "cite": "tests/fortuna/prob_model/test_step_schedule.py:24-29",
            "init_step_size": 1e-1, "total_steps": 10,
            "step_value": [[0, 1e-1], [10, 0.0], [20, 1e-1]],
        },
        "constant_schedule_with_cosine_burnin": {
            "cite": "tests/fortuna/prob_model/test_step_schedule.py:35-42",
            "init_step_size": 1e-1, "final_step_size": 1e-2, "burnin_steps": 10,
            "step_value": [[0, 1e-1], [10, 1e-2], [11, 1e-2]],
        },
        "cyclical_cosine_schedule_with_const_burnin": {
            "cite": "tests/fortuna/prob_model/test_step_schedule.py:44-50",
            "init_step_size": 1e-1, "burnin_steps": 10, "cycle_length": 10,
            "step_value": [[0, 1e-1], [1, 1e-1]], "step_differs_from_init": [12],
        },
        "classification_log_prob": {
            "cite": "tests/fortuna/test_prob_output_layer.py:38-43",
            "outputs": [[1.0, 1.0]], "target": [0], "expect": "1 - logsumexp([1, 1])",
            "value": float(1.0 - np.log(2 * np.e)),
        },
        "regression_log_prob": {
            "cite": "tests/fortuna/test_prob_output_layer.py:76-81",
            "outputs": [[1.0, 1.0]], "target": [[0.0]], "expect": "-0.5 * (log(2 pi) + 1 + exp(-1))",
            "value": float(-0.5 * (np.log(2 * np.pi) + 1.0 + np.exp(-1.0))),
        },
        "simple_conformal_sets": {
            "cite": "tests/fortuna/test_conformal_methods.py:37-53",
            "val_probs": [[0.1, 0.9], [0.8, 0.2], [0.3, 0.7]], "val_targets": [1, 1, 0],
            "test_probs": [[0.5, 0.5], [0.8, 0.2]], "error": 0.05,
            "must_contain": [[0, 1], [0]],
        },
    },
}


def oracle_vectors():
    from oracle import predictive as P
    from oracle import prng
    from oracle import sampler as S
    from oracle import scoring as Sc
    from tests import trees

    out = {}
    # --- PRNG stream
    out["bits_seed7_n9"] = prng.random_bits(prng.PRNGKey(7), 9)
    out["normal_seed7_n9"] = prng.normal(prng.PRNGKey(7), 9)
    out["choice_seed11_30of10"] = np.array([prng.choice(k, 10) for k in prng.split(prng.PRNGKey(11), 30)], np.int32)

    # --- sampler: 3 steps on the two-moons MLP tree, grad = 0.5 p + 0.01 (same rule as smoke())
    def run(kind, precond):
        params = trees.mlp_two_moons_tree(0)
        pre = S.Preconditioner(precond)
        st = S.sghmc_init(params, prng.PRNGKey(0), pre)
        p = params
        for _ in range(3):
            g = S.tree_map(lambda a: (a * 0.5 + 0.01).astype(np.float32), p)
            if kind == "sghmc":
                upd, st = S.sghmc_update(g, st, 0.01, S.constant_schedule(4e-6), pre)
            else:
                upd, st = S.cyclical_sgld_update(g, st, 1e-3, 4, 0.25, pre)
            p = S.apply_updates(p, upd)
        flat = np.concatenate([a.ravel() for _, a in S.tree_flatten(p)])
        mom = np.concatenate([a.ravel() for _, a in S.tree_flatten(st.momentum)])
        return flat, mom, st.rng_key

    for kind in ("sghmc", "csgld"):
        for precond in ("identity", "rmsprop"):
            flat, mom, key = run(kind, precond)
            out[f"{kind}_{precond}_params"] = flat
            out[f"{kind}_{precond}_momentum"] = mom
            out[f"{kind}_{precond}_key"] = key

    # --- predictive reductions on z[5, 12, 7]
    r = np.random.default_rng(123)
    z = (3 * r.standard_normal((5, 12, 7))).astype(np.float32)
    t = r.integers(0, 7, 12).astype(np.int32)
    out["z"], out["z_targets"] = z, t
    out["mean"] = P.cls_mean(z)
    out["mode"] = P.cls_mode(z).astype(np.int32)
    out["aleatoric_variance"] = P.cls_aleatoric_variance(z)
    out["epistemic_variance"] = P.cls_epistemic_variance(z)
    out["entropy"] = P.cls_entropy(z)
    out["aleatoric_entropy"] = P.cls_aleatoric_entropy(z)
    out["epistemic_entropy"] = P.cls_epistemic_entropy(z)
    out["log_prob"] = P.cls_log_prob(z, t)
    out["ensemble_log_prob"] = P.cls_ensemble_log_prob(z, t)

    # --- scoring on the mean (tie-heavy variant: probabilities rounded to 1/16)
    mean = out["mean"]
    tie = (np.round(mean * 16) / 16).astype(np.float32)
    for name, pr in (("aps", mean), ("aps_ties", tie)):
        out[f"{name}_perm"] = Sc.descending_perm(pr).astype(np.int32)
        out[f"{name}_all"] = Sc.aps_cumsums(pr)
        out[f"{name}_target"] = Sc.aps_score(pr, t)
    out["tie_probs"] = tie
    val_scores = Sc.aps_score(mean, t)
    conds, sizes, _ = Sc.conformal_sets(val_scores, Sc.aps_cumsums(mean), 0.2)
    out["conformal_mask"], out["conformal_sizes"] = conds.astype(np.uint8), sizes.astype(np.int32)
    counts, confs, accs, bin_idx = Sc.compute_counts_confs_accs(mean.argmax(-1), mean, t)
    out["ece_counts"], out["ece_confs"], out["ece_accs"] = counts.astype(np.int32), confs, accs
    out["ece_bin_idx"] = bin_idx.astype(np.int32)
    out["ece"] = np.float32(Sc.expected_calibration_error(mean.argmax(-1), mean, t))
    out["mce"] = np.float32(Sc.maximum_calibration_error(mean.argmax(-1), mean, t))
    out["quantile_err0.1"] = np.float32(Sc.conformal_quantile(val_scores, 0.1))

    # --- widened rows of the scope: diagnostics, conformal Bayes, target samples, MC entropies, calibration
    from oracle import calibration as Cal
    from oracle import diagnostic as D

    r = np.random.default_rng(321)
    xs = (1.0 + 0.1 * r.standard_normal((12, 9))).astype(np.float32)
    gs = r.standard_normal((12, 9)).astype(np.float32)
    out["diag_samples"], out["diag_grads"] = xs, gs
    out["ksd_imq"] = np.float32(D.kernel_stein_discrepancy_imq(xs, gs, 1.0, -0.5, dtype=np.float64))
    chain = np.cumsum(r.standard_normal((40, 6)), axis=0).astype(np.float32) * 0.3 + r.standard_normal((40, 6)).astype(np.float32)
    out["ess_chain"] = chain
    out["ess_direct"] = D.effective_sample_size_direct(chain, 0.0).astype(np.float32)
    ell = P.cls_ensemble_log_prob(z, t)                           # [5, 12] training log-probs of the fixture logits
    lsm = P.log_softmax_rows((2 * r.standard_normal((5, 4, 7))).astype(np.float32))
    out["cb_lsm_test"] = lsm
    out["cb_counts"] = Sc.conformal_bayes_counts(ell, lsm, dtype=np.float64)[0].astype(np.int32)
    out["cb_region_err0.3"] = Sc.conformal_bayes_region(ell, lsm, 0.3, dtype=np.float64).astype(np.uint8)
    y, idx = P.cls_sample_targets(z, prng.PRNGKey(5), 6)
    out["cls_sample_y"], out["cls_sample_idx"] = y, idx
    zr = np.concatenate([r.standard_normal((4, 10, 2)), 0.3 * r.standard_normal((4, 10, 2)) - 1.0], -1).astype(np.float32)
    out["reg_outputs"] = zr
    e = P.reg_mc_entropies(zr, prng.PRNGKey(9), 5, 4)
    for k, v in e.items():
        out[f"reg_{k}"] = v
    L, dL = Cal.calib_terms(z, t, np.float32(0.4))
    out["calib_L_lt0.4"], out["calib_dL_lt0.4"] = L, dL

    # --- output calibration models and the MaxCovFixPrec calibrator
    from oracle import maxcov as MC
    from oracle import output_calib as OC

    z0 = z[0]                                                      # [12, 7] logits of the fixture
    out["oc_focal_g2_lt0.3"] = np.array(OC.focal_loss(z0, t, 0.3, 2.0), np.float64)
    out["oc_focal_g0_lt0"] = np.array(OC.focal_loss(z0, t, 0.0, 0.0), np.float64)
    yr = r.standard_normal((10, 2)).astype(np.float32)
    out["oc_reg_targets"] = yr
    out["oc_scaled_mse_lt0.2"] = np.array(OC.scaled_mse(zr[0], yr, 0.2), np.float64)
    out["oc_cls_sample"] = OC.cls_output_sample(z0, prng.PRNGKey(13), 5, np.array([0.3], np.float32))
    out["oc_reg_sample"] = OC.reg_output_sample(zr[0], prng.PRNGKey(13), 3, np.array([0.2], np.float32))
    out["oc_traj_cls"] = OC.calibrate(z0, t, 10)[1]
    out["oc_traj_reg"] = OC.calibrate(zr[0], yr, 10, regression=True)[1]
    pb = r.uniform(0, 1, 200).astype(np.float32)
    yb = (r.uniform(0, 1, 200) < pb).astype(np.int32)              # labels drawn from the scores: a calibrated classifier
    out["mc_probs"], out["mc_targets"] = pb, yb
    tp, tn, vp, vn = MC.calibrate(yb, pb, 0.8, 0.8, n_taus=40)
    out["mc_taus"], out["mc_values_pos"], out["mc_values_neg"] = np.array([tp, tn], np.float32), vp, vn
    out["mc_patched"] = MC.apply_patches(pb, tp, tn)
    # --- calibration models (loss heads, Adam) and multivalid scorers
    from oracle import calib_model as CM
    from oracle import multivalid as MV

    zc = (2 * r.standard_normal((40, 6))).astype(np.float32)
    yc = r.integers(0, 6, 40).astype(np.int32)
    out["cm_logits"], out["cm_targets"] = zc, yc
    out["cm_focal_g2"] = np.array(CM.focal_loss(zc, yc, 2.0), np.float32)
    out["cm_focal_grad_g2"] = CM.focal_loss_grad(zc, yc, 2.0).astype(np.float32)
    ad = CM.Adam(1e-2)
    pa, ga = zc.ravel().copy(), CM.focal_loss_grad(zc, yc, 2.0).astype(np.float32).ravel()
    st = ad.init(pa)
    for _ in range(3):
        pa, st = ad.step(pa, ga, st)
    out["cm_adam3"] = pa
    sm0 = r.random(300).astype(np.float32)
    xm = np.clip(sm0 + 0.2, 0, 1).astype(np.float32)  # values biased upwards against the scores: something to patch
    sm = np.clip(sm0 + 0.1 * r.standard_normal(300), 0, 1).astype(np.float32)
    gm = r.random((300, 2)) < 0.5
    out["mv_values"], out["mv_scores"], out["mv_groups"] = xm, sm, gm
    out["mv_perm_seed2_n300"] = MV.permutation(2, 300).astype(np.int32)
    it = MV.Iterative("multicalibrator", seed=2)
    stt = it.calibrate(sm, gm, values=xm, n_rounds=4, n_buckets=8, eta=0.5)
    out["mv_mc_losses"] = np.array(stt["losses"], np.float32)
    out["mv_mc_patches"] = np.array(it.patches, np.float64)
    out["mv_mc_applied"] = it.apply_patches(gm, xm)
    bm = MV.Iterative("batchmvp", seed=2)
    stb = bm.calibrate(sm, gm, n_rounds=4, n_buckets=8, eta=0.5, coverage=0.9, bucket_types=(">=", "<="))
    out["mv_bm_losses"] = np.array(stb["losses"], np.float32)
    out["mv_bm_patches"] = np.array(bm.patches, np.float64)
    out["mv_oneshot_patches"] = MV.one_shot_patches(sm, np.round(xm, 1), n_buckets=11)
    return out


if __name__ == "__main__":
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(KAT, f, indent=1)
    vec = oracle_vectors()
    np.savez(os.path.join(HERE, "oracle_vectors.npz"), **vec)
    print("wrote kat.json and oracle_vectors.npz:", sum(v.nbytes for v in vec.values()), "bytes of arrays")
