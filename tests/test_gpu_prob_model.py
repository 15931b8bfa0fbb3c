"""The user API on BASELINE config 1 (two-moons MLP ProbClassifier, SGHMC / cyclical-SGLD posterior, predictive
mean / entropy), exercised like tests/fortuna/prob_model/test_train.py:49-60,206-389 does (train -> sample ->
predictive), plus value-level checks the reference's smoke tests do not make."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples"))


def _setup():
    from two_moons_sghmc import MLP, make_moons

    from fortuna_b200.data import DataLoader

    train, test = make_moons(500, 0.07, 0), make_moons(400, 0.07, 2)
    return MLP, train, test, DataLoader.from_array_data(train, batch_size=128, shuffle=True), DataLoader.from_array_data(test, batch_size=128)


def test_sghmc_two_moons_train_and_predict(cuda):
    from fortuna_b200.data import InputsLoader
    from fortuna_b200.metric.classification import accuracy
    from fortuna_b200.prob_model import FitConfig, FitOptimizer, ProbClassifier, SGHMCPosteriorApproximator
    from oracle import predictive as P

    MLP, train, test, train_loader, test_loader = _setup()
    torch.manual_seed(0)
    pm = ProbClassifier(model=MLP(), posterior_approximator=SGHMCPosteriorApproximator(
        n_samples=20, n_thinning=10, burnin_length=600, step_schedule=1e-3), seed=0)
    status = pm.train(train_loader, val_data_loader=test_loader, fit_config=FitConfig(optimizer=FitOptimizer(n_epochs=250)))
    assert status["fit_status"]["loss"].shape == (250,) and status["fit_status"]["val_loss"].shape == (250,)
    assert pm.posterior.state.n_filled == 20
    names = ("mean", "mode", "entropy", "aleatoric_entropy", "epistemic_entropy", "variance", "aleatoric_variance",
             "epistemic_variance", "log_prob", "ensemble_log_prob")
    key = pm.rng.get()
    st = pm.predictive.statistics(test_loader, names, n_posterior_samples=30, rng=key)
    assert st["mean"].shape == (400, 2) and st["ensemble_log_prob"].shape == (30, 400)
    assert float(accuracy(st["mode"], test[1])) > 0.97
    # the same key gives the same draws: each single-statistic method equals the joint pass
    inputs = test_loader.to_inputs_loader()
    assert torch.equal(pm.predictive.mean(inputs, 30, rng=key), st["mean"])
    assert torch.equal(pm.predictive.entropy(inputs, 30, rng=key), st["entropy"])
    assert torch.equal(pm.predictive.log_prob(test_loader, 30, rng=key), st["log_prob"])
    # everything against the oracle on the very logits the predictive used
    z = pm.predictive.sample_calibrated_outputs(inputs, 30, rng=key).cpu().numpy()
    np.testing.assert_allclose(st["mean"].cpu().numpy(), P.cls_mean(z), rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(st["entropy"].cpu().numpy(), P.cls_entropy(z), rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(st["epistemic_entropy"].cpu().numpy(), P.cls_epistemic_entropy(z), rtol=1e-5, atol=4e-6)
    np.testing.assert_allclose(st["log_prob"].cpu().numpy(), P.cls_log_prob(z, test[1]), rtol=1e-5, atol=2e-6)
    # the S draws are the reference's: idx_s = choice(split(key, S)[s], n_samples)
    from oracle import prng as oprng
    from fortuna_b200 import ops

    keys = oprng.split(ops.key_to_host(key), 30)
    want_idx = np.array([oprng.choice(k, 20) for k in keys])
    assert np.array_equal(pm.posterior.sample_indices(key, 30).cpu().numpy(), want_idx)
    # epistemic uncertainty is larger far from the data than on it
    grid = np.stack(np.meshgrid(np.linspace(-4, 5, 20), np.linspace(-4, 4, 20)), -1).reshape(-1, 2).astype(np.float32)
    far = grid[(np.abs(grid[:, 0] - 0.5) > 3) | (np.abs(grid[:, 1]) > 3)]
    ent_far = float(pm.predictive.epistemic_entropy(InputsLoader.from_array_inputs(far), 30).mean())
    assert ent_far > float(st["epistemic_entropy"].mean())
    sets = pm.predictive.credible_set(inputs, 30, error=0.05)
    assert len(sets) == 400 and all(1 <= len(s) <= 2 for s in sets)


def test_last_layer_posterior_uses_batched_contraction(cuda):
    """freeze_fun keeps only output_subnet trainable (examples/scaling_up_bayesian_inference.pct.py:90-104): the
    predictive then runs features once + K2 over the sample bank, and must equal one forward pass per drawn sample."""
    from fortuna_b200.prob_model import CyclicalSGLDPosteriorApproximator, FitConfig, FitOptimizer, ProbClassifier

    MLP, train, test, train_loader, test_loader = _setup()
    torch.manual_seed(1)
    pm = ProbClassifier(model=MLP(), posterior_approximator=CyclicalSGLDPosteriorApproximator(
        n_samples=8, n_thinning=5, cycle_length=40, init_step_size=1e-3, exploration_ratio=0.25), seed=3)
    freeze = lambda path, v: "trainable" if "output_subnet" in path else "frozen"
    pm.train(train_loader, fit_config=FitConfig(optimizer=FitOptimizer(n_epochs=40, freeze_fun=freeze)))
    bound = pm.posterior.bound
    assert bound.is_last_layer_only() and bound.params.n_params == 30 * 2 + 2
    assert pm.posterior.state.n_filled == 8
    key = pm.rng.get()
    inputs = test_loader.to_inputs_loader()
    z = pm.predictive.sample_calibrated_outputs(inputs, 12, rng=key)
    idx = pm.posterior.sample_indices(key, 12).cpu().numpy()
    x = torch.as_tensor(test[0], device=cuda)
    keep = bound.params.flat.clone()
    for s, i in enumerate(idx):
        bound.load_flat(pm.posterior.state.bank[int(i)])
        with torch.no_grad():
            ref = bound.model(x)
        np.testing.assert_allclose(z[s].cpu().numpy(), ref.cpu().numpy(), rtol=1e-4, atol=1e-5)
    bound.load_flat(keep)


def test_rejections(cuda):
    from fortuna_b200.prob_model import ProbClassifier, SGHMCPosteriorApproximator

    MLP = _setup()[0]
    with pytest.raises(NotImplementedError):
        ProbClassifier(model=MLP(), posterior_approximator="advi")
    with pytest.raises(ValueError):  # too few steps to ever collect n_samples (sghmc_callback.py:55-64)
        from fortuna_b200.data import DataLoader
        from fortuna_b200.prob_model import FitConfig, FitOptimizer

        _, train, _, train_loader, _ = _setup()
        ProbClassifier(model=MLP(), posterior_approximator=SGHMCPosteriorApproximator(n_samples=10, burnin_length=1000)).train(
            train_loader, fit_config=FitConfig(optimizer=FitOptimizer(n_epochs=2)))


@pytest.mark.parametrize("last_layer_only", [True, False])
def test_checkpoints_in_reference_layout_round_trip(cuda, tmp_path, last_layer_only):
    """FitCheckpointer(save_checkpoint_dir=...) writes ``<dir>/c/checkpoint_<step>`` (chain) and
    ``<dir>/<i>/checkpoint_<step>`` (samples) as flax-msgpack state dicts (sghmc_posterior.py:113-156,
    posterior_multi_state_repository.py:20-29); a fresh ProbClassifier.load_state(dir) predicts the same bits."""
    from fortuna_b200.prob_model import (
        FitCheckpointer,
        FitConfig,
        FitOptimizer,
        ProbClassifier,
        SGHMCPosteriorApproximator,
    )
    from fortuna_b200.training import checkpoint as ck

    MLP, train, test, train_loader, test_loader = _setup()
    torch.manual_seed(0)
    approx = dict(n_samples=5, n_thinning=4, burnin_length=20, step_schedule=1e-3)
    freeze = (lambda path, v: "trainable" if "output_subnet" in path else "frozen") if last_layer_only else None
    d = str(tmp_path / "ckpt")
    pm = ProbClassifier(model=MLP(), posterior_approximator=SGHMCPosteriorApproximator(**approx), seed=0)
    pm.train(train_loader, fit_config=FitConfig(
        optimizer=FitOptimizer(n_epochs=12, freeze_fun=freeze),
        checkpointer=FitCheckpointer(save_checkpoint_dir=d, save_every_n_steps=10, keep_top_n_checkpoints=2)))
    assert sorted(os.listdir(d)) == ["0", "1", "2", "3", "4", "c"]
    assert len(os.listdir(os.path.join(d, "c"))) == 2  # keep_top_n_checkpoints
    chain = ck.restore_checkpoint(os.path.join(d, "c"))
    assert ck.decode_name(chain) == "SGHMCState" and chain["step"] == 12 * len(train_loader)
    assert int(chain["opt_state"]["count"]) == chain["step"] and chain["opt_state"]["rng_key"].dtype == np.uint32
    s0 = ck.restore_checkpoint(os.path.join(d, "0"))
    assert s0["step"] == 24  # first stored step: burn-in 20, then every 4th
    leaves = s0["params"]["model"]["params"]
    if last_layer_only:
        assert set(leaves) == {"output_subnet"}  # only the sampled leaves are stored per sample (:139-157)
        assert len(chain["params"]["model"]["params"]) > 1  # the chain keeps the frozen part
        assert chain["_encoded_which_params"] is not None
    np.testing.assert_array_equal(leaves["output_subnet"]["weight"],
                                  pm.posterior.state.get(0).tree()["model"]["params"]["output_subnet"]["weight"].cpu().numpy())
    key = pm.rng.get()
    want = pm.predictive.mean(test_loader.to_inputs_loader(), 16, rng=key)
    torch.manual_seed(123)  # a differently initialised module: everything must come from the files
    pm2 = ProbClassifier(model=MLP(), posterior_approximator=SGHMCPosteriorApproximator(**approx), seed=7)
    pm2.load_state(d)
    assert pm2.posterior.state.n_filled == 5 and pm2.posterior.bound.is_last_layer_only() == last_layer_only
    assert torch.equal(pm2.posterior.state.bank, pm.posterior.state.bank)
    got = pm2.predictive.mean(test_loader.to_inputs_loader(), 16, rng=key)
    assert torch.equal(got, want)
    # save_state writes the same layout somewhere else
    d2 = str(tmp_path / "again")
    pm2.save_state(d2)
    assert sorted(os.listdir(d2)) == ["0", "1", "2", "3", "4", "c"]
    np.testing.assert_array_equal(ck.restore_checkpoint(os.path.join(d2, "3"))["params"]["model"]["params"]["output_subnet"]["bias"],
                                  ck.restore_checkpoint(os.path.join(d, "3"))["params"]["model"]["params"]["output_subnet"]["bias"])
    with pytest.raises(ValueError):
        pm2.load_state(str(tmp_path / "nothing_here"))


def test_sample_bank_loads_reference_style_files(cuda, tmp_path):
    """A directory as Fortuna's SGMCMCPosteriorStateRepository leaves it (flax nn.Dense leaf names, trainable leaves only)
    -> device bank -> LastLayerSamples for the S-batched contraction."""
    from fortuna_b200.prob_model.posterior.sgmcmc.sgmcmc_posterior import SGMCMCPosterior
    from fortuna_b200.prob_model.posterior.sgmcmc.sgmcmc_posterior_state_repository import SGMCMCPosteriorStateRepository
    from fortuna_b200.training import checkpoint as ck
    from fortuna_b200.utils.random import RandomNumberGenerator

    r = np.random.default_rng(0)
    n, D, C = 6, 16, 3
    ws = r.standard_normal((n, D, C)).astype(np.float32)
    bs = r.standard_normal((n, C)).astype(np.float32)
    for i in range(n):
        state = {"step": 100 + i, "params": {"model": {"params": {"output_subnet": {"last": {"kernel": ws[i], "bias": bs[i]}}}}},
                 "opt_state": None, "mutable": None, "calib_params": None, "calib_mutable": None,
                 "encoded_name": ck.to_state_dict(ck.encode_name("SGHMCState"))}
        ck.save_checkpoint(str(tmp_path / str(i)), state, step=100 + i)
    repo = SGMCMCPosteriorStateRepository(n, checkpoint_dir=str(tmp_path)).load_checkpoints()
    assert repo.n_filled == n and repo.steps == [100 + i for i in range(n)]
    post = SGMCMCPosterior(repo, RandomNumberGenerator(0, cuda))
    ll = post.last_layer()
    assert torch.equal(ll.kernel.cpu(), torch.from_numpy(ws)) and torch.equal(ll.bias.cpu(), torch.from_numpy(bs))
    lazy = SGMCMCPosteriorStateRepository(n, checkpoint_dir=str(tmp_path))
    assert torch.equal(lazy.get(4).tree()["model"]["params"]["output_subnet"]["last"]["bias"].cpu(), torch.from_numpy(bs[4]))


def test_fused_prior_gradient_equals_autograd_route(cuda):
    """run_sgmcmc(fuse_prior=True) - scaling and Gaussian-prior gradient inside the sampler kernel, log-prior value from
    the kernel's sum(theta^2) - follows the same chain as the route where autograd differentiates the whole log-joint."""
    from fortuna_b200 import ops
    from fortuna_b200.prob_model.posterior.sgmcmc import sgmcmc_preconditioner as pc
    from fortuna_b200.prob_model.posterior.sgmcmc import sgmcmc_step_schedule as sched
    from fortuna_b200.prob_model.posterior.sgmcmc.sghmc.sghmc_integrator import sghmc_integrator
    from fortuna_b200.prob_model.posterior.sgmcmc.sgmcmc_fit import run_sgmcmc
    from fortuna_b200.prob_model.posterior.sgmcmc.torch_model import BoundModel
    from fortuna_b200.prob_model.prior import IsotropicGaussianPrior
    from fortuna_b200.data import DataLoader

    MLP, train, _, _, _ = _setup()
    loader = DataLoader.from_array_data(train, batch_size=100)  # no shuffling: both runs see the same batches
    out = []
    for fuse in (True, False):
        torch.manual_seed(0)
        bound = BoundModel(MLP(), cuda)
        tx = sghmc_integrator(0.05, None, ops.PRNGKey(4, cuda), sched.constant_schedule(1e-4), pc.rmsprop_preconditioner())
        status, state = run_sgmcmc(bound, IsotropicGaussianPrior(log_var=-1.0), tx, [], loader, n_epochs=3, fuse_prior=fuse)
        out.append((status["loss"], bound.params.flat.clone(), state.opt_state.momentum.flat.clone()))
    np.testing.assert_allclose(out[0][0], out[1][0], rtol=1e-5)
    scale = float(out[1][1].abs().max())
    assert float((out[0][1] - out[1][1]).abs().max()) <= 2e-5 * scale  # 15 chaotic-free steps: rounding-level drift only
    assert float((out[0][2] - out[1][2]).abs().max()) <= 2e-4 * float(out[1][2].abs().max())


def test_conformal_set_of_the_predictive(cuda):
    """ClassificationPredictive.conformal_set (predictive/classification.py:485-626) end to end on two moons: sets are
    what the oracle builds from the very logits the predictive used, sizes are 1 for confident points."""
    from fortuna_b200 import ops
    from fortuna_b200.prob_model import FitConfig, FitOptimizer, ProbClassifier, SGHMCPosteriorApproximator
    from oracle import predictive as P
    from oracle import scoring as Sc

    MLP, train, test, train_loader, test_loader = _setup()
    torch.manual_seed(0)
    pm = ProbClassifier(model=MLP(), posterior_approximator=SGHMCPosteriorApproximator(
        n_samples=10, n_thinning=5, burnin_length=200, step_schedule=1e-3), seed=0)
    pm.train(train_loader, fit_config=FitConfig(optimizer=FitOptimizer(n_epochs=80)))
    key = pm.rng.get()
    inputs = test_loader.to_inputs_loader()
    sets, ess = pm.predictive.conformal_set(train_loader, inputs, n_posterior_samples=20, error=0.05, rng=key, return_ess=True)
    assert len(sets) == 400 and ess.shape == (400, 2, 1)
    z_train = pm.predictive.sample_calibrated_outputs(train_loader.to_inputs_loader(), 20, rng=key).cpu().numpy()
    z_test = pm.predictive.sample_calibrated_outputs(inputs, 20, rng=key).cpu().numpy()
    y = np.asarray(train_loader.to_array_targets())
    ell = np.take_along_axis(P.log_softmax_rows(z_train), np.broadcast_to(y[None, :, None], (20, len(y), 1)), axis=2)[..., 0]
    lsm = P.log_softmax_rows(z_test)
    cnt64, p_train, p_test = Sc.conformal_bayes_counts(ell, lsm, dtype=np.float64)
    amb = (np.abs(p_train - p_test[..., None]) <= 1e-5 * p_test[..., None]).sum(axis=2)
    thr = float(np.float32(0.05 * (len(y) + 1)))
    clear = np.abs(cnt64 + 1 - thr) > amb + 0.5
    want = Sc.conformal_bayes_region(ell, lsm, 0.05, dtype=np.float64)
    got = np.zeros_like(want)
    for i, s in enumerate(sets):
        got[i, s] = True
    assert clear.mean() > 0.95 and np.array_equal(got[clear], want[clear])
    np.testing.assert_allclose(ess[..., 0].cpu().numpy(), Sc.conformal_bayes_ess(lsm, dtype=np.float64), rtol=1e-4)
    assert np.mean([len(s) for s in sets]) < 1.5 and all(s == sorted(s) for s in sets)


def test_predictive_sample_classification(cuda):
    """ClassificationPredictive.sample through the user API == the oracle on the stored logits, batch by batch."""
    from fortuna_b200 import ops
    from fortuna_b200.prob_model import FitConfig, FitOptimizer, ProbClassifier, SGHMCPosteriorApproximator
    from oracle import predictive as P

    MLP, train, test, train_loader, test_loader = _setup()
    torch.manual_seed(0)
    pm = ProbClassifier(model=MLP(), posterior_approximator=SGHMCPosteriorApproximator(
        n_samples=6, n_thinning=5, burnin_length=100, step_schedule=1e-3), seed=0)
    pm.train(train_loader, fit_config=FitConfig(optimizer=FitOptimizer(n_epochs=40)))
    key = pm.rng.get()
    inputs = test_loader.to_inputs_loader()
    y = pm.predictive.sample(inputs, n_target_samples=5, rng=key)
    assert y.shape == (5, 400) and y.dtype == torch.int32
    want = []
    for xb in inputs:
        stored = pm.predictive._logits_fn(None, calibrated=False)(
            torch.as_tensor(np.asarray(xb), dtype=torch.float32, device=cuda)).cpu().numpy()
        want.append(P.cls_sample_targets(stored, ops.key_to_host(key), 5)[0])
    assert np.array_equal(y.cpu().numpy(), np.concatenate(want, axis=1))
    assert float((y[0].cpu() == torch.as_tensor(test[1])).float().mean()) > 0.9  # a trained model mostly samples the label


def test_train_with_calibration_loader(cuda):
    """ProbClassifier.train(calib_data_loader=...) fits the default temperature scaler on a fixed ensemble
    (fortuna/prob_model/base.py:94-235); afterwards every predictive statistic uses the fitted temperature."""
    from fortuna_b200.prob_model import (
        Adam, CalibConfig, CalibOptimizer, CalibProcessor, FitConfig, FitOptimizer, ProbClassifier, SGHMCPosteriorApproximator,
    )
    from oracle import predictive as P

    MLP, train, test, train_loader, test_loader = _setup()
    torch.manual_seed(0)
    pm = ProbClassifier(model=MLP(), posterior_approximator=SGHMCPosteriorApproximator(
        n_samples=8, n_thinning=5, burnin_length=150, step_schedule=1e-3), seed=0)
    status = pm.train(train_loader, val_data_loader=test_loader, calib_data_loader=test_loader,
                      fit_config=FitConfig(optimizer=FitOptimizer(n_epochs=60)),
                      calib_config=CalibConfig(optimizer=CalibOptimizer(Adam(1e-2), n_epochs=15), processor=CalibProcessor(10)))
    cs = status["calib_status"]
    assert cs["loss"].shape == (15,) and cs["val_loss"].shape == (15,) and cs["loss"][-1] <= cs["loss"][0]
    lt = pm.predictive.log_temp
    assert lt is not None and lt.shape == (1,) and float(lt.abs()[0]) > 0
    key = pm.rng.get()
    inputs = test_loader.to_inputs_loader()
    z = pm.predictive.sample_calibrated_outputs(inputs, 12, rng=key)
    keep, pm.predictive.log_temp = pm.predictive.log_temp, None
    z_raw = pm.predictive.sample_calibrated_outputs(inputs, 12, rng=key)
    pm.predictive.log_temp = keep
    np.testing.assert_allclose(z.cpu().numpy(), z_raw.cpu().numpy() * np.exp(-lt.cpu().numpy()[0]), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(pm.predictive.mean(inputs, 12, rng=key).cpu().numpy(), P.cls_mean(z.cpu().numpy()), rtol=1e-5, atol=1e-8)
    # no calibrator: nothing to calibrate
    pm2 = ProbClassifier(model=MLP(), output_calibrator=None, posterior_approximator=SGHMCPosteriorApproximator(
        n_samples=2, n_thinning=2, burnin_length=2, step_schedule=1e-3))
    st2 = pm2.train(train_loader, calib_data_loader=test_loader, fit_config=FitConfig(optimizer=FitOptimizer(n_epochs=2)))
    assert st2["calib_status"] is None and pm2.predictive.log_temp is None


def test_fitted_temperature_travels_with_the_checkpoints(cuda, tmp_path):
    """calib_params = {"output_calibrator": {"params": {"log_temp": f32[1]}}} is part of every stored state
    (fortuna/prob_model/base.py:213-215): a model restored with load_state predicts with it."""
    from fortuna_b200.prob_model import (
        Adam, CalibConfig, CalibOptimizer, CalibProcessor, FitCheckpointer, FitConfig, FitOptimizer, ProbClassifier,
        SGHMCPosteriorApproximator,
    )
    from fortuna_b200.training import checkpoint as ck

    MLP, train, test, train_loader, test_loader = _setup()
    torch.manual_seed(0)
    d = str(tmp_path / "ck")
    approx = dict(n_samples=4, n_thinning=4, burnin_length=20, step_schedule=1e-3)
    pm = ProbClassifier(model=MLP(), posterior_approximator=SGHMCPosteriorApproximator(**approx), seed=0)
    pm.train(train_loader, calib_data_loader=test_loader,
             fit_config=FitConfig(optimizer=FitOptimizer(n_epochs=12), checkpointer=FitCheckpointer(save_checkpoint_dir=d)),
             calib_config=CalibConfig(optimizer=CalibOptimizer(Adam(1e-2), n_epochs=5), processor=CalibProcessor(6)))
    lt = pm.predictive.log_temp
    for sub in ("c", "0", "3"):
        stored = ck.restore_checkpoint(os.path.join(d, sub))["calib_params"]["output_calibrator"]["params"]["log_temp"]
        np.testing.assert_array_equal(stored, lt.cpu().numpy())
    pm2 = ProbClassifier(model=MLP(), posterior_approximator=SGHMCPosteriorApproximator(**approx), seed=1)
    pm2.load_state(d)
    assert torch.equal(pm2.predictive.log_temp, lt)
    key = pm.rng.get()
    inputs = test_loader.to_inputs_loader()
    assert torch.equal(pm2.predictive.mean(inputs, 8, rng=key), pm.predictive.mean(inputs, 8, rng=key))


def _small_fit(last_layer_only, seed=0, n_samples=8):
    from fortuna_b200.prob_model import FitConfig, FitOptimizer, ProbClassifier, SGHMCPosteriorApproximator

    MLP, train, test, train_loader, test_loader = _setup()
    torch.manual_seed(seed)
    pm = ProbClassifier(model=MLP(), posterior_approximator=SGHMCPosteriorApproximator(
        n_samples=n_samples, n_thinning=5, burnin_length=100, step_schedule=1e-3), seed=seed)
    freeze = (lambda path, v: "trainable" if "output_subnet" in path else "frozen") if last_layer_only else None
    pm.train(train_loader, fit_config=FitConfig(optimizer=FitOptimizer(n_epochs=40, freeze_fun=freeze)))
    return pm, train, test


@pytest.mark.parametrize("last_layer_only", [True, False])
def test_predictive_streams_the_loader_batch_by_batch(cuda, last_layer_only):
    """prob_model.predictive.* walks the loader like the reference's _loop_fun_through_inputs_loader (base.py:946-1013):
    the result does not depend on the batch size (ragged last batch included), equals the oracle on the drawn samples'
    logits, and two DIFFERENT inputs of the same shape give different answers (the round-1 cache keyed on a freed
    tensor's address could return the first one's logits)."""
    from fortuna_b200.data import DataLoader, InputsLoader
    from oracle import predictive as P

    pm, train, test = _small_fit(last_layer_only)
    key = pm.rng.get()
    names = ("mean", "entropy", "aleatoric_variance", "epistemic_variance", "log_prob", "ensemble_log_prob", "mode")
    whole = pm.predictive.statistics(DataLoader.from_array_data(test, batch_size=None), names, 12, rng=key)
    for bs in (128, 37, 400):
        part = pm.predictive.statistics(DataLoader.from_array_data(test, batch_size=bs), names, 12, rng=key)
        for k in names:  # the torch backbone picks its GEMM kernel by batch size: equal to float32 rounding, not bits
            if k == "mode":
                assert float((part[k] != whole[k]).float().mean()) < 0.01
            else:
                np.testing.assert_allclose(part[k].cpu().numpy(), whole[k].cpu().numpy(), rtol=2e-5, atol=2e-6, err_msg=f"{k} {bs}")
    z = pm.predictive.sample_calibrated_outputs(InputsLoader.from_array_inputs(test[0], batch_size=90), 12, rng=key).cpu().numpy()
    np.testing.assert_allclose(whole["mean"].cpu().numpy(), P.cls_mean(z), rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(whole["log_prob"].cpu().numpy(), P.cls_log_prob(z, test[1]), rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(whole["ensemble_log_prob"].cpu().numpy(), P.cls_ensemble_log_prob(z, test[1]), rtol=1e-5, atol=2e-6)
    other = (test[0][::-1] * np.float32(0.5)).copy()
    m1 = pm.predictive.mean(InputsLoader.from_array_inputs(test[0]), 12, rng=key)
    m2 = pm.predictive.mean(InputsLoader.from_array_inputs(other), 12, rng=key)
    m1b = pm.predictive.mean(InputsLoader.from_array_inputs(test[0]), 12, rng=key)
    assert torch.equal(m1, m1b) and not torch.equal(m1, m2)
    z2 = pm.predictive.sample_calibrated_outputs(InputsLoader.from_array_inputs(other), 12, rng=key).cpu().numpy()
    np.testing.assert_allclose(m2.cpu().numpy(), P.cls_mean(z2), rtol=1e-5, atol=1e-8)


def test_predictive_variance_and_std_follow_the_reference_key_schedule(cuda):
    """base.py:838-944: variance = aleatoric_variance(key1) + epistemic_variance(key2) with rng, key1 = split(rng) and
    rng, key2 = split(rng) - two different sets of posterior draws; std = sqrt(variance); given estimates are used as
    they are."""
    from fortuna_b200 import ops
    from fortuna_b200.data import InputsLoader
    from oracle import predictive as P
    from oracle import prng as oprng

    pm, train, test = _small_fit(True, seed=2)
    inputs = InputsLoader.from_array_inputs(test[0], batch_size=150)
    rng = pm.rng.get()
    k = ops.key_to_host(rng)
    r1, key1 = oprng.split(k)
    r2, key2 = oprng.split(r1)
    idx1 = np.array([oprng.choice(q, 8) for q in oprng.split(key1, 9)])
    idx2 = np.array([oprng.choice(q, 8) for q in oprng.split(key2, 9)])
    assert not np.array_equal(idx1, idx2)
    stored = pm.predictive._logits_fn(None)(torch.as_tensor(test[0], device=cuda)).cpu().numpy()  # [8, B, C]
    want = (P.cls_aleatoric_variance(stored[idx1]) + P.cls_epistemic_variance(stored[idx2])).astype(np.float32)
    var = pm.predictive.variance(inputs, 9, rng=rng)
    np.testing.assert_allclose(var.cpu().numpy(), want, rtol=1e-5, atol=4e-7)
    std = pm.predictive.std(inputs, 9, rng=rng)
    np.testing.assert_allclose(std.cpu().numpy(), np.sqrt(want), rtol=1e-5, atol=1e-6)
    # given estimates: no draws at all
    a = pm.predictive.aleatoric_variance(inputs, 9, rng=rng)
    e = pm.predictive.epistemic_variance(inputs, 9, rng=rng)
    assert torch.equal(pm.predictive.variance(inputs, 9, aleatoric_variances=a, epistemic_variances=e), a + e)
    assert torch.equal(pm.predictive.std(inputs, variances=a + e), torch.sqrt(a + e))
    # only the aleatoric term given: the epistemic one is drawn with the FIRST split of rng (base.py:889-897)
    v2 = pm.predictive.variance(inputs, 9, aleatoric_variances=a, rng=rng)
    want2 = a.cpu().numpy() + P.cls_epistemic_variance(stored[idx1])
    np.testing.assert_allclose(v2.cpu().numpy(), want2, rtol=1e-5, atol=4e-7)


def test_predictive_bf16_last_layer_and_conformal_sets_in_one_pass(cuda):
    """The bench's route through the public API: bf16 last-layer contraction + bf16 logits (north_star: rtol 2e-2 on the
    GEMM), statistics and APS conformal sets from ONE pass over the loader; sets equal fb200_aps_conformal_sets on the
    returned mean bit for bit."""
    from fortuna_b200 import ops
    from fortuna_b200.data import DataLoader
    from fortuna_b200.prob_model.predictive import pipeline
    from fortuna_b200.prob_model.predictive.classification import ClassificationPredictive

    pm, train, test = _small_fit(True, seed=1)
    key = pm.rng.get()
    loader = DataLoader.from_array_data(test, batch_size=128)
    ref = pm.predictive.statistics(loader, ("mean", "entropy"), 16, rng=key)
    fast = ClassificationPredictive(pm.posterior, last_layer_dtype=torch.bfloat16, logits_dtype=torch.bfloat16)
    t = torch.as_tensor(test[1], dtype=torch.int32, device=cuda)
    val = ops.sort_f32_(ops.aps_scores(ref["mean"], t).clone())
    calib = pipeline.ConformalCalibration(val, len(test[1]), 0.1)
    out = fast.statistics(loader, ("mean", "entropy"), 16, rng=key, conformal=calib)
    np.testing.assert_allclose(out["mean"].cpu().numpy(), ref["mean"].cpu().numpy(), rtol=2e-2, atol=2e-3)
    m, sz = ops.aps_conformal_sets(out["mean"], calib.val_sorted, calib.threshold)
    assert torch.equal(out["conformal_mask"], m) and torch.equal(out["conformal_sizes"], sz)


def test_predictive_statistics_to_host_with_streamed_ece(cuda):
    """statistics(..., to_host=True, ece=True): pinned host results equal the device results, and the ECE accumulated
    batch by batch (thresholds of the WHOLE data set) equals the metric on the full mean."""
    from fortuna_b200 import ops
    from fortuna_b200.data import DataLoader
    from fortuna_b200.prob_model.predictive import pipeline
    from oracle import scoring as Sc

    pm, train, test = _small_fit(True, seed=4)
    key = pm.rng.get()
    loader = DataLoader.from_array_data(test, batch_size=96)
    names = ("mean", "entropy", "log_prob", "mode")
    dev_out = pm.predictive.statistics(loader, names, 10, rng=key, ece=True)
    host_out = pm.predictive.statistics(loader, names, 10, rng=key, ece=True, to_host=True)
    for k in names + ("ece",):
        assert not host_out[k].is_cuda and (k == "ece" or host_out[k].is_pinned())
        assert torch.equal(host_out[k], dev_out[k].cpu()), k
    mean = dev_out["mean"].cpu().numpy()
    want = Sc.expected_calibration_error(mean.argmax(-1), mean, test[1])
    np.testing.assert_allclose(float(dev_out["ece"][0]), want, rtol=1e-5, atol=1e-7)
    again = pm.predictive.statistics(loader, names, 10, rng=key, ece=True, to_host=host_out)  # buffers reused
    assert again["mean"].data_ptr() == host_out["mean"].data_ptr()


def test_last_layer_samples_from_a_torch_fit(cuda):
    """SGMCMCPosterior.last_layer() on a bank filled by a fit here (torch nn.Linear: weight [C, D]) gives the flax-layout
    kernel [n, D, C] the contraction kernels take - the same logits as the predictive's own route."""
    from fortuna_b200 import ops

    pm, train, test = _small_fit(True, seed=5, n_samples=6)
    ll = pm.posterior.last_layer()
    w = pm.posterior.state.leaf_matrix(("model", "params", "output_subnet", "weight"))
    assert ll.kernel.shape == (6, w.shape[2], w.shape[1]) and torch.equal(ll.kernel, w.transpose(1, 2))
    x = torch.as_tensor(test[0], device=cuda)
    with torch.no_grad():
        feats = pm.posterior.bound.model.dfe_subnet(x).contiguous()
    z = ops.last_layer_logits(feats, ll.kernel, ll.bias, path=1)  # W_SDC layout
    ref = pm.predictive._logits_fn(None, calibrated=False)(x)
    np.testing.assert_allclose(z.cpu().numpy(), ref.cpu().numpy(), rtol=1e-5, atol=1e-5)
