"""Output calibration models on the GPU: the fused loss-and-derivative kernels, the one-step-per-epoch
calibration loop, the S = 1 predictive maps and the output-layer target samples against oracle/output_calib.py; the
user flow mirrors the reference's own dry-run test (tests/fortuna/calib_model/test_output_calib_model.py:130-260:
calibrate, restore and continue, dump, load_state, save_state)."""
import functools
import os

import numpy as np
import pytest
import torch

from oracle import output_calib as oc
from oracle import predictive as P
from oracle import prng as oprng

pytestmark = pytest.mark.gpu


def _cls_data(n, C, seed, scale=3.0):
    r = np.random.default_rng(seed)
    y = np.arange(n) % C
    r.shuffle(y)
    z = (scale * r.standard_normal((n, C))).astype(np.float32)
    z[np.arange(n), y] += 2.0  # informative logits, so the calibration has something to find
    return z, y.astype(np.int32)


def _reg_data(n, d, seed):
    r = np.random.default_rng(seed)
    mu = r.standard_normal((n, d))
    lv = 0.5 * r.standard_normal((n, d)) - 1.0
    y = mu + 1.7 * np.exp(0.5 * lv) * r.standard_normal((n, d))  # under-dispersed outputs: log_temp should rise
    return np.concatenate([mu, lv], -1).astype(np.float32), y.astype(np.float32)


@pytest.mark.parametrize("n,C", [(1, 2), (257, 10), (4096, 1000), (33, 1500)])
@pytest.mark.parametrize("gamma", [0.0, 1.0, 2.0, 3.5])
def test_focal_terms_match_oracle(cuda, n, C, gamma):
    from fortuna_b200 import ops

    z, y = _cls_data(n, C, n + C)
    for lt in (None, 0.45, -0.3):
        ltt = None if lt is None else torch.tensor([lt], dtype=torch.float32, device=cuda)
        t = ops.output_calib_focal_loss_terms(torch.from_numpy(z).to(cuda), torch.from_numpy(y).to(cuda), ltt, gamma).cpu().numpy()
        l0 = 0.0 if lt is None else lt
        want_l, want_d = oc.focal_loss(z, y, l0, gamma)
        tau = np.exp(-l0)
        np.testing.assert_allclose(t[0] / n, want_l, rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose(-tau * t[1] / n, want_d, rtol=2e-4, atol=2e-6)


def test_focal_terms_bf16_and_loss_fn(cuda):
    from fortuna_b200 import ops
    from fortuna_b200.loss.classification.focal_loss import focal_loss_fn

    z, y = _cls_data(500, 37, 1)
    zb = torch.from_numpy(z).to(cuda).to(torch.bfloat16)
    t = ops.output_calib_focal_loss_terms(zb, torch.from_numpy(y).to(cuda), None, 2.0).cpu().numpy()
    want_l, _ = oc.focal_loss(zb.float().cpu().numpy(), y, 0.0, 2.0)
    np.testing.assert_allclose(t[0] / 500, want_l, rtol=2e-5)
    got = focal_loss_fn(z, y)  # the public loss function on host arrays
    assert got.dtype == torch.float32 and got.is_cuda
    np.testing.assert_allclose(got.item(), oc.focal_loss(z, y)[0], rtol=2e-5)
    np.testing.assert_allclose(focal_loss_fn(z, y, gamma=0.0).item(), oc.focal_loss(z, y, 0.0, 0.0)[0], rtol=2e-5)


@pytest.mark.parametrize("n,d", [(1, 1), (300, 2), (5000, 8)])
def test_scaled_mse_terms_match_oracle(cuda, n, d):
    from fortuna_b200 import ops
    from fortuna_b200.loss.regression.scaled_mse import scaled_mse_fn

    z, y = _reg_data(n, d, n + d)
    for lt in (None, 0.6, -0.2):
        ltt = None if lt is None else torch.tensor([lt], dtype=torch.float32, device=cuda)
        t = ops.output_calib_scaled_mse_terms(torch.from_numpy(z).to(cuda), torch.from_numpy(y).to(cuda), ltt).cpu().numpy()
        want_l, want_d = oc.scaled_mse(z, y, 0.0 if lt is None else lt)
        np.testing.assert_allclose(t[0] / n, want_l, rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose(t[1] / n, want_d, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(scaled_mse_fn(z, y).item(), oc.scaled_mse(z, y)[0], rtol=2e-5)


def test_classifier_calibration_follows_oracle_trajectory(cuda, tmp_path):
    from fortuna_b200.metric.classification import accuracy, brier_score
    from fortuna_b200.output_calib_model import Checkpointer, Config, Monitor, Optimizer, OutputCalibClassifier

    def brier(dummy, p, y):
        return brier_score(p, y)

    def acc(preds, p, y):
        return accuracy(preds, y)

    z, y = _cls_data(600, 10, 0)
    zv, yv = _cls_data(300, 10, 1)
    model = OutputCalibClassifier()
    status = model.calibrate(z, y, val_outputs=zv, val_targets=yv,
                             config=Config(optimizer=Optimizer(n_epochs=40), monitor=Monitor(metrics=(brier, acc), verbose=False),
                                           checkpointer=Checkpointer(save_checkpoint_dir=str(tmp_path))))
    want_loss, want_lt = oc.calibrate(z, y, 40)
    np.testing.assert_allclose(status["loss"], want_loss, rtol=3e-5)
    lt = model.predictive.state.get().log_temp
    assert lt.dtype == np.float32 and abs(float(lt[0]) - float(want_lt[-1])) < 2e-5
    assert set(status) == {"loss", "brier", "acc", "val_loss", "val_brier", "val_acc"} and all(len(v) == 40 for v in status.values())
    # metrics of epoch 0 are taken at log_temp = 0 (the step's own outputs); accuracy does not depend on the temperature
    p0 = P.softmax(z)
    np.testing.assert_allclose(status["brier"][0], np.mean(np.sum((p0 - np.eye(10)[y]) ** 2, -1)), rtol=1e-5)
    assert np.all(status["acc"] == status["acc"][0])
    np.testing.assert_allclose(status["val_loss"][-1], oc.focal_loss(zv, yv, float(lt[0]))[0], rtol=3e-5)
    assert os.path.exists(tmp_path / "checkpoint_40")  # on_train_end
    # predictive statistics with and without the temperature
    zc = (z * np.exp(-lt[0], dtype=np.float32)).astype(np.float32)
    for calibrated, zz in ((True, zc), (False, z)):
        p = P.softmax(zz)
        np.testing.assert_allclose(model.predictive.mean(z, calibrated=calibrated).cpu().numpy(), p, rtol=2e-5, atol=1e-7)
        assert np.array_equal(model.predictive.mode(z, calibrated=calibrated).cpu().numpy(), p.argmax(-1))
        np.testing.assert_allclose(model.predictive.variance(z, calibrated=calibrated).cpu().numpy(), p * (1 - p), rtol=2e-5, atol=1e-7)
        np.testing.assert_allclose(model.predictive.std(z, calibrated=calibrated).cpu().numpy(), np.sqrt(p * (1 - p)), rtol=2e-5, atol=2e-4)
        np.testing.assert_allclose(model.predictive.entropy(z, calibrated=calibrated).cpu().numpy(),
                                   -(p * np.log(np.maximum(p, 1e-38))).sum(-1), rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(model.predictive.log_prob(z, y, calibrated=calibrated).cpu().numpy(),
                                   P.log_softmax_rows(zz)[np.arange(600), y], rtol=2e-5, atol=2e-6)
    # restore and continue (the optimizer restarts, the step counter goes on), then dump
    status2 = model.calibrate(z, y, config=Config(optimizer=Optimizer(n_epochs=3), monitor=Monitor(verbose=False),
                                                  checkpointer=Checkpointer(restore_checkpoint_path=str(tmp_path))))
    assert model.predictive.state.get().step == 43 and len(status2["loss"]) == 3
    np.testing.assert_allclose(status2["loss"][0], oc.focal_loss(z, y, float(lt[0]))[0], rtol=3e-5)
    dump = tmp_path / "dump"
    model.calibrate(z, y, config=Config(optimizer=Optimizer(n_epochs=2), monitor=Monitor(verbose=False),
                                        checkpointer=Checkpointer(save_checkpoint_dir=str(dump), dump_state=True)))
    assert model.predictive.state.checkpoint_dir == str(dump)
    lt2 = model.predictive.state.get().log_temp
    other = OutputCalibClassifier()
    other.load_state(str(dump))
    assert np.array_equal(other.predictive.state.get().log_temp, lt2)
    other.save_state(str(tmp_path / "saved"))
    assert os.path.exists(tmp_path / "saved" / "checkpoint_2")
    np.testing.assert_array_equal(other.predictive.mean(z).cpu().numpy(), model.predictive.mean(z).cpu().numpy())


def test_classifier_gamma_partial_early_stopping_and_identity(cuda):
    from fortuna_b200.loss.classification.focal_loss import focal_loss_fn
    from fortuna_b200.output_calib_model import Config, Monitor, Optimizer, OutputCalibClassifier

    z, y = _cls_data(400, 5, 2)
    m = OutputCalibClassifier()
    st = m.calibrate(z, y, loss_fn=functools.partial(focal_loss_fn, gamma=0.0),
                     config=Config(optimizer=Optimizer(n_epochs=15), monitor=Monitor(verbose=False)))
    want_loss, want_lt = oc.calibrate(z, y, 15, gamma=0.0)
    np.testing.assert_allclose(st["loss"], want_loss, rtol=3e-5)
    assert abs(float(m.predictive.state.get().log_temp[0]) - float(want_lt[-1])) < 2e-5
    # early stopping on a validation loss that cannot improve by min_delta
    st = m.calibrate(z, y, val_outputs=z, val_targets=y,
                     config=Config(optimizer=Optimizer(n_epochs=50),
                                   monitor=Monitor(verbose=False, early_stopping_patience=2, early_stopping_min_delta=10.0)))
    assert len(st["val_loss"]) == 4  # first epoch improves (from inf), then patience 2 + the stopping epoch
    # no calibrator: outputs pass through, the loss is constant
    ident = OutputCalibClassifier(output_calibrator=None)
    st = ident.calibrate(z, y, config=Config(optimizer=Optimizer(n_epochs=3), monitor=Monitor(verbose=False)))
    assert np.all(st["loss"] == st["loss"][0])
    np.testing.assert_allclose(ident.predictive.mean(z).cpu().numpy(), P.softmax(z), rtol=2e-5, atol=1e-7)


@pytest.mark.parametrize("n,C,T", [(1, 2, 1), (77, 10, 30), (300, 1000, 7), (40, 33, 64)])
def test_classifier_output_samples_bit_exact(cuda, n, C, T):
    from fortuna_b200 import ops

    z, _ = _cls_data(n, C, n + T, scale=2.0)
    for lt in (None, np.array([0.3], np.float32)):
        want = oc.cls_output_sample(z, oprng.PRNGKey(5), T, lt)
        got = ops.cls_output_sample_targets(torch.from_numpy(z).to(cuda), ops.PRNGKey(5, cuda), T,
                                            None if lt is None else torch.from_numpy(lt).to(cuda)).cpu().numpy()
        assert got.shape == (T, n) and got.dtype == np.int32
        mism = int((got != want).sum())
        # the draw is a count of prefix sums below r: identical unless a float32 softmax differs in the last ulp exactly
        # at the threshold (oracle exp vs device expf); allow a vanishing fraction, require exactness on small C
        assert mism <= (0 if C <= 33 else max(1, got.size // 2000)), mism


def test_regressor_calibration_and_predictive(cuda, tmp_path):
    from fortuna_b200 import ops
    from fortuna_b200.output_calib_model import Checkpointer, Config, Monitor, Optimizer, OutputCalibRegressor

    def standard_error(m, v, y):
        return torch.sum((y - m) ** 2 / v) / m.shape[0]

    z, y = _reg_data(500, 1, 7)
    model = OutputCalibRegressor()
    status = model.calibrate(z, y, val_outputs=z, val_targets=y,
                             config=Config(optimizer=Optimizer(n_epochs=60), monitor=Monitor(metrics=(standard_error,), verbose=False),
                                           checkpointer=Checkpointer(save_checkpoint_dir=str(tmp_path), save_every_n_steps=20)))
    want_loss, want_lt = oc.calibrate(z, y, 60, regression=True)
    np.testing.assert_allclose(status["loss"], want_loss, rtol=3e-5, atol=1e-6)
    lt = model.predictive.state.get().log_temp
    assert abs(float(lt[0]) - float(want_lt[-1])) < 2e-5 and lt[0] > 0.3  # the outputs were under-dispersed
    assert status["standard_error"][0] > status["val_standard_error"][-1]
    mu, lv = z[:, :1], z[:, 1:] + lt[0]
    np.testing.assert_allclose(model.predictive.mean(z).cpu().numpy(), mu, rtol=1e-6)
    np.testing.assert_allclose(model.predictive.mode(z).cpu().numpy(), mu, rtol=1e-6)
    np.testing.assert_allclose(model.predictive.variance(z).cpu().numpy(), np.exp(lv), rtol=2e-5)
    np.testing.assert_allclose(model.predictive.variance(z, calibrated=False).cpu().numpy(), np.exp(z[:, 1:]), rtol=2e-5)
    np.testing.assert_allclose(model.predictive.std(z).cpu().numpy(), np.exp(0.5 * lv), rtol=2e-5)
    want_lp = -0.5 * (np.exp(-lv) * (y - mu) ** 2 + lv + np.log(2 * np.pi)).sum(-1)
    np.testing.assert_allclose(model.predictive.log_prob(z, y).cpu().numpy(), want_lp, rtol=2e-5, atol=2e-6)
    key = ops.PRNGKey(11, cuda)
    hk = ops.key_to_host(key)
    smp = model.predictive.sample(30, z, rng=key)
    want = oc.reg_output_sample(z, hk, 30, lt)
    assert smp.shape == (30, 500, 1)
    np.testing.assert_allclose(smp.cpu().numpy(), want, rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(model.predictive.entropy(z, n_target_samples=30, rng=key).cpu().numpy(),
                               oc.reg_output_entropy(z, hk, 30, lt), rtol=2e-5, atol=2e-5)
    q = model.predictive.quantile([0.1, 0.9], z, n_samples=30, rng=key)
    np.testing.assert_allclose(q.cpu().numpy(), P.quantile_axis0(want, [0.1, 0.9]), rtol=1e-5, atol=2e-6)
    assert model.predictive.quantile(0.5, z, rng=key).shape == (500, 1)
    ci = model.predictive.credible_interval(z, n_samples=30, error=0.05, rng=key)
    assert ci.shape == (500, 2) and bool((ci[:, 0] < ci[:, 1]).all())
    np.testing.assert_allclose(ci.cpu().numpy()[:, 1], P.quantile_axis0(want, [0.975])[0, :, 0], rtol=1e-5, atol=2e-6)
    assert model.predictive.credible_interval(z, interval_type="left-tailed", rng=key).shape == (500, 1)
    with pytest.raises(ValueError):
        model.predictive.credible_interval(z, interval_type="both")
    # multi-dimensional targets: samples and the refusal of credible intervals
    z3, y3 = _reg_data(64, 3, 8)
    m3 = OutputCalibRegressor()
    m3.calibrate(z3, y3, config=Config(optimizer=Optimizer(n_epochs=2), monitor=Monitor(verbose=False)))
    lt3 = m3.predictive.state.get().log_temp
    np.testing.assert_allclose(m3.predictive.sample(5, z3, rng=key).cpu().numpy(), oc.reg_output_sample(z3, hk, 5, lt3),
                               rtol=1e-5, atol=2e-6)
    with pytest.raises(ValueError):
        m3.predictive.credible_interval(z3, rng=key)
    # checkpoints written every 20 steps plus the final one, two kept (keep_top_n_checkpoints)
    assert sorted(os.listdir(tmp_path)) == ["checkpoint_41", "checkpoint_60"]
    other = OutputCalibRegressor()
    other.load_state(str(tmp_path))
    assert np.array_equal(other.predictive.state.get().log_temp, lt)


def test_classifier_custom_uncertainty_fn_sees_calibrated_logits(cuda):
    """``Monitor(uncertainty_fn=...)`` receives the CALIBRATED outputs (output_calib_model_calibrator.py:313-316 after
    loss.py:52-58); a softmax there reproduces the default uncertainty estimates."""
    from fortuna_b200.metric.classification import brier_score
    from fortuna_b200.output_calib_model import Config, Monitor, Optimizer, OutputCalibClassifier

    def brier(dummy, p, y):
        return brier_score(p, y)

    z, y = _cls_data(300, 6, 9)
    runs = []
    for fn in (None, lambda zc: torch.softmax(zc, dim=-1)):
        m = OutputCalibClassifier()
        st = m.calibrate(z, y, val_outputs=z, val_targets=y,
                         config=Config(optimizer=Optimizer(n_epochs=12), monitor=Monitor(metrics=(brier,), uncertainty_fn=fn, verbose=False)))
        runs.append(st)
    np.testing.assert_allclose(runs[0]["brier"], runs[1]["brier"], rtol=2e-5)
    np.testing.assert_allclose(runs[0]["val_brier"], runs[1]["val_brier"], rtol=2e-5)
    assert runs[0]["val_brier"][-1] != runs[0]["val_brier"][0]  # the temperature moved, so the calibrated outputs changed
