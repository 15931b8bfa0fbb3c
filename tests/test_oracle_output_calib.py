"""CPU checks of the output calibration models: the oracle's analytic derivatives against central
differences, a hand-computed focal-loss value, and the host logic that needs no GPU - the state's checkpoint round trip
in the reference's flax-msgpack layout, the repository, early stopping, loss dispatch and the argument errors of
fortuna/output_calib_model/{base,classification,regression}.py."""
import functools
import os

import numpy as np
import pytest

from oracle import output_calib as oc


def test_focal_loss_hand_value():
    # one row, two classes, logits (0, log 3): p = (1/4, 3/4); target 0 -> -(3/4)^2 log(1/4)
    l, _ = oc.focal_loss(np.array([[0.0, np.log(3.0)]]), [0], 0.0, 2.0)
    assert abs(l - (0.75**2) * np.log(4.0)) < 1e-12
    # gamma = 0 is the cross entropy; log_temp = log 2 halves the logits: p = (1, sqrt 3) / (1 + sqrt 3)
    l, _ = oc.focal_loss(np.array([[0.0, np.log(3.0)]]), [1], np.log(2.0), 0.0)
    assert abs(l + np.log(np.sqrt(3.0) / (1.0 + np.sqrt(3.0)))) < 1e-12


@pytest.mark.parametrize("gamma", [0.0, 1.0, 2.0, 3.5])
@pytest.mark.parametrize("lt", [-0.4, 0.0, 0.7])
def test_focal_derivative_matches_central_difference(gamma, lt):
    r = np.random.default_rng(3)
    z, y = 3.0 * r.standard_normal((64, 9)), r.integers(0, 9, 64)
    _, d = oc.focal_loss(z, y, lt, gamma)
    h = 1e-6
    fd = (oc.focal_loss(z, y, lt + h, gamma)[0] - oc.focal_loss(z, y, lt - h, gamma)[0]) / (2 * h)
    assert abs(fd - d) < 1e-6 * max(1.0, abs(d))


@pytest.mark.parametrize("lt", [-0.4, 0.0, 0.7])
def test_scaled_mse_derivative_matches_central_difference(lt):
    r = np.random.default_rng(4)
    z, y = r.standard_normal((40, 6)), r.standard_normal((40, 3))
    l, d = oc.scaled_mse(z, y, lt)
    h = 1e-6
    fd = (oc.scaled_mse(z, y, lt + h)[0] - oc.scaled_mse(z, y, lt - h)[0]) / (2 * h)
    assert abs(fd - d) < 1e-6 * max(1.0, abs(d))
    # scaled MSE = -2 * mean Gaussian log-likelihood - d log 2 pi
    from oracle.calibration import calib_terms_reg

    L, _ = calib_terms_reg(z[None], y, lt)
    assert abs(l - (-2.0 * L[0] / 40 - 3 * np.log(2 * np.pi))) < 1e-10


def test_output_samples_shapes_and_determinism():
    r = np.random.default_rng(5)
    z = r.standard_normal((6, 5)).astype(np.float32)
    a, b = oc.cls_output_sample(z, [1, 2], 4), oc.cls_output_sample(z, [1, 2], 4)
    assert a.shape == (4, 6) and a.dtype == np.int32 and np.array_equal(a, b) and a.min() >= 0 and a.max() < 5
    # a near-one-hot row always draws its class
    zz = np.full((1, 5), -50.0, np.float32)
    zz[0, 3] = 50.0
    assert (oc.cls_output_sample(zz, [7, 7], 9) == 3).all()
    y = oc.reg_output_sample(r.standard_normal((3, 4)).astype(np.float32), [1, 2], 7)
    assert y.shape == (7, 3, 2) and y.dtype == np.float32


# ---------------------------------------------------------------------------------------------- host logic
def _state(lt=0.3, step=4):
    from fortuna_b200.output_calib_model.config.base import Optimizer
    from fortuna_b200.output_calib_model.state import OutputCalibState

    opt = Optimizer().method
    s = OutputCalibState.init({"output_calibrator": {"params": {"log_temp": np.array([lt], np.float32)}}}, optimizer=opt,
                              step=step)
    return s.apply_gradients(np.array([0.5], np.float32)), opt


def test_state_checkpoint_round_trip(tmp_path):
    from fortuna_b200.output_calib_model.output_calib_state_repository import (OutputCalibStateRepository, restore_state,
                                                                               save_state)
    from fortuna_b200.training import checkpoint

    s, opt = _state()
    assert s.step == 5 and s.log_temp.dtype == np.float32 and abs(float(s.log_temp[0]) - 0.29) < 1e-6  # first Adam step = -lr
    save_state(s, tmp_path, keep=2)
    assert os.path.exists(tmp_path / "checkpoint_5")
    d = checkpoint.restore_checkpoint(str(tmp_path))
    # the reference's state dict (flax to_state_dict of OutputCalibState)
    assert list(d) == ["step", "params", "opt_state", "encoded_name", "frozen_params", "dynamic_scale", "mutable"]
    assert checkpoint.decode_name(d) == "OutputCalibState"
    assert d["params"]["output_calibrator"]["params"]["log_temp"].shape == (1,)
    assert set(d["opt_state"]["0"]) == {"count", "mu", "nu"} and d["opt_state"]["1"] == {}
    r = restore_state(tmp_path)  # no optimizer: the saved optimizer state is kept
    assert r.step == 5 and np.array_equal(r.log_temp, s.log_temp) and r.opt_state["count"] == 1
    np.testing.assert_array_equal(r.opt_state["mu"], s.opt_state["mu"])
    r2 = restore_state(tmp_path, optimizer=opt)  # a given optimizer restarts (state.py:49-57)
    assert r2.opt_state["count"] == 0 and r2.step == 5
    with pytest.raises(ValueError):
        restore_state(tmp_path / "missing")
    os.makedirs(tmp_path / "empty")
    with pytest.raises(ValueError):
        restore_state(tmp_path / "empty")
    repo = OutputCalibStateRepository()
    with pytest.raises(ValueError):
        repo.get()
    repo.put(s)
    assert repo.get() is s
    disk = OutputCalibStateRepository(str(tmp_path / "dump"))
    disk.put(s, keep=1)
    assert np.array_equal(disk.get().log_temp, s.log_temp)
    # saving an older step without force is refused like flax's save_checkpoint
    old, _ = _state(step=1)
    with pytest.raises(checkpoint.InvalidCheckpointError):
        save_state(old, tmp_path)
    save_state(old, tmp_path, force_save=True)
    assert restore_state(tmp_path).step == 2


def test_early_stopping_follows_flax():
    from fortuna_b200.output_calib_model.output_calib_model_calibrator import EarlyStopping

    es = EarlyStopping(min_delta=0.1, patience=2)
    seq = [(1.0, True, False), (0.95, False, False), (0.8, True, False), (0.79, False, False), (0.78, False, False),
           (0.77, False, True)]
    for metric, improved, stop in seq:
        es.update(metric)
        assert (es.has_improved, es.should_stop) == (improved, stop), metric


def test_loss_dispatch_and_argument_errors():
    from fortuna_b200.loss.classification.focal_loss import focal_loss_fn
    from fortuna_b200.loss.regression.scaled_mse import scaled_mse_fn
    from fortuna_b200.output_calib_model import Config, Monitor, OutputCalibClassifier, OutputCalibRegressor
    from fortuna_b200.output_calib_model.output_calib_model_calibrator import resolve_loss

    assert resolve_loss(focal_loss_fn, False) == ("focal", 2.0)
    assert resolve_loss(functools.partial(focal_loss_fn, gamma=0.0), False) == ("focal", 0.0)
    assert resolve_loss(scaled_mse_fn, True) == ("scaled_mse", None)
    for fn, reg in ((scaled_mse_fn, False), (focal_loss_fn, True), (lambda o, t: 0.0, False)):
        with pytest.raises(NotImplementedError):
            resolve_loss(fn, reg)
    with pytest.raises(ValueError):
        Monitor(metrics=[len])
    with pytest.raises(ValueError):
        Monitor(metrics=(3,))
    with pytest.raises(ValueError):
        Monitor(uncertainty_fn=3)
    clf = OutputCalibClassifier()
    z, y = np.zeros((8, 3), np.float32), np.arange(8) % 3
    with pytest.raises(ValueError):  # outputs.shape[1] != number of classes in the targets
        clf.calibrate(np.zeros((8, 4), np.float32), y)
    with pytest.raises(ValueError):  # validation needs both
        clf.calibrate(z, y, val_outputs=z)
    with pytest.raises(ValueError):  # nothing calibrated yet
        clf.predictive.mean(z)
    with pytest.raises(ValueError):
        clf.save_state("/tmp/nowhere")
    with pytest.raises(ValueError):
        clf.load_state("/tmp/definitely/not/a/checkpoint")
    reg = OutputCalibRegressor()
    with pytest.raises(ValueError):  # outputs must be [mu | log var]
        reg.calibrate(np.zeros((8, 3), np.float32), np.zeros((8, 2), np.float32))
    with pytest.raises(NotImplementedError):
        OutputCalibClassifier(output_calibrator=object())
    assert isinstance(Config().optimizer.n_epochs, int) and Config().checkpointer.keep_top_n_checkpoints == 2
