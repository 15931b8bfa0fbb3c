"""Calibration models: fused loss heads, fused Adam, CalibClassifier / CalibRegressor end to end against
the numpy oracle, and user-supplied losses."""
import functools

import numpy as np
import pytest
import torch

from oracle import calib_model as O
from oracle import predictive as P

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,C", [(64, 10), (1000, 1000), (33, 2), (7, 37)])
@pytest.mark.parametrize("gamma", [0.0, 1.0, 2.0, 3.5])
def test_focal_loss_head(cuda, n, C, gamma):
    from fortuna_b200 import ops

    r = np.random.default_rng(n + C)
    z = (2 * r.standard_normal((n, C))).astype(np.float32)
    y = r.integers(0, C, n).astype(np.int32)
    loss, grad, rows = ops.focal_loss_grad(torch.from_numpy(z).to(cuda), torch.from_numpy(y).to(cuda), gamma)
    np.testing.assert_allclose(rows.cpu().numpy(), O.focal_loss_rows(z, y, gamma, np.float64), rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(loss.item(), float(O.focal_loss(z, y, gamma, np.float64)), rtol=1e-5)
    want = O.focal_loss_grad(z, y, gamma)
    np.testing.assert_allclose(grad.cpu().numpy(), want, rtol=2e-4, atol=2e-6 * np.abs(want).max())
    # the loss only: no gradient buffer
    loss2, none, _ = ops.focal_loss_grad(torch.from_numpy(z).to(cuda), torch.from_numpy(y).to(cuda), gamma, want_grad=False)
    assert none is None and loss2.item() == loss.item()


@pytest.mark.parametrize("n,d", [(50, 1), (257, 3)])
def test_scaled_mse_head(cuda, n, d):
    from fortuna_b200 import ops

    r = np.random.default_rng(n)
    o = r.standard_normal((n, 2 * d)).astype(np.float32)
    t = r.standard_normal((n, d)).astype(np.float32)
    loss, grad, terms = ops.scaled_mse_loss_grad(torch.from_numpy(o).to(cuda), torch.from_numpy(t).to(cuda))
    np.testing.assert_allclose(loss.item(), float(O.scaled_mse(o, t, np.float64)), rtol=1e-5)
    np.testing.assert_allclose(grad.cpu().numpy(), O.scaled_mse_grad(o, t), rtol=1e-4, atol=1e-7)


def test_adam_step_bit_exact(cuda):
    from fortuna_b200 import ops

    r = np.random.default_rng(0)
    p = r.standard_normal(10_001).astype(np.float32)
    opt = O.Adam(1e-2)
    st = opt.init(p)
    pt = torch.from_numpy(p.copy()).to(cuda)
    mu, nu = torch.zeros_like(pt), torch.zeros_like(pt)
    for step in range(1, 6):
        g = (r.standard_normal(p.size) * 10.0 ** r.integers(-6, 2)).astype(np.float32)
        p, st = opt.step(p, g, st)
        ops.adam_step_(pt, torch.from_numpy(g).to(cuda), mu, nu, step, 1e-2)
        assert np.array_equal(pt.cpu().numpy(), p) and np.array_equal(mu.cpu().numpy(), st["mu"])
        assert np.array_equal(nu.cpu().numpy(), st["nu"])


def _toy(n=600, D=5, C=3, seed=0):
    r = np.random.default_rng(seed)
    x = r.standard_normal((n, D)).astype(np.float32)
    y = ((x[:, 0] > 0).astype(np.int32) + (x[:, 1] > 0.3).astype(np.int32)).astype(np.int32)
    return x, y


def test_calib_classifier_matches_oracle_loop(cuda):
    """A single Dense layer calibrated for 4 epochs of 3 batches: per-epoch losses and final weights against the numpy loop
    (same focal loss, same Adam), then every predictive map against the oracle's S = 1 maps of the model outputs."""
    from fortuna_b200.calib_model import CalibClassifier, Config, Monitor, Optimizer
    from fortuna_b200.data.loader import DataLoader

    x, y = _toy()
    torch.manual_seed(0)
    model = torch.nn.Linear(5, 3)
    W0, b0 = model.weight.detach().numpy().copy(), model.bias.detach().numpy().copy()
    loader = DataLoader.from_array_data((x, y), batch_size=200)
    cm = CalibClassifier(model, seed=0)
    status = cm.calibrate(loader, val_data_loader=DataLoader.from_array_data((x[:100], y[:100]), batch_size=50),
                          config=Config(optimizer=Optimizer(n_epochs=4), monitor=Monitor(verbose=False)))
    losses, W, b = O.linear_calibration(W0, b0, [(x[i : i + 200], y[i : i + 200]) for i in range(0, 600, 200)], n_epochs=4)
    np.testing.assert_allclose(status["loss"], losses, rtol=2e-5)
    np.testing.assert_allclose(model.weight.detach().cpu().numpy(), W, rtol=2e-4, atol=2e-6)
    np.testing.assert_allclose(model.bias.detach().cpu().numpy(), b, rtol=2e-4, atol=2e-6)
    assert len(status["val_loss"]) == 4
    z = (x @ W.T + b).astype(np.float32)[None]
    il = loader.to_inputs_loader()
    np.testing.assert_allclose(cm.predictive.mean(il).cpu().numpy(), P.cls_mean(z), rtol=1e-4, atol=1e-7)
    assert np.array_equal(cm.predictive.mode(il).cpu().numpy(), P.cls_mode(z))
    np.testing.assert_allclose(cm.predictive.entropy(il).cpu().numpy(), P.cls_entropy(z), rtol=1e-4, atol=4e-6)
    np.testing.assert_allclose(cm.predictive.log_prob(loader).cpu().numpy(), P.cls_log_prob(z, y), rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(cm.predictive.variance(il).cpu().numpy(), P.cls_aleatoric_variance(z), rtol=1e-4, atol=2e-7)
    np.testing.assert_allclose(cm.predictive.std(il).cpu().numpy() ** 2, P.cls_aleatoric_variance(z), rtol=2e-4, atol=4e-7)
    s = cm.predictive.sample(il, n_samples=4, rng=cm.rng.get())
    assert s.shape == (4, 600) and int(s.min()) >= 0 and int(s.max()) < 3


def test_calib_classifier_state_round_trip_and_errors(cuda, tmp_path):
    from fortuna_b200.calib_model import CalibClassifier, Checkpointer, Config, Monitor, Optimizer
    from fortuna_b200.data.loader import DataLoader

    x, y = _toy(120)
    loader = DataLoader.from_array_data((x, y), batch_size=60)
    torch.manual_seed(1)
    cm = CalibClassifier(torch.nn.Sequential(torch.nn.Linear(5, 8), torch.nn.Tanh(), torch.nn.Linear(8, 3)))
    with pytest.raises(ValueError):
        cm.predictive.mean(loader.to_inputs_loader())  # no state yet
    with pytest.raises(ValueError):
        cm.calibrate(loader, config=Config(checkpointer=Checkpointer(dump_state=True)))
    cm.calibrate(loader, config=Config(optimizer=Optimizer(n_epochs=2), monitor=Monitor(verbose=False)))
    want = cm.predictive.mean(loader.to_inputs_loader()).clone()
    cm.save_state(str(tmp_path / "ck"))
    torch.manual_seed(2)
    other = CalibClassifier(torch.nn.Sequential(torch.nn.Linear(5, 8), torch.nn.Tanh(), torch.nn.Linear(8, 3)))
    other.load_state(str(tmp_path / "ck"))
    assert torch.equal(other.predictive.mean(loader.to_inputs_loader()), want)
    bad = CalibClassifier(torch.nn.Linear(5, 7))
    with pytest.raises(ValueError):
        bad.calibrate(loader)  # 7 outputs for 3 classes


def test_user_supplied_losses(cuda):
    """A torch cross entropy handed in as ``loss_fn`` runs through autograd and follows the fused focal head at gamma = 0
    (the same loss) - in CalibClassifier and, through the temperature's chain rule, in OutputCalibClassifier."""
    from fortuna_b200.calib_model import CalibClassifier, Config, Monitor, Optimizer
    from fortuna_b200.data.loader import DataLoader
    from fortuna_b200.loss.classification.focal_loss import focal_loss_fn
    from fortuna_b200 import output_calib_model as oc
    from fortuna_b200.output_calib_model import OutputCalibClassifier

    def xent(outputs, targets):
        return torch.nn.functional.cross_entropy(outputs, targets.long())

    x, y = _toy(300)
    loader = DataLoader.from_array_data((x, y), batch_size=100)
    res = []
    for loss_fn in (xent, functools.partial(focal_loss_fn, gamma=0.0)):
        torch.manual_seed(3)
        cm = CalibClassifier(torch.nn.Linear(5, 3))
        st = cm.calibrate(loader, loss_fn=loss_fn, config=Config(optimizer=Optimizer(n_epochs=3), monitor=Monitor(verbose=False)))
        res.append(st["loss"])
    np.testing.assert_allclose(res[0], res[1], rtol=1e-5)
    r = np.random.default_rng(5)
    z = (3 * r.standard_normal((400, 3))).astype(np.float32)
    t = r.integers(0, 3, 400).astype(np.int32)
    res = []
    for loss_fn in (xent, functools.partial(focal_loss_fn, gamma=0.0)):
        m = OutputCalibClassifier()
        st = m.calibrate(z, t, loss_fn=loss_fn, config=oc.Config(optimizer=oc.Optimizer(n_epochs=5), monitor=oc.Monitor(verbose=False)))
        res.append((st["loss"], m.predictive.state.get().log_temp))
    np.testing.assert_allclose(res[0][0], res[1][0], rtol=1e-5)
    np.testing.assert_allclose(res[0][1], res[1][1], rtol=1e-4, atol=1e-6)


def test_calib_regressor(cuda):
    from fortuna_b200.calib_model import CalibRegressor, Config, Monitor, Optimizer
    from fortuna_b200.data.loader import DataLoader

    r = np.random.default_rng(7)
    x = r.standard_normal((300, 4)).astype(np.float32)
    y = (x[:, :1] * 2 + 0.1 * r.standard_normal((300, 1))).astype(np.float32)
    loader = DataLoader.from_array_data((x, y), batch_size=100)
    torch.manual_seed(4)
    cr = CalibRegressor(torch.nn.Linear(4, 1), torch.nn.Linear(4, 1))
    st = cr.calibrate(loader, config=Config(optimizer=Optimizer(n_epochs=30), monitor=Monitor(verbose=False)))
    assert st["loss"][-1] < st["loss"][0]
    il = loader.to_inputs_loader()
    with torch.no_grad():
        out = cr.model(torch.from_numpy(x).to(cuda)).cpu().numpy()
    np.testing.assert_allclose(cr.predictive.mean(il).cpu().numpy(), out[:, :1], rtol=1e-6)
    np.testing.assert_allclose(cr.predictive.variance(il).cpu().numpy(), np.exp(out[:, 1:]), rtol=1e-5)
    ci = cr.predictive.credible_interval(il, n_samples=50, rng=cr.rng.get())
    assert ci.shape == (300, 2) and bool((ci[:, 0] <= ci[:, 1]).all())
    assert cr.predictive.entropy(il, n_samples=20, rng=cr.rng.get()).shape == (300,)
