"""Calibration loop of the output calibration models - ``OutputCalibModelCalibrator.train``
(fortuna/output_calib_model/output_calib_model_calibrator.py:85-506).  As in the reference an epoch is ONE optimizer
step over the whole calibration set.  The reference differentiates ``Loss.__call__`` (loss.py:26-79) with autodiff; here
one kernel pass over the outputs returns the loss sum and its derivative in the temperature, the scalar chain rule and
the Adam step run on the host in float32."""
import collections
import functools
import logging
import math
from typing import Callable, Dict, Optional, Tuple

import numpy as np
import torch

from .. import ops
from ..loss.classification.focal_loss import focal_loss_fn
from ..loss.regression.scaled_mse import scaled_mse_fn
from .output_calib_state_repository import save_state
from .state import OutputCalibState


class EarlyStopping:
    """flax.training.early_stopping.EarlyStopping.update as used through WithEarlyStoppingMixin
    (fortuna/training/mixin.py:93-151)."""

    def __init__(self, min_delta: float = 0.0, patience: int = 0):
        self.min_delta, self.patience = min_delta, patience
        self.best_metric, self.patience_count = float("inf"), 0
        self.should_stop, self.has_improved = False, False

    def update(self, metric: float) -> "EarlyStopping":
        if math.isinf(self.best_metric) or self.best_metric - metric > self.min_delta:
            self.best_metric, self.patience_count, self.has_improved = metric, 0, True
        else:
            self.should_stop = self.patience_count >= self.patience or self.should_stop
            self.patience_count += 1
            self.has_improved = False
        return self


def resolve_loss(loss_fn: Callable, regression: bool):
    """The loss functions the fused loss-and-derivative kernels implement: ``focal_loss_fn`` (optionally a
    functools.partial fixing ``gamma``; gamma = 0 is the cross entropy) for classification, ``scaled_mse_fn`` for
    regression.  Anything else has no kernel and is refused rather than run on the host."""
    fn, kw = loss_fn, {}
    if isinstance(loss_fn, functools.partial):
        fn, kw = loss_fn.func, dict(loss_fn.keywords)
        if loss_fn.args:
            raise NotImplementedError("positional arguments bound to the loss function are not supported")
    if not regression and fn is focal_loss_fn:
        gamma = float(kw.pop("gamma", 2.0))
        if kw:
            raise NotImplementedError(f"unsupported focal_loss_fn arguments {sorted(kw)}")
        return ("focal", gamma)
    if regression and fn is scaled_mse_fn and not kw:
        return ("scaled_mse", None)
    raise NotImplementedError(
        "loss_fn must be fortuna_b200.loss.classification.focal_loss.focal_loss_fn (classification) or "
        "fortuna_b200.loss.regression.scaled_mse.scaled_mse_fn (regression): these are the losses with a fused "
        "loss-and-derivative kernel.")


def resolve_loss_or_user(loss_fn: Callable, regression: bool):
    """As ``resolve_loss``; any other callable ``loss_fn(calibrated_outputs, targets) -> scalar`` over torch tensors is
    accepted as ("user", loss_fn): user code is differentiated by autograd in the calibrated outputs, and the chain rule
    through the temperature is one ordered dot product on the device (``loss_and_grad``)."""
    try:
        return resolve_loss(loss_fn, regression)
    except NotImplementedError:
        if not callable(loss_fn) or isinstance(loss_fn, functools.partial) and loss_fn.func in (focal_loss_fn, scaled_mse_fn):
            raise
        return ("user", loss_fn)


def _user_loss_and_grad(fn, outputs: torch.Tensor, targets: torch.Tensor, log_temp, identity: bool, regression: bool):
    """Classification: z' = z exp(-log_temp), dloss/dlog_temp = -sum dloss/dz' . z' (output_calibrator/classification.py:
    7-17).  Regression: log var' = log var + log_temp, dloss/dlog_temp = sum dloss/dlog var' (regression.py:7-19)."""
    zc = outputs.clone()
    if not identity:
        lt = torch.as_tensor(np.array(log_temp, dtype=np.float32).reshape(1), device=outputs.device)
        if regression:
            d = outputs.shape[1] // 2
            zc[:, d:] += lt  # the calibrated log variance is a one-scalar shift
        else:
            ops.softmax_rows_(zc, ops.EPI_LOGITS, log_temp=lt)
    leaf = zc.requires_grad_(True)
    with torch.enable_grad():
        loss = fn(leaf, targets)
        loss.backward()
    g = leaf.grad.contiguous()
    if identity:
        return np.float32(loss.item()), np.float32(0.0)
    if regression:
        d = outputs.shape[1] // 2
        glv = g[:, d:].contiguous()
        dlt = ops.dot_f32(glv, torch.ones_like(glv)).item()
    else:
        dlt = -ops.dot_f32(g, zc.detach().contiguous()).item()
    return np.float32(loss.item()), np.float32(dlt)


def loss_and_grad(kind, outputs: torch.Tensor, targets: torch.Tensor, log_temp: np.ndarray, identity: bool = False):
    """(loss, dloss/dlog_temp) at log_temp, float32 scalars; the mean over the calibration inputs."""
    if kind[0] == "user":
        return _user_loss_and_grad(kind[1], outputs, targets, log_temp, identity, kind[2])
    lt = None if identity else torch.as_tensor(np.array(log_temp, dtype=np.float32).reshape(1), device=outputs.device)
    n = outputs.shape[0]
    if kind[0] == "focal":
        t = ops.output_calib_focal_loss_terms(outputs, targets, lt, kind[1]).cpu().numpy()
        tau = float(np.exp(-np.float64(log_temp[0])))
        return np.float32(t[0] / n), np.float32(0.0 if identity else -tau * t[1] / n)  # dtau/dlog_temp = -tau
    t = ops.output_calib_scaled_mse_terms(outputs, targets, lt).cpu().numpy()
    return np.float32(t[0] / n), np.float32(0.0 if identity else t[1] / n)


class OutputCalibModelCalibrator:
    def __init__(self, calib_outputs, calib_targets, val_outputs=None, val_targets=None, predict_fn: Callable = None,
                 uncertainty_fn: Callable = None, regression: bool = False, identity: bool = False,
                 save_checkpoint_dir=None, save_every_n_steps: Optional[int] = None, keep_top_n_checkpoints: int = 2,
                 disable_training_metrics_computation: bool = False, eval_every_n_epochs: int = 1,
                 early_stopping_monitor: str = "val_loss", early_stopping_min_delta: float = 0.0,
                 early_stopping_patience: Optional[int] = 0):
        self.regression, self.identity = regression, identity
        f32 = lambda x: None if x is None else (x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))).to(
            device="cuda", dtype=torch.float32).contiguous()
        tgt = (lambda x, n: None if x is None else f32(x).reshape(n, -1).contiguous()) if regression else (
            lambda x, n: None if x is None else (x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))).to(
                device="cuda", dtype=torch.int32).contiguous())
        self._calib_outputs = f32(calib_outputs)
        self._calib_targets = tgt(calib_targets, self._calib_outputs.shape[0])
        self._val_outputs = f32(val_outputs)
        self._val_targets = None if val_outputs is None else tgt(val_targets, self._val_outputs.shape[0])
        self.predict_fn, self.uncertainty_fn = predict_fn, uncertainty_fn
        self.save_checkpoint_dir = save_checkpoint_dir
        self.save_every_n_steps = save_every_n_steps
        self.keep_top_n_checkpoints = keep_top_n_checkpoints
        self.disable_training_metrics_computation = disable_training_metrics_computation
        self.eval_every_n_epochs = eval_every_n_epochs
        self.early_stopping_monitor = early_stopping_monitor
        self.early_stopping_patience = early_stopping_patience
        self._early_stopping = None
        if early_stopping_patience is not None and early_stopping_patience > 0:
            self._early_stopping = EarlyStopping(early_stopping_min_delta, early_stopping_patience)

    @property
    def is_early_stopping_active(self) -> bool:
        return self._early_stopping is not None

    def _metrics(self, metrics, outputs, targets, log_temp, prefix: str = "") -> Dict[str, float]:
        vals = {}
        if metrics:
            preds = self.predict_fn(outputs, log_temp)
            uncertainties = self.uncertainty_fn(outputs, log_temp)
            for metric in metrics:
                v = metric(preds, uncertainties, targets)
                vals[prefix + metric.__name__] = float(v.item() if hasattr(v, "item") else v)
        return vals

    def train(self, state: OutputCalibState, loss_kind, n_epochs: int = 1, metrics: Optional[Tuple[Callable, ...]] = None,
              verbose: bool = True) -> Tuple[OutputCalibState, Dict[str, np.ndarray]]:
        training, validation = collections.defaultdict(list), collections.defaultdict(list)
        for epoch in range(n_epochs):
            # training step: the whole calibration set is one batch (:170-257)
            loss, grad = loss_and_grad(loss_kind, self._calib_outputs, self._calib_targets, state.log_temp, self.identity)
            lt_before = state.log_temp
            state = state.apply_gradients(np.array([grad], np.float32))
            if self.save_checkpoint_dir and self.save_every_n_steps and epoch % self.save_every_n_steps == 0:
                save_state(state, self.save_checkpoint_dir, keep=self.keep_top_n_checkpoints)
            cur = {"loss": float(loss)}
            if not self.disable_training_metrics_computation and metrics is not None:
                # aux["outputs"] of the step: calibrated with the parameters BEFORE the update (:269-287)
                cur.update(self._metrics(metrics, self._calib_outputs, self._calib_targets, lt_before))
            for k, v in cur.items():
                training[k].append(v)
            if verbose:
                logging.info(f"Epoch: {epoch + 1} | " + " | ".join(f"{m}: {round(float(v), 5)}" for m, v in cur.items()))
            # validation (:336-406, :455-463)
            if self._val_outputs is not None and self.eval_every_n_epochs and self.eval_every_n_epochs > 0 \
                    and epoch % self.eval_every_n_epochs == 0:
                vloss, _ = loss_and_grad(loss_kind, self._val_outputs, self._val_targets, state.log_temp, self.identity)
                cur = {"val_loss": float(vloss)}
                cur.update(self._metrics(metrics, self._val_outputs, self._val_targets, state.log_temp, "val_"))
                if self.is_early_stopping_active:
                    improved = self._early_stopping.update(cur[self.early_stopping_monitor]).has_improved
                    if improved and self.save_checkpoint_dir:
                        save_state(state, self.save_checkpoint_dir, force_save=True)
                for k, v in cur.items():
                    validation[k].append(v)
                if verbose:
                    logging.info(f"Epoch: {epoch + 1} | " + " | ".join(f"{m}: {round(float(v), 5)}" for m, v in cur.items()))
                if self.is_early_stopping_active and self._early_stopping.should_stop:
                    logging.info("[Early Stopping] Stopping training...")
                    break
        status = {k: np.array(v, np.float32) for k, v in training.items()}
        status.update({k: np.array(v, np.float32) for k, v in validation.items()})
        # on_train_end (:479-486)
        save_state(state, self.save_checkpoint_dir, keep=self.keep_top_n_checkpoints, force_save=True)
        return state, status
