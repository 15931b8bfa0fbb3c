"""Calibration loop of the calibration models - ``CalibModelCalibrator`` (fortuna/calib_model/calib_model_calibrator.py:29-
110) inside ``TrainerABC.train`` (fortuna/training/trainer.py:97-580): epochs over the calibration loader, one optimizer
step per batch, per-epoch means of loss and metrics, validation every ``eval_every_n_epochs`` with early stopping.

What runs where: the backbone's forward / backward is the user's torch module (outside the hot path); the loss head is ONE
kernel over the batch outputs returning the loss and the gradient in the outputs (``fb200_focal_loss_grad`` /
``fb200_scaled_mse_loss_grad``) that seeds the backward; the optimizer is ONE kernel over the flat parameter vector
(``fb200_adam_step_f32`` = optax.adam, the reference's default).  A user-supplied ``loss_fn(outputs, targets)`` (a torch
callable) is differentiated by autograd instead of the fused head."""
import collections
import logging
from typing import Callable, Dict, Optional, Tuple

import numpy as np
import torch

from .. import ops
from ..output_calib_model.output_calib_model_calibrator import EarlyStopping, resolve_loss
from .state import CalibState, save_state


def resolve_calib_loss(loss_fn: Callable, regression: bool):
    """("focal", gamma) / ("scaled_mse", None) for the losses with a fused head, else ("user", loss_fn)."""
    try:
        return resolve_loss(loss_fn, regression)
    except NotImplementedError:
        if not callable(loss_fn):
            raise
        return ("user", loss_fn)


class CalibModelCalibrator:
    def __init__(self, predict_fn: Callable, uncertainty_fn: Callable, regression: bool = False,
                 save_checkpoint_dir=None, save_every_n_steps: Optional[int] = None, keep_top_n_checkpoints: int = 2,
                 disable_training_metrics_computation: bool = False, eval_every_n_epochs: int = 1,
                 early_stopping_monitor: str = "val_loss", early_stopping_min_delta: float = 0.0,
                 early_stopping_patience: Optional[int] = 0):
        self.predict_fn, self.uncertainty_fn, self.regression = predict_fn, uncertainty_fn, regression
        self.save_checkpoint_dir = save_checkpoint_dir
        self.save_every_n_steps = save_every_n_steps
        self.keep_top_n_checkpoints = keep_top_n_checkpoints
        self.disable_training_metrics_computation = disable_training_metrics_computation
        self.eval_every_n_epochs = eval_every_n_epochs
        self.early_stopping_monitor = early_stopping_monitor
        self._early_stopping = None
        if early_stopping_patience is not None and early_stopping_patience > 0:
            self._early_stopping = EarlyStopping(early_stopping_min_delta, early_stopping_patience)

    @property
    def is_early_stopping_active(self) -> bool:
        return self._early_stopping is not None

    # ------------------------------------------------------------------ one batch
    def _batch(self, state: CalibState, x, y):
        dev = state.bound.device
        xt = (x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))).to(device=dev, dtype=torch.float32)
        yt = y if isinstance(y, torch.Tensor) else torch.as_tensor(np.asarray(y))
        if self.regression:
            yt = yt.to(device=dev, dtype=torch.float32).reshape(xt.shape[0], -1).contiguous()
        else:
            yt = yt.to(device=dev, dtype=torch.int32).contiguous()
        return xt, yt

    def _loss_head(self, loss_kind, outputs: torch.Tensor, targets: torch.Tensor, want_grad: bool):
        """(loss as a python float, d loss / d outputs or None)."""
        z = outputs.detach().contiguous()
        if loss_kind[0] == "focal":
            loss, grad, _ = ops.focal_loss_grad(z, targets, loss_kind[1], want_grad)
            return float(loss.item()), grad
        if loss_kind[0] == "scaled_mse":
            loss, grad, _ = ops.scaled_mse_loss_grad(z, targets, want_grad)
            return float(loss.item()), grad
        leaf = z.clone().requires_grad_(want_grad)
        with torch.enable_grad():
            loss = loss_kind[1](leaf, targets)
            if want_grad:
                loss.backward()
        return float(loss.item()), (leaf.grad if want_grad else None)

    def training_step(self, state: CalibState, loss_kind, batch, optimizer) -> Tuple[float, torch.Tensor]:
        """training_loss_step + TrainState.apply_gradients (calib_model_calibrator.py:30-66, training/trainer.py:150-175)."""
        x, y = self._batch(state, *batch)
        state.bound.zero_grad()
        state.model.train()
        outputs = state.model(x)
        loss, grad = self._loss_head(loss_kind, outputs, y, True)
        outputs.backward(grad.reshape(outputs.shape))
        state.step += 1
        ops.adam_step_(state.bound.params.flat, state.bound.grads.flat, state.mu, state.nu, state.step,
                       optimizer.learning_rate, optimizer.b1, optimizer.b2, optimizer.eps)
        return loss, outputs.detach(), y

    def validation_step(self, state: CalibState, loss_kind, batch, metrics) -> Dict[str, float]:
        x, y = self._batch(state, *batch)
        state.model.eval()
        with torch.no_grad():
            outputs = state.model(x).float().contiguous()
        loss, _ = self._loss_head(loss_kind, outputs, y, False)
        out = {"val_loss": loss}
        out.update(self._metrics(metrics, outputs, y, "val_"))
        return out

    def _metrics(self, metrics, outputs, targets, prefix: str = "") -> Dict[str, float]:
        vals = {}
        if metrics:
            preds, unc = self.predict_fn(outputs), self.uncertainty_fn(outputs)
            for metric in metrics:
                v = metric(preds, unc, targets)
                vals[prefix + metric.__name__] = float(v.item() if hasattr(v, "item") else v)
        return vals

    # ------------------------------------------------------------------ the loop
    def train(self, state: CalibState, loss_kind, training_dataloader, optimizer, n_epochs: int = 1,
              metrics: Optional[Tuple[Callable, ...]] = None, validation_dataloader=None, verbose: bool = True,
              callbacks=None) -> Tuple[CalibState, Dict[str, np.ndarray]]:
        training, validation = collections.defaultdict(list), collections.defaultdict(list)
        for epoch in range(n_epochs):
            per_step = collections.defaultdict(list)
            for batch in training_dataloader:
                loss, outputs, y = self.training_step(state, loss_kind, batch, optimizer)
                per_step["loss"].append(loss)
                if not self.disable_training_metrics_computation and metrics is not None:
                    for k, v in self._metrics(metrics, outputs.float().contiguous(), y).items():
                        per_step[k].append(v)
                if self.save_checkpoint_dir and self.save_every_n_steps and state.step % self.save_every_n_steps == 0:
                    save_state(state, self.save_checkpoint_dir, keep=self.keep_top_n_checkpoints)
            cur = {k: float(np.mean(v)) for k, v in per_step.items()}  # _get_mean_losses_and_metrics
            for k, v in cur.items():
                training[k].append(v)
            if verbose:
                logging.info(f"Epoch: {epoch + 1} | " + " | ".join(f"{m}: {round(float(v), 5)}" for m, v in cur.items()))
            if validation_dataloader is not None and self.eval_every_n_epochs and self.eval_every_n_epochs > 0 \
                    and epoch % self.eval_every_n_epochs == 0:
                per_batch = collections.defaultdict(list)
                for batch in validation_dataloader:
                    for k, v in self.validation_step(state, loss_kind, batch, metrics).items():
                        per_batch[k].append(v)
                cur = {k: float(np.mean(v)) for k, v in per_batch.items()}
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
        save_state(state, self.save_checkpoint_dir, keep=self.keep_top_n_checkpoints, force_save=True)
        return state, status
