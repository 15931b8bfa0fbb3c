"""Temperature-scaling calibration of a fitted ProbClassifier - ``ProbModel._calibrate`` (fortuna/prob_model/base.py:
109-235) with ``ProbModelOutputCalibrator`` (fortuna/prob_model/prob_model_calibrator.py:30-61) for the default
``ClassificationTemperatureScaler`` (fortuna/output_calibrator/classification.py:7-17).

Like the reference, the ensemble of outputs on the calibration data is computed ONCE (S drawn posterior samples, uncalibrated
logits [S, n, C]); every optimisation step then needs the loss
    -( logsumexp_s( (n_data / batch) sum_b log softmax(exp(-log_temp) z[s,b])[y_b] ) - log S )
and its derivative in log_temp.  One kernel pass (fb200_calib_cls_loss_terms) returns the per-sample sums and their
derivatives in tau = exp(-log_temp); the scalar chain rule and the Adam step run on the host in float32."""
from typing import Dict, Optional

import numpy as np
import torch

from .. import ops
from .calib_config import CalibConfig


class ClassificationTemperatureScaler:
    """Marker for the reference's default output calibrator: logits * exp(-log_temp), log_temp initialised at 0."""


class RegressionTemperatureScaler:
    """Marker for the reference's default regression calibrator (fortuna/output_calibrator/regression.py:7-19):
    log variance + log_temp, log_temp initialised at 0."""


def loss_and_grad(z: torch.Tensor, targets: torch.Tensor, log_temp: np.float32, n_data: int, regression: bool = False):
    """(loss, dloss/dlog_temp) of one batch at log_temp, float32 scalars.  Classification: z [S, b, C] logits, int32
    targets; regression: z [S, b, 2d] = [mu | log var], float32 targets [b, d]."""
    lt = torch.tensor([float(log_temp)], dtype=torch.float32, device=z.device)
    if regression:
        terms = ops.calib_reg_loss_terms(z, targets, lt).cpu().numpy()  # [2, S] float64: L_s, dL_s/dlog_temp
    else:
        terms = ops.calib_cls_loss_terms(z, targets, lt).cpu().numpy()  # [2, S] float64: L_s, dL_s/dtau
    S, b = z.shape[0], z.shape[1]
    scale = n_data / b
    a = scale * terms[0]
    m = a.max()
    w = np.exp(a - m)
    log_pred = m + np.log(w.sum()) - np.log(S)
    w /= w.sum()
    dlogpred = float((w * scale * terms[1]).sum())
    if regression:
        return np.float32(-log_pred), np.float32(-dlogpred)
    tau = float(np.exp(-np.float64(log_temp)))
    # loss = -log_pred;  dtau/dlog_temp = -tau
    return np.float32(-log_pred), np.float32(dlogpred * tau)


def calibrate_temperature(z_calib: torch.Tensor, targets_calib, batch_size: int, calib_config: Optional[CalibConfig] = None,
                          z_val: Optional[torch.Tensor] = None, targets_val=None, val_batch_size: Optional[int] = None,
                          regression: bool = False) -> Dict[str, np.ndarray]:
    """Optimise log_temp over the fixed ensemble z_calib [S, n, C].  Batches are consecutive slices of ``batch_size``
    inputs (the calibration loader must not shuffle: fortuna/prob_model/base.py:128).  Returns the status dict
    (per-epoch mean ``loss`` / ``val_loss``) and the float32[1] ``log_temp``."""
    cfg = calib_config or CalibConfig()
    dev = z_calib.device
    n = z_calib.shape[1]
    tdt = torch.float32 if regression else torch.int32
    t_all = torch.as_tensor(np.asarray(targets_calib), dtype=tdt, device=dev)
    tv_all = None if z_val is None else torch.as_tensor(np.asarray(targets_val), dtype=tdt, device=dev)
    if regression:
        t_all = t_all.reshape(n, -1)
        tv_all = None if tv_all is None else tv_all.reshape(z_val.shape[1], -1)
    opt = cfg.optimizer.method
    log_temp = np.zeros(1, np.float32)
    state = opt.init(log_temp)
    losses, val_losses = [], []
    for _ in range(cfg.optimizer.n_epochs):
        ep = []
        for b0 in range(0, n, batch_size):
            zb = z_calib[:, b0 : b0 + batch_size].contiguous()
            loss, g = loss_and_grad(zb, t_all[b0 : b0 + batch_size].contiguous(), log_temp[0], n, regression)
            upd, state = opt.update(np.array([g], np.float32), state)
            log_temp = (log_temp + upd).astype(np.float32)
            ep.append(loss)
        losses.append(np.mean(ep))
        if z_val is not None:
            nv, vb = z_val.shape[1], val_batch_size or batch_size
            v = [loss_and_grad(z_val[:, b0 : b0 + vb].contiguous(), tv_all[b0 : b0 + vb].contiguous(), log_temp[0], nv, regression)[0]
                 for b0 in range(0, nv, vb)]
            val_losses.append(np.mean(v))
    status = {"loss": np.array(losses, np.float32)}
    if val_losses:
        status["val_loss"] = np.array(val_losses, np.float32)
    return status, log_temp
