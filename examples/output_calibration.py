#!/usr/bin/env python
"""Calibration from model outputs (the reference's "from model outputs" usage mode,
docs/source/usage_modes/model_outputs.rst): temperature-scale given logits with ``OutputCalibClassifier``, then read the
calibrated predictive statistics; plus the binary MaxCovFixPrec calibrator on the positive-class probabilities."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fortuna_b200.conformal import MaxCoverageFixedPrecisionBinaryClassificationCalibrator  # noqa: E402
from fortuna_b200.metric.classification import expected_calibration_error  # noqa: E402
from fortuna_b200.output_calib_model import Config, Monitor, Optimizer, OutputCalibClassifier  # noqa: E402


def overconfident_logits(n, n_classes, seed, sharpness=4.0):
    """Logits of a classifier that is right 80 % of the time but far too sure of itself."""
    r = np.random.default_rng(seed)
    y = r.integers(0, n_classes, n)
    guess = np.where(r.uniform(size=n) < 0.8, y, r.integers(0, n_classes, n))
    z = r.standard_normal((n, n_classes)).astype(np.float32)
    z[np.arange(n), guess] += sharpness * 2
    return z, y.astype(np.int32)


def main():
    calib_z, calib_y = overconfident_logits(5000, 10, 0)
    test_z, test_y = overconfident_logits(5000, 10, 1)
    model = OutputCalibClassifier()
    status = model.calibrate(calib_z, calib_y, val_outputs=test_z, val_targets=test_y,
                             config=Config(optimizer=Optimizer(n_epochs=200), monitor=Monitor(verbose=False)))
    log_temp = float(model.predictive.state.get().log_temp[0])
    for calibrated in (False, True):
        probs = model.predictive.mean(test_z, calibrated=calibrated)
        preds = model.predictive.mode(test_z, calibrated=calibrated)
        ece = float(expected_calibration_error(preds, probs, test_y))
        nll = -float(model.predictive.log_prob(test_z, test_y, calibrated=calibrated).mean())
        print(f"calibrated={calibrated}: ECE {ece:.4f}  NLL {nll:.4f}")
    print(f"temperature exp(log_temp) = {np.exp(log_temp):.3f}; focal loss {status['loss'][0]:.4f} -> {status['loss'][-1]:.4f}")

    # binary case: scale the scores so that predicted positives / negatives are right at least 90 % of the time
    r = np.random.default_rng(2)
    p = r.uniform(0, 1, 20000).astype(np.float32)
    y = (r.uniform(0, 1, 20000) < np.clip(1.4 * (p - 0.5) + 0.5, 0, 1)).astype(np.int32)
    cal = MaxCoverageFixedPrecisionBinaryClassificationCalibrator()
    patched = cal.calibrate(targets=y, probs=p, true_positive_precision_threshold=0.9,
                            true_negative_precision_threshold=0.9, test_probs=p)
    print("MaxCovFixPrec patches:", {k: float(v) for k, v in cal.patches.items()},
          "| TP precision", float(cal.true_positive_precision(patched, y, 0.9)),
          "coverage", float(cal.true_positive_coverage(patched, 0.9)),
          "| TN precision", float(cal.true_negative_precision(patched, y, 0.9)),
          "coverage", float(cal.true_negative_coverage(patched, 0.1)))


if __name__ == "__main__":
    main()
    torch.cuda.synchronize()
