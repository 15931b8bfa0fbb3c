"""``OutputCalibModel`` - fortuna/output_calib_model/base.py:37-177: calibrate given model outputs, load / save the
calibration state."""
import logging
from typing import Callable, Optional

import numpy as np

from ..utils.random import RandomNumberGenerator
from .config.base import Config
from .output_calib_model_calibrator import OutputCalibModelCalibrator, resolve_loss_or_user
from .output_calib_state_repository import OutputCalibStateRepository, restore_state
from .state import OutputCalibState


class OutputCalibManager:
    """fortuna/output_calibrator/output_calib_manager/base.py:12-115 for the temperature scalers: holds the calibrator;
    its ``apply`` lives inside the kernels (the ``log_temp`` argument of every entry point)."""

    def __init__(self, output_calibrator=None):
        self.output_calibrator = output_calibrator

    def init(self, output_dim: int = None):
        return {"params": {"log_temp": np.zeros(1, np.float32)}, "mutable": None}


class _LazyRNG:
    """The key chain of fortuna/utils/random.py:26-49, created on the device at first use."""

    def __init__(self, seed: int):
        self.seed, self._chain = seed, None

    def get(self):
        if self._chain is None:
            self._chain = RandomNumberGenerator(seed=self.seed)
        return self._chain.get()


class OutputCalibModel:
    _regression = False

    def __init__(self, seed: int = 0):
        self.rng = _LazyRNG(seed)
        self.output_calib_manager.rng = self.rng
        self.predictive.rng = self.rng

    def _calibrate(self, uncertainty_fn: Callable, loss_fn: Callable, calib_outputs, calib_targets, val_outputs=None,
                   val_targets=None, config: Optional[Config] = None):
        config = config or Config()
        if (val_targets is not None and val_outputs is None) or (val_targets is None and val_outputs is not None):
            raise ValueError("For validation, both `val_outputs` and `val_targets` must be passed as arguments.")
        loss_kind = resolve_loss_or_user(loss_fn, self._regression)
        if loss_kind[0] == "user":
            loss_kind = ("user", loss_kind[1], self._regression)
        identity = self.output_calibrator is None
        calibrator = OutputCalibModelCalibrator(
            calib_outputs=calib_outputs, calib_targets=calib_targets, val_outputs=val_outputs, val_targets=val_targets,
            predict_fn=self._predict_fn, uncertainty_fn=uncertainty_fn, regression=self._regression, identity=identity,
            save_checkpoint_dir=config.checkpointer.save_checkpoint_dir,
            save_every_n_steps=config.checkpointer.save_every_n_steps,
            keep_top_n_checkpoints=config.checkpointer.keep_top_n_checkpoints,
            disable_training_metrics_computation=config.monitor.disable_calibration_metrics_computation,
            eval_every_n_epochs=config.monitor.eval_every_n_epochs,
            early_stopping_monitor=config.monitor.early_stopping_monitor,
            early_stopping_min_delta=config.monitor.early_stopping_min_delta,
            early_stopping_patience=config.monitor.early_stopping_patience,
        )
        if config.checkpointer.restore_checkpoint_path is None:
            init = self.output_calib_manager.init(output_dim=np.shape(calib_outputs)[-1])
            state = OutputCalibState.init(params={"output_calibrator": {"params": init["params"]}},
                                          mutable={"output_calibrator": init["mutable"]}, optimizer=config.optimizer.method)
        else:
            state = restore_state(config.checkpointer.restore_checkpoint_path, optimizer=config.optimizer.method)
        if config.monitor.verbose:
            logging.info("Start calibration.")
        state, status = calibrator.train(state=state, loss_kind=loss_kind, n_epochs=config.optimizer.n_epochs,
                                         metrics=config.monitor.metrics, verbose=config.monitor.verbose)
        self.predictive.state = OutputCalibStateRepository(
            config.checkpointer.save_checkpoint_dir if config.checkpointer.dump_state is True else None)
        self.predictive.state.put(state, keep=config.checkpointer.keep_top_n_checkpoints)
        return status

    def load_state(self, checkpoint_path) -> None:
        try:
            restore_state(checkpoint_path)
        except ValueError:
            raise ValueError(f"No checkpoint was found in `checkpoint_path={checkpoint_path}`.")
        self.predictive.state = OutputCalibStateRepository(checkpoint_dir=checkpoint_path)

    def save_state(self, checkpoint_path, keep_top_n_checkpoints: int = 1) -> None:
        if self.predictive.state is None:
            raise ValueError("""No state available. You must first either calibrate the model, or load a saved checkpoint.""")
        return self.predictive.state.put(self.predictive.state.get(), checkpoint_path=checkpoint_path,
                                         keep=keep_top_n_checkpoints)
