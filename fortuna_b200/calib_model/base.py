"""``CalibModel`` - fortuna/calib_model/base.py:43-242: calibrate a model against a loss over a data loader, load / save
its state."""
import logging
from typing import Callable, Optional

import torch

from ..prob_model.posterior.sgmcmc.torch_model import BoundModel
from ..utils.random import RandomNumberGenerator
from .calib_model_calibrator import CalibModelCalibrator, resolve_calib_loss
from .config.base import Config
from .state import CalibState, CalibStateRepository, restore_into


class CalibModel:
    _regression = False

    def __init__(self, model: torch.nn.Module, seed: int = 0, device="cuda"):
        self.model = model
        self.device = torch.device(device)
        self.rng = RandomNumberGenerator(seed=seed, device=self.device)
        self.predictive.rng = self.rng

    def _new_state(self, freeze_fun=None) -> CalibState:
        self.model.to(self.device)
        bound = BoundModel(self.model, self.device, freeze_fun=freeze_fun)
        return CalibState(bound, torch.zeros_like(bound.params.flat), torch.zeros_like(bound.params.flat), 0)

    def _init_state(self, config: Config) -> CalibState:
        """base.py:214-242."""
        ck = config.checkpointer
        if ck.restore_checkpoint_path is None:
            if ck.start_from_current_state:
                if self.predictive.state is None:
                    raise ValueError("`start_from_current_state` needs a calibrated or loaded model.")
                return self.predictive.state.get()
            return self._new_state(config.optimizer.freeze_fun)
        if ck.start_from_current_state:
            logging.warning("`config.checkpointer.start_from_current_state` will be ignored since "
                            "`config.checkpointer.restore_checkpoint_path` is given.")
        return restore_into(self._new_state(config.optimizer.freeze_fun), ck.restore_checkpoint_path)

    def _calibrate(self, calib_data_loader, uncertainty_fn: Callable, loss_fn: Callable, val_data_loader=None,
                   config: Optional[Config] = None):
        """base.py:62-154."""
        config = config or Config()
        if config.checkpointer.dump_state is True and not config.checkpointer.save_checkpoint_dir:
            raise ValueError("`save_checkpoint_dir` must be passed when `dump_state` is set to True.")
        loss_kind = resolve_calib_loss(loss_fn, self._regression)
        trainer = CalibModelCalibrator(
            predict_fn=self._predict_fn, uncertainty_fn=uncertainty_fn, regression=self._regression,
            save_checkpoint_dir=config.checkpointer.save_checkpoint_dir,
            save_every_n_steps=config.checkpointer.save_every_n_steps,
            keep_top_n_checkpoints=config.checkpointer.keep_top_n_checkpoints,
            disable_training_metrics_computation=config.monitor.disable_calibration_metrics_computation,
            eval_every_n_epochs=config.monitor.eval_every_n_epochs,
            early_stopping_monitor=config.monitor.early_stopping_monitor,
            early_stopping_min_delta=config.monitor.early_stopping_min_delta,
            early_stopping_patience=config.monitor.early_stopping_patience,
        )
        state = self._init_state(config)
        if config.monitor.verbose:
            logging.info("Start calibration.")
        state, status = trainer.train(state=state, loss_kind=loss_kind, training_dataloader=calib_data_loader,
                                      optimizer=config.optimizer.method, n_epochs=config.optimizer.n_epochs,
                                      metrics=config.monitor.metrics, validation_dataloader=val_data_loader,
                                      verbose=config.monitor.verbose, callbacks=config.callbacks)
        self.predictive.state = CalibStateRepository(
            config.checkpointer.save_checkpoint_dir if config.checkpointer.dump_state is True else None)
        self.predictive.state.put(state, keep=config.checkpointer.keep_top_n_checkpoints)
        if config.monitor.verbose:
            logging.info("Calibration completed.")
        return status

    def load_state(self, checkpoint_path) -> None:
        """base.py:156-174: the parameters of the latest checkpoint under ``checkpoint_path`` into this model."""
        state = restore_into(self._new_state(), checkpoint_path)
        self.predictive.state = CalibStateRepository(checkpoint_dir=str(checkpoint_path))
        self.predictive.state._state = state

    def save_state(self, checkpoint_path, keep_top_n_checkpoints: int = 1) -> None:
        if self.predictive.state is None:
            raise ValueError("No state available. You must first either calibrate the model, or load a saved checkpoint.")
        return self.predictive.state.put(self.predictive.state.get(), checkpoint_path=checkpoint_path,
                                         keep=keep_top_n_checkpoints)
