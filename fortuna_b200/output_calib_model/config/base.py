"""Configuration objects of the output calibration models, field for field those of
fortuna/output_calib_model/config/{base,optimizer,checkpointer,monitor,processor}.py."""
from typing import Callable, Optional, Tuple

from ...prob_model.calib_config import Adam


class Optimizer:
    """config/optimizer.py:8-25: default ``optax.adam(1e-2)`` (restated in prob_model/calib_config.Adam), 100 epochs - an
    epoch is ONE step over the whole calibration set."""

    def __init__(self, method: Optional[Adam] = None, n_epochs: int = 100):
        self.method = method if method is not None else Adam(1e-2)
        self.n_epochs = n_epochs


class Checkpointer:
    """config/checkpointer.py:6-37."""

    def __init__(self, save_checkpoint_dir=None, restore_checkpoint_path=None, save_every_n_steps: Optional[int] = None,
                 keep_top_n_checkpoints: Optional[int] = 2, dump_state: bool = False):
        self.save_checkpoint_dir = save_checkpoint_dir
        self.save_every_n_steps = save_every_n_steps
        self.restore_checkpoint_path = restore_checkpoint_path
        self.keep_top_n_checkpoints = keep_top_n_checkpoints
        self.dump_state = dump_state


class Monitor:
    """config/monitor.py:13-80.  ``metrics`` are callables (predictions, uncertainty estimates, targets) -> value, applied
    to device tensors; ``uncertainty_fn`` maps calibrated outputs to the uncertainty estimates."""

    def __init__(self, metrics: Optional[Tuple[Callable, ...]] = None, uncertainty_fn: Optional[Callable] = None,
                 early_stopping_patience: int = 0, early_stopping_monitor: str = "val_loss",
                 early_stopping_min_delta: float = 0.0, eval_every_n_epochs: int = 1,
                 disable_calibration_metrics_computation: bool = False, verbose: bool = True):
        if metrics is not None:
            if type(metrics) != tuple:
                raise ValueError("`metrics` must be a tuple of callable metrics.")
            for metric in metrics:
                if not callable(metric):
                    raise ValueError(f"All metrics in `metrics` must be callable objects, but {metric} is not.")
        if uncertainty_fn is not None and not callable(uncertainty_fn):
            raise ValueError("`uncertainty_fn` must be a a callable function.")
        self.metrics = metrics
        self.uncertainty_fn = uncertainty_fn
        self.early_stopping_patience = early_stopping_patience
        self.early_stopping_monitor = early_stopping_monitor
        self.early_stopping_min_delta = early_stopping_min_delta
        self.eval_every_n_epochs = eval_every_n_epochs
        self.disable_calibration_metrics_computation = disable_calibration_metrics_computation
        self.verbose = verbose


class Processor:
    """config/processor.py:1-19.  One process drives one GPU here; ``devices`` / ``disable_jit`` are accepted and unused."""

    def __init__(self, devices: int = -1, disable_jit: bool = False):
        self.devices = devices
        self.disable_jit = disable_jit


class Config:
    """config/base.py:7-32."""

    def __init__(self, optimizer: Optional[Optimizer] = None, checkpointer: Optional[Checkpointer] = None,
                 monitor: Optional[Monitor] = None, processor: Optional[Processor] = None):
        self.optimizer = optimizer or Optimizer()
        self.checkpointer = checkpointer or Checkpointer()
        self.monitor = monitor or Monitor()
        self.processor = processor or Processor()
