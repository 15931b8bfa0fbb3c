"""Configuration of the calibration models, field for field fortuna/calib_model/config/{base,optimizer,checkpointer,
monitor,processor}.py.  Monitor / Processor are those of the output calibration models (same fields)."""
from typing import Callable, List, Optional

from ...output_calib_model.config.base import Monitor, Processor  # noqa: F401
from ...prob_model.calib_config import Adam


class Optimizer:
    """config/optimizer.py:16-40: default ``optax.adam(1e-2)`` (run as ONE fused kernel over the flat parameter vector,
    fb200_adam_step_f32), 100 epochs; ``freeze_fun(path, value) -> "trainable" | "frozen"`` as in
    fortuna/utils/freeze.py:84-148."""

    def __init__(self, method: Optional[Adam] = None, n_epochs: int = 100, freeze_fun: Optional[Callable] = None):
        self.method = method if method is not None else Adam(1e-2)
        self.n_epochs = n_epochs
        self.freeze_fun = freeze_fun


class Checkpointer:
    """config/checkpointer.py:6-44."""

    def __init__(self, save_checkpoint_dir=None, restore_checkpoint_path=None, start_from_current_state: bool = False,
                 save_every_n_steps: Optional[int] = None, keep_top_n_checkpoints: Optional[int] = 2,
                 dump_state: bool = False):
        self.save_checkpoint_dir = save_checkpoint_dir
        self.restore_checkpoint_path = restore_checkpoint_path
        self.start_from_current_state = start_from_current_state
        self.save_every_n_steps = save_every_n_steps
        self.keep_top_n_checkpoints = keep_top_n_checkpoints
        self.dump_state = dump_state


class Config:
    """config/base.py:10-43."""

    def __init__(self, optimizer: Optional[Optimizer] = None, checkpointer: Optional[Checkpointer] = None,
                 monitor: Optional[Monitor] = None, processor: Optional[Processor] = None,
                 callbacks: Optional[List] = None):
        self.optimizer = optimizer or Optimizer()
        self.checkpointer = checkpointer or Checkpointer()
        self.monitor = monitor or Monitor()
        self.processor = processor or Processor()
        self.callbacks = callbacks
