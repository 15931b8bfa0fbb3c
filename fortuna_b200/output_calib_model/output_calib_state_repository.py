"""``OutputCalibStateRepository`` - fortuna/output_calib_model/output_calib_state_repository.py:1-10 over
fortuna/training/train_state_repository.py:22-90 and the restore of
fortuna/output_calib_model/output_calib_mixin.py:15-40: the state lives in memory, or - with a ``checkpoint_dir`` - as
``checkpoint_<step>`` files in the reference's flax-msgpack format, restored on every ``get``."""
import os
from typing import Any, Optional

from ..training import checkpoint
from .state import OutputCalibState


def restore_state(path, optimizer: Any = None, prefix: str = "checkpoint_") -> OutputCalibState:
    if not os.path.isdir(path) and not os.path.isfile(path):
        raise ValueError(f"`restore_checkpoint_path={path}` was not found.")
    d = checkpoint.restore_checkpoint(str(path), prefix=prefix)
    if d is None:
        raise ValueError(f"No checkpoint was found in `restore_checkpoint_path={path}`.")
    return OutputCalibState.init_from_dict(d, optimizer)


def save_state(state: OutputCalibState, directory, keep: int = 1, force_save: bool = False, prefix: str = "checkpoint_"):
    """fortuna/training/mixin.py:32-57."""
    if directory:
        checkpoint.save_checkpoint(str(directory), state.to_state_dict(), state.step, prefix=prefix, keep=keep,
                                   overwrite=force_save)


class OutputCalibStateRepository:
    def __init__(self, checkpoint_dir: Optional[str] = None):
        self.checkpoint_dir = checkpoint_dir
        self.__state = None

    def get(self, checkpoint_path=None, optimizer: Any = None, prefix: str = "checkpoint_") -> OutputCalibState:
        if not checkpoint_path and self.checkpoint_dir:
            checkpoint_path = self.checkpoint_dir
        if checkpoint_path:
            return restore_state(checkpoint_path, optimizer, prefix)
        if self.__state is None:
            raise ValueError("No state available.")
        if optimizer is not None:
            return OutputCalibState.init(self.__state.params, self.__state.mutable, optimizer, step=self.__state.step)
        return self.__state

    def put(self, state: OutputCalibState, checkpoint_path=None, keep: int = 1, prefix: str = "checkpoint_") -> None:
        if not checkpoint_path and self.checkpoint_dir:
            checkpoint_path = self.checkpoint_dir
        if checkpoint_path:
            save_state(state, checkpoint_path, keep=keep, force_save=True, prefix=prefix)
        else:
            self.__state = state
