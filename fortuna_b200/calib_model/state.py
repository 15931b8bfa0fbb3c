"""``CalibState`` / ``CalibStateRepository`` - fortuna/calib_model/state.py:15-94 and calib_state_repository.py: the
trainable state of a calibration model (here: the torch module whose trainable parameters are views of one flat vector,
the Adam moments over that vector, the step count) and where it is kept / saved.  On disk: the reference's checkpoint
container (flax msgpack, ``checkpoint_<step>``, fortuna/training/mixin.py:32-85) holding ``params`` by module path,
``opt_state`` and ``step``; leaf names are the torch module's (a flax-written CalibState has flax leaf names: see
DESIGN 7)."""
import os
from typing import Optional

import numpy as np
import torch

from ..training import checkpoint as ckpt


class CalibState:
    def __init__(self, bound, mu: torch.Tensor, nu: torch.Tensor, step: int = 0):
        self.bound, self.mu, self.nu, self.step = bound, mu, nu, step

    @property
    def model(self) -> torch.nn.Module:
        return self.bound.model

    @property
    def params(self):
        """{("model", "params", <module path>...): tensor view} of the trainable leaves."""
        return dict(zip(self.bound.params.paths, self.bound.params.leaf_views()))

    def to_dict(self):
        tree = {}
        for path, v in self.params.items():
            d = tree
            for k in path[:-1]:
                d = d.setdefault(k, {})
            d[path[-1]] = v.detach().cpu().numpy()
        return {"params": tree, "step": np.int32(self.step),
                "opt_state": {"count": np.int32(self.step), "mu": self.mu.cpu().numpy(), "nu": self.nu.cpu().numpy()}}

    def load_dict(self, d) -> None:
        flat = self.bound.params
        for path, view in zip(flat.paths, flat.leaf_views()):
            node = d["params"]
            for k in path:
                node = node[k]
            view.copy_(torch.as_tensor(np.array(node)).reshape(view.shape))
        self.step = int(d.get("step", 0))
        opt = d.get("opt_state")
        if opt is not None:
            self.mu.copy_(torch.as_tensor(np.array(opt["mu"])))
            self.nu.copy_(torch.as_tensor(np.array(opt["nu"])))


class CalibStateRepository:
    def __init__(self, checkpoint_dir: Optional[str] = None):
        self.checkpoint_dir = checkpoint_dir
        self._state: Optional[CalibState] = None

    def get(self) -> CalibState:
        if self._state is None:
            raise ValueError("No state available.")
        return self._state

    def put(self, state: CalibState, checkpoint_path: Optional[str] = None, keep: int = 1) -> None:
        self._state = state
        path = checkpoint_path or self.checkpoint_dir
        if path:
            save_state(state, path, keep=keep, force_save=True)


def save_state(state: CalibState, directory, keep: int = 1, force_save: bool = False, prefix: str = "checkpoint_"):
    if not directory:
        return
    os.makedirs(str(directory), exist_ok=True)
    ckpt.save_checkpoint(str(directory), state.to_dict(), step=int(state.step), prefix=prefix, keep=keep or 1,
                         overwrite=force_save)


def restore_into(state: CalibState, path, prefix: str = "checkpoint_") -> CalibState:
    d = ckpt.restore_checkpoint(str(path), prefix=prefix)
    if d is None:
        raise ValueError(f"No checkpoint was found in `checkpoint_path={path}`.")
    state.load_dict(d)
    return state
