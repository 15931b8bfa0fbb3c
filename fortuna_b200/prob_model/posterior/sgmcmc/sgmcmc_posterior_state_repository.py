"""On-device sample bank - replaces SGMCMCPosteriorStateRepository
(fortuna/prob_model/posterior/sgmcmc/sgmcmc_posterior_state_repository.py:27-159).

The reference keeps each stored sample as a host pytree (or a msgpack checkpoint on disk) and fetches it
through ``jax.pure_callback`` for every predictive draw.  Here the ``size`` samples live in one
``[size, N]`` float32 CUDA matrix (N = trainable parameters, flat in tree-flatten order, i.e. only the
``which_params`` leaves when part of the model is frozen); ``put`` is a device-to-device row copy
(fb200_bank_put_f32) and ``get`` returns a zero-copy FlatTree view of a row.

With ``checkpoint_dir`` the bank is written through to disk in the reference's own layout - one flax-msgpack file
``<checkpoint_dir>/<i>/checkpoint_<step>`` per sample (posterior_multi_state_repository.py:20-29,
fortuna/training/mixin.py:32-58), holding the state dict of an ``SGHMCState`` / ``CyclicalSGLDState`` with only the
trainable leaves under ``params`` (:139-157) - and ``load_checkpoints`` fills the bank from such a directory, whether
Fortuna or this package wrote it.  The predictive path never touches the files: it reads the bank.
"""

import os
from typing import Dict, List, Optional

import numpy as np
import torch

from .... import ops
from ....training import checkpoint as ckpt
from ....utils.flat_tree import FlatTree


def _key_u32(key) -> np.ndarray:
    k = key.detach().cpu().numpy() if isinstance(key, torch.Tensor) else np.asarray(key)
    return k.view(np.uint32) if k.dtype == np.int32 else k.astype(np.uint32)


def _opt_state_dict(opt_state) -> Optional[Dict]:
    """OptaxSGHMCState / OptaxCyclicalSGLDState as flax's to_state_dict lays them out."""
    if opt_state is None:
        return None
    if hasattr(opt_state, "sgld_state"):  # optax.sgd(1.0) = chain(identity, scale(-1)): two array-free states
        return {"sgd_state": {"0": {}, "1": {}}, "sgld_state": _opt_state_dict(opt_state.sgld_state)}
    pre = opt_state.preconditioner_state
    pre_d = {f: ckpt.to_state_dict(getattr(pre, f)) for f in getattr(pre, "_fields", ())}
    return {
        "count": np.asarray(opt_state.count, dtype=np.int32),
        "rng_key": _key_u32(opt_state.rng_key),
        "momentum": ckpt.to_state_dict(opt_state.momentum),
        "preconditioner_state": pre_d,
    }


def posterior_state_dict(state, name: str = "SGHMCState", which_params=None, calib_params=None,
                         calib_mutable=None) -> Dict:
    """The dict ``flax.serialization.to_state_dict`` gives for the reference's SG-MCMC posterior state
    (fortuna/prob_model/posterior/state.py:24-36, sgmcmc/sghmc/sghmc_state.py:23-32): every struct field that is a
    pytree node, ``apply_fn`` / ``tx`` (static) left out."""
    params = state.params if hasattr(state, "params") else state
    enc = None
    if which_params is not None:
        enc = {str(i): {str(j): np.array([ord(c) for c in s]) for j, s in enumerate(path)}
               for i, path in enumerate(which_params)}
    return {
        "step": int(getattr(state, "step", 0)),
        "params": ckpt.to_state_dict(params),
        "opt_state": _opt_state_dict(getattr(state, "opt_state", None)),
        "encoded_name": ckpt.to_state_dict(ckpt.encode_name(name)),
        "frozen_params": None,
        "dynamic_scale": None,
        "mutable": ckpt.to_state_dict(getattr(state, "mutable", None)),
        "calib_params": ckpt.to_state_dict(calib_params),
        "calib_mutable": ckpt.to_state_dict(calib_mutable),
        "grad_accumulated": None,
        "_encoded_which_params": enc,
    }


class SGMCMCPosteriorStateRepository:
    def __init__(self, size: int, template: Optional[FlatTree] = None, checkpoint_dir=None, all_params=None,
                 which_params=None, state_name: str = "SGHMCState", device="cuda"):
        self.size = int(size)
        self.checkpoint_dir = str(checkpoint_dir) if checkpoint_dir else None
        self._template = template
        self._all_params = all_params
        self._which_params = which_params
        self._state_name = state_name
        self._device = device
        self.bank: Optional[torch.Tensor] = None
        self._filled = [False] * self.size
        self.steps: List[Optional[int]] = [None] * self.size
        if template is not None:
            self._allocate(template)

    def _allocate(self, template: FlatTree):
        self._template = template
        self.bank = torch.zeros((self.size, template.flat.numel()), dtype=torch.float32, device=template.flat.device)

    def _dir(self, i: int) -> str:
        return os.path.join(self.checkpoint_dir, str(i))

    def put(self, state, i: int, checkpoint_path=None, keep: int = 1, prefix: str = "checkpoint_", **_):
        params = state.params if hasattr(state, "params") else state
        if not isinstance(params, FlatTree):
            raise TypeError("the sample bank stores FlatTree parameters")
        if self.bank is None:
            self._allocate(params)
        if not (0 <= i < self.size):
            raise IndexError(i)
        ops.bank_put(self.bank, i, params.flat)
        self._filled[i] = True
        self.steps[i] = int(getattr(state, "step", 0))
        if checkpoint_path or self.checkpoint_dir:
            # trainer.save_checkpoint(force_save=True): fortuna/training/train_state_repository.py:46-60
            ckpt.save_checkpoint(
                checkpoint_path or self._dir(i),
                posterior_state_dict(state, self._state_name, self._which_params),
                step=self.steps[i], prefix=prefix, keep=keep, overwrite=True,
            )

    def _row_from_dict(self, d: Dict, i: int):
        tree = d["params"]
        if self.bank is None:
            self._allocate(FlatTree.from_tree(tree, device=self._device))
        row = FlatTree.from_tree(tree, device=self.bank.device)
        if not row.same_layout(self._template):
            raise ValueError(f"checkpoint {i} holds a different parameter tree than the bank")
        ops.bank_put(self.bank, i, row.flat)
        self._filled[i] = True
        self.steps[i] = int(d.get("step", 0))

    def load_checkpoints(self, checkpoint_dir=None, prefix: str = "checkpoint_") -> "SGMCMCPosteriorStateRepository":
        """Fill the bank from ``<dir>/<i>/<prefix><step>`` for i < size (the files PosteriorMultiStateRepository
        writes); the state class name inside each file must be an SG-MCMC state."""
        root = str(checkpoint_dir) if checkpoint_dir else self.checkpoint_dir
        if root is None:
            raise ValueError("No checkpoint directory available.")
        for i in range(self.size):
            d = ckpt.restore_checkpoint(os.path.join(root, str(i)), prefix=prefix)
            if d is None:
                raise ValueError(f"No checkpoint was found in `restore_checkpoint_path={os.path.join(root, str(i))}`.")
            name = ckpt.decode_name(d)
            if name not in ("SGHMCState", "CyclicalSGLDState"):
                raise ValueError(f"checkpoint {i} holds a `{name}`, not an SG-MCMC posterior state")
            self._row_from_dict(d, i)
        return self

    def get(self, i: Optional[int] = None, **_):
        if i is None:
            return [self.get(j) for j in range(self.size)]
        if (self.bank is None or not self._filled[i]) and self.checkpoint_dir:
            d = ckpt.restore_checkpoint(self._dir(i))
            if d is None:
                raise ValueError("No state available.")
            self._row_from_dict(d, i)
        if self.bank is None:
            raise ValueError("No state available.")
        t = self._template
        return FlatTree(self.bank[i], t.paths, t.shapes, t.offsets)

    def pull(self, i: int, prefix: str = "checkpoint_", **_):
        out = self.get(i).clone()
        self._filled[i] = False
        if self.checkpoint_dir:
            path = ckpt.latest_checkpoint(self._dir(i), prefix)
            if path:
                os.remove(path)
        return out

    @property
    def n_filled(self) -> int:
        return sum(self._filled)

    def leaf_matrix(self, path) -> torch.Tensor:
        """[size, *leaf_shape] contiguous copy of one leaf across all samples (e.g. the last-layer kernel)."""
        t = self._template
        k = t.paths.index(tuple(path))
        o, n, s = t.offsets[k], t.lens[k], t.shapes[k]
        return self.bank[:, o : o + n].reshape((self.size,) + tuple(s)).contiguous()
