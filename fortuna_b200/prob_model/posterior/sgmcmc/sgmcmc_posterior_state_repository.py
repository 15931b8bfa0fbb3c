"""On-device sample bank - replaces SGMCMCPosteriorStateRepository
(fortuna/prob_model/posterior/sgmcmc/sgmcmc_posterior_state_repository.py:27-159).

The reference keeps each stored sample as a host pytree (or a msgpack checkpoint on disk) and fetches it
through ``jax.pure_callback`` for every predictive draw.  Here the ``size`` samples live in one
``[size, N]`` float32 CUDA matrix (N = trainable parameters, flat in tree-flatten order, i.e. only the
``which_params`` leaves when part of the model is frozen); ``put`` is a device-to-device row copy
(fb200_bank_put_f32) and ``get`` returns a zero-copy FlatTree view of a row.
"""

from typing import List, Optional

import torch

from .... import ops
from ....utils.flat_tree import FlatTree


class SGMCMCPosteriorStateRepository:
    def __init__(self, size: int, template: Optional[FlatTree] = None, checkpoint_dir=None, all_params=None,
                 which_params=None):
        if checkpoint_dir is not None:
            raise NotImplementedError("msgpack checkpoints of the sample bank are outside the hot path (SURVEY 8f.1)")
        self.size = int(size)
        self._template = template
        self._all_params = all_params
        self._which_params = which_params
        self.bank: Optional[torch.Tensor] = None
        self._filled = [False] * self.size
        if template is not None:
            self._allocate(template)

    def _allocate(self, template: FlatTree):
        self._template = template
        self.bank = torch.zeros((self.size, template.flat.numel()), dtype=torch.float32, device=template.flat.device)

    def put(self, state, i: int, keep: int = 1, **_):
        params = state.params if hasattr(state, "params") else state
        if not isinstance(params, FlatTree):
            raise TypeError("the sample bank stores FlatTree parameters")
        if self.bank is None:
            self._allocate(params)
        if not (0 <= i < self.size):
            raise IndexError(i)
        ops.bank_put(self.bank, i, params.flat)
        self._filled[i] = True

    def get(self, i: Optional[int] = None, **_):
        if i is None:
            return [self.get(j) for j in range(self.size)]
        t = self._template
        return FlatTree(self.bank[i], t.paths, t.shapes, t.offsets)

    def pull(self, i: int, **_):
        out = self.get(i).clone()
        self._filled[i] = False
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
