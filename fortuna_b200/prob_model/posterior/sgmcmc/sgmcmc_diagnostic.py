"""SG-MCMC diagnostics - same names, arguments and errors as
fortuna/prob_model/posterior/sgmcmc/sgmcmc_diagnostic.py:16-140.

``samples`` / ``grads`` may be what the reference takes - a list of pytrees (nested dicts of arrays), one per MCMC
sample, or a 2-D array ``[n_samples, n_params]`` - or their device forms here: a list of ``FlatTree``s, or a 2-D CUDA
tensor such as the sample bank itself (``posterior.state.bank``: no copy, rows may be strided).  Leaves are
raveled in jax tree-flatten order (sorted keys), like ``jax.flatten_util.ravel_pytree``.

The arithmetic runs in fb200_ksd_imq / fb200_ess (csrc/diagnostic.cu); this module only lays the inputs out.
"""
from typing import Any, List, Optional

import numpy as np
import torch

from .... import ops
from ....utils.flat_tree import FlatTree, _flatten


def _ravel_one(x) -> np.ndarray:
    if isinstance(x, dict):
        return np.concatenate([np.asarray(v, dtype=np.float32).reshape(-1) for _, v in _flatten(x)])
    return np.asarray(x, dtype=np.float32).reshape(-1)


def _as_matrix(samples: Any, device) -> torch.Tensor:
    """[n_samples, n_params] float32 CUDA matrix (== ravel_pytree(samples)[0].reshape(len(samples), -1))."""
    if isinstance(samples, torch.Tensor):
        m = samples if samples.dim() == 2 else samples.reshape(samples.shape[0], -1)
        return m.to(device=device, dtype=torch.float32)
    if isinstance(samples, (list, tuple)) and len(samples) and isinstance(samples[0], FlatTree):
        # padding between leaves (16-byte alignment) must not count as parameters: gather the leaf ranges
        t = samples[0]
        cols = torch.cat([torch.arange(o, o + n) for o, n in zip(t.offsets, t.lens)]).to(samples[0].flat.device)
        return torch.stack([s.flat for s in samples]).index_select(1, cols).contiguous()
    if isinstance(samples, (list, tuple)):
        m = np.stack([_ravel_one(s) for s in samples])
    else:
        a = np.asarray(samples, dtype=np.float32)
        m = a.reshape(a.shape[0], -1)
    return torch.from_numpy(np.ascontiguousarray(m, dtype=np.float32)).to(device)


def _device_of(x):
    if isinstance(x, torch.Tensor) and x.is_cuda:
        return x.device
    if isinstance(x, (list, tuple)) and len(x) and isinstance(x[0], FlatTree):
        return x[0].flat.device
    return torch.device("cuda")


def kernel_stein_discrepancy_imq(samples: List, grads: List, c: float = 1.0, beta: float = -0.5) -> torch.Tensor:
    """Kernel Stein discrepancy with the inverse multiquadric kernel (Gorham & Mackey 2017); a 0-d device tensor."""
    if not c > 0:
        raise ValueError("`c` should be > 0.")
    if not beta < 0:
        raise ValueError("`beta` should be < 0.")
    dev = _device_of(samples)
    return ops.ksd_imq(_as_matrix(samples, dev), _as_matrix(grads, dev), c, beta).reshape(())


def effective_sample_size(samples: List, filter_threshold: Optional[float] = 0.0):
    """Parameter-wise effective sample size of a chain.  Returns a tensor ``[n_params]`` for array input, a nested
    dict of tensors shaped like one sample for pytree input (the reference's ``unravel_fn(ess)``)."""
    dev = _device_of(samples)
    out = ops.ess(_as_matrix(samples, dev), filter_threshold)
    first = samples[0] if isinstance(samples, (list, tuple)) and len(samples) else None
    if isinstance(first, FlatTree):
        off, res = 0, {}
        for path, n, shp in zip(first.paths, first.lens, first.shapes):
            d = res
            for k in path[:-1]:
                d = d.setdefault(k, {})
            d[path[-1]] = out[off : off + n].view(shp)
            off += n
        return res
    if isinstance(first, dict):
        res, off = {}, 0
        for path, leaf in _flatten(first):
            n = int(np.asarray(leaf).size)
            d = res
            for k in path[:-1]:
                d = d.setdefault(k, {})
            d[path[-1]] = out[off : off + n].view(tuple(np.asarray(leaf).shape))
            off += n
        return res
    return out
