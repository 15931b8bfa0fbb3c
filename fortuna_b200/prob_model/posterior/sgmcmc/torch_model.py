"""Glue between a torch ``nn.Module`` (the backbone; model definitions are outside the hot path) and the flat
parameter layout the sampler kernel works on.

Fortuna's params pytree is ``{"model": {"params": {<module path>: {...}}}}``.  Here the
module's ``named_parameters()`` give the paths (``"output_subnet.weight"`` -> ``("model", "params",
"output_subnet", "weight")``), ``freeze_fun(path, value) -> "trainable" | "frozen"`` selects the trainable sub-tree
exactly like ``FitOptimizer.freeze_fun`` (fortuna/utils/freeze.py:84-148), and every trainable parameter of the
module is re-pointed at a VIEW of the flat vector - so the sampler kernel's in-place update is what the next
forward pass sees, and autograd accumulates straight into a flat gradient vector (zero copies per step)."""
from typing import Callable, Dict, Optional, Tuple

import torch

from ....utils.flat_tree import FlatTree

PREFIX = ("model", "params")


def param_paths(model: torch.nn.Module) -> Dict[Tuple[str, ...], torch.nn.Parameter]:
    return {PREFIX + tuple(name.split(".")): p for name, p in model.named_parameters()}


def _nest(items):
    root = {}
    for path, v in items:
        d = root
        for k in path[:-1]:
            d = d.setdefault(k, {})
        d[path[-1]] = v
    return root


class BoundModel:
    """model + (params, grads) FlatTrees over its trainable parameters."""

    def __init__(self, model: torch.nn.Module, device, freeze_fun: Optional[Callable] = None,
                 flat_alloc: Optional[Callable] = None):
        """``flat_alloc(n, device)`` allocates the two flat vectors (e.g. in symmetric memory for the data-parallel
        step); default: ordinary device memory."""
        self.model = model.to(device=device, dtype=torch.float32)
        self.device = device
        all_p = param_paths(self.model)
        self.trainable = {pth: p for pth, p in all_p.items()
                          if freeze_fun is None or freeze_fun(pth, p) == "trainable"}
        if not self.trainable:
            raise ValueError("`freeze_fun` froze every parameter: nothing to sample.")
        for pth, p in all_p.items():
            p.requires_grad_(pth in self.trainable)
        self.params = FlatTree.from_tree(_nest((pth, p.detach()) for pth, p in self.trainable.items()), device=device)
        self.grads = self.params.zeros_like()
        if flat_alloc is not None:
            for name in ("params", "grads"):
                old = getattr(self, name)
                flat = flat_alloc(old.flat.numel(), device)
                flat.copy_(old.flat)
                setattr(self, name, FlatTree(flat, old.paths, old.shapes, old.offsets))
        self.dp = None  # PeerReplicas when the fit runs data-parallel
        by_path = dict(zip(self.params.paths, zip(self.params.leaf_views(), self.grads.leaf_views())))
        for pth, p in self.trainable.items():
            pv, gv = by_path[pth]
            p.data = pv   # the module now reads the flat vector
            p.grad = gv   # and autograd accumulates into the flat gradient

    def zero_grad(self):
        self.grads.flat.zero_()

    def load_flat(self, flat_row: torch.Tensor):
        """Make the module evaluate a stored posterior sample (a row of the sample bank)."""
        self.params.flat.copy_(flat_row)

    def is_last_layer_only(self) -> bool:
        """True when exactly ``output_subnet.{weight,bias}`` are trainable and the module exposes the reference's
        ``dfe_subnet`` / ``output_subnet`` split (fortuna/model/mlp.py:16-64) - the S-batched contraction applies."""
        m = self.model
        if not (hasattr(m, "dfe_subnet") and hasattr(m, "output_subnet") and isinstance(m.output_subnet, torch.nn.Linear)):
            return False
        want = {PREFIX + ("output_subnet", "weight")}
        if m.output_subnet.bias is not None:
            want.add(PREFIX + ("output_subnet", "bias"))
        return set(self.trainable) == want


def bind_for_fit(model: torch.nn.Module, device, freeze_fun: Optional[Callable] = None, processor=None) -> BoundModel:
    """BoundModel for a fit.  With ``processor.devices == -1`` (the reference's default, fit_config/processor.py) in a
    multi-rank NCCL job the flat vectors live in symmetric memory, every replica starts from rank 0's parameters, and
    ``bound.dp`` carries the peer mappings the data-parallel step kernel needs."""
    import torch.distributed as dist

    devices = getattr(processor, "devices", -1)
    if (devices == -1 and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
            and dist.get_backend() == "nccl"):
        from .dp_step import PeerReplicas, symmetric_empty

        bound = BoundModel(model, device, freeze_fun=freeze_fun, flat_alloc=symmetric_empty)
        dist.broadcast(bound.params.flat, 0)
        bound.dp = PeerReplicas(bound.params.flat, bound.grads.flat, dist.get_world_size(), dist.get_rank())
        return bound
    return BoundModel(model, device, freeze_fun=freeze_fun)
