"""Argument plumbing shared by the host-side mirrors: accept numpy arrays or torch tensors, hand contiguous CUDA
tensors of the dtype the C ABI expects to fortuna_b200.ops.  No arithmetic."""
import numpy as np
import torch


def device_of(*xs, default="cuda"):
    for x in xs:
        if isinstance(x, torch.Tensor) and x.is_cuda:
            return x.device
    return torch.device(default)


def as_device(x, dtype, device=None):
    if x is None:
        return None
    if not isinstance(x, torch.Tensor):
        x = torch.from_numpy(np.ascontiguousarray(np.asarray(x)))
    if device is None:
        device = x.device if x.is_cuda else torch.device("cuda")
    return x.to(device=device, dtype=dtype).contiguous()
