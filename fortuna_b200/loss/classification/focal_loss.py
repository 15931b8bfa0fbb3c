"""``focal_loss_fn`` - fortuna/loss/classification/focal_loss.py:10-36: -mean_b (1 - p_b)^gamma log p_b with
p_b = softmax(outputs_b)[y_b].  The value comes from one kernel pass (fb200_output_calib_focal_loss_terms); the output
calibration loop recognises this function (or a ``functools.partial`` of it fixing ``gamma``) and asks the same kernel
for the derivative in the temperature as well."""
import numpy as np
import torch

from ... import ops


def focal_loss_fn(outputs, targets, gamma: float = 2.0) -> torch.Tensor:
    z = torch.as_tensor(np.asarray(outputs) if not isinstance(outputs, torch.Tensor) else outputs, dtype=torch.float32).cuda().contiguous()
    y = torch.as_tensor(np.asarray(targets) if not isinstance(targets, torch.Tensor) else targets).to(device=z.device, dtype=torch.int32).contiguous()
    terms = ops.output_calib_focal_loss_terms(z, y, None, gamma)
    return (terms[0] / z.shape[0]).to(torch.float32)
