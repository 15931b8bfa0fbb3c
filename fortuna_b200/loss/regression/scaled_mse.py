"""``scaled_mse_fn`` - fortuna/loss/regression/scaled_mse.py:6-26: mean_b sum_k ( exp(-log var) (y - mu)^2 + log var )
for outputs = [mu | log var], through fb200_output_calib_scaled_mse_terms."""
import numpy as np
import torch

from ... import ops


def scaled_mse_fn(outputs, targets) -> torch.Tensor:
    z = torch.as_tensor(np.asarray(outputs) if not isinstance(outputs, torch.Tensor) else outputs, dtype=torch.float32).cuda().contiguous()
    y = torch.as_tensor(np.asarray(targets) if not isinstance(targets, torch.Tensor) else targets).to(device=z.device, dtype=torch.float32)
    terms = ops.output_calib_scaled_mse_terms(z, y.reshape(z.shape[0], -1).contiguous(), None)
    return (terms[0] / z.shape[0]).to(torch.float32)
