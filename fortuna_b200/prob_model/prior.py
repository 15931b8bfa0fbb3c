"""Isotropic Gaussian prior (fortuna/prob_model/prior/gaussian.py:12-40): log N(theta; 0, exp(log_var) I).
Its gradient is part of the user's loss, computed by autograd together with the likelihood - it is not part of the
sampler step."""
import math

import torch


class IsotropicGaussianPrior:
    def __init__(self, log_var: float = 0.0):
        self.log_var = float(log_var)

    def log_joint_prob(self, flat_params: torch.Tensor) -> torch.Tensor:
        n = flat_params.numel()
        return -0.5 * (flat_params.square().sum() * math.exp(-self.log_var) + n * (math.log(2 * math.pi) + self.log_var))
