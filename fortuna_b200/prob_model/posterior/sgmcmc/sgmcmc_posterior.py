"""SG-MCMC posterior over the last layer - the part of
fortuna/prob_model/posterior/sgmcmc/sgmcmc_posterior.py:25-59 the predictive path needs.

``sample(rng)`` draws ``idx = jax.random.choice(rng, n_samples)`` (uniform, with replacement) bit-exactly
on the device and returns a zero-copy view of that stored sample.  The batched predictive never
materialises drawn samples: it hands the index vector to the last-layer kernel, which reads
``w[idx[s]]`` directly (the reference goes through ``jax.pure_callback`` to host memory / disk per draw).
"""

from typing import Optional, Sequence

import torch

from .... import ops
from ....utils.random import RandomNumberGenerator, WithRNG
from .sgmcmc_posterior_state_repository import SGMCMCPosteriorStateRepository


class LastLayerSamples:
    """Dense views of the output layer across the sample bank, laid out for the contraction kernels:
    kernel [n, D, C] float32 (flax ``nn.Dense`` layout), bias [n, C] float32 and, built once on demand, the
    K-major bf16 copy [n, C, D] the tcgen05 path consumes."""

    def __init__(self, kernel: torch.Tensor, bias: Optional[torch.Tensor], log_temp: Optional[torch.Tensor] = None):
        if kernel.dim() != 3:
            raise ValueError("kernel must be [n_samples, D, C]")
        self.kernel = kernel.contiguous()
        self.bias = bias.contiguous() if bias is not None else None
        self.log_temp = log_temp
        self._kernel_t_bf16 = None

    @property
    def n_samples(self) -> int:
        return self.kernel.shape[0]

    @property
    def kernel_t_bf16(self) -> torch.Tensor:
        if self._kernel_t_bf16 is None:
            self._kernel_t_bf16 = ops.transpose_w_bf16(self.kernel)
        return self._kernel_t_bf16


class SGMCMCPosterior(WithRNG):
    def __init__(self, state_repository: SGMCMCPosteriorStateRepository, rng: RandomNumberGenerator,
                 last_layer_path: Sequence[str] = ("model", "params", "output_subnet", "last"),
                 log_temp: Optional[torch.Tensor] = None):
        self.state = state_repository
        self.rng = rng
        self.last_layer_path = tuple(last_layer_path)
        self.log_temp = log_temp
        self._last_layer = None

    @property
    def n_samples(self) -> int:
        return self.state.size

    def sample_index(self, rng: Optional[torch.Tensor] = None) -> torch.Tensor:
        """int32[1] device tensor: jax.random.choice(rng, n_samples)."""
        if rng is None:
            rng = self.rng.get()
        # choice(rng, n) == the single draw of sample_indices with the key itself (no split over draws)
        return ops.sample_indices_from_keys(rng.view(1, 2), self.n_samples)

    def sample(self, rng: Optional[torch.Tensor] = None, **kwargs):
        """A stored sample as a FlatTree (zero-copy row view).  Reads the drawn index back to the host; the
        batched predictive path uses ``sample_indices`` instead and never synchronises."""
        idx = int(self.sample_index(rng).item())
        return self.state.get(idx)

    def sample_indices(self, rng: torch.Tensor, n_posterior_samples: int) -> torch.Tensor:
        """idx[s] = choice(split(rng, S)[s], n_samples) for s < S - exactly the draws of
        fortuna/prob_model/predictive/base.py:632 + sgmcmc_posterior.py:51, one kernel, no host round trip."""
        return ops.sample_indices(rng, n_posterior_samples, self.n_samples)

    def last_layer(self) -> LastLayerSamples:
        if self._last_layer is None:
            kernel = self.state.leaf_matrix(self.last_layer_path + ("kernel",))
            try:
                bias = self.state.leaf_matrix(self.last_layer_path + ("bias",))
            except ValueError:
                bias = None
            self._last_layer = LastLayerSamples(kernel, bias, self.log_temp)
        return self._last_layer
