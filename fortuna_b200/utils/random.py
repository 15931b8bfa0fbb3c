"""RNG holder and per-leaf noise trees - same names as fortuna/utils/random.py:11-74, keys on the device.

A key is jax's uint32[2] threefry key stored as an int32[2] CUDA tensor (bit pattern); every split /
draw is a library kernel that is bit-identical to jax.random at jax 0.4.30 defaults.
"""

import torch

from .. import ops
from .flat_tree import FlatTree


def generate_rng_like_tree(rng: torch.Tensor, target: FlatTree) -> torch.Tensor:
    """One key per leaf, in tree-flatten order: jax.random.split(rng, num_leaves) -> int32[L,2]."""
    return ops.split(rng, len(target.lens))


def generate_random_normal_like_tree(rng: torch.Tensor, target: FlatTree) -> FlatTree:
    """Per-leaf jax.random.normal(key_l, leaf.shape) (fortuna/utils/random.py:17-23).  The sampler kernel
    never calls this (it draws the same numbers in registers); it exists for API parity and tests."""
    keys = generate_rng_like_tree(rng, target)
    out = target.empty_like()
    for key, view in zip(keys, out.leaf_views()):
        view.copy_(ops.random_normal(key.contiguous(), view.numel()).view(view.shape))
    return out


class RandomNumberGenerator:
    """fortuna/utils/random.py:26-49: ``get()`` advances the key chain, key <- split(key)[0]."""

    def __init__(self, seed: int, device="cuda"):
        self._rng = ops.PRNGKey(seed, device)

    def get(self) -> torch.Tensor:
        self._rng = ops.split(self._rng)[0].contiguous()
        return self._rng


class WithRNG:
    @property
    def rng(self) -> RandomNumberGenerator:
        return self._rng

    @rng.setter
    def rng(self, rng: RandomNumberGenerator):
        self._rng = rng
