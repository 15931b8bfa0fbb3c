"""Sampler preconditioners - same factories as
fortuna/prob_model/posterior/sgmcmc/sgmcmc_preconditioner.py:44-138.

In the reference a ``Preconditioner`` is a bundle of five tree_map closures that XLA fuses into the
update.  Here the arithmetic lives inside the fused sampler kernel (fortuna_b200/csrc/sgmcmc.cu), so a
``Preconditioner`` is a description of it (kind + constants) plus ``init``, which allocates the state in
the reference's layout (``RMSPropPreconditionerState(grad_moment_estimates=<params-shaped>)`` or the
empty ``IdentityPreconditionerState``).
"""

from typing import NamedTuple, Optional

from ....utils.flat_tree import FlatTree


class IdentityPreconditionerState(NamedTuple):
    pass


class RMSPropPreconditionerState(NamedTuple):
    grad_moment_estimates: FlatTree


class Preconditioner(NamedTuple):
    kind: str  # "identity" | "rmsprop"
    running_average_factor: float = 0.99
    eps: float = 1.0e-7

    def init(self, params: FlatTree):
        if self.kind == "identity":
            return IdentityPreconditionerState()
        return RMSPropPreconditionerState(grad_moment_estimates=params.zeros_like())

    @property
    def is_rmsprop(self) -> bool:
        return self.kind == "rmsprop"


def rmsprop_preconditioner(running_average_factor: float = 0.99, eps: float = 1.0e-7) -> Preconditioner:
    return Preconditioner("rmsprop", float(running_average_factor), float(eps))


def identity_preconditioner() -> Preconditioner:
    return Preconditioner("identity")


def precond_buffer(state) -> Optional["FlatTree"]:
    return state.grad_moment_estimates if isinstance(state, RMSPropPreconditionerState) else None
