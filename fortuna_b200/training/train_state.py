"""Minimal TrainState: the part of flax's ``TrainState`` that Fortuna's trainer uses around the sampler
(fortuna/training/train_state.py:17-28, call site fortuna/training/trainer.py:170).

``apply_gradients`` is the fused path: tx.apply updates params in place inside the sampler kernel."""

from typing import Any, NamedTuple

from ..utils.flat_tree import FlatTree, as_flat_tree


class TrainState(NamedTuple):
    step: int
    params: FlatTree
    tx: Any
    opt_state: Any
    mutable: Any = None

    @classmethod
    def create(cls, params, tx, mutable=None):
        params = as_flat_tree(params)
        return cls(step=0, params=params, tx=tx, opt_state=tx.init(params), mutable=mutable)

    def apply_gradients(self, grads, joint=None, dp=None, max_grad_norm=None, **kwargs):
        """flax TrainState.apply_gradients: tx.update + optax.apply_updates + step += 1, in one launch.
        ``joint`` (see sghmc_integrator.apply) additionally fuses the log-joint gradient assembly into that launch;
        ``max_grad_norm`` the trainer's gradient clipping (fortuna/training/trainer.py:158-169)."""
        extra = {"max_grad_norm": max_grad_norm} if max_grad_norm else {}
        if dp is not None:
            params, opt_state = self.tx.apply(self.params, grads, self.opt_state, joint=joint, dp=dp, **extra)
        elif joint is not None:
            params, opt_state = self.tx.apply(self.params, grads, self.opt_state, joint=joint, **extra)
        else:
            params, opt_state = self.tx.apply(self.params, grads, self.opt_state, **extra)
        return self._replace(step=self.step + 1, params=params, opt_state=opt_state, **kwargs)
