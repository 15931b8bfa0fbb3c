"""SG-MCMC posterior over the last layer - the part of
fortuna/prob_model/posterior/sgmcmc/sgmcmc_posterior.py:25-59 the predictive path needs.

``sample(rng)`` draws ``idx = jax.random.choice(rng, n_samples)`` (uniform, with replacement) bit-exactly
on the device and returns a zero-copy view of that stored sample.  The batched predictive never
materialises drawn samples: it hands the index vector to the last-layer kernel, which reads
``w[idx[s]]`` directly (the reference goes through ``jax.pure_callback`` to host memory / disk per draw).
"""

import os
from typing import Optional, Sequence

import numpy as np
import torch

from .... import ops
from ....training import checkpoint as ckpt
from ....utils.flat_tree import FlatTree
from ....utils.random import RandomNumberGenerator, WithRNG
from .sgmcmc_posterior_state_repository import SGMCMCPosteriorStateRepository, posterior_state_dict


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


class _Sample:
    """(step, params) pair in the shape ``posterior_state_dict`` reads."""

    def __init__(self, step: int, params: FlatTree):
        self.step, self.params, self.opt_state, self.mutable = step, params, None, None


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

    # ------------------------------------------------------------------ checkpoints (reference layout on disk)
    STATE_NAME = "SGHMCState"

    def _which_params(self):
        """Key paths of the sampled leaves when part of the model is frozen, else None (sghmc_posterior.py:137-143)."""
        bound = getattr(self, "bound", None)
        if bound is None or len(bound.trainable) == len(list(bound.model.parameters())):
            return None
        return tuple(list(p) for p in bound.params.paths)

    def _all_params_tree(self):
        """Every parameter of the bound module as a nested numpy dict (frozen ones included)."""
        from .torch_model import _nest, param_paths

        return _nest((pth, p.detach().cpu().numpy().copy()) for pth, p in param_paths(self.bound.model).items())

    def _calib_params(self):
        """The fitted output-calibrator parameters in the reference's pytree
        (``{"output_calibrator": {"params": {"log_temp": f32[1]}}}``) or None."""
        lt = getattr(self, "log_temp", None)
        if lt is None:
            return None
        return {"output_calibrator": {"params": {"log_temp": lt.detach().cpu().numpy().astype(np.float32).reshape(1)}}}

    def _fit_repository(self, fit_config) -> SGMCMCPosteriorStateRepository:
        """The repository the sampling callback fills (sghmc_posterior.py:151-156): written through to
        ``save_checkpoint_dir/<i>`` when a directory is configured."""
        c = fit_config.checkpointer
        dp = getattr(self.bound, "dp", None)
        # data-parallel fit: every rank fills its own device bank (the replicas are bit-identical), rank 0 alone writes
        # the sample files
        ckpt_dir = c.save_checkpoint_dir if (dp is None or dp.rank == 0) else None
        self.state = SGMCMCPosteriorStateRepository(
            self.posterior_approximator.n_samples, template=self.bound.params, checkpoint_dir=ckpt_dir,
            which_params=self._which_params(), state_name=self.STATE_NAME)
        self._last_layer = None
        return self.state

    def _chain_saver(self, fit_config):
        c = fit_config.checkpointer
        if not c.save_checkpoint_dir:
            return None
        chain_dir = os.path.join(str(c.save_checkpoint_dir), "c")
        which = self._which_params()

        def save(state):
            d = posterior_state_dict(state, self.STATE_NAME, which)
            if which is not None:  # the chain keeps the whole model: load_state reads the frozen part from it
                d["params"] = ckpt.to_state_dict(self._all_params_tree())
            ckpt.save_checkpoint(chain_dir, d, step=state.step, keep=c.keep_top_n_checkpoints, overwrite=True)

        return save

    def _restore_chain_params(self, fit_config) -> None:
        """``restore_checkpoint_path`` / ``start_from_current_state`` (sgmcmc_posterior.py:82-99): the chain restarts
        from stored parameters; the sampler state is re-initialised by ``tx.init`` like the reference does when an
        optimizer is passed to ``init_from_dict`` (posterior/state.py:77-81)."""
        c = fit_config.checkpointer
        if c.restore_checkpoint_path is not None:
            d = ckpt.restore_checkpoint(os.path.join(str(c.restore_checkpoint_path), "c"))
            if d is None:
                raise ValueError(f"No checkpoint was found in `restore_checkpoint_path={c.restore_checkpoint_path}`.")
            self._load_params_tree(d["params"])
        elif c.start_from_current_state and self.state is not None and self.state.bank is not None:
            self.bound.load_flat(self.state.get(self.state.size - 1).flat)

    def _load_params_tree(self, tree) -> None:
        from .torch_model import param_paths

        for pth, p in param_paths(self.bound.model).items():
            d = tree
            try:
                for k in pth:
                    d = d[k]
            except (KeyError, TypeError):
                continue  # sample checkpoints hold the trainable leaves only
            with torch.no_grad():
                p.copy_(torch.from_numpy(np.array(d)).to(device=p.device, dtype=p.dtype).view_as(p))

    def load_state(self, checkpoint_dir) -> None:
        """sgmcmc_posterior.py:61-76: chain state from ``<dir>/c`` (frozen parameters, which leaves are sampled), the
        samples from ``<dir>/<i>`` - here straight into the device bank."""
        try:
            d = ckpt.restore_checkpoint(os.path.join(str(checkpoint_dir), "c"))
        except ValueError:
            d = None
        if d is None:
            raise ValueError(f"No checkpoint was found in `checkpoint_dir={checkpoint_dir}`.")
        enc = d.get("_encoded_which_params")
        which = None
        if enc is not None:
            which = tuple(["".join(chr(int(o)) for o in np.asarray(v)) for v in path.values()] for path in enc.values())
        bound = getattr(self, "bound", None)
        if bound is not None:
            self._load_params_tree(d["params"])
        approx = getattr(self, "posterior_approximator", None)
        size = approx.n_samples if approx is not None else self.state.size
        self.state = SGMCMCPosteriorStateRepository(
            size, template=bound.params if bound is not None else None, checkpoint_dir=checkpoint_dir,
            which_params=which, state_name=ckpt.decode_name(d),
            device=bound.device if bound is not None else self.rng._rng.device)
        self.state.load_checkpoints()
        self._last_layer = None
        cp = d.get("calib_params")
        try:
            lt = np.asarray(cp["output_calibrator"]["params"]["log_temp"], dtype=np.float32).reshape(1)
            self.log_temp = torch.from_numpy(lt.copy()).to(self.state.bank.device)
        except (TypeError, KeyError):
            self.log_temp = None

    def save_state(self, checkpoint_dir, keep_top_n_checkpoints: int = 1) -> None:
        """sgmcmc_posterior.py:78-80, plus the chain file ``load_state`` looks for."""
        which = self._which_params()
        chain = getattr(self, "chain_state", None)
        for i in range(self.state.size):
            step = self.state.steps[i] if self.state.steps[i] is not None else i
            ckpt.save_checkpoint(os.path.join(str(checkpoint_dir), str(i)),
                                 posterior_state_dict(_Sample(step, self.state.get(i)), self.STATE_NAME, which,
                                                      calib_params=self._calib_params()),
                                 step=step, keep=keep_top_n_checkpoints, overwrite=True)
        d = posterior_state_dict(chain if chain is not None else _Sample(0, self.state.get(self.state.size - 1)),
                                 self.STATE_NAME, which, calib_params=self._calib_params())
        if which is not None and getattr(self, "bound", None) is not None:
            d["params"] = ckpt.to_state_dict(self._all_params_tree())
        ckpt.save_checkpoint(os.path.join(str(checkpoint_dir), "c"), d, step=d["step"], keep=keep_top_n_checkpoints,
                             overwrite=True)

    def last_layer(self) -> LastLayerSamples:
        """The stored output layers as ``LastLayerSamples`` (kernel [n, D, C], bias [n, C]).  Two leaf conventions exist in
        a bank: a flax module's ``<last_layer_path>/kernel`` [D, C] (checkpoints written by Fortuna) and a torch
        ``nn.Linear``'s ``output_subnet/weight`` [C, D] (a fit run here) - the latter is transposed into the former."""
        if self._last_layer is None:
            paths = [tuple(p) for p in self.state._template.paths]
            if self.last_layer_path + ("kernel",) in paths:
                kernel = self.state.leaf_matrix(self.last_layer_path + ("kernel",))
                bias = self.state.leaf_matrix(self.last_layer_path + ("bias",)) if self.last_layer_path + ("bias",) in paths else None
            else:
                from .torch_model import PREFIX

                w = PREFIX + ("output_subnet", "weight")
                if w not in paths:
                    raise ValueError(f"the sample bank holds neither {self.last_layer_path + ('kernel',)} nor {w}")
                kernel = self.state.leaf_matrix(w).transpose(1, 2).contiguous()  # [n, C, D] -> [n, D, C]
                b = PREFIX + ("output_subnet", "bias")
                bias = self.state.leaf_matrix(b) if b in paths else None
            self._last_layer = LastLayerSamples(kernel, bias, self.log_temp)
        return self._last_layer
