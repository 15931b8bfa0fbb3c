"""``prob_model.predictive`` for regression - method names / signatures of
fortuna/prob_model/predictive/regression.py:22-369 and fortuna/prob_model/predictive/base.py (mean, variances, std,
log_prob, sample, the Monte-Carlo entropies).  Model outputs are ``[mu, log sigma^2]``
(fortuna/prob_output_layer/regression.py:18-74)."""
from typing import List, Optional, Union

import numpy as np
import torch

from ... import ops


class RegressionPredictive:
    def __init__(self, posterior, log_temp: Optional[torch.Tensor] = None):
        self.posterior = posterior
        self.log_temp = log_temp

    def _stored_outputs(self, x: torch.Tensor) -> torch.Tensor:
        """[n_stored, B, 2d]: one forward pass per STORED posterior sample."""
        bound, bank = self.posterior.bound, self.posterior.state
        keep = bound.params.flat.clone()
        outs = []
        with torch.no_grad():
            for i in range(bank.size):
                bound.load_flat(bank.bank[i])
                outs.append(bound.model(x))
        bound.load_flat(keep)
        return torch.stack(outs).contiguous()

    def _x(self, loader) -> torch.Tensor:
        x = loader.to_array_inputs() if hasattr(loader, "to_array_inputs") else loader
        return torch.as_tensor(np.asarray(x), dtype=torch.float32, device=self.posterior.bound.device)

    @staticmethod
    def _batches(loader, with_targets: bool = False):
        """(inputs, targets or None) batches: a loader is walked batch by batch like the reference's
        ``_loop_fun_through_inputs_loader`` / ``_loop_fun_through_data_loader`` (predictive/base.py:946-1013), targets
        staying with their inputs; a bare array is one batch."""
        if hasattr(loader, "__iter__") and not isinstance(loader, (np.ndarray, torch.Tensor)):
            for b in loader:
                if isinstance(b, tuple):
                    yield (b[0], b[1] if with_targets else None)
                else:
                    yield (b, None)
        else:
            yield (loader, None)

    def _reduce(self, loader, names, n_posterior_samples, rng, targets=None):
        """One pass over the loader: per batch one forward pass per STORED sample, the gather of the drawn ones and the
        reduction kernels; batches are concatenated along the input axis (the per-sample axis of ``ensemble_log_prob``
        stays first).  ``distribute`` is accepted by the public methods and ignored: regression outputs are a few floats
        per input, every rank evaluates everything (replicas)."""
        if rng is None:
            rng = self.posterior.rng.get()
        idx = self.posterior.sample_indices(rng, n_posterior_samples).long()
        dev = self.posterior.bound.device
        parts = {n: [] for n in names}
        with_t = targets is not None
        array_targets = None
        if with_t and not (hasattr(loader, "__iter__") and not isinstance(loader, (np.ndarray, torch.Tensor))):
            array_targets = np.asarray(targets)
        b0 = 0
        for xb, tb in self._batches(loader, with_targets=with_t):
            x = torch.as_tensor(np.asarray(xb), dtype=torch.float32, device=dev)
            z = self._stored_outputs(x).index_select(0, idx).contiguous()
            t = None
            if with_t:
                tsrc = tb if tb is not None else array_targets[b0 : b0 + x.shape[0]]
                t = torch.as_tensor(np.asarray(tsrc), dtype=torch.float32, device=dev).reshape(z.shape[1], -1)
            out = ops.bma_reduce_reg(z, names, targets=t, log_temp=self.log_temp)
            for n in names:
                parts[n].append(out[n])
            b0 += x.shape[0]
        return {n: (v[0] if len(v) == 1 else torch.cat(v, dim=1 if n == "ensemble_log_prob" else 0)) for n, v in parts.items()}

    def mean(self, inputs_loader, n_posterior_samples: int = 30, rng=None, distribute: bool = True):
        return self._reduce(inputs_loader, ("mean",), n_posterior_samples, rng)["mean"]

    def mode(self, inputs_loader, n_posterior_samples: int = 30, means=None, rng=None, distribute: bool = True):
        return means if means is not None else self.mean(inputs_loader, n_posterior_samples, rng)

    def aleatoric_variance(self, inputs_loader, n_posterior_samples: int = 30, rng=None, distribute: bool = True):
        return self._reduce(inputs_loader, ("aleatoric_variance",), n_posterior_samples, rng)["aleatoric_variance"]

    def epistemic_variance(self, inputs_loader, n_posterior_samples: int = 30, rng=None, distribute: bool = True):
        return self._reduce(inputs_loader, ("epistemic_variance",), n_posterior_samples, rng)["epistemic_variance"]

    def variance(self, inputs_loader, n_posterior_samples: int = 30, aleatoric_variances=None, epistemic_variances=None,
                 rng=None, distribute: bool = True):
        """predictive/base.py:838-898 (inherited by RegressionPredictive): ``variance = aleatoric + epistemic`` where each
        term, unless given, is estimated from ITS OWN draws - ``rng, key = split(rng)`` before the aleatoric term and again
        before the epistemic one."""
        if rng is None:
            rng = self.posterior.rng.get()
        if aleatoric_variances is None:
            ks = ops.split(rng, 2)
            rng, key = ks[0].contiguous(), ks[1].contiguous()
            aleatoric_variances = self.aleatoric_variance(inputs_loader, n_posterior_samples, rng=key)
        if epistemic_variances is None:
            ks = ops.split(rng, 2)
            rng, key = ks[0].contiguous(), ks[1].contiguous()
            epistemic_variances = self.epistemic_variance(inputs_loader, n_posterior_samples, rng=key)
        dev = self.posterior.bound.device
        a = torch.as_tensor(aleatoric_variances, dtype=torch.float32, device=dev).contiguous()
        e = torch.as_tensor(epistemic_variances, dtype=torch.float32, device=dev).contiguous()
        return ops.variance_combine(a, e)

    def std(self, inputs_loader, n_posterior_samples: int = 30, variances=None, rng=None, distribute: bool = True):
        """predictive/base.py:900-944: ``sqrt(variance)``, the variance estimated as above unless given."""
        if variances is None:
            variances = self.variance(inputs_loader, n_posterior_samples, rng=rng)
        v = torch.as_tensor(variances, dtype=torch.float32, device=self.posterior.bound.device).contiguous()
        return ops.variance_combine(v, None, want_std=True)

    def log_prob(self, data_loader, n_posterior_samples: int = 30, rng=None, distribute: bool = True):
        return self._reduce(data_loader, ("log_prob",), n_posterior_samples, rng, self._targets(data_loader))["log_prob"]

    def ensemble_log_prob(self, data_loader, n_posterior_samples: int = 30, rng=None, distribute: bool = True):
        return self._reduce(data_loader, ("ensemble_log_prob",), n_posterior_samples, rng,
                            self._targets(data_loader))["ensemble_log_prob"]

    @staticmethod
    def _targets(data_loader):
        """Marker / array for ``_reduce``: a loader's targets are read from its (inputs, targets) batches."""
        if hasattr(data_loader, "__iter__") and not isinstance(data_loader, (np.ndarray, torch.Tensor)):
            return True
        return np.asarray(data_loader.to_array_targets())

    def _mc_entropy(self, name, inputs_loader, n_posterior_samples, n_target_samples, rng):
        """regression.py:51-267: ``key1, *keys = split(rng, 1 + S)``; the S posterior draws use key1, the target samples of
        draw i use keys[i]; one kernel (fb200_reg_mc_entropies) forms the Monte-Carlo mean without materialising them."""
        if rng is None:
            rng = self.posterior.rng.get()
        ks = ops.split(rng, 1 + n_posterior_samples)
        idx = self.posterior.sample_indices(ks[0].contiguous(), n_posterior_samples)
        tkeys = ks[1:].contiguous()
        dev = self.posterior.bound.device
        parts = []
        for xb, _ in self._batches(inputs_loader):  # the same keys for every batch, like the reference's batched loop
            z = self._stored_outputs(torch.as_tensor(np.asarray(xb), dtype=torch.float32, device=dev))
            parts.append(ops.reg_mc_entropies(z, idx, tkeys, n_target_samples, (name,), log_temp=self.log_temp)[name])
        return parts[0] if len(parts) == 1 else torch.cat(parts)

    def aleatoric_entropy(self, inputs_loader, n_posterior_samples: int = 30, n_target_samples: int = 30, rng=None,
                          distribute: bool = True):
        return self._mc_entropy("aleatoric_entropy", inputs_loader, n_posterior_samples, n_target_samples, rng)

    def epistemic_entropy(self, inputs_loader, n_posterior_samples: int = 30, n_target_samples: int = 30, rng=None,
                          distribute: bool = True):
        return self._mc_entropy("epistemic_entropy", inputs_loader, n_posterior_samples, n_target_samples, rng)

    def entropy(self, inputs_loader, n_posterior_samples: int = 30, n_target_samples: int = 30, rng=None,
                distribute: bool = True):
        return self._mc_entropy("entropy", inputs_loader, n_posterior_samples, n_target_samples, rng)

    def sample(self, inputs_loader, n_target_samples: int = 1, return_aux=None, rng=None, distribute: bool = True) -> torch.Tensor:
        """[T, B, d] target samples.  Like the reference, every input BATCH of the loader is sampled with the same
        key (the noise array has the batch's shape), so results depend on the loader's batch size exactly as there."""
        if return_aux:
            raise NotImplementedError("auxiliary outputs of `sample` are outside the hot path")
        if rng is None:
            rng = self.posterior.rng.get()
        parts = []
        for xb, _ in self._batches(inputs_loader):
            zb = self._stored_outputs(self._x(xb))
            parts.append(ops.reg_sample_targets(zb, rng, n_target_samples, log_temp=self.log_temp))
        return torch.cat(parts, dim=1)

    def quantile(self, q: Union[float, List, np.ndarray], inputs_loader, n_target_samples: int = 30, rng=None,
                 distribute: bool = True) -> torch.Tensor:
        samples = self.sample(inputs_loader, n_target_samples, rng=rng)
        scalar = np.ndim(q) == 0
        qq = torch.as_tensor(np.atleast_1d(np.asarray(q, dtype=np.float32)), device=samples.device)
        out = ops.quantile_axis0(samples.contiguous(), qq)
        return out[0] if scalar else out

    def credible_interval(self, inputs_loader, n_target_samples: int = 30, error: float = 0.05,
                          interval_type: str = "two-tailed", rng=None, distribute: bool = True) -> torch.Tensor:
        supported_types = ["two-tailed", "right-tailed", "left-tailed"]
        if interval_type not in supported_types:
            raise ValueError("`type={}` not recognised. Please choose among the following supported types: {}.".format(
                interval_type, supported_types))
        q = [0.5 * error, 1 - 0.5 * error] if interval_type == "two-tailed" else (error if interval_type == "left-tailed" else 1 - error)
        qq = self.quantile(q, inputs_loader, n_target_samples, rng)
        if qq.shape[-1] != 1:
            raise ValueError("""Credibility intervals are only supported for scalar target variables.""")
        if interval_type == "two-tailed":
            return qq.squeeze(2).transpose(0, 1).contiguous()  # [B, 2]: (lower, upper) per input
        return qq
