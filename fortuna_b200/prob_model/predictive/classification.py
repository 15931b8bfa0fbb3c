"""``prob_model.predictive`` for classification - the method names and signatures of
fortuna/prob_model/predictive/base.py:39-1022 and fortuna/prob_model/predictive/classification.py:22-483.

Every statistic is: (1) draw S posterior-sample indices exactly as the reference does
(``keys = split(rng, S)``; ``idx_s = choice(keys[s], n_samples)``, base.py:632 + sgmcmc_posterior.py:51) on the
device; (2) calibrated logits ``[S, B, C]`` - for a last-layer posterior the features are computed ONCE and the
S-batched contraction (K2) reads ``W[idx[s]]`` from the sample bank, otherwise the backbone is evaluated once per
STORED sample (not once per draw) and gathered; (3) one single-pass reduction kernel (K3/K4).  The reference runs
S full forward passes per statistic and fetches each drawn sample through a host callback."""
from typing import List, Optional

import numpy as np
import torch

from ... import ops
from ..posterior.sgmcmc.torch_model import PREFIX


class ClassificationPredictive:
    def __init__(self, posterior, log_temp: Optional[torch.Tensor] = None):
        self.posterior = posterior
        self.log_temp = log_temp
        self._stored_logits_cache = None  # (inputs id, [n, B, C])

    # ------------------------------------------------------------------ logits of the drawn samples
    def _inputs(self, inputs_loader) -> torch.Tensor:
        x = inputs_loader.to_array_inputs() if hasattr(inputs_loader, "to_array_inputs") else inputs_loader
        return torch.as_tensor(np.asarray(x), dtype=torch.float32, device=self.posterior.bound.device)

    def _indices(self, n_posterior_samples: int, rng) -> torch.Tensor:
        if rng is None:
            rng = self.posterior.rng.get()
        return self.posterior.sample_indices(rng, n_posterior_samples)

    def sample_calibrated_outputs(self, inputs_loader, n_output_samples: int = 1, rng=None, distribute: bool = True) -> torch.Tensor:
        """[S, B, C] calibrated logits of S drawn posterior samples (base.py:442-499)."""
        x = self._inputs(inputs_loader)
        idx = self._indices(n_output_samples, rng)
        bound, bank = self.posterior.bound, self.posterior.state
        if bound.is_last_layer_only():
            with torch.no_grad():
                feats = bound.model.dfe_subnet(x).contiguous()
            w = bank.leaf_matrix(PREFIX + ("output_subnet", "weight"))  # [n, C, D]: torch Linear rows are K-major
            has_bias = bound.model.output_subnet.bias is not None
            b = bank.leaf_matrix(PREFIX + ("output_subnet", "bias")) if has_bias else None
            return ops.last_layer_logits(feats, w, b, self.log_temp, w_layout=ops.W_SCD, sample_idx=idx, path=1)
        # full-network posterior: one forward pass per STORED sample, then gather the draws (with replacement)
        key = (x.data_ptr(), tuple(x.shape), bank.n_filled)
        if self._stored_logits_cache is None or self._stored_logits_cache[0] != key:
            keep = bound.params.flat.clone()
            outs = []
            with torch.no_grad():
                for i in range(bank.size):
                    bound.load_flat(bank.bank[i])
                    outs.append(bound.model(x))
            bound.load_flat(keep)
            self._stored_logits_cache = (key, torch.stack(outs).contiguous())
        stored = self._stored_logits_cache[1]
        z = stored.index_select(0, idx.long()).contiguous()
        if self.log_temp is not None:  # ClassificationTemperatureScaler: logits * exp(-log_temp), in place on the copy
            ops.softmax_rows_(z, ops.EPI_LOGITS, log_temp=self.log_temp)
        return z

    def _reduce(self, inputs_loader, names, n_posterior_samples, rng, targets=None):
        z = self.sample_calibrated_outputs(inputs_loader, n_posterior_samples, rng)
        t = None
        if targets is not None:
            t = torch.as_tensor(np.asarray(targets), dtype=torch.int32, device=z.device)
        return ops.bma_reduce_cls(z, names, targets=t)

    # ------------------------------------------------------------------ statistics (reference method names)
    def mean(self, inputs_loader, n_posterior_samples: int = 30, rng=None, distribute: bool = True):
        return self._reduce(inputs_loader, ("mean",), n_posterior_samples, rng)["mean"]

    def mode(self, inputs_loader, n_posterior_samples: int = 30, means=None, rng=None, distribute: bool = True):
        if means is not None:
            m = torch.as_tensor(means)
            return torch.argmax(m, dim=-1)  # given means: arg-max lookup only (classification.py:86)
        return self._reduce(inputs_loader, ("mode",), n_posterior_samples, rng)["mode"]

    def aleatoric_variance(self, inputs_loader, n_posterior_samples: int = 30, rng=None, distribute: bool = True):
        return self._reduce(inputs_loader, ("aleatoric_variance",), n_posterior_samples, rng)["aleatoric_variance"]

    def epistemic_variance(self, inputs_loader, n_posterior_samples: int = 30, rng=None, distribute: bool = True):
        return self._reduce(inputs_loader, ("epistemic_variance",), n_posterior_samples, rng)["epistemic_variance"]

    def variance(self, inputs_loader, n_posterior_samples: int = 30, rng=None, distribute: bool = True):
        """Same sample set for both terms (one pass).  NB the reference draws two different sample sets
        (base.py:883-897); pass ``aleatoric_variance`` / ``epistemic_variance`` separately to reproduce that."""
        return self._reduce(inputs_loader, ("variance",), n_posterior_samples, rng)["variance"]

    def std(self, inputs_loader, n_posterior_samples: int = 30, rng=None, distribute: bool = True):
        return self._reduce(inputs_loader, ("std",), n_posterior_samples, rng)["std"]

    def entropy(self, inputs_loader, n_posterior_samples: int = 30, rng=None, distribute: bool = True):
        return self._reduce(inputs_loader, ("entropy",), n_posterior_samples, rng)["entropy"]

    def aleatoric_entropy(self, inputs_loader, n_posterior_samples: int = 30, rng=None, distribute: bool = True):
        return self._reduce(inputs_loader, ("aleatoric_entropy",), n_posterior_samples, rng)["aleatoric_entropy"]

    def epistemic_entropy(self, inputs_loader, n_posterior_samples: int = 30, rng=None, distribute: bool = True):
        return self._reduce(inputs_loader, ("epistemic_entropy",), n_posterior_samples, rng)["epistemic_entropy"]

    def log_prob(self, data_loader, n_posterior_samples: int = 30, rng=None, distribute: bool = True):
        return self._reduce(data_loader, ("log_prob",), n_posterior_samples, rng, data_loader.to_array_targets())["log_prob"]

    def ensemble_log_prob(self, data_loader, n_posterior_samples: int = 30, rng=None, distribute: bool = True):
        return self._reduce(data_loader, ("ensemble_log_prob",), n_posterior_samples, rng,
                            data_loader.to_array_targets())["ensemble_log_prob"]

    def statistics(self, data_or_inputs_loader, names, n_posterior_samples: int = 30, rng=None):
        """All requested statistics from ONE sample set and ONE pass (no reference counterpart: the reference
        recomputes the S forward passes per statistic)."""
        targets = data_or_inputs_loader.to_array_targets() if hasattr(data_or_inputs_loader, "to_array_targets") else None
        return self._reduce(data_or_inputs_loader, tuple(names), n_posterior_samples, rng, targets)

    def credible_set(self, inputs_loader, n_posterior_samples: int = 30, error: float = 0.05, rng=None,
                     distribute: bool = True) -> List[List[int]]:
        mean = self.mean(inputs_loader, n_posterior_samples, rng)
        labels, size = ops.credible_sets(mean, error)
        labels, size = labels.cpu().numpy(), size.cpu().numpy()
        return [labels[i, : size[i]].tolist() for i in range(labels.shape[0])]

    def _stored_logits(self, x: torch.Tensor) -> torch.Tensor:
        """[n_stored, B, C] float32 logits of every stored posterior sample for one input batch."""
        bound, bank = self.posterior.bound, self.posterior.state
        if bound.is_last_layer_only():
            with torch.no_grad():
                feats = bound.model.dfe_subnet(x).contiguous()
            w = bank.leaf_matrix(PREFIX + ("output_subnet", "weight"))
            has_bias = bound.model.output_subnet.bias is not None
            bias = bank.leaf_matrix(PREFIX + ("output_subnet", "bias")) if has_bias else None
            return ops.last_layer_logits(feats, w, bias, None, w_layout=ops.W_SCD, path=1)
        keep = bound.params.flat.clone()
        outs = []
        with torch.no_grad():
            for i in range(bank.size):
                bound.load_flat(bank.bank[i])
                outs.append(bound.model(x))
        bound.load_flat(keep)
        return torch.stack(outs).float().contiguous()

    def sample(self, inputs_loader, n_target_samples: int = 1, return_aux=None, rng=None, distribute: bool = True) -> torch.Tensor:
        """int32 [T, B] class samples (base.py:318-440).  Like the reference, every input BATCH of the loader is sampled
        with the same key, so results depend on the loader's batch size exactly as there."""
        if return_aux:
            raise NotImplementedError("auxiliary outputs of `sample` are outside the hot path")
        if rng is None:
            rng = self.posterior.rng.get()
        parts = []
        batches = inputs_loader if hasattr(inputs_loader, "__iter__") and not isinstance(inputs_loader, (np.ndarray, torch.Tensor)) else [inputs_loader]
        for xb in batches:
            zb = self._stored_logits(self._inputs(xb))
            parts.append(ops.cls_sample_targets(zb, rng, n_target_samples, log_temp=self.log_temp))
        return torch.cat(parts, dim=1)

    def conformal_set(self, train_data_loader, test_inputs_loader, n_posterior_samples: int = 30, error: float = 0.05,
                      rng=None, distribute: bool = True, return_ess: bool = False):
        """Conformal Bayes sets by importance sampling (classification.py:485-626): the SAME posterior draws score the
        training data (``ensemble_log_prob`` [S, n_train], from the single-pass reduction kernel) and every class of
        every test input (log-softmax [S, n_test, C] - the reference evaluates the test grid class by class through
        ``ensemble_log_prob``, C redundant forward passes); the [n_test*C, S] x [S, n_train] contraction and its rank
        count run in fb200_conformal_bayes_sets without materialising the product.  Returns a list of label lists
        (ascending labels), plus the importance-weight ESS [n_test, S-free C, 1]-shaped like the reference's when asked."""
        if rng is None:
            rng = self.posterior.rng.get()
        z_train = self.sample_calibrated_outputs(train_data_loader.to_inputs_loader(), n_posterior_samples, rng)
        t = torch.as_tensor(np.asarray(train_data_loader.to_array_targets()), dtype=torch.int32, device=z_train.device)
        ell_train = ops.bma_reduce_cls(z_train, ("ensemble_log_prob",), targets=t)["ensemble_log_prob"]
        del z_train
        z_test = self.sample_calibrated_outputs(test_inputs_loader, n_posterior_samples, rng)  # same rng: same draws
        lsm_test = ops.softmax_rows_(z_test.float().contiguous(), ops.EPI_LOG_SOFTMAX)
        out = ops.conformal_bayes_sets(ell_train, lsm_test, error, want_ess=return_ess)
        region = out["region"].cpu().numpy().astype(bool)
        sets = [np.nonzero(row)[0].tolist() for row in region]
        if return_ess:
            return sets, out["ess"].unsqueeze(-1)  # lax.map over test points of [C, 1] arrays (:601-621)
        return sets
