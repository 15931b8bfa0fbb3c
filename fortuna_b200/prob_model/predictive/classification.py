"""``prob_model.predictive`` for classification - the method names and signatures of
fortuna/prob_model/predictive/base.py:39-1022 and fortuna/prob_model/predictive/classification.py:22-483.

Every statistic is: (1) draw S posterior-sample indices exactly as the reference does
(``keys = split(rng, S)``; ``idx_s = choice(keys[s], n_samples)``, base.py:632 + sgmcmc_posterior.py:51) on the
device; (2) walk the loader batch by batch like ``_loop_fun_through_inputs_loader`` (base.py:946-1013) - pinned staging,
H2D copies overlapped with the kernels of the previous batch (``streaming.stream_cls_reduce``); per batch the calibrated
logits ``[S, Bc, C]`` come from the features computed ONCE and the S-batched contraction K2 reading ``W[idx[s]]`` from the
sample bank (last-layer posterior), or from one backbone pass per STORED sample (not per draw) and a gather; (3) one
single-pass reduction kernel K3/K4 per batch writes the batch's rows of the result.  ``distribute=True`` (the reference's
default, there a ``pmap`` over the batch: base.py:364-366, 549-568) in a multi-rank NCCL job shards the S draws over the
ranks and exchanges partial sums over NVLink; every rank returns the whole arrays.  The reference runs S full forward
passes per statistic and fetches each drawn sample through a host callback."""
from typing import List, Optional

import numpy as np
import torch

from ... import ops
from ..posterior.sgmcmc.torch_model import PREFIX
from . import streaming


def _is_loader(x) -> bool:
    return hasattr(x, "__iter__") and not isinstance(x, (np.ndarray, torch.Tensor))


class ClassificationPredictive:
    def __init__(self, posterior, log_temp: Optional[torch.Tensor] = None, last_layer_dtype=torch.float32,
                 logits_dtype=torch.float32):
        """``last_layer_dtype=torch.bfloat16`` runs the last-layer contraction on the tcgen05 path (bf16 operands, fp32
        accumulation: north_star tolerance rtol 2e-2 on the logits) where the shape allows it; ``logits_dtype`` is what
        K2 hands to K3 (bf16 halves the traffic between them).  Defaults reproduce the reference's float32."""
        self.posterior = posterior
        self.log_temp = log_temp
        self.last_layer_dtype = last_layer_dtype
        self.logits_dtype = logits_dtype
        self._staging = {}
        self._shard = None
        self._bank_version = None
        self._wt = None  # (bank version, dtype) -> K-major copy of the stored last layers

    # ------------------------------------------------------------------ inputs
    def _device(self):
        return self.posterior.bound.device

    def _inputs(self, inputs_loader) -> torch.Tensor:
        x = inputs_loader.to_array_inputs() if hasattr(inputs_loader, "to_array_inputs") else inputs_loader
        return torch.as_tensor(np.asarray(x), dtype=torch.float32, device=self._device())

    @staticmethod
    def _input_batches(loader, with_targets: bool = False):
        """Batches of inputs (numpy) of an InputsLoader / DataLoader, or the array itself as one batch.  With
        ``with_targets`` a DataLoader's (inputs, targets) pairs are passed on whole, like the reference's
        ``_loop_fun_through_data_loader`` (base.py:977-1013): targets stay with their inputs whatever order the loader
        yields them in."""
        if _is_loader(loader):
            for b in loader:
                if isinstance(b, tuple):
                    yield b if with_targets else b[0]
                else:
                    yield b
        else:
            yield loader if isinstance(loader, torch.Tensor) else np.asarray(loader)

    @staticmethod
    def _n_rows(loader) -> int:
        if hasattr(loader, "size"):
            return int(loader.size)
        if hasattr(loader, "to_array_inputs"):
            return len(loader.to_array_inputs())
        return len(loader)

    def _indices(self, n_posterior_samples: int, rng) -> torch.Tensor:
        if rng is None:
            rng = self.posterior.rng.get()
        return self.posterior.sample_indices(rng, n_posterior_samples)

    # ------------------------------------------------------------------ logits of the drawn samples, one batch
    def _last_layer_weights(self):
        """Stored output layers as K2 wants them: W [n, C, D] (torch ``nn.Linear`` rows are K-major already) in the
        contraction dtype, bias [n, C] float32; rebuilt when the bank changes."""
        bound, bank = self.posterior.bound, self.posterior.state
        version = (id(bank), bank.n_filled, tuple(bank.steps), self.last_layer_dtype)
        if self._wt is None or self._wt[0] != version:
            w = bank.leaf_matrix(PREFIX + ("output_subnet", "weight"))  # [n, C, D] float32
            if self.last_layer_dtype == torch.bfloat16:
                n, Cn, D = w.shape  # fb200_transpose_w_bf16 on [n C, D, 1] is a pure cast to the same K-major layout
                w = ops.transpose_w_bf16(w.view(n * Cn, D, 1)).view(n, Cn, D)
            has_bias = bound.model.output_subnet.bias is not None
            b = bank.leaf_matrix(PREFIX + ("output_subnet", "bias")) if has_bias else None
            self._wt = (version, w, b)
        return self._wt[1], self._wt[2]

    def _logits_fn(self, idx: Optional[torch.Tensor], calibrated: bool = True):
        """x_dev [Bc, ...] -> [S, Bc, C] logits of the draws `idx` (None: every stored sample once), calibrated by the
        fitted temperature unless ``calibrated=False``."""
        bound, bank = self.posterior.bound, self.posterior.state
        log_temp = self.log_temp if calibrated else None
        if bound.is_last_layer_only():
            w, b = self._last_layer_weights()
            tc = self.last_layer_dtype == torch.bfloat16

            def last_layer(x):
                with torch.no_grad():
                    feats = bound.model.dfe_subnet(x)
                feats = feats.to(torch.bfloat16).contiguous() if tc else feats.float().contiguous()
                return ops.last_layer_logits(feats, w, b, log_temp, out_dtype=self.logits_dtype if tc else torch.float32,
                                             w_layout=ops.W_SCD, sample_idx=idx, path=0 if tc else 1)

            return last_layer

        # full-network posterior: one forward pass per STORED sample, then gather the draws (with replacement)
        def full_network(x):
            keep = bound.params.flat.clone()
            outs = []
            with torch.no_grad():
                for i in range(bank.size):
                    bound.load_flat(bank.bank[i])
                    outs.append(bound.model(x).float())
            bound.load_flat(keep)
            stored = torch.stack(outs)
            z = stored if idx is None else stored.index_select(0, idx.long())
            z = z.contiguous()
            if log_temp is not None:  # ClassificationTemperatureScaler: logits * exp(-log_temp), in place on the copy
                ops.softmax_rows_(z, ops.EPI_LOGITS, log_temp=log_temp)
            return z

        return full_network

    def sample_calibrated_outputs(self, inputs_loader, n_output_samples: int = 1, rng=None, distribute: bool = True) -> torch.Tensor:
        """[S, B, C] calibrated logits of S drawn posterior samples (base.py:442-499), the loader walked batch by batch."""
        idx = self._indices(n_output_samples, rng)
        fn = self._logits_fn(idx)
        dev = self._device()
        parts = [fn(torch.as_tensor(xb, dtype=torch.float32, device=dev)) for xb in self._input_batches(inputs_loader)]
        return parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)

    # ------------------------------------------------------------------ the streamed reduction
    def _reduce(self, loader, names, n_posterior_samples, rng, targets=None, distribute: bool = True, sets=None,
                to_host=False, sharded_io: bool = False, ece_total=None):
        idx = self._indices(n_posterior_samples, rng)
        world, rank = streaming.dist_world(distribute)
        S = int(n_posterior_samples)
        shard = None
        if world > 1 and S % world == 0:  # each rank evaluates S / world of the draws (same draws as one GPU would)
            if self._shard is None or self._shard.world != world:
                self._shard = streaming.ShardContext(world, rank, self._device())
            shard = self._shard
            idx = idx[rank * (S // world) : (rank + 1) * (S // world)].contiguous()
        return streaming.stream_cls_reduce(self._input_batches(loader, with_targets=targets is not None),
                                           self._logits_fn(idx), names, self._n_rows(loader), S, self._device(),
                                           targets=targets if not _is_loader(loader) else None, shard=shard, sets=sets,
                                           staging=self._staging, to_host=to_host, sharded_io=sharded_io,
                                           ece_total=ece_total)

    @staticmethod
    def _targets(data_loader):
        """Marker / array for ``_reduce``: a loader's targets are read from its (inputs, targets) batches."""
        return True if _is_loader(data_loader) else np.asarray(data_loader.to_array_targets())

    # ------------------------------------------------------------------ statistics (reference method names)
    def mean(self, inputs_loader, n_posterior_samples: int = 30, rng=None, distribute: bool = True):
        return self._reduce(inputs_loader, ("mean",), n_posterior_samples, rng, distribute=distribute)["mean"]

    def mode(self, inputs_loader, n_posterior_samples: int = 30, means=None, rng=None, distribute: bool = True):
        if means is not None:
            m = torch.as_tensor(means)
            return torch.argmax(m, dim=-1)  # given means: arg-max lookup only (classification.py:86)
        return self._reduce(inputs_loader, ("mode",), n_posterior_samples, rng, distribute=distribute)["mode"]

    def aleatoric_variance(self, inputs_loader, n_posterior_samples: int = 30, rng=None, distribute: bool = True):
        return self._reduce(inputs_loader, ("aleatoric_variance",), n_posterior_samples, rng,
                            distribute=distribute)["aleatoric_variance"]

    def epistemic_variance(self, inputs_loader, n_posterior_samples: int = 30, rng=None, distribute: bool = True):
        return self._reduce(inputs_loader, ("epistemic_variance",), n_posterior_samples, rng,
                            distribute=distribute)["epistemic_variance"]

    def variance(self, inputs_loader, n_posterior_samples: int = 30, aleatoric_variances=None, epistemic_variances=None,
                 rng=None, distribute: bool = True):
        """base.py:838-898: ``variance = aleatoric + epistemic`` where each term, unless given, is estimated from ITS OWN
        draws - ``rng, key = split(rng)`` before the aleatoric term and again before the epistemic one, so the two use
        different sample index sets.  (``statistics(..., ("variance",))`` is the one-pass variant on a single sample set.)"""
        if rng is None:
            rng = self.posterior.rng.get()
        if aleatoric_variances is None:
            ks = ops.split(rng, 2)
            rng, key = ks[0].contiguous(), ks[1].contiguous()
            aleatoric_variances = self.aleatoric_variance(inputs_loader, n_posterior_samples, rng=key, distribute=distribute)
        if epistemic_variances is None:
            ks = ops.split(rng, 2)
            rng, key = ks[0].contiguous(), ks[1].contiguous()
            epistemic_variances = self.epistemic_variance(inputs_loader, n_posterior_samples, rng=key, distribute=distribute)
        dev = self._device()
        a = torch.as_tensor(aleatoric_variances, dtype=torch.float32, device=dev).contiguous()
        e = torch.as_tensor(epistemic_variances, dtype=torch.float32, device=dev).contiguous()
        return ops.variance_combine(a, e)

    def std(self, inputs_loader, n_posterior_samples: int = 30, variances=None, rng=None, distribute: bool = True):
        """base.py:900-944: ``sqrt(variance)``, the variance estimated as above unless given."""
        if variances is None:
            variances = self.variance(inputs_loader, n_posterior_samples, rng=rng, distribute=distribute)
        v = torch.as_tensor(variances, dtype=torch.float32, device=self._device()).contiguous()
        return ops.variance_combine(v, None, want_std=True)

    def entropy(self, inputs_loader, n_posterior_samples: int = 30, rng=None, distribute: bool = True):
        return self._reduce(inputs_loader, ("entropy",), n_posterior_samples, rng, distribute=distribute)["entropy"]

    def aleatoric_entropy(self, inputs_loader, n_posterior_samples: int = 30, rng=None, distribute: bool = True):
        return self._reduce(inputs_loader, ("aleatoric_entropy",), n_posterior_samples, rng,
                            distribute=distribute)["aleatoric_entropy"]

    def epistemic_entropy(self, inputs_loader, n_posterior_samples: int = 30, rng=None, distribute: bool = True):
        return self._reduce(inputs_loader, ("epistemic_entropy",), n_posterior_samples, rng,
                            distribute=distribute)["epistemic_entropy"]

    def log_prob(self, data_loader, n_posterior_samples: int = 30, rng=None, distribute: bool = True):
        return self._reduce(data_loader, ("log_prob",), n_posterior_samples, rng, self._targets(data_loader),
                            distribute=distribute)["log_prob"]

    def ensemble_log_prob(self, data_loader, n_posterior_samples: int = 30, rng=None, distribute: bool = True):
        return self._reduce(data_loader, ("ensemble_log_prob",), n_posterior_samples, rng, self._targets(data_loader),
                            distribute=distribute)["ensemble_log_prob"]

    def statistics(self, data_or_inputs_loader, names, n_posterior_samples: int = 30, rng=None, distribute: bool = True,
                   conformal=None, to_host=False, sharded_io: bool = False, ece: bool = False):
        """All requested statistics from ONE sample set and ONE pass over the loader (no reference counterpart: the
        reference recomputes the S forward passes per statistic).  ``conformal`` (a
        ``pipeline.ConformalCalibration``): additionally the adaptive-prediction-set conformal sets of the mean
        (``conformal_mask`` / ``conformal_sizes``), scored batch by batch while the next batch is in flight.
        ``to_host`` (True or a dict of pinned tensors to reuse): results as pinned HOST tensors, copied back batch by batch
        on their own stream.  ``sharded_io`` (multi-rank, ``distribute=True``): the loader holds THIS rank's rows of every
        batch and the results are those rows' (see ``streaming.stream_cls_reduce``).  ``ece``: additionally ``ece`` =
        (ECE, MCE) of the mean against the loader's targets (fortuna/metric/classification.py:43-133)."""
        targets = self._targets(data_or_inputs_loader) if hasattr(data_or_inputs_loader, "to_array_targets") else None
        ece_total = None
        if ece:
            if targets is None:
                raise ValueError("the calibration error needs a data loader (targets)")
            world, _ = streaming.dist_world(distribute)
            ece_total = self._n_rows(data_or_inputs_loader) * (world if sharded_io and world > 1 else 1)
        return self._reduce(data_or_inputs_loader, tuple(names), n_posterior_samples, rng, targets, distribute=distribute,
                            sets=conformal, to_host=to_host, sharded_io=sharded_io, ece_total=ece_total)

    def credible_set(self, inputs_loader, n_posterior_samples: int = 30, error: float = 0.05, rng=None,
                     distribute: bool = True) -> List[List[int]]:
        mean = self.mean(inputs_loader, n_posterior_samples, rng, distribute=distribute)
        labels, size = ops.credible_sets(mean, error)
        labels, size = labels.cpu().numpy(), size.cpu().numpy()
        return [labels[i, : size[i]].tolist() for i in range(labels.shape[0])]

    def sample(self, inputs_loader, n_target_samples: int = 1, return_aux=None, rng=None, distribute: bool = True) -> torch.Tensor:
        """int32 [T, B] class samples (base.py:318-440).  Like the reference, every input BATCH of the loader is sampled
        with the same key, so results depend on the loader's batch size exactly as there."""
        if return_aux:
            raise NotImplementedError("auxiliary outputs of `sample` are outside the hot path")
        if rng is None:
            rng = self.posterior.rng.get()
        stored = self._logits_fn(None, calibrated=False)  # uncalibrated: the sample kernel applies the temperature
        dev = self._device()
        parts = []
        for xb in self._input_batches(inputs_loader):
            zb = stored(torch.as_tensor(xb, dtype=torch.float32, device=dev)).float().contiguous()
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
        ell_train = self.ensemble_log_prob(train_data_loader, n_posterior_samples, rng=rng, distribute=distribute)
        z_test = self.sample_calibrated_outputs(test_inputs_loader, n_posterior_samples, rng)  # same rng: same draws
        lsm_test = ops.softmax_rows_(z_test.float().contiguous(), ops.EPI_LOG_SOFTMAX)
        out = ops.conformal_bayes_sets(ell_train, lsm_test, error, want_ess=return_ess)
        region = out["region"].cpu().numpy().astype(bool)
        sets = [np.nonzero(row)[0].tolist() for row in region]
        if return_ess:
            return sets, out["ess"].unsqueeze(-1)  # lax.map over test points of [C, 1] arrays (:601-621)
        return sets
