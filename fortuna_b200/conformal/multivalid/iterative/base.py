"""``IterativeMultivalidMethod`` - fortuna/conformal/multivalid/iterative/base.py:22-593.

The reference's round is a four-deep ``vmap`` of ``_calibration_error`` over (bucket type, group, bucket, dimension) that
materialises ``[n_taus, G, n_buckets, D, n]`` membership arrays, an ``argmax``, a patch and two loss evaluations.  Here the
data stay on the device: ONE kernel pass (``fb200_multivalid_stats``) returns the member count and the two sums every
(tau, g, v, c) cell needs, the few thousand cells are scored on the host in float32 with the reference's own formulas and
``argmax`` order, and the chosen patch is applied by ``fb200_multivalid_patch`` (calibration and validation split), the
losses by ``fb200_multivalid_loss``.  Control flow (early stopping, tolerances, status) is the reference's, line for line."""
import abc
import logging
from typing import Dict, Optional, Tuple, Union

import numpy as np
import torch

from .... import ops
from ....utils.arrays import as_device
from ..base import MultivalidMethod, np_of, to_f32

F32 = np.float32


def jax_permutation(seed: int, n: int, device) -> np.ndarray:
    """jax.random.choice(PRNGKey(seed), n, shape=(n,), replace=False) = permutation(key, n) (jax/_src/random.py
    ``_shuffle``): ceil(3 ln n / ln(2^32 - 1)) rounds of  key, sub = split(key); x = x[stable argsort(bits(sub, n))]."""
    key = ops.PRNGKey(seed, device)
    rounds = int(np.ceil(3 * np.log(max(1, n)) / np.log(np.iinfo(np.uint32).max)))
    x = np.arange(n)
    for _ in range(rounds):
        ks = ops.split(key, 2)
        key, sub = ks[0].contiguous(), ks[1].contiguous()
        bits = ops.random_bits(sub, n).cpu().numpy().view(np.uint32)
        x = x[np.argsort(bits, kind="stable")]
    return x


class IterativeMultivalidMethod(MultivalidMethod):
    _mode = 0          # statistic of fb200_multivalid_stats: 0 the score, 1 [score <= value]
    _top_label = False

    def __init__(self, seed: int = 0):
        super().__init__(seed=seed)
        self._patches = []
        self._eta = None
        self._bucket_types = None
        self._patch_type = None

    # ------------------------------------------------------------------ hooks of the subclasses
    @abc.abstractmethod
    def _cell_errors(self, stats: np.ndarray, n: int, **kwargs) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """stats float64 [n_taus, G, n_buckets, D, 3] -> (error [n_taus, G, n_buckets, D], prob_b, additive patch) float32."""

    @abc.abstractmethod
    def _get_patch(self, cell, stats, n, values, scores, groups, buckets, patch_type, **kwargs) -> float:
        pass

    def _loss(self, values: torch.Tensor, scores: torch.Tensor) -> float:
        raise NotImplementedError

    # ------------------------------------------------------------------ device plumbing
    @staticmethod
    def _init_groups(groups, size: int, device) -> torch.Tensor:
        if groups is None:
            return torch.ones((size, 1), dtype=torch.uint8, device=device)
        return as_device(groups, torch.uint8, device).clone()

    @staticmethod
    def _maybe_check_groups(groups, test_groups=None):
        for g in (groups, test_groups):
            if g is not None and np.ndim(g) != 2:
                raise ValueError("`groups` must be a 2-dimensional array.")
        for g in (groups, test_groups):
            if g is not None and (g.dtype != torch.bool if isinstance(g, torch.Tensor) else np.asarray(g).dtype != bool):
                raise ValueError("All elements in `groups` must be a bool.")

    def _maybe_init_values(self, values, size: Optional[int], device) -> torch.Tensor:
        if values is None:
            if size is None:
                raise ValueError("If `values` is not provided, `size` must be provided.")
            return torch.zeros((size, 1), dtype=torch.float32, device=device)
        v = to_f32(values, device).clone()
        return v[:, None].contiguous() if v.ndim == 1 else v

    @staticmethod
    def _get_bucket_type_indices(bucket_types: Tuple[str]):
        taus = []
        if "=" in bucket_types:
            taus.append(0)
        if ">=" in bucket_types:
            taus.append(1)
        if "<=" in bucket_types:
            taus.append(2)
        return taus

    @staticmethod
    def _check_bucket_types(bucket_types: Tuple[str]):
        for bucket_type in bucket_types:
            if bucket_type not in ["=", ">=", "<="]:
                raise ValueError(f"`bucket_type='{bucket_type}'` not recognized.")

    def _stats(self, values, scores, groups, buckets_dev, taus) -> np.ndarray:
        full = ops.multivalid_stats(values, scores, groups, buckets_dev, self._top_label, self._mode, taus)
        return full.cpu().numpy()[list(taus)]

    def _apply(self, values, groups, patch_rec, eta, patch_type) -> torch.Tensor:
        taut, gt, vt, ct, patch = patch_rec
        return ops.multivalid_patch_(values, groups, gt, ct, taut, vt, self.n_buckets, eta, patch, self._top_label,
                                     patch_type == "multiplicative")

    # ------------------------------------------------------------------ calibrate (iterative/base.py:34-333)
    def calibrate(self, scores, groups=None, values=None, test_groups=None, test_values=None, atol: float = 1e-4,
                  rtol: float = 1e-6, min_prob_b: Union[float, str] = "auto", n_buckets: int = 100, n_rounds: int = 1000,
                  eta: float = 0.1, split: float = 0.8, bucket_types: Tuple[str, ...] = ("<=", ">="),
                  patch_type: str = "additive", **kwargs) -> Union[Dict, Tuple[torch.Tensor, Dict]]:
        if n_rounds < 1:
            raise ValueError("`n_rounds` must be at least 1.")
        if test_values is not None and values is None:
            raise ValueError("If `test_values is provided, `values` must also be provided.")
        if values is not None and test_groups is not None and test_values is None:
            raise ValueError("If `values` and `test_groups` are provided, `test_values` must also be provided.")
        if test_groups is not None and groups is None:
            raise ValueError("If `test_groups` is provided, `groups` must also be provided.")
        if groups is not None and test_values is not None and test_groups is None:
            raise ValueError("If `groups` and `test_values` are provided, `test_groups` must also be provided.")
        if eta <= 0 or eta > 1:
            raise ValueError("`eta` must be a float greater than 0 and less than or equal to 1.")
        if split <= 0 or split > 1:
            raise ValueError("`split` must be a float greater than 0 and less than 1.")
        if min_prob_b != "auto" and (min_prob_b < 0 or min_prob_b > 1):
            raise ValueError("`min_prob_b` must be greater than or equal to 0 and less than or equal to 1.")
        allowed_patch_types = ["additive", "multiplicative"]
        if patch_type not in allowed_patch_types:
            raise ValueError(f"`patch_type={patch_type}` not recognized. Please select one among the following options: "
                             f"{allowed_patch_types}.")
        self._check_scores(scores)
        scores = self._process_scores(scores)
        dev = scores.device
        n_dims = scores.shape[1]
        min_prob_b = self._maybe_init_min_prob_b(min_prob_b=min_prob_b, n_buckets=n_buckets, n_dims=n_dims)
        self._maybe_check_groups(groups, test_groups)
        groups = self._init_groups(groups, scores.shape[0], dev)
        values_given = values
        self._maybe_check_values(values_given, test_values)
        values = self._maybe_init_values(values, groups.shape[0], dev)
        self.n_buckets = n_buckets
        buckets = self._get_buckets(n_buckets)
        buckets_dev = torch.from_numpy(buckets).to(dev)
        self._check_bucket_types(bucket_types)
        taus = self._get_bucket_type_indices(bucket_types)
        self._bucket_types = bucket_types
        self._eta = eta
        self._patch_type = patch_type

        size = len(scores)
        calib_size = int(np.ceil(F32(split * size)))  # jnp.ceil of the python product: rounded to float32 first
        perm = torch.from_numpy(jax_permutation(self._seed, size, dev)).to(dev)
        tr, va = perm[:calib_size], perm[calib_size:]
        scores, val_scores = scores[tr].contiguous(), scores[va].contiguous()
        values, val_values = values[tr].contiguous(), values[va].contiguous()
        groups, val_groups = groups[tr].contiguous(), groups[va].contiguous()
        n = scores.shape[0]

        val_losses = [self._loss(val_values, val_scores)]
        old_losses = self._loss(values, scores)
        self._patches = []
        converged = False
        t = 0
        for t in range(n_rounds):
            stats = self._stats(values, scores, groups, buckets_dev, taus)
            err, prob_b, add_patch = self._cell_errors(stats, n, **kwargs)
            cond_b = prob_b >= F32(min_prob_b)
            if not np.count_nonzero(cond_b):
                logging.warning(f"Early stopping triggered after {t} rounds. No further sets with probability at least "
                                "`min_prob_b` left available.")
                break
            err = err * cond_b
            cell = np.unravel_index(int(np.argmax(err)), err.shape)  # (tau index, g, bucket index, c): first maximum
            taut, gt, vt, ct = taus[cell[0]], int(cell[1]), float(buckets[cell[2]]), int(cell[3])
            patch = self._get_patch(cell=cell, stats=stats, n=n, add_patch=add_patch, values=values, scores=scores,
                                    groups=groups, buckets=buckets, patch_type=patch_type, tau=taut, g=gt, v=vt, c=ct,
                                    **kwargs)
            rec = (taut, gt, vt, ct, float(patch))
            self._apply(values, groups, rec, eta, patch_type)
            val_members = self._apply(val_values, val_groups, rec, eta, patch_type) if len(val_values) else None
            losses = self._loss(values, scores)
            val_losses.append(self._loss(val_values, val_scores))
            if losses > old_losses:
                logging.warning("The algorithm cannot achieve the desired tolerance. Please try increasing `n_buckets`, or "
                                "decreasing `min_prob_b`.")
                break
            if val_losses[-1] <= atol:
                logging.info(f"Absolute tolerance satisfied after {t} rounds.")
                converged = True
                break
            if val_losses[-1] > val_losses[-2]:
                logging.warning(f"Early stopping triggered after {t} rounds. The loss started increasing on the "
                                "validation data.")
                break
            any_val = val_members is not None and int(val_members.item()) > 0
            if any_val and (val_losses[-2] - val_losses[-1]) / val_losses[-2] <= rtol:
                logging.info(f"Relative tolerance satisfied after {t} rounds.")
                converged = True
                break
            old_losses = losses
            self._patches.append(rec)
        status = dict(n_rounds=len(self.patches), losses=val_losses, converged=converged)
        if t == n_rounds - 1 and not converged:
            logging.warning("Maximum number of rounds reached without convergence. Consider adjusting the hyperparameters. "
                            "In particular: increasing `n_rounds`, increasing `n_buckets`, decreasing `min_prob_b` and "
                            "increasing `eta`.")
        if test_groups is not None or test_values is not None:
            return self.apply_patches(test_groups, test_values), status
        return status

    # ------------------------------------------------------------------ apply_patches (iterative/base.py:335-371)
    def apply_patches(self, groups=None, values=None) -> torch.Tensor:
        if groups is None and values is None:
            raise ValueError("At least one between `groups` and `values` must be provided.")
        if not len(self._patches):
            logging.warning("No patches available.")
            return values
        dev = torch.device("cuda") if not isinstance(values, torch.Tensor) or not values.is_cuda else values.device
        one_d = values is None or np.ndim(values) == 1
        self._maybe_check_values(values)
        values = self._maybe_init_values(values, len(groups) if groups is not None else None, dev)
        self._maybe_check_groups(groups)
        groups = self._init_groups(groups, values.shape[0], dev)
        for rec in self._patches:
            self._apply(values, groups, rec, self.eta, self.patch_type)
        return values[:, 0].contiguous() if one_d and values.shape[1] == 1 else values

    # ------------------------------------------------------------------ calibration_error (iterative/base.py:373-407)
    def calibration_error(self, scores, groups=None, values=None, n_buckets: int = 10, **kwargs) -> torch.Tensor:
        self._check_scores(scores)
        scores = self._process_scores(scores)
        dev = scores.device
        n_dims = scores.shape[1]
        self._maybe_check_groups(groups)
        self._maybe_check_values(values)
        groups = self._init_groups(groups, scores.shape[0], dev)
        values = self._maybe_init_values(values, scores.shape[0], dev)
        buckets_dev = torch.from_numpy(self._get_buckets(n_buckets)).to(dev)
        saved, self.n_buckets = self.n_buckets, n_buckets
        try:
            stats = self._stats(values, scores, groups, buckets_dev, [0])
        finally:
            self.n_buckets = saved
        err, _, _ = self._cell_errors(stats, scores.shape[0], **kwargs)
        error = err[0].sum(1, dtype=F32)                              # [G, D]
        gmean = (groups.sum(0, dtype=torch.float64).cpu().numpy() / scores.shape[0]).astype(F32)
        error = (error / gmean[:, None]).astype(F32)
        if n_dims == 1:
            error = error[:, 0]
        return torch.from_numpy(error).to(dev)

    @property
    def eta(self):
        return self._eta

    @eta.setter
    def eta(self, eta):
        self._eta = eta

    @property
    def patch_type(self):
        return self._patch_type

    @property
    def bucket_types(self):
        return self._bucket_types
