"""``OneShotMultivalidMethod`` - fortuna/conformal/multivalid/one_shot/base.py:17-153: one patch per (bucket, dimension)
from a single pass over the calibration data (``fb200_multivalid_stats`` with the "=" buckets and no groups), applied to
new values by bucket look-up."""
import abc
import logging
from typing import Optional, Union

import numpy as np
import torch

from .... import ops
from ..base import MultivalidMethod, to_f32

F32 = np.float32


class OneShotMultivalidMethod(MultivalidMethod):
    _top_label = False

    def __init__(self, seed: int = 0):
        super().__init__(seed=seed)
        self._patches = None
        self._n_buckets = None

    def calibrate(self, scores, values=None, test_values=None, n_buckets: int = 100,
                  min_prob_b: Union[float, str] = "auto") -> Optional[torch.Tensor]:
        if test_values is not None and values is None:
            raise ValueError("If `test_values is provided, `values` must also be provided.")
        if min_prob_b != "auto" and (min_prob_b < 0 or min_prob_b > 1):
            raise ValueError("`min_prob_b` must be greater than or equal to 0 and less than or equal to 1.")
        if values is None:
            raise ValueError("`values` must be provided.")
        self._check_scores(scores)
        scores = self._process_scores(scores)
        n_dims = scores.shape[1]
        min_prob_b = self._maybe_init_min_prob_b(min_prob_b=min_prob_b, n_buckets=n_buckets, n_dims=n_dims)
        self._maybe_check_values(values, test_values)
        v = to_f32(values, scores.device)
        v = v[:, None].contiguous() if v.ndim == 1 else v
        self.n_buckets = n_buckets
        buckets = self._get_buckets(n_buckets)
        stats = ops.multivalid_stats(v, scores, None, torch.from_numpy(buckets).to(v.device), self._top_label, 0, (0,))
        self._patches = self._get_patches(stats.cpu().numpy()[0, 0], scores.shape[0], buckets, min_prob_b)  # [n_buckets, D]
        if test_values is not None:
            return self.apply_patches(test_values)

    @abc.abstractmethod
    def _get_patches(self, stats: np.ndarray, n: int, buckets: np.ndarray, min_prob_b: float) -> np.ndarray:
        pass

    def apply_patches(self, values) -> torch.Tensor:
        """one_shot/base.py:68-126.  The reference visits the sorted unique values u of ``values`` and overwrites every
        entry within h of u (|x - u| < h, float32; for the top-label variant only the arg-max column) with the patch of the
        bucket nearest to u - so an entry ends up with the patch of the LARGEST unique value within h of it.  That look-up
        is done here with sorted searches on the device."""
        if self._patches is None:
            logging.warning("No patches available.")
            return values
        x = to_f32(values).clone()
        dev = x.device
        one_d = x.ndim == 1
        x2 = x[:, None] if one_d else x
        h = F32(0.5 / self.n_buckets)
        buckets = torch.from_numpy(self._get_buckets(self.n_buckets)).to(dev)
        patches = torch.from_numpy(np.asarray(self._patches, dtype=F32)).to(dev)
        uniq = torch.unique(x2)  # sorted, flattened like jnp.unique
        # bucket nearest to each unique value (first minimum of |buckets - u|, like jnp.argmin)
        idx_u = torch.argmin((buckets[None, :] - uniq[:, None]).abs(), dim=1)
        out = x2.clone()
        top = x2.argmax(1) if self._top_label else None
        for c in range(x2.shape[1]):
            col = x2[:, c].contiguous()
            # largest unique u with |col - u| < h: candidates lie just below col + h
            pos = torch.searchsorted(uniq, col + h, right=False) - 1
            found = torch.zeros_like(col, dtype=torch.bool)
            best = torch.zeros_like(pos)
            for back in range(3):  # float32 rounding of col + h can misplace the boundary by one slot either way
                p = (pos + 1 - back).clamp(0, uniq.numel() - 1)
                ok = ((col - uniq[p]).abs() < h) & ~found
                best = torch.where(ok, p, best)
                found |= ok
            sel = found if top is None else (found & (top == c))
            out[:, c] = torch.where(sel, patches[idx_u[best], c], out[:, c])
        return out[:, 0].contiguous() if one_d else out
