"""Quantile and one-dimensional-uncertainty conformal regressors (fortuna/conformal/regression/quantile.py,
onedim_uncertainty.py, base.py).  The conformal quantile ``jnp.quantile(scores, ceil((n+1)(1-error))/n)`` - a sort and an
interpolation over the validation scores - runs on the device (fb200_sort_f32 + fb200_quantile_sorted); ``score`` and
``conformal_interval`` are one elementwise expression each over validation-set-sized vectors, evaluated in float32 on the
host in the reference's operation order (small-n host algebra, not a kernel)."""
import numpy as np
import torch

from ... import ops
from ...utils.arrays import as_device, device_of


def _f32(a) -> np.ndarray:
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.asarray(a, dtype=np.float32)


def conformal_quantile_level(n: int, error: float) -> np.float32:
    return np.float32(np.ceil(np.float32((n + 1) * (1 - error))) / np.float32(n))


def conformal_quantile(scores, error: float) -> torch.Tensor:
    if error < 0 or error > 1:
        raise ValueError("""`error` must be a scalar between 0 and 1.""")
    s = as_device(scores, torch.float32, device_of(scores)).clone().view(-1)
    ops.sort_f32_(s)
    q = torch.tensor([conformal_quantile_level(s.numel(), error)], dtype=torch.float32, device=s.device)
    return ops.quantile_sorted(s, q)[0]


class ConformalRegressor:
    def is_in(self, values, conformal_intervals) -> np.ndarray:
        """fortuna/conformal/regression/base.py:5-27."""
        v, ci = _f32(values), _f32(conformal_intervals)
        return (v <= ci[:, 1]) * (v >= ci[:, 0])


class QuantileConformalRegressor(ConformalRegressor):
    def score(self, val_lower_bounds, val_upper_bounds, val_targets) -> np.ndarray:
        """quantile.py:10-52: max(lower - y, y - upper)."""
        lo, hi, y = _f32(val_lower_bounds), _f32(val_upper_bounds), _f32(val_targets)
        if lo.shape != hi.shape:
            raise ValueError("""The shapes of `val_lower_bounds` and `val_upper_bounds` must be the same, but shapes {} and {}
                were given, respectively.""".format(lo.shape, hi.shape))
        if lo.ndim != 1:
            raise ValueError("`val_lower_bounds` and `val_upper_bounds` must be one-dimensional arrays.")
        if y.ndim != 2 or y.shape[1] != 1:
            raise ValueError("""The second dimension of `val_targets` must have only one component.""")
        y = y.squeeze(1)
        return np.maximum(lo - y, y - hi)

    def quantile(self, val_lower_bounds=None, val_upper_bounds=None, val_targets=None, error: float = 0.05,
                 scores=None) -> torch.Tensor:
        if scores is None:
            scores = self.score(val_lower_bounds, val_upper_bounds, val_targets)
        return conformal_quantile(scores, error)

    def conformal_interval(self, val_lower_bounds, val_upper_bounds, test_lower_bounds, test_upper_bounds, val_targets,
                           error: float, quantile=None) -> np.ndarray:
        """quantile.py:86-179 -> float32 [n_test, 2] (lower - q, upper + q)."""
        vlo, vhi = _f32(val_lower_bounds), _f32(val_upper_bounds)
        tlo, thi = _f32(test_lower_bounds), _f32(test_upper_bounds)
        if vlo.shape != vhi.shape:
            raise ValueError("The shapes of `val_lower_bounds` and `val_upper_bounds` must be the same.")
        if tlo.shape != thi.shape:
            raise ValueError("The shapes of `test_lower_bounds` and `test_upper_bounds` must be the same.")
        for name, arr in (("val", vlo), ("test", tlo)):
            if arr.ndim not in (1, 2):
                raise ValueError(f"`{name}_lower_bounds` and `{name}_upper_bounds` must be one- or two-dimensional arrays. If "
                                 "two-dimensional, the second dimension must be 1.")
            if arr.ndim == 2 and arr.shape[1] != 1:
                raise ValueError(f"The second dimension of `{name}_lower_bounds` must have only one component, but"
                                 f"{arr.shape[1]} components were found.")
        if vlo.ndim == 2:
            vlo, vhi = vlo.squeeze(1), vhi.squeeze(1)
        if tlo.ndim == 2:
            tlo, thi = tlo.squeeze(1), thi.squeeze(1)
        if quantile is None:
            quantile = self.quantile(vlo, vhi, val_targets, error)
        q = np.float32(float(quantile))
        return np.stack([tlo - q, thi + q], axis=1)


class OneDimensionalUncertaintyConformalRegressor(ConformalRegressor):
    def score(self, val_preds, val_uncertainties, val_targets) -> np.ndarray:
        """onedim_uncertainty.py:10-52: |y - pred| / uncertainty."""
        p, u, y = _f32(val_preds), _f32(val_uncertainties), _f32(val_targets)
        if p.ndim != 2 or p.shape[1] != 1:
            raise ValueError("""`val_preds` must be a two-dimensional array. The second dimension must have only one component.""")
        if u.ndim != 2 or u.shape[1] != 1:
            raise ValueError("""`val_uncertainties` must be a two-dimensional array. The second dimension must have only one
                component.""")
        if (u <= 0).any():
            raise ValueError("""All elements in `val_uncertainties` must be strictly positive.""")
        return (np.abs(y - p) / u).squeeze(1)

    def quantile(self, val_preds=None, val_uncertainties=None, val_targets=None, error: float = 0.05,
                 scores=None) -> torch.Tensor:
        if scores is None:
            scores = self.score(val_preds, val_uncertainties, val_targets)
        return conformal_quantile(scores, error)

    def conformal_interval(self, val_preds, val_uncertainties, test_preds, test_uncertainties, val_targets, error: float,
                           quantile=None) -> np.ndarray:
        """onedim_uncertainty.py:92-150 -> float32 [n_test, 2] (pred -+ uncertainty * q)."""
        tp, tu = _f32(test_preds), _f32(test_uncertainties)
        if tp.ndim != 2 or tp.shape[1] != 1:
            raise ValueError("""`test_preds` must be a two-dimensional array. The second dimension must have only one component.""")
        if tu.ndim != 2 or tu.shape[1] != 1:
            raise ValueError("""`test_uncertainties` must be a two-dimensional array. The second dimension must have only one
                component.""")
        if (tu <= 0).any():
            raise ValueError("""All elements in `test_uncertainties` must be strictly positive.""")
        if quantile is None:
            quantile = self.quantile(val_preds, val_uncertainties, val_targets, error)
        q = np.float32(float(quantile))
        return np.stack([(tp - tu * q).squeeze(1), (tp + tu * q).squeeze(1)], axis=1)
