"""Calibration configuration with the reference's names (fortuna/prob_model/calib_config/*.py): the fields the
temperature-scaling calibration reads."""
from typing import Optional


class Adam:
    """optax.adam(learning_rate, b1=0.9, b2=0.999, eps=1e-8) restated for the calibrator's (tiny) parameter vector:
    mu <- b1 mu + (1-b1) g;  nu <- b2 nu + (1-b2) g^2;  update = -lr * (mu / (1-b1^t)) / (sqrt(nu / (1-b2^t)) + eps)."""

    def __init__(self, learning_rate: float = 1e-2, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8):
        self.learning_rate, self.b1, self.b2, self.eps = learning_rate, b1, b2, eps

    def init(self, params):
        import numpy as np

        p = np.asarray(params, dtype=np.float32)
        return {"count": 0, "mu": np.zeros_like(p), "nu": np.zeros_like(p)}

    def update(self, grad, state):
        import numpy as np

        f = np.float32
        g = np.asarray(grad, dtype=f)
        count = state["count"] + 1
        mu = (f(self.b1) * state["mu"] + f(1 - self.b1) * g).astype(f)
        nu = (f(self.b2) * state["nu"] + f(1 - self.b2) * g * g).astype(f)
        mu_hat = mu / f(1 - f(self.b1) ** count)
        nu_hat = nu / f(1 - f(self.b2) ** count)
        upd = (f(-self.learning_rate) * mu_hat / (np.sqrt(nu_hat) + f(self.eps))).astype(f)
        return upd, {"count": count, "mu": mu, "nu": nu}


class CalibOptimizer:
    """fortuna/prob_model/calib_config/optimizer.py:8-26: default ``optax.adam(1e-2)``, 100 epochs."""

    def __init__(self, method: Optional[Adam] = None, n_epochs: int = 100):
        self.method = method if method is not None else Adam(1e-2)
        self.n_epochs = n_epochs


class CalibProcessor:
    """fortuna/prob_model/calib_config/processor.py: posterior samples whose outputs form the fixed ensemble."""

    def __init__(self, n_posterior_samples: int = 30, **_ignored):
        self.n_posterior_samples = n_posterior_samples


class CalibConfig:
    def __init__(self, optimizer: CalibOptimizer = None, processor: CalibProcessor = None, **_ignored):
        self.optimizer = optimizer or CalibOptimizer()
        self.processor = processor or CalibProcessor()
