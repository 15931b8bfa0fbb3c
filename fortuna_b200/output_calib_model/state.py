"""``OutputCalibState`` - fortuna/output_calib_model/state.py:22-109: the calibration parameters
(``{"output_calibrator": {"params": {"log_temp": f32[1]}}}`` for the temperature scalers), the optimizer state and the
step counter, with the reference's checkpoint state dict (flax ``to_state_dict`` of the TrainState dataclass:
step, params, opt_state, encoded_name, frozen_params, dynamic_scale, mutable)."""
from typing import Any, Dict, Optional

import numpy as np

from ..training.checkpoint import encode_name


class OutputCalibState:
    encoded_name = encode_name("OutputCalibState")

    def __init__(self, params: Dict, mutable: Optional[Dict] = None, opt_state: Any = None, step: int = 0, tx: Any = None):
        self.params = params
        self.mutable = mutable if mutable is not None else {"output_calibrator": None}
        self.opt_state = opt_state
        self.step = int(step)
        self.tx = tx

    # ------------------------------------------------------------------ the temperature (a float32[1] host array)
    @property
    def log_temp(self) -> np.ndarray:
        return self.params["output_calibrator"]["params"]["log_temp"]

    @classmethod
    def init(cls, params: Dict, mutable: Optional[Dict] = None, optimizer: Any = None, **kwargs) -> "OutputCalibState":
        lt = np.asarray(params["output_calibrator"]["params"]["log_temp"], dtype=np.float32).reshape(1)
        params = {"output_calibrator": {"params": {"log_temp": lt}}}
        if optimizer is None:
            opt_state = kwargs.get("opt_state")
        else:
            opt_state = optimizer.init(lt)
        return cls(params=params, mutable=mutable, opt_state=opt_state, step=kwargs.get("step", 0), tx=optimizer)

    @classmethod
    def init_from_dict(cls, d: Dict, optimizer: Any = None, **kwargs) -> "OutputCalibState":
        """state.py:70-109: a restored checkpoint dict (or an init dict) -> state; a given optimizer restarts its state."""
        extra = {k: v for k, v in d.items() if k not in ("params", "mutable", "optimizer")}
        if "opt_state" in extra and isinstance(extra["opt_state"], dict) and "0" in extra["opt_state"]:
            a = extra["opt_state"]["0"]  # optax.adam: (ScaleByAdamState(count, mu, nu), EmptyState())
            leaf = lambda t: np.asarray(t["output_calibrator"]["params"]["log_temp"], dtype=np.float32).reshape(1)
            extra["opt_state"] = {"count": int(np.asarray(a["count"])), "mu": leaf(a["mu"]), "nu": leaf(a["nu"])}
        extra.update(kwargs)
        return cls.init(d["params"], d.get("mutable"), optimizer, **{k: extra[k] for k in ("opt_state", "step") if k in extra})

    def apply_gradients(self, grad: np.ndarray) -> "OutputCalibState":
        """flax TrainState.apply_gradients: tx.update, optax.apply_updates, step + 1."""
        upd, opt_state = self.tx.update(np.asarray(grad, dtype=np.float32).reshape(1), self.opt_state)
        lt = (self.log_temp + upd).astype(np.float32)
        return OutputCalibState({"output_calibrator": {"params": {"log_temp": lt}}}, self.mutable, opt_state, self.step + 1,
                                self.tx)

    def to_state_dict(self) -> Dict:
        tree = lambda v: {"output_calibrator": {"params": {"log_temp": np.asarray(v, dtype=np.float32).reshape(1)}}}
        if self.opt_state is None:
            opt = None
        else:
            opt = {"0": {"count": np.asarray(self.opt_state["count"], dtype=np.int32), "mu": tree(self.opt_state["mu"]),
                         "nu": tree(self.opt_state["nu"])}, "1": {}}
        return {
            "step": self.step,
            "params": tree(self.log_temp),
            "opt_state": opt,
            "encoded_name": {str(i): c for i, c in enumerate(self.encoded_name)},
            "frozen_params": None,
            "dynamic_scale": None,
            "mutable": {"output_calibrator": None},
        }
