"""Calibration from model outputs - the reference's ``fortuna.output_calib_model`` usage mode: the S = 1
special case of the predictive maps, driven by the same kernels."""
from .classification import OutputCalibClassifier  # noqa: F401
from .config.base import Checkpointer, Config, Monitor, Optimizer, Processor  # noqa: F401
from .regression import OutputCalibRegressor  # noqa: F401
