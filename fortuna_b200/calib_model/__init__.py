"""Calibration models - fortuna/calib_model/__init__.py: a model trained against a calibration loss, and the S = 1
predictive maps over its outputs."""
from .classification import CalibClassifier
from .config.base import Checkpointer, Config, Monitor, Optimizer, Processor
from .regression import CalibRegressor

__all__ = ["CalibClassifier", "CalibRegressor", "Config", "Optimizer", "Checkpointer", "Monitor", "Processor"]
