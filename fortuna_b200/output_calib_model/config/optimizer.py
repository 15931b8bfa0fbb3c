from .base import Optimizer  # noqa: F401
