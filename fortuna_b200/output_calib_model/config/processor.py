from .base import Processor  # noqa: F401
