from .base import Checkpointer, Config, Monitor, Optimizer, Processor  # noqa: F401
