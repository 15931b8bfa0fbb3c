from .base import Checkpointer  # noqa: F401
