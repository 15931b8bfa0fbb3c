from .base import Monitor  # noqa: F401
