from .loader import DataLoader, InputsLoader, TargetsLoader  # noqa: F401
