from .config import OptimizationConfig  # noqa: F401
