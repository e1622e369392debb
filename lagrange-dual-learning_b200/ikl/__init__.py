"""ikl -- B200-native inverse-kinematics Lagrangian kernels (FK + penalties + dual ascent) behind a C ABI.

Importing this package does not touch CUDA; the shared object is loaded (and, if missing, built) on first use.
"""
from .spec import ConstraintSpec  # noqa: F401
from .configs import standard_config, load_config, STANDARD_CONFIGS  # noqa: F401

__all__ = ["ConstraintSpec", "standard_config", "load_config", "STANDARD_CONFIGS"]
