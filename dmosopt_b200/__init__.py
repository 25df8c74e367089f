"""dmosopt_b200 -- B200-native surrogate-generation hot path for dmosopt.

Plugin import paths (dmosopt resolves them with ``config.import_object_by_path``):

    optimizer_name         = "dmosopt_b200.NSGA2" | "dmosopt_b200.AGEMOEA" | "dmosopt_b200.SMPSO" | "dmosopt_b200.CMAES"
    surrogate_method_name  = "dmosopt_b200.GPR_Matern" | "dmosopt_b200.GPR_RBF"

Importing the package does not touch CUDA; the first numerical call loads
``libdmosopt_b200.so`` and creates the context, and fails loudly when either
is unavailable (there is no CPU fallback).
"""

from .MOEA import MOEA as MOEABase  # noqa: F401
from .MOEA import Struct  # noqa: F401
from .NSGA2 import NSGA2  # noqa: F401
from .model import GPR_Matern, GPR_RBF, Model  # noqa: F401

try:  # optimizers that arrive later in the build keep the package importable
    from .AGEMOEA import AGEMOEA  # noqa: F401
except ImportError:  # pragma: no cover
    pass
try:
    from .SMPSO import SMPSO  # noqa: F401
except ImportError:  # pragma: no cover
    pass
try:
    from .CMAES import CMAES  # noqa: F401
except ImportError:  # pragma: no cover
    pass

__version__ = "0.1.0"
