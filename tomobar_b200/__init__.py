"""B200-native (sm_100a) parallel-beam reconstruction hot path behind the ToMoBAR
``RecToolsIRCuPy`` / ``RecToolsDIRCuPy`` interface.  Importing the package loads libtmb.so;
there is no CPU fallback."""

from tomobar_b200 import _lib  # noqa: F401  (fails loudly when the CUDA library is missing)
from tomobar_b200.projector import ProjTools3D  # noqa: F401
from tomobar_b200.regularisersCuPy import PD_TV_cupy, ROF_TV_cupy, prox_regul  # noqa: F401
from tomobar_b200.methodsIR_CuPy import RecToolsIRCuPy  # noqa: F401
from tomobar_b200.methodsDIR_CuPy import RecToolsDIRCuPy  # noqa: F401

__version__ = "0.1.0"
