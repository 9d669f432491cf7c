"""wannierberri_b200 -- B200-native k-grid evaluation hot path of wannier-berri
(wannierberri.fourier + data_K + formula + calculators.static) behind the reference's
System_R / run() / calculators interface.  Arithmetic runs in libwbgpu.so (hand-written sm_100a
CUDA behind a C-ABI, include/wbgpu.h); there is no CPU fallback."""
from . import calculators  # noqa: F401
from .system import System_R, SystemSOC, synthetic_system, kramers_system  # noqa: F401
from .grid import Grid  # noqa: F401
from .result import EnergyResult, ResultDict  # noqa: F401
from .engine import Engine  # noqa: F401
from .data_K import Data_K_R  # noqa: F401
from .run import run  # noqa: F401
from . import smoother  # noqa: F401
from .smoother import get_smoother  # noqa: F401

__version__ = "0.1.0"
