"""A/B of the fused axes-1+0 pass of the R->k transform (option fourier_method): stage time per 256k k-points of the
headline workload and the largest difference of the scan results to the default.  python profiles/ab_fourier.py"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wannierberri_b200 as wb
from wannierberri_b200 import _lib

fe = wb.System_R.from_npz(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "fe_system.npz"))
st = wb.calculators.static
Ef = np.linspace(12., 22., 2000)
specs = st.AHC(Efermi=Ef).specs() + st.DOS(Efermi=Ef).specs()
shifts, factors = wb.Grid(fe, NKdiv=[20] * 3, NKFFT=[20] * 3).K_arrays()
ref = None
for method in [int(a) for a in sys.argv[1:]] or [0, 2, 1]   # 0 = default (partial sums in registers), 2 = KC-accumulator fused kernel, 1 = two passes:
    eng = wb.Engine(fe)
    eng.set_option("fourier_method", method)
    eng.plan([20] * 3, [s.formula for s in specs])
    out = eng.scan(shifts[:32], factors[:32], specs)
    eng.set_option("timing", 1)
    for _ in range(3):
        out = eng.scan(shifts[:32], factors[:32], specs)
    ms, calls = (C.c_double * 5)(), (C.c_int64 * 5)()
    _lib.lib().wbgpu_stage_times(eng._ctx, ms, calls)
    if ref is None:
        ref = out
    err = max(np.abs(a - b).max() / np.abs(b).max() for a, b in zip(out, ref))
    print(f"fourier_method {method}: fourier {ms[0] / 3:.3f} ms per 256k k-points; max rel diff to the first variant {err:.1e}", flush=True)
    eng.close()
