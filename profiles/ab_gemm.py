"""A/B of the size-generic rotation (option gemm_stack: 1 = several channels per CTA for nw <= 32, 0 = one channel per
CTA) on the Te Fermi-surface workload of BASELINE config 3 and on Fe AHC + Morb through the generic path."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wannierberri_b200 as wb
from wannierberri_b200 import _lib

st = wb.calculators.static
te = wb.System_R.from_npz(os.path.join(ROOT, "tests", "golden", "te_system.npz"))
fe = wb.System_R.from_npz(os.path.join(ROOT, "tests", "golden", "fe_system.npz"))
Ef = np.linspace(4., 8., 401)
cases = [("Te BCD+GME Fermi surface", te, [10] * 3, st.BerryDipole_FermiSurf(Efermi=Ef).specs() + st.GME_orb_FermiSurf(Efermi=Ef).specs()
          + st.GME_spin_FermiSurf(Efermi=Ef).specs(), 16, 0),
         ("Fe AHC+Morb generic path", fe, [20] * 3, st.AHC(Efermi=np.linspace(12., 22., 2000)).specs() + st.Morb(Efermi=np.linspace(12., 22., 2000)).specs(), 16, 4)]
for name, system, div, specs, nb, rot in cases:
    shifts, factors = wb.Grid(system, NKdiv=div, NKFFT=[20] * 3).K_arrays()
    ref = None
    for stack in (0, 1):
        eng = wb.Engine(system)
        eng.set_option("gemm_stack", stack)
        eng.set_option("rotate_method", rot)
        eng.plan([20] * 3, [s.formula for s in specs])
        out = eng.scan(shifts[:nb], factors[:nb], specs)
        eng.set_option("timing", 1)
        for _ in range(2):
            out = eng.scan(shifts[:nb], factors[:nb], specs)
        ms, calls = (C.c_double * 5)(), (C.c_int64 * 5)()
        _lib.lib().wbgpu_stage_times(eng._ctx, ms, calls)
        ref = out if ref is None else ref
        err = max(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300) for a, b in zip(out, ref))
        print(f"{name}: gemm_stack {stack}: rotate {ms[2] / 2:.2f} ms per {nb * 8000} k-points, diff {err:.1e}", flush=True)
        eng.close()
