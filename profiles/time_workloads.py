#!/usr/bin/env python
"""Stage times (CUDA events inside the library) of the three BASELINE static workloads on one GPU.
   python profiles/time_workloads.py [blocks]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wannierberri_b200 as wb  # noqa: E402
from wannierberri_b200 import _lib  # noqa: E402

st = wb.calculators.static
blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 16
fe = wb.System_R.from_npz(os.path.join(ROOT, "tests", "golden", "fe_system.npz"))
te = wb.System_R.from_npz(os.path.join(ROOT, "tests", "golden", "te_system.npz"))
Ef = np.linspace(12.0, 22.0, 2000)
EfT = np.linspace(4.0, 8.0, 401)
cases = {
    "fe_ahc_dos": (fe, dict(ahc=st.AHC(Efermi=Ef), dos=st.DOS(Efermi=Ef))),
    "fe_ahc_morb": (fe, dict(ahc=st.AHC(Efermi=Ef), morb=st.Morb(Efermi=Ef))),
    "te_bcd_gme": (te, dict(bcd=st.BerryDipole_FermiSurf(Efermi=EfT), gme_orb=st.GME_orb_FermiSurf(Efermi=EfT),
                            gme_spin=st.GME_spin_FermiSurf(Efermi=EfT))),
}
for name, (system, calcs) in cases.items():
    specs = [s for c in calcs.values() for s in c.specs()]
    eng = wb.Engine(system, device=0)
    eng.plan([20, 20, 20], [s.formula for s in specs], external_terms=True)
    grid = wb.Grid(system, NKdiv=[20, 20, 20], NKFFT=[20, 20, 20])
    shifts, factors = grid.K_arrays()
    eng.scan(shifts[:blocks], factors[:blocks], specs)
    eng.set_option("timing", 1)
    nrep = 3
    for _ in range(nrep):
        eng.scan(shifts[:blocks], factors[:blocks], specs)
    ms = (C.c_double * 5)()
    calls = (C.c_int64 * 5)()
    _lib.check(_lib.lib().wbgpu_stage_times(eng._ctx, ms, calls))
    tot = sum(ms) / nrep
    nk = blocks * 8000
    print(f"{name:12s} nw={system.num_wann} {nk} k-points: " +
          " ".join(f"{n}={ms[i] / nrep:.2f}" for i, n in enumerate(["fourier", "eigh", "rotate", "identity", "scan"])) +
          f" total={tot:.2f} ms -> {nk / tot * 1e3:.3e} k/s")
    eng.close()
