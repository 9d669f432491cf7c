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

# ---- BASELINE config 4 in miniature: synthetic 32-WF model, Kubo optical conductivity 500 omega x 200 Efermi
import time  # noqa: E402
dyn = wb.calculators.dynamic
s32 = wb.synthetic_system(32, rmax=2, seed=20261017)
eng = wb.Engine(s32, device=0)
eng.plan([16, 16, 16], [_lib.IDENTITY, _lib.KUBO], external_terms=True)
grid = wb.Grid(s32, NKdiv=[8, 8, 8], NKFFT=[16, 16, 16])
shifts, factors = grid.K_arrays()
oc = dyn.OpticalConductivity(Efermi=np.linspace(-1, 1, 200), omega=np.linspace(0, 5, 500), smr_fixed_width=0.1)
nb = 4
eng.kubo_scan(shifts[:1], factors[:1], oc.spec(), oc.Efermi, oc.omega)
eng.set_option("timing", 1)
t0 = time.perf_counter()
eng.kubo_scan(shifts[:nb], factors[:nb], oc.spec(), oc.Efermi, oc.omega)
dt = time.perf_counter() - t0
ms = (C.c_double * 5)()
calls = (C.c_int64 * 5)()
_lib.check(_lib.lib().wbgpu_stage_times(eng._ctx, ms, calls))
nk = nb * 4096
print(f"kubo32_optcond nw=32 {nk} k-points, 500 omega x 200 Ef: " +
      " ".join(f"{n}={ms[i]:.2f}" for i, n in enumerate(["fourier", "eigh", "rotate", "identity", "entries+accumulate"])) +
      f" wall={dt * 1e3:.1f} ms -> {nk / dt:.3e} k/s")
eng.close()

# ---- BASELINE config 5 in miniature: synthetic 128-WF model (nR = 343), DOS + CumDOS + AHC
st = wb.calculators.static
s128 = wb.synthetic_system(128, rmax=3, seed=5)
Ef5 = np.linspace(-4, 4, 401)
for label, calcs in (("dos_cumdos", dict(dos=st.DOS(Efermi=Ef5), cumdos=st.CumDOS(Efermi=Ef5))),
                     ("ahc", dict(ahc=st.AHC(Efermi=Ef5)))):
    specs = [s for c in calcs.values() for s in c.specs()]
    eng = wb.Engine(s128, device=0)
    eng.plan([16, 16, 16], [s.formula for s in specs], external_terms=True)
    grid = wb.Grid(s128, NKdiv=[4, 4, 4], NKFFT=[16, 16, 16])
    shifts, factors = grid.K_arrays()
    eng.scan(shifts[:1], factors[:1], specs)
    eng.set_option("timing", 1)
    nb = 2
    eng.scan(shifts[:nb], factors[:nb], specs)
    _lib.check(_lib.lib().wbgpu_stage_times(eng._ctx, ms, calls))
    tot = sum(ms)
    nk = nb * 4096
    print(f"large128_{label} nw=128 {nk} k-points: " +
          " ".join(f"{n}={ms[i]:.2f}" for i, n in enumerate(["fourier", "eigh", "rotate", "identity", "scan"])) +
          f" total={tot:.2f} ms -> {nk / tot * 1e3:.3e} k/s")
    eng.close()
