#!/usr/bin/env python
"""Stage times (CUDA events inside the library) of the next-row-4 / Kubo additions on one GPU:
Te 24-WF Fermi-sea / product formulae (K-blocks of 20^3 of the 200^3 grid), 32-WF shift / injection current, 32-WF SHC.
   python profiles/time_workloads_fsea.py [blocks]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wannierberri_b200 as wb  # noqa: E402
from wannierberri_b200 import _lib  # noqa: E402

st, dyn = wb.calculators.static, wb.calculators.dynamic
blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 4
te = wb.System_R.from_npz(os.path.join(ROOT, "tests", "golden", "te_system.npz"))
EfT = np.linspace(4.0, 8.0, 401)
cases = {
    "te_bd_sea(DerOmega)": dict(a=st.BerryDipole_FermiSea(Efermi=EfT)),
    "te_gme_spin_sea(DerSpin)": dict(a=st.GME_spin_FermiSea(Efermi=EfT)),
    "te_ahc_zeeman_spin(OmegaS)": dict(a=st.AHC_Zeeman_spin(Efermi=EfT)),
    "te_nldrude_fsurf(MassVel)": dict(a=st.NLDrude_FermiSurf(Efermi=EfT)),
    "te_hall_sea(MassMass)": dict(a=st.Hall_classic_FermiSea(Efermi=EfT)),
    "te_shc_static(simple)": dict(a=st.SHC(Efermi=EfT, kwargs_formula=dict(spin_current_type="simple"))),
}
ms = (C.c_double * 5)()
calls = (C.c_int64 * 5)()
for name, calcs in cases.items():
    specs = [s for c in calcs.values() for s in c.specs()]
    eng = wb.Engine(te, device=0)
    eng.plan([20, 20, 20], [s.formula for s in specs], external_terms=True)
    shifts, factors = wb.Grid(te, NKdiv=[10, 10, 10], NKFFT=[20, 20, 20]).K_arrays()
    eng.scan(shifts[:blocks], factors[:blocks], specs)   # warm-up at the same size: the sub-batch buffers are grown here
    eng.set_option("timing", 1)
    eng.scan(shifts[:blocks], factors[:blocks], specs)
    _lib.check(_lib.lib().wbgpu_stage_times(eng._ctx, ms, calls))
    tot = sum(ms)
    nk = blocks * 8000
    print(f"{name:28s} nw=24 {nk} k-points: " +
          " ".join(f"{n}={ms[i]:.2f}" for i, n in enumerate(["fourier", "eigh", "rotate+formula", "identity", "scan"])) +
          f" total={tot:.2f} ms -> {nk / tot * 1e3:.3e} k/s", flush=True)
    eng.close()

s32 = wb.synthetic_system(32, rmax=2, seed=20261017, matrices=("Ham", "AA", "SS"))
shifts, factors = wb.Grid(s32, NKdiv=[8, 8, 8], NKFFT=[16, 16, 16]).K_arrays()
kw = dict(Efermi=np.linspace(-1, 1, 200), omega=np.linspace(0, 5, 500), smr_fixed_width=0.1)
for name, calc in (("kubo32_shift", dyn.ShiftCurrent(sc_eta=0.05, **kw)), ("kubo32_injection", dyn.InjectionCurrent(**kw)),
                   ("kubo32_shc_simple", dyn.SHC(SHC_type="simple", **kw))):
    eng = wb.Engine(s32, device=0)
    spec = calc.spec()
    eng.plan([16, 16, 16], [_lib.IDENTITY, spec.formula_flag], external_terms=True)
    nb = 2
    eng.kubo_scan(shifts[:nb], factors[:nb], spec, calc.Efermi, calc.omega)   # warm-up at the same size
    eng.set_option("timing", 1)
    t0 = time.perf_counter()
    eng.kubo_scan(shifts[:nb], factors[:nb], spec, calc.Efermi, calc.omega)
    dt = time.perf_counter() - t0
    _lib.check(_lib.lib().wbgpu_stage_times(eng._ctx, ms, calls))
    nk = nb * 4096
    print(f"{name:28s} nw=32 {nk} k-points, 500 omega x 200 Ef: " +
          " ".join(f"{n}={ms[i]:.2f}" for i, n in enumerate(["fourier", "eigh", "rotate+matrix", "identity", "entries+accumulate"])) +
          f" wall={dt * 1e3:.1f} ms -> {nk / dt:.3e} k/s", flush=True)
    eng.close()
