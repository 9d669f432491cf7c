#!/usr/bin/env python
"""A/B of the two accumulation kernels of the optical conductivity (option kubo_method) on BASELINE config 4 in miniature."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wannierberri_b200 as wb  # noqa: E402
from wannierberri_b200 import _lib  # noqa: E402

s32 = wb.synthetic_system(32, rmax=2, seed=20261017)
eng = wb.Engine(s32, device=0)
eng.plan([16, 16, 16], [_lib.IDENTITY, _lib.KUBO], external_terms=True)
shifts, factors = wb.Grid(s32, NKdiv=[8, 8, 8], NKFFT=[16, 16, 16]).K_arrays()
oc = wb.calculators.dynamic.OpticalConductivity(Efermi=np.linspace(-1, 1, 200), omega=np.linspace(0, 5, 500), smr_fixed_width=0.1)
nb = 4
ref = None
for method in (1, 0, 1, 0):
    eng.set_option("kubo_method", method)
    eng.kubo_scan(shifts[:nb], factors[:nb], oc.spec(), oc.Efermi, oc.omega)
    eng.set_option("timing", 1)
    t0 = time.perf_counter()
    out = eng.kubo_scan(shifts[:nb], factors[:nb], oc.spec(), oc.Efermi, oc.omega)
    dt = time.perf_counter() - t0
    ms = (C.c_double * 5)()
    calls = (C.c_int64 * 5)()
    _lib.check(_lib.lib().wbgpu_stage_times(eng._ctx, ms, calls))
    eng.set_option("timing", 0)
    if ref is None:
        ref = out
    print(f"kubo_method={method}: entries+accumulate={ms[4]:.2f} ms rotate={ms[2]:.2f} eigh={ms[1]:.2f} wall={dt * 1e3:.1f} ms -> "
          f"{nb * 4096 / dt:.3e} k/s; max rel diff vs first {np.abs(out - ref).max() / np.abs(ref).max():.1e}", flush=True)
