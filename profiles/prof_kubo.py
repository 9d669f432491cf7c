#!/usr/bin/env python
"""Driver for ncu captures of the Kubo path: synthetic 32-WF model (BASELINE config 4), 500 omega x 200 Efermi.
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/kubo_launches.csv \
       python profiles/prof_kubo.py 1"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wannierberri_b200 as wb  # noqa: E402
from wannierberri_b200 import _lib  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1
nfft = int(sys.argv[2]) if len(sys.argv) > 2 else 16
s32 = wb.synthetic_system(32, rmax=2, seed=20261017)
eng = wb.Engine(s32, device=0)
eng.plan([nfft] * 3, [_lib.IDENTITY, _lib.KUBO], external_terms=True)
grid = wb.Grid(s32, NKdiv=[8, 8, 8], NKFFT=[nfft] * 3)
shifts, factors = grid.K_arrays()
oc = wb.calculators.dynamic.OpticalConductivity(Efermi=np.linspace(-1, 1, 200), omega=np.linspace(0, 5, 500),
                                                smr_fixed_width=0.1)
out = eng.kubo_scan(shifts[:nb], factors[:nb], oc.spec(), oc.Efermi, oc.omega)
print("done", np.abs(out).max())
