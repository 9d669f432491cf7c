#!/usr/bin/env python
"""Fixture for the tetrahedron method with `hole_like` (inverse Fermi sea, der = -1) and with `Emin` / `Emax`
(grid/tetrahedron.py:197-198, 246-266; calculators/static.py:50-52, 84-91) from the unmodified reference.

    cd /tmp && PYTHONPATH=/root/reference:/root/repo/oracle/stubs python /root/repo/tests/golden/make_golden_tetra_holes.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import OUT, build_fe, run_ref, calc  # noqa: E402

fe = build_fe()
Ef = np.linspace(14.0, 20.0, 31)
st = calc.static
CASES = dict(ahc_holes=("AHC", dict(hole_like=True)), cumdos_holes=("CumDOS", dict(hole_like=True)),
             morb_holes=("Morb", dict(hole_like=True)), ahc_holes_emax=("AHC", dict(hole_like=True, Emax=30.)),
             ahc_emin=("AHC", dict(Emin=12.5)), cumdos_emin=("CumDOS", dict(Emin=11.)),
             ohmic_sea_emin=("Ohmic_FermiSea", dict(Emin=12.5, degen_thresh=0.05)), ahc_plain=("AHC", {}))
calcs = {k: getattr(st, name)(Efermi=Ef, tetra=True, **kw) for k, (name, kw) in CASES.items()}
grid, res = run_ref(fe, [4, 4, 4], [2, 2, 2], calcs)
out = dict(Efermi=Ef, NK=np.array([4, 4, 4]), NKFFT=np.array([2, 2, 2]))
for k in calcs:
    out[k] = res.results[k].data
    print(k, out[k].shape, np.abs(out[k]).max())
np.savez_compressed(os.path.join(OUT, "golden_fe_tetra_holes.npz"), **out)
