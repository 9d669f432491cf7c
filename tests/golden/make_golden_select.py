#!/usr/bin/env python
"""Fixture for `select_bands` (calculators/static.py:93-100, 129-136): the unmodified reference on the Fe 18-WF system,
Fermi-surface calculators restricted to a set of bands, with a degeneracy threshold that makes multi-band groups.

    cd /tmp && PYTHONPATH=/root/reference:/root/repo/oracle/stubs python /root/repo/tests/golden/make_golden_select.py
"""
import os
import sys

import numpy as np

OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, OUT)
import wannierberri as wberri  # noqa: E402
from make_golden import build_fe  # noqa: E402

system = build_fe()
Efermi = np.linspace(15.0, 19.0, 21)
grid = wberri.Grid(system, NK=[4, 4, 4], NKFFT=[2, 2, 2])
calc = wberri.calculators.static
sel = np.array([3, 4, 7, 8, 11])
calcs = dict(ohmic_sel=calc.Ohmic_FermiSurf(Efermi=Efermi, select_bands=sel, degen_thresh=0.3),
             dos_sel=calc.DOS(Efermi=Efermi, select_bands=sel),
             bcd_sel=calc.BerryDipole_FermiSurf(Efermi=Efermi, select_bands=sel, degen_thresh=0.3, degen_Kramers=True),
             gme_sel=calc.GME_orb_FermiSurf(Efermi=Efermi, select_bands=np.array([5])),
             ohmic_all=calc.Ohmic_FermiSurf(Efermi=Efermi, degen_thresh=0.3))
res = wberri.run(system, grid=grid, calculators=calcs, adpt_num_iter=0, use_irred_kpt=False, symmetrize=False,
                 parallel=False, fout_name="/tmp/select", print_progress_step_time=1e9)
out = dict(Efermi=Efermi, NK=np.array([4, 4, 4]), NKFFT=np.array([2, 2, 2]), select=np.array(sel))
for key in calcs:
    out[key] = res.results[key].data
    print(key, out[key].shape, np.abs(out[key]).max())
np.savez_compressed(os.path.join(OUT, "golden_select.npz"), **out)
