#!/usr/bin/env python
"""Fixture of the second-order calculators: NLDrude_Zeeman_{spin, orb_Omega, orb},
eMChA_FermiSurf, QuantumMetric_FermiSea / QuantumMetric_Vel_DQ evaluated by the unmodified reference (`wannierberri.run`
on its own Data_K_R, formulae Der2Spin / Der2Omega / Der2Morb / emcha_surf / tildeFab / tildeFab_d) on the Fe 18-WF
system, with and without external terms, and with wide band groups (degen_thresh = 0.3 eV, Kramers pairs).

    cd /tmp && PYTHONPATH=/root/reference:/root/repo/oracle/stubs python /root/repo/tests/golden/make_golden_second_order.py
"""
import os
import sys

import numpy as np

OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, OUT)
sys.path.insert(0, os.path.dirname(OUT))

import wannierberri as wberri  # noqa: E402
from wannierberri.calculators import static as rst  # noqa: E402
from make_golden import build_fe  # noqa: E402
from second_order_calcs import make_calculators  # noqa: E402


if __name__ == "__main__":
    system = build_fe()
    Efermi = np.linspace(15.0, 19.0, 9)
    grid = wberri.Grid(system, NK=[4, 4, 4], NKFFT=[2, 2, 2])
    calcs = make_calculators(rst, Efermi)
    res = wberri.run(system, grid=grid, calculators=calcs, adpt_num_iter=0, use_irred_kpt=False, symmetrize=False,
                     parallel=False, fout_name=os.path.join("/tmp", "second_order"), print_progress_step_time=1e9)
    out = dict(Efermi=Efermi, NK=np.array([4, 4, 4]), NKFFT=np.array([2, 2, 2]))
    for key in calcs:
        out[key] = res.results[key].data
        out[key + "_TR"] = res.results[key].transformTR.factor
        out[key + "_Inv"] = res.results[key].transformInv.factor
        print(key, out[key].shape, np.abs(out[key]).max())
    np.savez_compressed(os.path.join(OUT, "golden_second_order.npz"), **out)
