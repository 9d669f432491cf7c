#!/usr/bin/env python
"""Golden fixture of the Kubo calculators at kBT > 0 (Fermi-Dirac factor, utility.py:172-182) from the UNMODIFIED upstream
reference on its `random` system, NK = 6, NKFFT = 3.

    cd /tmp && PYTHONPATH=/root/reference:/root/repo/oracle/stubs python /root/repo/tests/golden/make_golden_kbt.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import REF, OUT, run_ref, System_R  # noqa: E402
from wannierberri.calculators import dynamic as dyn  # noqa: E402


def main():
    system = System_R.from_npz(path=os.path.join(REF, "tests", "data", "random"), legacy=True)
    p = dict(Efermi=np.linspace(-2, 2, 41), omega=np.arange(0.0, 7.1, 0.5), smr_fixed_width=0.20)
    calcs = dict(optcond=dyn.OpticalConductivity(kBT=0.05, **p), optcond_hot=dyn.OpticalConductivity(kBT=0.5, smr_type="Gaussian", **p),
                 jdos=dyn.JDOS(kBT=0.05, **p), shc_ryoo=dyn.SHC(SHC_type="ryoo", kBT=0.03, **p),
                 shift=dyn.ShiftCurrent(sc_eta=0.1, kBT=0.05, **p), injection=dyn.InjectionCurrent(kBT=0.01, **p),
                 optcond_thresh=dyn.OpticalConductivity(kBT=0.05, degen_thresh=0.3, **p))
    grid, res = run_ref(system, [6, 6, 6], [3, 3, 3], calcs)
    out = dict(NK=np.array([6, 6, 6]), NKFFT=np.array([3, 3, 3]), Efermi=p["Efermi"], omega=p["omega"])
    for q in calcs:
        out[q] = res.results[q].data
        print(q, out[q].shape, out[q].dtype, np.abs(out[q]).max())
    np.savez_compressed(os.path.join(OUT, "golden_random_kbt.npz"), **out)


if __name__ == "__main__":
    sys.exit(main())
