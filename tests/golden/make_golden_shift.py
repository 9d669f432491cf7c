#!/usr/bin/env python
"""Golden fixture of the Kubo shift current and injection current (calculators/dynamic.py:244-365) from the UNMODIFIED
upstream reference on its `random` system (no R <-> -R symmetry; AA present), NK = 6, NKFFT = 3.

    cd /tmp && PYTHONPATH=/root/reference:/root/repo/oracle/stubs python /root/repo/tests/golden/make_golden_shift.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import REF, OUT, run_ref, System_R  # noqa: E402
from wannierberri.calculators import dynamic as dyn  # noqa: E402


def main():
    system = System_R.from_npz(path=os.path.join(REF, "tests", "data", "random"), legacy=True)
    p = dict(Efermi=np.linspace(-2, 2, 9), omega=np.arange(0.0, 7.1, 0.5), smr_fixed_width=0.20)
    calcs = dict(shift=dyn.ShiftCurrent(sc_eta=0.1, smr_type="Lorentzian", **p),
                 shift_gauss=dyn.ShiftCurrent(sc_eta=0.04, smr_type="Gaussian", **p),
                 shift_int=dyn.ShiftCurrent(sc_eta=0.1, smr_type="Lorentzian", kwargs_formula=dict(external_terms=False), **p),
                 shift_thresh=dyn.ShiftCurrent(sc_eta=0.1, smr_type="Lorentzian", degen_thresh=0.3, **p),
                 injection=dyn.InjectionCurrent(smr_type="Lorentzian", **p),
                 injection_gauss=dyn.InjectionCurrent(smr_type="Gaussian", **p),
                 injection_int=dyn.InjectionCurrent(smr_type="Lorentzian", kwargs_formula=dict(external_terms=False), **p),
                 injection_thresh=dyn.InjectionCurrent(smr_type="Lorentzian", degen_thresh=0.3, **p))
    grid, res = run_ref(system, [6, 6, 6], [3, 3, 3], calcs)
    out = dict(NK=np.array([6, 6, 6]), NKFFT=np.array([3, 3, 3]), Efermi=p["Efermi"], omega=p["omega"])
    for q in calcs:
        out[q] = res.results[q].data
        print(q, out[q].shape, out[q].dtype, np.abs(out[q]).max())
    np.savez_compressed(os.path.join(OUT, "golden_random_shift.npz"), **out)


if __name__ == "__main__":
    sys.exit(main())
