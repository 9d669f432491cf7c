#!/usr/bin/env python
"""Golden fixture of Ohmic_FermiSurf (formula VelVel) from the UNMODIFIED upstream reference; asserts that the live
run reproduces the reference's own golden file Fe_W90-conductivity_ohmic_fsurf_iter-0000.npz.

    cd /tmp && PYTHONPATH=/root/reference:/root/repo/oracle/stubs \
        python /root/repo/tests/golden/make_golden_ohmic.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import REF, OUT, build_fe, run_ref, calc  # noqa: E402


def main():
    fe = build_fe()
    Ef = np.linspace(17, 18, 11)
    calcs = dict(ohmic_fsurf=calc.static.Ohmic_FermiSurf(Efermi=Ef),
                 ohmic_fsurf_thresh=calc.static.Ohmic_FermiSurf(Efermi=Ef, degen_thresh=0.05),
                 ohmic_fsurf_tetra=calc.static.Ohmic_FermiSurf(Efermi=Ef, tetra=True))
    grid, res = run_ref(fe, [4, 4, 4], [2, 2, 2], calcs)
    ref = np.load(os.path.join(REF, "tests/reference/integrate_files", "Fe_W90-conductivity_ohmic_fsurf_iter-0000.npz"))["data"]
    got = res.results["ohmic_fsurf"].data
    err = np.abs(got - ref).max() / np.abs(ref).max()
    print(f"Fe_W90-conductivity_ohmic_fsurf: live reference run vs reference golden file: rel err {err:.2e}")
    assert err < 1e-8
    out = dict(Efermi=Ef, NK=np.array([4, 4, 4]), NKFFT=np.array([2, 2, 2]), upstream_golden_ohmic_fsurf=ref)
    for q in calcs:
        out[q] = res.results[q].data
    np.savez_compressed(os.path.join(OUT, "golden_fe_ohmic.npz"), **out)
    print("written")


if __name__ == "__main__":
    sys.exit(main())
