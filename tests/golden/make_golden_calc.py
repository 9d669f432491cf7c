#!/usr/bin/env python
"""Collect the reference's own calculator regression files tests/reference/calculators/calculator-Fe-ident-fder={0..3}.npz
(tests/test_calc.py:71-77: StaticCalculator(Formula=Identity, fder=0..3) on one K-block dK=[0.1,0.2,0.3], NKFFT=3^3) into
one fixture, after checking that a live run of the unmodified reference reproduces them.

    cd /tmp && PYTHONPATH=/root/reference:/root/repo/oracle/stubs python /root/repo/tests/golden/make_golden_calc.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import REF, OUT, build_fe, wberri, Data_K_R, calc  # noqa: E402
from wannierberri.formula import covariant as frml  # noqa: E402


def main():
    fe = build_fe()
    Ef = np.linspace(17, 18, 11)
    grid = wberri.Grid(system=fe, NKFFT=[3, 3, 3], NKdiv=1)
    data = Data_K_R(fe, dK=[0.1, 0.2, 0.3], grid=grid, fftlib="numpy")
    out = dict(Efermi=Ef, dK=np.array([0.1, 0.2, 0.3]), NKFFT=np.array([3, 3, 3]))
    for fder in range(4):
        ref = np.load(os.path.join(REF, "tests/reference/calculators", f"calculator-Fe-ident-fder={fder}.npz"))["data"]
        got = calc.static.StaticCalculator(Formula=frml.Identity, Efermi=Ef, tetra=False, fder=fder)(data).data
        err = np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300)
        print(f"calculator-Fe-ident-fder={fder}: live reference vs reference file: rel err {err:.2e}")
        assert err < 1e-8
        out[f"upstream_ident_fder{fder}"] = ref
    np.savez_compressed(os.path.join(OUT, "golden_fe_calc_fder.npz"), **out)
    print("written")


if __name__ == "__main__":
    sys.exit(main())
