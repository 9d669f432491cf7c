#!/usr/bin/env python
"""Fixture of the reference's `random` system (tests/data/random: 6 WF, 20 random R-vectors WITHOUT the R <-> -R
symmetry, all matrices) and its own golden files tests/reference/integrate_files/random-*_iter-0000.npz
(tests/test_run.py:653-669): a live run of the unmodified reference must reproduce them; the system arrays and the
golden data are stored for the oracle and GPU parity tests.

    cd /tmp && PYTHONPATH=/root/reference:/root/repo/oracle/stubs python /root/repo/tests/golden/make_golden_random.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import REF, OUT, run_ref, dump_system, System_R, calc  # noqa: E402
from wannierberri.calculators import dynamic as dyn  # noqa: E402


def main():
    system = System_R.from_npz(path=os.path.join(REF, "tests", "data", "random"), legacy=True)
    dump_system(system, "random_system.npz", ("Ham", "AA", "BB", "CC", "SS"))
    Ef = np.linspace(-2, 2, 5)
    st = calc.static
    calcs = dict(ahc=st.AHC(Efermi=Ef), dos=st.DOS(Efermi=Ef), cumdos=st.CumDOS(Efermi=Ef), Morb=st.Morb(Efermi=Ef),
                 spin=st.Spin(Efermi=Ef), conductivity_ohmic_fsurf=st.Ohmic_FermiSurf(Efermi=Ef),
                 opt_conductivity=dyn.OpticalConductivity(Efermi=np.array([17.0, 18.0]), omega=np.arange(0.0, 7.1, 1.0),
                                                          smr_fixed_width=0.20, smr_type="Gaussian"),
                 opt_conductivity_in=dyn.OpticalConductivity(Efermi=np.linspace(-2, 2, 9), omega=np.arange(0.0, 7.1, 1.0),
                                                             smr_fixed_width=0.20, smr_type="Lorentzian"),
                 berry_dipole_fsurf=st.BerryDipole_FermiSurf(Efermi=Ef), gme_orb_fsurf=st.GME_orb_FermiSurf(Efermi=Ef),
                 gme_spin_fsurf=st.GME_spin_FermiSurf(Efermi=Ef), ahc_tetra=st.AHC(Efermi=Ef, tetra=True))
    grid, res = run_ref(system, [6, 6, 6], [3, 3, 3], calcs)
    out = dict(Efermi=Ef, NK=np.array([6, 6, 6]), NKFFT=np.array([3, 3, 3]), opt_Efermi=np.array([17.0, 18.0]),
               opt_omega=np.arange(0.0, 7.1, 1.0), opt_in_Efermi=np.linspace(-2, 2, 9))
    for q in ("ahc", "dos", "cumdos", "Morb", "spin", "conductivity_ohmic_fsurf", "opt_conductivity"):
        ref = np.load(os.path.join(REF, "tests/reference/integrate_files", f"random-{q}_iter-0000.npz"))["data"]
        got = res.results[q].data
        err = np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300)
        print(f"random-{q}: live reference run vs reference golden file: rel err {err:.2e} (max |ref| {np.abs(ref).max():.3e})")
        assert err < 1e-8 or np.abs(ref).max() < 1e-12, q
        out["upstream_golden_" + q] = ref
    for q in calcs:
        out[q] = res.results[q].data
    np.savez_compressed(os.path.join(OUT, "golden_random.npz"), **out)
    print("written")


if __name__ == "__main__":
    sys.exit(main())
