#!/usr/bin/env python
"""Golden fixture of the symmetric run mode (the reference's default: use_irred_kpt=True, symmetrize=True) from the
UNMODIFIED upstream reference; asserts that the live run reproduces the reference's own golden files
tests/reference/integrate_files/Fe_W90_sym-{ahc,dos,cumdos,Morb,spin}_iter-0000.npz.

    cd /tmp && PYTHONPATH=/root/reference:/root/repo/oracle/stubs \
        python /root/repo/tests/golden/make_golden_sym.py
"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import REF, OUT, build_fe, wberri, calc  # noqa: E402
from wannierberri.calculators import dynamic as dyn  # noqa: E402


def main():
    fe = build_fe()   # pointgroup: C4z, C2x*TimeReversal, Inversion (tests/common_systems.py:186)
    Ef = np.linspace(17, 18, 11)
    st = calc.static
    calcs = dict(ahc=st.AHC(Efermi=Ef), dos=st.DOS(Efermi=Ef), cumdos=st.CumDOS(Efermi=Ef), Morb=st.Morb(Efermi=Ef),
                 spin=st.Spin(Efermi=Ef), berry_dipole_fsurf=st.BerryDipole_FermiSurf(Efermi=Ef),
                 gme_orb_fsurf=st.GME_orb_FermiSurf(Efermi=Ef), gme_spin_fsurf=st.GME_spin_FermiSurf(Efermi=Ef),
                 ahc_tetra=st.AHC(Efermi=Ef, tetra=True), dos_tetra=st.DOS(Efermi=Ef, tetra=True),
                 opt_conductivity=dyn.OpticalConductivity(Efermi=np.array([17.0, 18.0]), omega=np.arange(0.0, 7.1, 1.0),
                                                          smr_fixed_width=0.20, smr_type="Gaussian"))
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            grid = wberri.Grid(fe, NK=[4, 4, 4], NKFFT=[2, 2, 2])
            res = wberri.run(fe, grid=grid, calculators=calcs, parallel=False, use_irred_kpt=True, symmetrize=True,
                             fout_name="g", print_progress_step_time=1e9, print_progress_step_percent=1000)
            K_list = grid.get_K_list(use_symmetry=True)
        finally:
            os.chdir(cwd)
    out = dict(Efermi=Ef, NK=np.array([4, 4, 4]), NKFFT=np.array([2, 2, 2]),
               K_list_Kp_fullBZ=np.array([K.Kp_fullBZ for K in K_list]), K_list_factor=np.array([K.factor for K in K_list]),
               opt_Efermi=np.array([17.0, 18.0]), opt_omega=np.arange(0.0, 7.1, 1.0))
    for q in ("ahc", "dos", "cumdos", "Morb", "spin", "opt_conductivity"):
        ref = np.load(os.path.join(REF, "tests/reference/integrate_files", f"Fe_W90_sym-{q}_iter-0000.npz"))["data"]
        got = res.results[q].data
        err = np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300)
        print(f"Fe_W90_sym-{q}: live reference run vs reference golden file: rel err {err:.2e}")
        assert err < 1e-8, q
        out["upstream_golden_" + q] = ref
    for q in calcs:
        out[q] = res.results[q].data
    np.savez_compressed(os.path.join(OUT, "golden_fe_sym.npz"), **out)
    print("written", os.path.join(OUT, "golden_fe_sym.npz"))


if __name__ == "__main__":
    sys.exit(main())
