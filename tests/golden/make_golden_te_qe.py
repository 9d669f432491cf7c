#!/usr/bin/env python
"""Golden fixture of the reference's Te test (tests/test_run.py:1101-1119: Te 24-WF spinor system, tetrahedron method,
symmetry-reduced K-list + symmetrisation, NK = [3,3,4], NKFFT = [1,1,4]) from the UNMODIFIED upstream reference;
asserts that the live run reproduces the reference's own golden files Te_QE-{dos,cumdos,GME_orb_FermiSurf}_iter-0000.npz.

    cd /tmp && PYTHONPATH=/root/reference:/root/repo/oracle/stubs python /root/repo/tests/golden/make_golden_te_qe.py
"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import REF, OUT, build_te, wberri, calc  # noqa: E402


def main():
    te = build_te()   # pointgroup C3z, C2x, TimeReversal (tests/common_systems.py:1022-1027)
    Ef = np.linspace(4, 8, 11)
    st = calc.static
    calcs = dict(dos=st.DOS(Efermi=Ef, tetra=True), cumdos=st.CumDOS(Efermi=Ef, tetra=True),
                 GME_orb_FermiSurf=st.GME_orb_FermiSurf(Efermi=Ef, tetra=True),
                 GME_spin_FermiSurf=st.GME_spin_FermiSurf(Efermi=Ef, tetra=True),
                 berry_dipole_fsurf=st.BerryDipole_FermiSurf(Efermi=Ef, tetra=True), ahc=st.AHC(Efermi=Ef, tetra=True))
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            grid = wberri.Grid(te, NK=[3, 3, 4], NKFFT=[1, 1, 4])
            res = wberri.run(te, grid=grid, calculators=calcs, parallel=False, use_irred_kpt=True, symmetrize=True,
                             fout_name="g", print_progress_step_time=1e9, print_progress_step_percent=1000)
            K_list = grid.get_K_list(use_symmetry=True)
        finally:
            os.chdir(cwd)
    out = dict(Efermi=Ef, NK=np.array([3, 3, 4]), NKFFT=np.array([1, 1, 4]),
               K_list_Kp_fullBZ=np.array([K.Kp_fullBZ for K in K_list]), K_list_factor=np.array([K.factor for K in K_list]))
    for q in ("dos", "cumdos", "GME_orb_FermiSurf"):
        ref = np.load(os.path.join(REF, "tests/reference/integrate_files", f"Te_QE-{q}_iter-0000.npz"))["data"]
        got = res.results[q].data
        err = np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300)
        print(f"Te_QE-{q}: live reference run vs reference golden file: rel err {err:.2e}")
        assert err < 1e-7, q
        out["upstream_golden_" + q] = ref
    for q in calcs:
        out[q] = res.results[q].data
    np.savez_compressed(os.path.join(OUT, "golden_te_qe.npz"), **out)
    print("written")


if __name__ == "__main__":
    sys.exit(main())
