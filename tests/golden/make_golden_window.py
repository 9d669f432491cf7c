#!/usr/bin/env python
"""Golden fixture for the Emin / Emax / hole_like arguments of StaticCalculator (calculators/static.py:26-52, 80-97)
from the UNMODIFIED upstream reference on the seeded 6-WF synthetic model of make_golden_adpt.py:

    cd /tmp && PYTHONPATH=/root/reference:/root/repo/oracle/stubs:/root/repo \
        python /root/repo/tests/golden/make_golden_window.py

Without the tetrahedron method the reference passes Emin / Emax on to Data_K.get_bands_in_range_groups, which does
not read them (data_K/data_K.py:172-185): the script asserts that, so the fixture also pins "no effect".
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import OUT, run_ref, System_R, Rvectors, calc  # noqa: E402
from wannierberri_b200.system import synthetic_system  # noqa: E402  (array generator only)


def main():
    g = synthetic_system(6, rmax=1, seed=4242)
    system = System_R(silent=True)
    system.set_real_lattice(g.real_lattice)
    system.num_wann = g.num_wann
    system.wannier_centers_cart = g.wannier_centers_cart
    system.rvec = Rvectors(g.real_lattice, iRvec=g.rvec.iRvec, shifts_left_red=system.wannier_centers_red)
    for k in ("Ham", "AA"):
        system.set_R_mat(k, g.get_R_mat(k))
    system._NKFFT_recommended = np.array([3, 3, 3])
    system.set_pointgroup([])
    Ef = np.linspace(-3., 3., 31)
    st = calc.static
    win = dict(Emin=-1., Emax=1.)
    calcs = dict(ahc=st.AHC(Efermi=Ef), cumdos=st.CumDOS(Efermi=Ef), dos=st.DOS(Efermi=Ef),
                 ahc_win_hole=st.AHC(Efermi=Ef, hole_like=True, **win),
                 cumdos_win_hole=st.CumDOS(Efermi=Ef, hole_like=True, **win),
                 dos_win_hole=st.DOS(Efermi=Ef, hole_like=True, **win),
                 ahc_tetra=st.AHC(Efermi=Ef, tetra=True), ahc_tetra_Emax=st.AHC(Efermi=Ef, tetra=True, Emax=1.),
                 dos_tetra_win=st.DOS(Efermi=Ef, tetra=True, **win))
    grid, res = run_ref(system, [6, 6, 6], [3, 3, 3], calcs)
    out = dict(Efermi=Ef, Emin=-1., Emax=1.)
    for q in calcs:
        out[q] = res.results[q].data
    assert np.array_equal(out["ahc_win_hole"], -out["ahc"])          # hole_like: sign of a Fermi-sea quantity
    assert np.array_equal(out["cumdos_win_hole"], -out["cumdos"])
    assert np.array_equal(out["dos_win_hole"], out["dos"])           # Fermi surface: no sign
    assert np.array_equal(out["ahc_tetra_Emax"], out["ahc_tetra"])   # Emax acts on the inverse Fermi sea only
    np.savez_compressed(os.path.join(OUT, "golden_synth_window.npz"), **out)
    print("written", os.path.join(OUT, "golden_synth_window.npz"))


if __name__ == "__main__":
    sys.exit(main())
