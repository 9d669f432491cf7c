#!/usr/bin/env python
"""Golden fixture of adaptive refinement (run(adpt_num_iter > 0), no symmetry) from the UNMODIFIED upstream
reference, on a seeded synthetic model WITHOUT symmetry (so that no two K-points tie in the selection):

    cd /tmp && PYTHONPATH=/root/reference:/root/repo/oracle/stubs:/root/repo \
        python /root/repo/tests/golden/make_golden_adpt.py
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
    Ef = np.linspace(-3., 3., 61)
    calcs = dict(ahc=calc.static.AHC(Efermi=Ef), dos=calc.static.DOS(Efermi=Ef))
    out = dict(Efermi=Ef, NK=np.array([6, 6, 6]), NKFFT=np.array([3, 3, 3]), seed=4242, num_wann=6)
    for n_iter in (0, 1, 3):
        grid, res = run_ref(system, [6, 6, 6], [3, 3, 3], calcs, adpt_num_iter=n_iter, adpt_fac=2, adpt_mesh=2)
        for q in calcs:
            out[f"iter{n_iter}_{q}"] = res.results[q].data
    np.savez_compressed(os.path.join(OUT, "golden_synth_adpt.npz"), **out)
    print("written", os.path.join(OUT, "golden_synth_adpt.npz"))
    # refinement driven by tetrahedron-method and Kubo calculators (per-K-point cells shrink with the refinement)
    from wannierberri.calculators import dynamic as dyn
    Ef2 = np.linspace(-3., 3., 13)
    calcs = dict(ahc_tetra=calc.static.AHC(Efermi=Ef2, tetra=True), dos_tetra=calc.static.DOS(Efermi=Ef2, tetra=True),
                 cumdos=calc.static.CumDOS(Efermi=Ef2),
                 optcond=dyn.OpticalConductivity(Efermi=Ef2[::4], omega=np.linspace(0., 4., 9), smr_fixed_width=0.2))
    out = dict(Efermi=Ef2, omega=np.linspace(0., 4., 9))
    for n_iter in (0, 2):
        grid, res = run_ref(system, [6, 6, 6], [3, 3, 3], calcs, adpt_num_iter=n_iter, adpt_fac=2, adpt_mesh=2)
        for q in calcs:
            out[f"iter{n_iter}_{q}"] = res.results[q].data
    np.savez_compressed(os.path.join(OUT, "golden_synth_adpt_tetra_kubo.npz"), **out)


if __name__ == "__main__":
    sys.exit(main())
