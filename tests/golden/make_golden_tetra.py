#!/usr/bin/env python
"""Golden fixtures of the tetrahedron method (StaticCalculator(tetra=True) on grid K-blocks: TetraWeightsParal)
from the UNMODIFIED upstream reference.  Same conventions as make_golden.py (build container only):

    cd /tmp && PYTHONPATH=/root/reference:/root/repo/oracle/stubs \
        python /root/repo/tests/golden/make_golden_tetra.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import OUT, build_fe, run_ref, wberri, Data_K_R, calc  # noqa: E402


def main():
    fe = build_fe()
    Ef = np.linspace(12.0, 22.0, 101)
    st = calc.static
    calcs = dict(
        ahc=st.AHC(Efermi=Ef, tetra=True), dos=st.DOS(Efermi=Ef, tetra=True), cumdos=st.CumDOS(Efermi=Ef, tetra=True),
        Morb=st.Morb(Efermi=Ef, tetra=True), spin=st.Spin(Efermi=Ef, tetra=True),
        ahc_thresh=st.AHC(Efermi=Ef, tetra=True, degen_thresh=0.05),
        bcd=st.BerryDipole_FermiSurf(Efermi=Ef, tetra=True),
        gme_spin=st.GME_spin_FermiSurf(Efermi=Ef, tetra=True),
    )
    out = dict(Efermi=Ef)
    # ---- one K-block (the block of golden_fe_block.npz) incl. the corner energies
    grid = wberri.Grid(fe, NK=[6, 6, 6], NKFFT=[3, 3, 3])
    Kp = grid.get_K_list(use_symmetry=False)[5]
    data = Data_K_R(fe, dK=Kp.Kp_fullBZ, grid=grid, Kpoint=Kp, fftlib="numpy")
    out.update(block_dK=np.array(Kp.Kp_fullBZ), block_NKFFT=np.array([3, 3, 3]), block_NKdiv=np.array([2, 2, 2]),
               block_dK_cell=np.array(Kp.dK_fullBZ), block_E_corners=np.array(data.E_K_corners_parallel()))
    for q, c in calcs.items():
        out["block_" + q] = c(data).data
    # ---- a whole (small) grid through run()
    grid, res = run_ref(fe, [4, 4, 4], [2, 2, 2], calcs)
    out.update(NK=np.array([4, 4, 4]), NKFFT=np.array([2, 2, 2]))
    for q in calcs:
        out["run_" + q] = res.results[q].data
    np.savez_compressed(os.path.join(OUT, "golden_fe_tetra.npz"), **out)
    print("written", os.path.join(OUT, "golden_fe_tetra.npz"))


if __name__ == "__main__":
    sys.exit(main())
