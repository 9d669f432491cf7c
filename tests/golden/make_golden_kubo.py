#!/usr/bin/env python
"""Golden fixtures of the Kubo path (DynamicCalculator: OpticalConductivity, JDOS) from the UNMODIFIED
upstream reference.  Same conventions as make_golden.py (run in the build container only):

    cd /tmp && PYTHONPATH=/root/reference:/root/repo/oracle/stubs \
        python /root/repo/tests/golden/make_golden_kubo.py

Asserts that the live reference run reproduces the reference's OWN golden file
tests/reference/integrate_files/Fe_W90-opt_conductivity_iter-0000.npz, then writes golden_fe_kubo.npz.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import REF, OUT, build_fe, run_ref, wberri, Data_K_R  # noqa: E402
from wannierberri.calculators import dynamic as dyn  # noqa: E402


def main():
    fe = build_fe()
    out = {}
    # ---- the reference's own test (tests/test_run.py:290-293)
    p_ref = dict(Efermi=np.array([17.0, 18.0]), omega=np.arange(0.0, 7.1, 1.0), smr_fixed_width=0.20, smr_type="Gaussian")
    # ---- wider scans, both smearing types, degenerate groups, internal terms only
    Ef = np.linspace(12.0, 22.0, 21)
    om = np.linspace(0.0, 5.0, 11)
    p_lor = dict(Efermi=Ef, omega=om, smr_fixed_width=0.1, smr_type="Lorentzian")
    calcs = dict(
        ref_optcond=dyn.OpticalConductivity(**p_ref),
        lor_optcond=dyn.OpticalConductivity(**p_lor),
        lor_optcond_thresh=dyn.OpticalConductivity(degen_thresh=0.05, **p_lor),
        lor_optcond_int=dyn.OpticalConductivity(kwargs_formula=dict(external_terms=False), **p_lor),
        gau_optcond=dyn.OpticalConductivity(Efermi=Ef, omega=om, smr_fixed_width=0.15, smr_type="Gaussian"),
        lor_jdos=dyn.JDOS(**p_lor),
        gau_jdos=dyn.JDOS(Efermi=Ef, omega=om, smr_fixed_width=0.15, smr_type="Gaussian"),
    )
    grid, res = run_ref(fe, [4, 4, 4], [2, 2, 2], calcs)
    ref = np.load(os.path.join(REF, "tests/reference/integrate_files", "Fe_W90-opt_conductivity_iter-0000.npz"))["data"]
    got = res.results["ref_optcond"].data
    err = np.abs(got - ref).max() / np.abs(ref).max()
    print(f"Fe_W90-opt_conductivity: live reference run vs reference golden file: rel err {err:.2e}")
    assert err < 1e-8
    out["upstream_golden_opt_conductivity"] = ref
    out.update(NK=np.array([4, 4, 4]), NKFFT=np.array([2, 2, 2]), ref_Efermi=p_ref["Efermi"], ref_omega=p_ref["omega"],
               Efermi=Ef, omega=om)
    for q in calcs:
        out["run_" + q] = res.results[q].data
    # ---- one K-block (the block of golden_fe_block.npz)
    grid = wberri.Grid(fe, NK=[6, 6, 6], NKFFT=[3, 3, 3])
    Kp = grid.get_K_list(use_symmetry=False)[5]
    data = Data_K_R(fe, dK=Kp.Kp_fullBZ, grid=grid, Kpoint=Kp, fftlib="numpy")
    out.update(block_dK=np.array(Kp.Kp_fullBZ), block_NKFFT=np.array([3, 3, 3]))
    for q, c in calcs.items():
        out["block_" + q] = c(data).data
    np.savez_compressed(os.path.join(OUT, "golden_fe_kubo.npz"), **out)
    print("written", os.path.join(OUT, "golden_fe_kubo.npz"))


if __name__ == "__main__":
    sys.exit(main())
