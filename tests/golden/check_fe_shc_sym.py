#!/usr/bin/env python
"""Container-only check (reads /root/reference, which does not travel to the GPU box; not collected by pytest):
the host path of run() -- symmetry-reduced K-list, K-block sharding, symmetrisation of the rank-3 complex SHC result --
with the oracle standing in for the engine, on the reference's Fe system INCLUDING its spin-current matrices (19 MB,
too large for a fixture), against the reference's own golden files
tests/reference/integrate_files/Fe_W90{,_sym}-opt_SHC{ryoo,qiao}_iter-0000.npz.

    python tests/golden/check_fe_shc_sym.py        # last run: all six comparisons <= 7e-14 relative
"""
import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
REFT = "/root/reference/tests/reference"
import wannierberri_b200 as wb
from wannierberri_b200 import _lib
from oracle import wb_oracle as orc
sys.path.insert(0, "/root/repo/tests")
import _gloo_worker as gw

d = os.path.join(REFT, "systems/Fe_W90")
load = lambda k: np.load(os.path.join(d, k + ".npz"))["arr_0"]
mats = {}
for k in ("Ham", "AA", "SS", "SA", "SHA", "SR", "SH", "SHR"):
    a = load(k)
    mats[k] = np.ascontiguousarray(a.transpose((2, 0, 1) + tuple(range(3, a.ndim))))
fe = wb.System_R(load("real_lattice"), load("iRvec"), load("wannier_centers_cart"))
for k, v in mats.items():
    fe.set_R_mat(k, v)
fe.set_pointgroup(["C4z", "C2x*TimeReversal", "Inversion"])
osys = orc.OracleSystem(fe.rvec.iRvec, fe.real_lattice, fe.wannier_centers_cart, mats)
eng = gw.OracleEngine(osys)
sys.modules["wannierberri_b200.run"].engine_for = lambda system, device=0: eng
p = dict(Efermi=np.array([17.0, 18.0]), omega=np.arange(0.0, 7.1, 1.0), smr_fixed_width=0.20, smr_type="Gaussian")
calcs = dict(ryoo=wb.calculators.dynamic.SHC(SHC_type="ryoo", **p), qiao=wb.calculators.dynamic.SHC(SHC_type="qiao", **p),
             oc=wb.calculators.dynamic.OpticalConductivity(**p))
grid = wb.Grid(fe, NK=[4, 4, 4], NKFFT=[2, 2, 2])
for sym in (False, True):
    res = wb.run(fe, grid, calcs, use_irred_kpt=sym, symmetrize=sym, device=0)
    for q, name in (("ryoo", "opt_SHCryoo"), ("qiao", "opt_SHCqiao"), ("oc", "opt_conductivity")):
        ref = np.load(os.path.join(REFT, "integrate_files", f"Fe_W90{'_sym' if sym else ''}-{name}_iter-0000.npz"))["data"]
        got = res.results[q].data
        print(sym, q, got.shape, ref.shape, np.abs(got - ref).max() / np.abs(ref).max())
