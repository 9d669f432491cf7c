#!/usr/bin/env python
"""Fixture of the Data_K_soc test: a synthetic SystemSOC (random scalar up / down systems of 3 Wannier functions with
DIFFERENT R-vector sets, a random hermitian spin-orbit term and spin matrix on a third set) evaluated by the unmodified
reference, whose run() picks Data_K_soc for it (data_K/__init__.py:10-18).  The Fe_gpaw SOC systems of the reference's
own tests need gpaw, which is not installed here.

    cd /tmp && PYTHONPATH=/root/reference:/root/repo/oracle/stubs python /root/repo/tests/golden/make_golden_soc.py
"""
import os

import numpy as np

import wannierberri as wberri
from wannierberri.system.system_R import System_R
from wannierberri.system.system_soc import SystemSOC
from wannierberri.fourier.rvectors import Rvectors
from wannierberri.data_K import get_data_k_class_from_system, Data_K_soc

OUT = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(20261018)
lattice = np.array([[3.1, 0.2, 0.0], [0.1, 2.9, 0.3], [0.0, 0.4, 3.3]])
nws = 3


def rset(rmax, drop):
    R = [(x, y, z) for x in range(-rmax, rmax + 1) for y in range(-rmax, rmax + 1) for z in range(-1, 2)]
    R = [r for r in R if r not in drop and tuple(-np.array(r)) not in drop]
    return np.array(R, dtype=int)


def herm_field(iRvec, nw, ncart, scale):
    idx = {tuple(R): i for i, R in enumerate(iRvec)}
    shape = (len(iRvec), nw, nw) + ((3,) if ncart else ())
    X = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) * scale
    X *= np.exp(-np.linalg.norm(iRvec, axis=1)).reshape((-1,) + (1,) * (len(shape) - 1))
    Xd = np.empty_like(X)
    for R, i in idx.items():
        Xd[i] = X[idx[tuple(-np.array(R))]].swapaxes(0, 1).conj()
    return 0.5 * (X + Xd)


def scalar_system(iRvec, centres):
    s = System_R(silent=True)
    s.set_real_lattice(lattice)
    s.num_wann = nws
    s.wannier_centers_cart = centres
    s.rvec = Rvectors(lattice, iRvec=iRvec, shifts_left_red=s.wannier_centers_red)
    s.set_R_mat("Ham", herm_field(iRvec, nws, 0, 1.0))
    s.set_R_mat("AA", herm_field(iRvec, nws, 1, 0.1))
    return s


c_up = rng.random((nws, 3)) @ lattice
c_dw = c_up + 0.05 * rng.standard_normal((nws, 3))
R_up, R_dw, R_soc = rset(1, {(1, 1, 1)}), rset(1, {(1, -1, 0)}), rset(1, {(1, 0, 1), (0, 1, -1)})
up, dw = scalar_system(R_up, c_up), scalar_system(R_dw, c_dw)
soc = SystemSOC(system_up=up, system_down=dw, silent=True)
soc.rvec = Rvectors(lattice, iRvec=R_soc, shifts_left_red=soc.wannier_centers_red)
soc.set_R_mat("Ham_SOC", herm_field(R_soc, 2 * nws, 0, 0.3))
soc.set_R_mat("SS", herm_field(R_soc, 2 * nws, 1, 0.5))
soc.has_soc = True
soc.set_pointgroup(symmetry_gen=[])
assert get_data_k_class_from_system(soc) is Data_K_soc

Efermi = np.linspace(-1.5, 1.5, 13)
calc = wberri.calculators.static
calcs = dict(dos=calc.DOS(Efermi=Efermi), cumdos=calc.CumDOS(Efermi=Efermi), ahc=calc.AHC(Efermi=Efermi),
             ahc_int=calc.AHC(Efermi=Efermi, kwargs_formula=dict(external_terms=False)), spin=calc.Spin(Efermi=Efermi),
             ohmic=calc.Ohmic_FermiSea(Efermi=Efermi), gme_spin=calc.GME_spin_FermiSurf(Efermi=Efermi))
grid = wberri.Grid(soc, NK=[6, 6, 4], NKFFT=[3, 3, 2])
res = wberri.run(soc, grid=grid, calculators=calcs, adpt_num_iter=0, use_irred_kpt=False, symmetrize=False, parallel=False,
                 fout_name="/tmp/soc", print_progress_step_time=1e9)
out = dict(Efermi=Efermi, NK=np.array([6, 6, 4]), NKFFT=np.array([3, 3, 2]), real_lattice=lattice, centres_up=c_up, centres_dw=c_dw,
           iRvec_up=R_up, iRvec_dw=R_dw, iRvec_soc=R_soc, Ham_SOC=soc.get_R_mat("Ham_SOC"), SS=soc.get_R_mat("SS"))
for tag, s in (("up", up), ("dw", dw)):
    for key in ("Ham", "AA"):
        out[f"{key}_{tag}"] = s.get_R_mat(key)
for key in calcs:
    out["res_" + key] = res.results[key].data
    print(key, out["res_" + key].shape, np.abs(out["res_" + key]).max())
np.savez_compressed(os.path.join(OUT, "golden_soc.npz"), **out)
