#!/usr/bin/env python
"""BASELINE configurations 3, 4 and 5 END TO END through run() on one GPU (configs 1 and 2: full_configs.py):
  config 3   Te 24-WF, BerryDipole_FermiSurf + GME_orb_FermiSurf + GME_spin_FermiSurf, 200^3 grid (NKdiv 10 x NKFFT 20),
             401 Fermi levels 4..8 eV -- full size
  config 4   synthetic 32-WF model (nR = 125), Kubo optical conductivity, 500 frequencies x 200 Fermi levels, 128^3 grid
             (NKdiv 8 x NKFFT 16) -- full size
  config 5   synthetic 128-WF model with ~4000 R-vectors (the shortest vectors of a cubic lattice, closed under R -> -R),
             eigenvalues + d_a H rotations: DOS + CumDOS + Ohmic_FermiSurf, K-blocks of 32^3 from the 512^3 grid
             (NKdiv 16): `blocks` K-blocks are run, the full grid is 4096 of them (per-block cost is constant)
    python profiles/full_configs_345.py [blocks5]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wannierberri_b200 as wb  # noqa: E402

st, dyn = wb.calculators.static, wb.calculators.dynamic


def timed_run(system, grid, calcs, **kw):
    wb.run(system, wb.Grid(system, NKdiv=[1, 1, 1], NKFFT=grid.FFT), calcs, **kw)   # plan + buffers
    t0 = time.perf_counter()
    res = wb.run(system, grid, calcs, **kw)
    return res, time.perf_counter() - t0


def main():
    blocks5 = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    te = wb.System_R.from_npz(os.path.join(ROOT, "tests", "golden", "te_system.npz"))
    Ef = np.linspace(4., 8., 401)
    calcs = dict(bcd=st.BerryDipole_FermiSurf(Efermi=Ef), gme_orb=st.GME_orb_FermiSurf(Efermi=Ef), gme_spin=st.GME_spin_FermiSurf(Efermi=Ef))
    res, dt = timed_run(te, wb.Grid(te, NKdiv=[10, 10, 10], NKFFT=[20, 20, 20]), calcs)
    print(f"config 3: Te BCD + GME (Fermi surface) 200^3 (8.0e6 k-points, 401 E_F): GPU run() {dt:.2f} s = {8.0e6 / dt:.3e} k-points/s; "
          f"max |D| = {np.abs(res.results['bcd'].data).max():.6e}", flush=True)

    s32 = wb.synthetic_system(32, rmax=2, seed=20261017)
    oc = dict(oc=dyn.OpticalConductivity(Efermi=np.linspace(-1, 1, 200), omega=np.linspace(0, 5, 500), smr_fixed_width=0.1))
    res, dt = timed_run(s32, wb.Grid(s32, NKdiv=[8, 8, 8], NKFFT=[16, 16, 16]), oc)
    nk = 128 ** 3
    print(f"config 4: 32-WF Kubo optical conductivity 128^3 ({nk} k-points, 500 omega x 200 E_F): GPU run() {dt:.2f} s = "
          f"{nk / dt:.3e} k-points/s; max |sigma| = {np.abs(res.results['oc'].data).max():.6e}", flush=True)

    # ---- config 5: ~4000 shortest R-vectors, 128 WF
    nw, rng = 128, np.random.default_rng(20261017)
    rr = np.arange(-10, 11)
    allR = np.array([[x, y, z] for x in rr for y in rr for z in rr], dtype=int)
    n2 = (allR ** 2).sum(axis=1)
    cut = np.sort(n2)[4000]   # whole shells only: the set stays closed under R -> -R
    iRvec = allR[n2 < cut]
    nR = len(iRvec)
    index = {tuple(R): i for i, R in enumerate(iRvec)}
    minus = np.array([index[tuple(-R)] for R in iRvec])
    t0 = time.perf_counter()
    H = (rng.standard_normal((nR, nw, nw)) + 1j * rng.standard_normal((nR, nw, nw))) * np.exp(-np.sqrt(n2[n2 < cut]))[:, None, None]
    H = 0.5 * (H + H[minus].transpose(0, 2, 1).conj())
    s128 = wb.System_R(np.eye(3) * 4.0, iRvec, rng.random((nw, 3)) * 4.0)
    s128.set_R_mat("Ham", H)
    print(f"config 5: model built on the host in {time.perf_counter() - t0:.1f} s: nw = {nw}, nR = {nR}", flush=True)
    Ef5 = np.linspace(-4, 4, 401)
    calcs = dict(dos=st.DOS(Efermi=Ef5), cumdos=st.CumDOS(Efermi=Ef5), ohmic=st.Ohmic_FermiSurf(Efermi=Ef5))
    eng = wb.Engine(s128)
    specs = [s for c in calcs.values() for s in c.specs()]
    eng.plan([32, 32, 32], [s.formula for s in specs])
    shifts, factors = wb.Grid(s128, NKdiv=[16, 16, 16], NKFFT=[32, 32, 32]).K_arrays()
    eng.scan(shifts[:1], factors[:1], specs)
    t0 = time.perf_counter()
    out = eng.scan(shifts[:blocks5], factors[:blocks5], specs)
    dt = time.perf_counter() - t0
    nk = blocks5 * 32 ** 3
    cum = out[1] * s128.cell_volume / factors[:blocks5].sum()
    print(f"config 5: 128 WF x {nR} R, DOS + CumDOS + Ohmic_FermiSurf, {blocks5} K-blocks of 32^3 ({nk} k-points): {dt:.2f} s = "
          f"{nk / dt:.3e} k-points/s per GPU -> 512^3 in {134217728 / (nk / dt) / 3600:.2f} GPU-hours; "
          f"CumDOS(top) = {cum[-1]:.6f} bands", flush=True)


if __name__ == "__main__":
    main()
