#!/usr/bin/env python
"""The BASELINE configurations END TO END through run() on one GPU, at their full sizes:

  config 1   Fe, AHC + DOS, 48^3 grid (NKdiv 4 x NKFFT 12), 1001 Fermi levels 12..22 eV -- GPU run() AND the CPU oracle on
             all host cores (same K-blocks), relative difference reported (the CPU-pinned case)
  config 2   Fe, AHC + Morb (and AHC + DOS), 400^3 grid (NKdiv 20 x NKFFT 20 = 6.4e7 k-points), 2000 Fermi levels -- GPU only

    python profiles/full_configs.py [--skip-oracle]
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
FE = os.path.join(ROOT, "tests", "golden", "fe_system.npz")
EF1 = np.linspace(12.0, 22.0, 1001)


def _init():
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"
    from oracle import wb_oracle as orc
    _block.sys = orc.OracleSystem.from_npz(FE)


def _block(args):
    from oracle import wb_oracle as orc
    dK, w = args
    data = orc.OracleDataK(_block.sys, dK, [12, 12, 12])
    return w * orc.AHC(data, EF1), w * orc.DOS(data, EF1)


def main():
    import wannierberri_b200 as wb
    st = wb.calculators.static
    fe = wb.System_R.from_npz(FE)
    # ---- config 1
    grid = wb.Grid(fe, NKdiv=[4, 4, 4], NKFFT=[12, 12, 12])
    calcs = dict(ahc=st.AHC(Efermi=EF1), dos=st.DOS(Efermi=EF1))
    wb.run(fe, grid, calcs)   # warm-up (context, plan)
    t0 = time.perf_counter()
    res = wb.run(fe, grid, calcs)
    t_gpu = time.perf_counter() - t0
    nk = 48 ** 3
    print(f"config 1: Fe AHC+DOS 48^3 ({nk} k-points, 1001 E_F): GPU run() {t_gpu * 1e3:.1f} ms = {nk / t_gpu:.3e} k-points/s")
    if "--skip-oracle" not in sys.argv:
        shifts, factors = grid.K_arrays()
        cores = len(os.sched_getaffinity(0))
        t0 = time.perf_counter()
        with mp.Pool(cores, initializer=_init) as pool:
            parts = pool.map(_block, list(zip(shifts, factors)))
        t_cpu = time.perf_counter() - t0
        ahc = sum(p[0] for p in parts)
        dos = sum(p[1] for p in parts)
        e1 = np.abs(res.results["ahc"].data - ahc).max() / np.abs(ahc).max()
        e2 = np.abs(res.results["dos"].data - dos).max() / np.abs(dos).max()
        print(f"          CPU oracle on {cores} cores: {t_cpu:.1f} s = {nk / t_cpu:.3e} k-points/s; "
              f"GPU vs CPU: AHC rel {e1:.2e}, DOS rel {e2:.2e}; speed-up {t_cpu / t_gpu:.0f}x")
    # ---- config 2
    EF2 = np.linspace(12.0, 22.0, 2000)
    grid = wb.Grid(fe, NKdiv=[20, 20, 20], NKFFT=[20, 20, 20])
    for name, calcs in (("AHC+DOS", dict(ahc=st.AHC(Efermi=EF2), dos=st.DOS(Efermi=EF2))),
                        ("AHC+Morb", dict(ahc=st.AHC(Efermi=EF2), morb=st.Morb(Efermi=EF2)))):
        t0 = time.perf_counter()
        res = wb.run(fe, grid, calcs)
        dt = time.perf_counter() - t0
        a = res.results["ahc"].data
        print(f"config 2: Fe {name} 400^3 (6.4e7 k-points, 2000 E_F): GPU run() {dt:.2f} s = {6.4e7 / dt:.3e} k-points/s; "
              f"sigma_z(E_F = 12.6 eV) = {a[np.argmin(abs(EF2 - 12.6)), 2]:.6e} S/m")


if __name__ == "__main__":
    main()
