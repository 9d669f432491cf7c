#!/usr/bin/env python
"""Small driver for ncu captures: the bench workload (Fe 18-WF, NKFFT=20^3 K-blocks of the 400^3 grid,
2000 Fermi levels), `--blocks` K-blocks per scan, `--scans` scans, no timing, no probes.

  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      python profiles/prof_driver.py --blocks 32 --scans 2
  ncu --set full --clock-control none --import-source on -k regex:'wb_' -s <first scan's launches> -c <n> \
      -o gpurun_out/prof python profiles/prof_driver.py --blocks 8 --scans 2
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=32)
    ap.add_argument("--scans", type=int, default=2)
    ap.add_argument("--workload", default="ahc_dos", choices=["ahc_dos", "ahc_morb", "te_fsurf"])
    ap.add_argument("--eig-method", type=int, default=0)
    args = ap.parse_args()
    import wannierberri_b200 as wb
    st = wb.calculators.static
    if args.workload == "te_fsurf":
        system = wb.System_R.from_npz(os.path.join(ROOT, "tests", "golden", "te_system.npz"))
        Ef = np.linspace(4.0, 8.0, 401)
        calcs = dict(bcd=st.BerryDipole_FermiSurf(Efermi=Ef), gme_orb=st.GME_orb_FermiSurf(Efermi=Ef),
                     gme_spin=st.GME_spin_FermiSurf(Efermi=Ef))
    else:
        system = wb.System_R.from_npz(os.path.join(ROOT, "tests", "golden", "fe_system.npz"))
        Ef = np.linspace(12.0, 22.0, 2000)
        if args.workload == "ahc_dos":
            calcs = dict(ahc=st.AHC(Efermi=Ef), dos=st.DOS(Efermi=Ef))
        else:
            calcs = dict(ahc=st.AHC(Efermi=Ef), morb=st.Morb(Efermi=Ef))
    specs = [s for c in calcs.values() for s in c.specs()]
    eng = wb.Engine(system, device=0)
    eng.set_option("eig_method", args.eig_method)
    eng.plan([20, 20, 20], [s.formula for s in specs], external_terms=True)
    grid = wb.Grid(system, NKdiv=[20, 20, 20], NKFFT=[20, 20, 20])
    shifts, factors = grid.K_arrays()
    for i in range(args.scans):
        out = eng.scan(shifts[:args.blocks], factors[:args.blocks], specs)
    print("launches", eng.kernel_launches, "checksum", float(sum(np.abs(o).sum() for o in out)))


if __name__ == "__main__":
    main()
