#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED upstream reference.

Run in the build container only (the reference tree does not travel to the GPU box):

    cd /tmp && PYTHONPATH=/root/reference:/root/repo/oracle/stubs \
        python /root/repo/tests/golden/make_golden.py

What it does
  1. builds the Fe 18-WF `System_R` from the saved R-space matrices
     (/root/reference/tests/reference/systems/Fe_W90, legacy [m,n,R] layout) and the Te 24-WF
     system (/root/reference/tests/data/Te_qe/system),
  2. runs the reference `wannierberri.run()` serially (numpy FFT, LAPACK eigh) on small grids,
  3. asserts that those runs reproduce the reference's OWN golden files
     (tests/reference/integrate_files/Fe_W90-*_iter-0000.npz) -- this is what pins the fixtures,
  4. writes compact fixtures: the input systems (fe_system.npz, te_system.npz) and the outputs
     (golden_*.npz) that tests/ compare the oracle and the CUDA path against.

Nothing here is imported by the product or by the tests; only the .npz outputs are used.
"""
import os
import sys
import tempfile

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

import wannierberri as wberri  # noqa: E402  (the upstream reference, via PYTHONPATH)
from wannierberri.system.system_R import System_R  # noqa: E402
from wannierberri.fourier.rvectors import Rvectors  # noqa: E402
from wannierberri.symmetry import point_symmetry as SYM  # noqa: E402
from wannierberri.data_K import Data_K_R  # noqa: E402
from wannierberri import calculators as calc  # noqa: E402


def build_fe():
    d = os.path.join(REF, "tests/reference/systems/Fe_W90")
    load = lambda k: np.load(os.path.join(d, k + ".npz"))["arr_0"]
    iRvec = load("iRvec")
    lattice = load("real_lattice")
    mats = {}
    for k in ("Ham", "AA", "BB", "CC", "SS"):
        a = load(k)
        mats[k] = np.ascontiguousarray(a.transpose((2, 0, 1) + tuple(range(3, a.ndim))))
    nw = mats["Ham"].shape[1]
    wcc = load("wannier_centers_cart")
    system = System_R(silent=True)
    system.set_real_lattice(lattice)
    system.num_wann = nw
    system.wannier_centers_cart = wcc
    system.rvec = Rvectors(lattice, iRvec=iRvec, shifts_left_red=system.wannier_centers_red)
    for k, v in mats.items():
        system.set_R_mat(k, v)
    system._NKFFT_recommended = np.array([3, 3, 3])
    system.set_pointgroup([SYM.C4z, SYM.C2x * SYM.TimeReversal, SYM.Inversion])
    return system


def build_te():
    system = System_R.from_npz(os.path.join(REF, "tests/data/Te_qe/system"), legacy=True)
    system.set_pointgroup(["C3z", "C2x", "TimeReversal"])
    return system


def dump_system(system, fname, keys):
    out = dict(
        iRvec=np.asarray(system.rvec.iRvec, dtype=np.int32),
        real_lattice=np.asarray(system.real_lattice, dtype=float),
        wannier_centers_cart=np.asarray(system.wannier_centers_cart, dtype=float),
        cRvec_shifted=np.asarray(system.rvec.cRvec_shifted, dtype=float),
        cell_volume=float(system.cell_volume),
    )
    for k in keys:
        out["XX_R_" + k] = np.asarray(system.get_R_mat(k))
    np.savez_compressed(os.path.join(OUT, fname), **out)


def run_ref(system, NK, NKFFT, calculators, **kw):
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            grid = wberri.Grid(system, NK=NK, NKFFT=NKFFT)
            res = wberri.run(system, grid=grid, calculators=calculators, parallel=False,
                             use_irred_kpt=False, symmetrize=False, fout_name="g",
                             print_progress_step_time=1e9, print_progress_step_percent=1000, **kw)
        finally:
            os.chdir(cwd)
    return grid, res


def main():
    fe = build_fe()
    dump_system(fe, "fe_system.npz", ("Ham", "AA", "BB", "CC", "SS"))

    # ---------------------------------------------------------------- Fe, the reference's test grid
    Ef = np.linspace(17, 18, 11)
    calcs = dict(ahc=calc.static.AHC(Efermi=Ef), dos=calc.static.DOS(Efermi=Ef),
                 cumdos=calc.static.CumDOS(Efermi=Ef), Morb=calc.static.Morb(Efermi=Ef),
                 spin=calc.static.Spin(Efermi=Ef),
                 ahc_int=calc.static.AHC(Efermi=Ef, kwargs_formula={"external_terms": False}),
                 berry_dipole_fsurf=calc.static.BerryDipole_FermiSurf(Efermi=Ef),
                 gme_orb_fsurf=calc.static.GME_orb_FermiSurf(Efermi=Ef),
                 gme_spin_fsurf=calc.static.GME_spin_FermiSurf(Efermi=Ef))
    grid, res = run_ref(fe, [4, 4, 4], [2, 2, 2], calcs)
    out = dict(Efermi=Ef, NK=np.array([4, 4, 4]), NKFFT=np.array([2, 2, 2]))
    for q in ("ahc", "dos", "cumdos", "Morb", "spin"):
        ref = np.load(os.path.join(REF, "tests/reference/integrate_files", f"Fe_W90-{q}_iter-0000.npz"))["data"]
        got = res.results[q].data
        err = np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300)
        print(f"Fe_W90-{q}: live reference run vs reference golden file: rel err {err:.2e}")
        assert err < 1e-8, q
        out["upstream_golden_" + q] = ref
    for q in calcs:
        out[q] = res.results[q].data
    # bit-exact k-grid fixtures: the K-block list and the k-points of block #5
    K_list = grid.get_K_list(use_symmetry=False)
    out["K_list_Kp_fullBZ"] = np.array([K.Kp_fullBZ for K in K_list])
    out["K_list_factor"] = np.array([K.factor for K in K_list])
    out["points_FFT"] = np.array(grid.points_FFT)
    np.savez_compressed(os.path.join(OUT, "golden_fe_nk4.npz"), **out)

    # ---------------------------------------------------------------- Fe, per-K-block stage dumps
    grid = wberri.Grid(fe, NK=[6, 6, 6], NKFFT=[3, 3, 3])
    K_list = grid.get_K_list(use_symmetry=False)
    Kp = K_list[5]
    data = Data_K_R(fe, dK=Kp.Kp_fullBZ, grid=grid, Kpoint=Kp, fftlib="numpy")
    Ef2 = np.linspace(12.0, 22.0, 101)
    out = dict(dK=np.array(Kp.Kp_fullBZ), NKFFT=np.array([3, 3, 3]), Efermi=Ef2,
               kpoints_all=np.array(data.kpoints_all), E_K=np.array(data.E_K),
               HH_K=np.array(data.HH_K),
               Xbar_Ham1_absdiag=np.abs(np.einsum("knna->kna", data.Xbar("Ham", 1))),
               delE_K=np.array(data.delE_K))
    # gauge-invariant probes: |X̄| is gauge dependent off the diagonal only through phases
    out["absV"] = np.abs(data.Xbar("Ham", 1))
    out["absA"] = np.abs(data.Xbar("AA"))
    for name, c in dict(ahc=calc.static.AHC(Efermi=Ef2), dos=calc.static.DOS(Efermi=Ef2),
                        cumdos=calc.static.CumDOS(Efermi=Ef2), Morb=calc.static.Morb(Efermi=Ef2),
                        ahc_kramers=calc.static.AHC(Efermi=Ef2, degen_Kramers=True),
                        ahc_thresh=calc.static.AHC(Efermi=Ef2, degen_thresh=0.05),
                        morb_thresh=calc.static.Morb(Efermi=Ef2, degen_thresh=0.05),
                        bcd_thresh=calc.static.BerryDipole_FermiSurf(Efermi=Ef2, degen_thresh=0.05),
                        gme_orb_thresh=calc.static.GME_orb_FermiSurf(Efermi=Ef2, degen_thresh=0.05),
                        gme_spin_thresh=calc.static.GME_spin_FermiSurf(Efermi=Ef2, degen_thresh=0.05),
                        ).items():
        out["block_" + name] = c(data).data
    np.savez_compressed(os.path.join(OUT, "golden_fe_block.npz"), **out)

    # ---------------------------------------------------------------- Fe, config-1-like scan (wide window)
    Ef3 = np.linspace(12.0, 22.0, 1001)
    calcs = dict(ahc=calc.static.AHC(Efermi=Ef3), dos=calc.static.DOS(Efermi=Ef3),
                 cumdos=calc.static.CumDOS(Efermi=Ef3))
    grid, res = run_ref(fe, [8, 8, 8], [4, 4, 4], calcs)
    out = dict(Efermi=Ef3, NK=np.array([8, 8, 8]), NKFFT=np.array([4, 4, 4]))
    for q in calcs:
        out[q] = res.results[q].data
    np.savez_compressed(os.path.join(OUT, "golden_fe_nk8.npz"), **out)

    # ---------------------------------------------------------------- Te (config 3 family, no tetra)
    te = build_te()
    dump_system(te, "te_system.npz", ("Ham", "AA", "BB", "CC", "SS"))
    Ef4 = np.linspace(4, 8, 41)
    calcs = dict(berry_dipole_fsurf=calc.static.BerryDipole_FermiSurf(Efermi=Ef4),
                 gme_orb_fsurf=calc.static.GME_orb_FermiSurf(Efermi=Ef4),
                 gme_spin_fsurf=calc.static.GME_spin_FermiSurf(Efermi=Ef4),
                 ahc=calc.static.AHC(Efermi=Ef4), dos=calc.static.DOS(Efermi=Ef4),
                 cumdos=calc.static.CumDOS(Efermi=Ef4), Morb=calc.static.Morb(Efermi=Ef4),
                 ahc_kramers=calc.static.AHC(Efermi=Ef4, degen_Kramers=True))
    grid, res = run_ref(te, [4, 4, 6], [2, 2, 3], calcs)
    out = dict(Efermi=Ef4, NK=np.array([4, 4, 6]), NKFFT=np.array([2, 2, 3]))
    for q in calcs:
        out[q] = res.results[q].data
    np.savez_compressed(os.path.join(OUT, "golden_te_nk4.npz"), **out)
    print("fixtures written to", OUT)


if __name__ == "__main__":
    sys.exit(main())
