"""Eigensolver stress / A-B on one GPU: eig_method 0 (twisted factorisation) vs 2 (QL with accumulated rotations) on
Fe K-blocks that contain the high-symmetry points, on Te, and on synthetic spectra with exact multiplets.
python profiles/eig_stress.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wannierberri_b200 as wb
from wannierberri_b200 import _lib

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def quality(eng, dK):
    E, U = eng.eig(dK, vectors=True)
    H = eng.xk(dK, "Ham")
    nw = E.shape[1]
    resid = np.abs(np.einsum("kij,kjn->kin", H, U) - U * E[:, None, :]).max(axis=(1, 2)) / np.abs(H).max()
    unit = np.abs(np.einsum("kin,kim->knm", U.conj(), U) - np.eye(nw)).max(axis=(1, 2))
    Eref = np.linalg.eigvalsh(H)
    return E, np.abs(E - Eref).max() / np.abs(Eref).max(), resid.max(), unit.max(), eng.last_eig_resolved


def case(name, system, NKFFT, dK):
    for method in (2, 0):
        eng = wb.Engine(system)
        eng.set_option("eig_method", method)
        eng.plan(NKFFT, [_lib.IDENTITY])
        E, de, r, u, nres = quality(eng, dK)
        print(f"{name:28s} method {method}: nk {E.shape[0]:6d}  dE {de:.1e}  resid {r:.1e}  unit {u:.1e}  jacobi {nres}", flush=True)


fe = wb.System_R.from_npz(os.path.join(GOLDEN, "fe_system.npz"))
te = wb.System_R.from_npz(os.path.join(GOLDEN, "te_system.npz"))
case("Fe 24^3 Gamma-centred", fe, [24, 24, 24], [0., 0., 0.])
case("Fe 20^3 shifted", fe, [20, 20, 20], [0.0125, 0.0375, 0.0025])
case("Te 16^3 Gamma-centred", te, [16, 16, 16], [0., 0., 0.])
for nw in (4, 8, 12, 18, 24):
    for deg in (False, True):
        s = wb.synthetic_system(nw, rmax=1, seed=nw, matrices=("Ham",), degenerate_pairs=deg)
        case(f"synthetic nw={nw} deg={deg}", s, [6, 6, 6], [0.03, 0.01, 0.2])
for nw in (6, 12, 18, 24):
    case(f"PT-symmetric (Kramers) nw={nw}", wb.kramers_system(nw, seed=nw), [8, 8, 8], [0.03, 0.01, 0.2])
