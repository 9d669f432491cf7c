"""Large-num_wann eigensolver check (33 .. 128 WF, with exactly degenerate spectra): method 0 (twisted factorisation, QL replay as
fallback) vs method 2 (replay only): errors against LAPACK and the number of matrices that took the fallback.
   python profiles/eig_large.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wannierberri_b200 as wb
from wannierberri_b200 import _lib
for nw, deg in ((33, False), (40, False), (40, True), (64, False), (96, True), (127, False), (128, False)):
    s = wb.synthetic_system(nw, rmax=1, seed=nw, matrices=("Ham",), degenerate_pairs=deg)
    for method in (0, 2):
        eng = wb.Engine(s); eng.set_option("eig_method", method); eng.plan([6, 6, 6], [_lib.IDENTITY])
        dK = [0.03, 0.01, 0.2]
        E, U = eng.eig(dK, vectors=True); H = eng.xk(dK, "Ham")
        resid = np.abs(np.einsum("kij,kjn->kin", H, U) - U * E[:, None, :]).max() / np.abs(H).max()
        unit = np.abs(np.einsum("kin,kim->knm", U.conj(), U) - np.eye(nw)).max()
        dE = np.abs(E - np.linalg.eigvalsh(H)).max() / np.abs(E).max()
        print(f"nw={nw} deg={deg} method={method}: dE {dE:.1e} resid {resid:.1e} unit {unit:.1e} replay/jacobi {eng.last_eig_resolved} of {E.shape[0]}", flush=True)
        eng.close()
