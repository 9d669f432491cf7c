import os, sys
import numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/scratch")
import wannierberri_b200 as wb
from wannierberri_b200 import _lib
s = wb.kramers_system(6, seed=6)
eng = wb.Engine(s); eng.set_option("eig_method", 1); eng.plan([8, 8, 8], [_lib.IDENTITY])
dK = [0.03, 0.01, 0.2]
E, U = eng.eig(dK, vectors=True); H = eng.xk(dK, "Ham")
print("sweeps", eng.last_eig_sweeps)
resid = np.abs(np.einsum("kij,kjn->kin", H, U) - U * E[:, None, :]).max(axis=(1, 2)) / np.abs(H).max()
unit = np.abs(np.einsum("kin,kim->knm", U.conj(), U) - np.eye(6)).max(axis=(1, 2))
bad = np.argsort(resid)[-5:]
print(bad, resid[bad], unit[bad])
ik = bad[-1]
np.set_printoptions(linewidth=200, precision=3)
A = U[ik].conj().T @ H[ik] @ U[ik]
print(np.abs(A - np.diag(np.diag(A))))
print(E[ik], np.linalg.eigvalsh(H[ik]))
print("herm err of H", np.abs(H[ik]-H[ik].conj().T).max(), "kpt", eng.kpoints(dK)[ik])
np.save("/root/repo/gpurun_out/badH.npy", H[ik])
