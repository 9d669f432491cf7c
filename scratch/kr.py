import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
import wannierberri_b200 as wb
from wannierberri_b200 import _lib
s = wb.kramers_system(24, seed=24)
out = {}
for method in (0, 1, 2):
    eng = wb.Engine(s); eng.set_option("eig_method", method); eng.plan([8, 8, 8], [_lib.IDENTITY])
    dK = [0.03, 0.01, 0.2]
    E, U = eng.eig(dK, vectors=True); H = eng.xk(dK, "Ham")
    resid = np.abs(np.einsum("kij,kjn->kin", H, U) - U * E[:, None, :]).max(axis=(1, 2)) / np.abs(H).max()
    out[method] = resid
    print(method, resid.max(), np.sort(resid)[-8:], eng.last_eig_resolved)
bad = np.argsort(out[0])[-60:]
print("method0 worst:", out[0][bad][-10:]); print("method1 same k:", out[1][bad][-10:])
print("corr:", np.corrcoef(np.log(out[0][bad]), np.log(out[1][bad]))[0, 1], (out[0] > 2e-14).sum())
