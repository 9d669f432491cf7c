import os, sys, time
import numpy as np
ROOT = "/root/repo"
sys.path[:0] = ["/root/reference", os.path.join(ROOT, "oracle", "stubs"), os.path.join(ROOT, "tests", "golden"), ROOT]
import torch
import wannierberri as wberri
from make_golden import build_fe
from wannierberri.data_K import Data_K_R as RefDataK
from wannierberri.formula import covariant as frml
from wannierberri_b200 import formula_gpu as fg

system = build_fe()
grid = wberri.Grid(system, NK=[4, 4, 4], NKFFT=[2, 2, 2])
dK = np.array([0.125, 0.0, 0.125])
dk = RefDataK(system, dK=dK, grid=grid)
nw = dk.num_wann
groups = dk.get_bands_in_range_groups(15., 19., degen_thresh=float(os.environ.get("DT","1e-4")), degen_Kramers=os.environ.get("KR")=="1", sea=(os.environ.get("SEA")=="1"))
kidx, a, b = [], [], []
for ik, g in enumerate(groups):
    for (x, y) in g:
        kidx.append(ik); a.append(x); b.append(y)
kidx, a, b = map(np.array, (kidx, a, b))
print("pairs", len(kidx))
names = sys.argv[1:] or list(fg.TRACES)
ref_cls = dict(NLDrude_Z_spin=frml.NLDrude_Z_spin, NLDrude_Z_orb_Omega=frml.NLDrude_Z_orb_Omega, NLDrude_Z_orb_Hplus=frml.NLDrude_Z_orb_Hplus,
               emcha_surf=frml.emcha_surf, QuantumMetric_ab=frml.QuantumMetric_ab, VelDQM=frml.VelDQM)
for ext in (True, False):
  for name in names:
    kw = dict(external_terms=ext)
    if name in ("QuantumMetric_ab", "VelDQM"): kw["FF_rotAA"] = True
    if name == "NLDrude_Z_spin" and not ext: continue
    t0 = time.time()
    f = ref_cls[name](dk, **kw)
    want = np.array([f.trace(ik, np.arange(x, y), np.concatenate((np.arange(0, x), np.arange(y, nw)))) for ik, x, y in zip(kidx, a, b)])
    t1 = time.time()
    alg = fg.BlockAlgebra(dk, kidx, a, b, "cpu", True, ext)
    got = fg.TRACES[name][0](alg).numpy()
    t2 = time.time()
    print(f"{name:22s} ext={ext} ref {t1-t0:6.1f}s mine {t2-t1:6.1f}s  max|want| {np.abs(want).max():.3e}  err {np.abs(got-want).max()/np.abs(want).max():.2e}", flush=True)
