import sys, os, ctypes as C
sys.path.insert(0, os.getcwd())
import numpy as np
import wannierberri_b200 as wb
from wannierberri_b200 import _lib
st = wb.calculators.static
fe = wb.System_R.from_npz("tests/golden/fe_system.npz")
Ef = np.linspace(12.0, 22.0, 2000)
for case, calcs in (("ahc_dos", dict(ahc=st.AHC(Efermi=Ef), dos=st.DOS(Efermi=Ef))), ("ahc_morb", dict(ahc=st.AHC(Efermi=Ef), morb=st.Morb(Efermi=Ef)))):
    specs = [s for c in calcs.values() for s in c.specs()]
    for opts in (dict(rot_r2=-1), dict(rot_r2=0), dict(rot_r2=1), dict(rot_r2=2), dict(rot_r2=3)):
        eng = wb.Engine(fe, device=0)
        for k, v in opts.items(): eng.set_option(k, v)
        eng.plan([20, 20, 20], [s.formula for s in specs], external_terms=True)
        grid = wb.Grid(fe, NKdiv=[20, 20, 20], NKFFT=[20, 20, 20])
        shifts, factors = grid.K_arrays()
        r0 = eng.scan(shifts[:32], factors[:32], specs)
        eng.set_option("timing", 1)
        for _ in range(3): eng.scan(shifts[:32], factors[:32], specs)
        ms = (C.c_double * 5)(); calls = (C.c_int64 * 5)()
        _lib.check(_lib.lib().wbgpu_stage_times(eng._ctx, ms, calls))
        print(case, opts, "rotate ms/256k = %.3f" % (ms[2] / 3), "checksum", float(np.abs(r0[0]).sum()))
        eng.close()
