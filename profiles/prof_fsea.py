"""Driver for ncu captures of the Fermi-sea formula kernels on Te (24 WF, K-blocks of 20^3):
   python profiles/prof_fsea.py {bd_sea|gme_orb_sea|nldrude_sea} [blocks]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wannierberri_b200 as wb
st = wb.calculators.static
te = wb.System_R.from_npz(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "te_system.npz"))
Ef = np.linspace(4.0, 8.0, 401)
which = sys.argv[1] if len(sys.argv) > 1 else "bd_sea"
calc = dict(bd_sea=st.BerryDipole_FermiSea, gme_orb_sea=st.GME_orb_FermiSea, nldrude_sea=st.NLDrude_FermiSea)[which](Efermi=Ef)
specs = calc.specs()
eng = wb.Engine(te, device=0)
eng.plan([20, 20, 20], [s.formula for s in specs], external_terms=True)
shifts, factors = wb.Grid(te, NKdiv=[10, 10, 10], NKFFT=[20, 20, 20]).K_arrays()
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 4
for i in range(2):
    out = eng.scan(shifts[:nb], factors[:nb], specs)
print("checksum", float(sum(np.abs(o).sum() for o in out)))
