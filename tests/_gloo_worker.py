"""Worker of tests/test_host.py::test_two_rank_gloo_run (one process per rank, gloo backend)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch.distributed as dist  # noqa: E402

import wannierberri_b200 as wb  # noqa: E402
from wannierberri_b200 import _lib  # noqa: E402
from oracle import wb_oracle as orc  # noqa: E402  (test double + checker)

GOLDEN = os.path.join(ROOT, "tests", "golden")


class OracleEngine:
    """Test double with the interface of wannierberri_b200.Engine; shards are evaluated by the oracle."""
    FORMULA = {_lib.IDENTITY: orc.Identity, _lib.OMEGA: orc.Omega}

    def __init__(self):
        self.osys = orc.OracleSystem.from_npz(os.path.join(GOLDEN, "fe_system.npz"))

    def plan(self, NKFFT, formulae, external_terms=True, max_kpoints_per_launch=0):
        self.NKFFT = np.array(NKFFT)

    def scan(self, dK, weight, specs):
        out = [np.zeros(s.shape) for s in specs]
        for d, w in zip(dK, weight):
            data = orc.OracleDataK(self.osys, d, self.NKFFT)
            for i, s in enumerate(specs):
                Ef = np.linspace(s.Ef_first, s.Ef_last, s.nEF)
                out[i] += w * orc.static_scan(data, self.FORMULA[s.formula], s.fder, Ef, degen_thresh=s.degen_thresh,
                                              degen_Kramers=bool(s.degen_Kramers), constant_factor=s.factor)
        return out


def main():
    rank, world = int(sys.argv[1]), int(sys.argv[2])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = OracleEngine()
    sys.modules["wannierberri_b200.run"].engine_for = lambda system, device=0: eng
    fe = wb.System_R.from_npz(os.path.join(GOLDEN, "fe_system.npz"))
    g = np.load(os.path.join(GOLDEN, "golden_fe_nk4.npz"))
    Ef = g["Efermi"]
    st = wb.calculators.static
    calcs = dict(ahc=st.AHC(Efermi=Ef), dos=st.DOS(Efermi=Ef), cumdos=st.CumDOS(Efermi=Ef))
    res = wb.run(fe, wb.Grid(fe, NK=[4, 4, 4], NKFFT=[2, 2, 2]), calcs, parallel=True, device=0)
    for q in calcs:
        err = np.abs(res.results[q].data - g["upstream_golden_" + q]).max() / np.abs(g["upstream_golden_" + q]).max()
        assert err < 1e-8, (q, err)
    dist.barrier()
    if rank == 0:
        print("PARITY OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
