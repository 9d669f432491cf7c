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

    def __init__(self, osys=None):
        self.osys = osys if osys is not None else orc.OracleSystem.from_npz(os.path.join(GOLDEN, "fe_system.npz"))

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

    def scan_blocks(self, dK, specs):
        return [np.array([self.scan([d], [1.], [s])[0] for d in dK]).reshape((len(dK),) + s.shape) for s in specs]

    def kubo_scan(self, dK, weight, spec, Efermi, omega):
        kind = {_lib.KUBO_OPTCOND: "opt_conductivity", _lib.KUBO_JDOS: "jdos", _lib.KUBO_SHC: "shc", _lib.KUBO_SHIFT: "shift_current",
                _lib.KUBO_INJECTION: "injection_current"}[int(spec.kind)]
        shc_type = {v: k for k, v in _lib.SHC_TYPES.items()}.get(int(spec.shc_type), "ryoo")
        out = 0
        for d, w in zip(dK, weight):
            data = orc.OracleDataK(self.osys, d, self.NKFFT)
            out = out + w * orc.kubo_scan(data, kind, Efermi, omega, kBT=spec.kBT, smr_fixed_width=spec.smr_fixed_width,
                                          smr_type="Lorentzian" if spec.smr_type == 0 else "Gaussian",
                                          degen_thresh=spec.degen_thresh, degen_Kramers=bool(spec.degen_Kramers),
                                          external_terms=bool(spec.external_terms), constant_factor=spec.factor,
                                          SHC_type=shc_type, sc_eta=spec.sc_eta)
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
    # Kubo scan sharded over the ranks (complex result through the float64 all-reduce)
    gk = np.load(os.path.join(GOLDEN, "golden_fe_kubo.npz"))
    oc = wb.calculators.dynamic.OpticalConductivity(Efermi=gk["ref_Efermi"], omega=gk["ref_omega"], smr_fixed_width=0.20,
                                                    smr_type="Gaussian")
    res = wb.run(fe, wb.Grid(fe, NK=[4, 4, 4], NKFFT=[2, 2, 2]), dict(oc=oc), parallel=True, device=0)
    ref = gk["upstream_golden_opt_conductivity"]
    assert np.abs(res.results["oc"].data - ref).max() / np.abs(ref).max() < 1e-8
    # spin Hall conductivity on the reference's `random` system (rank-3 complex result), K-blocks sharded over the ranks
    gs = np.load(os.path.join(GOLDEN, "golden_random_shc.npz"))
    rnd = wb.System_R.from_npz(os.path.join(GOLDEN, "random_system.npz"))
    eng3 = OracleEngine(orc.OracleSystem.from_npz(os.path.join(GOLDEN, "random_system.npz")))
    sys.modules["wannierberri_b200.run"].engine_for = lambda system, device=0: eng3
    shc = wb.calculators.dynamic.SHC(Efermi=gs["ref_Efermi"], omega=gs["omega"], smr_fixed_width=0.20, smr_type="Gaussian",
                                     SHC_type="qiao")
    res = wb.run(rnd, wb.Grid(rnd, NK=gs["NK"], NKFFT=gs["NKFFT"]), dict(shc=shc), parallel=True, device=0)
    ref = gs["upstream_golden_qiao"]
    assert np.abs(res.results["shc"].data - ref).max() / np.abs(ref).max() < 1e-8
    # adaptive refinement: the new K-points of every iteration are sharded, all ranks take the same decisions
    ga = np.load(os.path.join(GOLDEN, "golden_synth_adpt.npz"))
    sysg = wb.synthetic_system(6, rmax=1, seed=4242)
    eng2 = OracleEngine(orc.OracleSystem(sysg.rvec.iRvec, sysg.real_lattice, sysg.wannier_centers_cart,
                                         {k: sysg.get_R_mat(k) for k in ("Ham", "AA")}))
    sys.modules["wannierberri_b200.run"].engine_for = lambda system, device=0: eng2
    calcs = dict(ahc=st.AHC(Efermi=ga["Efermi"]), dos=st.DOS(Efermi=ga["Efermi"]))
    res = wb.run(sysg, wb.Grid(sysg, NKdiv=[2, 2, 2], NKFFT=[3, 3, 3]), calcs, adpt_num_iter=1, adpt_fac=2, adpt_mesh=2,
                 parallel=True, device=0)
    for q in calcs:
        err = np.abs(res.results[q].data - ga["iter1_" + q]).max() / np.abs(ga["iter1_" + q]).max()
        assert err < 1e-8, (q, err)
    dist.barrier()
    if rank == 0:
        print("PARITY OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
