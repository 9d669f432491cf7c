"""Parity of the CUDA path (through the C-ABI of libwbgpu.so) against the oracle and against the
fixtures generated from the unmodified reference.  Needs a B200: `pytest -m gpu`.

Tolerances: 1e-8 relative to the max-norm over the Fermi axis for integrated quantities and for
eigenvalues (BASELINE.json north_star; the reference's own regression tolerance,
tests/common_comparers.py:186-189); bit-exact for the k-grid."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

RTOL = 1e-8


def relerr(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def wb():
    import wannierberri_b200 as wb
    from wannierberri_b200 import _lib
    assert _lib.lib().wbgpu_device_count() > 0, "no CUDA device: the GPU tests cannot run (no CPU fallback exists)"
    return wb


@pytest.fixture(scope="module")
def orc():
    from oracle import wb_oracle
    return wb_oracle


@pytest.fixture(scope="module")
def fe(wb):
    return wb.System_R.from_npz(os.path.join(GOLDEN, "fe_system.npz"))


@pytest.fixture(scope="module")
def te(wb):
    return wb.System_R.from_npz(os.path.join(GOLDEN, "te_system.npz"))


@pytest.fixture(scope="module")
def fe_orc(orc):
    return orc.OracleSystem.from_npz(os.path.join(GOLDEN, "fe_system.npz"))


@pytest.fixture(scope="module")
def te_orc(orc):
    return orc.OracleSystem.from_npz(os.path.join(GOLDEN, "te_system.npz"))


ALL = None


def all_formulae():
    from wannierberri_b200 import _lib
    return [_lib.IDENTITY, _lib.OMEGA]


FORMULA_CASES = ["Omega", "Morb_Hpm", "Spin", "VelOmega", "VelHplus", "VelSpin"]


@pytest.mark.parametrize("rotate_method", [0, 1, 4])
@pytest.mark.parametrize("thresh", [1e-4, 0.05])
@pytest.mark.parametrize("name", FORMULA_CASES)
def test_formula_band_traces(wb, fe, fe_orc, orc, name, thresh, rotate_method):
    """Per-k, per-band-group values of every formula against the oracle's Formula classes evaluated with the
    reference's additive / non-additive loop (static.py:102-117), same groups."""
    from wannierberri_b200 import _lib
    from wannierberri_b200._lib import ScanSpec
    code = dict(Omega=_lib.OMEGA, Morb_Hpm=_lib.MORB_HPM, Spin=_lib.SPIN, VelOmega=_lib.VEL_OMEGA,
                VelHplus=_lib.VEL_HPLUS, VelSpin=_lib.VEL_SPIN)[name]
    b = np.load(os.path.join(GOLDEN, "golden_fe_block.npz"))
    NKFFT, dK = b["NKFFT"], b["dK"]
    fder = 0 if name in ("Omega", "Morb_Hpm", "Spin") else 1
    Ef = np.linspace(14., 20., 31)
    dEF = Ef[1] - Ef[0]
    spec = ScanSpec(formula=code, fder=fder, nEF=len(Ef), degen_Kramers=0, internal_terms=1, external_terms=1,
                    Ef_first=Ef[0], Ef_last=Ef[-1], dEF=dEF, degen_thresh=thresh, factor=1.0)
    eng = wb.Engine(fe)
    # 0 = automatic choice, 1 = shared-memory DFMA kernel, 4 = batched DMMA GEMM to global memory + formula kernel
    eng.set_option("rotate_method", rotate_method)
    eng.plan(NKFFT, [_lib.IDENTITY, code])
    lab, val = eng.band_traces(dK, spec)
    data = orc.OracleDataK(fe_orc, dK, NKFFT)
    form = getattr(orc, name)(data)
    nw = fe.num_wann
    EFmin, EFmax = Ef[0] - fder * dEF, Ef[-1] + fder * dEF
    worst = scale = 0.
    for ik in range(data.nk):
        groups = orc.band_groups(data.E_K[ik], EFmin, EFmax, thresh, False, sea=(fder == 0))
        if form.additive:
            vals = {n: form.trace(ik, np.arange(n[0], n[1]), np.concatenate((np.arange(0, n[0]), np.arange(n[1], nw))))
                    for n in groups}
        else:
            edge = {x: form.trace(ik, np.arange(0, x), np.arange(x, nw)) for n in groups for x in n}
            vals = {n: edge[n[1]] - edge[n[0]] for n in groups}
        for n, ref in vals.items():
            worst = max(worst, np.abs(val[ik, n[0]] - ref).max())
            scale = max(scale, np.abs(ref).max())
    assert worst < RTOL * scale


def test_kpoints_bit_exact(wb, fe, orc):
    b = np.load(os.path.join(GOLDEN, "golden_fe_block.npz"))
    eng = wb.Engine(fe)
    eng.plan(b["NKFFT"], all_formulae())
    assert np.array_equal(eng.kpoints(b["dK"]), b["kpoints_all"])
    for NKFFT, dK in (([4, 5, 6], [0.013, 0.0, 0.07]), ([12, 12, 12], [1 / 48, 2 / 48, 3 / 48]), ([1, 1, 7], [0.99, 0.5, 1 / 3])):
        eng.plan(NKFFT, all_formulae())
        assert np.array_equal(eng.kpoints(dK), orc.kpoints_all(NKFFT, dK))


@pytest.mark.parametrize("NKFFT,dK", [([3, 3, 3], [1 / 6, 0., 1 / 12]), ([2, 2, 2], [0., 0., 0.]),
                                      ([4, 6, 5], [0.01, 0.02, 0.03]), ([12, 12, 12], [1 / 48, 0, 3 / 48])])
def test_r_to_k(wb, fe, fe_orc, orc, NKFFT, dK):
    """Separable pruned DFT == the reference's scatter + ifftn (fourier/fft.py:133-192), all channels."""
    eng = wb.Engine(fe)
    eng.plan(NKFFT, all_formulae())
    data = orc.OracleDataK(fe_orc, dK, NKFFT)
    T = fe_orc.cRvec_shifted
    H = eng.xk(dK, "Ham")
    assert relerr(H, data.HH_K) < 1e-13
    dH = eng.xk(dK, "dHam")
    assert relerr(dH, data._R_to_k(orc.derivative(fe_orc.XX_R["Ham"], T), False)) < 1e-13
    A = eng.xk(dK, "AA")
    assert relerr(A, data._R_to_k(fe_orc.XX_R["AA"], True)) < 1e-13
    O = eng.xk(dK, "rotAA")
    assert relerr(O, data._R_to_k(data._rotAA_R(), True)) < 1e-13


def _lib_formula(name):
    from wannierberri_b200 import _lib
    return getattr(_lib, name)


@pytest.mark.parametrize("which", ["fe", "te"])
def test_eigh(wb, fe, te, fe_orc, te_orc, orc, which):
    sysg, syso = (fe, fe_orc) if which == "fe" else (te, te_orc)
    NKFFT, dK = [4, 4, 4], [0.02, 0.05, 0.11]
    eng = wb.Engine(sysg)
    eng.plan(NKFFT, all_formulae())
    E, U = eng.eig(dK, vectors=True)
    data = orc.OracleDataK(syso, dK, NKFFT)
    assert relerr(E, data.E_K) < 1e-12
    assert np.all(np.diff(E, axis=1) >= 0)
    H = data.HH_K
    resid = np.abs(np.einsum("kij,kjn->kin", H, U) - U * E[:, None, :]).max()
    assert resid < 1e-11 * np.abs(H).max()
    unit = np.abs(np.einsum("kin,kim->knm", U.conj(), U) - np.eye(sysg.num_wann)).max()
    assert unit < 1e-12


@pytest.mark.parametrize("nw,degenerate", [(1, False), (2, False), (3, False), (8, True), (9, False), (16, False),
                                           (17, False), (18, False), (18, True), (24, True), (32, False), (33, False),
                                           (40, False), (40, True), (64, False), (96, True), (127, False),
                                           (128, False)])
def test_eigh_sizes_and_degeneracies(wb, nw, degenerate):
    """The eigensolvers (Householder+QL: warp per k-point for nw <= 32, CTA per k-point for nw <= 128; Jacobi on
    request where it fits shared memory) against LAPACK on random Hermitian models, including exactly
    degenerate spectra."""
    sysg = wb.synthetic_system(nw, rmax=1, seed=nw, matrices=("Ham",), degenerate_pairs=degenerate)
    # nw = 18 also exercises the two-k-points-per-warp reduction (method 2) against the one-per-warp kernel
    # (method 3), on an odd number of k-points
    NKFFT, dK = ([3, 1, 3] if nw == 18 else [3, 2, 4]), [0.03, 0.01, 0.2]
    from wannierberri_b200 import _lib
    # method 0 = twisted-factorisation eigenvectors (nw <= 24), 2 / 3 = accumulated QL rotations, 4 / 5 = thread- /
    # lane-pair-per-matrix reduction, 1 = Jacobi
    for method in ((0, 2, 3, 4, 5, 1) if nw == 18 else (0, 5, 2, 1) if nw <= 20 else (0, 2, 1) if nw <= 40 else (0, 2)):   # nw > 32: 0 = twisted factorisation (QL replay as fallback), 2 = replay only
        eng = wb.Engine(sysg)
        eng.set_option("eig_method", method)
        eng.plan(NKFFT, [_lib.IDENTITY])
        E, U = eng.eig(dK, vectors=True)
        H = eng.xk(dK, "Ham")
        Eref = np.linalg.eigvalsh(H)
        assert relerr(E, Eref) < 1e-12, (nw, method)
        resid = np.abs(np.einsum("kij,kjn->kin", H, U) - U * E[:, None, :]).max()
        assert resid < 1e-11 * max(np.abs(H).max(), 1.), (nw, method, resid)
        unit = np.abs(np.einsum("kin,kim->knm", U.conj(), U) - np.eye(nw)).max()
        # (nw > 32, method 0: eigenvectors of T from twisted factorisations; neighbours beyond the 5e-3 |T| orthogonalisation
        # window keep ~4 eps |T| / gap: measured 2e-13 .. 9e-13 at nw = 128, against 1e-14 for the QL replay of method 2)
        assert unit < (3e-12 if nw > 32 and method == 0 else 1e-12), (nw, method, unit)


@pytest.mark.parametrize("nw", [6, 12, 18, 24, 40, 64])
def test_eigh_kramers_pairs(wb, nw):
    """PT-symmetric model: every level exactly doubly degenerate with the two halves of the basis coupled (the tridiagonal
    form is numerically reducible).  The twisted-factorisation solver hands the k-points it cannot resolve to the Jacobi
    kernel; both must deliver LAPACK-grade residuals (the Jacobi kernel stopped one sweep early on such spectra before)."""
    from wannierberri_b200 import _lib
    sysg = wb.kramers_system(nw, seed=nw)
    NKFFT, dK = [8, 8, 8], [0.03, 0.01, 0.2]
    # (nw > 32: method 0 = twisted factorisation, every matrix falls back to the QL replay here; 1 = Jacobi up to 40 WF)
    for method in ((0, 1, 2) if nw <= 40 else (0, 2)):
        eng = wb.Engine(sysg)
        eng.set_option("eig_method", method)
        eng.plan(NKFFT, [_lib.IDENTITY])
        E, U = eng.eig(dK, vectors=True)
        H = eng.xk(dK, "Ham")
        assert relerr(E, np.linalg.eigvalsh(H)) < 1e-13, (nw, method)
        assert np.abs(E[:, 0::2] - E[:, 1::2]).max() < 1e-13 * np.abs(E).max()
        resid = np.abs(np.einsum("kij,kjn->kin", H, U) - U * E[:, None, :]).max()
        assert resid < 2e-13 * np.abs(H).max(), (nw, method, resid)
        unit = np.abs(np.einsum("kin,kim->knm", U.conj(), U) - np.eye(nw)).max()
        assert unit < 5e-13, (nw, method, unit)


def test_eigh_golden_block(wb, fe):
    b = np.load(os.path.join(GOLDEN, "golden_fe_block.npz"))
    eng = wb.Engine(fe)
    eng.plan(b["NKFFT"], all_formulae())
    assert relerr(eng.eig(b["dK"]), b["E_K"]) < 1e-12


@pytest.mark.parametrize("rotate_method", [1, 2, 3, 4])
@pytest.mark.parametrize("kw", [dict(), dict(degen_thresh=0.05), dict(degen_Kramers=True),
                                dict(kwargs_formula=dict(external_terms=False)),
                                dict(kwargs_formula=dict(internal_terms=False))])
def test_omega_band_traces(wb, fe, fe_orc, orc, kw, rotate_method):
    """Per-k, per-band-group traces of Omega against the oracle's Formula.trace with the same groups."""
    b = np.load(os.path.join(GOLDEN, "golden_fe_block.npz"))
    NKFFT, dK, Ef = b["NKFFT"], b["dK"], b["Efermi"]
    calc = wb.calculators.static.AHC(Efermi=Ef, **kw)
    eng = wb.Engine(fe)
    eng.set_option("rotate_method", rotate_method)  # 1 = generic shared-memory kernel, 2 = runtime-nw DMMA kernel, 3 = compile-time-NW DMMA kernel
    eng.plan(NKFFT, all_formulae())
    lab, val = eng.band_traces(dK, calc.specs()[0])
    data = orc.OracleDataK(fe_orc, dK, NKFFT)
    form = orc.Omega(data, **kw.get("kwargs_formula", {}))
    nw = fe.num_wann
    scale = 0.
    worst = 0.
    for ik in range(data.nk):
        groups = orc.band_groups(data.E_K[ik], calc.EFmin, calc.EFmax, calc.degen_thresh, calc.degen_Kramers, sea=True)
        starts = {n[0] for n in groups}
        assert set(np.where(np.isfinite(lab[ik]) | (lab[ik] == -np.inf))[0]) == starts
        for (a, e), E in groups.items():
            inn = np.arange(a, e)
            out = np.concatenate((np.arange(0, a), np.arange(e, nw)))
            ref = form.trace(ik, inn, out)
            assert (lab[ik, a] == E) or abs(lab[ik, a] - E) < 1e-9
            worst = max(worst, np.abs(val[ik, a] - ref).max())
            scale = max(scale, np.abs(ref).max())
    assert worst < RTOL * scale


@pytest.mark.parametrize("rotate_method", [0, 1, 2, 4])
def test_dh_full_channel_path(wb, fe, rotate_method):
    """d_a H is hermitian in R-space for the Fe system, so the plan packs it as triangles; the full-matrix path
    (what a system without that symmetry gets, option dh_packed = 0) must give the same AHC / Morb."""
    b = np.load(os.path.join(GOLDEN, "golden_fe_block.npz"))
    st = wb.calculators.static
    specs = st.AHC(Efermi=b["Efermi"]).specs() + (st.Morb(Efermi=b["Efermi"]).specs() if rotate_method != 2 else [])
    out = []
    for packed in (1, 0):
        eng = wb.Engine(fe)
        eng.set_option("dh_packed", packed)
        eng.set_option("rotate_method", rotate_method)
        eng.plan(b["NKFFT"], [s.formula for s in specs])
        out.append(eng.scan(b["dK"][None, :], np.ones(1), specs))
        V = eng.xk(b["dK"], "dHam")
        assert np.abs(V - V.conj().transpose(0, 2, 1, 3)).max() < 1e-12 * np.abs(V).max()
        eng.close()
    for a, r in zip(*out):
        assert relerr(a, r) < 1e-11


def test_soc_system_vs_reference_data_k_soc(wb):
    """`run()` on a SOC system (scalar up / down systems + spin-orbit term on different R-vector sets; the reference
    evaluates it with Data_K_soc, data_K/data_K_soc.py:7-62) against the fixture of the unmodified reference
    (tests/golden/make_golden_soc.py): DOS, CumDOS, AHC with and without external terms, Spin, Ohmic, GME_spin."""
    from conftest import soc_system_from_fixture, SOC_CALCS
    g = np.load(os.path.join(GOLDEN, "golden_soc.npz"))
    soc = soc_system_from_fixture(wb, g)
    st = wb.calculators.static
    calcs = {k: getattr(st, name)(Efermi=g["Efermi"], **kw) for k, (name, kw) in SOC_CALCS.items()}
    res = wb.run(soc, wb.Grid(soc, NK=g["NK"], NKFFT=g["NKFFT"]), calcs, use_irred_kpt=False, symmetrize=False, write_files=False)
    for key in calcs:
        assert relerr(res.results[key].data, g["res_" + key]) < RTOL, key


def test_select_bands(wb, fe):
    """`select_bands` of the Fermi-surface calculators (static.py:93-100, 129-136) against the fixture of the unmodified
    reference; the band-resolved results add up to the unrestricted one (the reference's tests/test_calc.py:231-248); the
    Fermi sea and the tetrahedron method refuse it."""
    from test_oracle import SELECT_CASES, select_of
    g = np.load(os.path.join(GOLDEN, "golden_select.npz"))
    st = wb.calculators.static
    Ef = g["Efermi"]
    calcs = {k: getattr(st, name)(Efermi=Ef, select_bands=select_of(g, k), **kw) for k, (name, kw) in SELECT_CASES.items()}
    grid = wb.Grid(fe, NK=g["NK"], NKFFT=g["NKFFT"])
    res = wb.run(fe, grid, calcs, use_irred_kpt=False, symmetrize=False, write_files=False)
    for key in calcs:
        assert relerr(res.results[key].data, g[key]) < RTOL, key
    bands = {f"{i:02d}": st.Ohmic_FermiSurf(Efermi=Ef, select_bands=(i,)) for i in range(fe.num_wann)}
    bands["all"] = st.Ohmic_FermiSurf(Efermi=Ef)
    rb = wb.run(fe, grid, bands, use_irred_kpt=False, symmetrize=False, write_files=False)
    total = sum(rb.results[k].data for k in bands if k != "all")
    assert relerr(total, rb.results["all"].data) < 1e-10
    with pytest.raises(NotImplementedError):
        wb.run(fe, grid, dict(a=st.AHC(Efermi=Ef, select_bands=(1,))), write_files=False)
    with pytest.raises(NotImplementedError):
        st.Ohmic_FermiSurf(Efermi=Ef, select_bands=(1,), tetra=True)


def test_data_k_plugin_attributes(wb, fe, fe_orc, orc):
    """Plug-in hook #1 (data_K/data_K_R.py:69-97, data_K/data_K.py:211-326): what a user formula
    reads from `Data_K_R` -- `E_K`, `UU_K`, `Xbar(name, der)`, `D_H`, `dEig_inv`, `delE_K`, `kpoints_all`, band groups
    -- served by the CUDA kernels (`wbgpu_eig`, `wbgpu_xbar`).  Gauge-independent check of every matrix:
    U Xbar U^dagger must be the Wannier-gauge matrix of the oracle."""
    grid = wb.Grid(fe, NK=[6, 4, 4], NKFFT=[3, 2, 2])
    dK = np.array([0.05, 0.125, 0.0625])
    dk = wb.Data_K_R(fe, dK, grid)
    data = orc.OracleDataK(fe_orc, dK, grid.FFT)
    assert dk.nk == 12 and dk.num_wann == 18 and dk.nbands == 18
    assert np.array_equal(dk.kpoints_all, orc.kpoints_all(grid.FFT, dK))
    E, U = dk.E_K, dk.UU_K
    assert relerr(E, data.E_K) < 1e-12
    T = fe_orc.cRvec_shifted

    def wannier(name, der):
        X = data._rotAA_R() if name == "rotAA" else fe_orc.XX_R[name]
        for _ in range(der):
            X = orc.derivative(X, T)
        return data._R_to_k(X, name in ("AA", "SS", "rotAA"))

    for name, ders in (("Ham", (0, 1, 2, 3)), ("AA", (0, 1, 2)), ("rotAA", (0, 1, 2)), ("BB", (0, 1, 2)), ("CC", (0, 1, 2)), ("SS", (0, 1, 2))):
        for der in ders:
            Xbar = dk.Xbar(name, der)
            Xw = wannier(name, der)
            assert Xbar.shape == Xw.shape, (name, der)
            back = np.einsum("kia,kab...,kjb->kij...", U, Xbar, U.conj())
            assert relerr(back, Xw) < 1e-11, (name, der)
    assert relerr(np.einsum("kii->ki", dk.Xbar("Ham", 0)).real, E) < 1e-12
    ab = dk.Xbar("rotAAab", 1)                      # derived names: antisymmetric rank-2 forms of curl A / C
    assert ab.shape == (12, 18, 18, 3, 3, 3) and relerr(ab[:, :, :, 1, 2], -0.5j * dk.Xbar("rotAA", 1)[:, :, :, 0]) < 1e-15
    assert relerr(dk.Xbar("CCab_antisym")[:, :, :, 0, 2], 0.5j * dk.Xbar("CC")[:, :, :, 1]) < 1e-15
    dE = E[:, :, None] - E[:, None, :]
    inv = dk.dEig_inv
    assert np.all(inv[np.abs(dE) < 1e-7] == 0.) and np.allclose(inv[np.abs(dE) > 1e-7] * dE[np.abs(dE) > 1e-7], 1.)
    assert relerr(dk.D_H, -dk.Xbar("Ham", 1) * inv[..., None]) < 1e-14
    assert relerr(dk.delE_K, np.einsum("kiia->kia", dk.Xbar("Ham", 1)).real) == 0.
    for ik in (0, 7):
        for kw in (dict(sea=True), dict(degen_thresh=0.05, degen_Kramers=True), dict(degen_thresh=0.3)):
            got = dk.get_bands_in_range_groups_ik(ik, 16., 19., **kw)
            want = orc.band_groups(data.E_K[ik], 16., 19., kw.get("degen_thresh", -1), kw.get("degen_Kramers", False),
                                   sea=kw.get("sea", False))
            assert set(got) == set(want) and all(abs(got[g] - want[g]) < 1e-9 or got[g] == want[g] for g in got)


def test_plugin_formula_through_run(wb, fe):
    """A user-defined Formula class (tests/plugin_formula.py, written against the plug-in interface only) evaluated by
    `run()` on the GPU-resident Data_K_R, against the fixture that the UNMODIFIED reference produced with the same
    source on its own Data_K_R (tests/golden/make_golden_plugin.py): Fermi-sea, Fermi-surface and a non-additive case."""
    from plugin_formula import make_calculators
    g = np.load(os.path.join(GOLDEN, "golden_plugin.npz"))
    calcs = make_calculators(wb.calculators.static.StaticCalculator, g["Efermi"])
    calcs["ahc"] = wb.calculators.static.AHC(Efermi=g["Efermi"])   # a fused scan next to the plug-ins in one run
    res = wb.run(fe, wb.Grid(fe, NK=g["NK"], NKFFT=g["NKFFT"]), calcs, use_irred_kpt=False, symmetrize=False,
                 write_files=False)
    for key in ("user_sea", "user_surf", "user_weighted", "user_der2"):
        assert res.results[key].data.shape == g[key].shape
        assert relerr(res.results[key].data, g[key]) < RTOL, key
    g0 = np.load(os.path.join(GOLDEN, "golden_fe_nk4.npz"))
    if "Efermi" in g0.files and np.array_equal(g0["Efermi"], g["Efermi"]):
        assert relerr(res.results["ahc"].data, g0["ahc"]) < RTOL


def test_second_order_calculators_through_run(wb, fe):
    """NLDrude_Zeeman_{spin, orb_Omega, orb}, eMChA_FermiSurf, QuantumMetric_FermiSea / _Vel_DQ (Der2Spin / Der2Omega /
    Der2Morb / emcha_surf / tildeFab(_d)): this package's calculators -- matrices with up
    to three comma-derivatives from the CUDA kernels, batched block algebra of formula_gpu.py on the device -- through
    `run()` against the fixture of the UNMODIFIED reference (tests/golden/make_golden_second_order.py): with and without
    external terms, with wide (multi-band, Kramers-paired) groups, with use_factor=False."""
    from second_order_calcs import make_calculators
    g = np.load(os.path.join(GOLDEN, "golden_second_order.npz"))
    calcs = make_calculators(wb.calculators.static, g["Efermi"])
    res = wb.run(fe, wb.Grid(fe, NK=g["NK"], NKFFT=g["NKFFT"]), calcs, use_irred_kpt=False, symmetrize=False,
                 write_files=False)
    parity = {1: "ident", -1: "odd"}
    for key in calcs:
        r = res.results[key]
        assert r.data.shape == g[key].shape, key
        assert relerr(r.data, g[key]) < RTOL, (key, relerr(r.data, g[key]))
        assert r.transformTR == parity[int(g[key + "_TR"])] and r.transformInv == parity[int(g[key + "_Inv"])], key


def test_second_order_calculators_vs_oracle(wb, te, te_orc, orc):
    """the second-order calculators on ANOTHER system and grid than the fixture (Te, 24 WF, 2 x 2 x 3 K-blocks of 2 x 2 x 1):
    GPU path against the oracle's per-group restatement (oracle/wb_oracle.py: _Cov, Der2Spin, Der2Omega, emcha_surf, tildeFab,
    VelDQM), which is itself pinned on the reference fixture (tests/test_oracle.py::test_second_order_formulae_vs_reference)"""
    st = wb.calculators.static
    Ef = np.linspace(5.0, 7.0, 9)
    ff = dict(FF_rotAA=True)
    calcs = dict(z_spin=st.NLDrude_Zeeman_spin(Efermi=Ef), z_orb_omega=st.NLDrude_Zeeman_orb_Omega(Efermi=Ef),
                 z_orb=st.NLDrude_Zeeman_orb(Efermi=Ef),
                 emcha=st.eMChA_FermiSurf(Efermi=Ef), qmetric=st.QuantumMetric_FermiSea(Efermi=Ef, kwargs_formula=ff),
                 qmetric_dip=st.QuantumMetric_Vel_DQ(Efermi=Ef, kwargs_formula=ff))
    names = dict(z_spin="NLDrude_Zeeman_spin", z_orb_omega="NLDrude_Zeeman_orb_Omega", z_orb="NLDrude_Zeeman_orb", emcha="eMChA_FermiSurf",
                 qmetric="QuantumMetric_FermiSea", qmetric_dip="QuantumMetric_Vel_DQ")
    res = wb.run(te, wb.Grid(te, NKdiv=[2, 2, 3], NKFFT=[2, 2, 1]), calcs, use_irred_kpt=False, symmetrize=False, write_files=False)
    ref = orc.run(te_orc, [2, 2, 3], [2, 2, 1], {k: (names[k], Ef, {}) for k in calcs})
    for key in calcs:
        assert relerr(res.results[key].data, ref[key]) < RTOL, (key, relerr(res.results[key].data, ref[key]))


@pytest.mark.parametrize("nw", [4, 6, 8, 10, 12, 14, 16, 20, 22, 24])
def test_fused_rotation_kernel_sizes(wb, nw):
    """The compile-time-num_wann DMMA rotation + formula kernel (rotate_method 3; every even num_wann <= 24) against the
    size-generic GEMM + formula kernels (rotate_method 4): AHC + Morb with all R-matrices random, with the d_a H channels
    packed (columns trimmed to the bands below EFmax) and full, and on a spectrum of exactly degenerate pairs."""
    st = wb.calculators.static
    for degenerate in (False, True):
        sysg = wb.synthetic_system(nw, rmax=1, seed=100 + nw, matrices=("Ham", "AA", "BB", "CC"), degenerate_pairs=degenerate)
        eng = wb.Engine(sysg)
        eng.plan([2, 3, 2], [_lib_formula("IDENTITY")])
        E = eng.eig([0.01, 0.02, 0.03])
        eng.close()
        Ef = np.linspace(np.percentile(E, 20), np.percentile(E, 70), 41)
        # (AHC + Morb at num_wann >= 22 exceeds the shared memory of an SM: the automatic choice falls back to method 4)
        specs = st.AHC(Efermi=Ef).specs() + (st.Morb(Efermi=Ef).specs() if nw <= 20 else [])
        res = {}
        for method, packed in ((4, 1), (3, 1), (3, 0)):
            eng = wb.Engine(sysg)
            eng.set_option("rotate_method", method)
            eng.set_option("dh_packed", packed)
            eng.plan([2, 3, 2], [s.formula for s in specs])
            res[(method, packed)] = eng.scan(np.array([[0.01, 0.02, 0.03], [0.3, 0.1, 0.2]]), np.array([0.5, 0.5]), specs)
            eng.close()
        for key in ((3, 1), (3, 0)):
            for a, r in zip(res[key], res[(4, 1)]):
                assert relerr(a, r) < 1e-10, (nw, degenerate, key)


@pytest.mark.parametrize("nw,which", [(18, "fe"), (24, "te"), (10, "synth"), (40, "synth"), (64, "synth")])
def test_column_window_matches_full_rotation(wb, fe, te, nw, which):
    """Size-generic rotation with the per-k-point column window (only rows / columns of the bands of a band group are
    formed; hermitian channels mirrored; wb_rotate_gemm.cuh) against the full rotation (option rotate_trim = 0): the
    Fermi-surface product formulae, a Fermi-sea scan (window from band 0) and wide band groups, in both GEMM kernels
    (several channels per CTA for num_wann <= 32, one channel per CTA above)."""
    st = wb.calculators.static
    if which == "synth":
        sysg = wb.synthetic_system(nw, rmax=1, seed=300 + nw, matrices=("Ham", "AA", "BB", "CC", "SS"))
        eng = wb.Engine(sysg)
        eng.plan([2, 3, 2], [_lib_formula("IDENTITY")])
        E = eng.eig([0.01, 0.02, 0.03])
        eng.close()
        Ef = np.linspace(np.percentile(E, 45), np.percentile(E, 60), 31)
    else:
        sysg, Ef = (fe, np.linspace(15., 19., 31)) if which == "fe" else (te, np.linspace(4., 8., 31))
    calcs = [st.BerryDipole_FermiSurf(Efermi=Ef), st.GME_orb_FermiSurf(Efermi=Ef), st.GME_spin_FermiSurf(Efermi=Ef),
             st.Ohmic_FermiSurf(Efermi=Ef), st.Ohmic_FermiSea(Efermi=Ef), st.BerryDipole_FermiSurf(Efermi=Ef, degen_thresh=0.3),
             st.AHC(Efermi=Ef, degen_thresh=0.2, degen_Kramers=True)]
    dK = np.array([[0.01, 0.02, 0.03], [0.3, 0.1, 0.2]])
    for c in calcs:
        specs = c.specs()
        res = {}
        for trim in (0, 1):
            eng = wb.Engine(sysg)
            eng.set_option("rotate_method", 4)
            eng.set_option("rotate_trim", trim)
            eng.plan([3, 2, 3], [s.formula for s in specs])
            res[trim] = eng.scan(dK, np.array([0.5, 0.5]), specs)
            eng.close()
        for a, r in zip(res[1], res[0]):
            assert np.abs(r).max() > 0
            assert relerr(a, r) < 1e-11, (nw, which, type(c).__name__)


def test_per_block_tetra_and_kubo_scans(wb, fe):
    """wbgpu_static_scan_tetra_blocks / wbgpu_kubo_scan_blocks (per-K-block results of one batched call: what the refinement
    loop asks for) against one call per K-block of the summed entry points; K-blocks with cells of different sizes, as
    after a refinement step."""
    st, dyn = wb.calculators.static, wb.calculators.dynamic
    Ef = np.linspace(15., 19., 21)
    tcalcs = [st.AHC(Efermi=Ef, tetra=True), st.DOS(Efermi=Ef, tetra=True), st.BerryDipole_FermiSurf(Efermi=Ef, tetra=True)]
    specs = [s for c in tcalcs for s in c.specs()]
    dK = np.array([[0.01, 0.02, 0.03], [0.3, 0.1, 0.2], [0.15, 0.05, 0.4], [0.0, 0.25, 0.125], [0.4, 0.4, 0.1]])
    cells = np.array([[0.25, 0.25, 0.25]] * 3 + [[0.125, 0.125, 0.125]] * 2)
    for max_k in (0, 16):   # 16 k-points per launch = batches of two K-blocks: histograms cleared between the batches
        eng = wb.Engine(fe)
        eng.plan([2, 2, 2], [s.formula for s in specs], max_kpoints_per_launch=max_k)
        got = eng.scan_tetra_blocks(dK, cells, specs)
        for b in range(len(dK)):
            want = eng.scan_tetra(dK[b:b + 1], np.ones(1), cells[b], specs)
            for g, w in zip(got, want):
                assert relerr(g[b], w) < 1e-12, (max_k, b)
        eng.close()
    oc = dyn.OpticalConductivity(Efermi=Ef, omega=np.linspace(0., 3., 17), smr_fixed_width=0.2)
    shc = dyn.SHC(Efermi=Ef, omega=np.linspace(0., 3., 9), smr_fixed_width=0.2, SHC_type="simple")
    for c in (oc, shc):
        sp = c.spec()
        for max_k in (0, 16):
            eng = wb.Engine(fe)
            eng.plan([2, 2, 2], [_lib_formula("IDENTITY"), sp.formula_flag], max_kpoints_per_launch=max_k)
            got = eng.kubo_scan_blocks(dK, sp, c.Efermi, c.omega)
            for b in range(len(dK)):
                want = eng.kubo_scan(dK[b:b + 1], np.ones(1), sp, c.Efermi, c.omega)
                assert got[b].shape == want.shape and relerr(got[b], want) < 1e-11, (type(c).__name__, max_k, b)
            eng.close()


BLOCK_CASES = dict(
    ahc=("AHC", {}), dos=("DOS", {}), cumdos=("CumDOS", {}), Morb=("Morb", {}),
    ahc_kramers=("AHC", dict(degen_Kramers=True)), ahc_thresh=("AHC", dict(degen_thresh=0.05)),
    morb_thresh=("Morb", dict(degen_thresh=0.05)),
    bcd_thresh=("BerryDipole_FermiSurf", dict(degen_thresh=0.05)),
    gme_orb_thresh=("GME_orb_FermiSurf", dict(degen_thresh=0.05)),
    gme_spin_thresh=("GME_spin_FermiSurf", dict(degen_thresh=0.05)),
)


@pytest.mark.parametrize("case", sorted(BLOCK_CASES))
def test_block_calculators_vs_reference(wb, fe, case):
    """calc(Data_K_R) for one K-block against the reference's output (fixture)."""
    b = np.load(os.path.join(GOLDEN, "golden_fe_block.npz"))
    grid = wb.Grid(fe, NKdiv=[2, 2, 2], NKFFT=b["NKFFT"])
    data = wb.Data_K_R(fe, dK=b["dK"], grid=grid)
    name, kw = BLOCK_CASES[case]
    res = getattr(wb.calculators.static, name)(Efermi=b["Efermi"], **kw)(data)
    assert res.data.shape == b["block_" + case].shape
    assert relerr(res.data, b["block_" + case]) < RTOL


@pytest.mark.parametrize("nw", [5, 8, 12, 16, 20, 24, 27, 32])
def test_omega_synthetic_sizes(wb, orc, nw):
    """AHC scan on seeded random models of several sizes (DMMA kernel variants for nw <= 20, batched DMMA GEMM
    above) against the oracle; one K-block."""
    sysg = wb.synthetic_system(nw, rmax=1, seed=100 + nw)
    syso = orc.OracleSystem(sysg.rvec.iRvec, sysg.real_lattice, sysg.wannier_centers_cart,
                            {k: sysg.get_R_mat(k) for k in ("Ham", "AA")})
    NKFFT, dK = [3, 3, 2], [0.05, 0.11, 0.02]
    Ef = np.linspace(-3., 3., 61)
    grid = wb.Grid(sysg, NKdiv=[1, 1, 1], NKFFT=NKFFT)
    data = wb.Data_K_R(sysg, dK=dK, grid=grid)
    got = wb.calculators.static.AHC(Efermi=Ef)(data).data
    ref = orc.AHC(orc.OracleDataK(syso, dK, NKFFT), Ef)
    assert relerr(got, ref) < RTOL


@pytest.mark.parametrize("nw", [40, 64, 128])
def test_large_num_wann(wb, orc, nw):
    """BASELINE config 5 in miniature (many Wannier functions, Ham + AA): eigenvalues, DOS / CumDOS and the
    tensor-core rotation + Berry curvature (AHC) against the oracle; one small K-block."""
    sysg = wb.synthetic_system(nw, rmax=1, seed=500 + nw)
    syso = orc.OracleSystem(sysg.rvec.iRvec, sysg.real_lattice, sysg.wannier_centers_cart,
                            {k: sysg.get_R_mat(k) for k in ("Ham", "AA")})
    NKFFT, dK = [2, 1, 2], [0.05, 0.11, 0.02]
    Ef = np.linspace(-4., 4., 41)
    grid = wb.Grid(sysg, NKdiv=[1, 1, 1], NKFFT=NKFFT)
    data = wb.Data_K_R(sysg, dK=dK, grid=grid)
    odata = orc.OracleDataK(syso, dK, NKFFT)
    st = wb.calculators.static
    assert relerr(st.DOS(Efermi=Ef)(data).data, orc.DOS(odata, Ef)) < RTOL
    assert relerr(st.CumDOS(Efermi=Ef)(data).data, orc.CumDOS(odata, Ef)) < RTOL
    assert relerr(st.AHC(Efermi=Ef)(data).data, orc.AHC(odata, Ef)) < RTOL


# ---------------------------------------------------------------------------------------- Kubo path
KUBO_CASES = dict(
    ref_optcond=("OpticalConductivity", dict(smr_fixed_width=0.20, smr_type="Gaussian"), True),
    lor_optcond=("OpticalConductivity", dict(smr_fixed_width=0.1, smr_type="Lorentzian"), False),
    lor_optcond_thresh=("OpticalConductivity", dict(smr_fixed_width=0.1, smr_type="Lorentzian", degen_thresh=0.05), False),
    lor_optcond_int=("OpticalConductivity", dict(smr_fixed_width=0.1, smr_type="Lorentzian",
                                                 kwargs_formula=dict(external_terms=False)), False),
    gau_optcond=("OpticalConductivity", dict(smr_fixed_width=0.15, smr_type="Gaussian"), False),
    lor_jdos=("JDOS", dict(smr_fixed_width=0.1, smr_type="Lorentzian"), False),
    gau_jdos=("JDOS", dict(smr_fixed_width=0.15, smr_type="Gaussian"), False),
)


@pytest.mark.parametrize("case", sorted(KUBO_CASES))
def test_kubo_block_vs_reference(wb, fe, case):
    """DynamicCalculator.__call__(Data_K_R) for one K-block against the reference's output (fixture written by
    tests/golden/make_golden_kubo.py)."""
    g = np.load(os.path.join(GOLDEN, "golden_fe_kubo.npz"))
    name, kw, ref_axes = KUBO_CASES[case]
    Ef, om = (g["ref_Efermi"], g["ref_omega"]) if ref_axes else (g["Efermi"], g["omega"])
    grid = wb.Grid(fe, NKdiv=[2, 2, 2], NKFFT=g["block_NKFFT"])
    data = wb.Data_K_R(fe, dK=g["block_dK"], grid=grid)
    res = getattr(wb.calculators.dynamic, name)(Efermi=Ef, omega=om, **kw)(data)
    assert res.data.shape == g["block_" + case].shape
    assert relerr(res.data, g["block_" + case]) < RTOL


def test_kubo_run_vs_upstream_golden(wb, fe):
    """run() with Kubo and static calculators together on the reference's test grid, against the reference's own
    golden file Fe_W90-opt_conductivity_iter-0000.npz (tests/test_run.py:290-305) and the other fixtures."""
    g = np.load(os.path.join(GOLDEN, "golden_fe_kubo.npz"))
    g4 = np.load(os.path.join(GOLDEN, "golden_fe_nk4.npz"))
    dyn, st = wb.calculators.dynamic, wb.calculators.static
    calcs = {"ahc": st.AHC(Efermi=g4["Efermi"])}
    for case, (name, kw, ref_axes) in KUBO_CASES.items():
        Ef, om = (g["ref_Efermi"], g["ref_omega"]) if ref_axes else (g["Efermi"], g["omega"])
        calcs[case] = getattr(dyn, name)(Efermi=Ef, omega=om, **kw)
    grid = wb.Grid(fe, NKdiv=[2, 2, 2], NKFFT=[2, 2, 2])
    res = wb.run(fe, grid, calcs)
    assert relerr(res.results["ref_optcond"].data, g["upstream_golden_opt_conductivity"]) < RTOL
    assert relerr(res.results["ahc"].data, g4["upstream_golden_ahc"]) < RTOL
    for case in KUBO_CASES:
        assert relerr(res.results[case].data, g["run_" + case]) < RTOL, case


@pytest.mark.parametrize("nw,nEF,nom", [(32, 60, 40), (12, 1500, 3), (40, 5, 7)])
def test_kubo_synthetic(wb, orc, nw, nEF, nom):
    """BASELINE config 4 in miniature (synthetic 32-WF model, Lorentzian, kBT = 0) and the tiling corner cases of the
    accumulation kernel (several omega tiles; an Efermi axis that does not fit one shared-memory tile) against
    the oracle."""
    sysg = wb.synthetic_system(nw, rmax=1, seed=700 + nw)
    syso = orc.OracleSystem(sysg.rvec.iRvec, sysg.real_lattice, sysg.wannier_centers_cart,
                            {k: sysg.get_R_mat(k) for k in ("Ham", "AA")})
    NKFFT, dK = [2, 2, 2], [0.05, 0.11, 0.02]
    Ef, om = np.linspace(-1., 1., nEF), np.linspace(0., 5., nom)
    grid = wb.Grid(sysg, NKdiv=[1, 1, 1], NKFFT=NKFFT)
    data = wb.Data_K_R(sysg, dK=dK, grid=grid)
    odata = orc.OracleDataK(syso, dK, NKFFT)
    kw = dict(smr_fixed_width=0.1, smr_type="Lorentzian")
    got = wb.calculators.dynamic.OpticalConductivity(Efermi=Ef, omega=om, **kw)(data).data
    assert relerr(got, orc.OpticalConductivity(odata, Ef, omega=om, **kw)) < RTOL
    data.engine.set_option("kubo_method", 1)   # the per-(omega, re|im) accumulation kernel against the register-tiled one
    got1 = wb.calculators.dynamic.OpticalConductivity(Efermi=Ef, omega=om, **kw)(data).data
    data.engine.set_option("kubo_method", 0)
    assert relerr(got1, got) < 1e-12
    got = wb.calculators.dynamic.JDOS(Efermi=Ef, omega=om, **kw)(data).data
    assert relerr(got, orc.JDOS(odata, Ef, omega=om, **kw)) < RTOL


# ---------------------------------------------------------------------------------------- Kubo spin Hall conductivity
SHC_CASES = dict(ref_qiao=("qiao", "ref", {}), ref_ryoo=("ryoo", "ref", {}), in_qiao=("qiao", "in", {}), in_ryoo=("ryoo", "in", {}),
                 in_simple=("simple", "in", {}), in_ryoo_thresh=("ryoo", "in", dict(degen_thresh=0.3)))


def test_shc_random_system_vs_upstream_goldens(wb):
    """dynamic.SHC (Kubo spin Hall conductivity, spin currents of Ryoo / Qiao / {S, v}/2) on the reference's `random`
    system through run(), all six scans in one call: against the live reference run of make_golden_shc.py and the
    reference's own golden files random-opt_SHC{qiao,ryoo}_iter-0000.npz (tests/test_run.py:653-669)."""
    g = np.load(os.path.join(GOLDEN, "golden_random_shc.npz"))
    rnd = wb.System_R.from_npz(os.path.join(GOLDEN, "random_system.npz"))
    calcs = {}
    for case, (t, ax, extra) in SHC_CASES.items():
        calcs[case] = wb.calculators.dynamic.SHC(Efermi=g[ax + "_Efermi"], omega=g["omega"], smr_fixed_width=0.20,
                                                 smr_type="Gaussian" if ax == "ref" else "Lorentzian", SHC_type=t, **extra)
    calcs["abc"] = wb.calculators.dynamic.SHC(Efermi=g["in_Efermi"], omega=g["omega"], smr_fixed_width=0.20, SHC_type="ryoo",
                                              shc_abc=(1, 2, 3))
    res = wb.run(rnd, wb.Grid(rnd, NK=g["NK"], NKFFT=g["NKFFT"]), calcs)
    for case, (t, ax, extra) in SHC_CASES.items():
        got = res.results[case].data
        assert got.shape == g[case].shape, case
        assert relerr(got, g[case]) < RTOL, case
        if ax == "ref":
            assert relerr(got, g["upstream_golden_" + t]) < RTOL, case
    assert res.results["abc"].data.shape == g["in_ryoo"].shape[:2]
    assert relerr(res.results["abc"].data, g["in_ryoo"][:, :, 0, 1, 2]) < RTOL


@pytest.mark.parametrize("nw,nEF,nom,external", [(14, 40, 45, True), (33, 7, 5, False)])
def test_shc_synthetic(wb, orc, nw, nEF, nom, external):
    """SHC on a synthetic model with the R <-> -R symmetry (hermitian-packed d_aH, SS channels), random spin-current
    matrices, more bands than a warp and more frequencies than one omega tile, one K-block, against the oracle."""
    sysg = wb.synthetic_system(nw, rmax=1, seed=900 + nw, matrices=("Ham", "AA", "SS"))
    rng = np.random.default_rng(5 + nw)
    nR = sysg.rvec.nRvec
    for key, tail in (("SA", (3, 3)), ("SHA", (3, 3)), ("SR", (3, 3)), ("SH", (3,)), ("SHR", (3, 3))):
        shape = (nR, nw, nw) + tail
        sysg.set_R_mat(key, 0.1 * (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)))
    syso = orc.OracleSystem(sysg.rvec.iRvec, sysg.real_lattice, sysg.wannier_centers_cart,
                            {k: sysg.get_R_mat(k) for k in ("Ham", "AA", "SS", "SA", "SHA", "SR", "SH", "SHR")})
    NKFFT, dK = [2, 2, 2], [0.05, 0.11, 0.02]
    Ef, om = np.linspace(-1., 1., nEF), np.linspace(0., 5., nom)
    grid = wb.Grid(sysg, NKdiv=[1, 1, 1], NKFFT=NKFFT)
    data = wb.Data_K_R(sysg, dK=dK, grid=grid)
    odata = orc.OracleDataK(syso, dK, NKFFT)
    for t in ("ryoo", "qiao", "simple"):
        kw = dict(smr_fixed_width=0.1, smr_type="Lorentzian")
        got = wb.calculators.dynamic.SHC(Efermi=Ef, omega=om, SHC_type=t, kwargs_formula=dict(external_terms=external),
                                         **kw)(data).data
        ref = orc.SHC(odata, Ef, omega=om, SHC_type=t, external_terms=external, **kw)
        assert got.shape == ref.shape == (nEF, nom, 3, 3, 3)
        assert relerr(got, ref) < RTOL, t


def test_shc_errors(wb, fe):
    """unknown spin-current type (formula/covariant.py:696-697) and missing R-matrices (system_R.py:106-111)"""
    with pytest.raises(ValueError):
        wb.calculators.dynamic.SHC(Efermi=[0.], omega=[0.], SHC_type="other")
    grid = wb.Grid(fe, NKdiv=[1, 1, 1], NKFFT=[2, 2, 2])
    with pytest.raises(ValueError):   # the Fe fixture carries no SA / SHA
        wb.calculators.dynamic.SHC(Efermi=[17.], omega=[0., 1.])(wb.Data_K_R(fe, dK=[0, 0, 0], grid=grid))


# ---------------------------------------------------------------------------------------- Kubo shift / injection current
SHIFT_CASES = dict(
    shift=("ShiftCurrent", dict(sc_eta=0.1, smr_type="Lorentzian")), shift_gauss=("ShiftCurrent", dict(sc_eta=0.04, smr_type="Gaussian")),
    shift_int=("ShiftCurrent", dict(sc_eta=0.1, smr_type="Lorentzian", kwargs_formula=dict(external_terms=False))),
    shift_thresh=("ShiftCurrent", dict(sc_eta=0.1, smr_type="Lorentzian", degen_thresh=0.3)),
    injection=("InjectionCurrent", dict(smr_type="Lorentzian")), injection_gauss=("InjectionCurrent", dict(smr_type="Gaussian")),
    injection_int=("InjectionCurrent", dict(smr_type="Lorentzian", kwargs_formula=dict(external_terms=False))),
    injection_thresh=("InjectionCurrent", dict(smr_type="Lorentzian", degen_thresh=0.3)),
)


def test_shift_and_injection_current_random_system(wb):
    """Kubo shift current and injection current (calculators/dynamic.py:244-365) on the reference's `random` system through
    run(), against the live reference run of tests/golden/make_golden_shift.py."""
    g = np.load(os.path.join(GOLDEN, "golden_random_shift.npz"))
    rnd = wb.System_R.from_npz(os.path.join(GOLDEN, "random_system.npz"))
    dyn = wb.calculators.dynamic
    calcs = {k: getattr(dyn, name)(Efermi=g["Efermi"], omega=g["omega"], smr_fixed_width=0.20, **kw)
             for k, (name, kw) in SHIFT_CASES.items()}
    res = wb.run(rnd, wb.Grid(rnd, NK=g["NK"], NKFFT=g["NKFFT"]), calcs)
    for k in SHIFT_CASES:
        got = res.results[k].data
        assert got.shape == g[k].shape and got.dtype == g[k].dtype, k
        assert relerr(got, g[k]) < RTOL, k


@pytest.mark.parametrize("nw,nom", [(14, 45), (33, 5)])
def test_shift_and_injection_current_synthetic(wb, orc, nw, nom):
    """the same on a synthetic model with the R <-> -R symmetry (hermitian-packed channels), more bands than a warp and
    more frequencies than one omega tile, one K-block, against the oracle."""
    sysg = wb.synthetic_system(nw, rmax=1, seed=500 + nw)
    syso = orc.OracleSystem(sysg.rvec.iRvec, sysg.real_lattice, sysg.wannier_centers_cart,
                            {k: sysg.get_R_mat(k) for k in ("Ham", "AA")})
    NKFFT, dK = [2, 2, 2], [0.05, 0.11, 0.02]
    Ef, om = np.linspace(-1., 1., 12), np.linspace(0., 5., nom)
    data = wb.Data_K_R(sysg, dK=dK, grid=wb.Grid(sysg, NKdiv=[1, 1, 1], NKFFT=NKFFT))
    odata = orc.OracleDataK(syso, dK, NKFFT)
    kw = dict(smr_fixed_width=0.1, smr_type="Lorentzian")
    got = wb.calculators.dynamic.ShiftCurrent(Efermi=Ef, omega=om, sc_eta=0.05, **kw)(data).data
    assert relerr(got, orc.ShiftCurrent(odata, Ef, omega=om, sc_eta=0.05, **kw)) < RTOL
    got2 = wb.calculators.dynamic.InjectionCurrent(Efermi=Ef, omega=om, **kw)(data).data
    assert relerr(got2, orc.InjectionCurrent(odata, Ef, omega=om, **kw)) < RTOL
    # the per-(omega, component) accumulation kernel (kubo_method 1) against the register-tiled one (default), also at kBT > 0
    for kbt in (0., 0.05):
        res = {}
        for method in (0, 1):
            data.engine.set_option("kubo_method", method)
            res[method] = [wb.calculators.dynamic.ShiftCurrent(Efermi=Ef, omega=om, sc_eta=0.05, kBT=kbt, **kw)(data).data,
                           wb.calculators.dynamic.InjectionCurrent(Efermi=Ef, omega=om, kBT=kbt, **kw)(data).data,
                           wb.calculators.dynamic.SHC(Efermi=Ef, omega=om, SHC_type="simple", kBT=kbt, **kw)(data).data
                           if sysg.has_R_mat("SS") else np.ones(1)]
        data.engine.set_option("kubo_method", 0)
        for a, b in zip(res[0], res[1]):
            assert relerr(a, b) < 1e-11, kbt


KBT_CASES = dict(
    optcond=("OpticalConductivity", dict(kBT=0.05)), optcond_hot=("OpticalConductivity", dict(kBT=0.5, smr_type="Gaussian")),
    jdos=("JDOS", dict(kBT=0.05)), shc_ryoo=("SHC", dict(SHC_type="ryoo", kBT=0.03)),
    shift=("ShiftCurrent", dict(sc_eta=0.1, kBT=0.05)), injection=("InjectionCurrent", dict(kBT=0.01)),
    optcond_thresh=("OpticalConductivity", dict(kBT=0.05, degen_thresh=0.3)),
)


def test_kubo_finite_temperature_random_system(wb):
    """Kubo calculators at kBT > 0 (the owner value of a band group is spread over the Fermi levels within +- 30 kBT of its
    energy with the Fermi-Dirac differences) through run(), against the live reference run of make_golden_kbt.py."""
    g = np.load(os.path.join(GOLDEN, "golden_random_kbt.npz"))
    rnd = wb.System_R.from_npz(os.path.join(GOLDEN, "random_system.npz"))
    dyn = wb.calculators.dynamic
    calcs = {k: getattr(dyn, name)(Efermi=g["Efermi"], omega=g["omega"], smr_fixed_width=0.20, **kw)
             for k, (name, kw) in KBT_CASES.items()}
    res = wb.run(rnd, wb.Grid(rnd, NK=g["NK"], NKFFT=g["NKFFT"]), calcs)
    for k in KBT_CASES:
        assert res.results[k].data.shape == g[k].shape, k
        assert relerr(res.results[k].data, g[k]) < RTOL, k


def test_kubo_finite_temperature_synthetic(wb, orc):
    """32-WF model, non-uniform Fermi axis, kBT comparable to the level spacing, against the oracle"""
    sysg = wb.synthetic_system(32, rmax=1, seed=732)
    syso = orc.OracleSystem(sysg.rvec.iRvec, sysg.real_lattice, sysg.wannier_centers_cart,
                            {k: sysg.get_R_mat(k) for k in ("Ham", "AA")})
    NKFFT, dK = [2, 2, 2], [0.05, 0.11, 0.02]
    Ef, om = np.sort(np.concatenate([np.linspace(-1., 1., 30), [-0.333, 0.017, 0.018]])), np.linspace(0., 5., 40)
    data = wb.Data_K_R(sysg, dK=dK, grid=wb.Grid(sysg, NKdiv=[1, 1, 1], NKFFT=NKFFT))
    odata = orc.OracleDataK(syso, dK, NKFFT)
    kw = dict(smr_fixed_width=0.1, smr_type="Lorentzian", kBT=0.02)
    got = wb.calculators.dynamic.OpticalConductivity(Efermi=Ef, omega=om, **kw)(data).data
    assert relerr(got, orc.OpticalConductivity(odata, Ef, omega=om, **kw)) < RTOL
    got = wb.calculators.dynamic.JDOS(Efermi=Ef, omega=om, **kw)(data).data
    assert relerr(got, orc.JDOS(odata, Ef, omega=om, **kw)) < RTOL


# ---------------------------------------------------------------------------------------- tetrahedron method
TETRA_CASES = dict(
    ahc=("AHC", {}), dos=("DOS", {}), cumdos=("CumDOS", {}), Morb=("Morb", {}), spin=("Spin", {}),
    ahc_thresh=("AHC", dict(degen_thresh=0.05)), bcd=("BerryDipole_FermiSurf", {}), gme_spin=("GME_spin_FermiSurf", {}),
)
# Fermi-sea quantities (der = 0): the north-star tolerance.  Fermi-surface quantities (der = 1): the reference
# evaluates the derivative of the tetrahedron occupation as a cubic in E_F whose coefficients cancel by (E/dE)^2 ~ 1e7;
# a 1e-15 relative change of the corner energies moves ITS result by 1-3e-8 (tests/test_oracle.py::
# test_tetra_derivative_conditioning), and the corner eigenvalues of any two eigensolvers differ by more than that.
TETRA_RTOL = dict(dos=1e-6, bcd=1e-6, gme_spin=1e-6)


@pytest.mark.parametrize("case", sorted(TETRA_CASES))
def test_tetra_block_vs_reference(wb, fe, case):
    """StaticCalculator(tetra=True)(Data_K_R) for one K-block against the reference's output (fixture written by
    tests/golden/make_golden_tetra.py): corner energies, tetrahedron band groups and weights on the GPU."""
    g = np.load(os.path.join(GOLDEN, "golden_fe_tetra.npz"))
    name, kw = TETRA_CASES[case]
    grid = wb.Grid(fe, NKdiv=g["block_NKdiv"], NKFFT=g["block_NKFFT"])
    data = wb.Data_K_R(fe, dK=g["block_dK"], grid=grid)
    res = getattr(wb.calculators.static, name)(Efermi=g["Efermi"], tetra=True, **kw)(data)
    assert res.data.shape == g["block_" + case].shape
    assert relerr(res.data, g["block_" + case]) < TETRA_RTOL.get(case, RTOL)


def test_tetra_hole_like_and_emin(wb, fe):
    """tetra=True with hole_like (weights 1 - occupation and the group of the bands above the Fermi axis, up to Emax) and
    with Emin on Fermi-sea quantities (grid/tetrahedron.py:197-198, 246-266; static.py:50-52, 84-91) against the fixture of
    the unmodified reference (tests/golden/make_golden_tetra_holes.py)."""
    from test_oracle import TETRA_HOLE_CASES
    g = np.load(os.path.join(GOLDEN, "golden_fe_tetra_holes.npz"))
    st = wb.calculators.static
    calcs = {k: getattr(st, name)(Efermi=g["Efermi"], tetra=True, **kw) for k, (name, kw) in TETRA_HOLE_CASES.items()}
    res = wb.run(fe, wb.Grid(fe, NK=g["NK"], NKFFT=g["NKFFT"]), calcs, use_irred_kpt=False, symmetrize=False, write_files=False)
    for key in calcs:
        assert relerr(res.results[key].data, g[key]) < RTOL, key
    with pytest.raises(NotImplementedError):
        st.DOS(Efermi=g["Efermi"], tetra=True, hole_like=True)


def test_tetra_run_vs_reference(wb, fe):
    """run() mixing tetrahedron and plain calculators on a whole (small) grid against the reference's run()."""
    g = np.load(os.path.join(GOLDEN, "golden_fe_tetra.npz"))
    g4 = np.load(os.path.join(GOLDEN, "golden_fe_nk4.npz"))
    st = wb.calculators.static
    calcs = {case: getattr(st, name)(Efermi=g["Efermi"], tetra=True, **kw) for case, (name, kw) in TETRA_CASES.items()}
    calcs["plain_ahc"] = st.AHC(Efermi=g4["Efermi"])
    grid = wb.Grid(fe, NKdiv=[2, 2, 2], NKFFT=[2, 2, 2])
    res = wb.run(fe, grid, calcs)
    for case in TETRA_CASES:
        assert relerr(res.results[case].data, g["run_" + case]) < TETRA_RTOL.get(case, RTOL), case
    assert relerr(res.results["plain_ahc"].data, g4["upstream_golden_ahc"]) < RTOL


@pytest.mark.parametrize("NKFFT,nEF", [([1, 1, 1], 5), ([5, 3, 7], 40), ([2, 9, 1], 1), ([3, 3, 3], 9001)])
def test_odd_grid_shapes(wb, fe, fe_orc, orc, NKFFT, nEF):
    """Edge cases of the grid / Fermi axes: a single k-point per K-block, unequal odd FFT grids (R-vectors alias onto
    them, fourier/fft.py:177), one Fermi level (dEF = 0.001, static.py:55), a Fermi axis whose histogram does not
    fit shared memory; and an empty K-block list."""
    dK = np.array([0.013, 0.021, 0.007])
    Ef = np.linspace(15., 19., nEF) if nEF > 1 else np.array([17.3])
    st = wb.calculators.static
    grid = wb.Grid(fe, NKdiv=[1, 1, 1], NKFFT=NKFFT)
    data = wb.Data_K_R(fe, dK=dK, grid=grid)
    odata = orc.OracleDataK(fe_orc, dK, NKFFT)
    names = ("AHC", "DOS", "CumDOS") if nEF < 1000 else ("AHC", "DOS")
    for name in names:
        got = getattr(st, name)(Efermi=Ef)(data).data
        ref = orc.CALCULATORS[name](odata, Ef)
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() <= RTOL * max(np.abs(ref).max(), 1e-300), name
    eng = wb.Engine(fe)
    specs = st.AHC(Efermi=Ef).specs()
    eng.plan(NKFFT, [s.formula for s in specs])
    empty = eng.scan(np.zeros((0, 3)), np.zeros(0), specs)
    assert empty[0].shape == specs[0].shape and not empty[0].any()


@pytest.mark.parametrize("rmax,NKFFT", [(1, [4, 3, 5]), (2, [6, 6, 2]), (3, [5, 8, 3]), (4, [3, 2, 4])])
def test_r_to_k_variants(wb, orc, rmax, NKFFT):
    """The three implementations of the axes-1+0 pass of the R -> k transform (option fourier_method: 0 = partial sums in
    registers, R-box edges up to 8; 2 = array of accumulators; 1 = two separate passes) against the oracle's transform
    (fourier/fft.py:133-192) on R-boxes of 3, 5, 7 and 9 cells per edge (9: the default falls back to variant 2)."""
    nw = 5
    sysg = wb.synthetic_system(nw, rmax=rmax, seed=7 + rmax, matrices=("Ham", "AA"))
    osys = orc.OracleSystem(sysg.rvec.iRvec, sysg.real_lattice, sysg.wannier_centers_cart, sysg._XX_R)
    dK = np.array([0.11, 0.02, 0.31])
    data = orc.OracleDataK(osys, dK, NKFFT)
    T = osys.cRvec_shifted
    want = dict(Ham=data._R_to_k(osys.XX_R["Ham"], True), AA=data._R_to_k(osys.XX_R["AA"], True),
                dHam=data._R_to_k(orc.derivative(osys.XX_R["Ham"], T), False))
    from wannierberri_b200 import _lib
    for method in (0, 2, 1):
        eng = wb.Engine(sysg)
        eng.set_option("fourier_method", method)
        eng.plan(NKFFT, [_lib.OMEGA])
        for name, ref in want.items():
            assert relerr(eng.xk(dK, name), ref) < 1e-13, (method, name)
        eng.close()


def test_full_size_properties(wb, fe):
    """BASELINE-size launches (K-blocks of 20^3 from the 400^3 grid, 2000 Fermi levels, 16 blocks = 128 000 k-points
    per call) through size-independent properties: the band-counting sum rule, linearity in the K-block weights,
    additivity over shards, and agreement of the fused nw = 18 rotation kernel with the size-generic path."""
    st = wb.calculators.static
    Ef = np.linspace(5.0, 50.0, 2000)   # brackets the whole band structure (9.3 .. 47 eV)
    grid = wb.Grid(fe, NKdiv=[20, 20, 20], NKFFT=[20, 20, 20])
    shifts, factors = grid.K_arrays()
    sel = slice(1000, 1016)
    specs = st.CumDOS(Efermi=Ef).specs() + st.AHC(Efermi=Ef).specs()
    eng = wb.Engine(fe)
    eng.plan([20, 20, 20], [s.formula for s in specs])
    w = np.full(16, 1. / 16)
    cum, ahc = eng.scan(shifts[sel], w, specs)
    # every band is counted once per k-point: CumDOS * cell_volume = 18 above the top band, 0 below the bottom
    assert abs(cum[-1] * fe.cell_volume - 18.) < 1e-9 and cum[0] == 0.
    assert np.all(np.diff(cum) >= -1e-12)
    # the Berry curvature summed over ALL bands vanishes identically at every k-point
    assert np.abs(ahc[-1]).max() < 1e-9 * np.abs(ahc).max()
    # linearity / additivity
    cum2, ahc2 = eng.scan(shifts[sel], 3. * w, specs)
    assert relerr(ahc2, 3. * ahc) < 1e-12 and relerr(cum2, 3. * cum) < 1e-12
    parts = [eng.scan(shifts[1000 + 4 * i:1004 + 4 * i], w[:4], specs) for i in range(4)]
    assert relerr(sum(p[1] for p in parts), ahc) < 1e-11
    # fused kernel vs batched DMMA GEMM + formula kernel on the same 128 000 k-points
    eng.set_option("rotate_method", 4)
    ahc_g = eng.scan(shifts[sel], w, specs)[1]
    assert relerr(ahc_g, ahc) < 1e-10


def _oracle_blocks(path, calcs, dKs, NKFFT):
    """The oracle on a list of K-blocks, one process per block (each block is seconds of per-k Python loops):
    list over blocks of {key: data}."""
    import multiprocessing as mp
    jobs = [(path, calcs, np.asarray(dK, dtype=float), list(NKFFT)) for dK in dKs]
    if len(jobs) == 1:
        return [_oracle_block_job(jobs[0])]
    with mp.get_context("spawn").Pool(min(len(jobs), os.cpu_count() or 1)) as pool:
        return pool.map(_oracle_block_job, jobs)


def _oracle_block_job(job):
    path, calcs, dK, NKFFT = job
    from oracle import wb_oracle as o
    data = o.OracleDataK(o.OracleSystem.from_npz(path), dK, NKFFT)
    return {key: o.CALCULATORS[name](data, Ef, **kw) for key, (name, Ef, kw) in calcs.items()}


BENCH_EF = np.linspace(12.0, 22.0, 2000)   # bench.py EFERMI


@pytest.mark.parametrize("rotate_method", [0, 4])
def test_bench_config_block_vs_oracle(wb, fe, rotate_method):
    """The EXACT configuration bench.py times (BASELINE config 2 split: K-blocks of NKFFT = 20^3 of the 400^3 K-list,
    AHC + DOS, 2000 Fermi levels over 12..22 eV, default rotation method = the fused column-trimmed DMMA kernel)
    against the oracle on identical inputs at 1e-8: one K-block alone, and a call over 3 K-blocks with unequal weights
    (weighted sum and batch offsets)."""
    st = wb.calculators.static
    grid = wb.Grid(fe, NKdiv=[20, 20, 20], NKFFT=[20, 20, 20])
    shifts, factors = grid.K_arrays()
    sel = [0, 1234, 7999]
    w = np.array([0.5, 0.2, 0.3])
    calcs = dict(ahc=st.AHC(Efermi=BENCH_EF), dos=st.DOS(Efermi=BENCH_EF))
    specs = [s for c in calcs.values() for s in c.specs()]
    eng = wb.Engine(fe)
    eng.set_option("rotate_method", rotate_method)
    eng.plan([20, 20, 20], [s.formula for s in specs], external_terms=True)
    ref = _oracle_blocks(os.path.join(GOLDEN, "fe_system.npz"),
                         dict(ahc=("AHC", BENCH_EF, {}), dos=("DOS", BENCH_EF, {})), shifts[sel], [20, 20, 20])
    one = eng.scan(shifts[sel[1:2]], np.ones(1), specs)
    got = dict(ahc=calcs["ahc"].result(one[:1], fe.cell_volume).data, dos=calcs["dos"].result(one[1:], fe.cell_volume).data)
    for q in got:
        assert got[q].shape == ref[1][q].shape
        assert relerr(got[q], ref[1][q]) < RTOL, (q, "single block")
    three = eng.scan(shifts[sel], w, specs)
    got = dict(ahc=calcs["ahc"].result(three[:1], fe.cell_volume).data, dos=calcs["dos"].result(three[1:], fe.cell_volume).data)
    for q in got:
        want = sum(wi * r[q] for wi, r in zip(w, ref))
        assert relerr(got[q], want) < RTOL, (q, "three blocks")


def test_baseline_config1_block_vs_oracle(wb, fe):
    """BASELINE config 1 (48^3 grid as NKdiv = 4 x NKFFT = 12, AHC + DOS, 1001 Fermi levels 12..22 eV, no symmetry):
    two K-blocks of 12^3 through run()'s engine path against the oracle at 1e-8."""
    st = wb.calculators.static
    Ef = np.linspace(12.0, 22.0, 1001)
    shifts, factors = wb.Grid(fe, NKdiv=[4, 4, 4], NKFFT=[12, 12, 12]).K_arrays()
    sel = [5, 63]
    calcs = dict(ahc=st.AHC(Efermi=Ef), dos=st.DOS(Efermi=Ef))
    specs = [s for c in calcs.values() for s in c.specs()]
    eng = wb.Engine(fe)
    eng.plan([12, 12, 12], [s.formula for s in specs], external_terms=True)
    ref = _oracle_blocks(os.path.join(GOLDEN, "fe_system.npz"), dict(ahc=("AHC", Ef, {}), dos=("DOS", Ef, {})),
                         shifts[sel], [12, 12, 12])
    arr = eng.scan(shifts[sel], factors[sel], specs)
    got = dict(ahc=calcs["ahc"].result(arr[:1], fe.cell_volume).data, dos=calcs["dos"].result(arr[1:], fe.cell_volume).data)
    for q in got:
        assert relerr(got[q], sum(f * r[q] for f, r in zip(factors[sel], ref))) < RTOL, q


def test_baseline_config2_ahc_morb_block_vs_oracle(wb, fe):
    """BASELINE config 2 as written: AHC + Morb (BB, CC channels, non-additive Morb_Hpm) on a 20^3 K-block of the 400^3
    K-list, 2000 Fermi levels, against the oracle at 1e-8."""
    st = wb.calculators.static
    shifts, _ = wb.Grid(fe, NKdiv=[20, 20, 20], NKFFT=[20, 20, 20]).K_arrays()
    calcs = dict(ahc=st.AHC(Efermi=BENCH_EF), morb=st.Morb(Efermi=BENCH_EF))
    data = wb.Data_K_R(fe, shifts[4321], wb.Grid(fe, NKdiv=[20, 20, 20], NKFFT=[20, 20, 20]))
    ref = _oracle_blocks(os.path.join(GOLDEN, "fe_system.npz"),
                         dict(ahc=("AHC", BENCH_EF, {}), morb=("Morb", BENCH_EF, {})), shifts[4321:4322], [20, 20, 20])[0]
    for q, c in calcs.items():
        assert relerr(c(data).data, ref[q]) < RTOL, q


def test_baseline_config3_te_block_vs_oracle(wb, te):
    """BASELINE config 3: Te (24 WF, spinor), BerryDipole_FermiSurf + GME_orb_FermiSurf + GME_spin_FermiSurf, tetra=False,
    one 20^3 K-block of the 200^3 K-list (NKdiv = 10), 401 Fermi levels 4..8 eV, against the oracle at 1e-8."""
    st = wb.calculators.static
    Ef = np.linspace(4.0, 8.0, 401)
    grid = wb.Grid(te, NKdiv=[10, 10, 10], NKFFT=[20, 20, 20])
    shifts, _ = grid.K_arrays()
    names = ["BerryDipole_FermiSurf", "GME_orb_FermiSurf", "GME_spin_FermiSurf"]
    data = wb.Data_K_R(te, shifts[321], grid)
    ref = _oracle_blocks(os.path.join(GOLDEN, "te_system.npz"), {n: (n, Ef, {}) for n in names}, shifts[321:322],
                         [20, 20, 20])[0]
    for n in names:
        assert relerr(getattr(st, n)(Efermi=Ef)(data).data, ref[n]) < RTOL, n


def test_full_size_properties_next_rows(wb, te):
    """The kernels of the next-row formulae at BASELINE-size K-blocks (Te, 24 WF, NKFFT = 20^3): the rotated matrices of
    one call are processed in several sub-batches (33 .. 57 matrices per k-point), so additivity over the K-blocks of a
    call checks the sub-batch offsets; and physical identities: the generalised derivative of the Berry curvature summed
    over ALL bands vanishes, the rank-3 Kubo scans are additive as well."""
    st, dyn = wb.calculators.static, wb.calculators.dynamic
    Ef = np.linspace(-5.0, 25.0, 301)   # brackets the whole band structure of the Te model
    shifts, factors = wb.Grid(te, NKdiv=[10, 10, 10], NKFFT=[20, 20, 20]).K_arrays()
    sel = slice(500, 503)
    w = np.full(3, 1. / 3)
    calcs = [st.BerryDipole_FermiSea(Efermi=Ef), st.GME_orb_FermiSea(Efermi=Ef), st.NLDrude_FermiSea(Efermi=Ef),
             st.AHC_Zeeman_spin(Efermi=Ef), st.SHC(Efermi=Ef, kwargs_formula=dict(spin_current_type="simple"))]
    specs = [s for c in calcs for s in c.specs()]
    eng = wb.Engine(te)
    eng.plan([20, 20, 20], [s.formula for s in specs])
    whole = eng.scan(shifts[sel], w, specs)
    parts = [eng.scan(shifts[500 + i:501 + i], w[:1], specs) for i in range(3)]
    for i, s in enumerate(specs):
        assert relerr(sum(p[i] for p in parts), whole[i]) < 1e-11, i
    bd = whole[0]
    assert np.abs(bd[-1]).max() < 1e-9 * np.abs(bd).max()   # all bands occupied: d_d Omega_c summed over all bands = 0
    s32 = wb.synthetic_system(32, rmax=2, seed=20261017)
    sh32, _ = wb.Grid(s32, NKdiv=[8, 8, 8], NKFFT=[16, 16, 16]).K_arrays()
    sc = dyn.ShiftCurrent(Efermi=np.linspace(-1, 1, 20), omega=np.linspace(0, 5, 70), sc_eta=0.05, smr_fixed_width=0.1, kBT=0.02)
    e32 = wb.Engine(s32)
    e32.plan([16, 16, 16], [sc.spec().formula_flag], external_terms=True)
    whole = e32.kubo_scan(sh32[7:9], np.ones(2), sc.spec(), sc.Efermi, sc.omega)
    parts = [e32.kubo_scan(sh32[7 + i:8 + i], np.ones(1), sc.spec(), sc.Efermi, sc.omega) for i in range(2)]
    assert relerr(parts[0] + parts[1], whole) < 1e-11


def test_run_symmetric_vs_upstream_golden(wb):
    """The reference's default run mode (use_irred_kpt=True, symmetrize=True): symmetry-reduced K-list, results
    symmetrised over the 16 operations of the magnetic point group of bcc Fe -- against the reference's own golden
    files Fe_W90_sym-{ahc,dos,cumdos,Morb,spin,opt_conductivity}_iter-0000.npz (tests/test_run.py:432-459) and the
    other quantities of the fixture (tests/golden/make_golden_sym.py)."""
    g = np.load(os.path.join(GOLDEN, "golden_fe_sym.npz"))
    fe = wb.System_R.from_npz(os.path.join(GOLDEN, "fe_system.npz"), pointgroup=["C4z", "C2x*TimeReversal", "Inversion"])
    Ef = g["Efermi"]
    st, dyn = wb.calculators.static, wb.calculators.dynamic
    calcs = dict(ahc=st.AHC(Efermi=Ef), dos=st.DOS(Efermi=Ef), cumdos=st.CumDOS(Efermi=Ef), Morb=st.Morb(Efermi=Ef),
                 spin=st.Spin(Efermi=Ef), berry_dipole_fsurf=st.BerryDipole_FermiSurf(Efermi=Ef),
                 gme_orb_fsurf=st.GME_orb_FermiSurf(Efermi=Ef), gme_spin_fsurf=st.GME_spin_FermiSurf(Efermi=Ef),
                 ahc_tetra=st.AHC(Efermi=Ef, tetra=True), dos_tetra=st.DOS(Efermi=Ef, tetra=True),
                 opt_conductivity=dyn.OpticalConductivity(Efermi=g["opt_Efermi"], omega=g["opt_omega"],
                                                          smr_fixed_width=0.20, smr_type="Gaussian"))
    res = wb.run(fe, wb.Grid(fe, NK=[4, 4, 4], NKFFT=[2, 2, 2]), calcs, use_irred_kpt=True, symmetrize=True)
    for q in ("ahc", "dos", "cumdos", "Morb", "spin", "opt_conductivity"):
        assert relerr(res.results[q].data, g["upstream_golden_" + q]) < RTOL, q
    for q in calcs:
        assert res.results[q].data.shape == g[q].shape
        if np.abs(g[q]).max() < 1e-10:   # forbidden by symmetry (inversion: no Berry dipole / gyrotropy): noise in both
            assert np.abs(res.results[q].data).max() < 1e-10, q
        else:
            assert relerr(res.results[q].data, g[q]) < (1e-6 if q == "dos_tetra" else RTOL), q
    # symmetrising a full-grid run leaves only what the magnetic point group allows: sigma_z for the AHC
    full = wb.run(fe, wb.Grid(fe, NK=[4, 4, 4], NKFFT=[2, 2, 2]), dict(ahc=calcs["ahc"]), use_irred_kpt=False, symmetrize=True)
    a = full.results["ahc"].data
    assert np.abs(a[:, :2]).max() < 1e-12 * np.abs(a[:, 2]).max()


def test_run_defaults_are_the_references(wb, tmp_path):
    """run(system, grid, calculators) with NO further arguments is the reference's default mode (run_grid.py:118-142,
    233-234, 368): symmetry-irreducible K-list + symmetrised results (the upstream `Fe_W90_sym-*` goldens were made this
    way), `use_irred_kpt=True` forces `symmetrize` on, the result files `<fout_name>-<key>_iter-0000.{npz,dat}` are
    written, and a system without a point group runs as one with the identity group."""
    g = np.load(os.path.join(GOLDEN, "golden_fe_sym.npz"))
    fe = wb.System_R.from_npz(os.path.join(GOLDEN, "fe_system.npz"), pointgroup=["C4z", "C2x*TimeReversal", "Inversion"])
    Ef = g["Efermi"]
    st = wb.calculators.static
    calcs = dict(ahc=st.AHC(Efermi=Ef), Morb=st.Morb(Efermi=Ef, save_mode="bin"))
    grid = wb.Grid(fe, NK=[4, 4, 4], NKFFT=[2, 2, 2])
    res = wb.run(fe, grid, calcs, fout_name=str(tmp_path / "out"))
    for q in calcs:
        assert relerr(res.results[q].data, g["upstream_golden_" + q]) < RTOL, q
    forced = wb.run(fe, grid, calcs, use_irred_kpt=True, symmetrize=False, write_files=False)   # must not skip the symmetrisation
    for q in calcs:
        assert relerr(forced.results[q].data, g["upstream_golden_" + q]) < RTOL, q
    from wannierberri_b200.result import EnergyResult
    back = EnergyResult.from_npz(str(tmp_path / "out-ahc_iter-0000.npz"))
    assert np.array_equal(back.data, res.results["ahc"].data) and back.transformTR == "odd"
    assert (tmp_path / "out-ahc_iter-0000.dat").is_file() and (tmp_path / "out-Morb_iter-0000.npz").is_file()
    assert not (tmp_path / "out-Morb_iter-0000.dat").exists()
    bare = wb.System_R.from_npz(os.path.join(GOLDEN, "fe_system.npz"))     # no point group: identity
    a = wb.run(bare, wb.Grid(bare, NK=[4, 4, 4], NKFFT=[2, 2, 2]), dict(ahc=calcs["ahc"]), write_files=False)
    b = wb.run(bare, wb.Grid(bare, NK=[4, 4, 4], NKFFT=[2, 2, 2]), dict(ahc=calcs["ahc"]), use_irred_kpt=False, symmetrize=False,
               write_files=False)
    assert relerr(a.results["ahc"].data, b.results["ahc"].data) < 1e-12   # (the scan sums with atomics: not bit-reproducible)
    g0 = np.load(os.path.join(GOLDEN, "golden_fe_nk4.npz"))
    if "upstream_golden_ahc" in g0.files:
        assert relerr(a.results["ahc"].data, g0["upstream_golden_ahc"]) < RTOL


@pytest.mark.parametrize("fder", [0, 1, 2, 3])
def test_fder_stencils_vs_upstream_files(wb, fe, fder):
    """StaticCalculator(Formula=Identity, fder=0..3): the finite-difference stencils of the scan (static.py:137-147)
    against the reference's own files tests/reference/calculators/calculator-Fe-ident-fder=*.npz."""
    from wannierberri_b200 import _lib
    g = np.load(os.path.join(GOLDEN, "golden_fe_calc_fder.npz"))
    grid = wb.Grid(fe, NKdiv=[1, 1, 1], NKFFT=g["NKFFT"])
    data = wb.Data_K_R(fe, dK=g["dK"], grid=grid)
    got = wb.calculators.static.StaticCalculator(Formula=_lib.IDENTITY, Efermi=g["Efermi"], tetra=False, fder=fder)(data).data
    ref = g[f"upstream_ident_fder{fder}"]
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= RTOL * max(np.abs(ref).max(), 1e-300)


def test_random_system_vs_upstream_goldens(wb):
    """The reference's `random` test system (6 WF, 20 R-vectors WITHOUT the R <-> -R symmetry -> the plan keeps d_aH as
    full channels) through run(): static, tetrahedron and Kubo calculators against the reference's own golden files
    random-*_iter-0000.npz (tests/test_run.py:653-669) and the live reference run of make_golden_random.py."""
    g = np.load(os.path.join(GOLDEN, "golden_random.npz"))
    rnd = wb.System_R.from_npz(os.path.join(GOLDEN, "random_system.npz"))
    Ef = g["Efermi"]
    st, dyn = wb.calculators.static, wb.calculators.dynamic
    calcs = dict(ahc=st.AHC(Efermi=Ef), dos=st.DOS(Efermi=Ef), cumdos=st.CumDOS(Efermi=Ef), Morb=st.Morb(Efermi=Ef),
                 spin=st.Spin(Efermi=Ef), conductivity_ohmic_fsurf=st.Ohmic_FermiSurf(Efermi=Ef),
                 berry_dipole_fsurf=st.BerryDipole_FermiSurf(Efermi=Ef), gme_orb_fsurf=st.GME_orb_FermiSurf(Efermi=Ef),
                 gme_spin_fsurf=st.GME_spin_FermiSurf(Efermi=Ef), ahc_tetra=st.AHC(Efermi=Ef, tetra=True),
                 opt_conductivity=dyn.OpticalConductivity(Efermi=g["opt_Efermi"], omega=g["opt_omega"],
                                                          smr_fixed_width=0.20, smr_type="Gaussian"),
                 opt_conductivity_in=dyn.OpticalConductivity(Efermi=g["opt_in_Efermi"], omega=g["opt_omega"],
                                                             smr_fixed_width=0.20, smr_type="Lorentzian"))
    res = wb.run(rnd, wb.Grid(rnd, NK=[6, 6, 6], NKFFT=[3, 3, 3]), calcs)
    for q in calcs:
        assert res.results[q].data.shape == g[q].shape, q
        assert np.abs(res.results[q].data - g[q]).max() <= RTOL * max(np.abs(g[q]).max(), 1e-300), q
        if "upstream_golden_" + q in g.files:
            ref = g["upstream_golden_" + q]
            assert np.abs(res.results[q].data - ref).max() <= RTOL * max(np.abs(ref).max(), 1e-300), q
    # every rotation / eigensolver variant on this system
    for method in (1, 4):
        eng = wb.Engine(rnd)
        eng.set_option("rotate_method", method)
        specs = calcs["ahc"].specs() + calcs["Morb"].specs()
        eng.plan([3, 3, 3], [s.formula for s in specs])
        shifts, factors = wb.Grid(rnd, NK=[6, 6, 6], NKFFT=[3, 3, 3]).K_arrays()
        a = eng.scan(shifts, factors, specs)[0]
        assert relerr(a, g["ahc"]) < RTOL, method


@pytest.mark.parametrize("tag", ["fe", "random"])
@pytest.mark.parametrize("rotate_method", [0, 1])
def test_ohmic_fermi_sea_vs_upstream_golden(wb, fe, tag, rotate_method):
    """Ohmic_FermiSea (formula InvMass: second comma-derivative channels of H, rotated diagonals, generalised
    derivative) against the reference's own golden files {Fe_W90,random}-conductivity_ohmic_iter-0000.npz, with
    degenerate groups and with the tetrahedron method; hermitian-packed (Fe) and full (random) derivative channels."""
    g = np.load(os.path.join(GOLDEN, "golden_ohmic_sea.npz"))
    system = fe if tag == "fe" else wb.System_R.from_npz(os.path.join(GOLDEN, "random_system.npz"))
    NK, NKFFT = ([4, 4, 4], [2, 2, 2]) if tag == "fe" else ([6, 6, 6], [3, 3, 3])
    Ef = g[f"{tag}_Efermi"]
    st = wb.calculators.static
    calcs = dict(ohmic=st.Ohmic_FermiSea(Efermi=Ef), ohmic_thresh=st.Ohmic_FermiSea(Efermi=Ef, degen_thresh=0.05),
                 ohmic_tetra=st.Ohmic_FermiSea(Efermi=Ef, tetra=True), ahc=st.AHC(Efermi=Ef))
    from wannierberri_b200.data_K import engine_for
    engine_for(system, 0).set_option("rotate_method", rotate_method)
    try:
        res = wb.run(system, wb.Grid(system, NK=NK, NKFFT=NKFFT), calcs, device=0)
    finally:
        engine_for(system, 0).set_option("rotate_method", 0)
    assert relerr(res.results["ohmic"].data, g[f"{tag}_upstream_golden_ohmic"]) < RTOL
    for q in ("ohmic", "ohmic_thresh", "ohmic_tetra"):
        assert relerr(res.results[q].data, g[f"{tag}_{q}"]) < RTOL, q


def test_te_qe_tetra_symmetric_vs_upstream_goldens(wb):
    """The reference's Te test (tests/test_run.py:1101-1119; BASELINE config 3 family): 24-WF spinor system, tetrahedron
    method, symmetry-reduced K-list (point group C3z, C2x, TimeReversal) + symmetrisation, NK = [3,3,4] with
    NKFFT = [1,1,4] -- against the reference's own golden files Te_QE-{dos,cumdos,GME_orb_FermiSurf}_iter-0000.npz and
    the other quantities of the live reference run (tests/golden/make_golden_te_qe.py)."""
    g = np.load(os.path.join(GOLDEN, "golden_te_qe.npz"))
    te = wb.System_R.from_npz(os.path.join(GOLDEN, "te_system.npz"), pointgroup=["C3z", "C2x", "TimeReversal"])
    grid = wb.Grid(te, NK=[3, 3, 4], NKFFT=[1, 1, 4])
    shifts, factors = grid.K_arrays(use_symmetry=True)
    assert np.array_equal(shifts, g["K_list_Kp_fullBZ"]) and np.array_equal(factors, g["K_list_factor"])
    Ef = g["Efermi"]
    st = wb.calculators.static
    calcs = dict(dos=st.DOS(Efermi=Ef, tetra=True), cumdos=st.CumDOS(Efermi=Ef, tetra=True),
                 GME_orb_FermiSurf=st.GME_orb_FermiSurf(Efermi=Ef, tetra=True),
                 GME_spin_FermiSurf=st.GME_spin_FermiSurf(Efermi=Ef, tetra=True),
                 berry_dipole_fsurf=st.BerryDipole_FermiSurf(Efermi=Ef, tetra=True), ahc=st.AHC(Efermi=Ef, tetra=True))
    res = wb.run(te, grid, calcs, use_irred_kpt=True, symmetrize=True)
    der1 = ("dos", "GME_orb_FermiSurf", "GME_spin_FermiSurf", "berry_dipole_fsurf")   # der = 1 tetrahedron weights: 1e-6
    for q in calcs:
        tol = 1e-6 if q in der1 else RTOL
        scale = np.abs(g[q]).max()
        if q == "ahc":   # forbidden by time reversal: rounding noise (~1e-9 S/m) in the reference and here
            assert scale < 1e-6 and np.abs(res.results[q].data).max() < 1e-6, q
            continue
        assert np.abs(res.results[q].data - g[q]).max() <= tol * scale, q
        if "upstream_golden_" + q in g.files:
            ref = g["upstream_golden_" + q]
            assert np.abs(res.results[q].data - ref).max() <= tol * np.abs(ref).max(), q


# ---------------------------------------------------------------------------------------- Fermi-sea formulae of next-row 4
FSEA_CASES = dict(
    bd_sea=("BerryDipole_FermiSea", {}), bd_sea_int=("BerryDipole_FermiSea", dict(kwargs_formula=dict(external_terms=False))),
    bd_sea_thresh=("BerryDipole_FermiSea", dict(degen_thresh=0.3)), bd_sea_tetra=("BerryDipole_FermiSea", dict(tetra=True)),
    nlahc_sea=("NLAHC_FermiSea", {}),
    shc_ryoo=("SHC", dict(kwargs_formula=dict(spin_current_type="ryoo"))),
    shc_qiao=("SHC", dict(kwargs_formula=dict(spin_current_type="qiao"))),
    shc_simple=("SHC", dict(kwargs_formula=dict(spin_current_type="simple"))),
    shc_simple_int=("SHC", dict(kwargs_formula=dict(spin_current_type="simple", external_terms=False))),
    shc_ryoo_thresh=("SHC", dict(degen_thresh=0.3, kwargs_formula=dict(spin_current_type="ryoo"))),
    shc_qiao_tetra=("SHC", dict(tetra=True, kwargs_formula=dict(spin_current_type="qiao"))),
    gme_spin_sea=("GME_spin_FermiSea", {}), gme_spin_sea_tetra=("GME_spin_FermiSea", dict(tetra=True)),
    nldrude_fsurf=("NLDrude_FermiSurf", {}), nldrude_fder2=("NLDrude_Fermider2", {}),
    nldrude_fsurf_thresh=("NLDrude_FermiSurf", dict(degen_thresh=0.3)),
    hall_fsurf=("Hall_classic_FermiSurf", {}), hall_sea=("Hall_classic_FermiSea", {}),
    hall_fsurf_thresh=("Hall_classic_FermiSurf", dict(degen_thresh=0.3)), hall_sea_tetra=("Hall_classic_FermiSea", dict(tetra=True)),
    ahc_zeeman_spin=("AHC_Zeeman_spin", {}), ahc_zeeman_spin_thresh=("AHC_Zeeman_spin", dict(degen_thresh=0.3)),
    ahc_zeeman_spin_int=("AHC_Zeeman_spin", dict(kwargs_formula=dict(external_terms=False))),
    omegaomega=("OmegaOmega", {}), nlahc_fsurf=("NLAHC_FermiSurf", {}),
    nldrude_sea=("NLDrude_FermiSea", {}), nldrude_sea_thresh=("NLDrude_FermiSea", dict(degen_thresh=0.3)),
    nldrude_sea_tetra=("NLDrude_FermiSea", dict(tetra=True)),
    gme_orb_sea=("GME_orb_FermiSea", {}), gme_orb_sea_thresh=("GME_orb_FermiSea", dict(degen_thresh=0.3)),
    gme_orb_sea_int=("GME_orb_FermiSea", dict(kwargs_formula=dict(external_terms=False))),
    gme_orb_sea_tetra=("GME_orb_FermiSea", dict(tetra=True)),
    ahc_zeeman_orb=("AHC_Zeeman_orb", {}), ahc_zeeman_orb_thresh=("AHC_Zeeman_orb", dict(degen_thresh=0.3)),
    ahc_zeeman_orb_int=("AHC_Zeeman_orb", dict(kwargs_formula=dict(external_terms=False))),
)


def test_fermi_sea_formulae_random_system(wb):
    """DerOmega (BerryDipole_FermiSea, NLAHC_FermiSea; formula/covariant.py:212-259) and SpinOmega (static.SHC; :759-789)
    on the reference's `random` system (full d_aH channels, degenerate groups, tetrahedron method) through run(), all
    scans in one call, against the live reference run of tests/golden/make_golden_fsea.py."""
    g = np.load(os.path.join(GOLDEN, "golden_fsea.npz"))
    rnd = wb.System_R.from_npz(os.path.join(GOLDEN, "random_system.npz"))
    st = wb.calculators.static
    calcs = {k: getattr(st, name)(Efermi=g["rnd_Efermi"], **kw) for k, (name, kw) in FSEA_CASES.items()}
    res = wb.run(rnd, wb.Grid(rnd, NK=g["rnd_NK"], NKFFT=g["rnd_NKFFT"]), calcs)
    for k in FSEA_CASES:
        assert res.results[k].data.shape == g["rnd_" + k].shape, k
        assert relerr(res.results[k].data, g["rnd_" + k]) < RTOL, k


def test_te_qe_berry_dipole_fermi_sea_vs_upstream_goldens(wb):
    """The reference's Te test (tests/test_run.py:1101-1119): BerryDipole_FermiSea and NLAHC_FermiSea with the tetrahedron
    method on the symmetry-reduced K-list, against the reference's own golden files
    Te_QE-{BerryDipole_FermiSea,berry_dipole,AHC_Zeeman_spin}_iter-0000.npz; and without tetrahedra against the live
    reference run.  (AHC_Zeeman_spin: der = 1 tetrahedron weights, 1e-6 as in test_te_qe_tetra_symmetric_vs_upstream_goldens.)"""
    g = np.load(os.path.join(GOLDEN, "golden_fsea.npz"))
    te = wb.System_R.from_npz(os.path.join(GOLDEN, "te_system.npz"), pointgroup=["C3z", "C2x", "TimeReversal"])
    Ef = g["te_Efermi"]
    st = wb.calculators.static
    calcs = dict(BerryDipole_FermiSea=st.BerryDipole_FermiSea(Efermi=Ef, tetra=True),
                 berry_dipole=st.NLAHC_FermiSea(Efermi=Ef, tetra=True),
                 BerryDipole_FermiSea_notetra=st.BerryDipole_FermiSea(Efermi=Ef),
                 AHC_Zeeman_spin=st.AHC_Zeeman_spin(Efermi=Ef, tetra=True),
                 NLDrude_FermiSea=st.NLDrude_FermiSea(Efermi=Ef, tetra=True),
                 GME_orb_FermiSea=st.GME_orb_FermiSea(Efermi=Ef, tetra=True))
    res = wb.run(te, wb.Grid(te, NK=g["te_NK"], NKFFT=g["te_NKFFT"]), calcs, use_irred_kpt=True, symmetrize=True)
    for q in calcs:
        tol = 1e-6 if q == "AHC_Zeeman_spin" else RTOL
        if q == "NLDrude_FermiSea":   # odd under time reversal: vanishes on Te; ~1e-17 of rounding noise in the reference
            assert np.abs(g["te_" + q]).max() < 1e-14 and np.abs(res.results[q].data).max() < 1e-14   # (O(1) on `random`)
            continue
        assert relerr(res.results[q].data, g["te_" + q]) < tol, q
        if "te_upstream_golden_" + q in g.files:
            assert relerr(res.results[q].data, g["te_upstream_golden_" + q]) < tol, q


@pytest.mark.parametrize("nw,pairs", [(12, True), (35, False), (64, False)])
def test_fermi_sea_formulae_synthetic(wb, orc, nw, pairs):
    """DerOmega and SpinOmega on synthetic models with the R <-> -R symmetry (hermitian-packed derivative channels):
    exactly degenerate pairs (groups of two bands, D = 0 inside a group) and more bands than a warp, against the oracle."""
    sysg = wb.synthetic_system(nw, rmax=1, seed=300 + nw, matrices=("Ham", "AA", "SS", "BB", "CC"), degenerate_pairs=pairs)
    rng = np.random.default_rng(11 + nw)
    nR = sysg.rvec.nRvec
    for key, tail in (("SA", (3, 3)), ("SHA", (3, 3))):
        shape = (nR, nw, nw) + tail
        sysg.set_R_mat(key, 0.1 * (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)))
    syso = orc.OracleSystem(sysg.rvec.iRvec, sysg.real_lattice, sysg.wannier_centers_cart,
                            {k: sysg.get_R_mat(k) for k in ("Ham", "AA", "SS", "SA", "SHA", "BB", "CC")})
    NKFFT, dK = [2, 2, 2], [0.05, 0.11, 0.02]
    Ef = np.linspace(-1., 1., 21)
    data = wb.Data_K_R(sysg, dK=dK, grid=wb.Grid(sysg, NKdiv=[1, 1, 1], NKFFT=NKFFT))
    odata = orc.OracleDataK(syso, dK, NKFFT)
    st = wb.calculators.static
    got = st.BerryDipole_FermiSea(Efermi=Ef)(data).data
    assert relerr(got, orc.BerryDipole_FermiSea(odata, Ef)) < RTOL
    got = st.BerryDipole_FermiSea(Efermi=Ef, kwargs_formula=dict(external_terms=False))(data).data
    assert relerr(got, orc.BerryDipole_FermiSea(odata, Ef, kwargs_formula=dict(external_terms=False))) < RTOL
    got = st.SHC(Efermi=Ef, kwargs_formula=dict(spin_current_type="ryoo"))(data).data
    assert relerr(got, orc.SHC_static(odata, Ef, kwargs_formula=dict(spin_current_type="ryoo"))) < RTOL
    got = st.NLDrude_FermiSea(Efermi=Ef)(data).data
    assert relerr(got, orc.NLDrude_FermiSea(odata, Ef)) < RTOL
    got = st.GME_orb_FermiSea(Efermi=Ef)(data).data
    assert relerr(got, orc.GME_orb_FermiSea(odata, Ef)) < RTOL
    got = st.Hall_classic_FermiSea(Efermi=Ef)(data).data
    assert relerr(got, orc.Hall_classic_FermiSea(odata, Ef)) < RTOL
    got = st.AHC_Zeeman_spin(Efermi=Ef)(data).data
    assert relerr(got, orc.AHC_Zeeman_spin(odata, Ef)) < RTOL


def test_ohmic_fsurf_vs_upstream_golden(wb, fe):
    """Ohmic_FermiSurf (formula VelVel) against the reference's own golden file
    Fe_W90-conductivity_ohmic_fsurf_iter-0000.npz, with degenerate groups, and with the tetrahedron method."""
    g = np.load(os.path.join(GOLDEN, "golden_fe_ohmic.npz"))
    st = wb.calculators.static
    calcs = dict(ohmic_fsurf=st.Ohmic_FermiSurf(Efermi=g["Efermi"]),
                 ohmic_fsurf_thresh=st.Ohmic_FermiSurf(Efermi=g["Efermi"], degen_thresh=0.05),
                 ohmic_fsurf_tetra=st.Ohmic_FermiSurf(Efermi=g["Efermi"], tetra=True))
    res = wb.run(fe, wb.Grid(fe, NK=[4, 4, 4], NKFFT=[2, 2, 2]), calcs)
    assert relerr(res.results["ohmic_fsurf"].data, g["upstream_golden_ohmic_fsurf"]) < RTOL
    assert relerr(res.results["ohmic_fsurf_thresh"].data, g["ohmic_fsurf_thresh"]) < RTOL
    assert relerr(res.results["ohmic_fsurf_tetra"].data, g["ohmic_fsurf_tetra"]) < 1e-6   # der = 1 tetrahedron weights


def test_adaptive_refinement_symmetric(wb):
    """Refinement on the symmetry-reduced K-list (the reference's test_Fe_sym_refine, tests/test_run.py:557-578):
    per-K-point results symmetrised before the selection, children of a divided K-point merged with their symmetry
    equivalents (grid/Kpoint.py:146-216).  Which K-points get refined depends on ALL calculators of the run (through
    ResultDict.max): with the reference's own calculator set (minus its two `_test` duplicates of ahc / Morb) the
    reference's golden files Fe_W90_sym-*_iter-000{1,2,3}.npz are reproduced; with five calculators the selection
    differs from iteration 2 on and the fixture of the live reference run (make_golden_sym_adpt.py) is the check."""
    g = np.load(os.path.join(GOLDEN, "golden_fe_sym_adpt.npz"))
    fe = wb.System_R.from_npz(os.path.join(GOLDEN, "fe_system.npz"), pointgroup=["C4z", "C2x*TimeReversal", "Inversion"])
    Ef = g["Efermi"]
    st = wb.calculators.static
    for n_iter in (1, 2, 3):
        for full in (False, True):
            calcs = dict(ahc=st.AHC(Efermi=Ef), dos=st.DOS(Efermi=Ef), cumdos=st.CumDOS(Efermi=Ef), Morb=st.Morb(Efermi=Ef),
                         spin=st.Spin(Efermi=Ef))
            if full:
                calcs = dict(ahc=st.AHC(Efermi=Ef), conductivity_ohmic=st.Ohmic_FermiSea(Efermi=Ef),
                             conductivity_ohmic_fsurf=st.Ohmic_FermiSurf(Efermi=Ef), Morb=st.Morb(Efermi=Ef),
                             dos=st.DOS(Efermi=Ef), cumdos=st.CumDOS(Efermi=Ef), spin=st.Spin(Efermi=Ef))
            res = wb.run(fe, wb.Grid(fe, NK=[4, 4, 4], NKFFT=[2, 2, 2]), calcs, use_irred_kpt=True, symmetrize=True,
                         adpt_num_iter=n_iter)
            for q in calcs:
                assert relerr(res.results[q].data, g[f"{'full_' if full else ''}iter{n_iter}_{q}"]) < RTOL, (n_iter, full, q)
                if full or n_iter == 1:
                    assert relerr(res.results[q].data, g[f"upstream_golden_iter{n_iter}_{q}"]) < RTOL, (n_iter, full, q)


def test_adaptive_refinement(wb, fe, orc):
    """run(adpt_num_iter > 0): per-K-block results from the GPU + the reference's refinement loop, against the
    reference's own run() on a model without symmetry (fixture of tests/golden/make_golden_adpt.py); per-K-block
    results must also add up to the weighted scan."""
    g = np.load(os.path.join(GOLDEN, "golden_synth_adpt.npz"))
    sysg = wb.synthetic_system(6, rmax=1, seed=4242)
    st = wb.calculators.static
    for n_iter in (0, 1, 3):
        calcs = dict(ahc=st.AHC(Efermi=g["Efermi"]), dos=st.DOS(Efermi=g["Efermi"]))
        grid = wb.Grid(sysg, NKdiv=[2, 2, 2], NKFFT=[3, 3, 3])
        res = wb.run(sysg, grid, calcs, adpt_num_iter=n_iter, adpt_fac=2, adpt_mesh=2) if n_iter else wb.run(sysg, grid, calcs)
        for q in calcs:
            assert relerr(res.results[q].data, g[f"iter{n_iter}_{q}"]) < RTOL, (n_iter, q)
    # blocks add up (Fe, more K-blocks than one launch holds when max_kpoints_per_launch is small)
    specs = st.AHC(Efermi=g["Efermi"] + 17.).specs() + st.DOS(Efermi=g["Efermi"] + 17.).specs()
    eng = wb.Engine(fe)
    eng.plan([2, 2, 2], [s.formula for s in specs], max_kpoints_per_launch=24)
    shifts, factors = wb.Grid(fe, NKdiv=[2, 2, 2], NKFFT=[2, 2, 2]).K_arrays()
    whole = eng.scan(shifts, factors, specs)
    blocks = eng.scan_blocks(shifts, specs)
    for w, b in zip(whole, blocks):
        assert b.shape[0] == len(factors)
        assert relerr(np.tensordot(factors, b, axes=(0, 0)), w) < 1e-12


def test_adaptive_refinement_restart(wb, tmp_path):
    """Restart files (run_grid.py:271-298,343-363): a run with `allow_restart` followed by `restart=True` for more
    iterations gives the result of the straight run -- the reference's own test of this (tests/test_run.py:556-578)
    against the same refinement fixture; also restarting from an earlier iteration of the saved history."""
    g = np.load(os.path.join(GOLDEN, "golden_synth_adpt.npz"))
    sysg = wb.synthetic_system(6, rmax=1, seed=4242)
    st = wb.calculators.static
    mk = lambda: dict(ahc=st.AHC(Efermi=g["Efermi"]), dos=st.DOS(Efermi=g["Efermi"]))
    grid = lambda: wb.Grid(sysg, NKdiv=[2, 2, 2], NKFFT=[3, 3, 3])
    kl = str(tmp_path / "klist")
    kw = dict(adpt_fac=2, adpt_mesh=2, file_Klist_path=kl)
    res = wb.run(sysg, grid(), mk(), adpt_num_iter=1, allow_restart=True, **kw)
    for q in ("ahc", "dos"):
        assert relerr(res.results[q].data, g[f"iter1_{q}"]) < RTOL, q
    res = wb.run(sysg, grid(), mk(), adpt_num_iter=2, restart=True, allow_restart=True, **kw)
    for q in ("ahc", "dos"):
        assert relerr(res.results[q].data, g[f"iter3_{q}"]) < RTOL, q
    assert sorted(os.listdir(kl)) == ["K_list.pickle"] + [f"factors_iter-{i:08d}.npy" for i in range(4)]
    res = wb.run(sysg, grid(), mk(), adpt_num_iter=0, restart=True, restart_iteration=1, **kw)   # the state after iteration 1
    for q in ("ahc", "dos"):
        assert relerr(res.results[q].data, g[f"iter1_{q}"]) < RTOL, q
    res = wb.run(sysg, grid(), mk(), adpt_num_iter=0, allow_restart=True, file_Klist_path=kl)   # plain run that can be continued
    for q in ("ahc", "dos"):
        assert relerr(res.results[q].data, g[f"iter0_{q}"]) < RTOL, q


def test_adaptive_refinement_dump_results(wb, tmp_path):
    """`dump_results=True` (run_grid.py:62-63, 242-243; grid/Kpoint.py:56-67): every K-point's result goes to its own
    `_Kp-<ik>.pickle` and leaves the memory, the refinement re-reads the divided points' results from there; the run
    and its restart reproduce the reference's refinement fixture."""
    import pickle
    g = np.load(os.path.join(GOLDEN, "golden_synth_adpt.npz"))
    sysg = wb.synthetic_system(6, rmax=1, seed=4242)
    st = wb.calculators.static
    mk = lambda: dict(ahc=st.AHC(Efermi=g["Efermi"]), dos=st.DOS(Efermi=g["Efermi"]))
    grid = lambda: wb.Grid(sysg, NKdiv=[2, 2, 2], NKFFT=[3, 3, 3])
    kl = str(tmp_path / "klist")
    kw = dict(adpt_fac=2, adpt_mesh=2, file_Klist_path=kl, dump_results=True, parameters_K=dict(fftlib="numpy"),
              data_k_class=wb.Data_K_R)
    res = wb.run(sysg, grid(), mk(), adpt_num_iter=1, **kw)
    for q in ("ahc", "dos"):
        assert relerr(res.results[q].data, g[f"iter1_{q}"]) < RTOL, q
    files = sorted(os.listdir(kl))
    nK = sum(f.startswith("_Kp-") for f in files)
    assert nK > 8 and "K_list.pickle" in files and "factors_iter-00000001.npy" in files
    with open(os.path.join(kl, "_Kp-0.pickle"), "rb") as f:
        r0 = pickle.load(f)
    assert set(r0.results) == {"ahc", "dos"} and r0.results["ahc"].data.shape == (len(g["Efermi"]), 3)
    with open(os.path.join(kl, "K_list.pickle"), "rb") as f:
        part = pickle.load(f)
    assert all(K.result is None and K.res_dumped_flag and K.was_evaluated_flag for K in part)
    res = wb.run(sysg, grid(), mk(), adpt_num_iter=2, restart=True, **kw)
    for q in ("ahc", "dos"):
        assert relerr(res.results[q].data, g[f"iter3_{q}"]) < RTOL, q


def test_window_arguments_and_hole_like(wb):
    """Emin / Emax / hole_like of StaticCalculator against the reference's own run (fixture of
    tests/golden/make_golden_window.py): without the tetrahedron method Emin / Emax do not act (data_K.py:172-185),
    hole_like flips the sign of Fermi-sea quantities (static.py:51-52); with tetra=True Emax acts on the inverse Fermi
    sea only and neither edge on a Fermi-surface quantity."""
    g = np.load(os.path.join(GOLDEN, "golden_synth_window.npz"))
    sysg = wb.synthetic_system(6, rmax=1, seed=4242)
    st = wb.calculators.static
    Ef, win = g["Efermi"], dict(Emin=float(g["Emin"]), Emax=float(g["Emax"]))
    calcs = dict(ahc_win_hole=st.AHC(Efermi=Ef, hole_like=True, **win),
                 cumdos_win_hole=st.CumDOS(Efermi=Ef, hole_like=True, **win),
                 dos_win_hole=st.DOS(Efermi=Ef, hole_like=True, **win),
                 ahc_tetra_Emax=st.AHC(Efermi=Ef, tetra=True, Emax=win["Emax"]),
                 dos_tetra_win=st.DOS(Efermi=Ef, tetra=True, **win))
    res = wb.run(sysg, wb.Grid(sysg, NKdiv=[2, 2, 2], NKFFT=[3, 3, 3]), calcs)
    for q in calcs:
        tol = 1e-6 if q == "dos_tetra_win" else RTOL   # der = 1 tetrahedron weights (see TETRA_CASES)
        assert relerr(res.results[q].data, g[q]) < tol, q


def test_adaptive_refinement_tetra_and_kubo(wb):
    """run(adpt_num_iter = 2) driven by tetrahedron-method and Kubo calculators (evaluated one K-point per call, the
    tetrahedron cell of a refined K-point is its own dK / NKFFT, grid/Kpoint.py:107-109) next to a plain static one,
    against the reference's own run() (fixture of tests/golden/make_golden_adpt.py)."""
    g = np.load(os.path.join(GOLDEN, "golden_synth_adpt_tetra_kubo.npz"))
    sysg = wb.synthetic_system(6, rmax=1, seed=4242)
    st, dyn = wb.calculators.static, wb.calculators.dynamic
    Ef = g["Efermi"]
    for n_iter in (0, 2):
        calcs = dict(ahc_tetra=st.AHC(Efermi=Ef, tetra=True), dos_tetra=st.DOS(Efermi=Ef, tetra=True), cumdos=st.CumDOS(Efermi=Ef),
                     optcond=dyn.OpticalConductivity(Efermi=Ef[::4], omega=g["omega"], smr_fixed_width=0.2))
        grid = wb.Grid(sysg, NKdiv=[2, 2, 2], NKFFT=[3, 3, 3])
        res = wb.run(sysg, grid, calcs, adpt_num_iter=n_iter, adpt_fac=2, adpt_mesh=2) if n_iter else wb.run(sysg, grid, calcs)
        for q in calcs:
            tol = 1e-6 if q == "dos_tetra" else RTOL   # der = 1 tetrahedron weights (see TETRA_CASES)
            assert relerr(res.results[q].data, g[f"iter{n_iter}_{q}"]) < tol, (n_iter, q)


def test_run_fe_vs_upstream_golden(wb, fe):
    """run() on the reference's own test grid against the data of the reference's golden files
    tests/reference/integrate_files/Fe_W90-{ahc,dos,cumdos}_iter-0000.npz."""
    g = np.load(os.path.join(GOLDEN, "golden_fe_nk4.npz"))
    Ef = g["Efermi"]
    st = wb.calculators.static
    calcs = dict(ahc=st.AHC(Efermi=Ef), dos=st.DOS(Efermi=Ef), cumdos=st.CumDOS(Efermi=Ef),
                 Morb=st.Morb(Efermi=Ef), spin=st.Spin(Efermi=Ef))
    res = wb.run(fe, wb.Grid(fe, NK=[4, 4, 4], NKFFT=[2, 2, 2]), calcs)
    for q in calcs:
        assert relerr(res.results[q].data, g["upstream_golden_" + q]) < RTOL, q


def test_run_fe_other_quantities(wb, fe):
    """Quantities without an upstream golden file: against a live run of the reference (fixture)."""
    g = np.load(os.path.join(GOLDEN, "golden_fe_nk4.npz"))
    Ef = g["Efermi"]
    st = wb.calculators.static
    calcs = dict(ahc_int=st.AHC(Efermi=Ef, kwargs_formula={"external_terms": False}),
                 berry_dipole_fsurf=st.BerryDipole_FermiSurf(Efermi=Ef),
                 gme_orb_fsurf=st.GME_orb_FermiSurf(Efermi=Ef),
                 gme_spin_fsurf=st.GME_spin_FermiSurf(Efermi=Ef))
    res = wb.run(fe, wb.Grid(fe, NK=[4, 4, 4], NKFFT=[2, 2, 2]), calcs)
    for q in calcs:
        assert relerr(res.results[q].data, g[q]) < RTOL, q


def test_run_fe_wide_window(wb, fe):
    """BASELINE config-1 style scan (1001 Fermi levels over 12..22 eV) on an 8^3 grid."""
    g = np.load(os.path.join(GOLDEN, "golden_fe_nk8.npz"))
    Ef = g["Efermi"]
    st = wb.calculators.static
    calcs = dict(ahc=st.AHC(Efermi=Ef), dos=st.DOS(Efermi=Ef), cumdos=st.CumDOS(Efermi=Ef))
    res = wb.run(fe, wb.Grid(fe, NK=[8, 8, 8], NKFFT=[4, 4, 4]), calcs)
    for q in calcs:
        assert relerr(res.results[q].data, g[q]) < RTOL, q


def test_run_te(wb, te):
    g = np.load(os.path.join(GOLDEN, "golden_te_nk4.npz"))
    Ef = g["Efermi"]
    st = wb.calculators.static
    calcs = dict(dos=st.DOS(Efermi=Ef), cumdos=st.CumDOS(Efermi=Ef),
                 berry_dipole_fsurf=st.BerryDipole_FermiSurf(Efermi=Ef),
                 gme_orb_fsurf=st.GME_orb_FermiSurf(Efermi=Ef),
                 gme_spin_fsurf=st.GME_spin_FermiSurf(Efermi=Ef))
    res = wb.run(te, wb.Grid(te, NK=[4, 4, 6], NKFFT=[2, 2, 3]), calcs)
    for q in calcs:
        assert relerr(res.results[q].data, g[q]) < RTOL, q


def test_grid_split_independence(wb, fe):
    """Size-independent property: the result depends only on NKdiv x NKFFT (docs/source/benchmark.rst:26-30),
    and a shard-wise evaluation adds up to the whole (linearity in the K-block weights)."""
    Ef = np.linspace(12., 22., 201)
    st = wb.calculators.static
    calcs = dict(ahc=st.AHC(Efermi=Ef), cumdos=st.CumDOS(Efermi=Ef))
    ref = wb.run(fe, wb.Grid(fe, NKdiv=[1, 1, 1], NKFFT=[12, 12, 12]), calcs)
    for div, fft in (([2, 2, 2], [6, 6, 6]), ([4, 4, 4], [3, 3, 3]), ([3, 2, 6], [4, 6, 2])):
        res = wb.run(fe, wb.Grid(fe, NKdiv=div, NKFFT=fft), calcs)
        for q in calcs:
            assert relerr(res.results[q].data, ref.results[q].data) < 1e-9, (q, div)
    # sum rule: far above all bands CumDOS counts every band
    top = st.CumDOS(Efermi=np.array([100., 101.]))
    r = wb.run(fe, wb.Grid(fe, NKdiv=[2, 2, 2], NKFFT=[3, 3, 3]), dict(c=top))
    assert np.allclose(r.results["c"].data, fe.num_wann, rtol=0, atol=1e-12)


def test_shards_add_up(wb, fe):
    from wannierberri_b200.run import shard_bounds
    Ef = np.linspace(15., 19., 41)
    st = wb.calculators.static
    grid = wb.Grid(fe, NKdiv=[3, 3, 3], NKFFT=[3, 3, 3])
    specs = st.AHC(Efermi=Ef).specs() + st.DOS(Efermi=Ef).specs()
    eng = wb.Engine(fe)
    eng.plan(grid.FFT, [s.formula for s in specs], max_kpoints_per_launch=27 * 4)  # forces several launches
    shifts, factors = grid.K_arrays()
    whole = eng.scan(shifts, factors, specs)
    parts = None
    for r in range(4):
        lo, hi = shard_bounds(len(factors), r, 4)
        p = eng.scan(shifts[lo:hi], factors[lo:hi], specs)
        parts = p if parts is None else [a + b for a, b in zip(parts, p)]
    for a, b in zip(whole, parts):
        assert relerr(a, b) < 1e-11


def test_errors(wb, fe):
    from wannierberri_b200 import _lib
    st = wb.calculators.static
    eng = wb.Engine(fe)
    with pytest.raises(ValueError):
        eng.scan(np.zeros((1, 3)), np.ones(1), st.DOS(Efermi=np.linspace(0, 1, 3)).specs())  # not planned
    eng.plan([2, 2, 2], [_lib.IDENTITY])
    with pytest.raises(ValueError):
        eng.scan(np.zeros((1, 3)), np.ones(1), st.AHC(Efermi=np.linspace(0, 1, 3)).specs())  # not in the plan
    with pytest.raises(NotImplementedError):   # hole_like tetrahedron weights exist for Fermi-sea quantities only
        st.DOS(Efermi=np.linspace(0, 1, 3), tetra=True, hole_like=True)
    with pytest.raises(ValueError):
        wb.calculators.dynamic.OpticalConductivity(Efermi=np.linspace(0, 1, 3), omega=np.linspace(0, 1, 3), kBT=-0.01)
    with pytest.raises(NotImplementedError):   # the Fermi axis of a Kubo scan must be ascending
        wb.calculators.dynamic.OpticalConductivity(Efermi=np.array([1., 0.]), omega=np.linspace(0, 1, 3))
    with pytest.raises(ValueError):
        bare = wb.System_R(fe.real_lattice, fe.rvec.iRvec, fe.wannier_centers_cart)
        bare.get_R_mat("Ham")
