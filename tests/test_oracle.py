"""The oracle (oracle/wb_oracle.py) against fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle import wb_oracle as orc

from conftest import GOLDEN

RTOL = 1e-8  # the reference's own regression tolerance (tests/common_comparers.py:186-189)


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def fe():
    return orc.OracleSystem.from_npz(os.path.join(GOLDEN, "fe_system.npz"))


@pytest.fixture(scope="module")
def te():
    return orc.OracleSystem.from_npz(os.path.join(GOLDEN, "te_system.npz"))


def test_cRvec_shifted(fe):
    f = np.load(os.path.join(GOLDEN, "fe_system.npz"))
    # floating point (goes through reduced coordinates in the reference): not index work
    assert np.abs(fe.cRvec_shifted - f["cRvec_shifted"]).max() < 1e-13
    assert fe.cell_volume == pytest.approx(float(f["cell_volume"]), rel=1e-15)


def test_kgrid_bit_exact():
    g = np.load(os.path.join(GOLDEN, "golden_fe_nk4.npz"))
    shifts, factors = orc.K_list([2, 2, 2], [2, 2, 2])
    assert np.array_equal(shifts, g["K_list_Kp_fullBZ"])
    assert np.array_equal(factors, g["K_list_factor"])
    assert np.array_equal(orc.points_FFT([2, 2, 2]), g["points_FFT"])
    b = np.load(os.path.join(GOLDEN, "golden_fe_block.npz"))
    assert np.array_equal(orc.kpoints_all(b["NKFFT"], b["dK"]), b["kpoints_all"])


def test_block_stages(fe):
    b = np.load(os.path.join(GOLDEN, "golden_fe_block.npz"))
    data = orc.OracleDataK(fe, b["dK"], b["NKFFT"])
    assert relerr(data.HH_K, b["HH_K"]) < 1e-13
    assert relerr(data.E_K, b["E_K"]) < 1e-13
    assert relerr(data.delE_K, b["delE_K"]) < 1e-11
    assert relerr(np.abs(data.Xbar("Ham", 1)), b["absV"]) < 1e-9
    assert relerr(np.abs(data.Xbar("AA")), b["absA"]) < 1e-9
    # FFT path == explicit DFT (as tests/test_data_k.py does for the reference)
    slow = orc.r_to_k_slow(fe.XX_R["AA"], fe.iRvec, b["NKFFT"], b["dK"], True)
    fast = orc.r_to_k(fe.XX_R["AA"], fe.iRvec, b["NKFFT"], b["dK"], True)
    assert relerr(fast, slow) < 1e-13


BLOCK_CASES = dict(
    ahc=("AHC", {}), dos=("DOS", {}), cumdos=("CumDOS", {}), Morb=("Morb", {}),
    ahc_kramers=("AHC", dict(degen_Kramers=True)), ahc_thresh=("AHC", dict(degen_thresh=0.05)),
    morb_thresh=("Morb", dict(degen_thresh=0.05)),
    bcd_thresh=("BerryDipole_FermiSurf", dict(degen_thresh=0.05)),
    gme_orb_thresh=("GME_orb_FermiSurf", dict(degen_thresh=0.05)),
    gme_spin_thresh=("GME_spin_FermiSurf", dict(degen_thresh=0.05)),
)


@pytest.mark.parametrize("case", sorted(BLOCK_CASES))
def test_block_calculators(fe, case):
    b = np.load(os.path.join(GOLDEN, "golden_fe_block.npz"))
    data = orc.OracleDataK(fe, b["dK"], b["NKFFT"])
    name, kw = BLOCK_CASES[case]
    got = orc.CALCULATORS[name](data, b["Efermi"], **kw)
    assert got.shape == b["block_" + case].shape
    assert relerr(got, b["block_" + case]) < RTOL


def test_fe_run_vs_upstream_golden(fe):
    """Full run on the reference's own test grid; compared with the data of the reference's
    golden files Fe_W90-*_iter-0000.npz (copied into the fixture by make_golden.py)."""
    g = np.load(os.path.join(GOLDEN, "golden_fe_nk4.npz"))
    Ef = g["Efermi"]
    calcs = dict(ahc=("AHC", Ef, {}), dos=("DOS", Ef, {}), cumdos=("CumDOS", Ef, {}), Morb=("Morb", Ef, {}),
                 spin=("Spin", Ef, {}))
    res = orc.run(fe, [2, 2, 2], [2, 2, 2], calcs)
    for q in calcs:
        assert relerr(res[q], g["upstream_golden_" + q]) < RTOL, q
        assert relerr(res[q], g[q]) < RTOL, q


def test_fe_run_other_quantities(fe):
    g = np.load(os.path.join(GOLDEN, "golden_fe_nk4.npz"))
    Ef = g["Efermi"]
    calcs = dict(ahc_int=("AHC", Ef, dict(kwargs_formula=dict(external_terms=False))),
                 berry_dipole_fsurf=("BerryDipole_FermiSurf", Ef, {}),
                 gme_orb_fsurf=("GME_orb_FermiSurf", Ef, {}),
                 gme_spin_fsurf=("GME_spin_FermiSurf", Ef, {}))
    res = orc.run(fe, [2, 2, 2], [2, 2, 2], calcs)
    for q in calcs:
        assert relerr(res[q], g[q]) < RTOL, q


def test_te_run(te):
    g = np.load(os.path.join(GOLDEN, "golden_te_nk4.npz"))
    Ef = g["Efermi"]
    calcs = dict(berry_dipole_fsurf=("BerryDipole_FermiSurf", Ef, {}),
                 gme_orb_fsurf=("GME_orb_FermiSurf", Ef, {}),
                 gme_spin_fsurf=("GME_spin_FermiSurf", Ef, {}),
                 dos=("DOS", Ef, {}), cumdos=("CumDOS", Ef, {}))
    # (AHC and Morb vanish by time-reversal symmetry in Te: the fixture holds only rounding noise)
    res = orc.run(te, [2, 2, 2], [2, 2, 3], calcs)
    for q in calcs:
        assert relerr(res[q], g[q]) < RTOL, q


# ---------------------------------------------------------------------------------------- Kubo path
KUBO_CASES = dict(
    ref_optcond=("OpticalConductivity", dict(smr_fixed_width=0.20, smr_type="Gaussian"), True),
    lor_optcond=("OpticalConductivity", dict(smr_fixed_width=0.1, smr_type="Lorentzian"), False),
    lor_optcond_thresh=("OpticalConductivity", dict(smr_fixed_width=0.1, smr_type="Lorentzian", degen_thresh=0.05), False),
    lor_optcond_int=("OpticalConductivity", dict(smr_fixed_width=0.1, smr_type="Lorentzian", external_terms=False), False),
    gau_optcond=("OpticalConductivity", dict(smr_fixed_width=0.15, smr_type="Gaussian"), False),
    lor_jdos=("JDOS", dict(smr_fixed_width=0.1, smr_type="Lorentzian"), False),
    gau_jdos=("JDOS", dict(smr_fixed_width=0.15, smr_type="Gaussian"), False),
)


def kubo_axes(g, ref_axes):
    return (g["ref_Efermi"], g["ref_omega"]) if ref_axes else (g["Efermi"], g["omega"])


@pytest.mark.parametrize("case", sorted(KUBO_CASES))
def test_kubo_block_vs_reference(fe, case):
    """DynamicCalculator.__call__ for one K-block (fixture written by make_golden_kubo.py)."""
    g = np.load(os.path.join(GOLDEN, "golden_fe_kubo.npz"))
    name, kw, ref_axes = KUBO_CASES[case]
    Ef, om = kubo_axes(g, ref_axes)
    data = orc.OracleDataK(fe, g["block_dK"], g["block_NKFFT"])
    got = orc.CALCULATORS[name](data, Ef, omega=om, **kw)
    assert got.shape == g["block_" + case].shape
    assert relerr(got, g["block_" + case]) < RTOL


def test_kubo_run_vs_upstream_golden(fe):
    """The reference's own regression file Fe_W90-opt_conductivity_iter-0000.npz (tests/test_run.py:290-305)."""
    g = np.load(os.path.join(GOLDEN, "golden_fe_kubo.npz"))
    name, kw, _ = KUBO_CASES["ref_optcond"]
    res = orc.run(fe, [2, 2, 2], [2, 2, 2], dict(oc=(name, g["ref_Efermi"], dict(omega=g["ref_omega"], **kw))))
    assert relerr(res["oc"], g["upstream_golden_opt_conductivity"]) < RTOL
    assert relerr(res["oc"], g["run_ref_optcond"]) < RTOL


# ---------------------------------------------------------------------------------------- tetrahedron method
TETRA_CASES = dict(
    ahc=("AHC", {}), dos=("DOS", {}), cumdos=("CumDOS", {}), Morb=("Morb", {}), spin=("Spin", {}),
    ahc_thresh=("AHC", dict(degen_thresh=0.05)), bcd=("BerryDipole_FermiSurf", {}), gme_spin=("GME_spin_FermiSurf", {}),
)


def test_tetra_corner_energies(fe):
    g = np.load(os.path.join(GOLDEN, "golden_fe_tetra.npz"))
    data = orc.OracleDataK(fe, g["block_dK"], g["block_NKFFT"], NKdiv=g["block_NKdiv"])
    assert np.array_equal(data.dK_cell, g["block_dK_cell"])
    assert relerr(data.E_K_corners_parallel(), g["block_E_corners"]) < 1e-12


@pytest.mark.parametrize("case", sorted(TETRA_CASES))
def test_tetra_block_vs_reference(fe, case):
    """StaticCalculator(tetra=True).__call__ for one K-block (fixture written by make_golden_tetra.py)."""
    g = np.load(os.path.join(GOLDEN, "golden_fe_tetra.npz"))
    name, kw = TETRA_CASES[case]
    data = orc.OracleDataK(fe, g["block_dK"], g["block_NKFFT"], NKdiv=g["block_NKdiv"])
    got = orc.CALCULATORS[name](data, g["Efermi"], tetra=True, **kw)
    assert got.shape == g["block_" + case].shape
    assert relerr(got, g["block_" + case]) < RTOL


def test_tetra_derivative_conditioning(fe):
    """The reference's der >= 1 tetrahedron weights are a cubic in E_F with heavily cancelling coefficients
    (grid/tetrahedron.py:53-78): a relative perturbation of 1e-15 of the corner energies -- below what any two
    eigensolvers agree to -- changes a Fermi-surface integral by ~1e-8, a Fermi-sea integral (closed form) by ~1e-15.
    This is why the GPU parity tests use 1e-6 for the der = 1 tetrahedron cases and 1e-8 for everything else."""
    g = np.load(os.path.join(GOLDEN, "golden_fe_tetra.npz"))

    def run(pert):
        data = orc.OracleDataK(fe, g["block_dK"], g["block_NKFFT"], NKdiv=g["block_NKdiv"])
        E0 = data.E_K_corners_parallel()
        Ep = E0 * (1 + pert * np.random.default_rng(1).standard_normal(E0.shape))
        data.E_K_corners_parallel = lambda: Ep
        return (orc.CALCULATORS["GME_spin_FermiSurf"](data, g["Efermi"], tetra=True),
                orc.CALCULATORS["AHC"](data, g["Efermi"], tetra=True))
    (s0, a0), (s1, a1) = run(0.), run(1e-15)
    assert relerr(a1, a0) < 1e-12
    assert 1e-10 < relerr(s1, s0) < 1e-6


# ---------------------------------------------------------------------------------------- adaptive refinement
def synth_adpt_system():
    from wannierberri_b200.system import synthetic_system  # array generator only (no GPU involved)
    g = synthetic_system(6, rmax=1, seed=4242)
    return g, orc.OracleSystem(g.rvec.iRvec, g.real_lattice, g.wannier_centers_cart,
                               {k: g.get_R_mat(k) for k in ("Ham", "AA")})


def test_adaptive_refinement_vs_reference():
    """The refinement loop (run_grid.py:303-387) against the reference's run(adpt_num_iter = 0, 1, 3) on a model
    without symmetry (fixture written by tests/golden/make_golden_adpt.py)."""
    g = np.load(os.path.join(GOLDEN, "golden_synth_adpt.npz"))
    _, syso = synth_adpt_system()
    calcs = dict(ahc=("AHC", g["Efermi"], {}), dos=("DOS", g["Efermi"], {}))
    hist = orc.run_adaptive(syso, [2, 2, 2], [3, 3, 3], calcs, adpt_num_iter=3, adpt_mesh=2, adpt_fac=2)
    for n_iter in (0, 1, 3):
        for q in calcs:
            assert relerr(hist[n_iter][q], g[f"iter{n_iter}_{q}"]) < RTOL, (n_iter, q)


def test_ohmic_fsurf_vs_upstream_golden(fe):
    """Ohmic_FermiSurf (VelVel, calculators/static.py:393-403) on the reference's test grid against the reference's
    own golden file Fe_W90-conductivity_ohmic_fsurf_iter-0000.npz."""
    g = np.load(os.path.join(GOLDEN, "golden_fe_ohmic.npz"))
    res = orc.run(fe, [2, 2, 2], [2, 2, 2], dict(a=("Ohmic_FermiSurf", g["Efermi"], {}),
                                                  b=("Ohmic_FermiSurf", g["Efermi"], dict(degen_thresh=0.05))))
    assert relerr(res["a"], g["upstream_golden_ohmic_fsurf"]) < RTOL
    assert relerr(res["b"], g["ohmic_fsurf_thresh"]) < RTOL


@pytest.mark.parametrize("fder", [0, 1, 2, 3])
def test_fder_stencils_vs_upstream_files(fe, fder):
    """StaticCalculator(Formula=Identity, fder=0..3) (static.py:137-147) against the reference's own files
    tests/reference/calculators/calculator-Fe-ident-fder=*.npz (tests/test_calc.py:71-77)."""
    g = np.load(os.path.join(GOLDEN, "golden_fe_calc_fder.npz"))
    data = orc.OracleDataK(fe, g["dK"], g["NKFFT"])
    got = orc.static_scan(data, orc.Identity, fder, g["Efermi"])
    ref = g[f"upstream_ident_fder{fder}"]
    assert np.abs(got - ref).max() <= RTOL * max(np.abs(ref).max(), 1e-300)


RANDOM_CASES = dict(
    ahc=("AHC", {}), dos=("DOS", {}), cumdos=("CumDOS", {}), Morb=("Morb", {}), spin=("Spin", {}),
    conductivity_ohmic_fsurf=("Ohmic_FermiSurf", {}), berry_dipole_fsurf=("BerryDipole_FermiSurf", {}),
    gme_orb_fsurf=("GME_orb_FermiSurf", {}), gme_spin_fsurf=("GME_spin_FermiSurf", {}),
)


def test_random_system_vs_upstream_goldens():
    """The reference's `random` test system (6 WF, 20 R-vectors without the R <-> -R symmetry: H(k) is hermitised by the
    transform, d_aH is NOT hermitian) against the reference's own golden files random-*_iter-0000.npz
    (tests/test_run.py:653-669) and a live reference run (tests/golden/make_golden_random.py)."""
    g = np.load(os.path.join(GOLDEN, "golden_random.npz"))
    rnd = orc.OracleSystem.from_npz(os.path.join(GOLDEN, "random_system.npz"))
    calcs = {k: (name, g["Efermi"], kw) for k, (name, kw) in RANDOM_CASES.items()}
    calcs["opt_conductivity"] = ("OpticalConductivity", g["opt_Efermi"], dict(omega=g["opt_omega"], smr_fixed_width=0.20, smr_type="Gaussian"))
    calcs["opt_conductivity_in"] = ("OpticalConductivity", g["opt_in_Efermi"], dict(omega=g["opt_omega"], smr_fixed_width=0.20))
    res = orc.run(rnd, [2, 2, 2], [3, 3, 3], calcs)
    for q in calcs:
        scale = max(np.abs(g[q]).max(), 1e-300)
        assert np.abs(res[q] - g[q]).max() <= RTOL * scale, q
        if "upstream_golden_" + q in g.files:
            assert np.abs(res[q] - g["upstream_golden_" + q]).max() <= RTOL * max(np.abs(g["upstream_golden_" + q]).max(), 1e-300), q


@pytest.mark.parametrize("tag,div,fft", [("fe", [2, 2, 2], [2, 2, 2]), ("random", [2, 2, 2], [3, 3, 3])])
def test_ohmic_fermi_sea_vs_upstream_golden(tag, div, fft):
    """Ohmic_FermiSea (InvMass: Xbar('Ham', 2) + generalised derivative, elementary.py:28-34, formula.py:95-112)
    against the reference's own golden files {Fe_W90,random}-conductivity_ohmic_iter-0000.npz."""
    g = np.load(os.path.join(GOLDEN, "golden_ohmic_sea.npz"))
    system = orc.OracleSystem.from_npz(os.path.join(GOLDEN, "fe_system.npz" if tag == "fe" else "random_system.npz"))
    Ef = g[f"{tag}_Efermi"]
    res = orc.run(system, div, fft, dict(a=("Ohmic_FermiSea", Ef, {}), b=("Ohmic_FermiSea", Ef, dict(degen_thresh=0.05))))
    assert relerr(res["a"], g[f"{tag}_upstream_golden_ohmic"]) < RTOL
    assert relerr(res["b"], g[f"{tag}_ohmic_thresh"]) < RTOL


SHC_CASES = dict(ref_qiao=("qiao", "ref", {}), ref_ryoo=("ryoo", "ref", {}), in_qiao=("qiao", "in", {}), in_ryoo=("ryoo", "in", {}),
                 in_simple=("simple", "in", {}), in_ryoo_thresh=("ryoo", "in", dict(degen_thresh=0.3)))


@pytest.mark.parametrize("case", sorted(SHC_CASES))
def test_shc_random_system(case):
    """Kubo spin Hall conductivity (calculators/dynamic.py:204-237, formula/covariant.py:689-756) on the reference's
    `random` system: the three spin-current types against the live reference run of make_golden_shc.py and, for
    qiao / ryoo, the reference's own golden files random-opt_SHC{qiao,ryoo}_iter-0000.npz."""
    g = np.load(os.path.join(GOLDEN, "golden_random_shc.npz"))
    rnd = orc.OracleSystem.from_npz(os.path.join(GOLDEN, "random_system.npz"))
    t, ax, extra = SHC_CASES[case]
    kw = dict(omega=g["omega"], smr_fixed_width=0.20, smr_type="Gaussian" if ax == "ref" else "Lorentzian", SHC_type=t, **extra)
    res = orc.run(rnd, [2, 2, 2], [3, 3, 3], dict(a=("SHC", g[ax + "_Efermi"], kw)))
    assert relerr(res["a"], g[case]) < RTOL
    if ax == "ref":
        assert relerr(res["a"], g["upstream_golden_" + t]) < RTOL


FSEA_CASES = dict(
    bd_sea=("BerryDipole_FermiSea", {}), bd_sea_int=("BerryDipole_FermiSea", dict(kwargs_formula=dict(external_terms=False))),
    bd_sea_thresh=("BerryDipole_FermiSea", dict(degen_thresh=0.3)), bd_sea_tetra=("BerryDipole_FermiSea", dict(tetra=True)),
    nlahc_sea=("NLAHC_FermiSea", {}),
    shc_ryoo=("SHC_static", dict(kwargs_formula=dict(spin_current_type="ryoo"))),
    shc_qiao=("SHC_static", dict(kwargs_formula=dict(spin_current_type="qiao"))),
    shc_simple=("SHC_static", dict(kwargs_formula=dict(spin_current_type="simple"))),
    shc_simple_int=("SHC_static", dict(kwargs_formula=dict(spin_current_type="simple", external_terms=False))),
    shc_ryoo_thresh=("SHC_static", dict(degen_thresh=0.3, kwargs_formula=dict(spin_current_type="ryoo"))),
    shc_qiao_tetra=("SHC_static", dict(tetra=True, kwargs_formula=dict(spin_current_type="qiao"))),
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


def test_fermi_sea_formulae_random_system():
    """DerOmega (BerryDipole_FermiSea / NLAHC_FermiSea; formula/covariant.py:212-259) and SpinOmega (static.SHC; :759-789)
    on the reference's `random` system against the live reference run of make_golden_fsea.py (which also reproduces the
    upstream Te_QE-{BerryDipole_FermiSea,berry_dipole} golden files, checked on the GPU path with symmetry + tetra)."""
    g = np.load(os.path.join(GOLDEN, "golden_fsea.npz"))
    rnd = orc.OracleSystem.from_npz(os.path.join(GOLDEN, "random_system.npz"))
    res = orc.run(rnd, [2, 2, 2], [3, 3, 3], {k: (name, g["rnd_Efermi"], kw) for k, (name, kw) in FSEA_CASES.items()})
    for k in FSEA_CASES:
        assert res[k].shape == g["rnd_" + k].shape, k
        assert relerr(res[k], g["rnd_" + k]) < RTOL, k


SHIFT_CASES = dict(
    shift=("ShiftCurrent", dict(sc_eta=0.1, smr_type="Lorentzian")), shift_gauss=("ShiftCurrent", dict(sc_eta=0.04, smr_type="Gaussian")),
    shift_int=("ShiftCurrent", dict(sc_eta=0.1, smr_type="Lorentzian", external_terms=False)),
    shift_thresh=("ShiftCurrent", dict(sc_eta=0.1, smr_type="Lorentzian", degen_thresh=0.3)),
    injection=("InjectionCurrent", dict(smr_type="Lorentzian")), injection_gauss=("InjectionCurrent", dict(smr_type="Gaussian")),
    injection_int=("InjectionCurrent", dict(smr_type="Lorentzian", external_terms=False)),
    injection_thresh=("InjectionCurrent", dict(smr_type="Lorentzian", degen_thresh=0.3)),
)


def test_shift_and_injection_current_random_system():
    """Kubo shift current and injection current (calculators/dynamic.py:244-365) on the reference's `random` system against
    the live reference run of tests/golden/make_golden_shift.py."""
    g = np.load(os.path.join(GOLDEN, "golden_random_shift.npz"))
    rnd = orc.OracleSystem.from_npz(os.path.join(GOLDEN, "random_system.npz"))
    calcs = {k: (name, g["Efermi"], dict(omega=g["omega"], smr_fixed_width=0.20, **kw)) for k, (name, kw) in SHIFT_CASES.items()}
    res = orc.run(rnd, [2, 2, 2], [3, 3, 3], calcs)
    for k in SHIFT_CASES:
        assert res[k].shape == g[k].shape and res[k].dtype == g[k].dtype, k
        assert relerr(res[k], g[k]) < RTOL, k


KBT_CASES = dict(
    optcond=("OpticalConductivity", dict(kBT=0.05)), optcond_hot=("OpticalConductivity", dict(kBT=0.5, smr_type="Gaussian")),
    jdos=("JDOS", dict(kBT=0.05)), shc_ryoo=("SHC", dict(SHC_type="ryoo", kBT=0.03)),
    shift=("ShiftCurrent", dict(sc_eta=0.1, kBT=0.05)), injection=("InjectionCurrent", dict(kBT=0.01)),
    optcond_thresh=("OpticalConductivity", dict(kBT=0.05, degen_thresh=0.3)),
)


def test_kubo_finite_temperature_random_system():
    """Kubo calculators with the Fermi-Dirac factor at kBT > 0 (utility.py:172-182, dynamic.py:44-45,61-69) on the reference's
    `random` system against the live reference run of tests/golden/make_golden_kbt.py."""
    g = np.load(os.path.join(GOLDEN, "golden_random_kbt.npz"))
    rnd = orc.OracleSystem.from_npz(os.path.join(GOLDEN, "random_system.npz"))
    calcs = {k: (name, g["Efermi"], dict(omega=g["omega"], smr_fixed_width=0.20, **kw)) for k, (name, kw) in KBT_CASES.items()}
    res = orc.run(rnd, [2, 2, 2], [3, 3, 3], calcs)
    for k in KBT_CASES:
        assert res[k].shape == g[k].shape, k
        assert relerr(res[k], g[k]) < RTOL, k


def test_soc_system_flattening_vs_reference_data_k_soc():
    """Data_K_soc (data_K/data_K_soc.py:7-62): a SOC system -- scalar up / down systems on the even / odd Wannier
    functions plus a spin-orbit term, each on its own R-vector set -- is flattened by `wannierberri_b200.system.flatten_soc`
    into ONE System_R; the oracle on that system must reproduce what the unmodified reference computed with its
    Data_K_soc (fixture: tests/golden/make_golden_soc.py)."""
    import wannierberri_b200 as wb
    from wannierberri_b200.system import flatten_soc, as_system
    from conftest import soc_system_from_fixture, SOC_CALCS
    g = np.load(os.path.join(GOLDEN, "golden_soc.npz"))
    soc = soc_system_from_fixture(wb, g)
    flat = flatten_soc(soc)
    assert as_system(soc) is as_system(soc) and as_system(soc).num_wann == 6
    assert flat.num_wann == 6 and flat.rvec.nRvec == len({tuple(R) for k in ("up", "dw", "soc") for R in g["iRvec_" + k]})
    assert set(flat._XX_R) == {"Ham", "AA", "SS"}
    osys = orc.OracleSystem(flat.rvec.iRvec, flat.real_lattice, flat.wannier_centers_cart, flat._XX_R)
    NKdiv = (g["NK"] // g["NKFFT"]).tolist()
    res = orc.run(osys, NKdiv, g["NKFFT"].tolist(), {k: (name, g["Efermi"], kw) for k, (name, kw) in SOC_CALCS.items()})
    for key in SOC_CALCS:
        want = g["res_" + key]
        assert np.abs(res[key] - want).max() <= RTOL * np.abs(want).max(), key


SELECT_CASES = dict(ohmic_sel=("Ohmic_FermiSurf", dict(degen_thresh=0.3)), dos_sel=("DOS", {}),
                    bcd_sel=("BerryDipole_FermiSurf", dict(degen_thresh=0.3, degen_Kramers=True)),
                    gme_sel=("GME_orb_FermiSurf", {}), ohmic_all=("Ohmic_FermiSurf", dict(degen_thresh=0.3)))


def select_of(g, key):
    return None if key == "ohmic_all" else np.array([5]) if key == "gme_sel" else g["select"]


def test_select_bands_vs_reference():
    """`select_bands` (calculators/static.py:93-100, 129-136; utility.py:398-403): band groups count with the fraction of
    their bands that is selected; fixture from the unmodified reference (tests/golden/make_golden_select.py)."""
    g = np.load(os.path.join(GOLDEN, "golden_select.npz"))
    osys = orc.OracleSystem.from_npz(os.path.join(GOLDEN, "fe_system.npz"))
    NKdiv = (g["NK"] // g["NKFFT"]).tolist()
    calcs = {k: (name, g["Efermi"], dict(kw, **({} if select_of(g, k) is None else dict(select_bands=select_of(g, k)))))
             for k, (name, kw) in SELECT_CASES.items()}
    res = orc.run(osys, NKdiv, g["NKFFT"].tolist(), calcs)
    for key in calcs:
        assert np.abs(res[key] - g[key]).max() <= RTOL * np.abs(g[key]).max(), key
    data = orc.OracleDataK(osys, [0., 0., 0.], g["NKFFT"])
    with pytest.raises(NotImplementedError):
        orc.CALCULATORS["AHC"](data, g["Efermi"], select_bands=np.array([1]))


TETRA_HOLE_CASES = dict(ahc_holes=("AHC", dict(hole_like=True)), cumdos_holes=("CumDOS", dict(hole_like=True)),
                        morb_holes=("Morb", dict(hole_like=True)), ahc_holes_emax=("AHC", dict(hole_like=True, Emax=30.)),
                        ahc_emin=("AHC", dict(Emin=12.5)), cumdos_emin=("CumDOS", dict(Emin=11.)),
                        ohmic_sea_emin=("Ohmic_FermiSea", dict(Emin=12.5, degen_thresh=0.05)), ahc_plain=("AHC", {}))


def test_tetra_hole_like_and_emin_vs_reference():
    """tetra=True with hole_like (weights 1 - occupation and the group of the bands above the Fermi axis, up to Emax) and
    with Emin on Fermi-sea quantities (grid/tetrahedron.py:197-198, 246-266; static.py:50-52, 84-91) against the fixture of
    the unmodified reference (tests/golden/make_golden_tetra_holes.py)."""
    g = np.load(os.path.join(GOLDEN, "golden_fe_tetra_holes.npz"))
    osys = orc.OracleSystem.from_npz(os.path.join(GOLDEN, "fe_system.npz"))
    NKdiv = (g["NK"] // g["NKFFT"]).tolist()
    calcs = {k: (name, g["Efermi"], dict(kw, tetra=True)) for k, (name, kw) in TETRA_HOLE_CASES.items()}
    res = orc.run(osys, NKdiv, g["NKFFT"].tolist(), calcs)
    for key in calcs:
        assert np.abs(res[key] - g[key]).max() <= RTOL * np.abs(g[key]).max(), key


def test_second_order_formulae_vs_reference(fe):
    """Der2Spin / Der2Omega (Der2Dcov, Der2A, Der2O) / Der2Morb (Der2B, Der2H) / emcha_surf / tildeFab / tildeFab_d and the calculators built on them
    (NLDrude_Zeeman_spin, NLDrude_Zeeman_orb_Omega, NLDrude_Zeeman_orb, eMChA_FermiSurf, QuantumMetric_FermiSea, QuantumMetric_Vel_DQ) against the
    fixture of the unmodified reference (tests/golden/make_golden_second_order.py): external terms on and off, wide
    Kramers-paired band groups."""
    g = np.load(os.path.join(GOLDEN, "golden_second_order.npz"))
    Ef = g["Efermi"]
    cases = dict(z_spin=("NLDrude_Zeeman_spin", {}), z_orb_omega=("NLDrude_Zeeman_orb_Omega", {}), emcha=("eMChA_FermiSurf", {}),
                 qmetric=("QuantumMetric_FermiSea", {}), qmetric_dip=("QuantumMetric_Vel_DQ", {}),
                 z_orb=("NLDrude_Zeeman_orb", {}), z_orb_int=("NLDrude_Zeeman_orb", dict(kwargs_formula=dict(external_terms=False))),
                 emcha_int=("eMChA_FermiSurf", dict(kwargs_formula=dict(external_terms=False))),
                 qmetric_int=("QuantumMetric_FermiSea", dict(kwargs_formula=dict(external_terms=False))),
                 z_spin_wide=("NLDrude_Zeeman_spin", dict(degen_thresh=0.3, degen_Kramers=True)),
                 qmetric_wide=("QuantumMetric_FermiSea", dict(degen_thresh=0.3)))
    res = orc.run(fe, list(g["NK"] // g["NKFFT"]), list(g["NKFFT"]), {k: (n, Ef, kw) for k, (n, kw) in cases.items()})
    for key in cases:
        assert res[key].shape == g[key].shape
        assert relerr(res[key], g[key]) < RTOL, key
