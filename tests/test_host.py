"""CPU-only tests: the C-ABI library loads and exports what include/wbgpu.h declares, the host logic
(grid, calculators, sharding, multi-rank reduction over gloo) behaves like the reference's."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

import wannierberri_b200 as wb
from wannierberri_b200 import _lib
from wannierberri_b200.run import shard_bounds


def test_library_exports_header_symbols():
    header = open(os.path.join(ROOT, "include", "wbgpu.h")).read()
    declared = set(re.findall(r"\b(wbgpu_[A-Za-z0-9_]+)\s*\(", header))
    declared -= {"wbgpu_ctx"}
    assert declared, "no declarations parsed"
    L = _lib.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/wbgpu.h but not exported by libwbgpu.so"
    assert set(_lib.EXPORTED) <= declared
    assert L.wbgpu_version() >= 100


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly."""
    if _lib.lib().wbgpu_device_count() > 0:
        pytest.skip("a CUDA device is present")
    fe = wb.System_R.from_npz(os.path.join(GOLDEN, "fe_system.npz"))
    with pytest.raises(RuntimeError, match="no CUDA device"):
        wb.Engine(fe)
    with pytest.raises(RuntimeError):
        wb.run(fe, wb.Grid(fe, NK=4, NKFFT=2), dict(dos=wb.calculators.static.DOS(Efermi=np.linspace(0, 1, 3))))


def test_grid_bit_exact():
    fe = wb.System_R.from_npz(os.path.join(GOLDEN, "fe_system.npz"))
    g = np.load(os.path.join(GOLDEN, "golden_fe_nk4.npz"))
    grid = wb.Grid(fe, NK=[4, 4, 4], NKFFT=[2, 2, 2])
    assert list(grid.div) == [2, 2, 2] and list(grid.FFT) == [2, 2, 2]
    shifts, factors = grid.K_arrays()
    assert np.array_equal(shifts, g["K_list_Kp_fullBZ"])
    assert np.array_equal(factors, g["K_list_factor"])
    assert np.array_equal(grid.points_FFT, g["points_FFT"])
    K = grid.get_K_list()
    assert np.array_equal(np.array([k.Kp_fullBZ for k in K]), g["K_list_Kp_fullBZ"])
    # awkward sizes: same floating point operations as the oracle's restatement of grid.py
    from oracle import wb_oracle as orc
    for div, fft in (([3, 5, 7], [4, 6, 9]), ([20, 20, 20], [20, 20, 20]), ([1, 1, 1], [12, 12, 12])):
        s, f = wb.Grid(fe, NKdiv=div, NKFFT=fft).K_arrays()
        so, fo = orc.K_list(div, fft)
        assert np.array_equal(s, so) and np.array_equal(f, fo)
        assert np.array_equal(wb.Grid(fe, NKdiv=div, NKFFT=fft).points_FFT, orc.points_FFT(fft))


def test_grid_determineNK():
    fe = wb.System_R.from_npz(os.path.join(GOLDEN, "fe_system.npz"))
    assert list(fe.NKFFT_recommended) == [7, 7, 7]
    g = wb.Grid(fe, NK=48, NKFFT=12)
    assert list(g.div) == [4, 4, 4] and list(g.dense) == [48, 48, 48]
    with pytest.warns(UserWarning):
        g = wb.Grid(fe, NK=50, NKFFT=12)
    assert list(g.dense) == [48, 48, 48]
    with pytest.raises(ValueError):
        wb.Grid(fe)
    g = wb.Grid(fe, NK=56)
    assert np.all(g.dense == 56)


def test_calculator_windows_and_errors():
    st = wb.calculators.static
    Ef = np.linspace(12., 22., 1001)
    a = st.AHC(Efermi=Ef)
    s = a.specs()[0]
    assert (s.formula, s.fder, s.nEF) == (_lib.OMEGA, 0, 1001)
    assert a.EFmin == Ef[0] and a.nEF_extra == 1001
    d = st.DOS(Efermi=Ef)
    assert d.nEF_extra == 1003 and d.EFmin == Ef[0] - (Ef[1] - Ef[0])
    m = st.Morb(Efermi=Ef)
    assert [x.formula for x in m.specs()] == [_lib.MORB_HPM, _lib.OMEGA]
    assert st.AHC(Efermi=Ef, hole_like=True).constant_factor == -a.constant_factor
    assert st.AHC(Efermi=Ef, use_factor=False).specs()[0].factor == -1.0
    for bad in (dict(k_resolved=True), dict(select_bands=[1], tetra=True)):
        with pytest.raises(NotImplementedError):
            st.AHC(Efermi=Ef, **bad)
    with pytest.raises(NotImplementedError):   # hole_like weights of the tetrahedron method: Fermi-sea quantities only
        st.DOS(Efermi=Ef, tetra=True, hole_like=True)
    t = st.AHC(Efermi=Ef, tetra=True, hole_like=True, Emax=30.).specs()[0]
    assert t.tetra_flags == 5 and t.tetra_Emax == 30. and st.AHC(Efermi=Ef, tetra=True, Emin=13.).specs()[0].tetra_flags == 2
    assert st.AHC(Efermi=Ef, Emin=13.).specs()[0].tetra_flags == 0   # not read without tetra
    with pytest.raises(NotImplementedError):   # "Selection of bands for Fermi sea is not implemented" (data_K.py:179-180)
        st.AHC(Efermi=Ef, select_bands=[1]).specs()
    sel = st.Ohmic_FermiSurf(Efermi=Ef, select_bands=(3, 70, 3)).specs()[0]
    assert sel.use_select == 1 and sel.select_mask[0] == 1 << 3 and sel.select_mask[1] == 1 << 6
    with pytest.raises(ValueError):
        st.AHC(Efermi=5.0)
    w = st.AHC(Efermi=Ef, Emin=13., Emax=20.)   # read by the tetrahedron method only (as in the reference)
    assert (w.Emin, w.Emax) == (13., 20.) and w.specs()[0].factor == a.specs()[0].factor
    st.DOS(Efermi=Ef, tetra=True, Emin=13., Emax=20.)   # fder = 1: neither edge acts


def test_shard_bounds():
    for n in (0, 1, 7, 64, 8000):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                lo, hi = shard_bounds(n, r, world)
                got += list(range(lo, hi))
            assert got == list(range(n))


def test_result_algebra():
    E = np.linspace(0, 1, 5)
    a = wb.EnergyResult(E, np.ones((5, 3)))
    b = (a * 2 + a) - a
    assert np.allclose(b.data, 2.0)
    assert (None + a if False else a.__radd__(None)) is a
    with pytest.raises(RuntimeError):
        a + wb.EnergyResult(E + 1, np.ones((5, 3)))
    d = a.as_dict()
    assert set(d) >= {"data", "Energies_0", "E_titles", "rank", "transformTR", "transformInv", "comment"}


def test_two_rank_gloo_run():
    """world_size=2 over gloo on CPU: K-block sharding + one all-reduce reproduces the fixture.
    The GPU engine is replaced by a test double that evaluates shards with the oracle."""
    script = os.path.join(ROOT, "tests", "_gloo_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29577")
    procs = [subprocess.Popen([sys.executable, script, str(r), "2"], env=env, stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
    assert "PARITY OK" in outs[0]


FE_POINTGROUP = ["C4z", "C2x*TimeReversal", "Inversion"]   # tests/common_systems.py:186 of the reference


def test_symmetry_reduced_k_list_bit_exact():
    """Grid.get_K_list(use_symmetry=True) (grid/grid.py:149-189): the irreducible K-points and their absorbed weights
    equal the reference's, in the reference's order; the group has 16 operations and averages tensors like the
    reference (point_symmetry.py:104-119, 337-338)."""
    g = np.load(os.path.join(GOLDEN, "golden_fe_sym.npz"))
    fe = wb.System_R.from_npz(os.path.join(GOLDEN, "fe_system.npz"), pointgroup=FE_POINTGROUP)
    assert fe.pointgroup.size == 16
    shifts, factors = wb.Grid(fe, NK=[4, 4, 4], NKFFT=[2, 2, 2]).K_arrays(use_symmetry=True)
    assert np.array_equal(shifts, g["K_list_Kp_fullBZ"])
    assert np.array_equal(factors, g["K_list_factor"])
    assert abs(factors.sum() - 1.) < 1e-15
    # larger grid: weights still add up, fewer points
    s2, f2 = wb.Grid(fe, NKdiv=[6, 6, 6], NKFFT=[2, 2, 2]).K_arrays(use_symmetry=True)
    assert abs(f2.sum() - 1.) < 1e-13 and len(f2) < 6 ** 3 / 4
    # symmetrisation: an axial vector along z (odd under TR, even under inversion) survives, x / y components vanish
    from wannierberri_b200.result import EnergyResult
    r = EnergyResult(np.arange(3.), np.tile(np.array([1., 2., 3.]), (3, 1)), transformTR="odd", transformInv="ident")
    assert np.allclose(r.symmetrized(fe.pointgroup).data, np.tile(np.array([0., 0., 3.]), (3, 1)), atol=1e-14)
    with pytest.raises(ValueError):
        wb.System_R.from_npz(os.path.join(GOLDEN, "fe_system.npz"), pointgroup=["C3z"])   # not a symmetry of the bcc cell


def test_adapt_reference_calculator_instances():
    """run() accepts instances of the reference's own calculator classes: every static and Kubo
    class is recognised by name and converted with the same constant factor, formula and options.  Container-only: skipped
    where the reference tree is absent (GPU box)."""
    import sys
    ref = "/root/reference"
    if not os.path.isdir(os.path.join(ref, "wannierberri")):
        pytest.skip("reference tree not present")
    stubs = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "stubs")
    sys.path[:0] = [ref, stubs]
    try:
        from wannierberri import calculators as calc
    except Exception as err:   # pragma: no cover
        pytest.skip(f"reference not importable: {err}")
    finally:
        sys.path.remove(ref)
        sys.path.remove(stubs)
    from wannierberri_b200.calculators import static as st, dynamic as dy
    Ef, om = np.linspace(0, 1, 5), np.linspace(0, 1, 4)
    S, D = calc.static, calc.dynamic
    names = ["DOS", "CumDOS", "Spin", "AHC", "Morb", "BerryDipole_FermiSurf", "NLAHC_FermiSurf", "GME_orb_FermiSurf",
             "GME_spin_FermiSurf", "Ohmic_FermiSurf", "Ohmic_FermiSea", "BerryDipole_FermiSea", "NLAHC_FermiSea", "GME_spin_FermiSea",
             "GME_orb_FermiSea", "NLDrude_FermiSurf", "NLDrude_Fermider2", "NLDrude_FermiSea", "Hall_classic_FermiSurf",
             "Hall_classic_FermiSea", "AHC_Zeeman_spin", "AHC_Zeeman_orb", "OmegaOmega"]
    for name in names:
        c = getattr(S, name)(Efermi=Ef, tetra=(name == "BerryDipole_FermiSea"), degen_thresh=0.01)
        n = st.adapt(c)
        assert type(n).__name__ == name and n.tetra == c.tetra and n.degen_thresh == 0.01 and n.fder == c.fder
        assert abs(n.constant_factor - c.constant_factor) <= 1e-12 * abs(c.constant_factor), name
    n = st.adapt(S.SHC(Efermi=Ef, kwargs_formula={"spin_current_type": "qiao"}))
    assert [int(s.formula) for s in n.specs()] == [_lib.SHC_QIAO]
    for c in (D.OpticalConductivity(Efermi=Ef, omega=om, kBT=0.01), D.JDOS(Efermi=Ef, omega=om),
              D.SHC(Efermi=Ef, omega=om, SHC_type="qiao", shc_abc=(1, 2, 3)), D.ShiftCurrent(Efermi=Ef, omega=om, sc_eta=0.04),
              D.InjectionCurrent(Efermi=Ef, omega=om)):
        n = dy.adapt(c)
        assert type(n).__name__ == type(c).__name__ and n.kBT == c.kBT
        assert abs(n.constant_factor - c.constant_factor) <= 1e-12 * abs(c.constant_factor)
    assert dy.adapt(D.ShiftCurrent(Efermi=Ef, omega=om, sc_eta=0.04)).spec().sc_eta == 0.04
    assert dy.adapt(D.SHC(Efermi=Ef, omega=om, SHC_type="qiao")).spec().shc_type == _lib.SHC_QIAO


def test_smoothers():
    """wannierberri_b200.smoother against direct sums; EnergyResult.dataSmooth / .max use the smoother of every energy axis;
    the reference's own smoother classes give the same numbers (container-only part)."""
    import sys
    from wannierberri_b200.smoother import FermiDiracSmoother, GaussianSmoother, VoidSmoother, get_smoother
    rng = np.random.default_rng(3)
    E = np.linspace(-1., 1., 81)
    A = rng.standard_normal((81, 3, 2)) + 1j * rng.standard_normal((81, 3, 2))
    sm = GaussianSmoother(E, 0.06, maxdE=5)
    got = sm(A)
    for i in (0, 7, 40, 80):   # direct evaluation incl. the renormalised windows at both ends
        j = np.arange(max(0, i - sm.NE1), min(81, i + sm.NE1 + 1))
        w = np.exp(-((E[j] - E[i]) / 0.06) ** 2)
        assert np.allclose(got[i], np.tensordot(w / w.sum(), A[j], axes=(0, 0)), rtol=1e-13, atol=1e-15)
    assert np.allclose(sm(A.transpose(1, 0, 2), axis=1), got.transpose(1, 0, 2))
    assert np.allclose(sm(np.ones(81)), 1.)   # normalised
    assert isinstance(get_smoother(E, 0, "Gaussian"), VoidSmoother) and get_smoother(E, 300, "Fermi-Dirac") == FermiDiracSmoother(E, 300)
    with pytest.raises(ValueError):
        get_smoother(E, 1., "other")
    r = wb.EnergyResult(E, A.real, smoothers=[sm])
    assert np.allclose(r.dataSmooth, sm(A.real)) and np.allclose(r.max[0], np.abs(sm(A.real)).max())
    assert np.allclose((r + r * 2.).dataSmooth, 3 * sm(A.real))
    with pytest.raises(RuntimeError):
        r + wb.EnergyResult(E, A.real, smoothers=[GaussianSmoother(E, 0.1)])
    c = wb.calculators.static.AHC(Efermi=E, smoother=sm)
    assert c.result([np.zeros((81, 3))], 1.).smoothers == [sm]
    ref = "/root/reference"
    if not os.path.isdir(os.path.join(ref, "wannierberri")):
        return
    sys.path.insert(0, ref)
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location("_ref_smoother", os.path.join(ref, "wannierberri", "smoother.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(ref)
    assert np.allclose(mod.GaussianSmoother(E, 0.06, maxdE=5)(A), got, rtol=1e-12, atol=1e-14)
    assert np.allclose(mod.FermiDiracSmoother(E, 400.)(A.real), FermiDiracSmoother(E, 400.)(A.real), rtol=1e-12, atol=1e-14)
    # a reference smoother object is accepted as is
    r2 = wb.EnergyResult(E, A.real, smoothers=[mod.FermiDiracSmoother(E, 400.)])
    assert np.allclose(r2.dataSmooth, FermiDiracSmoother(E, 400.)(A.real), rtol=1e-12, atol=1e-14)


def test_kpoint_dump_result(tmp_path):
    """`dump_results` of run(): KpointBZparallel.dump_result / get_result (grid/Kpoint.py:40-67) -- the result leaves
    the memory, the K-point stays "evaluated", `max` and the pickled K-list keep working, absorb() reads it back."""
    import pickle
    from wannierberri_b200.grid import KpointBZparallel
    from wannierberri_b200.result import EnergyResult, ResultDict
    Ef = np.linspace(0, 1, 5)
    res = ResultDict(dict(q=EnergyResult(Ef, np.arange(15.).reshape(5, 3), rank=1)))
    K = KpointBZparallel([0.5, 0, 0], [1, 1, 1], [2, 2, 2], 0.25)
    with pytest.raises(RuntimeError):
        K.get_result()
    K.set_result(res)
    mx = np.array(K.max)
    K.set_storage_path(str(tmp_path / "_Kp-0.pickle"))
    K.dump_result()
    K.dump_result()   # idempotent
    assert K.result is None and K.res_dumped_flag and K.was_evaluated_flag
    assert np.array_equal(np.array(K.max), mx)
    K2 = pickle.loads(pickle.dumps(K))
    assert np.array_equal(K2.get_result().results["q"].data, res.results["q"].data)
    K3 = KpointBZparallel([0.5, 0, 0], [1, 1, 1], [2, 2, 2], 0.25)
    K3.absorb(K2)
    assert K3.factor == 0.5 and np.array_equal(K3.result.results["q"].data, res.results["q"].data)


def test_parameters_K_and_data_k_class():
    """run(parameters_K=..., data_k_class=...): `fftlib` is validated as in the reference (fourier/fft.py:63) and has
    no effect, defaults pass, anything that would need another code path raises before any GPU work."""
    from wannierberri_b200.data_K import check_parameters_K
    check_parameters_K(None)
    check_parameters_K(dict(fftlib="numpy", Emin=-np.inf, Emax=np.inf, random_gauge=False, degen_thresh_random_gauge=1e-4))
    with pytest.raises(ValueError):
        check_parameters_K(dict(fftlib="cufft"))
    for bad in (dict(random_gauge=True), dict(Emin=0.), dict(k_list=[[0, 0, 0]])):
        with pytest.raises(NotImplementedError):
            check_parameters_K(bad)
    fe = wb.System_R.from_npz(os.path.join(GOLDEN, "fe_system.npz"))
    calcs = dict(dos=wb.calculators.static.DOS(Efermi=np.linspace(17, 18, 3)))
    grid = wb.Grid(fe, NKdiv=[1, 1, 1], NKFFT=[2, 2, 2])
    with pytest.raises(NotImplementedError):
        wb.run(fe, grid, calcs, data_k_class=object)
    with pytest.raises(ValueError):
        wb.run(fe, grid, calcs, parameters_K=dict(fftlib="cufft"))


def _import_reference():
    """the reference package from /root/reference (container only; absent on the GPU box)"""
    ref = "/root/reference"
    if not os.path.isdir(os.path.join(ref, "wannierberri")):
        pytest.skip("reference tree not present")
    stubs = os.path.join(ROOT, "oracle", "stubs")
    sys.path[:0] = [ref, stubs]
    try:
        import wannierberri
        return wannierberri
    except Exception as err:   # pragma: no cover
        pytest.skip(f"reference not importable: {err}")
    finally:
        sys.path.remove(ref)
        sys.path.remove(stubs)


def test_result_files_interchangeable_with_reference(tmp_path):
    """`EnergyResult.savedata` writes what the reference writes (energyresult.py:224-248): `<prefix>-<key>_iter-NNNN.npz`
    with the transforms as `Transform.as_dict()` dictionaries, plus the `.dat` table when "txt" is in `save_mode`; the
    reference's `EnergyResult.from_npz` reads the file back with its symmetry behaviour intact, and this package reads
    a file written by the reference."""
    from wannierberri_b200.result import EnergyResult
    Ef = np.linspace(0., 1., 4)
    data = np.arange(12.).reshape(4, 3)
    r = EnergyResult(Ef, data, transformTR="odd", transformInv="ident", comment="two\nlines", save_mode="bin+txt")
    r.savedata("ahc", str(tmp_path / "res"), "tag", 3)
    npz, dat = tmp_path / "res-ahc-tag_iter-0003.npz", tmp_path / "res-ahc-tag_iter-0003.dat"
    assert npz.is_file() and dat.is_file()
    lines = dat.read_text().split("\n")
    assert lines[0] == "#### two" and lines[1] == "#### lines" and lines[2].startswith("#") and len(lines) == 3 + 4
    assert np.allclose(np.array(lines[3].split(), dtype=float), [0., 0., 1., 2., 0., 1., 2.])
    back = EnergyResult.from_npz(str(npz))
    assert back.transformTR == "odd" and back.transformInv == "ident" and np.array_equal(back.data, data)
    EnergyResult(Ef, data, transformTR="odd", transformInv="ident", save_mode="bin").savedata("x", str(tmp_path / "b"), "", 0)
    assert (tmp_path / "b-x_iter-0000.npz").is_file() and not (tmp_path / "b-x_iter-0000.dat").exists()
    wberri = _import_reference()
    from wannierberri.symmetry.point_symmetry import transform_odd, transform_ident
    ref = wberri.result.EnergyResult.from_npz(str(npz))
    assert ref.transformTR == transform_odd and ref.transformInv == transform_ident
    assert np.array_equal(ref.data, data) and ref.rank == 1
    ref.save(str(tmp_path / "fromref"))
    mine = EnergyResult.from_npz(str(tmp_path / "fromref.npz"))
    assert mine.transformTR == "odd" and mine.transformInv == "ident" and np.array_equal(mine.data, data)


def test_grid_respects_point_group_and_periodicity():
    """determineNK / autoNK (grid/grid.py:196-266): grids that break the point group are refused, the automatic FFT
    grid is searched among the symmetric ones, non-periodic directions get one k-point; against the reference's Grid
    where it is importable."""
    fe = wb.System_R.from_npz(os.path.join(GOLDEN, "fe_system.npz"), pointgroup=["C4z", "C2x*TimeReversal", "Inversion"])
    with pytest.raises(AssertionError):
        wb.Grid(fe, NKdiv=[2, 3, 2], NKFFT=[4, 4, 4])     # C4z maps x on y: 2 != 3
    with pytest.raises(AssertionError):
        wb.Grid(fe, NK=[8, 8, 8], NKFFT=[4, 2, 4])
    wb.Grid(fe, NKdiv=[2, 3, 2], NKFFT=[4, 4, 4], use_symmetry=False)   # trivial group: anything goes
    g = wb.Grid(fe, NK=[30, 30, 30])
    assert fe.pointgroup.symmetric_grid(g.FFT) and fe.pointgroup.symmetric_grid(g.div)
    te = wb.System_R.from_npz(os.path.join(GOLDEN, "te_system.npz"), pointgroup=["C3z", "C2x", "TimeReversal"])
    gt = wb.Grid(te, NK=[30, 30, 40])
    assert gt.FFT[0] == gt.FFT[1] and gt.div[0] == gt.div[1]
    slab = wb.System_R.from_npz(os.path.join(GOLDEN, "fe_system.npz"))
    slab.periodic = np.array([True, True, False])
    gs = wb.Grid(slab, NKdiv=[2, 2, 2], NKFFT=[4, 4, 4])
    assert list(gs.div) == [2, 2, 1] and list(gs.FFT) == [4, 4, 1]
    wberri = _import_reference()
    from wannierberri.grid.grid import determineNK
    from wannierberri.symmetry.point_symmetry import PointGroup
    for system, gens, NK in ((fe, ["C4z", "C2x*TimeReversal", "Inversion"], [30, 30, 30]), (fe, ["C4z", "C2x*TimeReversal", "Inversion"], [48, 48, 48]),
                             (te, ["C3z", "C2x", "TimeReversal"], [30, 30, 40]), (te, ["C3z", "C2x", "TimeReversal"], [17, 17, 9])):
        pg = PointGroup(gens, real_lattice=system.real_lattice)
        div, fft = determineNK(np.array([True] * 3), None, None, np.array(NK), np.array(system.NKFFT_recommended), pg)
        mine = wb.Grid(system, NK=NK)
        assert np.array_equal(mine.div, div) and np.array_equal(mine.FFT, fft), (NK, mine.div, mine.FFT, div, fft)


def test_tetra_needs_uniform_fermi_axis():
    st = wb.calculators.static
    st.DOS(Efermi=np.linspace(0, 1, 11), tetra=True)
    st.DOS(Efermi=np.array([0., 0.1, 0.3]))                 # plain scan: the reference does not check either
    with pytest.raises(NotImplementedError):
        st.DOS(Efermi=np.array([0., 0.1, 0.3]), tetra=True)
    assert wb.calculators.dynamic.JDOS(Efermi=np.linspace(0, 1, 3), omega=np.linspace(0, 1, 3)).spec().external_terms == 0


def test_reference_calculators_on_datak_host():
    """Plug-in hook #1: UNMODIFIED calculators and formulae of the reference evaluate on this
    package's `DataKHost` interface -- `E_K`, `Xbar`, `D_H`, `dEig_inv`, `Dcov`, `covariant()` with comma- and
    generalised derivatives, `get_bands_in_range_groups` -- and give what they give on the reference's own `Data_K_R`.
    Here (no GPU) the three primitives of `DataKHost` are taken from the reference's Data_K_R of the same K-block, so
    that the test isolates the host logic; on the GPU box the primitives of `wannierberri_b200.Data_K_R` are checked
    against the oracle and a user-defined formula runs end to end (tests/test_gpu_parity.py)."""
    wberri = _import_reference()
    sys.path[:0] = ["/root/reference", os.path.join(ROOT, "oracle", "stubs"), os.path.join(ROOT, "tests", "golden")]
    try:
        from make_golden import build_fe
        from wannierberri.data_K import Data_K_R as RefDataK
        from wannierberri.calculators import static as rst
    finally:
        del sys.path[:3]
    from wannierberri_b200.data_K import DataKHost
    from plugin_formula import make_calculators
    system = build_fe()
    grid = wberri.Grid(system, NK=[4, 4, 4], NKFFT=[2, 2, 2])
    dK = np.array([0.125, 0.0, 0.125])

    class Host(DataKHost):
        def __init__(self, ref):
            self.ref, self.system = ref, ref.system
            self.nk, self.num_wann, self.cell_volume = ref.nk, ref.num_wann, ref.cell_volume
            self.force_internal_terms_only = ref.force_internal_terms_only

        def _eig(self):
            return self.ref.E_K, self.ref.UU_K

        def _xbar(self, name, der):
            return self.ref.Xbar(name, der)

    Ef = np.linspace(15., 19., 9)
    calcs = dict(ahc=rst.AHC(Efermi=Ef), morb=rst.Morb(Efermi=Ef), bcd_sea=rst.BerryDipole_FermiSea(Efermi=Ef),
                 gme_spin_sea=rst.GME_spin_FermiSea(Efermi=Ef), ohmic_sea=rst.Ohmic_FermiSea(Efermi=Ef),
                 ahc_kramers=rst.AHC(Efermi=Ef, degen_thresh=0.05, degen_Kramers=True),
                 zeeman_spin=rst.AHC_Zeeman_spin(Efermi=Ef),
                 # second-order formulae of the reference (Der2Spin / Der2Omega / Der2Morb, eMChA, quantum metric with
                 # FF = rotAAab): second comma-derivatives through Xbar(name, 2), evaluated by the reference's own classes
                 nldrude_z_spin=rst.NLDrude_Zeeman_spin(Efermi=Ef), nldrude_z_orb=rst.NLDrude_Zeeman_orb(Efermi=Ef),
                 emcha=rst.eMChA_FermiSurf(Efermi=Ef),
                 qmetric=rst.QuantumMetric_FermiSea(Efermi=Ef, kwargs_formula=dict(FF_rotAA=True)),
                 qmetric_dip=rst.QuantumMetric_Vel_DQ(Efermi=Ef, kwargs_formula=dict(FF_rotAA=True)))
    calcs.update(make_calculators(rst.StaticCalculator, Ef))
    for key, c in calcs.items():
        want = c(RefDataK(system, dK=dK, grid=grid)).data
        got = c(Host(RefDataK(system, dK=dK, grid=grid))).data
        assert np.abs(got - want).max() <= 1e-12 * max(np.abs(want).max(), 1e-300), key
    # this package's plug-in calculator (host loop of StaticCalculator._call_plugin) against the reference's loop
    mine = make_calculators(wb.calculators.static.StaticCalculator, Ef)
    for key, c in mine.items():
        want = calcs[key](RefDataK(system, dK=dK, grid=grid)).data
        got = c(Host(RefDataK(system, dK=dK, grid=grid))).data
        assert np.abs(got - want).max() <= 1e-12 * max(np.abs(want).max(), 1e-300), key


def test_second_order_block_algebra_vs_reference(monkeypatch):
    """formula_gpu.py (full-matrix block algebra with partition masks, batched over the (k-point, band group) pairs of a
    K-block) against the reference's per-group numpy formulae -- Der2Spin / Der2Omega / Der2Morb with Der3E, Omega,
    Morb_Hpm (NLDrude_Z_*), emcha_surf, tildeFab / tildeFab_d (quantum metric) -- and this package's calculators of them
    against the reference's, on the reference's Data_K_R of one Fe K-block.  The tensors live on the CPU here (torch),
    the product evaluates the same code on the CUDA device (tests/test_gpu_parity.py: through run() vs a fixture)."""
    wberri = _import_reference()
    sys.path[:0] = ["/root/reference", os.path.join(ROOT, "oracle", "stubs"), os.path.join(ROOT, "tests", "golden")]
    try:
        from make_golden import build_fe
        from wannierberri.data_K import Data_K_R as RefDataK
        from wannierberri.calculators import static as rst
        from wannierberri.formula import covariant as frml
    finally:
        del sys.path[:3]
    import functools
    from wannierberri_b200 import formula_gpu as fg
    system = build_fe()
    grid = wberri.Grid(system, NK=[4, 4, 4], NKFFT=[2, 2, 2])
    dk = RefDataK(system, dK=np.array([0.125, 0.0, 0.125]), grid=grid)
    nw = dk.num_wann
    ref_cls = dict(NLDrude_Z_spin=frml.NLDrude_Z_spin, NLDrude_Z_orb_Omega=frml.NLDrude_Z_orb_Omega,
                   NLDrude_Z_orb_Hplus=frml.NLDrude_Z_orb_Hplus, emcha_surf=frml.emcha_surf,
                   QuantumMetric_ab=frml.QuantumMetric_ab, VelDQM=frml.VelDQM)
    # band sets: single bands, wide Kramers-paired groups, and the Fermi-sea set [0, bandmax)
    for kw_groups in (dict(degen_thresh=1e-4, sea=False), dict(degen_thresh=0.3, degen_Kramers=True, sea=True)):
        groups = dk.get_bands_in_range_groups(15., 19., **kw_groups)
        pairs = [(ik, g) for ik, gs in enumerate(groups) for g in gs][::3]
        kidx, a, b = (np.array(x) for x in zip(*[(ik, g[0], g[1]) for ik, g in pairs]))
        for ext in (True, False):
            for name, cls in ref_cls.items():
                kw = dict(external_terms=ext)
                if name in ("QuantumMetric_ab", "VelDQM"):
                    kw["FF_rotAA"] = True
                f = cls(dk, **kw)
                want = np.array([f.trace(ik, np.arange(x, y), np.concatenate((np.arange(0, x), np.arange(y, nw))))
                                 for ik, x, y in zip(kidx, a, b)])
                got = fg.batch_traces(name, dk, kidx, a, b, internal=True, external=ext, device="cpu")
                assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max(), (name, ext, kw_groups)
    # the calculators (scan, factors, the Hplus - 2 E_F Omega combination) on the same K-block
    from second_order_calcs import make_calculators
    Ef = np.linspace(15., 19., 9)
    monkeypatch.setattr(fg, "batch_traces", functools.partial(fg.batch_traces, device="cpu"))
    mine, theirs = make_calculators(wb.calculators.static, Ef), make_calculators(rst, Ef)
    for key in ("z_spin", "z_orb", "emcha_wide", "qmetric", "qmetric_dip"):
        want, got = theirs[key](dk).data, mine[key](dk).data
        assert np.abs(got - want).max() <= 1e-11 * np.abs(want).max(), key
    with pytest.raises(NotImplementedError):   # FF_R is not transformed on the GPU path
        wb.calculators.static.QuantumMetric_FermiSea(Efermi=Ef)


def test_run_plugin_loop_with_reference_calculators():
    """`run()` with calculators that are NOT scans of the library -- here unmodified objects of the reference
    (NLDrude_Zeeman_spin: second-order formula; a user class) -- loops over the K-blocks itself, calls
    `calc(data_k_class(system, dK, grid))`, weights, sums and symmetrises (run_grid.py:258-265, 59-72).  With a
    `DataKHost` subclass whose primitives come from the reference's Data_K_R no GPU is touched, so the loop is checked
    here against the reference's own `run()` with the same calculators (incl. the irreducible K-list + symmetrisation)."""
    wberri = _import_reference()
    sys.path[:0] = ["/root/reference", os.path.join(ROOT, "oracle", "stubs"), os.path.join(ROOT, "tests", "golden")]
    try:
        from make_golden import build_fe
        from wannierberri.data_K import Data_K_R as RefDataK
        from wannierberri.calculators import static as rst
    finally:
        del sys.path[:3]
    from wannierberri_b200.data_K import DataKHost
    from plugin_formula import make_calculators
    ref_system = build_fe()
    ref_grid = wberri.Grid(ref_system, NK=[4, 4, 4], NKFFT=[2, 2, 2])

    class Host(DataKHost):
        def __init__(self, system, dK, grid, device=0, **kw):
            self.ref = RefDataK(ref_system, dK=np.array(dK), grid=ref_grid)
            self.system, self.grid, self.dK = system, grid, np.array(dK)
            self.nk, self.num_wann, self.cell_volume = self.ref.nk, self.ref.num_wann, self.ref.cell_volume
            self.force_internal_terms_only = False

        def _eig(self):
            return self.ref.E_K, self.ref.UU_K

        def _xbar(self, name, der):
            return self.ref.Xbar(name, der)

    Ef = np.linspace(16., 18., 5)
    calcs = dict(z_spin=rst.NLDrude_Zeeman_spin(Efermi=Ef), ahc=rst.AHC(Efermi=Ef, kwargs_formula=dict(external_terms=False)))
    calcs.update({k: v for k, v in make_calculators(rst.StaticCalculator, Ef).items() if k == "user_surf"})
    fe = wb.System_R.from_npz(os.path.join(GOLDEN, "fe_system.npz"), pointgroup=["C4z", "C2x*TimeReversal", "Inversion"])
    scale = {}   # a quantity that vanishes by symmetry is compared on the scale of its unsymmetrised value
    for sym in (False, True):
        want = wberri.run(ref_system, grid=ref_grid, calculators=calcs, adpt_num_iter=0, use_irred_kpt=sym, symmetrize=sym,
                          parallel=False, fout_name=os.path.join("/tmp", "plugloop"), print_progress_step_time=1e9)
        # `ahc` and `z_spin` have the names of calculators of this package (which would replace them by its own: a scan of
        # the CUDA kernels / the batched device algebra): hand them over as plain callables so that they take the plug-in loop
        wrap = lambda c: (lambda data_K: c(data_K))   # noqa: E731
        mine = dict(calcs, ahc=wrap(calcs["ahc"]), z_spin=wrap(calcs["z_spin"]))
        got = wb.run(fe, wb.Grid(fe, NK=[4, 4, 4], NKFFT=[2, 2, 2]), mine, use_irred_kpt=sym, symmetrize=sym,
                     data_k_class=Host, write_files=False, parallel=False)
        for key in calcs:
            a, b = got.results[key].data, want.results[key].data
            scale.setdefault(key, np.abs(b).max())
            assert a.shape == b.shape and np.abs(a - b).max() <= 1e-10 * scale[key], (key, sym)
