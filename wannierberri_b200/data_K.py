"""GPU-resident per-K-block state with the constructor signature of the reference's `Data_K_R`
(data_K/data_K.py:73-83, data_K_R.py:11-22): `cls(system, dK=Kpoint.Kp_fullBZ, grid=grid, Kpoint=Kpoint, **parameters_K)`,
so it can be passed as `data_k_class` and calculators can be called on it one K-block at a time."""
import numpy as np

from .engine import Engine

_ENGINES = {}
_MAX_ENGINES = 3   # every engine owns device work space (wbgpu_plan: up to 48 GB): the least recently used one is closed


def engine_for(system, device=0):
    key = (id(system), device)
    eng = _ENGINES.pop(key, None)
    if eng is None or eng.system is not system:
        if eng is not None:
            eng.close()
        while len(_ENGINES) >= _MAX_ENGINES:
            _ENGINES.pop(next(iter(_ENGINES))).close()
        eng = Engine(system, device=device)
    _ENGINES[key] = eng   # most recently used last
    return eng


def check_parameters_K(parameters_K):
    """`parameters_K` of run() / keyword arguments of Data_K (data_K/data_K.py:73-83).  `fftlib` names the CPU FFT
    backend of the reference: there is one R->k transform here (the CUDA one), so a valid name is accepted and has no
    effect, an unknown one raises as in fourier/fft.py:63; defaults of the other parameters are accepted, anything
    else raises (no CPU fallback)."""
    parameters_K = dict(parameters_K or {})
    lib = parameters_K.pop("fftlib", "fftw")
    if str(lib).lower() not in ("fftw", "numpy", "slow"):
        raise ValueError(f"fftlib '{lib}' is unknown/not supported")
    defaults = dict(Emin=-np.inf, Emax=np.inf, random_gauge=False)
    for key in list(parameters_K):
        if key in defaults and parameters_K[key] == defaults[key]:
            parameters_K.pop(key)
    parameters_K.pop("degen_thresh_random_gauge", None)   # read only with random_gauge=True
    if parameters_K:
        raise NotImplementedError(f"parameters_K {sorted(parameters_K)} are not implemented on the GPU path")


class DataKHost:
    """What calculators and plug-in formulae read from a `Data_K` beyond the scans (data_K/data_K.py:138-334): derived
    host-side quantities of the three primitives `E_K`, `UU_K` and `Xbar(name, der)` that a subclass supplies."""

    _xbar_cache = None
    _cov_cache = None
    is_phonon = False
    Emin, Emax = -np.inf, np.inf

    # ---- primitives (subclass)
    def _eig(self):
        raise NotImplementedError

    def _xbar(self, name, der):
        raise NotImplementedError

    # ---- basic variables
    @property
    def nbands(self):
        return self.num_wann

    @property
    def real_lattice(self):
        return self.system.real_lattice

    @property
    def E_K(self):
        if getattr(self, "_E_K", None) is None:
            self._E_K, self._UU_K = self._eig()
        return self._E_K

    @property
    def UU_K(self):
        self.E_K
        return self._UU_K

    def Xbar(self, name, der=0):
        """U^dagger (d^der X) U, `[nk][nw][nw][3]^(ncart + der)` (data_K/data_K_R.py:69-97)"""
        if self._xbar_cache is None:
            self._xbar_cache = {}
        key = (name, der)
        if key not in self._xbar_cache:
            if name in ("rotAAab", "CCab_antisym"):
                # antisymmetric rank-2 forms of curl A / C (data_K_R.py:53-66): X_ab = -+ i/2 eps_abc X_c, linear, so the
                # same in the Hamiltonian gauge and after any number of comma-derivatives
                X = self.Xbar("rotAA" if name == "rotAAab" else "CC", der)
                res = np.zeros(X.shape[:3] + (3, 3) + X.shape[4:], dtype=complex)
                al, be = np.array([1, 2, 0]), np.array([2, 0, 1])
                res[:, :, :, al, be] = -0.5j * X
                res[:, :, :, be, al] = 0.5j * X
                self._xbar_cache[key] = res
            else:
                self._xbar_cache[key] = self._xbar(name, der)
        return self._xbar_cache[key]

    @property
    def delE_K(self):
        """band velocities: the diagonal of Xbar('Ham', 1) (data_K.py:236-242)"""
        d = np.einsum("klla->kla", self.Xbar("Ham", 1))
        worst = np.abs(d.imag).max()
        if worst > 1e-10:
            raise RuntimeError(f"The band derivatives have considerable imaginary part: {worst}")
        return d.real

    @property
    def dEig_inv(self):
        """1 / (E_m - E_n), zero where the two energies are closer than 1e-7 (data_K.py:290-298)"""
        dE = self.E_K[:, :, None] - self.E_K[:, None, :]
        close = np.abs(dE) < 1e-7
        dE[close] = 1.
        inv = 1. / dE
        inv[close] = 0.
        return inv

    @property
    def D_H(self):
        """D^H_a = -Vbar_a / (E_m - E_n) (data_K.py:324-326)"""
        return -self.Xbar("Ham", 1) * self.dEig_inv[:, :, :, None]

    def get_A_H(self, external_terms=True):
        """generalised Berry connection i D^H (+ Abar) (data_K.py:328-334)"""
        A = 1j * self.D_H
        if external_terms:
            A = A + self.Xbar("AA")
        return A

    @property
    def Dcov(self):
        from .formula import Dcov
        return Dcov(self)

    @property
    def V_covariant(self):
        from .formula import Velocity_ln
        return Velocity_ln(self.Xbar("Ham", 1))

    def covariant(self, name, commader=0, gender=0, save=True):
        """the matrix `name` with `commader` comma-derivatives, or its generalised derivative (`gender` = 1), as an
        object with `nn` / `ln` / `nl` / `ll` blocks (data_K.py:244-271)"""
        from . import formula
        assert commader * gender == 0, "cannot mix comm and generalized derivatives"
        if self._cov_cache is None:
            self._cov_cache = {}
        key = (name, commader, gender)
        if key in self._cov_cache:
            return self._cov_cache[key]
        if gender == 0:
            res = formula.Matrix_ln(self.Xbar(name, commader), transformTR=formula.transform_TR(name, commader),
                                    transformInv=formula.transform_Inv(name, commader))
        elif gender == 1:
            if name == "Ham":
                res = self.V_covariant
            else:
                res = formula.Matrix_GenDer_ln(self.covariant(name), self.covariant(name, commader=1), self.Dcov,
                                               transformTR=formula.transform_TR(name, gender),
                                               transformInv=formula.transform_Inv(name, gender))
        else:
            raise NotImplementedError()
        if save:
            self._cov_cache[key] = res
        return res

    def get_bands_in_range_groups_ik(self, ik, emin, emax, degen_thresh=-1, degen_Kramers=False, sea=False,
                                     Emin=-np.inf, Emax=np.inf, select_bands=None):
        from .formula import band_groups
        return band_groups(self.E_K[ik], emin, emax, degen_thresh, degen_Kramers, sea, select_bands)

    def get_bands_in_range_groups(self, emin, emax, degen_thresh=-1, degen_Kramers=False, sea=False, Emin=-np.inf,
                                  Emax=np.inf, select_bands=None):
        return [self.get_bands_in_range_groups_ik(ik, emin, emax, degen_thresh, degen_Kramers, sea, Emin, Emax, select_bands)
                for ik in range(self.nk)]


# which scan formula makes wbgpu_plan keep the channel that Xbar(name, der) reads
def _xbar_plan_formula(name, der):
    from . import _lib
    table = {("Ham", 0): _lib.IDENTITY, ("Ham", 1): _lib.VEL_VEL, ("Ham", 2): _lib.INV_MASS, ("Ham", 3): _lib.DER3E,
             ("AA", 0): _lib.OMEGA, ("AA", 1): _lib.DER_OMEGA, ("rotAA", 0): _lib.OMEGA, ("rotAA", 1): _lib.DER_OMEGA,
             ("BB", 0): _lib.MORB_HPM, ("BB", 1): _lib.DER_MORB, ("CC", 0): _lib.MORB_HPM, ("CC", 1): _lib.DER_MORB,
             ("SS", 0): _lib.SPIN, ("SS", 1): _lib.DER_SPIN,
             ("AA", 2): _lib.XBAR_DER2, ("rotAA", 2): _lib.XBAR_DER2, ("BB", 2): _lib.XBAR_DER2, ("CC", 2): _lib.XBAR_DER2,
             ("SS", 2): _lib.XBAR_DER2}
    if (name, der) not in table:
        raise NotImplementedError(f"Xbar('{name}', der={der}) is not available on the GPU path")
    return table[(name, der)]


class Data_K_R(DataKHost):

    def __init__(self, system, dK, grid, Kpoint=None, device=0, **parameters_K):
        check_parameters_K(parameters_K)
        self.system = system
        self.grid = grid
        self.Kpoint = Kpoint
        self.dK = np.array(dK, dtype=float)
        self.NKFFT = np.array(grid.FFT, dtype=int)
        self.nk = int(np.prod(self.NKFFT))
        self.num_wann = system.num_wann
        self.cell_volume = system.cell_volume
        self.force_internal_terms_only = getattr(system, "force_internal_terms_only", False)
        self.engine = engine_for(system, device)
        self._formulae = set()
        self._external = False

    def _plan(self, formulae, external_terms=True):
        from . import _lib
        self.engine.plan(self.NKFFT, set(formulae) | {_lib.IDENTITY}, external_terms=external_terms)

    def _plan_more(self, formula, external_terms):
        """plans of the probes accumulate, so that a formula that reads several matrices re-plans once per new channel"""
        self._formulae.add(formula)
        self._external = self._external or external_terms
        self._plan(self._formulae, self._external)

    # ---- primitives of DataKHost, evaluated by the CUDA kernels
    def _eig(self):
        self._plan_more(_xbar_plan_formula("Ham", 0), False)
        return self.engine.eig(self.dK, vectors=True)

    def _xbar(self, name, der):
        self._plan_more(_xbar_plan_formula(name, der), name in ("AA", "rotAA", "BB", "CC"))
        return self.engine.xbar(self.dK, name, der)

    @property
    def HH_K(self):
        self._plan_more(_xbar_plan_formula("Ham", 0), False)
        return self.engine.xk(self.dK, "Ham")

    def scan(self, specs, external_terms=True, tetra=False):
        if self.force_internal_terms_only:
            for s in specs:
                s.external_terms = 0
            external_terms = False
        self._plan([s.formula for s in specs], external_terms)
        if tetra:
            # KpointBZparallel.dK_fullBZ (grid/Kpoint.py:107-109)
            dK_cell = 1. / (np.array(self.grid.div, dtype=float) * np.array(self.grid.FFT, dtype=float))
            return self.engine.scan_tetra(self.dK[None, :], np.ones(1), dK_cell, specs)
        return self.engine.scan(self.dK[None, :], np.ones(1), specs)

    def kubo_scan(self, spec, Efermi, omega):
        from . import _lib
        if self.force_internal_terms_only:
            spec.external_terms = 0
        self._plan([spec.formula_flag], bool(spec.external_terms))
        return self.engine.kubo_scan(self.dK[None, :], np.ones(1), spec, Efermi, omega)

    @property
    def kpoints_all(self):
        self._plan([])
        return self.engine.kpoints(self.dK)

