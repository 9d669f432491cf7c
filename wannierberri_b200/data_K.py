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


class Data_K_R:

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

    def _plan(self, formulae, external_terms=True):
        from . import _lib
        self.engine.plan(self.NKFFT, set(formulae) | {_lib.IDENTITY}, external_terms=external_terms)

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

    @property
    def E_K(self):
        self._plan([])
        return self.engine.eig(self.dK)
