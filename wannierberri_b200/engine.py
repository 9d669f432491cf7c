"""Engine: one System_R replica resident on one GPU, behind the C-ABI of libwbgpu.so."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import ScanSpec, check, dptr, as_f64
from .system import as_system


class Engine:
    """Owns a `wbgpu_ctx`.  `plan()` fixes the FFT sub-grid and the formulae, `scan()` evaluates
    Fermi scans over a list of K-blocks (the GPU replacement of the K-block loop of
    run_grid.py:258-265 + StaticCalculator.__call__)."""

    def __init__(self, system, device=0, stream=None):
        self.system = as_system(system)
        L = _lib.lib()
        self._L = L
        self.nw = int(system.num_wann)
        iRvec = np.ascontiguousarray(system.rvec.iRvec, dtype=np.int32)
        T = as_f64(system.rvec.cRvec_shifted)
        self.nR = iRvec.shape[0]
        if T.shape != (self.nR, self.nw, self.nw, 3):
            raise ValueError(f"cRvec_shifted has shape {T.shape}")
        self._ctx = C.c_void_p()
        check(L.wbgpu_create(C.byref(self._ctx), int(device), self.nw, self.nR,
                             iRvec.ctypes.data_as(C.POINTER(C.c_int32)), dptr(T), float(system.cell_volume),
                             C.c_void_p(stream)))
        self.device = int(device)
        self._plan_key = None
        for key, idx in _lib.KEYS.items():
            if system.has_R_mat(key):
                X = np.ascontiguousarray(system.get_R_mat(key), dtype=np.complex128)
                ncart = _lib.KEY_NCART.get(key, 3)
                want = (self.nR, self.nw, self.nw) + {1: (), 3: (3,), 9: (3, 3)}[ncart]
                if X.shape != want:
                    raise ValueError(f"R-matrix {key} has shape {X.shape}, expected {want}")
                check(L.wbgpu_set_R_matrix(self._ctx, idx, dptr(X.view(np.float64)), ncart))

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self._L.wbgpu_destroy(self._ctx)
            self._ctx = C.c_void_p()

    __del__ = close

    # ------------------------------------------------------------------
    def plan(self, NKFFT, formulae, external_terms=True, max_kpoints_per_launch=0):
        NKFFT = np.ascontiguousarray(NKFFT, dtype=np.int32)
        mask = 0
        for f in formulae:
            mask |= 1 << int(f)
        key = (tuple(NKFFT), mask, bool(external_terms), int(max_kpoints_per_launch))
        if key == self._plan_key:
            return
        check(self._L.wbgpu_plan(self._ctx, NKFFT.ctypes.data_as(C.POINTER(C.c_int32)), mask, int(bool(external_terms)),
                                 int(max_kpoints_per_launch)))
        self._plan_key = key
        self.NKFFT = NKFFT.copy()
        self.nk = int(np.prod(NKFFT))

    def set_option(self, name, value):
        check(self._L.wbgpu_set_option(self._ctx, name.encode(), int(value)))

    def scan(self, dK, weight, specs):
        """Host buffers in, host arrays out: list of arrays `[nEF, 3^rank]`, one per spec."""
        dK = as_f64(dK).reshape(-1, 3)
        weight = as_f64(weight).reshape(-1)
        assert weight.shape[0] == dK.shape[0]
        arr = (ScanSpec * len(specs))(*specs)
        out = np.zeros(sum(s.size for s in specs))
        check(self._L.wbgpu_static_scan(self._ctx, dK.shape[0], dptr(dK), dptr(weight), arr, len(specs), dptr(out)))
        res, off = [], 0
        for s in specs:
            res.append(out[off:off + s.size].reshape(s.shape).copy())
            off += s.size
        return res

    def scan_blocks(self, dK, specs):
        """Per-K-block results: list (one per spec) of arrays `[nblocks, nEF, 3^rank]`."""
        dK = as_f64(dK).reshape(-1, 3)
        nb = dK.shape[0]
        arr = (ScanSpec * len(specs))(*specs)
        total = sum(s.size for s in specs)
        out = np.zeros((nb, total))
        check(self._L.wbgpu_static_scan_blocks(self._ctx, nb, dptr(dK), arr, len(specs), dptr(out)))
        res, off = [], 0
        for s in specs:
            res.append(out[:, off:off + s.size].reshape((nb,) + s.shape).copy())
            off += s.size
        return res

    def scan_tetra(self, dK, weight, dK_cell, specs):
        """`scan()` with the tetrahedron method: `dK_cell` = Kpoint.dK_fullBZ = 1 / (NKdiv * NKFFT)."""
        dK = as_f64(dK).reshape(-1, 3)
        weight = as_f64(weight).reshape(-1)
        dK_cell = as_f64(dK_cell).reshape(3)
        arr = (ScanSpec * len(specs))(*specs)
        out = np.zeros(sum(s.size for s in specs))
        check(self._L.wbgpu_static_scan_tetra(self._ctx, dK.shape[0], dptr(dK), dptr(weight), dptr(dK_cell), arr,
                                              len(specs), dptr(out)))
        res, off = [], 0
        for s in specs:
            res.append(out[off:off + s.size].reshape(s.shape).copy())
            off += s.size
        return res

    def scan_tetra_blocks(self, dK, dK_cell, specs):
        """Per-K-block tetrahedron scans (refinement loop): `dK_cell[nblocks][3]`, list of `[nblocks, nEF, 3^rank]`."""
        dK = as_f64(dK).reshape(-1, 3)
        nb = dK.shape[0]
        dK_cell = as_f64(np.broadcast_to(np.asarray(dK_cell, dtype=float), (nb, 3)))
        arr = (ScanSpec * len(specs))(*specs)
        total = sum(s.size for s in specs)
        out = np.zeros((nb, total))
        check(self._L.wbgpu_static_scan_tetra_blocks(self._ctx, nb, dptr(dK), dptr(dK_cell), arr, len(specs), dptr(out)))
        res, off = [], 0
        for s in specs:
            res.append(out[:, off:off + s.size].reshape((nb,) + s.shape).copy())
            off += s.size
        return res

    def scan_dev(self, dK_dev, weight_dev, specs, out_dev):
        """Device-resident variant: torch CUDA tensors (float64) for dK[nb,3], weight[nb], out[sum sizes];
        asynchronous on the context's stream."""
        nb = int(dK_dev.shape[0])
        arr = (ScanSpec * len(specs))(*specs)
        check(self._L.wbgpu_static_scan_dev(self._ctx, nb, C.c_void_p(dK_dev.data_ptr()),
                                            C.c_void_p(weight_dev.data_ptr()), arr, len(specs),
                                            C.c_void_p(out_dev.data_ptr())))

    def kubo_scan(self, dK, weight, spec, Efermi, omega):
        """One (Efermi x omega) scan over a list of K-blocks: `[nEF, nomega, 3, 3]` complex (optical conductivity),
        `[nEF, nomega, 3, 3, 3]` complex (spin Hall conductivity) or `[nEF, nomega]` float (JDOS)."""
        dK = as_f64(dK).reshape(-1, 3)
        weight = as_f64(weight).reshape(-1)
        Efermi, omega = as_f64(Efermi), as_f64(omega)
        assert weight.shape[0] == dK.shape[0] and len(Efermi) == spec.nEF and len(omega) == spec.nomega
        out = np.zeros(int(self._L.wbgpu_kubo_size(C.byref(spec))))
        check(self._L.wbgpu_kubo_scan(self._ctx, dK.shape[0], dptr(dK), dptr(weight), C.byref(spec), dptr(Efermi),
                                      dptr(omega), dptr(out)))
        if spec.is_complex:
            return out.view(np.complex128).reshape(spec.shape)
        return out.reshape(spec.shape)

    def kubo_scan_blocks(self, dK, spec, Efermi, omega):
        """Per-K-block Kubo scans (refinement loop): `[nblocks, nEF, nomega, ...]`."""
        dK = as_f64(dK).reshape(-1, 3)
        nb = dK.shape[0]
        Efermi, omega = as_f64(Efermi), as_f64(omega)
        n = int(self._L.wbgpu_kubo_size(C.byref(spec)))
        out = np.zeros((nb, n))
        check(self._L.wbgpu_kubo_scan_blocks(self._ctx, nb, dptr(dK), C.byref(spec), dptr(Efermi), dptr(omega), dptr(out)))
        if spec.is_complex:
            return [out[b].view(np.complex128).reshape(spec.shape).copy() for b in range(nb)]
        return [out[b].reshape(spec.shape).copy() for b in range(nb)]

    def kubo_scan_dev(self, dK_dev, weight_dev, spec, Efermi, omega, out_dev):
        """Device-resident variant of `kubo_scan`: torch CUDA tensors (float64) for dK[nb,3], weight[nb] and the flat
        output (complex as re, im pairs); asynchronous on the context's stream."""
        Efermi, omega = as_f64(Efermi), as_f64(omega)
        check(self._L.wbgpu_kubo_scan_dev(self._ctx, int(dK_dev.shape[0]), C.c_void_p(dK_dev.data_ptr()),
                                          C.c_void_p(weight_dev.data_ptr()), C.byref(spec), dptr(Efermi), dptr(omega),
                                          C.c_void_p(out_dev.data_ptr())))

    # ------------------------------------------------------------------ parity probes
    def kpoints(self, dK):
        dK = as_f64(dK)
        out = np.zeros((self.nk, 3))
        check(self._L.wbgpu_kpoints(self._ctx, dptr(dK), dptr(out)))
        return out

    def eig(self, dK, vectors=False):
        dK = as_f64(dK)
        E = np.zeros((self.nk, self.nw))
        U = np.zeros((self.nk, self.nw, self.nw), dtype=np.complex128) if vectors else None
        check(self._L.wbgpu_eig(self._ctx, dptr(dK), dptr(E), dptr(U.view(np.float64)) if vectors else None))
        return (E, U) if vectors else E

    def xk(self, dK, channel):
        dK = as_f64(dK)
        ncart = 1 if channel == "Ham" else 3
        X = np.zeros((self.nk, self.nw, self.nw) + ((3,) if ncart == 3 else ()), dtype=np.complex128)
        check(self._L.wbgpu_xk(self._ctx, dptr(dK), _lib.CHANNELS[channel], dptr(X.view(np.float64))))
        return X

    def xbar(self, dK, name, der=0):
        """Hamiltonian-gauge matrix U^dagger (d^der X) U of one K-block, `[nk][nw][nw][3]^(ncart + der)`"""
        dK = as_f64(dK)
        ncart = (0 if name == "Ham" else 1) + int(der)
        X = np.zeros((self.nk, self.nw, self.nw) + (3,) * ncart, dtype=np.complex128)
        check(self._L.wbgpu_xbar(self._ctx, dptr(dK), _lib.CHANNELS[name], int(der), dptr(X.view(np.float64))))
        return X

    def band_traces(self, dK, spec):
        dK = as_f64(dK)
        rank = _lib.FORMULA_RANK[int(spec.formula)]
        lab = np.zeros((self.nk, self.nw))
        val = np.zeros((self.nk, self.nw) + (3,) * rank)
        check(self._L.wbgpu_band_traces(self._ctx, dptr(dK), C.byref(spec), dptr(lab), dptr(val)))
        return lab, val

    @property
    def kernel_launches(self):
        return int(self._L.wbgpu_kernel_launches(self._ctx))

    @property
    def last_eig_sweeps(self):
        return int(self._L.wbgpu_last_eig_sweeps(self._ctx))

    @property
    def last_eig_resolved(self):
        """k-points of the last `eig` call that the fast eigensolver handed to the Jacobi kernel"""
        return int(self._L.wbgpu_last_eig_resolved(self._ctx))
