"""ctypes binding of libwbgpu.so (include/wbgpu.h).  No CPU fallback: if the shared library is
missing, or there is no CUDA device when a context is created, this raises."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwbgpu.so")

# enums of include/wbgpu.h
(IDENTITY, OMEGA, MORB_HPM, VEL_OMEGA, VEL_HPLUS, VEL_SPIN, SPIN, KUBO, VEL_VEL, INV_MASS, SHC_RYOO, SHC_QIAO,
 SHC_SIMPLE, DER_OMEGA, DER_SPIN, VEL_VEL_VEL, MASS_VEL, MASS_MASS, VEL_MASS_VEL, OMEGA_S, OMEGA_OMEGA, SHIFT_CURRENT, DER3E, DER_MORB, OMEGA_HPLUS,
 XBAR_DER2) = range(26)   # XBAR_DER2: no scan; keeps the second comma-derivatives for Data_K_R.Xbar(name, 2)
SHC_TYPES = {"ryoo": SHC_RYOO, "qiao": SHC_QIAO, "simple": SHC_SIMPLE}
KUBO_OPTCOND, KUBO_JDOS, KUBO_SHC, KUBO_SHIFT, KUBO_INJECTION = 0, 1, 2, 3, 4
FORMULA_RANK = {IDENTITY: 0, OMEGA: 1, MORB_HPM: 1, SPIN: 1, VEL_OMEGA: 2, VEL_HPLUS: 2, VEL_SPIN: 2, VEL_VEL: 2,
                INV_MASS: 2, DER_OMEGA: 2, SHC_RYOO: 3, SHC_QIAO: 3, SHC_SIMPLE: 3, DER_SPIN: 2, VEL_VEL_VEL: 3, MASS_VEL: 3,
                MASS_MASS: 4, VEL_MASS_VEL: 4, OMEGA_S: 2, OMEGA_OMEGA: 2, DER3E: 3, DER_MORB: 2, OMEGA_HPLUS: 2}
KEYS = {"Ham": 0, "AA": 1, "BB": 2, "CC": 3, "SS": 4, "SA": 5, "SHA": 6, "SR": 7, "SH": 8, "SHR": 9}
KEY_NCART = {"Ham": 1, "SA": 9, "SHA": 9, "SR": 9, "SHR": 9}   # default 3
CHANNELS = {"Ham": 0, "dHam": 1, "AA": 2, "rotAA": 3, "BB": 4, "CC": 5, "SS": 6}


class ScanSpec(C.Structure):
    _fields_ = [("formula", C.c_int32), ("fder", C.c_int32), ("nEF", C.c_int32), ("degen_Kramers", C.c_int32),
                ("internal_terms", C.c_int32), ("external_terms", C.c_int32),
                ("Ef_first", C.c_double), ("Ef_last", C.c_double), ("dEF", C.c_double),
                ("degen_thresh", C.c_double), ("factor", C.c_double),
                ("select_mask", C.c_uint64 * 2), ("use_select", C.c_int32), ("tetra_flags", C.c_int32),
                ("tetra_Emin", C.c_double), ("tetra_Emax", C.c_double)]

    @property
    def size(self):
        return int(self.nEF) * 3 ** FORMULA_RANK[int(self.formula)]

    @property
    def shape(self):
        return (int(self.nEF),) + (3,) * FORMULA_RANK[int(self.formula)]


class KuboSpec(C.Structure):
    _fields_ = [("kind", C.c_int32), ("nEF", C.c_int32), ("nomega", C.c_int32), ("smr_type", C.c_int32),
                ("degen_Kramers", C.c_int32), ("external_terms", C.c_int32), ("shc_type", C.c_int32),
                ("reserved", C.c_int32),
                ("smr_fixed_width", C.c_double), ("degen_thresh", C.c_double), ("factor", C.c_double),
                ("sc_eta", C.c_double), ("kBT", C.c_double)]

    @property
    def shape(self):
        return (int(self.nEF), int(self.nomega)) + {KUBO_OPTCOND: (3, 3), KUBO_SHC: (3, 3, 3), KUBO_SHIFT: (3, 3, 3),
                                                     KUBO_INJECTION: (3, 3, 3)}.get(int(self.kind), ())

    @property
    def is_complex(self):
        return int(self.kind) in (KUBO_OPTCOND, KUBO_SHC, KUBO_INJECTION)

    @property
    def formula_flag(self):
        """the plan flag (wbgpu_plan formula_mask bit) this scan needs"""
        return {KUBO_SHC: int(self.shc_type), KUBO_SHIFT: SHIFT_CURRENT}.get(int(self.kind), KUBO)


_lib = None


def lib():
    """Load libwbgpu.so (built in-tree by `__graft_entry__.build()` / `python -m wannierberri_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m wannierberri_b200.build` "
                           "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    pd = C.POINTER(C.c_double)
    L.wbgpu_last_error.restype = C.c_char_p
    L.wbgpu_version.restype = C.c_int
    L.wbgpu_device_count.restype = C.c_int
    L.wbgpu_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.POINTER(i32), pd, dbl, vp]
    L.wbgpu_destroy.argtypes = [vp]
    L.wbgpu_set_R_matrix.argtypes = [vp, C.c_int, pd, C.c_int]
    L.wbgpu_plan.argtypes = [vp, C.POINTER(i32), C.c_uint32, C.c_int, i64]
    L.wbgpu_static_scan.argtypes = [vp, C.c_int, pd, pd, C.POINTER(ScanSpec), C.c_int, pd]
    L.wbgpu_static_scan_dev.argtypes = [vp, C.c_int, vp, vp, C.POINTER(ScanSpec), C.c_int, vp]
    L.wbgpu_static_scan_blocks.argtypes = [vp, C.c_int, pd, C.POINTER(ScanSpec), C.c_int, pd]
    L.wbgpu_static_scan_tetra.argtypes = [vp, C.c_int, pd, pd, pd, C.POINTER(ScanSpec), C.c_int, pd]
    L.wbgpu_static_scan_tetra_blocks.argtypes = [vp, C.c_int, pd, pd, C.POINTER(ScanSpec), C.c_int, pd]
    L.wbgpu_spec_size.argtypes = [C.POINTER(ScanSpec)]
    L.wbgpu_spec_size.restype = i64
    L.wbgpu_kubo_size.argtypes = [C.POINTER(KuboSpec)]
    L.wbgpu_kubo_size.restype = i64
    L.wbgpu_kubo_scan.argtypes = [vp, C.c_int, pd, pd, C.POINTER(KuboSpec), pd, pd, pd]
    L.wbgpu_kubo_scan_dev.argtypes = [vp, C.c_int, vp, vp, C.POINTER(KuboSpec), pd, pd, vp]
    L.wbgpu_kubo_scan_blocks.argtypes = [vp, C.c_int, pd, C.POINTER(KuboSpec), pd, pd, pd]
    L.wbgpu_kpoints.argtypes = [vp, pd, pd]
    L.wbgpu_eig.argtypes = [vp, pd, pd, pd]
    L.wbgpu_xk.argtypes = [vp, pd, C.c_int, pd]
    L.wbgpu_xbar.argtypes = [vp, pd, C.c_int, C.c_int, pd]
    L.wbgpu_band_traces.argtypes = [vp, pd, C.POINTER(ScanSpec), pd, pd]
    L.wbgpu_kernel_launches.argtypes = [vp]
    L.wbgpu_kernel_launches.restype = i64
    L.wbgpu_last_eig_sweeps.argtypes = [vp]
    L.wbgpu_last_eig_resolved.argtypes = [vp]
    L.wbgpu_last_eig_resolved.restype = C.c_int
    L.wbgpu_set_option.argtypes = [vp, C.c_char_p, i64]
    L.wbgpu_stage_times.argtypes = [vp, pd, C.POINTER(i64)]
    L.wbgpu_fp64_peak.argtypes = [C.c_int, C.c_int, pd]
    for name in ("wbgpu_create", "wbgpu_destroy", "wbgpu_set_R_matrix", "wbgpu_plan", "wbgpu_static_scan",
                 "wbgpu_static_scan_dev", "wbgpu_kpoints", "wbgpu_eig", "wbgpu_xk", "wbgpu_xbar", "wbgpu_band_traces",
                 "wbgpu_last_eig_sweeps", "wbgpu_set_option", "wbgpu_stage_times", "wbgpu_fp64_peak", "wbgpu_kubo_scan",
                 "wbgpu_static_scan_tetra", "wbgpu_static_scan_blocks", "wbgpu_kubo_scan_dev", "wbgpu_static_scan_tetra_blocks",
                 "wbgpu_kubo_scan_blocks"):
        getattr(L, name).restype = C.c_int
    _lib = L
    return L


EXPORTED = ["wbgpu_last_error", "wbgpu_version", "wbgpu_device_count", "wbgpu_create", "wbgpu_destroy",
            "wbgpu_set_R_matrix", "wbgpu_plan", "wbgpu_static_scan", "wbgpu_static_scan_dev", "wbgpu_spec_size",
            "wbgpu_kpoints", "wbgpu_eig", "wbgpu_xk", "wbgpu_xbar", "wbgpu_band_traces", "wbgpu_kernel_launches",
            "wbgpu_last_eig_sweeps", "wbgpu_last_eig_resolved", "wbgpu_set_option", "wbgpu_stage_times", "wbgpu_fp64_peak",
            "wbgpu_kubo_size",
            "wbgpu_kubo_scan", "wbgpu_static_scan_tetra", "wbgpu_static_scan_blocks", "wbgpu_kubo_scan_dev",
            "wbgpu_static_scan_tetra_blocks", "wbgpu_kubo_scan_blocks"]


def check(status):
    """Map a non-zero status to the Python exception the reference would raise."""
    if status == 0:
        return
    msg = lib().wbgpu_last_error().decode()
    if "not implemented" in msg:
        raise NotImplementedError(msg)
    if msg.startswith("CUDA error") or "no CUDA device" in msg:
        raise RuntimeError(msg)
    raise ValueError(msg)


def dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)
