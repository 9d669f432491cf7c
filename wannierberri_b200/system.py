"""Input container mirroring the part of `System_R` that the hot path reads
(reference: system/system_R.py:83,106-175; system/system.py:173-181; fourier/rvectors.py:351-369).

Model construction (Wannier90 / tb.dat readers, symmetrisation, ...) is out of scope: build the
system with the reference and hand it to `run()` -- any object with `.rvec.iRvec`,
`.rvec.cRvec_shifted`, `.get_R_mat(key)`, `.has_R_mat(key)`, `.num_wann`, `.cell_volume` is accepted
(see `as_system`) -- or load the arrays with `System_R.from_arrays / from_npz`."""
import numpy as np


class _Rvec:
    """The two attributes of `Rvectors` that the path needs."""

    def __init__(self, iRvec, cRvec_shifted):
        self.iRvec = np.ascontiguousarray(iRvec, dtype=np.int32)
        self.cRvec_shifted = np.ascontiguousarray(cRvec_shifted, dtype=np.float64)
        self.nRvec = self.iRvec.shape[0]


class System_R:

    def __init__(self, real_lattice, iRvec, wannier_centers_cart, force_internal_terms_only=False, periodic=(True,) * 3):
        self.real_lattice = np.array(real_lattice, dtype=float)
        self.wannier_centers_cart = np.array(wannier_centers_cart, dtype=float)
        self.num_wann = self.wannier_centers_cart.shape[0]
        self.periodic = np.array(periodic, dtype=bool)
        self.force_internal_terms_only = force_internal_terms_only
        self.is_phonon = False
        iRvec = np.array(iRvec, dtype=int)
        t = self.wannier_centers_cart
        cR = iRvec.dot(self.real_lattice)
        # R + t_j - t_i   (rvectors.py:355-369)
        self.rvec = _Rvec(iRvec, cR[:, None, None, :] + (t[None, :, :] - t[:, None, :])[None])
        self._XX_R = {}
        self.pointgroup = None

    # --- system_R.py:106-175
    def set_R_mat(self, key, value, reset=False):
        value = np.ascontiguousarray(value, dtype=np.complex128)
        if value.shape[:3] != (self.rvec.nRvec, self.num_wann, self.num_wann):
            raise ValueError(f"R-matrix {key} has shape {value.shape}, expected "
                             f"({self.rvec.nRvec},{self.num_wann},{self.num_wann},...)")
        if key in self._XX_R and not reset:
            raise RuntimeError(f"setting {key} for the second time without explicit permission. smth is wrong")
        self._XX_R[key] = value

    def get_R_mat(self, key):
        try:
            return self._XX_R[key]
        except KeyError:
            raise ValueError(f"The real-space matrix elements '{key}' are not set in the system")

    def set_pointgroup(self, symmetry_gen=(), spacegroup=None):
        """system.py:264-300 (generators only): symmetries of the system, used for symmetry-reduced K-lists and for
        the symmetrisation of results.  Operations by name ("C4z", "C2x*TimeReversal", "Inversion", ...) or as
        `wannierberri_b200.symmetry.PointSymmetry`."""
        if spacegroup is not None:
            raise NotImplementedError("space groups from irrep are not available on the GPU path; give the generators")
        from .symmetry import PointGroup
        self.pointgroup = PointGroup(symmetry_gen, real_lattice=self.real_lattice)

    def has_R_mat(self, key):
        return key in self._XX_R

    @property
    def cell_volume(self):
        return abs(np.linalg.det(self.real_lattice))

    @property
    def NKFFT_recommended(self):
        """system_R.py:591-597: 1 + 2 max|R_i|."""
        return 2 * np.abs(self.rvec.iRvec).max(axis=0) + 1

    @classmethod
    def from_npz(cls, path, pointgroup=None):
        """Load the compact fixture format written by tests/golden/make_golden.py."""
        f = np.load(path)
        s = cls(f["real_lattice"], f["iRvec"], f["wannier_centers_cart"])
        if pointgroup is not None:
            s.set_pointgroup(pointgroup)
        for k in f.files:
            if k.startswith("XX_R_"):
                s.set_R_mat(k[5:], f[k])
        return s


def synthetic_system(num_wann, rmax=2, seed=20261017, lattice_const=4.0, matrices=("Ham", "AA"), degenerate_pairs=False):
    """Seeded random tight-binding model: all R with every
    component in -rmax..rmax on a cubic lattice, X_R = (G + iG') exp(-|R|) with X(-R) = X(R)^dagger
    enforced, AA scaled by 0.1 with zero on-site diagonal, centres uniform in the cell.
    `degenerate_pairs` doubles every level exactly (H = 1_2 (x) H_half) -- a stress test for the
    eigensolver and the degenerate-group logic."""
    rng = np.random.default_rng(seed)
    rr = np.arange(-rmax, rmax + 1)
    iRvec = np.array([[x, y, z] for x in rr for y in rr for z in rr], dtype=int)
    nR = len(iRvec)
    lattice = np.eye(3) * lattice_const
    nh = num_wann // 2 if degenerate_pairs else num_wann
    centres = rng.random((nh, 3)) * lattice_const
    if degenerate_pairs:
        assert num_wann % 2 == 0
        centres = np.concatenate([centres, centres])
    index = {tuple(R): i for i, R in enumerate(iRvec)}
    decay = np.exp(-np.linalg.norm(iRvec, axis=1))

    def herm_field(ncart):
        shape = (nR, nh, nh) + ((3,) if ncart == 3 else ())
        X = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) * decay.reshape((nR,) + (1,) * (len(shape) - 1))
        Xd = np.empty_like(X)
        for R, i in index.items():
            Xd[i] = X[index[tuple(-np.array(R))]].swapaxes(0, 1).conj()
        X = 0.5 * (X + Xd)
        if degenerate_pairs:
            big = np.zeros((nR, num_wann, num_wann) + shape[3:], dtype=complex)
            big[:, :nh, :nh] = X
            big[:, nh:, nh:] = X
            X = big
        return X

    s = System_R(lattice, iRvec, centres)
    for key in matrices:
        X = herm_field(1 if key == "Ham" else 3)
        if key != "Ham":
            X *= 0.1
            i0 = index[(0, 0, 0)]
            for n in range(num_wann):
                X[i0, n, n] = 0
        s.set_R_mat(key, X)
    return s


class SystemSOC(System_R):
    """Container with the attributes of the reference's `SystemSOC` that the K-block path reads (system/system_soc.py:
    14-70, 232-270): scalar `system_up` / `system_down` (the same object for nspin = 1), the spin-orbit term `Ham_SOC`
    and the spin matrix `SS` of the combined system on R-vectors of their own.  Wannier function 2i is the up copy of
    scalar function i, 2i + 1 the down copy."""

    def __init__(self, system_up, system_down=None, iRvec=None, Ham_SOC=None, SS=None):
        self.system_up = system_up
        self.system_down = system_up if system_down is None else system_down
        self.nspin = 1 if system_down is None else 2
        self.num_wann_scalar = system_up.num_wann
        centres = np.zeros((2 * self.num_wann_scalar, 3))
        centres[::2] = self.system_up.wannier_centers_cart
        centres[1::2] = self.system_down.wannier_centers_cart
        if iRvec is None:
            iRvec = np.zeros((1, 3), dtype=int)
        super().__init__(system_up.real_lattice, iRvec, centres,
                         force_internal_terms_only=(system_up.force_internal_terms_only or self.system_down.force_internal_terms_only),
                         periodic=system_up.periodic)
        self.has_soc = Ham_SOC is not None
        if Ham_SOC is not None:
            self.set_R_mat("Ham_SOC", Ham_SOC)
        if SS is not None:
            self.set_R_mat("SS", SS)


def is_soc_system(obj):
    """a system with non-self-consistent spin-orbit coupling on top of scalar up / down systems (system/system_soc.py)"""
    return all(hasattr(obj, a) for a in ("system_up", "system_down", "num_wann_scalar", "nspin"))


def flatten_soc(soc):
    """The System_R that is equivalent to a SOC system, for the K-block path of `Data_K_soc` (data_K/data_K_soc.py:7-62).

    Data_K_soc builds, per K-block, H(k) and every other matrix from three R-space sets: the scalar up system on the
    even Wannier functions, the scalar down system on the odd ones, and the spin-orbit term `Ham_SOC` (plus the spin
    matrix `SS`) of the combined system.  All of that is linear in R-space and the derivative factors R + t_j - t_i of
    the three sets agree (the combined Wannier centres interleave the up and down ones), so the K-block arithmetic is
    that of ONE system on the union of the three R-vector sets: Ham = up (+) down + Ham_SOC, AA / BB / CC / ... = up (+)
    down, SS = the combined system's.  Done once on the host; the GPU path then runs unchanged."""
    up, down = soc.system_up, soc.system_down
    parts = [soc.rvec, up.rvec, down.rvec] if getattr(soc, "rvec", None) is not None else [up.rvec, down.rvec]
    merged = sorted({tuple(int(x) for x in R) for rv in parts for R in rv.iRvec})
    index = {R: i for i, R in enumerate(merged)}
    maps = [np.array([index[tuple(int(x) for x in R)] for R in rv.iRvec]) for rv in parts]
    m_up, m_down = maps[-2], maps[-1]
    nw = int(soc.num_wann)
    flat = System_R(soc.real_lattice, np.array(merged, dtype=int), soc.wannier_centers_cart,
                    force_internal_terms_only=getattr(soc, "force_internal_terms_only", False),
                    periodic=getattr(soc, "periodic", (True,) * 3))
    flat.pointgroup = getattr(soc, "pointgroup", None)
    up_keys = [k for k in ("Ham", "AA", "BB", "CC") if up.has_R_mat(k) and down.has_R_mat(k)]
    for key in up_keys:
        Xu, Xd = np.asarray(up.get_R_mat(key)), np.asarray(down.get_R_mat(key))
        X = np.zeros((len(merged), nw, nw) + Xu.shape[3:], dtype=complex)
        X[m_up, ::2, ::2] += Xu
        X[m_down, 1::2, 1::2] += Xd
        if key == "Ham" and getattr(soc, "has_soc", False):
            X[maps[0]] += np.asarray(soc.get_R_mat("Ham_SOC"))
        flat.set_R_mat(key, X)
    if getattr(soc, "rvec", None) is not None and soc.has_R_mat("SS"):
        S = np.zeros((len(merged), nw, nw, 3), dtype=complex)
        S[maps[0]] = np.asarray(soc.get_R_mat("SS"))
        flat.set_R_mat("SS", S)
    return flat


_FLAT_SOC = {}


def kramers_system(num_wann, rmax=1, seed=20261018, lattice_const=4.0):
    """Seeded random model with PT symmetry: H(k) = [[A(k), B(k)], [-B(k)^*, A(k)^*]] with A hermitian and B antisymmetric, so
    every level of every k-point is EXACTLY doubly degenerate although the two halves of the basis are coupled (unlike
    `synthetic_system(degenerate_pairs=True)`, which is block diagonal) -- the hard case for an eigensolver that builds
    eigenvectors one eigenvalue at a time."""
    assert num_wann % 2 == 0
    nh = num_wann // 2
    rng = np.random.default_rng(seed)
    rr = np.arange(-rmax, rmax + 1)
    iRvec = np.array([[x, y, z] for x in rr for y in rr for z in rr], dtype=int)
    index = {tuple(R): i for i, R in enumerate(iRvec)}
    decay = np.exp(-np.linalg.norm(iRvec, axis=1))[:, None, None]
    A = (rng.standard_normal((len(iRvec), nh, nh)) + 1j * rng.standard_normal((len(iRvec), nh, nh))) * decay
    B = (rng.standard_normal((len(iRvec), nh, nh)) + 1j * rng.standard_normal((len(iRvec), nh, nh))) * decay
    B = B - B.swapaxes(1, 2)                                   # antisymmetric for every R
    minus = np.array([index[tuple(-R)] for R in iRvec])
    A = 0.5 * (A + A[minus].swapaxes(1, 2).conj())              # A(-R) = A(R)^dagger
    H = np.zeros((len(iRvec), num_wann, num_wann), dtype=complex)
    H[:, :nh, :nh] = A
    H[:, :nh, nh:] = B
    H[:, nh:, :nh] = -B[minus].conj()
    H[:, nh:, nh:] = A.swapaxes(1, 2)                          # A(-R)^* = A(R)^T
    centres = rng.random((nh, 3)) * lattice_const
    s = System_R(np.eye(3) * lattice_const, iRvec, np.concatenate([centres, centres]))
    s.set_R_mat("Ham", H)
    return s


def as_system(obj):
    """Accept this package's `System_R` or the reference's (duck typing on the attributes read by
    the path: SURVEY.md section 2, row 9); a SOC system (`SystemSOC`: scalar up / down systems + spin-orbit term) is
    flattened once into the equivalent System_R (`flatten_soc`)."""
    if is_soc_system(obj):
        key = id(obj)
        if key not in _FLAT_SOC or _FLAT_SOC[key][0] is not obj:
            _FLAT_SOC.clear()   # one at a time: the flattened copy holds full R-space matrices
            _FLAT_SOC[key] = (obj, flatten_soc(obj))
        return _FLAT_SOC[key][1]
    for attr in ("rvec", "get_R_mat", "has_R_mat", "num_wann", "cell_volume"):
        if not hasattr(obj, attr):
            raise ValueError(f"system object lacks attribute '{attr}' needed by the GPU path")
    return obj
