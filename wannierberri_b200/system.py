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


def as_system(obj):
    """Accept this package's `System_R` or the reference's (duck typing on the attributes read by
    the path: SURVEY.md section 2, row 9)."""
    for attr in ("rvec", "get_R_mat", "has_R_mat", "num_wann", "cell_volume"):
        if not hasattr(obj, attr):
            raise ValueError(f"system object lacks attribute '{attr}' needed by the GPU path")
    return obj
