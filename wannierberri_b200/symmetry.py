"""Point-group symmetries of reciprocal-space results: the part of the reference's `symmetry/point_symmetry.py` that
the k-grid path touches -- symmetry-reduced K-lists (`Grid.get_K_list(use_symmetry=True)`, grid/grid.py:149-189;
`PointGroup.star`, point_symmetry.py:417-424) and the symmetrisation of the summed result
(`PointGroup.symmetrize`, :337-338; `PointSymmetry.transform_tensor`, :104-119).  Host-side, tiny arrays."""
import numpy as np

SYMMETRY_PRECISION = 1e-6


class PointSymmetry:
    """point_symmetry.py:50-119: k transforms as iTR * iInv * (R @ k); `R` is kept proper (det = +1)."""

    def __init__(self, R, TR=False):
        R = np.array(R, dtype=float)
        self.TR = bool(TR)
        self.Inv = bool(np.linalg.det(R) < 0)
        self.R = R * (-1 if self.Inv else 1)
        self.iTR = -1 if self.TR else 1
        self.iInv = -1 if self.Inv else 1

    def __mul__(self, other):
        return PointSymmetry((self.R @ other.R) * (self.iInv * other.iInv), self.TR != other.TR)

    def __eq__(self, other):
        return np.linalg.norm(self.R - other.R) < 1e-12 and self.TR == other.TR and self.Inv == other.Inv

    def transform_reduced_vector(self, vec, basis):
        return vec @ (basis @ self.R.T @ np.linalg.inv(basis)) * (self.iTR * self.iInv)

    def transform_tensor(self, data, rank, transformTR, transformInv):
        """rotate the last `rank` axes by R; then the formula's behaviour under time reversal / inversion."""
        res = np.array(data, copy=True)
        dim = res.ndim
        for i in range(dim - rank, dim):
            res = np.moveaxis(np.moveaxis(res, i, -1) @ self.R.T, -1, i)
        if self.TR:
            res = apply_transform(transformTR, res)
        if self.Inv:
            res = apply_transform(transformInv, res)
        return res


def apply_transform(name, res):
    """the pre-defined transforms of point_symmetry.py:431-475, named as in calculators/*.py of this package"""
    if callable(name):
        return name(res)
    if name == "ident":
        return res
    if name == "odd":
        return -res
    if name == "trans":
        return np.swapaxes(res, -1, -2)
    if name == "odd_trans_021":   # Transform(factor=-1, transpose_axes=(0, 2, 1)) of a rank-3 tensor
        return -np.swapaxes(res, -1, -2)
    raise ValueError(f"unknown transform {name!r}")


def rotation(n, axis=(0, 0, 1)):
    """point_symmetry.py:122-143 (Rodrigues formula instead of scipy.spatial.transform)."""
    axis = np.array(axis, dtype=float)
    axis /= np.linalg.norm(axis)
    th = 2 * np.pi / n
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return PointSymmetry(np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K))


def mirror(axis=(0, 0, 1)):
    return PointSymmetry(-rotation(2, axis).R)


Identity = PointSymmetry(np.eye(3))
Inversion = PointSymmetry(-np.eye(3))
TimeReversal = PointSymmetry(np.eye(3), True)
DICT_SYM = {"Identity": Identity, "Inversion": Inversion, "TimeReversal": TimeReversal,
            "Mx": mirror([1, 0, 0]), "My": mirror([0, 1, 0]), "Mz": mirror([0, 0, 1]),
            "C2x": rotation(2, [1, 0, 0]), "C2y": rotation(2, [0, 1, 0]), "C2z": rotation(2, [0, 0, 1]),
            "C3z": rotation(3, [0, 0, 1]), "C4x": rotation(4, [1, 0, 0]), "C4y": rotation(4, [0, 1, 0]),
            "C4z": rotation(4, [0, 0, 1]), "C6z": rotation(6, [0, 0, 1])}


def from_string_prod(string):
    """point_symmetry.py:192-217: "C2x*TimeReversal" -> product of the named operations."""
    res = Identity
    for s in string.split("*")[::-1]:
        if s not in DICT_SYM:
            raise ValueError(f"The symmetry {s} is not defined")
        res = DICT_SYM[s] * res
    return res


class PointGroup:
    """point_symmetry.py:220-292, 313-338, 407-424: closure of the generators; star of a k-point; group average."""

    def __init__(self, generator_list=(), real_lattice=None):
        self.real_lattice = np.array(real_lattice, dtype=float)
        self.recip_lattice = 2 * np.pi * np.linalg.inv(self.real_lattice).T
        sym_list = [op if isinstance(op, PointSymmetry) else from_string_prod(op) for op in generator_list]
        if len(sym_list) == 0:
            sym_list = [Identity]
        while True:
            lenold = len(sym_list)
            for s1 in list(sym_list):
                for s2 in list(sym_list):
                    s3 = s1 * s2
                    if s3 not in sym_list:
                        sym_list.append(s3)
                        if len(sym_list) > 1000:
                            raise RuntimeError("Cannot define a finite group")
            if len(sym_list) == lenold:
                break
        self.symmetries = sym_list
        for basis, what in ((self.real_lattice, "real_lattice"), (self.recip_lattice, "recip_lattice")):
            if not self.check_basis_symmetry(basis):
                raise ValueError(f"{what} is not symmetric: check that the symmetries are consistent with the lattice")

    @property
    def size(self):
        return len(self.symmetries)

    def check_basis_symmetry(self, basis, tol=1e-6, rel_tol=None):
        if rel_tol is not None:
            tol = rel_tol * tol
        for sym in self.symmetries:
            rot = sym.transform_reduced_vector(np.eye(3), basis)
            if np.abs(np.round(rot) - rot).max() > tol:
                return False
        return True

    def symmetric_grid(self, nk):
        return self.check_basis_symmetry(self.recip_lattice / np.array(nk)[:, None], rel_tol=10)

    def star(self, k):
        return star(self, k)


def star(pointgroup, k):
    """distinct images of the reduced k-vector under the group (point_symmetry.py:417-424); works on this package's
    PointGroup and on the reference's (attributes `.symmetries[i].R/.iTR/.iInv`, `.recip_lattice`)."""
    basis = np.asarray(pointgroup.recip_lattice)
    binv = np.linalg.inv(basis)
    st = [np.asarray(k) @ (basis @ S.R.T @ binv) * (S.iTR * S.iInv) for S in pointgroup.symmetries]
    for i in range(len(st) - 1, 0, -1):
        diff = np.array(st[:i]) - np.array(st[i])[None, :]
        if np.linalg.norm(diff - diff.round(), axis=-1).min() < SYMMETRY_PRECISION:
            del st[i]
    return np.array(st)


def symmetrize_tensor(pointgroup, data, rank, transformTR, transformInv):
    """PointGroup.symmetrize on one array (point_symmetry.py:337-338, 407-415): mean over the group."""
    total = 0
    for S in pointgroup.symmetries:
        sym = S if isinstance(S, PointSymmetry) else PointSymmetry(np.asarray(S.R) * (-1 if S.Inv else 1), S.TR)
        total = total + sym.transform_tensor(data, rank, transformTR, transformInv)
    return total / len(pointgroup.symmetries)
