"""Plug-in formula interface (hook #2): the objects a user-defined `Formula_ln` -- or an
unmodified calculator of the reference -- asks a `Data_K` for.

The built-in calculators of this package never come here: they declare scans that the CUDA kernels evaluate.  This
module exists for formulae that only exist as the user's Python code: `Data_K_R.covariant()` / `.Dcov` hand out the
objects below, whose matrices were produced on the GPU (`wbgpu_eig`, `wbgpu_xbar`: transform, eigensolver and the
U^dagger X U rotation run in the CUDA kernels); what the user's `nn()` / `ln()` does with the band blocks is the user's
numpy.  Interface of the reference: formula/formula.py:9-118 (`ln`, `nn`, `nl`, `ll`, `trace`, `additive`, `ndim`,
`transformTR`, `transformInv`), formula/elementary.py:21-53 (`Dcov`, `DEinv_ln`).
"""
import numpy as np


def reference_transforms():
    """(transform_ident, transform_odd) as the reference's `Transform` objects when the reference is importable (so that
    its `FormulaProduct` / `TransformProduct` can combine them), else the names this package's results use."""
    try:
        from wannierberri.symmetry.point_symmetry import transform_ident, transform_odd
        return transform_ident, transform_odd
    except Exception:   # the reference is not installed: this package's own names
        return "ident", "odd"


def transform_TR(name, der=0):
    """behaviour under time reversal of a Hamiltonian-gauge matrix with `der` derivatives (data_K/data_K.py:31-46)"""
    ident, odd = reference_transforms()
    if name == "Ham":
        parity = 0
    elif name in ("CC", "FF", "OO", "GG", "SS", "rotAA", "rotAAab", "CCab_antisym"):
        parity = 1
    elif name in ("D", "AA", "BB", "CCab"):
        return None
    else:
        raise ValueError(f"parity under TR unknown for {name}")
    return odd if (parity + der) % 2 else ident


def transform_Inv(name, der=0):
    """behaviour under inversion (data_K/data_K.py:13-28)"""
    ident, odd = reference_transforms()
    if name in ("Ham", "CC", "FF", "OO", "GG", "SS", "rotAA", "rotAAab", "CCab_antisym"):
        parity = 0
    elif name in ("D", "AA", "BB", "CCab"):
        return None
    else:
        raise ValueError(f"parity under inversion unknown for {name}")
    return odd if (parity + der) % 2 else ident


class Formula_ln:
    """A quantity that is covariant under gauge changes inside the `inn` and inside the `out` band sets; subclasses
    give the blocks `nn(ik, inn, out)` and `ln(ik, inn, out)` (rows in `out`, columns in `inn`)."""

    ndim = 0
    transformTR = None
    transformInv = None
    additive = True   # trace(A u B) = trace(A) + trace(B); False e.g. for the orbital moment

    def __init__(self, data_K=None, internal_terms=True, external_terms=True, **kwargs):
        self.internal_terms = internal_terms
        self.external_terms = external_terms
        if data_K is not None and getattr(data_K, "force_internal_terms_only", False):
            self.external_terms = False

    def nn(self, ik, inn, out):
        raise NotImplementedError

    def ln(self, ik, inn, out):
        raise NotImplementedError

    def nl(self, ik, inn, out):
        return self.ln(ik, out, inn)

    def ll(self, ik, inn, out):
        return self.nn(ik, out, inn)

    def trace(self, ik, inn, out):
        return np.einsum("nn...->...", self.nn(ik, inn, out)).real


class Matrix_ln(Formula_ln):
    """blocks of a stored matrix `[nk][nw][nw][3]^ndim`"""

    def __init__(self, matrix, transformTR=None, transformInv=None):
        self.matrix = matrix
        self.ndim = matrix.ndim - 3
        self.transformTR, self.transformInv = transformTR, transformInv

    def nn(self, ik, inn, out):
        return self.matrix[ik][np.ix_(inn, inn)]

    def ln(self, ik, inn, out):
        return self.matrix[ik][np.ix_(out, inn)]


class Dcov(Matrix_ln):
    """D_ln = -V_ln / (E_l - E_n) between the two band sets (elementary.py:42-49)"""

    def __init__(self, data_K):
        super().__init__(data_K.D_H)

    def nn(self, ik, inn, out):
        raise ValueError("D is defined between the inner and the outer states only")


class DEinv_ln(Matrix_ln):
    """1 / (E_m - E_n) between the two band sets (elementary.py:21-28)"""

    def __init__(self, data_K):
        super().__init__(data_K.dEig_inv)

    def nn(self, ik, inn, out):
        raise NotImplementedError("1/(E_m - E_n) is not defined inside the inner states")


class Matrix_GenDer_ln(Formula_ln):
    """generalised derivative  X^{:d} = d_d X - [D^d, X]  of a covariant matrix (formula/formula.py:95-118)"""

    def __init__(self, matrix, matrix_comader, D, transformTR=None, transformInv=None):
        self.A, self.dA, self.D = matrix, matrix_comader, D
        self.ndim = matrix.ndim + 1
        self.transformTR, self.transformInv = transformTR, transformInv

    def nn(self, ik, inn, out):
        res = np.array(self.dA.nn(ik, inn, out))
        res -= np.einsum("mld,lnb...->mnb...d", self.D.nl(ik, inn, out), self.A.ln(ik, inn, out))
        res += np.einsum("mlb...,lnd->mnb...d", self.A.nl(ik, inn, out), self.D.ln(ik, inn, out))
        return res

    def ln(self, ik, inn, out):
        res = np.array(self.dA.ln(ik, inn, out))
        res -= np.einsum("mld,lnb...->mnb...d", self.D.ln(ik, inn, out), self.A.nn(ik, inn, out))
        res += np.einsum("mlb...,lnd->mnb...d", self.A.ll(ik, inn, out), self.D.ln(ik, inn, out))
        return res


class Velocity_ln(Matrix_ln):
    """covariant derivative of the Hamiltonian: d_a H inside a band set, zero between the sets (data_K.py:273-284)"""

    def __init__(self, matrix):
        _, odd = reference_transforms()
        super().__init__(matrix, transformTR=odd, transformInv=odd)

    def ln(self, ik, inn, out):
        return np.zeros((len(out), len(inn), 3), dtype=complex)


def band_groups(E, emin, emax, degen_thresh=-1, degen_Kramers=False, sea=False, select_bands=None):
    """{(ib1, ib2): energy label} of the band groups of one k-point that touch [emin, emax] (mean energy) and, with
    `sea`, the group (0, bandmax) of everything below (label -inf): Data_K.get_bands_in_range_groups_ik
    (data_K/data_K.py:172-189, grid/tetrahedron.py:131-162)."""
    nb = len(E)
    borders = [0] + [i for i in range(1, nb) if E[i] - E[i - 1] > degen_thresh] + [nb]
    if degen_Kramers:   # groups start and end on even band indices
        borders = [i for i in borders if i % 2 == 0]
    groups = {}
    first = None
    for a, b in zip(borders, borders[1:]):
        if select_bands is not None and not set(range(a, b)) & set(select_bands):
            continue
        if E[a:b].max() >= emin and E[a:b].min() <= emax:
            groups[(a, b)] = E[a:b].mean()
            if first is None:
                first = a
    if sea:
        if select_bands is not None:
            raise NotImplementedError("Selection of bands for Fermi sea is not implemented")
        below = np.where(E < emin)[0]
        bandmax = below[-1] + 1 if len(below) else 0
        if first is not None:
            bandmax = min(bandmax, first)
        if bandmax > 0:
            groups[(0, bandmax)] = -np.inf
    return groups
