"""k-grid: NK = NKdiv x NKFFT, the K-block list and the FFT sub-grid (bit-exact restatement of
grid/grid.py:68-75,149-172,196-266 and grid/Kpoint.py:75-77 of the reference; symmetry-reduced
K-lists follow grid/grid.py:169-189 with the star of symmetry.py)."""
import warnings

import pickle
import numpy as np


def _one2three(x):
    if x is None:
        return None
    if isinstance(x, (int, np.integer)):
        return np.array([x, x, x], dtype=int)
    x = np.array(x, dtype=int)
    assert x.shape == (3,)
    return x


def _iterate_vector(v1, v2):   # grid/grid.py:192-193
    return ((x, y, z) for x in range(v1[0], v2[0]) for y in range(v1[1], v2[1]) for z in range(v1[2], v2[2]))


class KpointBZ:
    """One K-block: `K` (reduced coordinates of the K-grid), shift of the FFT sub-grid `Kp_fullBZ`,
    weight `factor` (grid/Kpoint.py:20-38,75-77)."""

    def __init__(self, K, dK, NKFFT, factor, refinement_level=0):
        self.K = np.copy(K)
        self.dK = np.copy(dK)
        self.NKFFT = np.copy(NKFFT)
        self.factor = factor
        self.refinement_level = refinement_level

    @property
    def Kp_fullBZ(self):
        return self.K / self.NKFFT


class Grid:

    def __init__(self, system=None, NK=None, NKFFT=None, NKdiv=None, length=None, length_FFT=None, use_symmetry=True):
        """grid/grid.py:129-139 + determineNK / autoNK (:196-266): every given grid must respect the point group of the
        system (`PointGroup.symmetric_grid`), the automatic FFT grid is searched among the symmetric ones only, and
        non-periodic directions get one k-point."""
        NKdiv, NKFFT, NK = _one2three(NKdiv), _one2three(NKFFT), _one2three(NK)
        # GridAbstract.__init__ (grid.py:44-57): without `use_symmetry` the grid carries the trivial point group
        pointgroup = getattr(system, "pointgroup", None) if use_symmetry else None
        recip = 2 * np.pi * np.linalg.inv(system.real_lattice).T

        def symmetric(nk):
            return pointgroup is None or bool(pointgroup.symmetric_grid(nk))

        if length is not None:
            if NK is None:
                NK = np.array(np.round(length / (2 * np.pi) * np.linalg.norm(recip, axis=1)), dtype=int)
            else:
                warnings.warn("length is disregarded in presence of NK")
        if length_FFT is not None:
            if NKFFT is None:
                NKFFT = np.array(np.round(length_FFT / (2 * np.pi) * np.linalg.norm(recip, axis=1)), dtype=int)
            else:
                warnings.warn("length_FFT is disregarded in presence of NKFFT")
        for name, nk in (("NKdiv", NKdiv), ("NK", NK), ("NKFFT", NKFFT)):
            if nk is not None and not symmetric(nk):
                raise AssertionError(f" {name}={nk} is not consistent with the given symmetry ")
        if (NKdiv is not None) and (NKFFT is not None):
            if NK is not None:
                warnings.warn("NK is disregarded in presence of NKdiv,NKFFT")
                NK = None
        elif NK is not None:
            if NKdiv is not None:
                warnings.warn("NKdiv is disregarded in presence of NK or length")
            if NKFFT is None:
                # autoNK (grid.py:196-212): the smallest symmetric FFT grid in [rec, 3 rec), then among the symmetric
                # grids in [min, 2 min) the one whose multiple is closest to NK
                rec = np.array(system.NKFFT_recommended)
                sym1 = np.array([f for f in _iterate_vector(rec, rec * 3) if symmetric(f)])
                fmin = sym1[np.argmin(sym1.prod(axis=1))]
                cands = np.array([f for f in _iterate_vector(fmin, fmin * 2) if symmetric(f)])
                div = np.array(np.round(NK[None, :] / cands), dtype=int)
                div[div <= 0] = 1
                change = div * cands / NK[None, :]
                change[change > 1] = 1. / change[change > 1]
                NKFFT = cands[np.argmax(change.min(axis=1))]
            NKdiv = np.array(np.round(NK / NKFFT), dtype=int)
            NKdiv[NKdiv <= 0] = 1
        else:
            raise ValueError("you need to specify either NK or a pair (NKdiv,NKFFT) or (NK,NKFFT)."
                             f"found NK={NK}, NKdiv={NKdiv}, NKFFT={NKFFT} ")
        if NK is not None and not np.all(NK == NKFFT * NKdiv):
            warnings.warn(f" the requested k-grid {NK} was adjusted to {NKFFT * NKdiv}. ")
        notperiodic = np.logical_not(np.array(getattr(system, "periodic", (True, True, True)), dtype=bool))
        NKdiv, NKFFT = np.array(NKdiv), np.array(NKFFT)
        NKdiv[notperiodic] = 1
        NKFFT[notperiodic] = 1
        self.div = NKdiv
        self.FFT = NKFFT
        self.pointgroup = pointgroup

    @property
    def dense(self):
        return self.div * self.FFT

    @property
    def points_FFT(self):
        dkx, dky, dkz = 1. / self.FFT
        ix, iy, iz = np.meshgrid(np.arange(self.FFT[0]), np.arange(self.FFT[1]), np.arange(self.FFT[2]), indexing="ij")
        return np.stack([ix.ravel() * dkx, iy.ravel() * dky, iz.ravel() * dkz], axis=1)

    def K_arrays(self, use_symmetry=False):
        """(Kp_fullBZ[nK,3], factor[nK]) of the K-list, x-major / z-fastest, same floating point operations as the
        reference: K = [x,y,z] * (1./div); Kp_fullBZ = K / FFT.  With `use_symmetry` only one K-point of every star
        of the point group is kept and it absorbs the weights of the others, in the reference's order
        (grid/grid.py:169-189: the first point met in the z-outer / x-inner sweep survives)."""
        dK = 1. / self.div
        factor = 1. / np.prod(self.div)
        x, y, z = np.meshgrid(np.arange(self.div[0]), np.arange(self.div[1]), np.arange(self.div[2]), indexing="ij")
        K = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1) * dK[None, :]
        factors = np.full(K.shape[0], factor)
        if use_symmetry and self.pointgroup is not None:
            from .symmetry import star
            div = np.array(self.div)
            alive = np.ones(tuple(div), dtype=bool)
            fac = np.full(tuple(div), factor)
            for iz in range(div[2]):
                for iy in range(div[1]):
                    for ix in range(div[0]):
                        if not alive[ix, iy, iz]:
                            continue
                        st = np.array(np.round(star(self.pointgroup, np.array([ix, iy, iz]) * dK) * div), dtype=int) % div
                        for k in map(tuple, st):
                            if k != (ix, iy, iz) and alive[k]:
                                fac[ix, iy, iz] += fac[k]
                                alive[k] = False
            keep = alive.ravel()
            K, factors = K[keep], fac.ravel()[keep]
        return K / self.FFT[None, :], factors

    def get_K_list(self, use_symmetry=True, k_batch=None):
        dK = 1. / self.div
        shifts, factors = self.K_arrays(use_symmetry=use_symmetry)
        return [KpointBZ(K=s * self.FFT, dK=dK, NKFFT=self.FFT, factor=f) for s, f in zip(shifts, factors)]


class KpointBZparallel:
    """A K-point of the refinement loop with its cell of size dK (reference: grid/Kpoint.py:9-100, 104-184).
    `K` is in units of 1/NKFFT of the reciprocal cell: `Kp_fullBZ = K / NKFFT` is the shift of the FFT sub-grid."""

    def __init__(self, K, dK, NKFFT, factor, refinement_level=0, pointgroup=None):
        self.K = np.array(K, dtype=float)
        self.dK = np.array(dK, dtype=float)
        self.NKFFT = np.array(NKFFT)
        self.factor = factor
        self.refinement_level = refinement_level
        self.pointgroup = pointgroup
        self.result = None
        self.result_storage_path = None
        self.res_dumped_flag = False
        self._max = None
        self._star = None
        self._distGamma = None

    @property
    def Kp_fullBZ(self):
        return self.K / self.NKFFT

    @property
    def dK_fullBZ(self):
        return self.dK / self.NKFFT

    @property
    def was_evaluated_flag(self):
        return self.result is not None or self.res_dumped_flag

    def set_result(self, res):
        self.result = res
        self._max = res.max

    def set_storage_path(self, path):  # grid/Kpoint.py:31-32
        self.result_storage_path = path

    def dump_result(self):
        """`dump_results=True` of run(): the result of this K-point goes to its own pickle file and leaves the
        memory (grid/Kpoint.py:56-62); `get_result` reads it back when a later iteration re-weights the point."""
        if self.res_dumped_flag:
            return
        with open(self.result_storage_path, "wb") as f:
            pickle.dump(self.result, f)
        self.result = None
        self.res_dumped_flag = True

    def get_result(self):  # grid/Kpoint.py:40-50
        if self.result is not None:
            return self.result
        if self.res_dumped_flag:
            with open(self.result_storage_path, "rb") as f:
                return pickle.load(f)
        raise RuntimeError("result for a K-point is called, which was never evaluated")

    @property
    def max(self):
        return self._max * self.factor

    @property
    def star(self):  # grid/Kpoint.py:115-120
        if self._star is None:
            if self.pointgroup is None:
                self._star = np.array([self.K])
            else:
                from .symmetry import star
                self._star = star(self.pointgroup, self.K)
        return self._star

    @property
    def distGamma(self):  # grid/Kpoint.py:178-182
        if self._distGamma is None:
            sh = np.arange(-3, 4)
            corners = np.array([[x, y, z] for x in sh for y in sh for z in sh])
            self._distGamma = np.linalg.norm(((self.K % 1)[None, :] - corners).dot(self.pointgroup.recip_lattice), axis=1).min()
        return self._distGamma

    def equiv(self, other):  # grid/Kpoint.py:137-144
        from .symmetry import SYMMETRY_PRECISION
        if self.refinement_level != other.refinement_level:
            return False
        dif = self.star[:, None, :] - other.star[None, :, :]
        return bool(np.linalg.norm((dif - np.round(dif)), axis=2).min() < SYMMETRY_PRECISION)

    def absorb(self, other):  # grid/Kpoint.py:125-135
        if other is None:
            return
        if other.was_evaluated_flag:
            if self.was_evaluated_flag:
                raise RuntimeError("combining two K-points with calculated result should not happen")
            self.set_result(other.get_result())
        self.factor += other.factor

    def divide(self, ndiv, periodic=(True, True, True), use_symmetry=False):
        """grid/Kpoint.py:146-176: ndiv[0] x ndiv[1] x ndiv[2] children tiling this point's cell; this point dies."""
        ndiv = np.array(ndiv)
        ndiv[np.logical_not(np.array(periodic, dtype=bool))] = 1
        dK_adpt = self.dK / ndiv
        adpt_shift = (-self.dK + dK_adpt) / 2.
        newfac = self.factor / np.prod(ndiv)
        children = [KpointBZparallel(self.K + adpt_shift + dK_adpt * np.array([x, y, z]), dK_adpt, self.NKFFT, newfac,
                                     self.refinement_level + 1, self.pointgroup)
                    for x in range(ndiv[0]) for y in range(ndiv[1]) for z in range(ndiv[2])]
        self.factor = 0
        if use_symmetry and self.pointgroup is not None:
            exclude_equiv_points(children)
        return children


def exclude_equiv_points(K_list, new_points=None):
    """grid/Kpoint.py:185-216: among K-points at the same distance from Gamma, a later point that is symmetry-
    equivalent to an earlier one is absorbed by it (only pairs involving a new point are examined)."""
    n = len(K_list)
    if new_points is None:
        new_points = n
    length = np.array([K.distGamma for K in K_list])
    order = np.argsort(length)
    length = length[order]
    wall = [0] + list(np.where(length[1:] - length[:-1] > 1e-4)[0] + 1) + [len(K_list)]
    exclude = []
    for start, end in zip(wall[:-1], wall[1:]):
        for l in range(start, end):
            i = order[l]
            if i not in exclude:
                for m in range(start, end):
                    j = order[m]
                    if i >= j:
                        continue
                    if i < n - new_points and j < n - new_points:
                        continue
                    if j not in exclude:
                        if K_list[i].equiv(K_list[j]):
                            exclude.append(j)
                            K_list[i].absorb(K_list[j])
    for i in sorted(exclude)[-1::-1]:
        del K_list[i]
