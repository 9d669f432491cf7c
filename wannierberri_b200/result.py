"""Result containers with the interface `run()` hands back (reference: result/energyresult.py:11-278,
result/resultdict.py:18-66).  Only what the static hot path produces: an array over Fermi levels."""
import numpy as np


class EnergyResult:

    def __init__(self, Energies, data, transformTR=None, transformInv=None, rank=None, E_titles=("Efermi",),
                 comment="undocumented", save_mode="bin+txt", smoothers=(None,)):
        if not isinstance(Energies, (list, tuple)):
            Energies = [Energies]
        self.Energies = [np.asarray(E) for E in Energies]
        self.data = np.asarray(data)
        self.rank = self.data.ndim - len(self.Energies) if rank is None else rank
        self.transformTR, self.transformInv = transformTR, transformInv
        self.E_titles = list(E_titles)
        self.comment = comment
        # result/result.py:7-11: the set of "bin" / "txt" named in the string (or an already parsed set)
        self.save_mode = {m for m in ("bin", "txt") if m in save_mode}
        self.smoothers = list(smoothers)

    def __mul__(self, number):
        return EnergyResult(self.Energies, self.data * number, self.transformTR, self.transformInv, self.rank,
                            self.E_titles, self.comment, self.save_mode, self.smoothers)

    __rmul__ = __mul__

    def __truediv__(self, number):
        return self * (1. / number)

    def __add__(self, other):
        if other is None or other == 0:
            return self
        for a, b in zip(self.Energies, other.Energies):
            if not np.array_equal(a, b):
                raise RuntimeError("Adding results with different energies")
        for a, b in zip(self.smoothers, other.smoothers):   # energyresult.py:176-180
            if not ((a is None and b is None) or a == b):
                raise RuntimeError("Adding results with different smoothers")
        return EnergyResult(self.Energies, self.data + other.data, self.transformTR, self.transformInv, self.rank,
                            self.E_titles, self.comment, self.save_mode, self.smoothers)

    __radd__ = __add__

    def __sub__(self, other):
        return self + other * (-1)

    def mul_array(self, other, axes=None):
        """energyresult.py: multiply by an array along the energy axes."""
        other = np.asarray(other)
        shape = other.shape + (1,) * (self.data.ndim - other.ndim)
        return EnergyResult(self.Energies, self.data * other.reshape(shape), self.transformTR, self.transformInv,
                            self.rank, self.E_titles, self.comment, self.save_mode, self.smoothers)

    def transform(self, sym):
        """energyresult.py:266-278: the result seen after the point-group operation `sym`."""
        return EnergyResult(self.Energies, sym.transform_tensor(self.data, self.rank, self.transformTR, self.transformInv),
                            self.transformTR, self.transformInv, self.rank, self.E_titles, self.comment, self.save_mode,
                            self.smoothers)

    def symmetrized(self, pointgroup):
        """PointGroup.symmetrize(result) (point_symmetry.py:337-338): mean over the group."""
        from .symmetry import symmetrize_tensor
        return EnergyResult(self.Energies,
                            symmetrize_tensor(pointgroup, self.data, self.rank, self.transformTR, self.transformInv),
                            self.transformTR, self.transformInv, self.rank, self.E_titles, self.comment, self.save_mode,
                            self.smoothers)

    @property
    def dataSmooth(self):
        """energyresult.py:120-131: every energy axis smoothed by its smoother (None = as is)"""
        d = self.data
        for i, sm in enumerate(self.smoothers):
            if sm is not None:
                d = sm(d, axis=i)
        return d

    @property
    def max(self):
        """[max |data|, norm, norm of the first difference along the energy axis] (energyresult.py:245-264):
        the numbers by which adaptive refinement selects K-points."""
        d = self.dataSmooth
        return np.array([np.abs(d).max(), np.linalg.norm(d), np.linalg.norm(d[1:] - d[:-1])])

    def as_dict(self):
        """same keys as energyresult.py:224-239; the transforms as the dictionaries of `Transform.as_dict`
        (point_symmetry.py:459-460), so that the reference's `EnergyResult.from_npz` restores them."""
        d = {"E_titles": self.E_titles, "data": self.data, "rank": self.rank,
             "transformTR": transform_as_dict(self.transformTR), "transformInv": transform_as_dict(self.transformInv),
             "comment": self.comment}
        for i, E in enumerate(self.Energies):
            d[f"Energies_{i}"] = E
        return d

    def save(self, name):   # result/result.py:39-46
        with open(name + ".npz", "wb") as f:
            np.savez_compressed(f, **self.as_dict())

    def _write(self, data, datasm, i):   # energyresult.py:203-214
        if i == len(self.Energies):
            flat = list(data.reshape(-1)) + list(datasm.reshape(-1))
            if np.iscomplexobj(data):
                return ["    " + "    ".join(f"{x.real:15.6e} {x.imag:15.6e}" for x in flat)]
            return ["    " + "    ".join(f"{x:15.6e}" for x in flat)]
        return [f"{E:15.6e}    {line:s}" for j, E in enumerate(self.Energies[i]) for line in self._write(data[j], datasm[j], i + 1)]

    def savetxt(self, name):
        """the reference's text table (energyresult.py:216-224): energies, the data and the smoothed data per row"""
        frmt = "{0:^31s}" if np.iscomplexobj(self.data) else "{0:^15s}"
        head = "".join("#### " + line + "\n" for line in self.comment.split("\n"))
        head += "#" + "    ".join(f"{t:^15s}" for t in self.E_titles) + " " * 8 + "    ".join(
            frmt.format(b) for b in _get_head(self.rank) * 2) + "\n"
        with open(name, "w") as f:
            f.write(head + "\n".join(self._write(self.data, self.dataSmooth, 0)))

    def savedata(self, name, prefix, suffix, i_iter):   # energyresult.py:241-248
        suffix = "-" + suffix if len(suffix) > 0 else ""
        prefix = prefix + "-" if len(prefix) > 0 else ""
        filename = prefix + name + suffix + f"_iter-{i_iter:04d}"
        if "bin" in self.save_mode:
            self.save(filename)
        if "txt" in self.save_mode:
            self.savetxt(filename + ".dat")

    @classmethod
    def from_npz(cls, file_npz):
        """energyresult.py:83-112"""
        res = np.load(file_npz, allow_pickle=True)
        Energies = [res[f"Energies_{i}"] for i, _ in enumerate(res["E_titles"])]

        def tr(key):
            if key not in res.files or res[key] is None:
                return None
            return transform_from_dict(res[key].item())
        return cls(Energies, res["data"], transformTR=tr("transformTR"), transformInv=tr("transformInv"),
                   rank=int(res["rank"]), E_titles=list(res["E_titles"]),
                   comment=str(res["comment"]) if "comment" in res.files else "undocumented")


def _get_head(n):   # utility.py:96-100
    return ["  "] if n <= 0 else [a + b for a in "xyz" for b in _get_head(n - 1)]


# the pre-defined transforms of point_symmetry.py:504-509 under the names used in calculators/*.py of this package
_TRANSFORMS = {"ident": dict(factor=1, conj=False, transpose_axes=None, swap_axes=None),
               "odd": dict(factor=-1, conj=False, transpose_axes=None, swap_axes=None),
               "trans": dict(factor=1, conj=False, transpose_axes=(1, 0), swap_axes=None),
               "odd_trans_021": dict(factor=-1, conj=False, transpose_axes=(0, 2, 1), swap_axes=None)}


def transform_as_dict(t):
    """`Transform.as_dict()` (point_symmetry.py:459-460) of a transform given by name, or of the reference's object."""
    if t is None:
        return None
    if hasattr(t, "as_dict"):
        return t.as_dict()
    return dict(_TRANSFORMS[t])


def transform_from_dict(d):
    """inverse of `transform_as_dict` for the pre-defined transforms (point_symmetry.py:512-527)"""
    if d is None or isinstance(d, str):
        return None
    key = dict(factor=int(d.get("factor", 1)), conj=bool(d.get("conj", False)),
               transpose_axes=None if d.get("transpose_axes") is None else tuple(d["transpose_axes"]),
               swap_axes=None if d.get("swap_axes") is None else tuple(d["swap_axes"]))
    for name, val in _TRANSFORMS.items():
        if val == key:
            return name
    raise ValueError(f"transform {d} is not one of the pre-defined ones")


class ResultDict:

    def __init__(self, results):
        self.results = results

    def __mul__(self, number):
        return ResultDict({k: v * number for k, v in self.results.items()})

    def __add__(self, other):
        if other is None or other == 0:
            return self
        return ResultDict({k: self.results[k] + other.results[k] for k in self.results if k in other.results})

    __radd__ = __add__

    def savedata(self, prefix, suffix, i_iter):
        for k, v in self.results.items():
            v.savedata(k, prefix, suffix, i_iter)

    @property
    def max(self):
        return np.array([x for v in self.results.values() for x in v.max])
