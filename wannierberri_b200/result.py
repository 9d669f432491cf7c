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
        self.save_mode = save_mode
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
        """same keys as energyresult.py:224-239."""
        d = {"E_titles": self.E_titles, "data": self.data, "rank": self.rank,
             "transformTR": str(self.transformTR), "transformInv": str(self.transformInv), "comment": self.comment}
        for i, E in enumerate(self.Energies):
            d[f"Energies_{i}"] = E
        return d

    def save(self, name):
        np.savez_compressed(name + ".npz", **self.as_dict())

    def savedata(self, name, prefix, suffix, i_iter):
        suffix = "-" + suffix if len(suffix) > 0 else ""
        prefix = prefix + "-" if len(prefix) > 0 else ""
        self.save(prefix + name + suffix + f"_iter-{i_iter:04d}")


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
