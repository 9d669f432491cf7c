"""Smoothing of a result along its energy axis (host-side post-processing; reference: smoother.py:9-174).  A smoother is
any callable `smoother(A, axis=0)`; the reference's own smoother objects are accepted wherever this package takes one.
The two classes here give the reference's normalised, truncated convolutions as one banded matrix product."""
import numpy as np
from scipy.constants import Boltzmann, elementary_charge


class _KernelSmoother:
    """res[i] = sum_j A[j] w(E_j - E_i) / sum_j w(E_j - E_i) over the grid points within maxdE * smear of E_i
    (smoother.py:61-73: the weights are renormalised where the window is cut by the ends of the grid)."""

    def __init__(self, E, smear, maxdE=8):
        self.E = np.array(E, dtype=float)
        self.smear, self.maxdE = float(smear), maxdE
        self.dE = self.E[1] - self.E[0]
        self.NE = self.E.shape[0]
        self.NE1 = int(self.maxdE * self.smear / self.dE)
        i = np.arange(self.NE)
        off = i[None, :] - i[:, None]
        W = np.where(np.abs(off) <= self.NE1, self._kernel(off * self.dE), 0.)
        self._W = W / W.sum(axis=1, keepdims=True)

    def __call__(self, A, axis=0):
        A = np.asarray(A)
        assert A.shape[axis] == self.NE
        return np.moveaxis(np.tensordot(self._W, np.moveaxis(A, axis, 0), axes=(1, 0)), 0, axis)

    def __eq__(self, other):
        return (type(self) is type(other) and self.NE1 == other.NE1 and self.maxdE == other.maxdE and
                np.isclose(self.smear, other.smear) and self.E.shape == other.E.shape and np.allclose(self.E, other.E))

    __hash__ = None


class FermiDiracSmoother(_KernelSmoother):
    """minus the derivative of the Fermi-Dirac function at temperature `T_Kelvin` (smoother.py:76-98)"""

    def __init__(self, E, T_Kelvin, maxdE=8):
        self.T_Kelvin = T_Kelvin
        super().__init__(E, T_Kelvin * Boltzmann / elementary_charge, maxdE)

    def _kernel(self, x):
        return 0.25 / self.smear / np.cosh(x / (2 * self.smear)) ** 2


class GaussianSmoother(_KernelSmoother):
    """Gaussian of width `smear` eV (smoother.py:101-121)"""

    def _kernel(self, x):
        return np.exp(-(x / self.smear) ** 2) / self.smear / np.sqrt(np.pi)


class VoidSmoother:
    def __call__(self, A, axis=0):
        return A

    def __eq__(self, other):
        return isinstance(other, VoidSmoother)

    __hash__ = None


def get_smoother(energy, smear, mode=None):
    """smoother.py:139-174"""
    if energy is None or smear is None or smear <= 0 or len(energy) <= 1:
        return VoidSmoother()
    if mode == "Fermi-Dirac":
        return FermiDiracSmoother(energy, smear)
    if mode == "Gaussian":
        return GaussianSmoother(energy, smear)
    raise ValueError("Smoother mode not recognized.")
