"""Static (Fermi-level scan) calculators with the constructor interface of the reference
(calculators/static.py:26-58 and calculator.py:20): same class names, same keyword arguments, same
post-operations.  A calculator here does no arithmetic on k-points itself: it declares the scans
(`specs()`) that the GPU engine must evaluate and combines their outputs (`combine()`), exactly
mirroring what `__call__` of the reference class composes."""
from copy import copy

import numpy as np

from .. import factors
from .. import _lib
from .._lib import ScanSpec
from ..result import EnergyResult

_TR_INV = {  # transformTR, transformInv of the formulae (covariant.py)
    _lib.IDENTITY: ("ident", "ident"), _lib.OMEGA: ("odd", "ident"), _lib.MORB_HPM: ("odd", "ident"),
    _lib.SPIN: ("odd", "ident"), _lib.VEL_OMEGA: ("ident", "odd"), _lib.VEL_HPLUS: ("ident", "odd"),
    _lib.VEL_SPIN: ("ident", "odd"), _lib.VEL_VEL: ("ident", "ident"),
    _lib.INV_MASS: ("ident", "ident"), _lib.DER_OMEGA: ("ident", "odd"),
    _lib.SHC_RYOO: ("ident", "ident"), _lib.SHC_QIAO: ("ident", "ident"), _lib.SHC_SIMPLE: ("ident", "ident"),
    # products: TransformProduct of the factors (V: odd, odd; InvMass: ident, ident; Omega, Spin: odd, ident)
    _lib.DER_SPIN: ("ident", "odd"), _lib.VEL_VEL_VEL: ("odd", "odd"), _lib.MASS_VEL: ("odd", "odd"),
    _lib.MASS_MASS: ("ident", "ident"), _lib.VEL_MASS_VEL: ("ident", "ident"), _lib.OMEGA_S: ("ident", "ident"),
    _lib.OMEGA_OMEGA: ("ident", "ident"), _lib.DER3E: ("odd", "odd"), _lib.DER_MORB: ("ident", "odd"),
    _lib.OMEGA_HPLUS: ("ident", "ident"),
}
_ALPHA, _BETA = np.array([1, 2, 0]), np.array([2, 0, 1])   # utility.py:45-46


class Calculator:

    def __init__(self, degen_thresh=1e-4, degen_Kramers=False, save_mode="bin+txt", print_comment=False):
        self.degen_thresh = degen_thresh
        self.degen_Kramers = degen_Kramers
        self.save_mode = save_mode
        if not hasattr(self, "comment"):
            self.comment = self.__doc__ if self.__doc__ is not None else "calculator not described"
        if print_comment:
            print(self.comment)

    allow_path = False
    allow_grid = True


class StaticCalculator(Calculator):

    extra_kwargs_formula = ()

    def __init__(self, Efermi, tetra=False, smoother=None, constant_factor=1., use_factor=True, kwargs_formula=None,
                 Emin=-np.inf, Emax=np.inf, hole_like=False, k_resolved=False, Formula=None, fder=None,
                 select_bands=None, **kwargs):
        super().__init__(**kwargs)
        self.Efermi = np.array(Efermi)
        if self.Efermi.ndim != 1 or len(self.Efermi) < 1:
            raise ValueError("Efermi must be a 1-d array")
        if k_resolved:
            raise NotImplementedError("k_resolved=True is not implemented on the GPU path")
        if select_bands is not None:
            # static.py:93-100, 129-136: band groups count with the fraction of their bands that is selected
            select_bands = np.array(sorted(set(int(b) for b in np.atleast_1d(select_bands))), dtype=int)
            if tetra:
                raise NotImplementedError("select_bands with tetra=True is not implemented on the GPU path")
            if len(select_bands) and (select_bands[0] < 0 or select_bands[-1] >= 128):
                raise ValueError("select_bands: band indices must lie in [0, 128)")
        if smoother is not None and not callable(smoother):
            raise ValueError("smoother must be callable as smoother(A, axis=0) (wannierberri_b200.smoother or the reference's)")
        self.kwargs_formula = copy(kwargs_formula) if kwargs_formula is not None else {}
        if Formula is not None:
            self.Formula = Formula
        unknown = set(self.kwargs_formula) - {"internal_terms", "external_terms"} - set(self.extra_kwargs_formula)
        if unknown and (not self.is_plugin or isinstance(self.Formula, str)):
            raise NotImplementedError(f"kwargs_formula {sorted(unknown)} are not implemented on the GPU path")
        if tetra and self.is_plugin:
            raise NotImplementedError("tetra=True with a plug-in Formula class is not implemented")
        self.use_factor = use_factor
        self.hole_like = hole_like
        self.tetra, self.smoother, self.k_resolved, self.select_bands = tetra, smoother, k_resolved, select_bands
        if fder is not None:
            self.fder = fder
        assert hasattr(self, "fder"), "fder not set"
        assert hasattr(self, "Formula"), "Formula not set"
        if self.fder not in (0, 1, 2, 3):
            raise NotImplementedError(f"Derivatives  d^{self.fder}f/dE^{self.fder} is not implemented")
        # Emin / Emax: the reference hands them to the band grouping, which reads them in the tetrahedron method only
        # (data_K/data_K.py:172-185 ignores them; grid/tetrahedron.py:246-263: Emin = lower edge of the Fermi-sea
        # group for fder = 0, Emax = upper edge of the inverse Fermi sea of hole_like).  Same here: no effect without
        # tetra; the one case in which they would act raises.
        self.Emin, self.Emax = Emin, Emax
        if tetra and hole_like and self.fder != 0:
            # the reference takes the der = -1 weights whatever fder is (static.py:84-88); only the inverse Fermi SEA is
            # a defined quantity
            raise NotImplementedError("tetra=True with hole_like is implemented for Fermi-sea quantities (fder = 0) only")
        self.constant_factor = constant_factor
        if self.hole_like and self.fder == 0:
            self.constant_factor *= -1
        # static.py:54-58
        self.extraEf = 0 if self.fder == 0 else 1 if self.fder in (1, 2) else 2
        self.dEF = self.Efermi[1] - self.Efermi[0] if len(self.Efermi) > 1 else 0.001
        self.EFmin = self.Efermi[0] - self.extraEf * self.dEF
        self.EFmax = self.Efermi[-1] + self.extraEf * self.dEF
        self.nEF_extra = self.Efermi.shape[0] + 2 * self.extraEf
        # the tetrahedron kernels evaluate the weights at Efermi[0] + i * dEF; the reference's weights_tetra takes the
        # levels as they are (grid/tetrahedron.py:270-281), so anything but a uniform axis must not pass silently
        if tetra and len(self.Efermi) > 2 and not np.allclose(np.diff(self.Efermi), self.dEF, rtol=1e-9, atol=1e-12 * abs(self.dEF)):
            raise NotImplementedError("tetra=True needs uniformly spaced Efermi on the GPU path")

    @property
    def is_plugin(self):
        """`Formula` is a user's class `Formula(data_K, **kwargs_formula)` (plug-in hook #2 of SURVEY.md section 8(b))
        instead of one of the scans that the CUDA kernels evaluate"""
        return not isinstance(getattr(self, "Formula", 0), (int, np.integer))

    # ---- what the engine must evaluate
    def _spec(self, formula=None, fder=None):
        f = self.Formula if formula is None else formula
        factor = self.constant_factor if self.use_factor else float(np.sign(self.constant_factor))
        spec = ScanSpec(formula=f, fder=self.fder if fder is None else fder, nEF=len(self.Efermi),
                        degen_Kramers=int(bool(self.degen_Kramers)),
                        internal_terms=int(bool(self.kwargs_formula.get("internal_terms", True))),
                        external_terms=int(bool(self.kwargs_formula.get("external_terms", True))),
                        Ef_first=float(self.Efermi[0]), Ef_last=float(self.Efermi[-1]), dEF=float(self.dEF),
                        degen_thresh=float(self.degen_thresh), factor=float(factor))
        if self.tetra:   # read by the tetrahedron method only (grid/tetrahedron.py:246-266)
            spec.tetra_flags = (1 if self.hole_like else 0) | (2 if self.Emin != -np.inf else 0) | (4 if self.Emax != np.inf else 0)
            spec.tetra_Emin = float(self.Emin) if self.Emin != -np.inf else 0.
            spec.tetra_Emax = float(self.Emax) if self.Emax != np.inf else 0.
        if self.select_bands is not None:
            if spec.fder == 0:   # data_K.py:179-180
                raise NotImplementedError("Selection of bands for Fermi sea is not implemented")
            spec.use_select = 1
            for b in self.select_bands:
                spec.select_mask[int(b) // 64] |= 1 << (int(b) % 64)
        return spec

    def specs(self):
        return [self._spec()]

    def combine(self, arrays, cell_volume):
        return arrays[0]

    @property
    def external_terms(self):
        return bool(self.kwargs_formula.get("external_terms", True))

    def result(self, arrays, cell_volume):
        tr, inv = _TR_INV[self.Formula]
        return EnergyResult(self.Efermi, self.combine(arrays, cell_volume), transformTR=tr, transformInv=inv,
                            comment=self.comment, save_mode=self.save_mode, smoothers=[self.smoother])

    def __call__(self, data_K):
        """Per-K-block evaluation with the reference's calling convention `calc(data_K)`; `data_K` is a
        `wannierberri_b200.Data_K_R` (GPU resident).  static.py:60-169."""
        if self.is_plugin:
            return self._call_plugin(data_K)
        arrays = data_K.scan(self.specs(), external_terms=self.external_terms, tetra=self.tetra)
        return self.result(arrays, data_K.cell_volume)

    def _call_plugin(self, data_K):
        """Fermi scan of a user-supplied formula object on one K-block (the loop of static.py:60-169 over the band
        groups of every k-point).  Eigenvalues and Hamiltonian-gauge matrices come from the GPU through `data_K`; the
        formula's own `trace()` is the user's code."""
        from math import ceil
        nb = data_K.num_wann
        groups_k = data_K.get_bands_in_range_groups(self.EFmin, self.EFmax, degen_thresh=self.degen_thresh,
                                                    degen_Kramers=self.degen_Kramers, sea=(self.fder == 0),
                                                    Emin=self.Emin, Emax=self.Emax, select_bands=self.select_bands)
        batched = isinstance(self.Formula, str)   # a formula of formula_gpu.py: all (k-point, group) pairs in one go
        if batched:
            from .. import formula_gpu
            _, rank, (tTR, tInv) = formula_gpu.TRACES[self.Formula]
            shape = (3,) * rank
            pairs = [(ik, g) for ik, groups in enumerate(groups_k) for g in groups]
            vals = formula_gpu.batch_traces(self.Formula, data_K, np.array([p[0] for p in pairs], dtype=int),
                                            np.array([p[1][0] for p in pairs], dtype=int),
                                            np.array([p[1][1] for p in pairs], dtype=int),
                                            internal=bool(self.kwargs_formula.get("internal_terms", True)),
                                            external=bool(self.kwargs_formula.get("external_terms", True))
                                            and not getattr(data_K, "force_internal_terms_only", False))
            pair_value = {p: v for p, v in zip(pairs, vals)}
        else:
            formula = self.Formula(data_K, **self.kwargs_formula)
            shape = (3,) * formula.ndim
            tTR, tInv = formula.transformTR, formula.transformInv
        steps = np.zeros((self.nEF_extra + 1,) + shape)   # value added from level i on: summed up at the end
        for ik, groups in enumerate(groups_k):
            if batched:
                values = {g: pair_value[(ik, g)] for g in groups}
            elif getattr(formula, "additive", True):
                values = {g: formula.trace(ik, np.arange(g[0], g[1]),
                                           np.concatenate((np.arange(0, g[0]), np.arange(g[1], nb)))) for g in groups}
            else:   # e.g. the orbital moment: trace over [0, edge) at every group edge, a group = difference of its edges
                edge = {x: formula.trace(ik, np.arange(0, x), np.arange(x, nb)) for x in {e for g in groups for e in g}}
                values = {g: edge[g[1]] - edge[g[0]] for g in groups}
            for g, E in groups.items():
                w = 1.   # utility.py:398-403: fraction of the group's bands that is selected
                if self.select_bands is not None:
                    w = np.sum((self.select_bands >= g[0]) & (self.select_bands < g[1])) / (g[1] - g[0])
                if E < self.EFmin:
                    steps[0] += values[g] * w
                elif E <= self.EFmax:
                    steps[ceil((E - self.EFmin) / self.dEF)] += values[g] * w
        tot = np.cumsum(steps[:-1], axis=0)
        d = self.dEF
        if self.fder == 1:
            tot = (tot[2:] - tot[:-2]) / (2 * d)
        elif self.fder == 2:
            tot = (tot[2:] + tot[:-2] - 2 * tot[1:-1]) / d ** 2
        elif self.fder == 3:
            tot = (tot[4:] - tot[:-4] - 2 * (tot[3:-1] - tot[1:-3])) / (2 * d ** 3)
        tot = tot / (data_K.cell_volume * data_K.nk)
        tot = tot * (self.constant_factor if self.use_factor else np.sign(self.constant_factor))
        return EnergyResult(self.Efermi, tot, transformTR=tTR, transformInv=tInv,
                            comment=self.comment, save_mode=self.save_mode, smoothers=[self.smoother])


class _DOS(StaticCalculator):

    def __init__(self, fder, **kwargs):
        self.Formula = _lib.IDENTITY
        self.fder = fder
        super().__init__(**kwargs)

    def combine(self, arrays, cell_volume):  # static.py:199-200
        return arrays[0] * cell_volume


class DOS(_DOS):
    r"""Density of states"""

    def __init__(self, **kwargs):
        super().__init__(fder=1, **kwargs)


class CumDOS(_DOS):
    r"""Cumulative density of states"""

    def __init__(self, **kwargs):
        super().__init__(fder=0, **kwargs)


class Spin(StaticCalculator):
    r"""Spin per unit cell (dimensionless)"""

    def __init__(self, **kwargs):
        self.Formula = _lib.SPIN
        self.fder = 0
        super().__init__(**kwargs)

    def combine(self, arrays, cell_volume):
        return arrays[0] * cell_volume


class AHC(StaticCalculator):
    r"""Anomalous Hall conductivity (:math:`s^3 \cdot A^2 / (kg \cdot m^3) = S/m`)

        | Output: :math:`O = -e^2/\hbar \int [dk] \Omega f`"""

    def __init__(self, constant_factor=factors.factor_ahc, **kwargs):
        self.Formula = _lib.OMEGA
        self.fder = 0
        super().__init__(constant_factor=constant_factor, **kwargs)


class Morb(StaticCalculator):
    r"""Orbital magnetic moment per unit cell (:math:`\mu_B`)

        | Output: :math:`M = -\int [dk] (H + G - 2E_f \cdot \Omega) f`"""

    def __init__(self, constant_factor=-factors.eV_au / factors.bohr ** 2, **kwargs):
        self.Formula = _lib.MORB_HPM
        self.fder = 0
        super().__init__(constant_factor=constant_factor, **kwargs)

    def specs(self):  # static.py:238-247: Hplus and AHC(constant_factor=same)
        return [self._spec(), self._spec(formula=_lib.OMEGA)]

    def combine(self, arrays, cell_volume):
        Hplus, Om = arrays
        return (Hplus - 2 * Om * self.Efermi[:, None]) * cell_volume


class BerryDipole_FermiSurf(StaticCalculator):
    r"""Berry curvature dipole (dimensionless)

        | Output: :math:`D_{\beta\delta} = -\int [dk] v_\beta \Omega_\delta f'`"""

    def __init__(self, **kwargs):
        self.Formula = _lib.VEL_OMEGA
        self.fder = 1
        super().__init__(**kwargs)


class GME_orb_FermiSurf(StaticCalculator):
    r"""Gyrotropic tensor orbital part (:math:`A`), Fermi surface integral

        | Output: :math:`K^{orb}_{\alpha :\mu} = \int [dk] v_\alpha m^{orb}_\mu f'`"""

    def __init__(self, constant_factor=factors.factor_gme_orb, **kwargs):
        self.Formula = _lib.VEL_HPLUS
        self.fder = 1
        super().__init__(constant_factor=constant_factor, **kwargs)

    def specs(self):  # static.py:273-281
        return [self._spec(), self._spec(formula=_lib.VEL_OMEGA)]

    def combine(self, arrays, cell_volume):
        Hplus, Om = arrays
        return Hplus - 2 * Om * self.Efermi[:, None, None]


class GME_spin_FermiSurf(StaticCalculator):
    r"""Gyrotropic tensor spin part (:math:`A`), Fermi surface integral

        | Output: :math:`K^{spin}_{\alpha :\mu} = \tau \int [dk] v_\alpha s_\mu f'`"""

    def __init__(self, constant_factor=factors.factor_gme_spin, **kwargs):
        self.Formula = _lib.VEL_SPIN
        self.fder = 1
        super().__init__(constant_factor=constant_factor, **kwargs)


class Ohmic_FermiSurf(StaticCalculator):
    r"""Ohmic conductivity (:math:`S/m`), Fermi surface integral

        | Output: :math:`\sigma_{\alpha\beta} = -e^2/\hbar \tau \int [dk] v_\alpha v_\beta f'`"""

    def __init__(self, constant_factor=factors.factor_ohmic, **kwargs):
        self.Formula = _lib.VEL_VEL
        self.fder = 1
        super().__init__(constant_factor=constant_factor, **kwargs)


class Ohmic_FermiSea(StaticCalculator):
    r"""Ohmic conductivity (:math:`S/m`), Fermi sea integral

        | Output: :math:`\sigma_{\alpha\beta} = e^2/\hbar \tau \int [dk] \partial_\beta v_\alpha f`"""

    def __init__(self, constant_factor=factors.factor_ohmic, **kwargs):
        self.Formula = _lib.INV_MASS
        self.fder = 0
        super().__init__(constant_factor=constant_factor, **kwargs)


class BerryDipole_FermiSea(StaticCalculator):
    r"""Berry curvature dipole as a Fermi-sea integral of the generalised derivative of the Berry curvature
    (static.py:472-487, formula DerOmega): data `[Efermi, beta, delta]`, dimensionless."""

    def __init__(self, **kwargs):
        self.Formula = _lib.DER_OMEGA
        self.fder = 0
        super().__init__(**kwargs)

    def combine(self, arrays, cell_volume):  # static.py:483-487: axes swapped to (beta, delta)
        return np.ascontiguousarray(arrays[0].swapaxes(1, 2))


class NLAHC_FermiSea(BerryDipole_FermiSea):
    r"""BerryDipole_FermiSea in the units of the nonlinear anomalous Hall conductivity, S^2/A (static.py:490-498)."""

    def __init__(self, constant_factor=factors.factor_nlahc, **kwargs):
        super().__init__(constant_factor=constant_factor, **kwargs)


class SHC(StaticCalculator):
    r"""Spin Hall conductivity with dc (:math:`S/m`), Fermi sea integral of the spin Berry curvature (static.py:647-660);
    `kwargs_formula={'spin_current_type': 'ryoo' | 'qiao' | 'simple'}`.  Output `[Efermi, a, b, s]`."""
    extra_kwargs_formula = ("spin_current_type",)

    def __init__(self, constant_factor=-factors.factor_ahc / 2, **kwargs):
        t = (kwargs.get("kwargs_formula") or {}).get("spin_current_type", "ryoo")
        if t not in _lib.SHC_TYPES:
            raise ValueError(f"spin_current_type {t} not recognized")  # formula/covariant.py:696-697
        self.Formula = _lib.SHC_TYPES[t]
        self.fder = 0
        super().__init__(constant_factor=constant_factor, **kwargs)


class NLAHC_FermiSurf(BerryDipole_FermiSurf):
    r"""Nonlinear anomalous Hall conductivity (:math:`S^2/A`), Fermi surface integral (static.py:461-469)"""

    def __init__(self, constant_factor=factors.factor_nlahc, **kwargs):
        super().__init__(constant_factor=constant_factor, **kwargs)


class GME_spin_FermiSea(StaticCalculator):
    r"""Spin part of the gyrotropic tensor as a Fermi-sea integral of the generalised derivative of the spin
    (static.py:322-337, formula DerSpin): data `[Efermi, alpha, mu]`, Ampere."""

    def __init__(self, constant_factor=factors.factor_gme_spin, **kwargs):
        self.Formula = _lib.DER_SPIN
        self.fder = 0
        super().__init__(constant_factor=constant_factor, **kwargs)

    def combine(self, arrays, cell_volume):
        return np.ascontiguousarray(arrays[0].swapaxes(1, 2))


def _hall_classic_post(d):   # static.py:419-423
    d = d[:, :, :, _BETA, _ALPHA] - d[:, :, :, _ALPHA, _BETA]
    return np.ascontiguousarray(0.5 * (d[:, _ALPHA, _BETA, :] - d[:, _BETA, _ALPHA, :]))


class Hall_classic_FermiSurf(StaticCalculator):
    r"""Classic Hall conductivity (:math:`S/m/T`), Fermi surface integral (static.py:407-424)"""

    def __init__(self, constant_factor=factors.factor_hall_classic, **kwargs):
        self.Formula = _lib.VEL_MASS_VEL
        self.fder = 1
        super().__init__(constant_factor=constant_factor, **kwargs)

    def combine(self, arrays, cell_volume):
        return _hall_classic_post(arrays[0])


class Hall_classic_FermiSea(StaticCalculator):
    r"""Classic Hall conductivity (:math:`S/m/T`), Fermi sea integral (static.py:427-445)"""

    def __init__(self, constant_factor=factors.factor_hall_classic, **kwargs):
        self.Formula = _lib.MASS_MASS
        self.fder = 0
        super().__init__(constant_factor=constant_factor, **kwargs)

    def combine(self, arrays, cell_volume):
        return _hall_classic_post(arrays[0].transpose(0, 4, 1, 2, 3))


class NLDrude_FermiSurf(StaticCalculator):
    r"""Drude conductivity (:math:`S^2/A`), Fermi surface integral (static.py:532-542)"""

    def __init__(self, constant_factor=factors.factor_nldrude, **kwargs):
        self.Formula = _lib.MASS_VEL
        self.fder = 1
        super().__init__(constant_factor=constant_factor, **kwargs)


class GME_orb_FermiSea(StaticCalculator):
    r"""Orbital part of the gyrotropic tensor as a Fermi-sea integral (static.py:284-300): DerMorb scan minus
    2 E_F times the BerryDipole_FermiSea scan with the same prefactor; data `[Efermi, alpha, mu]`, Ampere."""

    def __init__(self, constant_factor=factors.factor_gme_orb, **kwargs):
        self.Formula = _lib.DER_MORB
        self.fder = 0
        super().__init__(constant_factor=constant_factor, **kwargs)

    def specs(self):   # DerMorb and BerryDipole_FermiSea(constant_factor = same)
        return [self._spec(), self._spec(formula=_lib.DER_OMEGA)]

    def combine(self, arrays, cell_volume):
        Hplus, Om = arrays
        return np.ascontiguousarray((Hplus - 2 * Om * self.Efermi[:, None, None]).swapaxes(1, 2))


class NLDrude_FermiSea(StaticCalculator):
    r"""Drude conductivity (:math:`S^2/A`), Fermi sea integral of the third derivative of the band energy (static.py:519-529)"""

    def __init__(self, constant_factor=factors.factor_nldrude, **kwargs):
        self.Formula = _lib.DER3E
        self.fder = 0
        super().__init__(constant_factor=constant_factor, **kwargs)


class NLDrude_Fermider2(StaticCalculator):
    r"""Drude conductivity (:math:`S^2/A`), second derivative of the distribution function (static.py:545-555)"""

    def __init__(self, constant_factor=factors.factor_nldrude / 2, **kwargs):
        self.Formula = _lib.VEL_VEL_VEL
        self.fder = 2
        super().__init__(constant_factor=constant_factor, **kwargs)


class AHC_Zeeman_spin(StaticCalculator):
    r"""AHC conductivity Zeeman correction term spin part (:math:`S/m/T`), Fermi surface integral (static.py:605-615)"""

    def __init__(self, constant_factor=factors.fac_spin_Z * factors.factor_ahc, **kwargs):
        self.Formula = _lib.OMEGA_S
        self.fder = 1
        super().__init__(constant_factor=constant_factor, **kwargs)


class AHC_Zeeman_orb(StaticCalculator):
    r"""AHC conductivity Zeeman correction term orbital part (:math:`S/m/T`), Fermi surface integral (static.py:626-645)"""

    def __init__(self, constant_factor=factors.fac_orb_Z * factors.factor_ahc, **kwargs):
        self.Formula = _lib.OMEGA_HPLUS
        self.fder = 1
        super().__init__(constant_factor=constant_factor, **kwargs)

    def specs(self):   # OmegaHplus and OmegaOmega(constant_factor = same)
        return [self._spec(), self._spec(formula=_lib.OMEGA_OMEGA)]

    def combine(self, arrays, cell_volume):
        Hplus, Om = arrays
        return Hplus - 2 * Om * self.Efermi[:, None, None]


class OmegaOmega(StaticCalculator):
    r"""static.py:618-623"""

    def __init__(self, **kwargs):
        self.Formula = _lib.OMEGA_OMEGA
        self.fder = 1
        super().__init__(**kwargs)


# ---- second-order formulae: evaluated K-block by K-block on the GPU-resident Data_K_R by
# the batched block algebra of formula_gpu.py (torch on the device, matrices with up to three comma-derivatives from the
# CUDA kernels); `Formula` is the name of the trace there

class eMChA_FermiSurf(StaticCalculator):
    r"""electrical magnetochiral anisotropy, Fermi-surface term (static.py:560-566; formula covariant.py:868-890)"""
    extra_kwargs_formula = ()

    def __init__(self, constant_factor=factors.factor_emcha, **kwargs):
        self.Formula = "emcha_surf"
        self.fder = 1
        super().__init__(constant_factor=constant_factor, **kwargs)


class NLDrude_Zeeman_spin(StaticCalculator):
    r"""Zeeman (spin) correction of the non-linear Drude conductivity (static.py:569-575; covariant.py:893-899)"""

    def __init__(self, constant_factor=factors.fac_spin_Z * factors.factor_nldrude, **kwargs):
        self.Formula = "NLDrude_Z_spin"
        self.fder = 1
        super().__init__(constant_factor=constant_factor, **kwargs)


class NLDrude_Zeeman_orb_Omega(StaticCalculator):
    r"""Berry-curvature part of the orbital Zeeman correction (static.py:578-584; covariant.py:909-914)"""

    def __init__(self, **kwargs):
        self.Formula = "NLDrude_Z_orb_Omega"
        self.fder = 1
        super().__init__(**kwargs)


class NLDrude_Zeeman_orb(StaticCalculator):
    r"""Zeeman (orbital) correction of the non-linear Drude conductivity: Hplus part - 2 E_F x Omega part
    (static.py:587-603; covariant.py:901-907)"""

    def __init__(self, **kwargs):
        self.Formula = "NLDrude_Z_orb_Hplus"
        self.fder = 1
        super().__init__(**kwargs)

    def __call__(self, data_K):
        Hplus = super().__call__(data_K)
        Om = NLDrude_Zeeman_orb_Omega(Efermi=self.Efermi, tetra=self.tetra, smoother=self.smoother, use_factor=False,
                                      kwargs_formula=self.kwargs_formula)(data_K).mul_array(self.Efermi)
        final_factor = factors.fac_orb_Z * factors.factor_nldrude
        if not self.use_factor:
            final_factor = np.sign(final_factor)
        return (Hplus - Om * 2.) * final_factor


class _QuantumMetric(StaticCalculator):
    extra_kwargs_formula = ("FF_rotAA",)

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        if self.kwargs_formula.get("external_terms", True) and not self.kwargs_formula.get("FF_rotAA", False):
            raise NotImplementedError("the external terms of the quantum metric need FF_R, which the GPU path does not "
                                      "transform: pass kwargs_formula=dict(FF_rotAA=True) or external_terms=False")


class QuantumMetric_FermiSea(_QuantumMetric):
    r"""quantum metric of the occupied states (static.py:663-670; basic.py:19-51, covariant.py:916-922)"""

    def __init__(self, constant_factor=1., **kwargs):
        self.Formula = "QuantumMetric_ab"
        self.fder = 0
        super().__init__(constant_factor=constant_factor, **kwargs)


class QuantumMetric_Vel_DQ(_QuantumMetric):
    r"""quantum-metric dipole with the velocity (static.py:673-680; basic.py:59-95, covariant.py:925-936)"""

    def __init__(self, constant_factor=1., **kwargs):
        self.Formula = "VelDQM"
        self.fder = 1
        super().__init__(constant_factor=constant_factor, **kwargs)


_BATCHED = (eMChA_FermiSurf, NLDrude_Zeeman_spin, NLDrude_Zeeman_orb_Omega, NLDrude_Zeeman_orb, QuantumMetric_FermiSea,
            QuantumMetric_Vel_DQ)

_BY_NAME = {c.__name__: c for c in _BATCHED + (DOS, CumDOS, Spin, AHC, Morb, BerryDipole_FermiSurf, GME_orb_FermiSurf,
                                    GME_spin_FermiSurf, Ohmic_FermiSurf, Ohmic_FermiSea, BerryDipole_FermiSea,
                                    NLAHC_FermiSea, SHC, NLAHC_FermiSurf, GME_spin_FermiSea, Hall_classic_FermiSurf,
                                    Hall_classic_FermiSea, NLDrude_FermiSurf, NLDrude_Fermider2, NLDrude_FermiSea, GME_orb_FermiSea, AHC_Zeeman_spin, AHC_Zeeman_orb,
                                    OmegaOmega)}


def adapt(calc):
    """Accept a calculator of this package, or an instance of the reference's class of the same name
    (recognised by class name; its public attributes are copied)."""
    if isinstance(calc, StaticCalculator):
        return calc
    name = type(calc).__name__
    if name not in _BY_NAME:
        raise KeyError(f"calculator {name} is not available on the GPU path")
    kw = dict(Efermi=np.array(calc.Efermi), tetra=calc.tetra, smoother=calc.smoother, use_factor=calc.use_factor,
              kwargs_formula=calc.kwargs_formula, hole_like=False, k_resolved=calc.k_resolved,
              Emin=getattr(calc, "Emin", -np.inf), Emax=getattr(calc, "Emax", np.inf),
              select_bands=calc.select_bands, degen_thresh=calc.degen_thresh, degen_Kramers=calc.degen_Kramers,
              save_mode=calc.save_mode)
    fixed = ("DOS", "CumDOS", "Spin", "BerryDipole_FermiSurf", "BerryDipole_FermiSea", "OmegaOmega", "NLDrude_Zeeman_orb_Omega",
             "NLDrude_Zeeman_orb")   # no constant_factor argument
    if name not in fixed:
        kw["constant_factor"] = calc.constant_factor  # hole_like sign already folded in by the reference
    new = _BY_NAME[name](**kw)
    if name in fixed:
        new.constant_factor = calc.constant_factor
    if calc.tetra and getattr(calc, "hole_like", False):   # (its sign is in constant_factor already; the weights need the flag)
        if new.fder != 0:
            raise NotImplementedError("tetra=True with hole_like is implemented for Fermi-sea quantities (fder = 0) only")
        new.hole_like = True
    return new
