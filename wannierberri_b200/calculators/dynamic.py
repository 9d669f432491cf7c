"""Kubo (Efermi x omega) calculators with the constructor interface of the reference
(calculators/dynamic.py:26-57): same class names and keyword arguments.  Like the static calculators they do no
arithmetic on k-points: `spec()` declares the scan that libwbgpu.so evaluates (wbgpu_kubo_scan)."""
from copy import copy

import numpy as np

from .. import factors
from .. import _lib
from .._lib import KuboSpec
from ..result import EnergyResult
from .static import Calculator


class DynamicCalculator(Calculator):

    kind = None
    dtype = complex
    extra_kwargs_formula = ()
    shc_type = 0
    sc_eta = 0.

    def __init__(self, Efermi=None, omega=None, kBT=0, smr_fixed_width=0.1, smr_type='Lorentzian',
                 kwargs_formula=None, dtype=None, **kwargs):
        super().__init__(**kwargs)
        self.Efermi = np.array(Efermi, dtype=float).reshape(-1)
        self.omega = np.array(omega, dtype=float).reshape(-1)
        if kBT < 0:
            raise ValueError("kBT must not be negative")
        if smr_type not in ("Lorentzian", "Gaussian"):
            raise ValueError(f"Invalid smearing type {smr_type}")  # dynamic.py:54
        if len(self.Efermi) > 1 and not np.all(np.diff(self.Efermi) > 0):
            raise NotImplementedError("Efermi must be strictly ascending on the GPU path")
        self.kBT = kBT
        self.smr_fixed_width = smr_fixed_width
        self.smr_type = smr_type
        self.kwargs_formula = copy(kwargs_formula) if kwargs_formula is not None else {}
        unknown = set(self.kwargs_formula) - {"external_terms"} - set(self.extra_kwargs_formula)
        if unknown:
            raise NotImplementedError(f"kwargs_formula {sorted(unknown)} are not implemented on the GPU path")
        self.constant_factor = 1.

    @property
    def external_terms(self):
        return bool(self.kwargs_formula.get("external_terms", True))

    def spec(self):
        return KuboSpec(kind=self.kind, nEF=len(self.Efermi), nomega=len(self.omega),
                        smr_type=0 if self.smr_type == "Lorentzian" else 1,
                        degen_Kramers=int(bool(self.degen_Kramers)), external_terms=int(self.external_terms),
                        shc_type=int(self.shc_type), smr_fixed_width=float(self.smr_fixed_width),
                        degen_thresh=float(self.degen_thresh), factor=float(self.constant_factor),
                        sc_eta=float(self.sc_eta), kBT=float(self.kBT))

    def result(self, data):
        return EnergyResult([self.Efermi, self.omega], data, transformTR=self.transformTR,
                            transformInv=self.transformInv, E_titles=("Efermi", "omega"), comment=self.comment,
                            save_mode=self.save_mode, smoothers=(None, None))

    def __call__(self, data_K):
        """`calc(data_K)` for one K-block (dynamic.py:71-114)."""
        return self.result(data_K.kubo_scan(self.spec(), self.Efermi, self.omega))


class JDOS(DynamicCalculator):
    r"""Joint Density of States"""
    kind = _lib.KUBO_JDOS
    dtype = float
    transformTR, transformInv = "ident", "ident"

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.sigma = self.smr_fixed_width

    @property
    def external_terms(self):
        """JDOS reads band energies only (dynamic.py:146-162): no AA channel is asked of the plan, so it runs on a
        system that holds nothing but Ham, as in the reference."""
        return False


class OpticalConductivity(DynamicCalculator):
    r"""Optical conductivity :math:`\sigma_{ab}(\omega)` (Kubo-Greenwood), S/m"""
    kind = _lib.KUBO_OPTCOND
    transformTR, transformInv = "trans", "ident"

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.constant_factor = factors.factor_opt


class SHC(DynamicCalculator):
    r"""Spin Hall conductivity :math:`\sigma^{s}_{ab}(\omega)` (Kubo; dynamic.py:224-237), data `[Efermi, omega, a, b, s]`
    with the spin current of Ryoo et al. (`SHC_type="ryoo"`, needs SA, SHA), of Qiao et al. (`"qiao"`: SR, SH, SHR)
    or `{S, v}/2` (`"simple"`).  `shc_abc=(a, b, c)` (1-based) keeps one component."""
    kind = _lib.KUBO_SHC
    transformTR, transformInv = "ident", "ident"
    extra_kwargs_formula = ("SHC_type", "shc_abc")

    def __init__(self, SHC_type="ryoo", shc_abc=None, **kwargs):
        super().__init__(**kwargs)
        SHC_type = self.kwargs_formula.get("SHC_type", SHC_type)
        shc_abc = self.kwargs_formula.get("shc_abc", shc_abc)
        if SHC_type not in _lib.SHC_TYPES:
            raise ValueError(f"spin_current_type {SHC_type} not recognized")  # formula/covariant.py:696-697
        if shc_abc is not None:
            assert len(shc_abc) == 3
        self.kwargs_formula.update(dict(SHC_type=SHC_type, shc_abc=shc_abc))
        self.SHC_type, self.shc_abc = SHC_type, shc_abc
        self.shc_type = _lib.SHC_TYPES[SHC_type]
        self.constant_factor = factors.factor_shc

    def result(self, data):
        if self.shc_abc is not None and data.ndim == 5:   # Formula_SHC, dynamic.py:212-216
            a, b, c = (x - 1 for x in self.shc_abc)
            data = np.ascontiguousarray(data[:, :, a, b, c])
        return super().result(data)


class ShiftCurrent(DynamicCalculator):
    r"""Shift current :math:`\sigma^{abc}(\omega)` (dynamic.py:244-322), data `[Efermi, omega, a, b, c]` real; `sc_eta`
    broadens the energy denominators of the generalised derivative of the Berry connection."""
    kind = _lib.KUBO_SHIFT
    dtype = float
    transformTR, transformInv = "ident", "odd"
    extra_kwargs_formula = ("sc_eta",)

    def __init__(self, sc_eta, **kwargs):
        super().__init__(**kwargs)
        self.sc_eta = float(self.kwargs_formula.get("sc_eta", sc_eta))
        self.kwargs_formula.update(dict(sc_eta=self.sc_eta))
        self.constant_factor = factors.factor_shift_current


class InjectionCurrent(DynamicCalculator):
    r"""Injection current (dynamic.py:330-365; Eq. (10) of Lihm and Park, PRB 105, 045201), data `[Efermi, omega, a, b, c]`"""
    kind = _lib.KUBO_INJECTION
    transformTR, transformInv = "odd_trans_021", "odd"

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.constant_factor = factors.factor_injection_current


_BY_NAME = {c.__name__: c for c in (JDOS, OpticalConductivity, SHC, ShiftCurrent, InjectionCurrent)}


def adapt(calc):
    """Accept a calculator of this package or an instance of the reference's class of the same name."""
    if isinstance(calc, DynamicCalculator):
        return calc
    name = type(calc).__name__
    if name not in _BY_NAME:
        raise ValueError(f"calculator {name} is not available on the GPU path")
    extra = dict(sc_eta=calc.kwargs_formula["sc_eta"]) if name == "ShiftCurrent" else {}
    new = _BY_NAME[name](Efermi=np.array(calc.Efermi), omega=np.array(calc.omega), kBT=calc.kBT, **extra,
                         smr_fixed_width=calc.smr_fixed_width, smr_type=calc.smr_type,
                         kwargs_formula=calc.kwargs_formula, degen_thresh=calc.degen_thresh,
                         degen_Kramers=calc.degen_Kramers, save_mode=calc.save_mode)
    new.constant_factor = calc.constant_factor
    return new
