"""A user-defined formula and calculator factory written against the plug-in interface only (SURVEY.md section 8(b),
hooks #1 and #2): it reads `data_K.Dcov`, `data_K.covariant(...)`, `data_K.E_K` and returns band blocks.  The same
source runs on the reference's `Data_K_R` (tests/golden/make_golden_plugin.py, which makes the fixture) and on this
package's GPU-resident `Data_K_R` (tests/test_gpu_parity.py)."""
import numpy as np


class UserFormula:
    """rank-2 covariant test quantity
        F^{ab}_{nn'} = scale * sum_{l in out} (D^a_{nl} A^b_{ln'} + conj. transpose) + S^{a:b}_{nn'} + w sum_m V^a_{nm} S^b_{mn'}
    (Berry-connection cross term, generalised derivative of the spin, velocity x spin inside the band set);
    `weighted`: multiply by the mean energy of the inner states -- makes the trace non-additive."""

    def __init__(self, data_K, scale=1.0, w=0.5, weighted=False, **unused):
        self.D = data_K.Dcov
        self.A = data_K.covariant("AA")
        self.S = data_K.covariant("SS")
        self.dS = data_K.covariant("SS", gender=1)
        self.V = data_K.covariant("Ham", commader=1)
        self.E = data_K.E_K
        self.scale, self.w, self.weighted = scale, w, weighted
        self.ndim = 2
        self.transformTR = self.dS.transformTR
        self.transformInv = self.dS.transformInv

    @property
    def additive(self):
        return not self.weighted

    def nn(self, ik, inn, out):
        x = np.einsum("mla,lnb->mnab", self.D.nl(ik, inn, out), self.A.ln(ik, inn, out))
        res = self.scale * (x + x.swapaxes(0, 1).conj())
        res = res + self.dS.nn(ik, inn, out)
        res = res + self.w * np.einsum("mla,lnb->mnab", self.V.nn(ik, inn, out), self.S.nn(ik, inn, out))
        if self.weighted:
            res = res * (self.E[ik][inn].mean() if len(inn) else 0.)
        return res

    def ln(self, ik, inn, out):
        raise NotImplementedError()

    def trace(self, ik, inn, out):
        return np.einsum("nn...->...", self.nn(ik, inn, out)).real


class UserFormula2:
    """rank-3 test quantity that reads SECOND comma-derivatives:
        G^{bde}_{nn'} = S^{b,de}_{nn'} + i sum_{l in out} (D^d_{nl} A^{b,e}_{ln'} - A^{b,e}_{nl} D^d_{ln'}) + w A^{b,de}_{nn'}"""

    def __init__(self, data_K, w=0.3, **unused):
        self.D = data_K.Dcov
        self.dA = data_K.covariant("AA", commader=1)
        self.ddA = data_K.covariant("AA", commader=2)
        self.ddS = data_K.covariant("SS", commader=2)
        self.w = w
        self.ndim = 3
        self.transformTR = self.ddS.transformTR
        self.transformInv = self.ddS.transformInv
        self.additive = True

    def nn(self, ik, inn, out):
        res = np.array(self.ddS.nn(ik, inn, out))
        res = res + 1j * np.einsum("mld,lnbe->mnbde", self.D.nl(ik, inn, out), self.dA.ln(ik, inn, out))
        res = res - 1j * np.einsum("mlbe,lnd->mnbde", self.dA.nl(ik, inn, out), self.D.ln(ik, inn, out))
        return res + self.w * self.ddA.nn(ik, inn, out)

    def ln(self, ik, inn, out):
        raise NotImplementedError()

    def trace(self, ik, inn, out):
        return np.einsum("nn...->...", self.nn(ik, inn, out)).real


def make_calculators(StaticCalculator, Efermi):
    """{key: calculator} on top of the given `StaticCalculator` base class (the reference's or this package's)"""

    def cls(fder_, name, frm=UserFormula):
        class C(StaticCalculator):
            def __init__(self, **kwargs):
                self.Formula = frm
                self.fder = fder_
                self.comment = name
                super().__init__(**kwargs)
        C.__name__ = name
        return C

    return dict(
        user_sea=cls(0, "UserSea")(Efermi=Efermi, kwargs_formula=dict(scale=0.7)),
        user_surf=cls(1, "UserSurf")(Efermi=Efermi, kwargs_formula=dict(scale=1.3, w=-0.2), degen_thresh=0.05),
        user_weighted=cls(0, "UserWeighted")(Efermi=Efermi, kwargs_formula=dict(weighted=True), constant_factor=2.5),
        user_der2=cls(1, "UserDer2", UserFormula2)(Efermi=Efermi, kwargs_formula=dict(w=0.3)),
    )
