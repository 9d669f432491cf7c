"""Second-order covariant formulae (Der2Omega / Der2Spin / Der2Morb, eMChA, the Zeeman
corrections of the non-linear Drude weight, quantum metric) evaluated ON THE GPU for a whole K-block at once.

The reference evaluates a `Formula_ln` one (k-point, band group) at a time with numpy blocks `nn / nl / ln / ll`
(formula/formula.py:9-118, formula/covariant.py, formula/elementary.py, formula/basic.py).  Here the same algebra is
written for FULL nw x nw matrices with the band partition (P = inner set, Q = its complement) carried by 0/1 masks, so
that all (k-point, group) pairs of a K-block form one batch dimension `z` of device tensors:

  * a block of a stored matrix is the matrix times a row mask and a column mask;
  * `D` only ever connects the two sets (elementary.py:42-49): `Dx = P D Q + Q D P`; then the generalised derivative
    (formula/formula.py:95-118) of ANY block is one expression,  X^{:d} = d_d X + X Dx^d - Dx^d X  -- its PP block is
    the reference's `nn`, its QP block the reference's `ln`, its QQ block the reference's `ll` (the roles of the two
    sets swapped), because the terms that would differ vanish with the zero blocks of `Dx`;
  * DerDcov / Der2Dcov (elementary.py:55-96) likewise, with the block-diagonal parts of the velocity and of the inverse
    mass: they, too, only live on the two off-diagonal blocks.

Inputs are the Hamiltonian-gauge matrices with up to three comma-derivatives that the CUDA kernels produce
(`Data_K_R.Xbar`: R->k transform, eigensolver, U^dagger X U); the contractions are batched `torch.einsum` calls on the
device (cuBLAS / ATen), i.e. library kernels -- these rarely used calculators are NOT part of the hand-written hot
path and are not tuned.  There is no CPU fallback: without a CUDA device the evaluation raises."""
import numpy as np

AL, BE = [1, 2, 0], [2, 0, 1]   # utility.py:45-46


def _torch():
    import torch
    return torch


class BlockAlgebra:
    """Matrices of a batch of (k-point, inner band set [a, b)) pairs, and the derived covariant objects, memoised."""

    def __init__(self, data_K, kidx, a, b, device, internal=True, external=True):
        torch = _torch()
        self.t = torch
        self.dev = device
        self.data_K = data_K
        self.internal, self.external = internal, external
        self.nw = data_K.num_wann
        self.kidx = torch.as_tensor(np.asarray(kidx), device=device, dtype=torch.long)
        band = torch.arange(self.nw, device=device)
        a = torch.as_tensor(np.asarray(a), device=device)
        b = torch.as_tensor(np.asarray(b), device=device)
        self.P = ((band[None, :] >= a[:, None]) & (band[None, :] < b[:, None])).to(torch.float64)
        self.Q = 1. - self.P
        self.memo = {}

    # ---- stored matrices
    def X(self, name, der=0):
        """Xbar(name, der) of the pairs' k-points, `[z][nw][nw][3]^...` complex128 on the device"""
        key = ("X", name, der)
        if key not in self.memo:
            cache = self.data_K.__dict__.setdefault("_device_xbar", {})
            if (name, der) not in cache:
                cache[(name, der)] = self.t.from_numpy(np.ascontiguousarray(self.data_K.Xbar(name, der))).to(self.dev)
            self.memo[key] = cache[(name, der)][self.kidx]
        return self.memo[key]

    def get(self, key, make):
        if key not in self.memo:
            self.memo[key] = make()
        return self.memo[key]

    @property
    def E(self):
        return self.get("E", lambda: self.t.from_numpy(np.ascontiguousarray(self.data_K.E_K)).to(self.dev)[self.kidx])

    @property
    def dEinv(self):
        """1 / (E_m - E_n), zero for |E_m - E_n| < 1e-7 (data_K.py:290-298)"""
        def make():
            dE = self.E[:, :, None] - self.E[:, None, :]
            close = dE.abs() < 1e-7
            return self.t.where(close, self.t.zeros_like(dE), 1. / self.t.where(close, self.t.ones_like(dE), dE))
        return self.get("dEinv", make)

    # ---- masks
    def _m(self, which, T, axis):
        m = self.P if which == "n" else self.Q
        shape = [1] * T.dim()
        shape[0], shape[axis] = m.shape[0], m.shape[1]
        return m.reshape(shape)

    def blk(self, T, rows, cols):
        """block of T: rows / cols in the inner set ('n') or in its complement ('l')"""
        return T * self._m(rows, T, 1) * self._m(cols, T, 2)

    def off(self, T):
        return self.blk(T, "n", "l") + self.blk(T, "l", "n")

    def dia(self, T):
        return self.blk(T, "n", "n") + self.blk(T, "l", "l")

    def mm(self, pattern, *ops):
        """einsum over band / Cartesian indices with the batch index prefixed to every operand and to the result"""
        lhs, rhs = pattern.split("->")
        return self.t.einsum(",".join("z" + s for s in lhs.split(",")) + "->z" + rhs, *ops)

    @staticmethod
    def hc(T):
        """conjugate transpose in the band indices"""
        return T.transpose(1, 2).conj()

    # ---- elementary objects
    @property
    def V(self):
        return self.X("Ham", 1)

    @property
    def Dx(self):
        """inter-set part of D^H_a = -V_a / (E_m - E_n) (data_K.py:324-326, elementary.py:42-49)"""
        return self.get("Dx", lambda: self.off(-self.V * self.dEinv[..., None]))

    def gender(self, T, Tc):
        """generalised derivative of a covariant matrix T with comma-derivative Tc; the new index comes last"""
        return Tc + self.mm("ml...,lnd->mn...d", T, self.Dx) - self.mm("mld,ln...->mn...d", self.Dx, T)

    def cov(self, name, gender=0, commader=0):
        if gender == 0:
            return self.X(name, commader)
        return self.get(("gen", name), lambda: self.gender(self.X(name, 0), self.X(name, 1)))

    @property
    def dV(self):
        """inverse mass = generalised derivative of the velocity (elementary.py:28-34)"""
        return self.get("dV", lambda: self.gender(self.V, self.X("Ham", 2)))

    @property
    def dW(self):
        """generalised derivative of d_b d_c H (elementary.py:36-43)"""
        return self.get("dW", lambda: self.gender(self.X("Ham", 2), self.X("Ham", 3)))

    @property
    def dD(self):
        """generalised derivative of D between the two sets (elementary.py:55-72)"""
        def make():
            Vd = self.dia(self.V)
            t = self.mm("lpb,pnd->lnbd", Vd, self.Dx) - self.mm("lmb,mnd->lnbd", self.Dx, Vd)
            s = self.X("Ham", 2) + t + t.transpose(3, 4)
            return self.off(-s * self.dEinv[..., None, None])
        return self.get("dD", make)

    @property
    def ddD(self):
        """second generalised derivative of D (elementary.py:75-96)"""
        def make():
            Vd, dVd, D, dD = self.dia(self.V), self.dia(self.dV), self.Dx, self.dD
            s = self.dW.clone()
            for sign, pat, x, y in ((+1, "lpbe,pnd->lnbde", dVd, D), (+1, "lpde,pnb->lnbde", dVd, D),
                                    (+1, "lpe,pnbd->lnbde", Vd, dD), (+1, "lpd,pnbe->lnbde", Vd, dD),
                                    (+1, "lpb,pnde->lnbde", Vd, dD), (-1, "lmde,mnb->lnbde", dD, Vd),
                                    (-1, "lmbd,mne->lnbde", dD, Vd), (-1, "lmbe,mnd->lnbde", dD, Vd),
                                    (-1, "lmb,mnde->lnbde", D, dVd), (-1, "lmd,mnbe->lnbde", D, dVd)):
                s += sign * self.mm(pat, x, y)
            return self.off(-s * self.dEinv[..., None, None, None])
        return self.get("ddD", make)

    def der2(self, name):
        """second generalised derivative of a stored matrix (Der2A / Der2B / Der2O / Der2H / Der2Spin,
        covariant.py:25-121, 345-372): indices (b..., d, e)"""
        def make():
            T, dT = self.X(name, 0), self.cov(name, gender=1)
            res = self.gender(self.X(name, 1), self.X(name, 2))
            res = res - self.mm("mlde,ln...->mn...de", self.dD, T) - self.mm("mld,ln...e->mn...de", self.Dx, dT)
            res = res + self.mm("ml...,lnde->mn...de", T, self.dD) + self.mm("ml...e,lnd->mn...de", dT, self.Dx)
            return res
        return self.get(("der2", name), make)

    # ---- formulae: the PP block of the returned tensor is the reference's nn(ik, inn, out)
    def nn(self, T):
        return self.blk(T, "n", "n")

    def Omega(self):
        """Berry curvature (covariant.py:161-203)"""
        def make():
            Dnl, Dln = self.blk(self.Dx, "n", "l"), self.blk(self.Dx, "l", "n")
            s = 0.
            if self.internal:
                s = s - 1j * self.mm("mlc,lnc->mnc", Dnl[..., AL], Dln[..., BE])
            if self.external:
                A = self.X("AA")
                Aln, Ann = self.blk(A, "l", "n"), self.nn(A)
                s = s + 0.5 * self.nn(self.X("rotAA"))
                s = s - self.mm("mlc,lnc->mnc", Dnl[..., AL], Aln[..., BE]) + self.mm("mlc,lnc->mnc", Dnl[..., BE], Aln[..., AL])
                s = s - 1j * self.mm("mlc,lnc->mnc", Ann[..., AL], Ann[..., BE])
            return s + self.hc(s)
        return self.get("Omega", make)

    def DerOmega(self):
        """generalised derivative of the Berry curvature (covariant.py:212-259)"""
        def make():
            Dnl, dD = self.blk(self.Dx, "n", "l"), self.dD
            dDln, dDnl = self.blk(dD, "l", "n"), self.blk(dD, "n", "l")
            s = 0.
            if self.external:
                A, dA = self.X("AA"), self.cov("AA", gender=1)
                Aln, Ann, dAln, dAnn = self.blk(A, "l", "n"), self.nn(A), self.blk(dA, "l", "n"), self.nn(dA)
                s = s + 0.5 * self.nn(self.cov("rotAA", gender=1))
            for sg, a, b in ((+1, AL, BE), (-1, BE, AL)):
                if self.internal:
                    s = s - 1j * sg * self.mm("mlc,lncd->mncd", Dnl[..., a], dDln[:, :, :, b])
                if self.external:
                    s = s - sg * self.mm("mlc,lncd->mncd", Dnl[..., a], dAln[:, :, :, b])
                    s = s - sg * self.mm("mlcd,lnc->mncd", dDnl[:, :, :, a], Aln[..., b])
                    s = s - 1j * sg * self.mm("mlc,lncd->mncd", Ann[..., a], dAnn[:, :, :, b])
            return s + self.hc(s)
        return self.get("DerOmega", make)

    def Der2Omega(self):
        """second generalised derivative of the Berry curvature (covariant.py:267-312)"""
        def make():
            D, dD, ddD = self.Dx, self.dD, self.ddD
            Dnl, dDnl, dDln = self.blk(D, "n", "l"), self.blk(dD, "n", "l"), self.blk(dD, "l", "n")
            ddDnl, ddDln = self.blk(ddD, "n", "l"), self.blk(ddD, "l", "n")
            s = 0.
            if self.external:
                A, dA, ddA = self.X("AA"), self.cov("AA", gender=1), self.der2("AA")
                Aln, Ann = self.blk(A, "l", "n"), self.nn(A)
                dAln, dAnn = self.blk(dA, "l", "n"), self.nn(dA)
                ddAln, ddAnn = self.blk(ddA, "l", "n"), self.nn(ddA)
                s = s + 0.5 * self.nn(self.der2("rotAA"))
            for sg, a, b in ((+1, AL, BE), (-1, BE, AL)):
                if self.internal:
                    s = s - 1j * sg * self.mm("mlce,lncd->mncde", dDnl[:, :, :, a], dDln[:, :, :, b])
                    s = s - 1j * sg * self.mm("mlc,lncde->mncde", Dnl[..., a], ddDln[:, :, :, b])
                if self.external:
                    s = s - sg * self.mm("mlce,lncd->mncde", dDnl[:, :, :, a], dAln[:, :, :, b])
                    s = s - sg * self.mm("mlc,lncde->mncde", Dnl[..., a], ddAln[:, :, :, b])
                    s = s - sg * self.mm("mlcde,lnc->mncde", ddDnl[:, :, :, a], Aln[..., b])
                    s = s - sg * self.mm("mlcd,lnce->mncde", dDnl[:, :, :, a], dAln[:, :, :, b])
                    s = s - 1j * sg * self.mm("mlce,lncd->mncde", dAnn[:, :, :, a], dAnn[:, :, :, b])
                    s = s - 1j * sg * self.mm("mlc,lncde->mncde", Ann[..., a], ddAnn[:, :, :, b])
            return s + self.hc(s)
        return self.get("Der2Omega", make)

    def Der3E(self):
        """third derivative of the band energies (covariant.py:126-151)"""
        def make():
            V, dV, D, dD = self.V, self.dV, self.Dx, self.dD
            s = self.nn(self.dW)
            s = s + self.mm("mlac,lnb->mnabc", self.blk(dV, "n", "l"), self.blk(D, "l", "n"))
            s = s + self.mm("mla,lnbc->mnabc", self.blk(V, "n", "l"), self.blk(dD, "l", "n"))
            s = s - self.mm("mlbc,lna->mnabc", self.blk(dD, "n", "l"), self.blk(V, "l", "n"))
            s = s - self.mm("mlb,lnac->mnabc", self.blk(D, "n", "l"), self.blk(dV, "l", "n"))
            return s
        return self.get("Der3E", make)

    def Morb_H(self):
        """<d_a u| H |d_b u> antisymmetrised (covariant.py:375-421)"""
        def make():
            E = self.E.to(self.t.complex128)
            Dnl, Dln = self.blk(self.Dx, "n", "l"), self.blk(self.Dx, "l", "n")
            s = 0.
            if self.internal:
                s = s - 1j * self.mm("mlc,lnc->mnc", Dnl[..., AL] * E[:, None, :, None], Dln[..., BE])
            if self.external:
                A, B = self.X("AA"), self.X("BB")
                Bln, Ann = self.blk(B, "l", "n"), self.nn(A)
                s = s + 0.5 * self.nn(self.X("CC"))
                s = s - self.mm("mlc,lnc->mnc", Dnl[..., AL], Bln[..., BE]) + self.mm("mlc,lnc->mnc", Dnl[..., BE], Bln[..., AL])
                s = s - 1j * self.mm("mlc,lnc->mnc", Ann[..., AL] * E[:, None, :, None], Ann[..., BE])
            return s + self.hc(s)
        return self.get("Morb_H", make)

    @property
    def Eav(self):
        return self.get("Eav", lambda: 0.5 * (self.E[:, :, None] + self.E[:, None, :]))

    def Hplus(self):
        """Morb_H + (E_n + E_m) / 2 Omega (covariant.py:424-449, sign = +1)"""
        return self.get("Hplus", lambda: self.Morb_H() + self.Eav[..., None] * self.Omega())

    def Der2Morb_H(self):
        """second generalised derivative of Morb_H (covariant.py:566-641)"""
        def make():
            E = self.E.to(self.t.complex128)
            El2, El3 = E[:, :, None, None, None], E[:, :, None, None, None, None]   # energy of the ROW index of `ln` operands
            D, dD, ddD, V, dV = self.Dx, self.dD, self.ddD, self.V, self.dV
            Dnl, Dln = self.blk(D, "n", "l"), self.blk(D, "l", "n")
            dDnl, dDln, ddDln = self.blk(dD, "n", "l"), self.blk(dD, "l", "n"), self.blk(ddD, "l", "n")
            Vll, dVll, Vnn, dVnn = self.blk(V, "l", "l"), self.blk(dV, "l", "l"), self.nn(V), self.nn(dV)
            s = 0.
            if self.internal:
                s = s - 2j * self.mm("mpc,plde,lnc->mncde", Dnl[..., AL], dVll, Dln[..., BE])
                for sg, a, b in ((+1, AL, BE), (-1, BE, AL)):
                    s = s - 1j * sg * self.mm("mpce,pld,lnc->mncde", dDnl[:, :, :, a], Vll, Dln[..., b])
                    s = s - 1j * sg * self.mm("mpc,pld,lnce->mncde", Dnl[..., a], Vll, dDln[:, :, :, b])
                    s = s - 2j * sg * self.mm("mlce,lncd->mncde", dDnl[:, :, :, a], El2 * dDln[:, :, :, b])
                    s = s - 2j * sg * self.mm("mlc,lncde->mncde", Dnl[..., a], El3 * ddDln[:, :, :, b])
                    s = s - 2j * sg * self.mm("mpc,ple,lncd->mncde", Dnl[..., a], Vll, dDln[:, :, :, b])
            if self.external:
                A, dA, ddA = self.X("AA"), self.cov("AA", gender=1), self.der2("AA")
                B, dB, ddB = self.X("BB"), self.cov("BB", gender=1), self.der2("BB")
                Ann, dAnn, ddAnn = self.nn(A), self.nn(dA), self.nn(ddA)
                Bln, dBln, ddBln = self.blk(B, "l", "n"), self.blk(dB, "l", "n"), self.blk(ddB, "l", "n")
                Ec = E[:, None, :, None]             # energy of the COLUMN index of an `nn`-type operand
                s = s + self.nn(self.der2("CC"))
                s = s - 2j * self.mm("mpc,plde,lnc->mncde", Ann[..., AL], dVnn, Ann[..., BE])
                for sg, a, b in ((+1, AL, BE), (-1, BE, AL)):
                    s = s - 1j * sg * self.mm("mpce,pld,lnc->mncde", dAnn[:, :, :, a], Vnn, Ann[..., b])
                    s = s - 1j * sg * self.mm("mpc,pld,lnce->mncde", Ann[..., a], Vnn, dAnn[:, :, :, b])
                    s = s - 2j * sg * self.mm("mlce,lncd->mncde", dAnn[:, :, :, a] * Ec[..., None], dAnn[:, :, :, b])
                    s = s - 2j * sg * self.mm("mlc,lncde->mncde", Ann[..., a] * Ec, ddAnn[:, :, :, b])
                    s = s - 2j * sg * self.mm("mlc,lpe,pncd->mncde", Ann[..., a], Vnn, dAnn[:, :, :, b])
                    s = s - 2 * sg * self.mm("mlce,lncd->mncde", dDnl[:, :, :, a], dBln[:, :, :, b])
                    s = s - 2 * sg * self.mm("mlc,lncde->mncde", Dnl[..., a], ddBln[:, :, :, b])
                    s = s - 2 * sg * self.mm("mlce,lncd->mncde", self.hc(dBln[:, :, :, a]), dDln[:, :, :, b])
                    s = s - 2 * sg * self.mm("mlc,lncde->mncde", self.hc(Bln[..., a]), ddDln[:, :, :, b])
            return 0.5 * (s + self.hc(s))
        return self.get("Der2Morb_H", make)

    def Der2Hplus(self):
        """Der2Morb with sign = +1 (covariant.py:644-683)"""
        def make():
            V, dV, O, dO, ddO = self.nn(self.V), self.nn(self.dV), self.nn(self.Omega()), self.nn(self.DerOmega()), \
                self.nn(self.Der2Omega())
            t = self.Eav[..., None, None, None] * ddO
            t = t + 0.5 * (self.mm("mlce,lnd->mncde", dO, V) + self.mm("mlcd,lne->mncde", dO, V)
                           + self.mm("mld,lnce->mncde", V, dO) + self.mm("mle,lncd->mncde", V, dO)
                           + self.mm("mlc,lnde->mncde", O, dV) + self.mm("mlde,lnc->mncde", dV, O))
            return self.Der2Morb_H() + 0.5 * (t + self.hc(t))
        return self.get("Der2Hplus", make)

    def tildeFab(self):
        """<d_a u|(1 - P)|d_b u> with FF = rotAAab (basic.py:19-51)"""
        def make():
            Dnl, Dln = self.blk(self.Dx, "n", "l"), self.blk(self.Dx, "l", "n")
            s = 0.
            if self.internal:
                s = s - self.mm("mla,lnb->mnab", Dnl, Dln)
            if self.external:
                A = self.X("AA")
                s = s + self.nn(self.X("rotAAab")) + 2j * self.mm("mla,lnb->mnab", Dnl, self.blk(A, "l", "n"))
                s = s - self.mm("mla,lnb->mnab", self.nn(A), self.nn(A))
            return 0.5 * (s + s.permute(0, 2, 1, 4, 3).conj())
        return self.get("tildeFab", make)

    def tildeFab_d(self):
        """generalised derivative of tildeFab (basic.py:59-95)"""
        def make():
            Dnl, dDln = self.blk(self.Dx, "n", "l"), self.blk(self.dD, "l", "n")
            s = 0.
            if self.internal:
                s = s - 2 * self.mm("mla,lnbd->mnabd", Dnl, dDln)
            if self.external:
                A, dA = self.X("AA"), self.cov("AA", gender=1)
                s = s + self.nn(self.cov("rotAAab", gender=1))
                s = s + 2j * self.mm("mla,lnbd->mnabd", Dnl, self.blk(dA, "l", "n"))
                s = s + 2j * self.mm("mla,lnbd->mnabd", self.blk(A, "n", "l"), dDln)
                s = s - 2 * self.mm("mla,lnbd->mnabd", self.nn(A), self.nn(dA))
            return 0.5 * (s + s.permute(0, 2, 1, 4, 3, 5).conj())
        return self.get("tildeFab_d", make)

    # ---- traces over the inner set
    def tr(self, T):
        return self.t.einsum("znn...->z...", self.nn(T)).real

    def tr_prod(self, pattern, *ops):
        """trace over the inner set of a product of nn blocks (FormulaProduct, formula/formula.py:121-152)"""
        return self.mm(pattern, *[self.nn(o) for o in ops]).real


# what each formula needs to be traced; the results are `[z][3]^rank` real tensors
def _nldrude_z(alg, first, second):
    """FormulaSum([Der3E x M, Der2M x V], [-1, +1], ['apsu', 'uaps']) (covariant.py:893-914)"""
    t1 = alg.tr_prod("mnaps,nmu->apsu", alg.Der3E(), first)
    t2 = alg.tr_prod("mnuap,nms->apsu", second, alg.V)
    return t2 - t1


def trace_NLDrude_Z_spin(alg):
    return _nldrude_z(alg, alg.X("SS"), alg.der2("SS"))


def trace_NLDrude_Z_orb_Omega(alg):
    return _nldrude_z(alg, alg.Omega(), alg.Der2Omega())


def trace_NLDrude_Z_orb_Hplus(alg):
    return _nldrude_z(alg, alg.Hplus(), alg.Der2Hplus())


def trace_emcha_surf(alg):
    """covariant.py:868-890"""
    torch = alg.t
    f1 = alg.tr_prod("mlab,lpc,pmd->abcd", alg.dV, alg.Omega(), alg.V)       # InvMass x Omega x velocity
    f2 = alg.tr_prod("mla,lpbc,pmd->abcd", alg.V, alg.DerOmega(), alg.V)     # velocity x DerOmega x velocity
    tmp = f2 + torch.einsum("zapus->zaups", f1)
    delta = torch.eye(3, dtype=f1.dtype, device=f1.device)
    return (2 * tmp - 2 * torch.einsum("us,zabbp->zaups", delta, f1) - torch.einsum("au,zbbps->zaups", delta, tmp)
            + torch.einsum("us,zabpb->zaups", delta, f2) - f2)


def trace_QuantumMetric_ab(alg):
    f = alg.tr(alg.tildeFab())
    return 0.5 * (f + f.transpose(1, 2))


def trace_VelDQM(alg):
    f = alg.tildeFab_d()
    f = 0.5 * (f + f.transpose(3, 4))
    return alg.tr_prod("mla,lmbcd->abcd", alg.V, f)


TRACES = dict(NLDrude_Z_spin=(trace_NLDrude_Z_spin, 4, ("ident", "odd")),
              NLDrude_Z_orb_Omega=(trace_NLDrude_Z_orb_Omega, 4, ("ident", "odd")),
              NLDrude_Z_orb_Hplus=(trace_NLDrude_Z_orb_Hplus, 4, ("ident", "odd")),
              emcha_surf=(trace_emcha_surf, 4, ("ident", "odd")),
              QuantumMetric_ab=(trace_QuantumMetric_ab, 2, ("ident", "ident")),
              VelDQM=(trace_VelDQM, 4, ("ident", "ident")))


def batch_traces(name, data_K, kidx, a, b, internal=True, external=True, device=None, max_bytes=3e9):
    """traces of formula `name` over the band sets [a_i, b_i) of k-points kidx_i: `[len(kidx)][3]^rank` (numpy)"""
    torch = _torch()
    if device is None:   # (an explicit device is passed by the host tests only)
        if not torch.cuda.is_available():
            raise RuntimeError("the second-order formulae are evaluated on the GPU: no CUDA device (there is no CPU fallback)")
        dev_index = getattr(getattr(data_K, "engine", None), "device", None)
        device = torch.device("cuda", torch.cuda.current_device() if dev_index is None else int(dev_index))
    fn, rank, _ = TRACES[name]
    nw = data_K.num_wann
    n = len(kidx)
    out = np.zeros((n,) + (3,) * rank)
    # ~40 live rank-3 tensors of nw x nw x 27 complex128 per pair
    chunk = max(8, int(max_bytes / (40. * nw * nw * 27 * 16)))
    for i0 in range(0, n, chunk):
        sl = slice(i0, min(n, i0 + chunk))
        alg = BlockAlgebra(data_K, kidx[sl], a[sl], b[sl], device, internal, external)
        out[sl] = fn(alg).cpu().numpy()
        del alg
    return out
