"""numpy emulation of wb_eigh_jacobi_kernel (round-robin two-sided Jacobi) on PT-symmetric (Kramers-degenerate) matrices: the
sequence of off-diagonal norms per sweep -- the evidence behind the convergence rule of the kernel (one sweep after 1e-10 |A| is not
enough: 3e-12 -> 3e-21 -> 3e-25 -> 3e-37 was seen).   python profiles/jacobi_emulation.py"""
import numpy as np, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wannierberri_b200.system import kramers_system
s = kramers_system(6, seed=6)
iR = s.rvec.iRvec; HR = s._XX_R["Ham"]
def Hk(k):
    ph = np.exp(2j*np.pi*(iR@k)); H = np.einsum("r,rij->ij", ph, HR); return 0.5*(H+H.conj().T)
def rr_pair(n,s,m):
    a = n-1 if m==0 else (s+m)%(n-1); b=(s+(n-1)-m)%(n-1); return min(a,b),max(a,b)
def jacobi(H, newcrit=True):
    nw=H.shape[0]; A=H.copy(); U=np.eye(nw,dtype=complex); nrm2=(abs(A)**2).sum()
    last=False; offprev=nrm2; hist=[]
    for sweep in range(30):
        for st in range(nw-1):
            for m in range(nw//2):
                p,q=rr_pair(nw,st,m)
                b=A[p,q]; ab=abs(b); app=A[p,p].real; aqq=A[q,q].real
                if ab>1e-300 and ab>1e-18*(abs(app)+abs(aqq)):
                    tau=(aqq-app)/(2*ab); t=(1. if tau>=0 else -1.)/(abs(tau)+np.sqrt(1+tau*tau)); c=1/np.sqrt(1+t*t); sn=t*c; ph=b/ab
                    sph=sn*ph
                    for M in (A,U):
                        ap=M[:,p].copy(); aq=M[:,q].copy()
                        M[:,p]=c*ap-np.conj(sph)*aq; M[:,q]=sph*ap+c*aq
                    ap=A[p,:].copy(); aq=A[q,:].copy()
                    A[p,:]=c*ap-sph*aq; A[q,:]=np.conj(sph)*ap+c*aq
        if last: break
        off=(abs(np.triu(A,1))**2).sum(); hist.append(off/nrm2)
        if newcrit:
            if 2*off<=4e-30*nrm2 or (2*off<=1e-20*nrm2 and off>0.25*offprev): last=True
        elif 2*off<=1e-20*nrm2: last=True
        offprev=off
    E=np.diag(A).real
    return E,U,hist,A
rng=np.random.default_rng(0)
worst=0
for i in range(300):
    k=rng.random(3); H=Hk(k)
    E,U,hist,A=jacobi(H)
    r=abs(H@U-U*E).max()/abs(H).max()
    if r>worst: worst=r; print(i, r, ["%.1e"%h for h in hist], "imagdiag", abs(np.diag(A).imag).max())
print("grid")
worst=0
for ix in range(8):
  for iy in range(8):
    for iz in range(8):
        k=np.array([ix/8+0.03, iy/8+0.01, iz/8+0.2]); H=Hk(k)
        E,U,hist,A=jacobi(H)
        r=abs(H@U-U*E).max()/abs(H).max()
        if r>worst: worst=r; print(ix,iy,iz, r, ["%.1e"%h for h in hist])
