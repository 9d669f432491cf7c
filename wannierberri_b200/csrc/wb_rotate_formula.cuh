// Eigenbasis rotation  Xbar = U^dagger X U  fused with the covariant formulae, generic path
// (any nw whose working set fits shared memory): one CTA per k-point.
//
// Reference: Data_K._rotate (data_K/data_K.py:130-132), Data_K.dEig_inv / D_H (:290-298, :324-326),
// Omega.nn (formula/covariant.py:175-200), Formula_ln.trace (formula/formula.py:76-79) and the
// per-group evaluation loop of StaticCalculator.__call__ (calculators/static.py:102-108).
//
// Output = "events": for every k-point and every band group that the reference would put into
// its `weights` dict, the group label energy and the traced formula value.  Slot n of
// ev_label[k][nw] / ev_val[k][nw][ncomp] is used iff a group starts at band n.
#pragma once
#include "wb_common.cuh"
#include "wb_groups.cuh"

struct WbFormulaFlags {
    int formula;    // WBGPU_OMEGA, ...
    int internal_terms;
    int external_terms;
};

// 1/(E_a - E_b) with the reference's cutoff (data_K.py:292-297)
__device__ __forceinline__ double wb_deinv(double Ea, double Eb) {
    double d = Ea - Eb;
    return (fabs(d) < 1e-7) ? 0. : 1. / d;
}

// C = U^dagger (X U) for one full matrix X held in Xs; result to Cs.  All nw x nw, row-major.
template <int NT>
__device__ __forceinline__ void wb_rotate_smem(const cplx* Us, const cplx* Xs, cplx* Ys, cplx* Cs, int nw) {
    for (int x = threadIdx.x; x < nw * nw; x += NT) {
        int i = x / nw, l = x % nw;
        cplx acc = cmake(0., 0.);
        for (int j = 0; j < nw; j++) cfma(acc, Xs[i * nw + j], Us[j * nw + l]);
        Ys[x] = acc;
    }
    __syncthreads();
    for (int x = threadIdx.x; x < nw * nw; x += NT) {
        int n = x / nw, l = x % nw;
        cplx acc = cmake(0., 0.);
        for (int i = 0; i < nw; i++) cfma_conj(acc, Us[i * nw + n], Ys[i * nw + l]);
        Cs[x] = acc;
    }
    __syncthreads();
}

template <int NT>
__device__ __forceinline__ void wb_load_channel(const cplx* __restrict__ rec, int off, bool herm, cplx* Xs, int nw) {
    for (int x = threadIdx.x; x < nw * nw; x += NT) {
        int i = x / nw, j = x % nw;
        Xs[x] = herm ? load_herm(rec, off, i, j, nw) : rec[off + x];
    }
    __syncthreads();
}

template <int NT>
__global__ void __launch_bounds__(NT)
wb_omega_events_kernel(const cplx* __restrict__ rec, WbLayout L, long nk, const double* __restrict__ Eall,
                       const cplx* __restrict__ Uall, WbWindow win, WbFormulaFlags fl,
                       double* __restrict__ ev_label, double* __restrict__ ev_val) {
    extern __shared__ cplx smem_r[];
    const int nw = L.nw, n2 = nw * nw;
    cplx* Us = smem_r;
    cplx* Xs = Us + n2;
    cplx* Ys = Xs + n2;
    cplx* Vb = Ys + n2;          // [3][n2]
    cplx* Ab = Vb + 3 * n2;      // [3][n2]
    cplx* Od = Ab + 3 * n2;      // [3][nw]  diagonal of Obar
    double* Es = (double*)(Od + 3 * nw);
    double* label = Es + nw;
    double* rows = label + nw;   // [3][nw] per-band sums
    short* g1 = (short*)(rows + 3 * nw);
    short* g2 = g1 + nw;
    double* Rc = (double*)Xs;    // [3][n2] pair terms, aliases Xs/Ys (free after the rotations)

    for (long ik = blockIdx.x; ik < nk; ik += gridDim.x) {
        const cplx* r = rec + ik * L.E;
        for (int x = threadIdx.x; x < n2; x += NT) Us[x] = Uall[ik * n2 + x];
        for (int x = threadIdx.x; x < nw; x += NT) Es[x] = Eall[ik * nw + x];
        __syncthreads();
        if (nw <= 32) {
            if (threadIdx.x < 32) wb_band_groups_warp(Es, nw, win, g1, g2, label, threadIdx.x);
        } else if (threadIdx.x == 0) wb_band_groups(Es, nw, win, g1, g2, label);
        // rotations (the block-wide barriers inside also publish the groups)
        for (int a = 0; a < 3; a++) {
            wb_load_channel<NT>(r, L.off_dH[a], false, Xs, nw);
            wb_rotate_smem<NT>(Us, Xs, Ys, Vb + a * n2, nw);
        }
        if (fl.external_terms) {
            for (int a = 0; a < 3; a++) {
                wb_load_channel<NT>(r, L.off_A[a], true, Xs, nw);
                wb_rotate_smem<NT>(Us, Xs, Ys, Ab + a * n2, nw);
            }
            for (int c = 0; c < 3; c++) {  // only the diagonal of Obar enters the trace
                wb_load_channel<NT>(r, L.off_O[c], true, Xs, nw);
                for (int x = threadIdx.x; x < n2; x += NT) {
                    int i = x / nw, l = x % nw;
                    cplx acc = cmake(0., 0.);
                    for (int j = 0; j < nw; j++) cfma(acc, Xs[i * nw + j], Us[j * nw + l]);
                    Ys[x] = acc;
                }
                __syncthreads();
                for (int n = threadIdx.x; n < nw; n += NT) {
                    cplx acc = cmake(0., 0.);
                    for (int i = 0; i < nw; i++) cfma_conj(acc, Us[i * nw + n], Ys[i * nw + n]);
                    Od[c * nw + n] = acc;
                }
                __syncthreads();
            }
        }
        // pair terms  Re[ -i D_nl,a D_ln,b - D_nl,a A_ln,b + D_nl,b A_ln,a ],  n in a group, l outside it
        for (int x = threadIdx.x; x < n2; x += NT) {
            int n = x / nw, l = x % nw;
            double R[3] = {0., 0., 0.};
            if (g1[n] >= 0 && (l < g1[n] || l >= g2[n])) {
                double inv_nl = wb_deinv(Es[n], Es[l]);
                double inv_ln = wb_deinv(Es[l], Es[n]);
                cplx Dnl[3], Dln[3];
                for (int a = 0; a < 3; a++) {
                    Dnl[a] = cscale(-inv_nl, Vb[a * n2 + n * nw + l]);
                    Dln[a] = cscale(-inv_ln, Vb[a * n2 + l * nw + n]);
                }
                for (int c = 0; c < 3; c++) {
                    int al = WB_ALPHA(c), be = WB_BETA(c);
                    double v = 0.;
                    if (fl.internal_terms) {  // Re(-i z) = Im z
                        cplx z = cmul(Dnl[al], Dln[be]);
                        v += z.y;
                    }
                    if (fl.external_terms) {
                        cplx z1 = cmul(Dnl[al], Ab[be * n2 + l * nw + n]);
                        cplx z2 = cmul(Dnl[be], Ab[al * n2 + l * nw + n]);
                        v += -z1.x + z2.x;
                    }
                    R[c] = v;
                }
            }
            for (int c = 0; c < 3; c++) Rc[c * n2 + x] = R[c];
        }
        __syncthreads();
        // per band: sum over l, plus the in-group terms  1/2 O_nn - i sum_n' A_nn',a A_n'n,b
        for (int x = threadIdx.x; x < 3 * nw; x += NT) {
            int c = x / nw, n = x % nw;
            double s = 0.;
            if (g1[n] >= 0) {
                for (int l = 0; l < nw; l++) s += Rc[c * n2 + n * nw + l];
                if (fl.external_terms) {
                    int al = WB_ALPHA(c), be = WB_BETA(c);
                    s += 0.5 * Od[c * nw + n].x;
                    for (int m = g1[n]; m < g2[n]; m++) {
                        cplx z = cmul(Ab[al * n2 + n * nw + m], Ab[be * n2 + m * nw + n]);
                        s += z.y;  // Re(-i z)
                    }
                }
            }
            rows[x] = s;
        }
        __syncthreads();
        // events: trace over the group, "summ + summ^dagger" doubles the real part (covariant.py:199)
        for (int x = threadIdx.x; x < nw; x += NT) {
            double lab = label[x];
            ev_label[ik * nw + x] = lab;
            if (lab != CUDART_INF) {
                int b = g2[x];
                for (int c = 0; c < 3; c++) {
                    double s = 0.;
                    for (int n = x; n < b; n++) s += rows[c * nw + n];
                    ev_val[(ik * nw + x) * 3 + c] = 2. * s;
                }
            }
        }
        __syncthreads();
    }
}
