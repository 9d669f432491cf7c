// Eigenbasis rotation  Xbar = U^dagger X U  on the FP64 tensor cores (mma.sync.m8n8k4.f64, "DMMA")
// fused with the Berry-curvature formula, for 2*nw <= 40 (nw <= 20: bcc Fe, 18 WF).
//
// Reference: Data_K._rotate (data_K/data_K.py:130-132) = einsum('kba,kbc...,kcd->kad...') followed
// by Omega.nn / trace (formula/covariant.py:175-200).
//
// One CTA (4 warps) per k-point.  Complex GEMMs are real-ified with interleaved (re, im) indices
// so that the record's complex128 rows ARE the real A-operand rows:
//   step 1   Y = X U     [Yr Yi] (3nw x 2nw) = [Xr Xi] (3nw x 2nw) . B1 (2nw x 2nw),  three Cartesian
//            components of one channel stacked along M;
//   step 2   C = U^H Y   [Cr; Ci] (2nw x 3nw) = A2 (2nw x 2nw) . [Yr; Yi] (2nw x 3nw), stacked along N.
// The fragments of B1 (built from U) stay in REGISTERS for the whole k-point and serve both steps:
// the A2 fragment of tile (m, k) is the B1 fragment of tile (k, m) up to a per-lane sign.
// So every DMMA needs at most one shared-memory operand load per 5 MMAs.
// curl(A) needs only its rotated diagonal (trace), which is a dot product after step 1.
#pragma once
#include "wb_common.cuh"
#include "wb_groups.cuh"
#include "wb_rotate_formula.cuh"

__device__ __forceinline__ void wb_dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

struct WbDmmaDims {
    int K2;    // 4*KS  (real-ified, padded inner dimension)
    int ldx;   // leading dimension of Xs (doubles)
    int ldy;   // leading dimension of Yp (doubles)
    int ldc;   // leading dimension of Cs rows (doubles)
    int mrows; // 3*nw rounded up to 8
};

__host__ __device__ inline int wb_pad_ld(int n) {  // smallest ld >= n with ld % 16 in {4, 12}: conflict-free 64-bit fragment loads
    int ld = (n + 3) / 4 * 4;
    while (ld % 16 != 4 && ld % 16 != 12) ld += 4;
    return ld;
}

__host__ __device__ inline WbDmmaDims wb_dmma_dims(int nw, int KS) {
    WbDmmaDims d;
    d.K2 = 4 * KS;
    d.ldx = wb_pad_ld(d.K2);
    d.ldy = wb_pad_ld(nw);
    d.ldc = (nw + 1) / 2 * 2;
    d.mrows = (3 * nw + 7) / 8 * 8;
    return d;
}

__host__ inline size_t wb_dmma_smem_bytes(int nw, int KS) {
    WbDmmaDims d = wb_dmma_dims(nw, KS);
    size_t dbl = 2 * (size_t)nw * nw                 // Us
                 + (size_t)d.mrows * d.ldx           // Xs
                 + (size_t)3 * d.K2 * d.ldy          // Yp
                 + (size_t)12 * nw * d.ldc           // Cs: 6 matrices, planar re/im
                 + 2 * 3 * (size_t)nw                // Od
                 + 5 * (size_t)nw;                   // Es, label, rows[3]
    return dbl * sizeof(double) + 2 * nw * sizeof(short) + 16;
}

template <int KS, int MT2>
__global__ void __launch_bounds__(128)
wb_omega_events_dmma_kernel(const cplx* __restrict__ rec, WbLayout L, long nk, const double* __restrict__ Eall,
                            const cplx* __restrict__ Uall, WbWindow win, WbFormulaFlags fl,
                            double* __restrict__ ev_label, double* __restrict__ ev_val) {
    extern __shared__ double smem_d[];
    constexpr int NT = 128;
    const int nw = L.nw, n2 = nw * nw;
    const WbDmmaDims D = wb_dmma_dims(nw, KS);
    cplx* Us = (cplx*)smem_d;
    double* Xs = smem_d + 2 * n2;
    double* Yp = Xs + D.mrows * D.ldx;
    double* Cs = Yp + 3 * D.K2 * D.ldy;
    cplx* Od = (cplx*)(Cs + 12 * nw * D.ldc);
    double* Es = (double*)(Od + 3 * nw);
    double* label = Es + nw;
    double* rows = label + nw;
    short* g1 = (short*)(rows + 3 * nw);
    short* g2 = g1 + nw;
    double* Rc = Xs;  // [3][n2] pair terms; aliases Xs/Yp, free after the rotations

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, q = lane & 3;
    // per-lane sign relating the A2 fragments (step 2) to the B1 fragments (step 1)
    const int signmask = ((g ^ q) & 1) ? (int)0x80000000 : 0;
    const int ntile3 = (3 * nw + 7) / 8;   // tiles over the stacked 3*nw dimension

    // zero the padding that enters the K (inner) dimension once; X/Y loads never touch it
    for (int x = threadIdx.x; x < D.mrows * D.ldx; x += NT) Xs[x] = 0.;
    for (int x = threadIdx.x; x < 3 * D.K2 * D.ldy; x += NT) Yp[x] = 0.;
    __syncthreads();

    for (long ik = blockIdx.x; ik < nk; ik += gridDim.x) {
        const cplx* r = rec + ik * L.E;
        for (int x = threadIdx.x; x < n2; x += NT) Us[x] = Uall[ik * n2 + x];
        for (int x = threadIdx.x; x < nw; x += NT) Es[x] = Eall[ik * nw + x];
        __syncthreads();
        if (threadIdx.x == 0) wb_band_groups(Es, nw, win, g1, g2, label);
        // B1 fragments: element (kk = 4s + q, nn = 8t + g)
        double Bf[KS][MT2];
#pragma unroll
        for (int s = 0; s < KS; s++)
#pragma unroll
            for (int t = 0; t < MT2; t++) {
                int kk = 4 * s + q, nn = 8 * t + g;
                int j = kk >> 1, l = nn >> 1;
                double v = 0.;
                if (j < nw && l < nw) {
                    cplx u = Us[j * nw + l];
                    int pk = kk & 1, pn = nn & 1;
                    v = (pk == pn) ? u.x : (pk ? -u.y : u.y);
                }
                Bf[s][t] = v;
            }

        for (int trip = 0; trip < 3; trip++) {
            if (trip > 0 && !fl.external_terms) break;
            const int* offs = trip == 0 ? L.off_dH : (trip == 1 ? L.off_A : L.off_O);
            const bool herm = trip > 0;
            // ---- load the three matrices of this channel as full real-ified rows
            if (!herm) {
                for (int x = threadIdx.x; x < 3 * n2; x += NT) {
                    int m = x / n2, e = x % n2;
                    int i = e / nw, j = e % nw;
                    cplx v = r[offs[m] + e];
                    *(double2*)&Xs[(m * nw + i) * D.ldx + 2 * j] = v;
                }
            } else {
                for (int x = threadIdx.x; x < 3 * L.ntri; x += NT) {
                    int m = x / L.ntri, e = x % L.ntri;
                    // invert the packed index: row i such that tri_index(i, i) <= e
                    int i = 0;
                    while (i + 1 < nw && tri_index(i + 1, i + 1, nw) <= e) i++;
                    int j = i + (e - tri_index(i, i, nw));
                    cplx v = r[offs[m] + e];
                    *(double2*)&Xs[(m * nw + i) * D.ldx + 2 * j] = v;
                    if (i != j) *(double2*)&Xs[(m * nw + j) * D.ldx + 2 * i] = cconj(v);
                }
            }
            __syncthreads();
            // ---- step 1: Y = X U
            for (int mt = warp; mt < ntile3; mt += 4) {
                double acc[MT2][2];
#pragma unroll
                for (int t = 0; t < MT2; t++) acc[t][0] = acc[t][1] = 0.;
                const double* arow = Xs + (8 * mt + g) * D.ldx + q;
#pragma unroll
                for (int s = 0; s < KS; s++) {
                    double a = arow[4 * s];
#pragma unroll
                    for (int t = 0; t < MT2; t++) wb_dmma(acc[t][0], acc[t][1], a, Bf[s][t]);
                }
                int row = 8 * mt + g;
                if (row < 3 * nw) {
                    int m = row / nw, i = row % nw;
                    double* y0 = Yp + (m * D.K2 + 2 * i) * D.ldy;
#pragma unroll
                    for (int t = 0; t < MT2; t++) {
                        int l = 4 * t + q;
                        if (l < nw) { y0[l] = acc[t][0]; y0[D.ldy + l] = acc[t][1]; }
                    }
                }
            }
            __syncthreads();
            if (trip == 2) {
                // diagonal of Obar:  sum_i conj(U[i][n]) Y_c[i][n]
                for (int x = threadIdx.x; x < 3 * nw; x += NT) {
                    int c = x / nw, n = x % nw;
                    cplx acc = cmake(0., 0.);
                    const double* y = Yp + (c * D.K2) * D.ldy + n;
                    for (int i = 0; i < nw; i++)
                        cfma_conj(acc, Us[i * nw + n], cmake(y[(2 * i) * D.ldy], y[(2 * i + 1) * D.ldy]));
                    Od[x] = acc;
                }
            } else {
                // ---- step 2: C = U^H Y, stacked along N
                for (int nt = warp; nt < ntile3; nt += 4) {
                    double acc[MT2][2];
#pragma unroll
                    for (int t = 0; t < MT2; t++) acc[t][0] = acc[t][1] = 0.;
                    int cc = 8 * nt + g;  // stacked column of the B operand held by this lane
                    int mb = min(cc / nw, 2), lb = cc - mb * nw;
                    if (lb >= D.ldy) lb = 0;  // columns beyond 3*nw: discarded outputs, any finite input
                    const double* bcol = Yp + (mb * D.K2 + q) * D.ldy + lb;
#pragma unroll
                    for (int s = 0; s < KS; s++) {
                        double b = bcol[4 * s * D.ldy];
#pragma unroll
                        for (int t = 0; t < MT2; t++) {
                            // sign flip on the high word (integer pipe, keeps the FP64 pipe for the MMAs)
                            double a = __hiloint2double(__double2hiint(Bf[s][t]) ^ signmask, __double2loint(Bf[s][t]));
                            wb_dmma(acc[t][0], acc[t][1], a, b);
                        }
                    }
                    // lane holds Ctilde[rr = 8t + g][cc0 = 8nt + 2q + {0,1}]
#pragma unroll
                    for (int t = 0; t < MT2; t++) {
                        int rr = 8 * t + g;
                        int n = rr >> 1, p = rr & 1;
                        if (n < nw) {
#pragma unroll
                            for (int h = 0; h < 2; h++) {
                                int c0 = 8 * nt + 2 * q + h;
                                if (c0 < 3 * nw) {
                                    int m = c0 / nw, l = c0 % nw;
                                    Cs[(((trip * 3 + m) * 2 + p) * nw + n) * D.ldc + l] = acc[t][h];
                                }
                            }
                        }
                    }
                }
            }
            __syncthreads();
        }

        // ---- formula: pair terms Re[-i D_nl,a D_ln,b - D_nl,a A_ln,b + D_nl,b A_ln,a], n in a group, l outside
#define WB_V(a, n, l) cmake(Cs[((((a)) * 2 + 0) * nw + (n)) * D.ldc + (l)], Cs[((((a)) * 2 + 1) * nw + (n)) * D.ldc + (l)])
#define WB_A(a, n, l) WB_V(3 + (a), n, l)
        for (int x = threadIdx.x; x < n2; x += NT) {
            int n = x / nw, l = x % nw;
            double R[3] = {0., 0., 0.};
            if (g1[n] >= 0 && (l < g1[n] || l >= g2[n])) {
                double inv_nl = wb_deinv(Es[n], Es[l]);
                double inv_ln = wb_deinv(Es[l], Es[n]);
                cplx Dnl[3], Dln[3];
                for (int a = 0; a < 3; a++) {
                    Dnl[a] = cscale(-inv_nl, WB_V(a, n, l));
                    Dln[a] = cscale(-inv_ln, WB_V(a, l, n));
                }
                for (int c = 0; c < 3; c++) {
                    int al = WB_ALPHA(c), be = WB_BETA(c);
                    double v = 0.;
                    if (fl.internal_terms) v += cmul(Dnl[al], Dln[be]).y;
                    if (fl.external_terms) v += -cmul(Dnl[al], WB_A(be, l, n)).x + cmul(Dnl[be], WB_A(al, l, n)).x;
                    R[c] = v;
                }
            }
            for (int c = 0; c < 3; c++) Rc[c * n2 + x] = R[c];
        }
        __syncthreads();
        for (int x = threadIdx.x; x < 3 * nw; x += NT) {
            int c = x / nw, n = x % nw;
            double s = 0.;
            if (g1[n] >= 0) {
                for (int l = 0; l < nw; l++) s += Rc[c * n2 + n * nw + l];
                if (fl.external_terms) {
                    int al = WB_ALPHA(c), be = WB_BETA(c);
                    s += 0.5 * Od[c * nw + n].x;
                    for (int m = g1[n]; m < g2[n]; m++) s += cmul(WB_A(al, n, m), WB_A(be, m, n)).y;
                }
            }
            rows[x] = s;
        }
        __syncthreads();
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
        // Rc aliased Xs/Yp: restore the zero padding of the K dimension before the next k-point
        __syncthreads();
        for (int x = threadIdx.x; x < 3 * n2; x += NT) Rc[x] = 0.;
        __syncthreads();
#undef WB_V
#undef WB_A
    }
}
