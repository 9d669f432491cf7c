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
#include "wb_tma.cuh"

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
    int stage; // doubles per staging buffer (one channel triple, full matrices)
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
    d.stage = 6 * nw * nw;
    return d;
}

__host__ inline size_t wb_dmma_smem_bytes(int nw, int KS) {
    WbDmmaDims d = wb_dmma_dims(nw, KS);
    size_t dbl = 2 * (size_t)d.stage                 // staging ring: 2 channel triples
                 + 2 * 2 * (size_t)nw * nw           // U, double buffered
                 + (size_t)d.mrows * d.ldx           // Xs
                 + (size_t)3 * d.K2 * d.ldy          // Yp
                 + (size_t)12 * nw * d.ldc           // Cs: 6 matrices, planar re/im
                 + 2 * 3 * (size_t)nw                // Od
                 + 5 * (size_t)nw                    // Es, label, rows[3]
                 + 8;                                // 4 mbarriers
    return dbl * sizeof(double) + 2 * nw * sizeof(short) + (size_t)(nw * (nw + 1) / 2) * 2 * sizeof(short) + 32;
}

// One CTA walks k-points blockIdx.x, blockIdx.x + gridDim.x, ...  Per k-point the "items" are the channel
// triples dH | A | curl A (contiguous in the record).  A two-deep ring of staging buffers is filled by TMA
// bulk copies issued two items ahead, so the HBM/L2 latency of item i+2 hides behind the MMAs of items i, i+1.
template <int KS, int MT2>
__global__ void __launch_bounds__(128, 2)
wb_omega_events_dmma_kernel(const cplx* __restrict__ rec, WbLayout L, long nk, const double* __restrict__ Eall,
                            const cplx* __restrict__ Uall, WbWindow win, WbFormulaFlags fl,
                            double* __restrict__ ev_label, double* __restrict__ ev_val) {
    extern __shared__ __align__(16) double smem_d[];
    constexpr int NT = 128;
    const int nw = L.nw, n2 = nw * nw;
    const WbDmmaDims D = wb_dmma_dims(nw, KS);
    double* stg = smem_d;                              // [2][stage]
    cplx* Ustg = (cplx*)(stg + 2 * D.stage);           // [2][n2]
    double* Xs = (double*)(Ustg + 2 * n2);
    double* Yp = Xs + D.mrows * D.ldx;
    double* Cs = Yp + 3 * D.K2 * D.ldy;
    cplx* Od = (cplx*)(Cs + 12 * nw * D.ldc);
    double* Es = (double*)(Od + 3 * nw);
    double* label = Es + nw;
    double* rows = label + nw;
    uint64_t* bars = (uint64_t*)(rows + 3 * nw);       // [0..1] staging ring, [2..3] U
    short* g1 = (short*)(bars + 4);
    short* g2 = g1 + nw;
    short* lut = g2 + nw;                              // packed index -> (i, j)
    double* Rc = Xs;  // [3][n2] pair terms; aliases Xs, free after the rotations

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, q = lane & 3;
    // per-lane sign relating the A2 fragments (step 2) to the B1 fragments (step 1)
    const int signmask = ((g ^ q) & 1) ? (int)0x80000000 : 0;
    const int ntile3 = (3 * nw + 7) / 8;   // tiles over the stacked 3*nw dimension
    const int nitem = fl.external_terms ? 3 : 1;
    const uint32_t item_bytes[3] = {(uint32_t)(3 * (L.dH_herm ? L.ntri : n2) * 16), (uint32_t)(3 * L.ntri * 16),
                                    (uint32_t)(3 * L.ntri * 16)};
    const int item_off[3] = {L.off_dH[0], L.off_A[0], L.off_O[0]};

    // zero the padding that enters the K (inner) dimension once; X/Y loads never touch it
    for (int x = threadIdx.x; x < D.mrows * D.ldx; x += NT) Xs[x] = 0.;
    for (int x = threadIdx.x; x < 3 * D.K2 * D.ldy; x += NT) Yp[x] = 0.;
    for (int i = threadIdx.x; i < nw; i += NT)
        for (int j = i; j < nw; j++) {
            lut[2 * tri_index(i, j, nw)] = (short)i;
            lut[2 * tri_index(i, j, nw) + 1] = (short)j;
        }
    if (threadIdx.x == 0) {
        for (int b = 0; b < 4; b++) wb_mbar_init(&bars[b], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    const long nmine = (blockIdx.x < nk) ? (nk - 1 - blockIdx.x) / gridDim.x + 1 : 0;  // k-points of this CTA
    const long nitems_total = nmine * nitem;
    // producer (thread 0): issue the copy of global item `gi` (k-point gi / nitem, triple gi % nitem)
    auto issue_item = [&](long gi) {
        long ik = blockIdx.x + (gi / nitem) * (long)gridDim.x;
        int tr = (int)(gi % nitem);
        int b = (int)(gi & 1);
        wb_mbar_expect_tx(&bars[b], item_bytes[tr]);
        wb_bulk_g2s(stg + (size_t)b * D.stage, rec + ik * L.E + item_off[tr], item_bytes[tr], &bars[b]);
    };
    auto issue_U = [&](long kl) {  // kl = local k-point counter
        long ik = blockIdx.x + kl * (long)gridDim.x;
        int b = (int)(kl & 1);
        wb_mbar_expect_tx(&bars[2 + b], (uint32_t)(n2 * 16));
        wb_bulk_g2s(Ustg + (size_t)b * n2, Uall + ik * n2, (uint32_t)(n2 * 16), &bars[2 + b]);
    };
    if (threadIdx.x == 0 && nmine > 0) {
        issue_U(0);
        issue_item(0);
        if (nitems_total > 1) issue_item(1);
    }
    double Enext = 0.;
    if (nmine > 0 && threadIdx.x < nw) Enext = Eall[(long)blockIdx.x * nw + threadIdx.x];

    long gi = 0;
    for (long kl = 0; kl < nmine; kl++) {
        const long ik = blockIdx.x + kl * (long)gridDim.x;
        const cplx* Us = Ustg + (size_t)(kl & 1) * n2;
        if (threadIdx.x < nw) Es[threadIdx.x] = Enext;
        if (kl + 1 < nmine) {  // prefetch for the next k-point: E into registers, U by TMA
            if (threadIdx.x < nw) Enext = Eall[(ik + gridDim.x) * nw + threadIdx.x];
            if (threadIdx.x == 0) issue_U(kl + 1);
        }
        wb_mbar_wait(&bars[2 + (kl & 1)], (uint32_t)((kl >> 1) & 1));
        __syncthreads();
        if (nw <= 32) {
            if (threadIdx.x < 32) wb_band_groups_warp(Es, nw, win, g1, g2, label, threadIdx.x);
        } else if (threadIdx.x == 0) wb_band_groups(Es, nw, win, g1, g2, label);
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
                    const double* u = (const double*)&Us[j * nw + l];
                    int pk = kk & 1, pn = nn & 1;
                    v = (pk == pn) ? u[0] : (pk ? -u[1] : u[1]);
                }
                Bf[s][t] = v;
            }

        for (int trip = 0; trip < nitem; trip++, gi++) {
            const bool herm = trip > 0 || L.dH_herm;
            const int b = (int)(gi & 1);
            const cplx* src = (const cplx*)(stg + (size_t)b * D.stage);
            wb_mbar_wait(&bars[b], (uint32_t)((gi >> 1) & 1));
            // ---- unpack the three matrices of this channel into full real-ified rows
            if (!herm) {
                for (int x = threadIdx.x; x < 3 * n2; x += NT) {
                    int m = x / n2, e = x - m * n2;
                    int i = e / nw, j = e - i * nw;
                    *(double2*)&Xs[(m * nw + i) * D.ldx + 2 * j] = src[x];
                }
            } else {
                for (int x = threadIdx.x; x < 3 * L.ntri; x += NT) {
                    int m = x / L.ntri, e = x - m * L.ntri;
                    int i = lut[2 * e], j = lut[2 * e + 1];
                    cplx v = src[x];
                    *(double2*)&Xs[(m * nw + i) * D.ldx + 2 * j] = v;
                    if (i != j) *(double2*)&Xs[(m * nw + j) * D.ldx + 2 * i] = cconj(v);
                }
            }
            __syncthreads();
            // staging buffer b is free again: refill it with the item two ahead
            if (threadIdx.x == 0 && gi + 2 < nitems_total) issue_item(gi + 2);
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
                    int m = row / nw, i = row - m * nw;
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
                    int c = x / nw, n = x - c * nw;
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
                        double bv = bcol[4 * s * D.ldy];
#pragma unroll
                        for (int t = 0; t < MT2; t++) {
                            // sign flip on the high word (integer pipe, keeps the FP64 pipe for the MMAs)
                            double a = __hiloint2double(__double2hiint(Bf[s][t]) ^ signmask, __double2loint(Bf[s][t]));
                            wb_dmma(acc[t][0], acc[t][1], a, bv);
                        }
                    }
                    // lane holds Ctilde[rr = 8t + g][cc0 = 8nt + 2q + {0,1}]
                    int c0 = 8 * nt + 2 * q;
                    int m0 = c0 / nw, l0 = c0 - m0 * nw;
                    int c1 = c0 + 1;
                    int m1 = c1 / nw, l1 = c1 - m1 * nw;
#pragma unroll
                    for (int t = 0; t < MT2; t++) {
                        int rr = 8 * t + g;
                        int n = rr >> 1, p = rr & 1;
                        if (n < nw) {
                            if (c0 < 3 * nw) Cs[(((trip * 3 + m0) * 2 + p) * nw + n) * D.ldc + l0] = acc[t][0];
                            if (c1 < 3 * nw) Cs[(((trip * 3 + m1) * 2 + p) * nw + n) * D.ldc + l1] = acc[t][1];
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
