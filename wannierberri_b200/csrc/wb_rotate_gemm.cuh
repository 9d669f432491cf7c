// Eigenbasis rotation  Xbar = U^dagger X U  for ANY num_wann (<= 128) as a batched complex GEMM pair on the FP64
// tensor cores (mma.sync.m8n8k4.f64), results to global memory:  xbar[k][channel][nw][nw].
//
// Reference: Data_K._rotate (data_K/data_K.py:130-132) = einsum('kba,kbc...,kcd->kad...').
//
// This is the size-generic companion of wb_rotate_mma.cuh (compile-time num_wann, everything in shared memory):
// it serves num_wann > 20, the rank-2 Fermi-surface formulae and the Kubo path, whose rotated matrices do not fit
// one CTA's shared memory.  One CTA = (column panel of TN = 8 NTL columns, channel, k-point):
//   phase 1   Y[:, panel] = X U[:, panel]           K loop over chunks of KC, operands staged in shared memory;
//             a hermitian channel (upper triangle packed in the record) is expanded by the loader;
//   phase 2   C[:, panel] = U^H Y[:, panel]         Y never leaves shared memory.
// A complex product is four real DMMAs on the (re, im) components of complex128 fragments, which are read as
// 16-byte shared-memory loads; the row strides (KC + 4, TN + 2 complex) make every fragment load conflict free.
// Warp w owns the 8-row tiles w, w + 4 of every 64-row chunk.
#pragma once
#include "wb_common.cuh"
#include "wb_groups.cuh"

struct WbChanList {
    int n;
    int off[64];    // record offset (complex elements) of the matrix
    int herm[64];   // upper-triangle packed
};

__device__ __forceinline__ void wb_dmma_acc(double& d0, double& d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(d0), "+d"(d1)
        : "d"(a), "d"(b));
}

template <int NTL, int KC>
__host__ __device__ constexpr int wb_gemm_ldp() { return 8 * NTL + 2; }

template <int NTL, int KC>
__host__ inline size_t wb_gemm_smem_bytes(int nw) {
    int nwp = (nw + KC - 1) / KC * KC;
    return sizeof(cplx) * ((size_t)nwp * wb_gemm_ldp<NTL, KC>() + 64 * (KC + 4) + (size_t)KC * wb_gemm_ldp<NTL, KC>());
}

template <int NTL, int KC>
__global__ void __launch_bounds__(128)
wb_rotate_gemm_kernel(const cplx* __restrict__ rec, long recE, WbChanList ch, int nw, long nk,
                      const cplx* __restrict__ Uall, cplx* __restrict__ xbar, const int2* __restrict__ colwin) {
    constexpr int TN = 8 * NTL, LDP = wb_gemm_ldp<NTL, KC>(), LDA = KC + 4;
    extern __shared__ __align__(16) cplx smem_r[];
    const int nwp = (nw + KC - 1) / KC * KC;
    cplx* Yp = smem_r;                      // [nwp][LDP]
    cplx* As = Yp + (size_t)nwp * LDP;      // [64][LDA]
    cplx* Bs = As + 64 * LDA;               // [KC][LDP]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    const int l0 = blockIdx.x * TN;
    const int ic = blockIdx.y;
    const int off = ch.off[ic];
    const bool herm = ch.herm[ic];
    const int n2 = nw * nw;

    for (long ik = blockIdx.z; ik < nk; ik += gridDim.z) {
        const cplx* r = rec + ik * recE;
        const cplx* U = Uall + ik * n2;
        // column window (see wb_rotate_gemm_cg_kernel): panels without a band of a band group are skipped, the others
        // are formed whole; hermitian channels are mirrored into the rows of the window
        int c0 = 0, c1 = nw;
        if (colwin) {
            const int2 w = colwin[ik];
            c0 = w.x; c1 = w.y;
            if (l0 >= c1 || l0 + TN <= c0) continue;   // (uniform over the CTA)
        }
        for (int x = threadIdx.x; x < (nwp - nw) * LDP; x += 128) Yp[(size_t)nw * LDP + x] = cmake(0., 0.);
#pragma unroll 1
        for (int phase = 0; phase < 2; phase++) {
#pragma unroll 1
            for (int m0 = 0; m0 < nw; m0 += 64) {
                double are[2][NTL][2], aim[2][NTL][2];
#pragma unroll
                for (int i = 0; i < 2; i++)
#pragma unroll
                    for (int t = 0; t < NTL; t++) are[i][t][0] = are[i][t][1] = aim[i][t][0] = aim[i][t][1] = 0.;
                const bool act0 = m0 + 8 * warp < nw, act1 = m0 + 8 * (warp + 4) < nw;
#pragma unroll 1
                for (int j0 = 0; j0 < nw; j0 += KC) {
                    __syncthreads();
                    if (phase == 0) {
                        // As[r][c] = X[m0 + r][j0 + c];  Bs[r][c] = U[j0 + r][l0 + c]
                        for (int x = threadIdx.x; x < 64 * KC; x += 128) {
                            const int c = x % KC, rr = x / KC;
                            const int m = m0 + rr, j = j0 + c;
                            cplx v = cmake(0., 0.);
                            if (m < nw && j < nw) v = herm ? load_herm(r, off, m, j, nw) : r[off + m * nw + j];
                            As[rr * LDA + c] = v;
                        }
                        for (int x = threadIdx.x; x < KC * TN; x += 128) {
                            const int c = x % TN, rr = x / TN;
                            const int j = j0 + rr, l = l0 + c;
                            Bs[rr * LDP + c] = (j < nw && l < nw) ? U[j * nw + l] : cmake(0., 0.);
                        }
                    } else {
                        // As[r][c] = conj(U[j0 + c][m0 + r])
                        for (int x = threadIdx.x; x < 64 * KC; x += 128) {
                            const int rr = x % 64, c = x / 64;
                            const int n = m0 + rr, i = j0 + c;
                            cplx v = cmake(0., 0.);
                            if (n < nw && i < nw) v = cconj(U[i * nw + n]);
                            As[rr * LDA + c] = v;
                        }
                    }
                    __syncthreads();
                    const cplx* Bsrc = (phase == 0) ? Bs : Yp + (size_t)j0 * LDP;
#pragma unroll
                    for (int kk = 0; kk < KC; kk += 4) {
                        cplx b[NTL];
#pragma unroll
                        for (int t = 0; t < NTL; t++) b[t] = Bsrc[(kk + q) * LDP + 8 * t + g];
#pragma unroll
                        for (int i = 0; i < 2; i++) {
                            if (i == 0 ? act0 : act1) {
                                const cplx a = As[(8 * (warp + 4 * i) + g) * LDA + kk + q];
#pragma unroll
                                for (int t = 0; t < NTL; t++) {
                                    wb_dmma_acc(are[i][t][0], are[i][t][1], a.x, b[t].x);
                                    wb_dmma_acc(are[i][t][0], are[i][t][1], -a.y, b[t].y);
                                    wb_dmma_acc(aim[i][t][0], aim[i][t][1], a.x, b[t].y);
                                    wb_dmma_acc(aim[i][t][0], aim[i][t][1], a.y, b[t].x);
                                }
                            }
                        }
                    }
                }
                // epilogue: lane holds rows m0 + 8 (warp + 4 i) + g, columns 8 t + 2 q + {0, 1}
                if (phase == 0) {
#pragma unroll
                    for (int i = 0; i < 2; i++) {
                        const int m = m0 + 8 * (warp + 4 * i) + g;
                        if (m < nw) {
#pragma unroll
                            for (int t = 0; t < NTL; t++) {
                                Yp[(size_t)m * LDP + 8 * t + 2 * q] = cmake(are[i][t][0], aim[i][t][0]);
                                Yp[(size_t)m * LDP + 8 * t + 2 * q + 1] = cmake(are[i][t][1], aim[i][t][1]);
                            }
                        }
                    }
                } else {
                    cplx* out = xbar + ((size_t)ik * ch.n + ic) * n2;
#pragma unroll
                    for (int i = 0; i < 2; i++) {
                        const int n = m0 + 8 * (warp + 4 * i) + g;
                        if (n < nw) {
                            // mirror: row l of a hermitian channel for the bands n outside the computed panels
                            const bool mir = colwin && herm && (n < (c0 / TN) * TN || n >= ((c1 + TN - 1) / TN) * TN);
#pragma unroll
                            for (int t = 0; t < NTL; t++) {
                                const int l = l0 + 8 * t + 2 * q;
                                const cplx v0 = cmake(are[i][t][0], aim[i][t][0]), v1 = cmake(are[i][t][1], aim[i][t][1]);
                                if (l < nw) {
                                    out[n * nw + l] = v0;
                                    if (mir) out[l * nw + n] = cconj(v0);
                                }
                                if (l + 1 < nw) {
                                    out[n * nw + l + 1] = v1;
                                    if (mir) out[(l + 1) * nw + n] = cconj(v1);
                                }
                            }
                        }
                    }
                }
            }
            __syncthreads();   // Y panel complete (phase 0) / everybody done with it (phase 1)
        }
    }
}

// Small matrices (nw <= 8 NTL <= 32: one column panel holds the whole matrix): a 64-row MMA tile of the kernel above is
// mostly padding -- 24 of 64 rows at nw = 24 -- while every CTA pays the same staging steps and barriers.  Here a CTA
// rotates CG channels of a k-point at once: step 1 stacks their rows along M ([X_1; X_2; ..] U, CG nw <= 64 rows), step 2
// stacks the Y panels along N (U^dagger [Y_1 | Y_2 | ..]); U is staged once per CG channels.
template <int NTL, int KC, int CG>
__host__ __device__ constexpr int wb_gemm_cg_ldp() { return 8 * NTL * CG + 2; }

template <int NTL, int KC, int CG>
__host__ inline size_t wb_gemm_cg_smem_bytes(int nw) {
    int nwp = (nw + KC - 1) / KC * KC;
    return sizeof(cplx) * ((size_t)nwp * wb_gemm_cg_ldp<NTL, KC, CG>() + 64 * (KC + 4) + (size_t)KC * (8 * NTL + 2));
}

// Column window (TRIM): the formula stage of wb_events_xbar_kernel reads a rotated matrix only in ROWS and COLUMNS of
// the bands that belong to a band group of the k-point (X_{nl}, X_{ln}, X_{nn'} with n, n' in a group) -- a "cross" of the
// matrix; for Fermi-surface quantities that is 8-12 of the 24 bands of Te.  With `colwin[ik] = [c0, c1)` (the range of
// those bands, wb_band_window_kernel) the CTA multiplies by the columns c0 .. c1-1 of U only (NTA <= NTL column tiles,
// compile-time variants of the body), writes the columns of the result and, for a hermitian channel, their mirror
// images as the rows; everything outside the cross is left untouched and never read.  Non-hermitian channels (B; C
// inside a group) are read by columns only.
template <int NTL, int KC, int CG, int NTA>
__device__ __forceinline__ void wb_gemm_cg_kpoint(const cplx* __restrict__ r, const cplx* __restrict__ U, const WbChanList& ch,
                                                  int nw, int ic0, int ncg, int c0, int ncols, bool mirror, cplx* Yp, cplx* As,
                                                  cplx* Bs, cplx* __restrict__ outk) {
    constexpr int TN = 8 * NTL, LDP = wb_gemm_cg_ldp<NTL, KC, CG>(), LDA = KC + 4, LDB = TN + 2;
    const int nwp = (nw + KC - 1) / KC * KC;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    const int n2 = nw * nw, M = ncg * nw;
    __syncthreads();
    for (int x = threadIdx.x; x < nwp * LDP; x += 128) Yp[x] = cmake(0., 0.);
    // ---- step 1: Y_cc = X_cc U[:, c0 ..], rows of the channels stacked
#pragma unroll 1
    for (int m0 = 0; m0 < M; m0 += 64) {
        double are[2][NTA][2], aim[2][NTA][2];
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int t = 0; t < NTA; t++) are[i][t][0] = are[i][t][1] = aim[i][t][0] = aim[i][t][1] = 0.;
        const bool act0 = m0 + 8 * warp < M, act1 = m0 + 8 * (warp + 4) < M;
#pragma unroll 1
        for (int j0 = 0; j0 < nw; j0 += KC) {
            __syncthreads();
            for (int x = threadIdx.x; x < 64 * KC; x += 128) {
                const int c = x % KC, rr = x / KC;
                const int R = m0 + rr, j = j0 + c;
                cplx v = cmake(0., 0.);
                if (R < M && j < nw) {
                    const int cc = R / nw, m = R - cc * nw;
                    const int off = ch.off[ic0 + cc];
                    v = ch.herm[ic0 + cc] ? load_herm(r, off, m, j, nw) : r[off + m * nw + j];
                }
                As[rr * LDA + c] = v;
            }
            for (int x = threadIdx.x; x < KC * 8 * NTA; x += 128) {
                const int c = x % (8 * NTA), rr = x / (8 * NTA);
                const int j = j0 + rr;
                Bs[rr * LDB + c] = (j < nw && c0 + c < nw) ? U[j * nw + c0 + c] : cmake(0., 0.);
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < KC; kk += 4) {
                cplx b[NTA];
#pragma unroll
                for (int t = 0; t < NTA; t++) b[t] = Bs[(kk + q) * LDB + 8 * t + g];
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    if (i == 0 ? act0 : act1) {
                        const cplx a = As[(8 * (warp + 4 * i) + g) * LDA + kk + q];
#pragma unroll
                        for (int t = 0; t < NTA; t++) {
                            wb_dmma_acc(are[i][t][0], are[i][t][1], a.x, b[t].x);
                            wb_dmma_acc(are[i][t][0], are[i][t][1], -a.y, b[t].y);
                            wb_dmma_acc(aim[i][t][0], aim[i][t][1], a.x, b[t].y);
                            wb_dmma_acc(aim[i][t][0], aim[i][t][1], a.y, b[t].x);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const int R = m0 + 8 * (warp + 4 * i) + g;
            if (R < M) {
                const int cc = R / nw, m = R - cc * nw;
                cplx* yrow = Yp + (size_t)m * LDP + cc * TN;
#pragma unroll
                for (int t = 0; t < NTA; t++) {
                    yrow[8 * t + 2 * q] = cmake(are[i][t][0], aim[i][t][0]);
                    yrow[8 * t + 2 * q + 1] = cmake(are[i][t][1], aim[i][t][1]);
                }
            }
        }
    }
    // ---- step 2: C_cc = U^dagger Y_cc, the Y panels stacked along N; the nw <= 32 rows n are one 8-row tile per warp
    {
        double are[NTA * CG][2], aim[NTA * CG][2];
#pragma unroll
        for (int t = 0; t < NTA * CG; t++) are[t][0] = are[t][1] = aim[t][0] = aim[t][1] = 0.;
        const bool act = 8 * warp < nw;
#pragma unroll 1
        for (int j0 = 0; j0 < nw; j0 += KC) {
            __syncthreads();   // (first pass: Y complete)
            for (int x = threadIdx.x; x < 32 * KC; x += 128) {
                const int rr = x % 32, c = x / 32;
                const int i = j0 + c;
                As[rr * LDA + c] = (rr < nw && i < nw) ? cconj(U[i * nw + rr]) : cmake(0., 0.);
            }
            __syncthreads();
            const cplx* Bsrc = Yp + (size_t)j0 * LDP;
            if (act) {
#pragma unroll
                for (int kk = 0; kk < KC; kk += 4) {
                    const cplx a = As[(8 * warp + g) * LDA + kk + q];
#pragma unroll
                    for (int t = 0; t < NTA * CG; t++) {
                        const cplx bt = Bsrc[(kk + q) * LDP + (t / NTA) * TN + 8 * (t % NTA) + g];
                        wb_dmma_acc(are[t][0], are[t][1], a.x, bt.x);
                        wb_dmma_acc(are[t][0], are[t][1], -a.y, bt.y);
                        wb_dmma_acc(aim[t][0], aim[t][1], a.x, bt.y);
                        wb_dmma_acc(aim[t][0], aim[t][1], a.y, bt.x);
                    }
                }
            }
        }
        const int n = 8 * warp + g;
        if (n < nw) {
            const bool n_outside = (n < c0) || (n >= c0 + ncols);
#pragma unroll
            for (int t = 0; t < NTA * CG; t++) {
                const int cc = t / NTA, l = 8 * (t - cc * NTA) + 2 * q;
                if (cc < ncg) {
                    cplx* outm = outk + (size_t)(ic0 + cc) * n2;
                    const bool mir = mirror && n_outside && ch.herm[ic0 + cc];
                    const cplx v0 = cmake(are[t][0], aim[t][0]), v1 = cmake(are[t][1], aim[t][1]);
                    if (l < ncols) {
                        outm[(size_t)n * nw + c0 + l] = v0;
                        if (mir) outm[(size_t)(c0 + l) * nw + n] = cconj(v0);
                    }
                    if (l + 1 < ncols) {
                        outm[(size_t)n * nw + c0 + l + 1] = v1;
                        if (mir) outm[(size_t)(c0 + l + 1) * nw + n] = cconj(v1);
                    }
                }
            }
        }
    }
}

template <int NTL, int KC, int CG>
__global__ void __launch_bounds__(128)
wb_rotate_gemm_cg_kernel(const cplx* __restrict__ rec, long recE, WbChanList ch, int nw, long nk,
                         const cplx* __restrict__ Uall, cplx* __restrict__ xbar, const int2* __restrict__ colwin) {
    constexpr int LDP = wb_gemm_cg_ldp<NTL, KC, CG>(), LDA = KC + 4;
    extern __shared__ __align__(16) cplx smem_rc[];
    const int nwp = (nw + KC - 1) / KC * KC;
    cplx* Yp = smem_rc;                     // [nwp][CG * TN (+2)]: Y of channel cc in columns cc TN ..
    cplx* As = Yp + (size_t)nwp * LDP;      // [64][LDA]
    cplx* Bs = As + 64 * LDA;               // [KC][LDB]
    const int ic0 = blockIdx.y * CG;
    const int ncg = (ch.n - ic0 < CG) ? (ch.n - ic0) : CG;   // channels of this CTA
    const int n2 = nw * nw;
    for (long ik = blockIdx.z; ik < nk; ik += gridDim.z) {
        const cplx* r = rec + ik * recE;
        const cplx* U = Uall + ik * n2;
        cplx* outk = xbar + (size_t)ik * ch.n * n2;
        int c0 = 0, ncols = nw;
        if (colwin) {
            const int2 w = colwin[ik];
            c0 = w.x;
            ncols = w.y - w.x;
            if (ncols <= 0) continue;   // no band group at this k-point: nothing is read (uniform over the CTA)
        }
        const int nta = (ncols + 7) >> 3;
        const bool mirror = colwin != nullptr;
        if (nta >= NTL) wb_gemm_cg_kpoint<NTL, KC, CG, NTL>(r, U, ch, nw, ic0, ncg, c0, ncols, mirror, Yp, As, Bs, outk);
        else if (nta == 1) wb_gemm_cg_kpoint<NTL, KC, CG, 1>(r, U, ch, nw, ic0, ncg, c0, ncols, mirror, Yp, As, Bs, outk);
        else if (nta == 2) wb_gemm_cg_kpoint<NTL, KC, CG, (NTL > 2 ? 2 : NTL)>(r, U, ch, nw, ic0, ncg, c0, ncols, mirror, Yp, As, Bs, outk);
        else wb_gemm_cg_kpoint<NTL, KC, CG, (NTL > 3 ? 3 : NTL)>(r, U, ch, nw, ic0, ncg, c0, ncols, mirror, Yp, As, Bs, outk);
    }
}

// [c0, c1) of wb_rotate_gemm_cg_kernel's column window: the range of the bands of k-point ik that belong to a band group
// (the groups of the formula stage: same functions, same window).  One warp per k-point.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
wb_band_window_kernel(const double* __restrict__ Eall, int nw, long nk, WbWindow win, int2* __restrict__ colwin) {
    extern __shared__ __align__(16) double smem_w[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* Es = smem_w + (size_t)warp * (2 * nw + (nw + 3) / 4 * 2);
    double* label = Es + nw;
    short* g1 = (short*)(label + nw);
    short* g2 = g1 + nw;
    for (long ik = (long)blockIdx.x * WARPS + warp; ik < nk; ik += (long)gridDim.x * WARPS) {
        __syncwarp();
        for (int x = lane; x < nw; x += 32) Es[x] = Eall[ik * nw + x];
        __syncwarp();
        if (win.Ebmin) {
            if (lane == 0) wb_band_groups_tetra(Es, win.Ebmin + ik * nw, win.Ebmax + ik * nw, nw, win, g1, g2, label);
        } else if (nw <= 32) wb_band_groups_warp(Es, nw, win, g1, g2, label, lane);
        else if (lane == 0) wb_band_groups(Es, nw, win, g1, g2, label);
        __syncwarp();
        int lo = nw, hi = 0;
        for (int m = lane; m < nw; m += 32)
            if (g1[m] >= 0) { lo = min(lo, m); hi = max(hi, m + 1); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (lane == 0) colwin[ik] = make_int2(lo, hi);
    }
}

// Formula stage on rotated matrices held in global memory: one CTA per k-point.
// xbar[k][nch][nw][nw]; the channel triples are ordered V | A | B | O | C | S (those present) and are read in place
// (L2): staging them in shared memory was measured and does not pay, the kernel is bound by its sums, not by latency.
template <int NT>
__global__ void __launch_bounds__(NT, 8)
wb_events_xbar_kernel(const cplx* __restrict__ xbar, int nch, int nw, long nk, const double* __restrict__ Eall, WbWindow win,
                      WbEventLayout ev, double* __restrict__ mx_scratch, double* __restrict__ ev_label,
                      double* __restrict__ ev_val) {
    extern __shared__ __align__(16) double smem_x[];
    const int n2 = nw * nw;
    double* const small = smem_x;
    WbNeeds need = wb_needs(ev.mask, ev.external_terms);
    // everything that is needed at all is rotated in full here
    need.Oblk = need.Oblk || need.Odiag;
    need.Cblk = need.Cblk || need.Cdiag;
    need.Sblk = need.Sblk || need.Sdiag;
    need.Odiag = need.Cdiag = need.Sdiag = false;
    need.Wblk = need.Wdiag;
    need.Wdiag = false;
    double* Es = small;
    double* label = Es + nw;
    double* rows = label + nw;
    double* prod = rows + 3 * nw;
    double* Tedge = prod + 45 * nw;
    double* invtab = Tedge + 3 * (nw + 1);   // [nw][nw]
    short* g1 = (short*)(invtab + n2);
    short* g2 = g1 + nw;
    for (long ik = blockIdx.x; ik < nk; ik += gridDim.x) {
        for (int x = threadIdx.x; x < nw; x += NT) Es[x] = Eall[ik * nw + x];
        __syncthreads();
        if (win.Ebmin) {
            if (threadIdx.x == 0) wb_band_groups_tetra(Es, win.Ebmin + ik * nw, win.Ebmax + ik * nw, nw, win, g1, g2, label);
        } else if (nw <= 32) {
            if (threadIdx.x < 32) wb_band_groups_warp(Es, nw, win, g1, g2, label, threadIdx.x);
        } else if (threadIdx.x == 0) wb_band_groups(Es, nw, win, g1, g2, label);
        __syncthreads();
        const cplx* gsrc = xbar + (size_t)ik * nch * n2;
        int it = 0;
        auto next = [&](bool present) {   // triple `it` of this k-point
            const cplx* q = gsrc + (size_t)it * 3 * n2;
            if (present) it++;
            return q;
        };
        WbRotated R;
        R.Vb = next(need.V);
        R.Ab = next(need.A);
        R.Bb = next(need.B);
        R.Ob = next(need.Oblk);
        R.Cb = next(need.Cblk);
        R.Sb = next(need.Sblk);
        R.Wb = gsrc + (size_t)it * 3 * n2;   // six matrices after the triples
        R.Wd = nullptr;
        R.Od = R.Cd = R.Sd = nullptr;
        R.Es = Es; R.label = label; R.rows = rows; R.prod = prod; R.Tedge = Tedge;
        R.Mx = mx_scratch + (size_t)blockIdx.x * 3 * n2;
        R.inv = invtab;
        R.g1 = g1; R.g2 = g2;
        wb_formula_events<NT>(R, need, nw, ik, ev, ev_label, ev_val);
        __syncthreads();
    }
}

__host__ inline size_t wb_xbar_events_smem_bytes(int nw) {
    return sizeof(double) * ((size_t)2 * nw + 3 * nw + 45 * nw + 3 * (nw + 1) + (size_t)nw * nw) +
           2 * nw * sizeof(short) + 64;
}
