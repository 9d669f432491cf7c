// Fermi-sea formulae of next-row 4 on the rotated matrices of wb_rotate_gemm.cuh:
//   wb_spinomega_events_kernel   SpinOmega (spin Berry curvature, static.SHC; formula/covariant.py:759-789), rank 3
//   wb_deromega_events_kernel    DerOmega  (generalised derivative of the Berry curvature, BerryDipole_FermiSea /
//                                NLAHC_FermiSea; formula/covariant.py:212-259 with DerDcov, formula/elementary.py:56-72,
//                                and Matrix_GenDer_ln, formula/formula.py:95-118), rank 2
// One CTA per k-point; band groups and the "event" output (one slot per band that starts a group, values of the trace
// over the group) are those of wb_events_generic.cuh, so the Fermi scan / tetrahedron stages are shared.  Each of the
// two formulae is evaluated in an event pass of its own (WbEventLayout with a single formula at offset 0).
#pragma once
#include "wb_common.cuh"
#include "wb_groups.cuh"
#include "wb_rotate_formula.cuh"

// positions (in units of nw x nw matrices of the rotated record of one k-point) of the matrices DerOmega needs
struct WbDerOmegaChans {
    int iV, iA, iO, iW, idA, idO;   // d_a H [3] | A [3] | rotA [3] | d_b d_d H [6, wb_sym6] | d_d A_b [9, 3 b + d] | d_d rotA_c [9]
};

__host__ inline size_t wb_fsea_smem_bytes(int nw, int ncomp) {
    return sizeof(double) * (2 * (size_t)nw + (size_t)nw * nw + (size_t)nw * ncomp) + 2 * nw * sizeof(short) + 64;
}

template <int NT>
__device__ __forceinline__ void wb_fsea_groups(const double* __restrict__ Eall, long ik, int nw, const WbWindow& win,
                                               double* Es, double* label, double* inv, short* g1, short* g2) {
    for (int x = threadIdx.x; x < nw; x += NT) Es[x] = Eall[ik * nw + x];
    __syncthreads();
    if (win.Ebmin) {
        if (threadIdx.x == 0) wb_band_groups_tetra(Es, win.Ebmin + ik * nw, win.Ebmax + ik * nw, nw, win, g1, g2, label);
    } else if (nw <= 32) {
        if (threadIdx.x < 32) wb_band_groups_warp(Es, nw, win, g1, g2, label, threadIdx.x);
    } else if (threadIdx.x == 0) wb_band_groups(Es, nw, win, g1, g2, label);
    for (int x = threadIdx.x; x < nw * nw; x += NT) inv[x] = wb_deinv(Es[x / nw], Es[x % nw]);   // dEig_inv (data_K.py:290-298)
    __syncthreads();
}

// The velocity matrices V_a and D_a = -V_a / (E_p - E_q) of a k-point are read O(nw) times per element by the generalised
// derivatives below; from global memory (L2) the kernels spent 65 % of their time waiting for those loads (ncu,
// profiles/r2/fsea_kernels.txt).  When the launch reserves 2 x 3 x (nw^2 + 1) complex of dynamic shared memory behind
// the arrays of wb_fsea_smem_bytes (`stage` != 0), both are staged there once per k-point; the component stride nw^2 + 1
// keeps the three Cartesian components in different banks.
__host__ __device__ inline size_t wb_fsea_stage_offset(int nw, int ncomp) {   // in bytes, 16-byte aligned
    return (sizeof(double) * (2 * (size_t)nw + (size_t)nw * nw + (size_t)nw * ncomp) + 2 * nw * sizeof(short) + 64 + 15) / 16 * 16;
}
__host__ inline size_t wb_fsea_stage_bytes(int nw) { return sizeof(cplx) * 6 * ((size_t)nw * nw + 1); }
// (the kernels are templated on STAGE so that the staged accesses compile to shared-memory loads, not generic ones)
template <int NT>
__device__ __forceinline__ void wb_fsea_stage(const cplx* __restrict__ Vg, const double* inv, int nw, cplx* Vs, cplx* Ds) {
    const int n2 = nw * nw;
    for (int x = threadIdx.x; x < 3 * n2; x += NT) {
        const int a = x / n2, e = x - a * n2;
        const cplx v = Vg[x];
        Vs[a * (n2 + 1) + e] = v;
        Ds[a * (n2 + 1) + e] = cscale(-inv[e], v);
    }
    __syncthreads();
}

// trace_G[a][b][s] = sum_{m in G} sum_{l notin G} -2 Im( J_ml^{as} / (E_m - E_l) * (D_lm^b - i A_lm^b) )
template <int NT>
__global__ void __launch_bounds__(NT)
wb_spinomega_events_kernel(const cplx* __restrict__ xbar, int nch, int nw, long nk, const double* __restrict__ Eall,
                           WbWindow win, int iV, int iA, const cplx* __restrict__ Jspin, double* __restrict__ ev_label,
                           double* __restrict__ ev_val) {
    extern __shared__ __align__(16) double smem_f[];
    const int n2 = nw * nw;
    double* Es = smem_f;
    double* label = Es + nw;
    double* inv = label + nw;
    double* vals = inv + n2;           // [nw][27]
    short* g1 = (short*)(vals + 27 * nw);
    short* g2 = g1 + nw;
    for (long ik = blockIdx.x; ik < nk; ik += gridDim.x) {
        __syncthreads();
        wb_fsea_groups<NT>(Eall, ik, nw, win, Es, label, inv, g1, g2);
        const cplx* X = xbar + (size_t)ik * nch * n2;
        const cplx* J = Jspin + (size_t)ik * 9 * n2;
        for (int x = threadIdx.x; x < nw * 27; x += NT) {
            const int m = x / 27, comp = x - 27 * m;
            const int a = comp / 9, b = (comp / 3) % 3, s = comp % 3;
            double acc = 0.;
            if (g1[m] >= 0) {
                const int ga = g1[m], gb = g2[m];
                const cplx* Jm = J + (size_t)(3 * a + s) * n2 + m * nw;
                const cplx* Vb = X + (size_t)(iV + b) * n2;
                for (int l = 0; l < nw; l++) {
                    if (l >= ga && l < gb) continue;
                    const double iml = inv[m * nw + l];
                    const cplx v = Vb[l * nw + m];
                    cplx vd = cmake(iml * v.x, iml * v.y);      // D_lm = -V_lm / (E_l - E_m) = V_lm / (E_m - E_l)
                    if (iA >= 0) {                               // - i A_lm
                        const cplx al = X[(size_t)(iA + b) * n2 + l * nw + m];
                        vd = cmake(vd.x + al.y, vd.y - al.x);
                    }
                    const cplx j = Jm[l];
                    acc += iml * (j.x * vd.y + j.y * vd.x);
                }
            }
            vals[x] = -2. * acc;
        }
        __syncthreads();
        for (int x = threadIdx.x; x < nw; x += NT) ev_label[ik * nw + x] = label[x];
        for (int x = threadIdx.x; x < nw * 27; x += NT) {
            const int n0 = x / 27, comp = x - 27 * n0;
            if (label[n0] != CUDART_INF) {
                double s = 0.;
                for (int n = n0; n < g2[n0]; n++) s += vals[n * 27 + comp];
                ev_val[((size_t)ik * nw + n0) * 27 + comp] = s;
            }
        }
    }
}

__host__ __device__ inline size_t wb_deromega_scratch_elems(int nw) { return (size_t)36 * nw * nw; }   // complex elements per CTA

template <int NT, bool STAGE>
__global__ void __launch_bounds__(NT)
wb_deromega_events_kernel(const cplx* __restrict__ xbar, int nch, int nw, long nk, const double* __restrict__ Eall,
                          WbWindow win, WbDerOmegaChans C, int internal, int external, cplx* __restrict__ scratch,
                          double* __restrict__ ev_label, double* __restrict__ ev_val) {
    extern __shared__ __align__(16) double smem_f[];
    const int n2 = nw * nw;
    double* Es = smem_f;
    double* label = Es + nw;
    double* inv = label + nw;
    double* vals = inv + n2;           // [nw][9]: rows of the current group (indexed by band)
    short* g1 = (short*)(vals + 9 * nw);
    short* g2 = g1 + nw;
    cplx* const T1 = scratch + (size_t)blockIdx.x * wb_deromega_scratch_elems(nw);   // dD_ln [l][n - ga][b][d]
    cplx* const T2 = T1 + (size_t)9 * n2;                                            // dD_ml [m - ga][l][a][d]
    cplx* const T3 = T2 + (size_t)9 * n2;                                            // dA_ln [l][n - ga][b][d]
    cplx* const T4 = T3 + (size_t)9 * n2;                                            // dA_pn [p - ga][n - ga][b][d]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = NT / 32;
    for (long ik = blockIdx.x; ik < nk; ik += gridDim.x) {
        __syncthreads();
        wb_fsea_groups<NT>(Eall, ik, nw, win, Es, label, inv, g1, g2);
        const cplx* X = xbar + (size_t)ik * nch * n2;
        const cplx* V = X + (size_t)C.iV * n2;
        cplx* const Vs = (cplx*)((char*)smem_f + wb_fsea_stage_offset(nw, 9));   // [3][n2 + 1], then D likewise (STAGE)
        cplx* const Ds = Vs + 3 * (n2 + 1);
        if (STAGE) wb_fsea_stage<NT>(V, inv, nw, Vs, Ds);
        auto Vel = [&](int a, int e) { return STAGE ? Vs[a * (n2 + 1) + e] : V[(size_t)a * n2 + e]; };
        const cplx* A = X + (size_t)C.iA * n2;
        const cplx* O = X + (size_t)C.iO * n2;
        const cplx* W = X + (size_t)C.iW * n2;
        const cplx* dAc = X + (size_t)C.idA * n2;
        const cplx* dOc = X + (size_t)C.idO * n2;
        auto Dm = [&](int a, int p, int q) {   // D_pq,a = -V_pq,a / (E_p - E_q)
            return STAGE ? Ds[a * (n2 + 1) + p * nw + q] : cscale(-inv[p * nw + q], V[(size_t)a * n2 + p * nw + q]);
        };
        for (int x = threadIdx.x; x < nw; x += NT) ev_label[ik * nw + x] = label[x];
        for (int ga = 0; ga < nw; ga++) {
            if (label[ga] == CUDART_INF) continue;   // uniform: no group starts here
            const int gb = g2[ga], g = gb - ga;
            auto in_G = [&](int q) { return q >= ga && q < gb; };
            // ---- DerDcov: dD_rc^{bd} = -1/(E_r - E_c) ( W_rc^{bd} + sum_{p in rows} (V_rp^b D_pc^d + V_rp^d D_pc^b)
            //                                              - sum_{q in cols} (D_rq^b V_qc^d + D_rq^d V_qc^b) )
            //      T1: rows = out, cols = G;  T2: rows = G, cols = out (needed by the external terms only)
            const int nT12 = external ? 2 : 1;
            // thread = one element (r, c) with all six symmetric (b, d): three V / three D loads per partner band serve nine
            // products (a thread per (r, c, b, d) re-read them for every component pair and computed b <-> d twice)
            for (int x = threadIdx.x; x < nT12 * nw * g; x += NT) {
                const int which = x / (nw * g);
                const int y = x - which * (nw * g);
                int r, cidx;
                if (which == 0) { cidx = ga + y % g; r = y / g; }    // r = l (any band), c in G
                else { cidx = y % nw; r = ga + y / nw; }             // r = m in G, c = l
                const bool rG = in_G(r), cG = in_G(cidx);
                const bool live = (rG != cG);
                cplx acc[6];
#pragma unroll
                for (int i = 0; i < 6; i++) acc[i] = cmake(0., 0.);
                if (live) {
                    for (int p = 0; p < nw; p++) {
                        const bool prow = (in_G(p) == rG);   // p in the row set: + V_rp D_pc, else: - D_rp V_pc
                        cplx Xv[3], Yv[3];
                        if (prow) {   // (uniform over the live threads of a warp)
#pragma unroll
                            for (int a3 = 0; a3 < 3; a3++) { Xv[a3] = Vel(a3, r * nw + p); Yv[a3] = Dm(a3, p, cidx); }
                        } else {
#pragma unroll
                            for (int a3 = 0; a3 < 3; a3++) {
                                const cplx dr = Dm(a3, r, p);
                                Xv[a3] = cmake(-dr.x, -dr.y);
                                Yv[a3] = Vel(a3, p * nw + cidx);
                            }
                        }
                        int i = 0;
#pragma unroll
                        for (int b3 = 0; b3 < 3; b3++)
#pragma unroll
                            for (int d3 = b3; d3 < 3; d3++, i++) {
                                cfma(acc[i], Xv[b3], Yv[d3]);
                                cfma(acc[i], Xv[d3], Yv[b3]);
                            }
                    }
                }
                const double mi = live ? -inv[r * nw + cidx] : 0.;
                cplx* const dst = (which == 0) ? T1 + ((size_t)r * g + (cidx - ga)) * 9 : T2 + ((size_t)(r - ga) * nw + cidx) * 9;
                int i = 0;
#pragma unroll
                for (int b3 = 0; b3 < 3; b3++)
#pragma unroll
                    for (int d3 = b3; d3 < 3; d3++, i++) {
                        cplx res = cmake(0., 0.);
                        if (live) res = cscale(mi, cadd(W[(size_t)i * n2 + r * nw + cidx], acc[i]));
                        dst[3 * b3 + d3] = res;
                        dst[3 * d3 + b3] = res;
                    }
            }
            // ---- Matrix_GenDer_ln of A:  T3 = dA.ln (l in out, n in G),  T4 = dA.nn (p, n in G)
            if (external) {
                for (int x = threadIdx.x; x < nw * g; x += NT) {   // thread = (l, n), all nine (b, d)
                    const int n = ga + x % g, l = x / g;
                    const bool lG = in_G(l);
                    cplx acc[9];
#pragma unroll
                    for (int bd = 0; bd < 9; bd++) acc[bd] = dAc[(size_t)bd * n2 + l * nw + n];
                    // l in out:  dA_ln = A,d_ln - sum_{m' in G} D_lm'^d A_m'n^b + sum_{p in out} A_lp^b D_pn^d
                    // l in G:    dA_pn (p = l) = A,d_pn - sum_{q in out} D_pq^d A_qn^b + sum_{q in out} A_pq^b D_qn^d
                    for (int p = 0; p < nw; p++) {
                        const bool pG = in_G(p);
                        if (lG && pG) continue;
                        const bool minus = lG || pG, plus = lG || !pG;
                        cplx Dl[3], Ap[3], Al[3], Dp[3];
#pragma unroll
                        for (int a3 = 0; a3 < 3; a3++) {
                            Dl[a3] = Dm(a3, l, p);
                            Ap[a3] = A[(size_t)a3 * n2 + p * nw + n];
                            Al[a3] = A[(size_t)a3 * n2 + l * nw + p];
                            Dp[a3] = Dm(a3, p, n);
                        }
#pragma unroll
                        for (int b3 = 0; b3 < 3; b3++)
#pragma unroll
                            for (int d3 = 0; d3 < 3; d3++) {
                                if (minus) {
                                    const cplx z = cmul(Dl[d3], Ap[b3]);
                                    acc[3 * b3 + d3] = cmake(acc[3 * b3 + d3].x - z.x, acc[3 * b3 + d3].y - z.y);
                                }
                                if (plus) cfma(acc[3 * b3 + d3], Al[b3], Dp[d3]);
                            }
                    }
                    cplx* const dst = lG ? T4 + ((size_t)(l - ga) * g + (n - ga)) * 9 : T3 + ((size_t)l * g + (n - ga)) * 9;
#pragma unroll
                    for (int bd = 0; bd < 9; bd++) dst[bd] = acc[bd];
                }
            }
            __syncthreads();
            // ---- trace: one warp per (m in G, c, d), lanes over the partner bands
            for (int item = warp; item < g * 9; item += nwarp) {
                const int m = ga + item / 9, cd = item % 9, c = cd / 3, d = cd - 3 * c;
                cplx S = cmake(0., 0.);
                for (int l = lane; l < nw; l += 32) {
                    const bool lG = in_G(l);
                    if (external && !lG) {   // 1/2 of the generalised derivative of rotA, off-diagonal part
                        cplx z = cmul(O[(size_t)c * n2 + m * nw + l], Dm(d, l, m));
                        const cplx z2 = cmul(Dm(d, m, l), O[(size_t)c * n2 + l * nw + m]);
                        S.x += 0.5 * (z.x - z2.x);
                        S.y += 0.5 * (z.y - z2.y);
                    }
#pragma unroll
                    for (int t = 0; t < 2; t++) {
                        const int a = t ? WB_BETA(c) : WB_ALPHA(c), b = t ? WB_ALPHA(c) : WB_BETA(c);
                        const double sg = t ? -1. : 1.;
                        if (!lG) {
                            const cplx Dml = Dm(a, m, l);
                            if (internal) {   // -i s D_ml^a dD_lm^{bd}
                                const cplx z = cmul(Dml, T1[((size_t)l * g + (m - ga)) * 9 + 3 * b + d]);
                                S.x += sg * z.y;
                                S.y -= sg * z.x;
                            }
                            if (external) {   // -s D_ml^a dA_lm^{b:d} - s dD_ml^{ad} A_lm^b
                                const cplx z = cmul(Dml, T3[((size_t)l * g + (m - ga)) * 9 + 3 * b + d]);
                                const cplx z2 = cmul(T2[((size_t)(m - ga) * nw + l) * 9 + 3 * a + d], A[(size_t)b * n2 + l * nw + m]);
                                S.x -= sg * (z.x + z2.x);
                                S.y -= sg * (z.y + z2.y);
                            }
                        } else if (external) {   // -i s A_mp^a dA_pm^{b:d},  p = l in G
                            const cplx z = cmul(A[(size_t)a * n2 + m * nw + l], T4[((size_t)(l - ga) * g + (m - ga)) * 9 + 3 * b + d]);
                            S.x += sg * z.y;
                            S.y -= sg * z.x;
                        }
                    }
                }
                double re = S.x;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) re += __shfl_xor_sync(0xffffffffu, re, o);
                if (lane == 0) {
                    if (external) re += 0.5 * dOc[(size_t)cd * n2 + m * nw + m].x;
                    vals[m * 9 + cd] = 2. * re;   // summ + summ^dagger, real part of the trace
                }
            }
            __syncthreads();
            for (int cd = threadIdx.x; cd < 9; cd += NT) {
                double s = 0.;
                for (int n = ga; n < gb; n++) s += vals[n * 9 + cd];
                ev_val[((size_t)ik * nw + ga) * 9 + cd] = s;
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// FormulaProduct (formula/formula.py:121-149) of up to three factors whose nn blocks are formed per band group:
//   kind 1  V    covariant('Ham', commader = 1)                     rank 1
//   kind 2  M    InvMass = generalised derivative of V               rank 2  (elementary.py:28-34, formula.py:95-112)
//   kind 3  Om   Omega                                              rank 1  (covariant.py:161-203)
//   kind 4  S    Spin                                               rank 1  (covariant.py:331-335)
//   kind 5  dS   DerSpin = generalised derivative of S               rank 2  (covariant.py:338-342)
//   kind 6  Hp   Morb_Hpm (sign = +1) = Morb_H + Eav Omega           rank 1  (covariant.py:375-449)
// trace over the group of F1 F2 (F3), cartesian indices appended in order.  Covers VelVelVel, MassVel, MassMass,
// VelMassVel, OmegaS, OmegaOmega (covariant.py:823-858) and the single-factor DerSpin.
struct WbProductSpec {
    int nf;
    int kind[3];
    int iV, iW, iA, iO, iS, idS;   // matrix positions in the rotated record: d_a H [3] | d_b d_d H [6] | A [3] | rotA [3] | S [3] | d_d S_s [9]
    int iB, iC;                    // B [3] | C [3] (kind 6 with external terms)
};
__host__ __device__ inline int wb_product_kind_ncomp(int kind) { return (kind == 2 || kind == 5) ? 9 : 3; }
__host__ __device__ inline size_t wb_product_scratch_elems(int nw) { return (size_t)56 * nw * nw; }

template <int NT, bool STAGE>
__global__ void __launch_bounds__(NT)
wb_product_events_kernel(const cplx* __restrict__ xbar, int nch, int nw, long nk, const double* __restrict__ Eall,
                         WbWindow win, WbProductSpec P, int internal, int external, int stage_off, cplx* __restrict__ scratch,
                         double* __restrict__ ev_label, double* __restrict__ ev_val) {
    extern __shared__ __align__(16) double smem_f[];
    const int n2 = nw * nw;
    double* Es = smem_f;
    double* label = Es + nw;
    double* inv = label + nw;
    short* g1 = (short*)(inv + n2);
    short* g2 = g1 + nw;
    cplx* const F0 = scratch + (size_t)blockIdx.x * wb_product_scratch_elems(nw);
    int nc[3] = {1, 1, 1};
    for (int f = 0; f < P.nf; f++) nc[f] = wb_product_kind_ncomp(P.kind[f]);
    const int NC = nc[0] * nc[1] * nc[2];
    for (long ik = blockIdx.x; ik < nk; ik += gridDim.x) {
        __syncthreads();
        wb_fsea_groups<NT>(Eall, ik, nw, win, Es, label, inv, g1, g2);
        const cplx* X = xbar + (size_t)ik * nch * n2;
        const cplx* V = X + (size_t)P.iV * n2;
        cplx* const Vs = (cplx*)((char*)smem_f + stage_off);   // [3][n2 + 1], then D likewise (STAGE: wb_fsea_stage)
        cplx* const Ds = Vs + 3 * (n2 + 1);
        if (STAGE) wb_fsea_stage<NT>(V, inv, nw, Vs, Ds);
        auto Vel = [&](int a, int e) { return STAGE ? Vs[a * (n2 + 1) + e] : V[(size_t)a * n2 + e]; };
        const cplx* A = X + (size_t)P.iA * n2;
        auto Dm = [&](int a, int p, int q) {
            return STAGE ? Ds[a * (n2 + 1) + p * nw + q] : cscale(-inv[p * nw + q], V[(size_t)a * n2 + p * nw + q]);
        };
        for (int x = threadIdx.x; x < nw; x += NT) ev_label[ik * nw + x] = label[x];
        for (int ga = 0; ga < nw; ga++) {
            if (label[ga] == CUDART_INF) continue;   // uniform
            const int gb = g2[ga], g = gb - ga, gg = g * g;
            auto in_G = [&](int q) { return q >= ga && q < gb; };
            cplx* Fb[3];
            Fb[0] = F0;
            Fb[1] = Fb[0] + (size_t)gg * nc[0];
            Fb[2] = Fb[1] + (size_t)gg * nc[1];
            cplx* const Pm = Fb[2] + (size_t)gg * nc[2];   // F1 F2 of a three-factor product: [g][g][nc0 nc1]
            // ---- factor blocks F_f[m - ga][n - ga][comp]
            const int tot = gg * (nc[0] + (P.nf > 1 ? nc[1] : 0) + (P.nf > 2 ? nc[2] : 0));
            for (int x = threadIdx.x; x < tot; x += NT) {
                int f = 0, y = x;
                while (y >= gg * nc[f]) { y -= gg * nc[f]; f++; }
                const int comp = y % nc[f];
                const int mn = y / nc[f], m = ga + mn / g, n = ga + mn % g;
                const int kind = P.kind[f];
                cplx val;
                if (kind == 1) val = Vel(comp, m * nw + n);
                else if (kind == 4) val = X[(size_t)(P.iS + comp) * n2 + m * nw + n];
                else if (kind == 2 || kind == 5) {
                    // X^{b:d}_mn = X,d_mn - sum_{l notin G} D_ml^d X_ln^b + sum_l X_ml^b D_ln^d
                    const int b = comp / 3, d = comp - 3 * b;
                    const cplx* Xb = (kind == 2) ? V + (size_t)b * n2 : X + (size_t)(P.iS + b) * n2;
                    val = (kind == 2) ? X[(size_t)(P.iW + wb_sym6(b, d)) * n2 + m * nw + n] : X[(size_t)(P.idS + comp) * n2 + m * nw + n];
                    for (int l = 0; l < nw; l++) {
                        if (in_G(l)) continue;
                        const cplx z = cmul(Dm(d, m, l), Xb[l * nw + n]);
                        val = cmake(val.x - z.x, val.y - z.y);
                        cfma(val, Xb[m * nw + l], Dm(d, l, n));
                    }
                } else {
                    // Omega_c[m, n] = S(m, n) + conj(S(n, m)),
                    // S(M, L) = -i sum_l D_Ml^al D_lL^be + 1/2 O_ML - sum_l D_Ml^al A_lL^be + sum_l D_Ml^be A_lL^al - i sum_{m' in G} A_Mm'^al A_m'L^be
                    // kind 6 adds Morb_H: the same sums weighted with E_l / E_m', B instead of A, C instead of O, and Eav Omega
                    const int al = WB_ALPHA(comp), be = WB_BETA(comp);
                    const bool hp = (kind == 6);
                    const cplx* Bm = X + (size_t)P.iB * n2;
                    val = cmake(0., 0.);
                    cplx valh = cmake(0., 0.);
#pragma unroll
                    for (int side = 0; side < 2; side++) {
                        const int M = side ? n : m, Lb = side ? m : n;
                        cplx S = cmake(0., 0.), Sh = cmake(0., 0.);
                        for (int l = 0; l < nw; l++) {
                            if (in_G(l)) {
                                if (external) {
                                    const cplx z = cmul(A[(size_t)al * n2 + M * nw + l], A[(size_t)be * n2 + l * nw + Lb]);
                                    S.x += z.y; S.y -= z.x;
                                    if (hp) { Sh.x += Es[l] * z.y; Sh.y -= Es[l] * z.x; }
                                }
                                continue;
                            }
                            const cplx DMa = Dm(al, M, l), DMb = Dm(be, M, l);
                            if (internal) {
                                const cplx z = cmul(DMa, Dm(be, l, Lb));
                                S.x += z.y; S.y -= z.x;
                                if (hp) { Sh.x += Es[l] * z.y; Sh.y -= Es[l] * z.x; }
                            }
                            if (external) {
                                const cplx z = csub(cmul(DMb, A[(size_t)al * n2 + l * nw + Lb]), cmul(DMa, A[(size_t)be * n2 + l * nw + Lb]));
                                S = cadd(S, z);
                                if (hp) Sh = cadd(Sh, csub(cmul(DMb, Bm[(size_t)al * n2 + l * nw + Lb]), cmul(DMa, Bm[(size_t)be * n2 + l * nw + Lb])));
                            }
                        }
                        if (external) {
                            const cplx o = X[(size_t)(P.iO + comp) * n2 + M * nw + Lb];
                            S.x += 0.5 * o.x; S.y += 0.5 * o.y;
                            if (hp) {
                                const cplx cc = X[(size_t)(P.iC + comp) * n2 + M * nw + Lb];
                                Sh.x += 0.5 * cc.x; Sh.y += 0.5 * cc.y;
                            }
                        }
                        val = side ? cmake(val.x + S.x, val.y - S.y) : S;
                        valh = side ? cmake(valh.x + Sh.x, valh.y - Sh.y) : Sh;
                    }
                    if (hp) {
                        const double eav = 0.5 * (Es[m] + Es[n]);
                        val = cmake(valh.x + eav * val.x, valh.y + eav * val.y);
                    }
                }
                Fb[f][y] = val;
            }
            __syncthreads();
            if (P.nf == 3) {
                const int n01 = nc[0] * nc[1];
                for (int x = threadIdx.x; x < gg * n01; x += NT) {
                    const int c01 = x % n01, mp = x / n01, m = mp / g, p = mp % g;
                    const int c0 = c01 / nc[1], c1 = c01 - c0 * nc[1];
                    cplx acc = cmake(0., 0.);
                    for (int n = 0; n < g; n++) cfma(acc, Fb[0][(size_t)(m * g + n) * nc[0] + c0], Fb[1][(size_t)(n * g + p) * nc[1] + c1]);
                    Pm[x] = acc;
                }
                __syncthreads();
            }
            // ---- trace
            for (int comp = threadIdx.x; comp < NC; comp += NT) {
                double tr = 0.;
                if (P.nf == 1) {
                    for (int m = 0; m < g; m++) tr += Fb[0][(size_t)(m * g + m) * nc[0] + comp].x;
                } else {
                    const int nlast = (P.nf == 2) ? nc[1] : nc[2];
                    const int cl = comp % nlast, cf = comp / nlast;
                    const cplx* L = (P.nf == 2) ? Fb[0] : Pm;
                    const cplx* Rr = (P.nf == 2) ? Fb[1] : Fb[2];
                    const int nfirst = NC / nlast;
                    for (int m = 0; m < g; m++)
                        for (int n = 0; n < g; n++) {
                            const cplx a = L[(size_t)(m * g + n) * nfirst + cf], b = Rr[(size_t)(n * g + m) * nlast + cl];
                            tr += a.x * b.x - a.y * b.y;
                        }
                }
                ev_val[((size_t)ik * nw + ga) * NC + comp] = tr;
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Der3E (formula/covariant.py:126-151): third derivative of the band energy, trace over a band group G of
//   W3_mn^{abc} - D_ml^c W_ln^{ab} + W_ml^{ab} D_ln^c        (DerWln.nn, elementary.py:36-43)
//   + dV_ml^{a:c} D_ln^b + V_ml^a dD_ln^{bc} - dD_ml^{bc} V_ln^a - D_ml^b dV_ln^{a:c}
// with the generalised derivatives of V (InvMass, elementary.py:28-34) and of D (DerDcov, :56-72) between G and its
// complement.  Channels: d_a H [3] | d_b d_d H [6, wb_sym6] | d_b d_c d_d H [10, wb_sym10].
__host__ __device__ __forceinline__ int wb_sym10(int a, int b, int c) {
    // index of the sorted triple among xxx xxy xxz xyy xyz xzz yyy yyz yzz zzz
    int lo = min(a, min(b, c)), hi = max(a, max(b, c)), mid = a + b + c - lo - hi;
    const int base[3] = {0, 6, 9};
    return base[lo] + ((lo == 0) ? (mid == 0 ? hi : (mid == 1 ? 2 + hi : 5)) : (lo == 1) ? (mid == 1 ? hi - 1 : 2) : 0);
}

template <int NT, bool STAGE>
__global__ void __launch_bounds__(NT)
wb_der3e_events_kernel(const cplx* __restrict__ xbar, int nch, int nw, long nk, const double* __restrict__ Eall, WbWindow win,
                       int iV, int iW, int iW3, int stage_off, cplx* __restrict__ scratch, double* __restrict__ ev_label,
                       double* __restrict__ ev_val) {
    extern __shared__ __align__(16) double smem_f[];
    const int n2 = nw * nw;
    double* Es = smem_f;
    double* label = Es + nw;
    double* inv = label + nw;
    double* vals = inv + n2;           // [nw][27]
    short* g1 = (short*)(vals + 27 * nw);
    short* g2 = g1 + nw;
    cplx* const T1 = scratch + (size_t)blockIdx.x * wb_deromega_scratch_elems(nw);   // dD_ln [l][n - ga][b][c]
    cplx* const T2 = T1 + (size_t)9 * n2;                                            // dD_ml [m - ga][l][b][c]
    cplx* const T5 = T2 + (size_t)9 * n2;                                            // dV_ml [m - ga][l][a][c]
    cplx* const T6 = T5 + (size_t)9 * n2;                                            // dV_ln [l][n - ga][a][c]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = NT / 32;
    for (long ik = blockIdx.x; ik < nk; ik += gridDim.x) {
        __syncthreads();
        wb_fsea_groups<NT>(Eall, ik, nw, win, Es, label, inv, g1, g2);
        const cplx* X = xbar + (size_t)ik * nch * n2;
        const cplx* V = X + (size_t)iV * n2;
        cplx* const Vs = (cplx*)((char*)smem_f + stage_off);   // [3][n2 + 1], then D likewise (STAGE: wb_fsea_stage)
        cplx* const Ds = Vs + 3 * (n2 + 1);
        if (STAGE) wb_fsea_stage<NT>(V, inv, nw, Vs, Ds);
        auto Vel = [&](int a, int e) { return STAGE ? Vs[a * (n2 + 1) + e] : V[(size_t)a * n2 + e]; };
        const cplx* W = X + (size_t)iW * n2;
        const cplx* W3 = X + (size_t)iW3 * n2;
        auto Dm = [&](int a, int p, int q) {
            return STAGE ? Ds[a * (n2 + 1) + p * nw + q] : cscale(-inv[p * nw + q], V[(size_t)a * n2 + p * nw + q]);
        };
        for (int x = threadIdx.x; x < nw; x += NT) ev_label[ik * nw + x] = label[x];
        for (int ga = 0; ga < nw; ga++) {
            if (label[ga] == CUDART_INF) continue;   // uniform
            const int gb = g2[ga], g = gb - ga;
            auto in_G = [&](int q) { return q >= ga && q < gb; };
            // which: 0 dD rows out cols G (T1) | 1 dD rows G cols out (T2) | 2 dV rows G cols out (T5) | 3 dV rows out cols G (T6)
            for (int x = threadIdx.x; x < 4 * nw * g * 9; x += NT) {
                const int which = x / (nw * g * 9);
                int y = x - which * (nw * g * 9);
                const int bd = y % 9; y /= 9;
                const int b = bd / 3, d = bd - 3 * b;
                const bool rows_out = (which == 0 || which == 3);
                int r, cidx;
                if (rows_out) { cidx = ga + y % g; r = y / g; }
                else { cidx = y % nw; r = ga + y / nw; }
                const bool rG = in_G(r), cG = in_G(cidx);
                cplx res = cmake(0., 0.);
                if (rG != cG) {
                    cplx sum = W[(size_t)wb_sym6(b, d) * n2 + r * nw + cidx];
                    if (which < 2) {   // DerDcov
                        for (int p = 0; p < nw; p++) {
                            if (in_G(p) == rG) {
                                cfma(sum, Vel(b, r * nw + p), Dm(d, p, cidx));
                                cfma(sum, Vel(d, r * nw + p), Dm(b, p, cidx));
                            } else {
                                const cplx z1 = cmul(Dm(b, r, p), Vel(d, p * nw + cidx));
                                const cplx z2 = cmul(Dm(d, r, p), Vel(b, p * nw + cidx));
                                sum = cmake(sum.x - z1.x - z2.x, sum.y - z1.y - z2.y);
                            }
                        }
                        res = cscale(-inv[r * nw + cidx], sum);
                    } else {           // InvMass: V^{b:d}_rc = W_rc^{bd} - sum_{q in cols} D_rq^d V_qc^b + sum_{p in rows} V_rp^b D_pc^d
                        for (int p = 0; p < nw; p++) {
                            if (in_G(p) == rG) cfma(sum, Vel(b, r * nw + p), Dm(d, p, cidx));
                            else {
                                const cplx z = cmul(Dm(d, r, p), Vel(b, p * nw + cidx));
                                sum = cmake(sum.x - z.x, sum.y - z.y);
                            }
                        }
                        res = sum;
                    }
                }
                cplx* T = (which == 0) ? T1 : (which == 1) ? T2 : (which == 2) ? T5 : T6;
                if (rows_out) T[((size_t)r * g + (cidx - ga)) * 9 + bd] = res;
                else T[((size_t)(r - ga) * nw + cidx) * 9 + bd] = res;
            }
            __syncthreads();
            // ---- trace: one warp per (m in G, a, b, c), lanes over the partner bands
            for (int item = warp; item < g * 27; item += nwarp) {
                const int m = ga + item / 27, abc = item % 27, a = abc / 9, b = (abc / 3) % 3, c = abc % 3;
                double re = 0.;
                for (int l = lane; l < nw; l += 32) {
                    if (in_G(l)) continue;
                    const cplx Dml_c = Dm(c, m, l), Dlm_c = Dm(c, l, m), Dlm_b = Dm(b, l, m), Dml_b = Dm(b, m, l);
                    const int ab6 = wb_sym6(a, b);
                    re += cmul(W[(size_t)ab6 * n2 + m * nw + l], Dlm_c).x - cmul(Dml_c, W[(size_t)ab6 * n2 + l * nw + m]).x;
                    re += cmul(T5[((size_t)(m - ga) * nw + l) * 9 + 3 * a + c], Dlm_b).x;
                    re += cmul(Vel(a, m * nw + l), T1[((size_t)l * g + (m - ga)) * 9 + 3 * b + c]).x;
                    re -= cmul(T2[((size_t)(m - ga) * nw + l) * 9 + 3 * b + c], Vel(a, l * nw + m)).x;
                    re -= cmul(Dml_b, T6[((size_t)l * g + (m - ga)) * 9 + 3 * a + c]).x;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) re += __shfl_xor_sync(0xffffffffu, re, o);
                if (lane == 0) vals[m * 27 + abc] = re + W3[(size_t)wb_sym10(a, b, c) * n2 + m * nw + m].x;
            }
            __syncthreads();
            for (int abc = threadIdx.x; abc < 27; abc += NT) {
                double s = 0.;
                for (int n = ga; n < gb; n++) s += vals[n * 27 + abc];
                ev_val[((size_t)ik * nw + ga) * 27 + abc] = s;
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// DerMorb (formula/covariant.py:463-547; GME_orb_FermiSea): DerMorb_H plus Eav * DerOmega + (Omega V + V Omega) / 2.
// Not additive (static.py:109-117): T(x) = trace with inn = [0, x), out = [x, nw) at every edge x of a band group, value
// of the group [a, b) = T(b) - T(a).  Per edge the generalised derivatives between inn and out are formed as in
// wb_deromega_events_kernel (dD, dA) plus dB.ln, the two V-weighted double sums and the Omega block of inn.
struct WbDerMorbChans {
    int iV, iW, iA, iO, idA, idO, iB, idB, iC, idC;   // [3] [6] [3] [3] [9] [9] [3] [9] [3] [9]
};
__host__ __device__ inline size_t wb_dermorb_scratch_elems(int nw) { return (size_t)66 * nw * nw; }
__host__ inline size_t wb_dermorb_smem_bytes(int nw) {
    return sizeof(double) * (2 * (size_t)nw + (size_t)nw * nw + 9 * (size_t)nw + 9 * ((size_t)nw + 1)) + 2 * nw * sizeof(short) +
           (nw + 1) * sizeof(int) + 64;
}

template <int NT, bool STAGE>
__global__ void __launch_bounds__(NT)
wb_dermorb_events_kernel(const cplx* __restrict__ xbar, int nch, int nw, long nk, const double* __restrict__ Eall,
                         WbWindow win, WbDerMorbChans C, int internal, int external, int stage_off, cplx* __restrict__ scratch,
                         double* __restrict__ ev_label, double* __restrict__ ev_val) {
    extern __shared__ __align__(16) double smem_f[];
    const int n2 = nw * nw;
    double* Es = smem_f;
    double* label = Es + nw;
    double* inv = label + nw;
    double* vals = inv + n2;                 // [nw][9]
    double* Tedge = vals + 9 * nw;           // [nw + 1][9]
    short* g1 = (short*)(Tedge + 9 * (nw + 1));
    short* g2 = g1 + nw;
    int* edge = (int*)(g2 + nw + (nw & 1) * 1);
    edge = (int*)(((uintptr_t)edge + 3) & ~(uintptr_t)3);
    cplx* const T1 = scratch + (size_t)blockIdx.x * wb_dermorb_scratch_elems(nw);   // dD_ln [l][n][b][d]
    cplx* const T2 = T1 + (size_t)9 * n2;    // dD_ml [m][l][a][d]
    cplx* const T3 = T2 + (size_t)9 * n2;    // dA_ln [l][n][b][d]
    cplx* const T4 = T3 + (size_t)9 * n2;    // dA_pn [p][n][b][d]
    cplx* const T7 = T4 + (size_t)9 * n2;    // dB_ln [l][n][b][d]
    cplx* const T8 = T7 + (size_t)9 * n2;    // sum_{l in out} V_pl^d D_lm^beta(c)  [p][m][d][c]   (p in out)
    cplx* const T9 = T8 + (size_t)9 * n2;    // sum_{l in inn} V_pl^d A_lm^beta(c)  [p][m][d][c]   (p in inn)
    cplx* const T10 = T9 + (size_t)9 * n2;   // Omega_ml^c [m][l][c]   (m, l in inn)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = NT / 32;
    for (long ik = blockIdx.x; ik < nk; ik += gridDim.x) {
        __syncthreads();
        wb_fsea_groups<NT>(Eall, ik, nw, win, Es, label, inv, g1, g2);
        for (int x = threadIdx.x; x <= nw; x += NT) edge[x] = 0;
        for (int x = threadIdx.x; x < 9 * (nw + 1); x += NT) Tedge[x] = 0.;
        __syncthreads();
        for (int x = threadIdx.x; x < nw; x += NT)
            if (label[x] != CUDART_INF) { edge[x] = 1; edge[g2[x]] = 1; }
        __syncthreads();
        const cplx* X = xbar + (size_t)ik * nch * n2;
        const cplx* V = X + (size_t)C.iV * n2;
        cplx* const Vs = (cplx*)((char*)smem_f + stage_off);   // [3][n2 + 1], then D likewise (STAGE: wb_fsea_stage)
        cplx* const Ds = Vs + 3 * (n2 + 1);
        if (STAGE) wb_fsea_stage<NT>(V, inv, nw, Vs, Ds);
        auto Vel = [&](int a, int e) { return STAGE ? Vs[a * (n2 + 1) + e] : V[(size_t)a * n2 + e]; };
        const cplx* W = X + (size_t)C.iW * n2;
        const cplx* A = X + (size_t)C.iA * n2;
        const cplx* O = X + (size_t)C.iO * n2;
        const cplx* dAc = X + (size_t)C.idA * n2;
        const cplx* dOc = X + (size_t)C.idO * n2;
        const cplx* B = X + (size_t)C.iB * n2;
        const cplx* dBc = X + (size_t)C.idB * n2;
        const cplx* Cm = X + (size_t)C.iC * n2;
        const cplx* dCc = X + (size_t)C.idC * n2;
        auto Dm = [&](int a, int p, int q) {
            return STAGE ? Ds[a * (n2 + 1) + p * nw + q] : cscale(-inv[p * nw + q], V[(size_t)a * n2 + p * nw + q]);
        };
        for (int x = threadIdx.x; x < nw; x += NT) ev_label[ik * nw + x] = label[x];
        for (int xe = 1; xe <= nw; xe++) {
            if (!edge[xe]) continue;   // uniform
            const int g = xe;          // inn = [0, g), out = [g, nw)
            auto in_G = [&](int q) { return q < g; };
            // ---- T1 / T2: DerDcov between out and inn (as in wb_deromega_events_kernel)
            for (int x = threadIdx.x; x < 2 * nw * g * 9; x += NT) {
                const int which = x / (nw * g * 9);
                int y = x - which * (nw * g * 9);
                const int bd = y % 9; y /= 9;
                const int b = bd / 3, d = bd - 3 * b;
                int r, cidx;
                if (which == 0) { cidx = y % g; r = y / g; }
                else { cidx = y % nw; r = y / nw; }
                const bool rG = in_G(r), cG = in_G(cidx);
                cplx res = cmake(0., 0.);
                if (rG != cG) {
                    cplx sum = W[(size_t)wb_sym6(b, d) * n2 + r * nw + cidx];
                    for (int p = 0; p < nw; p++) {
                        if (in_G(p) == rG) {
                            cfma(sum, Vel(b, r * nw + p), Dm(d, p, cidx));
                            cfma(sum, Vel(d, r * nw + p), Dm(b, p, cidx));
                        } else {
                            const cplx z1 = cmul(Dm(b, r, p), Vel(d, p * nw + cidx));
                            const cplx z2 = cmul(Dm(d, r, p), Vel(b, p * nw + cidx));
                            sum = cmake(sum.x - z1.x - z2.x, sum.y - z1.y - z2.y);
                        }
                    }
                    res = cscale(-inv[r * nw + cidx], sum);
                }
                if (which == 0) T1[((size_t)r * g + cidx) * 9 + bd] = res;
                else T2[((size_t)r * nw + cidx) * 9 + bd] = res;
            }
            // ---- T8 (p in out) / T9 (p in inn): sum_l V_pl^d X_lm^beta(c),  X = D over l in out / A over l in inn
            for (int x = threadIdx.x; x < nw * g * 9; x += NT) {
                int y = x;
                const int dc = y % 9; y /= 9;
                const int d = dc / 3, c = dc - 3 * d, be = WB_BETA(c);
                const int m = y % g, p = y / g;
                cplx sum = cmake(0., 0.);
                if (!in_G(p)) {
                    if (internal)
                        for (int l = g; l < nw; l++) cfma(sum, Vel(d, p * nw + l), Dm(be, l, m));
                    T8[((size_t)p * g + m) * 9 + dc] = sum;
                } else {
                    if (external)
                        for (int l = 0; l < g; l++) cfma(sum, Vel(d, p * nw + l), A[(size_t)be * n2 + l * nw + m]);
                    T9[((size_t)p * g + m) * 9 + dc] = sum;
                }
            }
            if (external) {
                // ---- T3 = dA.ln, T4 = dA.nn, T7 = dB.ln (Matrix_GenDer_ln)
                for (int x = threadIdx.x; x < 2 * nw * g * 9; x += NT) {
                    const int which = x / (nw * g * 9);   // 0: A, 1: B (ln block only)
                    int y = x - which * (nw * g * 9);
                    const int bd = y % 9; y /= 9;
                    const int b = bd / 3, d = bd - 3 * b;
                    const int n = y % g, l = y / g;
                    const cplx* Xb = (which == 0 ? A : B) + (size_t)b * n2;
                    cplx sum = (which == 0 ? dAc : dBc)[(size_t)bd * n2 + l * nw + n];
                    if (!in_G(l)) {
                        for (int p = 0; p < nw; p++) {
                            if (in_G(p)) {
                                const cplx z = cmul(Dm(d, l, p), Xb[p * nw + n]);
                                sum = cmake(sum.x - z.x, sum.y - z.y);
                            } else cfma(sum, Xb[l * nw + p], Dm(d, p, n));
                        }
                        (which == 0 ? T3 : T7)[((size_t)l * g + n) * 9 + bd] = sum;
                    } else if (which == 0) {
                        for (int q = g; q < nw; q++) {
                            const cplx z = cmul(Dm(d, l, q), Xb[q * nw + n]);
                            sum = cmake(sum.x - z.x, sum.y - z.y);
                            cfma(sum, Xb[l * nw + q], Dm(d, q, n));
                        }
                        T4[((size_t)l * g + n) * 9 + bd] = sum;
                    }
                }
            }
            // ---- T10: Omega block of inn (covariant.py:161-203)
            for (int x = threadIdx.x; x < g * g * 3; x += NT) {
                const int comp = x % 3, mn = x / 3, m = mn / g, n = mn % g;
                const int al = WB_ALPHA(comp), be = WB_BETA(comp);
                cplx val = cmake(0., 0.);
#pragma unroll
                for (int side = 0; side < 2; side++) {
                    const int M = side ? n : m, Lb = side ? m : n;
                    cplx S = cmake(0., 0.);
                    for (int l = 0; l < nw; l++) {
                        if (in_G(l)) {
                            if (external) {
                                const cplx z = cmul(A[(size_t)al * n2 + M * nw + l], A[(size_t)be * n2 + l * nw + Lb]);
                                S.x += z.y; S.y -= z.x;
                            }
                            continue;
                        }
                        const cplx DMa = Dm(al, M, l), DMb = Dm(be, M, l);
                        if (internal) {
                            const cplx z = cmul(DMa, Dm(be, l, Lb));
                            S.x += z.y; S.y -= z.x;
                        }
                        if (external) {
                            const cplx z = csub(cmul(DMb, A[(size_t)al * n2 + l * nw + Lb]), cmul(DMa, A[(size_t)be * n2 + l * nw + Lb]));
                            S = cadd(S, z);
                        }
                    }
                    if (external) {
                        const cplx o = O[(size_t)comp * n2 + M * nw + Lb];
                        S.x += 0.5 * o.x; S.y += 0.5 * o.y;
                    }
                    val = side ? cmake(val.x + S.x, val.y - S.y) : S;
                }
                T10[(size_t)(m * g + n) * 3 + comp] = val;
            }
            __syncthreads();
            // ---- trace: one warp per (m in inn, c, d)
            for (int item = warp; item < g * 9; item += nwarp) {
                const int m = item / 9, cd = item % 9, c = cd / 3, d = cd - 3 * c;
                const int dc = 3 * d + c, al = WB_ALPHA(c);
                double SH = 0., SO = 0., OV = 0.;   // Re tr of DerMorb_H, of the DerOmega diagonal (before the factor 2), of Omega V
                for (int l = lane; l < nw; l += 32) {
                    const bool lG = in_G(l);
                    if (!lG) {
                        if (internal) SH += 2. * cmul(Dm(al, m, l), T8[((size_t)l * g + m) * 9 + dc]).y;   // Re(-2i z) = 2 Im z
                        if (external) {
                            SH += cmul(Cm[(size_t)c * n2 + m * nw + l], Dm(d, l, m)).x - cmul(Dm(d, m, l), Cm[(size_t)c * n2 + l * nw + m]).x;
                            SO += 0.5 * (cmul(O[(size_t)c * n2 + m * nw + l], Dm(d, l, m)).x - cmul(Dm(d, m, l), O[(size_t)c * n2 + l * nw + m]).x);
                        }
                    } else {
                        if (external) SH += 2. * cmul(A[(size_t)al * n2 + m * nw + l], T9[((size_t)l * g + m) * 9 + dc]).y;
                        OV += cmul(T10[(size_t)(m * g + l) * 3 + c], Vel(d, l * nw + m)).x;
                    }
#pragma unroll
                    for (int t = 0; t < 2; t++) {
                        const int a = t ? WB_BETA(c) : WB_ALPHA(c), b = t ? WB_ALPHA(c) : WB_BETA(c);
                        const double sg = t ? -1. : 1.;
                        if (!lG) {
                            const cplx Dml = Dm(a, m, l);
                            const cplx t1 = T1[((size_t)l * g + m) * 9 + 3 * b + d];
                            if (internal) {
                                const double im = cmul(Dml, t1).y;
                                SH += 2. * sg * Es[l] * im;   // Re(-2i s E_l D dD)
                                SO += sg * im;                // Re(-i s D dD)
                            }
                            if (external) {
                                SH -= 2. * sg * cmul(Dml, T7[((size_t)l * g + m) * 9 + 3 * b + d]).x;
                                SH -= 2. * sg * cmulc(t1, B[(size_t)a * n2 + l * nw + m]).x;   // conj(B_lm^a) dD_lm
                                SO -= sg * (cmul(Dml, T3[((size_t)l * g + m) * 9 + 3 * b + d]).x +
                                            cmul(T2[((size_t)m * nw + l) * 9 + 3 * a + d], A[(size_t)b * n2 + l * nw + m]).x);
                            }
                        } else if (external) {
                            const double im = cmul(A[(size_t)a * n2 + m * nw + l], T4[((size_t)l * g + m) * 9 + 3 * b + d]).y;
                            SH += 2. * sg * Es[l] * im;       // Re(-2i s A_ml E_l dA_lm)
                            SO += sg * im;                    // Re(-i s A_ml dA_lm)
                        }
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    SH += __shfl_xor_sync(0xffffffffu, SH, o);
                    SO += __shfl_xor_sync(0xffffffffu, SO, o);
                    OV += __shfl_xor_sync(0xffffffffu, OV, o);
                }
                if (lane == 0) {
                    if (external) {
                        SH += dCc[(size_t)cd * n2 + m * nw + m].x;
                        SO += 0.5 * dOc[(size_t)cd * n2 + m * nw + m].x;
                    }
                    vals[m * 9 + cd] = SH + Es[m] * 2. * SO + OV;
                }
            }
            __syncthreads();
            for (int cd = threadIdx.x; cd < 9; cd += NT) {
                double s = 0.;
                for (int n = 0; n < g; n++) s += vals[n * 9 + cd];
                Tedge[xe * 9 + cd] = s;
            }
            __syncthreads();
        }
        for (int x = threadIdx.x; x < nw * 9; x += NT) {
            const int n0 = x / 9, cd = x - 9 * n0;
            if (label[n0] != CUDART_INF) ev_val[((size_t)ik * nw + n0) * 9 + cd] = Tedge[g2[n0] * 9 + cd] - Tedge[n0 * 9 + cd];
        }
    }
}
