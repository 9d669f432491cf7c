// Generic rotation + formula kernel: every formula of the static path, any nw whose working set fits
// shared memory.  One CTA per k-point; rotated matrices live in shared memory.
//
// Reference: Data_K._rotate (data_K/data_K.py:130-132); D_H / dEig_inv (:290-298,:324-326);
// Omega (formula/covariant.py:161-203), Morb_H (:375-421), Morb_Hpm (:424-449), Velocity (:321-328),
// Spin (:331-335), FormulaProduct (formula/formula.py:121-152), VelOmega/VelHplus/VelSpin
// (covariant.py:798-814); additive / non-additive evaluation loop of StaticCalculator.__call__
// (calculators/static.py:102-117).
//
// Output "events": slot n of ev_label[k][nw] / ev_val[k][nw][NC] is used iff a band group starts at band n.
// ev_val holds the values of ALL requested formulae side by side (offsets in WbEventLayout), so that one pass
// over the rotated matrices serves e.g. AHC + Morb, or BerryDipole + GME_orb + GME_spin.
#pragma once
#include "wb_common.cuh"
#include "wb_groups.cuh"
#include "wb_rotate_formula.cuh"

struct WbEventLayout {
    int mask;          // bit f set: formula f requested
    int off[32];       // offset of formula f inside the NC values of an event
    int NC;            // values per event
    int internal_terms, external_terms;
};

// which rotated matrices a formula mask needs
struct WbNeeds {
    bool V, A, B, Odiag, Cdiag, Sdiag, Oblk, Cblk, Sblk, D, Wdiag, Wblk;
};
__host__ __device__ inline WbNeeds wb_needs(int mask, int external) {
    auto has = [&](int f) { return (mask >> f) & 1; };
    bool omega = has(1), morb = has(2), vo = has(3), vh = has(4), vs = has(5), spin = has(6), vv = has(8), im = has(9);
    WbNeeds n;
    n.D = omega || morb || vo || vh;
    n.V = n.D || vs || vv || im;
    n.Wdiag = im;
    n.Wblk = false;
    n.A = n.D && external;
    n.B = (morb || vh) && external;
    n.Oblk = (vo || vh) && external;
    n.Cblk = vh && external;
    n.Sblk = vs;
    n.Odiag = (omega || morb) && external && !n.Oblk;
    n.Cdiag = morb && external && !n.Cblk;
    n.Sdiag = spin && !n.Sblk;
    return n;
}

// number of full nw x nw complex matrices kept in shared memory (besides U, X, Y)
__host__ __device__ inline int wb_generic_nfull(const WbNeeds& n) {
    return 3 * ((int)n.V + (int)n.A + (int)n.B + (int)n.Oblk + (int)n.Cblk + (int)n.Sblk);
}
__host__ inline size_t wb_generic_smem_bytes(int nw, int mask, int external) {
    WbNeeds n = wb_needs(mask, external);
    size_t cplx_el = (size_t)(3 + wb_generic_nfull(n)) * nw * nw + 15 * (size_t)nw;  // U, X, Y, fulls, diag O/C/S/W
    size_t dbl = 2 * (size_t)nw + 3 * (size_t)nw + 45 * (size_t)nw + 3 * (size_t)(nw + 1) + (size_t)nw * nw;  // Es, label, rows[3], prod[36], Tedge, inv
    return cplx_el * sizeof(cplx) + dbl * sizeof(double) + 2 * nw * sizeof(short) + 64;
}

// Pointers to the rotated matrices of one k-point ([3][nw][nw] blocks, or [3][nw] rotated diagonals) and the
// work arrays of the formula stage.  The matrices may live in shared memory (wb_events_generic_kernel) or in
// global memory (wb_events_xbar_kernel, after the batched DMMA rotation of wb_rotate_gemm.cuh).
struct WbRotated {
    const cplx *Vb, *Ab, *Bb, *Ob, *Cb, *Sb;   // full blocks
    const cplx *Od, *Cd, *Sd;                  // diagonals (used when the corresponding block is absent)
    const cplx *Wb, *Wd;                       // second derivative of H: [6][nw*nw] blocks or [6][nw] diagonals
    const double* Es;
    const double* label;
    double *rows, *prod, *Tedge;
    double* Mx;                                // [3][nw*nw] scratch of the non-additive Morb evaluation
    double* inv;                               // [nw][nw] 1/(E_n - E_l) with the reference's cutoff (dEig_inv), filled here
    const short *g1, *g2;
};

// Formula stage: band-group traces of every requested formula -> events of k-point ik.  All threads of the CTA
// call it after the rotated matrices are visible; it ends WITHOUT a trailing barrier.
template <int NT>
__device__ __forceinline__ void wb_formula_events(const WbRotated& R, const WbNeeds& need, int nw, long ik,
                                                  const WbEventLayout& ev, double* __restrict__ ev_label,
                                                  double* __restrict__ ev_val) {
    const int n2 = nw * nw;
    const cplx *Vb = R.Vb, *Ab = R.Ab, *Bb = R.Bb, *Ob = R.Ob, *Cb = R.Cb, *Sb = R.Sb, *Od = R.Od, *Cd = R.Cd, *Sd = R.Sd;
    const double *Es = R.Es, *label = R.label;
    double *rows = R.rows, *prod = R.prod, *Tedge = R.Tedge;
    const short *g1 = R.g1, *g2 = R.g2;
    const bool f_omega = (ev.mask >> 1) & 1, f_morb = (ev.mask >> 2) & 1, f_vo = (ev.mask >> 3) & 1,
               f_vh = (ev.mask >> 4) & 1, f_vs = (ev.mask >> 5) & 1, f_spin = (ev.mask >> 6) & 1, f_vv = (ev.mask >> 8) & 1,
               f_im = (ev.mask >> 9) & 1;
    const bool internal = ev.internal_terms, external = ev.external_terms;
    // diagonal accessors that work for both storage modes
    auto Odg = [&](int c, int n) { return need.Oblk ? Ob[c * n2 + n * nw + n] : Od[c * nw + n]; };
    auto Cdg = [&](int c, int n) { return need.Cblk ? Cb[c * n2 + n * nw + n] : Cd[c * nw + n]; };
    auto Sdg = [&](int c, int n) { return need.Sblk ? Sb[c * n2 + n * nw + n] : Sd[c * nw + n]; };
    auto Wdg = [&](int c6, int n) { return need.Wblk ? R.Wb[c6 * n2 + n * nw + n] : R.Wd[c6 * nw + n]; };
    // dEig_inv once per k-point (data_K.py:290-298): the formulae below read every element many times
    double* const inv = R.inv;
    for (int x = threadIdx.x; x < n2; x += NT) inv[x] = wb_deinv(Es[x / nw], Es[x % nw]);
    __syncthreads();
    auto Dm = [&](int a, int n, int l) {  // D_nl,a = -Vbar_nl,a / (E_n - E_l)
        return cscale(-inv[n * nw + l], Vb[a * n2 + n * nw + l]);
    };

    // ---- S-type sums for a pair (M, Lb) of one group [ga, gb), component c:
    //   S  = -i sum_l D_Ml,al D_lL,be + 1/2 O_ML - sum_l D_Ml,al A_lL,be + sum_l D_Ml,be A_lL,al - i sum_m A_Mm,al A_mL,be
    //   Sh = same with E_l, C, B, E_m  (Morb_H)
    // Evaluated COOPERATIVELY by a group of LG = 8 adjacent lanes: the sum over the partner bands l is strided over
    // the lanes, lane 0 adds the in-group and the diagonal terms, a butterfly over the group leaves the totals on
    // every lane.  (A thread per (band, component) with serial O(nw) sums left most of the CTA idle.)
    constexpr int LG = 8;
    const int lgid = threadIdx.x % LG;
    const unsigned gmask = 0xFFu << (8 * ((threadIdx.x & 31) / LG));
    auto S_pair = [&](int M, int Lb, int ga, int gb, int c, bool want_h, cplx& S, cplx& Sh) {
        const int al = WB_ALPHA(c), be = WB_BETA(c);
        S = cmake(0., 0.);
        Sh = cmake(0., 0.);
        for (int l = lgid; l < nw; l += LG) {   // (unrolling by 4 for more loads in flight was measured: slower, 37.5 -> 39.8 ms)
            if (l >= ga && l < gb) continue;
            cplx DMa = Dm(al, M, l), DMb = Dm(be, M, l);
            if (internal) {
                cplx z = cmul(DMa, Dm(be, l, Lb));  // -i z
                S.x += z.y; S.y -= z.x;
                if (want_h) { Sh.x += Es[l] * z.y; Sh.y -= Es[l] * z.x; }
            }
            if (external) {
                cplx z = csub(cmul(DMb, Ab[al * n2 + l * nw + Lb]), cmul(DMa, Ab[be * n2 + l * nw + Lb]));
                S = cadd(S, z);
                if (want_h) {
                    cplx zh = csub(cmul(DMb, Bb[al * n2 + l * nw + Lb]), cmul(DMa, Bb[be * n2 + l * nw + Lb]));
                    Sh = cadd(Sh, zh);
                }
            }
        }
        if (external && lgid == 0) {
            for (int m = ga; m < gb; m++) {
                cplx z = cmul(Ab[al * n2 + M * nw + m], Ab[be * n2 + m * nw + Lb]);  // -i z
                S.x += z.y; S.y -= z.x;
                if (want_h) { Sh.x += Es[m] * z.y; Sh.y -= Es[m] * z.x; }
            }
            cplx o = (M == Lb) ? Odg(c, M) : (need.Oblk ? Ob[c * n2 + M * nw + Lb] : cmake(0., 0.));
            S.x += 0.5 * o.x; S.y += 0.5 * o.y;
            if (want_h) {
                cplx cc = (M == Lb) ? Cdg(c, M) : (need.Cblk ? Cb[c * n2 + M * nw + Lb] : cmake(0., 0.));
                Sh.x += 0.5 * cc.x; Sh.y += 0.5 * cc.y;
            }
        }
#pragma unroll
        for (int o = 1; o < LG; o <<= 1) {
            S.x += __shfl_xor_sync(gmask, S.x, o);
            S.y += __shfl_xor_sync(gmask, S.y, o);
            if (want_h) {   // uniform
                Sh.x += __shfl_xor_sync(gmask, Sh.x, o);
                Sh.y += __shfl_xor_sync(gmask, Sh.y, o);
            }
        }
    };

    // ---- additive traces and products: one lane group per (band M, component c).  Only the bands of a group do work
    // and they form (a subset of) one contiguous range of the sorted bands: the items are dealt over that range, so that
    // all lane groups of the CTA are busy (Te, Fermi-surface window: 9 of 24 bands -- dealing all 3 nw items left 10 of
    // the 16 lane groups idle in every round).
    int mlo = nw, mhi = 0;
    for (int m = 0; m < nw; m++)
        if (g1[m] >= 0) { mlo = min(mlo, m); mhi = m + 1; }
    const int mcnt = max(mhi - mlo, 0);
    for (int x = threadIdx.x / LG; x < 3 * mcnt; x += NT / LG) {
        int c = x / mcnt, M = mlo + x % mcnt;
        double tr_omega = 0.;
        double pv[3] = {0., 0., 0.}, ph[3] = {0., 0., 0.}, ps[3] = {0., 0., 0.}, pw[3] = {0., 0., 0.}, pm[3] = {0., 0., 0.};
        if (g1[M] >= 0) {
            const int ga = g1[M], gb = g2[M];
            if (f_omega) {
                cplx S, Sh;
                S_pair(M, M, ga, gb, c, false, S, Sh);
                tr_omega = 2. * S.x;
            }
            if (f_vo || f_vh) {
                for (int Lb = ga; Lb < gb; Lb++) {
                    cplx S1, Sh1, S2, Sh2;
                    S_pair(M, Lb, ga, gb, c, f_vh, S1, Sh1);
                    if (Lb == M) { S2 = S1; Sh2 = Sh1; }
                    else S_pair(Lb, M, ga, gb, c, f_vh, S2, Sh2);
                    cplx Om = cadd(S1, cconj(S2));  // Omega_c[M, Lb]
                    cplx F = Om;
                    for (int a = 0; a < 3; a++) {  // sum_L V_LM,a F[M,L]  -> contribution of row M
                        cplx v = Vb[a * n2 + Lb * nw + M];
                        pv[a] += cmul(v, Om).x;
                    }
                    if (f_vh) {
                        cplx Hh = cadd(Sh1, cconj(Sh2));
                        double eav = 0.5 * (Es[M] + Es[Lb]);
                        F = cmake(Hh.x + eav * Om.x, Hh.y + eav * Om.y);
                        for (int a = 0; a < 3; a++) ph[a] += cmul(Vb[a * n2 + Lb * nw + M], F).x;
                    }
                }
            }
            if (f_vs) {
                for (int Lb = ga; Lb < gb; Lb++)
                    for (int a = 0; a < 3; a++) ps[a] += cmul(Vb[a * n2 + Lb * nw + M], Sb[c * n2 + M * nw + Lb]).x;
            }
            if (f_im) {
                // InvMass (elementary.py:28-34, formula.py:108-112), trace over the group, b = c:
                //   W_MM,bd + sum_{l notin group} ( -D_Ml,d V_lM,b + V_Ml,b D_lM,d ) = W_MM,bd + sum_l (V_Ml,d V_lM,b + V_Ml,b V_lM,d) / (E_M - E_l)
                for (int l = lgid; l < nw; l += LG) {
                    if (l >= ga && l < gb) continue;
                    const double iv = inv[M * nw + l];
                    const cplx VMlb = Vb[c * n2 + M * nw + l], VlMb = Vb[c * n2 + l * nw + M];
                    for (int d = 0; d < 3; d++)
                        pm[d] += iv * (cmul(Vb[d * n2 + M * nw + l], VlMb).x + cmul(VMlb, Vb[d * n2 + l * nw + M]).x);
                }
                for (int d = 0; d < 3; d++) {
#pragma unroll
                    for (int o = 1; o < LG; o <<= 1) pm[d] += __shfl_xor_sync(gmask, pm[d], o);
                    pm[d] += Wdg(wb_sym6(c, d), M).x;
                }
            }
            if (f_vv) {   // VelVel: sum_{L in group} V_LM,a V_ML,c
                for (int Lb = ga; Lb < gb; Lb++)
                    for (int a = 0; a < 3; a++) pw[a] += cmul(Vb[a * n2 + Lb * nw + M], Vb[c * n2 + M * nw + Lb]).x;
            }
        }
        if (lgid == 0) {
            rows[c * nw + M] = tr_omega;
            for (int a = 0; a < 3; a++) {  // component (a, b = c) of the rank-2 products
                prod[M * 45 + a * 3 + c] = pv[a];
                prod[M * 45 + 9 + a * 3 + c] = ph[a];
                prod[M * 45 + 18 + a * 3 + c] = ps[a];
                prod[M * 45 + 27 + a * 3 + c] = pw[a];
                prod[M * 45 + 36 + c * 3 + a] = pm[a];   // InvMass [b = c][d = a]
            }
        }
    }
    // ---- non-additive Morb_Hpm (static.py:109-117): T(x) = trace with inn = 0..x-1, out = x..nw-1,
    //      value of group (a, b) = T(b) - T(a).  With the pair quantities
    //        G[n,l]   = (E_l+E_n) Re(-i D_nl,al D_ln,be) + Re(-D_nl,al B_ln,be + D_nl,be B_ln,al)
    //                   + E_n Re(-D_nl,al A_ln,be + D_nl,be A_ln,al)                      (n in inn, l in out)
    //        Gin[n,m] = (E_m+E_n) Re(-i A_nm,al A_mn,be)                                  (n, m in inn)
    //        dg[n]    = 1/2 Re C_nn + 1/2 E_n Re O_nn
    //      T(x) = 2 [ sum_{n<x<=l} G[n,l] + sum_{n,m<x} Gin[n,m] + sum_{n<x} dg[n] ].
    //      One matrix Mx holds G in its upper triangle, Gin[n,m]+Gin[m,n] in the lower, Gin[n,n]+dg[n] on the diagonal.
    if (f_morb) {
        __syncthreads();
        double* Mx = R.Mx;
        for (int x = threadIdx.x; x < n2; x += NT) {
            int n = x / nw, l = x % nw;
            for (int c = 0; c < 3; c++) {
                const int al = WB_ALPHA(c), be = WB_BETA(c);
                double v = 0.;
                if (n < l) {
                    cplx Dna = Dm(al, n, l), Dnb = Dm(be, n, l);
                    if (internal) v += (Es[l] + Es[n]) * cmul(Dna, Dm(be, l, n)).y;
                    if (external) {
                        v += -cmul(Dna, Bb[be * n2 + l * nw + n]).x + cmul(Dnb, Bb[al * n2 + l * nw + n]).x;
                        v += Es[n] * (-cmul(Dna, Ab[be * n2 + l * nw + n]).x + cmul(Dnb, Ab[al * n2 + l * nw + n]).x);
                    }
                } else if (external) {
                    // here (row n, col l) with l <= n holds the in-part of the pair {l, n}
                    double q1 = (Es[l] + Es[n]) * cmul(Ab[al * n2 + n * nw + l], Ab[be * n2 + l * nw + n]).y;
                    if (n == l) v = q1 + 0.5 * Cdg(c, n).x + 0.5 * Es[n] * Odg(c, n).x;
                    else v = q1 + (Es[l] + Es[n]) * cmul(Ab[al * n2 + l * nw + n], Ab[be * n2 + n * nw + l]).y;
                }
                Mx[c * n2 + x] = v;
            }
        }
        __syncthreads();
        for (int y = threadIdx.x; y < 3 * (nw + 1); y += NT) {
            int c = y / (nw + 1), xe = y % (nw + 1);
            double t = 0.;
            bool is_edge = (xe == nw) || (xe < nw && g1[xe] == xe) || (xe > 0 && g2[xe - 1] == xe);
            if (is_edge) {
                for (int n = 0; n < xe; n++) {
                    for (int l = xe; l < nw; l++) t += Mx[c * n2 + n * nw + l];
                    for (int m = 0; m <= n; m++) t += Mx[c * n2 + n * nw + m];
                }
            }
            Tedge[c * (nw + 1) + xe] = 2. * t;
        }
    }
    __syncthreads();
    // ---- events
    for (int x = threadIdx.x; x < nw; x += NT) {
        double lab = label[x];
        ev_label[ik * nw + x] = lab;
        if (lab != CUDART_INF) {
            const int b = g2[x];
            double* out = ev_val + (ik * nw + x) * ev.NC;
            if (f_omega)
                for (int c = 0; c < 3; c++) {
                    double s = 0.;
                    for (int n = x; n < b; n++) s += rows[c * nw + n];
                    out[ev.off[1] + c] = s;
                }
            if (f_morb)
                for (int c = 0; c < 3; c++) out[ev.off[2] + c] = Tedge[c * (nw + 1) + b] - Tedge[c * (nw + 1) + x];
            if (f_spin)
                for (int c = 0; c < 3; c++) {
                    double s = 0.;
                    for (int n = x; n < b; n++) s += Sdg(c, n).x;
                    out[ev.off[6] + c] = s;
                }
            if (f_vo || f_vh || f_vs || f_vv || f_im)
                for (int ab = 0; ab < 9; ab++) {
                    double so = 0., sh = 0., ss = 0., sv = 0., sm = 0.;
                    for (int n = x; n < b; n++) {
                        so += prod[n * 45 + ab];
                        sh += prod[n * 45 + 9 + ab];
                        ss += prod[n * 45 + 18 + ab];
                        sv += prod[n * 45 + 27 + ab];
                        sm += prod[n * 45 + 36 + ab];
                    }
                    if (f_im) out[ev.off[9] + ab] = sm;
                    if (f_vo) out[ev.off[3] + ab] = so;
                    if (f_vh) out[ev.off[4] + ab] = sh;
                    if (f_vs) out[ev.off[5] + ab] = ss;
                    if (f_vv) out[ev.off[8] + ab] = sv;
                }
        }
    }
}

template <int NT>
__global__ void __launch_bounds__(NT)
wb_events_generic_kernel(const cplx* __restrict__ rec, WbLayout L, long nk, const double* __restrict__ Eall,
                         const cplx* __restrict__ Uall, WbWindow win, WbEventLayout ev,
                         double* __restrict__ ev_label, double* __restrict__ ev_val) {
    extern __shared__ cplx smem_g[];
    const int nw = L.nw, n2 = nw * nw;
    const WbNeeds need = wb_needs(ev.mask, ev.external_terms);
    cplx* Us = smem_g;
    cplx* Xs = Us + n2;
    cplx* Ys = Xs + n2;
    cplx* p = Ys + n2;
    cplx* Vb = p; if (need.V) p += 3 * n2;
    cplx* Ab = p; if (need.A) p += 3 * n2;
    cplx* Bb = p; if (need.B) p += 3 * n2;
    cplx* Ob = p; if (need.Oblk) p += 3 * n2;
    cplx* Cb = p; if (need.Cblk) p += 3 * n2;
    cplx* Sb = p; if (need.Sblk) p += 3 * n2;
    cplx* Od = p; p += 3 * nw;
    cplx* Cd = p; p += 3 * nw;
    cplx* Sd = p; p += 3 * nw;
    cplx* Wd = p; p += 6 * nw;
    double* Es = (double*)p;
    double* label = Es + nw;
    double* rows = label + nw;         // [3][nw]   per-band Omega trace terms
    double* prod = rows + 3 * nw;      // [nw][45]  per-band partial products: VelOmega 9 | VelHplus 9 | VelSpin 9 | VelVel 9 | InvMass 9
    double* Tedge = prod + 45 * nw;    // [3][nw+1] cumulative non-additive traces
    double* invtab = Tedge + 3 * (nw + 1);   // [nw][nw]
    short* g1 = (short*)(invtab + n2);
    short* g2 = g1 + nw;

    for (long ik = blockIdx.x; ik < nk; ik += gridDim.x) {
        const cplx* r = rec + ik * L.E;
        for (int x = threadIdx.x; x < n2; x += NT) Us[x] = Uall[ik * n2 + x];
        for (int x = threadIdx.x; x < nw; x += NT) Es[x] = Eall[ik * nw + x];
        __syncthreads();
        if (nw <= 32) {
            if (threadIdx.x < 32) wb_band_groups_warp(Es, nw, win, g1, g2, label, threadIdx.x);
        } else if (threadIdx.x == 0) wb_band_groups(Es, nw, win, g1, g2, label);
        // ---- rotations
        auto rotate_full = [&](const int* offs, bool herm, cplx* dst) {
            for (int a = 0; a < 3; a++) {
                wb_load_channel<NT>(r, offs[a], herm, Xs, nw);
                wb_rotate_smem<NT>(Us, Xs, Ys, dst + a * n2, nw);
            }
        };
        auto rotate_diag = [&](const int* offs, bool herm, cplx* dst, int ncomp = 3) {
            for (int c = 0; c < ncomp; c++) {
                wb_load_channel<NT>(r, offs[c], herm, Xs, nw);
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
                    dst[c * nw + n] = acc;
                }
                __syncthreads();
            }
        };
        if (need.V) rotate_full(L.off_dH, L.dH_herm, Vb);
        if (need.A) rotate_full(L.off_A, true, Ab);
        if (need.B) rotate_full(L.off_B, false, Bb);
        if (need.Oblk) rotate_full(L.off_O, true, Ob);
        if (need.Cblk) rotate_full(L.off_C, false, Cb);
        if (need.Sblk) rotate_full(L.off_S, true, Sb);
        if (need.Odiag) rotate_diag(L.off_O, true, Od);
        if (need.Cdiag) rotate_diag(L.off_C, false, Cd);
        if (need.Sdiag) rotate_diag(L.off_S, true, Sd);
        if (need.Wdiag) rotate_diag(L.off_W, L.dH_herm, Wd, 6);
        __syncthreads();
        WbRotated R;
        R.Vb = Vb; R.Ab = Ab; R.Bb = Bb; R.Ob = Ob; R.Cb = Cb; R.Sb = Sb; R.Od = Od; R.Cd = Cd; R.Sd = Sd;
        R.Wb = nullptr; R.Wd = Wd;
        R.Es = Es; R.label = label; R.rows = rows; R.prod = prod; R.Tedge = Tedge;
        R.Mx = (double*)Xs;   // [3][n2] doubles = 1.5 n2 complex: Xs and half of Ys, both free now
        R.inv = invtab;
        R.g1 = g1; R.g2 = g2;
        wb_formula_events<NT>(R, need, nw, ik, ev, ev_label, ev_val);
        __syncthreads();
    }
}
