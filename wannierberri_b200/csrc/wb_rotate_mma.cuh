// Eigenbasis rotation  Xbar = U^dagger X U  on the FP64 tensor cores (mma.sync.m8n8k4.f64, "DMMA") fused with
// the Berry-curvature / orbital-moment formulae; num_wann is a COMPILE-TIME constant (NW <= 24).
//
// Reference: Data_K._rotate (data_K/data_K.py:130-132) = einsum('kba,kbc...,kcd->kad...'); D_H / dEig_inv
// (data_K.py:290-298, 324-326); Omega.nn / trace (formula/covariant.py:161-203, formula/formula.py:76-79);
// Morb_H / Morb_Hpm (covariant.py:375-449) with the non-additive evaluation of StaticCalculator.__call__
// (calculators/static.py:109-117).
//
// One CTA (4 warps) walks k-points blockIdx.x, blockIdx.x + gridDim.x, ...  Per k-point the "items" are channel
// triples of the record (dH | A | B | curl A | C), each fetched by ONE TMA bulk copy into a two-deep staging ring
// (issued two items ahead) and consumed IN PLACE:
//   step 1   Y = X U      [Yr Yi] (3NW x 2NW) = [Xr Xi] (3NW x 2NW) . B1 (2NW x 2NW)
//            the three Cartesian components are stacked along M; the A fragments are read straight from the
//            staging buffer -- a full matrix's complex128 rows ARE the real-ified operand rows, a hermitian
//            (upper-triangle-packed) channel is gathered through per-lane offsets computed once per kernel;
//   step 2   C = U^H Y    [Cr; Ci] (2NW x 3NW) = A2 (2NW x 2NW) . [Yr; Yi] (2NW x 3NW), stacked along N.
// The fragments of B1 (built from U) stay in REGISTERS for the whole k-point and serve both steps: step 2 is
// evaluated as conj(U^T conj(Y)) and the A fragment of U^T's tile (m, k) is the B1 fragment of tile (k, m).  Channels of which only the rotated
// diagonal is needed (curl A, C) stop after step 1 (a dot product with conj(U) per band).
// Y is double buffered, so there is ONE CTA barrier per item.  U and E of the next k-point arrive by TMA as well, the
// band groups are double buffered: no barrier between k-points.
//
// Formula stage: thread = (band n, 1/7 of the partner bands l); partial sums meet in shared memory.
//   Omega:    value(group) = 2 sum_{n in g} [ sum_{l notin g} w(n,l) + sum_{m in g} Im(A_nm,al A_mn,be) + 1/2 Re O_nn ]
//   Morb_Hpm: T(x) of static.py:109-117 telescopes:  T(b) - T(a) = 2 sum_{a <= y < b} t[y],
//             t[y] = sum_{l>y} G[y,l] - sum_{l<y} G[l,y] + sum_{l<y} (Gin[y,l] + Gin[l,y]) + Gin[y,y] + dg[y]
//             where pairs inside one band group are dropped from the G sums (they cancel identically and the
//             reference never forms them: their energy denominators are the near-degenerate ones).
#pragma once
#include "wb_common.cuh"
#include "wb_groups.cuh"
#include "wb_rotate_formula.cuh"
#include "wb_rotate_dmma.cuh"   // wb_dmma, TMA / mbarrier helpers
#include "wb_events_generic.cuh" // WbEventLayout, WbNeeds
#include "wb_eigh_ql.cuh"         // warp_sum
#include <type_traits>

__device__ __forceinline__ void wb_dmma_nv(double& d0, double& d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(d0), "+d"(d1)
        : "d"(a), "d"(b));
}

template <int NW>
struct WbMma {
    static constexpr int N2 = NW * NW;
    static constexpr int NTRI = NW * (NW + 1) / 2;
    static constexpr int KS = (2 * NW + 3) / 4;    // k-steps of 4 over the real-ified inner dimension
    static constexpr int K2 = 4 * KS;
    static constexpr int MT2 = (2 * NW + 7) / 8;   // 8-wide tiles over the real-ified band dimension
    static constexpr int NT3 = (3 * NW + 7) / 8;   // 8-wide tiles over the three stacked components
    static constexpr int TPW = (NT3 + 3) / 4;      // stacked tiles per warp
    static constexpr int ldy_pad(int n) { return (n % 16 == 4 || n % 16 == 12) ? n : ldy_pad(n + 4); }
    static constexpr int LDY = ldy_pad((NT3 * 8 + 3) / 4 * 4);  // conflict-free 64-bit B-fragment loads
    static constexpr int LDC = NW | 1;             // odd: conflict-free transposed complex128 loads
    // staging buffer (doubles): a triple of full matrices; padded rows of the last stacked tile stay in bounds
    static constexpr int STAGE = (6 * N2 > NT3 * 8 * 2 * NW + 8 ? 6 * N2 : NT3 * 8 * 2 * NW + 8);
    // shared-memory layout (offsets in doubles): everything of fixed size first, so that the kernel addresses it
    // with immediates; then the rotated matrices Cs[nslot][NW][LDC] (complex), then the staging slots.
    static constexpr int OFF_U = 0;                          // [2][N2] complex
    static constexpr int OFF_Y = OFF_U + 4 * N2;             // [2][K2][LDY]
    static constexpr int OFF_DG = OFF_Y + 2 * K2 * LDY;      // [2 kinds][3][NW] complex
    static constexpr int OFF_ES = OFF_DG + 12 * NW;          // [2][NW]
    static constexpr int OFF_LABEL = OFF_ES + 2 * NW;        // [2][NW]
    static constexpr int OFF_ROWS = OFF_LABEL + 2 * NW;      // [3][NW]
    static constexpr int OFF_TM = OFF_ROWS + 3 * NW;         // [3][NW]
    static constexpr int OFF_BARS = OFF_TM + 3 * NW;         // 8 mbarriers
    static constexpr int OFF_G12 = OFF_BARS + 8;             // [2][2][NW] shorts
    static constexpr int OFF_CS = (OFF_G12 + (4 * NW + 3) / 4 + 1) / 2 * 2;
};

// items of one k-point (uniform over the grid)
struct WbMmaPlan {
    int nitem;
    int off[6];      // record offset (complex elements) of the triple
    int bytes[6];    // TMA size
    int herm[6];     // upper-triangle packed
    int slot[6];     // >= 0: full rotation into Cs slots slot..slot+2 ; < 0: diagonal only, kind = -1 - slot
    int soff[6];     // staging slot of the item (offset in doubles from the shared-memory base); every item owns
                     // one, refilled for the NEXT k-point
    int r1[6], r2[6]; // warp -> tile rotation of step 1 / step 2 of the item (balances the warps between barriers)
    int nstage;      // staging doubles in total
    int nslot;       // Cs matrices
};

template <int NW>
__host__ inline size_t wb_mma_smem_bytes(const WbMmaPlan& P) {
    using M = WbMma<NW>;
    return ((size_t)M::OFF_CS + 2 * (size_t)P.nslot * NW * M::LDC + (size_t)P.nstage) * sizeof(double);
}

// host: which items a formula mask needs (Omega and/or Morb_Hpm only)
template <int NW>
inline bool wb_mma_make_plan(const WbLayout& L, int mask, int external, WbMmaPlan* P) {
    const int allowed = (1 << 1) | (1 << 2);
    if (!mask || (mask & ~allowed)) return false;
    const bool morb = (mask >> 2) & 1;
    const int nw = L.nw, n2 = nw * nw;
    int n = 0, slot = 0;
    auto add = [&](int off, bool herm, int s) {
        P->off[n] = off;
        P->bytes[n] = 3 * (herm ? L.ntri : n2) * 16;
        P->herm[n] = herm;
        P->slot[n] = s;
        n++;
    };
    add(L.off_dH[0], L.dH_herm, slot); slot += 3;
    if (external) {
        add(L.off_A[0], true, slot); slot += 3;
        if (morb) { add(L.off_B[0], false, slot); slot += 3; }
        add(L.off_O[0], true, -1);
        if (morb) add(L.off_C[0], false, -2);
    }
    for (int i = 0; i < n; i++)
        if (P->off[i] < 0) return false;
    int so = WbMma<NW>::OFF_CS + 2 * slot * NW * WbMma<NW>::LDC;
    const int so0 = so;
    for (int i = 0; i < n; i++) {
        P->soff[i] = so;
        so += P->herm[i] ? 6 * WbMma<NW>::NTRI : WbMma<NW>::STAGE;
        so = (so + 1) / 2 * 2;
    }
    for (int i = 0; i < n; i++) { P->r1[i] = P->herm[i] ? 1 : 0; P->r2[i] = P->herm[i] ? 3 : 2; }
    P->nstage = so - so0;
    P->nitem = n;
    P->nslot = slot;
    return true;
}

// TRIM (needs every rotated matrix that the formulae read on both sides of the diagonal to be hermitian: d_a H packed,
// A hermitised): only the COLUMNS n < nhi of the rotated matrices are formed, nhi = end of the last band group that
// can carry an event (its first band lies below EFmax) -- step 1 multiplies by those columns of U only, step 2 has
// correspondingly fewer stacked column tiles, and the formulae read X_nl as conj(X_ln).  For Fe with E_F up to 22 eV
// that is 12 of 18 bands: 36 % fewer DMMAs.  The number of column tiles is dispatched to compile-time variants of the
// step-1 body (a predicated DMMA would break the interleaving of the accumulator chains).
// DBG (microbenchmarks only): bit 0 skip the formula stage, bit 1 skip the DMMAs, bit 2 skip the TMA waits
template <int NW, bool TRIM = false, int DBG = 0>
__global__ void __launch_bounds__(128, 2)
wb_events_mma_kernel(const cplx* __restrict__ rec, WbLayout L, WbMmaPlan P, long nk, const double* __restrict__ Eall,
                     const cplx* __restrict__ Uall, WbWindow win, WbEventLayout ev, double* __restrict__ ev_label,
                     double* __restrict__ ev_val) {
    using M = WbMma<NW>;
    static_assert(NW % 2 == 0 && NW <= 24, "even num_wann <= 24 (E rows are fetched by 16-byte-granular bulk copies)");
    constexpr int N2 = M::N2, NTRI = M::NTRI, KS = M::KS, K2 = M::K2, MT2 = M::MT2, NT3 = M::NT3, TPW = M::TPW,
                  LDY = M::LDY, LDC = M::LDC;
    // partner-band groups of the formula stage: thread = (band, group), NW x LG threads whose 6 partial sums fit Y[0]
    constexpr int LG_T = 128 / NW, LG_Y = (K2 * LDY) / (6 * NW);
    constexpr int LG = (LG_T < 7 ? LG_T : 7) < LG_Y ? (LG_T < 7 ? LG_T : 7) : LG_Y;
    static_assert(LG >= 1 && NW * LG <= 128 && NW * LG * 6 <= K2 * LDY, "formula stage mapping");
    extern __shared__ __align__(16) double smem_m[];
    cplx* const Ustg = (cplx*)(smem_m + M::OFF_U);
    double* const Yp = smem_m + M::OFF_Y;
    cplx* const Dg = (cplx*)(smem_m + M::OFF_DG);          // rotated diagonals of curl A, C
    double* const Es2 = smem_m + M::OFF_ES;
    double* const label2 = smem_m + M::OFF_LABEL;
    double* const rows = smem_m + M::OFF_ROWS;             // Omega per band
    double* const tm = smem_m + M::OFF_TM;                 // Morb t[y]
    uint64_t* const bars = (uint64_t*)(smem_m + M::OFF_BARS);   // [0..5] staging slots, [6..7] U + E
    short* const g12 = (short*)(smem_m + M::OFF_G12);
    cplx* const Cs = (cplx*)(smem_m + M::OFF_CS);          // [nslot][NW][LDC]
    double* const part = Yp;                               // [128][6] formula partial sums (Y[0] is idle then)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, q = lane & 3;
    const bool f_omega = (ev.mask >> 1) & 1, f_morb = (ev.mask >> 2) & 1;
    const bool internal = ev.internal_terms, external = ev.external_terms;
    const int nitem = P.nitem;

    // ---- one-time setup
    for (int x = threadIdx.x; x < 2 * K2 * LDY; x += 128) Yp[x] = 0.;   // K / N padding of Y stays finite
    for (int x = threadIdx.x; x < P.nstage; x += 128) smem_m[P.soff[0] + x] = 0.;
    if (threadIdx.x == 0) {
        for (int b = 0; b < 8; b++) wb_mbar_init(&bars[b], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    // gather offsets of the A fragments of hermitian (packed) triples: bit 15 = negate
    unsigned hoff[TPW][(KS + 1) / 2];   // two 16-bit entries per register
#pragma unroll
    for (int r = 0; r < TPW; r++)
#pragma unroll
        for (int s = 0; s < KS; s++) {
            int R = 8 * (((warp + 1) & 3) + 4 * r) + g;   // hermitian triples use tile rotation 1 (see below)
            if (R >= 3 * NW) R = 0;
            int m = R / NW, i = R - m * NW;
            int k = 4 * s + q, j = k >> 1, p = k & 1;
            if (j >= NW) j = NW - 1;
            int lo = min(i, j), hi = max(i, j);
            unsigned o = (unsigned)(2 * (m * NTRI + tri_index(lo, hi, NW)) + p) | ((p && j < i) ? 0x8000u : 0u);
            if (s & 1) hoff[r][s >> 1] |= o << 16;
            else hoff[r][s >> 1] = o;
        }
    // per-lane base of the B1 fragment loads from U (complex AoS as doubles)
    const int pk = q & 1, pn = g & 1;
    const int bf_base = 2 * ((q >> 1) * NW + (g >> 1)) + (pk ^ pn);
    const int bf_neg = (pk && !pn) ? (int)0x80000000 : 0;
    __syncthreads();

    const long nmine = (blockIdx.x < nk) ? (nk - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    // producer (thread 0): item `it` of local k-point kl into its own slot
    auto issue_item = [&](long kl, int it) {
        long ik = blockIdx.x + kl * (long)gridDim.x;
        wb_mbar_expect_tx(&bars[it], (uint32_t)P.bytes[it]);
        wb_bulk_g2s(smem_m + P.soff[it], rec + ik * L.E + P.off[it], (uint32_t)P.bytes[it], &bars[it]);
    };
    auto issue_UE = [&](long kl) {
        long ik = blockIdx.x + kl * (long)gridDim.x;
        int b = (int)(kl & 1);
        wb_mbar_expect_tx(&bars[6 + b], (uint32_t)(N2 * 16 + NW * 8));
        wb_bulk_g2s(Ustg + (size_t)b * N2, Uall + ik * N2, (uint32_t)(N2 * 16), &bars[6 + b]);
        wb_bulk_g2s(Es2 + b * NW, Eall + ik * NW, (uint32_t)(NW * 8), &bars[6 + b]);
    };
    if (threadIdx.x == 0 && nmine > 0) {
        issue_UE(0);
        for (int it = 0; it < nitem; it++) issue_item(0, it);
    }

    for (long kl = 0; kl < nmine; kl++) {
        const long ik = blockIdx.x + kl * (long)gridDim.x;
        const int kb = (int)(kl & 1);
        const cplx* Us = Ustg + (size_t)kb * N2;
        const double* Es = Es2 + kb * NW;
        double* label = label2 + kb * NW;
        short* g1 = g12 + kb * 2 * NW;
        short* g2 = g1 + NW;
        wb_mbar_wait(&bars[6 + kb], (uint32_t)((kl >> 1) & 1));
        // band groups: by the warp that has one tile less in the first phase
        if (warp == 3) wb_band_groups_warp(Es, NW, win, g1, g2, label, lane);

        // ---- columns of the rotated matrices that are needed at this k-point (every warp computes it for itself)
        int thi = MT2, NCB = NW;
        if (TRIM) {
            const double En = (lane < NW) ? Es[lane] : CUDART_INF;
            bool brd = (lane < NW) && (lane == 0 || En - Es[lane > 0 ? lane - 1 : 0] > win.degen_thresh);
            if (win.degen_Kramers && (lane & 1)) brd = false;
            const unsigned Bm = __ballot_sync(0xffffffffu, brd);
            const int a0 = 31 - __clz((int)(Bm & ((2u << lane) - 1u)));   // start of the group of band `lane`
            const bool ok = (lane < NW) && (Es[max(a0, 0)] <= win.EFmax);
            const unsigned okm = __ballot_sync(0xffffffffu, ok);
            const int nhi = okm ? 32 - __clz((int)okm) : 1;
            thi = min(MT2, (nhi + 3) >> 2);
            if (thi < MT2) NCB = 4 * thi;
        }
        const int ntl = TRIM ? (3 * NCB + 7) >> 3 : NT3;   // stacked column tiles of step 2

        // ---- B1 fragments: element (kk = 4s + q, nn = 8t + g) of the real-ified U
        double Bf[KS][MT2];
        {
            const double* ud = (const double*)Us + bf_base;
#pragma unroll
            for (int s = 0; s < KS; s++)
#pragma unroll
                for (int t = 0; t < MT2; t++) {
                    const bool safe = (2 * s + 1 < NW) && (4 * t + 3 < NW);
                    double v = 0.;
                    if (safe || ((2 * s + (q >> 1) < NW) && (4 * t + (g >> 1) < NW))) {
                        double u = ud[2 * (2 * s * NW + 4 * t)];
                        v = __hiloint2double(__double2hiint(u) ^ bf_neg, __double2loint(u));
                    }
                    Bf[s][t] = v;
                }
        }
        for (int it = 0; it < nitem; it++) {
            const double* src = smem_m + P.soff[it];
            double* Y = Yp + (it & 1) * K2 * LDY;
            const bool herm = P.herm[it];
            const int slot = P.slot[it];
            if (!(DBG & 4)) wb_mbar_wait(&bars[it], (uint32_t)(kl & 1));
            // ---- step 1: Y = X U on stacked row tiles mt = w1, w1 + 4, ...  The warp -> tile map is rotated from
            // phase to phase so that the warp with one tile less is a different one each time (the four tensor
            // pipes of the SM, shared with the co-resident CTA, stay evenly loaded).
            const int w1 = (warp + P.r1[it]) & 3, w2 = (warp + P.r2[it]) & 3;
            auto step1 = [&](auto thi_c) {
                constexpr int THI = decltype(thi_c)::value;   // column tiles of U that are multiplied
#pragma unroll
            for (int r = 0; r < TPW; r++) {
                const int mt = w1 + 4 * r;
                if (mt < NT3) {
                    double acc[THI][2];
#pragma unroll
                    for (int t = 0; t < THI; t++) acc[t][0] = acc[t][1] = 0.;
                    // all A fragments of the tile first (their shared-memory latency overlaps), then the DMMA chain
                    double av[KS];
                    if (herm) {
#pragma unroll
                        for (int s = 0; s < KS; s++) {
                            const unsigned o = (hoff[r][s >> 1] >> (16 * (s & 1))) & 0xffffu;
                            double a = src[o & 0x7fff];
                            av[s] = __hiloint2double(__double2hiint(a) ^ (int)((o & 0x8000u) << 16), __double2loint(a));
                        }
                    } else {
                        const double* arow = src + (8 * mt + g) * (2 * NW) + q;
#pragma unroll
                        for (int s = 0; s < KS; s++) av[s] = arow[4 * s];
                    }
#pragma unroll
                    for (int s = 0; s < KS; s++)
#pragma unroll
                        for (int t = 0; t < THI; t++)
                            if (!(DBG & 2)) wb_dmma_nv(acc[t][0], acc[t][1], av[s], Bf[s][t]);
                    const int R = 8 * mt + g;
                    if (R < 3 * NW) {
                        const int m = R / NW, i = R - m * NW;
                        double* y0 = Y + (2 * i) * LDY + m * NCB + q;
#pragma unroll
                        for (int t = 0; t < THI; t++)
                            if (4 * t + 3 < NW || 4 * t + q < NW) {
                                y0[4 * t] = acc[t][0];
                                y0[LDY + 4 * t] = -acc[t][1];   // conj(Y), see step 2
                            }
                    }
                }
            }
            };
            if (!TRIM) step1(std::integral_constant<int, MT2>{});
            else switch (thi) {
                case 1: step1(std::integral_constant<int, 1>{}); break;
                case 2: step1(std::integral_constant<int, (MT2 >= 2 ? 2 : MT2)>{}); break;
                case 3: step1(std::integral_constant<int, (MT2 >= 3 ? 3 : MT2)>{}); break;
                case 4: step1(std::integral_constant<int, (MT2 >= 4 ? 4 : MT2)>{}); break;
                case 5: step1(std::integral_constant<int, (MT2 >= 5 ? 5 : MT2)>{}); break;
                default: step1(std::integral_constant<int, MT2>{}); break;
            }
            __syncthreads();   // Y complete; the item's staging slot is free; everybody is done with item it-1
            if (threadIdx.x == 0 && kl + 1 < nmine) {
                issue_item(kl + 1, it);            // refill the slot for the next k-point
                if (it == 0) issue_UE(kl + 1);     // (its U / E buffer was last used by k-point kl-1)
            }
            if (slot < 0) {
                // ---- rotated diagonal only:  sum_i conj(U[i][n]) Y_c[i][n]
                cplx* dst = Dg + (size_t)(-1 - slot) * 3 * NW;
                for (int x = threadIdx.x; x < 3 * NW; x += 128) {
                    const int c = x / NW, n = x - c * NW;
                    cplx a0 = cmake(0., 0.), a1 = cmake(0., 0.);
                    if (!TRIM || n < NCB) {
                        const double* y = Y + c * NCB + n;
#pragma unroll
                        for (int i = 0; i < NW; i += 2) {
                            cfma(a0, Us[i * NW + n], cmake(y[(2 * i) * LDY], y[(2 * i + 1) * LDY]));
                            cfma(a1, Us[(i + 1) * NW + n], cmake(y[(2 * i + 2) * LDY], y[(2 * i + 3) * LDY]));
                        }
                    }
                    dst[x] = cconj(cadd(a0, a1));   // Y holds conj(Y): sum conj(U) Y = conj(sum U conj(Y))
                }
            } else {
                // ---- step 2: C = U^H Y = conj(U^T conj(Y)) on stacked column tiles nt = w2, w2 + 4, ...
                // The real-ified U^T is the transpose of B1: its A fragment of tile (m, k) IS the B1 fragment of
                // tile (k, m), so step 1 stores conj(Y) and the epilogue conjugates back -- no operand fix-up.
#pragma unroll
                for (int r = 0; r < TPW; r++) {
                    const int nt = w2 + 4 * r;
                    if (nt < ntl) {
                        double acc[MT2][2];
#pragma unroll
                        for (int t = 0; t < MT2; t++) acc[t][0] = acc[t][1] = 0.;
                        const double* bcol = Y + q * LDY + 8 * nt + g;
                        double bvv[KS];
#pragma unroll
                        for (int s = 0; s < KS; s++) bvv[s] = bcol[4 * s * LDY];
#pragma unroll
                        for (int s = 0; s < KS; s++)
#pragma unroll
                            for (int t = 0; t < MT2; t++) {
                                if (!(DBG & 2)) wb_dmma_nv(acc[t][0], acc[t][1], Bf[s][t], bvv[s]);
                            }
                        // lane holds C~[rr = 8t + g][8nt + 2q + {0,1}], rr = 2n + (re|im).  Pair the re / im lanes
                        // (g even / odd) so that each stores one complex128:
                        const int cc = 8 * nt + 2 * q + (g & 1);
                        const int m = TRIM ? ((cc >= 2 * NCB) ? 2 : (cc >= NCB) ? 1 : 0) : cc / NW;
                        const int l = cc - m * NCB;
                        cplx* crow = Cs + ((size_t)(slot + m) * NW + (g >> 1)) * LDC + l;
#pragma unroll
                        for (int t = 0; t < MT2; t++) {
                            double send = (g & 1) ? acc[t][0] : acc[t][1];
                            double recv = __shfl_xor_sync(0xffffffffu, send, 4);
                            cplx v = (g & 1) ? cmake(recv, -acc[t][1]) : cmake(acc[t][0], -recv);
                            if (cc < 3 * NCB && l < NW && (4 * t + 3 < NW || 4 * t + (g >> 1) < NW)) crow[(size_t)4 * t * LDC] = v;
                        }
                    }
                }
            }
        }
        __syncthreads();   // all rotated matrices are in Cs / Dg

        // ---- formula stage: thread = (band n, partner bands l = lg, lg + 7, ...)
#define WB_C(sl, a, n_, l_) Cs[((size_t)((sl) + (a)) * NW + (n_)) * LDC + (l_)]
#define WB_V(a, n_, l_) WB_C(0, a, n_, l_)
#define WB_A(a, n_, l_) WB_C(3, a, n_, l_)
#define WB_B(a, n_, l_) WB_C(6, a, n_, l_)
        if (!(DBG & 1)) {
            double om[3] = {0., 0., 0.}, tt[3] = {0., 0., 0.};
            const int n = threadIdx.x / LG, lg = threadIdx.x - LG * n;
            const int ga = (n < NW) ? g1[n] : -1, gb = (n < NW) ? g2[n] : -1;
            if (ga >= 0) {
                const double En = Es[n];
#pragma unroll
                for (int jj = 0; jj < (NW + LG - 1) / LG; jj++) {
                    const int l = lg + LG * jj;
                    if (l < NW) {
                        const bool same = (l >= ga && l < gb);
                        const double El = Es[l];
                        cplx Anl[3], Aln[3];
                        if (external) {
#pragma unroll
                            for (int a = 0; a < 3; a++) {
                                Aln[a] = WB_A(a, l, n);
                                Anl[a] = TRIM ? cconj(Aln[a]) : WB_A(a, n, l);
                            }
                        }
                        if (!same) {
                            const double inv = wb_deinv(En, El);   // 1/(E_n - E_l); D_nl = -V_nl inv, D_ln = +V_ln inv
                            cplx Dnl[3], Dln[3];
#pragma unroll
                            for (int a = 0; a < 3; a++) {
                                const cplx Vln = WB_V(a, l, n);
                                Dnl[a] = cscale(-inv, TRIM ? cconj(Vln) : WB_V(a, n, l));
                                Dln[a] = cscale(inv, Vln);
                            }
                            if (f_omega) {
#pragma unroll
                                for (int c = 0; c < 3; c++) {
                                    const int al = WB_ALPHA(c), be = WB_BETA(c);
                                    double v = 0.;
                                    if (internal) v += cmul(Dnl[al], Dln[be]).y;
                                    if (external) v += -cmul(Dnl[al], Aln[be]).x + cmul(Dnl[be], Aln[al]).x;
                                    om[c] += v;
                                }
                            }
                            if (f_morb) {
                                // pair {lo < hi}: G[lo,hi]; row n collects +G[n,l] (l > n) or -G[l,n] (l < n)
                                const bool up = l > n;
                                const double Elo = up ? En : El, Esum = En + El;
#pragma unroll
                                for (int c = 0; c < 3; c++) {
                                    const int al = WB_ALPHA(c), be = WB_BETA(c);
                                    const cplx Da = up ? Dnl[al] : Dln[al], Db = up ? Dnl[be] : Dln[be];
                                    const cplx Dtb = up ? Dln[be] : Dnl[be];
                                    double v = 0.;
                                    if (internal) v += Esum * cmul(Da, Dtb).y;
                                    if (external) {
                                        const cplx Bb = up ? WB_B(be, l, n) : WB_B(be, n, l);
                                        const cplx Ba = up ? WB_B(al, l, n) : WB_B(al, n, l);
                                        const cplx Ab = up ? Aln[be] : Anl[be];
                                        const cplx Aa = up ? Aln[al] : Anl[al];
                                        v += -cmul(Da, Bb).x + cmul(Db, Ba).x;
                                        v += Elo * (-cmul(Da, Ab).x + cmul(Db, Aa).x);
                                    }
                                    tt[c] += up ? v : -v;
                                }
                            }
                        }
                        if (external) {
#pragma unroll
                            for (int c = 0; c < 3; c++) {
                                const int al = WB_ALPHA(c), be = WB_BETA(c);
                                const double aa = cmul(Anl[al], Aln[be]).y;   // Im(A_nl,al A_ln,be)
                                if (f_omega && same) om[c] += aa;
                                if (f_morb) {
                                    const double Esum = En + El;
                                    if (l < n) tt[c] += Esum * (aa + cmul(Aln[al], Anl[be]).y);
                                    else if (l == n)
                                        tt[c] += Esum * aa + 0.5 * Dg[(3 + c) * NW + n].x + 0.5 * En * Dg[c * NW + n].x;
                                }
                            }
                        }
                    }
                }
            }
            if (n < NW) {
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    part[threadIdx.x * 6 + c] = om[c];
                    part[threadIdx.x * 6 + 3 + c] = tt[c];
                }
            }
        }
        __syncthreads();
        // totals per band: rows[c][n] (Omega), tm[c][n] (Morb)
        for (int x = threadIdx.x; x < 6 * NW; x += 128) {
            const int h = x / (3 * NW), y = x - h * 3 * NW;
            const int c = y / NW, n = y - c * NW;
            double s = 0.;
#pragma unroll
            for (int lg = 0; lg < LG; lg++) s += part[(n * LG + lg) * 6 + 3 * h + c];
            if (h == 0) rows[y] = s + (external ? 0.5 * Dg[c * NW + n].x : 0.);
            else tm[y] = s;
        }
        __syncthreads();
        // events: thread = (group start x, component c)
        for (int y = threadIdx.x; y < 3 * NW; y += 128) {
            const int x = y / 3, c = y - 3 * x;
            const double lab = label[x];
            if (c == 0) ev_label[ik * NW + x] = lab;
            if (lab != CUDART_INF) {
                const int bnd = g2[x];
                double* out = ev_val + (ik * NW + x) * ev.NC;
                if (f_omega) {
                    double s = 0.;
                    for (int n = x; n < bnd; n++) s += rows[c * NW + n];
                    out[ev.off[1] + c] = 2. * s;
                }
                if (f_morb) {
                    double s = 0.;
                    for (int n = x; n < bnd; n++) s += tm[c * NW + n];
                    out[ev.off[2] + c] = 2. * s;
                }
            }
        }
#undef WB_C
#undef WB_V
#undef WB_A
#undef WB_B
    }
}
