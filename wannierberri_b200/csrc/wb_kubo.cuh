// Kubo (frequency x Fermi-level) scans: OpticalConductivity and JDOS of the reference's DynamicCalculator.
//
// Reference: DynamicCalculator.__call__ (calculators/dynamic.py:71-114) evaluates, per k-point, a dense
//   restot[omega, Ef, ab] += sum_pairs factor_omega[omega, pair] * factor_Efermi[pair, Ef] * matrix_elements[pair, ab]
// over all ordered pairs of degenerate band groups.  At kBT = 0 the Fermi factor f(E2) - f(E1) of a pair with
// group energies lo < hi is -1 (+1 for the reversed pair) exactly on the Fermi levels lo <= Ef < hi and zero
// elsewhere: with s_g = first index of the (ascending) Efermi array with Ef >= E_g, the pair {i < j} contributes
//   X_ij(omega, ab) = -W1 M_ij[ab] + W2 M_ij[ba]      on the bins  s_i <= bin < s_j .
// In difference form along Efermi that is  +X_ij at bin s_i  and  -X_ij at bin s_j, i.e. per GROUP t one value
//   Y_t = sum_{u > t, s_u > s_t} X_tu - sum_{u < t, s_u < s_t} X_ut      added at bin s_t,
// which is accumulated in REGISTERS over the partners u and costs one global reduction per (group, omega, ab).
//   wb_kubo_entries_kernel     (CTA per k-point) band groups (data_K.py:172-186 with the window -inf..inf), the
//                              oriented visits (t, u) in owner order, their matrix elements
//                              M[ab] = i sum_{m in lo-group, n in hi-group} A_mn,a A_nm,b with the generalised Berry
//                              connection A = Abar + i D_H (data_K.py:328-334; Formula_OptCond, dynamic.py:170-181)
//                              from the rotated matrices of wb_rotate_gemm.cuh, sign and K-block weight folded in;
//   wb_kubo_accumulate_kernel  CTA = (tile of omega, slice of k-points); thread = (omega, ab, re|im).  Entries are
//                              staged in shared memory in chunks, the frequency factors W of (entry, omega) are
//                              computed once per chunk and shared by the 18 components;
//   wb_kubo_finalize_kernel    running sum over Efermi, scale, transpose to the reference's [Ef][omega][3][3].
// Optical conductivity: X_ij[ab] = -W1 M[ab] + W2 M[ba] couples a component only to its transpose, so the entries
// carry M in the symmetric / antisymmetric basis (slot ab = M[ab] + M[ba], slot ba = M[ab] - M[ba] for a < b, the
// diagonal as is) and the accumulator holds P = X[ab] + X[ba] = (W2 - W1) S and Q = X[ab] - X[ba] = -(W1 + W2) A:
// ONE complex multiply per (entry, omega, component) instead of two; the finalize kernel changes the basis back.
// The reference's dense contraction costs n_omega * n_pair * n_Ef * 9 complex multiply-adds per k-point; this form
// n_omega * n_pair * 9 * 2 (two visits per pair).
#pragma once
#include "wb_common.cuh"
#include "wb_groups.cuh"
#include "wb_rotate_formula.cuh"

constexpr int WB_KUBO_ENT = 20;    // doubles per entry: Delta | owner bin | M[9] complex (signed, weighted)
constexpr int WB_SHC_ENT = 56;     // spin Hall: Delta | owner bin | M_(lo,hi)[27] | M_(hi,lo)[27] real
constexpr int WB_KUBO_CHUNK = 32;  // entries staged per step of the accumulation kernel
constexpr int WB_KUBO_WT = 32;     // omega values per CTA of the accumulation kernel

// kinds (WbKuboParams::kind = WBGPU_KUBO_*): 0 optical conductivity, 1 JDOS, 2 spin Hall, 3 shift current, 4 injection
// current; the rank-3 kinds (>= 2) share the entry size and the thread layout of the accumulation kernel, whose template
// argument is 0, 1, 2 (real matrix elements: spin Hall, shift current) or 3 (complex matrix elements: injection current)
__host__ __device__ constexpr int wb_kubo_ent(int kind) { return kind >= 2 ? WB_SHC_ENT : WB_KUBO_ENT; }

// spin Hall conductivity: which rotated matrices (in units of nw x nw matrices of the rotated record) feed the
// spin-velocity matrix of formula/covariant.py:689-756
struct WbShcChans {
    int type;        // WBGPU_SHC_RYOO / _QIAO / _SIMPLE
    int iV, iA, iS;  // d_a H [3], A [3] (-1 without external terms), SS [3]
    int iX1, iX2, iX3;   // ryoo: SA [9], SHA [9], -;  qiao: SR [9], SH [3], SHR [9]
};

struct WbKuboParams {
    int kind;        // 0 = optical conductivity, 1 = JDOS, 2 = spin Hall conductivity, 3 = shift current, 4 = injection current
    double sc_eta;   // shift current: broadening of the denominators of the generalised derivative
    double kBT;      // 0: the Fermi factor of a group is a step at its owner bin; > 0: Fermi-Dirac (utility.py:172-182), the
                     // owner's value is spread over the bins within +- 30 kBT of its energy (entries then carry the energy)
    int smr_type;    // 0 = Lorentzian, 1 = Gaussian
    int external;    // external terms (Abar) in A_H
    int nEF, nomega;
    double eta;      // smr_fixed_width
    double EFmin, EFmax, wlo, whi;   // JDOS.nonzero (dynamic.py:160-162): Efermi.min/max, omega.min - 5 eta, omega.max + 5 eta
};

// first index i with Ef[i] >= E  (Ef ascending):  f(E) = (E <= Ef[i]) switches on there (utility.py:172-175)
__device__ __forceinline__ int wb_lower_bound(const double* __restrict__ Ef, int n, double E) {
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (Ef[mid] >= E) hi = mid;
        else lo = mid + 1;
    }
    return lo;
}

__host__ inline size_t wb_kubo_entries_smem_bytes(int nw) {
    int cap = nw * (nw - 1);
    return sizeof(double) * (3 * (size_t)nw) + sizeof(int) * ((size_t)nw + 4) + sizeof(short) * (4 * (size_t)nw) +
           sizeof(ushort2) * (size_t)(cap + 1) + 64;
}

template <int NT>
__global__ void __launch_bounds__(NT)
wb_kubo_entries_kernel(const cplx* __restrict__ xbar, int nch, int nw, long nk, long kofs, const double* __restrict__ Eall,
                       WbWindow win, WbKuboParams P, const double* __restrict__ Ef, const double* __restrict__ weight,
                       long nk_block, double* __restrict__ entries, int* __restrict__ count, int cap,
                       const cplx* __restrict__ Jspin) {
    extern __shared__ __align__(16) double smem_k[];
    double* Es = smem_k;
    double* label = Es + nw;
    double* gE = label + nw;
    int* gidx = (int*)(gE + nw);
    int* misc = gidx + nw;            // [0] = ng, [1] = nvalid
    short* g1 = (short*)(misc + 4);
    short* g2 = g1 + nw;
    short* gs = g2 + nw;
    short* ge = gs + nw;
    ushort2* plist = (ushort2*)(ge + nw + ((4 * nw) & 1));
    plist = (ushort2*)(((uintptr_t)plist + 3) & ~(uintptr_t)3);
    const int n2 = nw * nw;
    const int lane = threadIdx.x & 31;
    const int ENT = wb_kubo_ent(P.kind);
    for (long ik = blockIdx.x; ik < nk; ik += gridDim.x) {
        __syncthreads();
        for (int x = threadIdx.x; x < nw; x += NT) Es[x] = Eall[ik * nw + x];
        __syncthreads();
        if (threadIdx.x == 0) {
            wb_band_groups(Es, nw, win, g1, g2, label);
            int ng = 0;
            for (int n = 0; n < nw; n++)
                if (g1[n] == n) { gs[ng] = (short)n; ge[ng] = g2[n]; gE[ng] = label[n]; ng++; }
            misc[0] = ng;
        }
        __syncthreads();
        const int ng = misc[0];
        for (int g = threadIdx.x; g < ng; g += NT) gidx[g] = wb_lower_bound(Ef, P.nEF, gE[g]);
        __syncthreads();
        // ---- oriented visits (owner group t, partner u) with a non-empty Fermi interval, compacted in (t, u) order
        //      by one warp; owners whose bin lies beyond the last Fermi level never contribute
        if (threadIdx.x < 32) {
            int slot = 0;
            for (int t = 0; t < ng; t++) {
                const int st = gidx[t];
                if (P.kBT > 0.) {
                    if (gE[t] - 30. * P.kBT > P.EFmax) break;   // f = 0 on the whole Fermi axis, and for every later group
                } else if (st >= P.nEF) break;   // s is monotone in the group index
                for (int u0 = 0; u0 < ng; u0 += 32) {
                    const int u = u0 + lane;
                    bool valid = false;
                    if (u < ng && u != t) {
                        valid = (u > t) ? (st < gidx[u]) : (gidx[u] < st);
                        if (P.kBT > 0.) {   // DynamicCalculator.nonzero (dynamic.py:65-69)
                            const double e1max = P.EFmin - 30. * P.kBT, e0min = P.EFmax + 30. * P.kBT;
                            valid = !((gE[t] < e1max && gE[u] < e1max) || (gE[t] > e0min && gE[u] > e0min));
                        }
                        if (valid && P.kind == 1) {
                            const double lo = gE[min(t, u)], hi = gE[max(t, u)];
                            const double d = hi - lo, dm = lo - hi;
                            const bool nz1 = (hi < P.EFmax) && (lo > P.EFmin) && (P.wlo < d) && (d < P.whi);
                            const bool nz2 = (lo < P.EFmax) && (hi > P.EFmin) && (P.wlo < dm) && (dm < P.whi);
                            valid = nz1 || nz2;
                        }
                    }
                    const unsigned mask = __ballot_sync(0xffffffffu, valid);
                    if (valid) {
                        const int pos = slot + __popc(mask & ((1u << lane) - 1u));
                        if (pos < cap) plist[pos] = make_ushort2((unsigned short)t, (unsigned short)u);
                    }
                    slot += __popc(mask);
                }
            }
            if (lane == 0) misc[1] = min(slot, cap);
        }
        __syncthreads();
        const int nvalid = misc[1];
        if (threadIdx.x == 0) count[ik] = nvalid;
        const double wgt = weight[(kofs + ik) / nk_block];
        double* ent = entries + (size_t)ik * cap * ENT;
        const cplx* Vb = xbar + (size_t)ik * nch * n2;
        const cplx* Ab = Vb + 3 * n2;
        const cplx* Jb = Jspin + (size_t)ik * 9 * n2;
        const int per = (P.kind == 0) ? 9 : (P.kind >= 2) ? 27 : 1;
        for (int x = threadIdx.x; x < nvalid * per; x += NT) {
            const int slot = x / per, ab = x - slot * per;
            const ushort2 pr = plist[slot];
            const int i = min(pr.x, pr.y), j = max(pr.x, pr.y);   // lo, hi group
            const double sw = (pr.y > pr.x) ? wgt : -wgt;         // +X at the bin of the lower group, -X at the upper
            double* e = ent + (size_t)slot * ENT;
            if (ab == 0) {
                e[0] = gE[j] - gE[i];
                e[1] = (P.kBT > 0.) ? gE[pr.x] : (double)gidx[pr.x];
            }
            if (P.kind == 1) {
                const double lo = gE[i], hi = gE[j];
                const double d = hi - lo, dm = lo - hi;
                const bool nz1 = (hi < P.EFmax) && (lo > P.EFmin) && (P.wlo < d) && (d < P.whi);
                const bool nz2 = (lo < P.EFmax) && (hi > P.EFmin) && (P.wlo < dm) && (dm < P.whi);
                e[2] = sw * (double)((ge[i] - gs[i]) * (ge[j] - gs[j]));
                e[3] = (double)((nz1 ? 1 : 0) | (nz2 ? 2 : 0));
            } else if (P.kind == 2) {
                // Formula_SHC (dynamic.py:204-221): M_(1,2)[a, b, s] = sum_{m in 1, n in 2} Im( J_mn^{a s} B_nm^b ),
                // J = spin-velocity matrix, B = -i A_H
                const int a = ab / 9, b = (ab / 3) % 3, sp = ab % 3;
                const cplx* J = Jb + (size_t)(3 * a + sp) * n2;
                double mlh = 0., mhl = 0.;
                for (int m = gs[i]; m < ge[i]; m++)
                    for (int n = gs[j]; n < ge[j]; n++) {
                        const double inv = wb_deinv(Es[m], Es[n]);
                        const cplx Vmn = Vb[b * n2 + m * nw + n], Vnm = Vb[b * n2 + n * nw + m];
                        cplx Amn = cmake(inv * Vmn.y, -inv * Vmn.x);     // i D_mn,  D_mn = -V_mn / (E_m - E_n)
                        cplx Anm = cmake(-inv * Vnm.y, inv * Vnm.x);
                        if (P.external) {
                            Amn = cadd(Amn, Ab[b * n2 + m * nw + n]);
                            Anm = cadd(Anm, Ab[b * n2 + n * nw + m]);
                        }
                        const cplx Jmn = J[m * nw + n], Jnm = J[n * nw + m];
                        // Im(J * (-i A)) = -Re(J A)
                        mlh -= Jmn.x * Anm.x - Jmn.y * Anm.y;
                        mhl -= Jnm.x * Amn.x - Jnm.y * Amn.y;
                    }
                e[2 + ab] = sw * mlh;
                e[29 + ab] = sw * mhl;
            } else if (P.kind == 3) {
                // ShiftCurrentFormula (dynamic.py:298-303): Imn[n, m, a, b, c] = -Im( Agen_nm^{c a} A_mn^b ) + (b <-> c); the
                // frequency factor is even in the pair, the Fermi factor odd: entry = M_(hi,lo) - M_(lo,hi)
                const int a = ab / 9, b = (ab / 3) % 3, cc = ab % 3;
                double acc = 0.;
                for (int m = gs[i]; m < ge[i]; m++)
                    for (int n = gs[j]; n < ge[j]; n++) {
                        const double inv = wb_deinv(Es[m], Es[n]);
#pragma unroll
                        for (int t = 0; t < 2; t++) {
                            const int bb = t ? cc : b, c2 = t ? b : cc;   // term (b, c) and its (b <-> c) partner
                            const cplx Vmn = Vb[bb * n2 + m * nw + n], Vnm = Vb[bb * n2 + n * nw + m];
                            cplx Amn = cmake(inv * Vmn.y, -inv * Vmn.x), Anm = cmake(-inv * Vnm.y, inv * Vnm.x);
                            if (P.external) {
                                Amn = cadd(Amn, Ab[bb * n2 + m * nw + n]);
                                Anm = cadd(Anm, Ab[bb * n2 + n * nw + m]);
                            }
                            const cplx Gmn = Jb[(size_t)(3 * c2 + a) * n2 + m * nw + n], Gnm = Jb[(size_t)(3 * c2 + a) * n2 + n * nw + m];
                            // (lo, hi): n' = m, m' = n: -Im(G_mn A_nm);  (hi, lo): -Im(G_nm A_mn)
                            acc += (Gmn.x * Anm.y + Gmn.y * Anm.x) - (Gnm.x * Amn.y + Gnm.y * Amn.x);
                        }
                    }
                e[2 + ab] = sw * acc;
                e[29 + ab] = 0.;
            } else if (P.kind == 4) {
                // InjectionCurrentFormula (dynamic.py:336-343): Imn[m, n, a, b, c] = (v_m - v_n)^a A_mn^b A_nm^c, pair (lo, hi);
                // the reversed pair is -Imn[m, n, a, c, b]
                const int a = ab / 9, b = (ab / 3) % 3, cc = ab % 3;
                cplx acc = cmake(0., 0.);
                for (int m = gs[i]; m < ge[i]; m++)
                    for (int n = gs[j]; n < ge[j]; n++) {
                        const double inv = wb_deinv(Es[m], Es[n]);
                        const cplx Vmn = Vb[b * n2 + m * nw + n], Vnm = Vb[cc * n2 + n * nw + m];
                        cplx Amn = cmake(inv * Vmn.y, -inv * Vmn.x), Anm = cmake(-inv * Vnm.y, inv * Vnm.x);
                        if (P.external) {
                            Amn = cadd(Amn, Ab[b * n2 + m * nw + n]);
                            Anm = cadd(Anm, Ab[cc * n2 + n * nw + m]);
                        }
                        const double dv = Vb[a * n2 + m * nw + m].x - Vb[a * n2 + n * nw + n].x;
                        const cplx z = cmul(Amn, Anm);
                        acc.x += dv * z.x;
                        acc.y += dv * z.y;
                    }
                e[2 + 2 * ab] = sw * acc.x;
                e[3 + 2 * ab] = sw * acc.y;
            } else {
                const int a = ab / 3, b = ab - 3 * a;
                // M[ab] and M[ba] (slot ab stores their sum for a < b, their difference for a > b, M[aa] on the diagonal)
                cplx acc = cmake(0., 0.), act = cmake(0., 0.);
                for (int m = gs[i]; m < ge[i]; m++)
                    for (int n = gs[j]; n < ge[j]; n++) {
                        // A_mn,a = Abar_mn,a + i D_mn,a,  D_mn,a = -Vbar_mn,a / (E_m - E_n)
                        const double inv = wb_deinv(Es[m], Es[n]);
                        cplx Amn[2], Anm[2];
#pragma unroll
                        for (int x = 0; x < 2; x++) {
                            const int d = x ? b : a;
                            const cplx Vmn = Vb[d * n2 + m * nw + n], Vnm = Vb[d * n2 + n * nw + m];
                            Amn[x] = cmake(inv * Vmn.y, -inv * Vmn.x);     // i * (-inv V)
                            Anm[x] = cmake(-inv * Vnm.y, inv * Vnm.x);     // i * (+inv V): 1/(E_n - E_m) = -inv
                            if (P.external) {
                                Amn[x] = cadd(Amn[x], Ab[d * n2 + m * nw + n]);
                                Anm[x] = cadd(Anm[x], Ab[d * n2 + n * nw + m]);
                            }
                        }
                        cfma(acc, Amn[0], Anm[1]);   // A_mn,a A_nm,b -> M[ab]
                        cfma(act, Amn[1], Anm[0]);   // A_mn,b A_nm,a -> M[ba]
                    }
                cplx v = acc;
                if (a < b) v = cadd(acc, act);
                else if (a > b) v = csub(act, acc);   // slot (a > b) holds M[ba'] - M[ab'] with (a', b') = (b, a): M[a'b'] - M[b'a']
                e[2 + 2 * ab] = -sw * v.y;   // i * v
                e[3 + 2 * ab] = sw * v.x;
            }
        }
    }
}

// Spin-velocity matrix  J[k][3 a + s][m][n] = ( <m| v^a S^s |n> + h.c. ) / 2  in the Hamiltonian gauge
// (SpinVelocity, formula/covariant.py:689-756), from the rotated matrices of one k-point:
//   simple  J_mn = sum_l S_ml^s ( V_ln^a + i A_ln^a (E_l - E_n) )                                  (:700-707)
//   ryoo    J_mn = -i ( E_n SA_mn^{as} - SHA_mn^{as} ) + sum_l S_ml^s V_ln^a                        (:741-756)
//   qiao    J_mn = Re(V_nn^a) S_mn^s + E_n ( -i SR_mn^{as} + sum_l S_ml^s D_ln^a )
//                  - ( -i SHR_mn^{as} + sum_l SH_ml^s D_ln^a ),   D_ln = -V_ln / (E_l - E_n)        (:709-739)
// CTA per k-point; thread = (component, m <= n): the raw elements (m, n) and (n, m), then the hermitian part.
template <int NT>
__global__ void __launch_bounds__(NT)
wb_shc_spinvel_kernel(const cplx* __restrict__ xbar, int nch, int nw, long nk, const double* __restrict__ Eall, WbShcChans C,
                      int external, cplx* __restrict__ Jout) {
    extern __shared__ __align__(16) double smem_sv[];
    double* Es = smem_sv;
    const int n2 = nw * nw, ntri = nw * (nw + 1) / 2;
    for (long ik = blockIdx.x; ik < nk; ik += gridDim.x) {
        __syncthreads();
        for (int x = threadIdx.x; x < nw; x += NT) Es[x] = Eall[ik * nw + x];
        __syncthreads();
        const cplx* X = xbar + (size_t)ik * nch * n2;
        cplx* Jk = Jout + (size_t)ik * 9 * n2;
        for (int x = threadIdx.x; x < 9 * ntri; x += NT) {
            const int comp = x / ntri;
            int t = x - comp * ntri, m = 0;
            while (t >= nw - m) { t -= nw - m; m++; }
            const int n = m + t;
            const int a = comp / 3, sp = comp - 3 * a;
            const cplx* V = X + (size_t)(C.iV + a) * n2;
            const cplx* S = X + (size_t)(C.iS + sp) * n2;
            cplx raw[2];
#pragma unroll
            for (int side = 0; side < 2; side++) {
                const int r = side ? n : m, q = side ? m : n;   // element (r, q)
                const double Eq = Es[q];
                cplx acc = cmake(0., 0.);
                if (C.type == WBGPU_SHC_QIAO) {
                    const cplx* SH = X + (size_t)(C.iX2 + sp) * n2;
                    cplx k1 = cmake(0., 0.), l1 = cmake(0., 0.);
                    for (int l = 0; l < nw; l++) {
                        const double f = -wb_deinv(Es[l], Eq);
                        const cplx v = V[l * nw + q];
                        const cplx D = cmake(f * v.x, f * v.y);
                        cfma(k1, S[r * nw + l], D);
                        cfma(l1, SH[r * nw + l], D);
                    }
                    const cplx sr = X[(size_t)(C.iX1 + comp) * n2 + r * nw + q];
                    const cplx shr = X[(size_t)(C.iX3 + comp) * n2 + r * nw + q];
                    // -i z = (z.y, -z.x)
                    const cplx K = cmake(sr.y + k1.x, -sr.x + k1.y), Lq = cmake(shr.y + l1.x, -shr.x + l1.y);
                    const double dE = V[q * nw + q].x;
                    const cplx s = S[r * nw + q];
                    acc = cmake(dE * s.x + Eq * K.x - Lq.x, dE * s.y + Eq * K.y - Lq.y);
                } else {
                    const bool ext = C.type == WBGPU_SHC_SIMPLE && external;
                    const cplx* A = X + (size_t)(C.iA + a) * n2;
                    for (int l = 0; l < nw; l++) {
                        cplx v = V[l * nw + q];
                        if (ext) {   // + i A_lq (E_l - E_q)
                            const cplx al = A[l * nw + q];
                            const double d = Es[l] - Eq;
                            v = cmake(v.x - d * al.y, v.y + d * al.x);
                        }
                        cfma(acc, S[r * nw + l], v);
                    }
                    if (C.type == WBGPU_SHC_RYOO) {
                        const cplx sa = X[(size_t)(C.iX1 + comp) * n2 + r * nw + q];
                        const cplx sha = X[(size_t)(C.iX2 + comp) * n2 + r * nw + q];
                        const cplx z = cmake(Eq * sa.x - sha.x, Eq * sa.y - sha.y);
                        acc = cmake(acc.x + z.y, acc.y - z.x);
                    }
                }
                raw[side] = acc;
            }
            const cplx h = cmake(0.5 * (raw[0].x + raw[1].x), 0.5 * (raw[0].y - raw[1].y));
            Jk[(size_t)comp * n2 + m * nw + n] = h;
            Jk[(size_t)comp * n2 + n * nw + m] = cmake(h.x, -h.y);
        }
    }
}

// Generalised derivative of the Berry connection of the shift current (ShiftCurrentFormula, dynamic.py:247-296):
//   Agen[k][3 c + a][n][m] = i ( W_nm^{ca} + sum_HD + DV_bit ) / (E_m - E_n)  [ + A,a_nm^c + AD_bit - i AA_bit + sum_AD ]
// with P_lm^a = -V_lm^a (E_l - E_m) / ((E_l - E_m)^2 + sc_eta^2),
//   sum_XD = sum_l ( X_nl^c P_lm^a - P_nl^a X_lm^c ) - X_nn^c P_nm^a + P_nm^a X_mm^c        (X = V, A)
//   XD_bit = D_nm^c (X_nn^a - X_mm^a) + D_nm^a (X_nn^c - X_mm^c)                            (X = V, A)
//   AA_bit = (A_nn^a - A_mm^a) A_nm^c
template <int NT>
__global__ void __launch_bounds__(NT)
wb_shift_agen_kernel(const cplx* __restrict__ xbar, int nch, int nw, long nk, const double* __restrict__ Eall, int iW, int iA,
                     int idA, double sc_eta, cplx* __restrict__ Gout, int stage) {
    extern __shared__ __align__(16) double smem_sv[];
    double* Es = smem_sv;
    const int n2 = nw * nw;
    const double eta2 = sc_eta * sc_eta;
    for (long ik = blockIdx.x; ik < nk; ik += gridDim.x) {
        __syncthreads();
        for (int x = threadIdx.x; x < nw; x += NT) Es[x] = Eall[ik * nw + x];
        __syncthreads();
        const cplx* X = xbar + (size_t)ik * nch * n2;
        const cplx* V = X;   // d_a H is the first triple of the record
        cplx* Gk = Gout + (size_t)ik * 9 * n2;
        // P_lq^a once per k-point in shared memory when the launch reserved 3 nw^2 complex behind Es (`stage`): the inner
        // loop below evaluated it -- a division each -- twice per partner band
        cplx* const Ps = (cplx*)(Es + ((nw + 1) & ~1));
        if (stage) {
            for (int x = threadIdx.x; x < 3 * n2; x += NT) {
                const int lq = x % n2;
                const double d = Es[lq / nw] - Es[lq % nw];
                Ps[x] = cscale(-d / (d * d + eta2), V[x]);
            }
            __syncthreads();
        }
        auto Pv = [&](int a, int l, int q) {
            if (stage) return Ps[a * n2 + l * nw + q];
            const double d = Es[l] - Es[q];
            return cscale(-d / (d * d + eta2), V[(size_t)a * n2 + l * nw + q]);
        };
        for (int x = threadIdx.x; x < 9 * n2; x += NT) {
            const int ca = x / n2, nm = x - ca * n2, n = nm / nw, m = nm - n * nw;
            const int c = ca / 3, a = ca - 3 * c;
            const cplx Pnm = Pv(a, n, m);
            const double inm = wb_deinv(Es[n], Es[m]);
            const cplx Dc = cscale(-inm, V[(size_t)c * n2 + nm]), Da = cscale(-inm, V[(size_t)a * n2 + nm]);
            cplx sumV = cmake(0., 0.), sumA = cmake(0., 0.);
            for (int l = 0; l < nw; l++) {
                const cplx Plm = Pv(a, l, m), Pnl = Pv(a, n, l);
                cfma(sumV, V[(size_t)c * n2 + n * nw + l], Plm);
                const cplx z = cmul(Pnl, V[(size_t)c * n2 + l * nw + m]);
                sumV = cmake(sumV.x - z.x, sumV.y - z.y);
                if (iA >= 0) {
                    const cplx* Ac = X + (size_t)(iA + c) * n2;
                    cfma(sumA, Ac[n * nw + l], Plm);
                    const cplx z2 = cmul(Pnl, Ac[l * nw + m]);
                    sumA = cmake(sumA.x - z2.x, sumA.y - z2.y);
                }
            }
            const cplx Vnn_c = V[(size_t)c * n2 + n * nw + n], Vmm_c = V[(size_t)c * n2 + m * nw + m];
            const cplx Vnn_a = V[(size_t)a * n2 + n * nw + n], Vmm_a = V[(size_t)a * n2 + m * nw + m];
            cplx t = X[(size_t)(iW + wb_sym6(c, a)) * n2 + nm];
            t = cadd(t, sumV);
            t = cadd(t, cmul(Pnm, csub(Vmm_c, Vnn_c)));
            t = cadd(t, cmul(Dc, csub(Vnn_a, Vmm_a)));
            t = cadd(t, cmul(Da, csub(Vnn_c, Vmm_c)));
            const double imn = wb_deinv(Es[m], Es[n]);
            cplx g = cmake(-imn * t.y, imn * t.x);   // i t / (E_m - E_n)
            if (iA >= 0) {
                const cplx* Ac = X + (size_t)(iA + c) * n2;
                const cplx* Aa = X + (size_t)(iA + a) * n2;
                const cplx dAc = csub(Ac[n * nw + n], Ac[m * nw + m]), dAa = csub(Aa[n * nw + n], Aa[m * nw + m]);
                g = cadd(g, X[(size_t)(idA + 3 * c + a) * n2 + nm]);
                g = cadd(g, cadd(cmul(Dc, dAa), cmul(Da, dAc)));            // AD_bit
                const cplx aa = cmul(dAa, Ac[nm]);                            // AA_bit
                g = cmake(g.x + aa.y, g.y - aa.x);                            // - i AA_bit
                g = cadd(g, cadd(sumA, cmul(Pnm, cmake(-dAc.x, -dAc.y))));   // sum_AD: - A_nn^c P_nm^a + P_nm^a A_mm^c
            }
            Gk[x] = g;
        }
    }
}

// factor_omega of OpticalConductivity (dynamic.py:191-196) without the (E2 - E1) prefactor: 1/(d - i eta), the
// imaginary part replaced by pi * Gaussian(d) for smr_type != Lorentzian
__device__ __forceinline__ cplx wb_kubo_cfac(double d, double eta, int smr_type) {
    const double den = 1. / (d * d + eta * eta);
    if (smr_type == 0) return cmake(d * den, eta * den);
    double g = 0.;
    if (fabs(d) < eta * sqrt(200.0)) g = 1.0 / (sqrt(CUDART_PI) * eta) * exp(-(d / eta) * (d / eta));
    return cmake(d * den, CUDART_PI * g);
}
__device__ __forceinline__ double wb_kubo_smear(double x, double eta, int smr_type) {
    if (smr_type == 0) return 1.0 / (CUDART_PI * eta) * eta * eta / (x * x + eta * eta);
    if (fabs(x) < eta * sqrt(200.0)) return 1.0 / (sqrt(CUDART_PI) * eta) * exp(-(x / eta) * (x / eta));
    return 0.;
}

// threads per omega value of the accumulation kernel: optical conductivity (component, re | im) = 18, JDOS 1,
// spin Hall 27 components (each thread carries the real and the imaginary part)
__host__ __device__ constexpr int wb_kubo_tpw(int kind) { return kind == 0 ? 18 : kind >= 2 ? 27 : 1; }
__host__ __device__ constexpr int wb_kubo_nc(int kind) { return kind == 0 ? 18 : kind >= 2 ? 54 : 1; }

template <int KIND>
__global__ void __launch_bounds__(wb_kubo_tpw(KIND) * WB_KUBO_WT)
wb_kubo_accumulate_kernel(const double* __restrict__ entries, const int* __restrict__ count, int cap, long nk,
                          WbKuboParams P, const double* __restrict__ omega, const double* __restrict__ Ef,
                          double* __restrict__ Dglob) {
    constexpr int TPW = wb_kubo_tpw(KIND), NC = wb_kubo_nc(KIND), WT = WB_KUBO_WT, NT = TPW * WT, ENT = wb_kubo_ent(KIND);
    __shared__ __align__(16) double ent[WB_KUBO_CHUNK * ENT];
    __shared__ __align__(16) double Wb[WB_KUBO_CHUNK * WT * 4];
    const int w0 = blockIdx.x * WT, nwt = min(WT, P.nomega - w0);
    const int tid = threadIdx.x;
    const int iw = tid / TPW, c = tid - iw * TPW;
    const bool owner = iw < nwt;
    const int ab = c >> 1, ri = c & 1;
    const int wsel = ((ab / 3) > (ab % 3)) ? 2 : 0;   // antisymmetric slot (a > b): Wn, else Wd
    double* const col = Dglob + ((size_t)(w0 + iw) * P.nEF) * NC + (KIND >= 2 ? 2 * c : c);
    const int ct = (c / 9) * 9 + (c % 3) * 3 + (c / 3) % 3;   // KIND 3: component (a, c, b) of (a, b, c)
    for (long ik = blockIdx.y; ik < nk; ik += gridDim.y) {
        const int cnt = count[ik];
        const double* src = entries + (size_t)ik * cap * ENT;
        double Y = 0., Y2 = 0.;
        double curown = -CUDART_INF;   // owner of the values being accumulated: its bin (kBT = 0) or its energy
        bool have = false;
        // add the owner's value (Y, Y2) times its Fermi factor to the difference array
        auto flush = [&]() {
            if (!have || (Y == 0. && Y2 == 0.)) return;
            if (P.kBT == 0.) {
                const size_t o = (size_t)(int)curown * NC;
                if (Y != 0.) atomicAdd(col + o, Y);
                if (KIND >= 2 && Y2 != 0.) atomicAdd(col + o + 1, Y2);
                return;
            }
            const double E = curown, top = E + 30. * P.kBT;
            double prev = 0.;
            for (int i = wb_lower_bound(Ef, P.nEF, E - 30. * P.kBT); i < P.nEF; i++) {
                const double mu = Ef[i];
                const double f = (mu > top) ? 1. : 1. / (exp((E - mu) / P.kBT) + 1.);
                const double d = f - prev;
                prev = f;
                atomicAdd(col + (size_t)i * NC, Y * d);
                if (KIND >= 2) atomicAdd(col + (size_t)i * NC + 1, Y2 * d);
                if (mu > top) break;
            }
        };
        for (int p0 = 0; p0 < cnt; p0 += WB_KUBO_CHUNK) {
            const int np = min(WB_KUBO_CHUNK, cnt - p0);
            __syncthreads();
            for (int x = tid; x < np * ENT; x += NT) ent[x] = src[(size_t)p0 * ENT + x];
            __syncthreads();
            // ---- phase A: frequency factors of (entry, omega), shared by the components
            for (int x = tid; x < np * nwt; x += NT) {
                const int p = x / nwt, w = x - p * nwt;
                const double dl = ent[p * ENT], om = omega[w0 + w];
                double* o = Wb + (p * WT + w) * 4;
                if (KIND == 0) {
                    const cplx c1 = wb_kubo_cfac(dl - om, P.eta, P.smr_type);
                    const cplx c2 = wb_kubo_cfac(-dl - om, P.eta, P.smr_type);
                    // W1 = Delta cfac(Delta - omega): pair (lo, hi), Fermi factor -1;  W2 = -Delta cfac(-Delta - omega):
                    // pair (hi, lo), Fermi factor +1.  Stored: Wd = W2 - W1 (symmetric slots), Wn = -(W1 + W2) (antisymmetric)
                    const double w1x = dl * c1.x, w1y = dl * c1.y, w2x = -dl * c2.x, w2y = -dl * c2.y;
                    o[0] = w2x - w1x; o[1] = w2y - w1y;
                    o[2] = -(w1x + w2x); o[3] = -(w1y + w2y);
                } else if (KIND == 2 && P.kind == 3) {
                    // ShiftCurrent.factor_omega (dynamic.py:319-322), the same for both orders of the pair
                    o[0] = wb_kubo_smear(-dl - om, P.eta, P.smr_type) + wb_kubo_smear(dl - om, P.eta, P.smr_type);
                    o[1] = o[2] = o[3] = 0.;
                } else if (KIND == 2) {
                    // SHC.factor_omega (dynamic.py:232-237): cfac(E1 - E2 - omega) / 2; pair (lo, hi) with the Fermi
                    // factor -1, pair (hi, lo) with +1
                    const cplx c1 = wb_kubo_cfac(-dl - om, P.eta, P.smr_type);
                    const cplx c2 = wb_kubo_cfac(dl - om, P.eta, P.smr_type);
                    o[0] = -0.5 * c1.x; o[1] = -0.5 * c1.y;
                    o[2] = 0.5 * c2.x; o[3] = 0.5 * c2.y;
                } else if (KIND == 3) {
                    // InjectionCurrent.factor_omega (dynamic.py:363-365): smear(E1 - E2 - omega); pair (lo, hi) with the Fermi
                    // factor -1 on M[abc], pair (hi, lo) with +1 on -M[acb]
                    o[0] = -wb_kubo_smear(-dl - om, P.eta, P.smr_type);
                    o[1] = -wb_kubo_smear(dl - om, P.eta, P.smr_type);
                } else {
                    const int fl = (int)ent[p * ENT + 3];
                    o[0] = (fl & 1) ? wb_kubo_smear(dl - om, P.eta, P.smr_type) : 0.;    // (hi, lo): E1 - E2 = +Delta
                    o[1] = (fl & 2) ? wb_kubo_smear(-dl - om, P.eta, P.smr_type) : 0.;   // (lo, hi)
                }
            }
            __syncthreads();
            // ---- phase B: thread = (omega, component): register accumulation per owner bin
            if (owner) {
                for (int p = 0; p < np; p++) {
                    const double* e = ent + p * ENT;
                    const double own = e[1];
                    if (!have || own != curown) {   // uniform
                        flush();
                        curown = own;
                        have = true;
                        Y = 0.;
                        Y2 = 0.;
                    }
                    const double* W = Wb + (p * WT + iw) * 4;
                    if (KIND == 0) {
                        const double Mr = e[2 + 2 * ab], Mi = e[3 + 2 * ab], Wr = W[wsel], Wi = W[wsel + 1];
                        if (ri == 0) Y += Wr * Mr - Wi * Mi;
                        else Y += Wr * Mi + Wi * Mr;
                    } else if (KIND == 2) {
                        const double m1 = e[2 + c], m2 = e[29 + c];
                        Y += W[0] * m1 + W[2] * m2;
                        Y2 += W[1] * m1 + W[3] * m2;
                    } else if (KIND == 3) {
                        Y += W[0] * e[2 + 2 * c] + W[1] * e[2 + 2 * ct];
                        Y2 += W[0] * e[3 + 2 * c] + W[1] * e[3 + 2 * ct];
                    } else {
                        Y += (W[0] - W[1]) * e[2];
                    }
                }
            }
        }
        if (owner) flush();
    }
}

// Optical conductivity, register-tiled variant of the accumulation: thread = (group of 4 frequencies, complex component
// slot ab), CTA = 64 frequencies x 9 slots = 144 threads.  Per entry a thread loads its slot's M (one 16-byte load) and
// the frequency factor of its 4 frequencies (16-byte loads, broadcast among the 9 threads of the group) and does 4
// complex multiply-adds: 16 DFMA per 5 shared-memory loads instead of 2 per 4 (the per-(omega, re|im) kernel above is
// bound by its shared-memory loads).  Same entries, same difference array, same flush rule (owner bin / Fermi-Dirac).
constexpr int WB_KUBO_TW = 64;       // frequencies per CTA
constexpr int WB_KUBO_TCHUNK = 16;   // entries staged per step (18 = 8 full rounds of factor items was measured: slower, 30.3 -> 32.9 ms)

__global__ void __launch_bounds__(144)
wb_kubo_accumulate_optcond_tiled_kernel(const double* __restrict__ entries, const int* __restrict__ count, int cap, long nk,
                                        WbKuboParams P, const double* __restrict__ omega, const double* __restrict__ Ef,
                                        double* __restrict__ Dglob) {
    constexpr int WT = WB_KUBO_TW, CH = WB_KUBO_TCHUNK, NT = 144, ENT = WB_KUBO_ENT, NC = 18;
    static_assert(WT == 64, "layout of Wb: 16 frequency groups of 4");
    __shared__ __align__(16) double ent[CH * ENT];
    __shared__ __align__(16) double Wb[CH * WT * 4];
    const int w0 = blockIdx.x * WT, nwt = min(WT, P.nomega - w0);
    const int tid = threadIdx.x;
    const int grp = tid / 9, ab = tid - 9 * grp;          // frequencies w0 + 4 grp .. + 3
    const int wsel = ((ab / 3) > (ab % 3)) ? 2 : 0;       // antisymmetric slot (a > b): Wn, else Wd
    const int nmine = max(0, min(4, nwt - 4 * grp));      // frequencies of this thread inside the axis
    for (long ik = blockIdx.y; ik < nk; ik += gridDim.y) {
        const int cnt = count[ik];
        const double* src = entries + (size_t)ik * cap * ENT;
        double Yr[4] = {0., 0., 0., 0.}, Yi[4] = {0., 0., 0., 0.};
        double curown = -CUDART_INF;
        bool have = false;
        auto flush = [&]() {
            if (!have) return;
            for (int q = 0; q < nmine; q++) {
                if (Yr[q] == 0. && Yi[q] == 0.) continue;
                double* col = Dglob + ((size_t)(w0 + 4 * grp + q) * P.nEF) * NC + 2 * ab;
                if (P.kBT == 0.) {
                    const size_t o = (size_t)(int)curown * NC;
                    atomicAdd(col + o, Yr[q]);
                    atomicAdd(col + o + 1, Yi[q]);
                    continue;
                }
                const double E = curown, top = E + 30. * P.kBT;
                double prev = 0.;
                for (int i = wb_lower_bound(Ef, P.nEF, E - 30. * P.kBT); i < P.nEF; i++) {
                    const double mu = Ef[i];
                    const double f = (mu > top) ? 1. : 1. / (exp((E - mu) / P.kBT) + 1.);
                    const double d = f - prev;
                    prev = f;
                    atomicAdd(col + (size_t)i * NC, Yr[q] * d);
                    atomicAdd(col + (size_t)i * NC + 1, Yi[q] * d);
                    if (mu > top) break;
                }
            }
        };
        for (int p0 = 0; p0 < cnt; p0 += CH) {
            const int np = min(CH, cnt - p0);
            __syncthreads();
            for (int x = tid; x < np * ENT; x += NT) ent[x] = src[(size_t)p0 * ENT + x];
            __syncthreads();
            // frequency factors, zero beyond the end of the axis.  Layout: the factor of frequency w = 4 grp + q sits at
            // position 16 q + grp of its entry's row, so that the four frequency groups of a warp read ONE 128-byte line per
            // q (rows in frequency order put them 128 bytes apart: 4-way bank conflicts, 43 % of the wavefronts in ncu)
            for (int x = tid; x < np * WT; x += NT) {
                const int p = x / WT, pos = x - p * WT;
                const int w = ((pos & 15) << 2) + (pos >> 4);
                double* o = Wb + (p * WT + pos) * 4;
                if (w < nwt) {
                    const double dl = ent[p * ENT], om = omega[w0 + w];
                    cplx c1, c2;
                    if (P.smr_type == 0) {   // both Lorentzians from ONE division: 1/x1 = x2 / (x1 x2), 1/x2 = x1 / (x1 x2)
                        const double d1 = dl - om, d2 = -dl - om, e2 = P.eta * P.eta;
                        const double x1 = fma(d1, d1, e2), x2 = fma(d2, d2, e2);
                        const double r = 1. / (x1 * x2);
                        const double den1 = r * x2, den2 = r * x1;
                        c1 = cmake(d1 * den1, P.eta * den1);
                        c2 = cmake(d2 * den2, P.eta * den2);
                    } else {
                        c1 = wb_kubo_cfac(dl - om, P.eta, P.smr_type);
                        c2 = wb_kubo_cfac(-dl - om, P.eta, P.smr_type);
                    }
                    const double w1x = dl * c1.x, w1y = dl * c1.y, w2x = -dl * c2.x, w2y = -dl * c2.y;
                    o[0] = w2x - w1x; o[1] = w2y - w1y;
                    o[2] = -(w1x + w2x); o[3] = -(w1y + w2y);
                } else o[0] = o[1] = o[2] = o[3] = 0.;
            }
            __syncthreads();
            double owns[CH];   // owner bins of the chunk, loaded ahead of the loop (the branch below waited for each load)
#pragma unroll
            for (int p = 0; p < CH; p++) owns[p] = ent[(p < np ? p : 0) * ENT + 1];
#pragma unroll
            for (int p = 0; p < CH; p++) {
                if (p >= np) break;
                const double own = owns[p];
                if (!have || own != curown) {   // uniform
                    flush();
                    curown = own;
                    have = true;
#pragma unroll
                    for (int q = 0; q < 4; q++) Yr[q] = Yi[q] = 0.;
                }
                const double2 M = *reinterpret_cast<const double2*>(ent + p * ENT + 2 + 2 * ab);
                const double* Wp = Wb + (p * WT + grp) * 4 + wsel;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const double2 W = *reinterpret_cast<const double2*>(Wp + 64 * q);
                    Yr[q] += W.x * M.x - W.y * M.y;
                    Yi[q] += W.x * M.y + W.y * M.x;
                }
            }
        }
        flush();
    }
}

// Rank-3 kinds (spin Hall / shift current: KIND 2, injection current: KIND 3), register-tiled like the kernel above:
// thread = (group of 4 frequencies, component c of 27), CTA = 64 frequencies x 27 components = 432 threads; per entry a
// thread loads its component's two matrix elements once and the frequency factors of its 4 frequencies (group-minor
// layout, broadcast among the 27 threads of a group): 10 shared-memory loads per 16 DFMA instead of 4 per 4.
template <int KIND>
__global__ void __launch_bounds__(432)
wb_kubo_accumulate_rank3_tiled_kernel(const double* __restrict__ entries, const int* __restrict__ count, int cap, long nk,
                                      WbKuboParams P, const double* __restrict__ omega, const double* __restrict__ Ef,
                                      double* __restrict__ Dglob) {
    static_assert(KIND == 2 || KIND == 3, "spin Hall / shift current (2), injection current (3)");
    constexpr int WT = WB_KUBO_TW, CH = WB_KUBO_TCHUNK, NT = 432, ENT = WB_SHC_ENT, NC = 54;
    static_assert(WT == 64, "layout of Wb: 16 frequency groups of 4");
    __shared__ __align__(16) double ent[CH * ENT];
    __shared__ __align__(16) double Wb[CH * WT * 4];
    const int w0 = blockIdx.x * WT, nwt = min(WT, P.nomega - w0);
    const int tid = threadIdx.x;
    const int grp = tid / 27, c = tid - 27 * grp;            // frequencies w0 + 4 grp .. + 3
    const int ct = (c / 9) * 9 + (c % 3) * 3 + (c / 3) % 3;   // KIND 3: component (a, c, b) of (a, b, c)
    const int nmine = max(0, min(4, nwt - 4 * grp));
    for (long ik = blockIdx.y; ik < nk; ik += gridDim.y) {
        const int cnt = count[ik];
        const double* src = entries + (size_t)ik * cap * ENT;
        double Yr[4] = {0., 0., 0., 0.}, Yi[4] = {0., 0., 0., 0.};
        double curown = -CUDART_INF;
        bool have = false;
        auto flush = [&]() {
            if (!have) return;
            for (int q = 0; q < nmine; q++) {
                if (Yr[q] == 0. && Yi[q] == 0.) continue;
                double* col = Dglob + ((size_t)(w0 + 4 * grp + q) * P.nEF) * NC + 2 * c;
                if (P.kBT == 0.) {
                    const size_t o = (size_t)(int)curown * NC;
                    if (Yr[q] != 0.) atomicAdd(col + o, Yr[q]);
                    if (Yi[q] != 0.) atomicAdd(col + o + 1, Yi[q]);
                    continue;
                }
                const double E = curown, top = E + 30. * P.kBT;
                double prev = 0.;
                for (int i = wb_lower_bound(Ef, P.nEF, E - 30. * P.kBT); i < P.nEF; i++) {
                    const double mu = Ef[i];
                    const double f = (mu > top) ? 1. : 1. / (exp((E - mu) / P.kBT) + 1.);
                    const double d = f - prev;
                    prev = f;
                    atomicAdd(col + (size_t)i * NC, Yr[q] * d);
                    atomicAdd(col + (size_t)i * NC + 1, Yi[q] * d);
                    if (mu > top) break;
                }
            }
        };
        for (int p0 = 0; p0 < cnt; p0 += CH) {
            const int np = min(CH, cnt - p0);
            __syncthreads();
            for (int x = tid; x < np * ENT; x += NT) ent[x] = src[(size_t)p0 * ENT + x];
            __syncthreads();
            // frequency factors (the formulas of wb_kubo_accumulate_kernel), zero beyond the end of the axis; position
            // 16 q + grp of the row holds frequency 4 grp + q
            for (int x = tid; x < np * WT; x += NT) {
                const int p = x / WT, pos = x - p * WT;
                const int w = ((pos & 15) << 2) + (pos >> 4);
                double* o = Wb + (p * WT + pos) * 4;
                o[0] = o[1] = o[2] = o[3] = 0.;
                if (w < nwt) {
                    const double dl = ent[p * ENT], om = omega[w0 + w];
                    if (KIND == 2 && P.kind == 3) {          // ShiftCurrent.factor_omega (dynamic.py:319-322)
                        o[0] = wb_kubo_smear(-dl - om, P.eta, P.smr_type) + wb_kubo_smear(dl - om, P.eta, P.smr_type);
                    } else if (KIND == 2) {                  // SHC.factor_omega (dynamic.py:232-237)
                        const cplx c1 = wb_kubo_cfac(-dl - om, P.eta, P.smr_type);
                        const cplx c2 = wb_kubo_cfac(dl - om, P.eta, P.smr_type);
                        o[0] = -0.5 * c1.x; o[1] = -0.5 * c1.y;
                        o[2] = 0.5 * c2.x; o[3] = 0.5 * c2.y;
                    } else {                                 // InjectionCurrent.factor_omega (dynamic.py:363-365)
                        o[0] = -wb_kubo_smear(-dl - om, P.eta, P.smr_type);
                        o[1] = -wb_kubo_smear(dl - om, P.eta, P.smr_type);
                    }
                }
            }
            __syncthreads();
            if (grp < 16) {
                for (int p = 0; p < np; p++) {
                    const double* e = ent + p * ENT;
                    const double own = e[1];
                    if (!have || own != curown) {   // uniform
                        flush();
                        curown = own;
                        have = true;
#pragma unroll
                        for (int q = 0; q < 4; q++) Yr[q] = Yi[q] = 0.;
                    }
                    const double* Wp = Wb + (p * WT + grp) * 4;
                    if (KIND == 2) {
                        const double m1 = e[2 + c], m2 = e[29 + c];
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const double2 Wa = *reinterpret_cast<const double2*>(Wp + 64 * q);
                            const double2 Wc = *reinterpret_cast<const double2*>(Wp + 64 * q + 2);
                            Yr[q] += Wa.x * m1 + Wc.x * m2;
                            Yi[q] += Wa.y * m1 + Wc.y * m2;
                        }
                    } else {
                        const double2 M = *reinterpret_cast<const double2*>(e + 2 + 2 * c);
                        const double2 Mt = *reinterpret_cast<const double2*>(e + 2 + 2 * ct);
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const double2 Wa = *reinterpret_cast<const double2*>(Wp + 64 * q);
                            Yr[q] += Wa.x * M.x + Wa.y * Mt.x;
                            Yi[q] += Wa.x * M.y + Wa.y * Mt.y;
                        }
                    }
                }
            }
        }
        if (grp < 16) flush();
    }
}

// D holds differences along Efermi: value[iw][iEf][c] = sum_{f <= iEf} D[iw][f][c].  JDOS (NC = 1) and spin Hall
// (NC = 54 = [a][b][s][re | im]): out = scale * value.
// Optical conductivity (NC = 18): slots (ab, ba), a < b hold P = X[ab] + X[ba] and Q = X[ab] - X[ba]:
// out[iEf][iw][ab] = scale (P + Q) / 2, out[..][ba] = scale (P - Q) / 2; the diagonal slots hold X[aa].
__global__ void wb_kubo_finalize_kernel(const double* __restrict__ D, int nomega, int nEF, int NC, double scale,
                                        double* __restrict__ out, int real_only) {
    const long total = (long)nomega * NC;
    for (long x = (long)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (long)gridDim.x * blockDim.x) {
        const int c = (int)(x % NC), w = (int)(x / NC);
        if (NC != 18) {   // JDOS, rank-3 kinds: plain running sum (shift current: the real parts only)
            if (real_only && (c & 1)) continue;
            double run = 0.;
            for (int f = 0; f < nEF; f++) {
                run += D[((size_t)w * nEF + f) * NC + c];
                if (real_only) out[((size_t)f * nomega + w) * (NC / 2) + (c >> 1)] = scale * run;
                else out[((size_t)f * nomega + w) * NC + c] = scale * run;
            }
            continue;
        }
        const int ab = c >> 1, ri = c & 1, a = ab / 3, b = ab % 3;
        if (a > b) continue;   // handled by the thread of the transposed slot
        const int ct = 2 * (b * 3 + a) + ri;
        double run = 0., runt = 0.;
        for (int f = 0; f < nEF; f++) {
            run += D[((size_t)w * nEF + f) * NC + c];
            if (a == b) {
                out[((size_t)f * nomega + w) * NC + c] = scale * run;
            } else {
                runt += D[((size_t)w * nEF + f) * NC + ct];
                out[((size_t)f * nomega + w) * NC + c] = scale * 0.5 * (run + runt);
                out[((size_t)f * nomega + w) * NC + ct] = scale * 0.5 * (run - runt);
            }
        }
    }
}
