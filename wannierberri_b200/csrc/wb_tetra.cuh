// Tetrahedron method for Fermi-level scans on grid K-blocks (StaticCalculator(tetra=True) with KpointBZparallel).
//
// Reference: TetraWeightsParal / TetraWeights.weights_all_band_groups / weights_tetra (grid/tetrahedron.py:15-128,
// 165-268), Data_K_R.E_K_corners_parallel (data_K/data_K_R.py:120-141), the tetra branch of
// StaticCalculator.__call__ (calculators/static.py:84-91, 121-127).
//
// Every k-point owns the parallelepiped cell k +- dK_cell/2; a band's occupation weight at Fermi level ef is the
// average over 12 tetrahedra (cell centre + half a face) of the textbook tetrahedron occupation (or its der-th
// derivative).  A band group's value is added with weight w(ef) = mean over its bands -- dense along Efermi between
// the lowest and highest corner energy, identically 1 above (der = 0) and 0 below, so:
//   direct[ief]  += coef * w(ef_ief) * value      for the Fermi levels inside [e1, e4) of a tetrahedron,
//   suffix[ihi]  += coef * value                  (der = 0) at the first Fermi level >= e4; the Fermi-sea group goes
//                                                 to suffix[0];  a running sum over ief at the end.
#pragma once
#include "wb_common.cuh"
#include "wb_groups.cuh"

// Ebmin / Ebmax[k][n] = min / max over the centre and the 8 corners (TetraWeights.__init__, tetrahedron.py:180-183)
__global__ void wb_tetra_minmax_kernel(const double* __restrict__ Ec, const double* __restrict__ Ecorner, long nkl_stride,
                                       long n, double* __restrict__ Ebmin, double* __restrict__ Ebmax) {
    long x = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n) return;
    double mn = Ec[x], mx = mn;
    for (int c = 0; c < 8; c++) {
        double e = Ecorner[c * nkl_stride + x];
        mn = fmin(mn, e);
        mx = fmax(mx, e);
    }
    Ebmin[x] = mn;
    Ebmax[x] = mx;
}

// eigenvalues only: sort the QL output of k-points [k0, k0 + nk) into E (ascending); flagged k-points go to the
// Jacobi list.  Thread per k-point, rank by counting.
__global__ void wb_eig_sort_kernel(int nw, long k0, long nk, const double* __restrict__ dvals, const int* __restrict__ nsweep,
                                   double* __restrict__ Eout, int* __restrict__ fail_list, int* __restrict__ nfail) {
    long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nk) return;
    if (nsweep[t] < 0) {
        fail_list[atomicAdd(nfail, 1)] = (int)t;
        return;
    }
    const double* d = dvals + t * nw;
    for (int i = 0; i < nw; i++) {
        const double di = d[i];
        int rank = 0;
        for (int j = 0; j < nw; j++) {
            const double dj = d[j];
            rank += (dj < di) || (dj == di && j < i);
        }
        Eout[(k0 + t) * nw + rank] = di;
    }
}

struct WbTetra {
    double e1, e2, e3, e4;
    double c[3][3];   // der >= 1: coefficients c_{r,1..3} of the three pieces r = 1 (e1..e2), 2 (e2..e3), 3 (e3..e4)
};

// weights_tetra set-up (tetrahedron.py:15-28, 53-78): sorted corner energies with the 1e-12 separation fix
__device__ __forceinline__ void wb_tetra_setup(double a, double b, double c, double d, int der, WbTetra& T) {
#define WB_CSWAP(x, y) { const double lo = fmin(x, y), hi = fmax(x, y); x = lo; y = hi; }
    WB_CSWAP(a, b) WB_CSWAP(c, d) WB_CSWAP(a, c) WB_CSWAP(b, d) WB_CSWAP(b, c)
#undef WB_CSWAP
    const double diff_min = 1e-12;
    if (b - a < diff_min) b = a + diff_min;
    if (c - b < diff_min) c = b + diff_min;
    if (d - c < diff_min) d = c + diff_min;
    T.e1 = a; T.e2 = b; T.e3 = c; T.e4 = d;
    if (der == 0) return;
    // The polynomial coefficients are differences of large, nearly equal terms (energies ~ 20 eV over corner
    // spreads ~ 10 meV): their rounding noise is ~1e-7 relative, so parity with the reference needs ITS operation
    // order, with no FMA contraction -- hence the explicit round-to-nearest intrinsics.
#define M_(x, y) __dmul_rn(x, y)
#define A_(x, y) __dadd_rn(x, y)
#define S_(x, y) __dsub_rn(x, y)
    const double e1 = a, e2 = b, e3 = c, e4 = d;
    const double d41 = S_(e4, e1), d42 = S_(e4, e2), d43 = S_(e4, e3), d31 = S_(e3, e1), d32 = S_(e3, e2), d21 = S_(e2, e1);
    const double denom3 = __ddiv_rn(1., M_(M_(d41, d42), d43));
    const double denom2 = __ddiv_rn(1., M_(M_(M_(d31, d41), d32), d42));
    const double denom1 = __ddiv_rn(1., M_(M_(d21, d31), d41));
    T.c[0][2] = denom1;
    T.c[2][2] = denom3;
    T.c[1][2] = M_(denom2, S_(S_(A_(e1, e2), e3), e4));
    T.c[0][1] = M_(M_(-3., e1), denom1);
    // (((e3 - e2) * (e4 - e2)) - (e1 - e3) * (2 * e2 + e4) - (e3 + e1 + e2) * (e2 - e4)) * denom2
    T.c[1][1] = M_(S_(S_(M_(d32, d42), M_(S_(e1, e3), A_(M_(2., e2), e4))), M_(A_(A_(e3, e1), e2), S_(e2, e4))), denom2);
    T.c[2][1] = M_(M_(-3., e4), denom3);
    T.c[0][0] = M_(M_(3., M_(e1, e1)), denom1);
    // (-2 * e1 * ((e3 - e2) * (e4 - e2)) + (2 * e2 * e4 + e2 ** 2) * (e1 - e3) + (e1 * e2 + e2 * e3 + e1 * e3) * (e2 - e4)) * denom2
    T.c[1][0] = M_(A_(A_(M_(M_(-2., e1), M_(d32, d42)), M_(A_(M_(M_(2., e2), e4), M_(e2, e2)), S_(e1, e3))),
                      M_(A_(A_(M_(e1, e2), M_(e2, e3)), M_(e1, e3)), S_(e2, e4))), denom2);
    T.c[2][0] = M_(M_(3., M_(e4, e4)), denom3);
}

// occupation (der = 0, the "accurate" branch tetrahedron.py:33-50) or its derivative for e1 <= ef < e4
__device__ __forceinline__ double wb_tetra_weight(const WbTetra& T, double ef, int der) {
    const double e1 = T.e1, e2 = T.e2, e3 = T.e3, e4 = T.e4;
    if (der == 0) {
        if (ef >= e3) return 1 - ((ef - e4) / (e1 - e4)) * ((ef - e4) / (e2 - e4)) * ((ef - e4) / (e3 - e4));
        if (ef >= e2) {
            const double a13 = (ef - e1) / (e3 - e1), a14 = (ef - e1) / (e4 - e1);
            const double a23 = (ef - e2) / (e3 - e2), a24 = (ef - e2) / (e4 - e2);
            return a23 * a24 + a13 * (a14 * (1 - a24) + a24 * (1 - a23));
        }
        return ((ef - e1) / (e2 - e1)) * ((ef - e1) / (e3 - e1)) * ((ef - e1) / (e4 - e1));
    }
    const int r = (ef >= e3) ? 2 : (ef >= e2) ? 1 : 0;
    const double c1 = T.c[r][0], c2 = T.c[r][1], c3 = T.c[r][2];
    if (der == 1) return A_(c1, M_(ef, A_(M_(2., c2), M_(M_(3., c3), ef))));
    if (der == 2) return A_(M_(2., c2), M_(M_(6., c3), ef));
    return M_(6., c3);
}
#undef M_
#undef A_
#undef S_

// Fermi level i of the uniform grid, evaluated like numpy.linspace (start + i * step, two roundings)
__device__ __forceinline__ double wb_ef_at(double Ef0, double dEF, int i) { return __dadd_rn(Ef0, __dmul_rn((double)i, dEF)); }

// first i in [0, n] with Ef(i) >= e
__device__ __forceinline__ int wb_ef_lower_bound(double Ef0, double dEF, int n, double e) {
    double q = ceil((e - Ef0) / dEF);
    int i = (q <= 0.) ? 0 : (q >= (double)n) ? n : (int)q;
    while (i > 0 && wb_ef_at(Ef0, dEF, i - 1) >= e) i--;
    while (i < n && wb_ef_at(Ef0, dEF, i) < e) i++;
    return i;
}

__host__ inline size_t wb_tetra_acc_smem_bytes(int nw, int nEF, int ncomp, int use_smem) {
    return sizeof(double) * ((use_smem ? 2 * (size_t)nEF * ncomp : 0) + 2 * (size_t)nw) + sizeof(short) * 2 * nw + 16;
}

// hist = [direct[nEF][ncomp] | suffix[nEF][ncomp]];  one CTA walks k-points blockIdx.x, blockIdx.x + gridDim.x, ...
// `win` is the tetrahedron window of the spec (Ebmin / Ebmax rows are indexed with the k-point index of this launch).
__global__ void __launch_bounds__(128)
wb_tetra_accumulate_kernel(const double* __restrict__ ev_val, int ev_stride, int nw, long nk, long nk_block,
                           const double* __restrict__ Ecen, const double* __restrict__ Ecorner, long corner_stride,
                           WbWindow win, const double* __restrict__ weight, int ncomp, int der, int nEF, double Ef0,
                           double dEF, double* __restrict__ hist, int use_smem) {
    extern __shared__ __align__(16) double smem_tt[];
    const size_t hsz = (size_t)nEF * ncomp;
    double* p = smem_tt;
    double* hd = hist;
    if (use_smem) { hd = p; p += 2 * hsz; }
    double* hs = hd + hsz;
    double* Es = p;
    double* label = Es + nw;
    short* g1 = (short*)(label + nw);
    short* g2 = g1 + nw;
    if (use_smem) {
        for (size_t x = threadIdx.x; x < 2 * hsz; x += blockDim.x) hd[x] = 0.;
    }
    for (long ik = blockIdx.x; ik < nk; ik += gridDim.x) {
        __syncthreads();
        for (int x = threadIdx.x; x < nw; x += blockDim.x) Es[x] = Ecen[ik * nw + x];
        __syncthreads();
        if (threadIdx.x == 0) wb_band_groups_tetra(Es, win.Ebmin + ik * nw, win.Ebmax + ik * nw, nw, win, g1, g2, label);
        __syncthreads();
        const double wk = weight[ik / nk_block];
        // ---- work item = (band, tetrahedron)
        for (int x = threadIdx.x; x < nw * 12; x += blockDim.x) {
            const int ib = x / 12, j = x - 12 * ib;
            const int a = g1[ib];
            if (a < 0 || label[a] == -CUDART_INF) continue;   // not in a kept group / Fermi-sea group (below)
            const int b = g2[ib];
            const double coef = wk / (12. * (double)(b - a));
            // TetraWeightsParal.weight_1k1b_priv (tetrahedron.py:259-268): face `iface` of axis `ax`, half `tri`
            const int iface = j / 6, ax = (j % 6) >> 1, tri = j & 1;
            auto corner = [&](int pp, int qq) {
                const int c = (ax == 0) ? (iface * 4 + pp * 2 + qq) : (ax == 1) ? (pp * 4 + iface * 2 + qq) : (pp * 4 + qq * 2 + iface);
                return Ecorner[(size_t)c * corner_stride + ik * nw + ib];
            };
            WbTetra T;
            const int der_w = (der < 0) ? 0 : der;   // hole_like (der = -1): 1 - occupation (tetrahedron.py:197-198)
            wb_tetra_setup(Es[ib], corner(0, 0), tri == 0 ? corner(0, 1) : corner(1, 0), corner(1, 1), der_w, T);
            const int lo = wb_ef_lower_bound(Ef0, dEF, nEF, T.e1), hi = wb_ef_lower_bound(Ef0, dEF, nEF, T.e4);
            const double* v = ev_val + (size_t)(ik * nw + a) * ev_stride;
            for (int i = lo; i < hi; i++) {
                double w = wb_tetra_weight(T, wb_ef_at(Ef0, dEF, i), der_w);
                if (der < 0) w = 1. - w;
                w *= coef;
                for (int c = 0; c < ncomp; c++) atomicAdd(&hd[(size_t)i * ncomp + c], w * v[c]);
            }
            if (der == 0 && hi < nEF)
                for (int c = 0; c < ncomp; c++) atomicAdd(&hs[(size_t)hi * ncomp + c], coef * v[c]);
            if (der < 0 && lo > 0)   // weight one below e1: +X at level 0, -X at level lo of the running sum
                for (int c = 0; c < ncomp; c++) {
                    atomicAdd(&hs[c], coef * v[c]);
                    if (lo < nEF) atomicAdd(&hs[(size_t)lo * ncomp + c], -(coef * v[c]));
                }
        }
        // ---- Fermi-sea group / hole_like: group above the Fermi axis: weight one at every Fermi level
        for (int x = threadIdx.x; x < nw * ncomp; x += blockDim.x) {
            const int n = x / ncomp, cc = x - n * ncomp;
            if (label[n] == -CUDART_INF) atomicAdd(&hs[cc], wk * ev_val[(size_t)(ik * nw + n) * ev_stride + cc]);
        }
    }
    if (use_smem) {
        __syncthreads();
        for (size_t x = threadIdx.x; x < 2 * hsz; x += blockDim.x) {
            const double v = hd[x];
            if (v != 0.) atomicAdd(&hist[x], v);
        }
    }
}

// out[ief][c] = scale * (direct[ief][c] + sum_{j <= ief} suffix[j][c]);  one thread per component
__global__ void wb_tetra_finalize_kernel(const double* __restrict__ hist, int nEF, int ncomp, double scale,
                                         double* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncomp) return;
    const double* hd = hist;
    const double* hs = hist + (size_t)nEF * ncomp;
    double run = 0.;
    for (int i = 0; i < nEF; i++) {
        run += hs[(size_t)i * ncomp + c];
        out[(size_t)i * ncomp + c] = scale * (hd[(size_t)i * ncomp + c] + run);
    }
}
