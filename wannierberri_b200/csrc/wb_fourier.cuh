// R -> k transform as a pruned, separable 3-D DFT.
//
// Reference: Rvectors.apply_expdK / derivative / R_to_k (fourier/rvectors.py:477-506) and
// FFT_R_to_k.__call__ (fourier/fft.py:133-192).  Same mathematics, different algorithm:
//   X(k) = sum_R X_R exp(2 pi i R.(kappa + dK)),   kappa = (m0/N0, m1/N1, m2/N2)
// factorises over the three lattice directions because the K-block shift dK enters the phase
// exactly like kappa does.  The R-vectors live in a small box (n0 x n1 x n2, 7^3 for bcc Fe), so
// three passes of a dense n_d -> N_d transform along one axis cost ~(n0+...) complex MACs per
// output element instead of nR = 95 for the direct sum, with perfectly coalesced traffic: the
// innermost ("inner") index of every pass is the packed matrix-element index of the record.
//
// The R-space table is built once per plan (wb_build_rtable_kernel): derivative (x i(R+t_j-t_i)),
// curl of A and the hermitisation  X -> (X + X^dagger)/2  are linear and commute with the
// transform, so they are applied in R-space and only the upper triangle of hermitised channels
// is ever transformed.
#pragma once
#include "wb_common.cuh"

// ------------------------------------------------------------------------------------------
// R-space table:  table[(r0*n1 + r1)*n2 + r2][e],  r_d = R_d - rmin_d,  e = record element.
// One thread per (iR, i, j).  Accumulates with atomics (runs once per plan; duplicates in iRvec
// and the R <-> -R folding of the hermitisation both add up, like the scatter-add of fft.py:177).
// ------------------------------------------------------------------------------------------
struct WbRInputs {
    const cplx* Ham;   // [nR][nw][nw]
    const cplx* AA;    // [nR][nw][nw][3]
    const cplx* BB;
    const cplx* CC;
    const cplx* SS;
    const cplx *SA, *SHA, *SR, *SH, *SHR;   // [nR][nw][nw][3][3] ([3] for SH)
    const double* T;   // [nR][nw][nw][3]  cRvec_shifted
    const int* iRvec;  // [nR][3]
};

__device__ __forceinline__ void atomic_cadd(cplx* p, cplx v) {
    atomicAdd(&p->x, v.x);
    atomicAdd(&p->y, v.y);
}

// add X_R[i][j] to a hermitised channel: half to (R; i,j) if i<=j, half conj to (-R; j,i) if j<=i
__device__ __forceinline__ void add_herm(cplx* table, long cellR, long cellmR, int E, int off, int i, int j,
                                         int nw, cplx v) {
    if (i <= j) atomic_cadd(&table[cellR * E + off + tri_index(i, j, nw)], cscale(0.5, v));
    if (j <= i) atomic_cadd(&table[cellmR * E + off + tri_index(j, i, nw)], cscale(0.5, cconj(v)));
}

__global__ void wb_build_rtable_kernel(WbRInputs in, WbLayout L, int nR, int3 rmin, int3 nbox, cplx* table) {
    const int nw = L.nw;
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long total = (long)nR * nw * nw;
    if (idx >= total) return;
    int j = idx % nw;
    int i = (idx / nw) % nw;
    int iR = idx / ((long)nw * nw);
    int R0 = in.iRvec[3 * iR], R1 = in.iRvec[3 * iR + 1], R2 = in.iRvec[3 * iR + 2];
    long cellR = ((long)(R0 - rmin.x) * nbox.y + (R1 - rmin.y)) * nbox.z + (R2 - rmin.z);
    long cellmR = ((long)(-R0 - rmin.x) * nbox.y + (-R1 - rmin.y)) * nbox.z + (-R2 - rmin.z);
    const int E = L.E;
    double T[3];
    for (int a = 0; a < 3; a++) T[a] = in.T[idx * 3 + a];

    if (in.Ham) {
        cplx h = in.Ham[idx];
        add_herm(table, cellR, cellmR, E, L.off_H, i, j, nw, h);
        if (L.off_dH[0] >= 0) {
            for (int a = 0; a < 3; a++) {  // i * T_a * H   (rvectors.py:487-494)
                const cplx d = cmake(-T[a] * h.y, T[a] * h.x);
                if (L.dH_herm) add_herm(table, cellR, cellmR, E, L.off_dH[a], i, j, nw, d);
                else atomic_cadd(&table[cellR * E + L.off_dH[a] + i * nw + j], d);
            }
        }
        if (L.off_W[0] >= 0) {   // second comma-derivative: i T_d (i T_b H) = -T_b T_d H   (rvectors.py:487-494, twice)
            for (int b = 0; b < 3; b++)
                for (int d = b; d < 3; d++) {
                    const cplx w = cmake(-(T[d] * (T[b] * h.x)), -(T[d] * (T[b] * h.y)));
                    const int off = L.off_W[wb_sym6(b, d)];
                    if (L.dH_herm) add_herm(table, cellR, cellmR, E, off, i, j, nw, w);
                    else atomic_cadd(&table[cellR * E + off + i * nw + j], w);
                }
        }
    }
    if (in.Ham && L.off_W3[0] >= 0) {   // third comma-derivative: -i T_b T_c T_d H
        const cplx h = in.Ham[idx];
        int t = 0;
        for (int b = 0; b < 3; b++)
            for (int cc = b; cc < 3; cc++)
                for (int d = cc; d < 3; d++, t++) {
                    const double f = T[b] * T[cc] * T[d];
                    const cplx w = cmake(f * h.y, -f * h.x);
                    if (L.dH_herm) add_herm(table, cellR, cellmR, E, L.off_W3[t], i, j, nw, w);
                    else atomic_cadd(&table[cellR * E + L.off_W3[t] + i * nw + j], w);
                }
    }
    if (in.AA && L.off_A[0] >= 0) {
        cplx A[3];
        for (int a = 0; a < 3; a++) {
            A[a] = in.AA[idx * 3 + a];
            add_herm(table, cellR, cellmR, E, L.off_A[a], i, j, nw, A[a]);
        }
        if (L.off_O[0] >= 0) {
            for (int c = 0; c < 3; c++) {  // i (T_alpha A_beta - T_beta A_alpha)   (data_K_R.py:48-51)
                int al = WB_ALPHA(c), be = WB_BETA(c);
                cplx d = cmake(T[al] * A[be].x - T[be] * A[al].x, T[al] * A[be].y - T[be] * A[al].y);
                add_herm(table, cellR, cellmR, E, L.off_O[c], i, j, nw, cmake(-d.y, d.x));
            }
        }
        if (L.off_dA[0] >= 0) {   // i T_d A_b and i T_d rotA_c (rvectors.py:487-494)
            for (int b = 0; b < 3; b++) {
                const int al = WB_ALPHA(b), be = WB_BETA(b);
                const cplx r0 = cmake(T[al] * A[be].x - T[be] * A[al].x, T[al] * A[be].y - T[be] * A[al].y);
                const cplx rot = cmake(-r0.y, r0.x);
                for (int d = 0; d < 3; d++) {
                    add_herm(table, cellR, cellmR, E, L.off_dA[3 * b + d], i, j, nw, cmake(-T[d] * A[b].y, T[d] * A[b].x));
                    if (L.off_dO[0] >= 0)
                        add_herm(table, cellR, cellmR, E, L.off_dO[3 * b + d], i, j, nw, cmake(-T[d] * rot.y, T[d] * rot.x));
                }
            }
        }
    }
    if (in.BB && L.off_B[0] >= 0)
        for (int a = 0; a < 3; a++) atomic_cadd(&table[cellR * E + L.off_B[a] + i * nw + j], in.BB[idx * 3 + a]);
    if (in.CC && L.off_C[0] >= 0)
        for (int a = 0; a < 3; a++) atomic_cadd(&table[cellR * E + L.off_C[a] + i * nw + j], in.CC[idx * 3 + a]);
    if (in.BB && in.CC && L.off_dB[0] >= 0)   // i T_d B_b, i T_d C_c
        for (int a = 0; a < 3; a++) {
            const cplx bv = in.BB[idx * 3 + a], cv = in.CC[idx * 3 + a];
            for (int d = 0; d < 3; d++) {
                atomic_cadd(&table[cellR * E + L.off_dB[3 * a + d] + i * nw + j], cmake(-T[d] * bv.y, T[d] * bv.x));
                atomic_cadd(&table[cellR * E + L.off_dC[3 * a + d] + i * nw + j], cmake(-T[d] * cv.y, T[d] * cv.x));
            }
        }
    if (in.SS && L.off_S[0] >= 0)
        for (int a = 0; a < 3; a++) add_herm(table, cellR, cellmR, E, L.off_S[a], i, j, nw, in.SS[idx * 3 + a]);
    if (in.SS && L.off_dS[0] >= 0)
        for (int a = 0; a < 3; a++) {
            const cplx sv = in.SS[idx * 3 + a];
            for (int d = 0; d < 3; d++) add_herm(table, cellR, cellmR, E, L.off_dS[3 * a + d], i, j, nw, cmake(-T[d] * sv.y, T[d] * sv.x));
        }
    // second comma-derivatives  i T_e (i T_d X_b) = -T_d T_e X_b  (plug-in formulae; rvectors.py:487-494 twice)
    if (in.AA && L.off_d2A[0] >= 0)
        for (int b = 0; b < 3; b++) {
            const cplx a = in.AA[idx * 3 + b];
            const int al = WB_ALPHA(b), be = WB_BETA(b);
            const cplx Aal = in.AA[idx * 3 + al], Abe = in.AA[idx * 3 + be];
            const cplx r0 = cmake(T[al] * Abe.x - T[be] * Aal.x, T[al] * Abe.y - T[be] * Aal.y);
            const cplx rot = cmake(-r0.y, r0.x);   // (curl A)_b in R-space
            for (int d = 0; d < 3; d++)
                for (int e = d; e < 3; e++) {
                    const double f = -(T[d] * T[e]);
                    add_herm(table, cellR, cellmR, E, L.off_d2A[6 * b + wb_sym6(d, e)], i, j, nw, cscale(f, a));
                    if (L.off_d2O[0] >= 0) add_herm(table, cellR, cellmR, E, L.off_d2O[6 * b + wb_sym6(d, e)], i, j, nw, cscale(f, rot));
                }
        }
    if (in.SS && L.off_d2S[0] >= 0)
        for (int b = 0; b < 3; b++) {
            const cplx sv = in.SS[idx * 3 + b];
            for (int d = 0; d < 3; d++)
                for (int e = d; e < 3; e++)
                    add_herm(table, cellR, cellmR, E, L.off_d2S[6 * b + wb_sym6(d, e)], i, j, nw, cscale(-(T[d] * T[e]), sv));
        }
    if (in.BB && in.CC && L.off_d2B[0] >= 0)
        for (int b = 0; b < 3; b++) {
            const cplx bv = in.BB[idx * 3 + b], cv = in.CC[idx * 3 + b];
            for (int d = 0; d < 3; d++)
                for (int e = d; e < 3; e++) {
                    const double f = -(T[d] * T[e]);
                    atomic_cadd(&table[cellR * E + L.off_d2B[6 * b + wb_sym6(d, e)] + i * nw + j], cscale(f, bv));
                    atomic_cadd(&table[cellR * E + L.off_d2C[6 * b + wb_sym6(d, e)] + i * nw + j], cscale(f, cv));
                }
        }
    // spin-current matrices: not hermitised (data_K_R.py:84-87)
    if (in.SA && L.off_SA[0] >= 0)
        for (int a = 0; a < 9; a++) atomic_cadd(&table[cellR * E + L.off_SA[a] + i * nw + j], in.SA[idx * 9 + a]);
    if (in.SHA && L.off_SHA[0] >= 0)
        for (int a = 0; a < 9; a++) atomic_cadd(&table[cellR * E + L.off_SHA[a] + i * nw + j], in.SHA[idx * 9 + a]);
    if (in.SR && L.off_SR[0] >= 0)
        for (int a = 0; a < 9; a++) atomic_cadd(&table[cellR * E + L.off_SR[a] + i * nw + j], in.SR[idx * 9 + a]);
    if (in.SH && L.off_SH[0] >= 0)
        for (int a = 0; a < 3; a++) atomic_cadd(&table[cellR * E + L.off_SH[a] + i * nw + j], in.SH[idx * 3 + a]);
    if (in.SHR && L.off_SHR[0] >= 0)
        for (int a = 0; a < 9; a++) atomic_cadd(&table[cellR * E + L.off_SHR[a] + i * nw + j], in.SHR[idx * 9 + a]);
}

// Is  d_a H_R = i (R + t_j - t_i)_a H_R  hermitian, i.e. dH(R; i, j) == conj(dH(-R; j, i))?  The reference does not
// hermitise this channel (data_K_R.py:84-87), so it may only be packed as a triangle when the input has the
// symmetry.  cellmap[cell] = index of the R-vector in that cell of the bounding box (-1: none, -2: duplicated).
// out[0] = max |asymmetry|, out[1] = max |dH|  (as bit patterns of non-negative doubles -> atomicMax).
__global__ void wb_dH_herm_check_kernel(const cplx* __restrict__ Ham, const double* __restrict__ T, const int* __restrict__ iRvec,
                                        const int* __restrict__ cellmap, int nR, int nw, int3 rmin, int3 nbox,
                                        unsigned long long* out) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long total = (long)nR * nw * nw;
    if (idx >= total) return;
    int j = idx % nw;
    int i = (idx / nw) % nw;
    int iR = idx / ((long)nw * nw);
    int R0 = iRvec[3 * iR], R1 = iRvec[3 * iR + 1], R2 = iRvec[3 * iR + 2];
    long cellR = ((long)(R0 - rmin.x) * nbox.y + (R1 - rmin.y)) * nbox.z + (R2 - rmin.z);
    long cellmR = ((long)(-R0 - rmin.x) * nbox.y + (-R1 - rmin.y)) * nbox.z + (-R2 - rmin.z);
    const int jR = cellmap[cellmR];
    const bool dup = (cellmap[cellR] == -2) || (jR == -2);
    const cplx h = Ham[idx];
    double asym = 0., mag = 0.;
    for (int a = 0; a < 3; a++) {
        const double t = T[idx * 3 + a];
        const cplx d = cmake(-t * h.y, t * h.x);
        cplx dm = cmake(0., 0.);
        if (jR >= 0) {
            const long idm = ((long)jR * nw + j) * nw + i;
            const cplx hm = Ham[idm];
            const double tm = T[idm * 3 + a];
            dm = cmake(-tm * hm.y, tm * hm.x);
        }
        // d == conj(dm) ?
        asym = fmax(asym, fmax(fabs(d.x - dm.x), fabs(d.y + dm.y)));
        mag = fmax(mag, fmax(fabs(d.x), fabs(d.y)));
    }
    if (dup) asym = CUDART_INF;
    atomicMax(&out[0], (unsigned long long)__double_as_longlong(asym));
    atomicMax(&out[1], (unsigned long long)__double_as_longlong(mag));
}

// ------------------------------------------------------------------------------------------
// Twiddles  W_d[b][k][r] = exp(2 pi i (rmin_d + r) (k / N_d + dK[b][d]))
// The k/N part is reduced exactly in integers (mod N) before going to floating point.
// ------------------------------------------------------------------------------------------
__global__ void wb_twiddle_kernel(const double* __restrict__ dK, int nb, int N, int n, int rmin, int d,
                                  cplx* __restrict__ W) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long total = (long)nb * N * n;
    if (idx >= total) return;
    int r = idx % n;
    int k = (idx / n) % N;
    int b = idx / ((long)n * N);
    int R = rmin + r;
    int rk = (int)((((long)R * k) % N + N) % N);
    double frac = (double)rk / (double)N + (double)R * dK[3 * b + d];
    double s, c;
    sincospi(2.0 * frac, &s, &c);
    W[idx] = cmake(c, s);
}

// ------------------------------------------------------------------------------------------
// One pass of the separable transform along one axis:
//   out[b][o][k][t] = sum_r W[b][k][r] * in[b][o][r][t]      o < outer, r < n, k < N, t < S
// Thread = one inner index t and KC consecutive outputs k; the block shares (b, o, k-chunk), so
// the twiddles are a warp-uniform broadcast from shared memory and every global access is a fully
// coalesced 16-byte-per-lane stream.  4*KC DFMA per 16-byte load.
// ------------------------------------------------------------------------------------------
template <int KC>
__global__ void __launch_bounds__(256)
wb_axis_dft_kernel(const cplx* __restrict__ in, cplx* __restrict__ out, const cplx* __restrict__ W,
                   int n, int N, long S, int outer, long in_bstride, long out_bstride) {
    extern __shared__ cplx w_s[];  // [KC][n]
    const int nchunk = (N + KC - 1) / KC;
    const int o = blockIdx.y / nchunk;
    const int k0 = (blockIdx.y % nchunk) * KC;
    const int b = blockIdx.z;
    for (int x = threadIdx.x; x < KC * n; x += blockDim.x) {
        int kk = x / n, r = x % n;
        w_s[x] = (k0 + kk < N) ? W[((long)b * N + (k0 + kk)) * n + r] : cmake(0., 0.);
    }
    __syncthreads();
    long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= S) return;
    const cplx* src = in + (long)b * in_bstride + (long)o * n * S + t;
    cplx acc[KC];
#pragma unroll
    for (int kk = 0; kk < KC; kk++) acc[kk] = cmake(0., 0.);
    for (int r = 0; r < n; r++) {
        cplx v = __ldg(src + (long)r * S);
#pragma unroll
        for (int kk = 0; kk < KC; kk++) cfma(acc[kk], w_s[kk * n + r], v);
    }
    cplx* dst = out + (long)b * out_bstride + ((long)o * N + k0) * S + t;
#pragma unroll
    for (int kk = 0; kk < KC; kk++)
        if (k0 + kk < N) dst[(long)kk * S] = acc[kk];
}

// ------------------------------------------------------------------------------------------
// Axes 1 and 0 in ONE pass:
//   X[b][k0][k1][t] = sum_{r0} W0[b][k0][r0] * ( sum_{r1} W1[b][k1][r1] * Z[b][r0][r1][t] ),   t < S2 = N2 * E
// The intermediate Y[r0][k1][..] of the separable transform never goes to memory: a CTA stages the Z tile of TB
// inner indices (n0*n1*TB complex, read once, L2 resident) in shared memory and writes all N0*N1 outputs of those
// indices, so the only HBM stream is the k-space record X itself (the algorithmic 16*E bytes per k-point).
// Thread = one inner index; KC accumulators over k0; twiddles are warp-uniform shared-memory broadcasts.
// ------------------------------------------------------------------------------------------
template <int KC, int TB, int KSPLIT>
__global__ void __launch_bounds__(TB * KSPLIT)
wb_axis10_fused_kernel(const cplx* __restrict__ Z, cplx* __restrict__ X, const cplx* __restrict__ W1,
                       const cplx* __restrict__ W0, int n0, int n1, int N0, int N1, long S2, long z_bstride,
                       long x_bstride) {
    extern __shared__ cplx sm_f[];
    const int nchunk = (N0 + KC - 1) / KC;
    cplx* Zs = sm_f;                          // [n0*n1][TB]
    cplx* W1s = Zs + (size_t)n0 * n1 * TB;    // [N1][n1]
    cplx* W0s = W1s + N1 * n1;                // [nchunk*KC][n0], zero beyond N0
    // thread = (inner index el, share h of the k1 range): KSPLIT threads per inner index keep more warps in flight
    const int b = blockIdx.z;
    const int el = threadIdx.x % TB, h = threadIdx.x / TB;
    const long t = (long)blockIdx.x * TB + el;
    for (int x = threadIdx.x; x < N1 * n1; x += TB * KSPLIT) W1s[x] = W1[(long)b * N1 * n1 + x];
    for (int x = threadIdx.x; x < nchunk * KC * n0; x += TB * KSPLIT)
        W0s[x] = (x < N0 * n0) ? W0[(long)b * N0 * n0 + x] : cmake(0., 0.);
    const cplx* zsrc = Z + (long)b * z_bstride + t;
    for (int x = h; x < n0 * n1; x += KSPLIT) Zs[x * TB + el] = (t < S2) ? __ldg(zsrc + (long)x * S2) : cmake(0., 0.);
    __syncthreads();
    if (t >= S2) return;
    cplx* dst = X + (long)b * x_bstride + t;
    const int k1lo = (int)((long)N1 * h / KSPLIT), k1hi = (int)((long)N1 * (h + 1) / KSPLIT);
    for (int k1 = k1lo; k1 < k1hi; k1++) {
        const cplx* w1 = W1s + k1 * n1;
        for (int c0 = 0; c0 < N0; c0 += KC) {
            cplx acc[KC];
#pragma unroll
            for (int kk = 0; kk < KC; kk++) acc[kk] = cmake(0., 0.);
            const cplx* w0 = W0s + c0 * n0;
            for (int r0 = 0; r0 < n0; r0++) {
                const cplx* zrow = Zs + (size_t)r0 * n1 * TB + el;
                cplx y0 = cmake(0., 0.), y1 = cmake(0., 0.);
                int r1 = 0;
                for (; r1 + 1 < n1; r1 += 2) {
                    cfma(y0, w1[r1], zrow[r1 * TB]);
                    cfma(y1, w1[r1 + 1], zrow[(r1 + 1) * TB]);
                }
                if (r1 < n1) cfma(y0, w1[r1], zrow[r1 * TB]);
                const cplx y = cadd(y0, y1);
#pragma unroll
                for (int kk = 0; kk < KC; kk++) cfma(acc[kk], w0[kk * n0 + r0], y);
            }
#pragma unroll
            for (int kk = 0; kk < KC; kk++)
                if (c0 + kk < N0) dst[((long)(c0 + kk) * N1 + k1) * S2] = acc[kk];
        }
    }
}

// Same pass with the loops re-ordered for a small register footprint (more CTAs per SM): per k1 the n0 partial sums
//   y[r0] = sum_{r1} W1[k1][r1] Z[r0][r1][t]
// go to REGISTERS (n0 <= NR0, compile time), then every k0 is one short dot product  X[k0][k1][t] = sum_{r0} W0[k0][r0] y[r0]
// that is stored at once -- no array of N0 accumulators.  Same arithmetic, same order of the r1 / r0 sums.
template <int NR0, int TB, int KSPLIT, int MINB, int UK>
__global__ void __launch_bounds__(TB * KSPLIT, MINB)
wb_axis10_fused_y_kernel(const cplx* __restrict__ Z, cplx* __restrict__ X, const cplx* __restrict__ W1,
                         const cplx* __restrict__ W0, int n0, int n1, int N0, int N1, long S2, long z_bstride,
                         long x_bstride) {
    extern __shared__ cplx sm_fy[];
    cplx* Zs = sm_fy;                         // [n0*n1][TB]
    cplx* W1s = Zs + (size_t)n0 * n1 * TB;    // [N1][n1]
    cplx* W0s = W1s + N1 * n1;                // [N0][NR0], zero beyond n0
    const int b = blockIdx.z;
    const int el = threadIdx.x % TB, h = threadIdx.x / TB;
    const long t = (long)blockIdx.x * TB + el;
    for (int x = threadIdx.x; x < N1 * n1; x += TB * KSPLIT) W1s[x] = W1[(long)b * N1 * n1 + x];
    for (int x = threadIdx.x; x < N0 * NR0; x += TB * KSPLIT) {
        const int k = x / NR0, r = x - k * NR0;
        W0s[x] = (r < n0) ? W0[((long)b * N0 + k) * n0 + r] : cmake(0., 0.);
    }
    const cplx* zsrc = Z + (long)b * z_bstride + t;
    for (int x = h; x < n0 * n1; x += KSPLIT) Zs[x * TB + el] = (t < S2) ? __ldg(zsrc + (long)x * S2) : cmake(0., 0.);
    __syncthreads();
    if (t >= S2) return;
    cplx* dst = X + (long)b * x_bstride + t;
    const int k1lo = (int)((long)N1 * h / KSPLIT), k1hi = (int)((long)N1 * (h + 1) / KSPLIT);
    for (int k1 = k1lo; k1 < k1hi; k1++) {
        const cplx* w1 = W1s + k1 * n1;
        cplx y[NR0];
#pragma unroll
        for (int r0 = 0; r0 < NR0; r0++) {
            cplx y0 = cmake(0., 0.), y1 = cmake(0., 0.);
            if (r0 < n0) {
                const cplx* zrow = Zs + (size_t)r0 * n1 * TB + el;
                int r1 = 0;
                for (; r1 + 1 < n1; r1 += 2) {
                    cfma(y0, w1[r1], zrow[r1 * TB]);
                    cfma(y1, w1[r1 + 1], zrow[(r1 + 1) * TB]);
                }
                if (r1 < n1) cfma(y0, w1[r1], zrow[r1 * TB]);
            }
            y[r0] = cadd(y0, y1);
        }
        cplx* d1 = dst + (long)k1 * S2;
#pragma unroll UK
        for (int k0 = 0; k0 < N0; k0++) {
            const cplx* w0 = W0s + k0 * NR0;
            cplx a0 = cmake(0., 0.), a1 = cmake(0., 0.);
#pragma unroll
            for (int r0 = 0; r0 + 1 < NR0; r0 += 2) {
                cfma(a0, w0[r0], y[r0]);
                cfma(a1, w0[r0 + 1], y[r0 + 1]);
            }
            if (NR0 & 1) cfma(a0, w0[NR0 - 1], y[NR0 - 1]);
            d1[(long)k0 * N1 * S2] = cadd(a0, a1);
        }
    }
}

// k-points of a K-block, bit-identical to the reference:
//   points_FFT = ix * (1./N)   (grid/grid.py:68-75) ;  kpoints_all = (points_FFT + dK) % 1  (data_K.py:146-151)
__global__ void wb_kpoints_kernel(const double* __restrict__ dK, int3 N, double* __restrict__ kpts) {
    int ik = blockIdx.x * blockDim.x + threadIdx.x;
    int nk = N.x * N.y * N.z;
    if (ik >= nk) return;
    int iz = ik % N.z, iy = (ik / N.z) % N.y, ix = ik / (N.z * N.y);
    int idx[3] = {ix, iy, iz};
    int NN[3] = {N.x, N.y, N.z};
    for (int d = 0; d < 3; d++) {
        double dk = 1. / (double)NN[d];
        double v = __dmul_rn((double)idx[d], dk);
        v = __dadd_rn(v, dK[d]);
        // python's float % 1 for v >= 0 is fmod(v, 1); for v < 0 it adds 1 when the remainder is non-zero
        double m = fmod(v, 1.0);
        if (m < 0) m = __dadd_rn(m, 1.0);
        kpts[3 * ik + d] = m;
    }
}
