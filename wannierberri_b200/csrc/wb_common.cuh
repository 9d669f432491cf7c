// Common device helpers and the per-k-point record layout.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math_constants.h>

typedef double2 cplx;  // (re, im)

__host__ __device__ __forceinline__ cplx cmake(double r, double i) { return make_double2(r, i); }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return cmake(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return cmake(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return cmake(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * conj(b)
__device__ __forceinline__ cplx cmulc(cplx a, cplx b) {
    return cmake(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
// conj(a) * b
__device__ __forceinline__ cplx cconjmul(cplx a, cplx b) {
    return cmake(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ cplx cconj(cplx a) { return cmake(a.x, -a.y); }
__device__ __forceinline__ cplx cscale(double s, cplx a) { return cmake(s * a.x, s * a.y); }
// acc += a*b  (4 DFMA)
__device__ __forceinline__ void cfma(cplx& acc, cplx a, cplx b) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
}
// acc += conj(a)*b
__device__ __forceinline__ void cfma_conj(cplx& acc, cplx a, cplx b) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(-a.y, b.x, acc.y);
}

// Packed index of element (i, j), i <= j, of an upper triangle stored row by row.
__host__ __device__ __forceinline__ int tri_index(int i, int j, int nw) {
    return i * nw - (i * (i - 1)) / 2 + (j - i);
}

// Per-k-point record: all Wannier-gauge matrices of one k-point, contiguous.  Hermitised
// channels (H, A, curl A, S) hold the upper triangle only (nw(nw+1)/2 complex), the others the
// full nw*nw matrix.  Offsets are in complex elements; -1 = channel absent.
struct WbLayout {
    int nw;
    int ntri;        // nw(nw+1)/2
    int E;           // record length (complex elements)
    int off_H;       // hermitian
    int dH_herm;     // d_a H is hermitian in R-space (checked at plan time): its channels are packed like H
    int off_dH[3];   // full, or upper triangle if dH_herm
    int off_A[3];    // hermitian
    int off_O[3];    // hermitian  (curl A)
    int off_B[3];    // full
    int off_C[3];    // full
    int off_S[3];    // hermitian
    int off_W[6];    // d_b d_d H, (b, d) = xx xy xz yy yz zz; packed like d_a H (triangle iff dH_herm)
    // spin-current matrices (full): [3 velocity][3 spin] -> index 3 a + s; SH: [3 spin]
    int off_SA[9], off_SHA[9], off_SR[9], off_SH[3], off_SHR[9];
    // comma-derivatives d_d A_b and d_d rotA_c (index 3 b + d), hermitian like A and rotA (data_K_R.py:84-87)
    int off_dA[9], off_dO[9];
    int off_dS[9];   // d_d S_s, hermitian
    int off_dB[9], off_dC[9];   // d_d B_b, d_d C_c (full)
    int off_W3[10];  // d_b d_c d_d H, sorted triples xxx xxy xxz xyy xyz xzz yyy yyz yzz zzz; packed like d_a H
    // second comma-derivatives d_d d_e X_b (index 6 b + wb_sym6(d, e)) for plug-in formulae (Data_K_R.Xbar(name, 2)):
    // A, rotA, S hermitian-packed; B, C full
    int off_d2A[18], off_d2O[18], off_d2S[18], off_d2B[18], off_d2C[18];
};

// index of the symmetric pair (b, d) in off_W
__host__ __device__ __forceinline__ int wb_sym6(int b, int d) {
    const int lo = b < d ? b : d, hi = b < d ? d : b;
    return lo * 3 - (lo * (lo - 1)) / 2 + (hi - lo);
}

// load element (i,j) of a channel of a record
__device__ __forceinline__ cplx load_herm(const cplx* __restrict__ rec, int off, int i, int j, int nw) {
    if (i <= j) return rec[off + tri_index(i, j, nw)];
    return cconj(rec[off + tri_index(j, i, nw)]);
}

// Levi-Civita helper index arrays of the reference (utility.py:45-46)
#define WB_ALPHA(c) (((c) + 1) % 3)
#define WB_BETA(c) (((c) + 2) % 3)
