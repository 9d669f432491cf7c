// Shared-memory building blocks of the rotation kernels:  Xbar = U^dagger X U  (Data_K._rotate,
// data_K/data_K.py:130-132) and the energy denominators of D_H (data_K.py:290-298, 324-326).
#pragma once
#include "wb_common.cuh"
#include "wb_groups.cuh"

struct WbFormulaFlags {
    int formula;    // WBGPU_OMEGA, ...
    int internal_terms;
    int external_terms;
};

// 1/(E_a - E_b) with the reference's cutoff (data_K.py:292-297)
__device__ __forceinline__ double wb_deinv(double Ea, double Eb) {
    double d = Ea - Eb;
    return (fabs(d) < 1e-7) ? 0. : 1. / d;
}

// C = U^dagger (X U) for one full matrix X held in Xs; result to Cs.  All nw x nw, row-major.
template <int NT>
__device__ __forceinline__ void wb_rotate_smem(const cplx* Us, const cplx* Xs, cplx* Ys, cplx* Cs, int nw) {
    for (int x = threadIdx.x; x < nw * nw; x += NT) {
        int i = x / nw, l = x % nw;
        cplx acc = cmake(0., 0.);
        for (int j = 0; j < nw; j++) cfma(acc, Xs[i * nw + j], Us[j * nw + l]);
        Ys[x] = acc;
    }
    __syncthreads();
    for (int x = threadIdx.x; x < nw * nw; x += NT) {
        int n = x / nw, l = x % nw;
        cplx acc = cmake(0., 0.);
        for (int i = 0; i < nw; i++) cfma_conj(acc, Us[i * nw + n], Ys[i * nw + l]);
        Cs[x] = acc;
    }
    __syncthreads();
}

template <int NT>
__device__ __forceinline__ void wb_load_channel(const cplx* __restrict__ rec, int off, bool herm, cplx* Xs, int nw) {
    for (int x = threadIdx.x; x < nw * nw; x += NT) {
        int i = x / nw, j = x % nw;
        Xs[x] = herm ? load_herm(rec, off, i, j, nw) : rec[off + x];
    }
    __syncthreads();
}

