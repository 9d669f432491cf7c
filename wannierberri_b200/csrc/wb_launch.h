// Host-side launchers of kernels that live in translation units of their own (compiled in parallel with
// wbgpu_api.cu).  Return value: 0 = launched, -1 = this size has no instantiation (the caller takes its generic
// path), otherwise a cudaError_t.
#pragma once
#include <cuda_runtime.h>
#include "wb_common.cuh"

// wb_eigh_tpm.cuh: thread-per-matrix Householder tridiagonalisation (nw even, 4 <= nw <= 20, and odd sizes in between)
int wb_launch_tridiag_tpm(int nw, const cplx* rec, const WbLayout& L, long k0, long nk, double* d, double* e, cplx* tau,
                          cplx* V, cudaStream_t stream);

// the same with a pair of lanes per matrix (16 matrices per warp, half the shared memory per warp)
int wb_launch_tridiag_tpm2(int nw, const cplx* rec, const WbLayout& L, long k0, long nk, double* d, double* e, cplx* tau,
                           cplx* V, cudaStream_t stream);

// wb_eigh_tf.cuh: eigenvalues (implicit QL) and eigenvectors (twisted factorisation) of the tridiagonal matrices, thread
// per matrix, and the back-transformation with one lane per (matrix, eigenvector); 4 <= nw <= 24.
// d, e, tau, Z are indexed from the start of the chunk, E and VU from k0.
int wb_launch_trideig(int nw, bool vectors, long k0, long nk, const double* d, const double* e, double* E, double* Z,
                      int* fail_list, int* nfail, cudaStream_t stream);
int wb_launch_backtransform(int nw, long k0, long nk, const double* Z, const cplx* tau, cplx* VU, cudaStream_t stream);

// wb_rotate_mma.cuh: fused DMMA rotation + Omega / Morb_Hpm formulae, compile-time num_wann (even, 4 .. 24).
// Returns -1 when the kernel does not cover the request (formula mask, num_wann, shared memory).
struct WbWindow;
struct WbEventLayout;
int wb_launch_mma_events(int nw, bool trim, int rot_r2, const cplx* rec, const WbLayout& L, long nk, const double* E,
                         const cplx* U, const WbWindow& win, const WbEventLayout& ev, double* label, double* val,
                         int smem_optin, int sms, cudaStream_t stream);
