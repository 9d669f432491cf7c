// Translation unit of the thread-per-matrix eigensolver kernels (wb_eigh_tpm.cuh).
#include "wb_launch.h"
#include "wb_eigh_tpm.cuh"

template <int NW>
static int launch_tridiag_tpm(const cplx* rec, const WbLayout& L, long k0, long nk, double* d, double* e, cplx* tau, cplx* V,
                              cudaStream_t stream) {
    constexpr int smem = wb_tpm_smem_bytes<NW>();
    static bool configured = false;   // per process; the attribute is per function and device-independent in practice
    cudaError_t err = cudaFuncSetAttribute(wb_tridiag_tpm_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err != cudaSuccess) return (int)err;
    configured = true;
    (void)configured;
    wb_tridiag_tpm_kernel<NW><<<(unsigned)((nk + 31) / 32), 32, smem, stream>>>(rec, L, k0, nk, d, e, tau, V);
    return (int)cudaGetLastError();
}

template <int NW>
static int launch_tridiag_tpm2(const cplx* rec, const WbLayout& L, long k0, long nk, double* d, double* e, cplx* tau, cplx* V,
                               cudaStream_t stream) {
    constexpr int smem = wb_tpm2_smem_bytes<NW>();
    cudaError_t err = cudaFuncSetAttribute(wb_tridiag_tpm2_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err == cudaSuccess)
        err = cudaFuncSetAttribute(wb_tridiag_tpm2_kernel<NW>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (err != cudaSuccess) return (int)err;
    wb_tridiag_tpm2_kernel<NW><<<(unsigned)((nk + 15) / 16), 32, smem, stream>>>(rec, L, k0, nk, d, e, tau, V);
    return (int)cudaGetLastError();
}

int wb_launch_tridiag_tpm2(int nw, const cplx* rec, const WbLayout& L, long k0, long nk, double* d, double* e, cplx* tau,
                           cplx* V, cudaStream_t stream) {
    switch (nw) {
#define WB_CASE(N) case N: return launch_tridiag_tpm2<N>(rec, L, k0, nk, d, e, tau, V, stream);
        WB_CASE(4) WB_CASE(5) WB_CASE(6) WB_CASE(7) WB_CASE(8) WB_CASE(9) WB_CASE(10) WB_CASE(11) WB_CASE(12)
        WB_CASE(13) WB_CASE(14) WB_CASE(15) WB_CASE(16) WB_CASE(17) WB_CASE(18) WB_CASE(19) WB_CASE(20)
#undef WB_CASE
    }
    return -1;
}

int wb_launch_tridiag_tpm(int nw, const cplx* rec, const WbLayout& L, long k0, long nk, double* d, double* e, cplx* tau,
                          cplx* V, cudaStream_t stream) {
    switch (nw) {
#define WB_CASE(N) case N: return launch_tridiag_tpm<N>(rec, L, k0, nk, d, e, tau, V, stream);
        WB_CASE(4) WB_CASE(5) WB_CASE(6) WB_CASE(7) WB_CASE(8) WB_CASE(9) WB_CASE(10) WB_CASE(11) WB_CASE(12)
        WB_CASE(13) WB_CASE(14) WB_CASE(15) WB_CASE(16) WB_CASE(17) WB_CASE(18) WB_CASE(19) WB_CASE(20)
#undef WB_CASE
    }
    return -1;
}
