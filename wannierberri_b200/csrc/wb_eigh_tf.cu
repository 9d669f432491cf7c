// Translation unit of the twisted-factorisation eigensolver kernels (wb_eigh_tf.cuh).
#include "wb_launch.h"
#include "wb_eigh_tf.cuh"

template <int NW, bool VEC, int MINB_ = 0>
static int launch_trideig(long k0, long nk, const double* d, const double* e, double* E, double* Z, int* fail_list, int* nfail,
                          cudaStream_t stream) {
    constexpr int NT = 128;
    constexpr int MINB = MINB_ ? MINB_ : !VEC ? 4 : (NW <= 18) ? 4 : (NW <= 24) ? 3 : 2;   // 128 / 168 / 255 registers per thread
    constexpr int smem = (NT / 32) * wb_trideig_smem_doubles_per_warp<NW>() * 8;
    auto kern = wb_trideig_kernel<NW, NT, MINB, VEC>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err == cudaSuccess) err = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (err != cudaSuccess) return (int)err;
    kern<<<(unsigned)((nk + NT - 1) / NT), NT, smem, stream>>>(k0, nk, d, e, E, Z, fail_list, nfail);
    return (int)cudaGetLastError();
}

template <int NW, int MB, int MINB>
static int launch_backtransform_t(long k0, long nk, const double* Z, const cplx* tau, cplx* VU, cudaStream_t stream) {
    constexpr int smem = wb_backtransform_smem_bytes<NW, MB>();
    constexpr int NT = (MB * NW + 31) / 32 * 32;
    auto kern = wb_backtransform_kernel<NW, MB, MINB>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err == cudaSuccess) err = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (err != cudaSuccess) return (int)err;
    kern<<<(unsigned)((nk + MB - 1) / MB), NT, smem, stream>>>(k0, nk, Z, tau, VU);
    return (int)cudaGetLastError();
}

template <int NW>
static int launch_backtransform(long k0, long nk, const double* Z, const cplx* tau, cplx* VU, cudaStream_t stream) {
    // 4.5 warps per CTA; three CTAs per SM at 128 registers per thread (nw <= 20: the 18 complex components of a vector
    // and the staged reflector loads fit with ~0.4 KB of spills), measured faster than 7- and 9-warp CTAs at nw = 18
    return launch_backtransform_t<NW, (144 / NW > 0 ? 144 / NW : 1), (NW <= 20) ? 3 : 2>(k0, nk, Z, tau, VU, stream);
}

#define WB_TF_SIZES WB_CASE(4) WB_CASE(5) WB_CASE(6) WB_CASE(7) WB_CASE(8) WB_CASE(9) WB_CASE(10) WB_CASE(11) WB_CASE(12) \
    WB_CASE(13) WB_CASE(14) WB_CASE(15) WB_CASE(16) WB_CASE(17) WB_CASE(18) WB_CASE(19) WB_CASE(20) WB_CASE(21) WB_CASE(22)  \
    WB_CASE(23) WB_CASE(24)

int wb_launch_trideig(int nw, bool vectors, long k0, long nk, const double* d, const double* e, double* E, double* Z,
                      int* fail_list, int* nfail, cudaStream_t stream) {
    switch (nw) {
#define WB_CASE(N)                                                                                          \
    case N:                                                                                                 \
        return vectors ? launch_trideig<N, true>(k0, nk, d, e, E, Z, fail_list, nfail, stream)              \
                       : launch_trideig<N, false>(k0, nk, d, e, E, Z, fail_list, nfail, stream);
        WB_TF_SIZES
#undef WB_CASE
    }
    return -1;
}

int wb_launch_backtransform(int nw, long k0, long nk, const double* Z, const cplx* tau, cplx* VU, cudaStream_t stream) {
    switch (nw) {
#define WB_CASE(N) case N: return launch_backtransform<N>(k0, nk, Z, tau, VU, stream);
        WB_TF_SIZES
#undef WB_CASE
    }
    return -1;
}
