// Launcher template of the fused DMMA rotation + formula kernel (wb_rotate_mma.cuh); the sizes are instantiated in the
// translation units wb_rotate_mma_{a,b,c}.cu so that they compile in parallel.
#pragma once
#include "wb_launch.h"
#include "wb_rotate_mma.cuh"

template <int NW, bool TRIM>
static int wb_mma_launch_t(const cplx* rec, const WbLayout& L, const WbMmaPlan& P, long nk, const double* E, const cplx* U,
                           const WbWindow& win, const WbEventLayout& ev, double* label, double* val, int smem_optin, int sms,
                           cudaStream_t stream) {
    const size_t smem = wb_mma_smem_bytes<NW>(P);
    if ((int)smem > smem_optin) return -1;
    cudaError_t err = cudaFuncSetAttribute(wb_events_mma_kernel<NW, TRIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return (int)err;
    const long nblk = nk < (long)sms * 2 ? nk : (long)sms * 2;
    wb_events_mma_kernel<NW, TRIM><<<(unsigned)nblk, 128, smem, stream>>>(rec, L, P, nk, E, U, win, ev, label, val);
    return (int)cudaGetLastError();
}

template <int NW>
static int wb_mma_launch(bool trim, int rot_r2, const cplx* rec, const WbLayout& L, long nk, const double* E, const cplx* U,
                         const WbWindow& win, const WbEventLayout& ev, double* label, double* val, int smem_optin, int sms,
                         cudaStream_t stream) {
    WbMmaPlan P;
    if (!wb_mma_make_plan<NW>(L, ev.mask, ev.external_terms, &P)) return -1;
    trim = trim && L.dH_herm;
    // with trimmed columns step 2 has fewer stacked tiles than warps x rounds: give the warp that carries two of them the
    // single step-1 tile of the NEXT item, they run between the same pair of barriers
    if (trim) for (int i = 0; i < P.nitem; i++) P.r2[i] = 2;
    if (rot_r2 >= 0) for (int i = 0; i < P.nitem; i++) P.r2[i] = rot_r2;
    return trim ? wb_mma_launch_t<NW, true>(rec, L, P, nk, E, U, win, ev, label, val, smem_optin, sms, stream)
                : wb_mma_launch_t<NW, false>(rec, L, P, nk, E, U, win, ev, label, val, smem_optin, sms, stream);
}

#define WB_MMA_ARGS bool trim, int rot_r2, const cplx* rec, const WbLayout& L, long nk, const double* E, const cplx* U, \
    const WbWindow& win, const WbEventLayout& ev, double* label, double* val, int smem_optin, int sms, cudaStream_t stream
#define WB_MMA_PASS trim, rot_r2, rec, L, nk, E, U, win, ev, label, val, smem_optin, sms, stream
