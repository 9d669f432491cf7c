// Size dispatch of the fused DMMA rotation + formula kernel (instantiations: wb_rotate_mma_{a,b,c}.cu).
#include "wb_launch.h"
#include "wb_groups.cuh"
#include "wb_events_generic.cuh"

#define WB_MMA_ARGS bool trim, int rot_r2, const cplx* rec, const WbLayout& L, long nk, const double* E, const cplx* U, \
    const WbWindow& win, const WbEventLayout& ev, double* label, double* val, int smem_optin, int sms, cudaStream_t stream
int wb_launch_mma_events_a(int nw, WB_MMA_ARGS);
int wb_launch_mma_events_b(int nw, WB_MMA_ARGS);
int wb_launch_mma_events_c(int nw, WB_MMA_ARGS);

int wb_launch_mma_events(int nw, WB_MMA_ARGS) {
    if (nw & 1) return -1;
    if (nw <= 12) return wb_launch_mma_events_a(nw, trim, rot_r2, rec, L, nk, E, U, win, ev, label, val, smem_optin, sms, stream);
    if (nw <= 18) return wb_launch_mma_events_b(nw, trim, rot_r2, rec, L, nk, E, U, win, ev, label, val, smem_optin, sms, stream);
    return wb_launch_mma_events_c(nw, trim, rot_r2, rec, L, nk, E, U, win, ev, label, val, smem_optin, sms, stream);
}
