// Fused DMMA rotation + formula kernel, num_wann = 20, 22, 24.
#include "wb_rotate_mma_inst.cuh"

int wb_launch_mma_events_c(int nw, WB_MMA_ARGS) {
    switch (nw) {
        case 20: return wb_mma_launch<20>(WB_MMA_PASS);
        case 22: return wb_mma_launch<22>(WB_MMA_PASS);
        case 24: return wb_mma_launch<24>(WB_MMA_PASS);
    }
    return -1;
}
