// Fused DMMA rotation + formula kernel, num_wann = 14, 16, 18.
#include "wb_rotate_mma_inst.cuh"

int wb_launch_mma_events_b(int nw, WB_MMA_ARGS) {
    switch (nw) {
        case 14: return wb_mma_launch<14>(WB_MMA_PASS);
        case 16: return wb_mma_launch<16>(WB_MMA_PASS);
        case 18: return wb_mma_launch<18>(WB_MMA_PASS);
    }
    return -1;
}
