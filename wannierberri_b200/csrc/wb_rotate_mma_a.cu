// Fused DMMA rotation + formula kernel, num_wann = 4, 6, 8, 10, 12.
#include "wb_rotate_mma_inst.cuh"

int wb_launch_mma_events_a(int nw, WB_MMA_ARGS) {
    switch (nw) {
        case 4: return wb_mma_launch<4>(WB_MMA_PASS);
        case 6: return wb_mma_launch<6>(WB_MMA_PASS);
        case 8: return wb_mma_launch<8>(WB_MMA_PASS);
        case 10: return wb_mma_launch<10>(WB_MMA_PASS);
        case 12: return wb_mma_launch<12>(WB_MMA_PASS);
    }
    return -1;
}
