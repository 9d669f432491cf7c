// TMA 1-D bulk copies (cp.async.bulk -> SASS UBLKCP) and mbarrier helpers.  sm_90+ PTX, used on sm_100a.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t wb_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wb_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(wb_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void wb_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(wb_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void wb_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     wb_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(wb_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void wb_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok)
            : "r"(wb_smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}

