// Peak probes for the roofline denominators that MEASURED_PEAKS.json does not hold:
// FP64 vector FMA (DFMA) and FP64 mma.sync.m8n8k4 (DMMA) throughput of this GPU.
#pragma once
#include "wb_common.cuh"

// 8 independent FMA chains per thread, ITERS iterations: 16 flops * ITERS * 8 / 8 ...
template <int CHAINS>
__global__ void __launch_bounds__(256) wb_dfma_probe_kernel(double* out, int iters, double a, double b) {
    double acc[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) acc[c] = threadIdx.x * 1e-3 + c;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) acc[c] = fma(acc[c], a, b);
    }
    double s = 0.;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s += acc[c];
    if (s == 12345.678) out[0] = s;  // never true; keeps the chains alive
}

__device__ __forceinline__ void wb_dmma_m8n8k4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

template <int CHAINS>
__global__ void __launch_bounds__(256) wb_dmma_probe_kernel(double* out, int iters) {
    double d0[CHAINS], d1[CHAINS];
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * (threadIdx.x + 1);
#pragma unroll
    for (int c = 0; c < CHAINS; c++) { d0[c] = c; d1[c] = -c; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) wb_dmma_m8n8k4(d0[c], d1[c], a, b);
    }
    double s = 0.;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s += d0[c] + d1[c];
    if (s == 12345.678) out[0] = s;
}
