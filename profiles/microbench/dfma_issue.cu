// DFMA issue interval / latency seen by ONE warp, as a function of the independent chains per warp (ILP) and of the
// warps resident per SM.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_issue dfma_issue.cu && ./dfma_issue
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, int iters, double a, double b, long long* cyc) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i];
    if (s == 12345.678) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int ILP>
void run(int warps_per_block, int blocks_per_sm) {
    double* out; long long* cyc; cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
    int iters = 2000;
    k<ILP><<<148 * blocks_per_sm, 32 * warps_per_block>>>(out, iters, 1.0000001, 1e-9, cyc);
    cudaDeviceSynchronize();
    k<ILP><<<148 * blocks_per_sm, 32 * warps_per_block>>>(out, iters, 1.0000001, 1e-9, cyc);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("ILP %2d  warps/block %d blocks/SM %d : %.2f clk per DFMA per warp, %.2f clk per DFMA per SM-quarter-equivalent\n", ILP,
           warps_per_block, blocks_per_sm, (double)c / (iters * 8.0 * ILP), (double)c / (iters * 8.0 * ILP) / (warps_per_block * blocks_per_sm) * 4);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w : {1, 2, 4, 8}) {
        run<1>(w, 1); run<2>(w, 1); run<4>(w, 1); run<8>(w, 1); run<16>(w, 1);
    }
    run<8>(1, 2); run<8>(1, 4); run<16>(1, 2);
    return 0;
}
