// DMMA (mma.sync.m8n8k4.f64) issue/latency microbenchmark on one B200:
// achieved DMMA per clock per SM as a function of resident warps per SM and independent accumulator chains per warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_latency dmma_latency.cu && ./dmma_latency
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int CH>
__global__ void probe(double* out, int iters, long long* cyc) {
    double acc[CH][2];
    for (int c = 0; c < CH; c++) acc[c][0] = acc[c][1] = threadIdx.x * 1e-9;
    double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CH; c++) dmma(acc[c][0], acc[c][1], a, b);
    }
    long long t1 = clock64();
    double s = 0;
    for (int c = 0; c < CH; c++) s += acc[c][0] + acc[c][1];
    if (s == 12345.678) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
// DFMA chains for comparison
template <int CH>
__global__ void probe_fma(double* out, int iters, long long* cyc) {
    double acc[CH];
    for (int c = 0; c < CH; c++) acc[c] = threadIdx.x * 1e-9;
    double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CH; c++) acc[c] = fma(acc[c], a, b);
    }
    long long t1 = clock64();
    double s = 0;
    for (int c = 0; c < CH; c++) s += acc[c];
    if (s == 12345.678) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int CH>
void run(int warps_per_sm, double* d_out, long long* d_cyc, bool fma_) {
    const int iters = 20000;
    int threads = warps_per_sm * 32;
    int blocks = 148;
    if (threads > 1024) { blocks = 148 * (threads / 1024); threads = 1024; }
    for (int rep = 0; rep < 2; rep++) {
        if (fma_) probe_fma<CH><<<blocks, threads>>>(d_out, iters, d_cyc);
        else probe<CH><<<blocks, threads>>>(d_out, iters, d_cyc);
    }
    cudaDeviceSynchronize();
    long long cyc;
    cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
    double per_sm_per_clk = (double)iters * CH * warps_per_sm / (double)cyc;
    printf("%s warps/SM %2d chains %d : %8.1f clk per dependent step, %.4f warp-instr/clk/SM\n", fma_ ? "DFMA" : "DMMA",
           warps_per_sm, CH, (double)cyc / iters, per_sm_per_clk);
}
int main() {
    double* d_out; long long* d_cyc;
    cudaMalloc(&d_out, 8); cudaMalloc(&d_cyc, 8);
    for (int fma_ = 0; fma_ < 2; fma_++)
        for (int w : {4, 8, 16, 32}) {
            run<1>(w, d_out, d_cyc, fma_);
            run<2>(w, d_out, d_cyc, fma_);
            run<4>(w, d_out, d_cyc, fma_);
            run<5>(w, d_out, d_cyc, fma_);
            run<8>(w, d_out, d_cyc, fma_);
        }
    return 0;
}
