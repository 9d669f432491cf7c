// Standalone timing harness of the rotation+formula kernel with stages switched off (DBG bits):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../wannierberri_b200/csrc -o rotate_bench rotate_bench.cu
#include <cstdio>
#include <vector>
#include <cmath>
#include "wb_rotate_mma.cuh"
template <int DBG>
float run(const cplx* X, WbLayout L, WbMmaPlan P, long nk, const double* E, const cplx* U, double* lab, double* val, int ctas = 296,
          size_t extra_smem = 0) {
    constexpr int NW = 18;
    WbWindow win{12., 22., 0.005, 1e-4, 0, 1, 2000};
    WbEventLayout ev{};
    ev.mask = 2; ev.NC = 3; ev.internal_terms = 1; ev.external_terms = 1;
    size_t smem = wb_mma_smem_bytes<NW>(P) + extra_smem;
    cudaFuncSetAttribute(wb_events_mma_kernel<NW, false, DBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        wb_events_mma_kernel<NW, false, DBG><<<ctas, 128, smem>>>(X, L, P, nk, E, U, win, ev, lab, val);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = fminf(best, ms);
    }
    cudaError_t err = cudaGetLastError();
    printf("DBG=%d  %.3f ms for %ld k-points  (%s) smem %zu\n", DBG, best, nk, cudaGetErrorString(err), smem);
    return best;
}
int main() {
    setvbuf(stdout, NULL, _IONBF, 0);
    constexpr int NW = 18;
    long nk = 128000;
    WbLayout L; L.nw = NW; L.ntri = NW * (NW + 1) / 2;
    int off = 0;
    auto take = [&](bool h) { int o = off; off += h ? L.ntri : NW * NW; return o; };
    L.off_H = take(true);
    for (int a = 0; a < 3; a++) L.off_dH[a] = take(false);
    for (int a = 0; a < 3; a++) L.off_A[a] = take(true);
    for (int a = 0; a < 3; a++) L.off_O[a] = take(true);
    for (int a = 0; a < 3; a++) L.off_B[a] = L.off_C[a] = L.off_S[a] = -1;
    for (int a = 0; a < 6; a++) L.off_W[a] = -1;
    L.dH_herm = 0;
    L.E = off;
    WbMmaPlan P; wb_mma_make_plan<NW>(L, 2, 1, &P);
    std::vector<double> hx((size_t)4096 * L.E * 2);
    for (size_t i = 0; i < hx.size(); i++) hx[i] = sin(0.001 * i) * 0.5;
    cplx *X, *U; double *E, *lab, *val;
    cudaMalloc(&X, sizeof(cplx) * nk * L.E); cudaMalloc(&U, sizeof(cplx) * nk * NW * NW);
    cudaMalloc(&E, 8 * nk * NW); cudaMalloc(&lab, 8 * nk * NW); cudaMalloc(&val, 8 * nk * NW * 3);
    for (long k = 0; k < nk; k += 4096) cudaMemcpy(X + k * L.E, hx.data(), sizeof(cplx) * std::min(4096L, nk - k) * L.E, cudaMemcpyHostToDevice);
    std::vector<double> hu((size_t)nk * NW * NW * 2), he((size_t)nk * NW);
    for (size_t i = 0; i < hu.size(); i++) hu[i] = cos(0.003 * i) * 0.3;
    for (size_t i = 0; i < he.size(); i++) he[i] = 10. + (i % NW) * 2. + 0.3 * sin(0.01 * i);
    cudaMemcpy(U, hu.data(), hu.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(E, he.data(), he.size() * 8, cudaMemcpyHostToDevice);
    run<0>(X, L, P, nk, E, U, lab, val);
    run<1>(X, L, P, nk, E, U, lab, val);
    run<2>(X, L, P, nk, E, U, lab, val);
    run<3>(X, L, P, nk, E, U, lab, val);
    run<4>(X, L, P, nk, E, U, lab, val);
    printf("one CTA per SM:\n");
    run<0>(X, L, P, nk, E, U, lab, val, 148, 90 * 1024);
    run<1>(X, L, P, nk, E, U, lab, val, 148, 90 * 1024);
    run<3>(X, L, P, nk, E, U, lab, val, 148, 90 * 1024);
    return 0;
}
