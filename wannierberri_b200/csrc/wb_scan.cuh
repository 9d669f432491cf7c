// Fermi-level scan accumulation (StaticCalculator.__call__, calculators/static.py:128-155).
//
// The reference adds every group value to restot[iEf:] (or to all of restot when the label is
// below EFmin).  Here each event is added ONCE, to a histogram bin iEf (or to `below`), in a
// per-CTA shared-memory histogram that is flushed to the global one with atomics; the running
// sum over the Fermi-level axis, the finite-difference stencil for f', f'', f''' and the
// normalisation are linear and are applied once at the end (wb_scan_finalize_kernel).
#pragma once
#include "wb_common.cuh"
#include "wb_groups.cuh"

// Identity formula (covariant.py:10-22): trace over a group = number of bands in it.
__global__ void wb_identity_events_kernel(const double* __restrict__ Eall, int nw, long nk, WbWindow win,
                                          double* __restrict__ ev_label, double* __restrict__ ev_val) {
    extern __shared__ char smem_i[];
    // per-thread scratch in shared memory: E[nw], label[nw], g1[nw], g2[nw]
    const int per = nw * (2 * (int)sizeof(double) + 2 * (int)sizeof(short));
    char* base = smem_i + (size_t)threadIdx.x * ((per + 7) / 8 * 8);
    double* E = (double*)base;
    double* label = E + nw;
    short* g1 = (short*)(label + nw);
    short* g2 = g1 + nw;
    long ik = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ik >= nk) return;
    for (int n = 0; n < nw; n++) E[n] = Eall[ik * nw + n];
    if (win.Ebmin) wb_band_groups_tetra(E, win.Ebmin + ik * nw, win.Ebmax + ik * nw, nw, win, g1, g2, label);
    else wb_band_groups(E, nw, win, g1, g2, label);
    for (int n = 0; n < nw; n++) {
        ev_label[ik * nw + n] = label[n];
        if (label[n] != CUDART_INF) ev_val[ik * nw + n] = (double)(g2[n] - n);
    }
}

// hist[(nEFx + 1)][ncomp]: row 0 = "below" (label < EFmin), row 1 + iEf = bin iEf.
// per_block_stride > 0: one histogram PER K-BLOCK (blockIdx.y = K-block of this launch, weight 1) at
// hist + blockIdx.y * per_block_stride -- the per-K-point results that adaptive refinement needs (run_grid.py:59-72).
__global__ void __launch_bounds__(256)
wb_scan_accumulate_kernel(const double* __restrict__ ev_label, const double* __restrict__ ev_val, int ev_stride,
                          long nslots, int slots_per_block /* nk_block * nw */, const double* __restrict__ weight,
                          int ncomp, WbWindow win, double* __restrict__ hist, int use_smem, long per_block_stride,
                          const double* __restrict__ Eall = nullptr, int nw = 0, unsigned long long sel0 = 0ull,
                          unsigned long long sel1 = 0ull) {
    extern __shared__ double hist_s[];
    const int nrow = win.nEFx + 1;
    long s_begin = 0, s_end = nslots;
    if (per_block_stride > 0) {
        s_begin = (long)blockIdx.y * slots_per_block;
        s_end = s_begin + slots_per_block;
        hist += (size_t)blockIdx.y * per_block_stride;
    }
    if (use_smem) {
        for (int x = threadIdx.x; x < nrow * ncomp; x += blockDim.x) hist_s[x] = 0.;
        __syncthreads();
    }
    double* h = use_smem ? hist_s : hist;
    for (long s = s_begin + (long)blockIdx.x * blockDim.x + threadIdx.x; s < s_end; s += (long)gridDim.x * blockDim.x) {
        double E = ev_label[s];
        if (E == CUDART_INF) continue;
        int row;
        if (E < win.EFmin) row = 0;
        else if (E <= win.EFmax) {
            // iEf = ceil((E - EFmin) / dEF)   (static.py:135), same IEEE operations
            double q = __ddiv_rn(__dsub_rn(E, win.EFmin), win.dEF);
            int iEf = (int)ceil(q);
            if (iEf >= win.nEFx) continue;  // restot[iEf:] is empty
            row = 1 + iEf;
        } else continue;
        double w = (per_block_stride > 0) ? 1. : weight[s / slots_per_block];
        if (Eall) {
            // select_bands: the group that starts at band x ends at the next border (gap > degen_thresh, on an even band
            // with degen_Kramers: grid/tetrahedron.py:131-136); weight = selected bands of the group / its size
            const long k = s / nw;
            const int x = (int)(s - k * nw);
            const double* Ek = Eall + k * nw;
            int b = x + 1;
            while (b < nw && !((Ek[b] - Ek[b - 1] > win.degen_thresh) && (!win.degen_Kramers || (b & 1) == 0))) b++;
            int cnt = 0;
            for (int j = x; j < b; j++) cnt += (int)(((j < 64 ? sel0 >> j : sel1 >> (j - 64)) & 1ull));
            if (cnt == 0) continue;
            w *= (double)cnt / (double)(b - x);
        }
        for (int c = 0; c < ncomp; c++) atomicAdd(&h[row * ncomp + c], w * ev_val[s * ev_stride + c]);
    }
    if (use_smem) {
        __syncthreads();
        for (int x = threadIdx.x; x < nrow * ncomp; x += blockDim.x) {
            double v = hist_s[x];
            if (v != 0.) atomicAdd(&hist[x], v);
        }
    }
}

// restot[e] = below + sum_{e' <= e} hist[e'] ; stencil (static.py:137-147) ; * scale.
// One CTA per component: chunked running sum over the Fermi axis (thread = contiguous chunk, chunk offsets by a
// serial pass over 256 partial sums), then the finite-difference stencil.
// blockIdx.y = K-block in the per-K-block mode (strides of hist / cum / out between K-blocks), else 0.
__global__ void __launch_bounds__(256)
wb_scan_finalize_kernel(const double* __restrict__ hist, double* __restrict__ cum, int ncomp, int nEFx, int nEF,
                        int fder, double dEF, double scale, double* __restrict__ out, long hist_stride, long cum_stride,
                        long out_stride) {
    __shared__ double part[256];
    const int c = blockIdx.x;
    hist += (size_t)blockIdx.y * hist_stride;
    cum += (size_t)blockIdx.y * cum_stride;
    out += (size_t)blockIdx.y * out_stride;
    const int per = (nEFx + 255) / 256;
    const int e0 = threadIdx.x * per, e1 = min(e0 + per, nEFx);
    double s = 0.;
    for (int e = e0; e < e1; e++) s += hist[(1 + e) * ncomp + c];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double run = hist[c];  // "below" row
        for (int i = 0; i < 256; i++) { double v = part[i]; part[i] = run; run += v; }
    }
    __syncthreads();
    double run = part[threadIdx.x];
    for (int e = e0; e < e1; e++) {
        run += hist[(1 + e) * ncomp + c];
        cum[e * ncomp + c] = run;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nEF; e += 256) {
        double v;
        if (fder == 0) v = cum[e * ncomp + c];
        else if (fder == 1) v = (cum[(e + 2) * ncomp + c] - cum[e * ncomp + c]) / (2. * dEF);
        else if (fder == 2)
            v = (cum[(e + 2) * ncomp + c] + cum[e * ncomp + c] - 2. * cum[(e + 1) * ncomp + c]) / (dEF * dEF);
        else
            v = (cum[(e + 4) * ncomp + c] - cum[e * ncomp + c] -
                 2. * (cum[(e + 3) * ncomp + c] - cum[(e + 1) * ncomp + c])) / (2. * dEF * dEF * dEF);
        out[e * ncomp + c] = v * scale;
    }
}
