// Householder tridiagonalisation of a batch of small complex Hermitian matrices, ONE THREAD PER MATRIX.
//
// Replaces the first third of  E_K, UU_K = np.linalg.eigh(HH_K)  (data_K/data_K.py:211-218) for num_wann <= 20;
// same outputs as wb_tridiag_kernel / wb_tridiag2_kernel of wb_eigh_ql.cuh (d, e, tau and the Householder vectors of
// LAPACK's zhetd2 with UPLO = 'L'), so the QL and eigenvector kernels that follow are unchanged.
//
// Why a thread per matrix: with a warp (or half a warp) per matrix the lanes hold rows, 16..18 of 32 lanes work, every
// step pays two warp reductions, two barriers and a scalar zlarfg replicated over the lanes (measured: 34 % of the FP64
// pipe, 23 % of the warp slots).  Here a lane owns its matrix: no shuffles, no barriers, the scalar work of a step
// serves 32 matrices at once and all FP64 instructions are useful.
//
// Data layout: the lower triangle of the matrix lives in shared memory as S[e][lane] (complex128, element-major,
// lane-minor): a lane's element e is always in ITS bank group, so every access of the warp -- whatever element the
// lanes touch -- is a conflict-free 512-byte wavefront set.  NW(NW+1)/2 * 512 B per warp (87.5 KB at NW = 18: two
// warps per SM).  The Householder vector v and the work vector p / w live in REGISTERS with compile-time indices.
//
// To have the active block start at index 0 for every step (compile-time register indices, a runtime loop over the
// steps, no predicated-off work) the kernel runs the UPLO = 'U' recurrence -- which shrinks the active block from the
// bottom -- on the index-reversed matrix B = J A J; mirrored back, its reflectors and (d, e) ARE those of UPLO = 'L'
// on A:  k = NW-1-i,  e[k] = e'[i-1],  d[k] = d'[i],  tau[k] = tau'[i-1],  v_k[r] = v'_i[NW-1-r].
#pragma once
#include "wb_common.cuh"

template <int NW>
__host__ __device__ constexpr int wb_tpm_smem_bytes() { return NW * (NW + 1) / 2 * 32 * 16; }

template <int NW>
__global__ void __launch_bounds__(32)
wb_tridiag_tpm_kernel(const cplx* __restrict__ rec, WbLayout L, long k0, long nk, double* __restrict__ dout,
                      double* __restrict__ eout, cplx* __restrict__ tauout, cplx* __restrict__ Vout) {
    extern __shared__ __align__(16) cplx smem_t[];
    cplx* const S = smem_t + threadIdx.x;                 // element e of this lane's matrix: S[e * 32]
#define WB_S(r, c) S[(((r) * ((r) + 1)) / 2 + (c)) * 32]
    long t = (long)blockIdx.x * 32 + threadIdx.x;
    const bool live = t < nk;
    if (!live) t = nk - 1;                                // idle lanes mirror the last matrix and write nothing
    const long ik = k0 + t;
    {
        // B[r][c] = A[NW-1-r][NW-1-c]; for r >= c that is an element of the stored upper triangle of A
        const cplx* H = rec + ik * L.E + L.off_H;
#pragma unroll
        for (int r = 0; r < NW; r++)
#pragma unroll
            for (int c = 0; c <= r; c++) {
                cplx a = __ldg(&H[tri_index(NW - 1 - r, NW - 1 - c, NW)]);
                if (r == c) a.y = 0.;
                WB_S(r, c) = a;
            }
    }
    double* const d = dout + t * NW;
    double* const e = eout + t * NW;
    cplx* const tau_o = tauout + t * NW;
    cplx* const V = Vout + ik * NW * NW;

    cplx v[NW - 1], p[NW - 1];
#pragma unroll 1
    for (int i = NW - 1; i >= 1; i--) {                   // active block: rows / columns 0 .. i-1; column i is reduced
        const cplx* const Si = S + (size_t)((i * (i + 1)) / 2) * 32;   // row i of the lower triangle: conj of column i
        // ---- zlarfg: x = column i above the diagonal, alpha = x[i-1]
        double xnorm2 = 0.;
#pragma unroll
        for (int r = 0; r < NW - 1; r++) {
            v[r] = cmake(0., 0.);
            if (r < i) {
                const cplx a = Si[r * 32];
                v[r] = cmake(a.x, -a.y);
                if (r < i - 1) xnorm2 = fma(a.x, a.x, fma(a.y, a.y, xnorm2));
            }
        }
        cplx alpha = cmake(0., 0.);
#pragma unroll
        for (int r = 0; r < NW - 1; r++)
            if (r == i - 1) alpha = v[r];
        cplx tau = cmake(0., 0.);
        double beta = alpha.x;
        cplx scale = cmake(0., 0.);
        if (xnorm2 != 0. || alpha.y != 0.) {
            beta = -copysign(sqrt(fma(alpha.x, alpha.x, fma(alpha.y, alpha.y, xnorm2))), alpha.x);
            const double binv = 1. / beta;
            tau = cmake((beta - alpha.x) * binv, -alpha.y * binv);
            const cplx den = cmake(alpha.x - beta, alpha.y);
            const double dn = 1. / (den.x * den.x + den.y * den.y);
            scale = cmake(den.x * dn, -den.y * dn);
        }
        const int k = NW - 1 - i;                         // step of the UPLO = 'L' recurrence on A
#pragma unroll
        for (int r = 0; r < NW - 1; r++) {
            if (r < i - 1) {
                v[r] = cmul(v[r], scale);
                if (live) V[(size_t)(NW - 1 - r) * NW + k] = v[r];
            } else if (r == i - 1) v[r] = cmake(1., 0.);
        }
        if (live) { e[k] = beta; tau_o[k] = tau; }
        // ---- p = tau * B[0:i, 0:i] v   (a lane whose tau is 0 carries p = w = 0 through the update: no divergence)
#pragma unroll
        for (int r = 0; r < NW - 1; r++) p[r] = cmake(0., 0.);
#pragma unroll
        for (int r = 0; r < NW - 1; r++) {
            if (r < i) {
                const cplx vr = v[r];
                cplx acc = p[r];
#pragma unroll
                for (int c = 0; c < r; c++) {
                    const cplx a = WB_S(r, c);
                    cfma(acc, a, v[c]);                   // p_r += B[r][c] v_c
                    cfma_conj(p[c], a, vr);               // p_c += conj(B[r][c]) v_r
                }
                const double dr = WB_S(r, r).x;
                acc.x = fma(dr, vr.x, acc.x);
                acc.y = fma(dr, vr.y, acc.y);
                p[r] = acc;
            }
        }
        cplx dot = cmake(0., 0.);                         // (tau p)^H v
#pragma unroll
        for (int r = 0; r < NW - 1; r++)
            if (r < i) {
                p[r] = cmul(tau, p[r]);
                cfma_conj(dot, p[r], v[r]);
            }
        const cplx al2 = cscale(-0.5, cmul(tau, dot));
#pragma unroll
        for (int r = 0; r < NW - 1; r++)
            if (r < i) cfma(p[r], al2, v[r]);             // w = p + al2 v
        // ---- B[0:i, 0:i] -= v w^H + w v^H   (lower triangle)
#pragma unroll
        for (int r = 0; r < NW - 1; r++) {
            if (r < i) {
                const cplx vr = v[r], wr = p[r];
#pragma unroll
                for (int c = 0; c < r; c++) {
                    cplx a = WB_S(r, c);
                    const cplx wc = p[c], vc = v[c];     // a -= v_r conj(w_c) + w_r conj(v_c)
                    a.x = fma(-vr.x, wc.x, a.x);
                    a.x = fma(-vr.y, wc.y, a.x);
                    a.y = fma(-vr.y, wc.x, a.y);
                    a.y = fma(vr.x, wc.y, a.y);
                    a.x = fma(-wr.x, vc.x, a.x);
                    a.x = fma(-wr.y, vc.y, a.x);
                    a.y = fma(-wr.y, vc.x, a.y);
                    a.y = fma(wr.x, vc.y, a.y);
                    WB_S(r, c) = a;
                }
                cplx a = WB_S(r, r);
                a.x = fma(-2., fma(vr.x, wr.x, vr.y * wr.y), a.x);
                WB_S(r, r) = a;
            }
        }
        if (live) d[k] = Si[i * 32].x;                    // B[i][i], final since the previous step
    }
    if (live) {
        d[NW - 1] = S[0].x;
        e[NW - 1] = 0.;
        tau_o[NW - 1] = cmake(0., 0.);
    }
#undef WB_S
}

// ------------------------------------------------------------------------------------------ two lanes per matrix
// Same reduction with a matrix on a PAIR of lanes (16 matrices per warp): the thread-per-matrix kernel above keeps all
// FP64 work useful but its 87.5 KB of shared memory per warp leave two warps per SM (two of the four schedulers idle:
// measured 2.45 ms per 256k matrices against 2.04 ms for the two-matrices-per-warp kernel of wb_eigh_ql.cuh).  Here lane
// h = 0 / 1 of a pair owns the rows of the (index-reversed) lower triangle with r mod 2 = h -- balanced while the active
// block shrinks -- in its own lane-minor slots: 46 KB per warp at NW = 18, four warps per SM, one per scheduler.
// Per step the lanes share the reduced column (read from the owner's slots), each accumulates its rows' part of
// p = B v (both the row and the column contributions of an element), the two partial vectors meet through 4 NW
// shuffles, the O(NW) scalar work (zlarfg, w) is done twice.
template <int NW>
__host__ __device__ constexpr int wb_tpm2_slots() {   // elements of the lane with the longer rows (odd rows)
    return ((NW / 2) * (NW / 2 + 1) > ((NW + 1) / 2) * ((NW + 1) / 2)) ? (NW / 2) * (NW / 2 + 1) : ((NW + 1) / 2) * ((NW + 1) / 2);
}
template <int NW>
__host__ __device__ constexpr int wb_tpm2_smem_bytes() { return wb_tpm2_slots<NW>() * 32 * 16; }

template <int NW>
__global__ void __launch_bounds__(32)
wb_tridiag_tpm2_kernel(const cplx* __restrict__ rec, WbLayout L, long k0, long nk, double* __restrict__ dout,
                       double* __restrict__ eout, cplx* __restrict__ tauout, cplx* __restrict__ Vout) {
    extern __shared__ __align__(16) cplx smem_t2[];
    const int lane = threadIdx.x, h = lane & 1;
    cplx* const S = smem_t2 + lane;                       // my rows
    cplx* const Sp = smem_t2 + (lane & ~1);               // + parity of a row: the slots of its owner within my pair
    // first slot of row r in its owner's storage: rows 0, 2, 4, .. hold 1, 3, 5, .. elements; rows 1, 3, .. hold 2, 4, ..
#define WB_OFF(r) (((r) & 1) ? ((r) >> 1) * (((r) >> 1) + 1) : ((r) >> 1) * ((r) >> 1))
    constexpr int NQ = (NW + 1) / 2;                      // local rows per lane (the last one of lane 1 may not exist)
    long t = (long)blockIdx.x * 16 + (lane >> 1);
    const bool live = t < nk;
    if (!live) t = nk - 1;                                // idle pairs mirror the last matrix and write nothing
    const long ik = k0 + t;
    {
        // B[r][c] = A[NW-1-r][NW-1-c]; for r >= c that is an element of the stored upper triangle of A
        const cplx* H = rec + ik * L.E + L.off_H;
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            const int r = 2 * q + h;
            if (r < NW) {
#pragma unroll
                for (int c = 0; c <= 2 * q + 1; c++)
                    if (c <= r) {
                        cplx a = __ldg(&H[tri_index(NW - 1 - r, NW - 1 - c, NW)]);
                        if (r == c) a.y = 0.;
                        S[(WB_OFF(r) + c) * 32] = a;
                    }
            }
        }
    }
    __syncwarp();
    double* const d = dout + t * NW;
    double* const e = eout + t * NW;
    cplx* const tau_o = tauout + t * NW;
    cplx* const V = Vout + ik * NW * NW;
    const bool writer = live && (h == 0);

    cplx v[NW - 1], p[NW - 1];
#pragma unroll 1
    for (int i = NW - 1; i >= 1; i--) {                   // active block: rows / columns 0 .. i-1; column i is reduced
        const cplx* const Si = Sp + (i & 1) + (size_t)WB_OFF(i) * 32;   // row i of the lower triangle = conj of column i
        // ---- zlarfg: x = column i above the diagonal, alpha = x[i-1]  (both lanes)
        double xnorm2 = 0.;
        cplx alpha = cmake(0., 0.);
#pragma unroll
        for (int r = 0; r < NW - 1; r++) {
            v[r] = cmake(0., 0.);
            if (r < i) {
                const cplx a = Si[r * 32];
                v[r] = cmake(a.x, -a.y);
                if (r < i - 1) xnorm2 = fma(a.x, a.x, fma(a.y, a.y, xnorm2));
                else alpha = v[r];
            }
        }
        cplx tau = cmake(0., 0.);
        double beta = alpha.x;
        cplx scale = cmake(0., 0.);
        if (xnorm2 != 0. || alpha.y != 0.) {
            const double n2 = fma(alpha.x, alpha.x, fma(alpha.y, alpha.y, xnorm2));
            const double rinv = rsqrt(n2);
            beta = -copysign(n2 * rinv, alpha.x);
            const double binv = -copysign(rinv, alpha.x);
            tau = cmake((beta - alpha.x) * binv, -alpha.y * binv);
            const cplx den = cmake(alpha.x - beta, alpha.y);
            const double dn = __drcp_rn(fma(den.x, den.x, den.y * den.y));
            scale = cmake(den.x * dn, -den.y * dn);
        }
        const int k = NW - 1 - i;                         // step of the UPLO = 'L' recurrence on A
#pragma unroll
        for (int r = 0; r < NW - 1; r++) {
            if (r < i - 1) {
                v[r] = cmul(v[r], scale);
                if (live && ((r & 1) == h)) V[(size_t)(NW - 1 - r) * NW + k] = v[r];
            } else if (r == i - 1) v[r] = cmake(1., 0.);
        }
        if (writer) { e[k] = beta; tau_o[k] = tau; }
        // ---- my rows' part of B[0:i, 0:i] v
#pragma unroll
        for (int r = 0; r < NW - 1; r++) p[r] = cmake(0., 0.);
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            const int r = 2 * q + h;
            if (r < i) {
                const cplx* const Sr = S + (size_t)WB_OFF(r) * 32;
                const cplx vr = h ? v[(2 * q + 1 < NW - 1) ? 2 * q + 1 : 0] : v[(2 * q < NW - 1) ? 2 * q : 0];
                cplx acc0 = cmake(0., 0.), acc1 = cmake(0., 0.);
#pragma unroll
                for (int c = 0; c <= 2 * q; c++)
                    if (c < r) {
                        const cplx a = Sr[c * 32];
                        if (c & 1) cfma(acc1, a, v[c]);   // p_r += B[r][c] v_c
                        else cfma(acc0, a, v[c]);
                        cfma_conj(p[c], a, vr);           // p_c += conj(B[r][c]) v_r
                    }
                const double dr = Sr[r * 32].x;
                acc0.x = fma(dr, vr.x, acc0.x + acc1.x);
                acc0.y = fma(dr, vr.y, acc0.y + acc1.y);
                if (h) { if (2 * q + 1 < NW - 1) p[(2 * q + 1 < NW - 1) ? 2 * q + 1 : 0] = cadd(p[(2 * q + 1 < NW - 1) ? 2 * q + 1 : 0], acc0); }
                else { if (2 * q < NW - 1) p[(2 * q < NW - 1) ? 2 * q : 0] = cadd(p[(2 * q < NW - 1) ? 2 * q : 0], acc0); }
            }
        }
        // ---- the two halves of p meet
#pragma unroll
        for (int r = 0; r < NW - 1; r++)
            if (r < i) {
                p[r].x += __shfl_xor_sync(0xffffffffu, p[r].x, 1);
                p[r].y += __shfl_xor_sync(0xffffffffu, p[r].y, 1);
            }
        cplx dot = cmake(0., 0.);                         // (tau p)^H v
#pragma unroll
        for (int r = 0; r < NW - 1; r++)
            if (r < i) {
                p[r] = cmul(tau, p[r]);
                cfma_conj(dot, p[r], v[r]);
            }
        const cplx al2 = cscale(-0.5, cmul(tau, dot));
#pragma unroll
        for (int r = 0; r < NW - 1; r++)
            if (r < i) cfma(p[r], al2, v[r]);             // w = p + al2 v
        // ---- my rows of B[0:i, 0:i] -= v w^H + w v^H
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            const int r = 2 * q + h;
            if (r < i) {
                cplx* const Sr = S + (size_t)WB_OFF(r) * 32;
                const int rr = h ? ((2 * q + 1 < NW - 1) ? 2 * q + 1 : 0) : ((2 * q < NW - 1) ? 2 * q : 0);
                const cplx vr = h ? v[(2 * q + 1 < NW - 1) ? 2 * q + 1 : 0] : v[(2 * q < NW - 1) ? 2 * q : 0];
                const cplx wr = h ? p[(2 * q + 1 < NW - 1) ? 2 * q + 1 : 0] : p[(2 * q < NW - 1) ? 2 * q : 0];
                (void)rr;
#pragma unroll
                for (int c = 0; c <= 2 * q; c++)
                    if (c < r) {
                        cplx a = Sr[c * 32];
                        const cplx wc = p[c], vc = v[c];   // a -= v_r conj(w_c) + w_r conj(v_c)
                        a.x = fma(-vr.x, wc.x, a.x);
                        a.x = fma(-vr.y, wc.y, a.x);
                        a.y = fma(-vr.y, wc.x, a.y);
                        a.y = fma(vr.x, wc.y, a.y);
                        a.x = fma(-wr.x, vc.x, a.x);
                        a.x = fma(-wr.y, vc.y, a.x);
                        a.y = fma(-wr.y, vc.x, a.y);
                        a.y = fma(wr.x, vc.y, a.y);
                        Sr[c * 32] = a;
                    }
                cplx a = Sr[r * 32];
                a.x = fma(-2., fma(vr.x, wr.x, vr.y * wr.y), a.x);
                Sr[r * 32] = a;
            }
        }
        if (writer) d[k] = Si[i * 32].x;                  // B[i][i], final since the previous step
        __syncwarp();                                     // the next column is read from the partner's slots
    }
    if (writer) {
        d[NW - 1] = Sp[0].x;
        e[NW - 1] = 0.;
        tau_o[NW - 1] = cmake(0., 0.);
    }
#undef WB_OFF
}
