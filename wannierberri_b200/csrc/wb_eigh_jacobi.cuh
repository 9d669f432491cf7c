// Batched complex Hermitian eigensolver, generic path: parallel-order (round-robin) two-sided
// Jacobi, one warp per k-point, matrix and eigenvectors resident in shared memory.
//
// Replaces  E_K, UU_K = np.linalg.eigh(HH_K)   (data_K/data_K.py:211-218, 309-322).
// Eigenvalues ascending, eigenvectors in columns, like LAPACK zheevd.  Works for any nw that
// fits shared memory (2 * nw * (nw+1) * 16 B per warp); it is the fallback / cross-check for the
// Householder+QL path of wb_eigh_ql.cuh.
#pragma once
#include "wb_common.cuh"

#define WB_JACOBI_MAX_SWEEPS 30

// round-robin tournament: n players (n even, possibly one dummy), step s in [0, n-1), pair m in [0, n/2)
__device__ __forceinline__ void rr_pair(int n, int s, int m, int& p, int& q) {
    // player n-1 fixed, the others rotate
    int a = (m == 0) ? (n - 1) : ((s + m) % (n - 1));
    int b = (s + (n - 1) - m) % (n - 1);
    p = min(a, b);
    q = max(a, b);
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
wb_eigh_jacobi_kernel(const cplx* __restrict__ rec, WbLayout L, long k0, long nk_in, double* __restrict__ Eout,
                      cplx* __restrict__ Uout, int* __restrict__ max_sweeps, const int* __restrict__ list,
                      const int* __restrict__ nlist) {
    extern __shared__ cplx smem_j[];
    const int nw = L.nw;
    const int ld = nw + 1;  // padded leading dimension
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int npl = nw + (nw & 1);  // players (even)
    const int npair = npl / 2;
    // per-warp shared storage: A[nw][ld], U[nw][ld], rot[npair] (c, s, e^{i phi}) , ev[nw]
    cplx* A = smem_j + (size_t)warp * (2 * nw * ld + 2 * npair + nw);
    cplx* U = A + nw * ld;
    cplx* rot_cs = U + nw * ld;        // (c, s)
    cplx* rot_ph = rot_cs + npair;     // e^{i phi}
    double* ev = (double*)(rot_ph + npair);

    // either all k-points k0 .. k0+nk_in-1, or the ones listed in list[0 .. *nlist-1] (relative to k0)
    const long nk = list ? (long)(*nlist) : nk_in;
    for (long idx = (long)blockIdx.x * WARPS + warp; idx < nk; idx += (long)gridDim.x * WARPS) {
        const long ik = k0 + (list ? (long)list[idx] : idx);
        const cplx* H = rec + ik * L.E + L.off_H;
        double nrm2 = 0.;
        for (int x = lane; x < nw * nw; x += 32) {
            int i = x / nw, j = x % nw;
            cplx v = (i <= j) ? H[tri_index(i, j, nw)] : cconj(H[tri_index(j, i, nw)]);
            if (i == j) v.y = 0.;
            A[i * ld + j] = v;
            U[i * ld + j] = cmake(i == j ? 1. : 0., 0.);
            nrm2 += v.x * v.x + v.y * v.y;
        }
        for (int o = 16; o > 0; o >>= 1) nrm2 += __shfl_xor_sync(0xffffffffu, nrm2, o);
        __syncwarp();
        // converged: off-diagonal norm below 2e-15 ||A||_F, or -- once below 3e-14 ||A||_F -- no longer shrinking (rounding
        // floor).  "One more sweep after 1e-10" is not enough for numerically multiple eigenvalues (Kramers pairs), where
        // the off-diagonal norm can idle for a sweep (1e-10 -> 4e-11 -> 2e-13 -> 1e-16 ||A||_F was seen): the residual of
        // such k-points was 1e-12 |H|.
        const double tol_conv = 1e-27 * nrm2, tol_done = 4e-30 * nrm2;
        int sweep = 0;
        bool last = false;
        double off_prev = nrm2;
        for (; sweep < WB_JACOBI_MAX_SWEEPS; sweep++) {
            for (int s = 0; s < npl - 1; s++) {
                // (a) rotation parameters of the npair disjoint pairs
                for (int m = lane; m < npair; m += 32) {
                    int p, q;
                    rr_pair(npl, s, m, p, q);
                    double c = 1., sn = 0.;
                    cplx ph = cmake(1., 0.);
                    if (q < nw) {
                        cplx bpq = A[p * ld + q];
                        double absb = hypot(bpq.x, bpq.y);
                        double app = A[p * ld + p].x, aqq = A[q * ld + q].x;
                        if (absb > 1e-300 && absb > 1e-18 * (fabs(app) + fabs(aqq))) {
                            double tau = (aqq - app) / (2. * absb);
                            double t = (tau >= 0. ? 1. : -1.) / (fabs(tau) + sqrt(1. + tau * tau));
                            c = 1. / sqrt(1. + t * t);
                            sn = t * c;
                            ph = cmake(bpq.x / absb, bpq.y / absb);
                        }
                    }
                    rot_cs[m] = cmake(c, sn);
                    rot_ph[m] = ph;
                }
                __syncwarp();
                // (b) columns:  A <- A J,  U <- U J      J = [[c, s e^{i phi}], [-s e^{-i phi}, c]]
                for (int x = lane; x < 2 * nw * npair; x += 32) {
                    int m = x % npair;
                    int i = (x / npair) % nw;
                    cplx* M = (x >= nw * npair) ? U : A;
                    int p, q;
                    rr_pair(npl, s, m, p, q);
                    if (q >= nw) continue;
                    double c = rot_cs[m].x, sn = rot_cs[m].y;
                    cplx sph = cscale(sn, rot_ph[m]);  // s e^{i phi}
                    cplx ap = M[i * ld + p], aq = M[i * ld + q];
                    // new p = c ap - s e^{-i phi} aq ; new q = s e^{i phi} ap + c aq
                    cplx np_ = csub(cscale(c, ap), cconjmul(sph, aq));
                    cplx nq_ = cadd(cmul(sph, ap), cscale(c, aq));
                    M[i * ld + p] = np_;
                    M[i * ld + q] = nq_;
                }
                __syncwarp();
                // (c) rows:  A <- J^dagger A
                for (int x = lane; x < nw * npair; x += 32) {
                    int m = x / nw;
                    int j = x % nw;
                    int p, q;
                    rr_pair(npl, s, m, p, q);
                    if (q >= nw) continue;
                    double c = rot_cs[m].x, sn = rot_cs[m].y;
                    cplx sph = cscale(sn, rot_ph[m]);
                    cplx ap = A[p * ld + j], aq = A[q * ld + j];
                    // new row p = c ap - s e^{i phi} aq ; new row q = s e^{-i phi} ap + c aq
                    cplx np_ = csub(cscale(c, ap), cmul(sph, aq));
                    cplx nq_ = cadd(cconjmul(sph, ap), cscale(c, aq));
                    A[p * ld + j] = np_;
                    A[q * ld + j] = nq_;
                }
                __syncwarp();
            }
            if (last) { sweep++; break; }
            double off = 0.;
            for (int x = lane; x < nw * nw; x += 32) {
                int i = x / nw, j = x % nw;
                if (i < j) { cplx v = A[i * ld + j]; off += v.x * v.x + v.y * v.y; }
            }
            for (int o = 16; o > 0; o >>= 1) off += __shfl_xor_sync(0xffffffffu, off, o);
            if (2. * off <= tol_done || (2. * off <= tol_conv && off > 0.25 * off_prev)) last = true;
            off_prev = off;
        }
        // sort ascending (rank by counting; ties broken by index) and write out
        for (int i = lane; i < nw; i += 32) ev[i] = A[i * ld + i].x;
        __syncwarp();
        for (int i = lane; i < nw; i += 32) {
            double e = ev[i];
            int rank = 0;
            for (int j = 0; j < nw; j++) rank += (ev[j] < e) || (ev[j] == e && j < i);
            Eout[ik * nw + rank] = e;
            ((int*)rot_cs)[i] = rank;
        }
        __syncwarp();
        if (Uout) {
            for (int x = lane; x < nw * nw; x += 32) {
                int i = x / nw, j = x % nw;
                Uout[(ik * nw + i) * nw + ((int*)rot_cs)[j]] = U[i * ld + j];
            }
        }
        if (lane == 0 && max_sweeps) atomicMax(max_sweeps, sweep);
        __syncwarp();
    }
}
