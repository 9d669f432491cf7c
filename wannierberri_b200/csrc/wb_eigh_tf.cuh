// Eigenvalues and eigenvectors of the REAL symmetric tridiagonal matrices left by the Householder reduction, and the
// back-transformation of the eigenvectors -- the second and third thirds of
//     E_K, UU_K = np.linalg.eigh(HH_K)        (data_K/data_K.py:211-218)
// for num_wann <= 24, replacing the rotation-stream pair wb_tql_kernel / wb_eigvec_kernel of wb_eigh_ql.cuh.
//
// Why: accumulating the QL rotations into the eigenvector matrix is an O(n^3) job that only a warp per matrix can do
// (lane = row, 18 of 32 lanes busy, 6.5 KB of rotations per matrix through HBM).  Here the eigenvectors of T come from
// ONE twisted factorisation per eigenvalue (Fernando / Parlett-Dhillon):  T - sigma = N_r D_r N_r^T with the twist at
// the index r of smallest |gamma_r|, whose solution of  (T - sigma) z = gamma_r e_r  is the eigenvector -- O(n) per
// vector, no iteration, everything of a matrix in the registers of ONE THREAD, all 32 lanes busy.
//
//   K2  wb_trideig_kernel   thread per matrix: implicit QL for the eigenvalues (no rotations kept), sorting network,
//                           then per eigenvalue: top-down and bottom-up pivots D+, D-, gamma, the vector, modified
//                           Gram-Schmidt against the earlier vectors whose eigenvalues lie within ctol |T| (inverse
//                           iteration loses orthogonality as eps |T| / gap).  A vector that the Gram-Schmidt step
//                           cancels (numerically multiple eigenvalue: the best twist reproduces an earlier vector) is
//                           rebuilt from the other twist indices -- the factors do not depend on r -- keeping the one
//                           with the largest component outside the span of the earlier vectors.  Every vector carries
//                           its residual |gamma_r| / |z|; a matrix with a vector above 64 eps |T|, a cancelled vector
//                           or an unconverged QL is put on the list of the Jacobi kernel (wb_eigh_jacobi.cuh).
//   K3  wb_backtransform_kernel   item = (matrix, eigenvector): lane j applies the Householder reflectors of ITS matrix
//                           (shared memory, staged by TMA bulk copies; the 2-3 matrices of a warp sit in different
//                           banks, so the reads are broadcasts) to ITS real vector: all lanes busy.
#pragma once
#include "wb_common.cuh"
#include "wb_tma.cuh"

// ------------------------------------------------------------------------------------------ K2
template <int NW>
__host__ __device__ constexpr int wb_trideig_smem_doubles_per_warp() { return 3 * NW * 32 + 32 * (NW + 1); }

#define WB_TF_RESTOL 64.      // residual bound of a vector, in units of eps |T|
#define WB_TF_CTOL 0.03       // Gram-Schmidt window, in units of |T|
#define WB_TF_ACCEPT 0.5      // component outside the earlier vectors below which the other twists are tried
#define WB_TF_ACCEPT2 0.05    // ... below which the matrix goes to the Jacobi list

// the vector of the twist index r from the multipliers of the two factorisations (z_r = 1); returns |z|^2
template <int NW>
__device__ __forceinline__ double wb_tf_vector(const double (&lp)[NW - 1], const double (&um)[NW - 1], int r, double (&z)[NW]) {
#pragma unroll
    for (int i = 0; i < NW; i++) z[i] = (i == r) ? 1. : 0.;
#pragma unroll
    for (int i = NW - 2; i >= 0; i--)
        if (i < r) z[i] = -lp[i] * z[i + 1];
#pragma unroll
    for (int i = 0; i < NW - 1; i++)
        if (i >= r) z[i + 1] = -um[i] * z[i];
    double s0 = 0., s1 = 0.;
#pragma unroll
    for (int i = 0; i + 1 < NW; i += 2) {
        s0 = fma(z[i], z[i], s0);
        s1 = fma(z[i + 1], z[i + 1], s1);
    }
    if (NW & 1) s0 = fma(z[NW - 1], z[NW - 1], s0);
    return s0 + s1;
}

// VEC = false: eigenvalues only (Zout unused)
template <int NW, int NT, bool VEC>
__global__ void __launch_bounds__(NT)
wb_trideig_kernel(long k0, long nk, const double* __restrict__ din, const double* __restrict__ ein, double* __restrict__ Eout,
                  double* __restrict__ Zout, int* __restrict__ fail_list, int* __restrict__ nfail) {
    extern __shared__ double smem_tf[];
    constexpr int PW = wb_trideig_smem_doubles_per_warp<NW>();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* const base = smem_tf + (size_t)warp * PW;
    double* const dp = base + lane;                 // pristine diagonal, element i at dp[32 i]
    double* const ep = base + NW * 32 + lane;       // pristine off-diagonal
    double* const dw = base + 2 * NW * 32 + lane;   // QL working diagonal -> sorted eigenvalues
    double* const tile = base + 3 * NW * 32;        // [32][NW + 1]: QL working off-diagonal (as ew[32 i]), then the
    double* const ew = tile + lane;                 //               transposition buffer of the vector stores
    long t = (long)blockIdx.x * NT + threadIdx.x;
    const bool live = t < nk;
    if ((long)blockIdx.x * NT + (threadIdx.x & ~31) >= nk) return;   // a warp without matrices (no CTA-wide barriers here)
    if (!live) t = nk - 1;                          // idle lanes mirror the last matrix and write nothing
    double tnorm = 0.;
    {
        const double* dg = din + t * NW;
        const double* eg = ein + t * NW;
        double eprev = 0.;
#pragma unroll
        for (int i = 0; i < NW; i++) {
            const double di = dg[i], ei = (i < NW - 1) ? eg[i] : 0.;
            dp[i * 32] = di;
            ep[i * 32] = ei;
            dw[i * 32] = di;
            ew[i * 32] = ei;
            tnorm = fmax(tnorm, fabs(di) + fabs(ei) + fabs(eprev));
            eprev = ei;
        }
    }
    // ---- implicit QL, eigenvalues only (EISPACK tql1 / "tqli" recurrences; the rotations are not kept)
    bool fail = false;
    for (int l = 0; l < NW && !fail; l++) {
        int iter = 0;
        while (true) {
            int m = l;
            for (; m < NW - 1; m++) {
                const double dd = fabs(dw[m * 32]) + fabs(dw[(m + 1) * 32]);
                if (fabs(ew[m * 32]) + dd == dd) break;
            }
            if (m == l) break;
            if (++iter > 40) { fail = true; break; }
            const double el = ew[l * 32];
            double g = (dw[(l + 1) * 32] - dw[l * 32]) * (0.5 * __drcp_rn(el));
            double r = sqrt(fma(g, g, 1.));
            g = dw[m * 32] - dw[l * 32] + el * __drcp_rn(g + copysign(r, g));
            double s = 1., c = 1., p = 0.;
            int i = m - 1;
            for (; i >= l; i--) {
                const double ei = ew[i * 32];
                const double f = s * ei, b = c * ei;
                const double h2 = fma(f, f, g * g);
                double rinv;
                if (h2 > 1e-280 && h2 < 1e280) {
                    rinv = rsqrt(h2);
                    r = h2 * rinv;
                } else {                                 // out of the safe range of the plain formula
                    r = hypot(f, g);
                    rinv = (r != 0.) ? 1. / r : 0.;
                }
                ew[(i + 1) * 32] = r;
                if (r == 0.) {                           // recover from underflow
                    dw[(i + 1) * 32] -= p;
                    ew[m * 32] = 0.;
                    break;
                }
                s = f * rinv;
                c = g * rinv;
                g = dw[(i + 1) * 32] - p;
                r = (dw[i * 32] - g) * s + 2. * c * b;
                p = s * r;
                dw[(i + 1) * 32] = g + p;
                g = c * r - b;
            }
            if (i >= l) continue;
            dw[l * 32] -= p;
            ew[l * 32] = g;
            ew[m * 32] = 0.;
        }
    }
    // ---- sort ascending (odd-even transposition network on registers)
    {
        double w[NW];
#pragma unroll
        for (int i = 0; i < NW; i++) w[i] = dw[i * 32];
#pragma unroll
        for (int round = 0; round < NW; round++) {
#pragma unroll
            for (int i = (round & 1); i + 1 < NW; i += 2) {
                const double lo = fmin(w[i], w[i + 1]), hi = fmax(w[i], w[i + 1]);
                w[i] = lo;
                w[i + 1] = hi;
            }
        }
#pragma unroll
        for (int i = 0; i < NW; i++) {
            if (!(w[i] == w[i])) fail = true;        // NaN input
            dw[i * 32] = w[i];
            if (live) Eout[(k0 + t) * NW + i] = w[i];
        }
    }
    if (!VEC) {
        if (fail && live) fail_list[atomicAdd(nfail, 1)] = (int)t;
        return;
    }
    __syncwarp();   // the working off-diagonal is dead: its storage becomes the transposition buffer
    const double eps = 2.220446049250313e-16;
    const double pivmin = fmax(eps * tnorm * 0.0009765625, 1e-290);
    const double ctol = WB_TF_CTOL * tnorm;
    const double restol = WB_TF_RESTOL * eps * tnorm;
    double* const myrow = tile + lane * (NW + 1);
    const long blk0 = (long)blockIdx.x * NT + (threadIdx.x & ~31);   // first matrix of this warp (chunk-relative)
    double* const Zme = Zout + t * NW * NW;
#pragma unroll 1
    for (int j = 0; j < NW; j++) {
        const double sigma = dw[j * 32];
        double gam[NW], lp[NW - 1], um[NW - 1], z[NW];
        // top-down pivots D+ (kept in gam until gamma is formed) and multipliers l+
        {
            double dcur = dp[0] - sigma;
#pragma unroll
            for (int i = 0; i < NW - 1; i++) {
                if (fabs(dcur) < pivmin) dcur = -pivmin;
                gam[i] = dcur;
                const double ei = ep[i * 32];
                const double l = ei * __drcp_rn(dcur);
                lp[i] = l;
                dcur = fma(-l, ei, dp[(i + 1) * 32] - sigma);
            }
            gam[NW - 1] = dcur;
        }
        // bottom-up pivots D-, multipliers u-, gamma_i = D+_i + D-_i - (d_i - sigma); r = argmin |gamma|
        int r = NW - 1;
        double gr = gam[NW - 1];
        {
            double dcur = dp[(NW - 1) * 32] - sigma;
#pragma unroll
            for (int i = NW - 2; i >= 0; i--) {
                if (fabs(dcur) < pivmin) dcur = -pivmin;
                const double ei = ep[i * 32];
                const double u = ei * __drcp_rn(dcur);
                um[i] = u;
                const double di = dp[i * 32] - sigma;
                dcur = fma(-u, ei, di);
                const double g = gam[i] + (dcur - di);
                gam[i] = g;
                if (fabs(g) < fabs(gr)) { gr = g; r = i; }
            }
        }
        double n2 = wb_tf_vector<NW>(lp, um, r, z);
        double res = fabs(gr) * rsqrt(n2);           // |(T - sigma) z| / |z|
        double scale = rsqrt(n2);
#pragma unroll
        for (int i = 0; i < NW; i++) z[i] *= scale;
        // Gram-Schmidt window: earlier vectors with eigenvalues within ctol
        int p0 = j;
        while (p0 > 0 && sigma - dw[(p0 - 1) * 32] < ctol) p0--;
        double keep = 1.;                            // norm left after the projection
        if (p0 < j) {
            // one projection pass over the window; the latest vector is still in this lane's row of the tile
            auto project = [&](double(&v)[NW]) {
                for (int p = j - 1; p >= p0; p--) {
                    double dot = 0.;
                    if (p == j - 1) {
#pragma unroll
                        for (int i = 0; i < NW; i++) dot = fma(myrow[i], v[i], dot);
#pragma unroll
                        for (int i = 0; i < NW; i++) v[i] = fma(-dot, myrow[i], v[i]);
                    } else {
                        const double* zg = Zme + (size_t)p * NW;
#pragma unroll
                        for (int i = 0; i < NW; i++) dot = fma(__ldcg(zg + i), v[i], dot);
#pragma unroll
                        for (int i = 0; i < NW; i++) v[i] = fma(-dot, __ldcg(zg + i), v[i]);
                    }
                }
                double s = 0.;
#pragma unroll
                for (int i = 0; i < NW; i++) s = fma(v[i], v[i], s);
                return s;
            };
            double k2 = project(z);
            keep = sqrt(k2);
            if (!(keep >= WB_TF_ACCEPT)) {
                // numerically multiple eigenvalue: the vectors of the other twists span the rest of the eigenspace.
                // Try them by increasing |gamma| (decreasing weight in the eigenspace); keep the best.
                double best = (keep == keep) ? keep : 0., bres = res;
                int rbest = r;
                double glast = fabs(gr);
                int rlast = r;
                for (int trial = 1; trial < NW && best < WB_TF_ACCEPT; trial++) {
                    // next twist: smallest |gamma| above the last one (ties by index)
                    int rn = -1;
                    double gn = CUDART_INF;
#pragma unroll
                    for (int i = 0; i < NW; i++) {
                        const double a = fabs(gam[i]);
                        const bool after = (a > glast) || (a == glast && i > rlast);
                        if (after && (a < gn)) { gn = a; rn = i; }
                    }
                    if (rn < 0) break;
                    glast = gn;
                    rlast = rn;
                    const double m2 = wb_tf_vector<NW>(lp, um, rn, z);   // (z is rebuilt from the best twist below)
                    const double rc = gn * rsqrt(m2);
                    if (!(rc <= restol)) continue;
                    const double sc = rsqrt(m2);
#pragma unroll
                    for (int i = 0; i < NW; i++) z[i] *= sc;
                    const double kc = sqrt(project(z));
                    if (kc > best) { best = kc; bres = rc; rbest = rn; }
                }
                {
                    n2 = wb_tf_vector<NW>(lp, um, rbest, z);
                    scale = rsqrt(n2);
#pragma unroll
                    for (int i = 0; i < NW; i++) z[i] *= scale;
                    k2 = project(z);
                    res = bres;
                }
                keep = sqrt(k2);
                if (!(keep == keep)) keep = 0.;
                if (keep > 0.) {                         // second pass ("twice is enough") after a large cancellation
                    const double s1 = 1. / keep;
#pragma unroll
                    for (int i = 0; i < NW; i++) z[i] *= s1;
                    const double k3 = project(z);
                    const double s2 = rsqrt(k3);
#pragma unroll
                    for (int i = 0; i < NW; i++) z[i] *= s2;
                    res *= s1;
                }
            } else {
                const double s1 = rsqrt(k2);
#pragma unroll
                for (int i = 0; i < NW; i++) z[i] *= s1;
                res *= s1;
            }
        }
        if (!(keep >= WB_TF_ACCEPT2) || !(res <= restol)) fail = true;
        // ---- store vector j of the 32 matrices of the warp: Z[matrix][j][0..NW)
        __syncwarp();
#pragma unroll
        for (int i = 0; i < NW; i++) myrow[i] = z[i];
        __syncwarp();
#pragma unroll
        for (int q = 0; q < NW; q++) {
            const int idx = q * 32 + lane;
            const int mm = idx / NW, ii = idx - mm * NW;
            if (blk0 + mm < nk) Zout[((blk0 + mm) * NW + j) * NW + ii] = tile[mm * (NW + 1) + ii];
        }
    }
    if (fail && live) fail_list[atomicAdd(nfail, 1)] = (int)t;
}

// ------------------------------------------------------------------------------------------ K3
// U = H(0) H(1) ... H(NW-2) Z  for MB matrices per CTA; V (Householder vectors below the sub-diagonal, column k = vector k,
// LAPACK zhetd2 UPLO = 'L' as written by wb_tridiag*_kernel) and U share the buffer VU: the CTA stages V in shared memory
// before it writes U.  Z[matrix][j][i] real, eigenvector j ascending; U[matrix][i][j].
template <int NW, int MB>
__host__ __device__ constexpr int wb_backtransform_smem_bytes() { return MB * (NW * NW + 1) * 16 + MB * NW * 16 + 16; }

template <int NW, int MB>
__global__ void __launch_bounds__((MB * NW + 31) / 32 * 32)
wb_backtransform_kernel(long k0, long nk, const double* __restrict__ Z, const cplx* __restrict__ tauin, cplx* __restrict__ VU) {
    extern __shared__ __align__(16) cplx smem_bt[];
    constexpr int VS = NW * NW + 1;                    // stride of a matrix (16-byte units): 2-3 neighbours in distinct banks
    cplx* const Vs = smem_bt;
    cplx* const taus = smem_bt + MB * VS;
    uint64_t* const bar = (uint64_t*)(taus + MB * NW);
    const long m0 = (long)blockIdx.x * MB;             // first matrix of the CTA (chunk-relative)
    const int nm = (int)((nk - m0 < MB) ? (nk - m0) : MB);
    if (threadIdx.x == 0) {
        wb_mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        wb_mbar_expect_tx(bar, (uint32_t)(nm * (NW * NW * 16 + NW * 16)));
        for (int m = 0; m < nm; m++) {
            wb_bulk_g2s(Vs + m * VS, VU + (k0 + m0 + m) * NW * NW, NW * NW * 16, bar);
            wb_bulk_g2s(taus + m * NW, tauin + (m0 + m) * NW, NW * 16, bar);
        }
    }
    const int item = threadIdx.x;
    const int m = item / NW, j = item - m * NW;
    const bool live = (m < nm);
    cplx u[NW];
    if (live) {
        const double* zg = Z + ((m0 + m) * NW + j) * NW;
#pragma unroll
        for (int i = 0; i < NW; i++) u[i] = cmake(zg[i], 0.);
    } else {
#pragma unroll
        for (int i = 0; i < NW; i++) u[i] = cmake(0., 0.);
    }
    __syncthreads();
    wb_mbar_wait(bar, 0);
    const cplx* const V = Vs + (live ? m : 0) * VS;
    const cplx* const tm = taus + (live ? m : 0) * NW;
#pragma unroll
    for (int kk = 0; kk < NW - 1; kk++) {
        const int k = NW - 2 - kk;
        const cplx tau = tm[k];
        cplx sdot = u[k + 1], sdot1 = cmake(0., 0.);   // v[k+1] = 1
#pragma unroll
        for (int i = k + 2; i < NW; i++) {
            if ((i - k) & 1) cfma_conj(sdot1, V[i * NW + k], u[i]);
            else cfma_conj(sdot, V[i * NW + k], u[i]);
        }
        const cplx ts = cmul(tau, cadd(sdot, sdot1));   // tau = 0: identity
        u[k + 1] = csub(u[k + 1], ts);
#pragma unroll
        for (int i = k + 2; i < NW; i++) {
            const cplx vv = V[i * NW + k];
            u[i].x = fma(-ts.x, vv.x, u[i].x);
            u[i].x = fma(ts.y, vv.y, u[i].x);
            u[i].y = fma(-ts.x, vv.y, u[i].y);
            u[i].y = fma(-ts.y, vv.x, u[i].y);
        }
    }
    __syncthreads();   // (every thread is past its last read of the staged V of this CTA's own matrices)
    if (live) {
        cplx* Uo = VU + (k0 + m0 + m) * NW * NW;
#pragma unroll
        for (int i = 0; i < NW; i++) Uo[i * NW + j] = u[i];
    }
}
