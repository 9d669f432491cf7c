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
// Shared memory: two lane-minor arrays per warp (element i of lane's matrix at [32 i + lane]: conflict-free whatever
// element the lanes touch) -- the QL working copy of (d, e), then the pristine (d, e) for the factorisations.
template <int NW>
__host__ __device__ constexpr int wb_trideig_smem_doubles_per_warp() { return 2 * NW * 32; }

#define WB_TF_RESTOL 64.      // residual bound of a vector, in units of eps |T|
#define WB_TF_CTOL 0.03       // Gram-Schmidt window, in units of |T|
#define WB_TF_ACCEPT 0.5      // component outside the earlier vectors below which the other twists are tried
#define WB_TF_ACCEPT2 0.05    // ... below which the matrix goes to the Jacobi list

// Z is stored in pairs of components, lane-minor per group of 32 matrices:
//     Z[(((g NW + j) NP + i/2) 32 + lane) 2 + (i & 1)],   matrix = 32 g + lane from the start of the chunk, NP = ceil(NW/2),
// eigenvector j (ascending), component i.  The thread-per-matrix kernel moves a vector with fully coalesced 16-byte
// accesses (512 contiguous bytes per warp instruction); the back-transformation reads 16-byte pieces (half sectors).
__host__ __device__ __forceinline__ size_t wb_tf_zsize(long nk, int nw) {
    return (size_t)((nk + 31) / 32) * nw * ((nw + 1) / 2) * 64;
}
template <int NW>
__device__ __forceinline__ const double2* wb_tf_zvec(const double* Z, long t, int j) {
    constexpr int NP = (NW + 1) / 2;
    return reinterpret_cast<const double2*>(Z) + (((size_t)(t >> 5) * NW + j) * NP) * 32 + (t & 31);
}
template <int NW>
__device__ __forceinline__ void wb_tf_load_vec(const double2* __restrict__ zg, double (&v)[NW]) {
#pragma unroll
    for (int i = 0; i < NW; i += 2) {
        const double2 q = zg[(size_t)(i / 2) * 32];
        v[i] = q.x;
        if (i + 1 < NW) v[i + 1] = q.y;
    }
}
template <int NW>
__device__ __forceinline__ void wb_tf_store_vec(double2* __restrict__ zg, const double (&v)[NW]) {
#pragma unroll
    for (int i = 0; i < NW; i += 2) zg[(size_t)(i / 2) * 32] = make_double2(v[i], (i + 1 < NW) ? v[i + 1] : 0.);
}

// One twisted factorisation of T - sigma (d, e in shared memory, lane-minor) and the vector of the twist index r =
// argmin |gamma_i| among the indices that come AFTER (glast, rlast) in the order of increasing (|gamma|, index)
// (glast < 0: the overall minimum).  Returns r (-1: none left), |gamma_r| in *gabs, |z|^2 in *n2, z (z_r = 1) in A.
template <int NW>
__device__ __forceinline__ int wb_tf_attempt(const double* __restrict__ dp, const double* __restrict__ ep, double sigma,
                                             double pivmin, double glast, int rlast, double (&A)[NW], double (&B)[NW],
                                             double* gabs, double* n2) {
    // top-down: A[i] = l_i = e_i / D+_i, A[NW-1] = D+_{NW-1}
    {
        double dcur = dp[0] - sigma;
#pragma unroll
        for (int i = 0; i < NW - 1; i++) {
            if (fabs(dcur) < pivmin) dcur = -pivmin;
            const double ei = ep[i * 32];
            const double l = ei * __drcp_rn(dcur);
            A[i] = l;
            dcur = fma(-l, ei, dp[(i + 1) * 32] - sigma);
        }
        A[NW - 1] = dcur;
    }
    // bottom-up: B[i] = u_i = e_i / D-_{i+1};  gamma_i = D-_i - l_{i-1} e_{i-1}  (gamma_{NW-1} = D+_{NW-1})
    int r = -1;
    double gr = CUDART_INF;
    {
        const double g = fabs(A[NW - 1]);
        if ((g > glast) || (g == glast && NW - 1 > rlast)) { gr = g; r = NW - 1; }
        double dcur = dp[(NW - 1) * 32] - sigma;
#pragma unroll
        for (int i = NW - 2; i >= 0; i--) {
            if (fabs(dcur) < pivmin) dcur = -pivmin;
            const double ei = ep[i * 32];
            const double u = ei * __drcp_rn(dcur);
            B[i] = u;
            dcur = fma(-u, ei, dp[i * 32] - sigma);
            const double ga = fabs((i > 0) ? fma(-A[i - 1], ep[(i - 1) * 32], dcur) : dcur);
            const bool after = (ga > glast) || (ga == glast && i > rlast);
            if (after && ga <= gr) { gr = ga; r = i; }   // (<=: ties resolve to the lower index, which comes first)
        }
    }
    *gabs = gr;
    if (r < 0) return r;
    // the vector, in place: z_i (i < r) over l_i, z_{i+1} (i >= r) over u_i
    {
        double zn = 1.;
#pragma unroll
        for (int i = NW - 2; i >= 0; i--)
            if (i < r) { zn = -A[i] * zn; A[i] = zn; }
        zn = 1.;
#pragma unroll
        for (int i = 0; i < NW - 1; i++)
            if (i >= r) { zn = -B[i] * zn; B[i] = zn; }
#pragma unroll
        for (int i = NW - 1; i >= 1; i--) A[i] = (i < r) ? A[i] : ((i == r) ? 1. : B[i - 1]);
        if (r == 0) A[0] = 1.;
    }
    double s0 = 0., s1 = 0.;
#pragma unroll
    for (int i = 0; i + 1 < NW; i += 2) {
        s0 = fma(A[i], A[i], s0);
        s1 = fma(A[i + 1], A[i + 1], s1);
    }
    if (NW & 1) s0 = fma(A[NW - 1], A[NW - 1], s0);
    *n2 = s0 + s1;
    return r;
}

// VEC = false: eigenvalues only (Zout unused)
template <int NW, int NT, int MINB, bool VEC>
__global__ void __launch_bounds__(NT, MINB)
wb_trideig_kernel(long k0, long nk, const double* __restrict__ din, const double* __restrict__ ein, double* __restrict__ Eout,
                  double* __restrict__ Zout, int* __restrict__ fail_list, int* __restrict__ nfail) {
    static_assert(NW >= 2 && NW <= 31, "the split mask of the QL phase is one 32-bit word");
    extern __shared__ double smem_tf[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* const dw = smem_tf + (size_t)warp * wb_trideig_smem_doubles_per_warp<NW>() + lane;
    double* const ew = dw + NW * 32;
    long t = (long)blockIdx.x * NT + threadIdx.x;
    const bool live = t < nk;
    if ((long)blockIdx.x * NT + (threadIdx.x & ~31) >= nk) return;   // a warp without matrices (no CTA-wide barriers here)
    if (!live) t = nk - 1;                          // idle lanes mirror the last matrix through the QL phase
    const double* const dg = din + t * NW;
    const double* const eg = ein + t * NW;
    double tnorm = 0.;
    {
        double eprev = 0.;
#pragma unroll
        for (int i = 0; i < NW; i++) {
            const double di = dg[i], ei = (i < NW - 1) ? eg[i] : 0.;
            dw[i * 32] = di;
            ew[i * 32] = ei;
            tnorm = fmax(tnorm, fabs(di) + fabs(ei) + fabs(eprev));
            eprev = ei;
        }
    }
    // ---- implicit QL with Wilkinson shift, eigenvalues only (the recurrences of EISPACK tql1), organised in ROUNDS: in
    // a round every lane runs ONE sweep of its own matrix at its own deflation level l, the warp iterates over the
    // longest sweep.  A lane's split points (negligible e_i) live in a bit mask that the sweep itself updates -- the
    // rotation that produces e_{i+1} tests it -- so there is no scan of the off-diagonal between sweeps.
    bool fail = false;
    {
        unsigned mask = 1u << (NW - 1);             // bit i: e_i is negligible (bit NW-1: sentinel)
#pragma unroll
        for (int i = 0; i < NW - 1; i++) {
            const double dd = fabs(dw[i * 32]) + fabs(dw[(i + 1) * 32]);
            if (fabs(ew[i * 32]) + dd == dd) mask |= 1u << i;
        }
        int l = 0, iter = 0;
        while (true) {
            // deflate: d_l is an eigenvalue while e_l is negligible
            {
                const int adv = __ffs(~(mask >> l)) - 1;   // number of consecutive set bits from bit l
                if (adv > 0) { l += adv; iter = 0; }
            }
            bool active = (l < NW - 1) && !fail;
            if (active && ++iter > 40) { fail = true; active = false; }
            if (!__any_sync(0xffffffffu, active)) break;
            int m = l;
            double s = 1., c = 1., p = 0., g = 0., dnext = 0., dd2 = 0.;
            if (active) {
                m = l + __ffs(mask >> (l + 1));     // first split above l (the sentinel bounds it)
                const double el = ew[l * 32], dl = dw[l * 32];
                g = (dw[(l + 1) * 32] - dl) * (0.5 * __drcp_rn(el));
                const double r = sqrt(fma(g, g, 1.));
                dnext = dw[m * 32];                  // d_{i+1} of the first rotation
                g = dnext - dl + el * __drcp_rn(g + copysign(r, g));
                mask &= ~(((1u << m) - 1u) & ~((1u << l) - 1u));    // bits l .. m-1 are decided again by this sweep
            }
            const int len = __reduce_max_sync(0xffffffffu, active ? (m - l) : 0);
            int i = m - 1;
            for (int q = 0; q < len; q++, i--) {
                if (active && i >= l) {
                    const double ei = ew[i * 32], di = dw[i * 32];
                    const double f = s * ei, b = c * ei;
                    const double h2 = fma(f, f, g * g);
                    double r, rinv;
                    if (h2 > 1e-280 && h2 < 1e280) {
                        rinv = rsqrt(h2);
                        r = h2 * rinv;
                    } else {                         // out of the safe range of the plain formula
                        r = hypot(f, g);
                        rinv = (r != 0.) ? 1. / r : 0.;
                    }
                    ew[(i + 1) * 32] = r;
                    if (r == 0.) {                   // underflow: the sweep ends here (tql1)
                        dw[(i + 1) * 32] = dnext - p;
                        ew[m * 32] = 0.;
                        mask |= 1u << (i + 1);
                        active = false;              // (the next round starts from the same l)
                    } else {
                        s = f * rinv;
                        c = g * rinv;
                        g = dnext - p;
                        const double r2 = (di - g) * s + 2. * c * b;
                        p = s * r2;
                        const double dnew = g + p;
                        dw[(i + 1) * 32] = dnew;
                        g = c * r2 - b;
                        // e_{i+1} = r against the final d_{i+1}, d_{i+2} (i + 1 < m: e_m is set to zero below)
                        if (i + 1 < m) {
                            const double dd = fabs(dnew) + dd2;
                            if (r + dd == dd) mask |= 1u << (i + 1);
                        }
                        dd2 = fabs(dnew);
                        dnext = di;
                        if (i == l) {                // end of the sweep
                            const double dl = di - p;
                            dw[l * 32] = dl;
                            ew[l * 32] = g;
                            ew[m * 32] = 0.;
                            const double dd = fabs(dl) + dd2;
                            if (fabs(g) + dd == dd) mask |= 1u << l;
                        }
                    }
                }
            }
        }
    }
    // ---- sort ascending (odd-even transposition network on registers), write E
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
            if (live) Eout[(k0 + t) * NW + i] = w[i];
        }
    }
    if (!live) return;
    if (!VEC || fail) {
        if (fail) fail_list[atomicAdd(nfail, 1)] = (int)t;
        return;
    }
    // ---- eigenvectors.  The pristine (d, e) replace the QL working copy (this lane's slots only: no barrier).
    double* const dp = dw;
    double* const ep = ew;
#pragma unroll
    for (int i = 0; i < NW; i++) {
        dp[i * 32] = dg[i];
        ep[i * 32] = (i < NW - 1) ? eg[i] : 0.;
    }
    const double eps = 2.220446049250313e-16;
    const double pivmin = fmax(eps * tnorm * 0.0009765625, 1e-290);
    const double ctol = WB_TF_CTOL * tnorm;
    const double restol = WB_TF_RESTOL * eps * tnorm;
    const double* const Eme = Eout + (k0 + t) * NW;   // (this thread's own stores above)
    // One loop iteration = ONE factorisation (a single inlined copy of it keeps the register allocation of the common
    // path free of the rare one).  mode 0: the twist of smallest |gamma| for eigenvalue j; accepted unless the
    // projection cancels it.  mode 1 (numerically multiple eigenvalue): the vectors of the other twists span the rest
    // of the eigenspace -- try them by increasing |gamma| (decreasing weight in the eigenspace), remember the one with
    // the largest component outside the earlier vectors.  mode 2: rebuild that one and store it.
    int j = 0, mode = 0, rlast = -1, rbest = -1, ntrial = 0;
    double glast = -1., best = 0., gbest = -1.;
#pragma unroll 1
    while (j < NW) {
        const double sigma = Eme[j];
        double A[NW], B[NW];
        double gabs, n2;
        const int r = wb_tf_attempt<NW>(dp, ep, sigma, pivmin, glast, rlast, A, B, &gabs, &n2);
        bool final = (mode == 2);
        double keep = 1., res = 0.;
        int p0 = j;
        auto project = [&]() {                       // A <- A minus its components along the vectors p0 .. j-1; |A|^2
            for (int p = j - 1; p >= p0; p--) {
                wb_tf_load_vec<NW>(wb_tf_zvec<NW>(Zout, t, p), B);
                double dot = 0.;
#pragma unroll
                for (int i = 0; i < NW; i++) dot = fma(B[i], A[i], dot);
#pragma unroll
                for (int i = 0; i < NW; i++) A[i] = fma(-dot, B[i], A[i]);
            }
            double sq = 0.;
#pragma unroll
            for (int i = 0; i < NW; i++) sq = fma(A[i], A[i], sq);
            return sq;
        };
        if (r >= 0) {
            const double scale = rsqrt(n2);
            res = gabs * scale;                      // |(T - sigma) z| / |z|
#pragma unroll
            for (int i = 0; i < NW; i++) A[i] *= scale;
            // Gram-Schmidt window: earlier vectors with eigenvalues within ctol
            while (p0 > 0 && sigma - Eme[p0 - 1] < ctol) p0--;
            if (p0 < j) {
                keep = sqrt(project());
                if (!(keep == keep)) keep = 0.;
            }
            if (mode == 0 && keep >= WB_TF_ACCEPT) final = true;
        }
        if (final) {
            if (keep < 1.) {
                const double s1 = (keep > 0.) ? 1. / keep : 0.;
#pragma unroll
                for (int i = 0; i < NW; i++) A[i] *= s1;
                res *= s1;
                if (keep < WB_TF_ACCEPT && keep > 0.) {   // second pass ("twice is enough") after a large cancellation
                    const double s2 = rsqrt(project());
#pragma unroll
                    for (int i = 0; i < NW; i++) A[i] *= s2;
                }
            }
            // |gamma_r| / |z| is the residual only as far as the pivots of T - sigma carry no element growth, and the
            // projection adds |dot| |sigma_p - sigma| per earlier vector: measure |(T - sigma) z| of the vector that is
            // stored (z normalised; exactly degenerate, coupled spectra -- Kramers pairs -- are where the two differ)
            double r2 = 0., below = 0.;
#pragma unroll
            for (int i = 0; i < NW; i++) {
                const double ei = ep[i * 32];
                double v = fma(dp[i * 32] - sigma, A[i], below);
                if (i + 1 < NW) v = fma(ei, A[i + 1], v);
                below = ei * A[i];
                r2 = fma(v, v, r2);
            }
            if (r < 0 || !(keep >= WB_TF_ACCEPT2) || !(res <= restol) || !(r2 <= restol * restol)) fail = true;
            wb_tf_store_vec<NW>(const_cast<double2*>(wb_tf_zvec<NW>(Zout, t, j)), A);
            j++;
            mode = 0; glast = -1.; rlast = -1; best = 0.; rbest = -1; gbest = -1.; ntrial = 0;
        } else {
            bool exhausted = (r < 0);
            if (!exhausted) {
                if (res <= restol) {
                    if (keep > best) { best = keep; gbest = gabs; rbest = r; }
                } else if (mode == 1) exhausted = true;     // |gamma| only grows from here
                glast = gabs;
                rlast = r;
                mode = 1;
                if (++ntrial >= NW) exhausted = true;
            }
            if (exhausted || best >= WB_TF_ACCEPT) {
                // the attempt that follows (gbest, rbest - 1) in the order is (gbest, rbest); none found: the overall minimum
                mode = 2;
                glast = gbest;
                rlast = rbest - 1;
                if (rbest < 0) { glast = -1.; rlast = -1; fail = true; }
            }
        }
    }
    if (fail) fail_list[atomicAdd(nfail, 1)] = (int)t;
}

// ------------------------------------------------------------------------------------------ K3
// U = H(0) H(1) ... H(NW-2) Z  for MB matrices per CTA; V (Householder vectors below the sub-diagonal, column k = vector k,
// LAPACK zhetd2 UPLO = 'L' as written by wb_tridiag*_kernel) and U share the buffer VU: the CTA stages V in shared memory
// before it writes U.  Z real in the paired lane-minor layout above, eigenvector j ascending; U[matrix][i][j].
template <int NW, int MB>
__host__ __device__ constexpr int wb_backtransform_smem_bytes() { return MB * (NW * NW + 1) * 16 + MB * NW * 16 + 16; }

template <int NW, int MB, int MINB>
__global__ void __launch_bounds__((MB * NW + 31) / 32 * 32, MINB)
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
        double zr[NW];
        wb_tf_load_vec<NW>(wb_tf_zvec<NW>(Z, m0 + m, j), zr);
#pragma unroll
        for (int i = 0; i < NW; i++) u[i] = cmake(zr[i], 0.);
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
    if (live) {   // (in place: V of this CTA's own matrices was staged in shared memory before anything is written)
        cplx* Uo = VU + (k0 + m0 + m) * NW * NW;
#pragma unroll
        for (int i = 0; i < NW; i++) Uo[i * NW + j] = u[i];
    }
}
