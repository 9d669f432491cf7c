// Batched complex Hermitian eigensolver for 32 < nw <= 128: the same zhetd2 + tql2 + zunm2l sequence as
// wb_eigh_ql.cuh, with ONE CTA PER K-POINT for the two matrix phases (the matrix lives in shared memory):
//
//   K1L  wb_tridiag_cta_kernel   packed lower triangle of H(k) in shared memory (132 KB at nw = 128); Hermitian matvec (2 / 4 /
//                                8 threads per ACTIVE row) + rank-2 update (complementary row pairs) per Householder step.
//   K2   wb_tql_kernel           (wb_eigh_ql.cuh) thread per k-point implicit QL: eigenvalues; streams its Givens rotations
//                                (used by the fallback of K3L only).
//   K3L  wb_eigvec_cta_kernel    eigenvectors of T by twisted factorisation (thread per eigenvalue, Zt in shared memory,
//                                128 KB at nw = 128); matrices it cannot resolve: replay of the rotation stream (thread =
//                                row of Z); back-transformation by the Householder reflectors with the eigenvector slices
//                                in registers, panels of NT / 8 eigenvectors; writes E (ascending) and U.
//
// Replaces  E_K, UU_K = np.linalg.eigh(HH_K)   (data_K/data_K.py:211-218, 309-322).
#pragma once
#include "wb_common.cuh"
#include "wb_eigh_ql.cuh"

// start of column j of a packed lower triangle stored column by column (== tri_index(j, j, n))
__device__ __forceinline__ int wb_cs(int j, int n) { return j * n - (j * (j - 1)) / 2; }

// sum of (a, b) over the CTA; `red` holds 2 x 32 double2, `phase` alternates between the two halves so that
// back-to-back reductions need one barrier each
__device__ __forceinline__ double2 wb_block_sum2(double a, double b, double2* red, int& phase, int nwarps) {
    a = warp_sum(a);
    b = warp_sum(b);
    double2* buf = red + phase * 32;
    if ((threadIdx.x & 31) == 0) buf[threadIdx.x >> 5] = make_double2(a, b);
    __syncthreads();
    double2 s = make_double2(0., 0.);
    for (int w = 0; w < nwarps; w++) { s.x += buf[w].x; s.y += buf[w].y; }
    phase ^= 1;
    return s;
}

__host__ inline size_t wb_tridiag_cta_smem_bytes(int n) {
    return sizeof(cplx) * ((size_t)n * (n + 1) / 2 + 10 * (size_t)n) + sizeof(double2) * 64;
}

// Output: d[t][n], e[t][n] (e[n-1] = 0), tau[t][n], Vh[t][k][i] = component i (> k+1) of Householder vector k.
template <int NT>
__global__ void __launch_bounds__(NT)
wb_tridiag_cta_kernel(const cplx* __restrict__ rec, WbLayout L, long k0, long nk, double* __restrict__ dout,
                      double* __restrict__ eout, cplx* __restrict__ tauout, cplx* __restrict__ Vh) {
    extern __shared__ __align__(16) cplx smem_t[];
    const int n = L.nw, ntri = n * (n + 1) / 2;
    cplx* P = smem_t;                   // packed lower triangle, column major
    cplx* vs = P + ntri;                // [n]
    cplx* ws = vs + n;                  // [n]
    cplx* xs = ws + n;                  // [8][n] partial matvec sums
    double2* red = (double2*)(xs + 8 * n);
    const int nwarps = NT / 32;
    int phase = 0;
    for (long t = blockIdx.x; t < nk; t += gridDim.x) {
        const long ik = k0 + t;
        const cplx* H = rec + ik * L.E + L.off_H;
        __syncthreads();
        // record: H(i, j), i <= j at tri_index(i, j) = cs(i) + j - i; lower triangle L(r, j) = conj(H(j, r))
        for (int x = threadIdx.x; x < ntri; x += NT) P[x] = cconj(H[x]);
        __syncthreads();
        for (int j = threadIdx.x; j < n; j += NT) P[wb_cs(j, n)].y = 0.;
        double* d = dout + t * n;
        double* e = eout + t * n;
        cplx* tau_o = tauout + t * n;
        for (int k = 0; k < n - 1; k++) {
            __syncthreads();
            const int csk = wb_cs(k, n);
            // ---- zlarfg on column k, rows k+1..n-1
            double sq = 0.;
            for (int r = k + 2 + threadIdx.x; r < n; r += NT) {
                cplx x = P[csk + r - k];
                sq += x.x * x.x + x.y * x.y;
            }
            const double xnorm2 = wb_block_sum2(sq, 0., red, phase, nwarps).x;
            const cplx alpha = P[csk + 1];
            cplx tau = cmake(0., 0.);
            double beta = alpha.x;
            cplx scale = cmake(0., 0.);
            if (xnorm2 != 0. || alpha.y != 0.) {
                beta = -copysign(sqrt(alpha.x * alpha.x + alpha.y * alpha.y + xnorm2), alpha.x);
                const double binv = 1. / beta;
                tau = cmake((beta - alpha.x) * binv, -alpha.y * binv);
                cplx den = cmake(alpha.x - beta, alpha.y);
                double dn = 1. / (den.x * den.x + den.y * den.y);
                scale = cmake(den.x * dn, -den.y * dn);
            }
            const bool active = (tau.x != 0. || tau.y != 0.);  // uniform
            __syncthreads();   // everybody has read alpha / column k
            for (int r = k + 1 + threadIdx.x; r < n; r += NT) {
                cplx v = (r == k + 1) ? cmake(1., 0.) : cmul(P[csk + r - k], scale);
                vs[r] = v;
                if (r > k + 1) P[csk + r - k] = v;   // the Householder vector stays in place
            }
            if (threadIdx.x == 0) { d[k] = P[csk].x; e[k] = beta; tau_o[k] = tau; }
            if (!active) continue;
            __syncthreads();
            // ---- x = tau A v  (rows > k): row r = sum_{j=k+1..r} L(r,j) v_j + sum_{i>r} conj(L(i,r)) v_i
            const int m = n - k - 1;              // active rows k+1 .. n-1, m elements per row
            // S threads per row, S = 2 / 4 / 8 as the active block shrinks (m S <= NT): a fixed thread pair per matrix row
            // left the rows <= k idle -- half of the CTA on average
            const int S = (m > NT / 4) ? 2 : ((m > NT / 8) ? 4 : 8);
            {
                const int R = NT / S;
                const int rho = threadIdx.x % R, h = threadIdx.x / R;
                if (rho < m) {
                    const int r = k + 1 + rho;
                    const int Lr = (m + S - 1) / S;
                    const int e0 = h * Lr, e1 = min(m, e0 + Lr);
                    const int rowcnt = r - k;
                    cplx a0 = cmake(0., 0.), a1 = cmake(0., 0.);
                    {
                        const int ea = e0, eb = min(e1, rowcnt);
                        int j = k + 1 + ea;
                        int cj = wb_cs(j, n);
                        for (int ee = ea; ee < eb; ee++, j++) {
                            cfma((ee & 1) ? a1 : a0, P[cj + r - j], vs[j]);
                            cj += n - j;
                        }
                    }
                    {
                        const int ea = max(e0, rowcnt), eb = e1;
                        const int csr = wb_cs(r, n);
                        for (int ee = ea; ee < eb; ee++) {
                            const int i = r + 1 + ee - rowcnt;
                            cfma_conj((ee & 1) ? a1 : a0, P[csr + i - r], vs[i]);
                        }
                    }
                    xs[h * n + r] = cadd(a0, a1);
                }
            }
            __syncthreads();
            cplx xr[(128 + NT - 1) / NT];   // rows of this thread (n <= 128)
            double pdx = 0., pdy = 0.;
#pragma unroll
            for (int u = 0; u < (128 + NT - 1) / NT; u++) {
                const int r = threadIdx.x + u * NT;
                xr[u] = cmake(0., 0.);
                if (r > k && r < n) {
                    cplx xsum = cadd(xs[r], xs[n + r]);
                    for (int h = 2; h < S; h++) xsum = cadd(xsum, xs[h * n + r]);
                    xr[u] = cmul(tau, xsum);
                    cplx pd = cconjmul(xr[u], vs[r]);
                    pdx += pd.x;
                    pdy += pd.y;
                }
            }
            const double2 dot = wb_block_sum2(pdx, pdy, red, phase, nwarps);
            const cplx al2 = cscale(-0.5, cmul(tau, cmake(dot.x, dot.y)));
#pragma unroll
            for (int u = 0; u < (128 + NT - 1) / NT; u++) {
                const int r = threadIdx.x + u * NT;
                if (r > k && r < n) ws[r] = cadd(xr[u], cmul(al2, vs[r]));
            }
            __syncthreads();
            // ---- L(r, j) -= v_r conj(w_j) + w_r conj(v_j),  k < j <= r.  Row r has r - k elements: rows k+1+u and n-1-u are
            // paired (m + 1 elements together) and every pair is split over S2 = NT / R2 threads -- all threads busy, equal work
            {
                const int J = (m + 1) >> 1;
                const int R2 = (J > NT / 8) ? NT / 4 : ((J > NT / 16) ? NT / 8 : NT / 16);
                const int S2 = NT / R2;
                const int u = threadIdx.x % R2, h = threadIdx.x / R2;
                if (u < J) {
                    const int rA = k + 1 + u, rB = n - 1 - u;
                    const int cntA = u + 1, cntB = (rB > rA) ? (m - u) : 0;
                    const int C = cntA + cntB, L2 = (C + S2 - 1) / S2;
                    const int e0 = h * L2, e1 = min(C, e0 + L2);
                    auto update = [&](int r, int ja, int jb) {
                        const cplx v = vs[r], w = ws[r];
                        int cj = wb_cs(ja, n);
                        for (int j = ja; j < jb; j++) {
                            const cplx wj = ws[j], vj = vs[j];
                            cplx a = P[cj + r - j];
                            a.x = fma(-v.x, wj.x, a.x);
                            a.y = fma(-v.y, wj.x, a.y);
                            a.x = fma(-v.y, wj.y, a.x);
                            a.y = fma(v.x, wj.y, a.y);
                            a.x = fma(-w.x, vj.x, a.x);
                            a.y = fma(-w.y, vj.x, a.y);
                            a.x = fma(-w.y, vj.y, a.x);
                            a.y = fma(w.x, vj.y, a.y);
                            P[cj + r - j] = a;
                            cj += n - j;
                        }
                    };
                    if (e0 < cntA) update(rA, k + 1 + e0, k + 1 + min(e1, cntA));
                    if (cntB > 0 && e1 > cntA) update(rB, k + 1 + max(e0 - cntA, 0), k + 1 + (e1 - cntA));
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            d[n - 1] = P[wb_cs(n - 1, n)].x;
            e[n - 1] = 0.;
            tau_o[n - 1] = cmake(0., 0.);
        }
        cplx* Vt = Vh + (size_t)t * n * n;
        for (int k = 0; k < n - 2; k++) {
            const int csk = wb_cs(k, n);
            for (int i = k + 2 + threadIdx.x; i < n; i += NT) Vt[k * n + i] = P[csk + i - k];
        }
    }
}

__host__ inline size_t wb_eigvec_cta_smem_bytes(int n, int pw) {   // pw = eigenvectors per panel = threads / 8
    const int ldu = n | 1;
    return sizeof(double) * (((size_t)n * n + 1) & ~(size_t)1) + sizeof(cplx) * ((size_t)pw * ldu + 8 * n + n + 2 * 128) +
           sizeof(double) * 4 * n + sizeof(int) * 2 * n + 16;
}

template <int NT, int EMAX>
__global__ void __launch_bounds__(NT)
wb_eigvec_cta_kernel(int n, long k0, long nk, const double* __restrict__ dvals, const cplx* __restrict__ tauin,
                     const cplx* __restrict__ Vh, const double2* __restrict__ rot, int capR, const int* __restrict__ hdr,
                     int capS, const int* __restrict__ nsweep, int want_U, double* __restrict__ Eout,
                     cplx* __restrict__ Uout, int* __restrict__ nfail, const double* __restrict__ d0,
                     const double* __restrict__ e0, int* __restrict__ nreplay) {
    static_assert(NT >= 128 && NT % 32 == 0, "threads 0 .. n-1 = rows of Z (n <= 128); NT / 8 panel columns x 8 row slices");
    constexpr int PW = NT / 8;   // eigenvectors per panel of the back-transformation
    extern __shared__ __align__(16) double smem_e[];
    const int ldu = n | 1;
    double* Zs = smem_e;                          // [col][row]
    cplx* up = (cplx*)(Zs + (((size_t)n * n + 1) & ~(size_t)1));   // [PW][ldu]
    cplx* vbuf = up + PW * ldu;                   // [2][4][n]: two stages of four reflectors
    cplx* taus = vbuf + 8 * n;                    // [n]
    double2* rs = (double2*)(taus + n);           // [2][128]
    double* dsm = (double*)(rs + 2 * 128);        // [n]
    double* esm = dsm + n;                        // [n] eigenvalues, ascending
    double* td = esm + n;                         // [n] diagonal of the tridiagonal matrix (as reduced: d0)
    double* te = td + n;                          // [n] its off-diagonal (e0)
    int* rank = (int*)(te + n);                   // [n]
    int* inv = rank + n;                          // [n]
    const int tid = threadIdx.x;
    for (long t = blockIdx.x; t < nk; t += gridDim.x) {
        const long ik = k0 + t;
        const int nsr = nsweep[t];
        if (nsr < 0) {   // uniform
            if (tid == 0) atomicAdd(nfail, 1);
            continue;
        }
        const int ns = nsr & 4095;
        __syncthreads();
        if (tid < n) {
            dsm[tid] = dvals[t * n + tid];
            taus[tid] = tauin[t * n + tid];
            if (d0) { td[tid] = d0[t * n + tid]; te[tid] = (tid < n - 1) ? e0[t * n + tid] : 0.; }
        }
        __syncthreads();
        // ---- sort
        if (tid < n) {
            const double myd = dsm[tid];
            int rk = 0;
            for (int j = 0; j < n; j++) {
                const double dj = dsm[j];
                rk += (dj < myd) || (dj == myd && j < tid);
            }
            rank[tid] = rk;
            inv[rk] = tid;
            esm[rk] = myd;
            Eout[ik * n + rk] = myd;
        }
        if (!want_U) continue;
        __syncthreads();
        // ---- eigenvectors of T, first choice: ONE twisted factorisation per eigenvalue (Fernando; Parlett-Dhillon; the
        // scheme of wb_eigh_tf.cuh), thread j = eigenvalue j (ascending), workspace and result = column j of Zt[i][j]
        // (component-major: conflict free).  O(n^2) per matrix instead of the O(n^3) replay of the QL rotations.
        // Vectors of eigenvalues closer than 5e-3 |T| are orthogonalised against each other in the order of the
        // eigenvalues; a run of more than 16, a vector that the projection cancels (numerically multiple eigenvalue)
        // or a residual above 256 eps |T| sends the WHOLE matrix to the replay below (uniform decision).
        bool tf_done = false;
        if (d0) {
            int myfail = 0;
            double tnorm = 0.;
            for (int i = 0; i < n; i++) tnorm = fmax(tnorm, fabs(td[i]) + fabs(te[i]) + ((i > 0) ? fabs(te[i - 1]) : 0.));
            const double eps = 2.220446049250313e-16;
            const double pivmin = fmax(eps * tnorm * 0.0009765625, 1e-290);
            const double ctol = 5e-3 * tnorm;   // orthogonality of neighbours ~ 4 eps |T| / gap: 2e-13 at the edge of the window
            if (tid < n) {
                const double sigma = esm[tid];
                double* W = Zs + tid;   // element i at W[i * n]
                double dcur = td[0] - sigma;
                for (int i = 0; i < n - 1; i++) {          // top-down: l_i = e_i / D+_i
                    if (fabs(dcur) < pivmin) dcur = -pivmin;
                    const double ei = te[i];
                    const double l = ei / dcur;
                    W[i * n] = l;
                    dcur = fma(-l, ei, td[i + 1] - sigma);
                }
                int r = n - 1;
                double gr = fabs(dcur);                     // gamma_{n-1} = D+_{n-1}
                dcur = td[n - 1] - sigma;
                for (int i = n - 2; i >= 0; i--) {          // bottom-up: gamma_i = D-_i - l_{i-1} e_{i-1}; twist = argmin
                    if (fabs(dcur) < pivmin) dcur = -pivmin;
                    const double ei = te[i];
                    const double u = ei / dcur;
                    dcur = fma(-u, ei, td[i] - sigma);
                    const double ga = fabs((i > 0) ? fma(-W[(i - 1) * n], te[i - 1], dcur) : dcur);
                    if (ga <= gr) { gr = ga; r = i; }
                }
                dcur = td[n - 1] - sigma;
                for (int i = n - 2; i >= r; i--) {          // bottom-up again: u_i = e_i / D-_{i+1} kept for i >= r
                    if (fabs(dcur) < pivmin) dcur = -pivmin;
                    const double ei = te[i];
                    const double u = ei / dcur;
                    W[i * n] = u;
                    dcur = fma(-u, ei, td[i] - sigma);
                }
                // the vector in place: z_r = 1, z_i = -l_i z_{i+1} (i < r), z_{i+1} = -u_i z_i (i >= r)
                double nrm2 = 1.;
                {
                    double zn = 1.;
                    for (int i = r - 1; i >= 0; i--) { zn = -W[i * n] * zn; W[i * n] = zn; nrm2 = fma(zn, zn, nrm2); }
                    double ucur = (r <= n - 2) ? W[r * n] : 0.;
                    W[r * n] = 1.;
                    zn = 1.;
                    for (int i = r; i <= n - 2; i++) {
                        zn = -ucur * zn;
                        ucur = (i + 1 <= n - 2) ? W[(i + 1) * n] : 0.;
                        W[(i + 1) * n] = zn;
                        nrm2 = fma(zn, zn, nrm2);
                    }
                }
                if (!(nrm2 > 0.) || !(nrm2 < 1e300)) myfail = 1;
                const double sc = rsqrt(nrm2);
                for (int i = 0; i < n; i++) W[i * n] *= sc;
            }
            __syncthreads();
            // ---- clusters: position of eigenvalue j inside its run of gaps < ctol
            int pos = 0;
            if (tid < n) {
                while (pos < 17 && tid - pos > 0 && esm[tid - pos] - esm[tid - pos - 1] < ctol) pos++;
                if (pos > 16) myfail = 1;
            }
            for (int round = 1; round <= 16; round++) {
                if (!__syncthreads_or(tid < n && pos >= round)) break;   // (also orders the rounds)
                if (tid < n && pos == round && !myfail) {
                    double* W = Zs + tid;
                    for (int pass = 0; pass < 2; pass++) {             // "twice is enough"
                        for (int p = tid - round; p < tid; p++) {
                            const double* Wp = Zs + p;
                            double dot = 0.;
                            for (int i = 0; i < n; i++) dot = fma(W[i * n], Wp[i * n], dot);
                            for (int i = 0; i < n; i++) W[i * n] = fma(-dot, Wp[i * n], W[i * n]);
                        }
                        double nn = 0.;
                        for (int i = 0; i < n; i++) nn = fma(W[i * n], W[i * n], nn);
                        if (pass == 0 && !(nn > 0.0025)) myfail = 1;      // cancelled: numerically multiple eigenvalue
                        const double sc = rsqrt(nn);
                        if (!myfail)
                            for (int i = 0; i < n; i++) W[i * n] *= sc;
                    }
                }
            }
            // ---- residual |(T - sigma) z| of what is stored
            if (tid < n && !myfail) {
                const double sigma = esm[tid];
                const double* W = Zs + tid;
                double r2 = 0., below = 0., zi = W[0];
                for (int i = 0; i < n; i++) {
                    const double znext = (i + 1 < n) ? W[(i + 1) * n] : 0.;
                    const double v = fma(td[i] - sigma, zi, below) + te[i] * znext;
                    below = te[i] * zi;
                    zi = znext;
                    r2 = fma(v, v, r2);
                }
                const double restol = 256. * eps * tnorm;
                if (!(r2 <= restol * restol)) myfail = 1;
            }
            tf_done = !__syncthreads_or(myfail);
            if (!tf_done && tid == 0 && nreplay) atomicAdd(nreplay, 1);
        }
        if (!tf_done) {
            __syncthreads();
            for (int x = tid; x < n * n; x += NT) Zs[x] = 0.;
            __syncthreads();
            if (tid < n) Zs[tid * n + tid] = 1.;
            // ---- replay the rotation stream (thread = row of Z)
            const double2* myrot = rot + (size_t)t * capR;
            const int* myhdr = hdr + (size_t)t * capS;
            int r0 = 0;
            if (ns > 0) {
                const int h = myhdr[0];
                const int len = (h >> 8) - (h & 255);
                if (tid < len) rs[tid] = myrot[tid];
            }
            for (int s = 0; s < ns; s++) {
                const int h = myhdr[s];
                const int l = h & 255, m = h >> 8, len = m - l;
                __syncthreads();
                // the next sweep's rotations: loaded into a register here, stored to shared memory after this sweep (a
                // `smem = global` prefetch would stall the in-order warp at the store for the whole memory latency)
                double2 nextrot = make_double2(0., 0.);
                int len2 = 0;
                if (s + 1 < ns) {
                    const int h2 = myhdr[s + 1];
                    len2 = (h2 >> 8) - (h2 & 255);
                    if (tid < len2) nextrot = myrot[r0 + len + tid];
                }
                if (tid < n) {
                    // rotation i touches columns i, i+1 of this thread's row; the loads of a batch of four columns
                    // are issued before the stores of the batch (distinct addresses, but the compiler cannot know)
                    const double2* q = rs + (s & 1) * 128;
                    double carry = Zs[m * n + tid];
                    int i = m - 1;
                    for (; i - 3 >= l; i -= 4) {
                        const double z0 = Zs[i * n + tid], z1 = Zs[(i - 1) * n + tid], z2 = Zs[(i - 2) * n + tid],
                                     z3 = Zs[(i - 3) * n + tid];
                        const double2 c0 = q[m - 1 - i], c1 = q[m - i], c2 = q[m + 1 - i], c3 = q[m + 2 - i];
                        const double o0 = c0.y * z0 + c0.x * carry;
                        carry = c0.x * z0 - c0.y * carry;
                        const double o1 = c1.y * z1 + c1.x * carry;
                        carry = c1.x * z1 - c1.y * carry;
                        const double o2 = c2.y * z2 + c2.x * carry;
                        carry = c2.x * z2 - c2.y * carry;
                        const double o3 = c3.y * z3 + c3.x * carry;
                        carry = c3.x * z3 - c3.y * carry;
                        Zs[(i + 1) * n + tid] = o0;
                        Zs[i * n + tid] = o1;
                        Zs[(i - 1) * n + tid] = o2;
                        Zs[(i - 2) * n + tid] = o3;
                    }
                    for (; i >= l; i--) {
                        const double2 cs = q[m - 1 - i];
                        const double zi = Zs[i * n + tid];
                        Zs[(i + 1) * n + tid] = cs.y * zi + cs.x * carry;
                        carry = cs.x * zi - cs.y * carry;
                    }
                    Zs[l * n + tid] = carry;
                }
                if (tid < len2) rs[((s + 1) & 1) * 128 + tid] = nextrot;
                r0 += len;
            }
        }
        __syncthreads();
        // ---- back-transformation  u <- H(0) H(1) ... H(n-2) u  in panels of PW eigenvectors (sorted order)
        const cplx* Vt = Vh + (size_t)t * n * n;
        const int c = tid >> 3, sl = tid & 7;
        // Thread (c, sl) keeps the elements i = sl + 8 e of eigenvector c of the panel in REGISTERS for all n - 1 reflectors
        // (EMAX = ceil(n / 8) complex values) and reads every reflector element once per step: the shared-memory traffic
        // is the reflector only.  (With u in shared memory the kernel moved 80 bytes per 16 flops: 55 % of its time, ncu
        // profiles/r2/cfg5_eig_kernels.txt.)
        for (int p0 = 0; p0 < n; p0 += PW) {
            __syncthreads();
            cplx u[EMAX];
            {
                const bool colok = (p0 + c < n);
                // replay: Z[column of the QL order][row]; twisted factorisation: Zt[component][eigenvalue, ascending]
                const double* zcol = tf_done ? Zs + (colok ? p0 + c : 0) : Zs + (size_t)(colok ? inv[p0 + c] : 0) * n;
                const int zstride = tf_done ? n : 1;
#pragma unroll
                for (int e = 0; e < EMAX; e++) {
                    const int i = sl + 8 * e;
                    u[e] = cmake((colok && i < n) ? zcol[(size_t)i * zstride] : 0., 0.);
                }
            }
            // Reflectors are staged RB at a time (double buffer).  The next group is loaded into REGISTERS before the RB
            // steps of the current one and stored to shared memory after them: a warp issues in order, so a prefetch
            // written as `smem = global` stalls at the store and exposes the full memory latency in every step (what the
            // one-reflector-per-barrier loop did: ~1800 clocks per step, 55 % of the kernel).  n <= NT: element i = tid.
            constexpr int RB = 4;
            static_assert(NT >= 128, "one reflector element per thread");
            auto fetch_group = [&](int kb, cplx (&pre)[RB]) {
#pragma unroll
                for (int r = 0; r < RB; r++) {
                    const int k = kb - r;
                    pre[r] = cmake(0., 0.);
                    if (k >= 0 && tid < n) pre[r] = (tid > k + 1) ? Vt[(size_t)k * n + tid] : ((tid == k + 1) ? cmake(1., 0.) : cmake(0., 0.));
                }
            };
            auto store_group = [&](const cplx (&pre)[RB], cplx* dst) {
                if (tid < n) {
#pragma unroll
                    for (int r = 0; r < RB; r++) dst[r * n + tid] = pre[r];
                }
            };
            cplx pre[RB];
            fetch_group(n - 2, pre);
            store_group(pre, vbuf);
            int cur = 0;
            for (int kb = n - 2; kb >= 0; kb -= RB, cur ^= 1) {
                __syncthreads();
                const bool more = (kb - RB >= 0);
                if (more) fetch_group(kb - RB, pre);
#pragma unroll 1
                for (int r = 0; r < RB; r++) {
                    const int k = kb - r;
                    if (k < 0) break;
                    const cplx tau = taus[k];
                    if (tau.x == 0. && tau.y == 0.) continue;   // uniform
                    const cplx* v = vbuf + (cur * RB + r) * n;   // zero up to element k, one at k + 1
                    cplx vr[EMAX];
                    cplx s0 = cmake(0., 0.), s1 = cmake(0., 0.);
#pragma unroll
                    for (int e = 0; e < EMAX; e++) {
                        vr[e] = cmake(0., 0.);
                        if (8 * e + 7 > k) {   // (uniform: below, every element of the block is zero)
                            const int i = sl + 8 * e;
                            if (i < n) vr[e] = v[i];
                            if (e & 1) cfma_conj(s1, vr[e], u[e]);
                            else cfma_conj(s0, vr[e], u[e]);
                        }
                    }
                    cplx sd = cadd(s0, s1);
#pragma unroll
                    for (int o = 1; o < 8; o <<= 1) {
                        sd.x += __shfl_xor_sync(0xffffffffu, sd.x, o);
                        sd.y += __shfl_xor_sync(0xffffffffu, sd.y, o);
                    }
                    const cplx ts = cmul(tau, sd);
#pragma unroll
                    for (int e = 0; e < EMAX; e++) {
                        if (8 * e + 7 > k) {
                            u[e].x -= ts.x * vr[e].x - ts.y * vr[e].y;
                            u[e].y -= ts.x * vr[e].y + ts.y * vr[e].x;
                        }
                    }
                }
                if (more) store_group(pre, vbuf + (cur ^ 1) * RB * n);
            }
#pragma unroll
            for (int e = 0; e < EMAX; e++) {
                const int i = sl + 8 * e;
                if (i < n) up[c * ldu + i] = u[e];
            }
            __syncthreads();
            cplx* Uo = Uout + (size_t)ik * n * n;
            for (int x = tid; x < PW * n; x += NT) {
                const int cc = x % PW, i = x / PW;
                if (p0 + cc < n) Uo[i * n + p0 + cc] = up[cc * ldu + i];
            }
        }
    }
}
