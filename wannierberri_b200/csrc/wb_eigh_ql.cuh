// Batched complex Hermitian eigensolver for nw <= 32: Householder tridiagonalisation + implicit
// QL + back-transformation, split so that every phase keeps all 32 lanes (or as many as the
// matrix size allows) on USEFUL FP64 work and nothing scalar is replicated across a warp:
//
//   K1  wb_tridiag_kernel      warp per k-point, lane i = row i of A held in REGISTERS (static
//                              indexing, fully unrolled).  A -> (d, e, tau, Householder vectors).
//   K2  wb_tql_kernel          THREAD per k-point: implicit-shift QL on the real tridiagonal
//                              (d, e); the Givens rotations (c, s) of every sweep are streamed
//                              to memory instead of being applied to a matrix.
//   K3  wb_eigvec_kernel       warp per k-point, lane i = row i of Z (real, registers): replay the
//                              rotation stream; sort; transpose through shared memory; lane j = column j:
//                              apply the Householder reflectors in reverse; write E (ascending), U.
//
// Replaces  E_K, UU_K = np.linalg.eigh(HH_K)   (data_K/data_K.py:211-218, 309-322).  The maths is the
// classic zhetd2 + tql2 + zunm2l sequence (LAPACK / EISPACK), restated for this layout.
// A k-point whose QL iteration exceeds the stream capacity is flagged and re-solved by the
// Jacobi kernel (wb_eigh_jacobi.cuh).
#pragma once
#include "wb_common.cuh"
#include "wb_tma.cuh"

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ cplx shfl_c(cplx v, int src) {
    return cmake(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}

// ------------------------------------------------------------------------------------------ K1
// Output: d[ik][nw], e[ik][nw] (e[nw-1] = 0), tau[ik][nw], V = A rows (Householder vector k in
// column k, rows k+2..nw-1) written to Vout[ik][nw][nw].
// EXACT: nw == NWP is known at compile time, every bounds guard folds away.
template <int NWP, int WARPS, bool EXACT>
__global__ void __launch_bounds__(WARPS * 32)
wb_tridiag_kernel(const cplx* __restrict__ rec, WbLayout L, long k0, long nk, double* __restrict__ dout,
                  double* __restrict__ eout, cplx* __restrict__ tauout, cplx* __restrict__ Vout) {
    __shared__ cplx vs_all[WARPS][32];
    __shared__ cplx ws_all[WARPS][32];
    __shared__ cplx hs_all[WARPS][NWP * (NWP + 1) / 2];   // packed upper triangle of H, staged coalesced
    const int nw = EXACT ? NWP : L.nw;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    cplx* vs = vs_all[warp];
    cplx* ws = ws_all[warp];
    cplx* hs = hs_all[warp];
    long t = (long)blockIdx.x * WARPS + warp;
    if (t >= nk) return;
    long ik = k0 + t;
    const cplx* H = rec + ik * L.E + L.off_H;
    for (int x = lane; x < nw * (nw + 1) / 2; x += 32) hs[x] = H[x];
    __syncwarp();
    cplx a[NWP];
#pragma unroll
    for (int j = 0; j < NWP; j++) {
        a[j] = cmake(0., 0.);
        if (j < nw && lane < nw) {
            a[j] = (lane <= j) ? hs[tri_index(lane, j, nw)] : cconj(hs[tri_index(j, lane, nw)]);
            if (j == lane) a[j].y = 0.;
        }
    }
    double* d = dout + t * nw;
    double* e = eout + t * nw;
    cplx* tau_o = tauout + t * nw;
#pragma unroll
    for (int k = 0; k < NWP - 1; k++) {
        if (k < nw - 1) {  // uniform
            cplx xk = a[k];  // column k: element of row `lane`
            double sq = (lane > k + 1 && lane < nw) ? (xk.x * xk.x + xk.y * xk.y) : 0.;
            double xnorm2 = warp_sum(sq);
            cplx alpha = shfl_c(xk, k + 1);
            cplx tau = cmake(0., 0.);
            double beta = alpha.x;
            cplx v = cmake(0., 0.);
            if (xnorm2 != 0. || alpha.y != 0.) {  // zlarfg
                beta = -copysign(sqrt(alpha.x * alpha.x + alpha.y * alpha.y + xnorm2), alpha.x);
                const double binv = 1. / beta;
                tau = cmake((beta - alpha.x) * binv, -alpha.y * binv);
                cplx den = cmake(alpha.x - beta, alpha.y);
                double dn = 1. / (den.x * den.x + den.y * den.y);
                cplx scale = cmake(den.x * dn, -den.y * dn);
                if (lane > k + 1 && lane < nw) v = cmul(xk, scale);
            }
            if (lane == k + 1) v = cmake(1., 0.);
            const bool active = (tau.x != 0. || tau.y != 0.);  // uniform
            if (active) {
                vs[lane] = v;
                __syncwarp();
                // x = tau * A v   (rows > k)
                cplx x0 = cmake(0., 0.), x1 = cmake(0., 0.);  // two chains: halves the dependent-FMA latency
#pragma unroll
                for (int j = k + 1; j < NWP; j++)
                    if (j < nw) {
                        if ((j - k) & 1) cfma(x0, a[j], vs[j]);
                        else cfma(x1, a[j], vs[j]);
                    }
                cplx x = cmul(tau, cadd(x0, x1));
                if (lane <= k || lane >= nw) x = cmake(0., 0.);
                // dot = x^H v
                cplx pd = cconjmul(x, v);
                cplx dot = cmake(warp_sum(pd.x), warp_sum(pd.y));
                cplx al2 = cscale(-0.5, cmul(tau, dot));
                cplx w = cadd(x, cmul(al2, v));
                ws[lane] = w;
                __syncwarp();
                // A -= v w^H + w v^H
                if (lane > k && lane < nw) {
#pragma unroll
                    for (int j = k + 1; j < NWP; j++)
                        if (j < nw) {
                            const cplx wj = ws[j], vj = vs[j];   // a -= v conj(w_j) + w conj(v_j): 8 FMA
                            a[j].x = fma(-v.x, wj.x, a[j].x);
                            a[j].y = fma(-v.y, wj.x, a[j].y);
                            a[j].x = fma(-v.y, wj.y, a[j].x);
                            a[j].y = fma(v.x, wj.y, a[j].y);
                            a[j].x = fma(-w.x, vj.x, a[j].x);
                            a[j].y = fma(-w.y, vj.x, a[j].y);
                            a[j].x = fma(-w.y, vj.y, a[j].x);
                            a[j].y = fma(w.x, vj.y, a[j].y);
                        }
                }
                __syncwarp();
            }
            // store the Householder vector in place (rows > k+1 of column k)
            if (lane > k + 1) a[k] = v;
            if (lane == k) d[k] = a[k].x;
            if (lane == 0) { e[k] = beta; tau_o[k] = tau; }
        }
    }
    if (lane == nw - 1) {
#pragma unroll
        for (int j = 0; j < NWP; j++)
            if (j == nw - 1) d[j] = a[j].x;
    }
    if (lane == 0) { e[nw - 1] = 0.; tau_o[nw - 1] = cmake(0., 0.); }
    if (lane < nw) {
        cplx* Vrow = Vout + (ik * nw + lane) * nw;
#pragma unroll
        for (int j = 0; j < NWP; j++)
            if (j < nw) Vrow[j] = a[j];
    }
}

// ------------------------------------------------------------------------------------------ K1b
// Same reduction with TWO k-points per warp (16 lanes each) for 16 < NW <= 18: the last 16 rows of the matrix
// live on the 16 lanes; the R0 = NW - 16 leading rows are never needed explicitly -- row 0 is never active, row 1
// is active in step 0 only, where its elements are the conjugates of column 1 held by the other rows and all that
// survives of it is its updated diagonal.  Twice the lane utilisation of wb_tridiag_kernel.
__device__ __forceinline__ double half_sum(double v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int NW, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 4)
wb_tridiag2_kernel(const cplx* __restrict__ rec, WbLayout L, long k0, long nk, double* __restrict__ dout,
                   double* __restrict__ eout, cplx* __restrict__ tauout, cplx* __restrict__ Vout) {
    static_assert(NW > 16 && NW <= 18, "two leading rows at most are handled implicitly");
    constexpr int R0 = NW - 16, NTRI = NW * (NW + 1) / 2;
    __shared__ cplx vs_all[WARPS * 2][NW];
    __shared__ cplx ws_all[WARPS * 2][NW];
    __shared__ cplx hs_all[WARPS * 2][NTRI];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int half = lane >> 4, hl = lane & 15, hbase = lane & 16;
    const int row = hl + R0;
    cplx* vs = vs_all[warp * 2 + half];
    cplx* ws = ws_all[warp * 2 + half];
    cplx* hs = hs_all[warp * 2 + half];
    long t = ((long)blockIdx.x * WARPS + warp) * 2 + half;
    const bool live = t < nk;
    if (!live) t = nk - 1;   // the idle half mirrors the last k-point and writes nothing
    const long ik = k0 + t;
    const cplx* H = rec + ik * L.E + L.off_H;
    for (int x = hl; x < NTRI; x += 16) hs[x] = H[x];
    __syncwarp();
    cplx a[NW];
#pragma unroll
    for (int j = 0; j < NW; j++) {
        a[j] = (row <= j) ? hs[tri_index(row, j, NW)] : cconj(hs[tri_index(j, row, NW)]);
        if (j == row) a[j].y = 0.;
    }
    double a11 = (R0 == 2) ? hs[tri_index(1, 1, NW)].x : 0.;   // diagonal of the implicit row 1
    double* d = dout + t * NW;
    double* e = eout + t * NW;
    cplx* tau_o = tauout + t * NW;
    if (live && hl == 0) d[0] = hs[0].x;   // row 0 is never touched (R0 >= 1)
#pragma unroll
    for (int k = 0; k < NW - 1; k++) {
        const bool step0 = (R0 == 2 && k == 0);   // the pivot row k+1 = 1 is implicit
        cplx xk = a[k];
        double sq = (row > k + 1) ? (xk.x * xk.x + xk.y * xk.y) : 0.;
        double xnorm2 = half_sum(sq);
        cplx alpha;
        if (step0) alpha = cconj(hs[tri_index(0, 1, NW)]);
        else alpha = cmake(__shfl_sync(0xffffffffu, xk.x, hbase + (k + 1 - R0)), __shfl_sync(0xffffffffu, xk.y, hbase + (k + 1 - R0)));
        cplx tau = cmake(0., 0.);
        double beta = alpha.x;
        cplx v = cmake(0., 0.);
        if (xnorm2 != 0. || alpha.y != 0.) {  // zlarfg (rsqrt / reciprocal instead of sqrt and two divisions)
            const double n2 = alpha.x * alpha.x + alpha.y * alpha.y + xnorm2;
            const double rinv = rsqrt(n2);
            beta = -copysign(n2 * rinv, alpha.x);
            const double binv = -copysign(rinv, alpha.x);
            tau = cmake((beta - alpha.x) * binv, -alpha.y * binv);
            cplx den = cmake(alpha.x - beta, alpha.y);
            double dn = __drcp_rn(den.x * den.x + den.y * den.y);
            cplx scale = cmake(den.x * dn, -den.y * dn);
            if (row > k + 1) v = cmul(xk, scale);
        }
        if (row == k + 1) v = cmake(1., 0.);
        const bool active = (tau.x != 0. || tau.y != 0.);  // uniform within the half-warp
        // (both halves run the body; an inactive half works on tau = 0, which leaves its matrix unchanged)
        {
            __syncwarp();
            if (row > k) vs[row] = v;
            if (step0 && hl == 0) vs[1] = cmake(1., 0.);
            __syncwarp();
            // x = tau * A v   (rows > k)
            cplx x0 = cmake(0., 0.), x1 = cmake(0., 0.);
#pragma unroll
            for (int j = k + 1; j < NW; j++) {
                if ((j - k) & 1) cfma(x0, a[j], vs[j]);
                else cfma(x1, a[j], vs[j]);
            }
            cplx x = cmul(tau, cadd(x0, x1));
            if (row <= k) x = cmake(0., 0.);
            cplx xh = cmake(0., 0.);   // x of the implicit row 1 (step 0):  tau (a11 + sum_rows conj(a[row][1]) v[row])
            if (step0) {
                cplx p = cconjmul(a[1], v);
                xh = cmul(tau, cmake(a11 + half_sum(p.x), half_sum(p.y)));
            }
            // dot = x^H v
            cplx pd = cconjmul(x, v);
            cplx dot = cmake(half_sum(pd.x), half_sum(pd.y));
            if (step0) { dot.x += xh.x; dot.y -= xh.y; }   // conj(xh) * 1
            cplx al2 = cscale(-0.5, cmul(tau, dot));
            cplx w = cadd(x, cmul(al2, v));
            if (row > k) ws[row] = w;
            if (step0) {
                cplx wh = cadd(xh, al2);
                if (hl == 0) ws[1] = wh;
                a11 -= 2. * wh.x;
            }
            __syncwarp();
            // A -= v w^H + w v^H
            if (active && row > k) {
#pragma unroll
                for (int j = k + 1; j < NW; j++) {
                    const cplx wj = ws[j], vj = vs[j];
                    a[j].x = fma(-v.x, wj.x, a[j].x);
                    a[j].y = fma(-v.y, wj.x, a[j].y);
                    a[j].x = fma(-v.y, wj.y, a[j].x);
                    a[j].y = fma(v.x, wj.y, a[j].y);
                    a[j].x = fma(-w.x, vj.x, a[j].x);
                    a[j].y = fma(-w.y, vj.x, a[j].y);
                    a[j].x = fma(-w.y, vj.y, a[j].x);
                    a[j].y = fma(w.x, vj.y, a[j].y);
                }
            }
        }
        // store the Householder vector in place (rows > k+1 of column k)
        if (row > k + 1) a[k] = v;
        if (live) {
            if (row == k && k >= R0) d[k] = a[k].x;
            if (R0 == 2 && k == 1 && hl == 0) d[1] = a11;
            if (hl == 0) { e[k] = beta; tau_o[k] = tau; }
        }
    }
    if (live) {
        if (row == NW - 1) d[NW - 1] = a[NW - 1].x;
        if (hl == 0) { e[NW - 1] = 0.; tau_o[NW - 1] = cmake(0., 0.); }
        cplx* Vrow = Vout + (ik * NW + row) * NW;
#pragma unroll
        for (int j = 0; j < NW; j++) Vrow[j] = a[j];
    }
}

// ------------------------------------------------------------------------------------------ K2
// Implicit QL with Wilkinson-type shift (EISPACK tql2 / "tqli").  Thread per k-point.
// Stream: rot[ik][capR] (c, s) in application order; hdr[ik][capS] = l | (m << 8) per sweep;
// nsweep[ik] = number of sweeps, or -1 on overflow / non-convergence (=> Jacobi fallback).
template <int NT>
__global__ void __launch_bounds__(NT)
wb_tql_kernel(int nw, long nk, double* __restrict__ dio, const double* __restrict__ ein,
              double2* __restrict__ rot, int capR, int* __restrict__ hdr, int capS, int* __restrict__ nsweep) {
    extern __shared__ double smem_q[];  // d[nw][NT], e[nw][NT]
    double* d = smem_q + threadIdx.x;
    double* e = smem_q + (size_t)nw * NT + threadIdx.x;
    long t = (long)blockIdx.x * NT + threadIdx.x;
    if (t >= nk) return;
    for (int i = 0; i < nw; i++) {
        d[i * NT] = dio[t * nw + i];
        e[i * NT] = ein[t * nw + i];
    }
    double2* myrot = rot + (size_t)t * capR;
    int* myhdr = hdr + (size_t)t * capS;
    int nr = 0, ns = 0;
    bool fail = false;
    for (int l = 0; l < nw && !fail; l++) {
        int iter = 0;
        while (true) {
            int m = l;
            for (; m < nw - 1; m++) {
                double dd = fabs(d[m * NT]) + fabs(d[(m + 1) * NT]);
                if (fabs(e[m * NT]) + dd == dd) break;
            }
            if (m == l) break;
            if (++iter > 40 || ns >= capS || nr + (m - l) > capR) { fail = true; break; }
            double el = e[l * NT];
            double g = (d[(l + 1) * NT] - d[l * NT]) / (2. * el);
            double r = hypot(g, 1.);
            g = d[m * NT] - d[l * NT] + el / (g + copysign(r, g));
            double s = 1., c = 1., p = 0.;
            myhdr[ns++] = l | (m << 8);
            int i = m - 1;
            for (; i >= l; i--) {
                double f = s * e[i * NT];
                double b = c * e[i * NT];
                r = sqrt(f * f + g * g);
                if (!(r > 1e-140 && r < 1e140)) r = hypot(f, g);   // out of the safe range of the plain formula
                e[(i + 1) * NT] = r;
                if (r == 0.) {  // recover from underflow: identity rotations for the rest of the sweep
                    d[(i + 1) * NT] -= p;
                    e[m * NT] = 0.;
                    break;
                }
                const double rinv = 1. / r;
                s = f * rinv;
                c = g * rinv;
                g = d[(i + 1) * NT] - p;
                r = (d[i * NT] - g) * s + 2. * c * b;
                p = s * r;
                d[(i + 1) * NT] = g + p;
                g = c * r - b;
                myrot[nr++] = make_double2(c, s);
            }
            if (i >= l) {  // broke out early: pad the sweep with identity rotations
                for (; i >= l; i--) myrot[nr++] = make_double2(1., 0.);
                continue;
            }
            d[l * NT] -= p;
            e[l * NT] = g;
            e[m * NT] = 0.;
        }
    }
    for (int i = 0; i < nw; i++) dio[t * nw + i] = d[i * NT];
    nsweep[t] = fail ? -1 : (ns | (nr << 12));   // sweeps | rotations << 12
}

// ------------------------------------------------------------------------------------------ K3
// Per warp: ONE set of TMA bulk copies brings the Householder vectors, tau, the sweep headers and the whole
// rotation stream of the k-point into shared memory (a single exposed latency instead of one per sweep).
__host__ __device__ inline int wb_eigvec_smem_per_warp(int nw, int capR, int capS) {  // in 16-byte units
    return nw * nw + nw + capR + (capS + 3) / 4 + 1;   // (the transpose buffer aliases the rotation stream)
}

template <int NWP, int WARPS, bool EXACT>
__global__ void __launch_bounds__(WARPS * 32)
wb_eigvec_kernel(int nw_rt, long k0, long nk, const double* __restrict__ dvals, const cplx* __restrict__ tauin,
                 const double2* __restrict__ rot, int capR, const int* __restrict__ hdr, int capS,
                 const int* __restrict__ nsweep, double* __restrict__ Eout, cplx* __restrict__ VU,
                 int* __restrict__ fail_list, int* __restrict__ nfail) {
    extern __shared__ __align__(16) cplx smem_v[];
    // per warp: V[nw][nw] (Householder vectors), tau[nw], rot[capR], hdr[capS], mbarrier; the transpose buffer
    // Zs[nw][nw+1] doubles reuses the rotation stream, which is dead after the replay (nw(nw+1) <= 2 capR)
    const int nw = EXACT ? NWP : nw_rt;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per_warp = wb_eigvec_smem_per_warp(nw, capR, capS);
    cplx* V = smem_v + (size_t)warp * per_warp;
    cplx* taus = V + nw * nw;
    const double2* rots = (const double2*)(taus + nw);
    double* Zs = (double*)rots;
    const int* hdrs = (const int*)(rots + capR);
    uint64_t* bar = (uint64_t*)(hdrs + (capS + 3) / 4 * 4);
    long t = (long)blockIdx.x * WARPS + warp;
    if (t >= nk) return;
    long ik = k0 + t;
    const int nsr = nsweep[t];
    if (nsr < 0) {
        if (lane == 0) fail_list[atomicAdd(nfail, 1)] = (int)t;
        return;
    }
    const int ns = nsr & 4095, nr = nsr >> 12;
    if (lane == 0) {
        wb_mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        const uint32_t bV = (uint32_t)(nw * nw * 16), bT = (uint32_t)(nw * 16), bR = (uint32_t)(nr * 16),
                       bH = (uint32_t)((ns + 3) / 4 * 16);
        wb_mbar_expect_tx(bar, bV + bT + bR + bH);
        wb_bulk_g2s(V, VU + ik * nw * nw, bV, bar);
        wb_bulk_g2s(taus, tauin + t * nw, bT, bar);
        if (bR) wb_bulk_g2s((void*)rots, rot + (size_t)t * capR, bR, bar);
        if (bH) wb_bulk_g2s((void*)hdrs, hdr + (size_t)t * capS, bH, bar);
    }
    double myd = (lane < nw) ? dvals[t * nw + lane] : CUDART_INF;
    __syncwarp();
    wb_mbar_wait(bar, 0);
    // ---- replay the rotation stream on Z = I (lane = row)
    double z[NWP];
#pragma unroll
    for (int j = 0; j < NWP; j++) z[j] = (j == lane) ? 1. : 0.;
    int r0 = 0;
    for (int s = 0; s < ns; s++) {
        const int h = hdrs[s];
        const int l = h & 255, m = h >> 8;
        const double2* cs = rots + r0 + (m - 1);   // rotation of index i sits at cs[-i]
        r0 += m - l;
#pragma unroll
        for (int ii = 0; ii < NWP - 1; ii++) {
            const int i = NWP - 2 - ii;
            if (i < m && i >= l) {  // uniform
                double2 q = cs[-i];
                double f = z[i + 1];
                z[i + 1] = q.y * z[i] + q.x * f;
                z[i] = q.x * z[i] - q.y * f;
            }
        }
    }
    // ---- sort: rank of eigenvalue `lane`
    int rank = 0;
    for (int j = 0; j < nw; j++) {
        double dj = __shfl_sync(0xffffffffu, myd, j);
        rank += (dj < myd) || (dj == myd && j < lane);
    }
    if (lane < nw) Eout[ik * nw + rank] = myd;
    // ---- transpose: lane j takes column j of Z
    const int ldz = nw + 1;
    __syncwarp();
    if (lane < nw) {
#pragma unroll
        for (int j = 0; j < NWP; j++)
            if (j < nw) Zs[lane * ldz + j] = z[j];
    }
    __syncwarp();
    cplx u[NWP];
#pragma unroll
    for (int i = 0; i < NWP; i++) u[i] = cmake((i < nw && lane < nw) ? Zs[i * ldz + lane] : 0., 0.);
    // ---- back-transformation  u <- H(0) H(1) ... H(nw-2) u   (apply in reverse order)
#pragma unroll
    for (int kk = 0; kk < NWP - 1; kk++) {
        const int k = NWP - 2 - kk;
        if (k < nw - 1) {
            cplx tau = taus[k];
            if (tau.x != 0. || tau.y != 0.) {  // uniform
                cplx sdot = u[k + 1];  // v[k+1] = 1
                cplx sdot1 = cmake(0., 0.);
#pragma unroll
                for (int i = k + 2; i < NWP; i++)
                    if (i < nw) {
                        if ((i - k) & 1) cfma_conj(sdot1, V[i * nw + k], u[i]);
                        else cfma_conj(sdot, V[i * nw + k], u[i]);
                    }
                cplx ts = cmul(tau, cadd(sdot, sdot1));
                u[k + 1] = csub(u[k + 1], ts);
#pragma unroll
                for (int i = k + 2; i < NWP; i++)
                    if (i < nw) {
                        cplx vv = V[i * nw + k];
                        u[i].x -= ts.x * vv.x - ts.y * vv.y;
                        u[i].y -= ts.x * vv.y + ts.y * vv.x;
                    }
            }
        }
    }
    __syncwarp();
    if (lane < nw) {
        cplx* Uo = VU + ik * nw * nw;
#pragma unroll
        for (int i = 0; i < NWP; i++)
            if (i < nw) Uo[i * nw + rank] = u[i];
    }
}
