// libwbgpu.so -- C-ABI (include/wbgpu.h) and host-side orchestration of the sm_100a kernels.
//
// One context = one System_R replica resident on one GPU.  A scan call walks the K-block list in
// batches ("launches") sized to the workspace: twiddles -> 3 axis passes of the separable R->k
// transform -> eigensolver -> rotation+formula events -> histogram accumulation; the per-spec
// histograms stay on the device until the last batch, then one finalize kernel per spec.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/wbgpu.h"
#include "wb_common.cuh"
#include "wb_fourier.cuh"
#include "wb_eigh_jacobi.cuh"
#include "wb_eigh_ql.cuh"
#include "wb_eigh_large.cuh"
#include "wb_groups.cuh"
#include "wb_rotate_formula.cuh"
#include "wb_rotate_dmma.cuh"
#include "wb_events_generic.cuh"
#include "wb_rotate_gemm.cuh"
#include "wb_scan.cuh"
#include "wb_kubo.cuh"
#include "wb_fsea.cuh"
#include "wb_tetra.cuh"
#include "wb_probe.cuh"
#include "wb_launch.h"

static thread_local std::string g_err;

static int set_err(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}

#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            return set_err("CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), __FILE__, __LINE__,  \
                           cudaGetErrorString(e__));                                                 \
    } while (0)

struct wbgpu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int nw = 0, nR = 0;
    double cell_volume = 0;
    int* d_iRvec = nullptr;
    double* d_T = nullptr;
    cplx* d_XR[WBGPU_NKEYS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int3 rmin{0, 0, 0}, nbox{1, 1, 1};
    // plan
    bool planned = false;
    int N[3] = {1, 1, 1};
    uint32_t mask = 0;
    int external = 1;
    WbLayout L;
    cplx* d_table = nullptr;
    long nk_block = 1;
    int nb_max = 1;  // K-blocks per launch
    // workspace
    cplx* d_W[3] = {nullptr, nullptr, nullptr};
    cplx *d_Z = nullptr, *d_Y = nullptr, *d_X = nullptr, *d_U = nullptr;
    double *d_E = nullptr, *d_evlabel = nullptr, *d_evval = nullptr;
    double *d_hist = nullptr, *d_cum = nullptr;
    size_t hist_cap = 0;
    double *d_dK = nullptr, *d_weight = nullptr, *d_out = nullptr, *d_axes = nullptr;
    size_t dK_cap = 0, out_cap = 0, axes_cap = 0;
    int* d_sweeps = nullptr;
    // rotated matrices in global memory (generic DMMA GEMM path), per sub-batch of k-points
    double* d_xbar = nullptr;
    size_t xbar_cap = 0;
    double* d_mx = nullptr;
    size_t mx_cap = 0;
    double *d_d0 = nullptr, *d_e0 = nullptr;   // tridiagonal matrices as reduced (nw > 32: the QL kernel works in place)
    double* d_colwin = nullptr;   // int2 per k-point: band range of the band groups (column window of the GEMM rotation)
    size_t colwin_cap = 0;
    // Kubo path: entry lists of a sub-batch, global accumulator, axes
    double* d_kent = nullptr;
    size_t kent_cap = 0;
    double* d_kacc = nullptr;
    size_t kacc_cap = 0;
    double* d_shcJ = nullptr;   // spin-velocity matrices of a sub-batch [k][9][nw][nw] complex
    size_t shcJ_cap = 0;
    // tetrahedron method: H-only R-space table, corner energies [8][nkl][nw], per-band min / max over the cell
    cplx* d_tableH = nullptr;
    double* d_Ec = nullptr;
    size_t Ec_cap = 0;
    // Householder+QL eigensolver work space (per sub-batch of eig_chunk k-points)
    long eig_chunk = 0;
    int capR = 0, capS = 0;
    double *d_dw = nullptr, *d_ew = nullptr;
    cplx* d_tau = nullptr;
    cplx* d_Vh = nullptr;   // Householder vectors, reflector-major (nw > 32)
    double2* d_rot = nullptr;
    int *d_hdr = nullptr, *d_nsweep = nullptr, *d_faillist = nullptr, *d_nfail = nullptr;
    int64_t eig_fallbacks = 0;
    int last_sweeps = 0, last_resolved = 0;
    int64_t launches = 0;
    int eig_method = 0;
    int ev_ncmax = 1;
    int fourier_method = 0; // 0 = axes 1 and 0 fused (when the tile fits), 1 = three separate axis passes
    int rotate_method = 0;  // 0 = automatic, 1 = generic shared-memory DFMA kernel, 2 = DMMA kernel (Omega, nw <= 20),
                            // 3 = compile-time-NW DMMA kernel, 4 = batched DMMA GEMM to global memory + formula kernel
    int smem_optin = 0;
    int rot_r2 = -1;        // experiment: override of the step-2 warp rotation of the fused rotation kernel
    int kubo_method = 0;    // 0 = register-tiled accumulation kernels, 1 = per-(omega, component) kernel
    int rotate_trim = 1;    // 1 = form only the needed columns of the rotated matrices when they are hermitian
    int dh_packed = 1;      // 1 = pack d_a H as a triangle when it is hermitian in R-space, 0 = never
    int gemm_stack = 1;     // 1 = the size-generic rotation stacks several channels per CTA when nw <= 32
    std::vector<int> h_iRvec;
    // optional per-stage device timing (option "timing"): events around each stage of each batch
    int timing = 0;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<std::pair<int, int>> ev_used;  // (stage, index of start event); stop = index + 1
    size_t ev_next = 0;
    double stage_ms[WBGPU_NSTAGES] = {0, 0, 0, 0, 0};
    int64_t stage_calls[WBGPU_NSTAGES] = {0, 0, 0, 0, 0};
};

static void stage_begin(wbgpu_ctx* c, int stage) {
    if (!c->timing) return;
    while (c->ev_pool.size() < c->ev_next + 2) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        c->ev_pool.push_back(e);
    }
    c->ev_used.push_back({stage, (int)c->ev_next});
    cudaEventRecord(c->ev_pool[c->ev_next], c->stream);
}
static void stage_end(wbgpu_ctx* c) {
    if (!c->timing) return;
    cudaEventRecord(c->ev_pool[c->ev_next + 1], c->stream);
    c->ev_next += 2;
}
static void stage_collect(wbgpu_ctx* c) {
    if (!c->timing || c->ev_used.empty()) return;
    cudaStreamSynchronize(c->stream);
    for (auto& u : c->ev_used) {
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev_pool[u.second], c->ev_pool[u.second + 1]);
        c->stage_ms[u.first] += ms;
        c->stage_calls[u.first]++;
    }
    c->ev_used.clear();
    c->ev_next = 0;
}

static void free_plan(wbgpu_ctx* c) {
    cudaFree(c->d_table);
    cudaFree(c->d_tableH);
    c->d_tableH = nullptr;
    for (int d = 0; d < 3; d++) cudaFree(c->d_W[d]);
    cudaFree(c->d_Z); cudaFree(c->d_Y); cudaFree(c->d_X); cudaFree(c->d_U);
    cudaFree(c->d_E); cudaFree(c->d_evlabel); cudaFree(c->d_evval);
    cudaFree(c->d_dw); cudaFree(c->d_ew); cudaFree(c->d_tau); cudaFree(c->d_rot); cudaFree(c->d_hdr); cudaFree(c->d_Vh);
    c->d_Vh = nullptr;
    cudaFree(c->d_d0); cudaFree(c->d_e0);
    c->d_d0 = c->d_e0 = nullptr;
    cudaFree(c->d_nsweep); cudaFree(c->d_faillist); cudaFree(c->d_nfail);
    c->d_dw = c->d_ew = nullptr; c->d_tau = nullptr; c->d_rot = nullptr;
    c->d_hdr = c->d_nsweep = c->d_faillist = c->d_nfail = nullptr;
    c->d_table = nullptr;
    for (int d = 0; d < 3; d++) c->d_W[d] = nullptr;
    c->d_Z = c->d_Y = c->d_X = c->d_U = nullptr;
    c->d_E = c->d_evlabel = c->d_evval = nullptr;
    c->planned = false;
}

extern "C" const char* wbgpu_last_error(void) { return g_err.c_str(); }
extern "C" int wbgpu_version(void) { return 100; }
extern "C" int wbgpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

static int create_fill(wbgpu_ctx* c, int device, int nw, int nR, const int32_t* iRvec, const double* cRvec_shifted);

extern "C" int wbgpu_create(wbgpu_ctx** out, int device, int nw, int nR, const int32_t* iRvec,
                            const double* cRvec_shifted, double cell_volume, void* stream) {
    if (!out || !iRvec || !cRvec_shifted) return set_err("wbgpu_create: null pointer argument");
    if (nw < 1 || nw > 128) return set_err("wbgpu_create: num_wann=%d outside the supported range 1..128", nw);
    if (nR < 1) return set_err("wbgpu_create: nR=%d", nR);
    int ndev = wbgpu_device_count();
    if (ndev == 0) return set_err("wbgpu_create: no CUDA device available (there is no CPU fallback)");
    if (device < 0 || device >= ndev) return set_err("wbgpu_create: device %d out of range (have %d)", device, ndev);
    CK(cudaSetDevice(device));
    wbgpu_ctx* c = new wbgpu_ctx();
    c->device = device;
    c->stream = (cudaStream_t)stream;
    c->nw = nw;
    c->nR = nR;
    c->cell_volume = cell_volume;
    int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    for (int i = 0; i < nR; i++)
        for (int d = 0; d < 3; d++) {
            // the box is symmetric under R -> -R so that hermitisation can be done in R-space
            int a = abs(iRvec[3 * i + d]);
            hi[d] = std::max(hi[d], a);
            lo[d] = std::min(lo[d], -a);
        }
    c->rmin = make_int3(lo[0], lo[1], lo[2]);
    c->nbox = make_int3(hi[0] - lo[0] + 1, hi[1] - lo[1] + 1, hi[2] - lo[2] + 1);
    if (create_fill(c, device, nw, nR, iRvec, cRvec_shifted)) {   // nothing of a half-built context survives an error
        wbgpu_destroy(c);
        return 1;
    }
    *out = c;
    return 0;
}

static int create_fill(wbgpu_ctx* c, int device, int nw, int nR, const int32_t* iRvec, const double* cRvec_shifted) {
    CK(cudaMalloc(&c->d_iRvec, sizeof(int) * 3 * nR));
    CK(cudaMalloc(&c->d_T, sizeof(double) * 3 * (size_t)nR * nw * nw));
    CK(cudaMemcpy(c->d_iRvec, iRvec, sizeof(int) * 3 * nR, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_T, cRvec_shifted, sizeof(double) * 3 * (size_t)nR * nw * nw, cudaMemcpyHostToDevice));
    c->h_iRvec.assign(iRvec, iRvec + 3 * (size_t)nR);
    CK(cudaMalloc(&c->d_sweeps, 2 * sizeof(int)));   // [0] max Jacobi sweeps, [1] k-points re-solved by Jacobi
    CK(cudaDeviceGetAttribute(&c->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    return 0;
}

extern "C" int wbgpu_destroy(wbgpu_ctx* c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    free_plan(c);
    cudaFree(c->d_iRvec); cudaFree(c->d_T); cudaFree(c->d_sweeps);
    for (int k = 0; k < WBGPU_NKEYS; k++) cudaFree(c->d_XR[k]);
    cudaFree(c->d_xbar); cudaFree(c->d_mx); cudaFree(c->d_colwin); cudaFree(c->d_kent); cudaFree(c->d_kacc); cudaFree(c->d_shcJ); cudaFree(c->d_Ec);
    cudaFree(c->d_hist); cudaFree(c->d_cum); cudaFree(c->d_dK); cudaFree(c->d_weight); cudaFree(c->d_out); cudaFree(c->d_axes);
    delete c;
    return 0;
}

extern "C" int wbgpu_set_R_matrix(wbgpu_ctx* c, int key, const double* X_R, int ncart) {
    if (!c || !X_R) return set_err("wbgpu_set_R_matrix: null pointer argument");
    if (key < 0 || key >= WBGPU_NKEYS) return set_err("wbgpu_set_R_matrix: unknown key %d", key);
    int want = (key == WBGPU_HAM) ? 1 : (key == WBGPU_SA || key == WBGPU_SHA || key == WBGPU_SR || key == WBGPU_SHR) ? 9 : 3;
    if (ncart != want) return set_err("wbgpu_set_R_matrix: key %d needs ncart=%d, got %d", key, want, ncart);
    CK(cudaSetDevice(c->device));
    size_t bytes = sizeof(cplx) * (size_t)c->nR * c->nw * c->nw * ncart;
    if (!c->d_XR[key]) CK(cudaMalloc(&c->d_XR[key], bytes));
    CK(cudaMemcpy(c->d_XR[key], X_R, bytes, cudaMemcpyHostToDevice));
    c->planned = false;  // tables must be rebuilt
    return 0;
}

extern "C" int wbgpu_set_option(wbgpu_ctx* c, const char* name, int64_t value) {
    if (!c || !name) return set_err("wbgpu_set_option: null pointer argument");
    if (!strcmp(name, "eig_method")) { c->eig_method = (int)value; return 0; }
    if (!strcmp(name, "rotate_method")) { c->rotate_method = (int)value; return 0; }
    if (!strcmp(name, "fourier_method")) { c->fourier_method = (int)value; return 0; }
    if (!strcmp(name, "gemm_stack")) { c->gemm_stack = (int)value; return 0; }
    if (!strcmp(name, "rot_r2")) { c->rot_r2 = (int)value; return 0; }
    if (!strcmp(name, "rotate_trim")) { c->rotate_trim = (int)value; return 0; }
    if (!strcmp(name, "kubo_method")) { c->kubo_method = (int)value; return 0; }
    if (!strcmp(name, "dh_packed")) { c->dh_packed = (int)value; c->planned = false; return 0; }
    if (!strcmp(name, "timing")) {
        c->timing = (int)value;
        for (int i = 0; i < WBGPU_NSTAGES; i++) { c->stage_ms[i] = 0; c->stage_calls[i] = 0; }
        return 0;
    }
    return set_err("wbgpu_set_option: unknown option '%s'", name);
}

extern "C" int64_t wbgpu_kernel_launches(const wbgpu_ctx* c) { return c ? c->launches : 0; }
extern "C" int wbgpu_last_eig_sweeps(const wbgpu_ctx* c) { return c ? c->last_sweeps : 0; }
extern "C" int wbgpu_last_eig_resolved(const wbgpu_ctx* c) { return c ? c->last_resolved : 0; }

static int formula_rank(int f) {
    switch (f) {
        case WBGPU_IDENTITY: return 0;
        case WBGPU_OMEGA: case WBGPU_MORB_HPM: case WBGPU_SPIN: return 1;
        case WBGPU_VEL_OMEGA: case WBGPU_VEL_HPLUS: case WBGPU_VEL_SPIN: case WBGPU_VEL_VEL: case WBGPU_INV_MASS:
        case WBGPU_DER_OMEGA: case WBGPU_DER_SPIN: case WBGPU_OMEGA_S: case WBGPU_OMEGA_OMEGA: case WBGPU_DER_MORB:
        case WBGPU_OMEGA_HPLUS: return 2;
        case WBGPU_SHC_RYOO: case WBGPU_SHC_QIAO: case WBGPU_SHC_SIMPLE:   // SpinOmega
        case WBGPU_VEL_VEL_VEL: case WBGPU_MASS_VEL: case WBGPU_DER3E: return 3;
        case WBGPU_MASS_MASS: case WBGPU_VEL_MASS_VEL: return 4;
    }
    return -1;
}
static int formula_ncomp(int f) {
    int r = formula_rank(f), n = 1;
    for (int i = 0; i < r; i++) n *= 3;
    return n;
}
// formulae evaluated by a kernel of their own (wb_fsea.cuh): one event pass per spec
static bool formula_solo(int f) { return f >= WBGPU_SHC_RYOO && f < WBGPU_NFORMULA; }
// factor kinds of the FormulaProduct formulae (wb_fsea.cuh: 1 V, 2 InvMass, 3 Omega, 4 Spin, 5 DerSpin)
static bool formula_product(int f, WbProductSpec* P) {
    int k[3] = {0, 0, 0}, n = 0;
    switch (f) {
        case WBGPU_DER_SPIN: n = 1; k[0] = 5; break;
        case WBGPU_VEL_VEL_VEL: n = 3; k[0] = k[1] = k[2] = 1; break;
        case WBGPU_MASS_VEL: n = 2; k[0] = 2; k[1] = 1; break;
        case WBGPU_MASS_MASS: n = 2; k[0] = k[1] = 2; break;
        case WBGPU_VEL_MASS_VEL: n = 3; k[0] = 1; k[1] = 2; k[2] = 1; break;
        case WBGPU_OMEGA_S: n = 2; k[0] = 3; k[1] = 4; break;
        case WBGPU_OMEGA_OMEGA: n = 2; k[0] = k[1] = 3; break;
        case WBGPU_OMEGA_HPLUS: n = 2; k[0] = 3; k[1] = 6; break;
        default: return false;
    }
    if (P) { P->nf = n; for (int i = 0; i < 3; i++) P->kind[i] = k[i]; }
    return true;
}
static bool product_has(int f, int kind) {
    WbProductSpec P;
    if (!formula_product(f, &P)) return false;
    for (int i = 0; i < P.nf; i++)
        if (P.kind[i] == kind) return true;
    return false;
}
static int fder_extra(int fder) { return fder == 0 ? 0 : (fder <= 2 ? 1 : 2); }

extern "C" int64_t wbgpu_spec_size(const wbgpu_scan_spec* s) {
    if (!s || formula_rank(s->formula) < 0) return -1;
    return (int64_t)s->nEF * formula_ncomp(s->formula);
}

static WbWindow make_window(const wbgpu_scan_spec& s) {
    WbWindow w;
    int extra = fder_extra(s.fder);
    w.dEF = s.dEF;
    w.EFmin = s.Ef_first - extra * s.dEF;  // static.py:56-57
    w.EFmax = s.Ef_last + extra * s.dEF;
    w.nEFx = s.nEF + 2 * extra;
    w.degen_thresh = s.degen_thresh;
    w.degen_Kramers = s.degen_Kramers;
    w.sea = (s.fder == 0);
    w.Ebmin = w.Ebmax = nullptr;
    w.holes = 0;
    w.Emin_sea = -INFINITY;
    w.Emax_holes = INFINITY;
    return w;
}

// ------------------------------------------------------------------------------------------ plan
static int plan_impl(wbgpu_ctx* c, const int32_t NKFFT[3], uint32_t formula_mask, int external_terms, int64_t max_kpoints_per_launch);

extern "C" int wbgpu_plan(wbgpu_ctx* c, const int32_t NKFFT[3], uint32_t formula_mask, int external_terms,
                          int64_t max_kpoints_per_launch) {
    const int rc = plan_impl(c, NKFFT, formula_mask, external_terms, max_kpoints_per_launch);
    if (rc && c && !c->planned) {   // failed after the old plan was released: no half-allocated work space stays behind
        const std::string msg = g_err;
        cudaGetLastError();
        free_plan(c);
        c->planned = false;
        g_err = msg;
    }
    return rc;
}

static int plan_impl(wbgpu_ctx* c, const int32_t NKFFT[3], uint32_t formula_mask, int external_terms, int64_t max_kpoints_per_launch) {
    if (!c || !NKFFT) return set_err("wbgpu_plan: null pointer argument");
    for (int d = 0; d < 3; d++)
        if (NKFFT[d] < 1) return set_err("wbgpu_plan: NKFFT[%d]=%d", d, NKFFT[d]);
    if (!c->d_XR[WBGPU_HAM]) return set_err("wbgpu_plan: R-matrix 'Ham' is not set");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    free_plan(c);
    const int nw = c->nw;
    const uint32_t m = formula_mask;
    auto has = [&](int f) { return (m >> f) & 1u; };
    bool need_dH = has(WBGPU_OMEGA) || has(WBGPU_MORB_HPM) || has(WBGPU_VEL_OMEGA) || has(WBGPU_VEL_HPLUS) ||
                   has(WBGPU_VEL_SPIN) || has(WBGPU_KUBO) || has(WBGPU_VEL_VEL) || has(WBGPU_INV_MASS) || has(WBGPU_SHC_RYOO) ||
                   has(WBGPU_SHC_QIAO) || has(WBGPU_SHC_SIMPLE) || has(WBGPU_DER_OMEGA) || has(WBGPU_SHIFT_CURRENT);
    const bool shc = has(WBGPU_SHC_RYOO) || has(WBGPU_SHC_QIAO) || has(WBGPU_SHC_SIMPLE);
    bool prod_any = false, prod_mass = false, prod_omega = false, prod_spin = false;
    for (int f = WBGPU_DER_SPIN; f < WBGPU_NFORMULA; f++)
        if (has(f) && formula_product(f, nullptr)) {
            prod_any = true;
            prod_mass |= product_has(f, 2);
            prod_omega |= product_has(f, 3) || product_has(f, 6);
            prod_spin |= product_has(f, 4) || product_has(f, 5);
        }
    need_dH = need_dH || prod_any || has(WBGPU_DER3E) || has(WBGPU_DER_MORB);
    const bool der_om = has(WBGPU_DER_OMEGA) || has(WBGPU_DER_MORB);   // channels of DerOmega
    bool berry = has(WBGPU_OMEGA) || has(WBGPU_MORB_HPM) || has(WBGPU_VEL_OMEGA) || has(WBGPU_VEL_HPLUS) || has(WBGPU_KUBO);
    bool need_A = (berry || shc || der_om || prod_omega || has(WBGPU_SHIFT_CURRENT)) && external_terms;
    bool need_BC = (has(WBGPU_MORB_HPM) || has(WBGPU_VEL_HPLUS) || has(WBGPU_DER_MORB) || has(WBGPU_OMEGA_HPLUS)) && external_terms;
    bool need_S = has(WBGPU_SPIN) || has(WBGPU_VEL_SPIN) || shc || prod_spin;
    if (has(WBGPU_SHC_RYOO) && (!c->d_XR[WBGPU_SA] || !c->d_XR[WBGPU_SHA]))
        return set_err("wbgpu_plan: R-matrices 'SA','SHA' are not set (SHC_type='ryoo')");
    if (has(WBGPU_SHC_QIAO) && (!c->d_XR[WBGPU_SR] || !c->d_XR[WBGPU_SH] || !c->d_XR[WBGPU_SHR]))
        return set_err("wbgpu_plan: R-matrices 'SR','SH','SHR' are not set (SHC_type='qiao')");
    if (need_A && !c->d_XR[WBGPU_AA]) return set_err("wbgpu_plan: R-matrix 'AA' is not set (needed for external terms)");
    if (need_BC && (!c->d_XR[WBGPU_BB] || !c->d_XR[WBGPU_CC])) return set_err("wbgpu_plan: R-matrices 'BB','CC' are not set");
    if (need_S && !c->d_XR[WBGPU_SS]) return set_err("wbgpu_plan: R-matrix 'SS' is not set");

    // ---- is d_a H hermitian in R-space?  (then its three channels are transformed and stored as triangles)
    int dH_herm = 0;
    if (need_dH && c->dh_packed) {
        const size_t ncell0 = (size_t)c->nbox.x * c->nbox.y * c->nbox.z;
        std::vector<int> cellmap(ncell0, -1);
        for (int iR = 0; iR < c->nR; iR++) {
            const int* R = &c->h_iRvec[3 * (size_t)iR];
            size_t cell = ((size_t)(R[0] - c->rmin.x) * c->nbox.y + (R[1] - c->rmin.y)) * c->nbox.z + (R[2] - c->rmin.z);
            cellmap[cell] = (cellmap[cell] == -1) ? iR : -2;
        }
        int* d_map = nullptr;
        unsigned long long* d_res = nullptr;
        CK(cudaMalloc(&d_map, sizeof(int) * ncell0));
        CK(cudaMalloc(&d_res, 16));
        CK(cudaMemcpyAsync(d_map, cellmap.data(), sizeof(int) * ncell0, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemsetAsync(d_res, 0, 16, c->stream));
        long tot = (long)c->nR * nw * nw;
        wb_dH_herm_check_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, c->stream>>>(c->d_XR[WBGPU_HAM], c->d_T, c->d_iRvec, d_map, c->nR,
                                                                                   nw, c->rmin, c->nbox, d_res);
        c->launches++;
        double res[2] = {0, 0};
        CK(cudaMemcpyAsync(res, d_res, 16, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        cudaFree(d_map);
        cudaFree(d_res);
        dH_herm = (res[0] <= 1e-13 * res[1]) ? 1 : 0;
    }

    WbLayout L;
    L.nw = nw;
    L.ntri = nw * (nw + 1) / 2;
    L.dH_herm = dH_herm;
    int off = 0;
    auto take = [&](bool herm) { int o = off; off += herm ? L.ntri : nw * nw; return o; };
    L.off_H = take(true);
    for (int a = 0; a < 3; a++) L.off_dH[a] = need_dH ? take(dH_herm) : -1;
    for (int a = 0; a < 3; a++) L.off_A[a] = need_A ? take(true) : -1;
    bool need_O = need_A && (has(WBGPU_OMEGA) || has(WBGPU_MORB_HPM) || has(WBGPU_VEL_OMEGA) || has(WBGPU_VEL_HPLUS) ||
                             der_om || prod_omega);
    for (int a = 0; a < 3; a++) L.off_O[a] = need_O ? take(true) : -1;
    for (int a = 0; a < 3; a++) L.off_B[a] = need_BC ? take(false) : -1;
    for (int a = 0; a < 3; a++) L.off_C[a] = need_BC ? take(false) : -1;
    for (int a = 0; a < 3; a++) L.off_S[a] = need_S ? take(true) : -1;
    for (int a = 0; a < 6; a++)
        L.off_W[a] = (has(WBGPU_INV_MASS) || der_om || prod_mass || has(WBGPU_SHIFT_CURRENT) || has(WBGPU_DER3E))
                         ? take(dH_herm) : -1;
    for (int a = 0; a < 10; a++) L.off_W3[a] = has(WBGPU_DER3E) ? take(dH_herm) : -1;
    for (int a = 0; a < 9; a++) L.off_dS[a] = has(WBGPU_DER_SPIN) ? take(true) : -1;
    for (int a = 0; a < 9; a++) L.off_dA[a] = ((der_om || has(WBGPU_SHIFT_CURRENT)) && need_A) ? take(true) : -1;
    for (int a = 0; a < 9; a++) L.off_dO[a] = (der_om && need_A) ? take(true) : -1;
    for (int a = 0; a < 9; a++) L.off_dB[a] = (has(WBGPU_DER_MORB) && need_BC) ? take(false) : -1;
    for (int a = 0; a < 9; a++) L.off_dC[a] = (has(WBGPU_DER_MORB) && need_BC) ? take(false) : -1;
    // second comma-derivatives for Xbar(name, 2) of plug-in formulae: of the matrices that are set
    const bool x2 = has(WBGPU_XBAR_DER2);
    const bool x2A = x2 && external_terms && c->d_XR[WBGPU_AA], x2BC = x2 && external_terms && c->d_XR[WBGPU_BB] && c->d_XR[WBGPU_CC],
               x2S = x2 && c->d_XR[WBGPU_SS];
    for (int a = 0; a < 18; a++) L.off_d2A[a] = x2A ? take(true) : -1;
    for (int a = 0; a < 18; a++) L.off_d2O[a] = x2A ? take(true) : -1;
    for (int a = 0; a < 18; a++) L.off_d2S[a] = x2S ? take(true) : -1;
    for (int a = 0; a < 18; a++) L.off_d2B[a] = x2BC ? take(false) : -1;
    for (int a = 0; a < 18; a++) L.off_d2C[a] = x2BC ? take(false) : -1;
    for (int a = 0; a < 9; a++) L.off_SA[a] = has(WBGPU_SHC_RYOO) ? take(false) : -1;
    for (int a = 0; a < 9; a++) L.off_SHA[a] = has(WBGPU_SHC_RYOO) ? take(false) : -1;
    for (int a = 0; a < 9; a++) L.off_SR[a] = has(WBGPU_SHC_QIAO) ? take(false) : -1;
    for (int a = 0; a < 3; a++) L.off_SH[a] = has(WBGPU_SHC_QIAO) ? take(false) : -1;
    for (int a = 0; a < 9; a++) L.off_SHR[a] = has(WBGPU_SHC_QIAO) ? take(false) : -1;
    L.E = off;
    c->L = L;
    c->mask = m;
    c->external = external_terms;
    for (int d = 0; d < 3; d++) c->N[d] = NKFFT[d];
    c->nk_block = (long)NKFFT[0] * NKFFT[1] * NKFFT[2];

    // R-space table
    size_t ncell = (size_t)c->nbox.x * c->nbox.y * c->nbox.z;
    CK(cudaMalloc(&c->d_table, sizeof(cplx) * ncell * L.E));
    CK(cudaMemsetAsync(c->d_table, 0, sizeof(cplx) * ncell * L.E, c->stream));
    WbRInputs in;
    in.Ham = c->d_XR[WBGPU_HAM];
    in.AA = (need_A || x2A) ? c->d_XR[WBGPU_AA] : nullptr;
    in.BB = (need_BC || x2BC) ? c->d_XR[WBGPU_BB] : nullptr;
    in.CC = (need_BC || x2BC) ? c->d_XR[WBGPU_CC] : nullptr;
    in.SS = (need_S || x2S) ? c->d_XR[WBGPU_SS] : nullptr;
    in.SA = has(WBGPU_SHC_RYOO) ? c->d_XR[WBGPU_SA] : nullptr;
    in.SHA = has(WBGPU_SHC_RYOO) ? c->d_XR[WBGPU_SHA] : nullptr;
    in.SR = has(WBGPU_SHC_QIAO) ? c->d_XR[WBGPU_SR] : nullptr;
    in.SH = has(WBGPU_SHC_QIAO) ? c->d_XR[WBGPU_SH] : nullptr;
    in.SHR = has(WBGPU_SHC_QIAO) ? c->d_XR[WBGPU_SHR] : nullptr;
    in.T = c->d_T;
    in.iRvec = c->d_iRvec;
    long total = (long)c->nR * nw * nw;
    wb_build_rtable_kernel<<<(unsigned)((total + 127) / 128), 128, 0, c->stream>>>(in, L, c->nR, c->rmin, c->nbox, c->d_table);
    c->launches++;
    CK(cudaGetLastError());

    // workspace: bytes per k-point of one launch
    const int n0 = c->nbox.x, n1 = c->nbox.y, n2 = c->nbox.z;
    const double N0 = NKFFT[0], N1 = NKFFT[1];
    double per_k = 16.0 * L.E * (1.0 + n0 / N0 + (double)n0 * n1 / (N0 * N1)) + 16.0 * nw * nw + 8.0 * nw * 11 + 64;
    if (nw <= 32) per_k += 16.0 * (2 * nw * nw + 32) + 4.0 * (6 * nw + 12) + 32.0 * nw + 16;  // QL rotation stream etc.
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    // at most 45 % of the free memory and 48 GB: the sub-batched buffers of the scan calls (rotated matrices, Kubo
    // entries, eigensolver work space, formula scratch: <= 3 .. 6 GB each) and other contexts need the rest
    long kmax = (long)(std::min(0.45 * (double)free_b, 48.0e9) / per_k);
    long want = max_kpoints_per_launch > 0 ? (long)max_kpoints_per_launch : 262144;
    kmax = std::min(kmax, want);
    long nb = std::max(1L, kmax / c->nk_block);
    nb = std::min(nb, 65535L);
    c->nb_max = (int)nb;
    size_t nkl = (size_t)nb * c->nk_block;
    if ((double)nkl * per_k > 0.9 * (double)free_b)
        return set_err("wbgpu_plan: one K-block of %ld k-points needs %.1f GB, only %.1f GB free", c->nk_block,
                       nkl * per_k / 1e9, free_b / 1e9);
    int nbx[3] = {n0, n1, n2};
    for (int d = 0; d < 3; d++) CK(cudaMalloc(&c->d_W[d], sizeof(cplx) * nb * NKFFT[d] * nbx[d]));
    CK(cudaMalloc(&c->d_Z, sizeof(cplx) * nb * (size_t)n0 * n1 * NKFFT[2] * L.E));
    CK(cudaMalloc(&c->d_Y, sizeof(cplx) * nb * (size_t)n0 * NKFFT[1] * NKFFT[2] * L.E));
    CK(cudaMalloc(&c->d_X, sizeof(cplx) * nkl * L.E));
    CK(cudaMalloc(&c->d_U, sizeof(cplx) * nkl * nw * nw));
    CK(cudaMalloc(&c->d_E, sizeof(double) * nkl * nw));
    CK(cudaMalloc(&c->d_evlabel, sizeof(double) * nkl * nw));
    {
        int ncmax = 1;
        int sum = 0;
        for (int f = 1; f < WBGPU_NFORMULA; f++)
            if (formula_rank(f) >= 0 && ((m >> f) & 1u)) sum += formula_ncomp(f);
        ncmax = std::max(ncmax, sum);
        c->ev_ncmax = ncmax;
        CK(cudaMalloc(&c->d_evval, sizeof(double) * nkl * nw * ncmax));
    }
    {
        c->capR = 2 * nw * nw + 32;
        c->capS = (6 * nw + 8 + 3) / 4 * 4;  // 16-byte granular (bulk copies)
        if (nw <= 32) c->eig_chunk = std::min<long>((long)nkl, 262144);  // thread-per-k QL needs many k-points in flight
        else {
            // rotation stream + reflector-major Householder vectors: <= 6 GB per sub-batch
            double per = 16.0 * c->capR + 16.0 * nw * nw + 4.0 * c->capS + 40.0 * nw;
            c->eig_chunk = std::max(1L, std::min<long>((long)nkl, (long)(6.0e9 / per)));
        }
        size_t ch = (size_t)c->eig_chunk;
        CK(cudaMalloc(&c->d_dw, sizeof(double) * ch * nw));
        CK(cudaMalloc(&c->d_ew, sizeof(double) * ch * nw));
        CK(cudaMalloc(&c->d_tau, sizeof(cplx) * ch * nw));
        // (also holds the tridiagonal eigenvector matrices of wb_eigh_tf.cuh: groups of 32 k-points x nw x 2 ceil(nw/2) doubles)
        CK(cudaMalloc(&c->d_rot, std::max(sizeof(double2) * ch * c->capR, sizeof(double) * ((ch + 31) / 32) * 64 * nw * ((nw + 1) / 2))));
        CK(cudaMalloc(&c->d_hdr, sizeof(int) * ch * c->capS));
        CK(cudaMalloc(&c->d_nsweep, sizeof(int) * ch));
        CK(cudaMalloc(&c->d_faillist, sizeof(int) * ch));
        CK(cudaMalloc(&c->d_nfail, sizeof(int)));
        if (nw > 32) {
            CK(cudaMalloc(&c->d_Vh, sizeof(cplx) * ch * nw * nw));
            CK(cudaMalloc(&c->d_d0, sizeof(double) * ch * nw));
            CK(cudaMalloc(&c->d_e0, sizeof(double) * ch * nw));
        }
    }
    CK(cudaStreamSynchronize(c->stream));
    c->planned = true;
    return 0;
}

// ------------------------------------------------------------------------------------------ stages
static int pick_kc(int N) {
    for (int kc = 10; kc >= 4; kc--)
        if (N % kc == 0) return kc;
    if (N <= 10) return N;
    return 8;
}

template <int KC>
static void launch_axis(wbgpu_ctx* c, const cplx* in, cplx* out, const cplx* W, int n, int N, long S, int outer,
                        long in_bs, long out_bs, int nb) {
    int nchunk = (N + KC - 1) / KC;
    dim3 grid((unsigned)((S + 255) / 256), (unsigned)(outer * nchunk), (unsigned)nb);
    wb_axis_dft_kernel<KC><<<grid, 256, sizeof(cplx) * KC * n, c->stream>>>(in, out, W, n, N, S, outer, in_bs, out_bs);
    c->launches++;
}

static int axis_dft(wbgpu_ctx* c, const cplx* in, cplx* out, const cplx* W, int n, int N, long S, int outer,
                    long in_bs, long out_bs, int nb) {
    if ((long)outer * ((N + 3) / 4) > 65535) return set_err("axis_dft: grid.y overflow");
    switch (pick_kc(N)) {
#define WB_CASE(K) case K: launch_axis<K>(c, in, out, W, n, N, S, outer, in_bs, out_bs, nb); break;
        WB_CASE(1) WB_CASE(2) WB_CASE(3) WB_CASE(4) WB_CASE(5) WB_CASE(6) WB_CASE(7) WB_CASE(8) WB_CASE(9) WB_CASE(10)
#undef WB_CASE
        default: return set_err("axis_dft: bad KC");
    }
    CK(cudaGetLastError());
    return 0;
}

template <int KC, int TB>
static int launch_fused10(wbgpu_ctx* c, int nb, size_t smem) {
    constexpr int KSPLIT = 256 / TB;
    const int n0 = c->nbox.x, n1 = c->nbox.y;
    const int* N = c->N;
    const long S2 = (long)N[2] * c->L.E;
    CK(cudaFuncSetAttribute(wb_axis10_fused_kernel<KC, TB, KSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((S2 + TB - 1) / TB), 1, (unsigned)nb);
    wb_axis10_fused_kernel<KC, TB, KSPLIT><<<grid, TB * KSPLIT, smem, c->stream>>>(c->d_Z, c->d_X, c->d_W[1], c->d_W[0], n0, n1, N[0], N[1], S2,
                                                              (long)n0 * n1 * S2, (long)N[0] * N[1] * S2);
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

template <int NR0, int TB, int KSPLIT, int MINB, int UK>
static int launch_fused10_y(wbgpu_ctx* c, int nb) {
    const int n0 = c->nbox.x, n1 = c->nbox.y;
    const int* N = c->N;
    const long S2 = (long)N[2] * c->L.E;
    const size_t smem = sizeof(cplx) * ((size_t)n0 * n1 * TB + (size_t)N[1] * n1 + (size_t)N[0] * NR0);
    if ((int)smem > c->smem_optin) return -1;
    auto kern = wb_axis10_fused_y_kernel<NR0, TB, KSPLIT, MINB, UK>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    dim3 grid((unsigned)((S2 + TB - 1) / TB), 1, (unsigned)nb);
    kern<<<grid, TB * KSPLIT, smem, c->stream>>>(c->d_Z, c->d_X, c->d_W[1], c->d_W[0], n0, n1, N[0], N[1], S2,
                                                 (long)n0 * n1 * S2, (long)N[0] * N[1] * S2);
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

// returns -1 when the fused kernel does not apply
static int fused_axis10(wbgpu_ctx* c, int nb) {
    // default: the low-register kernel (partial sums over r1 in registers; 80 registers, three 256-thread CTAs per SM):
    // measured 2.34 ms against 2.88 ms per 256k k-points for the KC-accumulator kernel below (fourier_method 2)
    if (c->fourier_method != 2 && c->nbox.x <= 8) {
        int rc = c->nbox.x <= 4 ? launch_fused10_y<4, 64, 4, 3, 5>(c, nb)
               : c->nbox.x <= 6 ? launch_fused10_y<6, 64, 4, 3, 5>(c, nb) : launch_fused10_y<8, 64, 4, 3, 5>(c, nb);
        if (rc >= 0) return rc;
    }
    const int n0 = c->nbox.x, n1 = c->nbox.y;
    const int* N = c->N;
    const int KC = N[0] <= 8 ? 8 : N[0] <= 12 ? 12 : N[0] <= 16 ? 16 : 20;
    const int nchunk = (N[0] + KC - 1) / KC;
    auto smem_for = [&](int TB) { return sizeof(cplx) * ((size_t)n0 * n1 * TB + (size_t)N[1] * n1 + (size_t)nchunk * KC * n0); };
    int TB = 128;
    while (TB >= 32 && smem_for(TB) > 110 * 1024) TB /= 2;
    if (TB < 32) return -1;
    size_t smem = smem_for(TB);
#define WB_F10(KCV)                                                     \
    if (KC == KCV) {                                                    \
        if (TB == 128) return launch_fused10<KCV, 128>(c, nb, smem);    \
        if (TB == 64) return launch_fused10<KCV, 64>(c, nb, smem);      \
        return launch_fused10<KCV, 32>(c, nb, smem);                    \
    }
    WB_F10(8) WB_F10(12) WB_F10(16) WB_F10(20)
#undef WB_F10
    return -1;
}

// R->k for `nb` K-blocks whose shifts are dK_dev[nb][3]: fills c->d_X[nb][nk][E]
static int run_fourier(wbgpu_ctx* c, const double* dK_dev, int nb) {
    const int n[3] = {c->nbox.x, c->nbox.y, c->nbox.z};
    const int rm[3] = {c->rmin.x, c->rmin.y, c->rmin.z};
    const int* N = c->N;
    const long E = c->L.E;
    for (int d = 0; d < 3; d++) {
        long total = (long)nb * N[d] * n[d];
        wb_twiddle_kernel<<<(unsigned)((total + 127) / 128), 128, 0, c->stream>>>(dK_dev, nb, N[d], n[d], rm[d], d, c->d_W[d]);
        c->launches++;
    }
    CK(cudaGetLastError());
    // axis 2: table[r0][r1][r2][e] -> Z[b][r0][r1][k2][e]
    if (axis_dft(c, c->d_table, c->d_Z, c->d_W[2], n[2], N[2], E, n[0] * n[1], 0, (long)n[0] * n[1] * N[2] * E, nb)) return 1;
    // axes 1 and 0 fused: Z -> X[b][k0][k1][k2][e] without the intermediate Y (falls back to two passes when the
    // Z tile of even 32 inner indices does not fit shared memory)
    if (c->fourier_method != 1) {
        int rc = fused_axis10(c, nb);
        if (rc >= 0) return rc;
    }
    // axis 1: -> Y[b][r0][k1][k2][e]
    if (axis_dft(c, c->d_Z, c->d_Y, c->d_W[1], n[1], N[1], (long)N[2] * E, n[0], (long)n[0] * n[1] * N[2] * E,
                 (long)n[0] * N[1] * N[2] * E, nb)) return 1;
    // axis 0: -> X[b][k0][k1][k2][e]
    if (axis_dft(c, c->d_Y, c->d_X, c->d_W[0], n[0], N[0], (long)N[1] * N[2] * E, 1, (long)n[0] * N[1] * N[2] * E,
                 (long)N[0] * N[1] * N[2] * E, nb)) return 1;
    return 0;
}

__global__ void wb_count_resolved_kernel(const int* __restrict__ nlist, int* __restrict__ total) { *total += *nlist; }

static int launch_jacobi(wbgpu_ctx* c, long k0, long nk, bool want_U, const int* list, const int* nlist, long nblk_cap) {
    const int nw = c->nw;
    if (nlist) wb_count_resolved_kernel<<<1, 1, 0, c->stream>>>(nlist, c->d_sweeps + 1);
    constexpr int WARPS = 4;
    int npair = (nw + 1) / 2;
    size_t smem = sizeof(cplx) * WARPS * (size_t)(2 * nw * (nw + 1) + 2 * npair + nw);
    if ((int)smem > c->smem_optin) return set_err("eigh(Jacobi): num_wann=%d needs %zu B shared memory", nw, smem);
    CK(cudaFuncSetAttribute(wb_eigh_jacobi_kernel<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long nblk = std::min((nk + WARPS - 1) / WARPS, nblk_cap);
    wb_eigh_jacobi_kernel<WARPS><<<(unsigned)nblk, WARPS * 32, smem, c->stream>>>(c->d_X, c->L, k0, nk, c->d_E,
                                                                             want_U ? c->d_U : nullptr, c->d_sweeps, list, nlist);
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

template <int NWP, bool EXACT>
static int launch_ql(wbgpu_ctx* c, long k0, long nk, bool want_U) {
    const int nw = c->nw;
    constexpr int WARPS = 4;
    CK(cudaMemsetAsync(c->d_nfail, 0, sizeof(int), c->stream));
    // eig_method 0: tridiagonalisation, then eigenvalues + twisted-factorisation eigenvectors with a thread per matrix and
    // the back-transformation with a lane per (matrix, vector) (wb_eigh_tf.cuh; 4 <= nw <= 24); 2 / 3: the rotation-stream
    // QL kernels below (3 = one k-point per warp in the reduction); 4: as 0 with the thread-per-matrix reduction
    // (wb_eigh_tpm.cuh, nw <= 20; measured slower than the two-k-points-per-warp kernel)
    const bool tf = (c->eig_method == 0 || c->eig_method == 4 || c->eig_method == 5) && nw >= 4 && nw <= 24;
    int tpm = (c->eig_method == 4) ? wb_launch_tridiag_tpm(nw, c->d_X, c->L, k0, nk, c->d_dw, c->d_ew, c->d_tau, c->d_U, c->stream)
            : (c->eig_method == 5) ? wb_launch_tridiag_tpm2(nw, c->d_X, c->L, k0, nk, c->d_dw, c->d_ew, c->d_tau, c->d_U, c->stream) : -1;
    if (tpm > 0) return set_err("CUDA error %s launching the thread-per-matrix tridiagonalisation", cudaGetErrorName((cudaError_t)tpm));
    if (tpm == 0) {
    } else if constexpr (EXACT && NWP > 16 && NWP <= 18) {
        if (c->eig_method != 3)   // two k-points per warp
            wb_tridiag2_kernel<NWP, WARPS><<<(unsigned)((nk + 2 * WARPS - 1) / (2 * WARPS)), WARPS * 32, 0, c->stream>>>(
                c->d_X, c->L, k0, nk, c->d_dw, c->d_ew, c->d_tau, c->d_U);
        else
            wb_tridiag_kernel<NWP, WARPS, EXACT><<<(unsigned)((nk + WARPS - 1) / WARPS), WARPS * 32, 0, c->stream>>>(
                c->d_X, c->L, k0, nk, c->d_dw, c->d_ew, c->d_tau, c->d_U);
    } else {
        wb_tridiag_kernel<NWP, WARPS, EXACT><<<(unsigned)((nk + WARPS - 1) / WARPS), WARPS * 32, 0, c->stream>>>(
            c->d_X, c->L, k0, nk, c->d_dw, c->d_ew, c->d_tau, c->d_U);
    }
    CK(cudaGetLastError());
    if (tf) {
        // (the eigenvector matrix Z of the tridiagonal, nw^2 doubles per k-point, borrows the rotation-stream buffer)
        int rc = wb_launch_trideig(nw, want_U, k0, nk, c->d_dw, c->d_ew, c->d_E, (double*)c->d_rot, c->d_faillist, c->d_nfail,
                                   c->stream);
        if (rc) return set_err("CUDA error %s launching the tridiagonal eigensolver", cudaGetErrorName((cudaError_t)rc));
        c->launches += 2;
        if (want_U) {
            rc = wb_launch_backtransform(nw, k0, nk, (const double*)c->d_rot, c->d_tau, c->d_U, c->stream);
            if (rc) return set_err("CUDA error %s launching the back-transformation", cudaGetErrorName((cudaError_t)rc));
            c->launches++;
        }
        // multiple eigenvalues beyond what the twisted factorisation resolves, unconverged QL: Jacobi (normally none)
        return launch_jacobi(c, k0, nk, want_U, c->d_faillist, c->d_nfail, 148);
    }
    CK(cudaGetLastError());
    if (nw <= 24) {
        constexpr int NT2 = 128;
        wb_tql_kernel<NT2><<<(unsigned)((nk + NT2 - 1) / NT2), NT2, sizeof(double) * 2 * nw * NT2, c->stream>>>(
            nw, nk, c->d_dw, c->d_ew, c->d_rot, c->capR, c->d_hdr, c->capS, c->d_nsweep);
    } else {
        constexpr int NT2 = 64;
        wb_tql_kernel<NT2><<<(unsigned)((nk + NT2 - 1) / NT2), NT2, sizeof(double) * 2 * nw * NT2, c->stream>>>(
            nw, nk, c->d_dw, c->d_ew, c->d_rot, c->capR, c->d_hdr, c->capS, c->d_nsweep);
    }
    CK(cudaGetLastError());
    if (!want_U) {   // eigenvalues only: sort the QL output, no replay / back-transformation
        wb_eig_sort_kernel<<<(unsigned)((nk + 127) / 128), 128, 0, c->stream>>>(nw, k0, nk, c->d_dw, c->d_nsweep, c->d_E,
                                                                             c->d_faillist, c->d_nfail);
        c->launches += 3;
        CK(cudaGetLastError());
        return launch_jacobi(c, k0, nk, false, c->d_faillist, c->d_nfail, 148);
    }
    constexpr int EW = 2;  // warps per CTA of the eigenvector kernel (19.6 KB of shared memory per warp at nw = 18)
    size_t smem3 = sizeof(cplx) * EW * (size_t)wb_eigvec_smem_per_warp(nw, c->capR, c->capS);
    if ((int)smem3 > c->smem_optin) return set_err("eigh(QL): num_wann=%d needs %zu B shared memory", nw, smem3);
    CK(cudaFuncSetAttribute(wb_eigvec_kernel<NWP, EW, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
    wb_eigvec_kernel<NWP, EW, EXACT><<<(unsigned)((nk + EW - 1) / EW), EW * 32, smem3, c->stream>>>(
        nw, k0, nk, c->d_dw, c->d_tau, c->d_rot, c->capR, c->d_hdr, c->capS, c->d_nsweep, c->d_E, c->d_U, c->d_faillist,
        c->d_nfail);
    c->launches += 3;
    CK(cudaGetLastError());
    // k-points whose QL iteration did not fit the stream: re-solve with Jacobi (normally none)
    return launch_jacobi(c, k0, nk, true, c->d_faillist, c->d_nfail, 148);
}

// nw > 32: CTA-per-k-point tridiagonalisation and back-transformation around the thread-per-k-point QL
static int launch_ql_large(wbgpu_ctx* c, long k0, long nk, bool want_U) {
    const int nw = c->nw;
    CK(cudaMemsetAsync(c->d_nfail, 0, sizeof(int), c->stream));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    constexpr int NT1 = 256;
    size_t smem1 = wb_tridiag_cta_smem_bytes(nw);
    if ((int)smem1 > c->smem_optin) return set_err("eigh: num_wann=%d needs %zu B shared memory", nw, smem1);
    CK(cudaFuncSetAttribute(wb_tridiag_cta_kernel<NT1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    int per_sm1 = std::max(1, std::min(4, (int)((size_t)c->smem_optin / smem1)));
    wb_tridiag_cta_kernel<NT1><<<(unsigned)std::min(nk, (long)sms * per_sm1), NT1, smem1, c->stream>>>(
        c->d_X, c->L, k0, nk, c->d_dw, c->d_ew, c->d_tau, c->d_Vh);
    CK(cudaGetLastError());
    // eig_method 2 = eigenvectors by replaying the QL rotations only; default: twisted factorisation first (needs T as reduced)
    const bool tf = want_U && c->eig_method != 2;
    if (tf) {
        CK(cudaMemcpyAsync(c->d_d0, c->d_dw, sizeof(double) * nk * nw, cudaMemcpyDeviceToDevice, c->stream));
        CK(cudaMemcpyAsync(c->d_e0, c->d_ew, sizeof(double) * nk * nw, cudaMemcpyDeviceToDevice, c->stream));
    }
    constexpr int NT2 = 32;
    size_t smem2 = sizeof(double) * 2 * nw * NT2;
    CK(cudaFuncSetAttribute(wb_tql_kernel<NT2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    wb_tql_kernel<NT2><<<(unsigned)((nk + NT2 - 1) / NT2), NT2, smem2, c->stream>>>(nw, nk, c->d_dw, c->d_ew, c->d_rot, c->capR,
                                                                                 c->d_hdr, c->capS, c->d_nsweep);
    CK(cudaGetLastError());
    constexpr int NT3 = 256;   // replay: threads 0 .. nw-1 = rows; back-transformation: 32 eigenvectors x 8 slices
    size_t smem3 = wb_eigvec_cta_smem_bytes(nw, NT3 / 8);
    if ((int)smem3 > c->smem_optin) return set_err("eigh: num_wann=%d needs %zu B shared memory", nw, smem3);
    int per_sm3 = std::max(1, std::min(4, (int)((size_t)c->smem_optin / smem3)));
    // EMAX = elements of an eigenvector per thread of the back-transformation (8 threads per vector)
#define WB_EIGVEC_CTA(EM)                                                                                                     \
    do {                                                                                                                      \
        CK(cudaFuncSetAttribute(wb_eigvec_cta_kernel<NT3, EM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));     \
        wb_eigvec_cta_kernel<NT3, EM><<<(unsigned)std::min(nk, (long)sms * per_sm3), NT3, smem3, c->stream>>>(                \
            nw, k0, nk, c->d_dw, c->d_tau, c->d_Vh, c->d_rot, c->capR, c->d_hdr, c->capS, c->d_nsweep, want_U ? 1 : 0, c->d_E, \
            c->d_U, c->d_nfail, tf ? c->d_d0 : nullptr, tf ? c->d_e0 : nullptr, c->d_sweeps + 1);                             \
    } while (0)
    if (nw <= 48) WB_EIGVEC_CTA(6);
    else if (nw <= 64) WB_EIGVEC_CTA(8);
    else if (nw <= 96) WB_EIGVEC_CTA(12);
    else WB_EIGVEC_CTA(16);
#undef WB_EIGVEC_CTA
    c->launches += 3;
    CK(cudaGetLastError());
    // there is no second solver for these sizes: a QL iteration that did not converge is an error
    int nfail = 0;
    CK(cudaMemcpyAsync(&nfail, c->d_nfail, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (nfail) return set_err("eigh: the QL iteration did not converge for %d k-point(s)", nfail);
    return 0;
}

static int run_eigh(wbgpu_ctx* c, long nk, bool want_U) {
    const int nw = c->nw;
    CK(cudaMemsetAsync(c->d_sweeps, 0, 2 * sizeof(int), c->stream));
    if (nw > 32) {
        if (c->eig_method == 1) return launch_jacobi(c, 0, nk, want_U, nullptr, nullptr, 148L * 64);
        for (long k0 = 0; k0 < nk; k0 += c->eig_chunk)
            if (launch_ql_large(c, k0, std::min(c->eig_chunk, nk - k0), want_U)) return 1;
        return 0;
    }
    bool use_ql = (c->eig_method != 1);  // 2 = QL, 3 = QL with the one-k-point-per-warp reduction
    if (!use_ql) return launch_jacobi(c, 0, nk, want_U, nullptr, nullptr, 148L * 64);
    for (long k0 = 0; k0 < nk; k0 += c->eig_chunk) {
        long n = std::min(c->eig_chunk, nk - k0);
        int rc;
        if (nw == 18) rc = launch_ql<18, true>(c, k0, n, want_U);
        else if (nw == 16) rc = launch_ql<16, true>(c, k0, n, want_U);
        else if (nw == 24) rc = launch_ql<24, true>(c, k0, n, want_U);
        else if (nw == 32) rc = launch_ql<32, true>(c, k0, n, want_U);
        else if (nw <= 8) rc = launch_ql<8, false>(c, k0, n, want_U);
        else if (nw <= 16) rc = launch_ql<16, false>(c, k0, n, want_U);
        else if (nw <= 24) rc = launch_ql<24, false>(c, k0, n, want_U);
        else rc = launch_ql<32, false>(c, k0, n, want_U);
        if (rc) return rc;
    }
    return 0;
}

// Specs that share the Fermi window, the degeneracy rules and the term flags are evaluated from ONE pass over
// the rotated matrices ("event group"): e.g. AHC + Morb (Omega, Morb_Hpm, Omega) or the three Fermi-surface
// tensors of BASELINE config 3.
struct EvGroup {
    WbWindow win;
    WbEventLayout ev;
    std::vector<int> specs;
    bool solo = false;   // a formula with an event kernel of its own
};

static bool same_window(const WbWindow& a, const WbWindow& b) {
    return a.EFmin == b.EFmin && a.EFmax == b.EFmax && a.dEF == b.dEF && a.degen_thresh == b.degen_thresh &&
           a.degen_Kramers == b.degen_Kramers && a.sea == b.sea && a.nEFx == b.nEFx && a.Ebmin == b.Ebmin &&
           a.holes == b.holes && a.Emin_sea == b.Emin_sea && a.Emax_holes == b.Emax_holes;
}

static std::vector<EvGroup> make_groups(const wbgpu_scan_spec* specs, int nspec, bool tetra = false) {
    std::vector<EvGroup> groups;
    for (int i = 0; i < nspec; i++) {
        const wbgpu_scan_spec& s = specs[i];
        WbWindow w = make_window(s);
        if (tetra) {   // hole_like / Emin / Emax act in the tetrahedron method only (data_K.py:172-185)
            w.holes = s.tetra_flags & 1;
            if (s.tetra_flags & 2) w.Emin_sea = s.tetra_Emin;
            if (s.tetra_flags & 4) w.Emax_holes = s.tetra_Emax;
        }
        bool ident = (s.formula == WBGPU_IDENTITY);
        int found = -1;
        for (size_t g = 0; g < groups.size() && !formula_solo(s.formula); g++) {
            EvGroup& G = groups[g];
            bool g_ident = (G.ev.mask == 1);
            if (G.solo || g_ident != ident || !same_window(G.win, w)) continue;
            if (!ident && (G.ev.internal_terms != s.internal_terms || G.ev.external_terms != s.external_terms)) continue;
            found = (int)g;
            break;
        }
        if (found < 0) {
            EvGroup G;
            G.win = w;
            G.ev.mask = 0;
            G.ev.NC = 0;
            for (int f = 0; f < 32; f++) G.ev.off[f] = 0;
            G.solo = formula_solo(s.formula);
            G.ev.internal_terms = s.internal_terms;
            G.ev.external_terms = s.external_terms;
            groups.push_back(G);
            found = (int)groups.size() - 1;
        }
        EvGroup& G = groups[found];
        if (!((G.ev.mask >> s.formula) & 1)) {
            G.ev.mask |= 1 << s.formula;
            G.ev.off[s.formula] = G.ev.NC;
            G.ev.NC += formula_ncomp(s.formula);
        }
        G.specs.push_back(i);
    }
    return groups;
}

// fused DMMA rotation + formula kernel (translation units wb_rotate_mma_*.cu); returns -1 when it does not cover the request
static int launch_mma_events(wbgpu_ctx* c, const EvGroup& G, long nk) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    const int rc = wb_launch_mma_events(c->nw, c->rotate_trim != 0, c->rot_r2, c->d_X, c->L, nk, c->d_E, c->d_U, G.win, G.ev,
                                        c->d_evlabel, c->d_evval, c->smem_optin, sms, c->stream);
    if (rc > 0) return set_err("CUDA error %s launching the fused rotation kernel", cudaGetErrorName((cudaError_t)rc));
    if (rc == 0) c->launches++;
    return rc;
}

static int ensure(double** p, size_t* cap, size_t need);

template <int NTL, int KC>
static int launch_gemm(wbgpu_ctx* c, const WbChanList& ch, long k0, long n, const int2* colwin) {
    const int nw = c->nw;
    size_t smem = wb_gemm_smem_bytes<NTL, KC>(nw);
    if ((int)smem > c->smem_optin) return set_err("rotate(gemm): num_wann=%d needs %zu B shared memory", nw, smem);
    CK(cudaFuncSetAttribute(wb_rotate_gemm_kernel<NTL, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((nw + 8 * NTL - 1) / (8 * NTL)), (unsigned)ch.n, (unsigned)std::min(n, 16384L));
    wb_rotate_gemm_kernel<NTL, KC><<<grid, 128, smem, c->stream>>>(c->d_X + (size_t)k0 * c->L.E, (long)c->L.E, ch, nw, n,
                                                                 c->d_U + (size_t)k0 * nw * nw, (cplx*)c->d_xbar, colwin);
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

template <int NTL, int KC, int CG>
static int launch_gemm_cg(wbgpu_ctx* c, const WbChanList& ch, long k0, long n, const int2* colwin) {
    const int nw = c->nw;
    size_t smem = wb_gemm_cg_smem_bytes<NTL, KC, CG>(nw);
    if ((int)smem > c->smem_optin) return set_err("rotate(gemm): num_wann=%d needs %zu B shared memory", nw, smem);
    CK(cudaFuncSetAttribute(wb_rotate_gemm_cg_kernel<NTL, KC, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(1, (unsigned)((ch.n + CG - 1) / CG), (unsigned)std::min(n, 16384L));
    wb_rotate_gemm_cg_kernel<NTL, KC, CG><<<grid, 128, smem, c->stream>>>(c->d_X + (size_t)k0 * c->L.E, (long)c->L.E, ch, nw, n,
                                                                        c->d_U + (size_t)k0 * nw * nw, (cplx*)c->d_xbar, colwin);
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

// U^dagger X U of the listed channels for k-points [k0, k0 + n) -> c->d_xbar[n][ch.n][nw][nw]; with `colwin` (per k-point
// column range, nw <= 32) only the rows and columns of that range are formed
static int rotate_gemm(wbgpu_ctx* c, const WbChanList& ch, long k0, long n, const int2* colwin = nullptr) {
    const int nw = c->nw;
    if (c->gemm_stack) {   // several channels per CTA for small matrices (option gemm_stack, default on)
        // (NTL, KC, CG) by size; measured per 128k k-points: Te (24 WF, 18 channels) 33.6 -> 26.5 ms with (3, 8, 2) against
        // one channel per CTA -- (3, 24, 2) is slower, 3 CTAs per SM; Fe (18 WF, 15 channels) 33.7 -> 25.0 ms with (3, 20, 3)
        if (nw <= 8) return launch_gemm_cg<1, 8, 4>(c, ch, k0, n, colwin);
        if (nw <= 16) return launch_gemm_cg<2, 8, 4>(c, ch, k0, n, colwin);
        if (nw <= 20) return launch_gemm_cg<3, 20, 3>(c, ch, k0, n, colwin);
        if (nw <= 24) return launch_gemm_cg<3, 8, 2>(c, ch, k0, n, colwin);
        if (nw <= 32) return launch_gemm_cg<4, 16, 2>(c, ch, k0, n, colwin);
    }
    if (nw <= 8) return launch_gemm<1, 8>(c, ch, k0, n, colwin);
    if (nw <= 16) return launch_gemm<2, 8>(c, ch, k0, n, colwin);
    if (nw <= 24) return launch_gemm<3, 8>(c, ch, k0, n, colwin);
    return launch_gemm<4, 16>(c, ch, k0, n, colwin);
}

static long xbar_chunk(wbgpu_ctx* c, int nch, long nk) {
    double per_k = 16.0 * nch * c->nw * c->nw;
    long chunk = (long)(3.0e9 / per_k);
    return std::max(1L, std::min(chunk, nk));
}

// generic path: batched DMMA rotation to global memory, then the formula kernel
static int run_events_xbar(wbgpu_ctx* c, const EvGroup& G, long nk) {
    const int nw = c->nw;
    const WbLayout& L = c->L;
    WbNeeds need = wb_needs(G.ev.mask, G.ev.external_terms);
    WbChanList ch;
    ch.n = 0;
    auto add3 = [&](const int* offs, bool herm) {
        for (int a = 0; a < 3; a++) { ch.off[ch.n] = offs[a]; ch.herm[ch.n] = herm; ch.n++; }
    };
    if (need.V) add3(L.off_dH, L.dH_herm);
    if (need.A) add3(L.off_A, true);
    if (need.B) add3(L.off_B, false);
    if (need.Oblk || need.Odiag) add3(L.off_O, true);
    if (need.Cblk || need.Cdiag) add3(L.off_C, false);
    if (need.Sblk || need.Sdiag) add3(L.off_S, true);
    if (need.Wdiag)
        for (int a = 0; a < 6; a++) { ch.off[ch.n] = L.off_W[a]; ch.herm[ch.n] = L.dH_herm; ch.n++; }
    const long chunk = xbar_chunk(c, ch.n, nk);
    if (ensure(&c->d_xbar, &c->xbar_cap, sizeof(cplx) * (size_t)chunk * ch.n * nw * nw)) return 1;
    constexpr int NT = 128;
    const long nblk_max = 148L * 8;
    if (ensure(&c->d_mx, &c->mx_cap, sizeof(double) * (size_t)nblk_max * 3 * nw * nw)) return 1;
    size_t smem = wb_xbar_events_smem_bytes(nw);
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(wb_events_xbar_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // the formula stage reads rows / columns of the bands of a band group only (wb_rotate_gemm.cuh, TRIM); not so the
    // non-additive Morb evaluation (all pairs), nor a non-hermitian d_a H (its rows are not the mirror of its columns)
    const bool trim = c->rotate_trim && !((G.ev.mask >> 2) & 1) && (L.dH_herm || !need.V);
    if (trim && ensure(&c->d_colwin, &c->colwin_cap, sizeof(int2) * (size_t)chunk)) return 1;
    for (long k0 = 0; k0 < nk; k0 += chunk) {
        long n = std::min(chunk, nk - k0);
        WbWindow wloc = G.win;
        if (wloc.Ebmin) { wloc.Ebmin += k0 * nw; wloc.Ebmax += k0 * nw; }
        if (trim) {
            constexpr int WW = 4;
            const size_t smw = sizeof(double) * WW * (2 * (size_t)nw + (nw + 3) / 4 * 2);
            wb_band_window_kernel<WW><<<(unsigned)std::min((n + WW - 1) / WW, 148L * 16), WW * 32, smw, c->stream>>>(
                c->d_E + k0 * nw, nw, n, wloc, (int2*)c->d_colwin);
            c->launches++;
            CK(cudaGetLastError());
        }
        if (rotate_gemm(c, ch, k0, n, trim ? (const int2*)c->d_colwin : nullptr)) return 1;
        long nblk = std::min(n, nblk_max);
        wb_events_xbar_kernel<NT><<<(unsigned)nblk, NT, smem, c->stream>>>((const cplx*)c->d_xbar, ch.n, nw, n, c->d_E + k0 * nw, wloc, G.ev,
                                                                       c->d_mx, c->d_evlabel + k0 * nw,
                                                                       c->d_evval + (size_t)k0 * nw * G.ev.NC);
        c->launches++;
        CK(cudaGetLastError());
    }
    return 0;
}

// SpinOmega / DerOmega: batched DMMA rotation to global memory, then the formula's own event kernel (wb_fsea.cuh)
static int run_events_solo(wbgpu_ctx* c, const EvGroup& G, long nk) {
    const int nw = c->nw, n2 = nw * nw;
    const WbLayout& L = c->L;
    int formula = -1;
    for (int f = 0; f < WBGPU_NFORMULA; f++)
        if ((G.ev.mask >> f) & 1) formula = f;
    const bool ext = G.ev.external_terms;
    WbChanList ch;
    ch.n = 0;
    auto addn = [&](const int* offs, int n, int herm) {
        const int first = ch.n;
        for (int a = 0; a < n; a++) { ch.off[ch.n] = offs[a]; ch.herm[ch.n] = herm; ch.n++; }
        return first;
    };
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    constexpr int NT = 256;
    if (formula == WBGPU_DER_OMEGA) {
        constexpr int NT = 128;   // (shadows the 256 of the other kernels: 3 CTAs per SM, less idling in the small band groups)
        if (L.off_dH[0] < 0 || L.off_W[0] < 0 || (ext && (L.off_A[0] < 0 || L.off_O[0] < 0 || L.off_dA[0] < 0)))
            return set_err("scan: the plan does not hold the channels of DerOmega");
        WbDerOmegaChans C;
        C.iV = addn(L.off_dH, 3, L.dH_herm);
        C.iW = addn(L.off_W, 6, L.dH_herm);
        C.iA = C.iO = C.idA = C.idO = 0;
        if (ext) {
            C.iA = addn(L.off_A, 3, 1);
            C.iO = addn(L.off_O, 3, 1);
            C.idA = addn(L.off_dA, 9, 1);
            C.idO = addn(L.off_dO, 9, 1);
        }
        const long chunk = xbar_chunk(c, ch.n, nk);
        if (ensure(&c->d_xbar, &c->xbar_cap, sizeof(cplx) * (size_t)chunk * ch.n * n2)) return 1;
        const size_t per_cta = sizeof(cplx) * wb_deromega_scratch_elems(nw);
        const long nblk_max = std::max(32L, std::min((long)sms * 6, (long)(1.5e9 / (double)per_cta)));
        if (ensure(&c->d_mx, &c->mx_cap, per_cta * (size_t)nblk_max)) return 1;
        size_t smem = wb_fsea_smem_bytes(nw, 9);
        // V_a and D_a of the k-point staged in shared memory when two CTAs per SM still fit (wb_fsea_stage)
        const int stage = (wb_fsea_stage_offset(nw, 9) + wb_fsea_stage_bytes(nw) <= (size_t)c->smem_optin / 3 - 1024) ? 1 : 0;
        if (stage) smem = wb_fsea_stage_offset(nw, 9) + wb_fsea_stage_bytes(nw);
        if (smem > 48 * 1024)
            CK(cudaFuncSetAttribute(stage ? wb_deromega_events_kernel<NT, true> : wb_deromega_events_kernel<NT, false>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (long k0 = 0; k0 < nk; k0 += chunk) {
            const long n = std::min(chunk, nk - k0);
            if (rotate_gemm(c, ch, k0, n)) return 1;
            WbWindow wloc = G.win;
            if (wloc.Ebmin) { wloc.Ebmin += k0 * nw; wloc.Ebmax += k0 * nw; }
            if (stage)
                wb_deromega_events_kernel<NT, true><<<(unsigned)std::min(n, nblk_max), NT, smem, c->stream>>>(
                (const cplx*)c->d_xbar, ch.n, nw, n, c->d_E + k0 * nw, wloc, C, G.ev.internal_terms, ext ? 1 : 0, (cplx*)c->d_mx,
                c->d_evlabel + k0 * nw, c->d_evval + (size_t)k0 * nw * 9);
            else
                wb_deromega_events_kernel<NT, false><<<(unsigned)std::min(n, nblk_max), NT, smem, c->stream>>>(
                (const cplx*)c->d_xbar, ch.n, nw, n, c->d_E + k0 * nw, wloc, C, G.ev.internal_terms, ext ? 1 : 0, (cplx*)c->d_mx,
                c->d_evlabel + k0 * nw, c->d_evval + (size_t)k0 * nw * 9);
            c->launches++;
            CK(cudaGetLastError());
        }
        return 0;
    }
    if (formula == WBGPU_DER_MORB) {
        if (L.off_dH[0] < 0 || L.off_W[0] < 0 ||
            (ext && (L.off_A[0] < 0 || L.off_O[0] < 0 || L.off_dA[0] < 0 || L.off_B[0] < 0 || L.off_dB[0] < 0)))
            return set_err("scan: the plan does not hold the channels of DerMorb");
        WbDerMorbChans C;
        C.iA = C.iO = C.idA = C.idO = C.iB = C.idB = C.iC = C.idC = 0;
        C.iV = addn(L.off_dH, 3, L.dH_herm);
        C.iW = addn(L.off_W, 6, L.dH_herm);
        if (ext) {
            C.iA = addn(L.off_A, 3, 1); C.iO = addn(L.off_O, 3, 1); C.idA = addn(L.off_dA, 9, 1); C.idO = addn(L.off_dO, 9, 1);
            C.iB = addn(L.off_B, 3, 0); C.idB = addn(L.off_dB, 9, 0); C.iC = addn(L.off_C, 3, 0); C.idC = addn(L.off_dC, 9, 0);
        }
        const long chunk = xbar_chunk(c, ch.n, nk);
        if (ensure(&c->d_xbar, &c->xbar_cap, sizeof(cplx) * (size_t)chunk * ch.n * n2)) return 1;
        const size_t per_cta = sizeof(cplx) * wb_dermorb_scratch_elems(nw);
        const long nblk_max = std::max(32L, std::min((long)sms * 4, (long)(1.5e9 / (double)per_cta)));
        if (ensure(&c->d_mx, &c->mx_cap, per_cta * (size_t)nblk_max)) return 1;
        size_t smem = wb_dermorb_smem_bytes(nw);
        // V_a and D_a of the k-point staged in shared memory when two CTAs per SM still fit (wb_fsea_stage)
        const int stage_off = (int)((smem + 15) / 16 * 16);
        const bool stage = stage_off + wb_fsea_stage_bytes(nw) <= (size_t)c->smem_optin / 2 - 1024;
        if (stage) smem = stage_off + wb_fsea_stage_bytes(nw);
        if (smem > 48 * 1024)
            CK(cudaFuncSetAttribute(stage ? wb_dermorb_events_kernel<NT, true> : wb_dermorb_events_kernel<NT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (long k0 = 0; k0 < nk; k0 += chunk) {
            const long n = std::min(chunk, nk - k0);
            if (rotate_gemm(c, ch, k0, n)) return 1;
            WbWindow wloc = G.win;
            if (wloc.Ebmin) { wloc.Ebmin += k0 * nw; wloc.Ebmax += k0 * nw; }
            if (stage)
                wb_dermorb_events_kernel<NT, true><<<(unsigned)std::min(n, nblk_max), NT, smem, c->stream>>>(
                (const cplx*)c->d_xbar, ch.n, nw, n, c->d_E + k0 * nw, wloc, C, G.ev.internal_terms, ext ? 1 : 0, stage_off, (cplx*)c->d_mx,
                c->d_evlabel + k0 * nw, c->d_evval + (size_t)k0 * nw * 9);
            else
                wb_dermorb_events_kernel<NT, false><<<(unsigned)std::min(n, nblk_max), NT, smem, c->stream>>>(
                (const cplx*)c->d_xbar, ch.n, nw, n, c->d_E + k0 * nw, wloc, C, G.ev.internal_terms, ext ? 1 : 0, stage_off, (cplx*)c->d_mx,
                c->d_evlabel + k0 * nw, c->d_evval + (size_t)k0 * nw * 9);
            c->launches++;
            CK(cudaGetLastError());
        }
        return 0;
    }
    if (formula == WBGPU_DER3E) {
        if (L.off_dH[0] < 0 || L.off_W[0] < 0 || L.off_W3[0] < 0) return set_err("scan: the plan does not hold the channels of Der3E");
        const int iV = addn(L.off_dH, 3, L.dH_herm), iW = addn(L.off_W, 6, L.dH_herm), iW3 = addn(L.off_W3, 10, L.dH_herm);
        const long chunk = xbar_chunk(c, ch.n, nk);
        if (ensure(&c->d_xbar, &c->xbar_cap, sizeof(cplx) * (size_t)chunk * ch.n * n2)) return 1;
        const size_t per_cta = sizeof(cplx) * wb_deromega_scratch_elems(nw);
        const long nblk_max = std::max(32L, std::min((long)sms * 4, (long)(1.5e9 / (double)per_cta)));
        if (ensure(&c->d_mx, &c->mx_cap, per_cta * (size_t)nblk_max)) return 1;
        size_t smem = wb_fsea_smem_bytes(nw, 27);
        // V_a and D_a of the k-point staged in shared memory when two CTAs per SM still fit (wb_fsea_stage)
        const int stage_off = (int)((smem + 15) / 16 * 16);
        const bool stage = stage_off + wb_fsea_stage_bytes(nw) <= (size_t)c->smem_optin / 2 - 1024;
        if (stage) smem = stage_off + wb_fsea_stage_bytes(nw);
        if (smem > 48 * 1024)
            CK(cudaFuncSetAttribute(stage ? wb_der3e_events_kernel<NT, true> : wb_der3e_events_kernel<NT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (long k0 = 0; k0 < nk; k0 += chunk) {
            const long n = std::min(chunk, nk - k0);
            if (rotate_gemm(c, ch, k0, n)) return 1;
            WbWindow wloc = G.win;
            if (wloc.Ebmin) { wloc.Ebmin += k0 * nw; wloc.Ebmax += k0 * nw; }
            if (stage)
                wb_der3e_events_kernel<NT, true><<<(unsigned)std::min(n, nblk_max), NT, smem, c->stream>>>(
                (const cplx*)c->d_xbar, ch.n, nw, n, c->d_E + k0 * nw, wloc, iV, iW, iW3, stage_off, (cplx*)c->d_mx, c->d_evlabel + k0 * nw,
                c->d_evval + (size_t)k0 * nw * 27);
            else
                wb_der3e_events_kernel<NT, false><<<(unsigned)std::min(n, nblk_max), NT, smem, c->stream>>>(
                (const cplx*)c->d_xbar, ch.n, nw, n, c->d_E + k0 * nw, wloc, iV, iW, iW3, stage_off, (cplx*)c->d_mx, c->d_evlabel + k0 * nw,
                c->d_evval + (size_t)k0 * nw * 27);
            c->launches++;
            CK(cudaGetLastError());
        }
        return 0;
    }
    WbProductSpec P;
    if (formula_product(formula, &P)) {
        const bool hasM = product_has(formula, 2), hasO = product_has(formula, 3) || product_has(formula, 6),
                   hasS = product_has(formula, 4) || product_has(formula, 5), hasH = product_has(formula, 6);
        if (hasH && ext && (L.off_B[0] < 0 || L.off_C[0] < 0)) return set_err("scan: the plan does not hold the channels of formula %d", formula);
        if (L.off_dH[0] < 0 || (hasM && L.off_W[0] < 0) || (hasO && ext && (L.off_A[0] < 0 || L.off_O[0] < 0)) ||
            (hasS && L.off_S[0] < 0) || (formula == WBGPU_DER_SPIN && L.off_dS[0] < 0))
            return set_err("scan: the plan does not hold the channels of formula %d", formula);
        P.iW = P.iA = P.iO = P.iS = P.idS = P.iB = P.iC = 0;
        P.iV = addn(L.off_dH, 3, L.dH_herm);
        if (hasM) P.iW = addn(L.off_W, 6, L.dH_herm);
        if (hasO && ext) { P.iA = addn(L.off_A, 3, 1); P.iO = addn(L.off_O, 3, 1); }
        if (hasS) P.iS = addn(L.off_S, 3, 1);
        if (formula == WBGPU_DER_SPIN) P.idS = addn(L.off_dS, 9, 1);
        if (hasH && ext) { P.iB = addn(L.off_B, 3, 0); P.iC = addn(L.off_C, 3, 0); }
        const int NC = formula_ncomp(formula);
        const long chunk = xbar_chunk(c, ch.n, nk);
        if (ensure(&c->d_xbar, &c->xbar_cap, sizeof(cplx) * (size_t)chunk * ch.n * n2)) return 1;
        const size_t per_cta = sizeof(cplx) * wb_product_scratch_elems(nw);
        const long nblk_max = std::max(32L, std::min((long)sms * 4, (long)(1.5e9 / (double)per_cta)));
        if (ensure(&c->d_mx, &c->mx_cap, per_cta * (size_t)nblk_max)) return 1;
        size_t smem = wb_fsea_smem_bytes(nw, 0);
        // V_a and D_a of the k-point staged in shared memory when two CTAs per SM still fit (wb_fsea_stage)
        const int stage_off = (int)((smem + 15) / 16 * 16);
        const bool stage = stage_off + wb_fsea_stage_bytes(nw) <= (size_t)c->smem_optin / 2 - 1024;
        if (stage) smem = stage_off + wb_fsea_stage_bytes(nw);
        if (smem > 48 * 1024)
            CK(cudaFuncSetAttribute(stage ? wb_product_events_kernel<NT, true> : wb_product_events_kernel<NT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (long k0 = 0; k0 < nk; k0 += chunk) {
            const long n = std::min(chunk, nk - k0);
            if (rotate_gemm(c, ch, k0, n)) return 1;
            WbWindow wloc = G.win;
            if (wloc.Ebmin) { wloc.Ebmin += k0 * nw; wloc.Ebmax += k0 * nw; }
            if (stage)
                wb_product_events_kernel<NT, true><<<(unsigned)std::min(n, nblk_max), NT, smem, c->stream>>>(
                (const cplx*)c->d_xbar, ch.n, nw, n, c->d_E + k0 * nw, wloc, P, G.ev.internal_terms, ext ? 1 : 0, stage_off, (cplx*)c->d_mx,
                c->d_evlabel + k0 * nw, c->d_evval + (size_t)k0 * nw * NC);
            else
                wb_product_events_kernel<NT, false><<<(unsigned)std::min(n, nblk_max), NT, smem, c->stream>>>(
                (const cplx*)c->d_xbar, ch.n, nw, n, c->d_E + k0 * nw, wloc, P, G.ev.internal_terms, ext ? 1 : 0, stage_off, (cplx*)c->d_mx,
                c->d_evlabel + k0 * nw, c->d_evval + (size_t)k0 * nw * NC);
            c->launches++;
            CK(cudaGetLastError());
        }
        return 0;
    }
    // SpinOmega
    if (L.off_dH[0] < 0 || L.off_S[0] < 0 || (ext && L.off_A[0] < 0))
        return set_err("scan: the plan does not hold the channels of SpinOmega");
    WbShcChans sc;
    sc.type = formula; sc.iA = sc.iX1 = sc.iX2 = sc.iX3 = -1;
    sc.iV = addn(L.off_dH, 3, L.dH_herm);
    if (ext) sc.iA = addn(L.off_A, 3, 1);
    sc.iS = addn(L.off_S, 3, 1);
    if (formula == WBGPU_SHC_RYOO) { sc.iX1 = addn(L.off_SA, 9, 0); sc.iX2 = addn(L.off_SHA, 9, 0); }
    if (formula == WBGPU_SHC_QIAO) { sc.iX1 = addn(L.off_SR, 9, 0); sc.iX2 = addn(L.off_SH, 3, 0); sc.iX3 = addn(L.off_SHR, 9, 0); }
    const long chunk = xbar_chunk(c, ch.n, nk);
    if (ensure(&c->d_xbar, &c->xbar_cap, sizeof(cplx) * (size_t)chunk * ch.n * n2)) return 1;
    if (ensure(&c->d_shcJ, &c->shcJ_cap, sizeof(cplx) * (size_t)chunk * 9 * n2)) return 1;
    const size_t smem = wb_fsea_smem_bytes(nw, 27);
    if (smem > 48 * 1024)
        CK(cudaFuncSetAttribute(wb_spinomega_events_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (long k0 = 0; k0 < nk; k0 += chunk) {
        const long n = std::min(chunk, nk - k0);
        if (rotate_gemm(c, ch, k0, n)) return 1;
        wb_shc_spinvel_kernel<256><<<(unsigned)std::min(n, (long)sms * 8), 256, sizeof(double) * nw, c->stream>>>(
            (const cplx*)c->d_xbar, ch.n, nw, n, c->d_E + k0 * nw, sc, ext ? 1 : 0, (cplx*)c->d_shcJ);
        WbWindow wloc = G.win;
        if (wloc.Ebmin) { wloc.Ebmin += k0 * nw; wloc.Ebmax += k0 * nw; }
        wb_spinomega_events_kernel<NT><<<(unsigned)std::min(n, (long)sms * 8), NT, smem, c->stream>>>(
            (const cplx*)c->d_xbar, ch.n, nw, n, c->d_E + k0 * nw, wloc, sc.iV, sc.iA, (const cplx*)c->d_shcJ, c->d_evlabel + k0 * nw,
            c->d_evval + (size_t)k0 * nw * 27);
        c->launches += 2;
        CK(cudaGetLastError());
    }
    return 0;
}

static int run_events(wbgpu_ctx* c, const EvGroup& G, long nk) {
    const int nw = c->nw;
    const WbWindow& win = G.win;
    if (G.ev.NC > c->ev_ncmax) return set_err("scan: event buffer too small (plan/scan formula mismatch)");
    if (G.ev.mask == 1) {
        int per = ((nw * (2 * 8 + 2 * 2)) + 7) / 8 * 8;
        int nt = nw <= 32 ? 64 : 32;
        if ((size_t)per * nt > 48 * 1024)
            CK(cudaFuncSetAttribute(wb_identity_events_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, per * nt));
        wb_identity_events_kernel<<<(unsigned)((nk + nt - 1) / nt), nt, (size_t)per * nt, c->stream>>>(c->d_E, nw, nk, win, c->d_evlabel,
                                                                                                 c->d_evval);
        c->launches++;
        CK(cudaGetLastError());
        return 0;
    }
    if (G.solo) return run_events_solo(c, G, nk);
    WbNeeds need = wb_needs(G.ev.mask, G.ev.external_terms);
    const WbLayout& L = c->L;
    if ((need.V && L.off_dH[0] < 0) || (need.A && L.off_A[0] < 0) || (need.B && L.off_B[0] < 0) ||
        ((need.Oblk || need.Odiag) && L.off_O[0] < 0) || ((need.Cblk || need.Cdiag) && L.off_C[0] < 0) ||
        ((need.Sblk || need.Sdiag) && L.off_S[0] < 0) || (need.Wdiag && L.off_W[0] < 0))
        return set_err("scan: the plan does not hold the channels that formula mask 0x%x needs", G.ev.mask);
    if (G.win.Ebmin) return run_events_xbar(c, G, nk);   // tetrahedron band groups: size-generic path
    // compile-time-NW tensor-core kernel (Omega and/or Morb_Hpm)
    if (c->rotate_method == 0 || c->rotate_method == 3) {
        int rc = launch_mma_events(c, G, nk);
        if (rc >= 0) return rc;
        if (c->rotate_method == 3) return set_err("rotate: the NW-templated DMMA kernel does not cover num_wann=%d, mask 0x%x", nw, G.ev.mask);
    }
    long nblk = std::min(nk, 148L * 32);
    bool omega_only = (G.ev.mask == (1 << WBGPU_OMEGA));
    bool dmma = omega_only && (nw <= 20) && (c->rotate_method != 1) && (c->rotate_method != 4);
    if (c->rotate_method == 2 && !dmma) return set_err("rotate: the DMMA kernel covers Omega with num_wann <= 20 only");
    if (c->rotate_method == 4 || (c->rotate_method == 0 && !dmma)) return run_events_xbar(c, G, nk);
    if (dmma) {
        WbFormulaFlags fl{WBGPU_OMEGA, G.ev.internal_terms, G.ev.external_terms};
#define WB_DMMA_CASE(KS, MT2)                                                                                         \
    {                                                                                                                 \
        size_t smem = wb_dmma_smem_bytes(nw, KS);                                                                     \
        CK(cudaFuncSetAttribute(wb_omega_events_dmma_kernel<KS, MT2>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                (int)smem));                                                                          \
        wb_omega_events_dmma_kernel<KS, MT2><<<(unsigned)nblk, 128, smem, c->stream>>>(c->d_X, c->L, nk, c->d_E, c->d_U, \
                                                                                   win, fl, c->d_evlabel, c->d_evval); \
    }
        if (nw <= 4) WB_DMMA_CASE(2, 1)
        else if (nw <= 8) WB_DMMA_CASE(4, 2)
        else if (nw <= 12) WB_DMMA_CASE(6, 3)
        else if (nw <= 16) WB_DMMA_CASE(8, 4)
        else if (nw <= 18) WB_DMMA_CASE(9, 5)
        else WB_DMMA_CASE(10, 5)
#undef WB_DMMA_CASE
    } else {
        constexpr int NT = 128;
        size_t smem = wb_generic_smem_bytes(nw, G.ev.mask, G.ev.external_terms);
        if ((int)smem > c->smem_optin)
            return set_err("rotate(generic): num_wann=%d with formula mask 0x%x needs %zu B shared memory (max %d)", nw,
                           G.ev.mask, smem, c->smem_optin);
        CK(cudaFuncSetAttribute(wb_events_generic_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        wb_events_generic_kernel<NT><<<(unsigned)nblk, NT, smem, c->stream>>>(c->d_X, c->L, nk, c->d_E, c->d_U, win, G.ev,
                                                                          c->d_evlabel, c->d_evval);
    }
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

static int ensure(double** p, size_t* cap, size_t need) {
    if (*cap >= need) return 0;
    cudaFree(*p);
    *p = nullptr;
    CK(cudaMalloc(p, need));
    *cap = need;
    return 0;
}

static int check_spec(const wbgpu_ctx* c, const wbgpu_scan_spec& s) {
    if (formula_rank(s.formula) < 0) return set_err("scan: unknown formula %d", s.formula);
    if (s.fder < 0 || s.fder > 3) return set_err("scan: Derivatives d^%df/dE^%d is not implemented", s.fder, s.fder);
    if (s.nEF < 1) return set_err("scan: nEF=%d", s.nEF);
    if (!(s.dEF > 0)) return set_err("scan: dEF must be positive (Efermi must be an increasing uniform grid)");
    if (!((c->mask >> s.formula) & 1u)) return set_err("scan: formula %d was not declared in wbgpu_plan", s.formula);
    if (s.use_select && s.fder == 0) return set_err("scan: Selection of bands for Fermi sea is not implemented");
    return 0;
}

// out_dev: [total] (weighted sum over the K-blocks), or [nblocks][total] in the per-K-block mode (weights ignored)
static int static_scan_impl(wbgpu_ctx* c, int nblocks, const double* dK_dev, const double* weight_dev,
                            const wbgpu_scan_spec* specs, int nspec, double* out_dev, bool per_block) {
    if (!c || !dK_dev || !specs || !out_dev || (!per_block && !weight_dev)) return set_err("wbgpu_static_scan: null pointer argument");
    if (!c->planned) return set_err("wbgpu_static_scan: call wbgpu_plan first");
    if (nblocks < 0 || nspec < 1) return set_err("wbgpu_static_scan: nblocks=%d nspec=%d", nblocks, nspec);
    CK(cudaSetDevice(c->device));
    std::vector<size_t> hoff(nspec + 1, 0), ooffs(nspec + 1, 0);
    size_t cum_max = 0;
    for (int i = 0; i < nspec; i++) {
        if (check_spec(c, specs[i])) return 1;
        WbWindow w = make_window(specs[i]);
        size_t sz = (size_t)(w.nEFx + 1) * formula_ncomp(specs[i].formula);
        hoff[i + 1] = hoff[i] + sz;
        ooffs[i + 1] = ooffs[i] + (size_t)specs[i].nEF * formula_ncomp(specs[i].formula);
        cum_max = std::max(cum_max, sz);
    }
    const size_t nhist = per_block ? (size_t)c->nb_max : 1;   // histograms (and prefix-sum scratch rows) held at once
    size_t need = sizeof(double) * nhist * (hoff[nspec] + cum_max);
    if (ensure(&c->d_hist, &c->hist_cap, need)) return 1;
    CK(cudaMemsetAsync(c->d_hist, 0, sizeof(double) * nhist * hoff[nspec], c->stream));
    double* d_cum = c->d_hist + nhist * hoff[nspec];
    const int nw = c->nw;
    bool need_U = false;
    for (int i = 0; i < nspec; i++) need_U |= (specs[i].formula != WBGPU_IDENTITY);
    std::vector<EvGroup> groups = make_groups(specs, nspec);
    auto finalize = [&](int nb, double* out) {   // nb histograms -> out[nb][total] (nb = 1: the weighted sum)
        for (int i = 0; i < nspec; i++) {
            const wbgpu_scan_spec& s = specs[i];
            WbWindow w = make_window(s);
            int ncomp = formula_ncomp(s.formula);
            double scale = s.factor / (c->cell_volume * (double)c->nk_block);
            dim3 grid((unsigned)ncomp, (unsigned)nb);
            wb_scan_finalize_kernel<<<grid, 256, 0, c->stream>>>(c->d_hist + hoff[i], d_cum, ncomp, w.nEFx, s.nEF, s.fder, s.dEF,
                                                             scale, out + ooffs[i], (long)hoff[nspec], (long)cum_max,
                                                             (long)ooffs[nspec]);
            c->launches++;
        }
    };

    for (int b0 = 0; b0 < nblocks; b0 += c->nb_max) {
        int nb = std::min(c->nb_max, nblocks - b0);
        long nk = (long)nb * c->nk_block;
        stage_begin(c, WBGPU_STAGE_FOURIER);
        if (run_fourier(c, dK_dev + 3 * (size_t)b0, nb)) return 1;
        stage_end(c);
        stage_begin(c, WBGPU_STAGE_EIGH);
        if (run_eigh(c, nk, need_U)) return 1;
        stage_end(c);
        for (const EvGroup& G : groups) {
            stage_begin(c, G.ev.mask == 1 ? WBGPU_STAGE_IDENTITY : WBGPU_STAGE_ROTATE);
            if (run_events(c, G, nk)) return 1;
            stage_end(c);
            stage_begin(c, WBGPU_STAGE_SCAN);
            for (int i : G.specs) {
                const wbgpu_scan_spec& s = specs[i];
                int ncomp = formula_ncomp(s.formula);
                size_t hbytes = sizeof(double) * (size_t)(G.win.nEFx + 1) * ncomp;
                int use_smem = hbytes <= 96 * 1024;
                if (use_smem)
                    CK(cudaFuncSetAttribute(wb_scan_accumulate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
                long nslots = nk * nw;
                if (per_block) {
                    dim3 grid((unsigned)std::max(1L, std::min((c->nk_block * nw + 255) / 256, (long)(148 * 2 / std::max(nb, 1) + 1))),
                              (unsigned)nb);
                    wb_scan_accumulate_kernel<<<grid, 256, use_smem ? hbytes : 0, c->stream>>>(
                        c->d_evlabel, c->d_evval + G.ev.off[s.formula], G.ev.NC, nslots, (int)(c->nk_block * nw), nullptr, ncomp,
                        G.win, c->d_hist + hoff[i], use_smem, (long)hoff[nspec], s.use_select ? c->d_E : nullptr, nw,
                        (unsigned long long)s.select_mask[0], (unsigned long long)s.select_mask[1]);
                } else {
                    long nblk = std::min((nslots + 255) / 256, 148L * 2);
                    wb_scan_accumulate_kernel<<<(unsigned)nblk, 256, use_smem ? hbytes : 0, c->stream>>>(
                        c->d_evlabel, c->d_evval + G.ev.off[s.formula], G.ev.NC, nslots, (int)(c->nk_block * nw), weight_dev + b0,
                        ncomp, G.win, c->d_hist + hoff[i], use_smem, 0L, s.use_select ? c->d_E : nullptr, nw,
                        (unsigned long long)s.select_mask[0], (unsigned long long)s.select_mask[1]);
                }
                c->launches++;
            }
            stage_end(c);
            CK(cudaGetLastError());
        }
        if (per_block) {
            finalize(nb, out_dev + (size_t)b0 * ooffs[nspec]);
            CK(cudaMemsetAsync(c->d_hist, 0, sizeof(double) * nhist * hoff[nspec], c->stream));
        }
    }
    if (!per_block) finalize(1, out_dev);
    CK(cudaGetLastError());
    stage_collect(c);
    return 0;
}

extern "C" int wbgpu_static_scan_dev(wbgpu_ctx* c, int nblocks, const double* dK_dev, const double* weight_dev,
                                     const wbgpu_scan_spec* specs, int nspec, double* out_dev) {
    return static_scan_impl(c, nblocks, dK_dev, weight_dev, specs, nspec, out_dev, false);
}

extern "C" int wbgpu_stage_times(const wbgpu_ctx* c, double* ms, int64_t* calls) {
    if (!c || !ms || !calls) return set_err("wbgpu_stage_times: null pointer argument");
    for (int i = 0; i < WBGPU_NSTAGES; i++) { ms[i] = c->stage_ms[i]; calls[i] = c->stage_calls[i]; }
    return 0;
}

// FP64 peak probes: kind 0 = DFMA (vector pipe), 1 = DMMA (mma.sync.m8n8k4.f64).  Returns TFLOP/s.
extern "C" int wbgpu_fp64_peak(int device, int kind, double* tflops) {
    if (!tflops) return set_err("wbgpu_fp64_peak: null pointer argument");
    if (wbgpu_device_count() == 0) return set_err("wbgpu_fp64_peak: no CUDA device available");
    CK(cudaSetDevice(device));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    double* d_out;
    CK(cudaMalloc(&d_out, 8));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int iters = 4096, blocks = sms * 8, threads = 256;
    constexpr int CH = 8;
    double best = 0;
    for (int rep = 0; rep < 6; rep++) {
        CK(cudaEventRecord(e0));
        if (kind == 0) wb_dfma_probe_kernel<CH><<<blocks, threads>>>(d_out, iters, 1.0000001, 1e-9);
        else wb_dmma_probe_kernel<CH><<<blocks, threads>>>(d_out, iters);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        double flops = (kind == 0) ? 2.0 * CH * iters * (double)blocks * threads
                                   : 2.0 * 8 * 8 * 4 * CH * iters * (double)blocks * (threads / 32);
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    CK(cudaGetLastError());
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_out);
    *tflops = best;
    return 0;
}

extern "C" int wbgpu_static_scan(wbgpu_ctx* c, int nblocks, const double* dK, const double* weight,
                                 const wbgpu_scan_spec* specs, int nspec, double* out) {
    if (!c || !dK || !weight || !specs || !out) return set_err("wbgpu_static_scan: null pointer argument");
    CK(cudaSetDevice(c->device));
    size_t nout = 0;
    for (int i = 0; i < nspec; i++) {
        int64_t sz = wbgpu_spec_size(&specs[i]);
        if (sz < 0) return set_err("scan: unknown formula %d", specs[i].formula);
        nout += (size_t)sz;
    }
    if (ensure(&c->d_dK, &c->dK_cap, sizeof(double) * 4 * (size_t)std::max(nblocks, 1))) return 1;
    if (ensure(&c->d_out, &c->out_cap, sizeof(double) * nout)) return 1;
    double* d_w = c->d_dK + 3 * (size_t)std::max(nblocks, 1);
    CK(cudaMemcpyAsync(c->d_dK, dK, sizeof(double) * 3 * nblocks, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d_w, weight, sizeof(double) * nblocks, cudaMemcpyHostToDevice, c->stream));
    if (wbgpu_static_scan_dev(c, nblocks, c->d_dK, d_w, specs, nspec, c->d_out)) return 1;
    CK(cudaMemcpyAsync(out, c->d_out, sizeof(double) * nout, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}



// per-K-block results (HOST pointers): out[nblocks][total], total = sum of wbgpu_spec_size over the specs
extern "C" int wbgpu_static_scan_blocks(wbgpu_ctx* c, int nblocks, const double* dK, const wbgpu_scan_spec* specs, int nspec,
                                        double* out) {
    if (!c || !dK || !specs || !out) return set_err("wbgpu_static_scan_blocks: null pointer argument");
    CK(cudaSetDevice(c->device));
    size_t total = 0;
    for (int i = 0; i < nspec; i++) {
        int64_t sz = wbgpu_spec_size(&specs[i]);
        if (sz < 0) return set_err("scan: unknown formula %d", specs[i].formula);
        total += (size_t)sz;
    }
    const size_t nout = total * (size_t)std::max(nblocks, 0);
    if (ensure(&c->d_dK, &c->dK_cap, sizeof(double) * 4 * (size_t)std::max(nblocks, 1))) return 1;
    if (ensure(&c->d_out, &c->out_cap, sizeof(double) * std::max<size_t>(nout, 1))) return 1;
    CK(cudaMemcpyAsync(c->d_dK, dK, sizeof(double) * 3 * nblocks, cudaMemcpyHostToDevice, c->stream));
    if (static_scan_impl(c, nblocks, c->d_dK, nullptr, specs, nspec, c->d_out, true)) return 1;
    CK(cudaMemcpyAsync(out, c->d_out, sizeof(double) * nout, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

// ------------------------------------------------------------------------------------------ tetrahedron method
// per_block: one histogram per K-block (no weights), out[nblocks][total]; dK_cell[nblocks][3] then (cell_stride = 3): the
// K-points of a refinement iteration have cells of different sizes (grid/Kpoint.py:107-109, 145-175)
static int tetra_scan_impl(wbgpu_ctx* c, int nblocks, const double* dK, const double* weight, const double* dK_cell, int cell_stride,
                           const wbgpu_scan_spec* specs, int nspec, double* out, bool per_block) {
    if (!c || !dK || (!weight && !per_block) || !dK_cell || !specs || !out) return set_err("wbgpu_static_scan_tetra: null pointer argument");
    if (!c->planned) return set_err("wbgpu_static_scan_tetra: call wbgpu_plan first");
    if (nblocks < 0 || nspec < 1) return set_err("wbgpu_static_scan_tetra: nblocks=%d nspec=%d", nblocks, nspec);
    CK(cudaSetDevice(c->device));
    const int nw = c->nw;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    size_t nout = 0;
    std::vector<size_t> hoff(nspec + 1, 0);
    for (int i = 0; i < nspec; i++) {
        if (check_spec(c, specs[i])) return 1;
        if (specs[i].use_select) return set_err("scan: select_bands with the tetrahedron method is not implemented");
        if ((specs[i].tetra_flags & 1) && specs[i].fder != 0)
            return set_err("scan: hole_like with the tetrahedron method is not implemented for derivatives of the occupation (fder = %d)", specs[i].fder);
        size_t sz = (size_t)specs[i].nEF * formula_ncomp(specs[i].formula);
        hoff[i + 1] = hoff[i] + 2 * sz;   // direct | suffix
        nout += sz;
    }
    const size_t nhist = per_block ? (size_t)c->nb_max : 1;
    if (ensure(&c->d_hist, &c->hist_cap, sizeof(double) * nhist * hoff[nspec])) return 1;
    CK(cudaMemsetAsync(c->d_hist, 0, sizeof(double) * nhist * hoff[nspec], c->stream));
    if (ensure(&c->d_out, &c->out_cap, sizeof(double) * nout * (per_block ? std::max(nblocks, 1) : 1))) return 1;
    // H-only R-space table for the corner energies
    WbLayout LH;
    LH.nw = nw; LH.ntri = nw * (nw + 1) / 2; LH.E = LH.ntri; LH.off_H = 0; LH.dH_herm = 0;
    for (int a = 0; a < 3; a++) LH.off_dH[a] = LH.off_A[a] = LH.off_O[a] = LH.off_B[a] = LH.off_C[a] = LH.off_S[a] = -1;
    for (int a = 0; a < 6; a++) LH.off_W[a] = -1;
    for (int a = 0; a < 9; a++) LH.off_SA[a] = LH.off_SHA[a] = LH.off_SR[a] = LH.off_SHR[a] = LH.off_dA[a] = LH.off_dO[a] = LH.off_dS[a] = LH.off_dB[a] = LH.off_dC[a] = -1;
    for (int a = 0; a < 3; a++) LH.off_SH[a] = -1;
    for (int a = 0; a < 10; a++) LH.off_W3[a] = -1;
    const size_t ncell = (size_t)c->nbox.x * c->nbox.y * c->nbox.z;
    if (!c->d_tableH) {
        CK(cudaMalloc(&c->d_tableH, sizeof(cplx) * ncell * LH.E));
        CK(cudaMemsetAsync(c->d_tableH, 0, sizeof(cplx) * ncell * LH.E, c->stream));
        WbRInputs in;
        in.Ham = c->d_XR[WBGPU_HAM];
        in.AA = in.BB = in.CC = in.SS = nullptr;
        in.SA = in.SHA = in.SR = in.SH = in.SHR = nullptr;
        in.T = c->d_T;
        in.iRvec = c->d_iRvec;
        long total = (long)c->nR * nw * nw;
        wb_build_rtable_kernel<<<(unsigned)((total + 127) / 128), 128, 0, c->stream>>>(in, LH, c->nR, c->rmin, c->nbox, c->d_tableH);
        c->launches++;
        CK(cudaGetLastError());
    }
    // shifts of the centre and of the 8 corners k +- dK_cell/2 (data_K_R.py:120-141): [9][nblocks][3] | weight
    const int nbk = std::max(nblocks, 1);
    std::vector<double> hdk((size_t)27 * nbk + nbk);
    for (int b = 0; b < nblocks; b++) {
        for (int d = 0; d < 3; d++) hdk[3 * (size_t)b + d] = dK[3 * b + d];
        for (int cr = 0; cr < 8; cr++) {
            const int bit[3] = {(cr >> 2) & 1, (cr >> 1) & 1, cr & 1};
            for (int d = 0; d < 3; d++)
                hdk[3 * ((size_t)(1 + cr) * nbk + b) + d] = dK[3 * b + d] + (bit[d] ? 0.5 : -0.5) * dK_cell[(size_t)cell_stride * b + d];
        }
        hdk[(size_t)27 * nbk + b] = per_block ? 1. : weight[b];
    }
    if (ensure(&c->d_dK, &c->dK_cap, sizeof(double) * hdk.size())) return 1;
    CK(cudaMemcpyAsync(c->d_dK, hdk.data(), sizeof(double) * hdk.size(), cudaMemcpyHostToDevice, c->stream));
    const double* d_w = c->d_dK + (size_t)27 * nbk;
    const size_t nkl = (size_t)c->nb_max * c->nk_block;
    if (ensure(&c->d_Ec, &c->Ec_cap, sizeof(double) * 10 * nkl * nw)) return 1;
    double* d_Ebmin = c->d_Ec + 8 * nkl * nw;
    double* d_Ebmax = d_Ebmin + nkl * nw;
    bool need_U = false;
    for (int i = 0; i < nspec; i++) need_U |= (specs[i].formula != WBGPU_IDENTITY);

    // event groups with the tetrahedron window: [Efermi[0], Efermi[-1]], no widening (tetrahedron.py:232-241)
    std::vector<wbgpu_scan_spec> tsp(specs, specs + nspec);
    std::vector<EvGroup> groups = make_groups(tsp.data(), nspec, true);
    for (EvGroup& G : groups) {
        const wbgpu_scan_spec& s0 = specs[G.specs[0]];
        G.win.EFmin = s0.Ef_first;
        G.win.EFmax = s0.Ef_first + (s0.nEF - 1) * s0.dEF;
        if (s0.nEF > 1) G.win.EFmax = s0.Ef_last;
        G.win.nEFx = s0.nEF;
        G.win.Ebmin = d_Ebmin;
        G.win.Ebmax = d_Ebmax;
    }

    for (int b0 = 0; b0 < nblocks; b0 += c->nb_max) {
        const int nb = std::min(c->nb_max, nblocks - b0);
        const long nk = (long)nb * c->nk_block;
        // ---- corner energies: H-only transform + eigenvalues at the 8 shifted grids
        stage_begin(c, WBGPU_STAGE_EIGH);
        {
            const WbLayout Lsave = c->L;
            cplx* const tsave = c->d_table;
            c->L = LH;
            c->d_table = c->d_tableH;
            int rc = 0;
            for (int cr = 0; cr < 8 && !rc; cr++) {
                rc = run_fourier(c, c->d_dK + 3 * ((size_t)(1 + cr) * nbk + b0), nb);
                if (!rc) rc = run_eigh(c, nk, false);
                if (!rc && cudaMemcpyAsync(c->d_Ec + (size_t)cr * nkl * nw, c->d_E, sizeof(double) * nk * nw, cudaMemcpyDeviceToDevice,
                                           c->stream) != cudaSuccess)
                    rc = set_err("wbgpu_static_scan_tetra: device copy failed");
            }
            c->L = Lsave;
            c->d_table = tsave;
            if (rc) return 1;
        }
        stage_end(c);
        stage_begin(c, WBGPU_STAGE_FOURIER);
        if (run_fourier(c, c->d_dK + 3 * (size_t)b0, nb)) return 1;
        stage_end(c);
        stage_begin(c, WBGPU_STAGE_EIGH);
        if (run_eigh(c, nk, need_U)) return 1;
        wb_tetra_minmax_kernel<<<(unsigned)((nk * nw + 255) / 256), 256, 0, c->stream>>>(c->d_E, c->d_Ec, (long)(nkl * nw), nk * nw,
                                                                                       d_Ebmin, d_Ebmax);
        c->launches++;
        stage_end(c);
        for (const EvGroup& G : groups) {
            stage_begin(c, G.ev.mask == 1 ? WBGPU_STAGE_IDENTITY : WBGPU_STAGE_ROTATE);
            if (run_events(c, G, nk)) return 1;
            stage_end(c);
            stage_begin(c, WBGPU_STAGE_SCAN);
            for (int i : G.specs) {
                const wbgpu_scan_spec& s = specs[i];
                const int ncomp = formula_ncomp(s.formula);
                const int use_smem = 2 * sizeof(double) * (size_t)s.nEF * ncomp <= 96 * 1024;
                const size_t smem = wb_tetra_acc_smem_bytes(nw, s.nEF, ncomp, use_smem);
                if (smem > 48 * 1024)
                    CK(cudaFuncSetAttribute(wb_tetra_accumulate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                if (!per_block) {
                    const long nblk = std::min(nk, (long)sms * 2);
                    wb_tetra_accumulate_kernel<<<(unsigned)nblk, 128, smem, c->stream>>>(
                        c->d_evval + G.ev.off[s.formula], G.ev.NC, nw, nk, c->nk_block, c->d_E, c->d_Ec, (long)(nkl * nw), G.win, d_w + b0,
                        ncomp, (s.tetra_flags & 1) ? -1 : s.fder, s.nEF, s.Ef_first, s.dEF, c->d_hist + hoff[i], use_smem);
                    c->launches++;
                    continue;
                }
                // one launch per K-block on ITS k-points and ITS histogram (transforms, eigensolves and the event pass above
                // stay batched over the K-blocks)
                for (int bk = 0; bk < nb; bk++) {
                    const size_t ko = (size_t)bk * c->nk_block;
                    WbWindow wb = G.win;
                    wb.Ebmin += ko * nw;
                    wb.Ebmax += ko * nw;
                    const long nblk = std::min((long)c->nk_block, (long)sms * 2);
                    wb_tetra_accumulate_kernel<<<(unsigned)nblk, 128, smem, c->stream>>>(
                        c->d_evval + ko * nw * G.ev.NC + G.ev.off[s.formula], G.ev.NC, nw, (long)c->nk_block, c->nk_block, c->d_E + ko * nw,
                        c->d_Ec + ko * nw, (long)(nkl * nw), wb, d_w + b0 + bk, ncomp, (s.tetra_flags & 1) ? -1 : s.fder, s.nEF, s.Ef_first,
                        s.dEF, c->d_hist + (size_t)bk * hoff[nspec] + hoff[i], use_smem);
                    c->launches++;
                }
            }
            stage_end(c);
            CK(cudaGetLastError());
        }
        if (per_block) {   // the histograms of this batch -> out[b0 + bk][...], then cleared for the next batch
            for (int bk = 0; bk < nb; bk++) {
                size_t ooff = 0;
                for (int i = 0; i < nspec; i++) {
                    const wbgpu_scan_spec& s = specs[i];
                    const int ncomp = formula_ncomp(s.formula);
                    const double scale = s.factor / (c->cell_volume * (double)c->nk_block);
                    wb_tetra_finalize_kernel<<<(unsigned)((ncomp + 31) / 32), 32, 0, c->stream>>>(
                        c->d_hist + (size_t)bk * hoff[nspec] + hoff[i], s.nEF, ncomp, scale, c->d_out + (size_t)(b0 + bk) * nout + ooff);
                    c->launches++;
                    ooff += (size_t)s.nEF * ncomp;
                }
            }
            CK(cudaMemsetAsync(c->d_hist, 0, sizeof(double) * nhist * hoff[nspec], c->stream));
        }
    }
    size_t ooff = 0;
    for (int i = 0; i < nspec && !per_block; i++) {
        const wbgpu_scan_spec& s = specs[i];
        const int ncomp = formula_ncomp(s.formula);
        const double scale = s.factor / (c->cell_volume * (double)c->nk_block);
        wb_tetra_finalize_kernel<<<(unsigned)((ncomp + 31) / 32), 32, 0, c->stream>>>(c->d_hist + hoff[i], s.nEF, ncomp, scale, c->d_out + ooff);
        c->launches++;
        ooff += (size_t)s.nEF * ncomp;
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, c->d_out, sizeof(double) * nout * (per_block ? nblocks : 1), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    stage_collect(c);
    return 0;
}

extern "C" int wbgpu_static_scan_tetra(wbgpu_ctx* c, int nblocks, const double* dK, const double* weight, const double* dK_cell,
                                       const wbgpu_scan_spec* specs, int nspec, double* out) {
    return tetra_scan_impl(c, nblocks, dK, weight, dK_cell, 0, specs, nspec, out, false);
}

extern "C" int wbgpu_static_scan_tetra_blocks(wbgpu_ctx* c, int nblocks, const double* dK, const double* dK_cell,
                                              const wbgpu_scan_spec* specs, int nspec, double* out) {
    return tetra_scan_impl(c, nblocks, dK, nullptr, dK_cell, 3, specs, nspec, out, true);
}

// ------------------------------------------------------------------------------------------ Kubo
extern "C" int64_t wbgpu_kubo_size(const wbgpu_kubo_spec* s) {
    if (!s || s->nEF < 1 || s->nomega < 1) return -1;
    if (s->kind == WBGPU_KUBO_OPTCOND) return (int64_t)s->nEF * s->nomega * 18;
    if (s->kind == WBGPU_KUBO_JDOS) return (int64_t)s->nEF * s->nomega;
    if (s->kind == WBGPU_KUBO_SHC || s->kind == WBGPU_KUBO_INJECTION) return (int64_t)s->nEF * s->nomega * 54;
    if (s->kind == WBGPU_KUBO_SHIFT) return (int64_t)s->nEF * s->nomega * 27;
    return -1;
}

// dK_dev / weight_dev / out_dev: DEVICE pointers; Efermi / omega: HOST pointers (scan parameters, like the spec)
// per_block: one accumulator per K-block of a batch, out_dev[nblocks][nout] (weights must be 1): the refinement loop
static int kubo_scan_impl(wbgpu_ctx* c, int nblocks, const double* dK_dev, const double* weight_dev, const wbgpu_kubo_spec* spec,
                          const double* Efermi, const double* omega, double* out_dev, bool per_block = false) {
    if (!c->planned) return set_err("wbgpu_kubo_scan: call wbgpu_plan first");
    const int64_t nout = wbgpu_kubo_size(spec);
    if (nout < 0) return set_err("wbgpu_kubo_scan: bad spec (kind=%d nEF=%d nomega=%d)", spec->kind, spec->nEF, spec->nomega);
    if (spec->smr_type != 0 && spec->smr_type != 1) return set_err("wbgpu_kubo_scan: Invalid smearing type %d", spec->smr_type);
    if (!(spec->smr_fixed_width > 0)) return set_err("wbgpu_kubo_scan: smr_fixed_width must be positive");
    if (!(spec->kBT >= 0)) return set_err("wbgpu_kubo_scan: kBT must not be negative");
    const bool optcond = spec->kind == WBGPU_KUBO_OPTCOND, shc = spec->kind == WBGPU_KUBO_SHC;
    const bool shift = spec->kind == WBGPU_KUBO_SHIFT, inject = spec->kind == WBGPU_KUBO_INJECTION;
    const bool rotated = optcond || shc || shift || inject;   // needs eigenvectors and rotated matrices
    const WbLayout& L = c->L;
    if (shc) {
        const int t = spec->shc_type;
        if (t != WBGPU_SHC_RYOO && t != WBGPU_SHC_QIAO && t != WBGPU_SHC_SIMPLE)
            return set_err("wbgpu_kubo_scan: spin_current_type %d not recognized", t);   // covariant.py:696-697
        if (!((c->mask >> t) & 1u) || L.off_dH[0] < 0 || L.off_S[0] < 0 || (spec->external_terms && L.off_A[0] < 0))
            return set_err("wbgpu_kubo_scan: the plan does not hold the channels of this spin Hall scan (declare WBGPU_SHC_*)");
    }
    if (shift && (!((c->mask >> WBGPU_SHIFT_CURRENT) & 1u) || L.off_dH[0] < 0 || L.off_W[0] < 0 ||
                  (spec->external_terms && (L.off_A[0] < 0 || L.off_dA[0] < 0))))
        return set_err("wbgpu_kubo_scan: the plan does not hold the channels of the shift current (declare WBGPU_SHIFT_CURRENT)");
    if (shift && !(spec->sc_eta > 0)) return set_err("wbgpu_kubo_scan: sc_eta must be positive");
    if ((optcond || inject) && (!((c->mask >> WBGPU_KUBO) & 1u) || L.off_dH[0] < 0 || (spec->external_terms && L.off_A[0] < 0)))
        return set_err("wbgpu_kubo_scan: the plan does not hold the channels of the Kubo path (declare WBGPU_KUBO)");
    for (int i = 1; i < spec->nEF; i++)
        if (!(Efermi[i] > Efermi[i - 1])) return set_err("wbgpu_kubo_scan: Efermi must be strictly ascending");
    CK(cudaSetDevice(c->device));
    const int nw = c->nw, nEF = spec->nEF, nom = spec->nomega, NC = wb_kubo_nc(spec->kind), ENT = wb_kubo_ent(spec->kind);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);

    WbKuboParams P;
    P.kind = spec->kind; P.smr_type = spec->smr_type; P.external = spec->external_terms; P.nEF = nEF; P.nomega = nom;
    P.eta = spec->smr_fixed_width;
    P.sc_eta = spec->sc_eta;
    P.kBT = spec->kBT;
    P.EFmin = Efermi[0]; P.EFmax = Efermi[nEF - 1];
    double wmin = omega[0], wmax = omega[0];
    for (int i = 1; i < nom; i++) { wmin = std::min(wmin, omega[i]); wmax = std::max(wmax, omega[i]); }
    P.wlo = wmin - 5 * spec->smr_fixed_width;
    P.whi = wmax + 5 * spec->smr_fixed_width;
    WbWindow win;
    win.EFmin = -INFINITY; win.EFmax = INFINITY; win.dEF = 1.; win.degen_thresh = spec->degen_thresh;
    win.degen_Kramers = spec->degen_Kramers; win.sea = 0; win.nEFx = nEF;
    win.Ebmin = win.Ebmax = nullptr;
    win.holes = 0; win.Emin_sea = -INFINITY; win.Emax_holes = INFINITY;

    // device copies of the axes: Efermi | omega; accumulator D[nomega][nEF][NC] + output
    if (ensure(&c->d_axes, &c->axes_cap, sizeof(double) * ((size_t)nEF + nom))) return 1;
    const double* d_w = weight_dev;
    double* d_Ef = c->d_axes;
    double* d_om = d_Ef + nEF;
    CK(cudaMemcpyAsync(d_Ef, Efermi, sizeof(double) * nEF, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d_om, omega, sizeof(double) * nom, cudaMemcpyHostToDevice, c->stream));
    const size_t nacc = (size_t)nom * nEF * NC;
    // accumulators of the K-blocks of a batch (per_block: at most 1 GB of them at a time) + one array for the finalised values
    const int nbacc = per_block ? (int)std::max<size_t>(1, std::min<size_t>((size_t)c->nb_max, (size_t)(1.0e9 / (8.0 * nacc)))) : 1;
    if (ensure(&c->d_kacc, &c->kacc_cap, sizeof(double) * ((size_t)nbacc + 1) * nacc)) return 1;
    CK(cudaMemsetAsync(c->d_kacc, 0, sizeof(double) * (size_t)nbacc * nacc, c->stream));
    double* const d_fin = c->d_kacc + (size_t)nbacc * nacc;
    const double scale = spec->factor / (c->cell_volume * (double)c->nk_block);

    const int nwtile = (nom + WB_KUBO_WT - 1) / WB_KUBO_WT;
    const int nthreads = wb_kubo_tpw(spec->kind) * WB_KUBO_WT;

    WbChanList ch;
    ch.n = 0;
    WbShcChans sc;
    sc.type = spec->shc_type; sc.iV = 0; sc.iA = sc.iS = sc.iX1 = sc.iX2 = sc.iX3 = -1;
    auto addn = [&](const int* offs, int n, int herm) {
        const int first = ch.n;
        for (int a = 0; a < n; a++) { ch.off[ch.n] = offs[a]; ch.herm[ch.n] = herm; ch.n++; }
        return first;
    };
    if (rotated) {
        addn(L.off_dH, 3, L.dH_herm);
        if (spec->external_terms) sc.iA = addn(L.off_A, 3, 1);
    }
    if (shc) {
        sc.iS = addn(L.off_S, 3, 1);
        if (sc.type == WBGPU_SHC_RYOO) { sc.iX1 = addn(L.off_SA, 9, 0); sc.iX2 = addn(L.off_SHA, 9, 0); }
        if (sc.type == WBGPU_SHC_QIAO) { sc.iX1 = addn(L.off_SR, 9, 0); sc.iX2 = addn(L.off_SH, 3, 0); sc.iX3 = addn(L.off_SHR, 9, 0); }
    }
    int iW = 0, idA = 0;
    if (shift) {
        iW = addn(L.off_W, 6, L.dH_herm);
        if (spec->external_terms) idA = addn(L.off_dA, 9, 1);
    }
    const bool needJ = shc || shift;   // a [9][nw][nw] matrix per k-point in d_shcJ
    const int cap = std::max(1, nw * (nw - 1));
    const size_t smem_ent = wb_kubo_entries_smem_bytes(nw);
    if (smem_ent > 48 * 1024)
        CK(cudaFuncSetAttribute(wb_kubo_entries_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ent));

    const int nbstep = per_block ? std::min(c->nb_max, nbacc) : c->nb_max;
    for (int b0 = 0; b0 < nblocks; b0 += nbstep) {
        const int nb = std::min(nbstep, nblocks - b0);
        const long nk = (long)nb * c->nk_block;
        stage_begin(c, WBGPU_STAGE_FOURIER);
        if (run_fourier(c, dK_dev + 3 * (size_t)b0, nb)) return 1;
        stage_end(c);
        stage_begin(c, WBGPU_STAGE_EIGH);
        if (run_eigh(c, nk, rotated)) return 1;
        stage_end(c);
        long chunk = std::min(nk, std::max(1L, (long)(3.0e9 / (8.0 * ENT * cap))));
        if (rotated) chunk = std::min(chunk, xbar_chunk(c, ch.n, nk));
        if (rotated && ensure(&c->d_xbar, &c->xbar_cap, sizeof(cplx) * (size_t)chunk * ch.n * nw * nw)) return 1;
        if (needJ && ensure(&c->d_shcJ, &c->shcJ_cap, sizeof(cplx) * (size_t)chunk * 9 * nw * nw)) return 1;
        if (ensure(&c->d_kent, &c->kent_cap, sizeof(double) * (size_t)chunk * cap * ENT + sizeof(int) * (size_t)chunk + 16)) return 1;
        int* d_count = (int*)(c->d_kent + (size_t)chunk * cap * ENT);
        for (long k0 = 0; k0 < nk; k0 += chunk) {
            const long n = std::min(chunk, nk - k0);
            if (rotated) {
                stage_begin(c, WBGPU_STAGE_ROTATE);
                if (rotate_gemm(c, ch, k0, n)) return 1;
                if (shc) {
                    wb_shc_spinvel_kernel<256><<<(unsigned)std::min(n, (long)sms * 8), 256, sizeof(double) * nw, c->stream>>>(
                        (const cplx*)c->d_xbar, ch.n, nw, n, c->d_E + k0 * nw, sc, spec->external_terms, (cplx*)c->d_shcJ);
                    c->launches++;
                }
                if (shift) {
                    // (P_lq^a staged in shared memory when four CTAs per SM still fit)
                    const size_t sm_p = sizeof(double) * ((nw + 1) & ~1) + sizeof(cplx) * 3 * (size_t)nw * nw;
                    const int stage = sm_p <= (size_t)c->smem_optin / 4 - 1024;
                    const size_t sm_agen = stage ? sm_p : sizeof(double) * nw;
                    if (sm_agen > 48 * 1024)
                        CK(cudaFuncSetAttribute(wb_shift_agen_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_agen));
                    wb_shift_agen_kernel<256><<<(unsigned)std::min(n, (long)sms * 8), 256, sm_agen, c->stream>>>(
                        (const cplx*)c->d_xbar, ch.n, nw, n, c->d_E + k0 * nw, iW, sc.iA, idA, spec->sc_eta, (cplx*)c->d_shcJ, stage);
                    c->launches++;
                }
                stage_end(c);
            }
            stage_begin(c, WBGPU_STAGE_SCAN);
            wb_kubo_entries_kernel<128><<<(unsigned)std::min(n, (long)sms * 8), 128, smem_ent, c->stream>>>(
                (const cplx*)c->d_xbar, ch.n, nw, n, k0, c->d_E + k0 * nw, win, P, d_Ef, d_w + b0, c->nk_block, c->d_kent, d_count, cap,
                (const cplx*)c->d_shcJ);
            // accumulation of the k-points [ka, ka + na) of this chunk into `acc`
            auto accumulate = [&](long ka, long na, double* acc) {
                const double* ent = c->d_kent + (size_t)ka * cap * ENT;
                const int* cnt = d_count + ka;
                const int nsplit = (int)std::max(1L, std::min(na, (long)((6 * sms + nwtile - 1) / nwtile)));
                dim3 grid((unsigned)nwtile, (unsigned)nsplit);
                if (optcond && c->kubo_method != 1) {
                    const int ntile = (nom + WB_KUBO_TW - 1) / WB_KUBO_TW;
                    dim3 gridt((unsigned)ntile, (unsigned)std::max(1L, std::min(na, (long)((12 * sms + ntile - 1) / ntile))));
                    wb_kubo_accumulate_optcond_tiled_kernel<<<gridt, 144, 0, c->stream>>>(ent, cnt, cap, na, P, d_om, d_Ef, acc);
                } else if (optcond)
                    wb_kubo_accumulate_kernel<0><<<grid, nthreads, 0, c->stream>>>(ent, cnt, cap, na, P, d_om, d_Ef, acc);
                else if ((shc || shift || inject) && c->kubo_method != 1) {
                    const int ntile = (nom + WB_KUBO_TW - 1) / WB_KUBO_TW;
                    dim3 gridt((unsigned)ntile, (unsigned)std::max(1L, std::min(na, (long)((8 * sms + ntile - 1) / ntile))));
                    if (inject) wb_kubo_accumulate_rank3_tiled_kernel<3><<<gridt, 432, 0, c->stream>>>(ent, cnt, cap, na, P, d_om, d_Ef, acc);
                    else wb_kubo_accumulate_rank3_tiled_kernel<2><<<gridt, 432, 0, c->stream>>>(ent, cnt, cap, na, P, d_om, d_Ef, acc);
                } else if (shc || shift)
                    wb_kubo_accumulate_kernel<2><<<grid, nthreads, 0, c->stream>>>(ent, cnt, cap, na, P, d_om, d_Ef, acc);
                else if (inject)
                    wb_kubo_accumulate_kernel<3><<<grid, nthreads, 0, c->stream>>>(ent, cnt, cap, na, P, d_om, d_Ef, acc);
                else
                    wb_kubo_accumulate_kernel<1><<<grid, nthreads, 0, c->stream>>>(ent, cnt, cap, na, P, d_om, d_Ef, acc);
                c->launches++;
            };
            if (!per_block) accumulate(0, n, c->d_kacc);
            else   // the part of every K-block of the batch that lies in this chunk, into that block's accumulator
                for (long bk = k0 / c->nk_block; bk * c->nk_block < k0 + n; bk++) {
                    const long ka = std::max(k0, bk * (long)c->nk_block), kb = std::min(k0 + n, (bk + 1) * (long)c->nk_block);
                    accumulate(ka - k0, kb - ka, c->d_kacc + (size_t)bk * nacc);
                }
            c->launches++;
            stage_end(c);
            CK(cudaGetLastError());
        }
        if (per_block) {   // this batch's accumulators -> out_dev[b0 + bk][nout], then cleared
            for (int bk = 0; bk < nb; bk++) {
                wb_kubo_finalize_kernel<<<(unsigned)(((size_t)nom * NC + 127) / 128), 128, 0, c->stream>>>(
                    c->d_kacc + (size_t)bk * nacc, nom, nEF, NC, scale, d_fin, shift ? 1 : 0);
                c->launches++;
                CK(cudaMemcpyAsync(out_dev + (size_t)(b0 + bk) * nout, d_fin, sizeof(double) * (size_t)nout, cudaMemcpyDeviceToDevice, c->stream));
            }
            CK(cudaMemsetAsync(c->d_kacc, 0, sizeof(double) * (size_t)nbacc * nacc, c->stream));
        }
    }
    if (!per_block) {
        wb_kubo_finalize_kernel<<<(unsigned)(((size_t)nom * NC + 127) / 128), 128, 0, c->stream>>>(c->d_kacc, nom, nEF, NC, scale, d_fin,
                                                                                                        shift ? 1 : 0);
        c->launches++;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(out_dev, d_fin, sizeof(double) * (size_t)nout, cudaMemcpyDeviceToDevice, c->stream));
    }
    CK(cudaGetLastError());
    stage_collect(c);
    return 0;
}

extern "C" int wbgpu_kubo_scan_dev(wbgpu_ctx* c, int nblocks, const double* dK_dev, const double* weight_dev,
                                   const wbgpu_kubo_spec* spec, const double* Efermi, const double* omega, double* out_dev) {
    if (!c || !dK_dev || !weight_dev || !spec || !Efermi || !omega || !out_dev)
        return set_err("wbgpu_kubo_scan_dev: null pointer argument");
    return kubo_scan_impl(c, nblocks, dK_dev, weight_dev, spec, Efermi, omega, out_dev);
}

extern "C" int wbgpu_kubo_scan(wbgpu_ctx* c, int nblocks, const double* dK, const double* weight, const wbgpu_kubo_spec* spec,
                               const double* Efermi, const double* omega, double* out) {
    if (!c || !dK || !weight || !spec || !Efermi || !omega || !out) return set_err("wbgpu_kubo_scan: null pointer argument");
    const int64_t nout = wbgpu_kubo_size(spec);
    if (nout < 0) return set_err("wbgpu_kubo_scan: bad spec (kind=%d nEF=%d nomega=%d)", spec->kind, spec->nEF, spec->nomega);
    CK(cudaSetDevice(c->device));
    const int nbk = std::max(nblocks, 1);
    if (ensure(&c->d_dK, &c->dK_cap, sizeof(double) * 4 * (size_t)nbk)) return 1;
    if (ensure(&c->d_out, &c->out_cap, sizeof(double) * (size_t)nout)) return 1;
    double* d_w = c->d_dK + 3 * (size_t)nbk;
    CK(cudaMemcpyAsync(c->d_dK, dK, sizeof(double) * 3 * nblocks, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d_w, weight, sizeof(double) * nblocks, cudaMemcpyHostToDevice, c->stream));
    if (kubo_scan_impl(c, nblocks, c->d_dK, d_w, spec, Efermi, omega, c->d_out)) return 1;
    CK(cudaMemcpyAsync(out, c->d_out, sizeof(double) * (size_t)nout, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int wbgpu_kubo_scan_blocks(wbgpu_ctx* c, int nblocks, const double* dK, const wbgpu_kubo_spec* spec, const double* Efermi,
                                      const double* omega, double* out) {
    if (!c || !dK || !spec || !Efermi || !omega || !out) return set_err("wbgpu_kubo_scan_blocks: null pointer argument");
    const int64_t nout = wbgpu_kubo_size(spec);
    if (nout < 0) return set_err("wbgpu_kubo_scan: bad spec (kind=%d nEF=%d nomega=%d)", spec->kind, spec->nEF, spec->nomega);
    CK(cudaSetDevice(c->device));
    const int nbk = std::max(nblocks, 1);
    if (ensure(&c->d_dK, &c->dK_cap, sizeof(double) * 4 * (size_t)nbk)) return 1;
    if (ensure(&c->d_out, &c->out_cap, sizeof(double) * (size_t)nout * nbk)) return 1;
    double* d_w = c->d_dK + 3 * (size_t)nbk;
    std::vector<double> ones((size_t)nbk, 1.);
    CK(cudaMemcpyAsync(c->d_dK, dK, sizeof(double) * 3 * nblocks, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d_w, ones.data(), sizeof(double) * nblocks, cudaMemcpyHostToDevice, c->stream));
    if (kubo_scan_impl(c, nblocks, c->d_dK, d_w, spec, Efermi, omega, c->d_out, true)) return 1;
    CK(cudaMemcpyAsync(out, c->d_out, sizeof(double) * (size_t)nout * nblocks, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

// ------------------------------------------------------------------------------------------ probes
static int probe_prepare(wbgpu_ctx* c, const double dK[3]) {
    if (!c || !dK) return set_err("probe: null pointer argument");
    if (!c->planned) return set_err("probe: call wbgpu_plan first");
    CK(cudaSetDevice(c->device));
    if (ensure(&c->d_dK, &c->dK_cap, sizeof(double) * 4)) return 1;
    CK(cudaMemcpyAsync(c->d_dK, dK, sizeof(double) * 3, cudaMemcpyHostToDevice, c->stream));
    return 0;
}

extern "C" int wbgpu_kpoints(wbgpu_ctx* c, const double dK[3], double* kpoints) {
    if (!kpoints) return set_err("wbgpu_kpoints: null pointer argument");
    if (probe_prepare(c, dK)) return 1;
    long nk = c->nk_block;
    if (ensure(&c->d_out, &c->out_cap, sizeof(double) * 3 * nk)) return 1;
    wb_kpoints_kernel<<<(unsigned)((nk + 127) / 128), 128, 0, c->stream>>>(c->d_dK, make_int3(c->N[0], c->N[1], c->N[2]), c->d_out);
    c->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(kpoints, c->d_out, sizeof(double) * 3 * nk, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int wbgpu_eig(wbgpu_ctx* c, const double dK[3], double* E, double* U) {
    if (!E) return set_err("wbgpu_eig: null pointer argument");
    if (probe_prepare(c, dK)) return 1;
    long nk = c->nk_block;
    if (run_fourier(c, c->d_dK, 1)) return 1;
    if (run_eigh(c, nk, U != nullptr)) return 1;
    CK(cudaMemcpyAsync(E, c->d_E, sizeof(double) * nk * c->nw, cudaMemcpyDeviceToHost, c->stream));
    if (U) CK(cudaMemcpyAsync(U, c->d_U, sizeof(cplx) * nk * c->nw * c->nw, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(&c->last_sweeps, c->d_sweeps, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(&c->last_resolved, c->d_sweeps + 1, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int wbgpu_xk(wbgpu_ctx* c, const double dK[3], int channel, double* X) {
    if (!X) return set_err("wbgpu_xk: null pointer argument");
    if (probe_prepare(c, dK)) return 1;
    const WbLayout& L = c->L;
    const int nw = c->nw;
    const int* offs = nullptr;
    bool herm = false;
    int ncart = 3;
    switch (channel) {
        case WBGPU_CH_HAM: offs = &L.off_H; herm = true; ncart = 1; break;
        case WBGPU_CH_DHAM: offs = L.off_dH; herm = L.dH_herm; break;
        case WBGPU_CH_AA: offs = L.off_A; herm = true; break;
        case WBGPU_CH_ROTAA: offs = L.off_O; herm = true; break;
        case WBGPU_CH_BB: offs = L.off_B; break;
        case WBGPU_CH_CC: offs = L.off_C; break;
        case WBGPU_CH_SS: offs = L.off_S; herm = true; break;
        default: return set_err("wbgpu_xk: unknown channel %d", channel);
    }
    if (offs[0] < 0) return set_err("wbgpu_xk: channel %d is not part of the current plan", channel);
    long nk = c->nk_block;
    if (run_fourier(c, c->d_dK, 1)) return 1;
    std::vector<cplx> rec((size_t)nk * L.E);
    CK(cudaMemcpyAsync(rec.data(), c->d_X, sizeof(cplx) * rec.size(), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    cplx* out = (cplx*)X;  // unpacking of the parity probe's output (layout only, no arithmetic)
    for (long ik = 0; ik < nk; ik++)
        for (int i = 0; i < nw; i++)
            for (int j = 0; j < nw; j++)
                for (int a = 0; a < ncart; a++) {
                    const cplx* r = rec.data() + ik * L.E + offs[a];
                    cplx v;
                    if (!herm) v = r[i * nw + j];
                    else if (i <= j) v = r[tri_index(i, j, nw)];
                    else { v = r[tri_index(j, i, nw)]; v.y = -v.y; }
                    out[((ik * nw + i) * nw + j) * ncart + a] = v;
                }
    return 0;
}

// index of the sorted triple (b <= c <= d) in off_W3: xxx xxy xxz xyy xyz xzz yyy yyz yzz zzz
static int sym10(int b, int c, int d) {
    int t[3] = {b, c, d};
    std::sort(t, t + 3);
    static const int first[3] = {0, 6, 9};
    const int within = (t[1] - t[0]) * (2 * (3 - t[0]) - (t[1] - t[0]) + 1) / 2 + (t[2] - t[1]);   // pairs (c, d) >= t0, row-major
    return first[t[0]] + within;
}

// Parity / plug-in probe: the Hamiltonian-gauge matrix  Xbar(name, der) = U^dagger (d^der X) U  of one K-block
// (Data_K_R.Xbar, data_K/data_K_R.py:69-97; the rotation of data_K.py:130-132 runs in wb_rotate_gemm_kernel).
// X[nk][nw][nw][3]^ncart, ncart = (name != Ham) + der: value components first, derivative components after them.
extern "C" int wbgpu_xbar(wbgpu_ctx* c, const double dK[3], int channel, int der, double* X) {
    if (!X) return set_err("wbgpu_xbar: null pointer argument");
    if (probe_prepare(c, dK)) return 1;
    const WbLayout& L = c->L;
    const int nw = c->nw, n2 = nw * nw;
    std::vector<int> comp_off;   // record offset per output component
    int herm = 0;
    auto need = [&](const int* offs, int n, int h, const char* what) {
        if (offs[0] < 0) return set_err("wbgpu_xbar: %s is not part of the current plan", what);
        for (int a = 0; a < n; a++) comp_off.push_back(offs[a]);
        herm = h;
        return 0;
    };
    int rc = 0;
    if (channel == WBGPU_CH_HAM) {
        if (der == 0) rc = need(&L.off_H, 1, 1, "Ham");
        else if (der == 1) rc = need(L.off_dH, 3, L.dH_herm, "d Ham");
        else if (der == 2) {
            if (L.off_W[0] < 0) return set_err("wbgpu_xbar: d^2 Ham is not part of the current plan");
            for (int b = 0; b < 3; b++) for (int d = 0; d < 3; d++) comp_off.push_back(L.off_W[wb_sym6(b, d)]);
            herm = L.dH_herm;
        } else if (der == 3) {
            if (L.off_W3[0] < 0) return set_err("wbgpu_xbar: d^3 Ham is not part of the current plan");
            for (int b = 0; b < 3; b++) for (int cc = 0; cc < 3; cc++) for (int d = 0; d < 3; d++)
                comp_off.push_back(L.off_W3[sym10(b, cc, d)]);
            herm = L.dH_herm;
        } else return set_err("wbgpu_xbar: Ham has comma-derivatives up to order 3");
    } else {
        if (der == 2) {   // [b][d][e] from the 6 symmetric (d, e) pairs per value component
            const int* offs = nullptr;
            int h2 = 1;
            switch (channel) {
                case WBGPU_CH_AA: offs = L.off_d2A; break;
                case WBGPU_CH_ROTAA: offs = L.off_d2O; break;
                case WBGPU_CH_SS: offs = L.off_d2S; break;
                case WBGPU_CH_BB: offs = L.off_d2B; h2 = 0; break;
                case WBGPU_CH_CC: offs = L.off_d2C; h2 = 0; break;
                default: return set_err("wbgpu_xbar: unknown channel %d", channel);
            }
            if (offs[0] < 0) return set_err("wbgpu_xbar: second comma-derivatives of channel %d are not part of the current plan (WBGPU_XBAR_DER2)", channel);
            for (int b = 0; b < 3; b++) for (int d = 0; d < 3; d++) for (int e = 0; e < 3; e++) comp_off.push_back(offs[6 * b + wb_sym6(d, e)]);
            herm = h2;
        } else {
        if (der < 0 || der > 2) return set_err("wbgpu_xbar: comma-derivatives of order 0, 1 and 2 only");
        switch (channel) {
            case WBGPU_CH_AA: rc = der ? need(L.off_dA, 9, 1, "d AA") : need(L.off_A, 3, 1, "AA"); break;
            case WBGPU_CH_ROTAA: rc = der ? need(L.off_dO, 9, 1, "d rotAA") : need(L.off_O, 3, 1, "rotAA"); break;
            case WBGPU_CH_BB: rc = der ? need(L.off_dB, 9, 0, "d BB") : need(L.off_B, 3, 0, "BB"); break;
            case WBGPU_CH_CC: rc = der ? need(L.off_dC, 9, 0, "d CC") : need(L.off_C, 3, 0, "CC"); break;
            case WBGPU_CH_SS: rc = der ? need(L.off_dS, 9, 1, "d SS") : need(L.off_S, 3, 1, "SS"); break;
            default: return set_err("wbgpu_xbar: unknown channel %d", channel);
        }
        }
    }
    if (rc) return rc;
    // rotate every distinct record channel once
    WbChanList ch;
    ch.n = 0;
    std::vector<int> which(comp_off.size());
    for (size_t a = 0; a < comp_off.size(); a++) {
        int found = -1;
        for (int i = 0; i < ch.n; i++) if (ch.off[i] == comp_off[a]) found = i;
        if (found < 0) { found = ch.n; ch.off[ch.n] = comp_off[a]; ch.herm[ch.n] = herm; ch.n++; }
        which[a] = found;
    }
    const long nk = c->nk_block;
    const int ncomp = (int)comp_off.size();
    if (run_fourier(c, c->d_dK, 1)) return 1;
    if (run_eigh(c, nk, true)) return 1;
    const long chunk = xbar_chunk(c, ch.n, nk);
    if (ensure(&c->d_xbar, &c->xbar_cap, sizeof(cplx) * (size_t)chunk * ch.n * n2)) return 1;
    std::vector<cplx> host((size_t)chunk * ch.n * n2);
    cplx* out = (cplx*)X;
    for (long k0 = 0; k0 < nk; k0 += chunk) {
        const long n = std::min(chunk, nk - k0);
        if (rotate_gemm(c, ch, k0, n)) return 1;
        CK(cudaMemcpyAsync(host.data(), c->d_xbar, sizeof(cplx) * (size_t)n * ch.n * n2, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        for (long k = 0; k < n; k++)          // re-ordering of the probe's output (layout only, no arithmetic)
            for (int a = 0; a < ncomp; a++) {
                const cplx* src = host.data() + ((size_t)k * ch.n + which[a]) * n2;
                cplx* dst = out + (size_t)(k0 + k) * n2 * ncomp + a;
                for (int x = 0; x < n2; x++) dst[(size_t)x * ncomp] = src[x];
            }
    }
    return 0;
}

extern "C" int wbgpu_band_traces(wbgpu_ctx* c, const double dK[3], const wbgpu_scan_spec* spec, double* E_label,
                                 double* value) {
    if (!spec || !E_label || !value) return set_err("wbgpu_band_traces: null pointer argument");
    if (probe_prepare(c, dK)) return 1;
    if (check_spec(c, *spec)) return 1;
    long nk = c->nk_block;
    int ncomp = formula_ncomp(spec->formula);
    if (run_fourier(c, c->d_dK, 1)) return 1;
    if (run_eigh(c, nk, true)) return 1;
    CK(cudaMemsetAsync(c->d_evval, 0, sizeof(double) * nk * c->nw * ncomp, c->stream));
    std::vector<EvGroup> groups = make_groups(spec, 1);
    if (run_events(c, groups[0], nk)) return 1;
    CK(cudaMemcpyAsync(E_label, c->d_evlabel, sizeof(double) * nk * c->nw, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(value, c->d_evval, sizeof(double) * nk * c->nw * ncomp, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}
