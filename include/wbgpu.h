/*
 * wbgpu.h -- C-ABI of the B200-native k-grid evaluation library (libwbgpu.so).
 *
 * This is the drop-in boundary for the hot path of wannier-berri
 *   wannierberri.fourier + data_K + formula + calculators.static   (SURVEY.md section 8).
 * The reference is pure Python; the binding a maintainer adds is a ctypes.CDLL stub
 * (see INTEGRATION.md).  Plain pointers and sizes only; no torch / numpy types.
 *
 * Conventions
 *   - complex128 arrays are passed as `const double*` with (re, im) interleaved, C order;
 *   - every function returns 0 on success, non-zero on error; `wbgpu_last_error()` returns a
 *     human readable description of the last error on the calling thread;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails;
 *   - a context is used from one host thread at a time (the reference is single threaded per
 *     process, run_grid.py:258-265); different contexts may be used concurrently.
 *
 * Reference interfaces replaced (paths relative to /root/reference/wannierberri):
 *   wbgpu_create / wbgpu_set_R_matrix  <- System_R.get_R_mat, .rvec.iRvec, .rvec.cRvec_shifted
 *                                         (system/system_R.py:106-115, fourier/rvectors.py:351-369)
 *   wbgpu_plan                         <- Data_K_R.__init__ / Rvectors.set_fft_R_to_k
 *                                         (data_K/data_K_R.py:11-22, fourier/rvectors.py:443-475)
 *   wbgpu_static_scan(_dev)            <- paralfunc + StaticCalculator.__call__ over a list of K-blocks
 *                                         (run_grid.py:258-265,59-72; calculators/static.py:60-169)
 *   wbgpu_static_scan_tetra            <- the same with tetra=True (Data_K.tetraWeights, grid/tetrahedron.py)
 *   wbgpu_static_scan_blocks, wbgpu_static_scan_tetra_blocks, wbgpu_kubo_scan_blocks
 *                                      <- the same per K-block: Kpoint.set_result of the refinement loop
 *                                         (run_grid.py:59-72,343-375; grid/Kpoint.py:35-38,85-92,145-175)
 *   wbgpu_eig                          <- Data_K.E_K (data_K/data_K.py:211-218)
 *   wbgpu_xbar                         <- Data_K_R.Xbar / Data_K._rotate (data_K/data_K_R.py:69-97, data_K/data_K.py:130-132)
 *   wbgpu_xk                           <- Rvectors.R_to_k / FFT_R_to_k.__call__
 *                                         (fourier/rvectors.py:496-506, fourier/fft.py:133-192)
 *   wbgpu_band_traces                  <- Formula_ln.trace per band group (formula/formula.py:76-79)
 *   wbgpu_kubo_scan                    <- paralfunc + DynamicCalculator.__call__ over a list of K-blocks
 *                                         (run_grid.py:258-265; calculators/dynamic.py:26-114,146-196)
 */
#ifndef WBGPU_H
#define WBGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wbgpu_ctx wbgpu_ctx;

/* formulae (formula/covariant.py) */
enum {
    WBGPU_IDENTITY = 0,   /* Identity        covariant.py:10-22   rank 0 */
    WBGPU_OMEGA = 1,      /* Omega           covariant.py:161-203 rank 1 */
    WBGPU_MORB_HPM = 2,   /* Morb_Hpm(+1)    covariant.py:424-449 rank 1, non-additive */
    WBGPU_VEL_OMEGA = 3,  /* VelOmega        covariant.py:798-802 rank 2 */
    WBGPU_VEL_HPLUS = 4,  /* VelHplus        covariant.py:805-809 rank 2 */
    WBGPU_VEL_SPIN = 5,   /* VelSpin         covariant.py:812-814 rank 2 */
    WBGPU_SPIN = 6,       /* Spin            covariant.py:331-335 rank 1 */
    WBGPU_KUBO = 7,       /* plan flag only: channels of the Kubo path (dH, and A with external terms),
                             Formula_OptCond calculators/dynamic.py:170-181 */
    WBGPU_VEL_VEL = 8,    /* VelVel          covariant.py:817-820 rank 2 */
    WBGPU_INV_MASS = 9,   /* InvMass         elementary.py:28-34 (generalised derivative of the velocity, needs
                             the second comma-derivative of H) rank 2 */
    /* Spin Hall conductivity with the spin current of Ryoo et al. (R-matrices SA, SHA), of Qiao et al. (SR, SH, SHR) or
       {S, v}/2 (SS only).  As a plan flag: the channels of the Kubo scan dynamic.SHC (WBGPU_KUBO_SHC).  As the formula of
       a static scan: SpinOmega (spin Berry curvature, covariant.py:759-789; static.SHC), rank 3 [a][b][s]. */
    WBGPU_SHC_RYOO = 10,
    WBGPU_SHC_QIAO = 11,
    WBGPU_SHC_SIMPLE = 12,
    WBGPU_DER_OMEGA = 13, /* DerOmega        covariant.py:212-259 (generalised derivative of the Berry curvature:
                             BerryDipole_FermiSea, NLAHC_FermiSea), rank 2 [c][d]; needs d_b d_d H, d_d A_b, d_d rotA_c */
    WBGPU_DER_SPIN = 14,  /* DerSpin         covariant.py:338-342 (generalised derivative of the spin: GME_spin_FermiSea) rank 2 [s][d] */
    /* FormulaProduct (formula/formula.py:121-149) of Velocity, InvMass, Omega, Spin nn-blocks (covariant.py:823-858) */
    WBGPU_VEL_VEL_VEL = 15,   /* VelVelVel  rank 3 (NLDrude_Fermider2)      */
    WBGPU_MASS_VEL = 16,      /* MassVel    rank 3 (NLDrude_FermiSurf)      */
    WBGPU_MASS_MASS = 17,     /* MassMass   rank 4 (Hall_classic_FermiSea)  */
    WBGPU_VEL_MASS_VEL = 18,  /* VelMassVel rank 4 (Hall_classic_FermiSurf) */
    WBGPU_OMEGA_S = 19,       /* OmegaS     rank 2 (AHC_Zeeman_spin)        */
    WBGPU_OMEGA_OMEGA = 20,   /* OmegaOmega rank 2                          */
    WBGPU_SHIFT_CURRENT = 21, /* plan flag only: channels of the Kubo shift current (d_a H, d_b d_d H, A, d_d A_b) */
    WBGPU_DER3E = 22,         /* Der3E  covariant.py:126-151 (third derivative of the band energy: NLDrude_FermiSea) rank 3;
                                 needs d_b d_d H and d_b d_c d_d H */
    WBGPU_DER_MORB = 23,      /* DerMorb (sign = +1)  covariant.py:463-547 (generalised derivative of the orbital moment:
                                 GME_orb_FermiSea), rank 2 [c][d], not additive; needs the channels of DerOmega plus BB, CC and
                                 their comma-derivatives */
    WBGPU_OMEGA_HPLUS = 24,   /* OmegaHplus  covariant.py:861-865 (Omega x Morb_Hpm blocks: AHC_Zeeman_orb) rank 2 */
    WBGPU_XBAR_DER2 = 25,     /* no scan: makes wbgpu_plan keep the SECOND comma-derivatives of AA, rotAA, BB, CC, SS (of those that are
                                 set) for wbgpu_xbar(..., der = 2), i.e. Data_K_R.Xbar(name, 2) of plug-in formulae */
    WBGPU_NFORMULA = 26
};

/* R-space matrices a context can hold (System_R._XX_R keys) */
enum { WBGPU_HAM = 0, WBGPU_AA = 1, WBGPU_BB = 2, WBGPU_CC = 3, WBGPU_SS = 4,
       /* spin-current matrices of the Kubo spin Hall conductivity (formula/covariant.py:729-756) */
       WBGPU_SA = 5, WBGPU_SHA = 6, WBGPU_SR = 7, WBGPU_SH = 8, WBGPU_SHR = 9, WBGPU_NKEYS = 10 };

/* Wannier-gauge k-space channels that wbgpu_xk can return (parity probe) */
enum {
    WBGPU_CH_HAM = 0,     /* H(k)            hermitised   [nk][nw][nw]    */
    WBGPU_CH_DHAM = 1,    /* d_a H(k)        as is        [nk][nw][nw][3] */
    WBGPU_CH_AA = 2,      /* A_a(k)          hermitised   [nk][nw][nw][3] */
    WBGPU_CH_ROTAA = 3,   /* (curl A)_c(k)   hermitised   [nk][nw][nw][3] */
    WBGPU_CH_BB = 4,      /* B_a(k)          as is        [nk][nw][nw][3] */
    WBGPU_CH_CC = 5,      /* C_c(k)          as is        [nk][nw][nw][3] */
    WBGPU_CH_SS = 6       /* S_a(k)          hermitised   [nk][nw][nw][3] */
};

/* One Fermi-level scan = one StaticCalculator.__call__ (calculators/static.py:26-169),
 * tetra=False, k_resolved=False. */
typedef struct wbgpu_scan_spec {
    int32_t formula;         /* WBGPU_*                                                     */
    int32_t fder;            /* 0 Fermi sea, 1..3 derivatives of f (static.py:137-147)        */
    int32_t nEF;             /* len(Efermi)                                                  */
    int32_t degen_Kramers;   /* calculator.py:20                                             */
    int32_t internal_terms;  /* formula.py:11-29                                             */
    int32_t external_terms;
    double Ef_first;         /* Efermi[0]                                                    */
    double Ef_last;          /* Efermi[-1]                                                   */
    double dEF;              /* Efermi[1]-Efermi[0]  (0.001 if nEF == 1), static.py:55       */
    double degen_thresh;     /* calculator.py:20                                             */
    double factor;           /* constant_factor (or its sign), hole_like already applied     */
    /* select_bands (static.py:93-100, 129-136; utility.py:398-403): bit b of select_mask[b / 64] set = band b selected.
     * A band group counts with the fraction of its bands that is selected (groups without a selected band drop out).
     * use_select = 0: all bands.  Fermi-surface scans only (fder >= 1), as in the reference (data_K.py:179-180). */
    uint64_t select_mask[2];
    int32_t use_select;
    /* tetrahedron method only (wbgpu_static_scan_tetra; grid/tetrahedron.py:197-198, 246-266; static.py:84-91):
     * bit 0 = hole_like: weights 1 - occupation, plus the group of the bands above Efermi[-1] (up to tetra_Emax if bit 2);
     * bit 1 = tetra_Emin is set: the Fermi-sea group starts at the first band that reaches it;  bit 2 = tetra_Emax is set.
     * The plain scans ignore them, as the reference does (data_K.py:172-185). */
    int32_t tetra_flags;
    double tetra_Emin;
    double tetra_Emax;
} wbgpu_scan_spec;

const char* wbgpu_last_error(void);
int wbgpu_version(void);
/* number of CUDA devices visible (0 when there is none); never fails */
int wbgpu_device_count(void);

/* Create a context on CUDA device `device` for a system with `nw` Wannier functions and `nR`
 * lattice vectors.  iRvec[nR][3] (int32), cRvec_shifted[nR][nw][nw][3] = R + t_j - t_i (Angstrom).
 * All inputs are HOST pointers and are copied.  `stream` is a cudaStream_t (NULL = default). */
int wbgpu_create(wbgpu_ctx** ctx, int device, int nw, int nR, const int32_t* iRvec,
                 const double* cRvec_shifted, double cell_volume, void* stream);
int wbgpu_destroy(wbgpu_ctx* ctx);

/* Copy one R-space matrix to the device: X_R[nR][nw][nw][ncart] complex128, ncart = 1 (Ham), 3 (AA, BB, CC, SS, SH)
 * or 9 (SA, SHA, SR, SHR: [3 velocity][3 spin]). */
int wbgpu_set_R_matrix(wbgpu_ctx* ctx, int key, const double* X_R, int ncart);

/* Fix the FFT sub-grid NKFFT[3] and the set of formulae that will be evaluated
 * (formula_mask = OR of 1<<WBGPU_*), `external_terms` = whether any of them needs AA/BB/CC
 * channels.  Builds the device-resident Fourier tables and sizes the workspace for at most
 * `max_kpoints_per_launch` k-points per kernel batch (0 = choose from free memory). */
int wbgpu_plan(wbgpu_ctx* ctx, const int32_t NKFFT[3], uint32_t formula_mask, int external_terms,
               int64_t max_kpoints_per_launch);

/* Evaluate `nspec` Fermi scans over `nblocks` K-blocks:  out = sum_b weight[b] * scan(block b).
 * dK[nblocks][3] = Kpoint.Kp_fullBZ, weight[nblocks] = Kpoint.factor (HOST pointers).
 * out: concatenation over specs of data[nEF][3^rank] float64 (HOST pointer), already divided by
 * cell_volume and nk and multiplied by spec.factor (static.py:149-155).
 * Host<->device copies happen inside the call. */
int wbgpu_static_scan(wbgpu_ctx* ctx, int nblocks, const double* dK, const double* weight,
                      const wbgpu_scan_spec* specs, int nspec, double* out);
/* Same with dK / weight / out resident in device memory (DEVICE pointers); asynchronous on the
 * context's stream. */
int wbgpu_static_scan_dev(wbgpu_ctx* ctx, int nblocks, const double* dK_dev, const double* weight_dev,
                          const wbgpu_scan_spec* specs, int nspec, double* out_dev);
/* Per-K-block results, what adaptive refinement needs (Kpoint.set_result / .max, run_grid.py:59-72,343-375;
 * grid/Kpoint.py:35-38,85-92): out[nblocks][total], total = sum of wbgpu_spec_size over the specs; block b holds
 * the scans of K-block b alone (no weight).  HOST pointers. */
int wbgpu_static_scan_blocks(wbgpu_ctx* ctx, int nblocks, const double* dK, const wbgpu_scan_spec* specs, int nspec,
                             double* out);
/* number of float64 values one spec writes */
int64_t wbgpu_spec_size(const wbgpu_scan_spec* spec);
/* The same scans with the tetrahedron method on grid K-blocks, StaticCalculator(tetra=True) with KpointBZparallel
 * (calculators/static.py:84-91,121-127; grid/tetrahedron.py:15-128,165-268; data_K/data_K_R.py:120-141):
 * dK_cell[3] = Kpoint.dK_fullBZ, the edge of the cell around every k-point (1 / (NKdiv * NKFFT)); the Fermi levels
 * are Ef_first + i * dEF.  The band energies at the 8 cell corners are evaluated inside the call.  HOST pointers. */
int wbgpu_static_scan_tetra(wbgpu_ctx* ctx, int nblocks, const double* dK, const double* weight, const double* dK_cell,
                            const wbgpu_scan_spec* specs, int nspec, double* out);
/* Per-K-block results of the tetrahedron scans for the refinement loop (run_grid.py:343-375): out[nblocks][total] without
 * weights; dK_cell[nblocks][3] -- refined K-points have smaller cells (grid/Kpoint.py:107-109,145-175).  HOST pointers. */
int wbgpu_static_scan_tetra_blocks(wbgpu_ctx* ctx, int nblocks, const double* dK, const double* dK_cell,
                                   const wbgpu_scan_spec* specs, int nspec, double* out);

/* One (Efermi x omega) scan = one DynamicCalculator.__call__ (calculators/dynamic.py:26-114). */
enum { WBGPU_KUBO_OPTCOND = 0,  /* OpticalConductivity dynamic.py:184-196: complex128 data[nEF][nomega][3][3]    */
       WBGPU_KUBO_JDOS = 1,     /* JDOS                dynamic.py:146-162: float64    data[nEF][nomega]          */
       WBGPU_KUBO_SHC = 2,      /* SHC                 dynamic.py:204-237: complex128 data[nEF][nomega][3][3][3] */
       WBGPU_KUBO_SHIFT = 3,    /* ShiftCurrent        dynamic.py:244-322: float64    data[nEF][nomega][3][3][3] (spec.sc_eta) */
       WBGPU_KUBO_INJECTION = 4 /* InjectionCurrent    dynamic.py:330-365: complex128 data[nEF][nomega][3][3][3] */ };
typedef struct wbgpu_kubo_spec {
    int32_t kind;            /* WBGPU_KUBO_*                                   */
    int32_t nEF, nomega;
    int32_t smr_type;        /* 0 = Lorentzian, 1 = Gaussian (dynamic.py:49-54) */
    int32_t degen_Kramers;
    int32_t external_terms;  /* kwargs_formula (formula.py:11-29)               */
    int32_t shc_type;        /* SHC only: WBGPU_SHC_RYOO / _QIAO / _SIMPLE      */
    int32_t reserved;
    double smr_fixed_width;
    double degen_thresh;
    double factor;           /* constant_factor                                */
    double sc_eta;           /* ShiftCurrent only: broadening of the energy denominators of the generalised derivative */
    double kBT;              /* temperature of the Fermi-Dirac factor (utility.py:172-182); 0 = step function        */
} wbgpu_kubo_spec;
/* number of float64 values the scan writes (complex counted as 2) */
int64_t wbgpu_kubo_size(const wbgpu_kubo_spec* spec);
/* out = sum_b weight[b] * scan(block b).  Efermi[nEF] (ascending) and omega[nomega]: HOST pointers; out: HOST
 * pointer, layout of the reference's EnergyResult.data: [nEF][nomega][3][3] complex128 (re, im interleaved) or
 * [nEF][nomega] float64, already multiplied by factor / (nk * cell_volume) (dynamic.py:101).
 * The plan must include WBGPU_KUBO. */
int wbgpu_kubo_scan(wbgpu_ctx* ctx, int nblocks, const double* dK, const double* weight, const wbgpu_kubo_spec* spec,
                    const double* Efermi, const double* omega, double* out);

/* Same with dK / weight / out resident in device memory (DEVICE pointers; Efermi / omega stay HOST pointers: they are
 * parameters of the scan like the spec); asynchronous on the context's stream. */
int wbgpu_kubo_scan_dev(wbgpu_ctx* ctx, int nblocks, const double* dK_dev, const double* weight_dev,
                        const wbgpu_kubo_spec* spec, const double* Efermi, const double* omega, double* out_dev);
/* Per-K-block results of a Kubo scan for the refinement loop (run_grid.py:343-375): out[nblocks][wbgpu_kubo_size] without
 * weights.  HOST pointers. */
int wbgpu_kubo_scan_blocks(wbgpu_ctx* ctx, int nblocks, const double* dK, const wbgpu_kubo_spec* spec, const double* Efermi,
                           const double* omega, double* out);

/* Parity probes (HOST output pointers). One K-block each. */
int wbgpu_kpoints(wbgpu_ctx* ctx, const double dK[3], double* kpoints /*[nk][3]*/);
int wbgpu_eig(wbgpu_ctx* ctx, const double dK[3], double* E /*[nk][nw]*/, double* U /*[nk][nw][nw] c128 or NULL*/);
int wbgpu_xk(wbgpu_ctx* ctx, const double dK[3], int channel, double* X /*complex128, see enum*/);
/* Hamiltonian-gauge matrices of one K-block for plug-in formulae:  Xbar(name, der) = U^dagger (d^der X) U
 * (Data_K_R.Xbar, data_K/data_K_R.py:69-97).  channel: WBGPU_CH_HAM (der 0..3), WBGPU_CH_AA / ROTAA / BB / CC / SS
 * (der 0..2; der 2 needs WBGPU_XBAR_DER2 in the plan); the plan must hold the channel.  X[nk][nw][nw][3]^(ncart + der) complex128, value components first. */
int wbgpu_xbar(wbgpu_ctx* ctx, const double dK[3], int channel, int der, double* X);
/* per-k, per-band-group traces of a formula: E_label[nk][nw], value[nk][nw][3^rank]; slot b is
 * used iff a kept group starts at band b (label -inf for the Fermi-sea group), else label=+inf */
int wbgpu_band_traces(wbgpu_ctx* ctx, const double dK[3], const wbgpu_scan_spec* spec,
                      double* E_label, double* value);

/* stages of one batch of K-blocks, for wbgpu_stage_times */
enum {
    WBGPU_STAGE_FOURIER = 0,  /* twiddles + 3 axis passes of the R->k transform */
    WBGPU_STAGE_EIGH = 1,     /* batched Hermitian eigensolver */
    WBGPU_STAGE_ROTATE = 2,   /* U^dagger X U + formula -> band-group events */
    WBGPU_STAGE_IDENTITY = 3, /* band groups of the Identity formula */
    WBGPU_STAGE_SCAN = 4,     /* histogram accumulation (static) / entries + accumulation (Kubo) */
    WBGPU_NSTAGES = 5
};
/* With option "timing" = 1 every stage of every batch is bracketed by CUDA events on the context's
 * stream; accumulated device milliseconds and the number of timed stage instances since the option
 * was set.  ms[WBGPU_NSTAGES], calls[WBGPU_NSTAGES]. */
int wbgpu_stage_times(const wbgpu_ctx* ctx, double* ms, int64_t* calls);
/* FP64 peak probes for the roofline denominator: kind 0 = DFMA, 1 = DMMA (mma.sync m8n8k4). TFLOP/s. */
int wbgpu_fp64_peak(int device, int kind, double* tflops);

/* counters since context creation */
int64_t wbgpu_kernel_launches(const wbgpu_ctx* ctx);
/* last eigensolver launch: max Jacobi sweeps over k-points (diagnostic) */
int wbgpu_last_eig_sweeps(const wbgpu_ctx* ctx);
/* last wbgpu_eig call: k-points that the fast eigensolver handed to the Jacobi kernel (diagnostic) */
int wbgpu_last_eig_resolved(const wbgpu_ctx* ctx);
/* options; "eig_method": 0 = automatic (Householder + QL eigenvalues + twisted-factorisation eigenvectors for
 * num_wann <= 24, Householder + QL with accumulated rotations above), 1 = Jacobi, 2 / 3 = Householder + QL with
 * accumulated rotations (3: one k-point per warp in the reduction), 4 = as 0 with the thread-per-matrix reduction */
int wbgpu_set_option(wbgpu_ctx* ctx, const char* name, int64_t value);

#ifdef __cplusplus
}
#endif
#endif /* WBGPU_H */
