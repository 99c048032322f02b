/*
 * kreg_b200.h -- C ABI of libkreg_b200.so: the B200-native (sm_100a) KREG hot path.
 *
 * This is the drop-in boundary for the reference's native KREG layer.  Every entry point
 * replaces one routine of the reference's f2py module `KREG` (mlatom/fortran/KREG.f90, called
 * from mlatom/kreg_api.py) or of MLatomF's KRR engine (mlatom/MLatomF_src/A_KRR*.f90); the
 * replaced interface is cited at each declaration.  INTEGRATION.md shows the binding a
 * maintainer would put in mlatom/kreg_api.py.
 *
 * Conventions
 *  - All array arguments are DEVICE pointers to fp64 unless the name ends in `_host`.
 *    The caller owns every buffer (PyTorch tensors in the Python host layer).  No entry point
 *    allocates device memory; scratch space is passed in and its size is queried first.
 *  - Layouts are C order and byte-identical to the reference's Fortran arrays:
 *      xyz[p][a][t]   == Fortran XYZ(t,a,p)  (3,Nat,N)      (KREG.f90:299)
 *      X[p][d]        == Fortran X(d,p)      (D,N)          (KREG.f90:300)
 *      K[i*ldk + j]   == Fortran K(j,i) == K(i,j)           (KREG.f90:301; symmetric)
 *      alpha / y      : values first, then gradients in (point, atom, xyz) order, xyz fastest
 *                       (KREG.f90:273-291 get_itrgrxyz; kreg_api.py:51-72)
 *    D = Nat(Nat-1)/2; descriptor index of atom pair a<b (0-based): (2Nat-a-1)a/2 + (b-a) - 1
 *    (KREG.f90:48).
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls are
 *    asynchronous with respect to the host unless stated otherwise, re-entrant per stream.
 *  - Return value: 0 = ok; <0 = argument / capability error (see KREG_E_*); >0 = numerical
 *    failure reported by the factorisation (LAPACK `info`, kreg_solve only).  The library never
 *    calls exit()/abort() (the reference's stopKREG does: mlatom/fortran/stopper.f90:13).
 *  - There is no CPU fallback anywhere: every function needs an sm_100 device.
 */
#ifndef KREG_B200_H
#define KREG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KREG_OK 0
#define KREG_E_BADARG (-1)      /* null pointer, negative size, inconsistent counts */
#define KREG_E_UNSUPPORTED (-2) /* Natoms outside the compiled range, wrong device arch */
#define KREG_E_WORKSPACE (-3)   /* scratch buffer too small */
#define KREG_E_CUDA (-4)        /* a CUDA runtime / cuSOLVER call failed (see kreg_last_error) */

#define KREG_MIN_NATOMS 2
#define KREG_MAX_NATOMS 64 /* D = 2016.  The reference has no limit (Natoms is a run-time argument everywhere); this
                              one is where the shared-memory plans of the gradient blocks end, not an algorithmic one */
#define KREG_FUSED_MAX_NATOMS 24 /* up to here prediction keeps a test row's D + 1 accumulators in one warp's registers
                                    (one fused kernel); larger molecules take the tiled three-kernel path, which needs
                                    the scratch kreg_predict_workspace_bytes reports for EVERY batch size */

/* kreg_build_kernel_matrix `fill` argument: which part of the symmetric K is written */
#define KREG_FILL_LOWER 0 /* K[i][j], j <= i (row-major lower == column-major upper, potrf 'U') */
#define KREG_FILL_FULL 1  /* both triangles, as the reference does (KREG.f90:317-318)        */

/* kreg_predict `flags` */
#define KREG_PREDICT_ENERGY 1
#define KREG_PREDICT_GRADIENT 2

/* Library / device information ------------------------------------------------------- */

int kreg_abi_version(void);
/* Text of the last CUDA/cuSOLVER error seen by the calling thread ("" if none). */
const char *kreg_last_error(void);
/* Debug builds (make -C mlatom_b200/csrc debug -> libkreg_b200_dbg.so, -DKREG_DEBUG_BOUNDS) check every bulk-copy size /
 * alignment and every guarded global store inside the kernels; this returns the number of failed checks so far (and the
 * first failure's code and source line), synchronising the device.  Release builds return -1. */
long long kreg_debug_bounds_failures(int *first_code, int *first_line);
/* 0 if the current CUDA device can run the kernels (compute capability 10.x). */
int kreg_device_check(void);
/* Blocks the calling host thread until `stream` has drained (cudaStreamSynchronize): lets a ctypes caller finish a
 * latency-critical step (one MD step) without going through a framework's stream objects. */
int kreg_stream_synchronize(void *stream);

/* FP64 issue-rate probe used as the roofline denominator for the prediction kernel
 * (MEASURED_PEAKS.json holds no FP64 figure).  kind: 0 = DFMA, 1 = DMMA (mma.sync m8n8k4 f64).
 * Synchronous; writes TFLOP/s (best of `reps` launches of ~`ms_target` ms each). */
int kreg_measure_fp64_peak(int kind, int reps, double ms_target, double *tflops_host);

/* Descriptors -------------------------------------------------------------------------- */

/* Replaces KREG.kreg.get_xeq (KREG.f90:223-239): equilibrium pair distances.
 * req[Nat][3] -> xeq[D]. */
int kreg_xeq(int natoms, const double *req, double *xeq, void *stream);

/* Replaces KREG.kreg.calc_re_descriptors (KREG.f90:241-271; D_rel2eq.f90:17-83):
 * X[p][d] = xeq[d] / R_d(p).  Bit-identical to the reference (IEEE sqrt and divide, no FMA
 * contraction in R).  Needed for the model file; the assembly and prediction kernels below
 * build descriptors on the fly and never read X. */
int kreg_descriptors(int natoms, int64_t npoints, const double *xyz, const double *xeq, double *X,
                     void *stream);

/* Kernel-matrix assembly ------------------------------------------------------------------ */

/* Scratch bytes needed by kreg_build_kernel_matrix for a training set of `ntrain` geometries (centred descriptors,
 * and for gradient models the descriptor Jacobians plus 3*natoms+1 doubles per geometry pair).
 * kreg_kernel_block needs kreg_build_workspace_bytes(natoms, na, 0) + kreg_build_workspace_bytes(natoms, nb, 0). */
size_t kreg_build_workspace_bytes(int natoms, int64_t ntrain, int64_t ntrgr);
/* The same for a call that assembles only rows [row_begin, row_end) (a rank's slab in the sharded assembly): the
 * per-pair block then covers only the geometries whose gradient rows fall in the range, not all ntrain^2 pairs. */
size_t kreg_build_workspace_bytes_rows(int natoms, int64_t ntrain, int64_t ntrgr, int64_t row_begin, int64_t row_end);

/* Replaces KREG.kreg.calc_kernel_matrix + the diagonal shift of calcAlpha
 * (KREG.f90:293-346 and :616-622; A_KRR_kernel.f90:130-188, A_KRR.f90:905-918).
 * Builds rows [row_begin,row_end) of  K + diag(lambdav*1_NtrVal, lambdagradxyz*1_NtrGrXYZ)
 * directly from Cartesian coordinates (descriptor build fused into the call).
 *   ntrval  in {0, ntrain}; ntrgr in {0, 3*natoms*ntrain}  (kreg_api.py:52-72); M = ntrval+ntrgr.
 *   K points at element (0,0) of an M x ldk row-major matrix; only rows in the range are touched, so ranks can
 *   assemble disjoint row slabs of one matrix.  ldk >= M, except that with KREG_FILL_LOWER a slab may use
 *   a shorter row stride: at least row_end (rounded up to a whole 3*natoms block in the gradient part), the defined
 *   length of its last row -- slabs then travel between GPUs as trapezoids, not as M-wide rows.
 *   fill = KREG_FILL_LOWER writes j <= i only; KREG_FILL_FULL mirrors as the reference does (for gradient models
 *   only together with the whole row range; a partial range returns KREG_E_BADARG). */
int kreg_build_kernel_matrix(int natoms, int64_t ntrain, int64_t ntrval, int64_t ntrgr,
                             const double *xyz_train, const double *xeq, double sigma,
                             double lambdav, double lambdagradxyz, int64_t row_begin,
                             int64_t row_end, int fill, double *K, int64_t ldk, void *workspace,
                             size_t workspace_bytes, void *stream);

/* Rectangular value-value kernel block between two geometry sets (validation kernels for the
 * hyper-parameter search; MLatomF calc_Kvalidate, A_KRR_kernel.f90:190-223):
 * Kout[i][j] = k(X(xyz_a[i]), X(xyz_b[j])), row-major na x ldk. */
int kreg_kernel_block(int natoms, int64_t na, const double *xyz_a, int64_t nb, const double *xyz_b,
                      const double *xeq, double sigma, double *Kout, int64_t ldk, void *workspace,
                      size_t workspace_bytes, void *stream);

/* MLatomF's other kernel functions on the RE descriptor (Kfunk, A_KRR_kernel.f90:806-851; distance :2040-2074):
 *   KREG_KERNEL_GAUSSIAN     exp(-|Xi-Xj|_2^2 / (2 sigma^2))
 *   KREG_KERNEL_EXPONENTIAL  exp(-|Xi-Xj|_2 / sigma)
 *   KREG_KERNEL_MATERN       exp(-r/sigma) n!/(2n)! sum_k (n+k)!/(k!(n-k)!) (2r/sigma)^(n-k),  r = |Xi-Xj|_2,  n = nn <= 10
 *   KREG_KERNEL_LAPLACIAN    exp(-|Xi-Xj|_1 / sigma)
 * Value-value blocks only (calcKernel / calc_Kvalidate for models trained on values): Kout[i][j] = k(X(xyz_a[i]),
 * X(xyz_b[j])), row-major na x ldk, both triangles.  symmetric != 0 (then xyz_a and xyz_b must be the same set) puts
 * exactly 1 + lambda on the diagonal, i.e. the matrix kreg_solve factorises.  The three Euclidean kernels run on the
 * FP64 tensor-core tile kernel; the Laplacian's L1 distance is not a GEMM and has a plain tiled kernel. */
#define KREG_KERNEL_GAUSSIAN 0
#define KREG_KERNEL_EXPONENTIAL 1
#define KREG_KERNEL_MATERN 2
#define KREG_KERNEL_LAPLACIAN 3
size_t kreg_kernel_block_ex_workspace_bytes(int natoms, int64_t na, int64_t nb);
int kreg_kernel_block_ex(int kind, int nn, int natoms, int64_t na, const double *xyz_a, int64_t nb, const double *xyz_b,
                         const double *xeq, double sigma, int symmetric, double lambda, double *Kout, int64_t ldk,
                         void *workspace, size_t workspace_bytes, void *stream);

/* Models on those kernel functions (values-trained; MLatomF `kernel=exponential|Matern|Laplacian`, reached in the reference
 * through kreg(ml_program='MLatomF') + additional_mlatomf_args, models.py:494-495).  Training = kreg_kernel_block_ex
 * (symmetric, lambda) + kreg_solve.  Prediction of energies AND gradients for the three kernels of the Euclidean distance
 * runs on the fused two-GEMM tensor-core kernels up to KREG_FUSED_MAX_NATOMS atoms (the Gaussian model's kernels with a
 * different elementwise step) and on the tiled three-kernel path above:
 *   E[i] = prior + sum_j alpha_j k(r_ij)                                   (Yest_KRR, A_KRR_kernel.f90:225-245)
 *   G[i] = J_i^T sum_j alpha_j phi(r_ij) (X_j - X_i),  phi = -k'(r) / r     (YgradXYZest_KRR :731-744 -> dKdMiat_unpermuted
 *          :1407-1447 -> dKdXid :1178-1233)
 * Matern of order nn >= 1: the reference's analytic derivative; exponential (and Matern nn = 0): the reference takes a
 * central difference of Kfunk in descriptor space (:1222-1230), here the analytic phi = k / (sigma r), 0 at r = 0;
 * KREG_KERNEL_GAUSSIAN is accepted too (the same numbers as kreg_predict through this path).  The Laplacian kernel
 * (L1 distance, no GEMM form) predicts through kreg_kernel_block_ex + kreg_block_dot (energies) or
 * kreg_laplacian_predict (energies and analytic gradients). */
size_t kreg_radial_packed_model_bytes(int natoms, int64_t ntrain);
/* Packs the model in the layout kreg_predict_radial streams for this molecule size (the kernel width does not enter):
 * `packed` must be 256-byte aligned, kreg_radial_packed_model_bytes long; center[D] is written. */
int kreg_pack_model_radial(int natoms, int64_t ntrain, const double *xyz_train, const double *xeq,
                           const double *alpha_val, double *center, void *packed, size_t packed_bytes, void *stream);
/* Scratch for this batch: the three-kernel path's tiles above KREG_FUSED_MAX_NATOMS atoms; below, the partial
 * accumulators of the small-batch launch (the training set split over the chip, as in kreg_predict; 0 for batches
 * large enough to fill it; there a NULL or short workspace only means the batch runs unsplit). */
size_t kreg_predict_radial_workspace_bytes(int natoms, int64_t ntrain, int64_t npoints);
int kreg_predict_radial(int kind, int nn, int natoms, int64_t ntrain, const void *packed, const double *center,
                        const double *xeq, double sigma, double prior, int64_t npoints, const double *xyz, int flags,
                        double *E, double *G, void *workspace, size_t workspace_bytes, void *stream);
/* E[i] = prior + sum_j K[i][j] alpha[j] for a materialised kernel block (row-major na x ldk): deterministic. */
int kreg_block_dot(int64_t na, int64_t nb, const double *K, int64_t ldk, const double *alpha, double prior, double *E,
                   void *stream);
/* Laplacian-kernel model: energies and ANALYTIC Cartesian gradients from the materialised block K[i][j] =
 * exp(-|X_i - X_j|_1 / sigma) between `npoints` geometries and the training set (kreg_kernel_block_ex):
 *   E[i] = prior + sum_j alpha_j K_ij,   dE/dX_id = -(1/sigma) sum_j alpha_j K_ij sign(X_id - X_jd),   G = J_i^T dE/dX_i.
 * The reference has no closed form here: YgradXYZest_KRR (A_KRR_kernel.f90:731-744) reaches dKdXid's central difference of
 * Kfunk in descriptor space (:1222-1230), which approximates this derivative (identical away from the kinks X_id == X_jd).
 * X_train: DEVICE [ntrain][D] descriptors (kreg_descriptors); K row-major npoints x ldk. */
int kreg_laplacian_predict(int natoms, int64_t ntrain, const double *X_train, const double *alpha, const double *xeq,
                           double sigma, double prior, int64_t npoints, const double *xyz, const double *K, int64_t ldk,
                           int flags, double *E, double *G, void *stream);

/* Permutationally invariant kernel (values-trained models) --------------------------------------------------------
 * MLatomF `molDescrType=permuted permInvKernel` (A_KRR_kernel.f90:773-804 kernelFunction; descriptor calcX,
 * molDescr.f90:128-160 with permuteAtoms :197-220):
 *     k(M_i, M_j) = S_ij / (sqrt(S_ii) sqrt(S_jj)),   S_ij = sum_P exp(-|X_i - X_j^P|^2 / (2 sigma^2)),
 * X^P = RE descriptor of the geometry with its atoms permuted by the P-th permutation (and the equilibrium distances
 * permuted alike).  X^P is an entry permutation of X: X^P[d] = X[pidx[P][d]], pidx[P][d(a,c)] = d(p_P(a), p_P(c)) with
 * p_P the full 0-based atom map (mod_permInds, A_KRR_kernel.f90:944-950); pidx_inv is the table of the inverse maps.
 * pidx / pidx_inv: DEVICE int32 [nperm][D]; row 0 must be the identity (the reference pairs block 1 with every block).
 * nperm is limited by the shared memory of the self-term kernel, (2 D + nperm) doubles <= 200 KB (about 20 000 permutations
 * at 21 atoms); larger sets return KREG_E_UNSUPPORTED. */

/* Replaces calcX for the permuted descriptor: X[p][P][d] (the reference's X(:,p), Nperm blocks of D). */
int kreg_perm_descriptors(int natoms, int64_t npoints, const double *xyz, const double *xeq, int nperm,
                          const int32_t *pidx, double *X, void *stream);
/* S[p] = S_pp (KMiMi, A_KRR_kernel.f90:425-430) and, when dS != NULL, dS[p][a][t] = dS_pp / d xyz[p][a][t]
 * (dKMiMidMiat_part, :432-451). */
int kreg_perm_self_terms(int natoms, int64_t npoints, const double *xyz, const double *xeq, int nperm,
                         const int32_t *pidx, const int32_t *pidx_inv, double sigma, double *S, double *dS, void *stream);
/* Replaces calcKernel's value block / calc_Kvalidate with this kernel function (A_KRR_kernel.f90:149-159, 208-216):
 * Kout[i][j] = k(M_a[i], M_b[j]), row-major na x ldk, both triangles; symmetric != 0 (same set on both sides) puts
 * exactly 1 + lambda on the diagonal, i.e. the matrix kreg_solve factorises.  The S_ij^P terms run on the FP64 tensor-core
 * tile kernel in row slabs of at most 256 MB. */
size_t kreg_perm_kernel_block_workspace_bytes(int natoms, int64_t na, int64_t nb, int nperm);
int kreg_perm_kernel_block(int natoms, int64_t na, const double *xyz_a, int64_t nb, const double *xyz_b,
                           const double *xeq, int nperm, const int32_t *pidx, double sigma, int symmetric, double lambda,
                           double *Kout, int64_t ldk, void *workspace, size_t workspace_bytes, void *stream);
/* Prediction (Yest_KRR :225-245, YgradXYZest_KRR permInvKernel branch :414-503) runs through kreg_pack_model_x /
 * kreg_predict on the EXPANDED training set: the ntrain * nperm rows of kreg_perm_descriptors with the coefficients
 * alpha_expanded[j * nperm + P] = alpha[j] / sqrt(S_jj) written by kreg_perm_scale_alpha, prior 0, giving
 * E' = sum_j alpha_j S_ij / sqrt(S_jj) and G' = dE'/dxyz; kreg_perm_finish_prediction then forms, in place,
 *     E = prior + E' / sqrt(S_ii),    G = G' / sqrt(S_ii) - E' dS_ii / (2 S_ii^(3/2))        (:500-502).
 * E must hold E' also when only G is wanted; G may be NULL. */
int kreg_perm_scale_alpha(int64_t ntrain, int nperm, const double *alpha, const double *S, double *alpha_expanded,
                          void *stream);
int kreg_perm_finish_prediction(int natoms, int64_t npoints, const double *S, const double *dS, double prior, double *E,
                                double *G, void *stream);

/* Linear solve ------------------------------------------------------------------------------ */

/* Device / pinned-host scratch needed by kreg_solve for an m x m system. */
int kreg_solve_workspace_bytes(int64_t m, size_t *device_bytes, size_t *host_bytes);

/* Replaces mathUtils.solveSysLinEqs (mlatom/fortran/mathUtils.f90:8-104: dpotrf('U') +
 * dpotrs): in-place Cholesky of the matrix produced by kreg_build_kernel_matrix
 * (KREG_FILL_LOWER or _FULL) and solve for one right-hand side.  y is overwritten with alpha.
 * The only library call on the path (cuSOLVER potrf/potrs); timed separately by bench.py.
 * Synchronises the stream ONCE (to read potrf's `info`); the triangular solves are left running on `stream`.
 * Returns info (>0: leading minor not positive definite; K is destroyed) or a negative error. */
int kreg_solve(int64_t m, double *K, int64_t ldk, double *y_inout, void *device_workspace,
               size_t device_bytes, void *host_workspace, size_t host_bytes, void *stream);

/* Symmetric-indefinite fallback, replaces mathUtils.solveSysLinEqsBK (mlatom/fortran/mathUtils.f90:106-186: dsytrf +
 * dsytrs, entered from :24-26 when dpotrf fails).  K is a FRESHLY assembled matrix (kreg_build_kernel_matrix,
 * KREG_FILL_LOWER is enough) -- not the one kreg_solve destroyed: the reference hands Bunch-Kaufman the array potrf
 * already overwrote, which is not imitated.  y is overwritten with alpha, K is destroyed.
 *   method 1: Bunch-Kaufman, cuSOLVER sytrf + sytrs on the stored triangle (m <= 46000: sytrf has a 32-bit interface);
 *   method 2: pivoted LU, cuSOLVER getrf + getrs after mirroring the stored triangle in place;
 *   method 0: 1 when m allows it, else 2.
 * Returns 0, a negative error, or info > 0 (an exactly singular pivot). Synchronises the stream once to read info. */
int kreg_solve_indefinite_workspace_bytes(int64_t m, int method, size_t *device_bytes, size_t *host_bytes);
int kreg_solve_indefinite(int64_t m, double *K, int64_t ldk, double *y_inout, int method, void *device_workspace,
                          size_t device_bytes, void *host_workspace, size_t host_bytes, void *stream);

/* Prediction -------------------------------------------------------------------------------- */

/* Replaces nothing one-to-one: hoists the training-set half of the reference's gradient terms
 * out of the prediction loop:  v[j][d] = sum_{(a,t)} J_j[d,(a,t)] *
 * alpha_grad[j][a][t], J = descriptor Jacobian (KREG.f90:121,146). */
int kreg_precompute_v(int natoms, int64_t ntrain, const double *xyz_train, const double *xeq,
                      const double *alpha_grad, double *V, void *stream);

/* Size in bytes of the packed model consumed by kreg_predict. */
size_t kreg_packed_model_bytes(int natoms, int64_t ntrain, int has_gradient_terms);

/* Packs a trained model into the tensor-core fragment order the prediction kernel streams
 * through shared memory.  alpha_val may be NULL (gradient-only model: NtrVal = 0) and V may be
 * NULL (values-only model).  `center[D]` is written with the descriptor mean that is
 * subtracted from train and test descriptors alike (the kernel is invariant to it). */
int kreg_pack_model(int natoms, int64_t ntrain, const double *xyz_train, const double *xeq,
                    const double *alpha_val, const double *V, double sigma, double *center,
                    void *packed, size_t packed_bytes, void *stream);

/* Same as kreg_pack_model for a model that carries its training DESCRIPTORS X_train[ntrain][D] but no Cartesian
 * coordinates -- MLatomF's values-only .unf model files store exactly that (MLmodel.f90:880-886). */
int kreg_pack_model_x(int natoms, int64_t ntrain, const double *X_train, const double *alpha_val,
                      const double *V, double sigma, double *center, void *packed, size_t packed_bytes,
                      void *stream);

/* Scratch bytes kreg_predict can use for a batch of `npoints` geometries: small batches (fewer CTAs than half the
 * SMs -- the single-geometry MD / geometry-optimisation step) split the training set across the chip and reduce
 * the partial sums in a fixed order.  0 for batches large enough to fill the device. */
size_t kreg_predict_workspace_bytes(int natoms, int64_t ntrain, int has_gradient_terms, int64_t npoints);

/* Replaces KREG.kreg.predict (KREG.f90:534-596: Yest_KRR :348-366, YgradXYZest_KRR :368-391)
 * and MLatomF calcEst_KRR (A_KRR.f90:985-1029; A_KRR_kernel.f90:225-245, :327-412):
 *   E[i]       = prior + sum_j alpha_j k_ij + sum_j sum_bu alpha_j,bu dk_ij/dM_j,bu
 *   G[i][a][t] = dE_i / d xyz[i][a][t]            (gradient, not force: kreg_api.py:126-131)
 * for npoints geometries, energies and gradients in ONE pass over the training set.
 * flags: KREG_PREDICT_ENERGY | KREG_PREDICT_GRADIENT; E or G may be NULL when not requested.
 * workspace may be NULL (or smaller than kreg_predict_workspace_bytes): the batch then runs unsplit. */
int kreg_predict(int natoms, int64_t ntrain, const void *packed, const double *center,
                 const double *xeq, double sigma, double prior, int has_gradient_terms,
                 int64_t npoints, const double *xyz, int flags, double *E, double *G,
                 void *workspace, size_t workspace_bytes, void *stream);

/* Replaces YhessXYZest_KRR (KREG.f90:393-420 with d2KdMiatdMibu :180-221; predict :588-594): the Hessian of the
 * predicted energy exactly as the reference defines it -- values part of alpha only, products of first derivatives of
 * the descriptor only (no d2X/dM2 term).  X_train[ntrain][D] are the plain RE descriptors (kreg_descriptors),
 * alpha_val the first NtrVal coefficients (ntrain = 0 gives zeros, as the reference does for gradient-only models).
 * H[p][3Nat][3Nat], symmetric. */
int kreg_hessian(int natoms, int64_t ntrain, const double *X_train, const double *alpha_val, const double *xeq,
                 double sigma, int64_t npoints, const double *xyz, double *H, void *stream);

/* Batched molecular dynamics around kreg_predict (velocity Verlet exactly as mlatom/md.py:158-205; batch driver
 * md_parallel.py:155-200), all arrays device resident: xyz, vel, grad [ntraj][natoms][3]; acc_scale[3*natoms] = c_acc / m_a
 * and kin_scale[3*natoms] = c_kin * m_a / 2 per Cartesian component (the reference's unit factors, md.py:171 and
 * data.py:862, folded in by the caller); `active` (may be NULL) holds 1.0 for running and 0.0 for frozen trajectories.
 *   first half : a = -grad * acc_scale;  x += v dt + a dt^2/2;  v += a dt/2          (md.py:190-193)
 *   second half: v += a_new dt/2;  ekin[p] = sum kin_scale v^2                        (md.py:204, data.py:862)
 *   uncertainty: uq = |epot - eaux|; a running trajectory with uq > threshold is frozen and stop_step[p] = *step_no
 *                (the sampler's stop function, al_utils.py:246-322, :1418-1424) */
int kreg_md_first_half(int natoms, int64_t ntraj, double dt, const double *acc_scale, const double *grad, double *xyz,
                       double *vel, const double *active, void *stream);
int kreg_md_second_half(int natoms, int64_t ntraj, double dt, const double *acc_scale, const double *kin_scale,
                        const double *grad, double *vel, double *ekin, const double *active, void *stream);
int kreg_md_uncertainty(int64_t ntraj, const double *epot, const double *eaux, double threshold, const int64_t *step_no,
                        double *uq, double *active, int64_t *stop_step, void *stream);

/* Test geometries one CTA of kreg_predict handles for this molecule size / model kind (<0: unsupported size).
 * Callers that split a batch into chunks (host streaming, multi-stream pipelines) should use multiples of
 * SM count x this value so that every chunk fills whole waves. */
int kreg_predict_tile_rows(int natoms, int has_gradient_terms);

#ifdef __cplusplus
}
#endif
#endif /* KREG_B200_H */
