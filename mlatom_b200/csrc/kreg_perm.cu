// kreg_perm.cu -- MLatomF's permutationally invariant kernel on the permuted RE descriptor (values-trained models).
//
// Reference: molDescrType=permuted + permInvKernel.  The descriptor of a geometry is the concatenation of the RE
// descriptors of its Nperm atom-permuted copies (calcX, MLatomF_src/molDescr.f90:128-160; permuteAtoms :197-220; the
// equilibrium distances are permuted alike, :82-100) and the kernel is
//     k(M_i, M_j) = S_ij / (sqrt(S_ii) sqrt(S_jj)),   S_ij = sum_P exp(-|X_i - X_j^P|^2 / (2 sigma^2))
// (kernelFunction, A_KRR_kernel.f90:773-804), prediction Yest_KRR :225-245 and the permInvKernel branch of
// YgradXYZest_KRR :414-503.
//
// B200 formulation.  Block P of the permuted descriptor is an ENTRY PERMUTATION of the plain one:
// X^P[d(a,c)] = Req(p a, p c) / R(p a, p c) = X[pidx[P][d]] with pidx[P][d(a,c)] = d(p(a), p(c)), bit for bit.  So
//   * S_ij over a tile of geometries is the ordinary Gaussian block between the plain rows of set A and the Nperm-times
//     expanded rows of set B: it runs on the FP64 tensor-core tile kernel of the K assembly (values_ring_kernel), followed
//     by a reduction over P (summed in the reference's order) and the normalisation;
//   * a trained model predicts through the ordinary prediction kernels with an expanded training set: Ntrain * Nperm
//     descriptor rows X_j^P carrying alpha_j / sqrt(S_jj) each give E' = sum_j alpha_j S_ij / sqrt(S_jj) and its gradient G';
//     then E = prior + E' / sqrt(S_ii) and dE = G' / sqrt(S_ii) - E' dS_ii / (2 S_ii^(3/2))  (the quotient rule of :500-502),
//     with S_ii and dS_ii/dM_i from a small per-geometry kernel here.
#include "kreg_common.cuh"
#include "kreg_gemm.h"
#include "kreg_internal.h"

namespace kreg {

// Expanded, centred, zero-padded descriptor rows: row p * nperm + P, element d = X_p[pidx[P][d]] - center[d].
__global__ void __launch_bounds__(256) prep_descr_perm_kernel(int nat, int64_t n, const double *__restrict__ xyz,
                                                              const double *__restrict__ xeq,
                                                              const double *__restrict__ center, int XS, int nperm,
                                                              const int *__restrict__ pidx, double *__restrict__ Xc) {
  const int D = descr_size(nat);
  const int64_t total = n * nperm * XS;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / XS;
    const int d = (int)(e - r * XS);
    const int64_t p = r / nperm;
    const int P = (int)(r - p * nperm);
    double v = 0.0;
    if (d < D) {
      const int pd = pidx[P * D + d];
      int a, b;
      pair_from_index(nat, pd, a, b);
      const double *m = xyz + p * nat * 3;
      v = xeq[pd] / sqrt(rij2_exact(m + 3 * a, m + 3 * b)) - center[d];
    }
    Xc[e] = v;
  }
}

// The permuted descriptor itself, X[p][P][d] (MLatomF's X(:,p), molDescr.f90:146-158): what the model file stores and what
// the expanded prediction model is packed from.
__global__ void __launch_bounds__(256) perm_descriptors_kernel(int nat, int64_t n, const double *__restrict__ xyz,
                                                               const double *__restrict__ xeq, int nperm,
                                                               const int *__restrict__ pidx, double *__restrict__ X) {
  const int D = descr_size(nat);
  const int64_t total = n * nperm * D;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / D;
    const int d = (int)(e - r * D);
    const int64_t p = r / nperm;
    const int P = (int)(r - p * nperm);
    const int pd = pidx[P * D + d];
    int a, b;
    pair_from_index(nat, pd, a, b);
    const double *m = xyz + p * nat * 3;
    X[e] = xeq[pd] / sqrt(rij2_exact(m + 3 * a, m + 3 * b));
  }
}

// Self terms of one geometry per CTA: S = sum_P k_P, k_P = exp(-|X - X^P|^2 / (2 sigma^2)) (KMiMi, A_KRR_kernel.f90:425-430)
// and, optionally, dS/dM (dKMiMidMiat_part, :432-451):
//   dS/dx_(a,t) = 1/sigma^2 sum_{c != a} w_d X_d (x_c - x_a)_t / R_ac^2,   w_d = sum_P k_P (X[pi_P d] + X[pi_P^-1 d] - 2 X_d).
// Warps take permutations round-robin (warp-shuffle reductions, no CTA barrier per permutation); the sums over P run in
// the reference's order.
__global__ void __launch_bounds__(128) perm_self_kernel(int nat, int64_t n, const double *__restrict__ xyz,
                                                        const double *__restrict__ xeq, int nperm,
                                                        const int *__restrict__ pidx, const int *__restrict__ pinv,
                                                        double inv_2s2, double inv_s2, double *__restrict__ S,
                                                        double *__restrict__ dS) {
  extern __shared__ double sm[];
  const int D = descr_size(nat);
  double *X = sm;             // [D]
  double *kP = X + D;         // [nperm]
  double *ed = kP + nperm;    // [D]   w_d X_d / R_d^2 / sigma^2
  const int64_t p = blockIdx.x;
  const double *m = xyz + p * nat * 3;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  for (int d = tid; d < D; d += blockDim.x) {
    int a, b;
    pair_from_index(nat, d, a, b);
    X[d] = xeq[d] / sqrt(rij2_exact(m + 3 * a, m + 3 * b));
  }
  __syncthreads();
  for (int P = warp; P < nperm; P += nwarps) {
    const int *pi = pidx + (size_t)P * D;
    double q = 0.0;
    for (int d = lane; d < D; d += 32) {
      const double t = X[d] - X[pi[d]];
      q = fma(t, t, q);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (lane == 0) kP[P] = exp(-q * inv_2s2);
  }
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int P = 0; P < nperm; P++) s += kP[P];
    S[p] = s;
  }
  if (!dS) return;
  for (int d = tid; d < D; d += blockDim.x) {
    double w = 0.0;
    const double xd = X[d];
    for (int P = 0; P < nperm; P++) w = fma(kP[P], X[pidx[(size_t)P * D + d]] + X[pinv[(size_t)P * D + d]] - 2.0 * xd, w);
    int a, b;
    pair_from_index(nat, d, a, b);
    ed[d] = w * xd / rij2_exact(m + 3 * a, m + 3 * b) * inv_s2;
  }
  __syncthreads();
  for (int o = tid; o < 3 * nat; o += blockDim.x) {
    const int a = o / 3, t = o - 3 * a;
    double g = 0.0;
    for (int c = 0; c < nat; c++) {
      if (c == a) continue;
      const int d = a < c ? pair_index(nat, a, c) : pair_index(nat, c, a);
      g = fma(ed[d], m[3 * c + t] - m[3 * a + t], g);
    }
    dS[p * 3 * nat + o] = g;
  }
}

// K[i][j] = sum_P G[i][j * nperm + P] / (sqrt(S_i) sqrt(S_j)); exactly 1 + lambda where a symmetric block meets its diagonal.
__global__ void __launch_bounds__(256) perm_reduce_kernel(int64_t rows, int64_t r0, int64_t nb, int nperm,
                                                          const double *__restrict__ G, int64_t ldg,
                                                          const double *__restrict__ SA, const double *__restrict__ SB,
                                                          int symmetric, double lambda, double *__restrict__ K,
                                                          int64_t ldk) {
  const int64_t total = rows * nb;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / nb, j = e - i * nb;
    const double *g = G + i * ldg + j * nperm;
    double s = 0.0;
    for (int P = 0; P < nperm; P++) s += g[P];
    const int64_t gi = r0 + i;
    K[gi * ldk + j] = (symmetric && gi == j) ? 1.0 + lambda : s / (sqrt(SA[gi]) * sqrt(SB[j]));
  }
}

// scale[j * nperm + P] = alpha[j] / sqrt(S[j]): the expanded model's coefficients
__global__ void perm_scale_alpha_kernel(int64_t n, int nperm, const double *__restrict__ alpha,
                                        const double *__restrict__ S, double *__restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n * nperm) {
    const int64_t j = e / nperm;
    out[e] = alpha[j] / sqrt(S[j]);
  }
}

// One warp per geometry: E = prior + E' / sqrt(S),  dE = G' / sqrt(S) - E' dS / (2 S sqrt(S)), in place.
__global__ void __launch_bounds__(128) perm_finish_kernel(int nat, int64_t n, const double *__restrict__ S,
                                                          const double *__restrict__ dS, double prior,
                                                          double *__restrict__ E, double *__restrict__ G) {
  const int64_t p = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= n) return;
  const int lane = threadIdx.x & 31;
  const double s = S[p], rs = 1.0 / sqrt(s), e_raw = E[p];
  __syncwarp();
  if (G) {
    const double c = 0.5 * e_raw * rs / s;
    for (int o = lane; o < 3 * nat; o += 32) G[p * 3 * nat + o] = G[p * 3 * nat + o] * rs - c * dS[p * 3 * nat + o];
  }
  if (lane == 0) E[p] = prior + e_raw * rs;
}

struct PermWs {
  int XS, nchD;
  int64_t slab, ldg;
  size_t off_center, off_XA, off_XAch, off_bA, off_XB, off_XBch, off_bB, off_bBc, off_SA, off_SB, off_G, total;
};

constexpr size_t kPermGBytes = (size_t)256 << 20;  // rectangular S_ij^P block held at a time

static PermWs perm_ws(int nat, int64_t na, int64_t nb, int nperm) {
  PermWs w;
  const int D = descr_size(nat), KS = (D + 3) / 4;
  w.nchD = (KS + kGemmKC - 1) / kGemmKC;
  w.XS = w.nchD * 4 * kGemmKC;
  if (w.XS % 8 == 0) w.XS += 4;
  const int64_t nbe = nb * nperm;
  w.ldg = (nbe + 1) & ~(int64_t)1;
  int64_t slab = (int64_t)(kPermGBytes / ((size_t)w.ldg * sizeof(double))) / 128 * 128;
  if (slab < 128) slab = 128;
  if (slab > na) slab = na;
  w.slab = slab;
  size_t off = 0;
  auto take = [&](size_t doubles) {
    size_t o = off;
    off += (doubles * sizeof(double) + 255) / 256 * 256;
    return o;
  };
  w.off_center = take(D);
  w.off_XA = take((size_t)na * w.XS);
  w.off_XAch = take((size_t)w.nchD * na * 4 * kGemmKC);
  w.off_bA = take((size_t)na + 2);
  w.off_XB = take((size_t)nbe * w.XS);
  w.off_XBch = take((size_t)w.nchD * nbe * 4 * kGemmKC);
  w.off_bB = take((size_t)nbe + 2);
  w.off_bBc = take((size_t)nbe + 2);
  w.off_SA = take((size_t)na + 2);
  w.off_SB = take((size_t)nb + 2);
  w.off_G = take((size_t)slab * w.ldg);
  w.total = off;
  return w;
}

static int launch_perm_self(int nat, int64_t n, const double *xyz, const double *xeq, int nperm, const int *pidx,
                            const int *pinv, double sigma, double *S, double *dS, cudaStream_t st) {
  if (n <= 0) return KREG_OK;
  const int D = descr_size(nat);
  const size_t smem = (size_t)(2 * D + nperm) * sizeof(double);
  if (smem > 200 * 1024) return KREG_E_UNSUPPORTED;
  KREG_CUDA_TRY(cudaFuncSetAttribute(perm_self_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const double inv_s2 = 1.0 / (sigma * sigma);
  perm_self_kernel<<<(unsigned)n, 128, smem, st>>>(nat, n, xyz, xeq, nperm, pidx, pinv, 0.5 * inv_s2, inv_s2, S, dS);
  KREG_LAUNCH_CHECK();
  return KREG_OK;
}

static bool perm_args_ok(int natoms, int nperm, const int *pidx) {
  return natoms >= KREG_MIN_NATOMS && natoms <= KREG_MAX_NATOMS && nperm >= 1 && pidx != nullptr;
}

}  // namespace kreg

using namespace kreg;

extern "C" int kreg_perm_descriptors(int natoms, int64_t npoints, const double *xyz, const double *xeq, int nperm,
                                     const int32_t *pidx, double *X, void *stream) {
  if (natoms < KREG_MIN_NATOMS || natoms > KREG_MAX_NATOMS) return KREG_E_UNSUPPORTED;
  if (!perm_args_ok(natoms, nperm, pidx) || npoints < 0 || !xeq) return KREG_E_BADARG;
  if (npoints == 0) return KREG_OK;
  if (!xyz || !X) return KREG_E_BADARG;
  const int64_t total = npoints * nperm * descr_size(natoms);
  const int blocks = (int)(total / 256 + 1 < 148 * 16 ? total / 256 + 1 : 148 * 16);
  perm_descriptors_kernel<<<blocks, 256, 0, as_stream(stream)>>>(natoms, npoints, xyz, xeq, nperm, pidx, X);
  KREG_LAUNCH_CHECK();
  return KREG_OK;
}

extern "C" int kreg_perm_self_terms(int natoms, int64_t npoints, const double *xyz, const double *xeq, int nperm,
                                    const int32_t *pidx, const int32_t *pidx_inv, double sigma, double *S, double *dS,
                                    void *stream) {
  if (natoms < KREG_MIN_NATOMS || natoms > KREG_MAX_NATOMS) return KREG_E_UNSUPPORTED;
  if (!perm_args_ok(natoms, nperm, pidx) || npoints < 0 || !xeq || !(sigma > 0.0) || (dS && !pidx_inv)) return KREG_E_BADARG;
  if (npoints == 0) return KREG_OK;
  if (!xyz || !S) return KREG_E_BADARG;
  return launch_perm_self(natoms, npoints, xyz, xeq, nperm, pidx, pidx_inv, sigma, S, dS, as_stream(stream));
}

extern "C" size_t kreg_perm_kernel_block_workspace_bytes(int natoms, int64_t na, int64_t nb, int nperm) {
  if (natoms < KREG_MIN_NATOMS || natoms > KREG_MAX_NATOMS || na < 0 || nb < 0 || nperm < 1) return 0;
  return perm_ws(natoms, na, nb, nperm).total;
}

extern "C" int kreg_perm_kernel_block(int natoms, int64_t na, const double *xyz_a, int64_t nb, const double *xyz_b,
                                      const double *xeq, int nperm, const int32_t *pidx, double sigma, int symmetric,
                                      double lambda, double *Kout, int64_t ldk, void *workspace, size_t workspace_bytes,
                                      void *stream) {
  if (natoms < KREG_MIN_NATOMS || natoms > KREG_MAX_NATOMS) return KREG_E_UNSUPPORTED;
  if (!perm_args_ok(natoms, nperm, pidx) || na < 0 || nb < 0 || !xeq || !(sigma > 0.0) || ldk < nb) return KREG_E_BADARG;
  if (symmetric && na != nb) return KREG_E_BADARG;
  if (na == 0 || nb == 0) return KREG_OK;
  if (!xyz_a || !xyz_b || !Kout || !workspace) return KREG_E_BADARG;
  const PermWs w = perm_ws(natoms, na, nb, nperm);
  if (workspace_bytes < w.total) return KREG_E_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  char *ws = static_cast<char *>(workspace);
  auto at = [&](size_t off) { return reinterpret_cast<double *>(ws + off); };
  double *center = at(w.off_center);
  const double c1 = kExpScale / (sigma * sigma);
  const int64_t nbe = nb * nperm;
  int rc = launch_descriptor_mean(natoms, nb, xyz_b, xeq, center, st);
  if (rc) return rc;
  rc = launch_prep_rows(natoms, na, xyz_a, xeq, center, w.XS, c1, at(w.off_XA), at(w.off_bA), nullptr, st);
  if (rc) return rc;
  {
    const int64_t total = nbe * w.XS;
    const int blocks = (int)(total / 256 + 1 < 148 * 16 ? total / 256 + 1 : 148 * 16);
    prep_descr_perm_kernel<<<blocks, 256, 0, st>>>(natoms, nb, xyz_b, xeq, center, w.XS, nperm, pidx, at(w.off_XB));
    KREG_LAUNCH_CHECK();
  }
  rc = launch_prep_bias(natoms, nbe, w.XS, c1, at(w.off_XB), at(w.off_bB), at(w.off_bBc), st);
  if (rc) return rc;
  rc = launch_chunk_rows(na, w.XS, kGemmKC, w.nchD, 4 * kGemmKC, at(w.off_XA), at(w.off_XAch), st);
  if (rc) return rc;
  rc = launch_chunk_rows(nbe, w.XS, kGemmKC, w.nchD, 4 * kGemmKC, at(w.off_XB), at(w.off_XBch), st);
  if (rc) return rc;
  // self terms S_ii of both sets (pidx_inv is only needed for their gradients)
  rc = launch_perm_self(natoms, na, xyz_a, xeq, nperm, pidx, nullptr, sigma, at(w.off_SA), nullptr, st);
  if (rc) return rc;
  rc = launch_perm_self(natoms, nb, xyz_b, xeq, nperm, pidx, nullptr, sigma, at(w.off_SB), nullptr, st);
  if (rc) return rc;
  for (int64_t r0 = 0; r0 < na; r0 += w.slab) {
    const int64_t rows = na - r0 < w.slab ? na - r0 : w.slab;
    ValuesArgs a = {};
    a.XA = at(w.off_XA) + r0 * w.XS;
    a.XB = at(w.off_XB);
    a.XAch = at(w.off_XAch) + r0 * 4 * kGemmKC;
    a.XBch = at(w.off_XBch);
    a.XSC = 4 * kGemmKC;
    a.XS = w.XS;
    a.na = rows;
    a.nb = nbe;
    a.strideA = na;
    a.strideB = nbe;
    a.nchunks = w.nchD;
    a.c1 = c1;
    a.all_cols = 1;
    a.row_begin = 0;
    a.row_end = rows;
    a.K = at(w.off_G);
    a.ldk = w.ldg;
    a.biasA = at(w.off_bA) + r0;
    a.biasB = at(w.off_bBc);
    rc = launch_gemm_epi(a, EPI_K, st);
    if (rc) return rc;
    const int64_t total = rows * nb;
    const int blocks = (int)(total / 256 + 1 < 148 * 16 ? total / 256 + 1 : 148 * 16);
    perm_reduce_kernel<<<blocks, 256, 0, st>>>(rows, r0, nb, nperm, at(w.off_G), w.ldg, at(w.off_SA), at(w.off_SB),
                                               symmetric, lambda, Kout, ldk);
    KREG_LAUNCH_CHECK();
  }
  return KREG_OK;
}

extern "C" int kreg_perm_scale_alpha(int64_t ntrain, int nperm, const double *alpha, const double *S, double *alpha_expanded,
                                     void *stream) {
  if (ntrain < 0 || nperm < 1) return KREG_E_BADARG;
  if (ntrain == 0) return KREG_OK;
  if (!alpha || !S || !alpha_expanded) return KREG_E_BADARG;
  const int64_t total = ntrain * nperm;
  perm_scale_alpha_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(ntrain, nperm, alpha, S,
                                                                                         alpha_expanded);
  KREG_LAUNCH_CHECK();
  return KREG_OK;
}

extern "C" int kreg_perm_finish_prediction(int natoms, int64_t npoints, const double *S, const double *dS, double prior,
                                           double *E, double *G, void *stream) {
  if (natoms < KREG_MIN_NATOMS || natoms > KREG_MAX_NATOMS) return KREG_E_UNSUPPORTED;
  if (npoints < 0) return KREG_E_BADARG;
  if (npoints == 0) return KREG_OK;
  if (!S || !E || (G && !dS)) return KREG_E_BADARG;
  perm_finish_kernel<<<(unsigned)((npoints + 3) / 4), 128, 0, as_stream(stream)>>>(natoms, npoints, S, dS, prior, E, G);
  KREG_LAUNCH_CHECK();
  return KREG_OK;
}
