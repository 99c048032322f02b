// kreg_predict_big.cu -- energy + gradient prediction for molecules ABOVE the register-resident kernels' range
// (Natoms > KREG_FUSED_MAX_NATOMS = 24, D > 276).  The reference has no size limit anywhere on the path
// (mlatom/fortran/KREG.f90:34-52, 241-271, 534-596 take Natoms at run time); neither has this.
//
// Same algebra as kreg_predict.cu, but the per-test-row accumulator A_i (D + 1 doubles) no
// longer fits a warp's registers, so the pass over the training set is tiled in three kernels per tile of TN test
// geometries, all on the DMMA tile kernel of the K assembly (values_ring_kernel, kreg_kbuild.cu):
//     GEMM1   S = X'te . X'tr^T (K = D), epilogue w_ij = alpha_j exp(c1 (S - n_i/2 - n_j/2)) written chunk-major
//             (gradient-trained models: first C2 = alpha'_j + X'te . (V/s^2)^T, then w = k C2 and k itself is kept)
//     GEMM2   A = W . [X'tr | 1] (+ K . [V | 0])  (K = Ntrain, N = D + 1; the ones column accumulates E)
//     epilogue per geometry: E = prior + A[D], g = (A[:D] - X'_i A[D]) / s^2, dE/dM = J_i^T g   (KREG.f90:121 Jacobian)
// Fusing the two GEMMs buys nothing at this size: W makes one round trip through HBM (16 bytes per geometry pair)
// against 4 D >= 1200 tensor-core flops per pair -- 2.5 ps of HBM time against >= 32 ps of FP64 time.
#include "kreg_common.cuh"
#include "kreg_gemm.h"
#include "kreg_internal.h"

namespace kreg {

constexpr int kCW = 4 * kGemmKC;  // chunk width (36 doubles) of every chunk-major operand

struct BigModel {  // byte offsets inside the packed blob
  int D, XS, nchD, nbT, nchW;
  size_t off_Xc, off_Xch, off_bcol, off_scale, off_Vs, off_Vsch, off_XT, total;
};

static size_t pad256(size_t bytes) { return (bytes + 255) / 256 * 256; }

static BigModel big_model(int nat, int64_t ntr, bool has_v) {
  BigModel m;
  m.D = descr_size(nat);
  const int KS = (m.D + 3) / 4;
  m.nchD = (KS + kGemmKC - 1) / kGemmKC;
  m.XS = m.nchD * kCW;
  if (m.XS % 8 == 0) m.XS += 4;
  m.nbT = m.D + 1;
  m.nchW = (int)((ntr + kCW - 1) / kCW);
  size_t off = 0;
  auto take = [&](size_t doubles) {
    size_t o = off;
    off += pad256(doubles * sizeof(double));
    return o;
  };
  m.off_Xc = take((size_t)ntr * m.XS);
  m.off_Xch = take((size_t)m.nchD * ntr * kCW);
  m.off_bcol = take((size_t)ntr + 2);
  m.off_scale = take((size_t)ntr + 2);
  m.off_Vs = has_v ? take((size_t)ntr * m.XS) : off;
  m.off_Vsch = has_v ? take((size_t)m.nchD * ntr * kCW) : off;
  m.off_XT = take((size_t)m.nchW * (has_v ? 2 : 1) * m.nbT * kCW);
  m.total = off;
  return m;
}

struct BigWs {  // byte offsets inside the per-call scratch, sized for tiles of TN test geometries
  int64_t TN;
  int NPAD;
  int64_t ldc;
  size_t off_Xc, off_Xch, off_bias, off_bias2, off_zero, off_C2, off_W, off_acc, off_accE, total;
};

static BigWs big_ws(const BigModel &m, int64_t ntr, bool has_v, int64_t npoints, int sms, bool radial = false) {
  BigWs w;
  // GEMM2 runs (TN / 128) x ceil((D + 1) / 64) CTA tiles, two per SM: take whole waves of them, as many as keep the
  // chunk-major W (and K) under ~1.5 GB; small batches take what they need
  const int64_t ncol = (m.nbT + 63) / 64;
  int64_t row_tiles = (2 * (int64_t)sms) / ncol;
  if (row_tiles < 1) row_tiles = 1;
  const size_t w_row_bytes = (size_t)m.nchW * kCW * sizeof(double) * (has_v ? 3 : radial ? 2 : 1);  // W (+ K + C2) per test row
  int64_t waves = (int64_t)((size_t)1536 << 20) / (int64_t)(w_row_bytes * 128 * row_tiles);
  if (waves < 1) {
    waves = 1;
    row_tiles = (int64_t)((size_t)1536 << 20) / (int64_t)(w_row_bytes * 128);
    if (row_tiles < 1) row_tiles = 1;
  }
  if (waves > 4) waves = 4;
  int64_t TN = 128 * row_tiles * waves;
  const int64_t need = (npoints + 127) / 128 * 128;
  if (TN > need) TN = need;
  if (TN < 128) TN = 128;
  w.TN = TN;
  w.NPAD = (m.nbT + 1) & ~1;
  w.ldc = (ntr + 1) & ~(int64_t)1;
  size_t off = 0;
  auto take = [&](size_t doubles) {
    size_t o = off;
    off += pad256(doubles * sizeof(double));
    return o;
  };
  w.off_Xc = take((size_t)TN * m.XS);
  w.off_Xch = take((size_t)m.nchD * TN * kCW);
  w.off_bias = take((size_t)TN + 2);
  w.off_bias2 = radial ? take((size_t)TN + 2) : off;
  w.off_zero = take((size_t)m.nbT + 66);
  w.off_C2 = has_v ? take((size_t)TN * w.ldc) : off;
  w.off_W = take((size_t)m.nchW * ((has_v || radial) ? 2 : 1) * TN * kCW);
  w.off_acc = take((size_t)TN * w.NPAD);
  w.off_accE = radial ? take((size_t)TN * 2) : off;
  w.total = off;
  return w;
}

// Xc[j][d] = X[j][d] - center[d] (zero padded) for models that carry descriptors instead of geometries
__global__ void __launch_bounds__(256) big_center_rows_kernel(int D, int64_t n, int XS, const double *__restrict__ X,
                                                              const double *__restrict__ center, double *__restrict__ Xc) {
  const int64_t total = n * XS;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = e / XS;
    const int d = (int)(e - p * XS);
    Xc[e] = d < D ? X[p * D + d] - center[d] : 0.0;
  }
}

// gradient-trained models: Vs[j][d] = v_j[d] / s^2 (zero padded rows), scale[j] = alpha'_j = alpha_j - X'_j . v_j / s^2
__global__ void __launch_bounds__(256) big_prep_v_kernel(int D, int64_t n, int XS, const double *__restrict__ Xc,
                                                         const double *__restrict__ V, const double *__restrict__ alpha,
                                                         double inv_s2, double *__restrict__ Vs,
                                                         double *__restrict__ scale) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int d = 0; d < D; d++) {
      const double v = V[j * D + d];
      Vs[j * XS + d] = v * inv_s2;
      s += Xc[j * XS + d] * v;
    }
    for (int d = D; d < XS; d++) Vs[j * XS + d] = 0.0;
    scale[j] = (alpha ? alpha[j] : 0.0) - s * inv_s2;
  }
}

__global__ void __launch_bounds__(256) big_copy_scale_kernel(int64_t n, const double *__restrict__ alpha,
                                                             double *__restrict__ scale) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x)
    scale[j] = alpha ? alpha[j] : 0.0;
}

// B operand of GEMM2, chunk-major and transposed: XT[c][d][jj] = src[c * 36 + jj][d] for d < D; the row d == D holds
// `ones` (1 for the descriptor part: accumulates sum_j w_ij; 0 for the V part); zero beyond Ntrain.
__global__ void __launch_bounds__(256) big_transpose_kernel(int D, int64_t n, int nchW, const double *__restrict__ src,
                                                            int64_t ld, double ones, double *__restrict__ XT) {
  const int nbT = D + 1;
  const int64_t total = (int64_t)nchW * nbT * kCW;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int jj = (int)(e % kCW);
    const int64_t r = e / kCW;
    const int d = (int)(r % nbT);
    const int64_t j = (r / nbT) * kCW + jj;
    double v = 0.0;
    if (j < n) v = d < D ? src[j * ld + d] : ones;
    XT[e] = v;
  }
}

// One warp per test geometry, a lane per atom (atoms a, a + 32, ...): the same arithmetic, in the same order, as
// predict_row_epilogue_warp (kreg_predict_kernel.cuh).
__global__ void __launch_bounds__(256) big_epilogue_kernel(int nat, int64_t n, const double *__restrict__ xyz,
                                                           const double *__restrict__ xeq,
                                                           const double *__restrict__ Xc, int XS,
                                                           const double *__restrict__ acc, int ldacc, int ecol,
                                                           double inv_s2, double prior, double *__restrict__ E,
                                                           double *__restrict__ G,
                                                           const double *__restrict__ accE = nullptr, int ldaccE = 0) {
  const int lane = threadIdx.x & 31;
  const int64_t p = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (p >= n) return;
  const double *y = acc + p * ldacc;
  // Gaussian: the weights of the energy and of the gradient are the same numbers (w = alpha k); the radial kernels carry
  // the energy sums in an array of their own (accE) and the sum of the gradient weights in column `ecol`
  if (lane == 0 && E) E[p] = prior + (accE ? accE[p * ldaccE] : y[ecol]);
  if (!G) return;
  const double esum = y[ecol];
  const double *m = xyz + p * nat * 3;
  const double *xc = Xc + p * XS;
  for (int a = lane; a < nat; a += 32) {
    double gr[3] = {0.0, 0.0, 0.0};
    for (int c = 0; c < nat; c++) {
      if (c == a) continue;
      const int lo = a < c ? a : c, hi = a < c ? c : a;
      const int d = pair_index(nat, lo, hi);
      const double r2 = rij2_exact(m + 3 * lo, m + 3 * hi);
      const double x = __ldg(xeq + d) / sqrt(r2);
      const double gd = (y[d] - xc[d] * esum) * inv_s2;
      const double f = gd * x / r2;
#pragma unroll
      for (int t = 0; t < 3; t++) {
        const double dx = m[3 * hi + t] - m[3 * lo + t];
        gr[t] = fma(a == lo ? f : -f, dx, gr[t]);
      }
    }
    double *out = G + p * nat * 3 + 3 * a;
    out[0] = gr[0];
    out[1] = gr[1];
    out[2] = gr[2];
  }
}

static int grid_for(int64_t total) {
  const int64_t b = (total + 255) / 256;
  return (int)(b < 148 * 16 ? (b < 1 ? 1 : b) : 148 * 16);
}

size_t big_packed_model_bytes(int nat, int64_t ntr, bool has_v) { return big_model(nat, ntr, has_v).total; }

int big_pack_model(int nat, int64_t ntr, const double *xyz_train, const double *X_train, const double *xeq,
                   const double *alpha_val, const double *V, double sigma, double *center, void *packed,
                   size_t packed_bytes, cudaStream_t st) {
  const bool has_v = V != nullptr;
  const BigModel m = big_model(nat, ntr, has_v);
  if (packed_bytes < m.total) return KREG_E_WORKSPACE;
  char *base = static_cast<char *>(packed);
  double *Xc = reinterpret_cast<double *>(base + m.off_Xc), *Xch = reinterpret_cast<double *>(base + m.off_Xch);
  double *bcol = reinterpret_cast<double *>(base + m.off_bcol), *scale = reinterpret_cast<double *>(base + m.off_scale);
  double *XT = reinterpret_cast<double *>(base + m.off_XT);
  const double inv_s2 = 1.0 / (sigma * sigma), c1 = kExpScale * inv_s2;
  int rc;
  // bcol doubles as the scratch for the (unused) scaled row terms first: prep writes bias then bias_col
  if (xyz_train) {
    rc = launch_prep_rows(nat, ntr, xyz_train, xeq, center, m.XS, c1, Xc, scale, bcol, st);
    if (rc) return rc;
  } else {
    big_center_rows_kernel<<<grid_for(ntr * m.XS), 256, 0, st>>>(m.D, ntr, m.XS, X_train, center, Xc);
    KREG_LAUNCH_CHECK();
    rc = launch_prep_bias(nat, ntr, m.XS, c1, Xc, scale, bcol, st);
    if (rc) return rc;
  }
  rc = launch_chunk_rows(ntr, m.XS, kGemmKC, m.nchD, kCW, Xc, Xch, st);
  if (rc) return rc;
  if (has_v) {
    double *Vs = reinterpret_cast<double *>(base + m.off_Vs), *Vsch = reinterpret_cast<double *>(base + m.off_Vsch);
    big_prep_v_kernel<<<grid_for(ntr), 256, 0, st>>>(m.D, ntr, m.XS, Xc, V, alpha_val, inv_s2, Vs, scale);
    KREG_LAUNCH_CHECK();
    rc = launch_chunk_rows(ntr, m.XS, kGemmKC, m.nchD, kCW, Vs, Vsch, st);
    if (rc) return rc;
  } else {
    big_copy_scale_kernel<<<grid_for(ntr), 256, 0, st>>>(ntr, alpha_val, scale);
    KREG_LAUNCH_CHECK();
  }
  const int64_t tot = (int64_t)m.nchW * m.nbT * kCW;
  big_transpose_kernel<<<grid_for(tot), 256, 0, st>>>(m.D, ntr, m.nchW, Xc, m.XS, 1.0, XT);
  KREG_LAUNCH_CHECK();
  if (has_v) {
    big_transpose_kernel<<<grid_for(tot), 256, 0, st>>>(m.D, ntr, m.nchW, V, m.D, 0.0, XT + tot);
    KREG_LAUNCH_CHECK();
  }
  return KREG_OK;
}

size_t big_predict_workspace_bytes(int nat, int64_t ntr, bool has_v, int64_t npoints, int sms) {
  return big_ws(big_model(nat, ntr, has_v), ntr, has_v, npoints, sms).total;
}

size_t big_predict_radial_workspace_bytes(int nat, int64_t ntr, int64_t npoints, int sms) {
  return big_ws(big_model(nat, ntr, false), ntr, false, npoints, sms, true).total;
}

// Prediction with MLatomF's radial kernel functions (Gaussian / exponential / Matern; Kfunk A_KRR_kernel.f90:806-851, first
// derivatives dKdXid :1178-1233 inside dKdMiat_unpermuted :1407-1447) for a values-trained model, any molecule size, on
// the same three-kernel skeleton:
//   GEMM1   S = X'te . X'tr^T - n_j/2, epilogue (EPI_WR): r_ij from S, then w_ij = alpha_j phi(r_ij) and e_ij = alpha_j k(r_ij)
//   GEMM2   A = W . [X'tr | 1]   (gradient weights; skipped for energies only);    e_i = row sums of E (chunk_rowsum_kernel)
//   epilogue per geometry: E = prior + e, g = A[:D] - X'_i A[D], dE/dM = J_i^T g
// The packed model is the values-only layout of big_pack_model (sigma does not enter it).
int big_predict_radial(int kind, int nn, int nat, int64_t ntr, const void *packed, const double *center, const double *xeq,
                       double sigma, double prior, int64_t npoints, const double *xyz, double *E, double *G,
                       void *workspace, size_t workspace_bytes, int sms, cudaStream_t st) {
  const BigModel m = big_model(nat, ntr, false);
  const BigWs w = big_ws(m, ntr, false, npoints, sms, true);
  if (!workspace || workspace_bytes < w.total) return KREG_E_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0 || (reinterpret_cast<uintptr_t>(packed) & 255) != 0)
    return KREG_E_BADARG;
  const char *mb = static_cast<const char *>(packed);
  const double *Xch = reinterpret_cast<const double *>(mb + m.off_Xch);
  const double *bcol = reinterpret_cast<const double *>(mb + m.off_bcol);
  const double *scale = reinterpret_cast<const double *>(mb + m.off_scale);
  const double *XT = reinterpret_cast<const double *>(mb + m.off_XT);
  char *wb = static_cast<char *>(workspace);
  double *XcA = reinterpret_cast<double *>(wb + w.off_Xc), *XAch = reinterpret_cast<double *>(wb + w.off_Xch);
  double *biasS = reinterpret_cast<double *>(wb + w.off_bias), *biasU = reinterpret_cast<double *>(wb + w.off_bias2);
  double *zeros = reinterpret_cast<double *>(wb + w.off_zero), *Wch = reinterpret_cast<double *>(wb + w.off_W);
  double *acc = reinterpret_cast<double *>(wb + w.off_acc), *accE = reinterpret_cast<double *>(wb + w.off_accE);
  KREG_CUDA_TRY(cudaMemsetAsync(zeros, 0, (size_t)(m.nbT + 66) * sizeof(double), st));
  const bool want_g = G != nullptr;

  for (int64_t n0 = 0; n0 < npoints; n0 += w.TN) {
    const int64_t tn = npoints - n0 < w.TN ? npoints - n0 : w.TN;
    const double *x = xyz + n0 * nat * 3;
    // the radial epilogue wants the UNSCALED row term -n_i/2 (biasU); the scaled one (biasS) is a by-product
    int rc = launch_prep_rows(nat, tn, x, xeq, center, m.XS, 1.0, XcA, biasS, biasU, st);
    if (rc) return rc;
    rc = launch_chunk_rows(tn, m.XS, kGemmKC, m.nchD, kCW, XcA, XAch, st);
    if (rc) return rc;
    ValuesArgs a = {};
    a.XSC = kCW;
    a.XS = m.XS;
    a.all_cols = 1;
    a.row_begin = 0;
    a.row_end = tn;
    a.na = tn;
    a.strideA = tn;
    a.wkc = kCW;
    a.biasA = biasU;  // every mode of the tile kernel stages the row terms (the plain-GEMM mode does not use them)
    double *Efch = Wch + (size_t)m.nchW * tn * kCW;
    KREG_CUDA_TRY(cudaMemsetAsync(Wch + (size_t)(m.nchW - 1) * tn * kCW, 0, (size_t)tn * kCW * sizeof(double), st));
    KREG_CUDA_TRY(cudaMemsetAsync(Efch + (size_t)(m.nchW - 1) * tn * kCW, 0, (size_t)tn * kCW * sizeof(double), st));
    {
      ValuesArgs g1 = a;
      g1.XAch = XAch;
      g1.XBch = Xch;
      g1.nb = ntr;
      g1.strideB = ntr;
      g1.nchunks = m.nchD;
      g1.biasB = bcol;
      g1.colscale = scale;
      g1.Wch = Wch;
      g1.Kfch = Efch;
      g1.kind = kind;
      g1.nn = nn;
      g1.inv_sigma = 1.0 / sigma;
      g1.inv_s2 = 1.0 / (sigma * sigma);
      if (kind == KREG_KERNEL_MATERN) set_matern_coefficients(g1, nn, sigma);
      rc = launch_gemm_epi(g1, EPI_WR, st);
      if (rc) return rc;
    }
    if (want_g) {
      ValuesArgs g2 = a;
      g2.XAch = Wch;
      g2.XBch = XT;
      g2.nb = m.nbT;
      g2.strideB = m.nbT;
      g2.nchunks = m.nchW;
      g2.biasB = zeros;
      g2.K = acc;
      g2.ldk = w.NPAD;
      g2.njt = 1;
      rc = launch_gemm_epi(g2, EPI_LIN, st);
      if (rc) return rc;
    }
    // energies: row sums of the chunk-major alpha_j k_ij (a GEMM with the ones row would waste 63/64 of its tile)
    rc = launch_chunk_rowsum(tn, m.nchW, kCW, Efch, accE, 2, st);
    if (rc) return rc;
    big_epilogue_kernel<<<(unsigned)((tn + 7) / 8), 256, 0, st>>>(nat, tn, x, xeq, XcA, m.XS, want_g ? acc : accE,
                                                                  want_g ? w.NPAD : 2, want_g ? m.D : 0, 1.0, prior,
                                                                  E ? E + n0 : nullptr, G ? G + n0 * nat * 3 : nullptr,
                                                                  accE, 2);
    KREG_LAUNCH_CHECK();
  }
  return KREG_OK;
}

int big_predict(int nat, int64_t ntr, const void *packed, const double *center, const double *xeq, double sigma,
                double prior, bool has_v, int64_t npoints, const double *xyz, double *E, double *G, void *workspace,
                size_t workspace_bytes, int sms, cudaStream_t st) {
  const BigModel m = big_model(nat, ntr, has_v);
  const BigWs w = big_ws(m, ntr, has_v, npoints, sms);
  if (!workspace || workspace_bytes < w.total) return KREG_E_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0 || (reinterpret_cast<uintptr_t>(packed) & 255) != 0)
    return KREG_E_BADARG;  // bulk copies need 16-byte aligned sources; every sub-array sits on a 256-byte boundary
  const char *mb = static_cast<const char *>(packed);
  const double *Xch = reinterpret_cast<const double *>(mb + m.off_Xch);
  const double *bcol = reinterpret_cast<const double *>(mb + m.off_bcol);
  const double *scale = reinterpret_cast<const double *>(mb + m.off_scale);
  const double *Vsch = reinterpret_cast<const double *>(mb + m.off_Vsch);
  const double *XT = reinterpret_cast<const double *>(mb + m.off_XT);
  char *wb = static_cast<char *>(workspace);
  double *XcA = reinterpret_cast<double *>(wb + w.off_Xc), *XAch = reinterpret_cast<double *>(wb + w.off_Xch);
  double *biasA = reinterpret_cast<double *>(wb + w.off_bias), *zeros = reinterpret_cast<double *>(wb + w.off_zero);
  double *C2 = reinterpret_cast<double *>(wb + w.off_C2), *Wch = reinterpret_cast<double *>(wb + w.off_W);
  double *acc = reinterpret_cast<double *>(wb + w.off_acc);
  const double inv_s2 = 1.0 / (sigma * sigma), c1 = kExpScale * inv_s2;
  KREG_CUDA_TRY(cudaMemsetAsync(zeros, 0, (size_t)(m.nbT + 66) * sizeof(double), st));

  for (int64_t n0 = 0; n0 < npoints; n0 += w.TN) {
    const int64_t tn = npoints - n0 < w.TN ? npoints - n0 : w.TN;
    const double *x = xyz + n0 * nat * 3;
    int rc = launch_prep_rows(nat, tn, x, xeq, center, m.XS, c1, XcA, biasA, nullptr, st);
    if (rc) return rc;
    rc = launch_chunk_rows(tn, m.XS, kGemmKC, m.nchD, kCW, XcA, XAch, st);
    if (rc) return rc;
    ValuesArgs a = {};
    a.XSC = kCW;
    a.XS = m.XS;
    a.c1 = c1;
    a.all_cols = 1;
    a.row_begin = 0;
    a.row_end = tn;
    a.na = tn;
    a.strideA = tn;
    a.biasA = biasA;
    a.wkc = kCW;
    if (has_v) {
      // C2[i][j] = alpha'_j + X'_i . v_j / s^2
      ValuesArgs l = a;
      l.XAch = XAch;
      l.XBch = Vsch;
      l.nb = ntr;
      l.strideB = ntr;
      l.nchunks = m.nchD;
      l.biasB = scale;
      l.K = C2;
      l.ldk = w.ldc;
      rc = launch_gemm_epi(l, EPI_LIN, st);
      if (rc) return rc;
    }
    // the padding columns of the last W chunk (j >= Ntrain) must be exact zeros for GEMM2
    double *Kfch = Wch + (size_t)m.nchW * tn * kCW;
    KREG_CUDA_TRY(cudaMemsetAsync(Wch + (size_t)(m.nchW - 1) * tn * kCW, 0, (size_t)tn * kCW * sizeof(double), st));
    if (has_v)
      KREG_CUDA_TRY(cudaMemsetAsync(Kfch + (size_t)(m.nchW - 1) * tn * kCW, 0, (size_t)tn * kCW * sizeof(double), st));
    {
      ValuesArgs g1 = a;
      g1.XAch = XAch;
      g1.XBch = Xch;
      g1.nb = ntr;
      g1.strideB = ntr;
      g1.nchunks = m.nchD;
      g1.biasB = bcol;
      g1.colscale = scale;
      g1.Cin = C2;
      g1.ldc = w.ldc;
      g1.Wch = Wch;
      g1.Kfch = Kfch;
      rc = launch_gemm_epi(g1, has_v ? EPI_WV : EPI_W, st);
      if (rc) return rc;
    }
    const bool want_g = G != nullptr;
    {
      // A = W . [X'tr | 1] (+ K . [V | 0]); energies only: just the ones column (one column tile)
      ValuesArgs g2 = a;
      g2.XAch = Wch;
      g2.XBch = want_g ? XT : XT + (size_t)m.D * kCW;
      g2.nb = want_g ? m.nbT : 1;
      g2.strideB = m.nbT;
      g2.nchunks = m.nchW * ((has_v && want_g) ? 2 : 1);
      g2.biasB = zeros;
      g2.K = acc;
      g2.ldk = w.NPAD;
      g2.njt = 1;
      rc = launch_gemm_epi(g2, EPI_LIN, st);
      if (rc) return rc;
    }
    big_epilogue_kernel<<<(unsigned)((tn + 7) / 8), 256, 0, st>>>(nat, tn, x, xeq, XcA, m.XS, acc, w.NPAD,
                                                                  want_g ? m.D : 0, inv_s2, prior, E ? E + n0 : nullptr,
                                                                  G ? G + n0 * nat * 3 : nullptr);
    KREG_LAUNCH_CHECK();
  }
  return KREG_OK;
}

}  // namespace kreg
