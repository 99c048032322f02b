// kreg_gemm.h -- argument block and entry points of the DMMA tile kernel (values_ring_kernel, kreg_kbuild.cu) shared by
// the kernel-matrix assembly and the large-molecule prediction path (kreg_predict_big.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace kreg {

struct ValuesArgs {
  const double *XA;  // [na][XS] centred descriptors, zero padded
  const double *XB;  // [nb][XS]
  const double *XAch;  // multi-chunk descriptors only: [nchunks][na][XSC] chunk-major copy (contiguous chunk tiles)
  const double *XBch;  // [nchunks][nb][XSC]
  int XSC;             // row stride of the chunk-major copies = the kernel's conflict-free chunk stride
  const double *biasA;  // row term, already scaled: -c1 n_i / 2 (EPI_LIN: unused)
  const double *biasB;  // column term in units of S: -n_j / 2 (EPI_LIN: any additive column term); it is the initial value
                        // of the GEMM accumulators, so the epilogue needs no per-element add
  int64_t na, nb;
  int64_t strideA, strideB;  // rows per chunk of the chunk-major copies (na / nb unless an operand is a row sub-range)
  int XS;       // row stride (doubles), >= 4*KC*nchunks, = 4 mod 8
  int nchunks;  // K-chunks of 4*KC descriptor elements
  double c1;
  double lambda;  // added where i == j when `symmetric`
  int symmetric;  // 1: XA == XB (K build): exact 1+lambda on the diagonal, tiles above the diagonal skipped unless `all_cols`
  int all_cols;   // 1: compute every column tile directly (row-slab FULL mode / rectangular block)
  int mirror;     // 1: also write K[j][i] for tiles strictly below the diagonal blocks (whole-matrix FULL mode)
  int64_t row_begin, row_end;  // rows of A to produce
  double *K;
  int64_t ldk;
  int njt;  // column tiles walked by one CTA (set by the launcher)
  // EPI_W (prediction of large molecules, kreg_predict_big.cu): w_ij = k_ij * colscale[j] (value models) or
  // k_ij * Cin[i][j] (gradient models, which also keep k_ij), written chunk-major as the A operand of the second GEMM
  const double *colscale;
  const double *Cin;
  int64_t ldc;
  double *Wch, *Kfch;  // [nchunks_w][na][wkc]
  int wkc;             // columns per output chunk (the second GEMM's K chunk)
  // EPI_KR (MLatomF's other radial kernels, A_KRR_kernel.f90:806-851): k = f(|X_i - X_j| / sigma); biasA is then the
  // UNSCALED row term -n_i/2
  int kind;            // KREG_KERNEL_EXPONENTIAL / KREG_KERNEL_MATERN
  int nn;              // Matern order n (nu = n + 1/2)
  double inv_sigma;
  double mat_c[12];    // Matern polynomial coefficients in (2 r / sigma)^m, m = 0..nn
  // EPI_WR (prediction with the radial kernels, kreg_predict_big.cu): phi(r) = -k'(r) / r, the factor of (X_j - X_i) in
  // dk/dX_i (dKdXid, A_KRR_kernel.f90:1178-1233); Matern: exp(-r/sigma) sum_m mat_d[m] (2 r / sigma)^m, m = 0..nn-1
  double mat_d[12];
  double inv_s2;
};

constexpr int kGemmKC = 9;  // k-steps per chunk of the GEMM modes (chunk width 36 doubles, conflict-free stride 36)

// epilogue modes of values_ring_kernel
constexpr int EPI_K = 0;    // K-matrix / validation block: exp, diagonal, mirror
constexpr int EPI_W = 1;    // prediction GEMM1: exp * column scale -> chunk-major W
constexpr int EPI_WV = 2;   // prediction GEMM1, gradient models: exp * Cin -> chunk-major W and K
constexpr int EPI_LIN = 3;  // plain GEMM: C = A B^T + column term (row-major)
constexpr int EPI_KR = 4;   // radial kernels of the Euclidean distance: exponential, Matern
constexpr int EPI_WR = 5;   // prediction GEMM1 for the radial kernels: phi(r) * column scale -> chunk-major W, k(r) * column scale -> K

// kreg_kbuild.cu
void set_matern_coefficients(ValuesArgs &a, int nn, double sigma);
int launch_gemm_epi(const ValuesArgs &a, int epi, cudaStream_t st);
// centred, zero-padded descriptor rows Xc[n][XS] of the geometries xyz[n][nat][3]; bias[p] = -c1 |Xc_p|^2 / 2 and
// (optionally) bias_col[p] = -|Xc_p|^2 / 2
int launch_prep_rows(int nat, int64_t n, const double *xyz, const double *xeq, const double *center, int XS, double c1,
                     double *Xc, double *bias, double *bias_col, cudaStream_t st);
// the two bias arrays alone, for rows Xc that are already in place
int launch_prep_bias(int nat, int64_t n, int XS, double c1, const double *Xc, double *bias, double *bias_col,
                     cudaStream_t st);
// out[i * ldo] = sum of row i of a chunk-major array W[nch][n][wkc] (deterministic)
int launch_chunk_rowsum(int64_t n, int nch, int wkc, const double *W, double *out, int ldo, cudaStream_t st);
// chunk-major copy Xch[ch][p][XSC] = Xc[p][ch * 4 KC + (0 .. 4 KC)), zero padded
int launch_chunk_rows(int64_t n, int XS, int KC, int nchunks, int XSC, const double *Xc, double *Xch, cudaStream_t st);

}  // namespace kreg
