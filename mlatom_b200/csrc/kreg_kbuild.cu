// kreg_kbuild.cu -- kernel-matrix assembly: values block on FP64 tensor cores, gradient-augmented
// blocks as dense 3Nat x 3Nat tiles written at HBM rate.
//
// Replaces KREG.kreg.calc_kernel_matrix + the lambda shift of calcAlpha
// (mlatom/fortran/KREG.f90:293-346, :616-622; MLatomF_src/A_KRR_kernel.f90:130-188).
// Blocks (row-major K, M = Nv + Ng, Ng = 3 Nat Ntr, G3 = 3 Nat):
//   K[i][j]                         = k_ij                                   (values, KREG.f90:314-320)
//   K[Nv + p*G3 + q][i]             = (k_pi/s^2) (J_p^T (X_i - X_p))_q       (value-gradient, :326-331)
//   K[Nv + i*G3 + q][Nv + j*G3 + r] = (k_ij/s^2) [J_i^T J_j - u_i u_j^T/s^2]_{qr},
//                                     u_i = J_i^T D, u_j = J_j^T D, D = X_j - X_i (grad-grad, :337-343)
// J[d,(a,t)] = X_d (x_c - x_a)_t / R_ac^2 for d = (a,c) is the descriptor Jacobian (KREG.f90:121,146); the
// bracket is the O(D) expansion of the reference's (Nat-1)^2-term double loop.
#include "kreg_common.cuh"
#include "kreg_internal.h"
#include "kreg_gemm.h"

namespace kreg {

// =================================================================================================
// Values block / rectangular value-value block:  S = XA . XB^T on DMMA.8x8x4, then exp.
// =================================================================================================
constexpr int kTileI = 128;  // rows per CTA tile
constexpr int kTileJ = 64;   // columns per CTA tile



// The exponent c1 (X'_i.X'_j - n_i/2 - n_j/2) = -c1 |X_i - X_j|^2 / 2 is <= 0 up to rounding (a few ulp of c1 n);
// a positive rounding residue only makes k = 1 + O(1e-16), far inside the 1e-13 parity tolerance, so no clamp is
// needed (one FP64-pipe instruction per element saved; -DKREG_VALUES_CLAMP=1 restores it).
#ifndef KREG_VALUES_CLAMP
#define KREG_VALUES_CLAMP 0
#endif
#if KREG_VALUES_CLAMP
#define KREG_EXP_ARG(x) fmin((x), 0.0)
#else
#define KREG_EXP_ARG(x) (x)
#endif

// Epilogue of one 128 x 64 tile, shared by both values kernels: k = exp(c1 s + b_i) where the accumulator s already
// holds X'_i.X'_j - n_j/2 (it was initialised with the column term), diagonal exactly 1 + lambda, 16-byte stores (64 B
// runs per row).
__device__ __forceinline__ void values_tile_epilogue(const ValuesArgs &a, double (&S)[4][4][2], int64_t i0, int64_t j0,
                                                     const double *bA, const double *tab, int wi,
                                                     int wj, int g, int q, bool lower_only) {
  const bool diag_tile = a.symmetric && j0 + kTileJ - 1 >= i0;  // touches the diagonal block: computed directly
  // Row-slab FULL assembly of the values block computes K[i][j] and K[j][i] in different tiles (possibly on different
  // GPUs): there the exponent is formed symmetrically, c1 s + (b_i + b_j) with s = X'_i.X'_j alone (the accumulators were
  // started from zero), so that the two are bit-identical like the reference's mirrored copy (KREG.f90:317-318).
  const bool symx = a.symmetric && a.all_cols;
  // interior tile: every row wanted, every column valid, nothing on or above the diagonal, 16-byte aligned rows
  const bool interior = !diag_tile && !symx && i0 >= a.row_begin && i0 + kTileI <= a.row_end && i0 + kTileI <= a.na &&
                        j0 + kTileJ <= a.nb && (a.ldk & 1) == 0 && (reinterpret_cast<uintptr_t>(a.K) & 15) == 0;
  if (interior) {
    double *row = a.K + (i0 + wi * 32 + g) * a.ldk + j0 + wj * 32 + 2 * q;
#pragma unroll
    for (int mt = 0; mt < 4; mt++) {
      const double bi = bA[wi * 32 + mt * 8 + g];
#pragma unroll
      for (int nt = 0; nt < 4; nt++) {
        const int cj = wj * 32 + nt * 8 + 2 * q;
        const double k0 = exp_from_scaled(KREG_EXP_ARG(fma(S[mt][nt][0], a.c1, bi)), tab);
        const double k1 = exp_from_scaled(KREG_EXP_ARG(fma(S[mt][nt][1], a.c1, bi)), tab);
        KREG_CHECK_STORE(row + nt * 8 + 1, a.K + a.row_begin * a.ldk, (a.row_end - a.row_begin) * a.ldk);
        *reinterpret_cast<double2 *>(row + nt * 8) = make_double2(k0, k1);
        if (a.mirror) {
          double *t = a.K + (j0 + cj) * a.ldk + i0 + wi * 32 + mt * 8 + g;
          KREG_CHECK_STORE(t + a.ldk, a.K, a.na * a.ldk);
          t[0] = k0;
          t[a.ldk] = k1;
        }
      }
      row += 8 * a.ldk;
    }
    return;
  }
  // whole-matrix modes compute the tiles on and below the diagonal only: inside a diagonal tile the entries above the
  // diagonal are taken from their mirror images (written below), never computed a second time
  const bool tri = a.symmetric && !a.all_cols;
#pragma unroll
  for (int mt = 0; mt < 4; mt++) {
    const int ri = wi * 32 + mt * 8 + g;
    const int64_t i = i0 + ri;
    const bool row_ok = i < a.na && i >= a.row_begin && i < a.row_end;
    const double bi = bA[ri];
#pragma unroll
    for (int nt = 0; nt < 4; nt++) {
      const int cj = wj * 32 + nt * 8 + 2 * q;
      const int64_t j = j0 + cj;
      double e0 = bi, e1 = bi;
      if (symx) {
        e0 = bi + (j < a.nb ? __ldg(a.biasA + j) : 0.0);
        e1 = bi + (j + 1 < a.nb ? __ldg(a.biasA + j + 1) : 0.0);
      }
      double k0 = exp_from_scaled(KREG_EXP_ARG(fma(S[mt][nt][0], a.c1, e0)), tab);
      double k1 = exp_from_scaled(KREG_EXP_ARG(fma(S[mt][nt][1], a.c1, e1)), tab);
      if (a.symmetric) {
        if (i == j) k0 = 1.0 + a.lambda;
        if (i == j + 1) k1 = 1.0 + a.lambda;
      }
      if (row_ok) {
        double *dst = a.K + i * a.ldk + j;
        const bool w0 = j < a.nb && (!tri || j <= i);
        const bool w1 = j + 1 < a.nb && (!tri || j + 1 <= i);
        if (w0 || w1) KREG_CHECK_STORE(dst + (w1 ? 1 : 0), a.K + a.row_begin * a.ldk, (a.row_end - a.row_begin) * a.ldk);
        if (w0 && w1 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
          *reinterpret_cast<double2 *>(dst) = make_double2(k0, k1);
        } else {
          if (w0) dst[0] = k0;
          if (w1) dst[1] = k1;
        }
      }
      if (a.mirror && i < a.na) {
        if (j < a.nb && j < i) a.K[j * a.ldk + i] = k0;
        if (j + 1 < a.nb && j + 1 < i) a.K[(j + 1) * a.ldk + i] = k1;
      }
    }
  }
}

// Prediction GEMM1 epilogue (large molecules): w_ij = k_ij * colscale[j] (EPI_W) or k_ij * Cin[i][j] (EPI_WV, which
// also keeps k_ij), stored chunk-major -- [j / wkc][i][j % wkc] -- which is the layout the second GEMM streams its A
// operand in.  j is even and wkc is even, so a (j, j+1) pair never straddles a chunk and every store is 16 bytes.
template <int EPI>
__device__ __forceinline__ void w_tile_epilogue(const ValuesArgs &a, double (&S)[4][4][2], int64_t i0, int64_t j0,
                                                const double *bA, const double *tab, int wi, int wj, int g, int q) {
#pragma unroll
  for (int mt = 0; mt < 4; mt++) {
    const int ri = wi * 32 + mt * 8 + g;
    const int64_t i = i0 + ri;
    const bool row_ok = i < a.na;
    const double bi = bA[ri];
#pragma unroll
    for (int nt = 0; nt < 4; nt++) {
      const int64_t j = j0 + wj * 32 + nt * 8 + 2 * q;
      if (!row_ok || j >= a.nb) continue;
      double k0 = exp_from_scaled(KREG_EXP_ARG(fma(S[mt][nt][0], a.c1, bi)), tab);
      double k1 = exp_from_scaled(KREG_EXP_ARG(fma(S[mt][nt][1], a.c1, bi)), tab);
      const bool has1 = j + 1 < a.nb;
      if (!has1) k1 = 0.0;  // padding column of the last chunk: must be an exact zero for the second GEMM
      double w0, w1;
      if (EPI == EPI_WV) {
        const double *c = a.Cin + i * a.ldc + j;
        w0 = k0 * c[0];
        w1 = has1 ? k1 * c[1] : 0.0;
      } else {
        w0 = k0 * __ldg(a.colscale + j);
        w1 = has1 ? k1 * __ldg(a.colscale + j + 1) : 0.0;
      }
      const int64_t ch = j / a.wkc;
      const int64_t off = (ch * a.na + i) * a.wkc + (j - ch * a.wkc);
      KREG_CHECK(off >= 0 && off + 1 < (int64_t)((a.nb + a.wkc - 1) / a.wkc) * a.na * a.wkc && (off & 1) == 0, 3);
      *reinterpret_cast<double2 *>(a.Wch + off) = make_double2(w0, w1);
      if (EPI == EPI_WV) *reinterpret_cast<double2 *>(a.Kfch + off) = make_double2(k0, k1);
    }
  }
}

// Radial kernels of the Euclidean descriptor distance (MLatomF Kfunk, A_KRR_kernel.f90:806-851, with distance()
// :2040-2074): r^2 = n_i + n_j - 2 X'_i.X'_j = -2 (s + b_i) with the accumulator s = X'_i.X'_j - n_j/2 and the unscaled
// row term b_i = -n_i/2;  exponential k = exp(-r/sigma);  Matern k = exp(-r/sigma) sum_m c_m (2r/sigma)^m.
// The square root amplifies the cancellation error of r^2 (~1e-16 n) to ~1e-16 n / r in r: 1e-13 parity holds for
// r >~ 1e-3, i.e. for everything but near-duplicate geometries (the diagonal itself is set exactly).
__device__ __forceinline__ double radial_kernel_value(const ValuesArgs &a, double s_plus_bi) {
  const double r2 = fmax(-2.0 * s_plus_bi, 0.0);
  const double x = sqrt(r2) * a.inv_sigma;
  double p = 1.0;
  if (a.kind == KREG_KERNEL_MATERN) {
    const double y = 2.0 * x;
    p = a.mat_c[a.nn];
    for (int m = a.nn - 1; m >= 0; m--) p = fma(p, y, a.mat_c[m]);
  }
  return exp(-x) * p;
}

__device__ __forceinline__ void radial_tile_epilogue(const ValuesArgs &a, double (&S)[4][4][2], int64_t i0, int64_t j0,
                                                     const double *bA, int wi, int wj, int g, int q) {
#pragma unroll
  for (int mt = 0; mt < 4; mt++) {
    const int ri = wi * 32 + mt * 8 + g;
    const int64_t i = i0 + ri;
    if (i >= a.na) continue;
    const double bi = bA[ri];
#pragma unroll
    for (int nt = 0; nt < 4; nt++) {
      const int64_t j = j0 + wj * 32 + nt * 8 + 2 * q;
      double k0 = radial_kernel_value(a, S[mt][nt][0] + bi), k1 = radial_kernel_value(a, S[mt][nt][1] + bi);
      if (a.symmetric) {
        if (i == j) k0 = 1.0 + a.lambda;
        if (i == j + 1) k1 = 1.0 + a.lambda;
      }
      double *dst = a.K + i * a.ldk + j;
      if (j < a.nb) {
        KREG_CHECK_STORE(dst, a.K, a.na * a.ldk);
        dst[0] = k0;
      }
      if (j + 1 < a.nb) dst[1] = k1;
    }
  }
}

// Prediction GEMM1 epilogue for the radial kernels (kreg_predict_big.cu, radial models): from r^2 = -2 (s + b_i) both
//   w_ij = alpha_j phi(r_ij),  phi = -k'(r) / r   (the factor of X_j - X_i in dk/dX_i: dKdXid, A_KRR_kernel.f90:1178-1233)
//   e_ij = alpha_j k(r_ij)                         (Yest_KRR, A_KRR_kernel.f90:225-245)
// go out chunk-major, the A operands of the two second GEMMs.  kind: Gaussian (phi = k / s^2; exercises this path against
// the fused kernels), exponential (phi = k / (sigma r), 0 at r = 0 like the reference's central difference of exp(-|x|)),
// Matern of order nn >= 1 (the reference's analytic form, :1205-1219: a polynomial in r, regular at 0).
// exp(-x) through the table exponential of the Gaussian kernels (11 FP64-pipe instructions against ~30 for libm's exp).
template <int NN>
__device__ __forceinline__ void matern_pair_values(const ValuesArgs &a, double x, double ex, double &k, double &phi) {
  const double y = 2.0 * x;
  double p = a.mat_c[NN], d = a.mat_d[NN - 1];
#pragma unroll
  for (int m = NN - 1; m >= 0; m--) p = fma(p, y, a.mat_c[m]);
#pragma unroll
  for (int m = NN - 2; m >= 0; m--) d = fma(d, y, a.mat_d[m]);
  k = ex * p;
  phi = ex * d;
}

__device__ __forceinline__ void radial_pair_values(const ValuesArgs &a, double s_plus_bi, const double *tab, double &k,
                                                   double &phi) {
  const double r2 = fmax(-2.0 * s_plus_bi, 0.0);
  if (a.kind == KREG_KERNEL_GAUSSIAN) {
    k = exp_from_scaled(-0.5 * kExpScale * a.inv_s2 * r2, tab);
    phi = k * a.inv_s2;
    return;
  }
  const double r = sqrt(r2), x = r * a.inv_sigma, ex = exp_from_scaled(-kExpScale * x, tab);
  if (a.kind == KREG_KERNEL_MATERN && a.nn >= 1) {
    switch (a.nn) {  // the orders in use get straight-line polynomials; the rest a loop
      case 1:
        return matern_pair_values<1>(a, x, ex, k, phi);
      case 2:
        return matern_pair_values<2>(a, x, ex, k, phi);
      case 3:
        return matern_pair_values<3>(a, x, ex, k, phi);
      default:
        break;
    }
    const double y = 2.0 * x;
    double p = a.mat_c[a.nn], d = a.mat_d[a.nn - 1];
    for (int m = a.nn - 1; m >= 0; m--) p = fma(p, y, a.mat_c[m]);
    for (int m = a.nn - 2; m >= 0; m--) d = fma(d, y, a.mat_d[m]);
    k = ex * p;
    phi = ex * d;
    return;
  }
  k = ex;
  phi = r > 0.0 ? ex * a.inv_sigma / r : 0.0;
}

__device__ __forceinline__ void wr_tile_epilogue(const ValuesArgs &a, double (&S)[4][4][2], int64_t i0, int64_t j0,
                                                 const double *bA, const double *tab, int wi, int wj, int g, int q) {
  // the chunk-major position of a column does not depend on the row: four integer divisions per tile, not one per element
  int64_t coloff[4];
  double c0[4], c1[4];
  bool ok0[4], ok1[4];
#pragma unroll
  for (int nt = 0; nt < 4; nt++) {
    const int64_t j = j0 + wj * 32 + nt * 8 + 2 * q;
    const int64_t ch = j / a.wkc;
    coloff[nt] = ch * a.na * a.wkc + (j - ch * a.wkc);
    ok0[nt] = j < a.nb;
    ok1[nt] = j + 1 < a.nb;
    c0[nt] = ok0[nt] ? __ldg(a.colscale + j) : 0.0;
    c1[nt] = ok1[nt] ? __ldg(a.colscale + j + 1) : 0.0;  // padding column of the last chunk: exact zeros for the second GEMM
  }
#pragma unroll
  for (int mt = 0; mt < 4; mt++) {
    const int ri = wi * 32 + mt * 8 + g;
    const int64_t i = i0 + ri;
    if (i >= a.na) continue;
    const double bi = bA[ri];
#pragma unroll
    for (int nt = 0; nt < 4; nt++) {
      if (!ok0[nt]) continue;
      double k0, k1, p0, p1;
      radial_pair_values(a, S[mt][nt][0] + bi, tab, k0, p0);
      radial_pair_values(a, S[mt][nt][1] + bi, tab, k1, p1);
      const int64_t off = coloff[nt] + i * a.wkc;
      KREG_CHECK(off >= 0 && off + 1 < (int64_t)((a.nb + a.wkc - 1) / a.wkc) * a.na * a.wkc && (off & 1) == 0, 3);
      *reinterpret_cast<double2 *>(a.Wch + off) = make_double2(p0 * c0[nt], ok1[nt] ? p1 * c1[nt] : 0.0);
      *reinterpret_cast<double2 *>(a.Kfch + off) = make_double2(k0 * c0[nt], ok1[nt] ? k1 * c1[nt] : 0.0);
    }
  }
}

// Plain GEMM epilogue: C[i][j] = accumulator (A_i . B_j + column term), row-major.
__device__ __forceinline__ void lin_tile_epilogue(const ValuesArgs &a, double (&S)[4][4][2], int64_t i0, int64_t j0,
                                                  int wi, int wj, int g, int q) {
  const bool vec = (a.ldk & 1) == 0 && (reinterpret_cast<uintptr_t>(a.K) & 15) == 0;
#pragma unroll
  for (int mt = 0; mt < 4; mt++) {
    const int64_t i = i0 + wi * 32 + mt * 8 + g;
    if (i >= a.na) continue;
#pragma unroll
    for (int nt = 0; nt < 4; nt++) {
      const int64_t j = j0 + wj * 32 + nt * 8 + 2 * q;
      double *dst = a.K + i * a.ldk + j;
      if (j < a.nb) KREG_CHECK_STORE(dst + (j + 1 < a.nb ? 1 : 0), a.K, a.na * a.ldk);
      if (vec && j + 1 < a.nb) {
        *reinterpret_cast<double2 *>(dst) = make_double2(S[mt][nt][0], S[mt][nt][1]);
      } else {
        if (j < a.nb) dst[0] = S[mt][nt][0];
        if (j + 1 < a.nb) dst[1] = S[mt][nt][1];
      }
    }
  }
}

// The values block: CTA tile 128 x 64, 8 warps as 4 (rows) x 2 (cols), warp tile 32 x 32 = 4 x 4 DMMA tiles, <= 128
// registers so two CTAs share an SM.  A CTA walks `njt` consecutive column tiles; operands arrive through a ring of TMA
// bulk copies with full/empty mbarriers, like the prediction kernel's model stream, so there is no CTA-wide barrier in
// the steady state: a warp that has finished a tile's exponentials moves on to the next tile's tensor-core work while
// the other warps are still in their epilogues.
//   MULTI = false (KS <= 12, D <= 48): the A tile is loaded once; a stage = one column tile + its biases (the workspace
//           row stride is the conflict-free one, so a tile of rows is one contiguous block).
//   MULTI = true: the descriptor is walked in chunks of KC k-steps; a stage = the A chunk and the B chunk of one
//           (tile, chunk) step, both contiguous in the chunk-major copy of the workspace; column biases travel with chunk 0
//           into a slot selected by tile parity.
template <int KC, bool MULTI, int EPI = EPI_K>
__global__ void __launch_bounds__(256, 2) values_ring_kernel(const ValuesArgs a) {
  constexpr int XSP = KC * 4 + ((KC * 4) % 8 == 4 ? 0 : 4);
  constexpr int A_DBL = kTileI * XSP, B_DBL = kTileJ * XSP;
  constexpr int NST = MULTI ? 2 : (KC <= 9 ? 3 : 2);
  constexpr int STAGE = MULTI ? A_DBL + B_DBL : B_DBL;
  constexpr int NBIAS = MULTI ? 2 : NST;
  extern __shared__ __align__(128) double sm[];
  double *As = sm;                                  // !MULTI: [128][XSP]
  double *ring = As + (MULTI ? 0 : A_DBL);          // [NST][STAGE]  (MULTI: A chunk, then B chunk)
  double *bBs = ring + NST * STAGE;                 // [NBIAS][64]
  double *bA = bBs + NBIAS * kTileJ;                // [128]
  double *tab = bA + kTileI;                        // [64]
  uint64_t *full = reinterpret_cast<uint64_t *>(tab + 64);  // [NST], then empty[NST], then fullA
  uint64_t *empty = full + NST;
  uint64_t *fullA = empty + NST;

  const int64_t ti = (int64_t)blockIdx.y + a.row_begin / kTileI;
  const int64_t i0 = ti * kTileI;
  if (i0 >= a.na) return;
  int64_t tj_lo = (int64_t)blockIdx.x * a.njt;
  int64_t tj_hi = tj_lo + a.njt;
  const int64_t ntj = (a.nb + kTileJ - 1) / kTileJ;
  if (tj_hi > ntj) tj_hi = ntj;
  if (a.symmetric && !a.all_cols) {
    const int64_t last = (i0 + kTileI - 1) / kTileJ + 1;  // tiles entirely above the diagonal are not needed
    if (tj_hi > last) tj_hi = last;
  }
  if (tj_lo >= tj_hi) return;
  const int nch = MULTI ? a.nchunks : 1;
  const int nsteps = (int)(tj_hi - tj_lo) * nch;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int wi = warp >> 1, wj = warp & 1;

  if (tid == 0) {
    for (int s = 0; s < NST; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 8);
    }
    mbar_init(fullA, 1);
    mbar_fence_init();
  }
  fill_exp_table(tab, tid, 256);
  __syncthreads();

  // rows past the end of the matrix are not copied: their shared-memory rows hold whatever was there, their results
  // are never stored.  Byte counts stay multiples of 16 (bias ranges are rounded up to an even count; the workspace
  // arrays are padded to 256 bytes).
  const unsigned rowsA = (unsigned)(a.na - i0 < kTileI ? a.na - i0 : kTileI);
  auto issue = [&](int step) {  // thread 0 only
    const int s = step % NST, tile = step / nch, ch = step - tile * nch;
    const int64_t j0 = (tj_lo + tile) * kTileJ;
    const unsigned rows = (unsigned)(a.nb - j0 < kTileJ ? a.nb - j0 : kTileJ);
    const unsigned bx = rows * XSP * 8, bb = ((rows + 1) & ~1u) * 8, ba = rowsA * XSP * 8;
    double *st = ring + s * STAGE;
    if (MULTI) {
      mbar_arrive_expect_tx(&full[s], ba + bx + (ch == 0 ? bb : 0u));
      bulk_g2s(st, a.XAch + ((int64_t)ch * a.strideA + i0) * XSP, ba, &full[s]);
      bulk_g2s(st + A_DBL, a.XBch + ((int64_t)ch * a.strideB + j0) * XSP, bx, &full[s]);
      if (ch == 0) bulk_g2s(bBs + (tile & 1) * kTileJ, a.biasB + j0, bb, &full[s]);
    } else {
      mbar_arrive_expect_tx(&full[s], bx + bb);
      bulk_g2s(st, a.XB + j0 * a.XS, bx, &full[s]);
      bulk_g2s(bBs + s * kTileJ, a.biasB + j0, bb, &full[s]);
    }
  };
  if (tid == 0) {
    const unsigned bb = ((rowsA + 1) & ~1u) * 8, bx = MULTI ? 0u : rowsA * XSP * 8;
    mbar_arrive_expect_tx(fullA, bx + bb);
    if (!MULTI) bulk_g2s(As, a.XA + i0 * a.XS, bx, fullA);
    bulk_g2s(bA, a.biasA + i0, bb, fullA);
    for (int s = 0; s < NST && s < nsteps; s++) issue(s);
  }
  __syncwarp();

  double S[4][4][2];
  const bool lower_only = a.symmetric && !a.all_cols && !a.mirror;
  mbar_wait(fullA, 0);

  // accumulators start from the column term of their tile (-n_j/2; an arbitrary additive term for EPI_LIN), which
  // arrived with the tile's first stage: the epilogue then has no per-element bias add
  const bool zero_init = EPI == EPI_K && a.symmetric && a.all_cols;  // see values_tile_epilogue: symmetric exponent
  auto init_acc = [&](const double *bcol) {
#pragma unroll
    for (int nt = 0; nt < 4; nt++) {
      double2 b = *reinterpret_cast<const double2 *>(bcol + wj * 32 + nt * 8 + 2 * q);
      if (zero_init) b = make_double2(0.0, 0.0);
#pragma unroll
      for (int mt = 0; mt < 4; mt++) {
        S[mt][nt][0] = b.x;
        S[mt][nt][1] = b.y;
      }
    }
  };
  auto epilogue = [&](int64_t j0) {
    if (EPI == EPI_K)
      values_tile_epilogue(a, S, i0, j0, bA, tab, wi, wj, g, q, lower_only);
    else if (EPI == EPI_LIN)
      lin_tile_epilogue(a, S, i0, j0, wi, wj, g, q);
    else if (EPI == EPI_KR)
      radial_tile_epilogue(a, S, i0, j0, bA, wi, wj, g, q);
    else if (EPI == EPI_WR)
      wr_tile_epilogue(a, S, i0, j0, bA, tab, wi, wj, g, q);
    else
      w_tile_epilogue<EPI>(a, S, i0, j0, bA, tab, wi, wj, g, q);
  };

  for (int step = 0; step < nsteps; step++) {
    // producer duty: refill the slot every warp left in the previous iteration.  The branch is warp-uniform with the
    // single issuing lane inside it and an explicit reconvergence after it: with a bare `if (tid == 0)` the other
    // lanes of warp 0 ran ahead into the tile code while lane 0 was still issuing, and the uniform registers the two
    // paths share were clobbered -- warp 0's 32 x 32 sub-tile came out wrong in every tile after the first refill
    // (found with tools/debug_kvals.py once the accumulators started from the column term instead of zero).
    if (warp == 0 && step >= 1 && step - 1 + NST < nsteps) {
      if (lane == 0) {
        mbar_wait(&empty[(step - 1) % NST], ((step - 1) / NST) & 1);
        issue(step - 1 + NST);
      }
      __syncwarp();
    }
    const int s = step % NST;
    mbar_wait(&full[s], (step / NST) & 1);
    const double *At = MULTI ? ring + s * STAGE : As;
    const double *Bt = ring + s * STAGE + (MULTI ? A_DBL : 0);
    if (MULTI) {
      const int tile = step / nch;
      if (step - tile * nch == 0) init_acc(bBs + (tile & 1) * kTileJ);
    } else {
      init_acc(bBs + s * kTileJ);
    }
#pragma unroll
    for (int ks = 0; ks < KC; ks++) {
      double af[4], bf[4];
#pragma unroll
      for (int mt = 0; mt < 4; mt++) af[mt] = At[(wi * 32 + mt * 8 + g) * XSP + ks * 4 + q];
#pragma unroll
      for (int nt = 0; nt < 4; nt++) bf[nt] = Bt[(wj * 32 + nt * 8 + g) * XSP + ks * 4 + q];
#pragma unroll
      for (int mt = 0; mt < 4; mt++)
#pragma unroll
        for (int nt = 0; nt < 4; nt++) dmma884(S[mt][nt], af[mt], bf[nt]);
    }
    if (MULTI) {
      // the stage only holds operands: hand it back before the epilogue (the biases live in their own slot, which is
      // rewritten two tiles later, when every warp has long left this tile)
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
      const int tile = step / nch;
      if (step - tile * nch == nch - 1) epilogue((tj_lo + tile) * kTileJ);
    } else {
      epilogue((tj_lo + step) * kTileJ);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);  // this warp no longer reads the slot (tile and biases)
    }
  }
}

// chunk-major copy of the descriptor rows for multi-chunk descriptors: Xch[ch][p][XSC] = Xc[p][ch*4*KC + (0..4*KC)), zero padded
__global__ void __launch_bounds__(256) chunk_rows_kernel(int64_t n, int XS, int KC, int nchunks, int XSC,
                                                         const double *__restrict__ Xc, double *__restrict__ Xch) {
  const int64_t total = (int64_t)nchunks * n * XSC;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int dd = (int)(e % XSC);
    const int64_t r = e / XSC;
    const int64_t p = r % n;
    const int ch = (int)(r / n);
    const int d = ch * 4 * KC + dd;
    Xch[e] = (dd < 4 * KC && d < XS) ? Xc[p * XS + d] : 0.0;
  }
}

__global__ void set_diagonal_kernel(int64_t n, double *__restrict__ K, int64_t ldk, double v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) K[i * ldk + i] = v;
}

static int launch_set_diagonal(int64_t n, double *K, int64_t ldk, double v, cudaStream_t st) {
  set_diagonal_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, K, ldk, v);
  KREG_LAUNCH_CHECK();
  return KREG_OK;
}

int launch_chunk_rows(int64_t n, int XS, int KC, int nchunks, int XSC, const double *Xc, double *Xch, cudaStream_t st) {
  if (n <= 0) return KREG_OK;
  chunk_rows_kernel<<<1184, 256, 0, st>>>(n, XS, KC, nchunks, XSC, Xc, Xch);
  KREG_LAUNCH_CHECK();
  return KREG_OK;
}

template <int KC, int EPI = EPI_K>
static int launch_values_kc(ValuesArgs a, cudaStream_t st) {
  constexpr int XSP = KC * 4 + ((KC * 4) % 8 == 4 ? 0 : 4);
  const bool multi = a.nchunks > 1 || EPI != EPI_K;  // the GEMM modes always stream chunk-major operands
  if (multi ? (a.XSC != XSP || !a.XAch || !a.XBch) : (a.XS != XSP)) return KREG_E_BADARG;  // workspace layout contract
  if (!a.biasA || !a.biasB) return KREG_E_BADARG;  // both are staged by bulk copies in every mode (EPI_LIN ignores the row terms)
  const int64_t r1 = a.row_end < a.na ? a.row_end : a.na;
  if (r1 <= a.row_begin) return KREG_OK;
  const int64_t ti0 = a.row_begin / kTileI, ti1 = (r1 - 1) / kTileI;
  const int64_t ntj = (a.nb + kTileJ - 1) / kTileJ;
  const int64_t tiles = ntj * (ti1 - ti0 + 1) / ((a.symmetric && !a.all_cols) ? 2 : 1);
#ifdef KREG_NJT_OVERRIDE
  a.njt = KREG_NJT_OVERRIDE;
#else
  if (a.njt <= 0)
    a.njt = tiles > 200000 ? 16 : tiles > 50000 ? 8 : tiles > 1000 ? 4 : 2;  // measured: 10k rows 4 (0.25 vs 0.27 ms), 50k flat
#endif
  dim3 grid((unsigned)((ntj + a.njt - 1) / a.njt), (unsigned)(ti1 - ti0 + 1));
  if (multi) {
    const size_t smem = (size_t)(2 * (kTileI + kTileJ) * XSP + 2 * kTileJ + kTileI + 64 + 2 * 2 + 2) * sizeof(double);
    auto kern = values_ring_kernel<KC, true, EPI>;
    KREG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 256, smem, st>>>(a);
  } else if constexpr (EPI == EPI_K) {
    constexpr int NST = KC <= 9 ? 3 : 2;
    const size_t smem = (size_t)(kTileI * XSP + NST * kTileJ * XSP + NST * kTileJ + kTileI + 64 + 2 * NST + 2) * sizeof(double);
    auto kern = values_ring_kernel<KC, false, EPI_K>;
    KREG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 256, smem, st>>>(a);
  }
  KREG_LAUNCH_CHECK();
  return KREG_OK;
}

// The GEMM modes used by the large-molecule prediction path (kreg_predict_big.cu): always chunks of kGemmKC k-steps.
int launch_gemm_epi(const ValuesArgs &a, int epi, cudaStream_t st) {
  switch (epi) {
    case EPI_W:
      return launch_values_kc<kGemmKC, EPI_W>(a, st);
    case EPI_WV:
      return launch_values_kc<kGemmKC, EPI_WV>(a, st);
    case EPI_LIN:
      return launch_values_kc<kGemmKC, EPI_LIN>(a, st);
    case EPI_KR:
      return launch_values_kc<kGemmKC, EPI_KR>(a, st);
    case EPI_WR:
      return launch_values_kc<kGemmKC, EPI_WR>(a, st);
    case EPI_K:
      return launch_values_kc<kGemmKC, EPI_K>(a, st);
    default:
      return KREG_E_BADARG;
  }
}

// k-steps per staged chunk: one chunk up to 12; longer descriptors in chunks of at most 9, the largest whose
// double-buffered A and B tiles (3072 * XSP bytes, XSP <= 36) leave room for two CTAs per SM (aspirin, D = 210:
// 3.93 -> 3.25 ms for a 20k x 20k block against chunks of 11)
#ifndef KREG_MULTI_KC
#define KREG_MULTI_KC 9
#endif
static void chunking(int KS, int *KC, int *nchunks) {
  const int cap = KS <= 12 ? 12 : KREG_MULTI_KC;
  *nchunks = (KS + cap - 1) / cap;
  *KC = (KS + *nchunks - 1) / *nchunks;
}

static int launch_values(ValuesArgs a, int KC, cudaStream_t st) {
  a.njt = 0;
  a.strideA = a.na;
  a.strideB = a.nb;
  switch (KC) {
#define KREG_KC(N) \
  case N:          \
    return launch_values_kc<N>(a, st);
    KREG_KC(1) KREG_KC(2) KREG_KC(3) KREG_KC(4) KREG_KC(5) KREG_KC(6) KREG_KC(7) KREG_KC(8) KREG_KC(9) KREG_KC(10)
    KREG_KC(11) KREG_KC(12)
#undef KREG_KC
    default:
      return KREG_E_UNSUPPORTED;
  }
}

// Centred, zero-padded descriptors + exponent bias for one geometry set.
// Centred descriptors of every training geometry, one thread per (geometry, atom pair): a single thread walking all D
// pairs of a geometry is a chain of D square roots and divisions (82 us for 500 aspirin geometries).
__global__ void __launch_bounds__(256) prep_descr_kernel(int nat, int64_t n, const double *__restrict__ xyz,
                                                         const double *__restrict__ xeq,
                                                         const double *__restrict__ center, int XS,
                                                         double *__restrict__ Xc) {
  const int D = descr_size(nat);
  const int64_t total = n * XS;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = e / XS;
    const int d = (int)(e - p * XS);
    double v = 0.0;  // padding columns
    if (d < D) {
      int a, b;
      pair_from_index(nat, d, a, b);
      const double *m = xyz + p * nat * 3;
      v = xeq[d] / sqrt(rij2_exact(m + 3 * a, m + 3 * b)) - center[d];
    }
    Xc[e] = v;
  }
}

// Row terms -c1 |x|^2 / 2 (scaled exponent units) and column terms -|x|^2 / 2 (units of S), summed in descriptor
// order (one thread per geometry; the sum is a short FMA chain).
__global__ void __launch_bounds__(256) prep_bias_kernel(int nat, int64_t n, int XS, double c1,
                                                        const double *__restrict__ Xc, double *__restrict__ bias,
                                                        double *__restrict__ bias_col) {
  const int D = descr_size(nat);
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const double *x = Xc + p * XS;
    double nn = 0.0;
    for (int d = 0; d < D; d++) nn = fma(x[d], x[d], nn);
    bias[p] = -0.5 * c1 * nn;
    if (bias_col) bias_col[p] = -0.5 * nn;
  }
}

int launch_prep_bias(int nat, int64_t n, int XS, double c1, const double *Xc, double *bias, double *bias_col,
                     cudaStream_t st) {
  if (n <= 0) return KREG_OK;
  prep_bias_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(nat, n, XS, c1, Xc, bias, bias_col);
  KREG_LAUNCH_CHECK();
  return KREG_OK;
}

int launch_prep_rows(int nat, int64_t n, const double *xyz, const double *xeq, const double *center, int XS,
                     double c1, double *Xc, double *bias, double *bias_col, cudaStream_t st) {
  if (n <= 0) return KREG_OK;
  const int64_t total = n * XS;
  const int blocks = (int)(total / 256 + 1 < 148 * 16 ? total / 256 + 1 : 148 * 16);
  prep_descr_kernel<<<blocks, 256, 0, st>>>(nat, n, xyz, xeq, center, XS, Xc);
  KREG_LAUNCH_CHECK();
  prep_bias_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(nat, n, XS, c1, Xc, bias, bias_col);
  KREG_LAUNCH_CHECK();
  return KREG_OK;
}

// =================================================================================================
// Gradient-augmented blocks.  One CTA per (geometry i, block of TJ geometries j); thread = one matrix
// column (j, b, u); it walks the 3 Nat rows (a, t) of geometry i, so every row segment it helps write is
// TJ * 3 Nat contiguous doubles.  Per element: 2 shared loads + 3 FP64 ops -> HBM-write bound.
// =================================================================================================
struct GradArgs {
  const double *xyz;     // [ntr][NAT][3]
  const double *Xc;      // [ntr][XS] centred descriptors
  const double *bias;    // [ntr]
  const double *E;       // [ntr][NAT][NAT][3]  E(a,c)[t] = dX_{d_ac}/dx_{a,t} (zero for a == c)
  int64_t ntr;
  int64_t nv;            // 0 or ntr
  int XS;
  double c1, inv_s2, lambdag;
  int fill;              // KREG_FILL_LOWER / KREG_FILL_FULL
  int64_t row_begin, row_end;  // absolute row range of K to produce
  double *K;
  int64_t ldk;
  double *P;             // [i1 - i0][ntr][3 Nat + 1 (+pad)] per-pair terms (pair_terms_small_kernel -> grad_emit_kernel)
  int64_t p_i0;          // first geometry i whose gradient rows intersect [row_begin, row_end): row 0 of P
  int64_t p_i1;          // one past the last such geometry
  int64_t c_i0, c_i1;    // the geometries THIS launch covers (a chunk of [p_i0, p_i1): the assembly is pipelined in chunks)
  int ES;                // doubles per E block: Nat*Nat*3 rounded up to even (16-byte aligned blocks for bulk copies)
};

// E(a,c)[t] = X_d (x_c - x_a)_t / R^2   (KREG.f90:121)
__global__ void __launch_bounds__(128) jacobian_rows_kernel(int nat, int64_t n, const double *__restrict__ xyz,
                                                            const double *__restrict__ xeq, double *__restrict__ E,
                                                            int ES) {
  const int64_t total = n * nat * nat;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = e / (nat * nat);
    const int ac = (int)(e - p * nat * nat);
    const int a = ac / nat, c = ac % nat;
    double *out = E + p * ES + ac * 3;
    if (ac == 0 && (nat * nat * 3) < ES) E[p * ES + ES - 1] = 0.0;  // pad element
    if (a == c) {
      out[0] = out[1] = out[2] = 0.0;
      continue;
    }
    const double *m = xyz + p * nat * 3;
    const int lo = a < c ? a : c, hi = a < c ? c : a;
    const double r2 = rij2_exact(m + 3 * lo, m + 3 * hi);
    const double x = xeq[pair_index(nat, lo, hi)] / sqrt(r2);
    const double f = x / r2;
    for (int t = 0; t < 3; t++) out[t] = f * (m[3 * c + t] - m[3 * a + t]);
  }
}

// Up to 10 atoms the gradient-augmented blocks are produced by two kernels so that the store kernel has NO per-pair prologue
// (larger molecules: grad_fused_kernel below):
//   pair_terms_small_kernel : per geometry pair (i, j): k_ij/s^2 and u_i^(j) = J_i^T (X_j - X_i)  -> workspace P[i][j][PW]
//                       (PW = 3 Nat + 1 doubles, 4 % of the bytes of the block it describes); also emits the
//                       value-gradient rows.
//   grad_emit_kernel  : CTA = ONE geometry i x TJ geometries j, thread = matrix column (j, b, u); one round of global
//                       loads, one barrier, then 3 Nat stores per thread.  Small, short-lived CTAs (4 per SM) write
//                       flat 3 Nat-row strips -- the schedule that measured fastest for this layout
//                       (profiles/r01_store_pattern.txt).
// tuning knobs of grad_emit_kernel (tools/build_variant.sh builds A/B variants with other values)
#ifndef KREG_EMIT_NJB
#define KREG_EMIT_NJB 8
#endif
#ifndef KREG_EMIT_TJEVEN
#define KREG_EMIT_TJEVEN 1
#endif
#ifndef KREG_EMIT_MAXCOLS
#define KREG_EMIT_MAXCOLS 256
#endif
#ifndef KREG_EMIT_MINB
#define KREG_EMIT_MINB 2
#endif
template <int NAT>
struct GradCfg {
  static constexpr int G3 = 3 * NAT;
  static constexpr int G3P = (G3 + 1) & ~1;   // E_i rows padded to an even length
  static constexpr int PW = (G3 + 3) & ~1;    // doubles per pair in the workspace: [k/s^2, pad, u_i[0..G3)] (+pad)
  static constexpr int D = NAT * (NAT - 1) / 2;
  static constexpr int TJ0 = (KREG_EMIT_MAXCOLS / G3) < 2 ? 2 : (KREG_EMIT_MAXCOLS / G3 > 16 ? 16 : KREG_EMIT_MAXCOLS / G3);
  // geometries j per emit step: even, so that with an even ldk every CTA's row segments start and end on 32-byte
  // sectors (Nat = 9: 216 columns = 54 sectors; 243 columns left partial sectors at both ends of every segment and
  // cost 5 % of the whole assembly, profiles/r01_store_pattern.txt)
  static constexpr int TJ = KREG_EMIT_TJEVEN ? (TJ0 & ~1) : TJ0;
  static constexpr int COLS = TJ * G3;
  static constexpr int THREADS = ((COLS + 31) / 32) * 32;
  static constexpr int NJB = KREG_EMIT_NJB;   // steps (of TJ geometries) walked by one emit CTA
};

template <int NAT>
struct PairSmall {
  using C = GradCfg<NAT>;
  static constexpr int JT = 128;                                 // pairs (= threads) per CTA
  static constexpr int D = C::D;
  static constexpr int D2 = (D + 1) & ~1;                        // doubles copied per descriptor row (16-byte multiple)
  static constexpr int XP = ((D2 / 2) % 2) ? D2 : D2 + 2;        // shared row stride: odd count of 16-byte units
  static constexpr int ES = (NAT * NAT * 3 + 1) & ~1;
  static constexpr int PWS = C::PW | 1;
  static constexpr int OFF_XI = ES;
  static constexpr int OFF_XJ = OFF_XI + D2;
  static constexpr int OFF_US = OFF_XJ;                          // the pair-term block reuses the X_j rows (dead once in registers)
  static constexpr int STAGE = JT * XP > JT * PWS ? JT * XP : JT * PWS;
  static constexpr int OFF_TAB = (OFF_XJ + STAGE + 1) & ~1;
  static constexpr int OFF_BAR = OFF_TAB + 64;
  static constexpr int DOUBLES = OFF_BAR + 2;
};

template <int NAT>
__global__ void __launch_bounds__(128, 5) pair_terms_small_kernel(const GradArgs a) {
  using C = GradCfg<NAT>;
  using SM = PairSmall<NAT>;
  constexpr int G3 = C::G3, PW = C::PW, PWS = SM::PWS, D = C::D, D2 = SM::D2, XP = SM::XP, JT = SM::JT;
  extern __shared__ __align__(128) double qsm[];
  double *Ei = qsm, *Xi = qsm + SM::OFF_XI, *Xj = qsm + SM::OFF_XJ, *Us = qsm + SM::OFF_US, *tab = qsm + SM::OFF_TAB;
  uint64_t *bar = reinterpret_cast<uint64_t *>(qsm + SM::OFF_BAR);

  const int64_t i = (int64_t)blockIdx.y + a.c_i0;  // this launch covers geometries [c_i0, c_i1)
  const int64_t rbase = a.nv + i * G3;
  if (rbase + G3 <= a.row_begin || rbase >= a.row_end) return;
  const bool full = a.fill == KREG_FILL_FULL;
  const int64_t jend = (a.nv || full) ? a.ntr : i + 1;  // value-gradient rows need every j
  const int64_t j0 = (int64_t)blockIdx.x * JT;
  if (j0 >= jend) return;
  const int nj = (int)(jend - j0 < JT ? jend - j0 : JT);
  const int tid = threadIdx.x;
  const bool active = tid < nj;

  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_fence_init();
    mbar_arrive_expect_tx(&bar[0], (unsigned)((SM::ES + D2 + nj * D2) * 8));
    bulk_g2s(Ei, a.E + i * a.ES, SM::ES * 8, &bar[0]);
    bulk_g2s(Xi, a.Xc + i * a.XS, D2 * 8, &bar[0]);
  }
  fill_exp_table(tab, tid, JT);
  __syncthreads();  // barrier armed (and table filled) before anyone adds copies to it or waits on it
  if (active) bulk_g2s(Xj + tid * XP, a.Xc + (j0 + tid) * a.XS, D2 * 8, &bar[0]);  // one padded row per thread
  mbar_wait(&bar[0], 0);

  double dl[D2];
  if (active) {
    const double2 *row = reinterpret_cast<const double2 *>(Xj + tid * XP);
    const double2 *xi2 = reinterpret_cast<const double2 *>(Xi);
#pragma unroll
    for (int p = 0; p < D2 / 2; p++) {
      const double2 r = row[p], x = xi2[p];
      dl[2 * p] = r.x - x.x;
      dl[2 * p + 1] = r.y - x.y;
    }
  }
  __syncthreads();  // every X_j row is in registers: its shared-memory space becomes the output block
  if (active) {
    double d2 = 0.0;
#pragma unroll
    for (int d = 0; d < D; d++) d2 = fma(dl[d], dl[d], d2);
    const double k = (j0 + tid == i) ? 1.0 : exp_from_scaled(-0.5 * a.c1 * d2, tab);
    const double ks = k * a.inv_s2;
    double *us = Us + tid * PWS;
    us[0] = ks;
    us[1] = 0.0;
    double *kcol = a.K + rbase * a.ldk + j0 + tid;  // K[rbase + q][j]
    const bool interior = rbase >= a.row_begin && rbase + G3 <= a.row_end;
#pragma unroll
    for (int aa = 0; aa < NAT; aa++) {
      const double *ea = Ei + aa * NAT * 3;
      double s[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int c = 0; c < NAT; c++)
        if (c != aa) {
          const double v = dl[pair_index(NAT, aa < c ? aa : c, aa < c ? c : aa)];
          s[0] = fma(ea[c * 3 + 0], v, s[0]);
          s[1] = fma(ea[c * 3 + 1], v, s[1]);
          s[2] = fma(ea[c * 3 + 2], v, s[2]);
        }
#pragma unroll
      for (int t = 0; t < 3; t++) {
        const int q = aa * 3 + t;
        us[2 + q] = s[t];
        if (a.nv) {
          const int64_t row = rbase + q;
          if (interior || (row >= a.row_begin && row < a.row_end)) {
            const double v = ks * s[t];
            KREG_CHECK_STORE(kcol + (int64_t)q * a.ldk, a.K + a.row_begin * a.ldk, (a.row_end - a.row_begin) * a.ldk);
            kcol[(int64_t)q * a.ldk] = v;
            if (full) a.K[(j0 + tid) * a.ldk + row] = v;
          }
        }
      }
    }
  }
  __syncthreads();
  // pair terms -> workspace (coalesced); only pairs whose gradient-gradient block is emitted are read back
  double *P = a.P + ((i - a.p_i0) * a.ntr + j0) * PW;
  const int64_t jgg = full ? a.ntr : i + 1;
  const int np = (int)(jgg - j0 < nj ? (jgg - j0 > 0 ? jgg - j0 : 0) : nj);
  for (int e = tid; e < np * PW; e += JT) P[e] = Us[(e / PW) * PWS + e % PW];
}

// CTA = ONE geometry i x up to NJB*TJ geometries j, walked TJ at a time; thread = matrix column (j, b, u) of the current
// step.  This thread's 3 Nat values E_i(b, .)[.] live in registers for the whole strip; per step the pair terms, the X_j
// rows and the E_j blocks of the TJ geometries -- three contiguous global ranges -- arrive as three TMA bulk copies
// (cp.async.bulk, one elected thread, mbarrier completion) into the other shared buffer while the current step's
// 3 Nat stores per thread are issued.  One barrier per step; shared reads are 16-byte wide where the layout allows.
template <int NAT>
struct EmitSmem {
  using C = GradCfg<NAT>;
  static constexpr int EJ = (NAT * NAT * 3 + 1) & ~1;   // one E_j block (== GradArgs::ES)
  static constexpr int XSMAX = ((C::D + 3) / 4) * 4 + 48;  // >= XS for any chunking (XS = 4*KC*nchunks)
  static constexpr int OFF_P = 0;                                 // [2][TJ*PW]
  static constexpr int OFF_X = OFF_P + 2 * C::TJ * C::PW;         // [2][TJ*XS]   (XS taken at run time, <= XSMAX)
  static constexpr int OFF_E = OFF_X + 2 * C::TJ * XSMAX;         // [2][TJ*EJ]
  static constexpr int OFF_XI = OFF_E + 2 * C::TJ * EJ;           // [XSMAX]
  static constexpr int OFF_BAR = OFF_XI + XSMAX;                  // 2 mbarriers
  static constexpr int DOUBLES = OFF_BAR + 2;
};

template <int NAT>
__global__ void __launch_bounds__(GradCfg<NAT>::THREADS, (NAT <= 12 && GradCfg<NAT>::THREADS <= 256 ? KREG_EMIT_MINB : 1)) grad_emit_kernel(const GradArgs a) {
  using C = GradCfg<NAT>;
  using SM = EmitSmem<NAT>;
  constexpr int G3 = C::G3, G3P = C::G3P, PW = C::PW, TJ = C::TJ, NJB = C::NJB, D = C::D, EJ = SM::EJ;
  extern __shared__ __align__(128) double esm[];
  double *Pb = esm + SM::OFF_P, *Xj = esm + SM::OFF_X, *Ejs = esm + SM::OFF_E, *Xi = esm + SM::OFF_XI;
  uint64_t *bar = reinterpret_cast<uint64_t *>(esm + SM::OFF_BAR);
  const int XS = a.XS;

  const int64_t i = (int64_t)blockIdx.y + a.c_i0;  // this launch covers geometries [c_i0, c_i1)
  const bool full = a.fill == KREG_FILL_FULL;
  const int64_t jend = full ? a.ntr : i + 1;
  const int64_t jb0 = (int64_t)blockIdx.x * (NJB * TJ);
  if (jb0 >= jend) return;
  const int64_t rbase = a.nv + i * G3;
  if (rbase + G3 <= a.row_begin || rbase >= a.row_end) return;
  const int nsteps = (int)(((jend - jb0 < NJB * TJ ? jend - jb0 : NJB * TJ) + TJ - 1) / TJ);

  const int tid = threadIdx.x;
  const int jl = tid / G3, col = tid % G3;
  const int b = col / 3, u = col % 3;
  const bool colthread = tid < C::COLS;
  const bool interior = rbase >= a.row_begin && rbase + G3 <= a.row_end;
  const int64_t ld = a.ldk;

  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_fence_init();
  }
  for (int e = tid; e < D; e += C::THREADS) Xi[e] = a.Xc[i * XS + e];
  __syncthreads();

  auto prefetch = [&](int n) {  // called by thread 0 only
    const int64_t jb = jb0 + (int64_t)n * TJ;
    const unsigned npair = (unsigned)(a.ntr - jb < TJ ? a.ntr - jb : TJ);
    const int buf = n & 1;
    const unsigned bp = npair * PW * 8, bx = npair * (unsigned)XS * 8, be = npair * EJ * 8;
    mbar_arrive_expect_tx(&bar[buf], bp + bx + be);
    bulk_g2s(Pb + buf * TJ * PW, a.P + ((i - a.p_i0) * a.ntr + jb) * PW, bp, &bar[buf]);
    bulk_g2s(Xj + buf * TJ * SM::XSMAX, a.Xc + jb * XS, bx, &bar[buf]);
    bulk_g2s(Ejs + buf * TJ * EJ, a.E + jb * a.ES, be, &bar[buf]);
  };
  if (tid == 0) prefetch(0);
  __syncwarp();
  double e27[G3P];  // E_i(b, c)[t] for all (c, t);  E_i(a, b)[t] = -E_i(b, a)[t]
  if (colthread) {
#pragma unroll
    for (int z = 0; z < G3; z++) e27[z] = a.E[i * a.ES + b * G3 + z];
  }

  for (int n = 0; n < nsteps; n++) {
    __syncthreads();  // everyone left buffer (n+1)&1 (read during step n-1)
    if (tid == 0 && n + 1 < nsteps) prefetch(n + 1);
    __syncwarp();  // warp 0 reconverges before the tile code (uniform registers are shared with the issuing lane)
    mbar_wait(&bar[n & 1], (n >> 1) & 1);  // step n's operands have landed
    const int64_t jb = jb0 + (int64_t)n * TJ;
    const int64_t j = jb + jl;
    const bool work = colthread && j < jend;
    const int buf = n & 1;
    const double *pb = Pb + buf * TJ * PW + (colthread ? jl : 0) * PW;
    if (work) {
      double nej[NAT], kjd[3];
      const double *xj = Xj + buf * TJ * SM::XSMAX + jl * XS;
      const double *ejb = Ejs + buf * TJ * EJ + jl * EJ + b * G3 + u;  // E_j(b, c)[u] at ejb[3c]
#pragma unroll
      for (int c = 0; c < NAT; c++) nej[c] = -ejb[3 * c];
      const double kj = pb[0];
      double sj = 0.0;  // -u_j[col] = -sum_c E_j(b,c)[u] (X_j - X_i)[d_bc]
#pragma unroll
      for (int c = 0; c < NAT; c++)
        if (c != b) {
          const int d = pair_index(NAT, b < c ? b : c, b < c ? c : b);
          sj = fma(nej[c], xj[d] - Xi[d], sj);
        }
      const double c2 = (sj * a.inv_s2) * kj;  // -(k/s^2) u_j[col] / s^2
#pragma unroll
      for (int t = 0; t < 3; t++) {  // diagonal 3x3 block (a == b): (k/s^2) sum_c E_i(b,c)[t] E_j(b,c)[u]
        double acc = 0.0;
#pragma unroll
        for (int c = 0; c < NAT; c++) acc = fma(e27[c * 3 + t], nej[c], acc);
        kjd[t] = -acc * kj;
      }
      double *dst = a.K + rbase * ld + a.nv + j * G3 + col;
      if (interior && j != i) {
#pragma unroll
        for (int q = 0; q < G3; q++) {
          const int aa = q / 3, t = q % 3;
          double v = fma(pb[2 + q], c2, kj * (e27[q] * nej[aa]));
          if (aa == b) v += kjd[t];
          KREG_CHECK_STORE(dst, a.K + a.row_begin * ld, (a.row_end - a.row_begin) * ld);
          *dst = v;
          dst += ld;
        }
      } else {
        // boundary (rare): rows clipped to [row_begin, row_end), lambda on the matrix diagonal; operands re-read
        const double *eig = a.E + i * a.ES + b * G3;
#pragma unroll 1
        for (int q = 0; q < G3; q++) {
          const int aa = q / 3, t = q % 3;
          double v = fma(pb[2 + q], c2, kj * (eig[q] * -ejb[3 * aa]));
          if (aa == b) v += (t == 0 ? kjd[0] : t == 1 ? kjd[1] : kjd[2]);
          if (j == i && q == col) v += a.lambdag;
          const int64_t r = rbase + q;
          if (r >= a.row_begin && r < a.row_end) {
            KREG_CHECK_STORE(dst + (int64_t)q * ld, a.K + a.row_begin * ld, (a.row_end - a.row_begin) * ld);
            dst[(int64_t)q * ld] = v;
          }
        }
      }
    }
  }
}

// Fused version of the resident-block kernel: the pair terms are formed IN the emit step instead of by a kernel of their
// own (0.79 of 3.3 ms at 21 atoms, and a workspace of 3 Nat + 1 doubles per geometry pair written and read back).  With
// Dsq[j][c][b] = (X_j - X_i)[d_bc] in shared memory, the thread that owns column (j, b, u) forms in ONE loop over c
//     -u_j[(b,u)]     = sum_c E_j[c][b][u] Dsq[j][c][b]
//      u_i^(j)[(b,u)] = -sum_c E_i[c][b][u] Dsq[j][c][b]
//      |X_j - X_i|^2  = 1/2 sum_b sum_c Dsq[j][c][b]^2         (column b's share; summed over b in a fixed order)
//      m[t]           = sum_c E_j[c][b][u] E_i[c][b][t]        (the 3x3 diagonal blocks a == b are k/s^2 m[t] + rank-1 term)
// and the value-gradient entries of BOTH orders of the pair come out of the same step:
//      K[NtrVal + i*3Nat + (b,u)][j] = (k/s^2) u_i^(j)[(b,u)],     K[NtrVal + j*3Nat + (b,u)][i] = (k/s^2) u_j^(i)[(b,u)] = (k/s^2) sj.
// A rank that owns only a row slab needs the second kind for its geometries j and EVERY i > j, also those whose rows
// belong to other ranks: such steps run with the block emission switched off (`pair-only`).
// The kernel is bound by shared-memory wavefronts (84 % of the L1 data pipe at 21 atoms in its first version,
// profiles/r02_gradfused21_v1_ncu.txt: every 64-bit LDS with more than one address costs two), hence:
//  * NAT_CT > 0 (11..24 atoms): the thread's column of the RESIDENT E_j block, E_j[c][b][u] for all c, lives in Nat
//    registers for the CTA's lifetime; NAT_CT = 0 takes the size at run time and reads it from shared memory;
//  * EIG = false: E_i in shared memory is read with LDS (a run-time choice of pointer compiles to generic loads);
//  * one pass over the rows, strictly in order (the diagonal sums m[t] come from the prologue loop);
//  * at most 128 registers (two CTAs of 256 threads per SM): left alone, ptxas picks 128 for some sizes and 170-240 for
//    others (hoisted loads of the unrolled loops), which halves the resident warps -- 3.1 against 4.4 TB/s at 11 atoms;
//  * the value-gradient entries are 3 % of the bytes and 10 % of the time (small scattered writes: 5.0 TB/s without them,
//    4.5 with, at 21 atoms).  K[row of j][i] comes one entry per row per step; the thread stages eight steps in shared
//    memory and writes aligned 64-byte lines -- single 8-byte stores leave partial sectors behind that are evicted
//    before the next steps complete them (+0.25 ms at 21 atoms, +1.3 ms at 11).  Writing both kinds cooperatively after
//    the next step's barrier (TJ-wide runs; eight lanes per line) gained 1-3 % up to 21 atoms and cost an SM's second
//    CTA from 24 atoms on (4.62 -> 2.62 TB/s): not kept.
template <int NAT_CT, bool EIG>
__global__ void __launch_bounds__(256, 2) grad_fused_kernel(const GradArgs a, int nat_rt, int TJ, int NIB, int64_t i_end) {
  constexpr bool MFOLD = !EIG;  // diagonal sums from the prologue loop (E_i in shared memory: the extra loads are cheap)
  const int nat = NAT_CT > 0 ? NAT_CT : nat_rt;
  const int G3 = 3 * nat, D = descr_size(nat), NS = nat | 1, EJ = a.ES, XS = a.XS;
  const int T = blockDim.x;
  extern __shared__ __align__(128) double esm[];
  double *Ejs = esm;                                // [TJ*EJ]     resident
  double *Xj = Ejs + (size_t)TJ * EJ;               // [TJ*XS]     resident
  double *Eib = Xj + (size_t)TJ * XS;               // [2][EJ]     streamed (very large molecules: read through L1 instead)
  double *Xib = Eib + (EIG ? 0 : 2 * (size_t)EJ);   // [2][XS]
  double *Dsq = Xib + 2 * (size_t)XS;               // [TJ][nat*NS]
  double *Us = Dsq + (((size_t)TJ * nat * NS + 1) & ~(size_t)1);  // [TJ][G3]   u_i^(j)
  double *d2p = Us + (((size_t)TJ * G3 + 1) & ~(size_t)1);        // [TJ][nat]  per-column shares of |X_j - X_i|^2
  double *tab = d2p + (((size_t)TJ * nat + 1) & ~(size_t)1);      // [64]
  double *VG = tab + 64;                            // [8][T]      K[row of (j,b,u)][i] of up to 8 steps
  uint64_t *bar = reinterpret_cast<uint64_t *>(VG + 8 * T);            // [3]
  unsigned short *ptab = reinterpret_cast<unsigned short *>(bar + 4);  // [D] d -> (lo | hi << 8)

  const bool full = a.fill == KREG_FILL_FULL;
  const int64_t jb = (int64_t)blockIdx.x * TJ;
  const int npair = (int)(a.ntr - jb < TJ ? a.ntr - jb : TJ);
  // geometries i of this CTA: its chunk of [c_i0, i_end); steps with i >= c_i1 are pair-only and are needed only by j-blocks
  // that contain geometries of this launch's own range
  int64_t i_lo = (int64_t)blockIdx.y * NIB + a.c_i0, i_hi = i_lo + NIB;
  if (i_hi > i_end) i_hi = i_end;
  if (!full && i_lo < jb) i_lo = jb;
  if (i_lo >= a.c_i1 && (jb + npair <= a.c_i0 || jb >= a.c_i1)) return;
  if (i_lo >= i_hi) return;
  const int nsteps = (int)(i_hi - i_lo);

  const int tid = threadIdx.x;
  const int64_t ld = a.ldk;
  const bool okc = tid < npair * G3;
  const int r = okc ? tid : 0;
  const int jl = r / G3, cc = r - jl * G3, b = cc / 3, u = cc - 3 * b;
  const int64_t j = jb + jl;
  const int64_t row2 = a.nv + j * G3 + cc;
  const bool w2 = okc && a.nv > 0 && !full && row2 >= a.row_begin && row2 < a.row_end;
  double *const dst2 = a.K + row2 * ld;

  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_init(&bar[2], 1);
    mbar_fence_init();
  }
  fill_exp_table(tab, tid, T);
  __syncthreads();
  auto prefetch = [&](int n) {  // thread 0 only: operands of step n (geometry i_lo + n)
    const int64_t i = i_lo + n;
    const int buf = n & 1;
    const unsigned be = EIG ? 0u : (unsigned)EJ * 8, bx = (unsigned)XS * 8;
    mbar_arrive_expect_tx(&bar[buf], be + bx);
    if (!EIG) bulk_g2s(Eib + (size_t)buf * EJ, a.E + i * a.ES, be, &bar[buf]);
    bulk_g2s(Xib + (size_t)buf * XS, a.Xc + i * XS, bx, &bar[buf]);
  };
  if (tid == 0) {
    const unsigned be = (unsigned)npair * EJ * 8, bx = (unsigned)npair * XS * 8;
    mbar_arrive_expect_tx(&bar[2], be + bx);
    bulk_g2s(Ejs, a.E + jb * a.ES, be, &bar[2]);
    bulk_g2s(Xj, a.Xc + jb * XS, bx, &bar[2]);
    prefetch(0);
  }
  __syncwarp();
  for (int e = tid; e < D; e += T) {
    int lo, hi;
    pair_from_index(nat, e, lo, hi);
    ptab[e] = (unsigned short)(lo | (hi << 8));
  }
  for (int e = tid; e < TJ * nat; e += T) Dsq[(size_t)e * NS + (e % nat)] = 0.0;  // diagonals: never rewritten
  mbar_wait(&bar[2], 0);
  const double *ej = Ejs + (size_t)jl * EJ + cc;  // E_j[c][b][u] at ej[c * G3]  (= -E_j(b,c)[u])
  double ejr[NAT_CT > 0 ? NAT_CT : 1];
  if constexpr (NAT_CT > 0) {
#pragma unroll
    for (int c = 0; c < NAT_CT; c++) ejr[c] = ej[c * G3];
  }
  const double *dq = Dsq + (size_t)jl * nat * NS + b;
  const double *ui = Us + (size_t)jl * G3;
  const double half_c1 = -0.5 * a.c1;
#define KREG_EJ(c) (NAT_CT > 0 ? ejr[c] : ej[(c) * G3])

  for (int n = 0; n < nsteps; n++) {
    __syncthreads();  // everyone left the buffer about to be refilled and the shared arrays of the previous step
    if (tid == 0 && n + 1 < nsteps) prefetch(n + 1);
    __syncwarp();  // warp 0 reconverges before the tile code (uniform registers are shared with the issuing lane)
    const int buf = n & 1;
    mbar_wait(&bar[buf], (n >> 1) & 1);
    const int64_t i = i_lo + n;
    const double *Xi = Xib + (size_t)buf * XS;
    for (int e = tid; e < npair * D; e += T) {  // consecutive threads -> consecutive d: conflict-free reads
      const int jj = e / D, d = e - jj * D;
      const int lo = ptab[d] & 0xff, hi = ptab[d] >> 8;
      const double v = Xj[jj * XS + d] - Xi[d];
      double *q = Dsq + (size_t)jj * nat * NS;
      q[lo * NS + hi] = v;
      q[hi * NS + lo] = v;
    }
    __syncthreads();
    const bool want = okc && (full || j <= i);
    // E_i[c][b][t] at ei[c * G3 + t]  (= -E_i(b,c)[t])
    const double *ei;
    if constexpr (EIG) ei = a.E + i * a.ES + 3 * b;
    else ei = Eib + (size_t)buf * EJ + 3 * b;
    double sj = 0.0, uo = 0.0, dd = 0.0, m0 = 0.0, m1 = 0.0, m2 = 0.0;
    if (want) {
      if constexpr (MFOLD) {
#pragma unroll
        for (int c = 0; c < nat; c++) {
          const double dl = dq[c * NS], ejc = KREG_EJ(c);
          const double e0 = ei[c * G3], e1 = ei[c * G3 + 1], e2 = ei[c * G3 + 2];
          const double eu = u == 0 ? e0 : (u == 1 ? e1 : e2);
          sj = fma(ejc, dl, sj);   // -u_j[col]
          uo = fma(-eu, dl, uo);   // u_i^(j)[col]
          dd = fma(dl, dl, dd);
          m0 = fma(ejc, e0, m0);
          m1 = fma(ejc, e1, m1);
          m2 = fma(ejc, e2, m2);
        }
      } else {
        for (int c = 0; c < nat; c++) {
          const double dl = dq[c * NS];
          sj = fma(KREG_EJ(c), dl, sj);
          uo = fma(-ei[c * G3 + u], dl, uo);
          dd = fma(dl, dl, dd);
        }
      }
      Us[(size_t)jl * G3 + cc] = uo;
      if (u == 0) d2p[jl * nat + b] = dd;
    }
    __syncthreads();
    if (!want) continue;
    double d2 = 0.0;
    for (int c = 0; c < nat; c++) d2 += d2p[jl * nat + c];  // every thread of the pair sums in the same order
    const double kk = (j == i) ? 1.0 : exp_from_scaled(half_c1 * (0.5 * d2), tab);
    const double kj = kk * a.inv_s2;
    const int64_t rbase = a.nv + i * G3;
    // value-gradient entries of both orders of the pair
    if (a.nv) {
      const int64_t row1 = rbase + cc;
      if (row1 >= a.row_begin && row1 < a.row_end) {
        const double v = kj * uo;
        KREG_CHECK_STORE(a.K + row1 * ld + j, a.K + a.row_begin * ld, (a.row_end - a.row_begin) * ld);
        a.K[row1 * ld + j] = v;
        if (full) a.K[j * ld + row1] = v;
      }
      if (w2) {  // a slot per thread: no synchronisation
        const int slot = (int)((reinterpret_cast<uintptr_t>(dst2 + i) >> 3) & 7);  // position in the 64-byte line of K
        if (j != i) VG[slot * T + tid] = kj * sj;
        if (slot == 7 || n == nsteps - 1) {
          int64_t lo = i - slot;
          if (lo < i_lo) lo = i_lo;
          if (lo < j + 1) lo = j + 1;
          for (int64_t ik = lo; ik <= i; ik++) {
            KREG_CHECK_STORE(dst2 + ik, a.K + a.row_begin * ld, (a.row_end - a.row_begin) * ld);
            dst2[ik] = VG[(slot - (int)(i - ik)) * T + tid];
          }
        }
      }
    }
    if (i >= a.c_i1 || rbase + G3 <= a.row_begin || rbase >= a.row_end) continue;  // pair-only step
    const double c2 = (sj * a.inv_s2) * kj;  // -(k/s^2) u_j[col] / s^2
    m0 *= kj;
    m1 *= kj;
    m2 *= kj;
    double *dst = a.K + rbase * ld + a.nv + j * G3 + cc;
    if (rbase >= a.row_begin && rbase + G3 <= a.row_end && j != i) {
#pragma unroll
      for (int aa = 0; aa < nat; aa++) {
        const double knej = -(kj * KREG_EJ(aa));
        double p0 = __dmul_rn(knej, ei[aa * G3 + 0]), p1 = __dmul_rn(knej, ei[aa * G3 + 1]),
               p2 = __dmul_rn(knej, ei[aa * G3 + 2]);
        if constexpr (MFOLD) {
          if (aa == b) {
            p0 = m0;
            p1 = m1;
            p2 = m2;
          }
        } else {
          m0 = __dadd_rn(m0, p0);
          m1 = __dadd_rn(m1, p1);
          m2 = __dadd_rn(m2, p2);
          if (aa == b) continue;  // the three a == b rows are completed after the loop
        }
        double *d3 = dst + (int64_t)(3 * aa) * ld;
        KREG_CHECK_STORE(d3 + 2 * ld, a.K + a.row_begin * ld, (a.row_end - a.row_begin) * ld);
        d3[0] = fma(ui[3 * aa + 0], c2, p0);
        d3[ld] = fma(ui[3 * aa + 1], c2, p1);
        d3[2 * ld] = fma(ui[3 * aa + 2], c2, p2);
      }
      if constexpr (!MFOLD) {
        double *dg = dst + (int64_t)(3 * b) * ld;
        dg[0] = fma(ui[3 * b + 0], c2, -m0);
        dg[ld] = fma(ui[3 * b + 1], c2, -m1);
        dg[2 * ld] = fma(ui[3 * b + 2], c2, -m2);
      }
    } else {
      // boundary (rare): rows clipped to [row_begin, row_end), lambda on the matrix diagonal; the same numbers as above
      double mm[3] = {m0, m1, m2};
      if constexpr (!MFOLD) {
        for (int aa = 0; aa < nat; aa++) {
          const double knej = -(kj * KREG_EJ(aa));
          for (int t = 0; t < 3; t++) mm[t] = __dadd_rn(mm[t], __dmul_rn(knej, ei[aa * G3 + t]));
        }
        for (int t = 0; t < 3; t++) mm[t] = -mm[t];
      }
#pragma unroll
      for (int aa = 0; aa < nat; aa++) {
        const double knej = -(kj * KREG_EJ(aa));
#pragma unroll
        for (int t = 0; t < 3; t++) {
          double pt = __dmul_rn(knej, ei[aa * G3 + t]);
          if (aa == b) pt = t == 0 ? mm[0] : (t == 1 ? mm[1] : mm[2]);
          double val = fma(ui[3 * aa + t], c2, pt);
          const int q = 3 * aa + t;
          if (j == i && q == cc) val += a.lambdag;
          const int64_t rr = rbase + q;
          if (rr >= a.row_begin && rr < a.row_end) {
            KREG_CHECK_STORE(dst + (int64_t)q * ld, a.K + a.row_begin * ld, (a.row_end - a.row_begin) * ld);
            dst[(int64_t)q * ld] = val;
          }
        }
      }
    }
  }
#undef KREG_EJ
}

static size_t fused_smem_doubles(int nat, int XS, int EJ, int tj, int ei_global) {
  const int G3 = 3 * nat, D = descr_size(nat), NS = nat | 1;
  const size_t threads = (size_t)((tj * G3 + 31) / 32) * 32;
  size_t n = (size_t)tj * ((size_t)EJ + XS) + 2 * ((ei_global ? 0 : (size_t)EJ) + XS) + (((size_t)tj * nat * NS + 1) & ~(size_t)1) +
             (((size_t)tj * G3 + 1) & ~(size_t)1) + (((size_t)tj * nat + 1) & ~(size_t)1) + 64;
  n += 8 * threads;  // staged value-gradient entries
  n += 4 + (D + 3) / 4;  // barriers, pair table (unsigned short)
  return n;
}

#ifndef KREG_EMIT_TEMPLATE_MAX
#define KREG_EMIT_TEMPLATE_MAX 10  // largest molecule emitted by the register-resident grad_emit_kernel<NAT>
#endif
static int env_int(const char *name, int dflt) {
  const char *s = getenv(name);
  return (s && *s) ? atoi(s) : dflt;
}

// molecules above the register-resident emit kernel's range: pair terms formed inside the emit kernel (grad_fused_kernel)
static bool grad_fused(int nat) { return nat > KREG_EMIT_TEMPLATE_MAX; }

// (KREG_RT_H / KREG_RT_NIB / KREG_RT_EIG override the launch plan, for tuning runs)
static int launch_grad_fused(int nat, const GradArgs &a, cudaStream_t st) {
  if (a.c_i1 <= a.c_i0) return KREG_OK;
  const int G3 = 3 * nat, XS = a.XS, EJ = a.ES;
  // geometries j per CTA: as many thread groups of 3 Nat columns as 256 threads hold; two CTAs per SM when the resident
  // block allows it.  (Two geometries per thread -- half the E_i loads per output -- was measured and dropped: the 2 Nat
  // registers of E_j columns halve the resident warps, 4.19 against 4.38 TB/s at 21 atoms, 3.95 against 5.07 at 12.)
  int h = 256 / G3;
  if (h < 1) h = 1;
  if (h > 16) h = 16;
  const size_t two_ctas = 112 * 1024 / 8, one_cta = 224 * 1024 / 8;
  auto need = [&](int hh, int eig) { return fused_smem_doubles(nat, XS, EJ, hh, eig); };
  int h2 = h;
  while (h2 > 1 && need(h2, 0) > two_ctas) h2--;
  if (need(h2, 0) <= two_ctas && 2 * h2 >= h) {
    h = h2;
  } else {
    while (h > 1 && need(h, 0) > one_cta) h--;
  }
  h = env_int("KREG_RT_H", h);
  if (h < 1) h = 1;
  const int eig = env_int("KREG_RT_EIG", need(h, 0) > one_cta ? 1 : 0);
  const int tj = h;
  int nib = env_int("KREG_RT_NIB", 16);
  if (nib < 1) nib = 1;
  const int threads = ((h * G3 + 31) / 32) * 32;
  const size_t smem = need(h, eig) * sizeof(double);
  if (threads > 256 || smem > 227 * 1024) return KREG_E_UNSUPPORTED;
  // a row slab of a matrix with value rows also owns K[its gradient rows][i] for every LATER geometry i (pair-only steps)
  const bool full = a.fill == KREG_FILL_FULL;
  const int64_t i_end = (!full && a.nv > 0) ? a.ntr : a.c_i1;
  dim3 grid((unsigned)((a.ntr + tj - 1) / tj), (unsigned)((i_end - a.c_i0 + nib - 1) / nib));
#define KREG_LAUNCH_FUSED(KERN, NRT)                                                                             \
  do {                                                                                                           \
    KREG_CUDA_TRY(cudaFuncSetAttribute(KERN, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
    KERN<<<grid, threads, smem, st>>>(a, NRT, tj, nib, i_end);                                                   \
  } while (0)
  if (nat <= KREG_FUSED_MAX_NATOMS && !eig) {
    switch (nat) {
#define KREG_CASE(N)                                      \
  case N:                                                 \
    KREG_LAUNCH_FUSED((grad_fused_kernel<N, false>), N);  \
    break;
      KREG_CASE(11) KREG_CASE(12) KREG_CASE(13) KREG_CASE(14) KREG_CASE(15) KREG_CASE(16) KREG_CASE(17) KREG_CASE(18)
      KREG_CASE(19) KREG_CASE(20) KREG_CASE(21) KREG_CASE(22) KREG_CASE(23) KREG_CASE(24)
#undef KREG_CASE
      default:
        return KREG_E_UNSUPPORTED;
    }
  } else if (eig) {
    KREG_LAUNCH_FUSED((grad_fused_kernel<0, true>), nat);
  } else {
    KREG_LAUNCH_FUSED((grad_fused_kernel<0, false>), nat);
  }
#undef KREG_LAUNCH_FUSED
  KREG_LAUNCH_CHECK();
  return KREG_OK;
}

// up to KREG_EMIT_TEMPLATE_MAX atoms: pair terms (+ value-gradient rows) into the workspace, then the register-resident emit
template <int NAT>
static int launch_grad_nat(const GradArgs &a, cudaStream_t st) {
  using C = GradCfg<NAT>;
  using SM = PairSmall<NAT>;
  {
    auto kern = pair_terms_small_kernel<NAT>;
    const size_t smem = (size_t)SM::DOUBLES * sizeof(double);
    KREG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((a.ntr + SM::JT - 1) / SM::JT), (unsigned)(a.c_i1 - a.c_i0));
    kern<<<grid, SM::JT, smem, st>>>(a);
    KREG_LAUNCH_CHECK();
  }
  dim3 grid((unsigned)((a.ntr + C::NJB * C::TJ - 1) / (C::NJB * C::TJ)), (unsigned)(a.c_i1 - a.c_i0));
  auto emit = grad_emit_kernel<NAT>;
  const size_t esmem = (size_t)EmitSmem<NAT>::DOUBLES * sizeof(double);
  KREG_CUDA_TRY(cudaFuncSetAttribute(emit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)esmem));
  emit<<<grid, C::THREADS, esmem, st>>>(a);
  KREG_LAUNCH_CHECK();
  return KREG_OK;
}

// Gradient blocks of the geometries [g.p_i0, g.p_i1) on the caller's stream.  Which kernels run, by molecule size (measured on
// B200):
//   <= 10 atoms  pair_terms_small_kernel<NAT> + grad_emit_kernel<NAT> (operands in registers: 5.1 TB/s at 9 atoms; the
//                register version spills from 11 atoms up and stays at 3.2)
//   >= 11        grad_fused_kernel (j-block resident in shared memory, pair terms formed in the emit step: 4.4-4.9 TB/s
//                from 11 to 24 atoms, 3.2 at 32, 1.8 at 64)
// (Pipelining the two kernels of the small path over chunks of geometries on two side streams was measured and dropped:
// 5.14 -> 5.02 / 4.89 / 4.70 TB/s with 4 / 8 / 16 chunks at 9 atoms -- the emit kernel already fills every SM, so the second
// kernel only takes CTA slots away from it.)
static int launch_grad_stages(int nat, GradArgs g, cudaStream_t st) {
  if (g.p_i1 <= g.p_i0) return KREG_OK;
  g.c_i0 = g.p_i0;
  g.c_i1 = g.p_i1;
  if (grad_fused(nat)) return launch_grad_fused(nat, g, st);
  switch (nat) {
#define KREG_CASE(N) \
  case N:            \
    return launch_grad_nat<N>(g, st);
    KREG_CASE(2) KREG_CASE(3) KREG_CASE(4) KREG_CASE(5) KREG_CASE(6) KREG_CASE(7) KREG_CASE(8) KREG_CASE(9) KREG_CASE(10)
#undef KREG_CASE
    default:
      return KREG_E_UNSUPPORTED;
  }
}

// workspace carve-up helpers ---------------------------------------------------------------
struct BuildWs {
  size_t off_center, off_X, off_Xch, off_bias, off_biasc, off_E, off_P, total;
  int XS, KC, nchunks, XSC;
};

// geometries whose gradient rows intersect the row range [row_begin, row_end) of a matrix with nv value rows
static void grad_geom_range(int nat, int64_t n, int64_t nv, int64_t row_begin, int64_t row_end, int64_t *i0, int64_t *i1) {
  const int64_t G3 = 3 * (int64_t)nat;
  int64_t a = row_begin > nv ? (row_begin - nv) / G3 : 0;
  int64_t b = row_end > nv ? (row_end - nv + G3 - 1) / G3 : 0;
  if (b > n) b = n;
  if (a > b) a = b;
  *i0 = a;
  *i1 = b;
}

static BuildWs build_ws(int nat, int64_t n, bool with_grad, int64_t pair_rows = -1) {
  BuildWs w;
  if (pair_rows < 0) pair_rows = n;
  const TcShape sh(nat);
  chunking(sh.KS, &w.KC, &w.nchunks);
  w.XS = w.KC * w.nchunks * 4;
  if (w.XS % 8 == 0) w.XS += 4;  // row stride = 4 mod 8 doubles: rows can be bulk-copied into conflict-free shared tiles
  size_t off = 0;
  auto take = [&](size_t doubles) {
    size_t o = off;
    off += (doubles * sizeof(double) + 255) / 256 * 256;
    return o;
  };
  w.off_center = take(sh.D);
  w.off_X = take((size_t)n * w.XS);
  w.XSC = w.KC * 4 + ((w.KC * 4) % 8 == 4 ? 0 : 4);
  w.off_Xch = w.nchunks > 1 ? take((size_t)w.nchunks * n * w.XSC) : off;  // chunk-major copy for the values kernel
  w.off_bias = take((size_t)n);
  w.off_biasc = take((size_t)n);
  w.off_E = with_grad ? take((size_t)n * ((nat * nat * 3 + 1) & ~1)) : off;
  w.off_P = (with_grad && !grad_fused(nat)) ? take((size_t)pair_rows * n * ((3 * nat + 3) & ~1)) : off;
  w.total = off;
  return w;
}

}  // namespace kreg

using namespace kreg;

extern "C" size_t kreg_build_workspace_bytes(int natoms, int64_t ntrain, int64_t ntrgr) {
  if (natoms < KREG_MIN_NATOMS || natoms > KREG_MAX_NATOMS || ntrain < 0) return 0;
  return build_ws(natoms, ntrain, ntrgr > 0).total;
}

namespace kreg {

// Laplacian kernel exp(-|X_i - X_j|_1 / sigma) (A_KRR_kernel.f90:830-831 with the L1 distance of :2066-2069): not a GEMM;
// a 32 x 32 tile of pairs per CTA, descriptor chunks of 32 staged in shared memory (odd stride: conflict-free columns).
__global__ void __launch_bounds__(256) laplacian_block_kernel(int D, int64_t na, const double *__restrict__ XA, int64_t nb,
                                                              const double *__restrict__ XB, int XS, double inv_sigma,
                                                              int symmetric, double lambda, double *__restrict__ K,
                                                              int64_t ldk) {
  __shared__ double As[32][33], Bs[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int64_t i0 = (int64_t)blockIdx.y * 32, j0 = (int64_t)blockIdx.x * 32;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int d0 = 0; d0 < D; d0 += 32) {
    for (int r = ty; r < 32; r += 8) {
      const int d = d0 + tx;
      As[r][tx] = (i0 + r < na && d < D) ? XA[(i0 + r) * XS + d] : 0.0;
      Bs[r][tx] = (j0 + r < nb && d < D) ? XB[(j0 + r) * XS + d] : 0.0;
    }
    __syncthreads();
    for (int d = 0; d < 32; d++) {
      const double b = Bs[tx][d];
#pragma unroll
      for (int k = 0; k < 4; k++) acc[k] += fabs(As[ty + 8 * k][d] - b);
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int64_t i = i0 + ty + 8 * k, j = j0 + tx;
    if (i < na && j < nb) K[i * ldk + j] = (symmetric && i == j) ? 1.0 + lambda : exp(-acc[k] * inv_sigma);
  }
}

// Matern kernel of order nn as polynomials in y = 2 r / sigma times exp(-r / sigma):
//   k   = sum_m mat_c[m] y^m, m = 0..nn     (Kfunk, A_KRR_kernel.f90:837-849; MLatomF's summation index kk = nn - m)
//   phi = sum_m mat_d[m] y^m, m = 0..nn-1   (dKdXid, :1205-1219 = mod_MaternConsts(:,1), :966-977; kk = nn - 1 - m)
void set_matern_coefficients(ValuesArgs &a, int nn, double sigma) {
  auto fact = [](int n) {
    double f = 1.0;
    for (int i = 1; i <= n; i++) f *= i;
    return f;
  };
  for (int m = 0; m < 12; m++) a.mat_c[m] = a.mat_d[m] = 0.0;
  for (int m = 0; m <= nn; m++) {
    const int kk = nn - m;
    a.mat_c[m] = fact(nn) / fact(2 * nn) * (fact(nn + kk) / fact(kk) / fact(nn - kk));
  }
  for (int m = 0; m < nn; m++) {
    const int kk = nn - 1 - m;
    a.mat_d[m] = fact(nn) / fact(2 * nn) / (sigma * sigma) * 2.0 * (fact(nn + kk - 1) / fact(kk) / fact(nn - kk) * (nn - kk));
  }
}

// E[i] = prior + sum_j K[i][j] alpha[j] (Yest_KRR, A_KRR_kernel.f90:225-245, from a materialised kernel block): one warp
// per row, lanes stride the row, fixed-order shuffle tree -- run-to-run deterministic.
__global__ void __launch_bounds__(256) block_dot_kernel(int64_t na, int64_t nb, const double *__restrict__ K, int64_t ldk,
                                                        const double *__restrict__ alpha, double prior,
                                                        double *__restrict__ E) {
  const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= na) return;
  const int lane = threadIdx.x & 31;
  const double *row = K + i * ldk;
  double s = 0.0;
  for (int64_t j = lane; j < nb; j += 32) s = fma(row[j], alpha[j], s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) E[i] = prior + s;
}

// out[i * ldo] = sum over all chunks and columns of a chunk-major array W[ch][i][wkc]: the energies of the radial models
// (E_i = sum_j alpha_j k_ij) without a second GEMM whose single useful column would waste 63/64 of a tile.  One warp per
// row, chunks in order, fixed shuffle tree: deterministic.
__global__ void __launch_bounds__(256) chunk_rowsum_kernel(int64_t n, int nch, int wkc, const double *__restrict__ W,
                                                           double *__restrict__ out, int ldo) {
  const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= n) return;
  const int lane = threadIdx.x & 31;
  double s = 0.0;
  for (int ch = 0; ch < nch; ch++) {
    const double *row = W + ((int64_t)ch * n + i) * wkc;
    for (int jj = lane; jj < wkc; jj += 32) s += row[jj];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[i * ldo] = s;
}

int launch_chunk_rowsum(int64_t n, int nch, int wkc, const double *W, double *out, int ldo, cudaStream_t st) {
  if (n <= 0) return KREG_OK;
  chunk_rowsum_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(n, nch, wkc, W, out, ldo);
  KREG_LAUNCH_CHECK();
  return KREG_OK;
}

struct ExWs {
  int XS, nchD;
  size_t off_center, off_XA, off_XAch, off_bA, off_bAc, off_XB, off_XBch, off_bB, off_bBc, total;
};

static ExWs ex_ws(int nat, int64_t na, int64_t nb) {
  ExWs w;
  const int D = descr_size(nat), KS = (D + 3) / 4;
  w.nchD = (KS + kGemmKC - 1) / kGemmKC;
  w.XS = w.nchD * 4 * kGemmKC;
  if (w.XS % 8 == 0) w.XS += 4;
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
  w.off_bAc = take((size_t)na + 2);
  w.off_XB = take((size_t)nb * w.XS);
  w.off_XBch = take((size_t)w.nchD * nb * 4 * kGemmKC);
  w.off_bB = take((size_t)nb + 2);
  w.off_bBc = take((size_t)nb + 2);
  w.total = off;
  return w;
}

}  // namespace kreg

extern "C" size_t kreg_kernel_block_ex_workspace_bytes(int natoms, int64_t na, int64_t nb) {
  if (natoms < KREG_MIN_NATOMS || natoms > KREG_MAX_NATOMS || na < 0 || nb < 0) return 0;
  return ex_ws(natoms, na, nb).total;
}

extern "C" int kreg_kernel_block_ex(int kind, int nn, int natoms, int64_t na, const double *xyz_a, int64_t nb,
                                    const double *xyz_b, const double *xeq, double sigma, int symmetric, double lambda,
                                    double *Kout, int64_t ldk, void *workspace, size_t workspace_bytes, void *stream) {
  if (natoms < KREG_MIN_NATOMS || natoms > KREG_MAX_NATOMS) return KREG_E_UNSUPPORTED;
  if (kind < KREG_KERNEL_GAUSSIAN || kind > KREG_KERNEL_LAPLACIAN) return KREG_E_BADARG;
  if (kind == KREG_KERNEL_MATERN && (nn < 0 || nn > 10)) return KREG_E_UNSUPPORTED;
  if (na < 0 || nb < 0 || !xeq || !(sigma > 0.0) || ldk < nb) return KREG_E_BADARG;
  if (symmetric && na != nb) return KREG_E_BADARG;
  if (na == 0 || nb == 0) return KREG_OK;
  if (!xyz_a || !xyz_b || !Kout || !workspace) return KREG_E_BADARG;
  const ExWs w = ex_ws(natoms, na, nb);
  if (workspace_bytes < w.total) return KREG_E_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  char *ws = static_cast<char *>(workspace);
  auto at = [&](size_t off) { return reinterpret_cast<double *>(ws + off); };
  double *center = at(w.off_center);
  const double c1 = kExpScale / (sigma * sigma);
  int rc = launch_descriptor_mean(natoms, nb, xyz_b, xeq, center, st);  // the centre comes from set B (the training side)
  if (rc) return rc;
  rc = launch_prep_rows(natoms, na, xyz_a, xeq, center, w.XS, c1, at(w.off_XA), at(w.off_bA), at(w.off_bAc), st);
  if (rc) return rc;
  rc = launch_prep_rows(natoms, nb, xyz_b, xeq, center, w.XS, c1, at(w.off_XB), at(w.off_bB), at(w.off_bBc), st);
  if (rc) return rc;
  if (kind == KREG_KERNEL_LAPLACIAN) {
    dim3 grid((unsigned)((nb + 31) / 32), (unsigned)((na + 31) / 32));
    laplacian_block_kernel<<<grid, 256, 0, st>>>(descr_size(natoms), na, at(w.off_XA), nb, at(w.off_XB), w.XS, 1.0 / sigma,
                                                 symmetric, lambda, Kout, ldk);
    KREG_LAUNCH_CHECK();
    return KREG_OK;
  }
  rc = launch_chunk_rows(na, w.XS, kGemmKC, w.nchD, 4 * kGemmKC, at(w.off_XA), at(w.off_XAch), st);
  if (rc) return rc;
  rc = launch_chunk_rows(nb, w.XS, kGemmKC, w.nchD, 4 * kGemmKC, at(w.off_XB), at(w.off_XBch), st);
  if (rc) return rc;
  ValuesArgs a = {};
  a.XA = at(w.off_XA);
  a.XB = at(w.off_XB);
  a.XAch = at(w.off_XAch);
  a.XBch = at(w.off_XBch);
  a.XSC = 4 * kGemmKC;
  a.XS = w.XS;
  a.na = na;
  a.nb = nb;
  a.strideA = na;
  a.strideB = nb;
  a.nchunks = w.nchD;
  a.c1 = c1;
  a.lambda = lambda;
  a.all_cols = 1;
  a.row_begin = 0;
  a.row_end = na;
  a.K = Kout;
  a.ldk = ldk;
  a.biasB = at(w.off_bBc);
  if (kind == KREG_KERNEL_GAUSSIAN) {
    a.biasA = at(w.off_bA);
    rc = launch_gemm_epi(a, EPI_K, st);
    if (rc || !symmetric) return rc;
    // exact 1 + lambda on the diagonal of a symmetric block (the tile kernel's own diagonal logic belongs to the K build)
    return launch_set_diagonal(na, Kout, ldk, 1.0 + lambda, st);
  }
  a.symmetric = 0;  // every tile is computed; the diagonal is set in the epilogue through the flag below
  a.biasA = at(w.off_bAc);
  a.kind = kind;
  a.nn = nn;
  a.inv_sigma = 1.0 / sigma;
  if (kind == KREG_KERNEL_MATERN) set_matern_coefficients(a, nn, sigma);
  rc = launch_gemm_epi(a, EPI_KR, st);
  if (rc || !symmetric) return rc;
  return launch_set_diagonal(na, Kout, ldk, 1.0 + lambda, st);
}

extern "C" int kreg_block_dot(int64_t na, int64_t nb, const double *K, int64_t ldk, const double *alpha, double prior,
                              double *E, void *stream) {
  if (na < 0 || nb < 0 || ldk < nb) return KREG_E_BADARG;
  if (na == 0) return KREG_OK;
  if (!K || !alpha || !E) return KREG_E_BADARG;
  block_dot_kernel<<<(unsigned)((na + 7) / 8), 256, 0, as_stream(stream)>>>(na, nb, K, ldk, alpha, prior, E);
  KREG_LAUNCH_CHECK();
  return KREG_OK;
}

extern "C" size_t kreg_build_workspace_bytes_rows(int natoms, int64_t ntrain, int64_t ntrgr, int64_t row_begin,
                                                  int64_t row_end) {
  if (natoms < KREG_MIN_NATOMS || natoms > KREG_MAX_NATOMS || ntrain < 0 || row_begin < 0 || row_end < row_begin) return 0;
  if (ntrgr <= 0) return build_ws(natoms, ntrain, false).total;
  // the value block, when present, always has ntrain rows (kreg_build_kernel_matrix: ntrval in {0, ntrain}); a row range
  // is interpreted for both layouts and the larger requirement returned, so the caller need not pass ntrval
  int64_t i0, i1, j0, j1;
  grad_geom_range(natoms, ntrain, ntrain, row_begin, row_end, &i0, &i1);
  grad_geom_range(natoms, ntrain, 0, row_begin, row_end, &j0, &j1);
  const int64_t rows = (i1 - i0) > (j1 - j0) ? (i1 - i0) : (j1 - j0);
  return build_ws(natoms, ntrain, true, rows).total;
}

extern "C" int kreg_build_kernel_matrix(int natoms, int64_t ntrain, int64_t ntrval, int64_t ntrgr,
                                        const double *xyz_train, const double *xeq, double sigma, double lambdav,
                                        double lambdagradxyz, int64_t row_begin, int64_t row_end, int fill, double *K,
                                        int64_t ldk, void *workspace, size_t workspace_bytes, void *stream) {
  if (natoms < KREG_MIN_NATOMS || natoms > KREG_MAX_NATOMS) return KREG_E_UNSUPPORTED;
  const int64_t G3 = 3 * (int64_t)natoms;
  if (ntrain <= 0 || !xyz_train || !xeq || !K || !workspace || !(sigma > 0.0)) return KREG_E_BADARG;
  if (!(ntrval == 0 || ntrval == ntrain) || !(ntrgr == 0 || ntrgr == G3 * ntrain) || ntrval + ntrgr == 0)
    return KREG_E_BADARG;
  const int64_t M = ntrval + ntrgr;
  if (row_begin < 0 || row_end > M || row_begin > row_end) return KREG_E_BADARG;
  if (fill != KREG_FILL_LOWER && fill != KREG_FILL_FULL) return KREG_E_BADARG;
  {
    // FILL_LOWER writes columns j <= i only (whole 3 Nat-wide diagonal blocks in the gradient part), so a row-slab buffer
    // may have a row stride shorter than M: at least the defined length of its last row
    int64_t need = M;
    if (fill == KREG_FILL_LOWER && row_end < M)
      need = row_end <= ntrval ? row_end : ntrval + (row_end - ntrval + G3 - 1) / G3 * G3;
    if (ldk < need) return KREG_E_BADARG;
  }
  // gradient models: the mirrored (FULL) layout is only produced for the whole matrix
  if (fill == KREG_FILL_FULL && ntrgr > 0 && !(row_begin == 0 && row_end == M)) return KREG_E_BADARG;
  int64_t gi0 = 0, gi1 = 0;
  grad_geom_range(natoms, ntrain, ntrval, row_begin, row_end, &gi0, &gi1);
  const BuildWs w = build_ws(natoms, ntrain, ntrgr > 0, gi1 - gi0);
  if (workspace_bytes < w.total) return KREG_E_WORKSPACE;
  if (row_begin == row_end) return KREG_OK;
  cudaStream_t st = as_stream(stream);
  char *ws = static_cast<char *>(workspace);
  double *center = reinterpret_cast<double *>(ws + w.off_center);
  double *Xc = reinterpret_cast<double *>(ws + w.off_X);
  double *bias = reinterpret_cast<double *>(ws + w.off_bias);
  double *biasc = reinterpret_cast<double *>(ws + w.off_biasc);
  const double inv_s2 = 1.0 / (sigma * sigma);
  const double c1 = kExpScale * inv_s2;

  int rc = launch_descriptor_mean(natoms, ntrain, xyz_train, xeq, center, st);
  if (rc) return rc;
  rc = launch_prep_rows(natoms, ntrain, xyz_train, xeq, center, w.XS, c1, Xc, bias, biasc, st);
  if (rc) return rc;
  double *Xch = reinterpret_cast<double *>(ws + w.off_Xch);
  if (w.nchunks > 1 && ntrval > 0 && row_begin < ntrval) {
    chunk_rows_kernel<<<1184, 256, 0, st>>>(ntrain, w.XS, w.KC, w.nchunks, w.XSC, Xc, Xch);
    KREG_LAUNCH_CHECK();
  }
  if (ntrval > 0 && row_begin < ntrval) {
    ValuesArgs a;
    a.XA = a.XB = Xc;
    a.XAch = a.XBch = w.nchunks > 1 ? Xch : nullptr;
    a.XSC = w.XSC;
    a.biasA = bias;
    a.biasB = biasc;
    a.na = a.nb = ntrain;
    a.XS = w.XS;
    a.nchunks = w.nchunks;
    a.c1 = c1;
    a.lambda = lambdav;
    a.symmetric = 1;
    const bool whole = row_begin == 0 && row_end >= ntrval;
    a.all_cols = (fill == KREG_FILL_FULL && !whole) ? 1 : 0;
    a.mirror = (fill == KREG_FILL_FULL && whole) ? 1 : 0;
    a.row_begin = row_begin;
    a.row_end = row_end < ntrval ? row_end : ntrval;
    a.K = K;
    a.ldk = ldk;
    rc = launch_values(a, w.KC, st);
    if (rc) return rc;
  }
  if (ntrgr > 0 && row_end > ntrval) {
    double *E = reinterpret_cast<double *>(ws + w.off_E);
    {
      const int64_t total = ntrain * natoms * natoms;
      int blocks = (int)((total + 127) / 128);
      jacobian_rows_kernel<<<blocks, 128, 0, st>>>(natoms, ntrain, xyz_train, xeq, E, (natoms * natoms * 3 + 1) & ~1);
      KREG_LAUNCH_CHECK();
    }
    GradArgs g;
    g.xyz = xyz_train;
    g.Xc = Xc;
    g.bias = bias;
    g.E = E;
    g.ntr = ntrain;
    g.nv = ntrval;
    g.XS = w.XS;
    g.c1 = c1;
    g.inv_s2 = inv_s2;
    g.lambdag = lambdagradxyz;
    g.fill = fill;
    g.row_begin = row_begin;
    g.row_end = row_end;
    g.K = K;
    g.ldk = ldk;
    g.P = reinterpret_cast<double *>(ws + w.off_P);
    g.p_i0 = gi0;
    g.p_i1 = gi1;
    g.c_i0 = gi0;
    g.c_i1 = gi1;
    g.ES = (natoms * natoms * 3 + 1) & ~1;
    rc = launch_grad_stages(natoms, g, st);
    if (rc) return rc;
  }
  return KREG_OK;
}

extern "C" int kreg_kernel_block(int natoms, int64_t na, const double *xyz_a, int64_t nb, const double *xyz_b,
                                 const double *xeq, double sigma, double *Kout, int64_t ldk, void *workspace,
                                 size_t workspace_bytes, void *stream) {
  if (natoms < KREG_MIN_NATOMS || natoms > KREG_MAX_NATOMS) return KREG_E_UNSUPPORTED;
  if (na < 0 || nb < 0 || !xeq || !(sigma > 0.0) || ldk < nb) return KREG_E_BADARG;
  if (na == 0 || nb == 0) return KREG_OK;
  if (!xyz_a || !xyz_b || !Kout || !workspace) return KREG_E_BADARG;
  // workspace = build_ws(A) followed by build_ws(B); the centre is taken from set B (the training side)
  const BuildWs wa = build_ws(natoms, na, false), wb = build_ws(natoms, nb, false);
  if (workspace_bytes < wa.total + wb.total) return KREG_E_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  char *ws = static_cast<char *>(workspace);
  double *center = reinterpret_cast<double *>(ws + wa.total + wb.off_center);
  double *XA = reinterpret_cast<double *>(ws + wa.off_X), *bA = reinterpret_cast<double *>(ws + wa.off_bias);
  double *XB = reinterpret_cast<double *>(ws + wa.total + wb.off_X);
  double *bB = reinterpret_cast<double *>(ws + wa.total + wb.off_biasc);
  double *bBrow = reinterpret_cast<double *>(ws + wa.total + wb.off_bias);
  const double c1 = kExpScale / (sigma * sigma);
  int rc = launch_descriptor_mean(natoms, nb, xyz_b, xeq, center, st);
  if (rc) return rc;
  rc = launch_prep_rows(natoms, na, xyz_a, xeq, center, wa.XS, c1, XA, bA, nullptr, st);
  if (rc) return rc;
  rc = launch_prep_rows(natoms, nb, xyz_b, xeq, center, wb.XS, c1, XB, bBrow, bB, st);
  if (rc) return rc;
  double *XAch = reinterpret_cast<double *>(ws + wa.off_Xch);
  double *XBch = reinterpret_cast<double *>(ws + wa.total + wb.off_Xch);
  if (wa.nchunks > 1) {
    chunk_rows_kernel<<<1184, 256, 0, st>>>(na, wa.XS, wa.KC, wa.nchunks, wa.XSC, XA, XAch);
    KREG_LAUNCH_CHECK();
    chunk_rows_kernel<<<1184, 256, 0, st>>>(nb, wb.XS, wb.KC, wb.nchunks, wb.XSC, XB, XBch);
    KREG_LAUNCH_CHECK();
  }
  ValuesArgs a;
  a.XA = XA;
  a.XB = XB;
  a.XAch = wa.nchunks > 1 ? XAch : nullptr;
  a.XBch = wa.nchunks > 1 ? XBch : nullptr;
  a.XSC = wa.XSC;
  a.biasA = bA;
  a.biasB = bB;
  a.na = na;
  a.nb = nb;
  a.XS = wa.XS;
  a.nchunks = wa.nchunks;
  a.c1 = c1;
  a.lambda = 0.0;
  a.symmetric = 0;
  a.all_cols = 1;
  a.mirror = 0;
  a.row_begin = 0;
  a.row_end = na;
  a.K = Kout;
  a.ldk = ldk;
  return launch_values(a, wa.KC, st);
}
