// kreg_predict.cu -- fused KREG energy + analytic-gradient prediction on FP64 tensor cores.
//
// Replaces KREG.kreg.predict (mlatom/fortran/KREG.f90:534-596 with Yest_KRR :348-366 and
// YgradXYZest_KRR :368-391) and MLatomF's calcEst_KRR (A_KRR.f90:985-1029; "Optimized Code"
// A_KRR_kernel.f90:327-412).  Algebra, with mean-centred descriptors X':
//     k_ij = exp(-|X_i - X_j|^2 / 2s^2),  |X_i - X_j|^2 = n_i + n_j - 2 X'_i.X'_j
//     c_ij = alpha_j + (X_i - X_j).v_j / s^2          (v_j = J_j alpha_grad,j; 0 for value models)
//     E_i  = prior + sum_j k_ij c_ij
//     g_i  = [ sum_j w_ij X'_j - X'_i sum_j w_ij + sum_j k_ij v_j ] / s^2,   w_ij = k_ij c_ij
//     dE_i/dM_i = J_i^T g_i
// which is "attention without softmax": S = Xtest.Xtrain^T (GEMM1), an elementwise exp, and
// A += W.[Xtrain | 1] (GEMM2; the ones column accumulates sum_j w_ij = E_i for free).  Both GEMMs run
// on DMMA.8x8x4; a warp keeps its test-row fragments and the A accumulators in registers for the
// whole pass over the training set, and the training set arrives pre-swizzled in fragment order
// through a TMA bulk-copy ring in shared memory.  Energies and gradients come out of ONE pass (the
// reference makes two and recomputes k_ij 3*Nat+1 times).
#include <stdlib.h>

#include "kreg_common.cuh"
#include "kreg_internal.h"

namespace kreg {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;

// ------------------------------------------------------------------------------------------
// v_j = J_j alpha_grad,j  (J from KREG.f90:121,146)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) precompute_v_kernel(int nat, int64_t ntrain, const double *__restrict__ xyz,
                                                           const double *__restrict__ xeq,
                                                           const double *__restrict__ alpha_g,
                                                           double *__restrict__ V) {
  __shared__ unsigned short tab[descr_size(KREG_MAX_NATOMS)];
  for (int a = threadIdx.x; a < nat; a += blockDim.x)
    for (int b = a + 1; b < nat; b++) tab[pair_index(nat, a, b)] = (unsigned short)(a | (b << 8));
  __syncthreads();
  const int D = descr_size(nat);
  const int64_t total = ntrain * D;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t j = e / D;
    int d = (int)(e - j * D);
    int a = tab[d] & 0xff, c = tab[d] >> 8;
    const double *m = xyz + j * nat * 3;
    const double *ag = alpha_g + j * nat * 3;
    double r2 = rij2_exact(m + 3 * a, m + 3 * c);
    double x = xeq[d] / sqrt(r2);
    double s = 0.0;
#pragma unroll
    for (int t = 0; t < 3; t++) s += (m[3 * c + t] - m[3 * a + t]) * (ag[3 * a + t] - ag[3 * c + t]);
    V[e] = x * s / r2;
  }
}

// ------------------------------------------------------------------------------------------
// Model packing: one warp per group of 8 training points writes the group's chunk
//   [B1x | (B1v) | B2x | (B2v) | scal]
//   B1x[p][lane][e]  = X'[8g + lane/4][4(2p+e) + lane%4]              GEMM1 B fragments, k-step 2p+e
//   B1v              = same for v_j / s^2
//   B2x[nt][lane][s] = Xaug[8g + 2(lane%4) + s][8nt + lane/4]         GEMM2 B fragments (Xaug[.][D] = 1)
//   B2v              = same for v_j (no ones column)
//   scal[q][4]       = { a'(2q), b(2q), a'(2q+1), b(2q+1) },  a'_j = alpha_j - X'_j.v_j/s^2,
//                      b_j = -n_j/(2 s^2) * 64/ln2
// The k index of GEMM2 is the training point; fragment s of a lane pairs W's own accumulator element
// s (column 2q+s) with training point 2q+s, so W never leaves the registers it was computed in.
// ------------------------------------------------------------------------------------------
template <bool HAS_V>
__global__ void __launch_bounds__(32) pack_model_kernel(int nat, int64_t ntrain, const double *__restrict__ xyz,
                                                        const double *__restrict__ xeq,
                                                        const double *__restrict__ center,
                                                        const double *__restrict__ alpha,
                                                        const double *__restrict__ V, double c1, double inv_s2,
                                                        double *__restrict__ packed) {
  extern __shared__ double sm[];
  const TcShape sh(nat);
  const int TS = sh.NT * 8 > sh.KS * 4 ? sh.NT * 8 : sh.KS * 4;
  double *Xs = sm;            // [8][TS]
  double *Vs = sm + 8 * TS;   // [8][TS]
  double *nj = Vs + 8 * TS;   // [8]
  double *xv = nj + 8;        // [8]
  const int lane = threadIdx.x, g = lane >> 2, q = lane & 3;
  const int64_t grp = blockIdx.x;

  for (int i = lane; i < 16 * TS; i += 32) sm[i] = 0.0;
  __syncwarp();
  for (int a = 0; a < nat; a++)
    for (int b = a + 1; b < nat; b++) {
      int d = pair_index(nat, a, b);
      if (lane < 8) {
        int64_t j = grp * 8 + lane;
        if (j < ntrain) {
          const double *m = xyz + j * nat * 3;
          Xs[lane * TS + d] = xeq[d] / sqrt(rij2_exact(m + 3 * a, m + 3 * b)) - center[d];
          if (HAS_V) Vs[lane * TS + d] = V[j * sh.D + d];
        }
      }
    }
  __syncwarp();
  if (lane < 8) {
    double n = 0.0, s = 0.0;
    for (int d = 0; d < sh.D; d++) {
      n += Xs[lane * TS + d] * Xs[lane * TS + d];
      if (HAS_V) s += Xs[lane * TS + d] * Vs[lane * TS + d];
    }
    nj[lane] = n;
    xv[lane] = s;
  }
  __syncwarp();

  double *out = packed + grp * sh.chunk_doubles(HAS_V);
  // GEMM1 fragments
  for (int p = 0; p < sh.KS2; p++)
    for (int e = 0; e < 2; e++) {
      int ks = 2 * p + e;
      double vx = 0.0, vv = 0.0;
      if (ks < sh.KS) {
        vx = Xs[g * TS + 4 * ks + q];
        if (HAS_V) vv = Vs[g * TS + 4 * ks + q] * inv_s2;
      }
      out[(p * 32 + lane) * 2 + e] = vx;
      if (HAS_V) out[64 * sh.KS2 + (p * 32 + lane) * 2 + e] = vv;
    }
  // GEMM2 fragments
  double *o2x = out + 64 * sh.KS2 * (HAS_V ? 2 : 1);
  double *o2v = o2x + 64 * sh.NT;
  for (int nt = 0; nt < sh.NT; nt++)
    for (int s = 0; s < 2; s++) {
      int jl = 2 * q + s, col = 8 * nt + g;
      double vx = col < sh.D ? Xs[jl * TS + col] : (col == sh.D ? 1.0 : 0.0);
      o2x[(nt * 32 + lane) * 2 + s] = vx;
      if (HAS_V) o2v[(nt * 32 + lane) * 2 + s] = col < sh.D ? Vs[jl * TS + col] : 0.0;
    }
  double *osc = o2x + 64 * sh.NT * (HAS_V ? 2 : 1);
  if (lane < 8) {
    // scal[q][4] = { x(2q), s(2q), x(2q+1), s(2q+1) }:  s_j initialises the GEMM1 accumulator so that
    // c1 * (s_j + X'_i.X'_j) + bias_i is the scaled exponent (no per-element add in the kernel).
    //   gradient models: x_j = a'_j = alpha_j - X'_j.v_j/s^2,  s_j = -n_j/2
    //   value models   : x_j = +0.0 / -0.0 (sign of alpha_j),  s_j = -n_j/2 + s^2 ln|alpha_j|  -> w_ij comes out of
    //                    the exponential already multiplied by |alpha_j|; alpha_j = 0 (and padding) -> s_j = -1e300
    int64_t j = grp * 8 + lane;
    double a = (alpha && j < ntrain) ? alpha[j] : 0.0;
    double x, sj = -0.5 * nj[lane];
    if (HAS_V) {
      x = (j < ntrain) ? a - xv[lane] * inv_s2 : 0.0;
    } else {
      x = (a < 0.0) ? -0.0 : 0.0;
      sj = (a != 0.0) ? sj + log(fabs(a)) / inv_s2 : -1e300;
    }
    osc[(lane >> 1) * 4 + (lane & 1) * 2 + 0] = x;
    osc[(lane >> 1) * 4 + (lane & 1) * 2 + 1] = sj;
  }
}

// ------------------------------------------------------------------------------------------
// The prediction kernel
// ------------------------------------------------------------------------------------------
struct PredictArgs {
  const double *xyz;
  int64_t npoints;
  const double *xeq;
  const double *center;
  const double *packed;
  int ngroups;
  double c1;      // 64/ln2 / sigma^2
  double inv_s2;  // 1 / sigma^2
  double prior;
  double *E;
  double *G;
  // small-batch mode: the training set is split over gridDim.y slices of `stages_per_slice` ring stages; CTAs write
  // raw accumulators to `partial` and predict_finish_kernel reduces them in slice order (deterministic)
  int nslices;
  int stages_per_slice;
  double *partial;
};

template <int NAT, int MT, bool HAS_V, bool A_IN_SMEM, int GPS, int NSTAGE>
struct PredictCfg {
  static constexpr TcShape sh = TcShape(NAT);
  static constexpr int D = sh.D, KS = sh.KS, KS2 = sh.KS2, NT = sh.NT;
  static constexpr int ROWS = kWarps * MT * 8;
  static constexpr int TS = (NT * 8 > KS * 4 ? NT * 8 : KS * 4) + (A_IN_SMEM ? 4 : 0);  // +4: conflict-free A loads
  static constexpr int CHUNK = sh.chunk_doubles(HAS_V);
  static constexpr int STAGE = CHUNK * GPS;
  static constexpr int OFF_B1V = 64 * KS2;
  static constexpr int OFF_B2X = 64 * KS2 * (HAS_V ? 2 : 1);
  static constexpr int OFF_B2V = OFF_B2X + 64 * NT;
  static constexpr int OFF_SCAL = OFF_B2X + 64 * NT * (HAS_V ? 2 : 1);
  // shared memory carve-up (in doubles)
  static constexpr int SM_RING = 0;
  static constexpr int SM_TILE = SM_RING + NSTAGE * STAGE;
  static constexpr int SM_BIAS = SM_TILE + ROWS * TS;
  static constexpr int SM_TAB = SM_BIAS + ROWS;
  static constexpr int SM_BAR = SM_TAB + 64;
  static constexpr int SM_DOUBLES = SM_BAR + 2 * NSTAGE;
  static constexpr size_t SMEM_BYTES = (size_t)SM_DOUBLES * 8;
};

// Per-geometry epilogue: E = prior + A[D];  g = (A[:D] - X'_i A[D]) / s^2;  dE/dM = J_i^T g  (KREG.f90:121 Jacobian)
template <int NAT, bool WANT_G>
__device__ __forceinline__ void predict_row_epilogue(const PredictArgs &args, const double *y, int64_t p) {
  constexpr int D = NAT * (NAT - 1) / 2;
  const double esum = y[D];
  if (args.E) args.E[p] = args.prior + esum;
  if (WANT_G && args.G) {
    const double *m = args.xyz + p * (NAT * 3);
    double mm[NAT * 3], gr[NAT * 3];
#pragma unroll
    for (int i = 0; i < NAT * 3; i++) {
      mm[i] = m[i];
      gr[i] = 0.0;
    }
#pragma unroll
    for (int a = 0; a < NAT; a++)
#pragma unroll
      for (int b = a + 1; b < NAT; b++) {
        const int d = pair_index(NAT, a, b);
        const double r2 = rij2_exact(mm + 3 * a, mm + 3 * b);
        const double x = __ldg(args.xeq + d) / sqrt(r2);
        const double xc = x - __ldg(args.center + d);
        const double gd = (y[d] - xc * esum) * args.inv_s2;
        const double f = gd * x / r2;
#pragma unroll
        for (int t = 0; t < 3; t++) {
          const double dx = mm[3 * b + t] - mm[3 * a + t];
          gr[3 * a + t] = fma(f, dx, gr[3 * a + t]);
          gr[3 * b + t] = fma(-f, dx, gr[3 * b + t]);
        }
      }
    double *out = args.G + p * (NAT * 3);
#pragma unroll
    for (int i = 0; i < NAT * 3; i++) out[i] = gr[i];
  }
}

// Small-batch mode, second kernel: sum the slices' accumulators in slice order, then the usual epilogue.
template <int NAT, int ROWS, int TS>
__global__ void __launch_bounds__(kThreads) predict_finish_kernel(const PredictArgs args) {
  extern __shared__ __align__(16) double ftile[];  // [ROWS][TS]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t row0 = (int64_t)blockIdx.x * ROWS;
  const size_t slice_stride = (size_t)gridDim.x * ROWS * TS;
  for (int r = warp; r < ROWS; r += kWarps) {
    if (row0 + r >= args.npoints) break;
    const double *src = args.partial + ((size_t)blockIdx.x * ROWS + r) * TS;
    for (int c = lane; c < TS; c += 32) {
      double s = 0.0;
      for (int sl = 0; sl < args.nslices; sl++) s += src[(size_t)sl * slice_stride + c];
      ftile[r * TS + c] = s;
    }
  }
  __syncthreads();
  for (int r = tid; r < ROWS; r += kThreads) {
    const int64_t p = row0 + r;
    if (p < args.npoints) predict_row_epilogue<NAT, true>(args, ftile + r * TS, p);
  }
}

template <int NAT, int MT, bool HAS_V, bool A_IN_SMEM, int GPS, int NSTAGE, bool WANT_G>
__global__ void __launch_bounds__(kThreads, (MT <= 2 && !A_IN_SMEM && NAT <= 9) ? 2 : 1) predict_kernel(const PredictArgs args) {
  using C = PredictCfg<NAT, MT, HAS_V, A_IN_SMEM, GPS, NSTAGE>;
  constexpr int D = C::D, KS = C::KS, KS2 = C::KS2, NT = C::NT, TS = C::TS;
  extern __shared__ __align__(128) double smem[];
  double *ring = smem + C::SM_RING;
  double *tile = smem + C::SM_TILE;
  double *bias = smem + C::SM_BIAS;
  double *tab = smem + C::SM_TAB;
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + C::SM_BAR);
  uint64_t *empty_bar = full_bar + NSTAGE;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int64_t row0 = (int64_t)blockIdx.x * C::ROWS;
  const int all_stages = (args.ngroups + GPS - 1) / GPS;
  const int stage0 = args.nslices > 1 ? (int)blockIdx.y * args.stages_per_slice : 0;  // first ring stage of this slice
  const int total_stages = args.nslices > 1 ? min(args.stages_per_slice, all_stages - stage0) : all_stages;

  // ---- barrier init + first TMA bulk copies (overlap with the descriptor prologue) ----
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; s++) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kWarps);
    }
    mbar_fence_init();
  }
  fill_exp_table(tab, tid, kThreads);
  __syncthreads();
  if (tid == 0) {
    for (int s = 0; s < NSTAGE && s < total_stages; s++) {
      int ng = min(GPS, args.ngroups - (stage0 + s) * GPS);
      unsigned bytes = (unsigned)(ng * C::CHUNK * 8);
      mbar_arrive_expect_tx(&full_bar[s], bytes);
      bulk_g2s(ring + s * C::STAGE, args.packed + (size_t)(stage0 + s) * C::STAGE, bytes, &full_bar[s]);
    }
  }

  // ---- prologue: centred descriptors of this CTA's test rows -> shared tile ----
  for (int r = tid; r < C::ROWS; r += kThreads) {
    int64_t p = row0 + r;
    if (p >= args.npoints) p = args.npoints - 1;  // duplicate a valid row; its results are not stored
    const double *m = args.xyz + p * (NAT * 3);
    double mm[NAT * 3];
#pragma unroll
    for (int i = 0; i < NAT * 3; i++) mm[i] = m[i];
    double n = 0.0;
#pragma unroll
    for (int a = 0; a < NAT; a++)
#pragma unroll
      for (int b = a + 1; b < NAT; b++) {
        constexpr int dummy = 0;
        (void)dummy;
        const int d = pair_index(NAT, a, b);
        double x = __ldg(args.xeq + d) / sqrt(rij2_exact(mm + 3 * a, mm + 3 * b)) - __ldg(args.center + d);
        tile[r * TS + d] = x;
        n = fma(x, x, n);
      }
    for (int d = D; d < TS; d++) tile[r * TS + d] = 0.0;
    bias[r] = -0.5 * args.c1 * n;
  }
  __syncthreads();

  // ---- per-warp register state ----
  double afrag[A_IN_SMEM ? 1 : MT][A_IN_SMEM ? 1 : KS];
  double bi[MT];
  double acc[WANT_G ? MT : 1][WANT_G ? NT : 1][2];  // GEMM2 accumulators (energy + gradient passes)
  double esum[WANT_G ? 1 : MT];                      // energy-only passes: per-thread partial row sums of W
  const double *arow[MT];
#pragma unroll
  for (int mt = 0; mt < MT; mt++) {
    const int r = (warp * MT + mt) * 8 + g;
    arow[mt] = tile + r * TS + q;
    bi[mt] = bias[r];
    if (!A_IN_SMEM) {
#pragma unroll
      for (int ks = 0; ks < KS; ks++) afrag[mt][ks] = arow[mt][4 * ks];
    }
    if (WANT_G) {
#pragma unroll
      for (int nt = 0; nt < NT; nt++) acc[WANT_G ? mt : 0][WANT_G ? nt : 0][0] = acc[WANT_G ? mt : 0][WANT_G ? nt : 0][1] = 0.0;
    } else {
      esum[WANT_G ? 0 : mt] = 0.0;
    }
  }
  const double c1 = args.c1;

  // ---- main loop over the training set ----
  for (int it = 0; it < total_stages; it++) {
    // producer duty: refill the stage every warp finished in the previous iteration
    if (tid == 0 && it >= 1 && it - 1 + NSTAGE < total_stages) {
      const int sp = (it - 1) % NSTAGE;
      const int nxt = it - 1 + NSTAGE;
      mbar_wait(&empty_bar[sp], ((it - 1) / NSTAGE) & 1);
      int ng = min(GPS, args.ngroups - (stage0 + nxt) * GPS);
      unsigned bytes = (unsigned)(ng * C::CHUNK * 8);
      mbar_arrive_expect_tx(&full_bar[sp], bytes);
      bulk_g2s(ring + sp * C::STAGE, args.packed + (size_t)(stage0 + nxt) * C::STAGE, bytes, &full_bar[sp]);
    }
    const int s = it % NSTAGE;
    mbar_wait(&full_bar[s], (it / NSTAGE) & 1);
    const int ng = min(GPS, args.ngroups - (stage0 + it) * GPS);
    for (int gi = 0; gi < ng; gi++) {
      const double *cb = ring + s * C::STAGE + gi * C::CHUNK;
      // GEMM1: S = s_j + Xtest . Xtrain^T  (and S2 = Xtest . (V/s^2)^T); the accumulator starts from the column term
      const double4 sc = *reinterpret_cast<const double4 *>(cb + C::OFF_SCAL + q * 4);
      double S[MT][2], S2[HAS_V ? MT : 1][2];
#pragma unroll
      for (int mt = 0; mt < MT; mt++) {
        S[mt][0] = sc.y;
        S[mt][1] = sc.w;
        if (HAS_V) S2[mt][0] = S2[mt][1] = 0.0;
      }
#pragma unroll
      for (int p = 0; p < KS2; p++) {
        const double2 bx = *reinterpret_cast<const double2 *>(cb + (p * 32 + lane) * 2);
        double2 bv = make_double2(0.0, 0.0);
        if (HAS_V) bv = *reinterpret_cast<const double2 *>(cb + C::OFF_B1V + (p * 32 + lane) * 2);
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int ks = 2 * p + e;
          if (ks < KS) {
#pragma unroll
            for (int mt = 0; mt < MT; mt++) {
              const double a = A_IN_SMEM ? arow[mt][4 * ks] : afrag[A_IN_SMEM ? 0 : mt][A_IN_SMEM ? 0 : ks];
              dmma884(S[mt], a, e ? bx.y : bx.x);
              if (HAS_V) dmma884(S2[mt], a, e ? bv.y : bv.x);
            }
          }
        }
      }
      // elementwise: k = exp(c1 S + bias_i); gradient models w = k (a'_j + S2); value models w = sign(alpha_j) k
      // (|alpha_j| is inside the exponent), so no multiply at all
      double W[MT][2], Kf[HAS_V ? MT : 1][2];
#pragma unroll
      for (int mt = 0; mt < MT; mt++) {
        const double k0 = exp_from_scaled_cvt(fma(S[mt][0], c1, bi[mt]), tab);
        const double k1 = exp_from_scaled_cvt(fma(S[mt][1], c1, bi[mt]), tab);
        if (HAS_V) {
          W[mt][0] = k0 * (sc.x + S2[mt][0]);
          W[mt][1] = k1 * (sc.z + S2[mt][1]);
          Kf[mt][0] = k0;
          Kf[mt][1] = k1;
        } else {
          W[mt][0] = __hiloint2double(__double2hiint(k0) ^ __double2hiint(sc.x), __double2loint(k0));
          W[mt][1] = __hiloint2double(__double2hiint(k1) ^ __double2hiint(sc.z), __double2loint(k1));
        }
      }
      if (!WANT_G) {
        // energy only: E_i = sum_j w_ij needs no second GEMM, just the row sums of W
#pragma unroll
        for (int mt = 0; mt < MT; mt++) esum[WANT_G ? 0 : mt] += W[mt][0] + W[mt][1];
        continue;
      }
      // GEMM2: acc += W . [Xtrain | 1]  (+ K . V)
#pragma unroll
      for (int nt = 0; nt < NT; nt++) {
        const double2 bx = *reinterpret_cast<const double2 *>(cb + C::OFF_B2X + (nt * 32 + lane) * 2);
#pragma unroll
        for (int mt = 0; mt < MT; mt++) dmma884(acc[WANT_G ? mt : 0][WANT_G ? nt : 0], W[mt][0], bx.x);
#pragma unroll
        for (int mt = 0; mt < MT; mt++) dmma884(acc[WANT_G ? mt : 0][WANT_G ? nt : 0], W[mt][1], bx.y);
        if (HAS_V) {
          const double2 bv = *reinterpret_cast<const double2 *>(cb + C::OFF_B2V + (nt * 32 + lane) * 2);
#pragma unroll
          for (int mt = 0; mt < MT; mt++) dmma884(acc[WANT_G ? mt : 0][WANT_G ? nt : 0], Kf[mt][0], bv.x);
#pragma unroll
          for (int mt = 0; mt < MT; mt++) dmma884(acc[WANT_G ? mt : 0][WANT_G ? nt : 0], Kf[mt][1], bv.y);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
  }

  // ---- epilogue: accumulators -> shared tile, then one thread per test geometry ----
  __syncthreads();
  if (WANT_G) {
#pragma unroll
    for (int mt = 0; mt < MT; mt++) {
      const int r = (warp * MT + mt) * 8 + g;
#pragma unroll
      for (int nt = 0; nt < NT; nt++)
        *reinterpret_cast<double2 *>(tile + r * TS + 8 * nt + 2 * q) =
            make_double2(acc[WANT_G ? mt : 0][WANT_G ? nt : 0][0], acc[WANT_G ? mt : 0][WANT_G ? nt : 0][1]);
    }
  } else {
    // the four lanes q = 0..3 of a row hold disjoint column subsets: fixed-order butterfly, lane q == 0 stores
#pragma unroll
    for (int mt = 0; mt < MT; mt++) {
      double s = esum[WANT_G ? 0 : mt];
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (q == 0) tile[((warp * MT + mt) * 8 + g) * TS + D] = s;
    }
  }
  __syncthreads();
  if (WANT_G && args.nslices > 1) {
    // small-batch mode: hand the raw accumulators of the valid rows to predict_finish_kernel
    double *dst = args.partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * C::ROWS * TS;
    const int64_t left = args.npoints - row0;
    const int nvalid = (int)(left < C::ROWS ? left : C::ROWS);
    for (int e = tid; e < nvalid * TS; e += kThreads) dst[e] = tile[e];
    return;
  }
  for (int r = tid; r < C::ROWS; r += kThreads) {
    const int64_t p = row0 + r;
    if (p < args.npoints) predict_row_epilogue<NAT, WANT_G>(args, tile + r * TS, p);
  }
}

// ---- host-side dispatch -----------------------------------------------------------------------
// Small batches (fewer CTAs than half the SMs) split the training set over gridDim.y slices so the whole chip works
// on them: the single-geometry MD / geometry-optimisation step is latency-, not throughput-bound.
template <int GPS>
static void plan_slices(int ngroups, int64_t blocks, int sms, int *nslices, int *stages_per_slice) {
  const int all_stages = (ngroups + GPS - 1) / GPS;
  *nslices = 1;
  *stages_per_slice = all_stages;
  if (blocks * 2 > sms || all_stages < 8) return;
  int want = (int)((2 * sms + blocks - 1) / blocks);  // ~2 CTAs per SM in total
  int per = (all_stages + want - 1) / want;
  if (per < 4) per = 4;                               // keep the TMA ring busy
  *stages_per_slice = per;
  *nslices = (all_stages + per - 1) / per;
}

template <int NAT, int MT, bool HAS_V, bool A_IN_SMEM, int GPS, int NSTAGE, bool WANT_G>
static int launch_predict_g(PredictArgs a, cudaStream_t st) {
  using C = PredictCfg<NAT, MT, HAS_V, A_IN_SMEM, GPS, NSTAGE>;
  auto kern = predict_kernel<NAT, MT, HAS_V, A_IN_SMEM, GPS, NSTAGE, WANT_G>;
  static_assert(C::SMEM_BYTES <= 227 * 1024, "shared memory budget exceeded");
  KREG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
  const int64_t blocks = (a.npoints + C::ROWS - 1) / C::ROWS;
  dim3 grid((unsigned)blocks, (unsigned)(a.nslices > 1 ? a.nslices : 1));
  kern<<<grid, kThreads, C::SMEM_BYTES, st>>>(a);
  KREG_LAUNCH_CHECK();
  if (a.nslices > 1) {
    auto fin = predict_finish_kernel<NAT, C::ROWS, C::TS>;
    const size_t smem = (size_t)C::ROWS * C::TS * sizeof(double);
    KREG_CUDA_TRY(cudaFuncSetAttribute(fin, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fin<<<(unsigned)blocks, kThreads, smem, st>>>(a);
    KREG_LAUNCH_CHECK();
  }
  return KREG_OK;
}

// Workspace (bytes) the small-batch mode needs for this launch; 0 when the batch is large enough to run unsplit.
template <int NAT, int MT, bool HAS_V, bool A_IN_SMEM, int GPS, int NSTAGE>
static size_t predict_ws_cfg(int ngroups, int64_t npoints, int sms, int *nslices, int *stages_per_slice) {
  using C = PredictCfg<NAT, MT, HAS_V, A_IN_SMEM, GPS, NSTAGE>;
  const int64_t blocks = (npoints + C::ROWS - 1) / C::ROWS;
  plan_slices<GPS>(ngroups, blocks, sms, nslices, stages_per_slice);
  return *nslices > 1 ? (size_t)*nslices * blocks * C::ROWS * C::TS * sizeof(double) : 0;
}

template <int NAT, int MT, bool HAS_V, bool A_IN_SMEM, int GPS, int NSTAGE>
static int launch_predict_cfg(PredictArgs a, int sms, void *ws, size_t ws_bytes, bool query, size_t *need,
                              cudaStream_t st) {
  int nslices, per;
  const size_t bytes = predict_ws_cfg<NAT, MT, HAS_V, A_IN_SMEM, GPS, NSTAGE>(a.ngroups, a.npoints, sms, &nslices, &per);
  if (query) {
    *need = bytes;
    return KREG_OK;
  }
  a.nslices = 1;
  a.stages_per_slice = 0;
  a.partial = nullptr;
  if (nslices > 1 && ws && ws_bytes >= bytes) {  // without a workspace the batch simply runs unsplit
    a.nslices = nslices;
    a.stages_per_slice = per;
    a.partial = static_cast<double *>(ws);
  }
  // energy-only requests skip the second GEMM entirely (the split mode always carries the full accumulators)
  return (a.G || a.nslices > 1) ? launch_predict_g<NAT, MT, HAS_V, A_IN_SMEM, GPS, NSTAGE, true>(a, st)
                                : launch_predict_g<NAT, MT, HAS_V, A_IN_SMEM, GPS, NSTAGE, false>(a, st);
}

template <int NAT>
static int launch_predict_nat(const PredictArgs &a, bool has_v, int sms, void *ws, size_t ws_bytes, bool query,
                              size_t *need, cudaStream_t st) {
  constexpr TcShape sh(NAT);
  if constexpr (sh.D <= 36) {
    // small molecules: test-row fragments live in registers, 2 m-tiles per warp, two CTAs per SM (16 warps hide the
    // fixed DMMA issue latency better than one CTA with 4 m-tiles per warp: 51.5 vs 53.7 ms on cfg2)
    return has_v ? launch_predict_cfg<NAT, 2, true, false, 1, 3>(a, sms, ws, ws_bytes, query, need, st)
                 : launch_predict_cfg<NAT, 2, false, false, 3, 3>(a, sms, ws, ws_bytes, query, need, st);
  } else if constexpr (sh.D <= 66) {
    return has_v ? launch_predict_cfg<NAT, 1, true, false, 1, 4>(a, sms, ws, ws_bytes, query, need, st)
                 : launch_predict_cfg<NAT, 2, false, false, 2, 4>(a, sms, ws, ws_bytes, query, need, st);
  } else {
    // large molecules: accumulators fill the register file; test-row fragments stay in shared memory
    return has_v ? launch_predict_cfg<NAT, 1, true, true, 1, 2>(a, sms, ws, ws_bytes, query, need, st)
                 : launch_predict_cfg<NAT, 1, false, true, 1, 3>(a, sms, ws, ws_bytes, query, need, st);
  }
}

// test geometries handled by one CTA of the kernel launch_predict_nat<NAT> would pick
template <int NAT>
static int tile_rows_nat(bool has_v) {
  constexpr TcShape sh(NAT);
  const int mt = sh.D <= 36 ? 2 : sh.D <= 66 ? (has_v ? 1 : 2) : 1;
  return kWarps * mt * 8;
}

static int launch_predict(int nat, const PredictArgs &a, bool has_v, int sms, void *ws, size_t ws_bytes, bool query,
                          size_t *need, cudaStream_t st) {
  switch (nat) {
#define KREG_CASE(N) \
  case N:            \
    return launch_predict_nat<N>(a, has_v, sms, ws, ws_bytes, query, need, st);
    KREG_CASE(2) KREG_CASE(3) KREG_CASE(4) KREG_CASE(5) KREG_CASE(6) KREG_CASE(7) KREG_CASE(8) KREG_CASE(9)
    KREG_CASE(10) KREG_CASE(11) KREG_CASE(12) KREG_CASE(21)
#undef KREG_CASE
    default:
      return KREG_E_UNSUPPORTED;
  }
}

}  // namespace kreg

using namespace kreg;

extern "C" int kreg_precompute_v(int natoms, int64_t ntrain, const double *xyz_train, const double *xeq,
                                 const double *alpha_grad, double *V, void *stream) {
  if (natoms < KREG_MIN_NATOMS || natoms > KREG_MAX_NATOMS) return KREG_E_UNSUPPORTED;
  if (ntrain < 0 || (ntrain > 0 && (!xyz_train || !xeq || !alpha_grad || !V))) return KREG_E_BADARG;
  if (ntrain == 0) return KREG_OK;
  const int64_t total = ntrain * descr_size(natoms);
  int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  precompute_v_kernel<<<blocks, 256, 0, as_stream(stream)>>>(natoms, ntrain, xyz_train, xeq, alpha_grad, V);
  KREG_LAUNCH_CHECK();
  return KREG_OK;
}

extern "C" int kreg_predict_tile_rows(int natoms, int has_gradient_terms) {
  const bool v = has_gradient_terms != 0;
  switch (natoms) {
#define KREG_CASE(N) \
  case N:            \
    return tile_rows_nat<N>(v);
    KREG_CASE(2) KREG_CASE(3) KREG_CASE(4) KREG_CASE(5) KREG_CASE(6) KREG_CASE(7) KREG_CASE(8) KREG_CASE(9)
    KREG_CASE(10) KREG_CASE(11) KREG_CASE(12) KREG_CASE(21)
#undef KREG_CASE
    default:
      return KREG_E_UNSUPPORTED;
  }
}

extern "C" size_t kreg_packed_model_bytes(int natoms, int64_t ntrain, int has_gradient_terms) {
  if (natoms < KREG_MIN_NATOMS || natoms > KREG_MAX_NATOMS || ntrain < 0) return 0;
  const TcShape sh(natoms);
  const int64_t groups = (ntrain + 7) / 8;
  return (size_t)groups * sh.chunk_doubles(has_gradient_terms != 0) * sizeof(double);
}

extern "C" int kreg_pack_model(int natoms, int64_t ntrain, const double *xyz_train, const double *xeq,
                               const double *alpha_val, const double *V, double sigma, double *center,
                               void *packed, size_t packed_bytes, void *stream) {
  if (natoms < KREG_MIN_NATOMS || natoms > KREG_MAX_NATOMS) return KREG_E_UNSUPPORTED;
  if (ntrain <= 0 || !xyz_train || !xeq || !center || !packed || !(sigma > 0.0)) return KREG_E_BADARG;
  const bool has_v = V != nullptr;
  if (packed_bytes < kreg_packed_model_bytes(natoms, ntrain, has_v)) return KREG_E_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(packed) & 15) != 0) return KREG_E_BADARG;
  cudaStream_t st = as_stream(stream);
  int rc = launch_descriptor_mean(natoms, ntrain, xyz_train, xeq, center, st);
  if (rc) return rc;
  const TcShape sh(natoms);
  const int TS = sh.NT * 8 > sh.KS * 4 ? sh.NT * 8 : sh.KS * 4;
  const size_t smem = (size_t)(16 * TS + 16) * sizeof(double);
  const double inv_s2 = 1.0 / (sigma * sigma);
  const double c1 = kExpScale * inv_s2;
  const unsigned groups = (unsigned)((ntrain + 7) / 8);
  if (has_v)
    pack_model_kernel<true><<<groups, 32, smem, st>>>(natoms, ntrain, xyz_train, xeq, center, alpha_val, V, c1,
                                                      inv_s2, static_cast<double *>(packed));
  else
    pack_model_kernel<false><<<groups, 32, smem, st>>>(natoms, ntrain, xyz_train, xeq, center, alpha_val, nullptr,
                                                       c1, inv_s2, static_cast<double *>(packed));
  KREG_LAUNCH_CHECK();
  return KREG_OK;
}

static int device_sm_count(int *sms) {
  static thread_local int cached_dev = -1, cached_sms = 0;
  int dev = 0;
  KREG_CUDA_TRY(cudaGetDevice(&dev));
  if (dev != cached_dev) {
    KREG_CUDA_TRY(cudaDeviceGetAttribute(&cached_sms, cudaDevAttrMultiProcessorCount, dev));
    cached_dev = dev;
  }
  *sms = cached_sms;
  return KREG_OK;
}

extern "C" size_t kreg_predict_workspace_bytes(int natoms, int64_t ntrain, int has_gradient_terms, int64_t npoints) {
  if (natoms < KREG_MIN_NATOMS || natoms > KREG_MAX_NATOMS || ntrain <= 0 || npoints <= 0) return 0;
  int sms = 0;
  if (device_sm_count(&sms) != KREG_OK) return 0;
  PredictArgs a = {};
  a.npoints = npoints;
  a.ngroups = (int)((ntrain + 7) / 8);
  size_t need = 0;
  if (launch_predict(natoms, a, has_gradient_terms != 0, sms, nullptr, 0, true, &need, nullptr) != KREG_OK) return 0;
  return need;
}

extern "C" int kreg_predict(int natoms, int64_t ntrain, const void *packed, const double *center, const double *xeq,
                            double sigma, double prior, int has_gradient_terms, int64_t npoints, const double *xyz,
                            int flags, double *E, double *G, void *workspace, size_t workspace_bytes, void *stream) {
  if (natoms < KREG_MIN_NATOMS || natoms > KREG_MAX_NATOMS) return KREG_E_UNSUPPORTED;
  if (ntrain <= 0 || npoints < 0 || !packed || !center || !xeq || !(sigma > 0.0)) return KREG_E_BADARG;
  if (!(flags & (KREG_PREDICT_ENERGY | KREG_PREDICT_GRADIENT))) return KREG_E_BADARG;
  if (npoints == 0) return KREG_OK;
  if (!xyz) return KREG_E_BADARG;
  if ((flags & KREG_PREDICT_ENERGY) && !E) return KREG_E_BADARG;
  if ((flags & KREG_PREDICT_GRADIENT) && !G) return KREG_E_BADARG;
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc) return rc;
  PredictArgs a = {};
  a.xyz = xyz;
  a.npoints = npoints;
  a.xeq = xeq;
  a.center = center;
  a.packed = static_cast<const double *>(packed);
  a.ngroups = (int)((ntrain + 7) / 8);
  a.inv_s2 = 1.0 / (sigma * sigma);
  a.c1 = kExpScale * a.inv_s2;
  a.prior = prior;
  a.E = (flags & KREG_PREDICT_ENERGY) ? E : nullptr;
  a.G = (flags & KREG_PREDICT_GRADIENT) ? G : nullptr;
  size_t need = 0;
  return launch_predict(natoms, a, has_gradient_terms != 0, sms, workspace, workspace_bytes, false, &need,
                        as_stream(stream));
}
