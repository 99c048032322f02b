// kreg_predict_kernel.cuh -- the prediction kernel templates (see kreg_predict.cu for the algebra and the packed
// model layout).  Included by kreg_predict_inst.cu, which is compiled once per range of molecule sizes so that the
// per-Natoms instantiations build in parallel.
#pragma once
#include <stdlib.h>

#include "kreg_common.cuh"
#include "kreg_internal.h"
#include "kreg_predict_args.h"

namespace kreg {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;

template <int NAT, int MT, bool HAS_V, bool A_IN_SMEM, int GPS, int NSTAGE, int NW = kWarps>
struct PredictCfg {
  static constexpr TcShape sh = TcShape(NAT);
  static constexpr int D = sh.D, KS = sh.KS, KS2 = sh.KS2, NT = sh.NT;
  static constexpr int WARPS = NW, THREADS = NW * 32;
  static constexpr int ROWS = NW * MT * 8;
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
// (Radial kernels: the energy sum, sum_j alpha_j k_ij, is not the sum of the gradient weights, sum_j alpha_j phi_ij, that
// the ones column accumulates; it arrives through `esep`.)
template <int NAT, bool WANT_G>
__device__ __forceinline__ void predict_row_epilogue(const PredictArgs &args, const double *y, int64_t p,
                                                     const double *esep = nullptr) {
  constexpr int D = NAT * (NAT - 1) / 2;
  const double esum = y[D];
  if (args.E) args.E[p] = args.prior + (esep ? *esep : esum);
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

// The same epilogue with one LANE per atom (latency path: a single thread walking all D pairs is a chain of D square
// roots and divisions).  Atom a accumulates its partners in the order c = 0..Nat-1, which is the order in which the
// serial version reaches them, so both give bit-identical gradients.
template <int NAT>
__device__ __forceinline__ void predict_row_epilogue_warp(const PredictArgs &args, const double *y, int64_t p, int lane,
                                                          const double *esep = nullptr) {
  constexpr int D = NAT * (NAT - 1) / 2;
  const double esum = y[D];
  if (lane == 0 && args.E) args.E[p] = args.prior + (esep ? *esep : esum);
  if (!args.G || lane >= NAT) return;
  const double *m = args.xyz + p * (NAT * 3);
  const int a = lane;
  double gr[3] = {0.0, 0.0, 0.0};
  for (int c = 0; c < NAT; c++) {
    if (c == a) continue;
    const int lo = a < c ? a : c, hi = a < c ? c : a;
    const int d = pair_index(NAT, lo, hi);
    const double r2 = rij2_exact(m + 3 * lo, m + 3 * hi);
    const double x = __ldg(args.xeq + d) / sqrt(r2);
    const double xc = x - __ldg(args.center + d);
    const double gd = (y[d] - xc * esum) * args.inv_s2;
    const double f = gd * x / r2;
    // the serial version adds f (m_hi - m_lo) to the lower atom and -f (m_hi - m_lo) to the higher one
#pragma unroll
    for (int t = 0; t < 3; t++) {
      const double dx = m[3 * hi + t] - m[3 * lo + t];
      gr[t] = fma(a == lo ? f : -f, dx, gr[t]);
    }
  }
  double *out = args.G + p * (NAT * 3) + 3 * a;
  out[0] = gr[0];
  out[1] = gr[1];
  out[2] = gr[2];
}

// Small-batch mode, second kernel: sum the slices' accumulators, then the epilogue.  One CTA per geometry, so that a
// batch of a few dozen geometries is reduced by as many SMs; a row's TS accumulators are consecutive doubles in every
// slice.  The slice loop is split over P threads per element (thread k takes slices k, k+P, ...) and the P partial
// sums are added in the order k = 0..P-1: a fixed order for a given molecule size, and no chain of `nslices` dependent
// loads.  Epilogue: one warp, a lane per atom.
template <int NAT, int ROWS, int TS>
__global__ void __launch_bounds__(kThreads) predict_finish_kernel(const PredictArgs args) {
  __shared__ __align__(16) double ftile[TS];
  __shared__ double part[kThreads];
  __shared__ double e_sep;
  const int tid = threadIdx.x;
  const int64_t row = (int64_t)blockIdx.x * ROWS + blockIdx.y;
  if (row >= args.npoints) return;
  const size_t slice_stride = (size_t)gridDim.x * ROWS * TS;
  const double *__restrict__ src = args.partial + (size_t)row * TS;
  constexpr int P = TS >= kThreads ? 1 : (kThreads / TS > 16 ? 16 : kThreads / TS);
  if (P == 1) {
    for (int e = tid; e < TS; e += kThreads) {
      double s = 0.0;
#pragma unroll 4
      for (int sl = 0; sl < args.nslices; sl++) {
        KREG_CHECK_STORE(src + (size_t)sl * slice_stride + e, args.partial, args.partial_doubles);
        s += src[(size_t)sl * slice_stride + e];
      }
      ftile[e] = s;
    }
  } else {
    if (tid < TS * P) {
      const int e = tid % TS, k = tid / TS;
      double s = 0.0;
#pragma unroll 4
      for (int sl = k; sl < args.nslices; sl += P) {
        KREG_CHECK_STORE(src + (size_t)sl * slice_stride + e, args.partial, args.partial_doubles);
        s += src[(size_t)sl * slice_stride + e];
      }
      part[k * TS + e] = s;
    }
    __syncthreads();
    if (tid < TS) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < P; k++) s += part[k * TS + tid];
      ftile[tid] = s;
    }
  }
  if (args.radial && tid >= kThreads - 32) {
    // radial kernels: the slices' energy sums arrive beside the accumulators; lane k takes slices k, k + 32, ..., then a
    // fixed-order butterfly
    const double *es = args.partial_e + row;
    const size_t estride = (size_t)gridDim.x * ROWS;
    double s = 0.0;
    for (int sl = tid & 31; sl < args.nslices; sl += 32) {
      KREG_CHECK_STORE(es + (size_t)sl * estride, args.partial_e, args.partial_e_doubles);
      s += es[(size_t)sl * estride];
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((tid & 31) == 0) e_sep = s;
  }
  __syncthreads();
  if (tid < 32) predict_row_epilogue_warp<NAT>(args, ftile, row, tid, args.radial ? &e_sep : nullptr);
}

// SPLIT = small-batch (latency) mode: the training set is divided over gridDim.y slices and the raw accumulators go to
// predict_finish_kernel.  A separate instantiation, so that none of its code perturbs the throughput kernel.
// KIND = 1: values-trained model on one of MLatomF's radial kernel functions (Gaussian / exponential / Matern chosen at
// run time, args.kind): the elementwise step takes r_ij from the GEMM1 accumulator and forms the gradient weight
// alpha_j phi(r_ij) (phi = -k'(r)/r; dKdXid, A_KRR_kernel.f90:1178-1233) for GEMM2 and the energy term alpha_j k(r_ij),
// summed per thread like the energy-only pass does.  Everything else -- operand ring, fragments, GEMMs -- is shared.
template <int NAT, int MT, bool HAS_V, bool A_IN_SMEM, int GPS, int NSTAGE, bool WANT_G, int NW = kWarps, bool SPLIT = false,
          int KIND = 0>
__global__ void __launch_bounds__(NW * 32, (MT <= 2 && !A_IN_SMEM && NAT <= 9) ? 2 : 1) predict_kernel(const PredictArgs args) {
  static_assert(KIND == 0 || !HAS_V, "radial kernels: values-trained models");
  using C = PredictCfg<NAT, MT, HAS_V, A_IN_SMEM, GPS, NSTAGE, NW>;
  constexpr int kWarps = NW, kThreads = NW * 32;  // shadow the defaults: this kernel's own CTA shape
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
  const int stage0 = SPLIT ? (int)blockIdx.y * args.stages_per_slice : 0;  // first ring stage of this slice
  const int total_stages = SPLIT ? min(args.stages_per_slice, all_stages - stage0) : all_stages;

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
  __syncwarp();  // the issuing lane rejoins its warp before the prologue

  // ---- prologue: centred descriptors of this CTA's test rows -> shared tile ----
  const int64_t rows_left = args.npoints - row0;
  if (SPLIT && rows_left * 4 <= C::ROWS) {
    // sparse tile (latency mode: a handful of geometries): one thread per (row, atom pair) instead of a chain of D
    // square roots and divisions per thread; the norm is then summed in the same d order, so the numbers are the same
    const int nvalid = (int)rows_left;
    for (int e = tid; e < nvalid * D; e += kThreads) {
      const int r = e / D;
      int d = e - r * D, a = 0;
      while (d >= NAT - 1 - a) {
        d -= NAT - 1 - a;
        a++;
      }
      const int b = a + 1 + d;
      d = e - r * D;
      const double *m = args.xyz + (row0 + r) * (NAT * 3);
      tile[r * TS + d] = __ldg(args.xeq + d) / sqrt(rij2_exact(m + 3 * a, m + 3 * b)) - __ldg(args.center + d);
    }
    for (int e = tid; e < C::ROWS * TS; e += kThreads) {
      const int r = e / TS, d = e - r * TS;
      if (r >= nvalid || d >= D) tile[e] = 0.0;  // padding columns and unused rows (their results are not stored)
    }
    __syncthreads();
    for (int r = tid; r < C::ROWS; r += kThreads) {
      double n = 0.0;
      if (r < nvalid)
        for (int d = 0; d < D; d++) n = fma(tile[r * TS + d], tile[r * TS + d], n);
      bias[r] = -0.5 * args.c1 * n;
    }
  } else
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
  double esum[(WANT_G && KIND == 0) ? 1 : MT];       // energy-only passes (and radial kernels): per-thread partial row sums
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
    if (WANT_G && KIND != 0) esum[(WANT_G && KIND == 0) ? 0 : mt] = 0.0;
  }
  const double c1 = args.c1;

  // ---- main loop over the training set ----
  for (int it = 0; it < total_stages; it++) {
    // producer duty: refill the stage every warp finished in the previous iteration.  Warp-uniform branch with the one
    // issuing lane inside it and an explicit reconvergence after it: with a bare `if (tid == 0) { wait; issue }` the other
    // 31 lanes of warp 0 may run ahead into the tile code while lane 0 is still issuing, and the two paths share the
    // warp's uniform registers.  Costs 0.7 % on cfg2 (51.5 -> 51.9 ms).  A non-blocking variant
    // (mbarrier.test_wait, slots picked up one iteration later) was measured on the same box and dropped: equal at 9
    // atoms, 2-5 % slower at 16-24 atoms where the copy needs the whole iteration to land.
    if (warp == 0 && it >= 1 && it - 1 + NSTAGE < total_stages) {
      if (lane == 0) {
        const int sp = (it - 1) % NSTAGE;
        const int nxt = it - 1 + NSTAGE;
        mbar_wait(&empty_bar[sp], ((it - 1) / NSTAGE) & 1);
        int ng = min(GPS, args.ngroups - (stage0 + nxt) * GPS);
        unsigned bytes = (unsigned)(ng * C::CHUNK * 8);
        mbar_arrive_expect_tx(&full_bar[sp], bytes);
        bulk_g2s(ring + sp * C::STAGE, args.packed + (size_t)(stage0 + nxt) * C::STAGE, bytes, &full_bar[sp]);
      }
      __syncwarp();
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
      if constexpr (KIND != 0) {
        // radial kernels: r^2 = -2 (S + b_i) with S = X'_i.X'_j - n_j/2 and b_i = -n_i/2 (c1 = 1); sc.x / sc.z = alpha_j
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
          double kk[2], ph[2];
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const double r2 = fmax(-2.0 * (S[mt][e] + bi[mt]), 0.0);
            if (args.kind == KREG_KERNEL_GAUSSIAN) {
              kk[e] = exp_from_scaled_cvt(-0.5 * kExpScale * args.rinv_s2 * r2, tab);
              ph[e] = kk[e] * args.rinv_s2;
            } else {
              const double r = sqrt(r2), x = r * args.inv_sigma, ex = exp_from_scaled_cvt(-kExpScale * x, tab);
              if (args.kind == KREG_KERNEL_MATERN) {
                const double y = 2.0 * x;
                double pc, pd;
                if (args.nn <= 3) {
                  // the orders in use: fixed-length Horner forms (coefficients above the order are zero), so every
                  // coefficient is a constant-bank operand of its FMA -- no indexed loads, no loop
                  pc = fma(fma(fma(args.mat_c[3], y, args.mat_c[2]), y, args.mat_c[1]), y, args.mat_c[0]);
                  pd = fma(fma(args.mat_d[2], y, args.mat_d[1]), y, args.mat_d[0]);
                } else {
                  pc = args.mat_c[args.nn];
                  pd = args.mat_d[args.nn - 1];
                  for (int m = args.nn - 1; m >= 0; m--) pc = fma(pc, y, args.mat_c[m]);
                  for (int m = args.nn - 2; m >= 0; m--) pd = fma(pd, y, args.mat_d[m]);
                }
                kk[e] = ex * pc;
                ph[e] = ex * pd;
              } else {
                kk[e] = ex;
                ph[e] = r > 0.0 ? ex * args.inv_sigma / r : 0.0;
              }
            }
          }
          W[mt][0] = sc.x * ph[0];
          W[mt][1] = sc.z * ph[1];
          esum[(WANT_G && KIND == 0) ? 0 : mt] += sc.x * kk[0] + sc.z * kk[1];
        }
        if (!WANT_G) continue;
      } else {
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
      }  // KIND == 0
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
  if (WANT_G && KIND != 0) {
    // radial kernels: the energy sums travel beside the accumulators, in the (now dead) row-term array
#pragma unroll
    for (int mt = 0; mt < MT; mt++) {
      double s = esum[(WANT_G && KIND == 0) ? 0 : mt];
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (q == 0) bias[(warp * MT + mt) * 8 + g] = s;
    }
  }
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
  if (SPLIT) {
    // small-batch mode: hand the raw accumulators of the valid rows to predict_finish_kernel
    double *dst = args.partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * C::ROWS * TS;
    const int64_t left = args.npoints - row0;
    const int nvalid = (int)(left < C::ROWS ? left : C::ROWS);
    for (int e = tid; e < nvalid * TS; e += kThreads) {
      KREG_CHECK_STORE(dst + e, args.partial, args.partial_doubles);
      dst[e] = tile[e];
    }
    if (KIND != 0) {  // radial kernels: this slice's energy sums, [slice][geometry]
      double *de = args.partial_e + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * C::ROWS;
      for (int r = tid; r < nvalid; r += kThreads) {
        KREG_CHECK_STORE(de + r, args.partial_e, args.partial_e_doubles);
        de[r] = bias[r];
      }
    }
    return;
  }
  for (int r = tid; r < C::ROWS; r += kThreads) {
    const int64_t p = row0 + r;
    if (p < args.npoints) predict_row_epilogue<NAT, WANT_G>(args, tile + r * TS, p, (WANT_G && KIND != 0) ? bias + r : nullptr);
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

template <int NAT, int MT, bool HAS_V, bool A_IN_SMEM, int GPS, int NSTAGE, bool WANT_G, int NW, bool SPLIT = false, int KIND = 0>
static int launch_predict_g(PredictArgs a, cudaStream_t st) {
  using C = PredictCfg<NAT, MT, HAS_V, A_IN_SMEM, GPS, NSTAGE, NW>;
  auto kern = predict_kernel<NAT, MT, HAS_V, A_IN_SMEM, GPS, NSTAGE, WANT_G, NW, SPLIT, KIND>;
  static_assert(C::SMEM_BYTES <= 227 * 1024, "shared memory budget exceeded");
  KREG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
  const int64_t blocks = (a.npoints + C::ROWS - 1) / C::ROWS;
  dim3 grid((unsigned)blocks, (unsigned)(SPLIT ? a.nslices : 1));
  kern<<<grid, C::THREADS, C::SMEM_BYTES, st>>>(a);
  KREG_LAUNCH_CHECK();
  if (SPLIT) {
    auto fin = predict_finish_kernel<NAT, C::ROWS, C::TS>;
    fin<<<dim3((unsigned)blocks, C::ROWS), kThreads, 0, st>>>(a);
    KREG_LAUNCH_CHECK();
  }
  return KREG_OK;
}

// Workspace (bytes) the small-batch mode needs for this launch; 0 when the batch is large enough to run unsplit.
template <int NAT, int MT, bool HAS_V, bool A_IN_SMEM, int GPS, int NSTAGE, int NW>
static size_t predict_ws_cfg(int ngroups, int64_t npoints, int sms, int *nslices, int *stages_per_slice) {
  using C = PredictCfg<NAT, MT, HAS_V, A_IN_SMEM, GPS, NSTAGE, NW>;
  const int64_t blocks = (npoints + C::ROWS - 1) / C::ROWS;
  plan_slices<GPS>(ngroups, blocks, sms, nslices, stages_per_slice);
  return *nslices > 1 ? (size_t)*nslices * blocks * C::ROWS * C::TS * sizeof(double) : 0;
}

// the same configuration for a model on a radial kernel function (values-trained)
template <int NAT, int MT, bool A_IN_SMEM, int GPS, int NSTAGE, int NW = kWarps>
static int launch_predict_radial_cfg(PredictArgs a, int sms, void *ws, size_t ws_bytes, bool query, size_t *need,
                                     cudaStream_t st) {
  // small batches split the training set over the chip like the Gaussian model does; the slices' energy sums follow the
  // partial accumulators in the workspace
  using C = PredictCfg<NAT, MT, false, A_IN_SMEM, GPS, NSTAGE, NW>;
  int nslices = 1, per = 0;
  const size_t acc_bytes = predict_ws_cfg<NAT, MT, false, A_IN_SMEM, GPS, NSTAGE, NW>(a.ngroups, a.npoints, sms, &nslices, &per);
  const int64_t blocks = (a.npoints + C::ROWS - 1) / C::ROWS;
  const size_t bytes = nslices > 1 ? acc_bytes + (size_t)nslices * blocks * C::ROWS * sizeof(double) : 0;
  if (query) {
    *need = bytes;
    return KREG_OK;
  }
  a.nslices = 1;
  a.stages_per_slice = 0;
  a.partial = a.partial_e = nullptr;
  if (nslices > 1 && ws && ws_bytes >= bytes) {
    a.nslices = nslices;
    a.stages_per_slice = per;
    a.partial = static_cast<double *>(ws);
    a.partial_e = a.partial + acc_bytes / sizeof(double);
    a.partial_doubles = acc_bytes / sizeof(double);
    a.partial_e_doubles = (bytes - acc_bytes) / sizeof(double);
    return launch_predict_g<NAT, MT, false, A_IN_SMEM, GPS, NSTAGE, true, NW, true, 1>(a, st);
  }
  return a.G ? launch_predict_g<NAT, MT, false, A_IN_SMEM, GPS, NSTAGE, true, NW, false, 1>(a, st)
             : launch_predict_g<NAT, MT, false, A_IN_SMEM, GPS, NSTAGE, false, NW, false, 1>(a, st);
}

template <int NAT, int MT, bool HAS_V, bool A_IN_SMEM, int GPS, int NSTAGE, int NW = kWarps>
static int launch_predict_cfg(PredictArgs a, int sms, void *ws, size_t ws_bytes, bool query, size_t *need,
                              cudaStream_t st) {
  int nslices, per;
  const size_t bytes = predict_ws_cfg<NAT, MT, HAS_V, A_IN_SMEM, GPS, NSTAGE, NW>(a.ngroups, a.npoints, sms, &nslices, &per);
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
    a.partial_doubles = bytes / sizeof(double);
  }
  // energy-only requests skip the second GEMM entirely (the split mode always carries the full accumulators)
  if (a.nslices > 1) return launch_predict_g<NAT, MT, HAS_V, A_IN_SMEM, GPS, NSTAGE, true, NW, true>(a, st);
  return a.G ? launch_predict_g<NAT, MT, HAS_V, A_IN_SMEM, GPS, NSTAGE, true, NW>(a, st)
                                : launch_predict_g<NAT, MT, HAS_V, A_IN_SMEM, GPS, NSTAGE, false, NW>(a, st);
}

template <int NAT>
static int launch_predict_nat(const PredictArgs &a, bool has_v, int sms, void *ws, size_t ws_bytes, bool query,
                              size_t *need, cudaStream_t st) {
  constexpr TcShape sh(NAT);
  if (a.radial) {  // radial kernel functions: the value-model configuration of each size class
    if constexpr (sh.D <= 36) return launch_predict_radial_cfg<NAT, 2, false, 3, 3>(a, sms, ws, ws_bytes, query, need, st);
    else if constexpr (sh.D <= 66) return launch_predict_radial_cfg<NAT, 2, false, 2, 4>(a, sms, ws, ws_bytes, query, need, st);
    else if constexpr (sh.D <= 210) return launch_predict_radial_cfg<NAT, 1, true, 1, 3>(a, sms, ws, ws_bytes, query, need, st);
    else return launch_predict_radial_cfg<NAT, 1, true, 1, 2>(a, sms, ws, ws_bytes, query, need, st);
  }
  if constexpr (sh.D <= 36) {
    // small molecules: test-row fragments live in registers, 2 m-tiles per warp, two CTAs per SM (16 warps hide the
    // fixed DMMA issue latency better than one CTA with 4 m-tiles per warp: 51.5 vs 53.7 ms on cfg2)
    return has_v ? launch_predict_cfg<NAT, 2, true, false, 1, 3>(a, sms, ws, ws_bytes, query, need, st)
                 : launch_predict_cfg<NAT, 2, false, false, 3, 3>(a, sms, ws, ws_bytes, query, need, st);
  } else if constexpr (sh.D <= 66) {
    return has_v ? launch_predict_cfg<NAT, 1, true, false, 1, 4>(a, sms, ws, ws_bytes, query, need, st)
                 : launch_predict_cfg<NAT, 2, false, false, 2, 4>(a, sms, ws, ws_bytes, query, need, st);
  } else if constexpr (sh.D <= 210) {
    // large molecules: accumulators fill the register file; test-row fragments stay in shared memory
    return has_v ? launch_predict_cfg<NAT, 1, true, true, 1, 2>(a, sms, ws, ws_bytes, query, need, st)
                 : launch_predict_cfg<NAT, 1, false, true, 1, 3>(a, sms, ws, ws_bytes, query, need, st);
  } else {
    // 22-24 atoms (D <= 276): the 64-row test tile plus a two-stage ring no longer fit next to a gradient model's
    // chunks (72 KB each), so those run with four warps (32 test rows) per CTA; value models keep eight warps
    return has_v ? launch_predict_cfg<NAT, 1, true, true, 1, 2, 4>(a, sms, ws, ws_bytes, query, need, st)
                 : launch_predict_cfg<NAT, 1, false, true, 1, 2>(a, sms, ws, ws_bytes, query, need, st);
  }
}

// test geometries handled by one CTA of the kernel launch_predict_nat<NAT> would pick
template <int NAT>
static int tile_rows_nat(bool has_v) {
  constexpr TcShape sh(NAT);
  const int mt = sh.D <= 36 ? 2 : sh.D <= 66 ? (has_v ? 1 : 2) : 1;
  return (sh.D > 210 && has_v ? 4 : kWarps) * mt * 8;
}

}  // namespace kreg
