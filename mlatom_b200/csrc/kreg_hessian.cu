// kreg_hessian.cu -- Hessian of the predicted energy, as the reference defines it.
//
// Replaces YhessXYZest_KRR / d2KdMiatdMibu (mlatom/fortran/KREG.f90:180-221, :393-420, driven by predict :588-594):
// values part of alpha only, and only products of first derivatives of the descriptor (the reference has no
// d2X/dM2 term).  With D_j = X_j - X_i and u_j = J_i^T D_j the reference's double loop is
//     H_i = (1/s^4) sum_j w_ij u_j u_j^T  -  (1/s^2) (sum_j w_ij) J_i^T J_i,     w_ij = alpha_j k_ij
// (J_i^T J_i collapses to one product per 3x3 block, pairs of atoms share exactly one descriptor).
// One CTA per geometry; this is a frequency-calculation path (a handful of geometries), not a throughput path.
#include "kreg_common.cuh"
#include "kreg_internal.h"

namespace kreg {

template <int NAT>
__global__ void __launch_bounds__(256) hessian_kernel(int64_t ntrain, const double *__restrict__ Xtr,
                                                      const double *__restrict__ alpha, const double *__restrict__ xeq,
                                                      double inv_s2, const double *__restrict__ xyz,
                                                      double *__restrict__ H) {
  constexpr int D = NAT * (NAT - 1) / 2, G3 = 3 * NAT, TJ = 32, DP = D | 1, UP = G3 | 1;
  constexpr int NACC = (G3 * G3 + 255) / 256;
  extern __shared__ double hs[];
  double *Xi = hs;                      // [D]
  double *Ei = Xi + D;                  // [NAT*NAT*3]  E_i(a,c)[t] = X_d (x_c - x_a)_t / R^2
  double *Dl = Ei + NAT * NAT * 3;      // [TJ*DP]
  double *U = Dl + TJ * DP;             // [TJ*UP]
  double *w = U + TJ * UP;              // [TJ]
  double &esum_s = w[TJ];

  const int tid = threadIdx.x;
  const int64_t p = blockIdx.x;
  const double *m = xyz + p * NAT * 3;
  for (int e = tid; e < NAT * NAT; e += 256) {
    const int a = e / NAT, c = e % NAT;
    double *out = Ei + e * 3;
    if (a == c) {
      out[0] = out[1] = out[2] = 0.0;
    } else {
      const int lo = a < c ? a : c, hi = a < c ? c : a;
      const int d = pair_index(NAT, lo, hi);
      const double r2 = rij2_exact(m + 3 * lo, m + 3 * hi);
      const double x = xeq[d] / sqrt(r2);
      if (a < c) Xi[d] = x;
      const double f = x / r2;
      for (int t = 0; t < 3; t++) out[t] = f * (m[3 * c + t] - m[3 * a + t]);
    }
  }
  double acc[NACC];
#pragma unroll
  for (int k = 0; k < NACC; k++) acc[k] = 0.0;
  double esum = 0.0;
  __syncthreads();

  for (int64_t j0 = 0; j0 < ntrain; j0 += TJ) {
    const int nj = (int)(ntrain - j0 < TJ ? ntrain - j0 : TJ);
    for (int e = tid; e < TJ * D; e += 256) {
      const int jj = e / D, d = e % D;
      Dl[jj * DP + d] = jj < nj ? Xtr[(j0 + jj) * D + d] - Xi[d] : 0.0;
    }
    __syncthreads();
    {  // w_j = alpha_j exp(-|D_j|^2 / 2 s^2): eight lanes per training point, warp-uniform shuffles
      const int jj = tid >> 3, sub = tid & 7;
      const double *dd = Dl + jj * DP;
      double d2 = 0.0;
      for (int d = sub; d < D; d += 8) d2 = fma(dd[d], dd[d], d2);
      d2 += __shfl_xor_sync(0xffffffffu, d2, 1);
      d2 += __shfl_xor_sync(0xffffffffu, d2, 2);
      d2 += __shfl_xor_sync(0xffffffffu, d2, 4);
      if (sub == 0) w[jj] = jj < nj ? alpha[j0 + jj] * exp(-0.5 * inv_s2 * d2) : 0.0;
    }
    for (int e = tid; e < TJ * G3; e += 256) {  // u_j[(a,t)] = sum_c E_i(a,c)[t] D_j[d_ac]
      const int q = e / TJ, jj = e % TJ, a = q / 3, t = q % 3;
      const double *dd = Dl + jj * DP;
      double s = 0.0;
      for (int c = 0; c < NAT; c++)
        if (c != a) s = fma(Ei[(a * NAT + c) * 3 + t], dd[pair_index(NAT, a < c ? a : c, a < c ? c : a)], s);
      U[jj * UP + q] = s;
    }
    __syncthreads();
    if (tid == 0)
      for (int jj = 0; jj < TJ; jj++) esum += w[jj];
#pragma unroll
    for (int k = 0; k < NACC; k++) {
      const int e = tid + k * 256;
      if (e < G3 * G3) {
        const int q = e / G3, r = e % G3;
        double s = acc[k];
        for (int jj = 0; jj < TJ; jj++) s = fma(w[jj] * U[jj * UP + q], U[jj * UP + r], s);
        acc[k] = s;
      }
    }
    __syncthreads();
  }
  if (tid == 0) esum_s = esum;
  __syncthreads();
  const double es = esum_s * inv_s2, inv_s4 = inv_s2 * inv_s2;
  double *out = H + p * G3 * G3;
#pragma unroll
  for (int k = 0; k < NACC; k++) {
    const int e = tid + k * 256;
    if (e < G3 * G3) {
      const int q = e / G3, r = e % G3, a = q / 3, t = q % 3, b = r / 3, u = r % 3;
      double jj;  // (J_i^T J_i)[(a,t),(b,u)]
      if (a != b) {
        jj = Ei[(a * NAT + b) * 3 + t] * Ei[(b * NAT + a) * 3 + u];
      } else {
        jj = 0.0;
        for (int c = 0; c < NAT; c++) jj = fma(Ei[(a * NAT + c) * 3 + t], Ei[(a * NAT + c) * 3 + u], jj);
      }
      out[e] = acc[k] * inv_s4 - es * jj;
    }
  }
}

template <int NAT>
static int launch_hessian_nat(int64_t ntrain, const double *Xtr, const double *alpha, const double *xeq, double inv_s2,
                              int64_t npoints, const double *xyz, double *H, cudaStream_t st) {
  constexpr int D = NAT * (NAT - 1) / 2, G3 = 3 * NAT, TJ = 32;
  constexpr size_t smem = sizeof(double) * (D + NAT * NAT * 3 + TJ * (D | 1) + TJ * (G3 | 1) + TJ + 1);
  static_assert(smem <= 227 * 1024, "hessian tile does not fit");
  KREG_CUDA_TRY(cudaFuncSetAttribute(hessian_kernel<NAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  hessian_kernel<NAT><<<(unsigned)npoints, 256, smem, st>>>(ntrain, Xtr, alpha, xeq, inv_s2, xyz, H);
  KREG_LAUNCH_CHECK();
  return KREG_OK;
}

// Run-time-Natoms version for molecules above KREG_FUSED_MAX_NATOMS atoms: same sums in the same order; the (3 Nat)^2
// outputs of a geometry are split into slabs of 16 x 256 entries (one CTA each, blockIdx.y), the training tile TJ is
// chosen at launch so that the shared arrays fit.
constexpr int kHessAcc = 16;

__global__ void __launch_bounds__(256) hessian_rt_kernel(int nat, int TJ, int64_t ntrain, const double *__restrict__ Xtr,
                                                         const double *__restrict__ alpha, const double *__restrict__ xeq,
                                                         double inv_s2, const double *__restrict__ xyz,
                                                         double *__restrict__ H) {
  const int D = descr_size(nat), G3 = 3 * nat, DP = D | 1, UP = G3 | 1;
  extern __shared__ double hs[];
  double *Xi = hs;                         // [D]
  double *Ei = Xi + D;                     // [nat*nat*3]
  double *Dl = Ei + nat * nat * 3;         // [TJ*DP]
  double *U = Dl + (size_t)TJ * DP;        // [TJ*UP]
  double *w = U + (size_t)TJ * UP;         // [TJ] + esum

  const int tid = threadIdx.x;
  const int64_t p = blockIdx.x;
  const int e0 = blockIdx.y * (kHessAcc * 256);
  const double *m = xyz + p * nat * 3;
  for (int e = tid; e < nat * nat; e += 256) {
    const int a = e / nat, c = e % nat;
    double *out = Ei + e * 3;
    if (a == c) {
      out[0] = out[1] = out[2] = 0.0;
    } else {
      const int lo = a < c ? a : c, hi = a < c ? c : a;
      const int d = pair_index(nat, lo, hi);
      const double r2 = rij2_exact(m + 3 * lo, m + 3 * hi);
      const double x = xeq[d] / sqrt(r2);
      if (a < c) Xi[d] = x;
      const double f = x / r2;
      for (int t = 0; t < 3; t++) out[t] = f * (m[3 * c + t] - m[3 * a + t]);
    }
  }
  double acc[kHessAcc];
#pragma unroll
  for (int k = 0; k < kHessAcc; k++) acc[k] = 0.0;
  double esum = 0.0;
  __syncthreads();

  for (int64_t j0 = 0; j0 < ntrain; j0 += TJ) {
    const int nj = (int)(ntrain - j0 < TJ ? ntrain - j0 : TJ);
    for (int e = tid; e < TJ * D; e += 256) {
      const int jj = e / D, d = e % D;
      Dl[jj * DP + d] = jj < nj ? Xtr[(j0 + jj) * D + d] - Xi[d] : 0.0;
    }
    __syncthreads();
    for (int base = 0; base < TJ; base += 32) {  // eight lanes per training point, as the templated kernel sums
      const int jj = base + (tid >> 3), sub = tid & 7;
      const double *dd = Dl + (jj < TJ ? jj : 0) * DP;
      double d2 = 0.0;
      for (int d = sub; d < D; d += 8) d2 = fma(dd[d], dd[d], d2);
      d2 += __shfl_xor_sync(0xffffffffu, d2, 1);
      d2 += __shfl_xor_sync(0xffffffffu, d2, 2);
      d2 += __shfl_xor_sync(0xffffffffu, d2, 4);
      if (sub == 0 && jj < TJ) w[jj] = jj < nj ? alpha[j0 + jj] * exp(-0.5 * inv_s2 * d2) : 0.0;
    }
    for (int e = tid; e < TJ * G3; e += 256) {
      const int q = e / TJ, jj = e % TJ, a = q / 3, t = q % 3;
      const double *dd = Dl + jj * DP;
      double s = 0.0;
      for (int c = 0; c < nat; c++)
        if (c != a) s = fma(Ei[(a * nat + c) * 3 + t], dd[pair_index(nat, a < c ? a : c, a < c ? c : a)], s);
      U[jj * UP + q] = s;
    }
    __syncthreads();
    if (tid == 0)
      for (int jj = 0; jj < TJ; jj++) esum += w[jj];
#pragma unroll
    for (int k = 0; k < kHessAcc; k++) {
      const int e = e0 + tid + k * 256;
      if (e < G3 * G3) {
        const int q = e / G3, r = e % G3;
        double s = acc[k];
        for (int jj = 0; jj < TJ; jj++) s = fma(w[jj] * U[jj * UP + q], U[jj * UP + r], s);
        acc[k] = s;
      }
    }
    __syncthreads();
  }
  if (tid == 0) w[TJ] = esum;
  __syncthreads();
  const double es = w[TJ] * inv_s2, inv_s4 = inv_s2 * inv_s2;
  double *out = H + p * G3 * G3;
#pragma unroll
  for (int k = 0; k < kHessAcc; k++) {
    const int e = e0 + tid + k * 256;
    if (e < G3 * G3) {
      const int q = e / G3, r = e % G3, a = q / 3, t = q % 3, b = r / 3, u = r % 3;
      double jj;
      if (a != b) {
        jj = Ei[(a * nat + b) * 3 + t] * Ei[(b * nat + a) * 3 + u];
      } else {
        jj = 0.0;
        for (int c = 0; c < nat; c++) jj = fma(Ei[(a * nat + c) * 3 + t], Ei[(a * nat + c) * 3 + u], jj);
      }
      out[e] = acc[k] * inv_s4 - es * jj;
    }
  }
}

static int launch_hessian_rt(int nat, int64_t ntrain, const double *Xtr, const double *alpha, const double *xeq,
                             double inv_s2, int64_t npoints, const double *xyz, double *H, cudaStream_t st) {
  const int D = descr_size(nat), G3 = 3 * nat;
  const size_t fixed = (size_t)D + (size_t)nat * nat * 3 + 2;
  const size_t per_j = (size_t)(D | 1) + (G3 | 1) + 1;
  long tj = ((long)(200 * 1024 / 8) - (long)fixed) / (long)per_j;
  if (tj > 32) tj = 32;
  if (tj < 1) return KREG_E_UNSUPPORTED;
  const size_t smem = (fixed + per_j * tj) * sizeof(double);
  KREG_CUDA_TRY(cudaFuncSetAttribute(hessian_rt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned slabs = (unsigned)((G3 * G3 + kHessAcc * 256 - 1) / (kHessAcc * 256));
  hessian_rt_kernel<<<dim3((unsigned)npoints, slabs), 256, smem, st>>>(nat, (int)tj, ntrain, Xtr, alpha, xeq, inv_s2, xyz, H);
  KREG_LAUNCH_CHECK();
  return KREG_OK;
}

}  // namespace kreg

using namespace kreg;

extern "C" int kreg_hessian(int natoms, int64_t ntrain, const double *X_train, const double *alpha_val, const double *xeq,
                            double sigma, int64_t npoints, const double *xyz, double *H, void *stream) {
  if (natoms < KREG_MIN_NATOMS || natoms > KREG_MAX_NATOMS) return KREG_E_UNSUPPORTED;
  if (ntrain < 0 || npoints < 0 || !xeq || !(sigma > 0.0)) return KREG_E_BADARG;
  if (npoints == 0) return KREG_OK;
  if (!xyz || !H || (ntrain > 0 && (!X_train || !alpha_val))) return KREG_E_BADARG;
  const double inv_s2 = 1.0 / (sigma * sigma);
  cudaStream_t st = as_stream(stream);
  if (natoms > KREG_FUSED_MAX_NATOMS)
    return launch_hessian_rt(natoms, ntrain, X_train, alpha_val, xeq, inv_s2, npoints, xyz, H, st);
  switch (natoms) {
#define KREG_CASE(N) \
  case N:            \
    return launch_hessian_nat<N>(ntrain, X_train, alpha_val, xeq, inv_s2, npoints, xyz, H, st);
    KREG_CASE(2) KREG_CASE(3) KREG_CASE(4) KREG_CASE(5) KREG_CASE(6) KREG_CASE(7) KREG_CASE(8) KREG_CASE(9)
    KREG_CASE(10) KREG_CASE(11) KREG_CASE(12) KREG_CASE(13) KREG_CASE(14) KREG_CASE(15) KREG_CASE(16) KREG_CASE(17)
    KREG_CASE(18) KREG_CASE(19) KREG_CASE(20) KREG_CASE(21) KREG_CASE(22) KREG_CASE(23) KREG_CASE(24)
#undef KREG_CASE
    default:
      return KREG_E_UNSUPPORTED;
  }
}
