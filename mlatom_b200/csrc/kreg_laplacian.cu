// kreg_laplacian.cu -- energies and ANALYTIC Cartesian gradients of a values-trained model on MLatomF's Laplacian kernel
//   k(X_i, X_j) = exp(-|X_i - X_j|_1 / sigma)          (Kfunk, A_KRR_kernel.f90:830-831; L1 distance :2066-2069)
// from a materialised kernel block K[i][j] (kreg_kernel_block_ex):
//   E_i       = prior + sum_j alpha_j K_ij                                        (Yest_KRR, A_KRR_kernel.f90:225-245)
//   dE/dX_id  = -(1/sigma) sum_j alpha_j K_ij sign(X_id - X_jd)
//   dE/dM_ia  = sum_c dE/dX_d(a,c) * (-X_d / R_ac^2) (M_a - M_c)                  (RE descriptor Jacobian, KREG.f90:121)
// The reference has no closed form for this kernel: dKdXid takes a central difference of Kfunk in descriptor space
// (A_KRR_kernel.f90:1222-1230, step sqrt(eps) X_id), which approximates exactly this derivative away from the kinks
// X_id == X_jd (where the analytic form uses sign(0) = 0).
// The L1 distance has no GEMM form, so this is a plain FP64 kernel: a CTA owns TI test geometries, its 256 threads own the
// TI*D gradient elements (i, d) in registers and sweep the training set in tiles of 64 columns whose weights
// alpha_j K_ij are staged in shared memory; training descriptors are read straight from L1/L2 (consecutive threads,
// consecutive d).  Every sum runs in a fixed order: results are run-to-run identical.
#include "../../include/kreg_b200.h"
#include "kreg_common.cuh"
#include "kreg_internal.h"

namespace kreg {

constexpr int kLapThreads = 256;
constexpr int kLapTJ = 64;

template <int TI, int MAXE>
__global__ void __launch_bounds__(kLapThreads) laplacian_gradient_kernel(
    int nat, int64_t ntr, const double *__restrict__ Xtr, const double *__restrict__ alpha, const double *__restrict__ xeq,
    double inv_sigma, double prior, int64_t npoints, const double *__restrict__ xyz, const double *__restrict__ K, int64_t ldk,
    double *__restrict__ E, double *__restrict__ G) {
  extern __shared__ __align__(16) double lap_smem[];
  const int D = descr_size(nat), tid = threadIdx.x, n3 = nat * 3;
  double *xs = lap_smem;          // [TI][D]  test descriptors; after the sweep: dE/dX_d * (-X_d / R^2)
  double *fs = xs + TI * D;       // [TI][D]  -X_d / R_d^2
  double *ws = fs + TI * D;       // [TI][kLapTJ] alpha_j K_ij of the current tile
  double *ms = ws + TI * kLapTJ;  // [TI][3 Nat] coordinates
  const int64_t i0 = (int64_t)blockIdx.x * TI;
  const int nrows = (int)((npoints - i0) < TI ? (npoints - i0) : TI);
  for (int t = tid; t < nrows * n3; t += kLapThreads) ms[t] = xyz[i0 * n3 + t];
  __syncthreads();
  for (int t = tid; t < nrows * D; t += kLapThreads) {
    const int i = t / D, d = t - i * D;
    int a, b;
    pair_from_index(nat, d, a, b);
    const double r2 = rij2_exact(ms + i * n3 + 3 * a, ms + i * n3 + 3 * b);
    const double x = xeq[d] / sqrt(r2);
    xs[t] = x;
    fs[t] = -x / r2;
  }
  __syncthreads();
  // this thread's gradient elements e = tid + 256 k  ->  (i, d)
  double acc[MAXE], xi[MAXE];
  int wi[MAXE], dd[MAXE];
#pragma unroll
  for (int k = 0; k < MAXE; k++) {
    const int e = tid + kLapThreads * k;
    const bool live = e < nrows * D;
    const int i = live ? e / D : 0;
    wi[k] = live ? i * kLapTJ : -1;
    dd[k] = live ? e - i * D : 0;
    xi[k] = live ? xs[e] : 0.0;
    acc[k] = 0.0;
  }
  double esum = 0.0;
  for (int64_t j0 = 0; j0 < ntr; j0 += kLapTJ) {
    const int nj = (int)((ntr - j0) < kLapTJ ? (ntr - j0) : kLapTJ);
    __syncthreads();
    for (int t = tid; t < nrows * kLapTJ; t += kLapThreads) {
      const int i = t / kLapTJ, jj = t - i * kLapTJ;
      ws[t] = jj < nj ? alpha[j0 + jj] * K[(i0 + i) * ldk + j0 + jj] : 0.0;
    }
    __syncthreads();
    if (tid < nrows)
      for (int jj = 0; jj < nj; jj++) esum += ws[tid * kLapTJ + jj];
    const double *xt = Xtr + j0 * D;
#pragma unroll
    for (int k = 0; k < MAXE; k++) {
      if (wi[k] < 0) continue;
      const double *col = xt + dd[k];
      const double *w = ws + wi[k];
      const double x = xi[k];
      double s = acc[k];
#pragma unroll 4
      for (int jj = 0; jj < nj; jj++) {
        const double xj = __ldg(col + (size_t)jj * D);
        const double wj = w[jj];
        s += x > xj ? -wj : (x < xj ? wj : 0.0);
      }
      acc[k] = s;
    }
  }
  if (E && tid < nrows) E[i0 + tid] = prior + esum;
  if (!G) return;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < MAXE; k++) {
    const int e = tid + kLapThreads * k;
    if (wi[k] >= 0) xs[e] = acc[k] * inv_sigma * fs[e];
  }
  __syncthreads();
  // Cartesian gradient: atom a collects its partners in the order c = 0..Nat-1
  for (int t = tid; t < nrows * n3; t += kLapThreads) {
    const int i = t / n3, r = t - i * n3, a = r / 3, c3 = r - a * 3;
    const double *m = ms + i * n3;
    const double *cd = xs + i * D;
    double g = 0.0;
    for (int c = 0; c < nat; c++) {
      if (c == a) continue;
      const int d = a < c ? pair_index(nat, a, c) : pair_index(nat, c, a);
      g += cd[d] * (m[3 * a + c3] - m[3 * c + c3]);
    }
    G[(i0 + i) * n3 + r] = g;
  }
}

template <int TI, int MAXE>
static int launch_laplacian(int nat, int64_t ntr, const double *Xtr, const double *alpha, const double *xeq, double sigma,
                            double prior, int64_t npoints, const double *xyz, const double *K, int64_t ldk, double *E,
                            double *G, cudaStream_t st) {
  const int D = descr_size(nat);
  const size_t smem = ((size_t)2 * TI * D + (size_t)TI * kLapTJ + (size_t)TI * nat * 3) * sizeof(double);
  auto kern = laplacian_gradient_kernel<TI, MAXE>;
  KREG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)((npoints + TI - 1) / TI), kLapThreads, smem, st>>>(nat, ntr, Xtr, alpha, xeq, 1.0 / sigma, prior, npoints,
                                                                       xyz, K, ldk, E, G);
  KREG_LAUNCH_CHECK();
  return KREG_OK;
}

}  // namespace kreg

using namespace kreg;

extern "C" int kreg_laplacian_predict(int natoms, int64_t ntrain, const double *X_train, const double *alpha,
                                      const double *xeq, double sigma, double prior, int64_t npoints, const double *xyz,
                                      const double *K, int64_t ldk, int flags, double *E, double *G, void *stream) {
  if (natoms < KREG_MIN_NATOMS || natoms > KREG_MAX_NATOMS) return KREG_E_UNSUPPORTED;
  if (ntrain < 0 || npoints < 0 || !(sigma > 0.0) || ldk < ntrain) return KREG_E_BADARG;
  if (!(flags & (KREG_PREDICT_ENERGY | KREG_PREDICT_GRADIENT))) return KREG_E_BADARG;
  if (npoints == 0) return KREG_OK;
  if (!xeq || !xyz || (ntrain > 0 && (!X_train || !alpha || !K))) return KREG_E_BADARG;
  double *e = (flags & KREG_PREDICT_ENERGY) ? E : nullptr, *g = (flags & KREG_PREDICT_GRADIENT) ? G : nullptr;
  if (((flags & KREG_PREDICT_ENERGY) && !E) || ((flags & KREG_PREDICT_GRADIENT) && !G)) return KREG_E_BADARG;
  cudaStream_t st = as_stream(stream);
  const int D = descr_size(natoms);
  // elements per thread: TI * D / 256, rounded up
  if (8 * D <= 256 * 9) return launch_laplacian<8, 9>(natoms, ntrain, X_train, alpha, xeq, sigma, prior, npoints, xyz, K, ldk, e, g, st);
  if (4 * D <= 256 * 12) return launch_laplacian<4, 12>(natoms, ntrain, X_train, alpha, xeq, sigma, prior, npoints, xyz, K, ldk, e, g, st);
  return launch_laplacian<2, 16>(natoms, ntrain, X_train, alpha, xeq, sigma, prior, npoints, xyz, K, ldk, e, g, st);
}
