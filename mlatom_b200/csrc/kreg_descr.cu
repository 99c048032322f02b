// kreg_descr.cu -- RE descriptor kernels: X_d = Req_d / R_d over all atom pairs
// (reference: mlatom/fortran/KREG.f90:223-271; mlatom/MLatomF_src/D_rel2eq.f90:17-83).
#include "kreg_common.cuh"
#include "kreg_internal.h"

namespace kreg {

__global__ void xeq_kernel(int nat, const double *__restrict__ req, double *__restrict__ xeq) {
  for (int a = threadIdx.x; a < nat; a += blockDim.x)
    for (int b = a + 1; b < nat; b++) xeq[pair_index(nat, a, b)] = sqrt(rij2_exact(req + 3 * a, req + 3 * b));
}

// One thread per descriptor element; writes are fully coalesced, xyz reads hit L1.
// MODE 0: X = xeq / R (bit-identical to the reference).  MODE 1: X - center (used by the
// assembly / packing passes, which work on mean-centred descriptors).
template <int MODE>
__global__ void __launch_bounds__(256) descriptors_kernel(int nat, int64_t npoints, const double *__restrict__ xyz,
                                                          const double *__restrict__ xeq,
                                                          const double *__restrict__ center,
                                                          double *__restrict__ X) {
  const int D = descr_size(nat);
  const int64_t total = npoints * D;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = e / D;
    int d = (int)(e - p * D);
    int a, b;
    pair_from_index(nat, d, a, b);  // inverse of KREG.f90:48
    const double *m = xyz + p * nat * 3;
    double x = xeq[d] / sqrt(rij2_exact(m + 3 * a, m + 3 * b));
    if (MODE == 1) x -= center[d];
    X[e] = x;
  }
}

// Deterministic centre for the descriptors: mean of descriptor d over the first min(N, 2048) points (one block
// per d, fixed-shape tree).  Any centre is mathematically exact -- |X_i - X_j| does not depend on it -- it only
// keeps |X'|^2 small so that n_i + n_j - 2 X'_i.X'_j loses no digits; a subsample mean does that at O(1) cost.
__global__ void __launch_bounds__(256) descriptor_mean_kernel(int nat, int64_t npoints,
                                                              const double *__restrict__ xyz,
                                                              const double *__restrict__ xeq,
                                                              double *__restrict__ center) {
  __shared__ double red[256];
  const int d = blockIdx.x;
  int a, b;
  pair_from_index(nat, d, a, b);
  if (npoints > 2048) npoints = 2048;
  double s = 0.0;
  for (int64_t p = threadIdx.x; p < npoints; p += 256) {
    const double *m = xyz + p * nat * 3;
    s += xeq[d] / sqrt(rij2_exact(m + 3 * a, m + 3 * b));
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) center[d] = npoints > 0 ? red[0] / (double)npoints : 0.0;
}

int launch_descriptors(int nat, int64_t npoints, const double *xyz, const double *xeq, const double *center,
                       double *X, cudaStream_t st) {
  if (npoints == 0) return KREG_OK;
  const int64_t total = npoints * descr_size(nat);
  int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  if (center)
    descriptors_kernel<1><<<blocks, 256, 0, st>>>(nat, npoints, xyz, xeq, center, X);
  else
    descriptors_kernel<0><<<blocks, 256, 0, st>>>(nat, npoints, xyz, xeq, nullptr, X);
  KREG_LAUNCH_CHECK();
  return KREG_OK;
}

int launch_descriptor_mean(int nat, int64_t npoints, const double *xyz, const double *xeq, double *center,
                           cudaStream_t st) {
  descriptor_mean_kernel<<<descr_size(nat), 256, 0, st>>>(nat, npoints, xyz, xeq, center);
  KREG_LAUNCH_CHECK();
  return KREG_OK;
}

}  // namespace kreg

extern "C" int kreg_xeq(int natoms, const double *req, double *xeq, void *stream) {
  if (natoms < KREG_MIN_NATOMS || natoms > KREG_MAX_NATOMS) return KREG_E_UNSUPPORTED;
  if (!req || !xeq) return KREG_E_BADARG;
  kreg::xeq_kernel<<<1, 32, 0, kreg::as_stream(stream)>>>(natoms, req, xeq);
  KREG_LAUNCH_CHECK();
  return KREG_OK;
}

extern "C" int kreg_descriptors(int natoms, int64_t npoints, const double *xyz, const double *xeq, double *X,
                                void *stream) {
  if (natoms < KREG_MIN_NATOMS || natoms > KREG_MAX_NATOMS) return KREG_E_UNSUPPORTED;
  if (npoints < 0 || (npoints > 0 && (!xyz || !xeq || !X))) return KREG_E_BADARG;
  return kreg::launch_descriptors(natoms, npoints, xyz, xeq, nullptr, X, kreg::as_stream(stream));
}
