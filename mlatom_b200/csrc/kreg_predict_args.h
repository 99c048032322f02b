// kreg_predict_args.h -- argument block of the prediction kernels (shared by the dispatcher and the
// per-size instantiation objects).
#pragma once
#include <stdint.h>

namespace kreg {

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
  double *partial_e;  // radial-kernel models: the slices' energy sums, [slice][geometry] (behind `partial` in the workspace)
  size_t partial_doubles, partial_e_doubles;  // their extents (checked by the bounds-checking build)
  // models on MLatomF's radial kernel functions (predict_kernel<..., KIND = 1>; kreg_predict_radial): the packed model
  // then carries alpha_j itself and the plain column term -n_j/2, c1 is 1 (row term -n_i/2) and inv_s2 is 1
  int radial;         // 1: radial-kernel model (0: the Gaussian KREG model, everything below unused)
  int kind;           // KREG_KERNEL_GAUSSIAN / _EXPONENTIAL / _MATERN
  int nn;             // Matern order
  double inv_sigma;   // 1 / sigma
  double rinv_s2;     // 1 / sigma^2 (Gaussian through the radial path)
  double mat_c[12];   // k   = exp(-r/sigma) sum_m mat_c[m] (2r/sigma)^m
  double mat_d[12];   // phi = exp(-r/sigma) sum_m mat_d[m] (2r/sigma)^m
};

}  // namespace kreg
