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
};

}  // namespace kreg
