// kreg_predict_inst.cu -- instantiates the prediction kernels for one range of molecule sizes.
// Compiled five times with -DKREG_INST=0..4 (Makefile); each object exports one pair of range functions.
#include "kreg_predict_kernel.cuh"

#ifndef KREG_INST
#error "compile with -DKREG_INST=0..4"
#endif

namespace kreg {

#if KREG_INST == 0
#define KREG_RANGE_CASES KREG_CASE(2) KREG_CASE(3) KREG_CASE(4) KREG_CASE(5) KREG_CASE(6) KREG_CASE(7) KREG_CASE(8) KREG_CASE(9)
#define KREG_RANGE_NAME(x) x##_r0
#elif KREG_INST == 1
#define KREG_RANGE_CASES KREG_CASE(10) KREG_CASE(11) KREG_CASE(12) KREG_CASE(13) KREG_CASE(14) KREG_CASE(15)
#define KREG_RANGE_NAME(x) x##_r1
#elif KREG_INST == 2
#define KREG_RANGE_CASES KREG_CASE(16) KREG_CASE(17) KREG_CASE(18) KREG_CASE(19) KREG_CASE(20)
#define KREG_RANGE_NAME(x) x##_r2
#elif KREG_INST == 3
#define KREG_RANGE_CASES KREG_CASE(21) KREG_CASE(22)
#define KREG_RANGE_NAME(x) x##_r3
#else
#define KREG_RANGE_CASES KREG_CASE(23) KREG_CASE(24)
#define KREG_RANGE_NAME(x) x##_r4
#endif

// returns KREG_E_UNSUPPORTED when `nat` is not in this object's range
int KREG_RANGE_NAME(launch_predict)(int nat, const PredictArgs &a, bool has_v, int sms, void *ws, size_t ws_bytes,
                                    bool query, size_t *need, cudaStream_t st) {
  switch (nat) {
#define KREG_CASE(N) \
  case N:            \
    return launch_predict_nat<N>(a, has_v, sms, ws, ws_bytes, query, need, st);
    KREG_RANGE_CASES
#undef KREG_CASE
    default:
      return KREG_E_UNSUPPORTED;
  }
}

int KREG_RANGE_NAME(tile_rows)(int nat, bool has_v) {
  switch (nat) {
#define KREG_CASE(N) \
  case N:            \
    return tile_rows_nat<N>(has_v);
    KREG_RANGE_CASES
#undef KREG_CASE
    default:
      return KREG_E_UNSUPPORTED;
  }
}

}  // namespace kreg
