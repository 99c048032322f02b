// kreg_abi.cu -- library-level entry points: version, error text, device check, FP64 issue-rate probe.
#include <string.h>

#include "kreg_common.cuh"

namespace kreg {

#ifdef KREG_DEBUG_BOUNDS
__device__ unsigned long long g_debug_failures = 0ull;
__device__ int g_debug_first_code = 0;
__device__ int g_debug_first_line = 0;
#endif

static thread_local char tls_error[512] = "";

void set_last_error(const char *what, const char *detail) {
  snprintf(tls_error, sizeof tls_error, "%s: %s", what ? what : "?", detail ? detail : "?");
}

// Same kernels as tools/fp64_probe.cu, kept in the library so bench.py can take the roofline
// denominator live on the box it is running on.
__global__ void __launch_bounds__(256) probe_dfma_kernel(double *out, int iters, double a, double b) {
  double acc[16];
#pragma unroll
  for (int c = 0; c < 16; c++) acc[c] = threadIdx.x * 1e-9 + c;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int c = 0; c < 16; c++) acc[c] = fma(acc[c], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < 16; c++) s += acc[c];
  if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) probe_dmma_kernel(double *out, int iters) {
  double c[8][2];
  const double a = 1e-3 * threadIdx.x, b = -1e-3 * threadIdx.x;
#pragma unroll
  for (int k = 0; k < 8; k++) c[k][0] = c[k][1] = 0.0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int k = 0; k < 8; k++) dmma884(c[k], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s += c[k][0] + c[k][1];
  if (s == 123.456) out[0] = s;
}

}  // namespace kreg

using namespace kreg;

extern "C" int kreg_abi_version(void) { return 1; }

// Number of failed bounds / alignment checks since the library was loaded (synchronises the device); -1 in a build
// without -DKREG_DEBUG_BOUNDS.  first_code / first_line (may be NULL) identify the first failure (kreg_common.cuh).
extern "C" long long kreg_debug_bounds_failures(int *first_code, int *first_line) {
#ifdef KREG_DEBUG_BOUNDS
  unsigned long long n = 0;
  int code = 0, line = 0;
  if (cudaDeviceSynchronize() != cudaSuccess) return -2;
  if (cudaMemcpyFromSymbol(&n, g_debug_failures, sizeof n) != cudaSuccess) return -2;
  cudaMemcpyFromSymbol(&code, g_debug_first_code, sizeof code);
  cudaMemcpyFromSymbol(&line, g_debug_first_line, sizeof line);
  if (first_code) *first_code = code;
  if (first_line) *first_line = line;
  return (long long)n;
#else
  if (first_code) *first_code = 0;
  if (first_line) *first_line = 0;
  return -1;
#endif
}

extern "C" const char *kreg_last_error(void) { return tls_error; }

extern "C" int kreg_device_check(void) {
  int dev = 0;
  KREG_CUDA_TRY(cudaGetDevice(&dev));
  int major = 0;
  KREG_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) {
    set_last_error("kreg_device_check", "libkreg_b200 is built for sm_100a only");
    return KREG_E_UNSUPPORTED;
  }
  return KREG_OK;
}

extern "C" int kreg_stream_synchronize(void *stream) {
  KREG_CUDA_TRY(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  return KREG_OK;
}

extern "C" int kreg_measure_fp64_peak(int kind, int reps, double ms_target, double *tflops_host) {
  if (!tflops_host || reps < 1 || (kind != 0 && kind != 1)) return KREG_E_BADARG;
  int rc = kreg_device_check();
  if (rc) return rc;
  int dev = 0, sms = 0;
  KREG_CUDA_TRY(cudaGetDevice(&dev));
  KREG_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  double *out = nullptr;
  KREG_CUDA_TRY(cudaMalloc(&out, 64));
  cudaEvent_t e0, e1;
  KREG_CUDA_TRY(cudaEventCreate(&e0));
  KREG_CUDA_TRY(cudaEventCreate(&e1));
  const int blocks = sms * 8, threads = 256;
  // flops per launch iteration: DFMA 8*16 FMAs/thread, DMMA 4*8 tiles of 512 flop per warp
  const double flop_per_iter = kind == 0 ? 2.0 * blocks * threads * 8 * 16 : (double)blocks * threads / 32 * 4 * 8 * 512;
  int iters = 200;
  double best = 0.0;
  for (int r = -2; r < reps; r++) {
    KREG_CUDA_TRY(cudaEventRecord(e0));
    if (kind == 0)
      probe_dfma_kernel<<<blocks, threads>>>(out, iters, 0.999, 1e-3);
    else
      probe_dmma_kernel<<<blocks, threads>>>(out, iters);
    KREG_CUDA_TRY(cudaEventRecord(e1));
    KREG_CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    KREG_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    if (r == -2) {  // calibrate the launch length to ~ms_target
      double scale = ms_target / (ms > 1e-3 ? ms : 1e-3);
      iters = (int)(iters * scale) + 1;
      continue;
    }
    if (r >= 0) {
      double tf = flop_per_iter * iters / (ms * 1e-3) * 1e-12;
      if (tf > best) best = tf;
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  KREG_LAUNCH_CHECK();
  *tflops_host = best;
  return KREG_OK;
}
