// kreg_solve.cu -- (K + Lambda) alpha = y by Cholesky: the one library call on the KREG path.
// Replaces mathUtils.solveSysLinEqs (mlatom/fortran/mathUtils.f90:8-104: dpotrf('U') + dpotrs('U');
// MLatomF_src/mathUtils.f90:241-350) with cuSOLVER's 64-bit potrf/potrs, in place: the reference
// copies K into `InvMat` first (KREG.f90:613-616), doubling memory; here the matrix written by
// kreg_build_kernel_matrix already carries the lambda shift and is factorised where it lies.
#include <cusolverDn.h>

#include "kreg_common.cuh"

namespace kreg {

struct SolverCtx {
  cusolverDnHandle_t handle = nullptr;
  cusolverDnParams_t params = nullptr;
  int device = -1;
};

static thread_local SolverCtx tls_ctx;

static int get_ctx(SolverCtx **out) {
  int dev = 0;
  KREG_CUDA_TRY(cudaGetDevice(&dev));
  if (tls_ctx.handle && tls_ctx.device != dev) {
    cusolverDnDestroyParams(tls_ctx.params);
    cusolverDnDestroy(tls_ctx.handle);
    tls_ctx = SolverCtx();
  }
  if (!tls_ctx.handle) {
    if (cusolverDnCreate(&tls_ctx.handle) != CUSOLVER_STATUS_SUCCESS) {
      set_last_error("cusolverDnCreate", "failed");
      return KREG_E_CUDA;
    }
    if (cusolverDnCreateParams(&tls_ctx.params) != CUSOLVER_STATUS_SUCCESS) {
      set_last_error("cusolverDnCreateParams", "failed");
      return KREG_E_CUDA;
    }
    tls_ctx.device = dev;
  }
  *out = &tls_ctx;
  return KREG_OK;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace kreg

using namespace kreg;

// Row-major "lower" (j <= i) is column-major "upper": factorise K = U^T U like the reference's 'U'.
static const cublasFillMode_t kUplo = CUBLAS_FILL_MODE_UPPER;

extern "C" int kreg_solve_workspace_bytes(int64_t m, size_t *device_bytes, size_t *host_bytes) {
  if (m <= 0 || !device_bytes || !host_bytes) return KREG_E_BADARG;
  SolverCtx *ctx;
  int rc = get_ctx(&ctx);
  if (rc) return rc;
  size_t dws = 0, hws = 0;
  cusolverStatus_t st = cusolverDnXpotrf_bufferSize(ctx->handle, ctx->params, kUplo, m, CUDA_R_64F, nullptr, m,
                                                    CUDA_R_64F, &dws, &hws);
  if (st != CUSOLVER_STATUS_SUCCESS) {
    set_last_error("cusolverDnXpotrf_bufferSize", "failed");
    return KREG_E_CUDA;
  }
  *device_bytes = align_up(dws, 256) + 256;  // + info word
  *host_bytes = hws + 16;
  return KREG_OK;
}

extern "C" int kreg_solve(int64_t m, double *K, int64_t ldk, double *y_inout, void *device_workspace,
                          size_t device_bytes, void *host_workspace, size_t host_bytes, void *stream) {
  if (m <= 0 || !K || ldk < m || !y_inout || !device_workspace) return KREG_E_BADARG;
  SolverCtx *ctx;
  int rc = get_ctx(&ctx);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  if (cusolverDnSetStream(ctx->handle, st) != CUSOLVER_STATUS_SUCCESS) {
    set_last_error("cusolverDnSetStream", "failed");
    return KREG_E_CUDA;
  }
  size_t dws = 0, hws = 0;
  if (cusolverDnXpotrf_bufferSize(ctx->handle, ctx->params, kUplo, m, CUDA_R_64F, K, ldk, CUDA_R_64F, &dws, &hws) !=
      CUSOLVER_STATUS_SUCCESS) {
    set_last_error("cusolverDnXpotrf_bufferSize", "failed");
    return KREG_E_CUDA;
  }
  const size_t info_off = align_up(dws, 256);
  if (device_bytes < info_off + sizeof(int) || host_bytes < hws || (hws > 0 && !host_workspace))
    return KREG_E_WORKSPACE;
  int *d_info = reinterpret_cast<int *>(static_cast<char *>(device_workspace) + info_off);

  cusolverStatus_t cs = cusolverDnXpotrf(ctx->handle, ctx->params, kUplo, m, CUDA_R_64F, K, ldk, CUDA_R_64F,
                                         device_workspace, dws, host_workspace, hws, d_info);
  if (cs != CUSOLVER_STATUS_SUCCESS) {
    set_last_error("cusolverDnXpotrf", "failed");
    return KREG_E_CUDA;
  }
  int h_info = 0;
  KREG_CUDA_TRY(cudaMemcpyAsync(&h_info, d_info, sizeof(int), cudaMemcpyDeviceToHost, st));
  KREG_CUDA_TRY(cudaStreamSynchronize(st));
  if (h_info != 0) return h_info > 0 ? h_info : KREG_E_BADARG;

  cs = cusolverDnXpotrs(ctx->handle, ctx->params, kUplo, m, 1, CUDA_R_64F, K, ldk, CUDA_R_64F, y_inout, m, d_info);
  if (cs != CUSOLVER_STATUS_SUCCESS) {
    set_last_error("cusolverDnXpotrs", "failed");
    return KREG_E_CUDA;
  }
  KREG_CUDA_TRY(cudaMemcpyAsync(&h_info, d_info, sizeof(int), cudaMemcpyDeviceToHost, st));
  KREG_CUDA_TRY(cudaStreamSynchronize(st));
  return h_info == 0 ? KREG_OK : KREG_E_BADARG;
}
