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
  // the one synchronisation of the call: the caller has to know whether the factorisation succeeded before the
  // matrix is reused (pinned host word when the caller passed host scratch, else a stack word)
  int stack_info = 0;
  int *h_info_p = (host_workspace && host_bytes >= hws + sizeof(int))
                      ? reinterpret_cast<int *>(static_cast<char *>(host_workspace) + hws)
                      : &stack_info;
  KREG_CUDA_TRY(cudaMemcpyAsync(h_info_p, d_info, sizeof(int), cudaMemcpyDeviceToHost, st));
  KREG_CUDA_TRY(cudaStreamSynchronize(st));
  const int h_info = *h_info_p;
  if (h_info != 0) return h_info > 0 ? h_info : KREG_E_BADARG;

  cs = cusolverDnXpotrs(ctx->handle, ctx->params, kUplo, m, 1, CUDA_R_64F, K, ldk, CUDA_R_64F, y_inout, m, d_info);
  if (cs != CUSOLVER_STATUS_SUCCESS) {
    set_last_error("cusolverDnXpotrs", "failed");
    return KREG_E_CUDA;
  }
  // potrs' info only reports illegal arguments (checked above): no second read-back / synchronisation; the solve
  // stays asynchronous on `stream` from here on
  return KREG_OK;
}


// ------------------------------------------------------------------------------------------------
// Symmetric-indefinite fallback (the reference: solveSysLinEqsBK, mlatom/fortran/mathUtils.f90:106-186, dsytrf +
// dsytrs, entered from :24-26 when dpotrf reports a non-positive leading minor).
// ------------------------------------------------------------------------------------------------
namespace kreg {

__global__ void widen_pivots_kernel(int64_t m, const int *__restrict__ p32, int64_t *__restrict__ p64) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x)
    p64[i] = p32[i];
}

// K[j][i] = K[i][j] for j < i (row-major): fills the other triangle for the LU route
__global__ void __launch_bounds__(256) mirror_lower_kernel(int64_t m, double *__restrict__ K, int64_t ldk) {
  __shared__ double t[32][33];
  const int64_t bi = blockIdx.y, bj = blockIdx.x;
  if (bj > bi) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int64_t i = bi * 32 + r, j = bj * 32 + tx;
    t[r][tx] = (i < m && j < m) ? K[i * ldk + j] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int64_t j = bj * 32 + r, i = bi * 32 + tx;  // writes K[j][i], coalesced over i
    if (i < m && j < m && j < i) K[j * ldk + i] = t[tx][r];
  }
}

struct IndefPlan {
  int method;  // 1 = Bunch-Kaufman (sytrf/sytrs), 2 = pivoted LU (getrf/getrs)
  size_t off_piv32, off_piv64, off_work, off_info, work_bytes, total, host_bytes;
  int lwork_sytrf;
};

// sytrf is only offered through cuSOLVER's 32-bit interface: keep m * m inside int32 for it
static const int64_t kSytrfMaxM = 46000;

static int plan_indefinite(SolverCtx *ctx, int64_t m, int method, IndefPlan *pl) {
  if (method == 0) method = m <= kSytrfMaxM ? 1 : 2;
  if (method == 1 && m > kSytrfMaxM) return KREG_E_UNSUPPORTED;
  pl->method = method;
  size_t dws = 0, hws = 0, dws2 = 0, hws2 = 0;
  pl->lwork_sytrf = 0;
  if (method == 1) {
    int lwork = 0;
    if (cusolverDnDsytrf_bufferSize(ctx->handle, (int)m, nullptr, (int)m, &lwork) != CUSOLVER_STATUS_SUCCESS) {
      set_last_error("cusolverDnDsytrf_bufferSize", "failed");
      return KREG_E_CUDA;
    }
    pl->lwork_sytrf = lwork;
    dws = (size_t)lwork * sizeof(double);
    if (cusolverDnXsytrs_bufferSize(ctx->handle, kUplo, m, 1, CUDA_R_64F, nullptr, m, nullptr, CUDA_R_64F, nullptr, m,
                                    &dws2, &hws2) != CUSOLVER_STATUS_SUCCESS) {
      set_last_error("cusolverDnXsytrs_bufferSize", "failed");
      return KREG_E_CUDA;
    }
  } else {
    if (cusolverDnXgetrf_bufferSize(ctx->handle, ctx->params, m, m, CUDA_R_64F, nullptr, m, CUDA_R_64F, &dws, &hws) !=
        CUSOLVER_STATUS_SUCCESS) {
      set_last_error("cusolverDnXgetrf_bufferSize", "failed");
      return KREG_E_CUDA;
    }
  }
  pl->work_bytes = dws > dws2 ? dws : dws2;
  pl->host_bytes = (hws > hws2 ? hws : hws2) + 16;
  size_t off = 0;
  pl->off_piv32 = off;
  off += align_up((size_t)m * sizeof(int), 256);
  pl->off_piv64 = off;
  off += align_up((size_t)m * sizeof(int64_t), 256);
  pl->off_work = off;
  off += align_up(pl->work_bytes, 256);
  pl->off_info = off;
  off += 256;
  pl->total = off;
  return KREG_OK;
}

}  // namespace kreg

extern "C" int kreg_solve_indefinite_workspace_bytes(int64_t m, int method, size_t *device_bytes, size_t *host_bytes) {
  if (m <= 0 || !device_bytes || !host_bytes || method < 0 || method > 2) return KREG_E_BADARG;
  SolverCtx *ctx;
  int rc = get_ctx(&ctx);
  if (rc) return rc;
  IndefPlan pl;
  rc = plan_indefinite(ctx, m, method, &pl);
  if (rc) return rc;
  *device_bytes = pl.total;
  *host_bytes = pl.host_bytes;
  return KREG_OK;
}

extern "C" int kreg_solve_indefinite(int64_t m, double *K, int64_t ldk, double *y_inout, int method,
                                     void *device_workspace, size_t device_bytes, void *host_workspace,
                                     size_t host_bytes, void *stream) {
  if (m <= 0 || !K || ldk < m || !y_inout || !device_workspace || method < 0 || method > 2) return KREG_E_BADARG;
  SolverCtx *ctx;
  int rc = get_ctx(&ctx);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  if (cusolverDnSetStream(ctx->handle, st) != CUSOLVER_STATUS_SUCCESS) {
    set_last_error("cusolverDnSetStream", "failed");
    return KREG_E_CUDA;
  }
  IndefPlan pl;
  rc = plan_indefinite(ctx, m, method, &pl);
  if (rc) return rc;
  if (device_bytes < pl.total || host_bytes + 16 < pl.host_bytes || (pl.host_bytes > 16 && !host_workspace))
    return KREG_E_WORKSPACE;
  char *ws = static_cast<char *>(device_workspace);
  int *piv32 = reinterpret_cast<int *>(ws + pl.off_piv32);
  int64_t *piv64 = reinterpret_cast<int64_t *>(ws + pl.off_piv64);
  void *work = ws + pl.off_work;
  int *d_info = reinterpret_cast<int *>(ws + pl.off_info);
  int h_info = 0;
  if (pl.method == 1) {
    // row-major lower == column-major upper: the same triangle kreg_solve factorises
    if (cusolverDnDsytrf(ctx->handle, kUplo, (int)m, K, (int)ldk, piv32, static_cast<double *>(work), pl.lwork_sytrf,
                         d_info) != CUSOLVER_STATUS_SUCCESS) {
      set_last_error("cusolverDnDsytrf", "failed");
      return KREG_E_CUDA;
    }
    widen_pivots_kernel<<<(unsigned)((m + 255) / 256 < 1024 ? (m + 255) / 256 : 1024), 256, 0, st>>>(m, piv32, piv64);
    KREG_LAUNCH_CHECK();
    KREG_CUDA_TRY(cudaMemcpyAsync(&h_info, d_info, sizeof(int), cudaMemcpyDeviceToHost, st));
    KREG_CUDA_TRY(cudaStreamSynchronize(st));
    if (h_info != 0) return h_info > 0 ? h_info : KREG_E_BADARG;  // > 0: exactly singular block D(i,i)
    size_t dws2 = 0, hws2 = 0;
    if (cusolverDnXsytrs_bufferSize(ctx->handle, kUplo, m, 1, CUDA_R_64F, K, ldk, piv64, CUDA_R_64F, y_inout, m, &dws2,
                                    &hws2) != CUSOLVER_STATUS_SUCCESS ||
        dws2 > pl.work_bytes) {
      set_last_error("cusolverDnXsytrs_bufferSize", "failed");
      return KREG_E_CUDA;
    }
    if (cusolverDnXsytrs(ctx->handle, kUplo, m, 1, CUDA_R_64F, K, ldk, piv64, CUDA_R_64F, y_inout, m, work, dws2,
                         host_workspace, hws2, d_info) != CUSOLVER_STATUS_SUCCESS) {
      set_last_error("cusolverDnXsytrs", "failed");
      return KREG_E_CUDA;
    }
    return KREG_OK;
  }
  // pivoted LU of the full matrix (the other triangle is filled from the one the assembly wrote)
  {
    const unsigned nb = (unsigned)((m + 31) / 32);
    mirror_lower_kernel<<<dim3(nb, nb), 256, 0, st>>>(m, K, ldk);
    KREG_LAUNCH_CHECK();
  }
  size_t dws = 0, hws = 0;
  if (cusolverDnXgetrf_bufferSize(ctx->handle, ctx->params, m, m, CUDA_R_64F, K, ldk, CUDA_R_64F, &dws, &hws) !=
          CUSOLVER_STATUS_SUCCESS ||
      dws > pl.work_bytes) {
    set_last_error("cusolverDnXgetrf_bufferSize", "failed");
    return KREG_E_CUDA;
  }
  if (cusolverDnXgetrf(ctx->handle, ctx->params, m, m, CUDA_R_64F, K, ldk, piv64, CUDA_R_64F, work, dws, host_workspace,
                       hws, d_info) != CUSOLVER_STATUS_SUCCESS) {
    set_last_error("cusolverDnXgetrf", "failed");
    return KREG_E_CUDA;
  }
  KREG_CUDA_TRY(cudaMemcpyAsync(&h_info, d_info, sizeof(int), cudaMemcpyDeviceToHost, st));
  KREG_CUDA_TRY(cudaStreamSynchronize(st));
  if (h_info != 0) return h_info > 0 ? h_info : KREG_E_BADARG;
  // the matrix is symmetric, so the column-major view cuSOLVER factorised is the same matrix: no transpose needed
  if (cusolverDnXgetrs(ctx->handle, ctx->params, CUBLAS_OP_N, m, 1, CUDA_R_64F, K, ldk, piv64, CUDA_R_64F, y_inout, m,
                       d_info) != CUSOLVER_STATUS_SUCCESS) {
    set_last_error("cusolverDnXgetrs", "failed");
    return KREG_E_CUDA;
  }
  return KREG_OK;
}
