// kreg_common.cuh -- device helpers shared by the KREG sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/kreg_b200.h"

namespace kreg {

// ---- host-side error plumbing -------------------------------------------------------------
void set_last_error(const char *what, const char *detail);

#define KREG_CUDA_TRY(expr)                                   \
  do {                                                        \
    cudaError_t err__ = (expr);                               \
    if (err__ != cudaSuccess) {                               \
      kreg::set_last_error(#expr, cudaGetErrorString(err__)); \
      return KREG_E_CUDA;                                     \
    }                                                         \
  } while (0)

#define KREG_LAUNCH_CHECK() KREG_CUDA_TRY(cudaGetLastError())

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- descriptor indexing (KREG.f90:34-52 get_ac2dArray, 0-based) ---------------------------
__host__ __device__ constexpr int descr_size(int nat) { return nat * (nat - 1) / 2; }
__host__ __device__ constexpr int pair_index(int nat, int a, int b) {  // requires a < b
  return (2 * nat - a - 1) * a / 2 + (b - a) - 1;
}

// inverse of pair_index: descriptor element d -> atoms a < b (closed form + integer fix-up; no table, any Natoms)
__host__ __device__ inline void pair_from_index(int nat, int d, int &a, int &b) {
  const double t = 2.0 * nat - 1.0;
  int aa = (int)((t - sqrt(t * t - 8.0 * d)) * 0.5);
  if (aa < 0) aa = 0;
  if (aa > nat - 2) aa = nat - 2;
  while (aa > 0 && (2 * nat - aa - 1) * aa / 2 > d) aa--;
  while (aa < nat - 2 && (2 * nat - aa - 2) * (aa + 1) / 2 <= d) aa++;
  a = aa;
  b = d - (2 * nat - aa - 1) * aa / 2 + aa + 1;
}

// Squared distance exactly as the reference's Rij (KREG.f90:84): (dx^2 + dy^2) + dz^2 with every
// operation individually rounded (no FMA contraction), so descriptors are bit-identical.
__device__ __forceinline__ double rij2_exact(const double *pa, const double *pb) {
  double dx = __dsub_rn(pa[0], pb[0]), dy = __dsub_rn(pa[1], pb[1]), dz = __dsub_rn(pa[2], pb[2]);
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// ---- FP64 tensor-core tile: D(8x8) += A(8x4) * B(4x8)  (SASS: DMMA.8x8x4) ------------------
// Fragment ownership for lane l (g = l / 4, q = l % 4):
//   a = A[g][q]     b = B[q][g]     c[0], c[1] = C[g][2q], C[g][2q + 1]
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c[0]), "+d"(c[1])
      : "d"(a), "d"(b));
}

// ---- exp for the Gaussian kernel ----------------------------------------------------------
// Evaluates exp(x) given y = x * 64/ln2 (the callers fold that scale into their FMA constants).
//   y = 64 m + j + f,  j in [0,64), |f| <= 1/2  =>  exp(x) = 2^m * 2^(j/64) * exp(f ln2/64)
// 2^(j/64) comes from a 64-entry table (shared memory), exp(f ln2/64) - 1 from a degree-5
// polynomial (truncation 3.5e-17).  11 FP64-pipe instructions including the caller's FMA, versus
// ~20 for libm exp (profiles/r01_fp64_probe.txt).  Arguments x < -704 return 0.
constexpr double kExpScale = 92.332482616893658071;  // 64 / ln 2
__device__ __forceinline__ double exp_from_scaled(double y, const double *__restrict__ tab64) {
  constexpr double MAGIC = 6755399441055744.0;  // 1.5 * 2^52: adds with round-to-nearest-integer
  constexpr double C1 = 1.0830424696249145e-02;  // ln2/64
  constexpr double C2 = 5.8649049550561697e-05;  // (ln2/64)^2 / 2
  constexpr double C3 = 2.1173137155464775e-07;  // (ln2/64)^3 / 6
  constexpr double C4 = 5.7328516886404021e-10;  // (ln2/64)^4 / 24
  constexpr double C5 = 1.2417843701716925e-12;  // (ln2/64)^5 / 120
  double t = y + MAGIC;
  int n = __double2loint(t);
  double f = y - (t - MAGIC);
  double p = fma(C5, f, C4);
  p = fma(p, f, C3);
  p = fma(p, f, C2);
  p = fma(p, f, C1);
  p = p * f;
  double T = tab64[n & 63];
  double r = fma(T, p, T);
  int m = n >> 6;
  r = __hiloint2double(__double2hiint(r) + m * 1048576, __double2loint(r));
  // y < -65000 (x < -704): the true value is below 1e-305; also keeps n inside int32 and the
  // exponent arithmetic above inside the normal range.  Integer compare on the sign-magnitude bits.
  bool underflow = static_cast<unsigned>(__double2hiint(y)) > 0xC0EFBD00u;
  return underflow ? 0.0 : r;
}

// Same function with the round-to-integer done by the conversion unit (F2I / I2F) instead of two FP64 adds
// against 1.5*2^52: 2 fewer instructions on the FP64 pipe, which is the one the prediction kernel saturates.
__device__ __forceinline__ double exp_from_scaled_cvt(double y, const double *__restrict__ tab64) {
  constexpr double C1 = 1.0830424696249145e-02, C2 = 5.8649049550561697e-05, C3 = 2.1173137155464775e-07;
  constexpr double C4 = 5.7328516886404021e-10, C5 = 1.2417843701716925e-12;
  const bool underflow = static_cast<unsigned>(__double2hiint(y)) > 0xC0EFBD00u;  // y < -65000, -inf, NaN(-)
  const int n = __double2int_rn(underflow ? 0.0 : y);
  const double f = y - __int2double_rn(n);
  double p = fma(C5, f, C4);
  p = fma(p, f, C3);
  p = fma(p, f, C2);
  p = fma(p, f, C1);
  p = p * f;
  const double T = tab64[n & 63];
  double r = fma(T, p, T);
  r = __hiloint2double(__double2hiint(r) + (n >> 6) * 1048576, __double2loint(r));
  return underflow ? 0.0 : r;
}

__device__ __forceinline__ void fill_exp_table(double *tab64, int tid, int nthreads) {
  for (int j = tid; j < 64; j += nthreads) tab64[j] = exp2((double)j * (1.0 / 64.0));
}

// ---- debug build: every bulk-copy size / alignment and every guarded global store is checked --------------------
// compute-sanitizer is closed on this GPU pool; `make debug` builds libkreg_b200_dbg.so with -DKREG_DEBUG_BOUNDS, the
// parity tests are run against it once per round (KREG_LIB=...), and kreg_debug_bounds_failures() must stay 0.
// Failure codes: 1 bulk copy size not a positive multiple of 16; 2 bulk copy address misaligned; 3 store outside the
// extent the caller declared; 4 shared-memory destination outside the CTA's dynamic allocation.
#ifdef KREG_DEBUG_BOUNDS
extern __device__ unsigned long long g_debug_failures;
extern __device__ int g_debug_first_code;
extern __device__ int g_debug_first_line;
__device__ __forceinline__ void debug_fail(int code, int line) {
  if (atomicAdd(&g_debug_failures, 1ull) == 0ull) {
    g_debug_first_code = code;
    g_debug_first_line = line;
  }
}
#define KREG_CHECK(cond, code)               \
  do {                                       \
    if (!(cond)) kreg::debug_fail(code, __LINE__); \
  } while (0)
// store through a pointer that must lie inside [base, base + count) doubles
#define KREG_CHECK_STORE(ptr, base, count) \
  KREG_CHECK((ptr) >= (base) && (ptr) < (base) + (count), 3)
#else
#define KREG_CHECK(cond, code) ((void)0)
#define KREG_CHECK_STORE(ptr, base, count) ((void)0)
#endif

// ---- mbarrier + bulk async copy (TMA engine, SASS UBLKCP) ------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// non-blocking phase test: true once the phase with this parity has completed
__device__ __forceinline__ bool mbar_test(uint64_t *bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, uint64_t *bar) {
#ifdef KREG_DEBUG_BOUNDS
  {
    unsigned tot_size;
    asm("mov.u32 %0, %%total_smem_size;" : "=r"(tot_size));
    KREG_CHECK(bytes > 0 && (bytes & 15u) == 0, 1);
    KREG_CHECK((reinterpret_cast<uintptr_t>(src_gmem) & 15) == 0 && (smem_u32(dst_smem) & 15u) == 0, 2);
    // shared-window addresses of user data start after the 1 KB the system reserves per CTA on sm_90+
    // (cudaDevAttrReservedSharedMemoryPerBlock), which %total_smem_size does not include
    KREG_CHECK(smem_u32(dst_smem) + bytes <= tot_size + 1024u, 4);
  }
#endif
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

}  // namespace kreg
