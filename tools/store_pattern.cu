// store_pattern.cu -- how fast can the gradient-gradient block WRITE PATTERN go with no compute at all?
// Variant 0: the grad_blocks_kernel pattern (CTA = 8 geometries i x 9 geometries j, thread = column, walks 216 rows).
// Variant 1: same bytes, but a CTA owns 27 rows x (TJ*27*NJB) columns (longer runs per row, fewer rows in flight).
// Variant 2: plain streaming fill of the same number of bytes (upper bound).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__global__ void __launch_bounds__(256) v0(double *K, long ld, int ntr, int TI) {
  const int G3 = 27, TJ = 9;
  long ib = (long)blockIdx.y * TI, jb = (long)blockIdx.x * TJ;
  if (jb > ib + TI - 1) return;
  int tid = threadIdx.x; if (tid >= TJ * G3) return;
  int jl = tid / G3, col = tid % G3; long j = jb + jl;
  if (j >= ntr) return;
  for (int il = 0; il < TI; il++) {
    long i = ib + il; if (i >= ntr || j > i) continue;
    double *dst = K + (i * G3) * ld + j * G3 + col;
    double v = (double)(il + tid);
#pragma unroll
    for (int q = 0; q < G3; q++) { *dst = v + q; dst += ld; }
  }
}

// CTA = 1 geometry-row block i (27 rows) x NJB j-blocks of 9 geometries; threads sweep columns, rows outer
__global__ void __launch_bounds__(256) v1(double *K, long ld, int ntr, int NJB) {
  const int G3 = 27;
  long i = blockIdx.y; long c0 = (long)blockIdx.x * NJB * 243; long cmax = (i + 1) * G3;  // columns < cmax
  if (c0 >= cmax) return;
  long c1 = c0 + (long)NJB * 243; if (c1 > cmax) c1 = cmax;
  for (int q = 0; q < G3; q++) {
    double *row = K + (i * G3 + q) * ld;
    for (long c = c0 + threadIdx.x; c < c1; c += 256) row[c] = (double)(q + c);
  }
}

// v0 with the grid transposed: consecutive CTAs walk DOWN a column band (i-block fastest)
__global__ void __launch_bounds__(256) v0t(double *K, long ld, int ntr, int TI) {
  const int G3 = 27, TJ = 9;
  long ib = (long)blockIdx.x * TI, jb = (long)blockIdx.y * TJ;
  if (jb > ib + TI - 1) return;
  int tid = threadIdx.x; if (tid >= TJ * G3) return;
  int jl = tid / G3, col = tid % G3; long j = jb + jl;
  if (j >= ntr) return;
  for (int il = 0; il < TI; il++) {
    long i = ib + il; if (i >= ntr || j > i) continue;
    double *dst = K + (i * G3) * ld + j * G3 + col;
    double v = (double)(il + tid);
#pragma unroll
    for (int q = 0; q < G3; q++) { *dst = v + q; dst += ld; }
  }
}

// CTA block of TI geometries x 9 geometries written row by row by all threads (rows outer, columns inner)
__global__ void __launch_bounds__(256) v3(double *K, long ld, int ntr, int TI) {
  const int G3 = 27, TJ = 9;
  long ib = (long)blockIdx.y * TI, jb = (long)blockIdx.x * TJ;
  if (jb > ib + TI - 1) return;
  for (int il = 0; il < TI; il++) {
    long i = ib + il; if (i >= ntr) continue;
    long c0 = jb * G3, c1 = c0 + TJ * G3, cmax = (i + 1) * G3; if (c1 > cmax) c1 = cmax;
    for (int q = 0; q < G3; q++) {
      double *row = K + (i * G3 + q) * ld;
      for (long c = c0 + threadIdx.x; c < c1; c += 256) row[c] = (double)(q + c);
    }
  }
}


// v0 TI=1 with cache-hinted stores: MODE 0 default, 1 st.global.cs (streaming / evict-first), 2 st.global.wt, 3 st.global.cg
template <int MODE>
__device__ __forceinline__ void st_hint(double *p, double v) {
  if (MODE == 1) __stcs(p, v);
  else if (MODE == 2) __stwt(p, v);
  else if (MODE == 3) __stcg(p, v);
  else *p = v;
}
template <int MODE, int NJB>
__global__ void __launch_bounds__(256) v4(double *K, long ld, int ntr) {
  const int G3 = 27, TJ = 9;
  long i = blockIdx.y;
  int tid = threadIdx.x; if (tid >= TJ * G3) return;
  for (int n = 0; n < NJB; n++) {
    long jb = ((long)blockIdx.x * NJB + n) * TJ;
    if (jb > i) return;
    int jl = tid / G3, col = tid % G3; long j = jb + jl;
    if (j > i) continue;
    double *dst = K + (i * G3) * ld + j * G3 + col;
    double v = (double)(n + tid);
#pragma unroll
    for (int q = 0; q < G3; q++) { st_hint<MODE>(dst, v + q); dst += ld; }
  }
}
// two adjacent columns per thread (16-byte stores where the row offset allows, i.e. even element index)
template <int MODE>
__global__ void __launch_bounds__(128) v5(double *K, long ld, int ntr, int NJB) {
  const int G3 = 27;
  long i = blockIdx.y; long c0 = (long)blockIdx.x * NJB * 244; long cmax = (i + 1) * G3;
  for (int n = 0; n < NJB; n++) {
    long c = c0 + (long)n * 244 + 2 * threadIdx.x;
    if (threadIdx.x >= 122 || c >= cmax) continue;
    for (int q = 0; q < G3; q++) {
      double *row = K + (i * G3 + q) * ld;
      if (c + 1 < cmax) {
        double2 v = make_double2((double)q, (double)c);
        if (MODE == 1) __stcs(reinterpret_cast<double2 *>(row + c), v); else *reinterpret_cast<double2 *>(row + c) = v;
      } else row[c] = 1.0;
    }
  }
}

// v6: thread = column computes 9 rows at a time into a shared ring buffer; one thread hands each 9-row x 216-column group
// to the TMA engine as 9 bulk stores (cp.async.bulk.global.shared::cta), three groups in flight
__device__ __forceinline__ unsigned s_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int NJB>
__global__ void __launch_bounds__(224, 2) v6(double *K, long ld, int ntr) {
  const int G3 = 27, TJ = 8, COLS = TJ * G3, RG = 9;
  extern __shared__ __align__(128) double ring[];  // [3][RG][COLS]
  long i = blockIdx.y;
  int tid = threadIdx.x;
  int grp = 0;
  for (int n = 0; n < NJB; n++) {
    long jb = ((long)blockIdx.x * NJB + n) * TJ;
    if (jb + TJ - 1 >= i) {  // ragged / diagonal step: direct stores
      if (jb > i) return;
      int jl = tid / G3; long j = jb + jl;
      if (tid < COLS && j <= i) {
        double *dst = K + (i * G3) * ld + jb * G3 + tid;
        for (int q = 0; q < G3; q++) { *dst = (double)(q + tid); dst += ld; }
      }
      continue;
    }
    for (int g = 0; g < 3; g++, grp++) {
      double *buf = ring + (grp % 3) * RG * COLS;
      if (tid < COLS) {
#pragma unroll
        for (int r = 0; r < RG; r++) buf[r * COLS + tid] = (double)(r + tid + n);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      __syncthreads();
      if (tid == 0) {
        double *dst = K + (i * G3 + g * RG) * ld + jb * G3;
        for (int r = 0; r < RG; r++)
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + r * ld),
                       "r"(s_u32(buf + r * COLS)), "r"(COLS * 8)
                       : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// v7: same staging, copy-out by all threads with 16-byte stores
template <int NJB>
__global__ void __launch_bounds__(224, 2) v7(double *K, long ld, int ntr) {
  const int G3 = 27, TJ = 8, COLS = TJ * G3, RG = 9;
  extern __shared__ __align__(128) double ring[];  // [RG][COLS]
  long i = blockIdx.y;
  int tid = threadIdx.x;
  for (int n = 0; n < NJB; n++) {
    long jb = ((long)blockIdx.x * NJB + n) * TJ;
    if (jb + TJ - 1 >= i) {
      if (jb > i) return;
      int jl = tid / G3; long j = jb + jl;
      if (tid < COLS && j <= i) {
        double *dst = K + (i * G3) * ld + jb * G3 + tid;
        for (int q = 0; q < G3; q++) { *dst = (double)(q + tid); dst += ld; }
      }
      continue;
    }
    for (int g = 0; g < 3; g++) {
      __syncthreads();
      if (tid < COLS) {
#pragma unroll
        for (int r = 0; r < RG; r++) ring[r * COLS + tid] = (double)(r + tid + n);
      }
      __syncthreads();
      double *dst = K + (i * G3 + g * RG) * ld + jb * G3;
      for (int e = tid; e < RG * COLS / 2; e += 224) {
        int r = e / (COLS / 2), c2 = e % (COLS / 2);
        *reinterpret_cast<double2 *>(dst + r * ld + 2 * c2) = *reinterpret_cast<const double2 *>(ring + r * COLS + 2 * c2);
      }
    }
  }
}

// v8: thread = column, TJ geometries per step (COLS = 27 TJ), 8-byte stores; v9: lane pairs, even lane stores row q and
// odd lane row q+1 as 16-byte words (what a shuffle-based epilogue would issue)
template <int TJ, int NJB, bool PAIR>
__global__ void __launch_bounds__(256) v8(double *K, long ld, int ntr) {
  const int G3 = 27, COLS = TJ * G3;
  long i = blockIdx.y;
  int tid = threadIdx.x; if (tid >= COLS) return;
  for (int n = 0; n < NJB; n++) {
    long jb = ((long)blockIdx.x * NJB + n) * TJ;
    if (jb > i) return;
    int jl = tid / G3; long j = jb + jl;
    bool strip = jb + TJ - 1 < i;
    if (PAIR && strip) {
      bool odd = tid & 1;
      double *dst = K + (i * G3) * ld + jb * G3 + (tid & ~1) + (odd ? ld : 0);
      double v = (double)(n + tid);
#pragma unroll
      for (int q2 = 0; q2 < 13; q2++) *reinterpret_cast<double2 *>(dst + (long)(2 * q2) * ld) = make_double2(v + q2, v - q2);
      K[(i * G3 + 26) * ld + jb * G3 + tid] = v;
    } else {
      if (j > i) continue;
      double *dst = K + (i * G3) * ld + j * G3 + (tid % G3);
      double v = (double)(n + tid);
#pragma unroll
      for (int q = 0; q < G3; q++) { *dst = v + q; dst += ld; }
    }
  }
}

__global__ void __launch_bounds__(256) v2(double *K, size_t n) {
  for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < n; e += (size_t)gridDim.x * 256) K[e] = 1.0;
}

int main() {
  const int ntr = 2000, G3 = 27; const long M = (long)ntr * G3, ld = M;
  double *K; CK(cudaMalloc(&K, (size_t)M * M * 8));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const double bytes = 4.0 * M * (M + 1);
  auto run = [&](const char *name, auto f) {
    float best = 1e9;
    for (int r = 0; r < 5; r++) { CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (r && ms < best) best = ms; }
    CK(cudaGetLastError());
    printf("%-44s %8.3f ms  %8.1f GB/s\n", name, best, bytes / best * 1e-6);
  };
  for (int TI : {1, 2}) {
    char nm[64]; snprintf(nm, 64, "v0 grad pattern TI=%d", TI);
    run(nm, [&] { v0<<<dim3((ntr + 8) / 9, (ntr + TI - 1) / TI), 256>>>(K, ld, ntr, TI); });
  }
  for (int TI : {1, 8}) {
    char nm[64]; snprintf(nm, 64, "v0t transposed grid TI=%d", TI);
    run(nm, [&] { v0t<<<dim3((ntr + TI - 1) / TI, (ntr + 8) / 9), 256>>>(K, ld, ntr, TI); });
  }
  for (int TI : {1, 4, 8}) {
    char nm[64]; snprintf(nm, 64, "v3 row-by-row block TI=%d", TI);
    run(nm, [&] { v3<<<dim3((ntr + 8) / 9, (ntr + TI - 1) / TI), 256>>>(K, ld, ntr, TI); });
  }
  for (int TI : {4, 8, 16}) {
    char nm[64]; snprintf(nm, 64, "v0 grad pattern TI=%d", TI);
    run(nm, [&] { v0<<<dim3((ntr + 8) / 9, (ntr + TI - 1) / TI), 256>>>(K, ld, ntr, TI); });
  }
  for (int NJB : {1, 4, 16, 64}) {
    char nm[64]; snprintf(nm, 64, "v1 27 rows x %d cols per CTA", NJB * 243);
    run(nm, [&] { v1<<<dim3((unsigned)((M + NJB * 243 - 1) / (NJB * 243)), ntr), 256>>>(K, ld, ntr, NJB); });
  }
  run("v4 thread=column, 1 step, default stores", [&] { v4<0, 1><<<dim3((ntr + 8) / 9, ntr), 256>>>(K, ld, ntr); });
  run("v4 thread=column, 1 step, st.cs", [&] { v4<1, 1><<<dim3((ntr + 8) / 9, ntr), 256>>>(K, ld, ntr); });
  run("v4 thread=column, 1 step, st.wt", [&] { v4<2, 1><<<dim3((ntr + 8) / 9, ntr), 256>>>(K, ld, ntr); });
  run("v4 thread=column, 1 step, st.cg", [&] { v4<3, 1><<<dim3((ntr + 8) / 9, ntr), 256>>>(K, ld, ntr); });
  run("v4 thread=column, 8 steps, default stores", [&] { v4<0, 8><<<dim3((ntr + 71) / 72, ntr), 256>>>(K, ld, ntr); });
  run("v4 thread=column, 8 steps, st.cs", [&] { v4<1, 8><<<dim3((ntr + 71) / 72, ntr), 256>>>(K, ld, ntr); });
  run("v4 thread=column, 8 steps, st.wt", [&] { v4<2, 8><<<dim3((ntr + 71) / 72, ntr), 256>>>(K, ld, ntr); });
  run("v5 two columns per thread (16 B), 1 step", [&] { v5<0><<<dim3((unsigned)((M + 243) / 244), ntr), 128>>>(K, ld, ntr, 1); });
  run("v5 two columns per thread (16 B), 8 steps", [&] { v5<0><<<dim3((unsigned)((M + 8 * 244 - 1) / (8 * 244)), ntr), 128>>>(K, ld, ntr, 8); });
  run("v5 two columns per thread (16 B), 8 steps, st.cs", [&] { v5<1><<<dim3((unsigned)((M + 8 * 244 - 1) / (8 * 244)), ntr), 128>>>(K, ld, ntr, 8); });
  CK(cudaFuncSetAttribute(v6<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 9 * 216 * 8));
  CK(cudaFuncSetAttribute(v6<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 9 * 216 * 8));
  run("v6 smem ring + TMA bulk stores, 1 step", [&] { v6<1><<<dim3((ntr + 7) / 8, ntr), 224, 3 * 9 * 216 * 8>>>(K, ld, ntr); });
  run("v6 smem ring + TMA bulk stores, 8 steps", [&] { v6<8><<<dim3((ntr + 63) / 64, ntr), 224, 3 * 9 * 216 * 8>>>(K, ld, ntr); });
  run("v7 smem stage + 16 B thread stores, 1 step", [&] { v7<1><<<dim3((ntr + 7) / 8, ntr), 224, 9 * 216 * 8>>>(K, ld, ntr); });
  run("v7 smem stage + 16 B thread stores, 8 steps", [&] { v7<8><<<dim3((ntr + 63) / 64, ntr), 224, 9 * 216 * 8>>>(K, ld, ntr); });
  run("v8 thread=column TJ=9, 8 steps, 8 B", [&] { v8<9, 8, false><<<dim3((ntr + 71) / 72, ntr), 256>>>(K, ld, ntr); });
  run("v8 thread=column TJ=8, 8 steps, 8 B", [&] { v8<8, 8, false><<<dim3((ntr + 63) / 64, ntr), 224>>>(K, ld, ntr); });
  run("v8 thread=column TJ=8, 1 step, 8 B", [&] { v8<8, 1, false><<<dim3((ntr + 7) / 8, ntr), 224>>>(K, ld, ntr); });
  run("v9 lane pairs TJ=8, 8 steps, 16 B (2 rows x 256 B per warp store)", [&] { v8<8, 8, true><<<dim3((ntr + 63) / 64, ntr), 224>>>(K, ld, ntr); });
  run("v9 lane pairs TJ=8, 1 step, 16 B", [&] { v8<8, 1, true><<<dim3((ntr + 7) / 8, ntr), 224>>>(K, ld, ntr); });
  run("v9 lane pairs TJ=8, 2 steps, 16 B", [&] { v8<8, 2, true><<<dim3((ntr + 15) / 16, ntr), 224>>>(K, ld, ntr); });
  run("v2 streaming fill, same bytes", [&] { v2<<<148 * 16, 256>>>(K, (size_t)(bytes / 8)); });
  return 0;
}
