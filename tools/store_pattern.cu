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
  run("v2 streaming fill, same bytes", [&] { v2<<<148 * 16, 256>>>(K, (size_t)(bytes / 8)); });
  return 0;
}
