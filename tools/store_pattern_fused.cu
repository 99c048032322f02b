// Bare store pattern of grad_fused_kernel (kreg_kbuild.cu) for any molecule size: the same grid, CTA shape, row order and
// 8-byte-per-lane stores, NO arithmetic and no shared memory -- what the HBM write path allows for this access pattern.
//   tools/bin/store_pattern_fused            (sizes 9 12 16 21 24 32 48 64; M ~ 54 000 gradient rows, lower triangle)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void __launch_bounds__(256, 2) fused_pattern(double *K, long ld, int ntr, int nat, int TJ, int NIB) {
  const int G3 = 3 * nat;
  const long jb = (long)blockIdx.x * TJ;
  long i_lo = (long)blockIdx.y * NIB, i_hi = i_lo + NIB;
  if (i_hi > ntr) i_hi = ntr;
  if (i_lo < jb) i_lo = jb;
  if (i_lo >= i_hi) return;
  const int tid = threadIdx.x;
  if (tid >= TJ * G3) return;
  const int jl = tid / G3, cc = tid - jl * G3;
  const long j = jb + jl;
  if (j >= ntr) return;
  for (long i = i_lo; i < i_hi; i++) {
    if (j > i) continue;
    double *dst = K + (i * G3) * ld + j * G3 + cc;
    const double v = (double)(i + tid);
    for (int aa = 0; aa < nat; aa++) {
      dst[0] = v;
      dst[ld] = v + 1.0;
      dst[2 * ld] = v + 2.0;
      dst += 3 * ld;
    }
  }
}

// The kernel's own CTA shape (TJ whole geometries), but the thread index is shifted by the misalignment of the segment so
// that every warp's 32 columns start on a 256-byte boundary; the first and last warp of a segment are partial.
__global__ void __launch_bounds__(288, 1) shifted_pattern(double *K, long ld, int ntr, int nat, int TJ, int NIB) {
  const int G3 = 3 * nat;
  const long jb = (long)blockIdx.x * TJ;
  long i_lo = (long)blockIdx.y * NIB, i_hi = i_lo + NIB;
  if (i_hi > ntr) i_hi = ntr;
  if (i_lo < jb) i_lo = jb;
  if (i_lo >= i_hi) return;
  const int mis = (int)((jb * G3) & 31);
  const int tid = (int)threadIdx.x - mis;
  if (tid < 0 || tid >= TJ * G3) return;
  const int jl = tid / G3, cc = tid - jl * G3;
  const long j = jb + jl;
  if (j >= ntr) return;
  for (long i = i_lo; i < i_hi; i++) {
    if (j > i) continue;
    double *dst = K + (i * G3) * ld + j * G3 + cc;
    const double v = (double)(i + tid);
    for (int aa = 0; aa < nat; aa++) {
      dst[0] = v;
      dst[ld] = v + 1.0;
      dst[2 * ld] = v + 2.0;
      dst += 3 * ld;
    }
  }
}

// The same rows and steps, but a CTA owns SPAN consecutive columns starting at a multiple of SPAN (whole warps, every warp
// store on 128-byte lines), whatever geometries j those columns belong to.
__global__ void __launch_bounds__(256, 2) span_pattern(double *K, long ld, int ntr, int nat, int SPAN, int NIB) {
  const int G3 = 3 * nat;
  const long c = (long)blockIdx.x * SPAN + threadIdx.x;
  long i_lo = (long)blockIdx.y * NIB, i_hi = i_lo + NIB;
  if (i_hi > ntr) i_hi = ntr;
  const long jfirst = ((long)blockIdx.x * SPAN) / G3;  // first geometry j of the span: rows i < jfirst write nothing here
  if (i_lo < jfirst) i_lo = jfirst;
  if (i_lo >= i_hi || threadIdx.x >= SPAN) return;
  for (long i = i_lo; i < i_hi; i++) {
    if (c >= (i + 1) * G3) continue;  // lower triangle in whole 3 Nat blocks
    double *dst = K + (i * G3) * ld + c;
    const double v = (double)(i + c);
    for (int aa = 0; aa < nat; aa++) {
      dst[0] = v;
      dst[ld] = v + 1.0;
      dst[2 * ld] = v + 2.0;
      dst += 3 * ld;
    }
  }
}

__global__ void fill(double *K, size_t n) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) K[e] = 1.0;
}

int main() {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int sizes[] = {9, 12, 16, 21, 24, 28, 32, 48};
  for (int nat : sizes) {
    const int G3 = 3 * nat, ntr = 54000 / G3;
    const long M = (long)ntr * G3;
   for (long LDA : {16L, 32L}) {
    const long ld = (M + LDA - 1) / LDA * LDA;
    printf("row stride padded to a multiple of %ld doubles (ld = %ld):\n", LDA, ld);
    double *K;
    CK(cudaMalloc(&K, (size_t)M * ld * 8));
    int TJ = 256 / G3;
    if (TJ < 1) TJ = 1;
    if (TJ > 16) TJ = 16;
    if (nat <= 10 && TJ * G3 > 224) TJ = 224 / G3;  // the small-molecule emit kernel's strip (8 geometries at 9 atoms)
    const int NIB = 16, threads = ((TJ * G3 + 31) / 32) * 32;
    const double bytes = 4.0 * M * (M + G3);  // lower triangle in whole 3 Nat blocks
    float best = 1e9f, bestf = 1e9f;
    for (int r = 0; r < 5; r++) {
      CK(cudaEventRecord(e0));
      fused_pattern<<<dim3((ntr + TJ - 1) / TJ, (ntr + NIB - 1) / NIB), threads>>>(K, ld, ntr, nat, TJ, NIB);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (r && ms < best) best = ms;
      CK(cudaEventRecord(e0));
      fill<<<148 * 16, 256>>>(K, (size_t)(bytes / 8));
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (r && ms < bestf) bestf = ms;
    }
    CK(cudaGetLastError());
    printf("nat %2d  TJ %2d  threads %3d  M %6ld  pattern %7.3f ms  %7.1f GB/s   streaming fill of the same bytes %7.3f ms  %7.1f GB/s\n",
           nat, TJ, threads, M, best, bytes / best * 1e-6, bestf, bytes / bestf * 1e-6);
    for (int tj = TJ; tj >= TJ - 1 && tj >= 1; tj--) {
      const int thr = ((tj * G3 + 31 + 31) / 32) * 32;
      if (thr > 288) continue;
      float bs = 1e9f;
      for (int r = 0; r < 4; r++) {
        CK(cudaEventRecord(e0));
        shifted_pattern<<<dim3((ntr + tj - 1) / tj, (ntr + NIB - 1) / NIB), thr>>>(K, ld, ntr, nat, tj, NIB);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r && ms < bs) bs = ms;
      }
      CK(cudaGetLastError());
      printf("        geometry-aligned CTA, thread index shifted to 256-byte warp stores, TJ %d, %3d threads: %7.3f ms  %7.1f GB/s\n", tj, thr, bs, bytes / bs * 1e-6);
    }
    for (int SPAN : {256}) {
      float bs = 1e9f;
      for (int r = 0; r < 4; r++) {
        CK(cudaEventRecord(e0));
        span_pattern<<<dim3((unsigned)((M + SPAN - 1) / SPAN), (ntr + NIB - 1) / NIB), SPAN>>>(K, ld, ntr, nat, SPAN, NIB);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r && ms < bs) bs = ms;
      }
      CK(cudaGetLastError());
      printf("        aligned span of %3d columns per CTA: %7.3f ms  %7.1f GB/s\n", SPAN, bs, bytes / bs * 1e-6);
    }
    CK(cudaFree(K));
   }
  }
  return 0;
}
