// fp64_probe.cu -- measures the B200's FP64 issue rates that bound the KREG kernels:
//   DFMA throughput, DFMA dependent latency, DMMA (mma.sync f64) throughput for the three
//   shapes sm_90+ offers, DFMA+DMMA co-issue, and double-precision exp() throughput.
// MEASURED_PEAKS.json has no FP64 entry, so the roofline denominator for the
// prediction kernel comes from here.  Build: see tools/Makefile.  Run on the GPU box only.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <int CHAINS>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
  double acc[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; c++) acc[c] = threadIdx.x * 1e-9 + c;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int c = 0; c < CHAINS; c++) acc[c] = fma(acc[c], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) s += acc[c];
  if (s == 123.456) out[0] = s;
}

__global__ void dfma_latency_kernel(double* out, long long* cycles, int iters, double a, double b) {
  double acc = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 64; r++) acc = fma(acc, a, b);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[0] = t1 - t0;
  if (acc == 123.456) out[0] = acc;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double (&c)[4], const double (&a)[2], double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// SHAPE: 0 = m8n8k4, 1 = m16n8k4, 2 = m16n8k8, 3 = m16n8k16 ; CH independent accumulator sets
template <int SHAPE, int CH>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters) {
  double c[CH][4];
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = 1e-3 * (threadIdx.x + i);
#pragma unroll
  for (int i = 0; i < 4; i++) b[i] = 1e-3 * (threadIdx.x - i);
#pragma unroll
  for (int k = 0; k < CH; k++)
#pragma unroll
    for (int i = 0; i < 4; i++) c[k][i] = 0.0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int k = 0; k < CH; k++) {
        if (SHAPE == 0) dmma884(c[k][0], c[k][1], a[0], b[0]);
        if (SHAPE == 1) { double aa[2] = {a[0], a[1]}; dmma1684(c[k], aa, b[0]); }
        if (SHAPE == 2) { double aa[4] = {a[0], a[1], a[2], a[3]}; double bb[2] = {b[0], b[1]}; dmma1688(c[k], aa, bb); }
        if (SHAPE == 3) dmma16816(c[k], a, b);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < CH; k++)
#pragma unroll
    for (int i = 0; i < 4; i++) s += c[k][i];
  if (s == 123.456) out[0] = s;
}

// co-issue: per iteration NM m16n8k8 DMMAs (2048 flop each) + NF DFMAs per thread (64 flop/warp each)
template <int NM, int NF>
__global__ void __launch_bounds__(256) mixed_kernel(double* out, int iters, double fa, double fb) {
  double c[NM > 0 ? NM : 1][4];
  double acc[NF > 0 ? NF : 1];
  double a[4], b[2];
#pragma unroll
  for (int i = 0; i < 4; i++) a[i] = 1e-3 * (threadIdx.x + i);
  b[0] = 1e-3 * threadIdx.x; b[1] = 2e-3 * threadIdx.x;
#pragma unroll
  for (int k = 0; k < NM; k++)
#pragma unroll
    for (int i = 0; i < 4; i++) c[k][i] = 0.0;
#pragma unroll
  for (int k = 0; k < NF; k++) acc[k] = k + threadIdx.x * 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int k = 0; k < (NM > NF ? NM : NF); k++) {
        if (k < NM) dmma1688(c[k], a, b);
        if (k < NF) acc[k] = fma(acc[k], fa, fb);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < NM; k++)
#pragma unroll
    for (int i = 0; i < 4; i++) s += c[k][i];
#pragma unroll
  for (int k = 0; k < NF; k++) s += acc[k];
  if (s == 123.456) out[0] = s;
}

// exp throughput: MODE 0 = CUDA libm exp(), 1 = polynomial exp for x<=0 (the kernels' variant)
__device__ __forceinline__ double exp_neg_poly(double x) {
  // x <= 0.  n = rint(x*log2e), r = x - n*ln2 (two-part), p(r) degree-11 Taylor/Horner, scale by 2^n
  const double L2E = 1.4426950408889634074, LN2HI = 6.93147180369123816490e-01, LN2LO = 1.90821492927058770002e-10;
  double t = fma(x, L2E, 6755399441055744.0);
  int n = __double2loint(t);
  double fn = t - 6755399441055744.0;
  double r = fma(fn, -LN2HI, x);
  r = fma(fn, -LN2LO, r);
  double p = 2.5052108385441718775e-8;
  p = fma(p, r, 2.7557319223985890653e-7);
  p = fma(p, r, 2.7557319223985890653e-6);
  p = fma(p, r, 2.4801587301587301587e-5);
  p = fma(p, r, 1.9841269841269841270e-4);
  p = fma(p, r, 1.3888888888888888889e-3);
  p = fma(p, r, 8.3333333333333333333e-3);
  p = fma(p, r, 4.1666666666666666667e-2);
  p = fma(p, r, 1.6666666666666666667e-1);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  int hi = __double2hiint(p) + (n << 20);
  double res = __hiloint2double(hi, __double2loint(p));
  return (n < -1020) ? 0.0 : res;
}

template <int MODE, int CH>
__global__ void __launch_bounds__(256) exp_kernel(double* out, int iters, double step) {
  double x[CH], s[CH];
#pragma unroll
  for (int c = 0; c < CH; c++) { x[c] = -1e-3 * (threadIdx.x + c); s[c] = 0; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int c = 0; c < CH; c++) {
      s[c] += MODE == 0 ? exp(x[c]) : exp_neg_poly(x[c]);
      x[c] -= step;
    }
  }
  double t = 0;
#pragma unroll
  for (int c = 0; c < CH; c++) t += s[c];
  if (t == 123.456) out[0] = t;
}

template <typename F>
static float time_ms(F launch, int reps = 5) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  printf("device: %s  SMs=%d  cc=%d.%d  clock=%d MHz\n", prop.name, sms, prop.major, prop.minor, clk_khz / 1000);
  double* out; CK(cudaMalloc(&out, 64));
  long long* cyc; CK(cudaMalloc(&cyc, 8));
  const int blocks = sms * 8, threads = 256;
  const double nthreads = (double)blocks * threads;

  {
    int iters = 4000;
    float ms = time_ms([&] { dfma_kernel<8><<<blocks, threads>>>(out, iters, 0.999, 1e-3); });
    printf("DFMA  8 chains : %8.3f ms  %7.2f TFLOP/s\n", ms, 2.0 * nthreads * iters * 8 * 8 / ms * 1e-9);
    ms = time_ms([&] { dfma_kernel<16><<<blocks, threads>>>(out, iters, 0.999, 1e-3); });
    printf("DFMA 16 chains : %8.3f ms  %7.2f TFLOP/s\n", ms, 2.0 * nthreads * iters * 8 * 16 / ms * 1e-9);
    // sustained: ~2 s worth
    int big = 40000;
    ms = time_ms([&] { dfma_kernel<16><<<blocks, threads>>>(out, big, 0.999, 1e-3); }, 3);
    printf("DFMA 16 chains sustained (%.0f ms launches): %7.2f TFLOP/s\n", ms, 2.0 * nthreads * big * 8 * 16 / ms * 1e-9);
  }
  {
    int iters = 2000;
    dfma_latency_kernel<<<1, 32>>>(out, cyc, iters, 0.999, 1e-3);
    CK(cudaDeviceSynchronize());
    long long h; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("DFMA dependent latency: %.2f cycles\n", (double)h / (iters * 64.0));
  }
  {
    int iters = 2000;
    float ms;
    ms = time_ms([&] { dmma_kernel<0, 8><<<blocks, threads>>>(out, iters); });
    printf("DMMA m8n8k4   x8: %8.3f ms  %7.2f TFLOP/s\n", ms, (nthreads / 32) * iters * 4.0 * 8 * 512 / ms * 1e-9);
    ms = time_ms([&] { dmma_kernel<1, 8><<<blocks, threads>>>(out, iters); });
    printf("DMMA m16n8k4  x8: %8.3f ms  %7.2f TFLOP/s\n", ms, (nthreads / 32) * iters * 4.0 * 8 * 1024 / ms * 1e-9);
    ms = time_ms([&] { dmma_kernel<2, 8><<<blocks, threads>>>(out, iters); });
    printf("DMMA m16n8k8  x8: %8.3f ms  %7.2f TFLOP/s\n", ms, (nthreads / 32) * iters * 4.0 * 8 * 2048 / ms * 1e-9);
    ms = time_ms([&] { dmma_kernel<3, 8><<<blocks, threads>>>(out, iters); });
    printf("DMMA m16n8k16 x8: %8.3f ms  %7.2f TFLOP/s\n", ms, (nthreads / 32) * iters * 4.0 * 8 * 4096 / ms * 1e-9);
    ms = time_ms([&] { dmma_kernel<2, 2><<<blocks, threads>>>(out, iters); });
    printf("DMMA m16n8k8  x2: %8.3f ms  %7.2f TFLOP/s\n", ms, (nthreads / 32) * iters * 4.0 * 2 * 2048 / ms * 1e-9);
  }
  {
    int iters = 2000;
    float ms;
    ms = time_ms([&] { mixed_kernel<4, 0><<<blocks, threads>>>(out, iters, 0.999, 1e-3); });
    printf("mixed 4 DMMA16x8x8 + 0 DFMA : %8.3f ms  %7.2f TFLOP/s\n", ms, (nthreads / 32) * iters * 4.0 * (4 * 2048 + 0 * 64) / ms * 1e-9);
    ms = time_ms([&] { mixed_kernel<0, 16><<<blocks, threads>>>(out, iters, 0.999, 1e-3); });
    printf("mixed 0 DMMA + 16 DFMA      : %8.3f ms  %7.2f TFLOP/s\n", ms, (nthreads / 32) * iters * 4.0 * (0 * 2048 + 16 * 64) / ms * 1e-9);
    ms = time_ms([&] { mixed_kernel<4, 16><<<blocks, threads>>>(out, iters, 0.999, 1e-3); });
    printf("mixed 4 DMMA16x8x8 + 16 DFMA: %8.3f ms  %7.2f TFLOP/s (sum)\n", ms, (nthreads / 32) * iters * 4.0 * (4 * 2048 + 16 * 64) / ms * 1e-9);
    ms = time_ms([&] { mixed_kernel<1, 16><<<blocks, threads>>>(out, iters, 0.999, 1e-3); });
    printf("mixed 1 DMMA16x8x8 + 16 DFMA: %8.3f ms  %7.2f TFLOP/s (sum)\n", ms, (nthreads / 32) * iters * 4.0 * (1 * 2048 + 16 * 64) / ms * 1e-9);
    ms = time_ms([&] { mixed_kernel<2, 8><<<blocks, threads>>>(out, iters, 0.999, 1e-3); });
    printf("mixed 2 DMMA16x8x8 +  8 DFMA: %8.3f ms  %7.2f TFLOP/s (sum)\n", ms, (nthreads / 32) * iters * 4.0 * (2 * 2048 + 8 * 64) / ms * 1e-9);
  }
  {
    int iters = 4000;
    float ms;
    ms = time_ms([&] { exp_kernel<0, 4><<<blocks, threads>>>(out, iters, 1e-4); });
    printf("exp() libm      : %8.3f ms  %7.2f Gexp/s  (=%.1f DFMA-slots each at measured peak)\n", ms, nthreads * iters * 4 / ms * 1e-6, 0.0);
    ms = time_ms([&] { exp_kernel<1, 4><<<blocks, threads>>>(out, iters, 1e-4); });
    printf("exp_neg_poly    : %8.3f ms  %7.2f Gexp/s\n", ms, nthreads * iters * 4 / ms * 1e-6);
  }
  return 0;
}
