// kreg_predict.cu -- fused KREG energy + analytic-gradient prediction on FP64 tensor cores.
//
// Replaces KREG.kreg.predict (mlatom/fortran/KREG.f90:534-596 with Yest_KRR :348-366 and
// YgradXYZest_KRR :368-391) and MLatomF's calcEst_KRR (A_KRR.f90:985-1029; "Optimized Code"
// A_KRR_kernel.f90:327-412).  Algebra, with mean-centred descriptors X':
//     k_ij = exp(-|X_i - X_j|^2 / 2s^2),  |X_i - X_j|^2 = n_i + n_j - 2 X'_i.X'_j
//     c_ij = alpha_j + (X_i - X_j).v_j / s^2          (v_j = J_j alpha_grad,j; 0 for value models)
//     E_i  = prior + sum_j k_ij c_ij
//     g_i  = [ sum_j w_ij X'_j - X'_i sum_j w_ij + sum_j k_ij v_j ] / s^2,   w_ij = k_ij c_ij
//     dE_i/dM_i = J_i^T g_i
// which is "attention without softmax": S = Xtest.Xtrain^T (GEMM1), an elementwise exp, and
// A += W.[Xtrain | 1] (GEMM2; the ones column accumulates sum_j w_ij = E_i for free).  Both GEMMs run
// on DMMA.8x8x4; a warp keeps its test-row fragments and the A accumulators in registers for the
// whole pass over the training set, and the training set arrives pre-swizzled in fragment order
// through a TMA bulk-copy ring in shared memory.  Energies and gradients come out of ONE pass (the
// reference makes two and recomputes k_ij 3*Nat+1 times).
#include "kreg_common.cuh"
#include "kreg_gemm.h"
#include "kreg_internal.h"
#include "kreg_predict_args.h"

namespace kreg {

// ------------------------------------------------------------------------------------------
// v_j = J_j alpha_grad,j  (J from KREG.f90:121,146)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) precompute_v_kernel(int nat, int64_t ntrain, const double *__restrict__ xyz,
                                                           const double *__restrict__ xeq,
                                                           const double *__restrict__ alpha_g,
                                                           double *__restrict__ V) {
  const int D = descr_size(nat);
  const int64_t total = ntrain * D;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t j = e / D;
    int d = (int)(e - j * D);
    int a, c;
    pair_from_index(nat, d, a, c);
    const double *m = xyz + j * nat * 3;
    const double *ag = alpha_g + j * nat * 3;
    double r2 = rij2_exact(m + 3 * a, m + 3 * c);
    double x = xeq[d] / sqrt(r2);
    double s = 0.0;
#pragma unroll
    for (int t = 0; t < 3; t++) s += (m[3 * c + t] - m[3 * a + t]) * (ag[3 * a + t] - ag[3 * c + t]);
    V[e] = x * s / r2;
  }
}

// ------------------------------------------------------------------------------------------
// Model packing: one warp per group of 8 training points writes the group's chunk
//   [B1x | (B1v) | B2x | (B2v) | scal]
//   B1x[p][lane][e]  = X'[8g + lane/4][4(2p+e) + lane%4]              GEMM1 B fragments, k-step 2p+e
//   B1v              = same for v_j / s^2
//   B2x[nt][lane][s] = Xaug[8g + 2(lane%4) + s][8nt + lane/4]         GEMM2 B fragments (Xaug[.][D] = 1)
//   B2v              = same for v_j (no ones column)
//   scal[q][4]       = { a'(2q), b(2q), a'(2q+1), b(2q+1) },  a'_j = alpha_j - X'_j.v_j/s^2,
//                      b_j = -n_j/(2 s^2) * 64/ln2
// The k index of GEMM2 is the training point; fragment s of a lane pairs W's own accumulator element
// s (column 2q+s) with training point 2q+s, so W never leaves the registers it was computed in.
// ------------------------------------------------------------------------------------------
template <bool HAS_V>
__global__ void __launch_bounds__(32) pack_model_kernel(int nat, int64_t ntrain, const double *__restrict__ xyz,
                                                        const double *__restrict__ xeq,
                                                        const double *__restrict__ center,
                                                        const double *__restrict__ alpha,
                                                        const double *__restrict__ V, double c1, double inv_s2,
                                                        double *__restrict__ packed,
                                                        const double *__restrict__ Xin /* optional [ntrain][D] */,
                                                        int raw_alpha = 0 /* radial-kernel models: alpha_j as it is */) {
  extern __shared__ double sm[];
  const TcShape sh(nat);
  const int TS = sh.NT * 8 > sh.KS * 4 ? sh.NT * 8 : sh.KS * 4;
  double *Xs = sm;            // [8][TS]
  double *Vs = sm + 8 * TS;   // [8][TS]
  double *nj = Vs + 8 * TS;   // [8]
  double *xv = nj + 8;        // [8]
  const int lane = threadIdx.x, g = lane >> 2, q = lane & 3;
  const int64_t grp = blockIdx.x;

  for (int i = lane; i < 16 * TS; i += 32) sm[i] = 0.0;
  __syncwarp();
  for (int a = 0; a < nat; a++)
    for (int b = a + 1; b < nat; b++) {
      int d = pair_index(nat, a, b);
      if (lane < 8) {
        int64_t j = grp * 8 + lane;
        if (j < ntrain) {
          if (Xin) {
            Xs[lane * TS + d] = Xin[j * sh.D + d] - center[d];
          } else {
            const double *m = xyz + j * nat * 3;
            Xs[lane * TS + d] = xeq[d] / sqrt(rij2_exact(m + 3 * a, m + 3 * b)) - center[d];
          }
          if (HAS_V) Vs[lane * TS + d] = V[j * sh.D + d];
        }
      }
    }
  __syncwarp();
  if (lane < 8) {
    double n = 0.0, s = 0.0;
    for (int d = 0; d < sh.D; d++) {
      n += Xs[lane * TS + d] * Xs[lane * TS + d];
      if (HAS_V) s += Xs[lane * TS + d] * Vs[lane * TS + d];
    }
    nj[lane] = n;
    xv[lane] = s;
  }
  __syncwarp();

  double *out = packed + grp * sh.chunk_doubles(HAS_V);
  // GEMM1 fragments
  for (int p = 0; p < sh.KS2; p++)
    for (int e = 0; e < 2; e++) {
      int ks = 2 * p + e;
      double vx = 0.0, vv = 0.0;
      if (ks < sh.KS) {
        vx = Xs[g * TS + 4 * ks + q];
        if (HAS_V) vv = Vs[g * TS + 4 * ks + q] * inv_s2;
      }
      out[(p * 32 + lane) * 2 + e] = vx;
      if (HAS_V) out[64 * sh.KS2 + (p * 32 + lane) * 2 + e] = vv;
    }
  // GEMM2 fragments
  double *o2x = out + 64 * sh.KS2 * (HAS_V ? 2 : 1);
  double *o2v = o2x + 64 * sh.NT;
  for (int nt = 0; nt < sh.NT; nt++)
    for (int s = 0; s < 2; s++) {
      int jl = 2 * q + s, col = 8 * nt + g;
      double vx = col < sh.D ? Xs[jl * TS + col] : (col == sh.D ? 1.0 : 0.0);
      o2x[(nt * 32 + lane) * 2 + s] = vx;
      if (HAS_V) o2v[(nt * 32 + lane) * 2 + s] = col < sh.D ? Vs[jl * TS + col] : 0.0;
    }
  double *osc = o2x + 64 * sh.NT * (HAS_V ? 2 : 1);
  if (lane < 8) {
    // scal[q][4] = { x(2q), s(2q), x(2q+1), s(2q+1) }:  s_j initialises the GEMM1 accumulator so that
    // c1 * (s_j + X'_i.X'_j) + bias_i is the scaled exponent (no per-element add in the kernel).
    //   gradient models: x_j = a'_j = alpha_j - X'_j.v_j/s^2,  s_j = -n_j/2
    //   value models   : x_j = +0.0 / -0.0 (sign of alpha_j),  s_j = -n_j/2 + s^2 ln|alpha_j|  -> w_ij comes out of
    //                    the exponential already multiplied by |alpha_j|; alpha_j = 0 (and padding) -> s_j = -1e300
    int64_t j = grp * 8 + lane;
    double a = (alpha && j < ntrain) ? alpha[j] : 0.0;
    double x, sj = -0.5 * nj[lane];
    if (HAS_V) {
      x = (j < ntrain) ? a - xv[lane] * inv_s2 : 0.0;
    } else if (raw_alpha) {
      x = (j < ntrain) ? a : 0.0;  // radial kernels: the weight multiplies k and phi after the fact; s_j = -n_j/2
    } else {
      x = (a < 0.0) ? -0.0 : 0.0;
      sj = (a != 0.0) ? sj + log(fabs(a)) / inv_s2 : -1e300;
    }
    osc[(lane >> 1) * 4 + (lane & 1) * 2 + 0] = x;
    osc[(lane >> 1) * 4 + (lane & 1) * 2 + 1] = sj;
  }
}

// range functions exported by the kreg_predict_inst.cu objects
#define KREG_DECL_RANGE(r)                                                                                          \
  int launch_predict_##r(int nat, const PredictArgs &a, bool has_v, int sms, void *ws, size_t ws_bytes, bool query, \
                         size_t *need, cudaStream_t st);                                                            \
  int tile_rows_##r(int nat, bool has_v);
KREG_DECL_RANGE(r0) KREG_DECL_RANGE(r1) KREG_DECL_RANGE(r2) KREG_DECL_RANGE(r3) KREG_DECL_RANGE(r4)
#undef KREG_DECL_RANGE

static int launch_predict(int nat, const PredictArgs &a, bool has_v, int sms, void *ws, size_t ws_bytes, bool query,
                          size_t *need, cudaStream_t st) {
  if (nat <= 9) return launch_predict_r0(nat, a, has_v, sms, ws, ws_bytes, query, need, st);
  if (nat <= 15) return launch_predict_r1(nat, a, has_v, sms, ws, ws_bytes, query, need, st);
  if (nat <= 20) return launch_predict_r2(nat, a, has_v, sms, ws, ws_bytes, query, need, st);
  if (nat <= 22) return launch_predict_r3(nat, a, has_v, sms, ws, ws_bytes, query, need, st);
  return launch_predict_r4(nat, a, has_v, sms, ws, ws_bytes, query, need, st);
}

static int tile_rows(int nat, bool has_v) {
  if (nat <= 9) return tile_rows_r0(nat, has_v);
  if (nat <= 15) return tile_rows_r1(nat, has_v);
  if (nat <= 20) return tile_rows_r2(nat, has_v);
  if (nat <= 22) return tile_rows_r3(nat, has_v);
  return tile_rows_r4(nat, has_v);
}

}  // namespace kreg

using namespace kreg;

extern "C" int kreg_precompute_v(int natoms, int64_t ntrain, const double *xyz_train, const double *xeq,
                                 const double *alpha_grad, double *V, void *stream) {
  if (natoms < KREG_MIN_NATOMS || natoms > KREG_MAX_NATOMS) return KREG_E_UNSUPPORTED;
  if (ntrain < 0 || (ntrain > 0 && (!xyz_train || !xeq || !alpha_grad || !V))) return KREG_E_BADARG;
  if (ntrain == 0) return KREG_OK;
  const int64_t total = ntrain * descr_size(natoms);
  int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  precompute_v_kernel<<<blocks, 256, 0, as_stream(stream)>>>(natoms, ntrain, xyz_train, xeq, alpha_grad, V);
  KREG_LAUNCH_CHECK();
  return KREG_OK;
}

// centre = mean of the first min(N, 2048) descriptor rows (same rule as descriptor_mean_kernel)
__global__ void __launch_bounds__(256) descriptor_mean_x_kernel(int D, int64_t npoints, const double *__restrict__ X,
                                                                double *__restrict__ center) {
  __shared__ double red[256];
  const int d = blockIdx.x;
  if (npoints > 2048) npoints = 2048;
  double s = 0.0;
  for (int64_t p = threadIdx.x; p < npoints; p += 256) s += X[p * D + d];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) center[d] = npoints > 0 ? red[0] / (double)npoints : 0.0;
}

extern "C" int kreg_pack_model_x(int natoms, int64_t ntrain, const double *X_train, const double *alpha_val,
                                 const double *V, double sigma, double *center, void *packed, size_t packed_bytes,
                                 void *stream) {
  if (natoms < KREG_MIN_NATOMS || natoms > KREG_MAX_NATOMS) return KREG_E_UNSUPPORTED;
  if (ntrain <= 0 || !X_train || !center || !packed || !(sigma > 0.0)) return KREG_E_BADARG;
  const bool has_v = V != nullptr;
  if (packed_bytes < kreg_packed_model_bytes(natoms, ntrain, has_v)) return KREG_E_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(packed) & 15) != 0) return KREG_E_BADARG;
  cudaStream_t st = as_stream(stream);
  const TcShape sh(natoms);
  descriptor_mean_x_kernel<<<sh.D, 256, 0, st>>>(sh.D, ntrain, X_train, center);
  KREG_LAUNCH_CHECK();
  if (natoms > KREG_FUSED_MAX_NATOMS)
    return big_pack_model(natoms, ntrain, nullptr, X_train, nullptr, alpha_val, V, sigma, center, packed, packed_bytes, st);
  const int TS = sh.NT * 8 > sh.KS * 4 ? sh.NT * 8 : sh.KS * 4;
  const size_t smem = (size_t)(16 * TS + 16) * sizeof(double);
  const double inv_s2 = 1.0 / (sigma * sigma);
  const double c1 = kExpScale * inv_s2;
  const unsigned groups = (unsigned)((ntrain + 7) / 8);
  if (has_v)
    pack_model_kernel<true><<<groups, 32, smem, st>>>(natoms, ntrain, nullptr, nullptr, center, alpha_val, V, c1, inv_s2,
                                                      static_cast<double *>(packed), X_train);
  else
    pack_model_kernel<false><<<groups, 32, smem, st>>>(natoms, ntrain, nullptr, nullptr, center, alpha_val, nullptr, c1,
                                                       inv_s2, static_cast<double *>(packed), X_train);
  KREG_LAUNCH_CHECK();
  return KREG_OK;
}

extern "C" int kreg_predict_tile_rows(int natoms, int has_gradient_terms) {
  if (natoms < KREG_MIN_NATOMS || natoms > KREG_MAX_NATOMS) return KREG_E_UNSUPPORTED;
  if (natoms > KREG_FUSED_MAX_NATOMS) return 128;  // CTA tile rows of the tiled path
  return tile_rows(natoms, has_gradient_terms != 0);
}

extern "C" size_t kreg_packed_model_bytes(int natoms, int64_t ntrain, int has_gradient_terms) {
  if (natoms < KREG_MIN_NATOMS || natoms > KREG_MAX_NATOMS || ntrain < 0) return 0;
  if (natoms > KREG_FUSED_MAX_NATOMS) return big_packed_model_bytes(natoms, ntrain, has_gradient_terms != 0);
  const TcShape sh(natoms);
  const int64_t groups = (ntrain + 7) / 8;
  return (size_t)groups * sh.chunk_doubles(has_gradient_terms != 0) * sizeof(double);
}

extern "C" int kreg_pack_model(int natoms, int64_t ntrain, const double *xyz_train, const double *xeq,
                               const double *alpha_val, const double *V, double sigma, double *center,
                               void *packed, size_t packed_bytes, void *stream) {
  if (natoms < KREG_MIN_NATOMS || natoms > KREG_MAX_NATOMS) return KREG_E_UNSUPPORTED;
  if (ntrain <= 0 || !xyz_train || !xeq || !center || !packed || !(sigma > 0.0)) return KREG_E_BADARG;
  const bool has_v = V != nullptr;
  if (packed_bytes < kreg_packed_model_bytes(natoms, ntrain, has_v)) return KREG_E_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(packed) & 15) != 0) return KREG_E_BADARG;
  cudaStream_t st = as_stream(stream);
  int rc = launch_descriptor_mean(natoms, ntrain, xyz_train, xeq, center, st);
  if (rc) return rc;
  if (natoms > KREG_FUSED_MAX_NATOMS)
    return big_pack_model(natoms, ntrain, xyz_train, nullptr, xeq, alpha_val, V, sigma, center, packed, packed_bytes, st);
  const TcShape sh(natoms);
  const int TS = sh.NT * 8 > sh.KS * 4 ? sh.NT * 8 : sh.KS * 4;
  const size_t smem = (size_t)(16 * TS + 16) * sizeof(double);
  const double inv_s2 = 1.0 / (sigma * sigma);
  const double c1 = kExpScale * inv_s2;
  const unsigned groups = (unsigned)((ntrain + 7) / 8);
  if (has_v)
    pack_model_kernel<true><<<groups, 32, smem, st>>>(natoms, ntrain, xyz_train, xeq, center, alpha_val, V, c1,
                                                      inv_s2, static_cast<double *>(packed), nullptr);
  else
    pack_model_kernel<false><<<groups, 32, smem, st>>>(natoms, ntrain, xyz_train, xeq, center, alpha_val, nullptr,
                                                       c1, inv_s2, static_cast<double *>(packed), nullptr);
  KREG_LAUNCH_CHECK();
  return KREG_OK;
}

static int device_sm_count(int *sms) {
  static thread_local int cached_dev = -1, cached_sms = 0;
  int dev = 0;
  KREG_CUDA_TRY(cudaGetDevice(&dev));
  if (dev != cached_dev) {
    KREG_CUDA_TRY(cudaDeviceGetAttribute(&cached_sms, cudaDevAttrMultiProcessorCount, dev));
    cached_dev = dev;
  }
  *sms = cached_sms;
  return KREG_OK;
}

extern "C" size_t kreg_predict_workspace_bytes(int natoms, int64_t ntrain, int has_gradient_terms, int64_t npoints) {
  if (natoms < KREG_MIN_NATOMS || natoms > KREG_MAX_NATOMS || ntrain <= 0 || npoints <= 0) return 0;
  int sms = 0;
  if (device_sm_count(&sms) != KREG_OK) return 0;
  if (natoms > KREG_FUSED_MAX_NATOMS)
    return big_predict_workspace_bytes(natoms, ntrain, has_gradient_terms != 0, npoints, sms);
  PredictArgs a = {};
  a.npoints = npoints;
  a.ngroups = (int)((ntrain + 7) / 8);
  size_t need = 0;
  if (launch_predict(natoms, a, has_gradient_terms != 0, sms, nullptr, 0, true, &need, nullptr) != KREG_OK) return 0;
  return need;
}

extern "C" int kreg_predict(int natoms, int64_t ntrain, const void *packed, const double *center, const double *xeq,
                            double sigma, double prior, int has_gradient_terms, int64_t npoints, const double *xyz,
                            int flags, double *E, double *G, void *workspace, size_t workspace_bytes, void *stream) {
  if (natoms < KREG_MIN_NATOMS || natoms > KREG_MAX_NATOMS) return KREG_E_UNSUPPORTED;
  if (ntrain <= 0 || npoints < 0 || !packed || !center || !xeq || !(sigma > 0.0)) return KREG_E_BADARG;
  if (!(flags & (KREG_PREDICT_ENERGY | KREG_PREDICT_GRADIENT))) return KREG_E_BADARG;
  if (npoints == 0) return KREG_OK;
  if (!xyz) return KREG_E_BADARG;
  if ((flags & KREG_PREDICT_ENERGY) && !E) return KREG_E_BADARG;
  if ((flags & KREG_PREDICT_GRADIENT) && !G) return KREG_E_BADARG;
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc) return rc;
  if (natoms > KREG_FUSED_MAX_NATOMS)
    return big_predict(natoms, ntrain, packed, center, xeq, sigma, prior, has_gradient_terms != 0, npoints, xyz,
                       (flags & KREG_PREDICT_ENERGY) ? E : nullptr, (flags & KREG_PREDICT_GRADIENT) ? G : nullptr,
                       workspace, workspace_bytes, sms, as_stream(stream));
  PredictArgs a = {};
  a.xyz = xyz;
  a.npoints = npoints;
  a.xeq = xeq;
  a.center = center;
  a.packed = static_cast<const double *>(packed);
  a.ngroups = (int)((ntrain + 7) / 8);
  a.inv_s2 = 1.0 / (sigma * sigma);
  a.c1 = kExpScale * a.inv_s2;
  a.prior = prior;
  a.E = (flags & KREG_PREDICT_ENERGY) ? E : nullptr;
  a.G = (flags & KREG_PREDICT_GRADIENT) ? G : nullptr;
  size_t need = 0;
  return launch_predict(natoms, a, has_gradient_terms != 0, sms, workspace, workspace_bytes, false, &need,
                        as_stream(stream));
}

// ---- models on MLatomF's other radial kernel functions (values-trained) --------------------------------------------------
extern "C" size_t kreg_radial_packed_model_bytes(int natoms, int64_t ntrain) {
  if (natoms < KREG_MIN_NATOMS || natoms > KREG_MAX_NATOMS || ntrain <= 0) return 0;
  if (natoms <= KREG_FUSED_MAX_NATOMS) return kreg_packed_model_bytes(natoms, ntrain, 0);  // the fused kernels' chunks
  return big_packed_model_bytes(natoms, ntrain, false);
}

extern "C" int kreg_pack_model_radial(int natoms, int64_t ntrain, const double *xyz_train, const double *xeq,
                                      const double *alpha_val, double *center, void *packed, size_t packed_bytes,
                                      void *stream) {
  if (natoms < KREG_MIN_NATOMS || natoms > KREG_MAX_NATOMS) return KREG_E_UNSUPPORTED;
  if (ntrain <= 0 || !xyz_train || !xeq || !alpha_val || !center || !packed) return KREG_E_BADARG;
  if (packed_bytes < kreg_radial_packed_model_bytes(natoms, ntrain)) return KREG_E_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(packed) & 255) != 0) return KREG_E_BADARG;
  cudaStream_t st = as_stream(stream);
  int rc = launch_descriptor_mean(natoms, ntrain, xyz_train, xeq, center, st);
  if (rc) return rc;
  if (natoms <= KREG_FUSED_MAX_NATOMS) {
    // the fused kernels' fragment-ordered chunks with alpha_j itself and s_j = -n_j/2 in the scalar slots
    const TcShape sh(natoms);
    const int TS = sh.NT * 8 > sh.KS * 4 ? sh.NT * 8 : sh.KS * 4;
    const size_t smem = (size_t)(16 * TS + 16) * sizeof(double);
    pack_model_kernel<false><<<(unsigned)((ntrain + 7) / 8), 32, smem, st>>>(natoms, ntrain, xyz_train, xeq, center, alpha_val,
                                                                         nullptr, 1.0, 1.0, static_cast<double *>(packed),
                                                                         nullptr, 1);
    KREG_LAUNCH_CHECK();
    return KREG_OK;
  }
  // the kernel width does not enter the packed arrays (centred rows, -n_j/2, alpha_j, transposed rows)
  return big_pack_model(natoms, ntrain, xyz_train, nullptr, xeq, alpha_val, nullptr, 1.0, center, packed, packed_bytes, st);
}

extern "C" size_t kreg_predict_radial_workspace_bytes(int natoms, int64_t ntrain, int64_t npoints) {
  if (natoms < KREG_MIN_NATOMS || natoms > KREG_MAX_NATOMS || ntrain <= 0 || npoints <= 0) return 0;
  int sms = 0;
  if (device_sm_count(&sms) != KREG_OK) return 0;
  if (natoms <= KREG_FUSED_MAX_NATOMS) {  // small batches: partial accumulators of the split launch (0 otherwise)
    PredictArgs a = {};
    a.npoints = npoints;
    a.ngroups = (int)((ntrain + 7) / 8);
    a.radial = 1;
    size_t need = 0;
    if (launch_predict(natoms, a, false, sms, nullptr, 0, true, &need, nullptr) != KREG_OK) return 0;
    return need;
  }
  return big_predict_radial_workspace_bytes(natoms, ntrain, npoints, sms);
}

extern "C" int kreg_predict_radial(int kind, int nn, int natoms, int64_t ntrain, const void *packed, const double *center,
                                   const double *xeq, double sigma, double prior, int64_t npoints, const double *xyz,
                                   int flags, double *E, double *G, void *workspace, size_t workspace_bytes, void *stream) {
  if (natoms < KREG_MIN_NATOMS || natoms > KREG_MAX_NATOMS) return KREG_E_UNSUPPORTED;
  if (kind != KREG_KERNEL_GAUSSIAN && kind != KREG_KERNEL_EXPONENTIAL && kind != KREG_KERNEL_MATERN)
    return KREG_E_UNSUPPORTED;  // the Laplacian's L1 distance has no GEMM form: kreg_kernel_block_ex + kreg_block_dot
  if (kind == KREG_KERNEL_MATERN && (nn < 0 || nn > 10)) return KREG_E_UNSUPPORTED;
  if (ntrain <= 0 || npoints < 0 || !packed || !center || !xeq || !(sigma > 0.0)) return KREG_E_BADARG;
  if (!(flags & (KREG_PREDICT_ENERGY | KREG_PREDICT_GRADIENT))) return KREG_E_BADARG;
  if (npoints == 0) return KREG_OK;
  if (!xyz) return KREG_E_BADARG;
  if ((flags & KREG_PREDICT_ENERGY) && !E) return KREG_E_BADARG;
  if ((flags & KREG_PREDICT_GRADIENT) && !G) return KREG_E_BADARG;
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc) return rc;
  // Matern of order 0 is the exponential kernel (Kfunk :837-849 with nn = 0); the reference differentiates it numerically
  if (kind == KREG_KERNEL_MATERN && nn == 0) kind = KREG_KERNEL_EXPONENTIAL;
  if (natoms <= KREG_FUSED_MAX_NATOMS) {
    // the fused two-GEMM kernels (predict_kernel<..., KIND = 1>): same operand ring and fragments as the Gaussian model
    PredictArgs a = {};
    a.xyz = xyz;
    a.npoints = npoints;
    a.xeq = xeq;
    a.center = center;
    a.packed = static_cast<const double *>(packed);
    a.ngroups = (int)((ntrain + 7) / 8);
    a.c1 = 1.0;      // row term -n_i/2, unscaled
    a.inv_s2 = 1.0;  // phi carries every factor of the gradient
    a.prior = prior;
    a.E = (flags & KREG_PREDICT_ENERGY) ? E : nullptr;
    a.G = (flags & KREG_PREDICT_GRADIENT) ? G : nullptr;
    a.radial = 1;
    a.kind = kind;
    a.nn = nn;
    a.inv_sigma = 1.0 / sigma;
    a.rinv_s2 = 1.0 / (sigma * sigma);
    if (kind == KREG_KERNEL_MATERN) {
      ValuesArgs coef = {};
      set_matern_coefficients(coef, nn, sigma);
      for (int m = 0; m < 12; m++) {
        a.mat_c[m] = coef.mat_c[m];
        a.mat_d[m] = coef.mat_d[m];
      }
    }
    size_t need = 0;
    return launch_predict(natoms, a, false, sms, workspace, workspace_bytes, false, &need, as_stream(stream));
  }
  return big_predict_radial(kind, nn, natoms, ntrain, packed, center, xeq, sigma, prior, npoints, xyz,
                            (flags & KREG_PREDICT_ENERGY) ? E : nullptr, (flags & KREG_PREDICT_GRADIENT) ? G : nullptr,
                            workspace, workspace_bytes, sms, as_stream(stream));
}
