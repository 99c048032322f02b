// kreg_internal.h -- declarations shared between the translation units of libkreg_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace kreg {

// kreg_descr.cu
int launch_descriptors(int nat, int64_t npoints, const double *xyz, const double *xeq, const double *center,
                       double *X, cudaStream_t st);
int launch_descriptor_mean(int nat, int64_t npoints, const double *xyz, const double *xeq, double *center,
                           cudaStream_t st);

// kreg_predict_big.cu -- molecules above KREG_FUSED_MAX_NATOMS atoms (tiled three-kernel prediction path)
size_t big_packed_model_bytes(int nat, int64_t ntr, bool has_v);
int big_pack_model(int nat, int64_t ntr, const double *xyz_train, const double *X_train, const double *xeq,
                   const double *alpha_val, const double *V, double sigma, double *center, void *packed,
                   size_t packed_bytes, cudaStream_t st);
size_t big_predict_workspace_bytes(int nat, int64_t ntr, bool has_v, int64_t npoints, int sms);
int big_predict(int nat, int64_t ntr, const void *packed, const double *center, const double *xeq, double sigma,
                double prior, bool has_v, int64_t npoints, const double *xyz, double *E, double *G, void *workspace,
                size_t workspace_bytes, int sms, cudaStream_t st);

size_t big_predict_radial_workspace_bytes(int nat, int64_t ntr, int64_t npoints, int sms);
int big_predict_radial(int kind, int nn, int nat, int64_t ntr, const void *packed, const double *center, const double *xeq,
                       double sigma, double prior, int64_t npoints, const double *xyz, double *E, double *G,
                       void *workspace, size_t workspace_bytes, int sms, cudaStream_t st);

// Shapes of the tensor-core formulation for a molecule with NAT atoms.
//   GEMM1  S  = Xtest (rows x D) . Xtrain^T        : K dimension D, padded to 4*KS
//   GEMM2  A += W (rows x 8) . [Xtrain | 1] (8 x N) : N = D + 1 (ones column -> energy), padded to 8*NT
struct TcShape {
  int D, KS, KS2, NT;
  __host__ __device__ constexpr TcShape(int nat)
      : D(nat * (nat - 1) / 2), KS((D + 3) / 4), KS2((KS + 1) / 2), NT((D + 1 + 7) / 8) {}
  // doubles in one packed group (8 training points)
  __host__ __device__ constexpr int chunk_doubles(bool has_v) const {
    return 64 * (KS2 + NT) * (has_v ? 2 : 1) + 16;
  }
};

}  // namespace kreg
