/*
 * kreg_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C (C11 + OpenMP) CPU restatement of the reference's KREG hot path, written
 * from the published algorithm, each function citing the reference lines it follows.
 * It is the checker for the CUDA product path and the "port" CPU baseline of bench.py.
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load it; the
 * product (mlatom_b200/) never links, loads or calls anything in oracle/.
 *
 * Pinning: every function here is compared against the reference's own compiled library
 * (oracle/_ref/KREG.so, see oracle/ref_kreg.py) by tests/test_oracle_vs_ref.py, and against
 * the committed golden vectors in tests/golden/ that were generated from that library.
 *
 * Array conventions (identical bytes to the reference's Fortran arrays):
 *   xyz[p][a][t]  == Fortran XYZ(t,a,p)      (3,Nat,N)
 *   X[p][d]       == Fortran X(d,p)          (D,N)
 *   K[i*M + j]    == Fortran K(j,i) == K(i,j) (symmetric, both triangles written)
 *   alpha[M], y[M]: values first, then gradients ordered (point, atom, xyz) with xyz fastest
 *                   (KREG.f90:273-291 get_itrgrxyz; A_KRR.f90:357-424).
 * Descriptor index: d(a,b) for a<b, 0-based: (2*Nat-a-1)*a/2 + (b-a) - 1
 *   == KREG.f90:48 `(2*Natoms-ii)*(ii-1)/2 + jj-ii` with 1-based ii=a+1, jj=b+1, minus 1.
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- index helpers ------------------------------------------------------------------ */

/* KREG.f90:34-52 get_ac2dArray; A_KRR_kernel.f90:955-963 mod_ac2dArray; D_rel2eq.f90:44-50 */
static inline int pair_index(int nat, int a, int b) {
  if (a > b) { int t = a; a = b; b = t; }
  return (2 * nat - a - 1) * a / 2 + (b - a) - 1;
}

void kro_ac2d(int nat, int *ac2d /* [nat][nat], 1-based values like the reference, diag 0 */) {
  for (int a = 0; a < nat; a++) ac2d[a * nat + a] = 0;
  for (int a = 0; a < nat - 1; a++)
    for (int b = a + 1; b < nat; b++) {
      int v = pair_index(nat, a, b) + 1;
      ac2d[a * nat + b] = v;
      ac2d[b * nat + a] = v;
    }
}

/* KREG.f90:77-85 Rij; MLatomF_src/mathUtils.f90:743-758 */
static inline double rij(const double *pa, const double *pb) {
  double dx = pa[0] - pb[0], dy = pa[1] - pb[1], dz = pa[2] - pb[2];
  return sqrt(dx * dx + dy * dy + dz * dz);
}

/* KREG.f90:223-239 get_Xeq; D_rel2eq.f90:176-201 getReq */
void kro_xeq(int nat, const double *req /* [nat][3] */, double *xeq /* [D] */) {
  for (int a = 0; a < nat; a++)
    for (int b = a + 1; b < nat; b++) xeq[pair_index(nat, a, b)] = rij(req + 3 * a, req + 3 * b);
}

/* KREG.f90:241-271 RE_descriptor / calc_RE_descriptors; D_rel2eq.f90:17-83 calcrel2eq */
void kro_descriptors(int nat, long np, const double *xyz, const double *xeq, double *X) {
  const int D = nat * (nat - 1) / 2;
#pragma omp parallel for schedule(static)
  for (long p = 0; p < np; p++) {
    const double *m = xyz + (size_t)p * nat * 3;
    double *x = X + (size_t)p * D;
    for (int a = 0; a < nat; a++)
      for (int b = a + 1; b < nat; b++) {
        int d = pair_index(nat, a, b);
        x[d] = xeq[d] / rij(m + 3 * a, m + 3 * b);
      }
  }
}

/* ---- kernel function and its derivatives, literal forms ----------------------------- */

/* KREG.f90:54-66 distance2; A_KRR_kernel.f90:2040-2074 distance (Gaussian branch) */
static inline double distance2(int D, const double *xi, const double *xj) {
  double s = 0.0;
  for (int d = 0; d < D; d++) {
    double t = xi[d] - xj[d];
    s = s + t * t;
  }
  return s;
}

/* KREG.f90:87-95 kernel; A_KRR_kernel.f90:806-851 Kfunk (plain Gaussian branch) */
static inline double kernel(int D, const double *xi, const double *xj, double sigma) {
  return exp(-1.0 * distance2(D, xi, xj) / (2.0 * sigma * sigma));
}

/* KREG.f90:106-125 dKdMiat; A_KRR_kernel.f90:1407-1447 dKdMiat_unpermuted (RE branch).
 * d k(X_i, X_j) / d M_i,(a,t): derivative with respect to the FIRST geometry. */
static double dKdMiat(int nat, int D, int a, int t, const double *xi, const double *xj,
                      const double *xyzi, double sigma) {
  double s = 0.0;
  for (int b = 0; b < nat; b++) {
    if (a != b) {
      int d = pair_index(nat, a, b);
      double r = rij(xyzi + 3 * a, xyzi + 3 * b);
      s = s + (xj[d] - xi[d]) * xi[d] * (xyzi[3 * b + t] - xyzi[3 * a + t]) / (r * r);
    }
  }
  return s * kernel(D, xi, xj, sigma) / (sigma * sigma);
}

/* KREG.f90:127-165 d2KdMiatdMjbu; A_KRR_kernel.f90:1947-1975 (non-permuted RE branch) */
static double d2KdMiatdMjbu(int nat, int D, int a, int t, int b, int u, const double *xi,
                            const double *xj, const double *xyzi, const double *xyzj,
                            double sigma) {
  double out1 = 0.0, out2 = 0.0;
  for (int c = 0; c < nat; c++) {
    if (a != c) {
      int d = pair_index(nat, a, c);
      double ri = rij(xyzi + 3 * a, xyzi + 3 * c);
      double dXdMi = xi[d] * (xyzi[3 * c + t] - xyzi[3 * a + t]) / (ri * ri);
      double in1 = 0.0, in2 = 0.0;
      for (int g = 0; g < nat; g++) {
        if (b != g) {
          int e = pair_index(nat, b, g);
          double rj = rij(xyzj + 3 * b, xyzj + 3 * g);
          double dXdMj = xj[e] * (xyzj[3 * g + u] - xyzj[3 * b + u]) / (rj * rj);
          in1 = in1 + dXdMj * (xj[d] - xi[d]) * (xi[e] - xj[e]);
          if (d == e) in2 = in2 + dXdMj;
        }
      }
      out1 = out1 + in1 * dXdMi;
      out2 = out2 + in2 * dXdMi;
    }
  }
  out1 = out1 / (sigma * sigma);
  return (out1 + out2) * kernel(D, xi, xj, sigma) / (sigma * sigma);
}

/* ---- kernel matrix ------------------------------------------------------------------ */

/* KREG.f90:293-346 calc_kernel_matrix; A_KRR_kernel.f90:130-188 calcKernel.
 * ntrval is 0 or ntr; ntrgr is 0 or 3*nat*ntr (kreg_api.py:52-72).  Literal (slow) form. */
void kro_kernel_matrix(int nat, int ntr, int ntrval, int ntrgr, const double *xyz, const double *X,
                       double sigma, double *K) {
  const int D = nat * (nat - 1) / 2;
  const size_t M = (size_t)ntrval + (size_t)ntrgr;
  const int G = 3 * nat;
  memset(K, 0, M * M * sizeof(double));

#pragma omp parallel for schedule(dynamic, 8)
  for (int i = 0; i < ntrval; i++) {
    K[(size_t)i * M + i] = kernel(D, X + (size_t)i * D, X + (size_t)i * D, sigma);
    for (int j = i + 1; j < ntrval; j++) {
      double k = kernel(D, X + (size_t)i * D, X + (size_t)j * D, sigma);
      K[(size_t)i * M + j] = k;
      K[(size_t)j * M + i] = k;
    }
  }
  /* value-gradient: K(ii, NtrVal+jj) = dKdMiat(a_jj, t_jj, X(:,p_jj), X(:,ii), XYZ(:,:,p_jj)) */
#pragma omp parallel for schedule(dynamic, 8)
  for (int i = 0; i < ntrval; i++) {
    for (int jj = 0; jj < ntrgr; jj++) {
      int p = jj / G, a = (jj % G) / 3, t = jj % 3;
      double v = dKdMiat(nat, D, a, t, X + (size_t)p * D, X + (size_t)i * D,
                         xyz + (size_t)p * nat * 3, sigma);
      K[(size_t)i * M + ntrval + jj] = v;
      K[((size_t)ntrval + jj) * M + i] = v;
    }
  }
  /* gradient-gradient */
#pragma omp parallel for schedule(dynamic, 4)
  for (int ii = 0; ii < ntrgr; ii++) {
    int pi = ii / G, a = (ii % G) / 3, t = ii % 3;
    for (int jj = ii; jj < ntrgr; jj++) {
      int pj = jj / G, b = (jj % G) / 3, u = jj % 3;
      double v = d2KdMiatdMjbu(nat, D, a, t, b, u, X + (size_t)pi * D, X + (size_t)pj * D,
                               xyz + (size_t)pi * nat * 3, xyz + (size_t)pj * nat * 3, sigma);
      K[((size_t)ntrval + ii) * M + ntrval + jj] = v;
      K[((size_t)ntrval + jj) * M + ntrval + ii] = v;
    }
  }
}

/* KREG.f90:616-622 calcAlpha (diagonal shift only; the solve is LAPACK, done by the caller) */
void kro_add_lambda(int ntrval, int ntrgr, double lambdav, double lambdag, double *K) {
  const size_t M = (size_t)ntrval + (size_t)ntrgr;
  for (int j = 0; j < ntrval; j++) K[(size_t)j * M + j] += lambdav;
  for (int j = 0; j < ntrgr; j++) K[((size_t)ntrval + j) * M + ntrval + j] += lambdag;
}

/* ---- prediction: literal KREG.f90 form ---------------------------------------------- */

/* KREG.f90:534-596 predict with Yest_KRR :348-366 and YgradXYZest_KRR :368-391.
 * No prior is added (the Python adapter does that, kreg_api.py:122).  O(3 Nat Nat D) per pair. */
void kro_predict_kregf90(int nat, int ntr, int ntrval, int ntrgr, const double *xyz_tr,
                         const double *X_tr, const double *alpha, double sigma, long np,
                         const double *xyz_te, const double *X_te, int calc_val, int calc_grad,
                         double *E, double *G) {
  const int D = nat * (nat - 1) / 2;
  const int G3 = 3 * nat;
  if (calc_val) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < np; i++) {
      const double *xp = X_te + (size_t)i * D;
      double e = 0.0;
      for (int j = 0; j < ntrval; j++) e = e + alpha[j] * kernel(D, xp, X_tr + (size_t)j * D, sigma);
      for (int jj = 0; jj < ntrgr; jj++) {
        int p = jj / G3, a = (jj % G3) / 3, t = jj % 3;
        e = e + alpha[ntrval + jj] * dKdMiat(nat, D, a, t, X_tr + (size_t)p * D, xp,
                                             xyz_tr + (size_t)p * nat * 3, sigma);
      }
      E[i] = e;
    }
  }
  if (calc_grad) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < np; i++) {
      const double *xp = X_te + (size_t)i * D;
      const double *mp = xyz_te + (size_t)i * nat * 3;
      for (int a = 0; a < nat; a++)
        for (int t = 0; t < 3; t++) {
          double g = 0.0;
          for (int j = 0; j < ntrval; j++)
            g = g + alpha[j] * dKdMiat(nat, D, a, t, xp, X_tr + (size_t)j * D, mp, sigma);
          for (int jj = 0; jj < ntrgr; jj++) {
            int p = jj / G3, b = (jj % G3) / 3, u = jj % 3;
            g = g + alpha[ntrval + jj] * d2KdMiatdMjbu(nat, D, a, t, b, u, xp, X_tr + (size_t)p * D,
                                                       mp, xyz_tr + (size_t)p * nat * 3, sigma);
          }
          G[(size_t)i * G3 + 3 * a + t] = g;
        }
    }
  }
}

/* KREG.f90:180-221 d2KdMiatdMibu: second derivative of k(X_i, X_j) with respect to two coordinates of the SAME geometry
 * M_i.  As in the reference only products of first derivatives of the descriptor appear (no d2X/dM2 term). */
static double d2KdMiatdMibu(int nat, int D, int a, int t, int b, int u, const double *xi, const double *xj,
                            const double *xyzi, double sigma) {
  double out1 = 0.0, out2 = 0.0;
  for (int c = 0; c < nat; c++) {
    if (a != c) {
      int d = pair_index(nat, a, c);
      double ri = rij(xyzi + 3 * a, xyzi + 3 * c);
      double dXdMiat = xi[d] * (xyzi[3 * c + t] - xyzi[3 * a + t]) / (ri * ri);
      double in1 = 0.0, in2 = 0.0;
      for (int g = 0; g < nat; g++) {
        if (b != g) {
          int e = pair_index(nat, b, g);
          double rg = rij(xyzi + 3 * b, xyzi + 3 * g);
          double dXdMibu = xi[e] * (xyzi[3 * g + u] - xyzi[3 * b + u]) / (rg * rg);
          in1 = in1 + dXdMibu * (xj[d] - xi[d]) * (xj[e] - xi[e]);
          if (d == e) in2 = in2 - dXdMibu;
        }
      }
      out1 = out1 + in1 * dXdMiat;
      out2 = out2 + in2 * dXdMiat;
    }
  }
  out1 = out1 / (sigma * sigma);
  return (out1 + out2) * kernel(D, xi, xj, sigma) / (sigma * sigma);
}

/* KREG.f90:393-420 YhessXYZest_KRR driven by predict :588-594: values part of alpha only.
 * H[p][dim2][dim1] == Fortran YhessXYZest(dim1, dim2, p), dim = 3*atom + xyz. */
void kro_hessian_kregf90(int nat, int ntrval, const double *X_tr, const double *alpha, double sigma, long np,
                         const double *xyz_te, const double *X_te, double *H) {
  const int D = nat * (nat - 1) / 2;
  const int G3 = 3 * nat;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < np; i++) {
    const double *xp = X_te + (size_t)i * D;
    const double *mp = xyz_te + (size_t)i * nat * 3;
    for (int a = 0; a < nat; a++)
      for (int t = 0; t < 3; t++)
        for (int b = 0; b < nat; b++)
          for (int u = 0; u < 3; u++) {
            double h = 0.0;
            for (int j = 0; j < ntrval; j++)
              h = h + alpha[j] * d2KdMiatdMibu(nat, D, a, t, b, u, xp, X_tr + (size_t)j * D, mp, sigma);
            H[(size_t)i * G3 * G3 + (size_t)(3 * b + u) * G3 + (3 * a + t)] = h;
          }
  }
}

/* ---- prediction: MLatomF "Optimized Code" form -------------------------------------- */

/* A_KRR.f90:985-1029 calcEst_KRR driving Yest_KRR (A_KRR_kernel.f90:225-245) and
 * YgradXYZest_KRR's Gaussian/RE/non-permuted branch (A_KRR_kernel.f90:327-412).  Loop structure
 * kept as in the reference (the Jacobian terms are re-evaluated inside the training-point loop).
 * `prior` is added to E as MLatomF does (A_KRR.f90:1006). */
void kro_predict_mlatomf(int nat, int ntr, int ntrval, int ntrgr, const double *xyz_tr,
                         const double *X_tr, const double *alpha, double sigma, double prior,
                         long np, const double *xyz_te, const double *X_te, int calc_val,
                         int calc_grad, double *E, double *G) {
  const int D = nat * (nat - 1) / 2;
  const int G3 = 3 * nat;
  const double inv_s2 = 1.0 / (sigma * sigma); /* dKdXid_const1, d2KdXiddXje_const1/3 */
  if (calc_val) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < np; i++) {
      const double *xp = X_te + (size_t)i * D;
      double e = 0.0;
      for (int j = 0; j < ntrval; j++) e = e + alpha[j] * kernel(D, xp, X_tr + (size_t)j * D, sigma);
      for (int jj = 0; jj < ntrgr; jj++) {
        int p = jj / G3, a = (jj % G3) / 3, t = jj % 3;
        e = e + alpha[ntrval + jj] * dKdMiat(nat, D, a, t, X_tr + (size_t)p * D, xp,
                                             xyz_tr + (size_t)p * nat * 3, sigma);
      }
      E[i] = prior + e;
    }
  }
  if (calc_grad) {
#pragma omp parallel
    {
      double *acc1 = (double *)malloc(sizeof(double) * G3);
      double *acc2 = (double *)malloc(sizeof(double) * G3);
      double *in1 = (double *)malloc(sizeof(double) * G3);
      double *in2 = (double *)malloc(sizeof(double) * G3);
      double *in3 = (double *)malloc(sizeof(double) * G3);
#pragma omp for schedule(static)
      for (long i = 0; i < np; i++) {
        const double *xp = X_te + (size_t)i * D;
        const double *mp = xyz_te + (size_t)i * nat * 3;
        for (int q = 0; q < G3; q++) acc1[q] = acc2[q] = in1[q] = in2[q] = in3[q] = 0.0;
        /* A_KRR_kernel.f90:333-352 */
        for (int j = 0; j < ntrval; j++) {
          const double *xj = X_tr + (size_t)j * D;
          double cmf = alpha[j] * kernel(D, xp, xj, sigma);
          for (int a = 0; a < nat - 1; a++)
            for (int c = a + 1; c < nat; c++) {
              int d = pair_index(nat, a, c);
              double r = rij(mp + 3 * a, mp + 3 * c);
              double temp1 = (xj[d] - xp[d]) * xp[d] / (r * r);
              for (int t = 0; t < 3; t++) {
                double temp2 = mp[3 * c + t] - mp[3 * a + t];
                in1[3 * a + t] = in1[3 * a + t] + temp2 * temp1;
                in1[3 * c + t] = in1[3 * c + t] - temp2 * temp1;
              }
            }
          for (int q = 0; q < G3; q++) { acc1[q] = acc1[q] + in1[q] * cmf; in1[q] = 0.0; }
        }
        for (int q = 0; q < G3; q++) acc1[q] = acc1[q] * inv_s2;
        /* A_KRR_kernel.f90:365-399 */
        for (int jj = 0; jj < ntrgr; jj++) {
          int p = jj / G3, b = (jj % G3) / 3, u = jj % 3;
          const double *xj = X_tr + (size_t)p * D;
          const double *mj = xyz_tr + (size_t)p * nat * 3;
          double cmf = alpha[ntrval + jj] * kernel(D, xp, xj, sigma);
          for (int g = 0; g < nat; g++) {
            if (b != g) {
              int e = pair_index(nat, b, g);
              double rj = rij(mj + 3 * b, mj + 3 * g);
              double temp1 = xj[e] * (mj[3 * g + u] - mj[3 * b + u]) / (rj * rj);
              for (int a = 0; a < nat - 1; a++)
                for (int c = a + 1; c < nat; c++) {
                  int d = pair_index(nat, a, c);
                  double ri = rij(mp + 3 * a, mp + 3 * c);
                  double temp2 = xp[d] / (ri * ri);
                  double temp3 = (xj[d] - xp[d]) * (xp[e] - xj[e]) * inv_s2;
                  for (int t = 0; t < 3; t++) {
                    double temp4 = mp[3 * c + t] - mp[3 * a + t];
                    double v = temp4 * temp2 * temp3 + ((d == e) ? temp4 : 0.0) * temp2;
                    in2[3 * a + t] = in2[3 * a + t] + v;
                    in2[3 * c + t] = in2[3 * c + t] - v;
                  }
                }
              for (int q = 0; q < G3; q++) { in3[q] = in3[q] + in2[q] * temp1; in2[q] = 0.0; }
            }
          }
          for (int q = 0; q < G3; q++) { acc2[q] = acc2[q] + in3[q] * cmf; in3[q] = 0.0; }
        }
        for (int q = 0; q < G3; q++) G[(size_t)i * G3 + q] = acc1[q] + acc2[q] * inv_s2;
      }
      free(acc1); free(acc2); free(in1); free(in2); free(in3);
    }
  }
}

/* ---- prediction: factored O(D) form -------------------------- */

/* v_j = J_j alpha_g,j  with J[d,(a,t)] = X_d (x_c - x_a)_t / R_ac^2 for d=(a,c)
 * (KREG.f90:121,146).  Same algebra as the CUDA path, evaluated on the CPU in the plain
 * left-to-right summation order.  Used as a fast checker at sizes where the literal forms
 * above would take hours, after it has itself been checked against them at small sizes. */
void kro_precompute_v(int nat, int ntr, const double *xyz_tr, const double *X_tr,
                      const double *alpha_g /* [ntr][nat][3] */, double *V /* [ntr][D] */) {
  const int D = nat * (nat - 1) / 2;
#pragma omp parallel for schedule(static)
  for (int j = 0; j < ntr; j++) {
    const double *m = xyz_tr + (size_t)j * nat * 3;
    const double *x = X_tr + (size_t)j * D;
    const double *ag = alpha_g + (size_t)j * nat * 3;
    for (int a = 0; a < nat; a++)
      for (int c = a + 1; c < nat; c++) {
        int d = pair_index(nat, a, c);
        double r = rij(m + 3 * a, m + 3 * c);
        double s = 0.0;
        for (int t = 0; t < 3; t++) s += (m[3 * c + t] - m[3 * a + t]) * (ag[3 * a + t] - ag[3 * c + t]);
        V[(size_t)j * D + d] = x[d] * s / (r * r);
      }
  }
}

/* f_i = prior + sum_j k_ij c_ij,  c_ij = alpha_j + (X_i - X_j).v_j / sigma^2
 * g_i[d] = (1/sigma^2) sum_j k_ij [ c_ij (X_j,d - X_i,d) + v_j[d] ],  grad_M f_i = J_i^T g_i
 * alpha_v may be NULL (gradient-only model), V may be NULL (values-only model). */
void kro_predict_factored(int nat, int ntr, const double *X_tr, const double *alpha_v,
                          const double *V, double sigma, double prior, long np,
                          const double *xyz_te, const double *X_te, int calc_grad, double *E,
                          double *G, double *Escale /* optional: sum_j |k_ij c_ij| */) {
  const int D = nat * (nat - 1) / 2;
  const int G3 = 3 * nat;
  const double inv_s2 = 1.0 / (sigma * sigma);
#pragma omp parallel
  {
    double *g = (double *)malloc(sizeof(double) * D);
#pragma omp for schedule(static)
    for (long i = 0; i < np; i++) {
      const double *xp = X_te + (size_t)i * D;
      double e = 0.0, es = 0.0;
      for (int d = 0; d < D; d++) g[d] = 0.0;
      for (int j = 0; j < ntr; j++) {
        const double *xj = X_tr + (size_t)j * D;
        double k = kernel(D, xp, xj, sigma);
        double c = alpha_v ? alpha_v[j] : 0.0;
        if (V) {
          const double *vj = V + (size_t)j * D;
          double s = 0.0;
          for (int d = 0; d < D; d++) s += (xp[d] - xj[d]) * vj[d];
          c += s * inv_s2;
        }
        double w = k * c;
        e += w;
        es += fabs(w);
        if (calc_grad) {
          if (V) {
            const double *vj = V + (size_t)j * D;
            for (int d = 0; d < D; d++) g[d] += w * (xj[d] - xp[d]) + k * vj[d];
          } else {
            for (int d = 0; d < D; d++) g[d] += w * (xj[d] - xp[d]);
          }
        }
      }
      if (E) E[i] = prior + e;
      if (Escale) Escale[i] = es;
      if (calc_grad) {
        const double *m = xyz_te + (size_t)i * nat * 3;
        double *out = G + (size_t)i * G3;
        for (int q = 0; q < G3; q++) out[q] = 0.0;
        for (int a = 0; a < nat; a++)
          for (int c = a + 1; c < nat; c++) {
            int d = pair_index(nat, a, c);
            double r = rij(m + 3 * a, m + 3 * c);
            double f = g[d] * inv_s2 * xp[d] / (r * r);
            for (int t = 0; t < 3; t++) {
              double dx = m[3 * c + t] - m[3 * a + t];
              out[3 * a + t] += f * dx;
              out[3 * c + t] -= f * dx;
            }
          }
      }
    }
    free(g);
  }
}

/* ---- kernel matrix: factored form (fast checker for large gradient-augmented K) ------ */

/* Same matrix as kro_kernel_matrix, assembled per geometry pair from
 *   dk/dM_j      = -(k/s^2) J_j^T D,        D = X_j - X_i     (value-gradient row i, point j)
 *   d2k/dMi dMj  =  (k/s^2) [ J_i^T J_j - (J_i^T D)(J_j^T D)^T / s^2 ]
 * (expansion of KREG.f90:117-124 and :143-164). */
void kro_kernel_matrix_factored(int nat, int ntr, int ntrval, int ntrgr, const double *xyz,
                                const double *X, double sigma, double *K) {
  const int D = nat * (nat - 1) / 2;
  const int G3 = 3 * nat;
  const size_t M = (size_t)ntrval + (size_t)ntrgr;
  const double inv_s2 = 1.0 / (sigma * sigma);
  /* dense Jacobians J[p][d][q], q=(atom,xyz) */
  double *J = (double *)calloc((size_t)ntr * D * G3, sizeof(double));
#pragma omp parallel for schedule(static)
  for (int p = 0; p < ntr; p++) {
    const double *m = xyz + (size_t)p * nat * 3;
    for (int a = 0; a < nat; a++)
      for (int c = a + 1; c < nat; c++) {
        int d = pair_index(nat, a, c);
        double r = rij(m + 3 * a, m + 3 * c);
        double f = X[(size_t)p * D + d] / (r * r);
        for (int t = 0; t < 3; t++) {
          double dx = m[3 * c + t] - m[3 * a + t];
          J[((size_t)p * D + d) * G3 + 3 * a + t] = f * dx;
          J[((size_t)p * D + d) * G3 + 3 * c + t] = -f * dx;
        }
      }
  }
  memset(K, 0, M * M * sizeof(double));
#pragma omp parallel
  {
    double *delta = (double *)malloc(sizeof(double) * D);
    double *ui = (double *)malloc(sizeof(double) * G3);
    double *uj = (double *)malloc(sizeof(double) * G3);
#pragma omp for schedule(dynamic, 4)
    for (int i = 0; i < ntr; i++) {
      for (int j = 0; j < ntr; j++) {
        const double *xi = X + (size_t)i * D, *xj = X + (size_t)j * D;
        double k = kernel(D, xi, xj, sigma);
        if (ntrval) K[(size_t)i * M + j] = k;
        if (!ntrgr) continue;
        const double *Ji = J + (size_t)i * D * G3, *Jj = J + (size_t)j * D * G3;
        for (int d = 0; d < D; d++) delta[d] = xj[d] - xi[d];
        for (int q = 0; q < G3; q++) {
          double si = 0.0, sj = 0.0;
          for (int d = 0; d < D; d++) { si += Ji[d * G3 + q] * delta[d]; sj += Jj[d * G3 + q] * delta[d]; }
          ui[q] = si; uj[q] = sj;
        }
        if (ntrval) {
          for (int q = 0; q < G3; q++) {
            double v = -k * inv_s2 * uj[q];
            K[(size_t)i * M + ntrval + (size_t)j * G3 + q] = v;
            K[((size_t)ntrval + (size_t)j * G3 + q) * M + i] = v;
          }
        }
        for (int q = 0; q < G3; q++)
          for (int r = 0; r < G3; r++) {
            double jj = 0.0;
            for (int d = 0; d < D; d++) jj += Ji[d * G3 + q] * Jj[d * G3 + r];
            K[((size_t)ntrval + (size_t)i * G3 + q) * M + ntrval + (size_t)j * G3 + r] =
                k * inv_s2 * (jj - ui[q] * uj[r] * inv_s2);
          }
      }
    }
    free(delta); free(ui); free(uj);
  }
  free(J);
}

/* ---- MLatomF's other kernel functions on the same descriptor -------------------------------
 * distance (A_KRR_kernel.f90:2040-2074): sum of squares for Gaussian, its square root for exponential and Matern,
 * sum of absolute differences for Laplacian.  Kfunk (A_KRR_kernel.f90:806-851): Gaussian exp(-d/(2 sigma^2)),
 * Laplacian / exponential exp(-d/sigma), Matern exp(-d/sigma) n!/(2n)! sum_k (n+k)!/(k!(n-k)!) (2d/sigma)^(n-k).
 * kind: 0 Gaussian, 1 exponential, 2 Matern (order nn), 3 Laplacian.  PARITY UNPINNED: the MLatomF binary is absent
 * from the reference checkout and KREG.so only has the Gaussian kernel, so nothing executable pins these three. */
static double factorial_rprec(int nn) { /* mathUtils.f90:577-590 */
  double f = 1.0;
  for (int ii = 1; ii <= nn; ii++) f = f * ii;
  return f;
}

static double mlatomf_distance(int kind, int D, const double *xi, const double *xj) {
  double d = 0.0;
  if (kind == 3) {
    for (int k = 0; k < D; k++) d = d + fabs(xi[k] - xj[k]);
  } else {
    for (int k = 0; k < D; k++) d = d + (xi[k] - xj[k]) * (xi[k] - xj[k]);
    if (kind == 1 || kind == 2) d = sqrt(d);
  }
  return d;
}

static double mlatomf_kfunk(int kind, int nn, int D, const double *xi, const double *xj, double sigma) {
  const double dist = mlatomf_distance(kind, D, xi, xj);
  if (kind == 0) return exp(-1.0 * dist / (2 * sigma * sigma));
  if (kind == 1 || kind == 3) return exp(-1.0 * dist / sigma);
  double k = 0.0;
  for (int kk = 0; kk <= nn; kk++)
    k = k + factorial_rprec(nn + kk) / factorial_rprec(kk) / factorial_rprec(nn - kk) * pow(2 * dist / sigma, nn - kk);
  return exp(-1.0 * dist / sigma) * factorial_rprec(nn) / factorial_rprec(2 * nn) * k;
}

/* K[i][j] = Kfunk(Xa[i], Xb[j]), row-major na x nb (calcKernel / calc_Kvalidate loops, A_KRR_kernel.f90:130-223) */
void kro_kernel_block_generic(int kind, int nn, int D, long na, const double *Xa, long nb, const double *Xb,
                              double sigma, double *K) {
#pragma omp parallel for schedule(static)
  for (long i = 0; i < na; i++)
    for (long j = 0; j < nb; j++) K[i * nb + j] = mlatomf_kfunk(kind, nn, D, Xa + i * D, Xb + j * D, sigma);
}

/* dKdXid (A_KRR_kernel.f90:1178-1233): Gaussian and Matern (nn > 0) analytic, everything else a central difference of
 * Kfunk with h = sqrt(epsilon) * Xi(dd). */
static double mlatomf_dKdXid(int kind, int nn, int D, const double *Xi, const double *Xj, int dd, double sigma) {
  if (kind == 0) return mlatomf_kfunk(0, nn, D, Xi, Xj, sigma) * (Xj[dd] - Xi[dd]) / (sigma * sigma);
  if (kind == 2 && nn > 0) {
    const double dist = mlatomf_distance(kind, D, Xi, Xj);
    double v = 0.0;
    for (int kk = 0; kk <= nn - 1; kk++)
      v = v + factorial_rprec(nn + kk - 1) / factorial_rprec(kk) / factorial_rprec(nn - kk) * (nn - kk) *
                  pow(2.0 * dist / sigma, nn - kk - 1) * (Xj[dd] - Xi[dd]);
    return v * exp(-1.0 * dist / sigma) * factorial_rprec(nn) / factorial_rprec(2 * nn) / (sigma * sigma) * 2.0;
  }
  const double eps = 2.220446049250313e-16; /* epsilon(Xi(dd)) */
  double h = sqrt(eps);
  if (fabs(Xi[dd]) > 0.0 + 2.0 * eps) h = h * Xi[dd];
  double *Xph = (double *)malloc(sizeof(double) * 2 * D), *Xmh = Xph + D;
  memcpy(Xph, Xi, sizeof(double) * D);
  memcpy(Xmh, Xi, sizeof(double) * D);
  Xph[dd] = Xph[dd] + h;
  Xmh[dd] = Xmh[dd] - h;
  const double fxph = mlatomf_kfunk(kind, nn, D, Xph, Xj, sigma), fxmh = mlatomf_kfunk(kind, nn, D, Xmh, Xj, sigma);
  free(Xph);
  return (fxph - fxmh) / (2.0 * h);
}

/* dKdMiat_unpermuted, RE branch (A_KRR_kernel.f90:1407-1447) with dKdXid_changing / _const2 / _const1 (:1235-1301) */
static double mlatomf_dKdMiat(int kind, int nn, int nat, int D, const double *Xi, const double *Xj, const double *xyzi,
                              int a, int tt, double sigma) {
  const int analytic = (kind == 0) || (kind == 2 && nn > 0);
  double v = 0.0;
  for (int bb = 0; bb < nat; bb++) {
    if (a != bb) {
      const int dd = pair_index(nat, a, bb);
      const double changing = analytic ? Xj[dd] - Xi[dd] : mlatomf_dKdXid(kind, nn, D, Xi, Xj, dd, sigma);
      const double R = rij(xyzi + 3 * a, xyzi + 3 * bb);
      v = v + changing * Xi[dd] * (xyzi[3 * bb + tt] - xyzi[3 * a + tt]) / (R * R);
    }
  }
  double const2 = 1.0, const1 = 1.0;
  if (kind == 0) {
    const2 = mlatomf_kfunk(0, nn, D, Xi, Xj, sigma);
    const1 = 1.0 / (sigma * sigma);
  } else if (kind == 2 && nn > 0) {
    const double dist = mlatomf_distance(kind, D, Xi, Xj);
    const double cst = factorial_rprec(nn) / factorial_rprec(2 * nn) / (sigma * sigma) * 2.0; /* mod_MaternConsts(:,1), :966-977 */
    const2 = 0.0;
    for (int kk = 0; kk <= nn - 1; kk++)
      const2 = const2 + factorial_rprec(nn + kk - 1) / factorial_rprec(kk) / factorial_rprec(nn - kk) * (nn - kk) *
                            pow(2.0 / sigma, nn - kk - 1) * cst * pow(dist, nn - kk - 1);
    const2 = const2 * exp(-1.0 * dist / sigma);
  }
  return v * const2 * const1;
}

/* Yest_KRR (A_KRR_kernel.f90:225-245) and the generic branch of YgradXYZest_KRR (:731-744) for a values-trained model on
 * any of the kernel functions; prior added as calcEst_KRR does (A_KRR.f90:1006). */
void kro_predict_generic(int kind, int nn, int nat, long ntr, const double *X_tr, const double *alpha, double sigma,
                         double prior, long np, const double *xyz_te, const double *X_te, int calc_grad, double *E,
                         double *G) {
  const int D = nat * (nat - 1) / 2;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < np; i++) {
    const double *Xi = X_te + i * D, *xyzi = xyz_te + (size_t)i * nat * 3;
    double y = 0.0;
    for (long jj = 0; jj < ntr; jj++) y = y + alpha[jj] * mlatomf_kfunk(kind, nn, D, Xi, X_tr + jj * D, sigma);
    E[i] = y + prior;
    if (!calc_grad) continue;
    for (int a = 0; a < nat; a++)
      for (int tt = 0; tt < 3; tt++) {
        double g = 0.0;
        for (long jj = 0; jj < ntr; jj++)
          g = g + alpha[jj] * mlatomf_dKdMiat(kind, nn, nat, D, Xi, X_tr + jj * D, xyzi, a, tt, sigma);
        G[((size_t)i * nat + a) * 3 + tt] = g;
      }
  }
}

/* ---- permutationally invariant kernel on the permuted RE descriptor ------------------------
 * MLatomF with molDescrType=permuted permInvKernel: the descriptor of a geometry is the concatenation of the RE
 * descriptors of its Nperm atom-permuted copies (calcX, molDescr.f90:128-160; permuteAtoms :197-220: permutedXYZ(:,ii) =
 * xyz(:,perm(ii)); the equilibrium distances are permuted the same way, :82-100), and
 *     k(M_i, M_j) = sum_P Kfunk(X_i[block 1], X_j[block P]) / sqrt(sum_P Kfunk(X_i[1], X_i[P]) sum_P Kfunk(X_j[1], X_j[P]))
 * (kernelFunction, A_KRR_kernel.f90:773-804).  PARITY UNPINNED: the MLatomF binary is absent, KREG.so has no such path.
 * pmap[P][ii] is the full 0-based atom map of permutation P (mod_permInds, A_KRR_kernel.f90:944-950). */
void kro_perm_descriptors(int nat, long np, const double *xyz, const double *req, int nperm, const int *pmap,
                          double *X /* [np][nperm * D] */) {
  const int D = nat * (nat - 1) / 2;
#pragma omp parallel for schedule(static)
  for (long p = 0; p < np; p++) {
    double *pxyz = (double *)malloc(sizeof(double) * 3 * nat), *peq = (double *)malloc(sizeof(double) * 3 * nat);
    double *reqd = (double *)malloc(sizeof(double) * D);
    for (int P = 0; P < nperm; P++) {
      for (int ii = 0; ii < nat; ii++) /* permuteAtoms, molDescr.f90:210-218 */
        for (int t = 0; t < 3; t++) {
          pxyz[3 * ii + t] = xyz[((size_t)p * nat + pmap[P * nat + ii]) * 3 + t];
          peq[3 * ii + t] = req[3 * pmap[P * nat + ii] + t];
        }
      kro_xeq(nat, peq, reqd);                                             /* getReq on the permuted eq. geometry */
      kro_descriptors(nat, 1, pxyz, reqd, X + ((size_t)p * nperm + P) * D); /* calcMolDescr with ReqMod = that block */
    }
    free(pxyz);
    free(peq);
    free(reqd);
  }
}

/* kernelFunction, permInvKernel branch (A_KRR_kernel.f90:783-795); Xi, Xj are whole permuted descriptors */
static double perm_kernel_function(int Xsize, int nperm, const double *Xi, const double *Xj, double sigma) {
  double KMiMj = 0.0, KMiMi = 0.0, KMjMj = 0.0;
  for (int PP = 0; PP < nperm; PP++) {
    KMiMj = KMiMj + kernel(Xsize, Xi, Xj + (size_t)Xsize * PP, sigma);
    KMiMi = KMiMi + kernel(Xsize, Xi, Xi + (size_t)Xsize * PP, sigma);
    KMjMj = KMjMj + kernel(Xsize, Xj, Xj + (size_t)Xsize * PP, sigma);
  }
  return KMiMj / (sqrt(KMiMi) * sqrt(KMjMj));
}

/* calcKernel values block / calc_Kvalidate with that kernel function (A_KRR_kernel.f90:149-159, 208-216) */
void kro_perm_kernel_block(int Xsize, int nperm, long na, const double *Xa, long nb, const double *Xb, double sigma,
                           double *K) {
  const size_t XV = (size_t)Xsize * nperm;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < na; i++)
    for (long j = 0; j < nb; j++) K[i * nb + j] = perm_kernel_function(Xsize, nperm, Xa + i * XV, Xb + j * XV, sigma);
}

/* Yest_KRR (A_KRR_kernel.f90:225-245, values part) and the first-order part of YgradXYZest_KRR's permInvKernel branch
 * (:414-503) for a values-trained model.  inv_pmap[P][a] = mod_inversePermInds(a,P).  E gets prior added
 * (A_KRR.f90:1006).  Loop structure and association of the reference kept. */
void kro_perm_predict(int nat, int nperm, const int *inv_pmap, long ntr, const double *X_tr /* [ntr][nperm*D] */,
                      const double *alpha, double sigma, double prior, long np, const double *xyz_te,
                      const double *X_te /* [np][nperm*D] */, int calc_grad, double *E, double *G /* [np][nat][3] */) {
  const int Xsize = nat * (nat - 1) / 2;
  const size_t XV = (size_t)Xsize * nperm;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < np; i++) {
    const double *Xi = X_te + i * XV, *xyz = xyz_te + (size_t)i * nat * 3;
    double y = 0.0;
    for (long jj = 0; jj < ntr; jj++) y = y + alpha[jj] * perm_kernel_function(Xsize, nperm, Xi, X_tr + jj * XV, sigma);
    E[i] = y + prior;
    if (!calc_grad) continue;
    double *tmp = (double *)calloc((size_t)3 * nat * 4, sizeof(double));
    double *temp_Array1 = tmp, *in1 = tmp + 3 * nat, *dKMiMi = tmp + 6 * nat, *dKMiMj = tmp + 9 * nat;
    double KMiMi = 0.0;
    for (int PP = 0; PP < nperm; PP++) KMiMi = KMiMi + kernel(Xsize, Xi, Xi + (size_t)Xsize * PP, sigma);
    /* dKMiMidMiat_part (:432-451) */
    for (int PP = 0; PP < nperm; PP++) {
      const size_t lb = (size_t)Xsize * PP;
      const double cmf = kernel(Xsize, Xi, Xi + lb, sigma);
      for (int a = 0; a < nat - 1; a++)
        for (int cc = a + 1; cc < nat; cc++) {
          const int dd = pair_index(nat, a, cc);
          const int Pdd = pair_index(nat, inv_pmap[PP * nat + a], inv_pmap[PP * nat + cc]);
          const double R = rij(xyz + 3 * a, xyz + 3 * cc);
          const double temp1 = (Xi[lb + dd] + Xi[Pdd] - Xi[dd] * 2) * Xi[dd] / (R * R);
          for (int tt = 0; tt < 3; tt++) {
            const double temp2 = xyz[3 * cc + tt] - xyz[3 * a + tt];
            in1[3 * a + tt] = in1[3 * a + tt] + temp1 * temp2;
            in1[3 * cc + tt] = in1[3 * cc + tt] - temp1 * temp2;
          }
        }
      for (int k = 0; k < 3 * nat; k++) {
        dKMiMi[k] = dKMiMi[k] + cmf * in1[k];
        in1[k] = 0.0;
      }
    }
    for (int k = 0; k < 3 * nat; k++) dKMiMi[k] = dKMiMi[k] / (sigma * sigma);
    /* loop over the training points (:454-503) */
    for (long jj = 0; jj < ntr; jj++) {
      const double *Xj = X_tr + jj * XV;
      double KMiMj = 0.0, KMjMj = 0.0;
      for (int PP = 0; PP < nperm; PP++) {
        KMiMj = KMiMj + kernel(Xsize, Xi, Xj + (size_t)Xsize * PP, sigma);
        KMjMj = KMjMj + kernel(Xsize, Xj, Xj + (size_t)Xsize * PP, sigma);
      }
      for (int k = 0; k < 3 * nat; k++) dKMiMj[k] = 0.0;
      const double cmf = alpha[jj] / sqrt(KMiMi * KMjMj);
      for (int PP = 0; PP < nperm; PP++) {
        const size_t lb = (size_t)Xsize * PP;
        const double cmf1 = kernel(Xsize, Xi, Xj + lb, sigma);
        for (int a = 0; a < nat - 1; a++)
          for (int cc = a + 1; cc < nat; cc++) {
            const int dd = pair_index(nat, a, cc);
            const double R = rij(xyz + 3 * a, xyz + 3 * cc);
            const double temp1 = (Xj[lb + dd] - Xi[dd]) * Xi[dd] / (R * R);
            for (int tt = 0; tt < 3; tt++) {
              const double temp2 = xyz[3 * cc + tt] - xyz[3 * a + tt];
              in1[3 * a + tt] = in1[3 * a + tt] + temp1 * temp2;
              in1[3 * cc + tt] = in1[3 * cc + tt] - temp1 * temp2;
            }
          }
        for (int k = 0; k < 3 * nat; k++) {
          dKMiMj[k] = dKMiMj[k] + cmf1 * in1[k];
          in1[k] = 0.0;
        }
      }
      for (int k = 0; k < 3 * nat; k++) {
        dKMiMj[k] = dKMiMj[k] / (sigma * sigma);
        temp_Array1[k] = temp_Array1[k] + cmf * (dKMiMj[k] - dKMiMi[k] * KMiMj / KMiMi / 2.0);
      }
    }
    for (int k = 0; k < 3 * nat; k++) G[(size_t)i * nat * 3 + k] = temp_Array1[k];
    free(tmp);
  }
}

int kro_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void kro_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
