// kreg_md.cu -- velocity-Verlet bookkeeping around kreg_predict for batched, device-resident molecular dynamics
// (SURVEY.md 8f row 3).  The reference propagates one trajectory at a time in numpy (mlatom/md.py:158-205; batch
// driver md_parallel.py:155-200): per step
//     a = -dE/dx / m * c_acc                     (md.py:171: forces / mass / ram2au * Bohr2Angstrom^2 * fs2au^2)
//     x += v dt + a dt^2 / 2 ;  v += a dt / 2    (md.py:190-193)
//     predict E, dE/dx at the new x              (md.py:198)
//     v += a_new dt / 2                          (md.py:204)
//     E_kin = sum m v^2 / 2 * c_kin              (data.py:862)
// Here the two halves are one kernel each (instead of ~10 elementwise framework kernels per step), so a step of B
// trajectories is: first-half kernel, kreg_predict, second-half kernel -- three nodes of one CUDA graph.
// `active[p]` (optional, 1.0 / 0.0) freezes trajectories the active-learning sampler has stopped (al_utils.py:246-322).
#include "kreg_common.cuh"

namespace kreg {

__global__ void __launch_bounds__(256) md_first_half_kernel(int g3, int64_t total, double dt, double half_dt2, double half_dt,
                                                            const double *__restrict__ acc_scale,  // [3 Nat] c_acc / m_a
                                                            const double *__restrict__ grad, double *__restrict__ xyz,
                                                            double *__restrict__ vel, const double *__restrict__ active) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = e / g3;
    const int c = (int)(e - p * g3);
    if (active && active[p] == 0.0) continue;
    const double a = -grad[e] * acc_scale[c];
    const double v = vel[e];
    xyz[e] = xyz[e] + v * dt + a * half_dt2;  // the reference's order of operations (md.py:190)
    vel[e] = v + a * half_dt;
  }
}

// one warp per trajectory: second half kick, then the kinetic energy by a fixed-shape reduction (run-to-run deterministic)
__global__ void __launch_bounds__(256) md_second_half_kernel(int g3, int64_t ntraj, double half_dt,
                                                             const double *__restrict__ acc_scale,
                                                             const double *__restrict__ kin_scale,  // [3 Nat] c_kin m_a / 2
                                                             const double *__restrict__ grad, double *__restrict__ vel,
                                                             double *__restrict__ ekin, const double *__restrict__ active) {
  const int lane = threadIdx.x & 31;
  const int64_t p = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= ntraj) return;
  const bool frozen = active && active[p] == 0.0;
  double s = 0.0;
  for (int c = lane; c < g3; c += 32) {
    const int64_t e = p * g3 + c;
    double v = vel[e];
    if (!frozen) {
      v = v + (-grad[e] * acc_scale[c]) * half_dt;
      vel[e] = v;
    }
    s = fma(kin_scale[c] * v, v, s);
  }
#pragma unroll
  for (int w = 16; w > 0; w >>= 1) s += __shfl_xor_sync(0xffffffffu, s, w);
  if (lane == 0) ekin[p] = s;
}

// uq = |E - E_aux| (al_utils.py:1418-1424); running trajectories above the threshold stop at this step
__global__ void __launch_bounds__(256) md_uncertainty_kernel(int64_t ntraj, const double *__restrict__ epot,
                                                             const double *__restrict__ eaux, double threshold,
                                                             const int64_t *__restrict__ step_no, double *__restrict__ uq,
                                                             double *__restrict__ active, int64_t *__restrict__ stop_step) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ntraj) return;
  const double u = fabs(epot[p] - eaux[p]);
  uq[p] = u;
  if (u > threshold && active[p] > 0.0) {
    stop_step[p] = *step_no;
    active[p] = 0.0;
  }
}

}  // namespace kreg

using namespace kreg;

extern "C" int kreg_md_first_half(int natoms, int64_t ntraj, double dt, const double *acc_scale, const double *grad,
                                  double *xyz, double *vel, const double *active, void *stream) {
  if (natoms < 1 || ntraj < 0 || !acc_scale || !grad || !xyz || !vel) return KREG_E_BADARG;
  if (ntraj == 0) return KREG_OK;
  const int64_t total = ntraj * 3 * natoms;
  const int64_t blocks = (total + 255) / 256;
  md_first_half_kernel<<<(unsigned)(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, as_stream(stream)>>>(
      3 * natoms, total, dt, 0.5 * dt * dt, 0.5 * dt, acc_scale, grad, xyz, vel, active);
  KREG_LAUNCH_CHECK();
  return KREG_OK;
}

extern "C" int kreg_md_second_half(int natoms, int64_t ntraj, double dt, const double *acc_scale, const double *kin_scale,
                                   const double *grad, double *vel, double *ekin, const double *active, void *stream) {
  if (natoms < 1 || ntraj < 0 || !acc_scale || !kin_scale || !grad || !vel || !ekin) return KREG_E_BADARG;
  if (ntraj == 0) return KREG_OK;
  md_second_half_kernel<<<(unsigned)((ntraj + 7) / 8), 256, 0, as_stream(stream)>>>(3 * natoms, ntraj, 0.5 * dt, acc_scale,
                                                                                    kin_scale, grad, vel, ekin, active);
  KREG_LAUNCH_CHECK();
  return KREG_OK;
}

extern "C" int kreg_md_uncertainty(int64_t ntraj, const double *epot, const double *eaux, double threshold,
                                   const int64_t *step_no, double *uq, double *active, int64_t *stop_step, void *stream) {
  if (ntraj < 0 || !epot || !eaux || !step_no || !uq || !active || !stop_step) return KREG_E_BADARG;
  if (ntraj == 0) return KREG_OK;
  md_uncertainty_kernel<<<(unsigned)((ntraj + 255) / 256), 256, 0, as_stream(stream)>>>(ntraj, epot, eaux, threshold, step_no,
                                                                                        uq, active, stop_step);
  KREG_LAUNCH_CHECK();
  return KREG_OK;
}
