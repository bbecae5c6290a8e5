// TEST INFRASTRUCTURE - NOT PRODUCT CODE, never loaded by the package.
//
// The product's generic step kernel, tiny-differentiable-simulator_b200/csrc/tds_stepw.cu, compiled FOR THE HOST: the CUDA
// built-ins it uses (threadIdx / blockIdx, warp votes, __syncwarp, clock64, extern shared memory) are given single-lane host
// meanings and the kernel body is called as an ordinary function, one "thread" after the other.  The kernel keeps one
// lane per environment and its warp collectives only steer warp-uniform shortcuts (skip the contact solve when no lane
// touches), so lane-by-lane execution computes the same numbers.  This lets the CPU test-suite (no GPU in the build
// container) execute the very source the GPU runs - kinematics, ABA, contacts, CRBA, the blocked solves, the spring-damper
// branch, the dual-number instance - against the reference.  It is a checker for the kernel source, not a fallback:
// nothing outside tests/ builds or loads it.
//   g++ -std=c++17 -O1 -shared -fPIC -I<csrc> -I<include> -I/usr/local/cuda/include tests/cpp/stepw_host.cpp -o tests/cpp/_stepw_host.so
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define TDS_B200_EXACT_RCP 1
#define TDS_STEPW_KERNEL_ONLY 1
struct EmuDim { unsigned x, y, z; };
static thread_local EmuDim emu_threadIdx, emu_blockIdx, emu_blockDim, emu_gridDim;
#define threadIdx emu_threadIdx
#define blockIdx emu_blockIdx
#define blockDim emu_blockDim
#define gridDim emu_gridDim
#define __any_sync(mask, pred) ((pred) ? 1 : 0)
#define __reduce_max_sync(mask, v) (v)
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
#define __syncwarp() ((void)0)
#define clock64() (0LL)
#undef __shared__
#define __shared__
#undef __grid_constant__
#define __grid_constant__
#undef __global__
#define __global__
#undef __launch_bounds__
#define __launch_bounds__(...)
alignas(16) char smem_raw[16];

#include "tds_model.h"
#include "../../tiny-differentiable-simulator_b200/csrc/tds_stepw.cu"

namespace {
template <typename RA, typename RC, typename RS, typename RQ>
void run_grid(const DevModel& M, const SimParams& P, const EnvParams& E, const StepIO& io, int mode, int use_pd, int n_dirs, char* scratch) {
  const int warps = (io.n + 31) / 32;
  emu_blockDim = {32, 1, 1};
  emu_gridDim = {(unsigned)warps, (unsigned)n_dirs, 1};
  for (unsigned by = 0; by < (unsigned)n_dirs; ++by)
    for (unsigned bx = 0; bx < (unsigned)warps; ++bx)
      for (unsigned t = 0; t < 32; ++t) {
        if ((int)(bx * 32 + t) >= io.n) continue;        // (padding lanes recompute the last environment on the GPU)
        emu_blockIdx = {bx, by, 0};
        emu_threadIdx = {t, 0, 0};
        tdsw::tds_stepw_kernel<RA, RC, RS, RQ, false>(M, P, E, io, mode, use_pd, scratch);
      }
}
}  // namespace

extern "C" {

// params: dt, g[3], friction, restitution, erp, cfm, pgs_iterations, keep_all, contact_model, spring_k, damper_d, exponent_n,
//         v_transition, hard_contact (16 doubles);  env: n_act, start_link, kp, kd, max_force, action_limit, poses[n_act]
// precision: 0 mixed, 1 fp64, 2 fp32.  q [n][n_q], qd [n][n_qd], tau_or_action [n][n_tau | n_act] (may be null).
// outputs (may be null): q_out, qd_out, qdd_out [n][n_qd], contact_dist [n][n_points], jac [n][rows][cols] (dual numbers).
int tdsemu_stepw(const double* model, int n_model, const double* params, const double* env, int precision, int mode, int use_pd,
                 int n, const double* q, const double* qd, const double* tau, double* q_out, double* qd_out, double* qdd_out,
                 double* contact_dist, double* jac) {
  DevModel* D = new DevModel;
  int rc = tds_build_dev_model(model, n_model, D);
  if (rc) { delete D; return rc; }
  const int sizes[3][3] = {{4, 8, 4}, {8, 8, 8}, {4, 4, 4}};
  const bool ad = jac != nullptr;
  if (ad) tds_build_layout_w(D, 16, 16, 16, -1, 16);
  else tds_build_layout_w(D, sizes[precision][0], sizes[precision][1], sizes[precision][2], -1);
  SimParams P;
  memset(&P, 0, sizeof(P));
  P.dt = params[0]; P.inv_dt = 1.0 / params[0];
  for (int k = 0; k < 3; ++k) P.gravity[k] = params[1 + k];
  P.friction = params[4]; P.restitution = params[5]; P.erp = params[6]; P.cfm = params[7];
  P.pgs_iterations = (int)params[8]; P.keep_all_points = (int)params[9];
  P.contact_model = (int)params[10]; P.spring_k = params[11]; P.damper_d = params[12]; P.exponent_n = params[13];
  P.v_transition = params[14]; P.hard_contact_condition = (int)params[15];
  EnvParams E;
  memset(&E, 0, sizeof(E));
  if (env) {   // tds_b200_set_env (tds_capi.cu): action k drives the k-th non-fixed link at or after start_link
    E.n_act = (int)env[0]; E.start_link = (int)env[1];
    E.kp = (float)env[2]; E.kd = (float)env[3]; E.max_force = (float)env[4]; E.action_limit = (float)env[5];
    int k = 0;
    for (int i = D->floating ? 0 : E.start_link; i < D->n_links && k < E.n_act; ++i) {
      if (D->flags[i] & TDS_LF_FIXED) continue;
      E.act_link[k] = i; E.initial_poses[k] = (float)env[6 + k]; ++k;
    }
  }
  const int ns = (n + 31) & ~31, n_q = D->n_q, n_qd = D->n_qd;
  const int n_tau = n_qd - (D->floating ? 6 : 0), n_in = use_pd ? E.n_act : n_tau;
  std::vector<float> sq((size_t)(n_q > 0 ? n_q : 1) * ns), sqd((size_t)(n_qd > 0 ? n_qd : 1) * ns), st((size_t)(n_in > 0 ? n_in : 1) * ns, 0.f);
  std::vector<float> oq(sq.size()), oqd(sqd.size()), oqdd(sqd.size()), ocd((size_t)(D->max_contacts + D->n_pair_points + 1) * ns);
  for (int e = 0; e < n; ++e) {
    for (int k = 0; k < n_q; ++k) sq[(size_t)k * ns + e] = (float)q[(size_t)e * n_q + k];
    for (int k = 0; k < n_qd; ++k) sqd[(size_t)k * ns + e] = (float)qd[(size_t)e * n_qd + k];
    if (tau) for (int k = 0; k < n_in; ++k) st[(size_t)k * ns + e] = (float)tau[(size_t)e * n_in + k];
  }
  StepIO io;
  memset(&io, 0, sizeof(io));
  io.q_in = sq.data(); io.qd_in = sqd.data(); io.tau_in = (tau || use_pd) ? st.data() : nullptr;
  io.q_out = oq.data(); io.qd_out = oqd.data(); io.qdd_out = oqdd.data();
  io.contact_dist = contact_dist ? ocd.data() : nullptr;
  io.n = n; io.n_stride = ns;
  const int rows = mode == 0 ? n_qd : n_q + n_qd, cols = n_q + n_qd + (use_pd ? E.n_act + 3 : n_tau);
  std::vector<double> jbuf;
  int n_dirs = 1;
  if (ad) { jbuf.assign((size_t)rows * cols * ns, 0.0); io.jac = jbuf.data(); io.jac_n_in = cols; io.jac_dir0 = 0; n_dirs = cols; }
  std::vector<char> scratch((size_t)n_dirs * ((n + 31) / 32) * D->x_total * 32 * 4 + 64);
  typedef tds::Dual<double> DD;
  if (ad) run_grid<DD, DD, DD, DD>(*D, P, E, io, mode, use_pd, n_dirs, scratch.data());
  else if (precision == 0) run_grid<float, double, float, float>(*D, P, E, io, mode, use_pd, 1, scratch.data());
  else if (precision == 1) run_grid<double, double, double, float>(*D, P, E, io, mode, use_pd, 1, scratch.data());
  else run_grid<float, float, float, float>(*D, P, E, io, mode, use_pd, 1, scratch.data());
  for (int e = 0; e < n; ++e) {
    if (q_out) for (int k = 0; k < n_q; ++k) q_out[(size_t)e * n_q + k] = oq[(size_t)k * ns + e];
    if (qd_out) for (int k = 0; k < n_qd; ++k) qd_out[(size_t)e * n_qd + k] = oqd[(size_t)k * ns + e];
    if (qdd_out) for (int k = 0; k < n_qd; ++k) qdd_out[(size_t)e * n_qd + k] = oqdd[(size_t)k * ns + e];
    if (contact_dist) for (int k = 0; k < D->max_contacts + D->n_pair_points; ++k) contact_dist[(size_t)e * (D->max_contacts + D->n_pair_points) + k] = ocd[(size_t)k * ns + e];
    if (ad) for (int k = 0; k < rows * cols; ++k) jac[(size_t)e * rows * cols + k] = jbuf[(size_t)k * ns + e];
  }
  const int npts = D->max_contacts + D->n_pair_points;   // plane candidates, then the candidates between multibodies
  delete D;
  return ad ? rows * 1000 + cols : npts;
}

}  // extern "C"
