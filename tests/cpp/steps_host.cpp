// TEST INFRASTRUCTURE - NOT PRODUCT CODE, never loaded by the package.
//
// The product's headline kernel, tiny-differentiable-simulator_b200/csrc/tds_steps.cu (model-specialised, one warp per tree
// role, lane = environment), compiled FOR THE HOST.  A tile (CTA) is executed by 128 host threads - four role warps x 32 lanes - that meet at a
// host barrier wherever the kernel has __syncthreads / __syncthreads_or; shared memory is one static array.  As for tests/cpp/stepw_host.cpp:
// a checker of the kernel SOURCE for a container without a GPU, not a fallback - nothing outside tests/ builds or loads it.
//   g++ -std=c++17 -O1 -pthread -shared -fPIC -I<csrc> -I<include> -I/usr/local/cuda/include tests/cpp/steps_host.cpp -o tests/cpp/_steps_host.so
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#define TDS_B200_EXACT_RCP 1
#define TDS_STEPS_KERNEL_ONLY 1

namespace emu {
struct Dim { unsigned x, y, z; };
thread_local Dim tIdx, bIdx;
struct Barrier {   // reusable barrier for the four role threads, with an OR-reduction
  std::mutex m; std::condition_variable cv; int count = 0, gen = 0, n = 4; int acc = 0, result = 0;
  int sync_or(int pred) {
    std::unique_lock<std::mutex> lk(m);
    acc |= pred ? 1 : 0;
    const int g = gen;
    if (++count == n) { result = acc; acc = 0; count = 0; ++gen; cv.notify_all(); return result; }
    cv.wait(lk, [&] { return gen != g; });
    return result;
  }
};
Barrier* g_bar = nullptr;
int g_force_or = 0;   // 1: every __syncthreads_or is true, as when ANOTHER lane of the tile has a contact
alignas(16) char* g_smem = nullptr;
}  // namespace emu
#define threadIdx emu::tIdx
#define blockIdx emu::bIdx
#define __syncthreads() ((void)emu::g_bar->sync_or(0))
#define __syncthreads_or(p) (emu::g_bar->sync_or(((p) ? 1 : 0) | emu::g_force_or))
#define __shfl_sync(mask, v, lane) (v)
#define clock64() (0LL)
#undef __shared__
#define __shared__
#undef __constant__
#define __constant__
#undef __grid_constant__
#define __grid_constant__
#undef __global__
#define __global__
#undef __launch_bounds__
#define __launch_bounds__(...)
#define smem_raw emu_smem_raw

#ifdef TDSEMU_SPEC_TEST   // -I<dir with spec_test.h>: a model compiled by csrc/gen_spec.cpp for this build (random robots of the tests)
#include "spec_test.h"
#endif
#include "../../tiny-differentiable-simulator_b200/csrc/tds_steps.cu"
#ifdef TDSEMU_SPEC_TEST
TDS_SPEC_TABLES(SpecTest, test)
typedef SpecTest SpecB;   // spec index 1
#else
typedef SpecAnt SpecB;
#endif

namespace tdss { alignas(16) char emu_smem_raw[1024 * 1024]; }   // the tile's shared memory (block-scope extern in the kernel)

namespace {
int g_whole_tile = 0;
// whole_tile 1: a tile = one CTA of 128 host threads, all alive at once - exact (the host-layout instance stages the tile's actions /
// observations cooperatively; __syncthreads_or is tile-wide) but slow on the host.  0: one lane at a time, its four role threads only
// (the lanes of the device-layout instances never talk to each other, only the roles do; the tile-wide OR is covered by force_or).
int g_rc = 0;
template <class SP, typename RA, typename RC, typename RS, int VAR>
void run(const SimParams& P, const EnvParams& E, const StepIO& io, int mode, int use_pd) {
  // (instances whose tile exceeds the 227 KB a CTA can have on the device are never selected by the library - see
  // tdsemu_steps_tile_bytes - but their arithmetic can still be checked here: the host "shared memory" is 1 MB)
  if ((size_t)tdss::Lay<SP, RA, RC, RS>::TOTAL * 32 * 4 > sizeof(tdss::emu_smem_raw)) { g_rc = -2; return; }
  const int tiles = (io.n + 31) / 32;
  for (int t = 0; t < tiles; ++t)
    for (int lane0 = 0; lane0 < (g_whole_tile ? 1 : 32); ++lane0) {
      if (!g_whole_tile && t * 32 + lane0 >= io.n) continue;
      emu::Barrier bar;
      bar.n = g_whole_tile ? 128 : 4;
      emu::g_bar = &bar;
      std::vector<std::thread> th;
      th.reserve(bar.n);
      for (int i = 0; i < bar.n; ++i) {
        const int tid = g_whole_tile ? i : i * 32 + lane0;
        th.emplace_back([&, tid] {
          emu::tIdx = {(unsigned)tid, 0, 0};
          emu::bIdx = {(unsigned)t, 0, 0};
          tdss::tds_step_spec_kernel<SP, RA, RC, RS, VAR, 1>(P, E, io, mode, use_pd);
        });
      }
      for (auto& x : th) x.join();
    }
}
}  // namespace

extern "C" {
// spec: 0 Laikago, 1 Ant.  params: dt, g[3], friction, restitution, erp, cfm, pgs_iterations, keep_all (10 doubles);
// env (null without PD): n_act, start_link, kp, kd, max_force, action_limit, reward_kind, auto_reset, poses[n_act], reset_q[n_q].
// precision 0 mixed / 1 fp64 / 2 fp32.  var: 0 general (qdd, contact distances, link transforms), 1 lean, 2 lean with the host
// layouts (actions [n][n_act] in, observations [n][n_q + n_qd] | reward [n] | done [n] out; needs whole_tile).
// flags: bit 0 force_or (see emu::g_force_or), bit 1 whole_tile (see run()).  Outputs may be null.
int tdsemu_steps(int spec, const double* params, const double* env, int precision, int var, int mode, int use_pd, int flags, int n,
                 const double* q, const double* qd, const double* tau, double* q_out, double* qd_out, double* qdd_out,
                 double* reward, double* done, double* contact_dist, double* link_xf, double* obs_aos, double* obs_tail) {
  emu::g_force_or = flags & 1;
  g_rc = 0;
  g_whole_tile = (flags >> 1) & 1;
  if (var == 2 && !g_whole_tile) return -1;
  SimParams P;
  memset(&P, 0, sizeof(P));
  P.dt = params[0]; P.inv_dt = 1.0 / params[0];
  for (int k = 0; k < 3; ++k) P.gravity[k] = params[1 + k];
  P.friction = params[4]; P.restitution = params[5]; P.erp = params[6]; P.cfm = params[7];
  P.pgs_iterations = (int)params[8]; P.keep_all_points = (int)params[9];
  EnvParams E;
  memset(&E, 0, sizeof(E));
  const int n_q = spec == 0 ? SpecLaikago::N_Q : SpecB::N_Q, n_qd = spec == 0 ? SpecLaikago::N_QD : SpecB::N_QD;
  const int n_act_model = spec == 0 ? SpecLaikago::N_ACT : SpecB::N_ACT;
  const int n_cand = spec == 0 ? SpecLaikago::N_CAND : SpecB::N_CAND, n_links = spec == 0 ? SpecLaikago::N_LINKS : SpecB::N_LINKS;
  if (env) {
    E.n_act = (int)env[0]; E.start_link = (int)env[1];
    E.kp = (float)env[2]; E.kd = (float)env[3]; E.max_force = (float)env[4]; E.action_limit = (float)env[5];
    E.reward_kind = (int)env[6]; E.auto_reset = (int)env[7];
    for (int k = 0; k < E.n_act; ++k) { E.initial_poses[k] = (float)env[8 + k]; E.act_link[k] = spec == 0 ? SpecLaikago::ACT_LINK[k] : SpecB::ACT_LINK[k]; }
    for (int k = 0; k < n_q; ++k) E.reset_q[k] = (float)env[8 + E.n_act + k];
  }
  const int floating = spec == 0 ? SpecLaikago::FLOATING : SpecB::FLOATING;
  const int ns = (n + 31) & ~31, n_in = use_pd ? n_act_model : n_qd - (floating ? 6 : 0), n_obs = n_q + n_qd;
  std::vector<float> sq((size_t)n_q * ns), sqd((size_t)n_qd * ns), st((size_t)n_in * ns, 0.f), oq(sq.size()), oqd(sqd.size()), oqdd(sqd.size()), orew(ns), odone(ns);
  std::vector<float> ocd((size_t)n_cand * ns), oxf((size_t)n_links * 12 * ns), aos((size_t)n_in * ns, 0.f), oobs((size_t)n_obs * ns), otail(2 * (size_t)ns);
  for (int e = 0; e < n; ++e) {
    for (int k = 0; k < n_q; ++k) sq[(size_t)k * ns + e] = (float)q[(size_t)e * n_q + k];
    for (int k = 0; k < n_qd; ++k) sqd[(size_t)k * ns + e] = (float)qd[(size_t)e * n_qd + k];
    if (tau) for (int k = 0; k < n_in; ++k) { st[(size_t)k * ns + e] = (float)tau[(size_t)e * n_in + k]; aos[(size_t)e * n_in + k] = (float)tau[(size_t)e * n_in + k]; }
  }
  StepIO io;
  memset(&io, 0, sizeof(io));
  io.q_in = sq.data(); io.qd_in = sqd.data(); io.tau_in = st.data();
  io.q_out = oq.data(); io.qd_out = oqd.data(); io.qdd_out = var == 0 ? oqdd.data() : nullptr;
  io.reward = orew.data(); io.done = odone.data();
  if (var == 0) { io.contact_dist = contact_dist ? ocd.data() : nullptr; io.link_xf = link_xf ? oxf.data() : nullptr; }
  if (var == 2) { io.act_aos = aos.data(); io.obs_aos = oobs.data(); io.obs_tail = otail.data(); }
  io.n = n; io.n_stride = ns;
#ifdef TDSEMU_SPEC_TEST   // (only the general instance, to keep the build of a test-time model short)
#define TDSEMU_LEAN(SP, A, C, S) else g_rc = -3;
#else
#define TDSEMU_LEAN(SP, A, C, S) else if (var == 1) run<SP, A, C, S, 1>(P, E, io, mode, use_pd); else run<SP, A, C, S, 2>(P, E, io, mode, use_pd);
#endif
#define RUN3(SP, A, C, S)                                                                         \
  do {                                                                                            \
    if (var == 0) run<SP, A, C, S, 0>(P, E, io, mode, use_pd);                                    \
    TDSEMU_LEAN(SP, A, C, S)                                                                      \
  } while (0)
#define RUN(SP)                                                                                   \
  do {                                                                                            \
    if (precision == 0) RUN3(SP, float, double, float);                                           \
    else if (precision == 1) RUN3(SP, double, double, double);                                    \
    else RUN3(SP, float, float, float);                                                           \
  } while (0)
  if (spec == 0) RUN(SpecLaikago); else RUN(SpecB);
#undef RUN
#undef RUN3
  for (int e = 0; e < n; ++e) {
    if (q_out) for (int k = 0; k < n_q; ++k) q_out[(size_t)e * n_q + k] = oq[(size_t)k * ns + e];
    if (qd_out) for (int k = 0; k < n_qd; ++k) qd_out[(size_t)e * n_qd + k] = oqd[(size_t)k * ns + e];
    if (qdd_out && var == 0) for (int k = 0; k < n_qd; ++k) qdd_out[(size_t)e * n_qd + k] = oqdd[(size_t)k * ns + e];
    if (reward) reward[e] = orew[e];
    if (done) done[e] = odone[e];
    if (contact_dist && var == 0) for (int k = 0; k < n_cand; ++k) contact_dist[(size_t)e * n_cand + k] = ocd[(size_t)k * ns + e];
    if (link_xf && var == 0) for (int k = 0; k < n_links * 12; ++k) link_xf[(size_t)e * n_links * 12 + k] = oxf[(size_t)k * ns + e];
    if (obs_aos && var == 2) for (int k = 0; k < n_obs; ++k) obs_aos[(size_t)e * n_obs + k] = oobs[(size_t)e * n_obs + k];
    if (obs_tail && var == 2) { obs_tail[e] = otail[e]; obs_tail[n + e] = otail[n + e]; }
  }
  return g_rc ? g_rc : n_cand;
}
// bytes of shared memory one tile of the instance needs (the library selects the kernel only if this fits the device's opt-in limit)
long tdsemu_steps_tile_bytes(int spec, int precision) {
#define TB(SP) (precision == 0 ? (long)tdss::Lay<SP, float, double, float>::TOTAL : precision == 1 ? (long)tdss::Lay<SP, double, double, double>::TOTAL : (long)tdss::Lay<SP, float, float, float>::TOTAL) * 32 * 4
  return spec == 0 ? TB(SpecLaikago) : TB(SpecB);
#undef TB
}
}  // extern "C"
