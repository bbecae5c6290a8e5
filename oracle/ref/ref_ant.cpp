// TEST INFRASTRUCTURE - NOT PRODUCT CODE.
//
// C-ABI shim around the UNMODIFIED reference Ant environment step (the second vectorized environment the
// reference binds, python/pytinydiffsim_includes.h:58-141), compiled in place from /root/reference
// (see oracle/build_ref.sh).  Same shape as ref_laikago.cpp:
//   impl 0: LocomotionContactSimulation::step_forward_original (examples/environments/locomotion_contact_simulation.h:151-304)
//           through AntContactSimulation2 (examples/environments/ant_environment2.h:28-105): dt 0.01, kp 15, kd 0.3, max 3
//   impl 1: omp_model_ant_forward_zero_kernel<double> (examples/environments/omp_model_ant_forward_zero.h)
// Input 31 doubles  q14 | qd14 | ... as the env defines them (queried at run time), URDF from the embedded string header.
#include <cstdio>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "math/tiny/tiny_double_utils.h"
#include "math/tiny/tiny_algebra.hpp"
#include "environments/ant_environment2.h"

typedef TinyAlgebra<double, TINY::DoubleUtils> A64;

namespace {
struct AntRef {
  std::vector<AntContactSimulation2<A64>*> sims;  // one per thread: step mutates mb_
  int in_dim = 0, out_dim = 0, state_dim = 0, act_dim = 0;
};
}  // namespace

extern "C" {

void* tdsref_ant_create(int num_threads) {
  if (num_threads < 1) num_threads = 1;
  AntRef* r = new AntRef;
  for (int t = 0; t < num_threads; ++t)
    r->sims.push_back(new AntContactSimulation2<A64>(false, "gym/ant_org_xyz_xyzrot.urdf", ant_org_xyz_xyzrot,
                                                     AntContactSimulation2<A64>::get_initial_poses(), false));
  r->in_dim = r->sims[0]->input_dim_with_action_and_variables();
  r->out_dim = r->sims[0]->output_dim();
  r->state_dim = r->sims[0]->input_dim();
  r->act_dim = r->sims[0]->action_dim();
  return r;
}

void tdsref_ant_destroy(void* p) {
  AntRef* r = (AntRef*)p;
  if (!r) return;
  for (auto* s : r->sims) delete s;
  delete r;
}

int tdsref_ant_input_dim(void* p) { return ((AntRef*)p)->in_dim; }
int tdsref_ant_output_dim(void* p) { return ((AntRef*)p)->out_dim; }
int tdsref_ant_state_dim(void* p) { return ((AntRef*)p)->state_dim; }
int tdsref_ant_action_dim(void* p) { return ((AntRef*)p)->act_dim; }
int tdsref_ant_num_threads(void* p) { return (int)((AntRef*)p)->sims.size(); }

void tdsref_ant_step(void* p, int impl, int n, const double* input, double* output) {
  AntRef* r = (AntRef*)p;
  const int in_dim = r->in_dim, out_dim = r->out_dim;
  const int nt = (int)r->sims.size();
#pragma omp parallel for num_threads(nt) schedule(static)
  for (int i = 0; i < n; ++i) {
#ifdef _OPENMP
    int t = omp_get_thread_num();
#else
    int t = 0;
#endif
    if (impl == 0) {
      std::vector<double> v(input + (size_t)i * in_dim, input + (size_t)(i + 1) * in_dim);
      std::vector<double> out(out_dim, 0.0);
      r->sims[t]->step_forward_original(v, out);
      for (int k = 0; k < out_dim; ++k) output[(size_t)i * out_dim + k] = out[k];
    } else {
      omp_model_ant_forward_zero_kernel<double>(1, output + (size_t)i * out_dim, input + (size_t)i * in_dim);
    }
  }
}

// reward / done of the reference env (ant_environment2.h:75-105); needs the previous and the current state
void tdsref_ant_reward_done(void* p, const double* prev_state, const double* cur_state, double* reward, int* done) {
  AntRef* r = (AntRef*)p;
  std::vector<double> prev(prev_state, prev_state + r->state_dim), cur(cur_state, cur_state + r->state_dim);
  double rew = 0;
  bool d = false;
  r->sims[0]->compute_reward_done(prev, cur, rew, d);
  *reward = rew;
  *done = d ? 1 : 0;
}

}  // extern "C"
