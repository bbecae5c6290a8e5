// TEST INFRASTRUCTURE - NOT PRODUCT CODE.
//
// C-ABI shim around the UNMODIFIED reference Laikago environment step, compiled in place
// from /root/reference (see oracle/build_ref.sh).  It exposes the two CPU implementations
// of the north-star hot path that the reference itself ships:
//   (1) LocomotionContactSimulation::step_forward_original
//       examples/environments/locomotion_contact_simulation.h:151-304   (templated path)
//   (2) omp_model_laikago_forward_zero_kernel<double>
//       examples/environments/omp_model_laikago_forward_zero.h          (its codegen path,
//       what OpenMPForwardStepper runs, examples/ars/ars_vectorized_environment.h:110-137)
// Both take the 51-double input  q18 | qd18 | action12 | kp,kd,max_force  and write the
// 411-double output  q18 | qd18 | 17 x (pos3, quat4) | up.z | zeros.
// The URDFs come from the string headers the reference embeds
// (laikago_toes_zup_xyz_xyzrot.h, plane_implicit_urdf.h), so this works without data files.
#include <cstdio>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "math/tiny/tiny_double_utils.h"
#include "math/tiny/tiny_algebra.hpp"
#include "environments/laikago_environment2.h"

typedef TinyAlgebra<double, TINY::DoubleUtils> A64;

namespace {
struct LaikagoRef {
  std::vector<LaikagoContactSimulation<A64>*> sims;  // one per thread: step mutates mb_
  int in_dim = 0, out_dim = 0;
};
}  // namespace

extern "C" {

void* tdsref_laikago_create(int num_threads) {
  if (num_threads < 1) num_threads = 1;
  LaikagoRef* r = new LaikagoRef;
  for (int t = 0; t < num_threads; ++t) {
    r->sims.push_back(new LaikagoContactSimulation<A64>(
        false, "laikago/laikago_toes_zup_xyz_xyzrot.urdf", laikago_toes_zup_xyz_xyzrot,
        LaikagoContactSimulation<A64>::get_initial_poses(), false));
  }
  r->in_dim = r->sims[0]->input_dim_with_action_and_variables();
  r->out_dim = r->sims[0]->output_dim();
  return r;
}

void tdsref_laikago_destroy(void* p) {
  LaikagoRef* r = (LaikagoRef*)p;
  if (!r) return;
  for (auto* s : r->sims) delete s;
  delete r;
}

int tdsref_laikago_input_dim(void* p) { return ((LaikagoRef*)p)->in_dim; }
int tdsref_laikago_output_dim(void* p) { return ((LaikagoRef*)p)->out_dim; }
int tdsref_laikago_num_threads(void* p) { return (int)((LaikagoRef*)p)->sims.size(); }

// impl 0 = templated step_forward_original, impl 1 = the reference's codegen kernel.
// input  [n][in_dim], output [n][out_dim]  (AoS, fp64: the layout of C-ABI v1,
// src/utils/cuda_codegen.hpp:164-266).
void tdsref_laikago_step(void* p, int impl, int n, const double* input, double* output) {
  LaikagoRef* r = (LaikagoRef*)p;
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
      omp_model_laikago_forward_zero_kernel<double>(1, output + (size_t)i * out_dim,
                                                    input + (size_t)i * in_dim);
    }
  }
}

// Export the flat model of the Laikago robot as loaded by the reference (see ref_core.cpp
// for the exporter; duplicated minimal version is avoided by re-exporting through ref_core's
// URDF path in the container).  Here: reward/done of the reference env
// (examples/environments/laikago_environment2.h:130-171).
void tdsref_laikago_reward_done(void* p, const double* cur_state, double* reward, int* done) {
  LaikagoRef* r = (LaikagoRef*)p;
  std::vector<double> prev(r->out_dim, 0.0), cur(cur_state, cur_state + r->out_dim);
  double rew = 0;
  bool d = false;
  r->sims[0]->compute_reward_done(prev, cur, rew, d);
  *reward = rew;
  *done = d ? 1 : 0;
}

}  // extern "C"
