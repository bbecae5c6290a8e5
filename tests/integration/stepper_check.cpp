// TEST INFRASTRUCTURE.  The reference's own vectorized environment (examples/ars/ars_vectorized_environment.h) stepped
// through ITS plugin point - VectorizedEnvironment::CustomForwardDynamicsStepper, installed in default_stepper_ - once
// with the reference's serial CPU stepper and once with a stepper that forwards to libtds_b200.so; the two outputs
// are compared.  Compiled in the build container against the reference headers (oracle/build_integration.sh), the
// binary travels to the GPU box; it reads the flat model from a file (the URDFs do not exist there).
//   stepper_check.bin <laikago_model.bin> [n_envs]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "math/tiny/tiny_double_utils.h"
#include "math/tiny/tiny_algebra.hpp"
#include "environments/laikago_environment2.h"
#include "ars/ars_vectorized_environment.h"
#include "tds_b200.h"

// the profiling hooks of the reference's visualizer utilities (tiny_logging.h:31-32) are not part of this build
void TinyEnterProfileZone(const char*) {}
void TinyLeaveProfileZone() {}

typedef TinyAlgebra<double, TINY::DoubleUtils> Algebra;
typedef LaikagoContactSimulation<Algebra> Sim;
typedef VectorizedEnvironment<Algebra, Sim> VecEnv;

// the stepper INTEGRATION.md section 2 shows, fed with a prebuilt flat model instead of URDF paths
struct B200Stepper : VecEnv::CustomForwardDynamicsStepper {
  typedef double Scalar;
  tds_b200_sim* sim = nullptr;
  int n = 0, nq = 0, nqd = 0, nact = 0;
  std::vector<double> q, qd, act, q2, qd2;
  B200Stepper(const std::vector<double>& model, int batch, const Sim& s) : n(batch) {
    sim = tds_b200_create(model.data(), (int)model.size(), batch, /*device*/ 0);
    if (!sim) { fprintf(stderr, "tds_b200_create: %s\n", tds_b200_last_error()); exit(2); }
    const double g[3] = {0, 0, -9.81};
    tds_b200_set_params(sim, 1e-3, g, /*friction*/ 1.0, 0.0, 0.2, 1e-5, 1, /*keep_all_points*/ 1);
    std::vector<double> poses(s.initial_poses_.begin(), s.initial_poses_.end());
    tds_b200_set_env(sim, (int)poses.size(), poses.data(), s.base_dof_, 100, 2, 50, 0.4, 1);
    int d[8];
    tds_b200_get_dims(sim, d);
    nq = d[2]; nqd = d[3]; nact = d[7];
    q.resize((size_t)n * nq); qd.resize((size_t)n * nqd); act.resize((size_t)n * nact); q2 = q; qd2 = qd;
  }
  ~B200Stepper() { tds_b200_destroy(sim); }
  void step(const std::vector<std::vector<Scalar>>& in, std::vector<std::vector<Scalar>>& out, std::vector<bool>& dones, int,
            const std::vector<Scalar>&) override {
    (void)dones;
    for (int i = 0; i < n; ++i) {   // in[i] = q | qd | action | kp, kd, max_force
      std::copy(in[i].begin(), in[i].begin() + nq, q.begin() + (size_t)i * nq);
      std::copy(in[i].begin() + nq, in[i].begin() + nq + nqd, qd.begin() + (size_t)i * nqd);
      std::copy(in[i].begin() + nq + nqd, in[i].begin() + nq + nqd + nact, act.begin() + (size_t)i * nact);
    }
    if (tds_b200_step_host(sim, TDS_B200_MODE_FULL, /*use_pd*/ 1, q.data(), qd.data(), act.data(), q2.data(), qd2.data(), nullptr,
                           nullptr)) {
      fprintf(stderr, "tds_b200_step_host: %s\n", tds_b200_last_error());
      exit(3);
    }
    for (int i = 0; i < n; ++i) {
      std::copy(q2.begin() + (size_t)i * nq, q2.begin() + (size_t)(i + 1) * nq, out[i].begin());
      std::copy(qd2.begin() + (size_t)i * nqd, qd2.begin() + (size_t)(i + 1) * nqd, out[i].begin() + nq);
    }
  }
};

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: stepper_check.bin <laikago_model.bin> [n_envs]\n"); return 2; }
  const int n = argc > 2 ? atoi(argv[2]) : 64;
  std::vector<double> model;
  {
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    double v;
    while (fread(&v, sizeof(double), 1, f) == 1) model.push_back(v);
    fclose(f);
  }
  Sim sim(false, "laikago/laikago_toes_zup_xyz_xyzrot.urdf", laikago_toes_zup_xyz_xyzrot, Sim::get_initial_poses(), false);
  VecEnv vec_env(sim, n);
  // perturbed standing states, the same for both steppers
  const int in_dim = sim.input_dim_with_action_and_variables(), out_dim = sim.output_dim();
  std::vector<std::vector<double>> inputs(n, std::vector<double>(in_dim, 0.0));
  srand(12345);
  auto U = [](double a, double b) { return a + (b - a) * (rand() / (double)RAND_MAX); };
  for (int i = 0; i < n; ++i) {
    std::vector<double>& x = inputs[i];
    x[0] = U(-0.2, 0.2); x[1] = U(-0.2, 0.2); x[2] = U(0.36, 0.50);
    for (int k = 3; k < 6; ++k) x[k] = U(-0.25, 0.25);
    for (int k = 0; k < 12; ++k) x[6 + k] = sim.initial_poses_[k] + U(-0.15, 0.15);
    for (int k = 0; k < 18; ++k) x[18 + k] = U(-1, 1);
    for (int k = 0; k < 12; ++k) x[36 + k] = U(-0.5, 0.5);
    x[48] = 100; x[49] = 2; x[50] = 50;
  }
  std::vector<bool> dones(n, false);
  std::vector<std::vector<double>> ref_out(n, std::vector<double>(out_dim, 0.0)), our_out = ref_out;
  vec_env.default_stepper_ = &vec_env.serial_stepper_;          // the reference's CPU stepper
  vec_env.default_stepper_->step(inputs, ref_out, dones);
  B200Stepper b200(model, n, sim);
  vec_env.default_stepper_ = &b200;                              // the plugin point (ars_vectorized_environment.h:154)
  vec_env.default_stepper_->step(inputs, our_out, dones);
  double worst = 0;
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 36; ++k) {
      const double e = std::fabs(our_out[i][k] - ref_out[i][k]) / std::fmax(1.0, std::fabs(ref_out[i][k]));
      if (e > worst) worst = e;
    }
  printf("stepper_check: %d envs, max rel err of q | qd vs the reference's serial stepper = %.3e\n", n, worst);
  return worst <= 1e-5 ? 0 : 1;
}
