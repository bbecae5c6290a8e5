// TEST INFRASTRUCTURE - NOT PRODUCT CODE, never loaded by the package.
// The rigid-body world kernel (tiny-differentiable-simulator_b200/csrc/tds_rigid.cu: one lane per world, no cooperation between lanes) compiled
// FOR THE HOST and called world after world, like tests/cpp/stepw_host.cpp does for the generic step kernel.
//   g++ -std=c++17 -O1 -shared -fPIC -I<csrc> -I<include> -I/usr/local/cuda/include tests/cpp/rigid_host.cpp -o tests/cpp/_rigid_host.so
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>
#include <vector>

#define TDS_B200_EXACT_RCP 1
#define TDS_RIGID_KERNEL_ONLY 1
namespace emu { struct Dim { unsigned x, y, z; }; static Dim tIdx, bIdx, bDim; }
#define threadIdx emu::tIdx
#define blockIdx emu::bIdx
#define blockDim emu::bDim
#undef __global__
#define __global__
#undef __grid_constant__
#define __grid_constant__
#undef __launch_bounds__
#define __launch_bounds__(...)

#include "../../tiny-differentiable-simulator_b200/csrc/tds_rigid.cu"

extern "C" {
// desc [n_bodies][6]; params: dt, g[3], friction, restitution, erp, iterations; state [n][n_bodies][13]; force [n][n_bodies][3] or null.
// jac (or null): [n][13 nb][16 nb] by the dual-number instance.
int tdsemu_rigid(const double* desc, int nb, const double* params, int n, const double* state, const double* force, int steps,
                 double* state_out, double* jac) {
  RigidWorld W;
  { const int rcw = tds_rigid_world_from_desc(desc, nb, &W); if (rcw) return rcw; }   // as tds_b200_rigid_create refuses
  W.dt = params[0]; for (int k = 0; k < 3; ++k) W.gravity[k] = params[1 + k];
  W.friction = params[4]; W.restitution = params[5]; W.erp = params[6]; W.num_solver_iterations = (int)params[7];
  const int ns = (n + 31) & ~31, rows = 13 * nb, cols = 16 * nb;
  std::vector<double> s((size_t)rows * ns, 0.0), o((size_t)rows * ns, 0.0), f((size_t)3 * nb * ns, 0.0), J;
  for (int e = 0; e < n; ++e) {
    for (int k = 0; k < rows; ++k) s[(size_t)k * ns + e] = state[(size_t)e * rows + k];
    if (force) for (int k = 0; k < 3 * nb; ++k) f[(size_t)k * ns + e] = force[(size_t)e * 3 * nb + k];
  }
  emu::bDim = {1, 1, 1};
  if (jac) J.assign((size_t)rows * cols * ns, 0.0);
  for (int e = 0; e < n; ++e) {
    emu::tIdx = {0, 0, 0};
    if (!jac) {
      emu::bIdx = {(unsigned)e, 0, 0};
      tdsrb::tds_rigid_step_kernel<double, double>(W, s.data(), o.data(), force ? f.data() : nullptr, steps, n, ns, nullptr, 0);
    } else {
      for (int d = 0; d < cols; ++d) {
        emu::bIdx = {(unsigned)e, (unsigned)d, 0};
        tdsrb::tds_rigid_step_kernel<tds::Dual<double>, double>(W, s.data(), o.data(), f.data(), steps, n, ns, J.data(), 0);
      }
    }
  }
  for (int e = 0; e < n; ++e) {
    if (state_out) for (int k = 0; k < rows; ++k) state_out[(size_t)e * rows + k] = o[(size_t)k * ns + e];
    if (jac) for (int k = 0; k < rows * cols; ++k) jac[(size_t)e * rows * cols + k] = J[(size_t)k * ns + e];
  }
  return 0;
}
}
