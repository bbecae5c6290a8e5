// TEST INFRASTRUCTURE - NOT PRODUCT CODE.
//
// The RigidBody path of the UNMODIFIED reference (compiled in place from /root/reference/src by oracle/build_ref.sh into
// oracle/_ref/libtds_ref.so): a World of RigidBodys with one collision shape each, stepped by the reference's own
//   tds::World::step                               src/world.hpp:293-363
//   tds::RigidBody::apply_central_force            src/rigid_body.hpp:91
// This file only #includes the reference headers and calls them.  Checker of csrc/tds_rigid.cu (tests/, golden generation).
#include <vector>

#include "math/tiny/tiny_double_utils.h"
#include "math/tiny/tiny_algebra.hpp"
#include "world.hpp"

#include "tds_b200_model.h"

using namespace tds;

namespace {
typedef TinyAlgebra<double, TINY::DoubleUtils> A;
typedef A::Vector3 Vector3;

struct RefRigid {
  World<A> world;
  std::vector<RigidBody<A>*> bodies;
  std::vector<Geometry<A>*> owned;
  double dt = 1.0 / 60.0;
  ~RefRigid() { for (auto* g : owned) delete g; }
};
}  // namespace

extern "C" {
// desc [n_bodies][6]: mass, shape (TINY_*_TYPE), p0..p3 (sphere: radius; capsule: radius, length; box: extents; plane: normal, constant)
void* tdsrefr_create(const double* desc, int n_bodies) {
  RefRigid* w = new RefRigid;
  for (int i = 0; i < n_bodies; ++i) {
    const double* d = desc + i * 6;
    Geometry<A>* g = nullptr;
    switch ((int)d[1]) {
      case TINY_SPHERE_TYPE: g = w->world.create_sphere(d[2]); break;
      case TINY_CAPSULE_TYPE: g = w->world.create_capsule(d[2], d[3]); break;
      case TINY_BOX_TYPE: g = w->world.create_box(Vector3(d[2], d[3], d[4])); break;
      case TINY_PLANE_TYPE: g = new Plane<A>(Vector3(d[2], d[3], d[4]), d[5]); w->owned.push_back(g); break;
      default: delete w; return nullptr;
    }
    w->bodies.push_back(w->world.create_rigid_body(d[0], g));
  }
  return w;
}
void tdsrefr_destroy(void* h) { delete (RefRigid*)h; }
void tdsrefr_set_params(void* h, double dt, const double* gravity, double friction, double restitution, double erp, int iterations) {
  RefRigid* w = (RefRigid*)h;
  w->dt = dt;
  w->world.set_gravity(Vector3(gravity[0], gravity[1], gravity[2]));
  w->world.default_friction = friction;
  w->world.default_restitution = restitution;
  w->world.get_rb_constraint_solver()->erp_ = erp;
  w->world.num_solver_iterations = iterations;
}
// state [n_bodies][13]: position, orientation xyzw, linear velocity, angular velocity; force [n_bodies][3] or NULL (applied before
// the first step); `steps` calls of World::step.  n_contacts (or NULL): contacts of the LAST step.
void tdsrefr_step(void* h, const double* state, const double* force, int steps, double* state_out, int* n_contacts) {
  RefRigid* w = (RefRigid*)h;
  for (size_t b = 0; b < w->bodies.size(); ++b) {
    RigidBody<A>* rb = w->bodies[b];
    const double* s = state + b * 13;
    rb->world_pose_.position_ = Vector3(s[0], s[1], s[2]);
    rb->world_pose_.orientation_ = A::quat_from_xyzw(s[3], s[4], s[5], s[6]);
    rb->linear_velocity_ = Vector3(s[7], s[8], s[9]);
    rb->angular_velocity_ = Vector3(s[10], s[11], s[12]);
    rb->clear_forces();
    if (force) rb->apply_central_force(Vector3(force[b * 3], force[b * 3 + 1], force[b * 3 + 2]));
  }
  for (int i = 0; i < steps; ++i) w->world.step(w->dt);
  if (n_contacts) *n_contacts = (int)w->world.rb_contacts_.size();
  for (size_t b = 0; b < w->bodies.size(); ++b) {
    RigidBody<A>* rb = w->bodies[b];
    double* s = state_out + b * 13;
    for (int k = 0; k < 3; ++k) { s[k] = rb->world_pose_.position_[k]; s[7 + k] = rb->linear_velocity_[k]; s[10 + k] = rb->angular_velocity_[k]; }
    s[3] = rb->world_pose_.orientation_.x(); s[4] = rb->world_pose_.orientation_.y(); s[5] = rb->world_pose_.orientation_.z(); s[6] = rb->world_pose_.orientation_.w();
  }
}
}  // extern "C"
