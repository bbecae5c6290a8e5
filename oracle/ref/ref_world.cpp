// TEST INFRASTRUCTURE - NOT PRODUCT CODE.
//
// A world of SEVERAL multibodies of the UNMODIFIED reference (compiled in place from /root/reference/src by
// oracle/build_ref.sh into oracle/_ref/libtds_ref.so), built from a flat model whose header says TDSM_H_NBODIES = K > 1
// (include/tds_b200_model.h): the ground plane first (multibody 0 of the world, as the reference's environments create it),
// then one fixed-base MultiBody per root link.  This file only #includes the reference headers and calls
//   tds::forward_dynamics      src/dynamics/forward_dynamics.hpp:11
//   tds::integrate_euler_qdd   src/dynamics/integrator.hpp:141
//   tds::World::step           src/world.hpp:293   (contacts between every pair of multibodies :206-282, solved list after list :351-355)
//   tds::integrate_euler       src/dynamics/integrator.hpp:10
// Used by tests/ and scripts/make_golden*.py only (checker of the kernels' multibody-vs-multibody contact stage).
#include <cstring>
#include <vector>

#include "math/tiny/tiny_double_utils.h"
#include "math/tiny/tiny_algebra.hpp"
#include "world.hpp"
#include "dynamics/forward_dynamics.hpp"
#include "dynamics/integrator.hpp"

#include "tds_b200_model.h"

using namespace tds;

namespace {
typedef TinyAlgebra<double, TINY::DoubleUtils> A;
typedef A::Vector3 Vector3;
typedef A::Matrix3 Matrix3;

Matrix3 mat_from(const double* r) { return Matrix3(r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8]); }

struct RefWorld {
  World<A> world;
  std::vector<MultiBody<A>*> bodies;
  std::vector<int> q_off, qd_off, link_off;   // first coordinate / velocity / link of a multibody in the flat model
  int n_q = 0, n_qd = 0;
  double dt = 1e-3;

  bool build(const double* m, int n) {
    if (n < TDSM_HEADER || (int)m[TDSM_H_MAGIC] != TDSM_MAGIC) return false;
    const int n_links = (int)m[TDSM_H_NLINKS], n_geoms = (int)m[TDSM_H_NGEOMS];
    if ((int)m[TDSM_H_FLOATING] || (int)m[TDSM_H_NBODIES] < 2) return false;
    n_q = (int)m[TDSM_H_NQ]; n_qd = (int)m[TDSM_H_NQD];
    if ((int)m[TDSM_H_HASPLANE]) {
      MultiBody<A>* plane = world.create_multi_body("plane");
      Plane<A>* geom = world.create_plane();
      geom->set_normal(Vector3(m[TDSM_H_PLANE_N], m[TDSM_H_PLANE_N + 1], m[TDSM_H_PLANE_N + 2]));
      plane->collision_geometries().push_back(geom);
      Transform<A> ident;
      ident.set_identity();
      plane->collision_transforms().push_back(ident);
      plane->initialize();
    }
    const double* L = m + TDSM_HEADER + TDSM_BASE;
    const double* G = L + (size_t)n_links * TDSM_LINK;
    MultiBody<A>* mb = nullptr;
    int first = 0;
    for (int i = 0; i < n_links; ++i) {
      const double* l = L + (size_t)i * TDSM_LINK;
      const int parent = (int)l[TDSM_L_PARENT];
      if (parent < 0) {
        if (mb) { mb->set_floating_base(false); mb->initialize(); }
        mb = world.create_multi_body("body");
        bodies.push_back(mb);
        first = i;
        link_off.push_back(i);
        q_off.push_back(-1); qd_off.push_back(-1);
      }
      Link<A> link;
      const Vector3 axis(l[TDSM_L_AXIS], l[TDSM_L_AXIS + 1], l[TDSM_L_AXIS + 2]);
      const JointType jt = (JointType)(int)l[TDSM_L_JTYPE];
      if (jt == JOINT_REVOLUTE_AXIS || jt == JOINT_PRISMATIC_AXIS) link.set_joint_type(jt, axis); else link.set_joint_type(jt);
      link.X_T.rotation = mat_from(l + TDSM_L_XT_R);
      link.X_T.translation = Vector3(l[TDSM_L_XT_T], l[TDSM_L_XT_T + 1], l[TDSM_L_XT_T + 2]);
      link.rbi = RigidBodyInertia<A>(l[TDSM_L_MASS], Vector3(l[TDSM_L_COM], l[TDSM_L_COM + 1], l[TDSM_L_COM + 2]), mat_from(l + TDSM_L_INERTIA));
      link.stiffness = l[TDSM_L_STIFFNESS];
      link.damping = l[TDSM_L_DAMPING];
      for (int g = 0; g < n_geoms; ++g) {
        const double* gg = G + (size_t)g * TDSM_GEOM;
        if ((int)gg[TDSM_G_LINK] != i) continue;
        Transform<A> x;
        x.rotation = mat_from(gg + TDSM_G_R);
        x.translation = Vector3(gg[TDSM_G_T], gg[TDSM_G_T + 1], gg[TDSM_G_T + 2]);
        Geometry<A>* geom = nullptr;
        switch ((int)gg[TDSM_G_TYPE]) {
          case TINY_SPHERE_TYPE: geom = world.create_sphere(gg[TDSM_G_P]); break;
          case TINY_CAPSULE_TYPE: geom = world.create_capsule(gg[TDSM_G_P], gg[TDSM_G_P + 1]); break;
          case TINY_BOX_TYPE: geom = world.create_box(Vector3(gg[TDSM_G_P], gg[TDSM_G_P + 1], gg[TDSM_G_P + 2])); break;
          case TINY_PLANE_TYPE: {
            Plane<A>* pl = world.create_plane();
            pl->set_normal(Vector3(gg[TDSM_G_P], gg[TDSM_G_P + 1], gg[TDSM_G_P + 2]));
            geom = pl;
            break;
          }
          default: continue;
        }
        link.collision_geometries.push_back(geom);
        link.X_collisions.push_back(x);
      }
      if (jt != JOINT_FIXED && q_off.back() < 0) { q_off.back() = (int)l[TDSM_L_QIDX]; qd_off.back() = (int)l[TDSM_L_QDIDX]; }
      mb->attach(link, parent < 0 ? -1 : parent - first);
    }
    if (mb) { mb->set_floating_base(false); mb->initialize(); }
    for (size_t k = 0; k < bodies.size(); ++k) if (q_off[k] < 0) { q_off[k] = 0; qd_off[k] = 0; }
    return (int)bodies.size() == (int)m[TDSM_H_NBODIES];
  }

  // mode 2: FD -> clear_forces -> integrate_euler_qdd (every multibody) -> World::step -> integrate_euler (every multibody);
  // mode 3: World::step alone.  contact_idx: (list, link_a, link_b) per contact of World::mb_contacts_, in order;
  // contact_data: normal on b [3], point on a [3], point on b [3], distance.
  void step(int mode, const double* q, const double* qd, const double* tau, double* q_out, double* qd_out, int* n_contacts,
            int* contact_idx, double* contact_data, int cap) {
    for (size_t k = 0; k < bodies.size(); ++k) {
      MultiBody<A>* mb = bodies[k];
      for (int i = 0; i < mb->dof(); ++i) mb->q(i) = q[q_off[k] + i];
      for (int i = 0; i < mb->dof_qd(); ++i) mb->qd(i) = qd[qd_off[k] + i];
      for (int i = 0; i < mb->dof_actuated(); ++i) mb->tau(i) = tau ? tau[qd_off[k] + i] : 0.0;
    }
    if (mode == 2)
      for (MultiBody<A>* mb : bodies) {
        forward_dynamics(*mb, world.get_gravity());
        mb->clear_forces();
        integrate_euler_qdd(*mb, dt);
      }
    else
      for (MultiBody<A>* mb : bodies) forward_kinematics(*mb, mb->q());
    world.step(dt);
    int nc = 0, list = 0;
    for (auto& lst : world.mb_contacts_) {
      for (auto& cp : lst) {
        if (nc < cap) {
          if (contact_idx) { contact_idx[nc * 3] = list; contact_idx[nc * 3 + 1] = cp.link_a; contact_idx[nc * 3 + 2] = cp.link_b; }
          if (contact_data) {
            double* d = contact_data + (size_t)nc * 10;
            for (int k = 0; k < 3; ++k) { d[k] = cp.world_normal_on_b[k]; d[3 + k] = cp.world_point_on_a[k]; d[6 + k] = cp.world_point_on_b[k]; }
            d[9] = cp.distance;
          }
        }
        ++nc;
      }
      ++list;
    }
    if (n_contacts) *n_contacts = nc;
    if (mode == 2)
      for (MultiBody<A>* mb : bodies) integrate_euler(*mb, dt);
    for (size_t k = 0; k < bodies.size(); ++k) {
      MultiBody<A>* mb = bodies[k];
      if (q_out) for (int i = 0; i < mb->dof(); ++i) q_out[q_off[k] + i] = mb->q(i);
      if (qd_out) for (int i = 0; i < mb->dof_qd(); ++i) qd_out[qd_off[k] + i] = mb->qd(i);
    }
  }
};
}  // namespace

extern "C" {
void* tdsrefw_create(const double* model, int n) {
  RefWorld* w = new RefWorld;
  if (!w->build(model, n)) { delete w; return nullptr; }
  return w;
}
void tdsrefw_destroy(void* h) { delete (RefWorld*)h; }
int tdsrefw_num_bodies(void* h) { return (int)((RefWorld*)h)->bodies.size(); }
void tdsrefw_set_params(void* h, double dt, const double* gravity, double friction, double restitution, int keep_all_points,
                        int pgs_iterations, double erp, double cfm) {
  RefWorld* w = (RefWorld*)h;
  w->dt = dt;
  w->world.set_gravity(Vector3(gravity[0], gravity[1], gravity[2]));
  w->world.default_friction = friction;
  w->world.default_restitution = restitution;
  auto* sol = w->world.get_mb_constraint_solver();
  sol->keep_all_points_ = keep_all_points != 0;
  sol->pgs_iterations_ = pgs_iterations;
  sol->erp_ = erp;
  sol->cfm_ = cfm;
}
void tdsrefw_step(void* h, int mode, const double* q, const double* qd, const double* tau, double* q_out, double* qd_out,
                  int* n_contacts, int* contact_idx, double* contact_data, int cap) {
  ((RefWorld*)h)->step(mode, q, qd, tau, q_out, qd_out, n_contacts, contact_idx, contact_data, cap);
}
}  // extern "C"
