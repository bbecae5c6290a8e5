// TEST INFRASTRUCTURE - NOT PRODUCT CODE.
//
// Thin C-ABI shim around the UNMODIFIED reference (erwincoumans/tiny-differentiable-simulator)
// headers, compiled *in place* from /root/reference/src by oracle/build_ref.sh into
// oracle/_ref/libtds_ref.so.  No reference source is copied into this repository; this file
// only #includes the reference headers and calls their public functions:
//   tds::forward_dynamics          src/dynamics/forward_dynamics.hpp:11
//   tds::integrate_euler_qdd       src/dynamics/integrator.hpp:141
//   tds::World::step               src/world.hpp:293
//   tds::integrate_euler           src/dynamics/integrator.hpp:10
//   tds::mass_matrix               src/dynamics/mass_matrix.hpp:13
//   tds::point_jacobian2           src/dynamics/jacobian.hpp:85
//   tds::UrdfCache::construct      src/urdf/urdf_cache.hpp:74
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load the resulting library (as the checker / the CPU baseline).
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "math/tiny/tiny_double_utils.h"
#include "math/tiny/tiny_float_utils.h"
#include "math/tiny/tiny_algebra.hpp"
#include "world.hpp"
#include "urdf/urdf_cache.hpp"
#include "dynamics/forward_dynamics.hpp"
#include "dynamics/integrator.hpp"
#include "dynamics/mass_matrix.hpp"
#include "dynamics/jacobian.hpp"

#include "tds_b200_model.h"

using namespace tds;

namespace {

template <typename A>
struct RefSim {
  using Scalar = typename A::Scalar;
  using Vector3 = typename A::Vector3;
  using Matrix3 = typename A::Matrix3;
  World<A> world;
  UrdfCache<A> cache;
  MultiBody<A>* plane_mb = nullptr;
  MultiBody<A>* mb = nullptr;
  double dt = 1e-3;

  static Matrix3 mat_from(const double* r) {
    return Matrix3(Scalar(r[0]), Scalar(r[1]), Scalar(r[2]), Scalar(r[3]), Scalar(r[4]),
                   Scalar(r[5]), Scalar(r[6]), Scalar(r[7]), Scalar(r[8]));
  }
  static void mat_to(const Matrix3& m, double* r) {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) r[i * 3 + j] = (double)m(i, j);
  }

  void add_plane(const double* normal) {
    plane_mb = world.create_multi_body("plane");
    Plane<A>* geom = world.create_plane();
    geom->set_normal(Vector3(Scalar(normal[0]), Scalar(normal[1]), Scalar(normal[2])));
    plane_mb->collision_geometries().push_back(geom);
    Transform<A> ident;
    ident.set_identity();
    plane_mb->collision_transforms().push_back(ident);
    plane_mb->initialize();
  }

  // Rebuild a reference MultiBody from the flat model (see include/tds_b200_model.h).
  bool build_from_flat(const double* m, int n) {
    if (n < TDSM_HEADER || (int)m[TDSM_H_MAGIC] != TDSM_MAGIC) return false;
    int n_links = (int)m[TDSM_H_NLINKS];
    int floating = (int)m[TDSM_H_FLOATING];
    int n_geoms = (int)m[TDSM_H_NGEOMS];
    int n_vis = (int)m[TDSM_H_NVIS];
    if ((int)m[TDSM_H_HASPLANE]) add_plane(m + TDSM_H_PLANE_N);
    mb = world.create_multi_body("robot");
    const double* b = m + TDSM_HEADER;
    mb->base_rbi() = RigidBodyInertia<A>(Scalar(b[0]), Vector3(Scalar(b[1]), Scalar(b[2]), Scalar(b[3])),
                                         mat_from(b + 4));
    const double* L = b + TDSM_BASE;
    const double* G = L + (size_t)n_links * TDSM_LINK;
    const double* V = G + (size_t)n_geoms * TDSM_GEOM;
    auto attach_geoms = [&](int link_index, std::vector<const Geometry<A>*>& geoms,
                            std::vector<Transform<A>>& xs) {
      for (int g = 0; g < n_geoms; ++g) {
        const double* gg = G + (size_t)g * TDSM_GEOM;
        if ((int)gg[TDSM_G_LINK] != link_index) continue;
        Transform<A> x;
        x.rotation = mat_from(gg + TDSM_G_R);
        x.translation = Vector3(Scalar(gg[TDSM_G_T]), Scalar(gg[TDSM_G_T + 1]), Scalar(gg[TDSM_G_T + 2]));
        Geometry<A>* geom = nullptr;
        switch ((int)gg[TDSM_G_TYPE]) {
          case TINY_SPHERE_TYPE: geom = world.create_sphere(Scalar(gg[TDSM_G_P])); break;
          case TINY_CAPSULE_TYPE: geom = world.create_capsule(Scalar(gg[TDSM_G_P]), Scalar(gg[TDSM_G_P + 1])); break;
          case TINY_BOX_TYPE:
            geom = world.create_box(Vector3(Scalar(gg[TDSM_G_P]), Scalar(gg[TDSM_G_P + 1]), Scalar(gg[TDSM_G_P + 2])));
            break;
          case TINY_PLANE_TYPE: {   // a plane shape on a link (urdf_to_multi_body.hpp:264-270)
            Plane<A>* pl = world.create_plane();
            pl->set_normal(Vector3(Scalar(gg[TDSM_G_P]), Scalar(gg[TDSM_G_P + 1]), Scalar(gg[TDSM_G_P + 2])));
            geom = pl;
            break;
          }
          default: continue;
        }
        geoms.push_back(geom);
        xs.push_back(x);
      }
    };
    attach_geoms(-1, mb->collision_geometries(), mb->collision_transforms());
    for (int i = 0; i < n_links; ++i) {
      const double* l = L + (size_t)i * TDSM_LINK;
      Link<A> link;
      Vector3 axis = Vector3(Scalar(l[TDSM_L_AXIS]), Scalar(l[TDSM_L_AXIS + 1]), Scalar(l[TDSM_L_AXIS + 2]));
      JointType jt = (JointType)(int)l[TDSM_L_JTYPE];
      if (jt == JOINT_REVOLUTE_AXIS || jt == JOINT_PRISMATIC_AXIS)
        link.set_joint_type(jt, axis);
      else
        link.set_joint_type(jt);
      link.X_T.rotation = mat_from(l + TDSM_L_XT_R);
      link.X_T.translation = Vector3(Scalar(l[TDSM_L_XT_T]), Scalar(l[TDSM_L_XT_T + 1]), Scalar(l[TDSM_L_XT_T + 2]));
      link.rbi = RigidBodyInertia<A>(Scalar(l[TDSM_L_MASS]),
                                     Vector3(Scalar(l[TDSM_L_COM]), Scalar(l[TDSM_L_COM + 1]), Scalar(l[TDSM_L_COM + 2])),
                                     mat_from(l + TDSM_L_INERTIA));
      link.stiffness = Scalar(l[TDSM_L_STIFFNESS]);
      link.damping = Scalar(l[TDSM_L_DAMPING]);
      attach_geoms(i, link.collision_geometries, link.X_collisions);
      for (int v = 0; v < n_vis; ++v) {
        const double* vv = V + (size_t)v * TDSM_VIS;
        if ((int)vv[TDSM_V_LINK] != i) continue;
        Transform<A> x;
        x.rotation = mat_from(vv + TDSM_V_R);
        x.translation = Vector3(Scalar(vv[TDSM_V_T]), Scalar(vv[TDSM_V_T + 1]), Scalar(vv[TDSM_V_T + 2]));
        link.X_visuals.push_back(x);
      }
      mb->attach(link, (int)l[TDSM_L_PARENT]);
    }
    mb->set_floating_base(floating != 0);
    mb->initialize();
    return true;
  }

  // Flatten the reference's own MultiBody (as produced by its URDF loader).
  int export_flat(double* out, int cap) const {
    int n_links = (int)mb->num_links();
    std::vector<double> geoms, vis;
    auto push_geoms = [&](int link_index, const std::vector<const Geometry<A>*>& gs,
                          const std::vector<Transform<A>>& xs) {
      for (size_t g = 0; g < gs.size(); ++g) {
        double rec[TDSM_GEOM] = {0};
        rec[TDSM_G_LINK] = link_index;
        rec[TDSM_G_TYPE] = gs[g]->get_type();
        switch (gs[g]->get_type()) {
          case TINY_SPHERE_TYPE: rec[TDSM_G_P] = (double)((const Sphere<A>*)gs[g])->get_radius(); break;
          case TINY_CAPSULE_TYPE:
            rec[TDSM_G_P] = (double)((const Capsule<A>*)gs[g])->get_radius();
            rec[TDSM_G_P + 1] = (double)((const Capsule<A>*)gs[g])->get_length();
            break;
          case TINY_BOX_TYPE: {
            auto e = ((const Box<A>*)gs[g])->get_extents();
            rec[TDSM_G_P] = (double)e[0]; rec[TDSM_G_P + 1] = (double)e[1]; rec[TDSM_G_P + 2] = (double)e[2];
            break;
          }
          case TINY_PLANE_TYPE: {
            auto nn = ((const Plane<A>*)gs[g])->get_normal();
            rec[TDSM_G_P] = (double)nn[0]; rec[TDSM_G_P + 1] = (double)nn[1]; rec[TDSM_G_P + 2] = (double)nn[2];
            break;
          }
          default: break;
        }
        mat_to(xs[g].rotation, rec + TDSM_G_R);
        for (int k = 0; k < 3; ++k) rec[TDSM_G_T + k] = (double)xs[g].translation[k];
        geoms.insert(geoms.end(), rec, rec + TDSM_GEOM);
      }
    };
    push_geoms(-1, mb->collision_geometries(), mb->collision_transforms());
    for (int i = 0; i < n_links; ++i) {
      const Link<A>& l = (*mb)[i];
      push_geoms(i, l.collision_geometries, l.X_collisions);
      for (size_t v = 0; v < l.X_visuals.size(); ++v) {
        double rec[TDSM_VIS] = {0};
        rec[TDSM_V_LINK] = i;
        mat_to(l.X_visuals[v].rotation, rec + TDSM_V_R);
        for (int k = 0; k < 3; ++k) rec[TDSM_V_T + k] = (double)l.X_visuals[v].translation[k];
        vis.insert(vis.end(), rec, rec + TDSM_VIS);
      }
    }
    int n_geoms = (int)(geoms.size() / TDSM_GEOM), n_vis = (int)(vis.size() / TDSM_VIS);
    int total = TDSM_HEADER + TDSM_BASE + n_links * TDSM_LINK + n_geoms * TDSM_GEOM + n_vis * TDSM_VIS;
    if (!out || cap < total) return total;
    std::memset(out, 0, sizeof(double) * total);
    out[TDSM_H_MAGIC] = TDSM_MAGIC;
    out[TDSM_H_NLINKS] = n_links;
    out[TDSM_H_FLOATING] = mb->is_floating() ? 1 : 0;
    out[TDSM_H_NQ] = mb->dof();
    out[TDSM_H_NQD] = mb->dof_qd();
    out[TDSM_H_NGEOMS] = n_geoms;
    out[TDSM_H_NVIS] = n_vis;
    out[TDSM_H_HASPLANE] = plane_mb ? 1 : 0;
    if (plane_mb) {
      const Plane<A>* p = (const Plane<A>*)plane_mb->collision_geometries()[0];
      for (int k = 0; k < 3; ++k) out[TDSM_H_PLANE_N + k] = (double)p->get_normal()[k];
      out[TDSM_H_PLANE_C] = (double)p->get_constant();
    }
    double* b = out + TDSM_HEADER;
    b[0] = (double)mb->base_rbi().mass;
    for (int k = 0; k < 3; ++k) b[1 + k] = (double)mb->base_rbi().com[k];
    mat_to(mb->base_rbi().inertia, b + 4);
    double* L = b + TDSM_BASE;
    for (int i = 0; i < n_links; ++i) {
      const Link<A>& l = (*mb)[i];
      double* r = L + (size_t)i * TDSM_LINK;
      r[TDSM_L_PARENT] = l.parent_index;
      r[TDSM_L_JTYPE] = (int)l.joint_type;
      r[TDSM_L_QIDX] = l.q_index;
      r[TDSM_L_QDIDX] = l.qd_index;
      bool rev = l.joint_type >= JOINT_REVOLUTE_X && l.joint_type <= JOINT_REVOLUTE_AXIS;
      for (int k = 0; k < 3; ++k) r[TDSM_L_AXIS + k] = (double)(rev ? l.S.top[k] : l.S.bottom[k]);
      mat_to(l.X_T.rotation, r + TDSM_L_XT_R);
      for (int k = 0; k < 3; ++k) r[TDSM_L_XT_T + k] = (double)l.X_T.translation[k];
      r[TDSM_L_MASS] = (double)l.rbi.mass;
      for (int k = 0; k < 3; ++k) r[TDSM_L_COM + k] = (double)l.rbi.com[k];
      mat_to(l.rbi.inertia, r + TDSM_L_INERTIA);
      r[TDSM_L_STIFFNESS] = (double)l.stiffness;
      r[TDSM_L_DAMPING] = (double)l.damping;
    }
    std::memcpy(L + (size_t)n_links * TDSM_LINK, geoms.data(), sizeof(double) * geoms.size());
    std::memcpy(L + (size_t)n_links * TDSM_LINK + geoms.size(), vis.data(), sizeof(double) * vis.size());
    return total;
  }

  void set_state(const double* q, const double* qd, const double* tau) {
    for (int i = 0; i < mb->dof(); ++i) mb->q(i) = Scalar(q[i]);
    for (int i = 0; i < mb->dof_qd(); ++i) mb->qd(i) = Scalar(qd ? qd[i] : 0.0);
    for (int i = 0; i < mb->dof_actuated(); ++i) mb->tau(i) = Scalar(tau ? tau[i] : 0.0);
  }

  // One step of the caller-defined pipeline (examples/environments/locomotion_contact_simulation.h:261-269
  // or cartpole_environment2.h:86-93) from raw (q, qd, tau).
  //   mode 0: forward_dynamics only (qdd out)
  //   mode 1: FD -> clear_forces -> integrate_euler                       (cartpole)
  //   mode 2: FD -> clear_forces -> integrate_euler_qdd -> world.step -> integrate_euler
  void step(int mode, const double* q, const double* qd, const double* tau, double* q_out,
            double* qd_out, double* qdd_out, double* qd_pre_contact, int* n_contacts,
            int* contact_idx, double* contact_data, int contact_cap) {
    set_state(q, qd, tau);
    forward_dynamics(*mb, world.get_gravity());
    if (qdd_out)
      for (int i = 0; i < mb->dof_qd(); ++i) qdd_out[i] = (double)mb->qdd(i);
    if (mode >= 1) {
      mb->clear_forces();
      if (mode == 2) {
        integrate_euler_qdd(*mb, Scalar(dt));
        if (qd_pre_contact)
          for (int i = 0; i < mb->dof_qd(); ++i) qd_pre_contact[i] = (double)mb->qd(i);
        world.step(Scalar(dt));
        int nc = 0;
        for (auto& list : world.mb_contacts_) {
          for (auto& cp : list) {
            if (nc < contact_cap) {
              if (contact_idx) {
                // (link_a, link_b): body indices are implied (plane = body A, robot = body B)
                contact_idx[nc * 2 + 0] = cp.link_a;
                contact_idx[nc * 2 + 1] = cp.link_b;
              }
              if (contact_data) {
                double* d = contact_data + (size_t)nc * 10;
                for (int k = 0; k < 3; ++k) {
                  d[k] = (double)cp.world_normal_on_b[k];
                  d[3 + k] = (double)cp.world_point_on_a[k];
                  d[6 + k] = (double)cp.world_point_on_b[k];
                }
                d[9] = (double)cp.distance;
              }
            }
            ++nc;
          }
        }
        if (n_contacts) *n_contacts = nc;
      }
      integrate_euler(*mb, Scalar(dt));
    }
    if (q_out)
      for (int i = 0; i < mb->dof(); ++i) q_out[i] = (double)mb->q(i);
    if (qd_out)
      for (int i = 0; i < mb->dof_qd(); ++i) qd_out[i] = (double)mb->qd(i);
  }

  void link_transforms(double* out) const {  // per link: R[9] row-major, t[3]
    for (size_t i = 0; i < mb->num_links(); ++i) {
      mat_to((*mb)[i].X_world.rotation, out + i * 12);
      for (int k = 0; k < 3; ++k) out[i * 12 + 9 + k] = (double)(*mb)[i].X_world.translation[k];
    }
  }

  void mass_matrix_of(const double* q, double* M_out) {
    set_state(q, nullptr, nullptr);
    int n = mb->dof_qd();
    typename A::MatrixX M(n, n);
    mass_matrix(*mb, &M);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) M_out[i * n + j] = (double)M(i, j);
  }

  void jacobian_of(const double* q, int link, const double* point, double* J_out) {
    set_state(q, nullptr, nullptr);
    forward_kinematics(*mb, mb->q());  // sets base_X_world for floating bodies (jacobian.hpp:39)
    int n = mb->dof_qd();
    auto J = point_jacobian2(*mb, link, Vector3(Scalar(point[0]), Scalar(point[1]), Scalar(point[2])), false);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < n; ++j) J_out[i * n + j] = (double)J(i, j);
  }
};

typedef TinyAlgebra<double, TINY::DoubleUtils> A64;
typedef TinyAlgebra<float, TINY::FloatUtils> A32;

struct Handle {
  int prec;  // 64 or 32
  RefSim<A64>* s64 = nullptr;
  RefSim<A32>* s32 = nullptr;
};

}  // namespace

#define DISPATCH(h, expr64, expr32) ((h)->prec == 64 ? (expr64) : (expr32))

extern "C" {

void* tdsref_create_from_urdf(const char* plane_urdf, const char* urdf, int floating, int prec) {
  Handle* h = new Handle;
  h->prec = prec == 32 ? 32 : 64;
  if (h->prec == 64) {
    auto* s = new RefSim<A64>;
    if (plane_urdf && plane_urdf[0]) s->plane_mb = s->cache.construct(plane_urdf, s->world, false, false);
    s->mb = s->cache.construct(urdf, s->world, false, floating != 0);
    h->s64 = s;
  } else {
    auto* s = new RefSim<A32>;
    if (plane_urdf && plane_urdf[0]) s->plane_mb = s->cache.construct(plane_urdf, s->world, false, false);
    s->mb = s->cache.construct(urdf, s->world, false, floating != 0);
    h->s32 = s;
  }
  return h;
}

void* tdsref_create_from_model(const double* model, int n, int prec) {
  Handle* h = new Handle;
  h->prec = prec == 32 ? 32 : 64;
  bool ok;
  if (h->prec == 64) {
    h->s64 = new RefSim<A64>;
    ok = h->s64->build_from_flat(model, n);
  } else {
    h->s32 = new RefSim<A32>;
    ok = h->s32->build_from_flat(model, n);
  }
  if (!ok) { delete h->s64; delete h->s32; delete h; return nullptr; }
  return h;
}

void tdsref_destroy(void* hv) {
  Handle* h = (Handle*)hv;
  if (!h) return;
  delete h->s64;
  delete h->s32;
  delete h;
}

int tdsref_export_model(void* hv, double* out, int cap) {
  Handle* h = (Handle*)hv;
  return DISPATCH(h, h->s64->export_flat(out, cap), h->s32->export_flat(out, cap));
}

int tdsref_dof_q(void* hv) { Handle* h = (Handle*)hv; return DISPATCH(h, h->s64->mb->dof(), h->s32->mb->dof()); }
int tdsref_dof_qd(void* hv) { Handle* h = (Handle*)hv; return DISPATCH(h, h->s64->mb->dof_qd(), h->s32->mb->dof_qd()); }
int tdsref_dof_tau(void* hv) { Handle* h = (Handle*)hv; return DISPATCH(h, h->s64->mb->dof_actuated(), h->s32->mb->dof_actuated()); }
int tdsref_num_links(void* hv) { Handle* h = (Handle*)hv; return (int)DISPATCH(h, h->s64->mb->num_links(), h->s32->mb->num_links()); }

void tdsref_set_params(void* hv, double dt, const double* gravity, double friction, double restitution,
                       int keep_all_points, int pgs_iterations, double erp, double cfm) {
  Handle* h = (Handle*)hv;
  if (h->prec == 64) {
    auto* s = h->s64;
    s->dt = dt;
    s->world.set_gravity(A64::Vector3(gravity[0], gravity[1], gravity[2]));
    s->world.default_friction = friction;
    s->world.default_restitution = restitution;
    auto* sol = s->world.get_mb_constraint_solver();
    sol->keep_all_points_ = keep_all_points != 0;
    sol->pgs_iterations_ = pgs_iterations;
    sol->erp_ = erp;
    sol->cfm_ = cfm;
  } else {
    auto* s = h->s32;
    s->dt = dt;
    s->world.set_gravity(A32::Vector3((float)gravity[0], (float)gravity[1], (float)gravity[2]));
    s->world.default_friction = (float)friction;
    s->world.default_restitution = (float)restitution;
    auto* sol = s->world.get_mb_constraint_solver();
    sol->keep_all_points_ = keep_all_points != 0;
    sol->pgs_iterations_ = pgs_iterations;
    sol->erp_ = (float)erp;
    sol->cfm_ = (float)cfm;
  }
}

void tdsref_step(void* hv, int mode, const double* q, const double* qd, const double* tau, double* q_out,
                 double* qd_out, double* qdd_out, double* qd_pre_contact, int* n_contacts, int* contact_idx,
                 double* contact_data, int contact_cap) {
  Handle* h = (Handle*)hv;
  if (h->prec == 64)
    h->s64->step(mode, q, qd, tau, q_out, qd_out, qdd_out, qd_pre_contact, n_contacts, contact_idx, contact_data, contact_cap);
  else
    h->s32->step(mode, q, qd, tau, q_out, qd_out, qdd_out, qd_pre_contact, n_contacts, contact_idx, contact_data, contact_cap);
}

// n steps back to back (timing loops of bench.py's reference arm: no per-step foreign-call overhead); fp64 instance.
// q [n][dof], qd [n][dof_qd], tau [n][dof_actuated] or NULL; outputs [n][...] or NULL.
void tdsref_step_batch(void* hv, int mode, int n, const double* q, const double* qd, const double* tau, double* q_out,
                       double* qd_out, double* qdd_out) {
  Handle* h = (Handle*)hv;
  auto* s = h->s64;
  if (!s) return;
  const int nq = s->mb->dof(), nqd = s->mb->dof_qd(), nt = s->mb->dof_actuated();
  for (int i = 0; i < n; ++i)
    s->step(mode, q + (size_t)i * nq, qd + (size_t)i * nqd, tau ? tau + (size_t)i * nt : nullptr,
            q_out ? q_out + (size_t)i * nq : nullptr, qd_out ? qd_out + (size_t)i * nqd : nullptr,
            qdd_out ? qdd_out + (size_t)i * nqd : nullptr, nullptr, nullptr, nullptr, nullptr, 0);
}

void tdsref_link_transforms(void* hv, double* out) {
  Handle* h = (Handle*)hv;
  if (h->prec == 64) h->s64->link_transforms(out); else h->s32->link_transforms(out);
}

void tdsref_mass_matrix(void* hv, const double* q, double* M_out) {
  Handle* h = (Handle*)hv;
  if (h->prec == 64) h->s64->mass_matrix_of(q, M_out); else h->s32->mass_matrix_of(q, M_out);
}

void tdsref_point_jacobian(void* hv, const double* q, int link, const double* point, double* J_out) {
  Handle* h = (Handle*)hv;
  if (h->prec == 64) h->s64->jacobian_of(q, link, point, J_out); else h->s32->jacobian_of(q, link, point, J_out);
}

}  // extern "C"
