/* TEST INFRASTRUCTURE - NOT PRODUCT CODE.  C-ABI of the plain-C fp64 restatement of the
 * reference hot path (oracle/tds_oracle.c).  Built into oracle/libtds_oracle.so by
 * oracle/build_oracle.sh.  Parity of this restatement is pinned against the unmodified
 * reference (oracle/_ref/libtds_ref.so) and the golden vectors in tests/golden/. */
#ifndef TDS_ORACLE_H
#define TDS_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

#define TDSO_MAX_LINKS 64
#define TDSO_MAX_QD 64
#define TDSO_MAX_CONTACTS 48

#define TDSO_MODE_FD 0        /* forward_dynamics only */
#define TDSO_MODE_NOCONTACT 1 /* FD -> integrate_euler (cartpole_environment2.h:86-93) */
#define TDSO_MODE_FULL 2      /* FD -> integrate_euler_qdd -> World::step -> integrate_euler */

typedef struct {
  double dt;
  double gravity[3];
  double friction;    /* World::default_friction, src/world.hpp:68 */
  double restitution; /* World::default_restitution, src/world.hpp:69 */
  double erp;         /* MultiBodyConstraintSolver::erp_, src/mb_constraint_solver.hpp:64 */
  double cfm;         /* ::cfm_ :65 */
  int pgs_iterations; /* ::pgs_iterations_ :60 */
  int keep_all_points; /* ::keep_all_points_ :59 */
  /* Spring-damper contact law (contact_model = 1).  PARITY UNPINNED: MultiBodyConstraintSolverSpring is absent from the
   * reference snapshot (only its parameter names survive, python/pytinydiffsim.inl:825-856); this restates the law
   * specified in DESIGN.md "Spring-damper contacts", not a reference source. */
  int contact_model;     /* 0: LCP / PGS (the reference's solver), 1: spring-damper */
  double spring_k, damper_d, exponent_n, v_transition;
  int hard_contact_condition;
} TdsoParams;

int tdso_step(const double* model, const TdsoParams* P, int mode, const double* q, const double* qd,
              const double* tau, double* q_out, double* qd_out, double* qdd_out, int* n_contacts,
              int* contact_idx, double* contact_data, int contact_cap, double* link_xf_out);
int tdso_mass_matrix(const double* model, const double* q, double* M_out);
int tdso_locomotion_step(const double* model, const TdsoParams* P, const double* initial_poses, int n_act,
                         int base_dof, const double* input, double* output, int output_dim);
#ifdef __cplusplus
}
#endif
#endif
