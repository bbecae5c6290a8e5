/* Flat ("compiled") model description shared by the host model compiler, the CUDA
 * simulator and the test oracles.
 *
 * A model is one contiguous array of doubles.  It is what the reference's URDF loader
 * (src/urdf/urdf_to_multi_body.hpp:41-220 -> MultiBody<Algebra>, src/multi_body.hpp:13,
 * Link<Algebra>, src/link.hpp:24) produces, flattened: per link the constant joint
 * transform X_T, the joint type/axis, the rigid-body inertia, plus collision / visual
 * shapes and an optional static ground plane (the "plane_implicit.urdf" body that the
 * reference's locomotion environments create first, locomotion_contact_simulation.h:108).
 *
 * Layout (all entries are doubles, integers are stored as exact doubles):
 *   header[TDSM_HEADER] | base[TDSM_BASE] | links[n_links][TDSM_LINK] |
 *   geoms[n_geoms][TDSM_GEOM] | visuals[n_vis][TDSM_VIS]
 * Matrices are 3x3 row-major.  Geoms are listed in the reference's contact enumeration
 * order (base geoms first, then link 0, 1, ...; src/world.hpp:212-281).
 */
#ifndef TDS_B200_MODEL_H
#define TDS_B200_MODEL_H

#define TDSM_MAGIC 20250200 /* layout version tag */

/* ---- header ---- */
#define TDSM_HEADER 16
#define TDSM_H_MAGIC 0
#define TDSM_H_NLINKS 1
#define TDSM_H_FLOATING 2
#define TDSM_H_NQ 3       /* MultiBody::dof()    (7 + joints when floating) */
#define TDSM_H_NQD 4      /* MultiBody::dof_qd() (6 + joints when floating) */
#define TDSM_H_NGEOMS 5
#define TDSM_H_NVIS 6
#define TDSM_H_HASPLANE 7 /* static ground plane = multibody 0 (body A of every contact) */
#define TDSM_H_PLANE_N 8  /* plane normal [3] (normalised, geometry.hpp:179) */
#define TDSM_H_PLANE_C 11 /* plane constant (always 0: urdf_to_multi_body.hpp:266-271) */
#define TDSM_H_NBODIES 12 /* 0 / 1: the links are ONE multibody.  K > 1: a world of K fixed-base multibodies, every root link
                             (parent -1) starts one, links and coordinates of a multibody contiguous; geoms of different
                             multibodies collide (src/world.hpp:206-282), geoms of one multibody never do */

/* ---- floating/fixed base rigid-body inertia (MultiBody::base_rbi_) ---- */
#define TDSM_BASE 13 /* mass, com[3], inertia[9] */

/* ---- per link ---- */
#define TDSM_LINK 34
#define TDSM_L_PARENT 0
#define TDSM_L_JTYPE 1 /* tds::JointType value, src/link.hpp:9-21 */
#define TDSM_L_QIDX 2
#define TDSM_L_QDIDX 3
#define TDSM_L_AXIS 4   /* [3] S.top (revolute) or S.bottom (prismatic) */
#define TDSM_L_XT_R 7   /* [9] X_T.rotation */
#define TDSM_L_XT_T 16  /* [3] X_T.translation */
#define TDSM_L_MASS 19
#define TDSM_L_COM 20     /* [3] */
#define TDSM_L_INERTIA 23 /* [9] */
#define TDSM_L_STIFFNESS 32
#define TDSM_L_DAMPING 33

/* ---- per collision geom ---- */
#define TDSM_GEOM 18
#define TDSM_G_LINK 0 /* -1 = base */
#define TDSM_G_TYPE 1 /* tds::GeometryTypes value, src/geometry.hpp:30-38 */
#define TDSM_G_P 2    /* [3] radius | radius,length | extents */
#define TDSM_G_R 5    /* [9] local rotation (X_collisions) */
#define TDSM_G_T 14   /* [3] local translation */

/* ---- per visual (link visuals only, in link order; locomotion_contact_simulation.h:281) ---- */
#define TDSM_VIS 13
#define TDSM_V_LINK 0
#define TDSM_V_R 1
#define TDSM_V_T 10

/* joint types (values of tds::JointType, src/link.hpp:9-21) */
#define TDSJ_FIXED (-1)
#define TDSJ_PRISMATIC_X 0
#define TDSJ_PRISMATIC_Y 1
#define TDSJ_PRISMATIC_Z 2
#define TDSJ_PRISMATIC_AXIS 3
#define TDSJ_REVOLUTE_X 4
#define TDSJ_REVOLUTE_Y 5
#define TDSJ_REVOLUTE_Z 6
#define TDSJ_REVOLUTE_AXIS 7
#define TDSJ_SPHERICAL 8

/* geometry types (values of tds::GeometryTypes, src/geometry.hpp:30-38) */
#define TDSG_SPHERE 0
#define TDSG_PLANE 1
#define TDSG_CAPSULE 2
#define TDSG_MESH 3
#define TDSG_BOX 4

#endif /* TDS_B200_MODEL_H */
