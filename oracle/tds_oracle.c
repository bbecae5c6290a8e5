/* TEST INFRASTRUCTURE - NOT PRODUCT CODE.
 *
 * Plain-C (fp64) restatement of the reference's per-step rigid-body hot path
 * (erwincoumans/tiny-differentiable-simulator @ 8381b8c), driven by the flat model of
 * include/tds_b200_model.h.  Every function cites the reference file:line it follows.
 * The restatement deliberately keeps the reference's operation structure (dense 6x6
 * X^T Ia X products, CRBA that re-runs kinematics, full Cholesky inverse, one PGS sweep,
 * ...) so that it documents the reference algorithm; the CUDA product path is free to
 * restructure the arithmetic.
 *
 * PARITY PIN: this file is checked against the unmodified reference compiled in place
 * (oracle/_ref/libtds_ref.so, oracle/build_ref.sh) by tests/test_oracle.py, and against the
 * committed golden vectors in tests/golden/ that were generated from that library
 * (tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this file.
 * Scope: 1-DoF joints (prismatic/revolute X/Y/Z/axis), fixed joints, fixed or floating
 * base, plane-vs-{sphere,capsule} contacts, LCP/PGS solver.  Spherical joints are out of
 * scope.
 */
#include <math.h>
#include <string.h>

#include "tds_b200_model.h"
#include "tds_oracle.h"

#define MAXL TDSO_MAX_LINKS
#define MAXD TDSO_MAX_QD
#define MAXC TDSO_MAX_CONTACTS

typedef struct { double R[9]; double t[3]; } Xf;           /* tds::Transform, src/math/transform.hpp:13 */
typedef struct { double top[3]; double bot[3]; } Sv;        /* Motion/ForceVector, src/math/spatial_vector.hpp:94,165 */
typedef struct { double I[9], H[9], M[9]; } Abi;            /* ArticulatedBodyInertia, src/math/inertia.hpp:95 */

/* ---------- small helpers ---------- */
static void m3_mul(const double* A, const double* B, double* C) {
  double r[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
  memcpy(C, r, sizeof r);
}
static void m3_t(const double* A, double* T) {
  double r[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r[i * 3 + j] = A[j * 3 + i];
  memcpy(T, r, sizeof r);
}
static void m3_v(const double* A, const double* v, double* o) {
  double r[3];
  for (int i = 0; i < 3; ++i) r[i] = A[i * 3] * v[0] + A[i * 3 + 1] * v[1] + A[i * 3 + 2] * v[2];
  memcpy(o, r, sizeof r);
}
static void m3t_v(const double* A, const double* v, double* o) {
  double r[3];
  for (int i = 0; i < 3; ++i) r[i] = A[i] * v[0] + A[3 + i] * v[1] + A[6 + i] * v[2];
  memcpy(o, r, sizeof r);
}
static void cross3(const double* a, const double* b, double* o) {
  double r[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
  memcpy(o, r, sizeof r);
}
static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
/* TinyVectorCrossMatrix, src/math/tiny/tiny_matrix3x3.h:1015 */
static void cross_matrix(const double* v, double* m) {
  m[0] = 0; m[1] = -v[2]; m[2] = v[1];
  m[3] = v[2]; m[4] = 0; m[5] = -v[0];
  m[6] = -v[1]; m[7] = v[0]; m[8] = 0;
}
static void eye3(double* m) { memset(m, 0, 9 * sizeof(double)); m[0] = m[4] = m[8] = 1.0; }

/* TinyMatrix3x3::setRotation (right-associative build), src/math/tiny/tiny_matrix3x3.h:315-340 */
static void quat_to_matrix(double x, double y, double z, double w, double* m) {
  double d = x * x + y * y + z * z + w * w;
  if (d == 0.0) return;
  double s = 2.0 / d;
  double xs = x * s, ys = y * s, zs = z * s;
  double wx = w * xs, wy = w * ys, wz = w * zs;
  double xx = x * xs, xy = x * ys, xz = x * zs;
  double yy = y * ys, yz = y * zs, zz = z * zs;
  m[0] = 1.0 - (yy + zz); m[1] = xy - wz; m[2] = xz + wy;
  m[3] = xy + wz; m[4] = 1.0 - (xx + zz); m[5] = yz - wx;
  m[6] = xz - wy; m[7] = yz + wx; m[8] = 1.0 - (xx + yy);
}

/* TinyMatrix3x3::getRotation (non-CppAD branch), src/math/tiny/tiny_matrix3x3.h:434-466 */
static void matrix_to_quat(const double* m, double* q /* x y z w */) {
  double trace = m[0] + m[4] + m[8];
  double temp[4];
  if (trace < 0.0) {
    int i = m[0] < m[4] ? (m[4] < m[8] ? 2 : 1) : (m[0] < m[8] ? 2 : 0);
    int j = (i + 1) % 3, k = (i + 2) % 3;
    double tmp = ((m[i * 3 + i] - m[j * 3 + j]) - m[k * 3 + k]) + 1.0;
    double s = sqrt(tmp);
    temp[i] = s * 0.5;
    s = 0.5 / s;
    temp[3] = (m[j * 3 + k] - m[k * 3 + j]) * s;
    temp[j] = (m[i * 3 + j] + m[j * 3 + i]) * s;
    temp[k] = (m[i * 3 + k] + m[k * 3 + i]) * s;
  } else {
    double s = sqrt(trace + 1.0);
    temp[3] = s * 0.5;
    s = 0.5 / s;
    temp[0] = (m[1 * 3 + 2] - m[2 * 3 + 1]) * s;
    temp[1] = (m[2 * 3 + 0] - m[0 * 3 + 2]) * s;
    temp[2] = (m[0 * 3 + 1] - m[1 * 3 + 0]) * s;
  }
  q[0] = temp[0]; q[1] = temp[1]; q[2] = temp[2]; q[3] = -temp[3];
}

/* q * v * q^-1, TinyQuaternion::rotate, src/math/tiny/tiny_quaternion.h:170-176,306-345 */
static void quat_rotate(const double* q, const double* v, double* o) {
  double x = q[0], y = q[1], z = q[2], w = q[3];
  /* t = q * v (vector as pure quaternion) */
  double tx = w * v[0] + y * v[2] - z * v[1];
  double ty = w * v[1] + z * v[0] - x * v[2];
  double tz = w * v[2] + x * v[1] - y * v[0];
  double tw = -x * v[0] - y * v[1] - z * v[2];
  /* t *= conj(q) */
  double ix = -x, iy = -y, iz = -z, iw = w;
  o[0] = tw * ix + tx * iw + ty * iz - tz * iy;
  o[1] = tw * iy + ty * iw + tz * ix - tx * iz;
  o[2] = tw * iz + tz * iw + tx * iy - ty * ix;
}

/* Transform::operator* (right-associative), src/math/transform.hpp:123-131 */
static void xf_mul(const Xf* a, const Xf* b, Xf* o) {
  Xf r;
  double rt[3];
  m3_v(a->R, b->t, rt);
  for (int k = 0; k < 3; ++k) r.t[k] = a->t[k] + rt[k];
  m3_mul(a->R, b->R, r.R);
  *o = r;
}
/* Transform::apply(MotionVector): E = R^T, src/math/transform.hpp:210-226 */
static void xf_apply_motion(const Xf* x, const Sv* in, Sv* out) {
  double rxw[3], v_rxw[3];
  Sv r;
  cross3(x->t, in->top, rxw);
  for (int k = 0; k < 3; ++k) v_rxw[k] = in->bot[k] - rxw[k];
  m3t_v(x->R, in->top, r.top);
  m3t_v(x->R, v_rxw, r.bot);
  *out = r;
}
/* Transform::apply_inverse(MotionVector), src/math/transform.hpp:232-243 */
static void xf_apply_inverse_motion(const Xf* x, const Sv* in, Sv* out) {
  Sv r;
  double c[3];
  m3_v(x->R, in->top, r.top);
  m3_v(x->R, in->bot, r.bot);
  cross3(x->t, r.top, c);
  for (int k = 0; k < 3; ++k) r.bot[k] += c[k];
  *out = r;
}
/* Transform::apply(ForceVector): X^T F, src/math/transform.hpp:249-262 */
static void xf_apply_force(const Xf* x, const Sv* in, Sv* out) {
  Sv r;
  double c[3];
  m3_v(x->R, in->bot, r.bot);
  m3_v(x->R, in->top, r.top);
  cross3(x->t, r.bot, c);
  for (int k = 0; k < 3; ++k) r.top[k] += c[k];
  *out = r;
}
/* Transform::matrix(), src/math/transform.hpp:72-87 (6x6 row-major) */
static void xf_matrix(const Xf* x, double* m) {
  double E[9], rx[9], mErx[9];
  m3_t(x->R, E);
  cross_matrix(x->t, rx);
  m3_mul(E, rx, mErx);
  for (int k = 0; k < 9; ++k) mErx[k] = -mErx[k];
  memset(m, 0, 36 * sizeof(double));
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      m[i * 6 + j] = E[i * 3 + j];
      m[(i + 3) * 6 + j] = mErx[i * 3 + j];
      m[(i + 3) * 6 + j + 3] = E[i * 3 + j];
    }
}

/* ArticulatedBodyInertia = RigidBodyInertia, src/math/inertia.hpp:121-130 */
static void abi_from_rbi(double mass, const double* com, const double* inertia, Abi* a) {
  double H[9], Ht[9], HHt[9];
  cross_matrix(com, H);
  m3_t(H, Ht);
  m3_mul(H, Ht, HHt);
  for (int k = 0; k < 9; ++k) a->I[k] = inertia[k] + HHt[k] * mass;
  memset(a->M, 0, sizeof a->M);
  a->M[0] = a->M[4] = a->M[8] = mass;
  for (int k = 0; k < 9; ++k) a->H[k] = H[k] * mass;
}
/* ArticulatedBodyInertia::operator*(MotionVector), src/math/inertia.hpp:205-210 */
static void abi_mul(const Abi* a, const Sv* v, Sv* out) {
  Sv r;
  double t1[3], t2[3];
  m3_v(a->I, v->top, t1);
  m3_v(a->H, v->bot, t2);
  for (int k = 0; k < 3; ++k) r.top[k] = t1[k] + t2[k];
  m3_v(a->M, v->bot, t1);
  m3t_v(a->H, v->top, t2);
  for (int k = 0; k < 3; ++k) r.bot[k] = t1[k] + t2[k];
  *out = r;
}
/* ArticulatedBodyInertia::matrix(), src/math/inertia.hpp:152-160 */
static void abi_matrix(const Abi* a, double* m) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      m[i * 6 + j] = a->I[i * 3 + j];
      m[i * 6 + j + 3] = a->H[i * 3 + j];
      m[(i + 3) * 6 + j] = a->H[j * 3 + i];
      m[(i + 3) * 6 + j + 3] = a->M[i * 3 + j];
    }
}
/* X^T * Ia * X as a dense 6x6 product, then keep blocks (0,0),(0,3),(3,3):
 * src/dynamics/forward_dynamics.hpp:187-189, src/dynamics/mass_matrix.hpp:45-46,
 * src/math/inertia.hpp:138-143 (the lower-left block is dropped, not symmetrised). */
static void xt_abi_x(const Xf* x, const Abi* a, Abi* out) {
  double X[36], A[36], T[36], R[36];
  xf_matrix(x, X);
  abi_matrix(a, A);
  /* T = X^T * A */
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      double s = 0;
      for (int k = 0; k < 6; ++k) s += X[k * 6 + i] * A[k * 6 + j];
      T[i * 6 + j] = s;
    }
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      double s = 0;
      for (int k = 0; k < 6; ++k) s += T[i * 6 + k] * X[k * 6 + j];
      R[i * 6 + j] = s;
    }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      out->I[i * 3 + j] = R[i * 6 + j];
      out->H[i * 3 + j] = R[i * 6 + j + 3];
      out->M[i * 3 + j] = R[(i + 3) * 6 + j + 3];
    }
}
static void abi_add(Abi* a, const Abi* b) {
  for (int k = 0; k < 9; ++k) { a->I[k] += b->I[k]; a->H[k] += b->H[k]; a->M[k] += b->M[k]; }
}

/* 3x3 inverse (adjugate / determinant), TinyMatrix3x3::inverse, src/math/tiny/tiny_matrix3x3.h */
static void m3_inv(const double* m, double* o) {
  double co0 = m[4] * m[8] - m[5] * m[7];
  double co1 = m[5] * m[6] - m[3] * m[8];
  double co2 = m[3] * m[7] - m[4] * m[6];
  double det = m[0] * co0 + m[1] * co1 + m[2] * co2;
  double s = 1.0 / det;
  double r[9] = {co0 * s, (m[2] * m[7] - m[1] * m[8]) * s, (m[1] * m[5] - m[2] * m[4]) * s,
                 co1 * s, (m[0] * m[8] - m[2] * m[6]) * s, (m[2] * m[3] - m[0] * m[5]) * s,
                 co2 * s, (m[1] * m[6] - m[0] * m[7]) * s, (m[0] * m[4] - m[1] * m[3]) * s};
  memcpy(o, r, sizeof r);
}
/* ArticulatedBodyInertia::inv_mul with the reference's block inverse (note C = -H, not H^T),
 * src/math/inertia.hpp:302-328 */
static void abi_inv_mul(const Abi* a, const Sv* f, Sv* out) {
  double Ainv[9], C[9], CAinv[9], CAinvB[9], S[9], D[9], AinvB[9], AinvBD[9], T[9], Ii[9], Hi[9];
  m3_inv(a->I, Ainv);
  for (int k = 0; k < 9; ++k) C[k] = -a->H[k];
  m3_mul(C, Ainv, CAinv);
  m3_mul(CAinv, a->H, CAinvB);
  for (int k = 0; k < 9; ++k) S[k] = a->M[k] - CAinvB[k];
  m3_inv(S, D);
  m3_mul(Ainv, a->H, AinvB);
  m3_mul(AinvB, D, AinvBD);
  m3_mul(AinvBD, C, T);
  m3_mul(T, Ainv, T);
  for (int k = 0; k < 9; ++k) { Ii[k] = Ainv[k] + T[k]; Hi[k] = -AinvBD[k]; }
  double t1[3], t2[3];
  Sv r;
  m3_v(Ii, f->top, t1); m3_v(Hi, f->bot, t2);
  for (int k = 0; k < 3; ++k) r.top[k] = t1[k] + t2[k];
  m3_v(D, f->bot, t1); m3t_v(Hi, f->top, t2);
  for (int k = 0; k < 3; ++k) r.bot[k] = t1[k] + t2[k];
  *out = r;
}

/* ---------- model access ---------- */
typedef struct {
  const double* m;
  int n_links, floating, n_q, n_qd, n_geoms, n_vis, has_plane;
  const double *base, *links, *geoms, *vis;
} Model;

static int model_open(const double* m, Model* M) {
  if ((int)m[TDSM_H_MAGIC] != TDSM_MAGIC) return -1;
  M->m = m;
  M->n_links = (int)m[TDSM_H_NLINKS];
  M->floating = (int)m[TDSM_H_FLOATING];
  M->n_q = (int)m[TDSM_H_NQ];
  M->n_qd = (int)m[TDSM_H_NQD];
  M->n_geoms = (int)m[TDSM_H_NGEOMS];
  M->n_vis = (int)m[TDSM_H_NVIS];
  M->has_plane = (int)m[TDSM_H_HASPLANE];
  M->base = m + TDSM_HEADER;
  M->links = M->base + TDSM_BASE;
  M->geoms = M->links + (size_t)M->n_links * TDSM_LINK;
  M->vis = M->geoms + (size_t)M->n_geoms * TDSM_GEOM;
  if (M->n_links > MAXL || M->n_qd > MAXD) return -2;
  return 0;
}
#define LNK(M, i) ((M)->links + (size_t)(i)*TDSM_LINK)

/* per-step scratch = the mutable members of tds::Link / tds::MultiBody */
typedef struct {
  Xf X_parent[MAXL], X_world[MAXL], base_X_world;
  Sv S[MAXL], vJ[MAXL], v[MAXL], c[MAXL], a[MAXL], pA[MAXL], U[MAXL];
  Abi abi[MAXL], base_abi;
  double D[MAXL], u[MAXL];
  Sv base_velocity, base_acc, base_bias_force;
} State;

static void link_S(const double* l, Sv* S) {
  memset(S, 0, sizeof *S);
  int jt = (int)l[TDSM_L_JTYPE];
  switch (jt) { /* Link::set_joint_type, src/link.hpp:125-193 */
    case TDSJ_PRISMATIC_X: S->bot[0] = 1; break;
    case TDSJ_PRISMATIC_Y: S->bot[1] = 1; break;
    case TDSJ_PRISMATIC_Z: S->bot[2] = 1; break;
    case TDSJ_PRISMATIC_AXIS: memcpy(S->bot, l + TDSM_L_AXIS, 3 * sizeof(double)); break;
    case TDSJ_REVOLUTE_X: S->top[0] = 1; break;
    case TDSJ_REVOLUTE_Y: S->top[1] = 1; break;
    case TDSJ_REVOLUTE_Z: S->top[2] = 1; break;
    case TDSJ_REVOLUTE_AXIS: memcpy(S->top, l + TDSM_L_AXIS, 3 * sizeof(double)); break;
    default: break;
  }
}

/* Link::jcalc(q, &X_J, &X_parent), src/link.hpp:229-287 */
static void jcalc(const double* l, double q, Xf* X_parent) {
  Xf XJ, XT;
  eye3(XJ.R);
  XJ.t[0] = XJ.t[1] = XJ.t[2] = 0;
  int jt = (int)l[TDSM_L_JTYPE];
  double c = cos(q), s = sin(q);
  switch (jt) {
    case TDSJ_PRISMATIC_X: XJ.t[0] = q; break;
    case TDSJ_PRISMATIC_Y: XJ.t[1] = q; break;
    case TDSJ_PRISMATIC_Z: XJ.t[2] = q; break;
    case TDSJ_PRISMATIC_AXIS:
      for (int k = 0; k < 3; ++k) XJ.t[k] = l[TDSM_L_AXIS + k] * q;
      break;
    case TDSJ_REVOLUTE_X: /* tiny_matrix3x3.h:218-234 */
      XJ.R[4] = c; XJ.R[5] = -s; XJ.R[7] = s; XJ.R[8] = c; break;
    case TDSJ_REVOLUTE_Y:
      XJ.R[0] = c; XJ.R[2] = s; XJ.R[6] = -s; XJ.R[8] = c; break;
    case TDSJ_REVOLUTE_Z:
      XJ.R[0] = c; XJ.R[1] = -s; XJ.R[3] = s; XJ.R[4] = c; break;
    case TDSJ_REVOLUTE_AXIS: { /* TinyQuaternion::setRotation(axis, angle), tiny_quaternion.h:178-183 */
      const double* ax = l + TDSM_L_AXIS;
      double d = sqrt(dot3(ax, ax));
      double sh = sin(q * 0.5) / d;
      quat_to_matrix(ax[0] * sh, ax[1] * sh, ax[2] * sh, cos(q * 0.5), XJ.R);
      break;
    }
    default: break; /* fixed: identity */
  }
  memcpy(XT.R, l + TDSM_L_XT_R, 9 * sizeof(double));
  memcpy(XT.t, l + TDSM_L_XT_T, 3 * sizeof(double));
  xf_mul(&XT, &XJ, X_parent);
}

static void sv_cross_motion(const Sv* a, const Sv* b, Sv* o) { /* tiny_algebra.hpp:101-105 */
  Sv r;
  double t1[3], t2[3];
  cross3(a->top, b->top, r.top);
  cross3(a->top, b->bot, t1);
  cross3(a->bot, b->top, t2);
  for (int k = 0; k < 3; ++k) r.bot[k] = t1[k] + t2[k];
  *o = r;
}
static void sv_cross_force(const Sv* a, const Sv* b, Sv* o) { /* tiny_algebra.hpp:112-115 */
  Sv r;
  double t1[3], t2[3];
  cross3(a->top, b->top, t1);
  cross3(a->bot, b->bot, t2);
  for (int k = 0; k < 3; ++k) r.top[k] = t1[k] + t2[k];
  cross3(a->top, b->bot, r.bot);
  *o = r;
}
static double sv_dot(const Sv* a, const Sv* b) { return dot3(a->top, b->top) + dot3(a->bot, b->bot); }

/* forward_kinematics(mb, q, qd), src/dynamics/kinematics.hpp:18-148.
 * qd == NULL reproduces the "empty qd" call made by mass_matrix (mass_matrix.hpp:36). */
static void forward_kinematics(const Model* M, State* st, const double* q, const double* qd) {
  if (M->floating) { /* kinematics.hpp:35-62 */
    quat_to_matrix(q[0], q[1], q[2], q[3], st->base_X_world.R);
    st->base_X_world.t[0] = q[4]; st->base_X_world.t[1] = q[5]; st->base_X_world.t[2] = q[6];
    for (int k = 0; k < 3; ++k) {
      st->base_velocity.top[k] = qd ? qd[k] : 0.0;
      st->base_velocity.bot[k] = qd ? qd[3 + k] : 0.0;
    }
    abi_from_rbi(M->base[0], M->base + 1, M->base + 4, &st->base_abi);
    double Rt[9], RI[9], Iw[9], t[3], gyro[3];
    m3_t(st->base_X_world.R, Rt);
    m3_mul(st->base_X_world.R, M->base + 4, RI);
    m3_mul(RI, Rt, Iw);
    m3_v(Iw, st->base_velocity.top, t);
    cross3(st->base_velocity.top, t, gyro);
    for (int k = 0; k < 3; ++k) { st->base_bias_force.top[k] = gyro[k]; st->base_bias_force.bot[k] = 0.0; }
  } else {
    eye3(st->base_X_world.R);
    st->base_X_world.t[0] = st->base_X_world.t[1] = st->base_X_world.t[2] = 0.0; /* set_identity, locomotion_contact_simulation.h:131 */
  }
  for (int i = 0; i < M->n_links; ++i) {
    const double* l = LNK(M, i);
    int parent = (int)l[TDSM_L_PARENT];
    int jt = (int)l[TDSM_L_JTYPE];
    double qv = (jt == TDSJ_FIXED) ? 0.0 : q[(int)l[TDSM_L_QIDX]];
    double qdv = (jt == TDSJ_FIXED || !qd) ? 0.0 : qd[(int)l[TDSM_L_QDIDX]];
    link_S(l, &st->S[i]);
    jcalc(l, qv, &st->X_parent[i]);
    for (int k = 0; k < 3; ++k) { st->vJ[i].top[k] = st->S[i].top[k] * qdv; st->vJ[i].bot[k] = st->S[i].bot[k] * qdv; }
    if (parent >= 0 || M->floating) { /* kinematics.hpp:76-87 */
      const Xf* pxw = parent >= 0 ? &st->X_world[parent] : &st->base_X_world;
      const Sv* pv = parent >= 0 ? &st->v[parent] : &st->base_velocity;
      Sv xv;
      xf_mul(pxw, &st->X_parent[i], &st->X_world[i]);
      xf_apply_motion(&st->X_parent[i], pv, &xv);
      for (int k = 0; k < 3; ++k) { st->v[i].top[k] = xv.top[k] + st->vJ[i].top[k]; st->v[i].bot[k] = xv.bot[k] + st->vJ[i].bot[k]; }
    } else { /* kinematics.hpp:88-95 */
      xf_mul(&st->base_X_world, &st->X_parent[i], &st->X_world[i]);
      st->v[i] = st->vJ[i];
    }
    sv_cross_motion(&st->v[i], &st->vJ[i], &st->c[i]); /* kinematics.hpp:96-97 (cJ = 0) */
    abi_from_rbi(l[TDSM_L_MASS], l + TDSM_L_COM, l + TDSM_L_INERTIA, &st->abi[i]); /* :99 */
    Sv Iv;
    abi_mul(&st->abi[i], &st->v[i], &Iv);
    sv_cross_force(&st->v[i], &Iv, &st->pA[i]); /* :132 (f_ext = 0 after clear_forces) */
  }
}

/* forward_dynamics (ABA), src/dynamics/forward_dynamics.hpp:11-326 */
static void forward_dynamics(const Model* M, State* st, const double* q, const double* qd, const double* tau,
                             const double* gravity, double* qdd) {
  forward_kinematics(M, st, q, qd);
  for (int i = M->n_links - 1; i >= 0; --i) { /* forward_dynamics.hpp:50-216 */
    const double* l = LNK(M, i);
    int parent = (int)l[TDSM_L_PARENT];
    int jt = (int)l[TDSM_L_JTYPE];
    abi_mul(&st->abi[i], &st->S[i], &st->U[i]); /* :111 */
    st->D[i] = sv_dot(&st->S[i], &st->U[i]);    /* :115 */
    double tau_val = 0.0;
    if (jt != TDSJ_FIXED) {
      int qdi = (int)l[TDSM_L_QDIDX];
      tau_val = tau ? tau[qdi - (M->floating ? 6 : 0)] : 0.0; /* multi_body.hpp:557-570 */
      tau_val -= l[TDSM_L_STIFFNESS] * q[(int)l[TDSM_L_QIDX]];   /* :122 */
      tau_val -= l[TDSM_L_DAMPING] * qd[qdi];                     /* :123 */
    }
    st->u[i] = tau_val - sv_dot(&st->S[i], &st->pA[i]); /* :129 */
    double invD = (jt == TDSJ_FIXED) ? 0.0 : 1.0 / st->D[i]; /* :153 */
    Abi Ia = st->abi[i];
    Sv UinvD;
    for (int k = 0; k < 3; ++k) { UinvD.top[k] = st->U[i].top[k] * invD; UinvD.bot[k] = st->U[i].bot[k] * invD; }
    for (int r = 0; r < 3; ++r)
      for (int cc = 0; cc < 3; ++cc) { /* mul_transpose, inertia.hpp:333-348; Ia = abi - U (U/D)^T, :160-168 */
        Ia.I[r * 3 + cc] -= st->U[i].top[r] * UinvD.top[cc];
        Ia.H[r * 3 + cc] -= st->U[i].top[r] * UinvD.bot[cc];
        Ia.M[r * 3 + cc] -= st->U[i].bot[r] * UinvD.bot[cc];
      }
    Sv Ia_c, pa, dpA;
    abi_mul(&Ia, &st->c[i], &Ia_c); /* :171 */
    double uD = st->u[i] * invD;
    for (int k = 0; k < 3; ++k) { /* :173 */
      pa.top[k] = st->pA[i].top[k] + Ia_c.top[k] + st->U[i].top[k] * uD;
      pa.bot[k] = st->pA[i].bot[k] + Ia_c.bot[k] + st->U[i].bot[k] * uD;
    }
    xf_apply_force(&st->X_parent[i], &pa, &dpA); /* :181 */
    Abi dI;
    xt_abi_x(&st->X_parent[i], &Ia, &dI); /* :187-189 */
    if (parent >= 0) {
      for (int k = 0; k < 3; ++k) { st->pA[parent].top[k] += dpA.top[k]; st->pA[parent].bot[k] += dpA.bot[k]; }
      abi_add(&st->abi[parent], &dI);
    } else if (M->floating) {
      for (int k = 0; k < 3; ++k) { st->base_bias_force.top[k] += dpA.top[k]; st->base_bias_force.bot[k] += dpA.bot[k]; }
      abi_add(&st->base_abi, &dI);
    }
  }
  if (M->floating) { /* :218-233 */
    Sv r;
    abi_inv_mul(&st->base_abi, &st->base_bias_force, &r);
    for (int k = 0; k < 3; ++k) { st->base_acc.top[k] = -r.top[k]; st->base_acc.bot[k] = -r.bot[k]; }
  } else { /* :237-243 */
    for (int k = 0; k < 3; ++k) { st->base_acc.top[k] = 0.0; st->base_acc.bot[k] = -gravity[k]; }
  }
  for (int i = 0; i < M->n_links; ++i) { /* :245-302 */
    const double* l = LNK(M, i);
    int parent = (int)l[TDSM_L_PARENT];
    int jt = (int)l[TDSM_L_JTYPE];
    const Sv* ap = parent >= 0 ? &st->a[parent] : &st->base_acc;
    Sv xa;
    xf_apply_motion(&st->X_parent[i], ap, &xa);
    for (int k = 0; k < 3; ++k) { st->a[i].top[k] = xa.top[k] + st->c[i].top[k]; st->a[i].bot[k] = xa.bot[k] + st->c[i].bot[k]; }
    if (jt != TDSJ_FIXED) {
      double invD = 1.0 / st->D[i];
      double qddv = invD * (st->u[i] - sv_dot(&st->U[i], &st->a[i]));
      qdd[(int)l[TDSM_L_QDIDX]] = qddv;
      for (int k = 0; k < 3; ++k) { st->a[i].top[k] += st->S[i].top[k] * qddv; st->a[i].bot[k] += st->S[i].bot[k] * qddv; }
    }
  }
  if (M->floating) { /* :317-322 */
    for (int k = 0; k < 3; ++k) { st->base_acc.bot[k] += gravity[k]; }
    for (int k = 0; k < 3; ++k) { qdd[k] = st->base_acc.top[k]; qdd[3 + k] = st->base_acc.bot[k]; }
  }
}

/* mass_matrix (CRBA), src/dynamics/mass_matrix.hpp:13-127.  Mm is n_qd x n_qd row-major. */
static void mass_matrix(const Model* M, State* st, const double* q, double* Mm) {
  int n = M->n_qd;
  forward_kinematics(M, st, q, 0); /* :36 */
  memset(Mm, 0, sizeof(double) * n * n);
  for (int i = M->n_links - 1; i >= 0; --i) {
    const double* l = LNK(M, i);
    int parent = (int)l[TDSM_L_PARENT];
    Abi dI;
    xt_abi_x(&st->X_parent[i], &st->abi[i], &dI); /* :45-46 */
    if (parent >= 0) abi_add(&st->abi[parent], &dI);
    else if (M->floating) abi_add(&st->base_abi, &dI);
    if ((int)l[TDSM_L_JTYPE] == TDSJ_FIXED) continue;
    int qd_i = (int)l[TDSM_L_QDIDX];
    Sv Fi;
    abi_mul(&st->abi[i], &st->S[i], &Fi); /* :86 */
    Mm[qd_i * n + qd_i] = sv_dot(&st->S[i], &Fi);
    int j = i;
    while ((int)LNK(M, j)[TDSM_L_PARENT] != -1) { /* :90-105 */
      xf_apply_force(&st->X_parent[j], &Fi, &Fi);
      j = (int)LNK(M, j)[TDSM_L_PARENT];
      if ((int)LNK(M, j)[TDSM_L_JTYPE] == TDSJ_FIXED) continue;
      int qd_j = (int)LNK(M, j)[TDSM_L_QDIDX];
      double h = sv_dot(&Fi, &st->S[j]);
      Mm[qd_i * n + qd_j] = h;
      Mm[qd_j * n + qd_i] = h;
    }
    if (M->floating) { /* :107-111 */
      xf_apply_force(&st->X_parent[j], &Fi, &Fi);
      for (int k = 0; k < 3; ++k) {
        Mm[k * n + qd_i] = Fi.top[k]; Mm[(3 + k) * n + qd_i] = Fi.bot[k];
        Mm[qd_i * n + k] = Fi.top[k]; Mm[qd_i * n + 3 + k] = Fi.bot[k];
      }
    }
  }
  if (M->floating) { /* :114-120 */
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        Mm[r * n + c] = st->base_abi.I[r * 3 + c];
        Mm[r * n + 3 + c] = st->base_abi.H[r * 3 + c];
        Mm[(3 + r) * n + c] = st->base_abi.H[c * 3 + r];
        Mm[(3 + r) * n + 3 + c] = st->base_abi.M[r * 3 + c];
      }
  }
}

/* TinyMatrixXxX::inversed via Cholesky, src/math/tiny/tiny_matrix_x.h:240-344.
 * Returns 0 if not positive definite. */
static int symmetric_inverse(const double* A, double* a, int n) {
  double diag[MAXD];
  memcpy(a, A, sizeof(double) * n * n);
  for (int i = 0; i < n; i++) {
    for (int j = i; j < n; j++) {
      double sum = a[i * n + j];
      for (int k = i - 1; k >= 0; k--) sum -= a[i * n + k] * a[j * n + k];
      if (i == j) {
        if (sum <= 0.0) return 0;
        diag[i] = sqrt(sum);
      } else {
        a[j * n + i] = sum / diag[i];
      }
    }
  }
  for (int i = 0; i < n; i++) {
    a[i * n + i] = 1.0 / diag[i];
    for (int j = i + 1; j < n; j++) {
      double sum = 0.0;
      for (int k = i; k < j; k++) sum -= a[j * n + k] * a[k * n + i];
      a[j * n + i] = sum / diag[j];
    }
  }
  for (int i = 0; i < n; i++)
    for (int j = i + 1; j < n; j++) a[i * n + j] = 0.0;
  for (int i = 0; i < n; i++) {
    a[i * n + i] = a[i * n + i] * a[i * n + i];
    for (int k = i + 1; k < n; k++) a[i * n + i] += a[k * n + i] * a[k * n + i];
    for (int j = i + 1; j < n; j++)
      for (int k = j; k < n; k++) a[i * n + j] += a[k * n + i] * a[k * n + j];
  }
  for (int i = 0; i < n; i++)
    for (int j = 0; j < i; j++) a[i * n + j] = a[j * n + i];
  return 1;
}

/* point_jacobian (world point, is_local_point=false), src/dynamics/jacobian.hpp:13-83.
 * J is 3 x n_qd row-major.  Uses X_world from the last kinematics pass (same q). */
static void point_jacobian(const Model* M, const State* st, int link_index, const double* point, double* J) {
  int n = M->n_qd;
  memset(J, 0, sizeof(double) * 3 * n);
  if (M->floating) { /* :39-58 */
    double bp[3], cr[9];
    for (int k = 0; k < 3; ++k) bp[k] = point[k] - st->base_X_world.t[k];
    cross_matrix(bp, cr);
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) J[r * n + c] = cr[c * 3 + r];
    J[0 * n + 3] = 1.0; J[1 * n + 4] = 1.0; J[2 * n + 5] = 1.0;
  }
  int i = link_index;
  while (i >= 0) { /* :63-80 */
    const double* l = LNK(M, i);
    if ((int)l[TDSM_L_JTYPE] != TDSJ_FIXED) {
      Sv st_, xs;
      Xf ptf;
      xf_apply_inverse_motion(&st->X_world[i], &st->S[i], &st_);
      eye3(ptf.R);
      memcpy(ptf.t, point, 3 * sizeof(double));
      xf_apply_motion(&ptf, &st_, &xs);
      int c = (int)l[TDSM_L_QDIDX];
      for (int r = 0; r < 3; ++r) J[r * n + c] = xs.bot[r];
    }
    i = (int)l[TDSM_L_PARENT];
  }
}

/* MultiBodyConstraintSolver::plane_space, src/mb_constraint_solver.hpp:506-520 */
static void plane_space(const double* n, double* p, double* q) {
  double n_sqr = n[2] * n[2];
  int mz = n_sqr > 0.5;
  double a = n[1] * n[1] + (mz ? n_sqr : n[0] * n[0]);
  double k = sqrt(a);
  p[0] = mz ? 0.0 : -n[1] * k;
  p[1] = mz ? -n[2] * k : n[0] * k;
  p[2] = mz ? n[1] * k : n[1] * k;
  q[0] = mz ? a * k : -n[2] * p[1];
  q[1] = mz ? -n[0] * p[2] : n[2] * p[0];
  q[2] = mz ? n[0] * p[1] : a * k;
}

typedef struct { double normal[3], pa[3], pb[3], dist; int link_b, geom; } Contact;

/* contact_plane_sphere, src/contact_point.hpp:97-125 */
static void plane_sphere(const double* n, double cst, const double* pos, double radius, Contact* c) {
  double mn[3] = {-n[0], -n[1], -n[2]};
  double t = -(dot3(pos, mn) + cst);
  for (int k = 0; k < 3; ++k) {
    c->pa[k] = pos[k] + t * mn[k];
    c->pb[k] = pos[k] - radius * n[k];
    c->normal[k] = mn[k];
  }
  c->dist = t - radius;
}

/* World::compute_contacts_multi_body_internal for (plane = body A, robot = body B),
 * src/world.hpp:206-282; dispatch src/contact_point.hpp:445-506. */
static int compute_contacts(const Model* M, const State* st, Contact* out, int cap) {
  if (!M->has_plane) return 0;
  const double* pn = M->m + TDSM_H_PLANE_N;
  double pc = M->m[TDSM_H_PLANE_C];
  int nc = 0;
  for (int jj = -1; jj < M->n_links; ++jj) {
    const Xf* xw = jj >= 0 ? &st->X_world[jj] : &st->base_X_world;
    for (int g = 0; g < M->n_geoms; ++g) {
      const double* gg = M->geoms + (size_t)g * TDSM_GEOM;
      if ((int)gg[TDSM_G_LINK] != jj) continue;
      Xf loc, tr;
      memcpy(loc.R, gg + TDSM_G_R, 9 * sizeof(double));
      memcpy(loc.t, gg + TDSM_G_T, 3 * sizeof(double));
      xf_mul(xw, &loc, &tr);
      int type = (int)gg[TDSM_G_TYPE];
      if (type == TDSG_SPHERE) {
        if (nc < cap) { plane_sphere(pn, pc, tr.t, gg[TDSM_G_P], &out[nc]); out[nc].link_b = jj; out[nc].geom = g; }
        ++nc;
      } else if (type == TDSG_CAPSULE) { /* contact_plane_capsule, contact_point.hpp:128-161 */
        double orn[4], len;
        matrix_to_quat(tr.R, orn);
        len = sqrt(orn[0] * orn[0] + orn[1] * orn[1] + orn[2] * orn[2] + orn[3] * orn[3]);
        for (int k = 0; k < 4; ++k) orn[k] /= len; /* normalize, world.hpp:231 */
        for (int e = 0; e < 2; ++e) {
          double off[3] = {0, 0, (e == 0 ? 0.5 : -0.5) * gg[TDSM_G_P + 1]}, ro[3], pos[3];
          quat_rotate(orn, off, ro);
          for (int k = 0; k < 3; ++k) pos[k] = tr.t[k] + ro[k];
          if (nc < cap) { plane_sphere(pn, pc, pos, gg[TDSM_G_P], &out[nc]); out[nc].link_b = jj; out[nc].geom = g; }
          ++nc;
        }
      }
      else if (type == TDSG_BOX) { /* contact_plane_box, contact_point.hpp:164-198: a sphere at each of the 8 corners */
        double orn[4], len, r = 1e-2; /* max(1e-2, Box::radius); the URDF loader leaves the box radius at 0 */
        matrix_to_quat(tr.R, orn);
        len = sqrt(orn[0] * orn[0] + orn[1] * orn[1] + orn[2] * orn[2] + orn[3] * orn[3]);
        for (int k = 0; k < 4; ++k) orn[k] /= len;
        const double d[3] = {0.5 * gg[TDSM_G_P] - r, 0.5 * gg[TDSM_G_P + 1] - r, 0.5 * gg[TDSM_G_P + 2] - r};
        for (int c = 0; c < 8; ++c) { /* Box::get_corner_points order, geometry.hpp:244-260: x outermost, z innermost */
          double off[3] = {(c & 4) ? -d[0] : d[0], (c & 2) ? -d[1] : d[1], (c & 1) ? -d[2] : d[2]}, ro[3], pos[3];
          quat_rotate(orn, off, ro);
          for (int k = 0; k < 3; ++k) pos[k] = tr.t[k] + ro[k];
          if (nc < cap) { plane_sphere(pn, pc, pos, r, &out[nc]); out[nc].link_b = jj; out[nc].geom = g; }
          ++nc;
        }
      }
      /* meshes vs plane: not restated (no config uses them). */
    }
  }
  return nc;
}

/* MultiBodyConstraintSolver::resolve_collision(_internal) with body A = static plane (n_a = 0),
 * src/mb_constraint_solver.hpp:169-498; solve_pgs :101-142. */
static void resolve_collision(const Model* M, State* st, const TdsoParams* P, const Contact* all, int n_all,
                              const double* q, double* qd) {
  Contact cps[MAXC];
  int n_c = 0;
  for (int i = 0; i < n_all; ++i)
    if (P->keep_all_points || all[i].dist < 0.0) { if (n_c < MAXC) cps[n_c] = all[i]; ++n_c; } /* :169-180 */
  if (n_c == 0 || n_c > MAXC) return;
  int n = M->n_qd;
  if (n == 0) return;
  static double Mm[MAXD * MAXD], Minv[MAXD * MAXD];
  static double Jc[3 * MAXC * MAXD], JM[3 * MAXC * MAXD], A[9 * MAXC * MAXC];
  double b[3 * MAXC], x[3 * MAXC], lo[3 * MAXC], hi[3 * MAXC];
  int dep[3 * MAXC];
  mass_matrix(M, st, q, Mm);              /* :225-226 (mb_b) */
  if (!symmetric_inverse(Mm, Minv, n)) return; /* :230-231; reference asserts */
  int rows = 3 * n_c;
  memset(Jc, 0, sizeof(double) * rows * n);
  memset(b, 0, sizeof b);
  for (int i = 0; i < n_c; ++i) { /* :271-388 */
    const Contact* cp = &cps[i];
    double collision = cp->dist < 0.0 ? 1.0 : 0.0;
    double J[3 * MAXD];
    point_jacobian(M, st, cp->link_b, cp->pb, J); /* :286 */
    double nrm[3], f1[3], f2[3], vel_b[3] = {0, 0, 0}, rel[3];
    for (int k = 0; k < 3; ++k) nrm[k] = cp->normal[k] * collision;
    for (int c = 0; c < n; ++c) Jc[i * n + c] = J[c] * nrm[0] + J[n + c] * nrm[1] + J[2 * n + c] * nrm[2];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < n; ++c) vel_b[r] += J[r * n + c] * qd[c];
    for (int k = 0; k < 3; ++k) rel[k] = 0.0 - vel_b[k]; /* vel_a = 0 (plane has no dofs) :299-303 */
    double nrv = dot3(cp->normal, rel);
    double baum = P->erp * cp->dist / P->dt;
    b[i] = (-(1.0 + P->restitution) * nrv - baum) * collision; /* :306-310 */
    plane_space(cp->normal, f1, f2);                           /* :330-333 */
    for (int k = 0; k < 3; ++k) { f1[k] *= collision; f2[k] *= collision; }
    b[n_c + i] = -dot3(f1, rel);
    b[2 * n_c + i] = -dot3(f2, rel);
    for (int c = 0; c < n; ++c) {
      Jc[(n_c + i) * n + c] = J[c] * f1[0] + J[n + c] * f1[1] + J[2 * n + c] * f1[2];
      Jc[(2 * n_c + i) * n + c] = J[c] * f2[0] + J[n + c] * f2[1] + J[2 * n + c] * f2[2];
    }
  }
  if (P->contact_model == 1) {
    /* Spring-damper law (DESIGN.md "Spring-damper contacts"; PARITY UNPINNED, no reference source): per penetrating
     * point, Hunt-Crossley normal force f_n = k x^n + d x^n xdot (x = -distance, xdot = n_b . v_b = approach speed),
     * clamped at 0 with hard_contact_condition; friction mu f_n tanh(|v_t| / v_transition) against the tangential
     * velocity; applied as the impulse f dt through qd -= M^-1 Jc^T p, like the LCP impulses above. */
    double jtp[MAXD];
    memset(jtp, 0, sizeof jtp);
    for (int i = 0; i < n_c; ++i) {
      const Contact* cp = &cps[i];
      if (!(cp->dist < 0.0)) continue;
      double J[3 * MAXD], f1[3], f2[3], vel_b[3] = {0, 0, 0};
      point_jacobian(M, st, cp->link_b, cp->pb, J);
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < n; ++c) vel_b[r] += J[r * n + c] * qd[c];
      plane_space(cp->normal, f1, f2);
      const double x = -cp->dist, vn = dot3(cp->normal, vel_b), v1 = dot3(f1, vel_b), v2 = dot3(f2, vel_b);
      const double xn = pow(x, P->exponent_n);
      double fn = P->spring_k * xn + P->damper_d * xn * vn;
      if (P->hard_contact_condition && fn < 0.0) fn = 0.0;
      const double vt = sqrt(v1 * v1 + v2 * v2);
      const double sc = vt > 1e-12 ? P->friction * fn * tanh(vt / P->v_transition) / vt * P->dt : 0.0;
      const double p0 = fn * P->dt, p1 = sc * v1, p2 = sc * v2;
      for (int c = 0; c < n; ++c)
        jtp[c] += (J[c] * cp->normal[0] + J[n + c] * cp->normal[1] + J[2 * n + c] * cp->normal[2]) * p0 +
                  (J[c] * f1[0] + J[n + c] * f1[1] + J[2 * n + c] * f1[2]) * p1 +
                  (J[c] * f2[0] + J[n + c] * f2[1] + J[2 * n + c] * f2[2]) * p2;
    }
    for (int r = 0; r < n; ++r) {
      double s_ = 0;
      for (int c = 0; c < n; ++c) s_ += Minv[r * n + c] * jtp[c];
      qd[r] -= s_;
    }
    return;
  }
  /* lcp_A = jac_con * mass_matrix_inv * jac_con^T + cfm, :392-412 */
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < n; ++c) {
      double s = 0;
      for (int k = 0; k < n; ++k) s += Jc[r * n + k] * Minv[k * n + c];
      JM[r * n + c] = s;
    }
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < rows; ++c) {
      double s = 0;
      for (int k = 0; k < n; ++k) s += JM[r * n + k] * Jc[c * n + k];
      A[r * rows + c] = s;
    }
  for (int r = 0; r < rows; ++r) A[r * rows + r] += P->cfm;
  for (int i = 0; i < n_c; ++i) { /* :417-436 */
    x[i] = x[n_c + i] = x[2 * n_c + i] = 0.0;
    lo[i] = 0.0; hi[i] = 100000.0; dep[i] = -1;
    lo[n_c + i] = -P->friction; hi[n_c + i] = P->friction; dep[n_c + i] = i;
    lo[2 * n_c + i] = -P->friction; hi[2 * n_c + i] = P->friction; dep[2 * n_c + i] = i;
  }
  for (int it = 0; it < P->pgs_iterations; ++it) /* solve_pgs :101-142 */
    for (int i = 0; i < rows; ++i) {
      double delta = 0.0;
      for (int j = 0; j < i; ++j) delta += A[i * rows + j] * x[j];
      for (int j = i + 1; j < rows; ++j) delta += A[i * rows + j] * x[j];
      x[i] = (b[i] - delta) / A[i * rows + i];
      double s = 1.0;
      if (dep[i] >= 0) { s = x[dep[i]]; if (s < 0.0) s = 0.0; }
      if (x[i] < lo[i] * s) x[i] = lo[i] * s;
      if (x[i] > hi[i] * s) x[i] = hi[i] * s;
    }
  /* qd_b -= Minv * (Jn^T p_n) + Minv * (Jf1^T p_f1) + Minv * (Jf2^T p_f2), :476-497 */
  for (int blk = 0; blk < 3; ++blk) {
    double jtp[MAXD];
    for (int c = 0; c < n; ++c) {
      double s = 0;
      for (int i = 0; i < n_c; ++i) s += Jc[(blk * n_c + i) * n + c] * x[blk * n_c + i];
      jtp[c] = s;
    }
    for (int r = 0; r < n; ++r) {
      double s = 0;
      for (int c = 0; c < n; ++c) s += Minv[r * n + c] * jtp[c];
      qd[r] -= s;
    }
  }
}

/* integrate_euler with qdd == 0 (integrate_euler_qdd zeroed it), src/dynamics/integrator.hpp:10-133 */
static void integrate_q(const Model* M, State* st, double* q, const double* qd, double dt) {
  if (M->floating) {
    double qx = q[0], qy = q[1], qz = q[2], qw = q[3];
    const double* w = qd;
    double h = 0.5 * dt;
    /* quat_velocity, src/math/tiny/tiny_algebra.hpp:604-614 */
    double ww = (-qx * w[0] - qy * w[1] - qz * w[2]) * h;
    double xx = (qw * w[0] + qz * w[1] - qy * w[2]) * h;
    double yy = (qw * w[1] + qx * w[2] - qz * w[0]) * h;
    double zz = (qw * w[2] + qy * w[0] - qx * w[1]) * h;
    qx += xx; qy += yy; qz += zz; qw += ww;
    double len = sqrt(qx * qx + qy * qy + qz * qz + qw * qw);
    q[0] = qx / len; q[1] = qy / len; q[2] = qz / len; q[3] = qw / len;
    quat_to_matrix(q[0], q[1], q[2], q[3], st->base_X_world.R); /* integrator.hpp:67 (rotation only) */
    q[4] += qd[3] * dt; q[5] += qd[4] * dt; q[6] += qd[5] * dt;
  }
  for (int i = 0; i < M->n_links; ++i) {
    const double* l = LNK(M, i);
    if ((int)l[TDSM_L_JTYPE] == TDSJ_FIXED) continue;
    q[(int)l[TDSM_L_QIDX]] += qd[(int)l[TDSM_L_QDIDX]] * dt;
  }
}

static State g_state; /* single-threaded test oracle */

int tdso_step(const double* model, const TdsoParams* P, int mode, const double* q_in, const double* qd_in,
              const double* tau, double* q_out, double* qd_out, double* qdd_out, int* n_contacts,
              int* contact_idx, double* contact_data, int contact_cap, double* link_xf_out) {
  Model M;
  int rc = model_open(model, &M);
  if (rc) return rc;
  State* st = &g_state;
  double q[MAXD + 1], qd[MAXD], qdd[MAXD];
  memcpy(q, q_in, sizeof(double) * M.n_q);
  memcpy(qd, qd_in, sizeof(double) * M.n_qd);
  memset(qdd, 0, sizeof qdd);
  forward_dynamics(&M, st, q, qd, tau, P->gravity, qdd);
  if (qdd_out) memcpy(qdd_out, qdd, sizeof(double) * M.n_qd);
  if (link_xf_out)
    for (int i = 0; i < M.n_links; ++i) {
      memcpy(link_xf_out + i * 12, st->X_world[i].R, 9 * sizeof(double));
      memcpy(link_xf_out + i * 12 + 9, st->X_world[i].t, 3 * sizeof(double));
    }
  if (n_contacts) *n_contacts = 0;
  if (mode == TDSO_MODE_NOCONTACT) { /* cartpole_environment2.h:86-93: integrate_euler with live qdd */
    for (int i = 0; i < M.n_qd; ++i) qd[i] += qdd[i] * P->dt;
    integrate_q(&M, st, q, qd, P->dt);
  } else if (mode == TDSO_MODE_FULL) { /* locomotion_contact_simulation.h:261-269 */
    for (int i = 0; i < M.n_qd; ++i) qd[i] += qdd[i] * P->dt; /* integrate_euler_qdd */
    Contact cs[MAXC];
    int nc = compute_contacts(&M, st, cs, MAXC); /* X_world from the ABA kinematics pass */
    if (nc > MAXC) return -3;
    if (n_contacts) *n_contacts = nc;
    for (int i = 0; i < nc && i < contact_cap; ++i) {
      if (contact_idx) { contact_idx[i * 2] = -1; contact_idx[i * 2 + 1] = cs[i].link_b; }
      if (contact_data) {
        double* d = contact_data + (size_t)i * 10;
        memcpy(d, cs[i].normal, 3 * sizeof(double));
        memcpy(d + 3, cs[i].pa, 3 * sizeof(double));
        memcpy(d + 6, cs[i].pb, 3 * sizeof(double));
        d[9] = cs[i].dist;
      }
    }
    resolve_collision(&M, st, P, cs, nc, q, qd);
    integrate_q(&M, st, q, qd, P->dt);
  }
  if (q_out) memcpy(q_out, q, sizeof(double) * M.n_q);
  if (qd_out) memcpy(qd_out, qd, sizeof(double) * M.n_qd);
  return 0;
}

int tdso_mass_matrix(const double* model, const double* q, double* M_out) {
  Model M;
  int rc = model_open(model, &M);
  if (rc) return rc;
  mass_matrix(&M, &g_state, q, M_out);
  return 0;
}

/* LocomotionContactSimulation::step_forward_original, locomotion_contact_simulation.h:151-304.
 * input  = q | qd | action[n_act] | kp, kd, max_force ; output = q | qd | visuals (pos3,quat4) | up.z */
int tdso_locomotion_step(const double* model, const TdsoParams* P, const double* initial_poses, int n_act,
                         int base_dof, const double* input, double* output, int output_dim) {
  Model M;
  int rc = model_open(model, &M);
  if (rc) return rc;
  const double* q = input;
  const double* qd = input + M.n_q;
  const double* act = qd + M.n_qd;
  double kp = act[n_act], kd = act[n_act + 1], max_force = act[n_act + 2];
  double tau[MAXD];
  memset(tau, 0, sizeof tau);
  int pose_index = 0;
  int start_link = M.floating ? 0 : base_dof; /* :181 */
  for (int i = start_link; i < M.n_links; ++i) { /* :168-258 */
    const double* l = LNK(&M, i);
    if ((int)l[TDSM_L_JTYPE] == TDSJ_FIXED) continue;
    if (pose_index >= n_act) return -4;
    int qi = (int)l[TDSM_L_QIDX], qdi = (int)l[TDSM_L_QDIDX];
    int ti = M.floating ? qdi - 6 : qdi;
    double a = act[pose_index];
    if (a > 0.4) a = 0.4;
    if (a < -0.4) a = -0.4;
    double q_des = initial_poses[pose_index++] + a;
    double force = kp * (q_des - q[qi]) + kd * (0.0 - qd[qdi]);
    if (force < -max_force) force = -max_force;
    if (force > max_force) force = max_force;
    tau[ti] = force;
  }
  double xf[MAXL * 12];
  memset(output, 0, sizeof(double) * output_dim);
  rc = tdso_step(model, P, TDSO_MODE_FULL, q, qd, tau, output, output + M.n_q, 0, 0, 0, 0, 0, xf);
  if (rc) return rc;
  int j = M.n_q + M.n_qd;
  for (int v = 0; v < M.n_vis; ++v) { /* :279-298, X_world is the one of the step's kinematics pass */
    const double* vv = M.vis + (size_t)v * TDSM_VIS;
    int li = (int)vv[TDSM_V_LINK];
    Xf xw, loc, r;
    memcpy(xw.R, xf + li * 12, 9 * sizeof(double));
    memcpy(xw.t, xf + li * 12 + 9, 3 * sizeof(double));
    memcpy(loc.R, vv + TDSM_V_R, 9 * sizeof(double));
    memcpy(loc.t, vv + TDSM_V_T, 3 * sizeof(double));
    xf_mul(&xw, &loc, &r);
    double orn[4];
    matrix_to_quat(r.R, orn);
    if (j + 7 > output_dim) return -5;
    output[j++] = r.t[0]; output[j++] = r.t[1]; output[j++] = r.t[2];
    output[j++] = orn[0]; output[j++] = orn[1]; output[j++] = orn[2]; output[j++] = orn[3];
  }
  if (j < output_dim) output[j++] = g_state.base_X_world.R[8]; /* :301-303 up_dot_world_z */
  return 0;
}
