// Batched rigid-body env-step kernel for sm_100a: one thread owns one environment, one warp owns a
// tile of 32 consecutive environments, per-environment scratch is staged in shared memory
// ([word][lane] interleaved -> bank-conflict free) or, for models that do not fit, in an
// L2-resident global arena with the same addressing.
//
// The kernel fuses the whole per-step hot path of the reference
//   PD torques            examples/environments/locomotion_contact_simulation.h:168-258
//   forward_kinematics    src/dynamics/kinematics.hpp:18-148
//   forward_dynamics(ABA) src/dynamics/forward_dynamics.hpp:11-326
//   integrate_euler_qdd   src/dynamics/integrator.hpp:141-195
//   contact detection     src/world.hpp:206-282, src/contact_point.hpp:97-161
//   mass_matrix (CRBA)    src/dynamics/mass_matrix.hpp:13-127
//   LCP assembly + PGS    src/mb_constraint_solver.hpp:101-142,191-498
//   integrate_euler       src/dynamics/integrator.hpp:10-133
// but is restructured for the GPU (this is not a translation):
//   * one kinematics pass instead of the reference's 1 + 2 + 2*contacts passes;
//   * child->parent articulated inertias are carried in registers along chains
//     (parent == i-1) and only branch points own a shared-memory accumulator;
//   * X^T Ia X in 3x3 block form with symmetric storage instead of dense 6x6 products;
//   * CRBA rides in the same leaf->root sweep as ABA pass 2, on rigid-body (10-float) composites;
//   * Cholesky factor L (no explicit inverse), Y = L^-1 Jc^T, matrix-free projected Gauss-Seidel on
//     w = Y p (A = Y^T Y + cfm is never formed), dqd = L^-T w;
//   * non-penetrating contact rows (masked to zero by the reference, :285-291) are skipped - their
//     impulse is exactly 0 in the reference's sweep as well;
//   * three scalar types: RA for the ABA, RC for kinematics / contact geometry, RS for CRBA + the
//     contact solve.  Default "mixed" = (fp32, fp64, fp32): erp/dt = 200 amplifies fp32 *position*
//     round-off beyond the 1e-5 parity budget, so world transforms, contact distances and the LCP
//     right-hand side are fp64; M is well conditioned (scaled cond ~10 for Laikago), fp32 suffices there.
#include <cuda_runtime.h>
#include <stdio.h>

#include "tds_math.cuh"
#include "tds_types.h"
#include "tds_model.h"

namespace tds {

// ---- per-environment scratch arena ------------------------------------------------------------
// word w (4 bytes) of environment column `col` lives at blk + (w*stride + col)*4; 8-byte elements
// start on even words.
struct Arena {
  char* blk;
  int stride;
  int col;
  template <typename T> TDS_D T& at(int word, int k) const {
    if (sizeof(T) == 4) return ((T*)blk)[(size_t)(word + k) * stride + col];
    return ((T*)blk)[(size_t)((word >> 1) + k) * stride + col];
  }
};

template <typename T> TDS_D void st_v3(const Arena& A, int w, int k, const V3<T>& v) {
  A.at<T>(w, k) = v.x; A.at<T>(w, k + 1) = v.y; A.at<T>(w, k + 2) = v.z;
}
template <typename T> TDS_D V3<T> ld_v3(const Arena& A, int w, int k) {
  return v3<T>(A.at<T>(w, k), A.at<T>(w, k + 1), A.at<T>(w, k + 2));
}
template <typename T> TDS_D void st_m3(const Arena& A, int w, int k, const M3<T>& m) {
  A.at<T>(w, k) = m.xx; A.at<T>(w, k + 1) = m.xy; A.at<T>(w, k + 2) = m.xz;
  A.at<T>(w, k + 3) = m.yx; A.at<T>(w, k + 4) = m.yy; A.at<T>(w, k + 5) = m.yz;
  A.at<T>(w, k + 6) = m.zx; A.at<T>(w, k + 7) = m.zy; A.at<T>(w, k + 8) = m.zz;
}
template <typename T> TDS_D M3<T> ld_m3(const Arena& A, int w, int k) {
  M3<T> m;
  m.xx = A.at<T>(w, k); m.xy = A.at<T>(w, k + 1); m.xz = A.at<T>(w, k + 2);
  m.yx = A.at<T>(w, k + 3); m.yy = A.at<T>(w, k + 4); m.yz = A.at<T>(w, k + 5);
  m.zx = A.at<T>(w, k + 6); m.zy = A.at<T>(w, k + 7); m.zz = A.at<T>(w, k + 8);
  return m;
}
template <typename T> TDS_D void st_xf(const Arena& A, int w, const Xf<T>& X) { st_m3(A, w, 0, X.R); st_v3(A, w, 9, X.t); }
template <typename T> TDS_D Xf<T> ld_xf(const Arena& A, int w) { Xf<T> X; X.R = ld_m3<T>(A, w, 0); X.t = ld_v3<T>(A, w, 9); return X; }
template <typename T> TDS_D void st_sv(const Arena& A, int w, int k, const Sv<T>& s) { st_v3(A, w, k, s.top); st_v3(A, w, k + 3, s.bot); }
template <typename T> TDS_D Sv<T> ld_sv(const Arena& A, int w, int k) { Sv<T> s; s.top = ld_v3<T>(A, w, k); s.bot = ld_v3<T>(A, w, k + 3); return s; }

// per-link region (element offsets, in units of RA)
enum { LK_XP = 0, LK_VC = 12, LK_U = 18, LK_INVD = 24, LK_u = 25, LK_SIZE = 26 };
// accumulator slot: abi 21 + pA 6 (units of RA), then Ic 10 (units of RC) at word offset M.acc_ic_word;
// the slot stride in words is M.acc_words (both computed on the host, tds_build_layout).
enum { AC_ABI = 0, AC_PA = 21, AC_NRA = 27, AC_NIC = 10 };
// contact record, RC part: pb 3, dist, link (the RS part b[3], x[3] lives at M.w_conS)
enum { CN_PB = 0, CN_DIST = 3, CN_LINK = 4, CN_SIZE = 5 };

template <typename T> TDS_D void acc_add_abi(const Arena& A, int w, const Abi<T>& a, const Sv<T>& p) {
  T* dummy = nullptr; (void)dummy;
  A.at<T>(w, 0) += a.I.xx; A.at<T>(w, 1) += a.I.xy; A.at<T>(w, 2) += a.I.xz; A.at<T>(w, 3) += a.I.yy; A.at<T>(w, 4) += a.I.yz; A.at<T>(w, 5) += a.I.zz;
  A.at<T>(w, 6) += a.H.xx; A.at<T>(w, 7) += a.H.xy; A.at<T>(w, 8) += a.H.xz; A.at<T>(w, 9) += a.H.yx; A.at<T>(w, 10) += a.H.yy; A.at<T>(w, 11) += a.H.yz;
  A.at<T>(w, 12) += a.H.zx; A.at<T>(w, 13) += a.H.zy; A.at<T>(w, 14) += a.H.zz;
  A.at<T>(w, 15) += a.M.xx; A.at<T>(w, 16) += a.M.xy; A.at<T>(w, 17) += a.M.xz; A.at<T>(w, 18) += a.M.yy; A.at<T>(w, 19) += a.M.yz; A.at<T>(w, 20) += a.M.zz;
  A.at<T>(w, AC_PA + 0) += p.top.x; A.at<T>(w, AC_PA + 1) += p.top.y; A.at<T>(w, AC_PA + 2) += p.top.z;
  A.at<T>(w, AC_PA + 3) += p.bot.x; A.at<T>(w, AC_PA + 4) += p.bot.y; A.at<T>(w, AC_PA + 5) += p.bot.z;
}
template <typename T> TDS_D void acc_add_rbi(const Arena& A, int w, const Rbi<T>& r) {
  A.at<T>(w, 0) += r.m; A.at<T>(w, 1) += r.h.x; A.at<T>(w, 2) += r.h.y; A.at<T>(w, 3) += r.h.z;
  A.at<T>(w, 4) += r.I.xx; A.at<T>(w, 5) += r.I.xy; A.at<T>(w, 6) += r.I.xz;
  A.at<T>(w, 7) += r.I.yy; A.at<T>(w, 8) += r.I.yz; A.at<T>(w, 9) += r.I.zz;
}
template <typename T> TDS_D Rbi<T> acc_load_rbi(const Arena& A, int w) {
  Rbi<T> r;
  r.m = A.at<T>(w, 0); r.h = ld_v3<T>(A, w, 1);
  r.I.xx = A.at<T>(w, 4); r.I.xy = A.at<T>(w, 5); r.I.xz = A.at<T>(w, 6);
  r.I.yy = A.at<T>(w, 7); r.I.yz = A.at<T>(w, 8); r.I.zz = A.at<T>(w, 9);
  return r;
}
template <typename T> TDS_D void acc_load(const Arena& A, int w, Abi<T>& a, Sv<T>& p) {
  a.I.xx = A.at<T>(w, 0); a.I.xy = A.at<T>(w, 1); a.I.xz = A.at<T>(w, 2); a.I.yy = A.at<T>(w, 3); a.I.yz = A.at<T>(w, 4); a.I.zz = A.at<T>(w, 5);
  a.H.xx = A.at<T>(w, 6); a.H.xy = A.at<T>(w, 7); a.H.xz = A.at<T>(w, 8); a.H.yx = A.at<T>(w, 9); a.H.yy = A.at<T>(w, 10); a.H.yz = A.at<T>(w, 11);
  a.H.zx = A.at<T>(w, 12); a.H.zy = A.at<T>(w, 13); a.H.zz = A.at<T>(w, 14);
  a.M.xx = A.at<T>(w, 15); a.M.xy = A.at<T>(w, 16); a.M.xz = A.at<T>(w, 17); a.M.yy = A.at<T>(w, 18); a.M.yz = A.at<T>(w, 19); a.M.zz = A.at<T>(w, 20);
  p = ld_sv<T>(A, w, AC_PA);
}

template <typename T> TDS_D Rbi<T> model_rbi(const double* r) {
  Rbi<T> o;
  o.m = T(r[0]); o.h = v3<T>(T(r[1]), T(r[2]), T(r[3]));
  o.I.xx = T(r[4]); o.I.xy = T(r[5]); o.I.xz = T(r[6]); o.I.yy = T(r[7]); o.I.yz = T(r[8]); o.I.zz = T(r[9]);
  return o;
}

// Link::jcalc, src/link.hpp:229-287: X_parent = X_T * X_J(q)
template <typename T> TDS_D Xf<T> jcalc(const DevModel& M, int i, T q) {
  Xf<T> XT;
  const double* xt = M.XT[i];
  XT.R.xx = T(xt[0]); XT.R.xy = T(xt[1]); XT.R.xz = T(xt[2]); XT.R.yx = T(xt[3]); XT.R.yy = T(xt[4]); XT.R.yz = T(xt[5]);
  XT.R.zx = T(xt[6]); XT.R.zy = T(xt[7]); XT.R.zz = T(xt[8]);
  XT.t = v3<T>(T(xt[9]), T(xt[10]), T(xt[11]));
  const int jt = M.jtype[i];
  if (jt == TDSJ_FIXED) return XT;
  Xf<T> XJ;
  XJ.R = m3_identity<T>();
  XJ.t = v3<T>(T(0), T(0), T(0));
  if (jt <= TDSJ_PRISMATIC_AXIS) {
    XJ.t = v3<T>(T(M.axis[i][0]) * q, T(M.axis[i][1]) * q, T(M.axis[i][2]) * q);
    Xf<T> r; r.R = XT.R; r.t = XT.t + mul(XT.R, XJ.t);
    return r;
  }
  if (jt == TDSJ_REVOLUTE_AXIS) {
    // TinyQuaternion::setRotation(axis, angle), src/math/tiny/tiny_quaternion.h:178-183
    T ax = T(M.axis[i][0]), ay = T(M.axis[i][1]), az = T(M.axis[i][2]);
    T d = sqrt_t(ax * ax + ay * ay + az * az);
    T s, c;
    sincos_t(q * T(0.5), &s, &c);
    s = s / d;
    XJ.R = quat_to_matrix<T>(ax * s, ay * s, az * s, c);
  } else {
    T s, c;
    sincos_t(q, &s, &c);
    if (jt == TDSJ_REVOLUTE_X) { XJ.R.yy = c; XJ.R.yz = -s; XJ.R.zy = s; XJ.R.zz = c; }
    else if (jt == TDSJ_REVOLUTE_Y) { XJ.R.xx = c; XJ.R.xz = s; XJ.R.zx = -s; XJ.R.zz = c; }
    else { XJ.R.xx = c; XJ.R.xy = -s; XJ.R.yx = s; XJ.R.yy = c; }
  }
  Xf<T> r; r.R = mul(XT.R, XJ.R); r.t = XT.t;
  return r;
}

template <typename T> TDS_D Sv<T> link_S(const DevModel& M, int i) {
  Sv<T> S;
  V3<T> ax = v3<T>(T(M.axis[i][0]), T(M.axis[i][1]), T(M.axis[i][2]));
  V3<T> z = v3<T>(T(0), T(0), T(0));
  const int fl = M.flags[i];
  S.top = (fl & TDS_LF_REVOLUTE) ? ax : z;
  S.bot = (fl & TDS_LF_PRISMATIC) ? ax : z;
  return S;
}

TDS_D int tri(int r, int c) { return r * (r + 1) / 2 + c; }  // lower triangle, r >= c

enum StepMode { MODE_FD = 0, MODE_NOCONTACT = 1, MODE_FULL = 2 };

// ---- strided small-vector helpers for the contact solve (element k of a vector lives at p[k*s]) ----
// Dot products use four independent accumulators so that consecutive shared-memory loads and FMAs
// overlap (a single dependent chain costs one LDS + FMA latency per element on a lone warp).
template <typename T> TDS_D T sdot(const T* a, const T* b, int s, int len) {
  T s0 = T(0), s1 = T(0), s2 = T(0), s3 = T(0);
  int k = 0;
  for (; k + 3 < len; k += 4) {
    s0 += a[k * s] * b[k * s];
    s1 += a[(k + 1) * s] * b[(k + 1) * s];
    s2 += a[(k + 2) * s] * b[(k + 2) * s];
    s3 += a[(k + 3) * s] * b[(k + 3) * s];
  }
  for (; k < len; ++k) s0 += a[k * s] * b[k * s];
  return (s0 + s1) + (s2 + s3);
}

template <typename RA, typename RC, typename RS, bool SMEM>
__global__ void __launch_bounds__(128, 1)
tds_step_kernel(const __grid_constant__ DevModel M, const __grid_constant__ SimParams P,
                const __grid_constant__ EnvParams E, const StepIO io, const int mode, const int use_pd,
                char* __restrict__ gscratch) {
  extern __shared__ __align__(16) char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp_in_blk = threadIdx.x >> 5;
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = env < io.n;
  const int e = live ? env : io.n - 1;  // dead lanes shadow the last env (no stores)
  Arena A;
  if (SMEM) {
    A.blk = smem_raw + (size_t)warp_in_blk * M.w_total * 32 * 4;
    A.stride = 32;
    A.col = lane;
  } else {
    A.blk = gscratch + (size_t)(env >> 5) * M.w_total * 32 * 4;   // per-warp block, same addressing as shared memory
    A.stride = 32;
    A.col = lane;
  }
  const int ST = A.stride;
  const int ns = io.n_stride;
  const int n_links = M.n_links;
  int phase_id = 0;
#define TDS_PHASE() do { if (io.phase_clk && lane == 0) io.phase_clk[(size_t)(env >> 5) * 16 + (phase_id++)] = clock64(); } while (0)
  TDS_PHASE();
  const int n = M.n_qd;
  const RA dtA = RA(P.dt);
  constexpr int RAW = (int)(sizeof(RA) / 4), RCW = (int)(sizeof(RC) / 4), RSW = (int)(sizeof(RS) / 4);
  float* const qv = &A.at<float>(M.w_q, 0);     // q, qd, tau live as fp32 (the HBM state type)
  float* const qdv = &A.at<float>(M.w_qd, 0);
  float* const tauv = &A.at<float>(M.w_tau, 0);

  // ---- load state ------------------------------------------------------------------------------
  for (int k = 0; k < M.n_q; ++k) qv[k * ST] = io.q_in[(size_t)k * ns + e];
  for (int k = 0; k < n; ++k) qdv[k * ST] = io.qd_in[(size_t)k * ns + e];
  for (int k = 0; k < n; ++k) tauv[k * ST] = 0.f;
  if (use_pd) {
    // PD torques, locomotion_contact_simulation.h:168-258
    for (int k = 0; k < E.n_act; ++k) {
      const int li = E.act_link[k];
      float a = io.tau_in[(size_t)k * ns + e];
      a = fminf(a, E.action_limit);
      a = fmaxf(a, -E.action_limit);
      const float q_des = E.initial_poses[k] + a;
      const float qa = qv[M.q_idx[li] * ST];
      const float qda = qdv[M.qd_idx[li] * ST];
      float f = E.kp * (q_des - qa) + E.kd * (0.f - qda);
      f = fminf(fmaxf(f, -E.max_force), E.max_force);
      tauv[M.qd_idx[li] * ST] = f;
    }
  } else if (io.tau_in) {
    const int off = M.floating ? 6 : 0;
    for (int k = off; k < n; ++k) tauv[k * ST] = io.tau_in[(size_t)(k - off) * ns + e];
  }
  for (int s = 0; s < M.n_acc; ++s) {
    for (int k = 0; k < AC_NRA; ++k) A.at<RA>(M.w_acc + s * M.acc_words, k) = RA(0);
    for (int k = 0; k < AC_NIC; ++k) A.at<RS>(M.w_acc + s * M.acc_words + M.acc_ic_word, k) = RS(0);
  }
  TDS_PHASE();  // 1: state loaded, PD done

  // ---- pass 1: kinematics root -> leaf (kinematics.hpp:18-148), positions in RC ---------------------
  Xf<RC> Xw_prev;
  Sv<RA> v_prev;
  M3<RC> baseR = m3_identity<RC>();
  if (M.floating) {
    baseR = quat_to_matrix<RC>(RC(qv[0]), RC(qv[ST]), RC(qv[2 * ST]), RC(qv[3 * ST]));
    Xw_prev.R = baseR;
    Xw_prev.t = v3<RC>(RC(qv[4 * ST]), RC(qv[5 * ST]), RC(qv[6 * ST]));
    v_prev.top = v3<RA>(RA(qdv[0]), RA(qdv[ST]), RA(qdv[2 * ST]));
    v_prev.bot = v3<RA>(RA(qdv[3 * ST]), RA(qdv[4 * ST]), RA(qdv[5 * ST]));
  } else {
    Xw_prev.R = baseR;
    Xw_prev.t = v3<RC>(RC(0), RC(0), RC(0));
    v_prev.top = v3<RA>(RA(0), RA(0), RA(0));
    v_prev.bot = v_prev.top;
  }
  const Xf<RC> Xw_base = Xw_prev;
  const Sv<RA> v_base = v_prev;
  st_xf<RC>(A, M.w_xw, Xw_base);
  constexpr int XWW = 12 * RCW;
  const int LW = M.link_words;
  for (int i = 0; i < n_links; ++i) {
    const int p = M.parent[i];
    const int fl = M.flags[i];
    Xf<RC> Xw_p;
    Sv<RA> v_p;
    if (fl & TDS_LF_PARENT_ADJ) { Xw_p = Xw_prev; v_p = v_prev; }
    else if (p >= 0) { Xw_p = ld_xf<RC>(A, M.w_xw + (p + 1) * XWW); v_p = ld_sv<RA>(A, M.w_link + p * LW, LK_VC); }
    else { Xw_p = Xw_base; v_p = v_base; }
    RC qi = (fl & TDS_LF_FIXED) ? RC(0) : RC(qv[M.q_idx[i] * ST]);
    Xf<RC> Xp = jcalc<RC>(M, i, qi);
    Xf<RC> Xw = xf_mul(Xw_p, Xp);
    st_xf<RC>(A, M.w_xw + (i + 1) * XWW, Xw);
    Xf<RA> XpA; XpA.R = cvt<RA>(Xp.R); XpA.t = cvt<RA>(Xp.t);
    st_xf<RA>(A, M.w_link + i * LW, XpA);
    Sv<RA> v = xf_apply_motion(XpA, v_p);
    if (!(fl & TDS_LF_FIXED)) {
      RA qdi = RA(qdv[M.qd_idx[i] * ST]);
      Sv<RA> S = link_S<RA>(M, i);
      v.top = v.top + S.top * qdi;
      v.bot = v.bot + S.bot * qdi;
    }
    st_sv<RA>(A, M.w_link + i * LW, LK_VC, v);
    if (io.link_xf && live) {
      float* o = io.link_xf + (size_t)i * 12 * ns + e;
      o[0] = (float)Xw.R.xx; o[(size_t)1 * ns] = (float)Xw.R.xy; o[(size_t)2 * ns] = (float)Xw.R.xz;
      o[(size_t)3 * ns] = (float)Xw.R.yx; o[(size_t)4 * ns] = (float)Xw.R.yy; o[(size_t)5 * ns] = (float)Xw.R.yz;
      o[(size_t)6 * ns] = (float)Xw.R.zx; o[(size_t)7 * ns] = (float)Xw.R.zy; o[(size_t)8 * ns] = (float)Xw.R.zz;
      o[(size_t)9 * ns] = (float)Xw.t.x; o[(size_t)10 * ns] = (float)Xw.t.y; o[(size_t)11 * ns] = (float)Xw.t.z;
    }
    Xw_prev = Xw;
    v_prev = v;
  }
  TDS_PHASE();  // 2: pass 1 done

  // ---- contact detection (world.hpp:206-282, contact_point.hpp:97-161) ----------------------------
  // Every sphere / capsule end emits one candidate point in the reference; only penetrating points
  // produce non-zero LCP rows, so only those are recorded for the solve.
  int n_active = 0;
  if (mode == MODE_FULL && M.has_plane) {
    const V3<RC> pn = v3<RC>(RC(M.plane_n[0]), RC(M.plane_n[1]), RC(M.plane_n[2]));
    int pt = 0;
    for (int g = 0; g < M.n_geoms; ++g) {
      const int L = M.g_link[g];
      const int ty = M.g_type[g];
      if (ty != TDSG_SPHERE && ty != TDSG_CAPSULE) continue;
      Xf<RC> Xw = ld_xf<RC>(A, M.w_xw + (L + 1) * XWW);
      V3<RC> c = Xw.t + mul(Xw.R, v3<RC>(RC(M.g_t[g][0]), RC(M.g_t[g][1]), RC(M.g_t[g][2])));
      const RC rad = RC(M.g_radius[g]);
      const int npts = (ty == TDSG_CAPSULE) ? 2 : 1;
      V3<RC> half = v3<RC>(RC(0), RC(0), RC(0));
      if (ty == TDSG_CAPSULE) half = mul(Xw.R, v3<RC>(RC(M.g_half[g][0]), RC(M.g_half[g][1]), RC(M.g_half[g][2])));
      for (int k = 0; k < npts; ++k) {
        V3<RC> pos = (ty == TDSG_CAPSULE) ? (k == 0 ? c + half : c - half) : c;
        const RC t = dot(pos, pn) - RC(M.plane_c);   // contact_point.hpp:112
        const RC dist = t - rad;
        if (io.contact_dist && live) io.contact_dist[(size_t)pt * ns + e] = (float)dist;
        ++pt;
        if (dist < RC(0) && n_active < M.max_contacts) {
          const int w = M.w_con + n_active * CN_SIZE * RCW;
          st_v3<RC>(A, w, CN_PB, pos - pn * rad);      // world_point_on_b
          A.at<RC>(w, CN_DIST) = dist;
          A.at<RC>(w, CN_LINK) = RC(L);
          ++n_active;
        }
      }
    }
  }
  const bool any_contact = __any_sync(0xffffffffu, n_active > 0);
  TDS_PHASE();  // 3: contacts detected

  // ---- pass 2: leaf -> root.  ABA (forward_dynamics.hpp:50-216) + CRBA (mass_matrix.hpp:39-125) ---
  RS* const Mm = &A.at<RS>(M.w_M, 0);          // lower triangle of M, later of its Cholesky factor
  Abi<RA> cA;
  Sv<RA> cP;
  Rbi<RS> cC;   // composite rigid-body inertia of the CRBA
  if (any_contact)
    for (int k = 0; k < n * (n + 1) / 2; ++k) Mm[k * ST] = RS(0);
  for (int i = n_links - 1; i >= 0; --i) {
    const int p = M.parent[i];
    const int fl = M.flags[i];
    const int wl = M.w_link + i * LW;
    const Xf<RA> Xp = ld_xf<RA>(A, wl);
    const Sv<RA> v = ld_sv<RA>(A, wl, LK_VC);
    const Rbi<RA> rb = model_rbi<RA>(M.rbi[i]);
    Rbi<RS> Ic = model_rbi<RS>(M.rbi[i]);
    Abi<RA> Ai = abi_from_rbi(rb);
    Sv<RA> pA = cross_mf(v, rbi_mul(rb, v));        // kinematics.hpp:132
    if (fl & TDS_LF_CHILD_ADJ) { abi_add(Ai, cA); pA = pA + cP; rbi_add(Ic, cC); }
    if (M.acc_slot[i] >= 0) {
      Abi<RA> sa; Sv<RA> sp;
      const int ws = M.w_acc + M.acc_slot[i] * M.acc_words;
      acc_load<RA>(A, ws, sa, sp);
      abi_add(Ai, sa); pA = pA + sp; rbi_add(Ic, acc_load_rbi<RS>(A, ws + M.acc_ic_word));
    }
    Sv<RA> pa = pA;
    Abi<RA> Ia = Ai;
    if (fl & TDS_LF_FIXED) {
      Sv<RA> z; z.top = v3<RA>(RA(0), RA(0), RA(0)); z.bot = z.top;
      st_sv<RA>(A, wl, LK_VC, z);
      st_sv<RA>(A, wl, LK_U, z);
      A.at<RA>(wl, LK_INVD) = RA(0);
      A.at<RA>(wl, LK_u) = RA(0);
    } else {
      const Sv<RA> S = link_S<RA>(M, i);
      const int qdi = M.qd_idx[i];
      const RA qdj = RA(qdv[qdi * ST]);
      Sv<RA> vJ; vJ.top = S.top * qdj; vJ.bot = S.bot * qdj;
      const Sv<RA> c = cross_mm(v, vJ);               // kinematics.hpp:96-97
      const Sv<RA> U = abi_mul(Ai, S);                // forward_dynamics.hpp:111
      const RA D = dot(S, U);
      const RA invD = RA(1) / D;
      RA tau = RA(tauv[qdi * ST]);
      tau -= RA(M.stiffness[i]) * RA(qv[M.q_idx[i] * ST]);
      tau -= RA(M.damping[i]) * qdj;
      const RA u = tau - dot(S, pA);                  // :129
      st_sv<RA>(A, wl, LK_VC, c);
      st_sv<RA>(A, wl, LK_U, U);
      A.at<RA>(wl, LK_INVD) = invD;
      A.at<RA>(wl, LK_u) = u;
      // Ia = abi - U (U/D)^T, :160-168
      const V3<RA> ut = U.top * invD, ub = U.bot * invD;
      Ia.I.xx -= U.top.x * ut.x; Ia.I.xy -= U.top.x * ut.y; Ia.I.xz -= U.top.x * ut.z;
      Ia.I.yy -= U.top.y * ut.y; Ia.I.yz -= U.top.y * ut.z; Ia.I.zz -= U.top.z * ut.z;
      Ia.H.xx -= U.top.x * ub.x; Ia.H.xy -= U.top.x * ub.y; Ia.H.xz -= U.top.x * ub.z;
      Ia.H.yx -= U.top.y * ub.x; Ia.H.yy -= U.top.y * ub.y; Ia.H.yz -= U.top.y * ub.z;
      Ia.H.zx -= U.top.z * ub.x; Ia.H.zy -= U.top.z * ub.y; Ia.H.zz -= U.top.z * ub.z;
      Ia.M.xx -= U.bot.x * ub.x; Ia.M.xy -= U.bot.x * ub.y; Ia.M.xz -= U.bot.x * ub.z;
      Ia.M.yy -= U.bot.y * ub.y; Ia.M.yz -= U.bot.y * ub.z; Ia.M.zz -= U.bot.z * ub.z;
      const Sv<RA> Iac = abi_mul(Ia, c);              // :171
      const RA uD = u * invD;
      pa.top = pA.top + Iac.top + U.top * uD;         // :173
      pa.bot = pA.bot + Iac.bot + U.bot * uD;
      // CRBA column of this joint, mass_matrix.hpp:86-111 (only needed when some lane has contacts)
      if (any_contact) {
        const Sv<RS> Sc = link_S<RS>(M, i);
        Sv<RS> F = rbi_mul(Ic, Sc);
        RS* const row = Mm + tri(qdi, 0) * ST;
        row[qdi * ST] = dot(Sc, F);
        int j = i;
        Xf<RS> Xj; Xj.R = cvt<RS>(Xp.R); Xj.t = cvt<RS>(Xp.t);
        while (true) {
          F = xf_apply_force(Xj, F);
          j = M.parent[j];
          if (j < 0) break;
          if (!(M.flags[j] & TDS_LF_FIXED)) row[M.qd_idx[j] * ST] = dot(F, link_S<RS>(M, j));
          const Xf<RA> Xa = ld_xf<RA>(A, M.w_link + j * LW);
          Xj.R = cvt<RS>(Xa.R); Xj.t = cvt<RS>(Xa.t);
        }
        if (M.floating) {
          row[0] = F.top.x; row[ST] = F.top.y; row[2 * ST] = F.top.z;
          row[3 * ST] = F.bot.x; row[4 * ST] = F.bot.y; row[5 * ST] = F.bot.z;
        }
      }
    }
    // propagate to the parent: register carry along chains, accumulator at branch points
    const Abi<RA> dA = xt_abi_x(Xp, Ia);               // :187-189
    const Sv<RA> dP = xf_apply_force(Xp, pa);          // :181
    Rbi<RS> dC = Ic;
    if (any_contact) { Xf<RS> Xc; Xc.R = cvt<RS>(Xp.R); Xc.t = cvt<RS>(Xp.t); dC = xt_rbi_x(Xc, Ic); }  // mass_matrix.hpp:45-46
    if (fl & TDS_LF_PARENT_ADJ) { cA = dA; cP = dP; cC = dC; }
    else {
      const int slot = (p >= 0) ? M.acc_slot[p] : M.base_acc;
      if (slot >= 0) {
        const int w = M.w_acc + slot * M.acc_words;
        acc_add_abi<RA>(A, w, dA, dP);
        acc_add_rbi<RS>(A, w + M.acc_ic_word, dC);
      }
    }
  }
  TDS_PHASE();  // 4: pass 2 (ABA + CRBA) done

  // ---- base acceleration (forward_dynamics.hpp:218-243) -----------------------------------------
  Sv<RA> a_prev;
  Sv<RC> base_acc;
  if (M.floating) {
    Rbi<RS> Ib = model_rbi<RS>(M.base_rbi);
    Abi<RA> Ab = abi_from_rbi(model_rbi<RA>(M.base_rbi));
    // gyroscopic bias, kinematics.hpp:54-61
    M3<RA> Rb = cvt<RA>(baseR);
    M3<RA> Ic0;
    Ic0.xx = RA(M.base_inertia_com[0]); Ic0.xy = RA(M.base_inertia_com[1]); Ic0.xz = RA(M.base_inertia_com[2]);
    Ic0.yx = RA(M.base_inertia_com[3]); Ic0.yy = RA(M.base_inertia_com[4]); Ic0.yz = RA(M.base_inertia_com[5]);
    Ic0.zx = RA(M.base_inertia_com[6]); Ic0.zy = RA(M.base_inertia_com[7]); Ic0.zz = RA(M.base_inertia_com[8]);
    M3<RA> Iw = rot_gen(Rb, Ic0);
    Sv<RA> pb; pb.top = cross(v_base.top, mul(Iw, v_base.top)); pb.bot = v3<RA>(RA(0), RA(0), RA(0));
    if (n_links > 0 && M.parent[0] < 0) { abi_add(Ab, cA); pb = pb + cP; rbi_add(Ib, cC); }
    if (M.base_acc >= 0) {
      Abi<RA> sa; Sv<RA> sp;
      const int ws = M.w_acc + M.base_acc * M.acc_words;
      acc_load<RA>(A, ws, sa, sp);
      abi_add(Ab, sa); pb = pb + sp; rbi_add(Ib, acc_load_rbi<RS>(A, ws + M.acc_ic_word));
    }
    if (any_contact) {  // mass_matrix.hpp:114-120: base block = composite inertia [I  hx; hx^T  m1]
      const RS z = RS(0);
      Mm[tri(0, 0) * ST] = Ib.I.xx; Mm[tri(1, 0) * ST] = Ib.I.xy; Mm[tri(1, 1) * ST] = Ib.I.yy;
      Mm[tri(2, 0) * ST] = Ib.I.xz; Mm[tri(2, 1) * ST] = Ib.I.yz; Mm[tri(2, 2) * ST] = Ib.I.zz;
      Mm[tri(3, 0) * ST] = z;        Mm[tri(3, 1) * ST] = Ib.h.z;  Mm[tri(3, 2) * ST] = -Ib.h.y;
      Mm[tri(4, 0) * ST] = -Ib.h.z;  Mm[tri(4, 1) * ST] = z;       Mm[tri(4, 2) * ST] = Ib.h.x;
      Mm[tri(5, 0) * ST] = Ib.h.y;   Mm[tri(5, 1) * ST] = -Ib.h.x; Mm[tri(5, 2) * ST] = z;
      Mm[tri(3, 3) * ST] = Ib.m; Mm[tri(4, 3) * ST] = z; Mm[tri(4, 4) * ST] = Ib.m;
      Mm[tri(5, 3) * ST] = z; Mm[tri(5, 4) * ST] = z; Mm[tri(5, 5) * ST] = Ib.m;
    }
    // -base_abi.inv_mul(bias) with the reference's block inverse (C = -H), inertia.hpp:302-328
    {
      M3<RC> I3, H3, M3m;
      I3.xx = Ab.I.xx; I3.xy = Ab.I.xy; I3.xz = Ab.I.xz; I3.yx = Ab.I.xy; I3.yy = Ab.I.yy; I3.yz = Ab.I.yz; I3.zx = Ab.I.xz; I3.zy = Ab.I.yz; I3.zz = Ab.I.zz;
      H3 = cvt<RC>(Ab.H);
      M3m.xx = Ab.M.xx; M3m.xy = Ab.M.xy; M3m.xz = Ab.M.xz; M3m.yx = Ab.M.xy; M3m.yy = Ab.M.yy; M3m.yz = Ab.M.yz; M3m.zx = Ab.M.xz; M3m.zy = Ab.M.yz; M3m.zz = Ab.M.zz;
      auto inv3 = [](const M3<RC>& m) {
        M3<RC> o;
        RC c0 = m.yy * m.zz - m.yz * m.zy, c1 = m.yz * m.zx - m.yx * m.zz, c2 = m.yx * m.zy - m.yy * m.zx;
        RC s = RC(1) / (m.xx * c0 + m.xy * c1 + m.xz * c2);
        o.xx = c0 * s; o.xy = (m.xz * m.zy - m.xy * m.zz) * s; o.xz = (m.xy * m.yz - m.xz * m.yy) * s;
        o.yx = c1 * s; o.yy = (m.xx * m.zz - m.xz * m.zx) * s; o.yz = (m.xz * m.yx - m.xx * m.yz) * s;
        o.zx = c2 * s; o.zy = (m.xy * m.zx - m.xx * m.zy) * s; o.zz = (m.xx * m.yy - m.xy * m.yx) * s;
        return o;
      };
      auto neg = [](M3<RC> m) { m.xx = -m.xx; m.xy = -m.xy; m.xz = -m.xz; m.yx = -m.yx; m.yy = -m.yy; m.yz = -m.yz; m.zx = -m.zx; m.zy = -m.zy; m.zz = -m.zz; return m; };
      auto sub = [](M3<RC> a, const M3<RC>& b) { a.xx -= b.xx; a.xy -= b.xy; a.xz -= b.xz; a.yx -= b.yx; a.yy -= b.yy; a.yz -= b.yz; a.zx -= b.zx; a.zy -= b.zy; a.zz -= b.zz; return a; };
      auto add = [](M3<RC> a, const M3<RC>& b) { a.xx += b.xx; a.xy += b.xy; a.xz += b.xz; a.yx += b.yx; a.yy += b.yy; a.yz += b.yz; a.zx += b.zx; a.zy += b.zy; a.zz += b.zz; return a; };
      M3<RC> Ainv = inv3(I3);
      M3<RC> C = neg(H3);
      M3<RC> D = inv3(sub(M3m, mul(mul(C, Ainv), H3)));
      M3<RC> AinvBD = mul(mul(Ainv, H3), D);
      M3<RC> Ii = add(Ainv, mul(mul(AinvBD, C), Ainv));
      M3<RC> Hi = neg(AinvBD);
      V3<RC> ft = cvt<RC>(pb.top), fb = cvt<RC>(pb.bot);
      V3<RC> at = mul(Ii, ft) + mul(Hi, fb);
      V3<RC> ab = mul(D, fb) + mulT(Hi, ft);
      base_acc.top = v3<RC>(-at.x, -at.y, -at.z);
      base_acc.bot = v3<RC>(-ab.x, -ab.y, -ab.z);
    }
  } else {
    base_acc.top = v3<RC>(RC(0), RC(0), RC(0));
    base_acc.bot = v3<RC>(RC(-P.gravity[0]), RC(-P.gravity[1]), RC(-P.gravity[2]));
  }
  a_prev.top = cvt<RA>(base_acc.top);
  a_prev.bot = cvt<RA>(base_acc.bot);
  const Sv<RA> a_base = a_prev;
  TDS_PHASE();  // 5: base done

  // ---- pass 3: root -> leaf accelerations (forward_dynamics.hpp:245-302) + integrate_euler_qdd ----
  for (int i = 0; i < n_links; ++i) {
    const int p = M.parent[i];
    const int fl = M.flags[i];
    const int wl = M.w_link + i * LW;
    Sv<RA> a_p;
    if (fl & TDS_LF_PARENT_ADJ) a_p = a_prev;
    else if (p >= 0) a_p = ld_sv<RA>(A, M.w_link + p * LW, LK_VC);
    else a_p = a_base;
    const Xf<RA> Xp = ld_xf<RA>(A, wl);
    Sv<RA> a = xf_apply_motion(Xp, a_p);
    if (!(fl & TDS_LF_FIXED)) {
      const Sv<RA> c = ld_sv<RA>(A, wl, LK_VC);
      const Sv<RA> U = ld_sv<RA>(A, wl, LK_U);
      a = a + c;
      const RA qdd = A.at<RA>(wl, LK_INVD) * (A.at<RA>(wl, LK_u) - dot(U, a));
      const Sv<RA> S = link_S<RA>(M, i);
      a.top = a.top + S.top * qdd;
      a.bot = a.bot + S.bot * qdd;
      const int qdi = M.qd_idx[i];
      if (mode == MODE_FD) { if (live && io.qdd_out) io.qdd_out[(size_t)qdi * ns + e] = (float)qdd; }
      else qdv[qdi * ST] = (float)(RA(qdv[qdi * ST]) + qdd * dtA);
    }
    st_sv<RA>(A, wl, LK_VC, a);
    a_prev = a;
  }
  if (M.floating) {  // forward_dynamics.hpp:317-322, integrator.hpp:153-163
    RC qb[6] = {base_acc.top.x, base_acc.top.y, base_acc.top.z, base_acc.bot.x + RC(P.gravity[0]),
                base_acc.bot.y + RC(P.gravity[1]), base_acc.bot.z + RC(P.gravity[2])};
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      if (mode == MODE_FD) { if (live && io.qdd_out) io.qdd_out[(size_t)k * ns + e] = (float)qb[k]; }
      else qdv[k * ST] = (float)(RC(qdv[k * ST]) + qb[k] * RC(P.dt));
    }
  }
  TDS_PHASE();  // 6: pass 3 done
  if (mode == MODE_FD) return;

  // ---- contact solve (RS arithmetic; only the right-hand side b needs the RC positions) ------------
  if (mode == MODE_FULL && any_contact) {
    RS* const invd = &A.at<RS>(M.w_invd, 0);
    RS* const wv = &A.at<RS>(M.w_w, 0);
    // Cholesky M = L L^T in place, row by row (L[i][j] needs rows i and j only; both are contiguous in the
    // row-major lower triangle).  The reference inverts M (tiny_matrix_x.h:240-344); only products with
    // M^-1 are needed here.
    for (int i = 0; i < n; ++i) {
      RS* const ri = Mm + tri(i, 0) * ST;
      for (int j = 0; j < i; ++j) {
        const RS* rj = Mm + tri(j, 0) * ST;
        ri[j * ST] = (ri[j * ST] - sdot(ri, rj, ST, j)) * invd[j * ST];
      }
      const RS d = ri[i * ST] - sdot(ri, ri, ST, i);
      const RS l = sqrt_t(d);
      ri[i * ST] = l;
      invd[i * ST] = RS(1) / l;
    }
    TDS_PHASE();  // 7: Cholesky done
    const int max_active = __reduce_max_sync(0xffffffffu, n_active);
    const V3<RC> nb = v3<RC>(RC(-M.plane_n[0]), RC(-M.plane_n[1]), RC(-M.plane_n[2]));  // world_normal_on_b
    const V3<RC> f1 = v3<RC>(RC(M.fr1[0]), RC(M.fr1[1]), RC(M.fr1[2]));
    const V3<RC> f2 = v3<RC>(RC(M.fr2[0]), RC(M.fr2[1]), RC(M.fr2[2]));
    for (int c = 0; c < max_active; ++c) {
      if (c < n_active) {
        const int wc = M.w_con + c * CN_SIZE * RCW;
        RS* const y0 = &A.at<RS>(M.w_Y, 0) + (c * 3) * n * ST;
        RS* const y1 = y0 + n * ST;
        RS* const y2 = y1 + n * ST;
        const V3<RC> pb = ld_v3<RC>(A, wc, CN_PB);
        const RC dist = A.at<RC>(wc, CN_DIST);
        const int L = (int)A.at<RC>(wc, CN_LINK);
        for (int k = 0; k < 3 * n; ++k) y0[k * ST] = RS(0);
        V3<RC> vel = v3<RC>(RC(0), RC(0), RC(0));   // vel_b = J qd
        if (M.floating) {  // jacobian.hpp:39-58
          const V3<RC> r = pb - Xw_base.t;
          // J[:,0:3] = cross_matrix(r)^T, J[:,3:6] = 1
          const V3<RC> c0 = v3<RC>(RC(0), -r.z, r.y), c1 = v3<RC>(r.z, RC(0), -r.x), c2 = v3<RC>(-r.y, r.x, RC(0));
          const V3<RC> e0 = v3<RC>(RC(1), RC(0), RC(0)), e1 = v3<RC>(RC(0), RC(1), RC(0)), e2 = v3<RC>(RC(0), RC(0), RC(1));
          const V3<RC> cols[6] = {c0, c1, c2, e0, e1, e2};
#pragma unroll
          for (int k = 0; k < 6; ++k) {
            y0[k * ST] = RS(dot(nb, cols[k]));
            y1[k * ST] = RS(dot(f1, cols[k]));
            y2[k * ST] = RS(dot(f2, cols[k]));
            vel = vel + cols[k] * RC(qdv[k * ST]);
          }
        }
        for (int j = L; j >= 0; j = M.parent[j]) {  // jacobian.hpp:63-80
          if (M.flags[j] & TDS_LF_FIXED) continue;
          const Xf<RC> Xw = ld_xf<RC>(A, M.w_xw + (j + 1) * XWW);
          const Sv<RC> S = link_S<RC>(M, j);
          const V3<RC> wv3 = mul(Xw.R, S.top);
          const V3<RC> col = mul(Xw.R, S.bot) + cross(wv3, pb - Xw.t);
          const int qj = M.qd_idx[j];
          y0[qj * ST] = RS(dot(nb, col));
          y1[qj * ST] = RS(dot(f1, col));
          y2[qj * ST] = RS(dot(f2, col));
          vel = vel + col * RC(qdv[qj * ST]);
        }
        // rel_vel = vel_a - vel_b = -vel ; mb_constraint_solver.hpp:299-345
        const RC nrv = -dot(nb, vel);
        RS* const cs = &A.at<RS>(M.w_conS + c * 6 * RSW, 0);   // b[3], x[3]
        cs[0] = RS(-(RC(1) + RC(P.restitution)) * nrv - RC(P.erp) * dist / RC(P.dt));
        cs[ST] = RS(dot(f1, vel));
        cs[2 * ST] = RS(dot(f2, vel));
        cs[3 * ST] = RS(0); cs[4 * ST] = RS(0); cs[5 * ST] = RS(0);
        // Y rows = L^-1 * Jc rows: forward substitution, the three right-hand sides share the loads of L
        for (int i = 0; i < n; ++i) {
          const RS* ri = Mm + tri(i, 0) * ST;
          RS a0 = RS(0), a1 = RS(0), a2 = RS(0), b0 = RS(0), b1 = RS(0), b2 = RS(0);
          int k = 0;
          for (; k + 1 < i; k += 2) {
            const RS l0 = ri[k * ST], l1 = ri[(k + 1) * ST];
            a0 += l0 * y0[k * ST]; a1 += l0 * y1[k * ST]; a2 += l0 * y2[k * ST];
            b0 += l1 * y0[(k + 1) * ST]; b1 += l1 * y1[(k + 1) * ST]; b2 += l1 * y2[(k + 1) * ST];
          }
          if (k < i) { const RS l0 = ri[k * ST]; a0 += l0 * y0[k * ST]; a1 += l0 * y1[k * ST]; a2 += l0 * y2[k * ST]; }
          const RS inv = invd[i * ST];
          y0[i * ST] = (y0[i * ST] - (a0 + b0)) * inv;
          y1[i * ST] = (y1[i * ST] - (a1 + b1)) * inv;
          y2[i * ST] = (y2[i * ST] - (a2 + b2)) * inv;
        }
      }
    }
    TDS_PHASE();  // 8: Jacobians + Y done
    // matrix-free projected Gauss-Seidel on w = Y p; row order normals | friction-1 | friction-2
    // (solve_pgs, mb_constraint_solver.hpp:101-142; bounds :417-436)
    for (int k = 0; k < n; ++k) wv[k * ST] = RS(0);
    const RS cfm = RS(P.cfm), mu = RS(P.friction);
    for (int it = 0; it < P.pgs_iterations; ++it) {
      for (int blk = 0; blk < 3; ++blk) {
        for (int c = 0; c < max_active; ++c) {
          if (c < n_active) {
            RS* const cs = &A.at<RS>(M.w_conS + c * 6 * RSW, 0);
            const RS* y = &A.at<RS>(M.w_Y, 0) + (c * 3 + blk) * n * ST;
            RS yy0 = RS(0), yy1 = RS(0), yw0 = RS(0), yw1 = RS(0);
            int k = 0;
            for (; k + 1 < n; k += 2) {
              const RS ya = y[k * ST], yb = y[(k + 1) * ST];
              yy0 += ya * ya; yw0 += ya * wv[k * ST];
              yy1 += yb * yb; yw1 += yb * wv[(k + 1) * ST];
            }
            if (k < n) { const RS ya = y[k * ST]; yy0 += ya * ya; yw0 += ya * wv[k * ST]; }
            const RS yy = yy0 + yy1, yw = yw0 + yw1;
            const RS x_old = cs[(3 + blk) * ST];
            RS x = (cs[blk * ST] - yw + yy * x_old) / (yy + cfm);
            if (blk == 0) {
              x = x < RS(0) ? RS(0) : x;
              x = x > RS(100000) ? RS(100000) : x;
            } else {
              RS s = cs[3 * ST];
              s = s < RS(0) ? RS(0) : s;
              const RS lim = mu * s;
              x = x < -lim ? -lim : x;
              x = x > lim ? lim : x;
            }
            cs[(3 + blk) * ST] = x;
            const RS dx = x - x_old;
            for (int k2 = 0; k2 < n; ++k2) wv[k2 * ST] += dx * y[k2 * ST];
          }
        }
      }
    }
    TDS_PHASE();  // 9: PGS done
    // qd_b -= M^-1 Jc^T p = L^-T w   (mb_constraint_solver.hpp:476-497)
    for (int i = n - 1; i >= 0; --i) {
      RS s0 = wv[i * ST], s1 = RS(0);
      int k = i + 1;
      for (; k + 1 < n; k += 2) {
        s0 -= Mm[(tri(k, 0) + i) * ST] * wv[k * ST];
        s1 -= Mm[(tri(k + 1, 0) + i) * ST] * wv[(k + 1) * ST];
      }
      if (k < n) s0 -= Mm[(tri(k, 0) + i) * ST] * wv[k * ST];
      const RS z = (s0 + s1) * invd[i * ST];
      wv[i * ST] = z;
      if (n_active > 0) qdv[i * ST] = (float)(RS(qdv[i * ST]) - z);
    }
  }
  TDS_PHASE();  // 10: impulses applied

  // ---- integrate_euler with qdd = 0 (integrator.hpp:10-133) ---------------------------------------
  RC up_z = RC(1);
  if (M.floating) {
    const RC h = RC(0.5) * RC(P.dt);
    RC qx = RC(qv[0]), qy = RC(qv[ST]), qz = RC(qv[2 * ST]), qw = RC(qv[3 * ST]);
    const RC w0 = RC(qdv[0]), w1 = RC(qdv[ST]), w2 = RC(qdv[2 * ST]);
    const RC dw = (-qx * w0 - qy * w1 - qz * w2) * h;
    const RC dx = (qw * w0 + qz * w1 - qy * w2) * h;
    const RC dy = (qw * w1 + qx * w2 - qz * w0) * h;
    const RC dz = (qw * w2 + qy * w0 - qx * w1) * h;
    qx += dx; qy += dy; qz += dz; qw += dw;
    const RC len = sqrt_t(qx * qx + qy * qy + qz * qz + qw * qw);
    qx /= len; qy /= len; qz /= len; qw /= len;
    qv[0] = (float)qx; qv[ST] = (float)qy; qv[2 * ST] = (float)qz; qv[3 * ST] = (float)qw;
    for (int k = 0; k < 3; ++k)
      qv[(4 + k) * ST] = (float)(RC(qv[(4 + k) * ST]) + RC(qdv[(3 + k) * ST]) * RC(P.dt));
    up_z = RC(1) - RC(2) * (qx * qx + qy * qy) / (qx * qx + qy * qy + qz * qz + qw * qw);
  }
  for (int i = 0; i < n_links; ++i) {
    if (M.flags[i] & TDS_LF_FIXED) continue;
    const int qi = M.q_idx[i];
    qv[qi * ST] = (float)(RC(qv[qi * ST]) + RC(qdv[M.qd_idx[i] * ST]) * RC(P.dt));
  }
  TDS_PHASE();  // 11: integrated

  // ---- reward / done / auto-reset, write back --------------------------------------------------------
  if (live) {
    bool done = false;
    if (E.reward_kind == 1) {
      // laikago_environment2.h:130-171, fixed-base emulation: x = q0, z = q2, rpy = q3..5;
      // up.z of quat_to_matrix(quat_from_euler_rpy(rpy)) = cos(roll) cos(pitch)
      const float x = qv[0], z = qv[2 * ST];
      const float upz = cosf(qv[3 * ST]) * cosf(qv[4 * ST]);
      done = (upz < 0.6f) || (z < 0.2f);
      if (io.reward) io.reward[e] = done ? 0.f : x;
    } else if (E.reward_kind == 2) {
      const float x = qv[4 * ST], z = qv[6 * ST];
      done = ((float)up_z < 0.6f) || (z < 0.2f);
      if (io.reward) io.reward[e] = done ? 0.f : x;
    } else if (E.reward_kind == 3) {   // ant_environment2.h:75-105: done = z < 0.26, reward = (x' - x)/dt, which integrate_euler makes the x velocity
      done = qv[2 * ST] < 0.26f;
      if (io.reward) io.reward[e] = done ? 0.f : qdv[0];
    }
    if (io.done && E.reward_kind) io.done[e] = done ? 1.f : 0.f;
    if (done && E.auto_reset) {
      // VectorizedEnvironment auto_reset_when_done (ars_vectorized_environment.h:262-283): back to the
      // reset pose (deterministic; the host-side reset() adds the reference's joint noise and settle steps)
      for (int k = 0; k < M.n_q; ++k) io.q_out[(size_t)k * ns + e] = E.reset_q[k];
      for (int k = 0; k < n; ++k) io.qd_out[(size_t)k * ns + e] = 0.f;
    } else {
      for (int k = 0; k < M.n_q; ++k) io.q_out[(size_t)k * ns + e] = qv[k * ST];
      for (int k = 0; k < n; ++k) io.qd_out[(size_t)k * ns + e] = qdv[k * ST];
    }
  }
}

}  // namespace tds

// ---- host launchers -------------------------------------------------------------------------------
extern "C" int tds_launch_step(const DevModel* M, const SimParams* P, const EnvParams* E, const StepIO* io,
                               int mode, int use_pd, int precision, char* gscratch, int use_smem,
                               int warps_per_block, cudaStream_t stream) {
  using namespace tds;
  const int threads = 32 * warps_per_block;
  const int blocks = (io->n + threads - 1) / threads;
  const size_t smem = use_smem ? (size_t)warps_per_block * M->w_total * 32 * 4 : 0;
  cudaError_t err = cudaSuccess;
#define TDS_LAUNCH(RA, RC, RS, SM)                                                                      \
  do {                                                                                                  \
    auto k = tds_step_kernel<RA, RC, RS, SM>;                                                           \
    static size_t smem_set = 0; /* opt-in once per instantiation, not per launch */                     \
    if (smem > 48 * 1024 && smem > smem_set) {                                                          \
      err = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);            \
      if (err == cudaSuccess) smem_set = smem;                                                          \
    }                                                                                                   \
    if (err == cudaSuccess) {                                                                           \
      k<<<blocks, threads, smem, stream>>>(*M, *P, *E, *io, mode, use_pd, gscratch);                    \
      err = cudaGetLastError();                                                                         \
    }                                                                                                   \
  } while (0)
  if (precision == 0) {        // mixed: fp32 ABA + fp32 contact solve, fp64 kinematics / contact geometry
    if (use_smem) TDS_LAUNCH(float, double, float, true); else TDS_LAUNCH(float, double, float, false);
  } else if (precision == 1) { // all fp64
    if (use_smem) TDS_LAUNCH(double, double, double, true); else TDS_LAUNCH(double, double, double, false);
  } else {                     // all fp32
    if (use_smem) TDS_LAUNCH(float, float, float, true); else TDS_LAUNCH(float, float, float, false);
  }
#undef TDS_LAUNCH
  return (int)err;
}
