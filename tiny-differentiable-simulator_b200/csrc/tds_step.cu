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
//   * mixed precision: ABA in RA (fp32), kinematics + contact solve in RC (fp64) because
//     erp/dt = 200 amplifies fp32 position round-off beyond the 1e-5 parity budget.
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
// contact record (units of RC): pb 3, dist, link, b[3], x[3]
enum { CN_PB = 0, CN_DIST = 3, CN_LINK = 4, CN_B = 5, CN_X = 8, CN_SIZE = 11 };

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

template <typename RA, typename RC, bool SMEM>
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
    A.blk = gscratch;
    A.stride = io.n_stride;
    A.col = e;
  }
  const int ns = io.n_stride;
  const int n_links = M.n_links;
  int phase_id = 0;
#define TDS_PHASE() do { if (io.phase_clk && lane == 0) io.phase_clk[(size_t)(env >> 5) * 16 + (phase_id++)] = clock64(); } while (0)
  TDS_PHASE();
  const int n = M.n_qd;
  const RA dtA = RA(P.dt);

  // ---- load state ------------------------------------------------------------------------------
  for (int k = 0; k < M.n_q; ++k) A.at<float>(M.w_q, k) = io.q_in[(size_t)k * ns + e];
  for (int k = 0; k < n; ++k) A.at<float>(M.w_qd, k) = io.qd_in[(size_t)k * ns + e];
  for (int k = 0; k < n; ++k) A.at<float>(M.w_tau, k) = 0.f;
  if (use_pd) {
    // PD torques, locomotion_contact_simulation.h:168-258
    for (int k = 0; k < E.n_act; ++k) {
      const int li = E.act_link[k];
      float a = io.tau_in[(size_t)k * ns + e];
      a = fminf(a, E.action_limit);
      a = fmaxf(a, -E.action_limit);
      const float q_des = E.initial_poses[k] + a;
      const float qa = A.at<float>(M.w_q, M.q_idx[li]);
      const float qda = A.at<float>(M.w_qd, M.qd_idx[li]);
      float f = E.kp * (q_des - qa) + E.kd * (0.f - qda);
      f = fminf(fmaxf(f, -E.max_force), E.max_force);
      A.at<float>(M.w_tau, M.qd_idx[li]) = f;
    }
  } else if (io.tau_in) {
    const int off = M.floating ? 6 : 0;
    for (int k = off; k < n; ++k) A.at<float>(M.w_tau, k) = io.tau_in[(size_t)(k - off) * ns + e];
  }
  for (int s = 0; s < M.n_acc; ++s) {
    for (int k = 0; k < AC_NRA; ++k) A.at<RA>(M.w_acc + s * M.acc_words, k) = RA(0);
    for (int k = 0; k < AC_NIC; ++k) A.at<RC>(M.w_acc + s * M.acc_words + M.acc_ic_word, k) = RC(0);
  }

  TDS_PHASE();  // 1: state loaded, PD done
  // ---- pass 1: kinematics root -> leaf (kinematics.hpp:18-148) -----------------------------------
  Xf<RC> Xw_prev;
  Sv<RA> v_prev;
  M3<RC> baseR = m3_identity<RC>();
  if (M.floating) {
    baseR = quat_to_matrix<RC>(RC(A.at<float>(M.w_q, 0)), RC(A.at<float>(M.w_q, 1)), RC(A.at<float>(M.w_q, 2)), RC(A.at<float>(M.w_q, 3)));
    Xw_prev.R = baseR;
    Xw_prev.t = v3<RC>(RC(A.at<float>(M.w_q, 4)), RC(A.at<float>(M.w_q, 5)), RC(A.at<float>(M.w_q, 6)));
    v_prev.top = v3<RA>(RA(A.at<float>(M.w_qd, 0)), RA(A.at<float>(M.w_qd, 1)), RA(A.at<float>(M.w_qd, 2)));
    v_prev.bot = v3<RA>(RA(A.at<float>(M.w_qd, 3)), RA(A.at<float>(M.w_qd, 4)), RA(A.at<float>(M.w_qd, 5)));
  } else {
    Xw_prev.R = baseR;
    Xw_prev.t = v3<RC>(RC(0), RC(0), RC(0));
    v_prev.top = v3<RA>(RA(0), RA(0), RA(0));
    v_prev.bot = v_prev.top;
  }
  const Xf<RC> Xw_base = Xw_prev;
  const Sv<RA> v_base = v_prev;
  st_xf<RC>(A, M.w_xw, Xw_base);
  const int XWW = 12 * (int)(sizeof(RC) / 4);
  const int LW = M.link_words;
  for (int i = 0; i < n_links; ++i) {
    const int p = M.parent[i];
    const int fl = M.flags[i];
    Xf<RC> Xw_p;
    Sv<RA> v_p;
    if (fl & TDS_LF_PARENT_ADJ) { Xw_p = Xw_prev; v_p = v_prev; }
    else if (p >= 0) { Xw_p = ld_xf<RC>(A, M.w_xw + (p + 1) * XWW); v_p = ld_sv<RA>(A, M.w_link + p * LW, LK_VC); }
    else { Xw_p = Xw_base; v_p = v_base; }
    RC qv = (fl & TDS_LF_FIXED) ? RC(0) : RC(A.at<float>(M.w_q, M.q_idx[i]));
    Xf<RC> Xp = jcalc<RC>(M, i, qv);
    Xf<RC> Xw = xf_mul(Xw_p, Xp);
    st_xf<RC>(A, M.w_xw + (i + 1) * XWW, Xw);
    Xf<RA> XpA; XpA.R = cvt<RA>(Xp.R); XpA.t = cvt<RA>(Xp.t);
    st_xf<RA>(A, M.w_link + i * LW, XpA);
    Sv<RA> v = xf_apply_motion(XpA, v_p);
    if (!(fl & TDS_LF_FIXED)) {
      RA qdv = RA(A.at<float>(M.w_qd, M.qd_idx[i]));
      Sv<RA> S = link_S<RA>(M, i);
      v.top = v.top + S.top * qdv;
      v.bot = v.bot + S.bot * qdv;
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
          const int w = M.w_con + n_active * CN_SIZE * (int)(sizeof(RC) / 4);
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
  Abi<RA> cA;
  Sv<RA> cP;
  Rbi<RC> cC;   // composite rigid-body inertia (CRBA) is carried in RC: see the precision note above
  if (any_contact)
    for (int k = 0; k < n * (n + 1) / 2; ++k) A.at<RC>(M.w_M, k) = RC(0);
  for (int i = n_links - 1; i >= 0; --i) {
    const int p = M.parent[i];
    const int fl = M.flags[i];
    const int wl = M.w_link + i * LW;
    const Xf<RA> Xp = ld_xf<RA>(A, wl);
    const Sv<RA> v = ld_sv<RA>(A, wl, LK_VC);
    const Rbi<RA> rb = model_rbi<RA>(M.rbi[i]);
    Rbi<RC> Ic = model_rbi<RC>(M.rbi[i]);
    Abi<RA> Ai = abi_from_rbi(rb);
    Sv<RA> pA = cross_mf(v, rbi_mul(rb, v));        // kinematics.hpp:132
    if (fl & TDS_LF_CHILD_ADJ) { abi_add(Ai, cA); pA = pA + cP; rbi_add(Ic, cC); }
    if (M.acc_slot[i] >= 0) {
      Abi<RA> sa; Sv<RA> sp;
      const int ws = M.w_acc + M.acc_slot[i] * M.acc_words;
      acc_load<RA>(A, ws, sa, sp);
      abi_add(Ai, sa); pA = pA + sp; rbi_add(Ic, acc_load_rbi<RC>(A, ws + M.acc_ic_word));
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
      const RA qdv = RA(A.at<float>(M.w_qd, qdi));
      Sv<RA> vJ; vJ.top = S.top * qdv; vJ.bot = S.bot * qdv;
      const Sv<RA> c = cross_mm(v, vJ);               // kinematics.hpp:96-97
      const Sv<RA> U = abi_mul(Ai, S);                // forward_dynamics.hpp:111
      const RA D = dot(S, U);
      const RA invD = RA(1) / D;
      RA tau = RA(A.at<float>(M.w_tau, qdi));
      tau -= RA(M.stiffness[i]) * RA(A.at<float>(M.w_q, M.q_idx[i]));
      tau -= RA(M.damping[i]) * qdv;
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
        const Sv<RC> Sc = link_S<RC>(M, i);
        Sv<RC> F = rbi_mul(Ic, Sc);
        A.at<RC>(M.w_M, tri(qdi, qdi)) = dot(Sc, F);
        int j = i;
        Xf<RC> Xj; Xj.R = cvt<RC>(Xp.R); Xj.t = cvt<RC>(Xp.t);
        while (true) {
          F = xf_apply_force(Xj, F);
          j = M.parent[j];
          if (j < 0) break;
          if (!(M.flags[j] & TDS_LF_FIXED)) A.at<RC>(M.w_M, tri(qdi, M.qd_idx[j])) = dot(F, link_S<RC>(M, j));
          const Xf<RA> Xa = ld_xf<RA>(A, M.w_link + j * LW);
          Xj.R = cvt<RC>(Xa.R); Xj.t = cvt<RC>(Xa.t);
        }
        if (M.floating) {
          A.at<RC>(M.w_M, tri(qdi, 0)) = F.top.x; A.at<RC>(M.w_M, tri(qdi, 1)) = F.top.y; A.at<RC>(M.w_M, tri(qdi, 2)) = F.top.z;
          A.at<RC>(M.w_M, tri(qdi, 3)) = F.bot.x; A.at<RC>(M.w_M, tri(qdi, 4)) = F.bot.y; A.at<RC>(M.w_M, tri(qdi, 5)) = F.bot.z;
        }
      }
    }
    // propagate to the parent: register carry along chains, accumulator at branch points
    const Abi<RA> dA = xt_abi_x(Xp, Ia);               // :187-189
    const Sv<RA> dP = xf_apply_force(Xp, pa);          // :181
    Rbi<RC> dC = Ic;
    if (any_contact) { Xf<RC> Xc; Xc.R = cvt<RC>(Xp.R); Xc.t = cvt<RC>(Xp.t); dC = xt_rbi_x(Xc, Ic); }  // mass_matrix.hpp:45-46
    if (fl & TDS_LF_PARENT_ADJ) { cA = dA; cP = dP; cC = dC; }
    else {
      const int slot = (p >= 0) ? M.acc_slot[p] : M.base_acc;
      if (slot >= 0) {
        const int w = M.w_acc + slot * M.acc_words;
        acc_add_abi<RA>(A, w, dA, dP);
        acc_add_rbi<RC>(A, w + M.acc_ic_word, dC);
      }
    }
  }

  TDS_PHASE();  // 4: pass 2 (ABA + CRBA) done
  // ---- base acceleration (forward_dynamics.hpp:218-243) -----------------------------------------
  Sv<RA> a_prev;
  Sv<RC> base_acc;
  if (M.floating) {
    Rbi<RC> Ib = model_rbi<RC>(M.base_rbi);
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
      abi_add(Ab, sa); pb = pb + sp; rbi_add(Ib, acc_load_rbi<RC>(A, ws + M.acc_ic_word));
    }
    if (any_contact) {  // mass_matrix.hpp:114-120: base block = composite inertia
      const RC z = RC(0);
      A.at<RC>(M.w_M, tri(0, 0)) = Ib.I.xx; A.at<RC>(M.w_M, tri(1, 0)) = Ib.I.xy; A.at<RC>(M.w_M, tri(1, 1)) = Ib.I.yy;
      A.at<RC>(M.w_M, tri(2, 0)) = Ib.I.xz; A.at<RC>(M.w_M, tri(2, 1)) = Ib.I.yz; A.at<RC>(M.w_M, tri(2, 2)) = Ib.I.zz;
      // rows 3..5: [H^T | M], H = h x
      A.at<RC>(M.w_M, tri(3, 0)) = z;            A.at<RC>(M.w_M, tri(3, 1)) = RC(Ib.h.z);  A.at<RC>(M.w_M, tri(3, 2)) = RC(-Ib.h.y);
      A.at<RC>(M.w_M, tri(4, 0)) = RC(-Ib.h.z);  A.at<RC>(M.w_M, tri(4, 1)) = z;           A.at<RC>(M.w_M, tri(4, 2)) = RC(Ib.h.x);
      A.at<RC>(M.w_M, tri(5, 0)) = RC(Ib.h.y);   A.at<RC>(M.w_M, tri(5, 1)) = RC(-Ib.h.x); A.at<RC>(M.w_M, tri(5, 2)) = z;
      A.at<RC>(M.w_M, tri(3, 3)) = RC(Ib.m); A.at<RC>(M.w_M, tri(4, 3)) = z; A.at<RC>(M.w_M, tri(4, 4)) = RC(Ib.m);
      A.at<RC>(M.w_M, tri(5, 3)) = z; A.at<RC>(M.w_M, tri(5, 4)) = z; A.at<RC>(M.w_M, tri(5, 5)) = RC(Ib.m);
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
      else A.at<float>(M.w_qd, qdi) = (float)(RA(A.at<float>(M.w_qd, qdi)) + qdd * dtA);
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
      else A.at<float>(M.w_qd, k) = (float)(RC(A.at<float>(M.w_qd, k)) + qb[k] * RC(P.dt));
    }
  }
  TDS_PHASE();  // 6: pass 3 done
  if (mode == MODE_FD) return;

  // ---- contact solve ------------------------------------------------------------------------------
  if (mode == MODE_FULL && any_contact) {
    const int RCW = (int)(sizeof(RC) / 4);
    // Cholesky M = L L^T in place (lower triangle).  The reference inverts M
    // (tiny_matrix_x.h:240-344); only products with M^-1 are needed.
    for (int j = 0; j < n; ++j) {
      RC d = A.at<RC>(M.w_M, tri(j, j));
      for (int k = 0; k < j; ++k) { const RC l = A.at<RC>(M.w_M, tri(j, k)); d -= l * l; }
      const RC ljj = sqrt_t(d);
      const RC inv = RC(1) / ljj;
      A.at<RC>(M.w_M, tri(j, j)) = ljj;
      for (int i = j + 1; i < n; ++i) {
        RC s = A.at<RC>(M.w_M, tri(i, j));
        for (int k = 0; k < j; ++k) s -= A.at<RC>(M.w_M, tri(i, k)) * A.at<RC>(M.w_M, tri(j, k));
        A.at<RC>(M.w_M, tri(i, j)) = s * inv;
      }
    }
    TDS_PHASE();  // 7: Cholesky done
    const int max_active = __reduce_max_sync(0xffffffffu, n_active);
    const V3<RC> nb = v3<RC>(RC(-M.plane_n[0]), RC(-M.plane_n[1]), RC(-M.plane_n[2]));  // world_normal_on_b
    const V3<RC> f1 = v3<RC>(RC(M.fr1[0]), RC(M.fr1[1]), RC(M.fr1[2]));
    const V3<RC> f2 = v3<RC>(RC(M.fr2[0]), RC(M.fr2[1]), RC(M.fr2[2]));
    for (int c = 0; c < max_active; ++c) {
      if (c < n_active) {
        const int wc = M.w_con + c * CN_SIZE * RCW;
        const int wy = M.w_Y + c * 3 * n * RCW;
        const V3<RC> pb = ld_v3<RC>(A, wc, CN_PB);
        const RC dist = A.at<RC>(wc, CN_DIST);
        const int L = (int)A.at<RC>(wc, CN_LINK);
        for (int k = 0; k < 3 * n; ++k) A.at<RC>(wy, k) = RC(0);
        V3<RC> vel = v3<RC>(RC(0), RC(0), RC(0));   // vel_b = J qd
        if (M.floating) {  // jacobian.hpp:39-58
          const V3<RC> r = pb - Xw_base.t;
          // J[:,0:3] = cross_matrix(r)^T, J[:,3:6] = 1
          const V3<RC> c0 = v3<RC>(RC(0), -r.z, r.y), c1 = v3<RC>(r.z, RC(0), -r.x), c2 = v3<RC>(-r.y, r.x, RC(0));
          const V3<RC> e0 = v3<RC>(RC(1), RC(0), RC(0)), e1 = v3<RC>(RC(0), RC(1), RC(0)), e2 = v3<RC>(RC(0), RC(0), RC(1));
          const V3<RC> cols[6] = {c0, c1, c2, e0, e1, e2};
#pragma unroll
          for (int k = 0; k < 6; ++k) {
            A.at<RC>(wy, k) = dot(nb, cols[k]);
            A.at<RC>(wy, n + k) = dot(f1, cols[k]);
            A.at<RC>(wy, 2 * n + k) = dot(f2, cols[k]);
            vel = vel + cols[k] * RC(A.at<float>(M.w_qd, k));
          }
        }
        for (int j = L; j >= 0; j = M.parent[j]) {  // jacobian.hpp:63-80
          if (M.flags[j] & TDS_LF_FIXED) continue;
          const Xf<RC> Xw = ld_xf<RC>(A, M.w_xw + (j + 1) * XWW);
          const Sv<RC> S = link_S<RC>(M, j);
          const V3<RC> wv = mul(Xw.R, S.top);
          const V3<RC> col = mul(Xw.R, S.bot) + cross(wv, pb - Xw.t);
          const int qj = M.qd_idx[j];
          A.at<RC>(wy, qj) = dot(nb, col);
          A.at<RC>(wy, n + qj) = dot(f1, col);
          A.at<RC>(wy, 2 * n + qj) = dot(f2, col);
          vel = vel + col * RC(A.at<float>(M.w_qd, qj));
        }
        // rel_vel = vel_a - vel_b = -vel ; mb_constraint_solver.hpp:299-345
        const RC nrv = -dot(nb, vel);
        A.at<RC>(wc, CN_B + 0) = -(RC(1) + RC(P.restitution)) * nrv - RC(P.erp) * dist / RC(P.dt);
        A.at<RC>(wc, CN_B + 1) = dot(f1, vel);
        A.at<RC>(wc, CN_B + 2) = dot(f2, vel);
        A.at<RC>(wc, CN_X + 0) = RC(0); A.at<RC>(wc, CN_X + 1) = RC(0); A.at<RC>(wc, CN_X + 2) = RC(0);
        // Y rows = L^-1 * Jc rows (forward substitution, three right-hand sides share the loads of L)
        for (int i = 0; i < n; ++i) {
          RC s0 = A.at<RC>(wy, i), s1 = A.at<RC>(wy, n + i), s2 = A.at<RC>(wy, 2 * n + i);
          for (int k = 0; k < i; ++k) {
            const RC l = A.at<RC>(M.w_M, tri(i, k));
            s0 -= l * A.at<RC>(wy, k); s1 -= l * A.at<RC>(wy, n + k); s2 -= l * A.at<RC>(wy, 2 * n + k);
          }
          const RC inv = RC(1) / A.at<RC>(M.w_M, tri(i, i));
          A.at<RC>(wy, i) = s0 * inv; A.at<RC>(wy, n + i) = s1 * inv; A.at<RC>(wy, 2 * n + i) = s2 * inv;
        }
      }
    }
    TDS_PHASE();  // 8: Jacobians + Y done
    // matrix-free projected Gauss-Seidel on w = Y p; row order normals | friction-1 | friction-2
    // (solve_pgs, mb_constraint_solver.hpp:101-142; bounds :417-436)
    for (int k = 0; k < n; ++k) A.at<RC>(M.w_w, k) = RC(0);
    for (int it = 0; it < P.pgs_iterations; ++it) {
      for (int blk = 0; blk < 3; ++blk) {
        for (int c = 0; c < max_active; ++c) {
          if (c < n_active) {
            const int wc = M.w_con + c * CN_SIZE * RCW;
            const int wy = M.w_Y + (c * 3 + blk) * n * RCW;
            RC yy = RC(0), yw = RC(0);
            for (int k = 0; k < n; ++k) { const RC y = A.at<RC>(wy, k); yy += y * y; yw += y * A.at<RC>(M.w_w, k); }
            const RC x_old = A.at<RC>(wc, CN_X + blk);
            RC x = (A.at<RC>(wc, CN_B + blk) - yw + yy * x_old) / (yy + RC(P.cfm));
            if (blk == 0) {
              x = x < RC(0) ? RC(0) : x;
              x = x > RC(100000) ? RC(100000) : x;
            } else {
              RC s = A.at<RC>(wc, CN_X + 0);
              s = s < RC(0) ? RC(0) : s;
              const RC lim = RC(P.friction) * s;
              x = x < -lim ? -lim : x;
              x = x > lim ? lim : x;
            }
            A.at<RC>(wc, CN_X + blk) = x;
            const RC dx = x - x_old;
            for (int k = 0; k < n; ++k) A.at<RC>(M.w_w, k) += dx * A.at<RC>(wy, k);
          }
        }
      }
    }
    TDS_PHASE();  // 9: PGS done
    // qd_b -= M^-1 Jc^T p = L^-T w   (mb_constraint_solver.hpp:476-497)
    for (int i = n - 1; i >= 0; --i) {
      RC s = A.at<RC>(M.w_w, i);
      for (int k = i + 1; k < n; ++k) s -= A.at<RC>(M.w_M, tri(k, i)) * A.at<RC>(M.w_w, k);
      s = s / A.at<RC>(M.w_M, tri(i, i));
      A.at<RC>(M.w_w, i) = s;
      if (n_active > 0) A.at<float>(M.w_qd, i) = (float)(RC(A.at<float>(M.w_qd, i)) - s);
    }
  }

  TDS_PHASE();  // 10: impulses applied
  // ---- integrate_euler with qdd = 0 (integrator.hpp:10-133) ---------------------------------------
  RC up_z = RC(1);
  if (M.floating) {
    const RC h = RC(0.5) * RC(P.dt);
    RC qx = RC(A.at<float>(M.w_q, 0)), qy = RC(A.at<float>(M.w_q, 1)), qz = RC(A.at<float>(M.w_q, 2)), qw = RC(A.at<float>(M.w_q, 3));
    const RC w0 = RC(A.at<float>(M.w_qd, 0)), w1 = RC(A.at<float>(M.w_qd, 1)), w2 = RC(A.at<float>(M.w_qd, 2));
    const RC dw = (-qx * w0 - qy * w1 - qz * w2) * h;
    const RC dx = (qw * w0 + qz * w1 - qy * w2) * h;
    const RC dy = (qw * w1 + qx * w2 - qz * w0) * h;
    const RC dz = (qw * w2 + qy * w0 - qx * w1) * h;
    qx += dx; qy += dy; qz += dz; qw += dw;
    const RC len = sqrt_t(qx * qx + qy * qy + qz * qz + qw * qw);
    qx /= len; qy /= len; qz /= len; qw /= len;
    A.at<float>(M.w_q, 0) = (float)qx; A.at<float>(M.w_q, 1) = (float)qy; A.at<float>(M.w_q, 2) = (float)qz; A.at<float>(M.w_q, 3) = (float)qw;
    for (int k = 0; k < 3; ++k)
      A.at<float>(M.w_q, 4 + k) = (float)(RC(A.at<float>(M.w_q, 4 + k)) + RC(A.at<float>(M.w_qd, 3 + k)) * RC(P.dt));
    up_z = RC(1) - RC(2) * (qx * qx + qy * qy) / (qx * qx + qy * qy + qz * qz + qw * qw);
  }
  for (int i = 0; i < n_links; ++i) {
    if (M.flags[i] & TDS_LF_FIXED) continue;
    const int qi = M.q_idx[i];
    A.at<float>(M.w_q, qi) = (float)(RC(A.at<float>(M.w_q, qi)) + RC(A.at<float>(M.w_qd, M.qd_idx[i])) * RC(P.dt));
  }

  TDS_PHASE();  // 11: integrated
  // ---- write back -----------------------------------------------------------------------------------
  if (live) {
    for (int k = 0; k < M.n_q; ++k) io.q_out[(size_t)k * ns + e] = A.at<float>(M.w_q, k);
    for (int k = 0; k < n; ++k) io.qd_out[(size_t)k * ns + e] = A.at<float>(M.w_qd, k);
    if (io.reward && E.reward_kind == 1) {
      // laikago_environment2.h:130-171, fixed-base emulation: x = q0, z = q2, rpy = q3..5
      const float x = A.at<float>(M.w_q, 0), z = A.at<float>(M.w_q, 2);
      const float roll = A.at<float>(M.w_q, 3), pitch = A.at<float>(M.w_q, 4);
      // up.z of quat_to_matrix(quat_from_euler_rpy(rpy)) = cos(roll) cos(pitch)
      const float upz = cosf(roll) * cosf(pitch);
      const bool done = (upz < 0.6f) || (z < 0.2f);
      io.reward[e] = done ? 0.f : x;
      if (io.done) io.done[e] = done ? 1.f : 0.f;
    } else if (io.reward && E.reward_kind == 2) {
      const float x = A.at<float>(M.w_q, 4), z = A.at<float>(M.w_q, 6);
      const bool done = ((float)up_z < 0.6f) || (z < 0.2f);
      io.reward[e] = done ? 0.f : x;
      if (io.done) io.done[e] = done ? 1.f : 0.f;
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
#define TDS_LAUNCH(RA, RC, SM)                                                                          \
  do {                                                                                                  \
    auto k = tds_step_kernel<RA, RC, SM>;                                                               \
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
  if (precision == 0) {        // mixed: fp32 ABA, fp64 kinematics + contact
    if (use_smem) TDS_LAUNCH(float, double, true); else TDS_LAUNCH(float, double, false);
  } else if (precision == 1) { // all fp64
    if (use_smem) TDS_LAUNCH(double, double, true); else TDS_LAUNCH(double, double, false);
  } else {                     // all fp32
    if (use_smem) TDS_LAUNCH(float, float, true); else TDS_LAUNCH(float, float, false);
  }
#undef TDS_LAUNCH
  return (int)err;
}
