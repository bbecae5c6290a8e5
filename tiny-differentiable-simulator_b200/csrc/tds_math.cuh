// Device-side spatial algebra for the batched rigid-body step kernels (sm_100a).
//
// Conventions follow the reference's right-associative transforms
// (src/math/transform.hpp:6-8,123-131): a Transform (R, t) stores the child->parent rotation
// and the child origin expressed in the parent frame; X_world = parent_X_world * X_parent.
// Everything is held in scalar struct members so it lives in registers (no local arrays).
#pragma once
#include <cuda_runtime.h>

namespace tds {

template <typename T> struct V3 { T x, y, z; };
template <typename T> struct M3 { T xx, xy, xz, yx, yy, yz, zx, zy, zz; };   // row-major
template <typename T> struct S3 { T xx, xy, xz, yy, yz, zz; };               // symmetric 3x3
template <typename T> struct Xf { M3<T> R; V3<T> t; };                       // tds::Transform
template <typename T> struct Sv { V3<T> top, bot; };                         // Motion/ForceVector
// Articulated-body inertia [I H; H^T M] with I, M symmetric (src/math/inertia.hpp:89-95).
template <typename T> struct Abi { S3<T> I; M3<T> H; S3<T> M; };
// Rigid-body (composite) inertia about the link origin: mass, h = m*com, I (symmetric).
template <typename T> struct Rbi { T m; V3<T> h; S3<T> I; };

#define TDS_D __device__ __forceinline__

template <typename T> TDS_D V3<T> v3(T x, T y, T z) { V3<T> r; r.x = x; r.y = y; r.z = z; return r; }
template <typename T> TDS_D V3<T> operator+(V3<T> a, V3<T> b) { return v3<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T> TDS_D V3<T> operator-(V3<T> a, V3<T> b) { return v3<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T> TDS_D V3<T> operator*(V3<T> a, T s) { return v3<T>(a.x * s, a.y * s, a.z * s); }
template <typename T> TDS_D T dot(V3<T> a, V3<T> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename T> TDS_D V3<T> cross(V3<T> a, V3<T> b) {
  return v3<T>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
template <typename T> TDS_D V3<T> mul(const M3<T>& A, V3<T> v) {
  return v3<T>(A.xx * v.x + A.xy * v.y + A.xz * v.z, A.yx * v.x + A.yy * v.y + A.yz * v.z,
               A.zx * v.x + A.zy * v.y + A.zz * v.z);
}
template <typename T> TDS_D V3<T> mulT(const M3<T>& A, V3<T> v) {  // A^T v
  return v3<T>(A.xx * v.x + A.yx * v.y + A.zx * v.z, A.xy * v.x + A.yy * v.y + A.zy * v.z,
               A.xz * v.x + A.yz * v.y + A.zz * v.z);
}
template <typename T> TDS_D V3<T> mul(const S3<T>& A, V3<T> v) {
  return v3<T>(A.xx * v.x + A.xy * v.y + A.xz * v.z, A.xy * v.x + A.yy * v.y + A.yz * v.z,
               A.xz * v.x + A.yz * v.y + A.zz * v.z);
}
template <typename T> TDS_D M3<T> mul(const M3<T>& A, const M3<T>& B) {
  M3<T> C;
  C.xx = A.xx * B.xx + A.xy * B.yx + A.xz * B.zx; C.xy = A.xx * B.xy + A.xy * B.yy + A.xz * B.zy; C.xz = A.xx * B.xz + A.xy * B.yz + A.xz * B.zz;
  C.yx = A.yx * B.xx + A.yy * B.yx + A.yz * B.zx; C.yy = A.yx * B.xy + A.yy * B.yy + A.yz * B.zy; C.yz = A.yx * B.xz + A.yy * B.yz + A.yz * B.zz;
  C.zx = A.zx * B.xx + A.zy * B.yx + A.zz * B.zx; C.zy = A.zx * B.xy + A.zy * B.yy + A.zz * B.zy; C.zz = A.zx * B.xz + A.zy * B.yz + A.zz * B.zz;
  return C;
}
template <typename T> TDS_D M3<T> m3_identity() {
  M3<T> R; R.xx = T(1); R.xy = T(0); R.xz = T(0); R.yx = T(0); R.yy = T(1); R.yz = T(0); R.zx = T(0); R.zy = T(0); R.zz = T(1);
  return R;
}
template <typename TO, typename TI> TDS_D M3<TO> cvt(const M3<TI>& a) {
  M3<TO> r; r.xx = TO(a.xx); r.xy = TO(a.xy); r.xz = TO(a.xz); r.yx = TO(a.yx); r.yy = TO(a.yy); r.yz = TO(a.yz); r.zx = TO(a.zx); r.zy = TO(a.zy); r.zz = TO(a.zz);
  return r;
}
template <typename TO, typename TI> TDS_D V3<TO> cvt(const V3<TI>& a) { return v3<TO>(TO(a.x), TO(a.y), TO(a.z)); }

// R * S * R^T for symmetric S (result symmetric).
template <typename T> TDS_D S3<T> rot_sym(const M3<T>& R, const S3<T>& S) {
  // T = R * S
  T txx = R.xx * S.xx + R.xy * S.xy + R.xz * S.xz, txy = R.xx * S.xy + R.xy * S.yy + R.xz * S.yz, txz = R.xx * S.xz + R.xy * S.yz + R.xz * S.zz;
  T tyx = R.yx * S.xx + R.yy * S.xy + R.yz * S.xz, tyy = R.yx * S.xy + R.yy * S.yy + R.yz * S.yz, tyz = R.yx * S.xz + R.yy * S.yz + R.yz * S.zz;
  T tzx = R.zx * S.xx + R.zy * S.xy + R.zz * S.xz, tzy = R.zx * S.xy + R.zy * S.yy + R.zz * S.yz, tzz = R.zx * S.xz + R.zy * S.yz + R.zz * S.zz;
  S3<T> o;
  o.xx = txx * R.xx + txy * R.xy + txz * R.xz;
  o.xy = txx * R.yx + txy * R.yy + txz * R.yz;
  o.xz = txx * R.zx + txy * R.zy + txz * R.zz;
  o.yy = tyx * R.yx + tyy * R.yy + tyz * R.yz;
  o.yz = tyx * R.zx + tyy * R.zy + tyz * R.zz;
  o.zz = tzx * R.zx + tzy * R.zy + tzz * R.zz;
  return o;
}
// R * A * R^T for general A.
template <typename T> TDS_D M3<T> rot_gen(const M3<T>& R, const M3<T>& A) {
  M3<T> Tm = mul(R, A);
  M3<T> o;
  o.xx = Tm.xx * R.xx + Tm.xy * R.xy + Tm.xz * R.xz; o.xy = Tm.xx * R.yx + Tm.xy * R.yy + Tm.xz * R.yz; o.xz = Tm.xx * R.zx + Tm.xy * R.zy + Tm.xz * R.zz;
  o.yx = Tm.yx * R.xx + Tm.yy * R.xy + Tm.yz * R.xz; o.yy = Tm.yx * R.yx + Tm.yy * R.yy + Tm.yz * R.yz; o.yz = Tm.yx * R.zx + Tm.yy * R.zy + Tm.yz * R.zz;
  o.zx = Tm.zx * R.xx + Tm.zy * R.xy + Tm.zz * R.xz; o.zy = Tm.zx * R.yx + Tm.zy * R.yy + Tm.zz * R.yz; o.zz = Tm.zx * R.zx + Tm.zy * R.zy + Tm.zz * R.zz;
  return o;
}

// TinyMatrix3x3::setRotation, src/math/tiny/tiny_matrix3x3.h:315-340 (right-associative build).
template <typename T> TDS_D M3<T> quat_to_matrix(T x, T y, T z, T w) {
  T d = x * x + y * y + z * z + w * w;
  T s = T(2) / d;
  T xs = x * s, ys = y * s, zs = z * s;
  T wx = w * xs, wy = w * ys, wz = w * zs;
  T xx = x * xs, xy = x * ys, xz = x * zs;
  T yy = y * ys, yz = y * zs, zz = z * zs;
  M3<T> m;
  m.xx = T(1) - (yy + zz); m.xy = xy - wz; m.xz = xz + wy;
  m.yx = xy + wz; m.yy = T(1) - (xx + zz); m.yz = yz - wx;
  m.zx = xz - wy; m.zy = yz + wx; m.zz = T(1) - (xx + yy);
  return m;
}

// Transform::operator*, src/math/transform.hpp:123-131.
template <typename T> TDS_D Xf<T> xf_mul(const Xf<T>& a, const Xf<T>& b) {
  Xf<T> r;
  r.t = a.t + mul(a.R, b.t);
  r.R = mul(a.R, b.R);
  return r;
}
// Transform::apply(MotionVector): (R^T w, R^T (v - t x w)), src/math/transform.hpp:210-226.
template <typename T> TDS_D Sv<T> xf_apply_motion(const Xf<T>& X, const Sv<T>& m) {
  Sv<T> r;
  r.top = mulT(X.R, m.top);
  r.bot = mulT(X.R, m.bot - cross(X.t, m.top));
  return r;
}
// Transform::apply(ForceVector) = X^T F: (R n + t x (R f), R f), src/math/transform.hpp:249-262.
template <typename T> TDS_D Sv<T> xf_apply_force(const Xf<T>& X, const Sv<T>& f) {
  Sv<T> r;
  r.bot = mul(X.R, f.bot);
  r.top = mul(X.R, f.top) + cross(X.t, r.bot);
  return r;
}
template <typename T> TDS_D Sv<T> operator+(const Sv<T>& a, const Sv<T>& b) { Sv<T> r; r.top = a.top + b.top; r.bot = a.bot + b.bot; return r; }
template <typename T> TDS_D T dot(const Sv<T>& a, const Sv<T>& b) { return dot(a.top, b.top) + dot(a.bot, b.bot); }
// motion x motion, src/math/tiny/tiny_algebra.hpp:101-105
template <typename T> TDS_D Sv<T> cross_mm(const Sv<T>& a, const Sv<T>& b) {
  Sv<T> r; r.top = cross(a.top, b.top); r.bot = cross(a.top, b.bot) + cross(a.bot, b.top); return r;
}
// motion x* force, src/math/tiny/tiny_algebra.hpp:112-115
template <typename T> TDS_D Sv<T> cross_mf(const Sv<T>& a, const Sv<T>& b) {
  Sv<T> r; r.top = cross(a.top, b.top) + cross(a.bot, b.bot); r.bot = cross(a.top, b.bot); return r;
}

// Ia * v = (I w + H v, M v + H^T w), src/math/inertia.hpp:205-210.
template <typename T> TDS_D Sv<T> abi_mul(const Abi<T>& A, const Sv<T>& v) {
  Sv<T> r;
  r.top = mul(A.I, v.top) + mul(A.H, v.bot);
  r.bot = mul(A.M, v.bot) + mulT(A.H, v.top);
  return r;
}
// Rigid-body inertia as an articulated inertia (H = h x, M = m 1), src/math/inertia.hpp:121-130.
template <typename T> TDS_D Abi<T> abi_from_rbi(const Rbi<T>& r) {
  Abi<T> A;
  A.I = r.I;
  A.H.xx = T(0); A.H.xy = -r.h.z; A.H.xz = r.h.y;
  A.H.yx = r.h.z; A.H.yy = T(0); A.H.yz = -r.h.x;
  A.H.zx = -r.h.y; A.H.zy = r.h.x; A.H.zz = T(0);
  A.M.xx = r.m; A.M.xy = T(0); A.M.xz = T(0); A.M.yy = r.m; A.M.yz = T(0); A.M.zz = r.m;
  return A;
}
// Rbi * motion: (I w + h x v, m v - h x w)
template <typename T> TDS_D Sv<T> rbi_mul(const Rbi<T>& r, const Sv<T>& v) {
  Sv<T> o;
  o.top = mul(r.I, v.top) + cross(r.h, v.bot);
  o.bot = v.bot * r.m - cross(r.h, v.top);
  return o;
}
template <typename T> TDS_D void abi_add(Abi<T>& a, const Abi<T>& b) {
  a.I.xx += b.I.xx; a.I.xy += b.I.xy; a.I.xz += b.I.xz; a.I.yy += b.I.yy; a.I.yz += b.I.yz; a.I.zz += b.I.zz;
  a.H.xx += b.H.xx; a.H.xy += b.H.xy; a.H.xz += b.H.xz; a.H.yx += b.H.yx; a.H.yy += b.H.yy; a.H.yz += b.H.yz; a.H.zx += b.H.zx; a.H.zy += b.H.zy; a.H.zz += b.H.zz;
  a.M.xx += b.M.xx; a.M.xy += b.M.xy; a.M.xz += b.M.xz; a.M.yy += b.M.yy; a.M.yz += b.M.yz; a.M.zz += b.M.zz;
}
template <typename T> TDS_D void rbi_add(Rbi<T>& a, const Rbi<T>& b) {
  a.m += b.m; a.h = a.h + b.h;
  a.I.xx += b.I.xx; a.I.xy += b.I.xy; a.I.xz += b.I.xz; a.I.yy += b.I.yy; a.I.yz += b.I.yz; a.I.zz += b.I.zz;
}

// X^T * Ia * X for X = (R, t): the (0,0), (0,3), (3,3) blocks of the reference's dense 6x6 triple
// product (src/dynamics/forward_dynamics.hpp:187-189), evaluated in block form:
//   M' = R M R^T,  H' = R H R^T + t x M',  I' = R I R^T + B + B^T - (t x M') tx,  B = t x (R H R^T)^T
template <typename T> TDS_D Abi<T> xt_abi_x(const Xf<T>& X, const Abi<T>& A) {
  Abi<T> o;
  o.M = rot_sym(X.R, A.M);
  M3<T> Hp = rot_gen(X.R, A.H);
  const T tx = X.t.x, ty = X.t.y, tz = X.t.z;
  // K = tx * M'   (tx = cross matrix of t)
  M3<T> K;
  K.xx = -tz * o.M.xy + ty * o.M.xz; K.xy = -tz * o.M.yy + ty * o.M.yz; K.xz = -tz * o.M.yz + ty * o.M.zz;
  K.yx = tz * o.M.xx - tx * o.M.xz;  K.yy = tz * o.M.xy - tx * o.M.yz;  K.yz = tz * o.M.xz - tx * o.M.zz;
  K.zx = -ty * o.M.xx + tx * o.M.xy; K.zy = -ty * o.M.xy + tx * o.M.yy; K.zz = -ty * o.M.xz + tx * o.M.yz;
  o.H.xx = Hp.xx + K.xx; o.H.xy = Hp.xy + K.xy; o.H.xz = Hp.xz + K.xz;
  o.H.yx = Hp.yx + K.yx; o.H.yy = Hp.yy + K.yy; o.H.yz = Hp.yz + K.yz;
  o.H.zx = Hp.zx + K.zx; o.H.zy = Hp.zy + K.zy; o.H.zz = Hp.zz + K.zz;
  // B = tx * Hp^T : B_ij = sum_k tx_ik Hp_jk
  T bxx = -tz * Hp.xy + ty * Hp.xz, bxy = -tz * Hp.yy + ty * Hp.yz, bxz = -tz * Hp.zy + ty * Hp.zz;
  T byx = tz * Hp.xx - tx * Hp.xz,  byy = tz * Hp.yx - tx * Hp.yz,  byz = tz * Hp.zx - tx * Hp.zz;
  T bzx = -ty * Hp.xx + tx * Hp.xy, bzy = -ty * Hp.yx + tx * Hp.yy, bzz = -ty * Hp.zx + tx * Hp.zy;
  // C = K * tx : C_ij = sum_k K_ik tx_kj ; tx columns: col x = (0, tz, -ty), col y = (-tz, 0, tx), col z = (ty, -tx, 0)
  T cxx = K.xy * tz - K.xz * ty, cxy = -K.xx * tz + K.xz * tx, cxz = K.xx * ty - K.xy * tx;
  T cyy = -K.yx * tz + K.yz * tx, cyz = K.yx * ty - K.yy * tx;
  T czz = K.zx * ty - K.zy * tx;
  S3<T> RI = rot_sym(X.R, A.I);
  o.I.xx = RI.xx + (bxx + bxx) - cxx;
  o.I.xy = RI.xy + (bxy + byx) - cxy;
  o.I.xz = RI.xz + (bxz + bzx) - cxz;
  o.I.yy = RI.yy + (byy + byy) - cyy;
  o.I.yz = RI.yz + (byz + bzy) - cyz;
  o.I.zz = RI.zz + (bzz + bzz) - czz;
  return o;
}

// X^T * Ic * X for a rigid-body (composite) inertia: m' = m, h' = R h + m t,
// I' = R I R^T - tx (Rh)x - (h')x tx   (src/math/transform.hpp:409-428, apply_transpose(rbi)).
template <typename T> TDS_D Rbi<T> xt_rbi_x(const Xf<T>& X, const Rbi<T>& r) {
  Rbi<T> o;
  o.m = r.m;
  V3<T> Rh = mul(X.R, r.h);
  o.h = Rh + X.t * r.m;
  S3<T> RI = rot_sym(X.R, r.I);
  // -(a x)(b x) = (a.b) 1 - b a^T ;  -tx (Rh)x - (h')x tx = (t.Rh)1 - Rh t^T + (h'.t)1 - t h'^T
  const V3<T> t = X.t;
  T d = dot(t, Rh) + dot(o.h, t);
  o.I.xx = RI.xx + d - Rh.x * t.x - t.x * o.h.x;
  o.I.yy = RI.yy + d - Rh.y * t.y - t.y * o.h.y;
  o.I.zz = RI.zz + d - Rh.z * t.z - t.z * o.h.z;
  o.I.xy = RI.xy - Rh.x * t.y - t.x * o.h.y;
  o.I.xz = RI.xz - Rh.x * t.z - t.x * o.h.z;
  o.I.yz = RI.yz - Rh.y * t.z - t.y * o.h.z;
  return o;
}

TDS_D void sincos_t(float a, float* s, float* c) { sincosf(a, s, c); }
// fp64 sine / cosine for joint angles (|a| up to a few thousand radians): Cody-Waite reduction by pi/2 with a
// three-part constant, degree-15/14 Taylor kernels on [-pi/4, pi/4] (truncation < 5e-15).  About 30 fp64
// instructions; the libdevice sincos() with its huge-argument path costs 3-4x that on a lone warp.
TDS_D void sincos_t(double a, double* s, double* c) {
  const double kq = rint(a * 0.63661977236758134308);      // 2/pi
  double r = fma(kq, -1.57079632673412561417e+00, a);       // pi/2 split in three parts
  r = fma(kq, -6.07710050650619224932e-11, r);
  r = fma(kq, -2.02226624879595063154e-21, r);
  const double r2 = r * r;
  double ps = 1.0 / 1307674368000.0;
  ps = fma(ps, -r2, 1.0 / 6227020800.0);
  ps = fma(ps, -r2, 1.0 / 39916800.0);
  ps = fma(ps, -r2, 1.0 / 362880.0);
  ps = fma(ps, -r2, 1.0 / 5040.0);
  ps = fma(ps, -r2, 1.0 / 120.0);
  ps = fma(ps, -r2, 1.0 / 6.0);
  const double sn = fma(ps * r2, -r, r);
  double pc = 1.0 / 87178291200.0;
  pc = fma(pc, -r2, 1.0 / 479001600.0);
  pc = fma(pc, -r2, 1.0 / 3628800.0);
  pc = fma(pc, -r2, 1.0 / 40320.0);
  pc = fma(pc, -r2, 1.0 / 720.0);
  pc = fma(pc, -r2, 1.0 / 24.0);
  pc = fma(pc, -r2, 0.5);
  const double cs = fma(pc, -r2, 1.0);
  const int q = (int)kq & 3;
  const double ss = (q & 1) ? cs : sn, cc = (q & 1) ? sn : cs;
  *s = (q & 2) ? -ss : ss;
  *c = ((q + 1) & 2) ? -cc : cc;
}
// value / derivative accessors that also accept the forward-mode dual numbers of tds_dual.cuh
TDS_D double val_of(float a) { return (double)a; }
TDS_D double val_of(double a) { return a; }
TDS_D float min_t(float a, float b) { return fminf(a, b); }
TDS_D float max_t(float a, float b) { return fmaxf(a, b); }
TDS_D double min_t(double a, double b) { return fmin(a, b); }
TDS_D double max_t(double a, double b) { return fmax(a, b); }
TDS_D float pow_t(float a, float b) { return powf(a, b); }
TDS_D double pow_t(double a, double b) { return pow(a, b); }
TDS_D float atan2_t(float y, float x) { return atan2f(y, x); }
TDS_D double atan2_t(double y, double x) { return atan2(y, x); }
TDS_D float tanh_t(float a) { return tanhf(a); }
TDS_D double tanh_t(double a) { return tanh(a); }
// reciprocal / reciprocal square root of the fp32 solver quantities (1 / D of ABA, inverted Cholesky diagonals, 1 / A_ii):
// IEEE division and sqrt + division cost ~8-10 instructions each with a slow-path call; the single-instruction MUFU
// approximations (<= 1 ulp / 2 ulp, far inside the fp32 round-off the mixed arithmetic already carries) shrink the
// specialised kernel from 7528 to 6944 instructions - and the kernel is instruction-fetch bound, so 13.5 instead of
// 14.5 us / step at 4096 environments, 750 instead of 677 M env-steps/s at 65536 (profiles/r02_experiments.md #28).
// -DTDS_B200_EXACT_RCP restores the IEEE forms.  fp64 always keeps the exact forms.
#ifndef TDS_B200_EXACT_RCP
TDS_D float inv_t(float a) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
TDS_D float rsqrt_t(float a) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
#else
TDS_D float inv_t(float a) { return 1.0f / a; }
TDS_D float rsqrt_t(float a) { return 1.0f / sqrtf(a); }
#endif
TDS_D double inv_t(double a) { return 1.0 / a; }
TDS_D double rsqrt_t(double a) { return 1.0 / sqrt(a); }
TDS_D float sqrt_t(float a) { return sqrtf(a); }
TDS_D double sqrt_t(double a) { return sqrt(a); }

}  // namespace tds
