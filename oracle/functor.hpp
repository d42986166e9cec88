// oracle/functor.hpp — TEST INFRASTRUCTURE (CPU oracle). Not part of the product path.
//
// CPU restatement of the reference's residual functors, templated on the scalar type so that the
// same code runs on double (cost only) and on Dual<N> (autodiff Jacobian), exactly as Ceres calls
// the reference functors with double and Jet<double,N>.
//
//   reprojection_residual  <-  OurCostFunctionBundle::operator_function<T>
//                              /root/reference/src/BundleAdjustment/BundleAdjustment.h:120-195
//                              (config decode :26-102, arity selection :199-222)
//   micro_image_projection <-  CameraModel::projectPoint<T>      src/CameraModel.h:87-199
//   radial_shift           <-  CameraModel::radialDistortion<T>  src/CameraModel.h:205-223
//   tangential_shift       <-  CameraModel::tangentialDistortion<T> src/CameraModel.h:228-241
//   pose_matrix            <-  RigidBody::getTransformationMatrix<T> src/CameraModel.h:246-264
//                              (Eigen: AngleAxis*AngleAxis*AngleAxis is evaluated as a quaternion
//                               product and converted with Quaternion::toRotationMatrix)
//   distance_residual      <-  OurConstraintFunctionBundle::operator() BundleAdjustment.h:262-267
//
// The operation ORDER follows the reference line by line so that rounding agrees; the code itself
// is written from scratch. Pinned against the reference's own headers by oracle/ref_bridge.cpp.
#pragma once
#include <cstdint>

#include "dual.hpp"

namespace lfba_oracle {

// Decoded `config` bit mask (BundleAdjustment.h:28-79, CameraCalibration.cpp:778-814).
struct FunctorConfig {
  int n_radial = 0;       // config & 0x3
  bool tangential = false;  // 0x004
  bool refine_poses = false;  // 0x100 (and no fixed view passed)
  bool robust = false;        // 0x200
  bool refine_points = false;  // 0x400
  bool ml_adjust = false;      // 0x800
  int idx_radial = -1;         // first radial parameter inside the camera block
  int idx_tangential = -1;
  int n_camera = 5;  // live camera parameters: 5 + n_radial + 2*tangential

  static FunctorConfig decode(uint32_t config) {
    FunctorConfig c;
    c.n_radial = static_cast<int>(config & 0x3u);
    c.n_camera = 5;
    if (c.n_radial > 0) {
      c.idx_radial = c.n_camera;
      c.n_camera += c.n_radial;
    }
    c.tangential = (config & 0x4u) != 0;
    if (c.tangential) {
      c.idx_tangential = c.n_camera;
      c.n_camera += 2;
    }
    c.refine_poses = (config & 0x100u) != 0;
    c.robust = (config & 0x200u) != 0;
    c.refine_points = (config & 0x400u) != 0;
    c.ml_adjust = (config & 0x800u) != 0;
    return c;
  }
};

// Per-observation constants held by one reference functor object (BundleAdjustment.h:81-101).
struct ObservationConstants {
  double obs_x, obs_y;  // observed micro-image point (raw px)
  double ml_x, ml_y;    // micro-lens centre (raw px)
  double sx, sy;        // raw pixel size spx/scale, spy/scale (mm)   (:86-87)
  double scale;         // depth_to_raw_im_scale                       (:88)
};

// ---- src/CameraModel.h:205-223 ------------------------------------------------------------------
template <class T>
inline void radial_shift(const T& x, const T& y, T& dx, T& dy, const T* k, int nk) {
  T r[5];
  if (nk > 5) nk = 5;
  r[0] = x * x + y * y;
  T dr = k[0] * r[0];
  for (int i = 1; i < nk; ++i) {
    r[i] = r[i - 1] * r[0];
    dr += k[i] * r[i];
  }
  dx = x * dr;
  dy = y * dr;
}

// ---- src/CameraModel.h:228-241 ------------------------------------------------------------------
template <class T>
inline void tangential_shift(const T& x, const T& y, T& dx, T& dy, const T* t) {
  if (t == nullptr) {
    dx = T(0.0);
    dy = T(0.0);
    return;
  }
  T r2 = x * x + y * y;
  dx = t[0] * (r2 + T(2.0) * x * x) + T(2.0) * t[1] * x * y;
  dy = t[1] * (r2 + T(2.0) * y * y) + T(2.0) * t[0] * x * y;
}

// ---- src/CameraModel.h:246-264 ------------------------------------------------------------------
// R = Rx(a0) Ry(a1) Rz(a2) the way Eigen evaluates AngleAxis products: each AngleAxis becomes a unit
// quaternion (cos(a/2), sin(a/2)*axis), the quaternions are multiplied, the product is converted to
// a rotation matrix. M is row-major 3x4 = [R | t] (the reference's 4x4 minus its constant last row).
template <class T>
struct Quat {
  T w, x, y, z;
};
template <class T>
inline Quat<T> quat_mul(const Quat<T>& a, const Quat<T>& b) {
  Quat<T> q;
  q.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  q.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  q.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  q.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return q;
}
template <class T>
inline Quat<T> quat_axis(const T& angle, int axis) {
  T ha = T(0.5) * angle;
  T s = sin(ha);
  Quat<T> q;
  q.w = cos(ha);
  q.x = s * T(axis == 0 ? 1.0 : 0.0);
  q.y = s * T(axis == 1 ? 1.0 : 0.0);
  q.z = s * T(axis == 2 ? 1.0 : 0.0);
  return q;
}
template <class T>
inline void pose_matrix(const T* view, T M[12]) {
  Quat<T> q = quat_mul(quat_mul(quat_axis(view[0], 0), quat_axis(view[1], 1)), quat_axis(view[2], 2));
  const T tx = T(2.0) * q.x, ty = T(2.0) * q.y, tz = T(2.0) * q.z;
  const T twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const T txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const T tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  M[0] = T(1.0) - (tyy + tzz);
  M[1] = txy - twz;
  M[2] = txz + twy;
  M[4] = txy + twz;
  M[5] = T(1.0) - (txx + tzz);
  M[6] = tyz - twx;
  M[8] = txz - twy;
  M[9] = tyz + twx;
  M[10] = T(1.0) - (txx + tyy);
  M[3] = view[3];
  M[7] = view[4];
  M[11] = view[5];
}

// RT * [X;1] for the first three rows (BundleAdjustment.h:169-172). Eigen's fixed-size 4-term dot
// is reduced pairwise: (m0 p0 + m1 p1) + (m2 p2 + m3 p3).
template <class T>
inline void transform_point(const T M[12], const T p[4], T out[3]) {
  for (int i = 0; i < 3; ++i) {
    out[i] = (M[4 * i + 0] * p[0] + M[4 * i + 1] * p[1]) + (M[4 * i + 2] * p[2] + M[4 * i + 3] * p[3]);
  }
}

// ---- src/CameraModel.h:87-199 -------------------------------------------------------------------
// pc: point in camera coordinates; k/t may be nullptr (no radial / no tangential).
template <class T>
inline void micro_image_projection(T& out_x, T& out_y, const T pc[3], const T& sx, const T& sy, const T& fL,
                                   const T& bL0, const T& B, const T c_raw[2], const T ml[2], const T* k,
                                   int nk, const T* t, bool ml_adjust) {
  // micro-lens centre on the MLA plane in mm, then 10 fixed-point steps of undistortion (:93-125)
  T cd[2] = {(ml[0] - c_raw[0]) * sx, (ml[1] - c_raw[1]) * sy};
  T cu[2] = {cd[0], cd[1]};
  const bool any_dist = (nk > 0) || (t != nullptr);
  if (any_dist) {
    T rx = T(0.0), ry = T(0.0), tx = T(0.0), ty = T(0.0);
    for (int it = 0; it < 10; ++it) {
      if (nk > 0) radial_shift<T>(cu[0], cu[1], rx, ry, k, nk);
      if (t != nullptr) tangential_shift<T>(cu[0], cu[1], tx, ty, t);
      cu[0] = cd[0] - rx - tx;
      cu[1] = cd[1] - ry - ty;
    }
  }
  if (ml_adjust) {  // :127-131
    cu[0] = cu[0] / (bL0 + B) * bL0;
    cu[1] = cu[1] / (bL0 + B) * bL0;
  }
  T zC0 = fL * bL0 / (fL - bL0);                                        // :133
  T pML[2] = {-cu[0] * fL / (fL - bL0), -cu[1] * fL / (fL - bL0)};      // :135-137
  T q[3] = {pc[0] - pML[0], pc[1] - pML[1], pc[2] + zC0};               // :139-142
  const T qz = q[2];
  q[0] = q[0] / qz;  // :144 (Eigen's vector /= scalar divides component-wise)
  q[1] = q[1] / qz;
  T pm[2] = {(q[0] - cu[0] / fL) * fL * B / (fL - bL0), (q[1] - cu[1] / fL) * fL * B / (fL - bL0)};  // :146-148
  T wx, wy;
  if (ml_adjust) {  // :152-176
    wx = pm[0] + cu[0];
    wy = pm[1] + cu[1];
    if (any_dist) {
      T rx = T(0.0), ry = T(0.0), tx = T(0.0), ty = T(0.0);
      if (nk > 0) radial_shift<T>(wx, wy, rx, ry, k, nk);
      if (t != nullptr) tangential_shift<T>(wx, wy, tx, ty, t);
      wx += rx + tx;
      wy += ry + ty;
    }
  } else {  // :177-192 (the inner mlCenterAdjustment branch there is unreachable)
    T one = T(1.0);
    wx = pm[0] * one + cd[0];
    wy = pm[1] * one + cd[1];
  }
  out_x = wx / sx + c_raw[0];  // :194-195
  out_y = wy / sy + c_raw[1];
}

// ---- BundleAdjustment.h:120-195 -----------------------------------------------------------------
// camera: 17-wide block; view: 6 or nullptr; point: 3 or nullptr.
// When the pose is not refined, `fixed_pc` holds hCamCoord (RT*X computed once in double, :94-101).
// When the point is not refined but the pose is, `fixed_point` holds the stored object point (:157).
template <class T>
inline bool reprojection_residual(const FunctorConfig& cfg, const ObservationConstants& oc, const T* camera,
                                  const T* view, const T* point, const double* fixed_point,
                                  const double* fixed_pc, T* residuals) {
  T fL = camera[0];
  if (fL < T(0.0)) fL = -fL;
  T bL0 = camera[1];
  if (bL0 < T(0.0)) bL0 = -bL0;
  T B = camera[2];
  if (B < T(0.0)) B = -B;
  T c_raw[2];
  c_raw[0] = (camera[3] + T(0.5)) * T(oc.scale) - T(0.5);
  c_raw[1] = (camera[4] + T(0.5)) * T(oc.scale) - T(0.5);
  if (c_raw[0] < T(0.0)) c_raw[0] = -c_raw[0];
  if (c_raw[1] < T(0.0)) c_raw[1] = -c_raw[1];

  const T* k = cfg.n_radial > 0 ? camera + cfg.idx_radial : nullptr;
  const T* t = cfg.tangential ? camera + cfg.idx_tangential : nullptr;

  T pw[4];
  if (cfg.refine_points) {
    for (int i = 0; i < 3; ++i) pw[i] = point[i];
  } else {
    for (int i = 0; i < 3; ++i) pw[i] = T(fixed_point ? fixed_point[i] : 0.0);
  }
  pw[3] = T(1.0);

  T pc[3];
  if (cfg.refine_poses) {
    T M[12];
    pose_matrix<T>(view, M);
    transform_point<T>(M, pw, pc);
  } else {
    for (int i = 0; i < 3; ++i) pc[i] = T(fixed_pc[i]);
  }

  T ml[2] = {T(oc.ml_x), T(oc.ml_y)};
  T px, py;
  micro_image_projection<T>(px, py, pc, T(oc.sx), T(oc.sy), fL, bL0, B, c_raw, ml, k, cfg.n_radial, t,
                            cfg.ml_adjust);
  residuals[0] = px - T(oc.obs_x);
  residuals[1] = py - T(oc.obs_y);
  return true;
}

// ---- BundleAdjustment.h:262-267 -----------------------------------------------------------------
template <class T>
inline bool distance_residual(double distance, double sigma, const T* p1, const T* p2, T* residual) {
  T d2 = (p1[0] - p2[0]) * (p1[0] - p2[0]) + (p1[1] - p2[1]) * (p1[1] - p2[1]) +
         (p1[2] - p2[2]) * (p1[2] - p2[2]);
  residual[0] = (pow(d2, T(0.5)) - T(distance)) / (T(sigma) + T(0.000001));
  return true;
}

}  // namespace lfba_oracle
