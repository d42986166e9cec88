// oracle/ref_bridge.cpp — TEST INFRASTRUCTURE.
//
// Compiles the reference's OWN hot-path headers, unmodified and in place,
//     /root/reference/src/BundleAdjustment/BundleAdjustment.h
//     /root/reference/src/CameraModel.h
// against the stand-in headers in oracle/shim/ (Eigen, Ceres, glog and OpenCV/COLMAP are absent from this
// image) and exports the reference functors through a C ABI.  Output: oracle/_ref/libref_functor.so
// (git-ignored; it travels to the GPU box with the snapshot, /root/reference does not).
// Purpose: pin the oracle's restated functor (oracle/functor.hpp) to the reference's actual code —
// tests/test_oracle_ref_pin.py compares residuals and Jet Jacobians on random inputs.
// No reference source is copied: this file only #includes it and mirrors the call sites
// src/CameraCalibration.cpp:879-912 (which Create() overload is used for which flag combination).
#include <cstdint>

#include "BundleAdjustment/BundleAdjustment.h"
// the reference's own EpiPolarLine class, compiled in place (no dependencies): pins the web primitives of the N2 oracle
#include "MicroLensGrid/EpiPolarLine.cpp"

extern "C" {

// One reprojection residual block, evaluated by the reference functor through (shim) AutoDiffCostFunction.
//   view / point non-NULL  -> they are parameter blocks (refinePoses / refine3Dpoints)
//   fixed_point            -> the stored object point (used when the point is not a parameter block)
//   fixed_view             -> the stored view (used when the pose is not refined: hCamCoord precomputed)
// jac: 2 x 26 row-major [camera 17 | view 6 | point 3]; absent blocks are left zero. jac may be NULL.
int ref_block_eval(uint32_t config, const double* obs, const double* ml, double spx, double spy, double scale,
                   const double* camera, const double* view, const double* point, const double* fixed_point,
                   const double* fixed_view, double* residual, double* jac) {
  Eigen::Vector2d mlc(ml[0], ml[1]);
  ceres::CostFunction* cf = nullptr;
  int nblocks = 0;
  const double* params[3] = {camera, nullptr, nullptr};
  if (fixed_view == nullptr) {
    if (point != nullptr) {  // src/CameraCalibration.cpp:882
      cf = OurCostFunctionBundle::Create((int)config, obs[0], obs[1], spx, spy, scale, mlc);
      params[1] = view;
      params[2] = point;
      nblocks = 3;
    } else {  // :885
      Eigen::Vector3d op(fixed_point[0], fixed_point[1], fixed_point[2]);
      cf = OurCostFunctionBundle::Create((int)config, obs[0], obs[1], spx, spy, scale, mlc, &op);
      params[1] = view;
      nblocks = 2;
    }
  } else {  // :906
    Eigen::Vector3d op(fixed_point[0], fixed_point[1], fixed_point[2]);
    double v[6];
    for (int i = 0; i < 6; ++i) v[i] = fixed_view[i];
    cf = OurCostFunctionBundle::Create((int)config, obs[0], obs[1], spx, spy, scale, mlc, &op, v);
    nblocks = 1;
  }
  bool ok;
  if (jac == nullptr) {
    ok = cf->Evaluate(params, residual, nullptr);
  } else {
    double jc[34], jv[12], jp[6];
    double* jacs[3] = {jc, jv, jp};
    ok = cf->Evaluate(params, residual, jacs);
    for (int i = 0; i < 52; ++i) jac[i] = 0.0;
    for (int r = 0; r < 2; ++r) {
      for (int j = 0; j < 17; ++j) jac[26 * r + j] = jc[17 * r + j];
      if (nblocks >= 2)
        for (int j = 0; j < 6; ++j) jac[26 * r + 17 + j] = jv[6 * r + j];
      if (nblocks >= 3)
        for (int j = 0; j < 3; ++j) jac[26 * r + 23 + j] = jp[3 * r + j];
    }
  }
  delete cf;
  return ok ? 1 : 0;
}

// Distance constraint block (src/CameraCalibration.cpp:922-923); jac6 = [d/dp1 (3) | d/dp2 (3)] or NULL.
int ref_distance_eval(double distance, double sigma, const double* p1, const double* p2, double* residual,
                      double* jac6) {
  ceres::CostFunction* cf = OurConstraintFunctionBundle::Create(distance, sigma);
  const double* params[2] = {p1, p2};
  bool ok;
  if (jac6 == nullptr) {
    ok = cf->Evaluate(params, residual, nullptr);
  } else {
    double* jacs[2] = {jac6, jac6 + 3};
    ok = cf->Evaluate(params, residual, jacs);
  }
  delete cf;
  return ok ? 1 : 0;
}

// CameraModel::projectPoint<double> as called from calcReprojectionError (src/CameraCalibration.cpp:1078).
void ref_project_point(const double* pc, double spx_raw, double spy_raw, double fL, double bL0, double B,
                       const double* c_raw, const double* ml, const double* radial, int n_radial,
                       const double* tangential, int ml_adjust, double* out_xy) {
  Eigen::Matrix<double, 3, 1> p(pc[0], pc[1], pc[2]);
  double c[2] = {c_raw[0], c_raw[1]}, m[2] = {ml[0], ml[1]};
  CameraModel::projectPoint<double>(out_xy[0], out_xy[1], p, spx_raw, spy_raw, fL, bL0, B, c, m,
                                    n_radial > 0 ? radial : nullptr, n_radial, tangential, ml_adjust != 0);
}

// RigidBody::getTransformationMatrix<double>, row-major 4x4 out.
void ref_pose_matrix(const double* view, double* rt16) {
  Eigen::Matrix<double, 3, 1> a(view[0], view[1], view[2]), t(view[3], view[4], view[5]);
  Eigen::Matrix<double, 4, 4> RT = RigidBody::getTransformationMatrix<double>(a, t);
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) rt16[4 * i + j] = RT(i, j);
}

// EpiPolarLine::EpiPolarLine / EpiPolarLine::add of the reference (src/MicroLensGrid/EpiPolarLine.cpp:16-46)
void ref_epi_make(double x, double y, double dist, double out[3]) {
  EpiPolarLine e(x, y, dist, 2.0f);
  out[0] = e.epiLine[0];
  out[1] = e.epiLine[1];
  out[2] = e.baseLineDist;
}
void ref_epi_add(const double a[3], const double b[3], double out[3]) {
  EpiPolarLine ea(1, 0, 1, 2.0f), eb(1, 0, 1, 2.0f);
  ea.epiLine[0] = a[0]; ea.epiLine[1] = a[1]; ea.baseLineDist = a[2];
  eb.epiLine[0] = b[0]; eb.epiLine[1] = b[1]; eb.baseLineDist = b[2];
  EpiPolarLine* r = ea.add(eb);
  out[0] = r->epiLine[0];
  out[1] = r->epiLine[1];
  out[2] = r->baseLineDist;
  delete r;
}
}
