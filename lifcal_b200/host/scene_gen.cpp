// lifcal_b200/host/scene_gen.cpp — seeded synthetic plenoptic scenes (host only; see include/lfba_scene.h).
//
// Produces the inputs of CameraCalibration::performBundleAdjustment() (reference:
// src/CameraCalibration.cpp:859-925) for a Raytrix-like focused plenoptic camera:
//   raw 2048x2048 px, pixelSize 0.0055 mm read as float (src/Utility/Settings.cpp:178), scale 2,
//   fL=35.0, bL0=33.07, B=0.57 (virtual depth 4..8 for Z in 0.5..3.5 m), c=(511.3,512.9),
//   k=(2e-4,-3e-7), t=(1e-5,-2e-5); hexagonal MLA, lens diameter 23 px, rotation 0.003 rad, centres held
//   as float32 (src/MicroLensGrid/MicroLens.h:22-23), validity radius D/2-1 (MicroLensGrid.cpp:108-111).
// A (point, frame) pair contributes one observation per micro lens whose micro image sees the virtual
// image point (src/CameraCalibration.cpp:655-764): virtual depth 2<v<20, |x_V - c| < v*(D/2-1).
// The observed position is the forward model at ground truth (the plenoptic projection of
// src/CameraModel.h:87-199 written out in plain double below) plus Gaussian noise, rounded through float32
// like the reference's xR,yR (:748-761).
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/lfba_scene.h"

namespace {

// ---- counter-based RNG: splitmix64 over (seed, stream, index) ----
inline uint64_t mix64(uint64_t z) {
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
struct Rng {
  uint64_t key;
  uint64_t ctr = 0;
  Rng(uint64_t seed, uint64_t stream, uint64_t index) {
    key = mix64(seed ^ mix64(stream * 0x100000001b3ull + 0x51ed270b)) ^ mix64(index + 0x2545f4914f6cdd1dull);
  }
  uint64_t next() { return mix64(key + (ctr++) * 0xd1342543de82ef95ull); }
  double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }  // [0,1)
  double normal() {
    double u1 = uniform(), u2 = uniform();
    if (u1 < 1e-300) u1 = 1e-300;
    return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
  }
};

struct Camera {
  double fL, bL0, B, cx, cy;  // cx, cy in total-focus px
  double k[2], t[2];
  int n_radial;
  bool tangential, ml_adjust;
  double spx, scale;  // pixelSize_totFoc, depth_to_raw_im_scale
  double s_raw() const { return spx / scale; }
  double craw_x() const { return (cx + 0.5) * scale - 0.5; }
  double craw_y() const { return (cy + 0.5) * scale - 0.5; }
};

inline void lens_shift(const Camera& c, double x, double y, double& dx, double& dy) {
  const double r2 = x * x + y * y;
  dx = dy = 0.0;
  if (c.n_radial > 0) {
    double d = c.k[0] * r2;
    if (c.n_radial > 1) d += c.k[1] * r2 * r2;
    dx += x * d;
    dy += y * d;
  }
  if (c.tangential) {
    dx += c.t[0] * (r2 + 2.0 * x * x) + 2.0 * c.t[1] * x * y;
    dy += c.t[1] * (r2 + 2.0 * y * y) + 2.0 * c.t[0] * x * y;
  }
}

// forward model: camera-frame point -> position in the micro image of the lens centred at (mlx, mly)
inline void project_truth(const Camera& c, const double pc[3], double mlx, double mly, double& ox, double& oy) {
  const double s = c.s_raw();
  const double crx = c.craw_x(), cry = c.craw_y();
  const double cdx = (mlx - crx) * s, cdy = (mly - cry) * s;
  double ux = cdx, uy = cdy;
  const bool dist = c.n_radial > 0 || c.tangential;
  if (dist)
    for (int i = 0; i < 10; ++i) {
      double dx, dy;
      lens_shift(c, ux, uy, dx, dy);
      ux = cdx - dx;
      uy = cdy - dy;
    }
  if (c.ml_adjust) {
    ux = ux / (c.bL0 + c.B) * c.bL0;
    uy = uy / (c.bL0 + c.B) * c.bL0;
  }
  const double D = c.fL - c.bL0;
  const double zC0 = c.fL * c.bL0 / D;
  const double qz = pc[2] + zC0;
  const double qx = (pc[0] + ux * c.fL / D) / qz, qy = (pc[1] + uy * c.fL / D) / qz;
  const double px = (qx - ux / c.fL) * c.fL * c.B / D, py = (qy - uy / c.fL) * c.fL * c.B / D;
  double wx, wy;
  if (c.ml_adjust) {
    wx = px + ux;
    wy = py + uy;
    if (dist) {
      double dx, dy;
      lens_shift(c, wx, wy, dx, dy);
      wx += dx;
      wy += dy;
    }
  } else {
    wx = px + cdx;
    wy = py + cdy;
  }
  ox = wx / s + crx;
  oy = wy / s + cry;
}

// R = Rx(a0) Ry(a1) Rz(a2), row-major
inline void euler_xyz(const double a[3], double R[9]) {
  const double cx = std::cos(a[0]), sx = std::sin(a[0]);
  const double cy = std::cos(a[1]), sy = std::sin(a[1]);
  const double cz = std::cos(a[2]), sz = std::sin(a[2]);
  R[0] = cy * cz;
  R[1] = -cy * sz;
  R[2] = sy;
  R[3] = sx * sy * cz + cx * sz;
  R[4] = -sx * sy * sz + cx * cz;
  R[5] = -sx * cy;
  R[6] = -cx * sy * cz + sx * sz;
  R[7] = cx * sy * sz + sx * cz;
  R[8] = cx * cy;
}

const int kRaw = 2048;
const double kLensDiameter = 23.0;
const double kGridRot = 0.003;

struct Obs {
  float ox, oy, mx, my;
  int32_t frame;
};

}  // namespace

struct lfba_scene {
  lfba_scene_spec spec;
  Camera truth;
  std::vector<double> camera_true, camera_init, views_true, views_init, points_true, points_init;
  std::vector<double> obs_x, obs_y, ml_x, ml_y;
  std::vector<int32_t> point_idx, frame_idx;
  std::vector<int32_t> c_p1, c_p2;
  std::vector<double> c_dist, c_sigma;
  int64_t n_tracks = 0;
};

extern "C" void lfba_scene_spec_init(lfba_scene_spec* s) {
  std::memset(s, 0, sizeof(*s));
  s->seed = 20240910ull;
  s->n_points = 500;
  s->n_frames = 10;
  s->window = 0;
  s->config = 2u | LFBA_CFG_TANGENTIAL | LFBA_CFG_REFINE_POSES | LFBA_CFG_ROBUST | LFBA_CFG_REFINE_POINTS |
              LFBA_CFG_MLADJ;
  s->calib_type = LFBA_CALIBRATION_ARUCO;
  s->n_constraints = 0;
  s->max_lenses = 64;
  s->order = 0;
  s->noise_px = 0.1;
  s->outlier_fraction = -1.0;  // resolved from the robust flag
  s->outlier_px = 5.0;
  s->init_intrinsics_rel = 2e-4;
  s->init_center_px = 1.0;
  s->init_angle_rad = 1e-3;
  s->init_trans_mm = 0.5;
  s->init_point_mm = 1.0;
}

extern "C" int lfba_scene_spec_preset(lfba_scene_spec* s, int cfg) {
  lfba_scene_spec_init(s);
  s->seed = 20240910ull + (uint64_t)cfg;
  switch (cfg) {
    case 1:  // calib_marker: 500 points x 10 frames, 3 distance constraints
      s->n_points = 500;
      s->n_frames = 10;
      s->n_constraints = 3;
      return 0;
    case 2:  // recalib: 5k points x 20 frames, fL and B fixed, bounds, no constraints
      s->n_points = 5000;
      s->n_frames = 20;
      s->calib_type = LFBA_RECALIBRATION;
      return 0;
    case 3:  // full calibration: 50k points x 100 frames, visibility window 20
      s->n_points = 50000;
      s->n_frames = 100;
      s->window = 20;
      s->n_constraints = 3;
      return 0;
    case 4:  // scaled scene: 1M points x 1000 frames, window 4, ~1e8 observations
      s->n_points = 1000000;
      s->n_frames = 1000;
      s->window = 4;
      s->order = 1;
      return 0;
    default:
      return -1;
  }
}

extern "C" lfba_scene* lfba_scene_create(const lfba_scene_spec* spec_in) {
  lfba_scene* sc = new lfba_scene();
  sc->spec = *spec_in;
  lfba_scene_spec& sp = sc->spec;
  const int P = sp.n_points, F = sp.n_frames;
  if (P <= 0 || F <= 0) {
    delete sc;
    return nullptr;
  }
  const int W = (sp.window <= 0 || sp.window >= F) ? F : sp.window;
  if (sp.max_lenses <= 0) sp.max_lenses = 64;
  const bool robust = (sp.config & LFBA_CFG_ROBUST) != 0;
  const double outlier_fraction = sp.outlier_fraction < 0 ? (robust ? 0.02 : 0.0) : sp.outlier_fraction;
  const int pb = (sp.point_begin == 0 && sp.point_end == 0) ? 0 : sp.point_begin;
  const int pe = (sp.point_begin == 0 && sp.point_end == 0) ? P : std::min(P, sp.point_end);
  const int nthreads = sp.num_threads > 0 ? sp.num_threads : omp_get_max_threads();

  Camera& cam = sc->truth;
  cam.spx = 2.0 * (double)(float)0.0055;  // depth_to_raw_im_scale * pixelSize, pixelSize read as float
  cam.scale = 2.0;
  cam.fL = 35.0;
  cam.bL0 = 33.07;
  cam.B = 0.57;
  cam.cx = 511.3;
  cam.cy = 512.9;
  cam.n_radial = (int)(sp.config & LFBA_CFG_NRADIAL_MASK);
  cam.tangential = (sp.config & LFBA_CFG_TANGENTIAL) != 0;
  cam.ml_adjust = (sp.config & LFBA_CFG_MLADJ) != 0;
  cam.k[0] = cam.n_radial > 0 ? 2e-4 : 0.0;
  cam.k[1] = cam.n_radial > 1 ? -3e-7 : 0.0;
  cam.t[0] = cam.tangential ? 1e-5 : 0.0;
  cam.t[1] = cam.tangential ? -2e-5 : 0.0;

  // ---- camera block: truth and initial guess (src/CameraCalibration.cpp:832-853: distortion starts at 0) ----
  sc->camera_true.assign(17, 0.0);
  sc->camera_init.assign(17, 0.0);
  {
    double* ct = sc->camera_true.data();
    ct[0] = cam.fL;
    ct[1] = cam.bL0;
    ct[2] = cam.B;
    ct[3] = cam.cx;
    ct[4] = cam.cy;
    int idx = 5;
    for (int i = 0; i < cam.n_radial; ++i) ct[idx++] = cam.k[i];
    if (cam.tangential) {
      ct[idx++] = cam.t[0];
      ct[idx++] = cam.t[1];
    }
    Rng r(sp.seed, 1, 0);
    double* ci = sc->camera_init.data();
    const bool recalib = sp.calib_type == LFBA_RECALIBRATION;
    // recalib: f and B come from the fixed-parameter file and stay constant (:930-940)
    ci[0] = recalib ? cam.fL : cam.fL * (1.0 + sp.init_intrinsics_rel * r.normal());
    ci[1] = cam.bL0 * (1.0 + sp.init_intrinsics_rel * r.normal());
    ci[2] = recalib ? cam.B : cam.B * (1.0 + sp.init_intrinsics_rel * r.normal());
    ci[3] = cam.cx + sp.init_center_px * r.normal();
    ci[4] = cam.cy + sp.init_center_px * r.normal();
  }

  // ---- poses: smooth trajectory, Euler XYZ + translation (world -> camera) ----
  sc->views_true.assign((size_t)6 * F, 0.0);
  sc->views_init.assign((size_t)6 * F, 0.0);
  for (int f = 0; f < F; ++f) {
    const double ph = 6.283185307179586 * (double)f / (double)std::min(F, 40);
    const double A = 0.02;  // ~1.1 degrees
    double a[3] = {A * std::sin(ph + 1.0), A * std::cos(1.3 * ph), 0.5 * A * std::sin(0.7 * ph)};
    const double C[3] = {60.0 * std::cos(ph), 60.0 * std::sin(ph), 25.0 * std::sin(0.5 * ph)};
    double R[9];
    euler_xyz(a, R);
    double* v = &sc->views_true[(size_t)6 * f];
    for (int i = 0; i < 3; ++i) {
      v[i] = a[i];
      v[3 + i] = -(R[3 * i] * C[0] + R[3 * i + 1] * C[1] + R[3 * i + 2] * C[2]);
    }
    Rng r(sp.seed, 2, (uint64_t)f);
    double* vi = &sc->views_init[(size_t)6 * f];
    for (int i = 0; i < 3; ++i) {
      vi[i] = v[i] + sp.init_angle_rad * r.normal();
      vi[3 + i] = v[3 + i] + sp.init_trans_mm * r.normal();
    }
  }

  // ---- points: sampled in the frustum of the centre frame of their visibility window ----
  sc->points_true.assign((size_t)3 * P, 0.0);
  sc->points_init.assign((size_t)3 * P, 0.0);
  std::vector<int32_t> win0((size_t)P);
  const double s_raw = cam.s_raw();
  const double crx = cam.craw_x(), cry = cam.craw_y();
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int p = 0; p < P; ++p) {
    const int f0 = (W >= F) ? 0 : (int)(((int64_t)p * (F - W + 1)) / P);
    win0[p] = f0;
    const int fc = std::min(F - 1, f0 + W / 2);
    Rng r(sp.seed, 3, (uint64_t)p);
    const double margin = 260.0;
    const double u = margin + (kRaw - 1 - 2 * margin) * r.uniform();
    const double v = margin + (kRaw - 1 - 2 * margin) * r.uniform();
    // virtual depth uniform in [4.2, 7.8] (Z ~ 0.52..3.0 m): ~0.76 v^2 = 13..46 micro images per view
    const double vdepth = 4.2 + 3.6 * r.uniform();
    const double bL = cam.bL0 + vdepth * cam.B;
    const double Z = cam.fL * bL / (bL - cam.fL);
    const double pc[3] = {(u - crx) * s_raw * Z / bL, (v - cry) * s_raw * Z / bL, Z};
    double R[9];
    const double* vw = &sc->views_true[(size_t)6 * fc];
    euler_xyz(vw, R);
    double d[3] = {pc[0] - vw[3], pc[1] - vw[4], pc[2] - vw[5]};
    double* pt = &sc->points_true[(size_t)3 * p];
    for (int i = 0; i < 3; ++i) pt[i] = R[i] * d[0] + R[3 + i] * d[1] + R[6 + i] * d[2];  // R^T d
    double* pi = &sc->points_init[(size_t)3 * p];
    for (int i = 0; i < 3; ++i) pi[i] = pt[i] + sp.init_point_mm * r.normal();
  }

  // ---- observations, generated per point (parallel), then laid out in the requested order ----
  const double gcx = 0.5 * (kRaw - 1), gcy = 0.5 * (kRaw - 1);  // grid rotation centre
  const double cr = std::cos(kGridRot), sr = std::sin(kGridRot);
  const double pitch_y = kLensDiameter * 0.8660254037844386;
  const double valid_r = kLensDiameter * 0.5 - 1.0;
  std::vector<std::vector<Obs>> per_point((size_t)(pe - pb));
  std::vector<int32_t> tracks_per_point((size_t)(pe - pb), 0);
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads)
  for (int p = pb; p < pe; ++p) {
    std::vector<Obs>& out = per_point[(size_t)(p - pb)];
    const double* pw = &sc->points_true[(size_t)3 * p];
    const int f0 = win0[p], f1 = std::min(F, f0 + W);
    int ntr = 0;
    for (int f = f0; f < f1; ++f) {
      const double* vw = &sc->views_true[(size_t)6 * f];
      double R[9];
      euler_xyz(vw, R);
      double pc[3];
      for (int i = 0; i < 3; ++i) pc[i] = R[3 * i] * pw[0] + R[3 * i + 1] * pw[1] + R[3 * i + 2] * pw[2] + vw[3 + i];
      if (pc[2] <= 1.2 * cam.fL) continue;
      const double bL = cam.fL * pc[2] / (pc[2] - cam.fL);
      const double vd = (bL - cam.bL0) / cam.B;  // virtual depth
      if (!(vd > 2.0 && vd < 20.0)) continue;     // src/CameraCalibration.cpp:655
      const double xv = pc[0] / pc[2] * bL / s_raw + crx, yv = pc[1] / pc[2] * bL / s_raw + cry;
      if (!(xv >= 0 && xv <= kRaw - 1 && yv >= 0 && yv <= kRaw - 1)) continue;
      // candidate lenses: un-rotate the virtual image point into grid coordinates
      const double rad = vd * valid_r;
      const double gx = cr * (xv - gcx) + sr * (yv - gcy), gy = -sr * (xv - gcx) + cr * (yv - gcy);
      const int j0 = (int)std::floor((gy - rad) / pitch_y) - 1, j1 = (int)std::ceil((gy + rad) / pitch_y) + 1;
      int count = 0;
      Rng r(sp.seed, 4, (uint64_t)p * (uint64_t)F + (uint64_t)f);
      for (int j = j0; j <= j1 && count < sp.max_lenses; ++j) {
        const double offx = (j & 1) ? 0.5 * kLensDiameter : 0.0;
        const int i0 = (int)std::floor((gx - rad - offx) / kLensDiameter) - 1;
        const int i1 = (int)std::ceil((gx + rad - offx) / kLensDiameter) + 1;
        for (int i = i0; i <= i1 && count < sp.max_lenses; ++i) {
          const double lx = i * kLensDiameter + offx, ly = j * pitch_y;
          const float mlx = (float)(cr * lx - sr * ly + gcx), mly = (float)(sr * lx + cr * ly + gcy);
          if (!(mlx >= 0 && mlx <= kRaw - 1 && mly >= 0 && mly <= kRaw - 1)) continue;
          const double dx = xv - (double)mlx, dy = yv - (double)mly;
          if (dx * dx + dy * dy >= rad * rad) continue;  // micro image does not see the point (:759)
          double ox, oy;
          project_truth(cam, pc, (double)mlx, (double)mly, ox, oy);
          ox += sp.noise_px * r.normal();
          oy += sp.noise_px * r.normal();
          if (outlier_fraction > 0 && r.uniform() < outlier_fraction) {
            ox += sp.outlier_px * r.normal();
            oy += sp.outlier_px * r.normal();
          }
          if (!(ox >= 0 && ox <= kRaw - 1 && oy >= 0 && oy <= kRaw - 1)) continue;  // :751
          out.push_back(Obs{(float)ox, (float)oy, mlx, mly, f});
          ++count;
        }
      }
      if (count > 0) ++ntr;
    }
    tracks_per_point[(size_t)(p - pb)] = ntr;
  }
  int64_t N = 0;
  std::vector<int64_t> pt_off((size_t)(pe - pb) + 1, 0);
  for (int p = pb; p < pe; ++p) {
    pt_off[(size_t)(p - pb) + 1] = pt_off[(size_t)(p - pb)] + (int64_t)per_point[(size_t)(p - pb)].size();
    sc->n_tracks += tracks_per_point[(size_t)(p - pb)];
  }
  N = pt_off[(size_t)(pe - pb)];
  sc->obs_x.resize((size_t)N);
  sc->obs_y.resize((size_t)N);
  sc->ml_x.resize((size_t)N);
  sc->ml_y.resize((size_t)N);
  sc->point_idx.resize((size_t)N);
  sc->frame_idx.resize((size_t)N);
  if (sp.order == 1) {
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads)
    for (int p = pb; p < pe; ++p) {
      const std::vector<Obs>& v = per_point[(size_t)(p - pb)];
      int64_t o = pt_off[(size_t)(p - pb)];
      for (const Obs& e : v) {
        sc->obs_x[(size_t)o] = e.ox;
        sc->obs_y[(size_t)o] = e.oy;
        sc->ml_x[(size_t)o] = e.mx;
        sc->ml_y[(size_t)o] = e.my;
        sc->point_idx[(size_t)o] = p;
        sc->frame_idx[(size_t)o] = e.frame;
        ++o;
      }
    }
  } else {
    // frame-major, within a frame by point then lens: the order of the reference's nested loops (:859-871)
    std::vector<int64_t> foff((size_t)F + 1, 0);
    for (int p = pb; p < pe; ++p)
      for (const Obs& e : per_point[(size_t)(p - pb)]) foff[(size_t)e.frame + 1]++;
    for (int f = 0; f < F; ++f) foff[(size_t)f + 1] += foff[(size_t)f];
    std::vector<int64_t> cur(foff.begin(), foff.end() - 1);
    for (int p = pb; p < pe; ++p)
      for (const Obs& e : per_point[(size_t)(p - pb)]) {
        const int64_t o = cur[(size_t)e.frame]++;
        sc->obs_x[(size_t)o] = e.ox;
        sc->obs_y[(size_t)o] = e.oy;
        sc->ml_x[(size_t)o] = e.mx;
        sc->ml_y[(size_t)o] = e.my;
        sc->point_idx[(size_t)o] = p;
        sc->frame_idx[(size_t)o] = e.frame;
      }
  }

  // ---- distance constraints between "marker" points: lines `id1 id2 distance sigma` (Constraints.cpp:41-54) ----
  const bool use_c = sp.n_constraints > 0 && (sp.config & LFBA_CFG_REFINE_POINTS) &&
                     sp.calib_type != LFBA_RECALIBRATION && P >= sp.n_constraints + 1;
  if (use_c) {
    const int nm = sp.n_constraints + 1;
    std::vector<int> markers((size_t)nm);
    for (int m = 0; m < nm; ++m) markers[(size_t)m] = (int)(((int64_t)m * (P - 1)) / (nm - 1));
    for (int k = 0; k < sp.n_constraints; ++k) {
      const int a = markers[(size_t)k], b = markers[(size_t)k + 1];
      const double* pa = &sc->points_true[(size_t)3 * a];
      const double* pbp = &sc->points_true[(size_t)3 * b];
      const double d = std::sqrt((pa[0] - pbp[0]) * (pa[0] - pbp[0]) + (pa[1] - pbp[1]) * (pa[1] - pbp[1]) +
                                 (pa[2] - pbp[2]) * (pa[2] - pbp[2]));
      sc->c_p1.push_back(a);
      sc->c_p2.push_back(b);
      sc->c_dist.push_back(d);
      sc->c_sigma.push_back(0.1);  // README.md:180
    }
  }
  return sc;
}

extern "C" void lfba_scene_destroy(lfba_scene* s) { delete s; }

extern "C" void lfba_scene_problem(const lfba_scene* s, lfba_problem* out) {
  std::memset(out, 0, sizeof(*out));
  out->config = s->spec.config;
  out->calib_type = s->spec.calib_type;
  out->spx = out->spy = s->truth.spx;
  out->scale = s->truth.scale;
  out->n_obs = (int64_t)s->obs_x.size();
  out->n_frames = s->spec.n_frames;
  out->n_points = s->spec.n_points;
  out->obs_x = s->obs_x.data();
  out->obs_y = s->obs_y.data();
  out->ml_x = s->ml_x.data();
  out->ml_y = s->ml_y.data();
  out->point_idx = s->point_idx.data();
  out->frame_idx = s->frame_idx.data();
  out->n_constraints = (int32_t)s->c_p1.size();
  out->c_p1 = s->c_p1.data();
  out->c_p2 = s->c_p2.data();
  out->c_dist = s->c_dist.data();
  out->c_sigma = s->c_sigma.data();
}
extern "C" const double* lfba_scene_camera_init(const lfba_scene* s) { return s->camera_init.data(); }
extern "C" const double* lfba_scene_views_init(const lfba_scene* s) { return s->views_init.data(); }
extern "C" const double* lfba_scene_points_init(const lfba_scene* s) { return s->points_init.data(); }
extern "C" const double* lfba_scene_camera_true(const lfba_scene* s) { return s->camera_true.data(); }
extern "C" const double* lfba_scene_views_true(const lfba_scene* s) { return s->views_true.data(); }
extern "C" const double* lfba_scene_points_true(const lfba_scene* s) { return s->points_true.data(); }
extern "C" int64_t lfba_scene_num_tracks(const lfba_scene* s) { return s->n_tracks; }
