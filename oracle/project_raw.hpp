// oracle/project_raw.hpp — TEST INFRASTRUCTURE (CPU oracle). Not part of the product path.
//
// CPU restatement of the step BEFORE the LF-BA solve: the epipolar-line web
//   CameraCalibration::defineEpiPolarLines      /root/reference/src/CameraCalibration.cpp:521-634
//   EpiPolarLine (ctor, add)                    /root/reference/src/MicroLensGrid/EpiPolarLine.cpp:16-46
// and the projection of every total-focus feature into all micro images that see it
//   CameraCalibration::projectPointsToRawImage  /root/reference/src/CameraCalibration.cpp:640-769.
// The reference computes this in FLOAT32 with a few double sub-expressions; the statement order and every float/double
// conversion below follow the reference line by line (compiled with -ffp-contract=off: a stock x86-64 build of the
// reference has no fused multiply-add). PARITY UNPINNED for the loop itself (it is a member of the OpenCV/COLMAP-bound
// driver class and cannot be compiled here); the EpiPolarLine class is pinned against the reference's own
// EpiPolarLine.cpp compiled in place (oracle/ref_bridge.cpp, tests/test_project_raw_cpu.py).
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace lfba_oracle {

struct EpiLine {
  double ex, ey, dist;  // unit direction, base-line length (px)
};
// EpiPolarLine::EpiPolarLine (:16-31): normalise unless the squared length equals 1.0f
inline EpiLine epi_make(double x, double y, double dist) {
  EpiLine e{x, y, dist};
  const double l2 = e.ex * e.ex + e.ey * e.ey;
  if (l2 != 1.0f) {
    const double l = std::sqrt(l2);
    e.ex /= l;
    e.ey /= l;
  }
  return e;
}
// EpiPolarLine::add (:38-46)
inline EpiLine epi_add(const EpiLine& a, const EpiLine& b) {
  const double x = a.ex * a.dist + b.ex * b.dist;
  const double y = a.ey * a.dist + b.ey * b.dist;
  return epi_make(x, y, std::sqrt(x * x + y * y));
}

// The web: groups of lines with (float-)equal base-line length, ascending. (:521-634)
struct EpiWeb {
  std::vector<EpiLine> lines;     // grouped
  std::vector<int32_t> group_begin;  // [n_groups + 1]
};

typedef EpiLine (*epi_make_fn)(double, double, double);
typedef EpiLine (*epi_add_fn)(const EpiLine&, const EpiLine&);

inline EpiWeb build_epi_web(float lens_diameter, float rotation, bool rotation_on_grid, epi_make_fn mk = epi_make,
                            epi_add_fn add = epi_add) {
  const float maxDist = lens_diameter * 10;
  EpiLine epl_0 = mk(1, 0, lens_diameter), epl_1 = mk(0.5, std::sqrt(0.75), lens_diameter);
  EpiLine n_epl_1 = mk(-0.5, -std::sqrt(0.75), lens_diameter), epl_2 = mk(0.5, -std::sqrt(0.75), lens_diameter);
  EpiLine n_epl_2 = mk(-0.5, std::sqrt(0.75), lens_diameter);
  if (rotation_on_grid) {
    const double ca = std::cos(rotation), sa = std::sin(rotation);  // cos(float) promotes to double as in the reference
    for (EpiLine* e : {&epl_0, &epl_1, &n_epl_1, &epl_2, &n_epl_2}) {
      const double ex = e->ex, ey = e->ey;
      e->ex = ex * ca + ey * sa;
      e->ey = -ex * sa + ey * ca;
    }
  }
  std::vector<EpiLine> L;
  L.push_back(epl_1);
  L.push_back(epl_2);
  int i = 0;
  while (L.back().dist < maxDist) {
    EpiLine a, b;
    if (i % 2 == 0) {
      a = add(L[i * 2], n_epl_2);
      b = add(L[i * 2 + 1], n_epl_1);
    } else {
      a = add(L[i * 2], epl_1);
      b = add(L[i * 2 + 1], epl_2);
    }
    L.push_back(a);
    L.push_back(b);
    ++i;
  }
  L.push_back(epl_0);
  const int init_len = (int)L.size();
  for (int k = 0; k < init_len; ++k) {
    EpiLine last = L[k];
    while (last.dist < maxDist) {
      L.push_back(add(last, epl_0));
      last = L.back();
    }
  }
  // insertion into the distance-sorted web (:596-632); groups compare their FIRST line's length, equality in float
  std::vector<std::vector<EpiLine>> web;
  web.push_back({L[0]});
  for (size_t k = 1; k < L.size(); ++k) {
    if (L[k].ey == -1.0f || L[k].dist > maxDist) continue;
    bool smaller = false, equal = false;
    size_t ii;
    for (ii = 0; ii < web.size() && !(smaller || equal); ++ii) {
      if ((float)web[ii][0].dist == (float)L[k].dist) equal = true;
      else if (web[ii][0].dist > L[k].dist) smaller = true;
    }
    if (smaller) web.insert(web.begin() + (ii - 1), std::vector<EpiLine>{L[k]});
    else if (equal) web[ii - 1].push_back(L[k]);
    else web.push_back({L[k]});
  }
  EpiWeb w;
  w.group_begin.push_back(0);
  for (auto& g : web) {
    for (auto& e : g) w.lines.push_back(e);
    w.group_begin.push_back((int32_t)w.lines.size());
  }
  return w;
}

struct LensGrid {
  int raw_width, raw_height, scale;
  float lens_diameter, lens_validity_radius_2;
  const float *cx, *cy;        // lens centres (MicroLens::centerX/Y are float)
  const int32_t* map_next;     // mapNextMl: nearest lens of every raw pixel (-1 = NULL)
  const int32_t* map_ml;       // mapMlPointer: lens whose valid micro image covers the pixel (-1 = NULL)
};

struct RawObs {
  double x, y, mlx, mly;
  int32_t feature;
};

// :640-769 for ONE feature; appends to out. Returns the number of observations appended.
inline int project_feature(const LensGrid& g, const EpiWeb& web, double img_x, double img_y, double vdepth_d, int32_t feature,
                           std::vector<RawObs>* out) {
  const float vdepth = (float)vdepth_d;
  if (!(vdepth > 2.0 && vdepth < 20.0)) return 0;
  const float x = (float)img_x, y = (float)img_y;
  const float radius = g.lens_diameter * 0.5f * vdepth + 2.0f;
  const float radius_2 = radius * radius;
  const float xUps = ((float)g.scale) * (x + 0.5f) - 0.5f;
  const float yUps = ((float)g.scale) * (y + 0.5f) - 0.5f;
  int xi = (int)(xUps + 0.5f);
  if (xi >= g.raw_width) xi = g.raw_width - 1;
  int yi = (int)(yUps + 0.5f);
  if (yi >= g.raw_height) yi = g.raw_height - 1;
  if (xi < 0 || yi < 0) return 0;  // (the reference would index out of bounds; features are inside the image)
  const int32_t ml0 = g.map_next[xi + g.raw_width * yi];
  if (ml0 < 0) return 0;
  const float cxn = g.cx[ml0], cyn = g.cy[ml0];
  const float dxn = cxn - xUps, dyn = cyn - yUps;
  const float d2n = dxn * dxn + dyn * dyn;
  if (d2n > radius_2) return 0;
  std::vector<int32_t> lenses;  // std::vector<MicroLens*> microLenses (:682)
  lenses.push_back(ml0);
  const int ng = (int)web.group_begin.size() - 1;
  for (int gi = 0; gi < ng; ++gi) {
    if (web.lines[web.group_begin[gi]].dist > radius) break;  // double > float
    for (int li = web.group_begin[gi]; li < web.group_begin[gi + 1]; ++li) {
      for (int s = 0; s < 2; ++s) {
        const float bl = (float)web.lines[li].dist;
        const float epx = s == 0 ? (float)web.lines[li].ex : (float)(-web.lines[li].ex);
        const float epy = s == 0 ? (float)web.lines[li].ey : (float)(-web.lines[li].ey);
        const float cx = cxn + bl * epx;
        const float cy = cyn + bl * epy;
        const float ddx = cx - xUps, ddy = cy - yUps;
        const float dd2 = ddx * ddx + ddy * ddy;
        if (dd2 > radius_2) continue;
        int cxi = (int)(cx + 0.5), cyi = (int)(cy + 0.5);  // double arithmetic, truncation
        if (cxi < 0) cxi = 0;
        if (cxi >= g.raw_width) cxi = g.raw_width - 1;
        if (cyi < 0) cyi = 0;
        if (cyi >= g.raw_height) cyi = g.raw_height - 1;
        const int32_t ml = g.map_ml[cxi + cyi * g.raw_width];
        if (ml < 0) continue;
        lenses.push_back(ml);
      }
    }
  }
  int n = 0;
  for (size_t k = 0; k < lenses.size(); ++k) {
    const float cx = g.cx[lenses[k]], cy = g.cy[lenses[k]];
    const float xR = (xUps - cx) / vdepth + cx;
    const float yR = (yUps - cy) / vdepth + cy;
    if (!(xR >= 0 && xR <= g.raw_width - 1 && yR >= 0 && yR <= g.raw_height - 1)) continue;
    const float tx = xR - cx, ty = yR - cy;
    const float t2 = tx * tx + ty * ty;
    if (t2 >= g.lens_validity_radius_2) continue;
    if (out) out->push_back(RawObs{(double)xR, (double)yR, (double)cx, (double)cy, feature});
    ++n;
  }
  return n;
}

}  // namespace lfba_oracle
