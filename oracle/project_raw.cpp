// oracle/project_raw.cpp — TEST INFRASTRUCTURE: C ABI of the N2 oracle (see project_raw.hpp).
#include "project_raw.hpp"

#include <cstring>

using namespace lfba_oracle;

extern "C" {

// Web of epipolar lines. Two-call pattern: with lines == NULL only the counts are returned.
// lines: [n_lines][3] = (ex, ey, dist) grouped; group_begin: [n_groups + 1].
int oracle_epi_web(float lens_diameter, float rotation, int rotation_on_grid, int32_t* n_lines, int32_t* n_groups,
                   double* lines, int32_t* group_begin) {
  const EpiWeb w = build_epi_web(lens_diameter, rotation, rotation_on_grid != 0);
  *n_lines = (int32_t)w.lines.size();
  *n_groups = (int32_t)w.group_begin.size() - 1;
  if (lines)
    for (size_t i = 0; i < w.lines.size(); ++i) {
      lines[3 * i] = w.lines[i].ex;
      lines[3 * i + 1] = w.lines[i].ey;
      lines[3 * i + 2] = w.lines[i].dist;
    }
  if (group_begin) std::memcpy(group_begin, w.group_begin.data(), w.group_begin.size() * sizeof(int32_t));
  return 0;
}

void oracle_epi_make(double x, double y, double dist, double out[3]) {
  const EpiLine e = epi_make(x, y, dist);
  out[0] = e.ex; out[1] = e.ey; out[2] = e.dist;
}
void oracle_epi_add(const double a[3], const double b[3], double out[3]) {
  const EpiLine e = epi_add(EpiLine{a[0], a[1], a[2]}, EpiLine{b[0], b[1], b[2]});
  out[0] = e.ex; out[1] = e.ey; out[2] = e.dist;
}

// projectPointsToRawImage over a feature list (frame-major like the reference's loops). Outputs may be NULL (count only).
int64_t oracle_project_to_raw(int raw_width, int raw_height, int scale, float lens_diameter, float lens_validity_radius_2,
                              float rotation, int rotation_on_grid, int32_t n_lenses, const float* lens_cx,
                              const float* lens_cy, const int32_t* map_next, const int32_t* map_ml, int64_t n_features,
                              const double* feat_x, const double* feat_y, const double* vdepth, int64_t capacity,
                              double* obs_x, double* obs_y, double* ml_x, double* ml_y, int64_t* feature_of_obs) {
  (void)n_lenses;
  const EpiWeb web = build_epi_web(lens_diameter, rotation, rotation_on_grid != 0);
  const LensGrid g{raw_width, raw_height, scale, lens_diameter, lens_validity_radius_2, lens_cx, lens_cy, map_next, map_ml};
  std::vector<RawObs> out;
  int64_t n = 0;
  for (int64_t f = 0; f < n_features; ++f) {
    out.clear();
    const int k = project_feature(g, web, feat_x[f], feat_y[f], vdepth[f], 0, &out);
    for (int j = 0; j < k; ++j, ++n) {
      if (n < capacity) {
        if (obs_x) obs_x[n] = out[j].x;
        if (obs_y) obs_y[n] = out[j].y;
        if (ml_x) ml_x[n] = out[j].mlx;
        if (ml_y) ml_y[n] = out[j].mly;
        if (feature_of_obs) feature_of_obs[n] = f;
      }
    }
  }
  return n;
}

}  // extern "C"
