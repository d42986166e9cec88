/*
 * lfba_scene.h — seeded synthetic plenoptic scenes for the LF-BA hot path.
 *
 * Test/bench input generator (host only, no CUDA): it produces exactly the arrays that
 * CameraCalibration::performBundleAdjustment() consumes (src/CameraCalibration.cpp:859-925) —
 * micro-image observations with their micro-lens centres, point/frame indices, initial camera / views /
 * points and distance constraints — the way the reference's front end would have produced them:
 *   - hexagonal micro-lens grid with float32 centres  (src/MicroLensGrid/MicroLensGrid.cpp:186-270,
 *     src/MicroLensGrid/MicroLens.h:22-23)
 *   - every micro lens whose micro image sees the virtual image point, virtual depth 2 < v < 20, points
 *     closer than lensValidityRadius to the lens centre  (src/CameraCalibration.cpp:655-764)
 *   - observations and lens centres rounded through float32  (:748-762)
 * The same bytes feed the CPU oracle and the CUDA path.
 */
#ifndef LFBA_SCENE_H_
#define LFBA_SCENE_H_
#include <stdint.h>

#include "lfba.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lfba_scene_spec {
  uint64_t seed;
  int32_t n_points;
  int32_t n_frames;
  int32_t window;        /* frames in which a point is visible; <= 0 or >= n_frames: all frames */
  uint32_t config;       /* LFBA_CFG_* used for the truth model (distortion terms present or not) and flags */
  int32_t calib_type;
  int32_t n_constraints; /* distance constraints between "marker" points (0 = none) */
  int32_t max_lenses;    /* cap on micro images per (point, frame); default 64 */
  int32_t order;         /* 0: frame-major (reference order, :859-871); 1: point-major (generation order) */
  /* shard: generate observations only for points in [point_begin, point_end); (0,0) = all */
  int32_t point_begin, point_end;
  double noise_px;         /* observation noise sigma, default 0.1 */
  double outlier_fraction; /* default 0.02 when LFBA_CFG_ROBUST else 0 */
  double outlier_px;       /* default 5.0 */
  double init_intrinsics_rel; /* relative perturbation of fL,bL0,B for the initial guess, default 2e-4 */
  double init_center_px;      /* default 1.0 */
  double init_angle_rad;      /* default 1e-3 */
  double init_trans_mm;       /* default 0.5 */
  double init_point_mm;       /* default 1.0 */
  int32_t num_threads;        /* 0 = all */
  int32_t reserved[7];
} lfba_scene_spec;

typedef struct lfba_scene lfba_scene;

void lfba_scene_spec_init(lfba_scene_spec* spec); /* defaults of SURVEY.md 8(d) */
/* Preset BASELINE.json configs: cfg = 1..4 (sizes in BASELINE.md section 4). */
int lfba_scene_spec_preset(lfba_scene_spec* spec, int cfg);
lfba_scene* lfba_scene_create(const lfba_scene_spec* spec);
void lfba_scene_destroy(lfba_scene* s);
/* Problem view into the scene's buffers (valid until destroy). */
void lfba_scene_problem(const lfba_scene* s, lfba_problem* out);
/* Initial guess (what performBundleAdjustment starts from) and ground truth. Arrays of 17, 6F, 3P. */
const double* lfba_scene_camera_init(const lfba_scene* s);
const double* lfba_scene_views_init(const lfba_scene* s);
const double* lfba_scene_points_init(const lfba_scene* s);
const double* lfba_scene_camera_true(const lfba_scene* s);
const double* lfba_scene_views_true(const lfba_scene* s);
const double* lfba_scene_points_true(const lfba_scene* s);
int64_t lfba_scene_num_tracks(const lfba_scene* s);

#ifdef __cplusplus
}
#endif
#endif
