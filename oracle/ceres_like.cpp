// oracle/ceres_like.cpp — TEST INFRASTRUCTURE (CPU oracle). Not part of the product path.
//
// CPU restatement of what the reference does between src/CameraCalibration.cpp:858 (ceres::Problem)
// and :965 (ceres::Solve): problem assembly (:859-953), the 6 solver options (:955-962), and the
// Ceres 2.1.0 machinery behind ceres::Solve for this configuration
//   TRUST_REGION / LEVENBERG_MARQUARDT / DENSE_SCHUR / EIGEN dense Cholesky / jacobi_scaling,
//   CauchyLoss(0.5) + Corrector, SubsetManifold{0,2}, box bounds + projected ARMIJO line search.
// Ceres is an un-vendored dependency (pin: /root/reference/installation/Dockerfile:105); what is restated
// here is its published algorithm (internal/ceres/trust_region_minimizer.cc,
// levenberg_marquardt_strategy.cc, schur_eliminator_impl.h, corrector.cc, loss_function.cc,
// line_search.cc, program_evaluator.h), summarised in SURVEY.md Appendix B.  PARITY UNPINNED for the
// LM loop (no Ceres, no golden vectors in the reference); the functor arithmetic IS pinned, see
// oracle/ref_bridge.cpp.
//
// It is also the timed CPU baseline ("Ceres-2.1.0-equivalent": Jet<26> autodiff, Jacobian materialised
// in Ceres' block layout, dense Schur, dense LLT), OpenMP over residual blocks like Ceres' num_threads.
#include <omp.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

#include "functor.hpp"
#include "oracle.h"

namespace lfba_oracle {

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// -------------------------------------------------------------------------------------------------
// Problem in "Ceres program" form
// -------------------------------------------------------------------------------------------------
struct Program {
  FunctorConfig cfg;
  uint32_t config_bits = 0;
  bool recalib = false;
  bool use_constraints = false;
  double spx = 0, spy = 0, scale = 1;
  int64_t N = 0;
  int F = 0, P = 0, K = 0;
  const double *ox = nullptr, *oy = nullptr, *mx = nullptr, *my = nullptr;
  const int32_t *pidx = nullptr, *fidx = nullptr;
  const int32_t *c1 = nullptr, *c2 = nullptr;
  const double *cdist = nullptr, *csigma = nullptr;
  bool has_views = false, has_points = false;

  // tangent layout: [camera tangent | active views | active points]
  std::vector<int> cam_cols;  // tangent camera column -> index in the 17-wide block (SubsetManifold)
  int n_cam_t = 17;
  std::vector<int> view_slot, active_frames;
  std::vector<int> point_slot, active_points;
  std::vector<char> coupled;        // point is touched by a distance constraint -> stays in the reduced system
  std::vector<int> coupled_slot;    // point -> index among coupled points
  std::vector<int> coupled_points;
  int n_t = 0;   // tangent size
  int n_red = 0; // reduced system size
  // CSR: observations grouped by point (stable in input order)
  std::vector<int64_t> pt_begin;
  std::vector<int64_t> pt_obs;
  // bounds on the camera block (recalib, :943-951)
  bool constrained = false;
  double lower[17], upper[17];

  int off_view(int slot) const { return n_cam_t + 6 * slot; }
  int off_point(int slot) const { return n_cam_t + 6 * (int)active_frames.size() + 3 * slot; }
  int red_view(int slot) const { return n_cam_t + 6 * slot; }
  int red_point(int cslot) const { return n_cam_t + 6 * (int)active_frames.size() + 3 * cslot; }
};

static bool build_program(const lfba_problem* pb, const double* camera0, Program& g) {
  g.cfg = FunctorConfig::decode(pb->config);
  g.config_bits = pb->config;
  g.recalib = pb->calib_type == LFBA_RECALIBRATION;
  g.spx = pb->spx;
  g.spy = pb->spy;
  g.scale = pb->scale;
  g.N = pb->n_obs;
  g.F = pb->n_frames;
  g.P = pb->n_points;
  g.ox = pb->obs_x;
  g.oy = pb->obs_y;
  g.mx = pb->ml_x;
  g.my = pb->ml_y;
  g.pidx = pb->point_idx;
  g.fidx = pb->frame_idx;
  g.has_views = g.cfg.refine_poses;
  g.has_points = g.cfg.refine_points;
  if (!g.cfg.refine_poses && g.cfg.refine_points) return false;  // null-deref in the reference (SURVEY C-2)
  // constraints only when points are refined and not in recalibration mode (:916)
  g.use_constraints = g.cfg.refine_points && !g.recalib && pb->n_constraints > 0;
  g.K = g.use_constraints ? pb->n_constraints : 0;
  g.c1 = pb->c_p1;
  g.c2 = pb->c_p2;
  g.cdist = pb->c_dist;
  g.csigma = pb->c_sigma;

  g.cam_cols.clear();
  for (int j = 0; j < 17; ++j) {
    if (g.recalib && (j == 0 || j == 2)) continue;  // SubsetManifold(17,{0,2}) (:930-940)
    g.cam_cols.push_back(j);
  }
  g.n_cam_t = (int)g.cam_cols.size();
  g.constrained = g.recalib;
  for (int j = 0; j < 17; ++j) {
    g.lower[j] = -std::numeric_limits<double>::max();
    g.upper[j] = std::numeric_limits<double>::max();
  }
  if (g.recalib) {  // :943-951, bounds from the INITIAL values
    const int bj[3] = {1, 3, 4};
    for (int b = 0; b < 3; ++b) {
      g.lower[bj[b]] = 0.7 * camera0[bj[b]];
      g.upper[bj[b]] = 1.3 * camera0[bj[b]];
    }
  }

  g.view_slot.assign(g.F, -1);
  g.point_slot.assign(g.P, -1);
  g.coupled.assign(g.P, 0);
  g.coupled_slot.assign(g.P, -1);
  std::vector<char> fseen(g.F, 0), pseen(g.P, 0);
  for (int64_t i = 0; i < g.N; ++i) {
    if (g.fidx[i] < 0 || g.fidx[i] >= g.F || g.pidx[i] < 0 || g.pidx[i] >= g.P) return false;
    fseen[g.fidx[i]] = 1;
    pseen[g.pidx[i]] = 1;
  }
  for (int k = 0; k < g.K; ++k) {
    if (g.c1[k] < 0 || g.c1[k] >= g.P || g.c2[k] < 0 || g.c2[k] >= g.P) return false;
    pseen[g.c1[k]] = pseen[g.c2[k]] = 1;
    g.coupled[g.c1[k]] = g.coupled[g.c2[k]] = 1;
  }
  g.active_frames.clear();
  g.active_points.clear();
  g.coupled_points.clear();
  if (g.has_views)
    for (int f = 0; f < g.F; ++f)
      if (fseen[f]) {
        g.view_slot[f] = (int)g.active_frames.size();
        g.active_frames.push_back(f);
      }
  if (g.has_points)
    for (int p = 0; p < g.P; ++p)
      if (pseen[p]) {
        g.point_slot[p] = (int)g.active_points.size();
        g.active_points.push_back(p);
        if (g.coupled[p]) {
          g.coupled_slot[p] = (int)g.coupled_points.size();
          g.coupled_points.push_back(p);
        }
      }
  g.n_t = g.n_cam_t + 6 * (int)g.active_frames.size() + 3 * (int)g.active_points.size();
  g.n_red = g.n_cam_t + 6 * (int)g.active_frames.size() + 3 * (int)g.coupled_points.size();

  g.pt_begin.assign((size_t)g.P + 1, 0);
  for (int64_t i = 0; i < g.N; ++i) g.pt_begin[g.pidx[i] + 1]++;
  for (int p = 0; p < g.P; ++p) g.pt_begin[p + 1] += g.pt_begin[p];
  g.pt_obs.resize((size_t)g.N);
  {
    std::vector<int64_t> cur(g.pt_begin.begin(), g.pt_begin.end() - 1);
    for (int64_t i = 0; i < g.N; ++i) g.pt_obs[cur[g.pidx[i]]++] = i;
  }
  return true;
}

// -------------------------------------------------------------------------------------------------
// One residual block: value (double functor) or value + autodiff Jacobian (Dual<26>), raw (no loss)
// -------------------------------------------------------------------------------------------------
static inline void fixed_camera_point(const double* view, const double* point, double pc[3]) {
  // hCamCoord = RT * X.homogeneous() in double, BundleAdjustment.h:94-101
  double M[12], pw[4] = {point[0], point[1], point[2], 1.0};
  pose_matrix<double>(view, M);
  transform_point<double>(M, pw, pc);
}

static inline void block_value(const Program& g, int64_t i, const double* camera, const double* views,
                               const double* points, double r[2], oracle_block_fn fn) {
  const int f = g.fidx[i], p = g.pidx[i];
  if (fn) {
    const double obs[2] = {g.ox[i], g.oy[i]}, ml[2] = {g.mx[i], g.my[i]};
    fn(g.config_bits, obs, ml, g.spx, g.spy, g.scale, camera, g.has_views ? views + 6 * f : nullptr,
       g.has_points ? points + 3 * p : nullptr, points + 3 * p, g.has_views ? nullptr : views + 6 * f, r,
       nullptr);
    return;
  }
  ObservationConstants oc{g.ox[i], g.oy[i], g.mx[i], g.my[i], g.spx / g.scale, g.spy / g.scale, g.scale};
  double pc[3];
  if (!g.has_views) fixed_camera_point(views + 6 * f, points + 3 * p, pc);
  reprojection_residual<double>(g.cfg, oc, camera, views + 6 * f, points + 3 * p, points + 3 * p, pc, r);
}

// Jacobian rows: jc[2*17], jv[2*6], jp[2*3] (row-major per block, Ceres layout)
static inline void block_autodiff(const Program& g, int64_t i, const double* camera, const double* views,
                                  const double* points, double r[2], double* jc, double* jv, double* jp,
                                  oracle_block_fn fn) {
  const int f = g.fidx[i], p = g.pidx[i];
  if (fn) {
    const double obs[2] = {g.ox[i], g.oy[i]}, ml[2] = {g.mx[i], g.my[i]};
    double jac[52];
    fn(g.config_bits, obs, ml, g.spx, g.spy, g.scale, camera, g.has_views ? views + 6 * f : nullptr,
       g.has_points ? points + 3 * p : nullptr, points + 3 * p, g.has_views ? nullptr : views + 6 * f, r,
       jac);
    for (int row = 0; row < 2; ++row) {
      for (int j = 0; j < 17; ++j) jc[17 * row + j] = jac[26 * row + j];
      if (jv)
        for (int j = 0; j < 6; ++j) jv[6 * row + j] = jac[26 * row + 17 + j];
      if (jp)
        for (int j = 0; j < 3; ++j) jp[3 * row + j] = jac[26 * row + 23 + j];
    }
    return;
  }
  ObservationConstants oc{g.ox[i], g.oy[i], g.mx[i], g.my[i], g.spx / g.scale, g.spy / g.scale, g.scale};
  // Ceres' AutoDiff seeds one Jet<double, 17+6+3> per parameter; blocks that are absent simply have no
  // partial slot (Jet<23>, Jet<17>).  Using the 26-wide dual for all arities gives identical values.
  typedef Dual<26> D;
  D cam[17], view[6], pt[3], res[2];
  for (int j = 0; j < 17; ++j) cam[j] = D(camera[j], j);
  if (g.has_views)
    for (int j = 0; j < 6; ++j) view[j] = D(views[6 * f + j], 17 + j);
  if (g.has_points)
    for (int j = 0; j < 3; ++j) pt[j] = D(points[3 * p + j], 23 + j);
  double pc[3];
  if (!g.has_views) fixed_camera_point(views + 6 * f, points + 3 * p, pc);
  reprojection_residual<D>(g.cfg, oc, cam, view, pt, points + 3 * p, pc, res);
  for (int row = 0; row < 2; ++row) {
    r[row] = res[row].a;
    for (int j = 0; j < 17; ++j) jc[17 * row + j] = res[row].v[j];
    if (jv)
      for (int j = 0; j < 6; ++j) jv[6 * row + j] = res[row].v[17 + j];
    if (jp)
      for (int j = 0; j < 3; ++j) jp[3 * row + j] = res[row].v[23 + j];
  }
}

// ceres::CauchyLoss(a)::Evaluate (loss_function.cc): b = a^2, c = 1/b
static inline void cauchy(double a, double s, double rho[3]) {
  const double b = a * a, c = 1.0 / b;
  const double sum = 1.0 + s * c;
  const double inv = 1.0 / sum;
  rho[0] = b * std::log(sum);
  rho[1] = std::max(std::numeric_limits<double>::min(), inv);
  rho[2] = -c * (inv * inv);
}

// -------------------------------------------------------------------------------------------------
// Evaluator: cost, corrected residuals, corrected + column-scaled Jacobian, tangent gradient
// (program_evaluator.h + residual_block.cc + corrector.cc)
// -------------------------------------------------------------------------------------------------
struct Evaluator {
  const Program& g;
  double loss_a;
  int nthreads;
  oracle_block_fn fn;
  // Jacobian storage (Ceres block layout, 416 B per observation with all three blocks)
  std::vector<double> Jc, Jv, Jp, Jk;  // Jk: constraints, [K][2][3] (point1 | point2)
  std::vector<double> r;               // 2N + K   (streaming: K, the constraint residuals only)
  int64_t num_jac_evals = 0, num_cost_evals = 0;
  int64_t num_block_recomputes = 0;    // streaming: passes over all reprojection blocks (autodiff re-evaluations)

  // ---- streaming ("Jacobian-free", block-recompute) mode --------------------------------------------------------
  // Ceres stores the block-sparse Jacobian (416 B per observation: 41.6 GB at the 1M x 1000 scene). In streaming mode
  // nothing O(N) is stored: the evaluator keeps the linearisation point and the column scale, and the corrected,
  // scaled block rows of ONE point's observations are recomputed into a thread-local chunk wherever the stored
  // Jacobian would be read (evaluation, Schur elimination, back-substitution / model cost change). Same arithmetic
  // per block; only the summation order of the camera / pose sums differs (by points instead of by input index).
  bool streaming = false;
  std::vector<double> lin_camera, lin_views, lin_points;  // point at which the Jacobian is defined
  std::vector<double> col_scale;                          // Jacobi scale applied on recompute (tangent layout)
  std::vector<double> cn2;                                // squared column norms of the corrected, unscaled Jacobian
  struct Chunk {
    std::vector<double> jc, jv, jp, r;
    int64_t b0 = 0;
  };
  struct Rows {
    const double *jc, *jv, *jp, *r;
  };

  Evaluator(const Program& prog, double a, int nt, oracle_block_fn f) : g(prog), loss_a(a), nthreads(nt), fn(f) {}

  void alloc_jacobian() {
    if (streaming && !g.has_points) streaming = false;  // nothing to chunk by: the stored form is small anyway
    Jk.resize((size_t)g.K * 6);
    if (streaming) {
      r.resize((size_t)g.K);
      col_scale.assign((size_t)g.n_t, 1.0);
      return;
    }
    Jc.resize((size_t)g.N * 34);
    if (g.has_views) Jv.resize((size_t)g.N * 12);
    if (g.has_points) Jp.resize((size_t)g.N * 6);
    r.resize((size_t)g.N * 2 + g.K);
  }
  double& rk(int k) { return streaming ? r[(size_t)k] : r[(size_t)2 * g.N + k]; }
  double rk(int k) const { return streaming ? r[(size_t)k] : r[(size_t)2 * g.N + k]; }

  // corrected (and, if `scaled`, column-scaled) block rows of every observation of point p at the linearisation
  // point; returns the point's share of the cost
  double fill_chunk(int p, bool scaled, Chunk& ch) const {
    const int64_t b0 = g.pt_begin[p], b1 = g.pt_begin[p + 1];
    const size_t m = (size_t)(b1 - b0);
    ch.b0 = b0;
    if (ch.jc.size() < m * 34) {
      ch.jc.resize(m * 34);
      ch.jv.resize(m * 12);
      ch.jp.resize(m * 6);
      ch.r.resize(m * 2);
    }
    double cost = 0.0;
    for (int64_t e = b0; e < b1; ++e) {
      const int64_t i = g.pt_obs[e];
      const size_t k = (size_t)(e - b0);
      double* jc = &ch.jc[k * 34];
      double* jv = g.has_views ? &ch.jv[k * 12] : nullptr;
      double* jp = g.has_points ? &ch.jp[k * 6] : nullptr;
      double rr[2];
      block_autodiff(g, i, lin_camera.data(), lin_views.data(), lin_points.data(), rr, jc, jv, jp, fn);
      const double s = rr[0] * rr[0] + rr[1] * rr[1];
      if (g.cfg.robust) {
        double rho[3];
        cauchy(loss_a, s, rho);
        cost += 0.5 * rho[0];
        const double sq = std::sqrt(rho[1]);
        if (s == 0.0 || rho[2] <= 0.0) {
          for (int j = 0; j < 34; ++j) jc[j] *= sq;
          if (jv)
            for (int j = 0; j < 12; ++j) jv[j] *= sq;
          if (jp)
            for (int j = 0; j < 6; ++j) jp[j] *= sq;
          rr[0] *= sq;
          rr[1] *= sq;
        }
      } else {
        cost += 0.5 * s;
      }
      ch.r[2 * k] = rr[0];
      ch.r[2 * k + 1] = rr[1];
      if (scaled) {
        const double* sc = col_scale.data();
        for (int j = 0; j < g.n_cam_t; ++j) {
          const int c = g.cam_cols[j];
          jc[c] *= sc[j];
          jc[17 + c] *= sc[j];
        }
        if (jv) {
          const int o = g.off_view(g.view_slot[g.fidx[i]]);
          for (int j = 0; j < 6; ++j) {
            jv[j] *= sc[(size_t)o + j];
            jv[6 + j] *= sc[(size_t)o + j];
          }
        }
        if (jp) {
          const int o = g.off_point(g.point_slot[g.pidx[i]]);
          for (int j = 0; j < 3; ++j) {
            jp[j] *= sc[(size_t)o + j];
            jp[3 + j] *= sc[(size_t)o + j];
          }
        }
      }
    }
    return cost;
  }
  // block rows of the e-th observation in point order (e indexes pt_obs)
  inline Rows rows(int64_t e, const Chunk* ch) const {
    if (!streaming) {
      const size_t i = (size_t)g.pt_obs[e];
      return Rows{&Jc[i * 34], g.has_views ? &Jv[i * 12] : nullptr, g.has_points ? &Jp[i * 6] : nullptr, &r[2 * i]};
    }
    const size_t k = (size_t)(e - ch->b0);
    return Rows{&ch->jc[k * 34], g.has_views ? &ch->jv[k * 12] : nullptr, g.has_points ? &ch->jp[k * 6] : nullptr,
                &ch->r[2 * k]};
  }

  // streaming evaluate: cost, tangent gradient and cn2 in one pass over the points
  double evaluate_streaming(const double* camera, const double* views, const double* points, std::vector<double>& grad) {
    lin_camera.assign(camera, camera + 17);
    lin_views.assign(views, views + (size_t)6 * g.F);
    lin_points.assign(points, points + (size_t)3 * g.P);
    ++num_block_recomputes;
    const int nv = (int)g.active_frames.size();
    const int ncv = g.n_cam_t + 6 * nv;
    const int ne = (int)g.active_points.size();
    std::vector<std::vector<double>> gth((size_t)nthreads, std::vector<double>((size_t)2 * ncv, 0.0));
    grad.assign((size_t)g.n_t, 0.0);
    cn2.assign((size_t)g.n_t, 0.0);
    double cost = 0.0;
#pragma omp parallel num_threads(nthreads)
    {
      std::vector<double>& gl = gth[(size_t)omp_get_thread_num()];
      double* nl = gl.data() + ncv;
      Chunk ch;
      double lcost = 0.0;
#pragma omp for schedule(static)
      for (int s = 0; s < ne; ++s) {
        const int p = g.active_points[s];
        lcost += fill_chunk(p, false, ch);
        double gp[3] = {0, 0, 0}, np_[3] = {0, 0, 0};
        for (int64_t e = g.pt_begin[p]; e < g.pt_begin[p + 1]; ++e) {
          const Rows w = rows(e, &ch);
          for (int j = 0; j < g.n_cam_t; ++j) {
            const int c = g.cam_cols[j];
            gl[j] += w.jc[c] * w.r[0] + w.jc[17 + c] * w.r[1];
            nl[j] += w.jc[c] * w.jc[c] + w.jc[17 + c] * w.jc[17 + c];
          }
          if (w.jv) {
            const int o = g.off_view(g.view_slot[g.fidx[g.pt_obs[e]]]);
            for (int j = 0; j < 6; ++j) {
              gl[o + j] += w.jv[j] * w.r[0] + w.jv[6 + j] * w.r[1];
              nl[o + j] += w.jv[j] * w.jv[j] + w.jv[6 + j] * w.jv[6 + j];
            }
          }
          for (int j = 0; j < 3; ++j) {
            gp[j] += w.jp[j] * w.r[0] + w.jp[3 + j] * w.r[1];
            np_[j] += w.jp[j] * w.jp[j] + w.jp[3 + j] * w.jp[3 + j];
          }
        }
        for (int j = 0; j < 3; ++j) {
          grad[(size_t)g.off_point(s) + j] = gp[j];
          cn2[(size_t)g.off_point(s) + j] = np_[j];
        }
      }
#pragma omp atomic
      cost += lcost;
    }
    for (int t = 0; t < nthreads; ++t)
      for (int j = 0; j < ncv; ++j) {
        grad[j] += gth[t][j];
        cn2[j] += gth[t][(size_t)ncv + j];
      }
    for (int k = 0; k < g.K; ++k) {
      typedef Dual<6> D;
      D p1[3], p2[3], res;
      for (int j = 0; j < 3; ++j) {
        p1[j] = D(points[3 * g.c1[k] + j], j);
        p2[j] = D(points[3 * g.c2[k] + j], 3 + j);
      }
      distance_residual<D>(g.cdist[k], g.csigma[k], p1, p2, &res);
      for (int j = 0; j < 6; ++j) Jk[(size_t)6 * k + j] = res.v[j];
      rk(k) = res.a;
      cost += 0.5 * res.a * res.a;
      const int o1 = g.off_point(g.point_slot[g.c1[k]]), o2 = g.off_point(g.point_slot[g.c2[k]]);
      for (int j = 0; j < 3; ++j) {
        grad[(size_t)o1 + j] += res.v[j] * res.a;
        grad[(size_t)o2 + j] += res.v[3 + j] * res.a;
        cn2[(size_t)o1 + j] += res.v[j] * res.v[j];
        cn2[(size_t)o2 + j] += res.v[3 + j] * res.v[3 + j];
      }
    }
    std::fill(col_scale.begin(), col_scale.end(), 1.0);  // a fresh evaluation is unscaled until scale_columns()
    return cost;
  }

  double cost_only(const double* camera, const double* views, const double* points) {
    ++num_cost_evals;
    double cost = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : cost) num_threads(nthreads)
    for (int64_t i = 0; i < g.N; ++i) {
      double rr[2];
      block_value(g, i, camera, views, points, rr, fn);
      const double s = rr[0] * rr[0] + rr[1] * rr[1];
      if (g.cfg.robust) {
        double rho[3];
        cauchy(loss_a, s, rho);
        cost += 0.5 * rho[0];
      } else {
        cost += 0.5 * s;
      }
    }
    for (int k = 0; k < g.K; ++k) {
      double rr;
      distance_residual<double>(g.cdist[k], g.csigma[k], points + 3 * g.c1[k], points + 3 * g.c2[k], &rr);
      cost += 0.5 * rr * rr;
    }
    return cost;
  }

  // Fills r, J (corrected, NOT yet column-scaled) and the tangent gradient.
  double evaluate(const double* camera, const double* views, const double* points, std::vector<double>& grad) {
    ++num_jac_evals;
    if (streaming) return evaluate_streaming(camera, views, points, grad);
    double cost = 0.0;
    const int nv = (int)g.active_frames.size();
    const int ncv = g.n_cam_t + 6 * nv;  // camera + views part of the gradient is accumulated per thread
    std::vector<std::vector<double>> gth((size_t)nthreads, std::vector<double>((size_t)ncv, 0.0));
#pragma omp parallel num_threads(nthreads)
    {
      std::vector<double>& gl = gth[(size_t)omp_get_thread_num()];
      double lcost = 0.0;
#pragma omp for schedule(static)
      for (int64_t i = 0; i < g.N; ++i) {
        double* jc = &Jc[(size_t)i * 34];
        double* jv = g.has_views ? &Jv[(size_t)i * 12] : nullptr;
        double* jp = g.has_points ? &Jp[(size_t)i * 6] : nullptr;
        double rr[2];
        block_autodiff(g, i, camera, views, points, rr, jc, jv, jp, fn);
        const double s = rr[0] * rr[0] + rr[1] * rr[1];
        if (g.cfg.robust) {
          double rho[3];
          cauchy(loss_a, s, rho);
          lcost += 0.5 * rho[0];
          // Corrector: rho'' < 0 for Cauchy -> scale residual and Jacobian by sqrt(rho') (corrector.cc)
          const double sq = std::sqrt(rho[1]);
          if (s == 0.0 || rho[2] <= 0.0) {
            for (int j = 0; j < 34; ++j) jc[j] *= sq;
            if (jv)
              for (int j = 0; j < 12; ++j) jv[j] *= sq;
            if (jp)
              for (int j = 0; j < 6; ++j) jp[j] *= sq;
            rr[0] *= sq;
            rr[1] *= sq;
          }
        } else {
          lcost += 0.5 * s;
        }
        r[(size_t)2 * i] = rr[0];
        r[(size_t)2 * i + 1] = rr[1];
        for (int j = 0; j < g.n_cam_t; ++j) {
          const int c = g.cam_cols[j];
          gl[j] += jc[c] * rr[0] + jc[17 + c] * rr[1];
        }
        if (jv) {
          const int o = g.off_view(g.view_slot[g.fidx[i]]);
          for (int j = 0; j < 6; ++j) gl[o + j] += jv[j] * rr[0] + jv[6 + j] * rr[1];
        }
      }
#pragma omp atomic
      cost += lcost;
    }
    grad.assign((size_t)g.n_t, 0.0);
    for (int t = 0; t < nthreads; ++t)
      for (int j = 0; j < ncv; ++j) grad[j] += gth[t][j];
    if (g.has_points) {
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads)
      for (int s = 0; s < (int)g.active_points.size(); ++s) {
        const int p = g.active_points[s];
        double gp[3] = {0, 0, 0};
        for (int64_t e = g.pt_begin[p]; e < g.pt_begin[p + 1]; ++e) {
          const int64_t i = g.pt_obs[e];
          const double* jp = &Jp[(size_t)i * 6];
          for (int j = 0; j < 3; ++j) gp[j] += jp[j] * r[(size_t)2 * i] + jp[3 + j] * r[(size_t)2 * i + 1];
        }
        for (int j = 0; j < 3; ++j) grad[(size_t)g.off_point(s) + j] = gp[j];
      }
    }
    // distance constraints: AutoDiffCostFunction<.., 1, 3, 3>, no loss (:923)
    for (int k = 0; k < g.K; ++k) {
      typedef Dual<6> D;
      D p1[3], p2[3], res;
      for (int j = 0; j < 3; ++j) {
        p1[j] = D(points[3 * g.c1[k] + j], j);
        p2[j] = D(points[3 * g.c2[k] + j], 3 + j);
      }
      distance_residual<D>(g.cdist[k], g.csigma[k], p1, p2, &res);
      for (int j = 0; j < 6; ++j) Jk[(size_t)6 * k + j] = res.v[j];
      rk(k) = res.a;
      cost += 0.5 * res.a * res.a;
      const int o1 = g.off_point(g.point_slot[g.c1[k]]), o2 = g.off_point(g.point_slot[g.c2[k]]);
      for (int j = 0; j < 3; ++j) {
        grad[(size_t)o1 + j] += res.v[j] * res.a;
        grad[(size_t)o2 + j] += res.v[3 + j] * res.a;
      }
    }
    return cost;
  }

  // squared column norms of the stored Jacobian (tangent columns)
  void squared_column_norms(std::vector<double>& out) const {
    if (streaming) {  // |J s|^2 column-wise = s^2 |J|^2 (J unscaled at the linearisation point)
      out.resize((size_t)g.n_t);
      for (int j = 0; j < g.n_t; ++j) out[j] = cn2[j] * col_scale[j] * col_scale[j];
      return;
    }
    const int nv = (int)g.active_frames.size();
    const int ncv = g.n_cam_t + 6 * nv;
    out.assign((size_t)g.n_t, 0.0);
    std::vector<std::vector<double>> th((size_t)nthreads, std::vector<double>((size_t)ncv, 0.0));
#pragma omp parallel num_threads(nthreads)
    {
      std::vector<double>& l = th[(size_t)omp_get_thread_num()];
#pragma omp for schedule(static)
      for (int64_t i = 0; i < g.N; ++i) {
        const double* jc = &Jc[(size_t)i * 34];
        for (int j = 0; j < g.n_cam_t; ++j) {
          const int c = g.cam_cols[j];
          l[j] += jc[c] * jc[c] + jc[17 + c] * jc[17 + c];
        }
        if (g.has_views) {
          const double* jv = &Jv[(size_t)i * 12];
          const int o = g.off_view(g.view_slot[g.fidx[i]]);
          for (int j = 0; j < 6; ++j) l[o + j] += jv[j] * jv[j] + jv[6 + j] * jv[6 + j];
        }
      }
    }
    for (int t = 0; t < nthreads; ++t)
      for (int j = 0; j < ncv; ++j) out[j] += th[t][j];
    if (g.has_points) {
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads)
      for (int s = 0; s < (int)g.active_points.size(); ++s) {
        const int p = g.active_points[s];
        double a[3] = {0, 0, 0};
        for (int64_t e = g.pt_begin[p]; e < g.pt_begin[p + 1]; ++e) {
          const double* jp = &Jp[(size_t)g.pt_obs[e] * 6];
          for (int j = 0; j < 3; ++j) a[j] += jp[j] * jp[j] + jp[3 + j] * jp[3 + j];
        }
        for (int j = 0; j < 3; ++j) out[(size_t)g.off_point(s) + j] = a[j];
      }
      for (int k = 0; k < g.K; ++k) {
        const int o1 = g.off_point(g.point_slot[g.c1[k]]), o2 = g.off_point(g.point_slot[g.c2[k]]);
        for (int j = 0; j < 3; ++j) {
          out[(size_t)o1 + j] += Jk[(size_t)6 * k + j] * Jk[(size_t)6 * k + j];
          out[(size_t)o2 + j] += Jk[(size_t)6 * k + 3 + j] * Jk[(size_t)6 * k + 3 + j];
        }
      }
    }
  }

  // J <- J * diag(scale)   (jacobian_->ScaleColumns)
  void scale_columns(const std::vector<double>& sc) {
    if (streaming) {  // applied when a chunk is recomputed
      col_scale = sc;
      for (int k = 0; k < g.K; ++k) {
        const int o1 = g.off_point(g.point_slot[g.c1[k]]), o2 = g.off_point(g.point_slot[g.c2[k]]);
        for (int j = 0; j < 3; ++j) {
          Jk[(size_t)6 * k + j] *= sc[(size_t)o1 + j];
          Jk[(size_t)6 * k + 3 + j] *= sc[(size_t)o2 + j];
        }
      }
      return;
    }
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int64_t i = 0; i < g.N; ++i) {
      double* jc = &Jc[(size_t)i * 34];
      for (int j = 0; j < g.n_cam_t; ++j) {
        const int c = g.cam_cols[j];
        jc[c] *= sc[j];
        jc[17 + c] *= sc[j];
      }
      if (g.has_views) {
        double* jv = &Jv[(size_t)i * 12];
        const int o = g.off_view(g.view_slot[g.fidx[i]]);
        for (int j = 0; j < 6; ++j) {
          jv[j] *= sc[(size_t)o + j];
          jv[6 + j] *= sc[(size_t)o + j];
        }
      }
      if (g.has_points) {
        double* jp = &Jp[(size_t)i * 6];
        const int o = g.off_point(g.point_slot[g.pidx[i]]);
        for (int j = 0; j < 3; ++j) {
          jp[j] *= sc[(size_t)o + j];
          jp[3 + j] *= sc[(size_t)o + j];
        }
      }
    }
    for (int k = 0; k < g.K; ++k) {
      const int o1 = g.off_point(g.point_slot[g.c1[k]]), o2 = g.off_point(g.point_slot[g.c2[k]]);
      for (int j = 0; j < 3; ++j) {
        Jk[(size_t)6 * k + j] *= sc[(size_t)o1 + j];
        Jk[(size_t)6 * k + 3 + j] *= sc[(size_t)o2 + j];
      }
    }
  }

  // m = J * step  and  -m.(r + m/2)   (trust_region_minimizer.cc, ComputeTrustRegionStep)
  double model_cost_change(const std::vector<double>& step) const {
    double acc = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : acc) num_threads(nthreads)
    for (int64_t i = 0; i < g.N; ++i) {
      double m[2] = {0, 0};
      const double* jc = &Jc[(size_t)i * 34];
      for (int j = 0; j < g.n_cam_t; ++j) {
        const int c = g.cam_cols[j];
        m[0] += jc[c] * step[j];
        m[1] += jc[17 + c] * step[j];
      }
      if (g.has_views) {
        const double* jv = &Jv[(size_t)i * 12];
        const int o = g.off_view(g.view_slot[g.fidx[i]]);
        for (int j = 0; j < 6; ++j) {
          m[0] += jv[j] * step[(size_t)o + j];
          m[1] += jv[6 + j] * step[(size_t)o + j];
        }
      }
      if (g.has_points) {
        const double* jp = &Jp[(size_t)i * 6];
        const int o = g.off_point(g.point_slot[g.pidx[i]]);
        for (int j = 0; j < 3; ++j) {
          m[0] += jp[j] * step[(size_t)o + j];
          m[1] += jp[3 + j] * step[(size_t)o + j];
        }
      }
      acc += m[0] * (r[(size_t)2 * i] + m[0] / 2.0) + m[1] * (r[(size_t)2 * i + 1] + m[1] / 2.0);
    }
    for (int k = 0; k < g.K; ++k) {
      const int o1 = g.off_point(g.point_slot[g.c1[k]]), o2 = g.off_point(g.point_slot[g.c2[k]]);
      double m = 0;
      for (int j = 0; j < 3; ++j)
        m += Jk[(size_t)6 * k + j] * step[(size_t)o1 + j] + Jk[(size_t)6 * k + 3 + j] * step[(size_t)o2 + j];
      acc += m * (rk(k) + m / 2.0);
    }
    return -acc;
  }
};

// -------------------------------------------------------------------------------------------------
// Dense Cholesky (lower, row-major, in place) — stands in for Eigen::LLT used by Ceres' DENSE_SCHUR
// -------------------------------------------------------------------------------------------------
static bool llt_lower(double* A, int n, int nthreads) {
  const int nb = 64;
  for (int k0 = 0; k0 < n; k0 += nb) {
    const int k1 = std::min(n, k0 + nb);
    // unblocked factorisation of the diagonal block
    for (int j = k0; j < k1; ++j) {
      double d = A[(size_t)j * n + j];
      for (int t = k0; t < j; ++t) d -= A[(size_t)j * n + t] * A[(size_t)j * n + t];
      if (!(d > 0.0)) return false;  // Eigen::NumericalIssue -> LINEAR_SOLVER_FAILURE
      d = std::sqrt(d);
      A[(size_t)j * n + j] = d;
      for (int i = j + 1; i < k1; ++i) {
        double s = A[(size_t)i * n + j];
        for (int t = k0; t < j; ++t) s -= A[(size_t)i * n + t] * A[(size_t)j * n + t];
        A[(size_t)i * n + j] = s / d;
      }
    }
    // panel: rows below the block
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int i = k1; i < n; ++i) {
      for (int j = k0; j < k1; ++j) {
        double s = A[(size_t)i * n + j];
        for (int t = k0; t < j; ++t) s -= A[(size_t)i * n + t] * A[(size_t)j * n + t];
        A[(size_t)i * n + j] = s / A[(size_t)j * n + j];
      }
    }
    // trailing update (lower triangle only)
#pragma omp parallel for schedule(dynamic, 8) num_threads(nthreads)
    for (int i = k1; i < n; ++i) {
      const double* li = &A[(size_t)i * n + k0];
      for (int j = k1; j <= i; ++j) {
        const double* lj = &A[(size_t)j * n + k0];
        double s = 0.0;
        for (int t = 0; t < k1 - k0; ++t) s += li[t] * lj[t];
        A[(size_t)i * n + j] -= s;
      }
    }
  }
  return true;
}
static void llt_solve(const double* L, int n, double* b) {
  for (int i = 0; i < n; ++i) {
    double s = b[i];
    for (int t = 0; t < i; ++t) s -= L[(size_t)i * n + t] * b[t];
    b[i] = s / L[(size_t)i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int t = i + 1; t < n; ++t) s -= L[(size_t)t * n + i] * b[t];
    b[i] = s / L[(size_t)i * n + i];
  }
}
// InvertPSDMatrix<3>: LLT of the 3x3 block, solve against identity (small_blas / invert_psd_matrix.h)
static bool invert_psd3(const double m[9], double inv[9]) {
  double L[9] = {0};
  for (int j = 0; j < 3; ++j) {
    double d = m[3 * j + j];
    for (int t = 0; t < j; ++t) d -= L[3 * j + t] * L[3 * j + t];
    if (!(d > 0.0)) return false;
    d = std::sqrt(d);
    L[3 * j + j] = d;
    for (int i = j + 1; i < 3; ++i) {
      double s = m[3 * i + j];
      for (int t = 0; t < j; ++t) s -= L[3 * i + t] * L[3 * j + t];
      L[3 * i + j] = s / d;
    }
  }
  for (int c = 0; c < 3; ++c) {
    double b[3] = {c == 0 ? 1.0 : 0.0, c == 1 ? 1.0 : 0.0, c == 2 ? 1.0 : 0.0};
    llt_solve(L, 3, b);
    for (int i = 0; i < 3; ++i) inv[3 * i + c] = b[i];
  }
  return true;
}

// -------------------------------------------------------------------------------------------------
// DENSE_SCHUR: eliminate the (uncoupled) point blocks, dense LLT of the reduced system, back-substitute
// (schur_eliminator_impl.h, schur_complement_solver.cc). Solves  min |J y - r|^2 + |D y|^2.
// -------------------------------------------------------------------------------------------------
struct SchurSolver {
  const Program& g;
  const Evaluator& ev;
  int nthreads;
  std::vector<double> S, rhs;
  double model_cost_change = 0.0;  // streaming mode: -m.(r + m/2), m = J (-y), formed during back-substitution

  SchurSolver(const Program& prog, const Evaluator& e, int nt) : g(prog), ev(e), nthreads(nt) {}

  // f-blocks of one observation row: reduced offsets + pointers into the Jacobian
  struct RowBlocks {
    int nblk;
    int off[3];
    int width[3];
    const double* j[3];
    int stride[3];
  };
  inline void row_blocks(int64_t i, const Evaluator::Rows& w, RowBlocks& rb) const {
    rb.nblk = 0;
    rb.off[rb.nblk] = 0;
    rb.width[rb.nblk] = g.n_cam_t;
    rb.j[rb.nblk] = w.jc;
    rb.stride[rb.nblk] = 17;
    rb.nblk++;
    if (g.has_views) {
      rb.off[rb.nblk] = g.red_view(g.view_slot[g.fidx[i]]);
      rb.width[rb.nblk] = 6;
      rb.j[rb.nblk] = w.jv;
      rb.stride[rb.nblk] = 6;
      rb.nblk++;
    }
    if (g.has_points && g.coupled[g.pidx[i]]) {
      rb.off[rb.nblk] = g.red_point(g.coupled_slot[g.pidx[i]]);
      rb.width[rb.nblk] = 3;
      rb.j[rb.nblk] = w.jp;
      rb.stride[rb.nblk] = 3;
      rb.nblk++;
    }
  }
  inline Evaluator::Rows rows_of_input(int64_t i) const {  // stored mode only: rows of input observation i
    return Evaluator::Rows{&ev.Jc[(size_t)i * 34], g.has_views ? &ev.Jv[(size_t)i * 12] : nullptr,
                           g.has_points ? &ev.Jp[(size_t)i * 6] : nullptr, &ev.r[(size_t)2 * i]};
  }
  inline double jval(const RowBlocks& rb, int b, int row, int c) const {
    // camera block: tangent column c -> ambient column
    return b == 0 ? rb.j[0][17 * row + g.cam_cols[c]] : rb.j[b][rb.stride[b] * row + c];
  }

  bool solve(const std::vector<double>& D, std::vector<double>& y) {
    const int n = g.n_red;
    const int nv = (int)g.active_frames.size();
    const size_t nn = (size_t)n * n;
    std::vector<std::vector<double>> Sth((size_t)nthreads), Rth((size_t)nthreads);
    for (int t = 0; t < nthreads; ++t) {
      Sth[t].assign(nn, 0.0);
      Rth[t].assign((size_t)n, 0.0);
    }
    const int ne = g.has_points ? (int)g.active_points.size() : 0;
    std::vector<double> inv_ete((size_t)ne * 9, 0.0);
    bool ok = true;
    if (ev.streaming) const_cast<Evaluator&>(ev).num_block_recomputes += 2;  // elimination + back-substitution

#pragma omp parallel num_threads(nthreads)
    {
      double* Sl = Sth[(size_t)omp_get_thread_num()].data();
      double* Rl = Rth[(size_t)omp_get_thread_num()].data();
      std::vector<double> buf((size_t)3 * n);  // E^T F for the current chunk, 3 x n (dense row, sparse use)
      std::vector<int> touched;
      std::vector<char> mark((size_t)n, 0);
      Evaluator::Chunk chunk;

      // F^T F and F^T r of one row into the thread-local lower triangle
      auto add_row = [&](const RowBlocks& rb, const double rr[2], bool add_rhs) {
        for (int a = 0; a < rb.nblk; ++a)
          for (int b = 0; b <= a; ++b)
            for (int ca = 0; ca < rb.width[a]; ++ca) {
              const double xa0 = jval(rb, a, 0, ca), xa1 = jval(rb, a, 1, ca);
              double* dst = Sl + (size_t)(rb.off[a] + ca) * n + rb.off[b];
              const int wb = (a == b) ? ca + 1 : rb.width[b];
              for (int cb = 0; cb < wb; ++cb) dst[cb] += xa0 * jval(rb, b, 0, cb) + xa1 * jval(rb, b, 1, cb);
            }
        if (add_rhs)
          for (int a = 0; a < rb.nblk; ++a)
            for (int ca = 0; ca < rb.width[a]; ++ca)
              Rl[rb.off[a] + ca] += jval(rb, a, 0, ca) * rr[0] + jval(rb, a, 1, ca) * rr[1];
      };

      if (!g.has_points) {
        // no e-blocks: the "reduced" system is the whole normal system
#pragma omp for schedule(static)
        for (int64_t i = 0; i < g.N; ++i) {
          RowBlocks rb;
          const Evaluator::Rows w = rows_of_input(i);
          row_blocks(i, w, rb);
          add_row(rb, w.r, true);
        }
      } else {
#pragma omp for schedule(dynamic, 64)
        for (int s = 0; s < ne; ++s) {
          const int p = g.active_points[s];
          const int64_t b0 = g.pt_begin[p], b1 = g.pt_begin[p + 1];
          if (ev.streaming) ev.fill_chunk(p, true, chunk);
          if (g.coupled[p]) {  // rows without an e-block (NoEBlockRowsUpdate)
            for (int64_t e = b0; e < b1; ++e) {
              RowBlocks rb;
              const Evaluator::Rows w = ev.rows(e, &chunk);
              row_blocks(g.pt_obs[e], w, rb);
              add_row(rb, w.r, true);
            }
            continue;
          }
          // chunk of e-block p: ete = E^T E + D_e^2, g_e = E^T r, buffer = E^T F
          const int op = g.off_point(s);
          double ete[9] = {D[op] * D[op], 0, 0, 0, D[op + 1] * D[op + 1], 0, 0, 0, D[op + 2] * D[op + 2]};
          double ge[3] = {0, 0, 0};
          touched.clear();
          for (int64_t e = b0; e < b1; ++e) {
            const int64_t i = g.pt_obs[e];
            const Evaluator::Rows w = ev.rows(e, &chunk);
            const double* jp = w.jp;
            const double* rr = w.r;
            for (int a = 0; a < 3; ++a) {
              for (int b = 0; b < 3; ++b) ete[3 * a + b] += jp[a] * jp[b] + jp[3 + a] * jp[3 + b];
              ge[a] += jp[a] * rr[0] + jp[3 + a] * rr[1];
            }
            RowBlocks rb;
            row_blocks(i, w, rb);
            for (int bb = 0; bb < rb.nblk; ++bb)
              for (int c = 0; c < rb.width[bb]; ++c) {
                const int col = rb.off[bb] + c;
                if (!mark[col]) {
                  mark[col] = 1;
                  touched.push_back(col);
                  buf[col] = buf[(size_t)n + col] = buf[(size_t)2 * n + col] = 0.0;
                }
                const double f0 = jval(rb, bb, 0, c), f1 = jval(rb, bb, 1, c);
                for (int a = 0; a < 3; ++a) buf[(size_t)a * n + col] += jp[a] * f0 + jp[3 + a] * f1;
              }
            add_row(rb, rr, false);  // F^T F of the chunk's rows
          }
          double inv[9];
          if (!invert_psd3(ete, inv)) {
#pragma omp atomic write
            ok = false;
            for (int col : touched) mark[col] = 0;
            continue;
          }
          for (int a = 0; a < 9; ++a) inv_ete[(size_t)9 * s + a] = inv[a];
          // rhs += F^T (r - E inv g_e) row by row (UpdateRhs)
          double ig[3];
          for (int a = 0; a < 3; ++a) ig[a] = inv[3 * a] * ge[0] + inv[3 * a + 1] * ge[1] + inv[3 * a + 2] * ge[2];
          for (int64_t e = b0; e < b1; ++e) {
            const int64_t i = g.pt_obs[e];
            const Evaluator::Rows w = ev.rows(e, &chunk);
            const double* jp = w.jp;
            double sj[2];
            for (int row = 0; row < 2; ++row)
              sj[row] = w.r[row] - (jp[3 * row] * ig[0] + jp[3 * row + 1] * ig[1] + jp[3 * row + 2] * ig[2]);
            RowBlocks rb;
            row_blocks(i, w, rb);
            for (int bb = 0; bb < rb.nblk; ++bb)
              for (int c = 0; c < rb.width[bb]; ++c)
                Rl[rb.off[bb] + c] += jval(rb, bb, 0, c) * sj[0] + jval(rb, bb, 1, c) * sj[1];
          }
          // lhs -= buffer^T inv buffer (ChunkOuterProduct), lower triangle
          std::sort(touched.begin(), touched.end());
          for (size_t ia = 0; ia < touched.size(); ++ia) {
            const int ca = touched[ia];
            double w[3];
            for (int a = 0; a < 3; ++a)
              w[a] = inv[a] * buf[ca] + inv[3 + a] * buf[(size_t)n + ca] + inv[6 + a] * buf[(size_t)2 * n + ca];
            double* dst = Sl + (size_t)ca * n;
            for (size_t ib = 0; ib <= ia; ++ib) {
              const int cb = touched[ib];
              dst[cb] -= w[0] * buf[cb] + w[1] * buf[(size_t)n + cb] + w[2] * buf[(size_t)2 * n + cb];
            }
          }
          for (int col : touched) mark[col] = 0;
        }
      }
    }
    if (!ok) return false;
    S.assign(nn, 0.0);
    rhs.assign((size_t)n, 0.0);
    for (int t = 0; t < nthreads; ++t) {
      const double* s = Sth[t].data();
#pragma omp parallel for schedule(static) num_threads(nthreads)
      for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) S[(size_t)i * n + j] += s[(size_t)i * n + j];
      for (int i = 0; i < n; ++i) rhs[i] += Rth[t][i];
      std::vector<double>().swap(Sth[t]);  // release early: n^2 doubles per thread
    }
    // constraint rows (no e-block): F = [coupled point1 | coupled point2]
    for (int k = 0; k < g.K; ++k) {
      const int o[2] = {g.red_point(g.coupled_slot[g.c1[k]]), g.red_point(g.coupled_slot[g.c2[k]])};
      const double* jk = &ev.Jk[(size_t)6 * k];
      const double rk = ev.rk(k);
      for (int a = 0; a < 2; ++a)
        for (int ca = 0; ca < 3; ++ca) {
          rhs[o[a] + ca] += jk[3 * a + ca] * rk;
          for (int b = 0; b < 2; ++b)
            for (int cb = 0; cb < 3; ++cb) {
              const int ri = o[a] + ca, ci = o[b] + cb;
              if (ci <= ri) S[(size_t)ri * n + ci] += jk[3 * a + ca] * jk[3 * b + cb];
            }
        }
    }
    // D_f^2 on the diagonal of the reduced system
    for (int j = 0; j < g.n_cam_t + 6 * nv; ++j) S[(size_t)j * n + j] += D[j] * D[j];
    for (size_t c = 0; c < g.coupled_points.size(); ++c) {
      const int ot = g.off_point(g.point_slot[g.coupled_points[c]]);
      const int orr = g.red_point((int)c);
      for (int j = 0; j < 3; ++j) S[(size_t)(orr + j) * n + orr + j] += D[ot + j] * D[ot + j];
    }
    if (!llt_lower(S.data(), n, nthreads)) return false;
    llt_solve(S.data(), n, rhs.data());

    // scatter the reduced solution, then back-substitute the eliminated points
    y.assign((size_t)g.n_t, 0.0);
    for (int j = 0; j < g.n_cam_t + 6 * nv; ++j) y[j] = rhs[j];
    for (size_t c = 0; c < g.coupled_points.size(); ++c) {
      const int ot = g.off_point(g.point_slot[g.coupled_points[c]]);
      for (int j = 0; j < 3; ++j) y[(size_t)ot + j] = rhs[g.red_point((int)c) + j];
    }
    double mcc_acc = 0.0;  // streaming: sum over rows of m.(r + m/2) with m = J (-y)
    if (g.has_points) {
#pragma omp parallel num_threads(nthreads) reduction(+ : mcc_acc)
      {
        Evaluator::Chunk chunk;
#pragma omp for schedule(dynamic, 64)
        for (int s = 0; s < ne; ++s) {
          const int p = g.active_points[s];
          if (g.coupled[p] && !ev.streaming) continue;
          const int64_t b0 = g.pt_begin[p], b1 = g.pt_begin[p + 1];
          if (ev.streaming) ev.fill_chunk(p, true, chunk);
          const int op = g.off_point(s);
          if (!g.coupled[p]) {
            double acc[3] = {0, 0, 0};
            for (int64_t e = b0; e < b1; ++e) {
              const int64_t i = g.pt_obs[e];
              const Evaluator::Rows w = ev.rows(e, &chunk);
              RowBlocks rb;
              row_blocks(i, w, rb);
              double sj[2] = {w.r[0], w.r[1]};
              for (int bb = 0; bb < rb.nblk; ++bb)
                for (int c = 0; c < rb.width[bb]; ++c) {
                  const double yy = rhs[rb.off[bb] + c];
                  sj[0] -= jval(rb, bb, 0, c) * yy;
                  sj[1] -= jval(rb, bb, 1, c) * yy;
                }
              const double* jp = w.jp;
              for (int a = 0; a < 3; ++a) acc[a] += jp[a] * sj[0] + jp[3 + a] * sj[1];
            }
            const double* inv = &inv_ete[(size_t)9 * s];
            for (int a = 0; a < 3; ++a)
              y[(size_t)op + a] = inv[3 * a] * acc[0] + inv[3 * a + 1] * acc[1] + inv[3 * a + 2] * acc[2];
          }
          if (ev.streaming) {  // model cost change rows of this point (trust_region_minimizer.cc: model_residuals = J step)
            for (int64_t e = b0; e < b1; ++e) {
              const int64_t i = g.pt_obs[e];
              const Evaluator::Rows w = ev.rows(e, &chunk);
              double m[2] = {0, 0};
              for (int j = 0; j < g.n_cam_t; ++j) {
                const int c = g.cam_cols[j];
                m[0] -= w.jc[c] * y[j];
                m[1] -= w.jc[17 + c] * y[j];
              }
              if (g.has_views) {
                const int o = g.off_view(g.view_slot[g.fidx[i]]);
                for (int j = 0; j < 6; ++j) {
                  m[0] -= w.jv[j] * y[(size_t)o + j];
                  m[1] -= w.jv[6 + j] * y[(size_t)o + j];
                }
              }
              for (int j = 0; j < 3; ++j) {
                m[0] -= w.jp[j] * y[(size_t)op + j];
                m[1] -= w.jp[3 + j] * y[(size_t)op + j];
              }
              mcc_acc += m[0] * (w.r[0] + m[0] / 2.0) + m[1] * (w.r[1] + m[1] / 2.0);
            }
          }
        }
      }
    }
    if (ev.streaming) {
      for (int k = 0; k < g.K; ++k) {
        const int o1 = g.off_point(g.point_slot[g.c1[k]]), o2 = g.off_point(g.point_slot[g.c2[k]]);
        double m = 0;
        for (int j = 0; j < 3; ++j)
          m -= ev.Jk[(size_t)6 * k + j] * y[(size_t)o1 + j] + ev.Jk[(size_t)6 * k + 3 + j] * y[(size_t)o2 + j];
        mcc_acc += m * (ev.rk(k) + m / 2.0);
      }
      model_cost_change = -mcc_acc;
    }
    for (double v : y)
      if (!std::isfinite(v)) return false;
    return true;
  }
};

// -------------------------------------------------------------------------------------------------
// Parameter state helpers (Program::Plus with manifold + bounds; norms over the blocks in the problem)
// -------------------------------------------------------------------------------------------------
struct Params {
  std::vector<double> camera, views, points;
};
static void plus(const Program& g, const Params& x, const std::vector<double>& delta, Params& out) {
  out = x;
  for (int j = 0; j < g.n_cam_t; ++j) out.camera[g.cam_cols[j]] = x.camera[g.cam_cols[j]] + delta[j];
  if (g.constrained)
    for (int j = 0; j < 17; ++j) {
      out.camera[j] = std::max(out.camera[j], g.lower[j]);
      out.camera[j] = std::min(out.camera[j], g.upper[j]);
    }
  for (size_t s = 0; s < g.active_frames.size(); ++s)
    for (int j = 0; j < 6; ++j)
      out.views[(size_t)6 * g.active_frames[s] + j] =
          x.views[(size_t)6 * g.active_frames[s] + j] + delta[(size_t)g.off_view((int)s) + j];
  for (size_t s = 0; s < g.active_points.size(); ++s)
    for (int j = 0; j < 3; ++j)
      out.points[(size_t)3 * g.active_points[s] + j] =
          x.points[(size_t)3 * g.active_points[s] + j] + delta[(size_t)g.off_point((int)s) + j];
}
static double sqnorm_x(const Program& g, const Params& x) {
  double s = 0;
  for (int j = 0; j < 17; ++j) s += x.camera[j] * x.camera[j];
  for (int f : g.active_frames)
    for (int j = 0; j < 6; ++j) s += x.views[(size_t)6 * f + j] * x.views[(size_t)6 * f + j];
  for (int p : g.active_points)
    for (int j = 0; j < 3; ++j) s += x.points[(size_t)3 * p + j] * x.points[(size_t)3 * p + j];
  return s;
}
// |a - b| (2-norm and max-norm) over the blocks in the problem
static void diff_norms(const Program& g, const Params& a, const Params& b, double& n2, double& ninf) {
  double s = 0, m = 0;
  auto acc = [&](double d) {
    s += d * d;
    m = std::max(m, std::fabs(d));
  };
  for (int j = 0; j < 17; ++j) acc(a.camera[j] - b.camera[j]);
  for (int f : g.active_frames)
    for (int j = 0; j < 6; ++j) acc(a.views[(size_t)6 * f + j] - b.views[(size_t)6 * f + j]);
  for (int p : g.active_points)
    for (int j = 0; j < 3; ++j) acc(a.points[(size_t)3 * p + j] - b.points[(size_t)3 * p + j]);
  n2 = std::sqrt(s);
  ninf = m;
}

// -------------------------------------------------------------------------------------------------
// Projected ARMIJO line search with CUBIC interpolation (line_search.cc, polynomial.cc), used by the
// trust-region loop only when the problem has bounds (recalib).
// -------------------------------------------------------------------------------------------------
struct Sample {
  double x = 0, value = 0, gradient = 0;
  bool value_valid = false, gradient_valid = false;
};
// Fit the polynomial through the given samples (values and, where valid, gradients); minimise on [lo, hi].
static double minimize_interpolant(const std::vector<Sample>& smp, double lo, double hi) {
  int ncon = 0;
  for (const Sample& s : smp) ncon += (s.value_valid ? 1 : 0) + (s.gradient_valid ? 1 : 0);
  const int deg = ncon - 1;
  std::vector<double> A((size_t)ncon * ncon, 0.0), b((size_t)ncon, 0.0);
  int row = 0;
  for (const Sample& s : smp) {
    if (s.value_valid) {
      for (int j = 0; j <= deg; ++j) A[(size_t)row * ncon + j] = std::pow(s.x, deg - j);
      b[row++] = s.value;
    }
    if (s.gradient_valid) {
      for (int j = 0; j < deg; ++j) A[(size_t)row * ncon + j] = (deg - j) * std::pow(s.x, deg - j - 1);
      b[row++] = s.gradient;
    }
  }
  // Gaussian elimination with partial pivoting
  for (int c = 0; c < ncon; ++c) {
    int piv = c;
    for (int i = c + 1; i < ncon; ++i)
      if (std::fabs(A[(size_t)i * ncon + c]) > std::fabs(A[(size_t)piv * ncon + c])) piv = i;
    for (int j = 0; j < ncon; ++j) std::swap(A[(size_t)c * ncon + j], A[(size_t)piv * ncon + j]);
    std::swap(b[c], b[piv]);
    const double d = A[(size_t)c * ncon + c];
    if (d == 0.0) continue;
    for (int i = c + 1; i < ncon; ++i) {
      const double f = A[(size_t)i * ncon + c] / d;
      for (int j = c; j < ncon; ++j) A[(size_t)i * ncon + j] -= f * A[(size_t)c * ncon + j];
      b[i] -= f * b[c];
    }
  }
  std::vector<double> coef((size_t)ncon, 0.0);  // highest degree first
  for (int i = ncon - 1; i >= 0; --i) {
    double s = b[i];
    for (int j = i + 1; j < ncon; ++j) s -= A[(size_t)i * ncon + j] * coef[j];
    coef[i] = A[(size_t)i * ncon + i] != 0.0 ? s / A[(size_t)i * ncon + i] : 0.0;
  }
  auto poly = [&](double x) {
    double v = 0;
    for (int j = 0; j <= deg; ++j) v = v * x + coef[j];
    return v;
  };
  auto dpoly = [&](double x) {
    double v = 0;
    for (int j = 0; j < deg; ++j) v = v * x + (deg - j) * coef[j];
    return v;
  };
  // candidates: end points and the stationary points inside (sign changes of p' on a fine grid, bisected)
  double best_x = lo, best_v = poly(lo);
  if (poly(hi) < best_v) {
    best_v = poly(hi);
    best_x = hi;
  }
  const int G = 4096;
  double xa = lo, da = dpoly(lo);
  for (int i = 1; i <= G; ++i) {
    const double xb = lo + (hi - lo) * i / G, db = dpoly(xb);
    if ((da <= 0 && db >= 0) || (da >= 0 && db <= 0)) {
      double l = xa, h = xb, dl = da;
      for (int it = 0; it < 100; ++it) {
        const double m = 0.5 * (l + h), dm = dpoly(m);
        if ((dl <= 0 && dm <= 0) || (dl >= 0 && dm >= 0)) {
          l = m;
          dl = dm;
        } else {
          h = m;
        }
      }
      const double xm = 0.5 * (l + h), vm = poly(xm);
      if (vm < best_v) {
        best_v = vm;
        best_x = xm;
      }
    }
    xa = xb;
    da = db;
  }
  return best_x;
}

}  // namespace lfba_oracle

using namespace lfba_oracle;

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" int oracle_max_threads(void) { return omp_get_max_threads(); }

extern "C" int oracle_eval(const lfba_problem* pb, const double* camera, const double* views,
                           const double* points, double* residuals, double* jac_camera, double* jac_view,
                           double* jac_point, double* cost, lfba_reproj_stats* stats, double inlier_threshold,
                           int num_threads, oracle_block_fn fn) {
  Program g;
  if (!build_program(pb, camera, g)) return LFBA_INVALID_ARGUMENT;
  const int nt = num_threads > 0 ? num_threads : omp_get_max_threads();
  const bool want_j = jac_camera || jac_view || jac_point;
  double c = 0, sx2 = 0, sy2 = 0, mx = 0, my = 0;
  int64_t inl = 0;
#pragma omp parallel for schedule(static) num_threads(nt) reduction(+ : c, sx2, sy2, inl) reduction(max : mx, my)
  for (int64_t i = 0; i < g.N; ++i) {
    double rr[2], jc[34], jv[12], jp[6];
    if (want_j) {
      for (int j = 0; j < 12; ++j) jv[j] = 0;
      for (int j = 0; j < 6; ++j) jp[j] = 0;
      block_autodiff(g, i, camera, views, points, rr, jc, g.has_views ? jv : nullptr,
                     g.has_points ? jp : nullptr, fn);
      if (jac_camera) std::memcpy(jac_camera + (size_t)i * 34, jc, sizeof(jc));
      if (jac_view) std::memcpy(jac_view + (size_t)i * 12, jv, sizeof(jv));
      if (jac_point) std::memcpy(jac_point + (size_t)i * 6, jp, sizeof(jp));
    } else {
      block_value(g, i, camera, views, points, rr, fn);
    }
    if (residuals) {
      residuals[(size_t)2 * i] = rr[0];
      residuals[(size_t)2 * i + 1] = rr[1];
    }
    const double s = rr[0] * rr[0] + rr[1] * rr[1];
    if (g.cfg.robust) {
      double rho[3];
      cauchy(0.5, s, rho);
      c += 0.5 * rho[0];
    } else {
      c += 0.5 * s;
    }
    // calcReprojectionError, src/CameraCalibration.cpp:1080-1093
    sx2 += rr[0] * rr[0];
    sy2 += rr[1] * rr[1];
    mx = std::max(mx, std::fabs(rr[0]));
    my = std::max(my, std::fabs(rr[1]));
    if (s <= inlier_threshold * inlier_threshold) ++inl;
  }
  for (int k = 0; k < g.K; ++k) {
    double rr;
    distance_residual<double>(g.cdist[k], g.csigma[k], points + 3 * g.c1[k], points + 3 * g.c2[k], &rr);
    c += 0.5 * rr * rr;
  }
  if (cost) *cost = c;
  if (stats) {
    stats->std_x = std::sqrt(sx2 / (double)g.N);
    stats->std_y = std::sqrt(sy2 / (double)g.N);
    stats->mae_x = mx;
    stats->mae_y = my;
    stats->num_points = g.N;
    stats->num_inliers = inl;
  }
  return LFBA_OK;
}

extern "C" int oracle_time_eval(const lfba_problem* pb, const double* camera, const double* views,
                                const double* points, int reps, int num_threads, double* seconds) {
  Program g;
  if (!build_program(pb, camera, g)) return LFBA_INVALID_ARGUMENT;
  const int nt = num_threads > 0 ? num_threads : omp_get_max_threads();
  Evaluator ev(g, 0.5, nt, nullptr);
  ev.alloc_jacobian();
  std::vector<double> grad;
  ev.evaluate(camera, views, points, grad);  // warm-up (page faults of the Jacobian)
  const double t0 = now_s();
  for (int k = 0; k < reps; ++k) ev.evaluate(camera, views, points, grad);
  *seconds = now_s() - t0;
  return LFBA_OK;
}

static int oracle_solve_impl(const lfba_problem* pb, const lfba_options* opt, double* camera17, double* views6F,
                             double* points3P, lfba_summary* sum, int num_threads, double max_seconds,
                             oracle_block_fn fn, bool streaming) {
  const double t_start = now_s();
  Program g;
  if (!build_program(pb, camera17, g)) return LFBA_INVALID_ARGUMENT;
  const int nt = num_threads > 0 ? num_threads : omp_get_max_threads();
  Evaluator ev(g, opt->loss_scale, nt, fn);
  ev.streaming = streaming;
  ev.alloc_jacobian();
  SchurSolver schur(g, ev, nt);

  Params x, cand;
  x.camera.assign(camera17, camera17 + 17);
  x.views.assign(views6F, views6F + (size_t)6 * g.F);
  x.points.assign(points3P, points3P + (size_t)3 * g.P);
  const int n_t = g.n_t;

  std::vector<lfba_iteration> log;
  std::vector<double> gradient, scale((size_t)n_t, 1.0), diagonal, lm_diag((size_t)n_t), step, delta((size_t)n_t);
  double x_cost = 0, x_norm = 0, candidate_cost = 0, model_cost_change = 0;
  double radius = opt->initial_trust_region_radius, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  int num_consecutive_invalid = 0;
  int n_success = 0, n_fail = 0;
  lfba_iteration it;
  std::memset(&it, 0, sizeof(it));
  int termination = LFBA_NO_CONVERGENCE, stop = LFBA_STOP_NONE;
  double t_iter = now_s();

  if (opt->minimizer_progress_to_stdout)
    std::printf("iter      cost      cost_change  |gradient|   |step|    tr_ratio  tr_radius  ls_iter  iter_time  total_time\n");

  // EvaluateGradientAndJacobian: cost, residuals, gradient, Jacobian; Jacobi scaling; projected gradient norms
  auto evaluate_gradient_and_jacobian = [&](int iteration) {
    x_cost = ev.evaluate(x.camera.data(), x.views.data(), x.points.data(), gradient);
    it.cost = x_cost;
    if (iteration == 0) {
      std::vector<double> cn;
      ev.squared_column_norms(cn);
      for (int j = 0; j < n_t; ++j) scale[j] = 1.0 / (1.0 + std::sqrt(cn[j]));
    }
    ev.scale_columns(scale);
    std::vector<double> neg((size_t)n_t);
    for (int j = 0; j < n_t; ++j) neg[j] = -gradient[j];
    Params proj;
    plus(g, x, neg, proj);
    diff_norms(g, x, proj, it.gradient_norm, it.gradient_max_norm);
  };

  // ---- IterationZero ----
  if (g.constrained) {
    std::vector<double> zero((size_t)n_t, 0.0);
    plus(g, x, zero, cand);  // projects the start point into the box
    x = cand;
  }
  x_norm = std::sqrt(sqnorm_x(g, x));
  it.iteration = 0;
  evaluate_gradient_and_jacobian(0);
  it.step_is_valid = 1;
  it.step_is_successful = 1;
  const double initial_cost = x_cost;
  double minimum_cost = std::numeric_limits<double>::max();
  Params best = x;

  // FinalizeIterationAndCheckIfMinimizerCanContinue
  auto finalize_and_continue = [&]() -> bool {
    if (it.step_is_successful) {
      ++n_success;
      if (x_cost < minimum_cost) {
        minimum_cost = x_cost;
        best = x;
      }
    } else {
      ++n_fail;
    }
    it.trust_region_radius = radius;
    const double t = now_s();
    it.iteration_time_s = t - t_iter;
    it.cumulative_time_s = t - t_start;
    log.push_back(it);
    if (opt->minimizer_progress_to_stdout)
      std::printf("% 4d % 8e   % 3.2e   % 3.2e  % 3.2e  % 3.2e % 3.2e     % 4d   % 3.2e   % 3.2e\n", it.iteration,
                  it.cost, it.cost_change, it.gradient_max_norm, it.step_norm, it.relative_decrease,
                  it.trust_region_radius, 1, it.iteration_time_s, it.cumulative_time_s);
    if (max_seconds > 0 && t - t_start > max_seconds) {
      termination = LFBA_NO_CONVERGENCE;
      stop = LFBA_STOP_NONE;
      return false;
    }
    if (it.iteration >= opt->max_num_iterations) {
      termination = LFBA_NO_CONVERGENCE;
      stop = LFBA_STOP_MAX_ITERATIONS;
      return false;
    }
    if (it.step_is_successful && it.gradient_max_norm <= opt->gradient_tolerance) {
      termination = LFBA_CONVERGENCE;
      stop = LFBA_STOP_GRADIENT_TOLERANCE;
      return false;
    }
    if (it.trust_region_radius <= opt->min_trust_region_radius) {
      termination = LFBA_CONVERGENCE;
      stop = LFBA_STOP_MIN_RADIUS;
      return false;
    }
    return true;
  };

  int status = LFBA_OK;
  while (finalize_and_continue()) {
    t_iter = now_s();
    const double prev_gnorm = it.gradient_norm, prev_gmax = it.gradient_max_norm;
    const int iteration = log.back().iteration + 1;
    std::memset(&it, 0, sizeof(it));
    it.iteration = iteration;

    // ---- ComputeTrustRegionStep: LevenbergMarquardtStrategy::ComputeStep + model cost change ----
    if (!reuse_diagonal) {
      ev.squared_column_norms(diagonal);
      for (int j = 0; j < n_t; ++j)
        diagonal[j] = std::min(std::max(diagonal[j], opt->min_lm_diagonal), opt->max_lm_diagonal);
    }
    for (int j = 0; j < n_t; ++j) lm_diag[j] = std::sqrt(diagonal[j] / radius);
    const bool solved = schur.solve(lm_diag, step);
    reuse_diagonal = true;
    it.step_is_valid = 0;
    if (solved) {
      for (double& v : step) v = -v;
      model_cost_change = ev.streaming ? schur.model_cost_change : ev.model_cost_change(step);
      it.step_is_valid = model_cost_change > 0.0;
    }
    if (!it.step_is_valid) {
      // HandleInvalidStep
      if (++num_consecutive_invalid >= opt->max_num_consecutive_invalid_steps) {
        termination = LFBA_TERM_FAILURE;
        stop = LFBA_STOP_INVALID_STEPS;
        status = LFBA_FAILURE;
        break;
      }
      radius = radius / decrease_factor;  // StepIsInvalid == StepRejected(0)
      decrease_factor *= 2.0;
      reuse_diagonal = true;
      it.cost = x_cost;
      it.cost_change = 0.0;
      it.gradient_max_norm = log.back().gradient_max_norm;
      it.gradient_norm = log.back().gradient_norm;
      it.step_norm = 0.0;
      it.relative_decrease = 0.0;
      it.step_is_successful = 0;
      continue;
    }
    num_consecutive_invalid = 0;
    for (int j = 0; j < n_t; ++j) delta[j] = step[j] * scale[j];

    // ---- DoLineSearch (bounds only) ----
    if (g.constrained) {
      double gd = 0;
      for (int j = 0; j < n_t; ++j) gd += gradient[j] * delta[j];
      double dmax = 0;
      for (int j = 0; j < n_t; ++j) dmax = std::max(dmax, std::fabs(delta[j]));
      Sample initial, previous, current;
      initial.x = 0;
      initial.value = x_cost;
      initial.gradient = gd;
      initial.value_valid = initial.gradient_valid = true;
      auto eval_at = [&](double a, Sample& s) {
        std::vector<double> sd((size_t)n_t);
        for (int j = 0; j < n_t; ++j) sd[j] = a * delta[j];
        Params xa;
        plus(g, x, sd, xa);
        // CUBIC interpolation => Ceres evaluates value AND gradient at every trial point.  The Jacobian
        // scratch is shared with the LM state in this restatement, so evaluate into a private evaluator.
        Evaluator le(g, opt->loss_scale, nt, fn);
        le.streaming = ev.streaming;
        le.alloc_jacobian();
        std::vector<double> gr;
        s.x = a;
        s.value = le.evaluate(xa.camera.data(), xa.views.data(), xa.points.data(), gr);
        ev.num_jac_evals += 1;
        s.value_valid = std::isfinite(s.value);
        double d = 0;
        for (int j = 0; j < n_t; ++j) d += gr[j] * delta[j];
        s.gradient = d;
        s.gradient_valid = s.value_valid && std::isfinite(d);
      };
      int ls_iters = 0;
      bool success = true;
      eval_at(1.0, current);
      const bool ls_debug = std::getenv("LFBA_DEBUG") != nullptr;
      if (ls_debug)
        std::printf("[oracle dbg] iter %d line search: a=%.17g phi=%.17g dphi=%.17g | phi0=%.17g dphi0=%.17g\n", iteration,
                    current.x, current.value, current.gradient, initial.value, initial.gradient);
      while (!current.value_valid || current.value > initial.value + 1e-4 * initial.gradient * current.x) {
        if (++ls_iters >= 20) {
          success = false;
          break;
        }
        double a_new;
        if (!current.value_valid) {
          a_new = std::min(std::max(current.x * 0.5, 1e-3 * current.x), 0.6 * current.x);
        } else {
          std::vector<Sample> smp{initial, current};
          if (previous.value_valid) smp.push_back(previous);
          a_new = minimize_interpolant(smp, 1e-3 * current.x, 0.6 * current.x);
        }
        if (a_new * dmax < 1e-9) {
          success = false;
          break;
        }
        previous = current;
        eval_at(a_new, current);
        if (ls_debug)
          std::printf("[oracle dbg] iter %d line search: a=%.17g phi=%.17g dphi=%.17g\n", iteration, current.x, current.value,
                      current.gradient);
      }
      it.line_search_iterations = ls_iters;
      if (success)
        for (int j = 0; j < n_t; ++j) delta[j] *= current.x;
    }

    // ---- ComputeCandidatePointAndEvaluateCost ----
    plus(g, x, delta, cand);
    candidate_cost = ev.cost_only(cand.camera.data(), cand.views.data(), cand.points.data());
    if (!std::isfinite(candidate_cost)) candidate_cost = std::numeric_limits<double>::max();

    // ---- ParameterToleranceReached ----
    double dn2, dninf;
    diff_norms(g, x, cand, dn2, dninf);
    it.step_norm = dn2;
    if (it.step_norm <= opt->parameter_tolerance * (x_norm + opt->parameter_tolerance)) {
      termination = LFBA_CONVERGENCE;
      stop = LFBA_STOP_PARAMETER_TOLERANCE;
      break;
    }
    // ---- FunctionToleranceReached ----
    it.cost_change = x_cost - candidate_cost;
    if (std::fabs(it.cost_change) <= opt->function_tolerance * x_cost) {
      termination = LFBA_CONVERGENCE;
      stop = LFBA_STOP_FUNCTION_TOLERANCE;
      break;
    }
    // ---- IsStepSuccessful (monotonic steps: StepQuality = (f - f+)/model_cost_change) ----
    it.relative_decrease = candidate_cost >= std::numeric_limits<double>::max()
                               ? std::numeric_limits<double>::lowest()
                               : (x_cost - candidate_cost) / model_cost_change;
    if (it.relative_decrease > opt->min_relative_decrease) {
      // HandleSuccessfulStep
      x = cand;
      x_norm = std::sqrt(sqnorm_x(g, x));
      evaluate_gradient_and_jacobian(iteration);
      it.step_is_successful = 1;
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * it.relative_decrease - 1.0, 3));
      radius = std::min(opt->max_trust_region_radius, radius);
      decrease_factor = 2.0;
      reuse_diagonal = false;
    } else {
      it.step_is_successful = 0;
      it.cost = candidate_cost;
      it.gradient_norm = prev_gnorm;
      it.gradient_max_norm = prev_gmax;
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
      reuse_diagonal = true;
    }
  }

  // user parameters = state at the lowest cost seen among successful steps
  std::memcpy(camera17, best.camera.data(), 17 * sizeof(double));
  std::memcpy(views6F, best.views.data(), (size_t)6 * g.F * sizeof(double));
  std::memcpy(points3P, best.points.data(), (size_t)3 * g.P * sizeof(double));
  if (sum) {
    sum->termination_type = termination;
    sum->stop_reason = stop;
    sum->num_iterations = (int)log.size();
    sum->num_successful_steps = n_success;
    sum->num_unsuccessful_steps = n_fail;
    sum->reduced_system_size = g.n_red;
    sum->initial_cost = initial_cost;
    sum->final_cost = minimum_cost;
    sum->num_jacobian_evals = ev.num_jac_evals;
    sum->num_observations = g.N;
    sum->num_tracks = 0;
    sum->num_lenses = ev.num_block_recomputes;  // streaming mode: passes over all blocks (0 in stored mode); oracle-only use of this slot
    sum->gpu_launches = 0;
    sum->setup_time_s = 0;
    sum->solve_time_s = now_s() - t_start;
    sum->solve_gpu_ms = 0;
    if (sum->iterations)
      for (int i = 0; i < std::min((int)log.size(), sum->iterations_capacity); ++i) sum->iterations[i] = log[i];
  }
  return status;
}

extern "C" int oracle_solve(const lfba_problem* pb, const lfba_options* opt, double* camera17, double* views6F,
                            double* points3P, lfba_summary* sum, int num_threads, double max_seconds,
                            oracle_block_fn fn) {
  return oracle_solve_impl(pb, opt, camera17, views6F, points3P, sum, num_threads, max_seconds, fn, false);
}

// Same solve without the O(N) Jacobian storage (block-recompute): for scenes whose Jacobian in Ceres' layout would not
// fit the host (416 B per observation). Same algorithm and per-block arithmetic; sums over blocks are taken in point
// order, so results agree with oracle_solve to rounding (tests/test_oracle_solver.py).
extern "C" int oracle_solve_streaming(const lfba_problem* pb, const lfba_options* opt, double* camera17,
                                      double* views6F, double* points3P, lfba_summary* sum, int num_threads,
                                      double max_seconds, oracle_block_fn fn) {
  return oracle_solve_impl(pb, opt, camera17, views6F, points3P, sum, num_threads, max_seconds, fn, true);
}
