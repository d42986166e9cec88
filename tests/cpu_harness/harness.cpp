// tests/cpu_harness/harness.cpp — TEST INFRASTRUCTURE.
// Runs the product's __host__ __device__ per-item math (lifcal_b200/csrc/lfba_math.cuh — the very functions the
// CUDA kernels call) on the CPU, so the analytic Jacobian and the per-track chain rule can be checked against
// the oracle in the no-GPU test tier. It is NOT a fallback: it is built only by the tests, lives outside the
// package and is never loaded by lifcal_b200/.
#include <cstdint>
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/lfba.h"
#include "../../lifcal_b200/csrc/lfba_math.cuh"

using namespace lfba;

static double g_feature_worst = 0.0;
static double g_feature9_worst = 0.0;  // the form k_eval_rows uses: NC features, weighted Gram, gram9_expand
static double g_cancel_sum[3] = {0,0,0}, g_cancel_n[3] = {0,0,0}, g_cancel_max[3] = {0,0,0};
extern "C" void harness_cancel(double* out) { for (int c = 0; c < 3; ++c) { out[2*c] = g_cancel_sum[c] / (g_cancel_n[c] + 1e-300); out[2*c+1] = g_cancel_max[c]; } }
extern "C" double harness_feature_worst(void) { return g_feature_worst; }
extern "C" double harness_feature9_worst(void) { return g_feature9_worst; }
extern "C" void harness_feature_reset(void) { g_feature_worst = 0.0; g_feature9_worst = 0.0; }

template <int NC, int NRAD>
static void eval_all(const lfba_problem* pb, const CamModel& m, const double* views, const double* points,
                     double* res, double* jc, double* jv, double* jp) {
  const bool rp = pb->config & LFBA_CFG_REFINE_POSES, r3 = pb->config & LFBA_CFG_REFINE_POINTS;
  for (int64_t i = 0; i < pb->n_obs; ++i) {
    double le[kLensStride], fe[kFrameStride], Pc[3];
    lens_entry(m, pb->ml_x[i], pb->ml_y[i], le);
    frame_entry(views + 6 * pb->frame_idx[i], fe);
    const double* X = points + 3 * pb->point_idx[i];
    track_point(fe, X, Pc);
    TrackCtx t;
    track_setup(m, Pc, t);
    double r[2], G[6], J[2 * NC];
    obs_eval<NC, NRAD>(m, t, le, pb->obs_x[i], pb->obs_y[i], r, G, J);
    double r2[2];
    obs_residual(m, t, le, pb->obs_x[i], pb->obs_y[i], r2);
    {  // feature form must reproduce the same Jacobian blocks through the Gram expansion (single observation)
      constexpr int NF = NC + 1, NQ = NF * (NF + 1) / 2;
      double rf[2], F[2 * NF], g[NQ + NF];
      obs_features<NC, NRAD>(m, t, le, pb->obs_x[i], pb->obs_y[i], rf, F);
      int q = 0;
      for (int a = 0; a < NF; ++a)
        for (int b = 0; b <= a; ++b) g[q++] = F[a] * F[b] + F[NF + a] * F[NF + b];
      for (int a = 0; a < NF; ++a) g[NQ + a] = F[a] * rf[0] + F[NF + a] * rf[1];
      double worst = 0.0;
      auto chk = [&](double a, double b) {
        const double den = fabs(a) > fabs(b) ? fabs(a) : fabs(b);
        if (den > 0 && fabs(a - b) / den > worst) worst = fabs(a - b) / den;
      };
      if (rf[0] != r[0] || rf[1] != r[1]) worst = 1.0;
      for (int c1 = 0; c1 < NC; ++c1)
        for (int c2 = 0; c2 <= c1; ++c2)
          chk(GramMap<NC>::cc(t, g, c1, c2), J[c1] * J[c2] + J[NC + c1] * J[NC + c2]);
      for (int a = 0; a < 3; ++a) {
        for (int c = 0; c < NC; ++c) chk(GramMap<NC>::gcam(t, g, a, c), G[a] * J[c] + G[3 + a] * J[NC + c]);
        for (int b = 0; b < 3; ++b) chk(GramMap<NC>::gg(t, g, a, b), G[a] * G[b] + G[3 + a] * G[3 + b]);
      }
      if (worst > g_feature_worst) g_feature_worst = worst;
      {  // what the fused kernel does: NC features (f2 dropped), Gram weighted as (w f_a).f_b, rebuilt by gram9_expand
        constexpr int NF9 = Feat9Dims<NC>::NF, NQ9 = Feat9Dims<NC>::NQ;
        double r9[2], F9[2 * NF9], g9[NQ9 + NF9], go[NQ + NF];
        obs_features9<NC, NRAD>(m, t, le, pb->obs_x[i], pb->obs_y[i], r9, F9, m.ml_adjust != 0, m.any_dist != 0);
        const double s9 = r9[0] * r9[0] + r9[1] * r9[1];
        const double w = 1.0 / (1.0 + s9 * m.loss_c);  // rho' of the Cauchy loss
        int q9 = 0;
        for (int a = 0; a < NF9; ++a) {
          const double wx = w * F9[a], wy = w * F9[NF9 + a];
          for (int b = 0; b <= a; ++b) g9[q9++] = wx * F9[b] + wy * F9[NF9 + b];
          g9[NQ9 + a] = wx * r9[0] + wy * r9[1];
        }
        gram9_expand<NC>(t, t.a1 * m.gamma, g9, go);
        double worst9 = 0.0;
        auto chk9 = [&](double a, double b, double scale) {
          if (scale > 0 && fabs(a - b) / scale > worst9) worst9 = fabs(a - b) / scale;
        };
        if (r9[0] != r[0] || r9[1] != r[1]) worst9 = 1.0;
        // compare against w * (Jacobian block products); scale = product of the column norms (Cauchy-Schwarz bound)
        auto nrm = [&](const double* v, int i0, int i1) { return sqrt(v[i0] * v[i0] + v[i1] * v[i1]); };
        for (int c1 = 0; c1 < NC; ++c1)
          for (int c2 = 0; c2 <= c1; ++c2)
            chk9(GramMap<NC>::cc(t, go, c1, c2), w * (J[c1] * J[c2] + J[NC + c1] * J[NC + c2]),
                 w * nrm(J, c1, NC + c1) * nrm(J, c2, NC + c2));
        for (int a = 0; a < 3; ++a) {
          for (int c = 0; c < NC; ++c)
            chk9(GramMap<NC>::gcam(t, go, a, c), w * (G[a] * J[c] + G[3 + a] * J[NC + c]), w * nrm(G, a, 3 + a) * nrm(J, c, NC + c));
          for (int b = 0; b < 3; ++b)
            chk9(GramMap<NC>::gg(t, go, a, b), w * (G[a] * G[b] + G[3 + a] * G[3 + b]), w * nrm(G, a, 3 + a) * nrm(G, b, 3 + b));
          chk9((a == 2 ? -t.g1 : t.g1) * go[NQ + a], w * (G[a] * r[0] + G[3 + a] * r[1]), w * nrm(G, a, 3 + a) * sqrt(s9));
        }
        if (worst9 > g_feature9_worst) g_feature9_worst = worst9;
      }
      for (int c = 0; c < 3; ++c) {
        double a, b;
        GramMap<NC>::geo(t, c, a, b);
        for (int row = 0; row < 2; ++row) {
          const double t1 = (a + b * t.a1 * m.gamma) * F[row * NF + 3], t2 = b * (t.Px * F[row * NF + 0] + t.Py * F[row * NF + 1]);
          const double ratio = (fabs(t1) + fabs(t2)) / (fabs(t1 + t2) + 1e-300);
          g_cancel_sum[c] += (fabs(t1) + fabs(t2)) * (fabs(t1) + fabs(t2)); g_cancel_n[c] += (t1 + t2) * (t1 + t2); if (ratio > g_cancel_max[c]) g_cancel_max[c] = ratio;
        }
      }
    }
    res[2 * i] = r[0];
    res[2 * i + 1] = r[1];
    if (r2[0] != r[0] || r2[1] != r[1]) res[2 * i] = 1e300;  // the two code paths must agree exactly
    for (int row = 0; row < 2; ++row) {
      for (int c = 0; c < 17; ++c) jc[34 * i + 17 * row + c] = c < NC ? J[NC * row + c] : 0.0;
      // pose block: [G dR0 X, G dR1 X, G dR2 X, G]
      for (int k = 0; k < 3; ++k) {
        double mk[3];
        mat3_vec(fe + 9 + 9 * k, X, mk);
        jv[12 * i + 6 * row + k] = rp ? G[3 * row] * mk[0] + G[3 * row + 1] * mk[1] + G[3 * row + 2] * mk[2] : 0.0;
        jv[12 * i + 6 * row + 3 + k] = rp ? G[3 * row + k] : 0.0;
      }
      // point block: G R
      for (int k = 0; k < 3; ++k)
        jp[6 * i + 3 * row + k] =
            r3 ? G[3 * row] * fe[k] + G[3 * row + 1] * fe[3 + k] + G[3 * row + 2] * fe[6 + k] : 0.0;
    }
  }
}

extern "C" int harness_eval(const lfba_problem* pb, const double* camera, const double* views, const double* points,
                            double* res, double* jc, double* jv, double* jp) {
  CamModel m;
  cam_model_init(m, camera, pb->config, pb->spx, pb->spy, pb->scale, 0.5);
  switch (m.n_radial * 2 + m.tangential) {
    case 0: eval_all<5, 0>(pb, m, views, points, res, jc, jv, jp); break;
    case 1: eval_all<7, 0>(pb, m, views, points, res, jc, jv, jp); break;
    case 2: eval_all<6, 1>(pb, m, views, points, res, jc, jv, jp); break;
    case 3: eval_all<8, 1>(pb, m, views, points, res, jc, jv, jp); break;
    case 4: eval_all<7, 2>(pb, m, views, points, res, jc, jv, jp); break;
    case 5: eval_all<9, 2>(pb, m, views, points, res, jc, jv, jp); break;
    default: return 1;
  }
  return 0;
}

extern "C" int harness_spd3_inverse(const double* a6, double* inv6) { return spd3_inverse(a6, inv6) ? 1 : 0; }
extern "C" void harness_distance(const double* p1, const double* p2, double d, double s, double* r, double* j) {
  distance_eval(p1, p2, d, s, *r, j);
}
extern "C" void harness_robust(const double* camera, uint32_t config, double s, double* scale, double* rho) {
  CamModel m;
  cam_model_init(m, camera, config, 0.011, 0.011, 2.0, 0.5);
  *scale = robust_scale(m, s, *rho);
}
