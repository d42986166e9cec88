// lfba_project.cu — the step BEFORE the LF-BA solve on the device: every total-focus feature with a
// valid virtual depth is projected into all micro images that see it.
//
// Reference being replaced: CameraCalibration::projectPointsToRawImage (src/CameraCalibration.cpp:640-769) with the
// web of epipolar lines of CameraCalibration::defineEpiPolarLines (:521-634, EpiPolarLine.cpp:16-46). The reference runs
// three nested host loops (frames x features x web) and pushes into std::vectors; here one thread handles one feature,
// a counting pass + exclusive scan fixes every feature's output range, and a second pass writes the observations in
// EXACTLY the reference's order (feature order, nearest lens first, then the web in ascending base-line order, +/-
// direction) straight into the SoA arrays lfba_solve() consumes.
// Arithmetic: the reference computes in float32 with a few double sub-expressions; every operation below is the same
// IEEE operation in the same order (__fmul_rn / __fadd_rn / ... keep nvcc from contracting a*b+c into an FMA the
// reference's x86-64 build does not have), so lens selection is bit-exact and coordinates are float32-exact.
#include <cmath>
#include <cstring>
#include <vector>

#include <algorithm>

#include "lfba_setup.cuh"

namespace lfba {

namespace {

// ---- host: the web (a few hundred lines, built once per call; double arithmetic as in the reference) ----
struct WebLine {
  double ex, ey, dist;
};
WebLine web_line(double x, double y, double dist) {
  WebLine e{x, y, dist};
  const double l2 = x * x + y * y;
  if (l2 != 1.0f) {  // the reference compares the double squared length with the float literal
    const double l = std::sqrt(l2);
    e.ex = x / l;
    e.ey = y / l;
  }
  return e;
}
WebLine web_sum(const WebLine& a, const WebLine& b) {
  const double x = a.ex * a.dist + b.ex * b.dist, y = a.ey * a.dist + b.ey * b.dist;
  return web_line(x, y, std::sqrt(x * x + y * y));
}
void build_web(float D, float rotation, bool rotate, std::vector<WebLine>& lines, std::vector<int32_t>& group_begin) {
  const float max_dist = D * 10;
  const double h = std::sqrt(0.75);
  WebLine base[5] = {web_line(1, 0, D), web_line(0.5, h, D), web_line(-0.5, -h, D), web_line(0.5, -h, D), web_line(-0.5, h, D)};
  if (rotate) {
    const double ca = std::cos(rotation), sa = std::sin(rotation);
    for (WebLine& e : base) {
      const double ex = e.ex, ey = e.ey;
      e.ex = ex * ca + ey * sa;
      e.ey = -ex * sa + ey * ca;
    }
  }
  const WebLine &e0 = base[0], &e1 = base[1], &m1 = base[2], &e2 = base[3], &m2 = base[4];
  std::vector<WebLine> all{e1, e2};
  for (int i = 0; all.back().dist < max_dist; ++i) {  // two zig-zag chains away from the centre
    const WebLine a = web_sum(all[2 * i], i % 2 == 0 ? m2 : e1), b = web_sum(all[2 * i + 1], i % 2 == 0 ? m1 : e2);
    all.push_back(a);
    all.push_back(b);
  }
  all.push_back(e0);
  const size_t n0 = all.size();
  for (size_t k = 0; k < n0; ++k)  // every chain element extended along the x base line
    for (WebLine last = all[k]; last.dist < max_dist;) {
      last = web_sum(last, e0);
      all.push_back(last);
    }
  // groups of float-equal length in ascending order; a new length is inserted before the first longer group
  std::vector<std::vector<WebLine>> groups{{all[0]}};
  for (size_t k = 1; k < all.size(); ++k) {
    const WebLine& e = all[k];
    if (e.ey == -1.0f || e.dist > max_dist) continue;
    size_t g = 0;
    bool eq = false, lt = false;
    for (; g < groups.size(); ++g) {
      if ((float)groups[g][0].dist == (float)e.dist) { eq = true; break; }
      if (groups[g][0].dist > e.dist) { lt = true; break; }
    }
    if (eq) groups[g].push_back(e);
    else if (lt) groups.insert(groups.begin() + g, std::vector<WebLine>{e});
    else groups.push_back({e});
  }
  lines.clear();
  group_begin.assign(1, 0);
  for (auto& g : groups) {
    lines.insert(lines.end(), g.begin(), g.end());
    group_begin.push_back((int32_t)lines.size());
  }
}

struct GridDev {
  int W, H, scale;
  float D, valid_r2;
  const float *cx, *cy;
  const int32_t *map_next, *map_ml;
  const double* web;          // [n_lines][3]
  const int32_t* group_begin; // [n_groups + 1]
  int n_groups;
};

// One feature: visits its micro lenses in the reference's order and calls emit(xR, yR, cx, cy) for every observation.
template <class Emit>
__device__ __forceinline__ int project_one(const GridDev& g, double img_x, double img_y, double vd, Emit emit) {
  const float v = (float)vd;
  if (!((double)v > 2.0 && (double)v < 20.0)) return 0;
  const float x = (float)img_x, y = (float)img_y;
  const float radius = __fadd_rn(__fmul_rn(__fmul_rn(g.D, 0.5f), v), 2.0f);
  const float radius2 = __fmul_rn(radius, radius);
  const float sc = (float)g.scale;
  const float xu = __fsub_rn(__fmul_rn(sc, __fadd_rn(x, 0.5f)), 0.5f);
  const float yu = __fsub_rn(__fmul_rn(sc, __fadd_rn(y, 0.5f)), 0.5f);
  int xi = (int)__fadd_rn(xu, 0.5f), yi = (int)__fadd_rn(yu, 0.5f);
  if (xi >= g.W) xi = g.W - 1;
  if (yi >= g.H) yi = g.H - 1;
  if (xi < 0 || yi < 0) return 0;  // outside the image: the reference would index out of bounds
  const int ml0 = g.map_next[xi + g.W * yi];
  if (ml0 < 0) return 0;
  const float cxn = g.cx[ml0], cyn = g.cy[ml0];
  {
    const float dx = __fsub_rn(cxn, xu), dy = __fsub_rn(cyn, yu);
    if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) > radius2) return 0;
  }
  int n = 0;
  auto visit = [&](int ml) {
    const float cx = g.cx[ml], cy = g.cy[ml];
    const float xr = __fadd_rn(__fdiv_rn(__fsub_rn(xu, cx), v), cx);
    const float yr = __fadd_rn(__fdiv_rn(__fsub_rn(yu, cy), v), cy);
    if (!(xr >= 0 && xr <= (float)(g.W - 1) && yr >= 0 && yr <= (float)(g.H - 1))) return;
    const float tx = __fsub_rn(xr, cx), ty = __fsub_rn(yr, cy);
    if (__fadd_rn(__fmul_rn(tx, tx), __fmul_rn(ty, ty)) >= g.valid_r2) return;
    emit(n, xr, yr, cx, cy);
    ++n;
  };
  visit(ml0);
  for (int gi = 0; gi < g.n_groups; ++gi) {
    const int lb = g.group_begin[gi], le = g.group_begin[gi + 1];
    if (g.web[3 * lb + 2] > (double)radius) break;
    for (int li = lb; li < le; ++li) {
      const float bl = (float)g.web[3 * li + 2];
      const double ex = g.web[3 * li], ey = g.web[3 * li + 1];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const float epx = s == 0 ? (float)ex : (float)(-ex), epy = s == 0 ? (float)ey : (float)(-ey);
        const float cx = __fadd_rn(cxn, __fmul_rn(bl, epx)), cy = __fadd_rn(cyn, __fmul_rn(bl, epy));
        const float dx = __fsub_rn(cx, xu), dy = __fsub_rn(cy, yu);
        if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) > radius2) continue;
        int cxi = (int)__dadd_rn((double)cx, 0.5), cyi = (int)__dadd_rn((double)cy, 0.5);
        cxi = cxi < 0 ? 0 : (cxi >= g.W ? g.W - 1 : cxi);
        cyi = cyi < 0 ? 0 : (cyi >= g.H ? g.H - 1 : cyi);
        const int ml = g.map_ml[cxi + cyi * g.W];
        if (ml >= 0) visit(ml);
      }
    }
  }
  return n;
}

__global__ void k_project_count(GridDev g, const double* fx, const double* fy, const double* vd, int32_t* counts, int64_t M) {
  const int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (f >= M) return;
  counts[f] = project_one(g, fx[f], fy[f], vd[f], [](int, float, float, float, float) {});
}
__global__ void k_project_write(GridDev g, const double* fx, const double* fy, const double* vd, const int32_t* fframe,
                                const int32_t* fpoint, const int64_t* offs, double* ox, double* oy, double* mx, double* my,
                                int32_t* opoint, int32_t* oframe, int64_t M) {
  const int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (f >= M) return;
  const int64_t o = offs[f];
  const int32_t fr = fframe ? fframe[f] : 0, pt = fpoint ? fpoint[f] : (int32_t)f;
  project_one(g, fx[f], fy[f], vd[f], [&](int k, float xr, float yr, float cx, float cy) {
    ox[o + k] = (double)xr;
    oy[o + k] = (double)yr;
    mx[o + k] = (double)cx;
    my[o + k] = (double)cy;
    if (opoint) opoint[o + k] = pt;
    if (oframe) oframe[o + k] = fr;
  });
}
__global__ void k_counts_to_i64(const int32_t* c, int64_t* out, int64_t M) {
  const int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (f < M) out[f] = c[f];
}

}  // namespace
}  // namespace lfba

using namespace lfba;

extern "C" int lfba_project_to_raw(const lfba_lens_grid* grid, int64_t n_features, const double* feat_x, const double* feat_y,
                                   const double* vdepth, const int32_t* frame_idx, const int32_t* point_idx, int64_t capacity,
                                   double* obs_x, double* obs_y, double* ml_x, double* ml_y, int32_t* out_point_idx,
                                   int32_t* out_frame_idx, int64_t* n_obs, int32_t device) {
  if (!grid || !n_obs || n_features < 0 || (n_features > 0 && (!feat_x || !feat_y || !vdepth))) return LFBA_INVALID_ARGUMENT;
  if (grid->raw_width <= 0 || grid->raw_height <= 0 || grid->n_lenses <= 0 || !grid->lens_cx || !grid->lens_cy ||
      !grid->map_next || !grid->map_ml)
    return LFBA_INVALID_ARGUMENT;
  try {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return LFBA_NO_DEVICE;
    if (device >= 0) LFBA_CUDA(cudaSetDevice(device));
    cudaStream_t s;
    LFBA_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    struct SG { cudaStream_t s; ~SG() { cudaStreamSynchronize(s); cudaStreamDestroy(s); } } sg{s};
    alloc_stream() = s;
    int rc = LFBA_OK;
    {
      std::vector<WebLine> lines;
      std::vector<int32_t> gb;
      build_web(grid->lens_diameter, grid->rotation, grid->rotation_on_grid != 0, lines, gb);
      const size_t WH = (size_t)grid->raw_width * grid->raw_height, M = (size_t)n_features;
      DevBuf<float> cx(grid->n_lenses), cy(grid->n_lenses);
      DevBuf<int32_t> mnext(WH), mml(WH), gbd(gb.size()), counts(M + 1), ffr(frame_idx ? M : 0), fpt(point_idx ? M : 0);
      DevBuf<double> web(lines.size() * 3), fx(M), fy(M), vd(M);
      DevBuf<int64_t> c64(M + 1), offs(M + 1);
      cx.upload(grid->lens_cx, grid->n_lenses, s);
      cy.upload(grid->lens_cy, grid->n_lenses, s);
      mnext.upload(grid->map_next, WH, s);
      mml.upload(grid->map_ml, WH, s);
      gbd.upload(gb.data(), gb.size(), s);
      web.upload(&lines[0].ex, lines.size() * 3, s);
      fx.upload(feat_x, M, s);
      fy.upload(feat_y, M, s);
      vd.upload(vdepth, M, s);
      if (frame_idx) ffr.upload(frame_idx, M, s);
      if (point_idx) fpt.upload(point_idx, M, s);
      GridDev g{grid->raw_width, grid->raw_height, grid->scale, grid->lens_diameter, grid->lens_validity_radius_2,
                cx.p, cy.p, mnext.p, mml.p, web.p, gbd.p, (int)gb.size() - 1};
      counts.zero(s);
      const unsigned nb = (unsigned)((M + 127) / 128);
      int64_t total = 0;
      if (M > 0) {
        k_project_count<<<nb, 128, 0, s>>>(g, fx.p, fy.p, vd.p, counts.p, (int64_t)M);
        k_counts_to_i64<<<(unsigned)((M + 1 + 255) / 256), 256, 0, s>>>(counts.p, c64.p, (int64_t)M + 1);
        size_t bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, bytes, c64.p, offs.p, (int)(M + 1), s);
        DevBuf<unsigned char> tmp(bytes);
        LFBA_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, c64.p, offs.p, (int)(M + 1), s));
        LFBA_CUDA(cudaMemcpyAsync(&total, offs.p + M, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
        LFBA_CUDA(cudaStreamSynchronize(s));
      }
      *n_obs = total;
      if (capacity > 0 && total > 0) {
        if (capacity < total || !obs_x || !obs_y || !ml_x || !ml_y) {
          rc = LFBA_INVALID_ARGUMENT;
        } else {
          DevBuf<double> ox((size_t)total), oy((size_t)total), mx((size_t)total), my((size_t)total);
          DevBuf<int32_t> op(out_point_idx ? (size_t)total : 0), of(out_frame_idx ? (size_t)total : 0);
          k_project_write<<<nb, 128, 0, s>>>(g, fx.p, fy.p, vd.p, frame_idx ? ffr.p : nullptr, point_idx ? fpt.p : nullptr,
                                               offs.p, ox.p, oy.p, mx.p, my.p, op.p, of.p, (int64_t)M);
          LFBA_CUDA(cudaGetLastError());
          ox.download(obs_x, (size_t)total, s);
          oy.download(obs_y, (size_t)total, s);
          mx.download(ml_x, (size_t)total, s);
          my.download(ml_y, (size_t)total, s);
          if (out_point_idx) op.download(out_point_idx, (size_t)total, s);
          if (out_frame_idx) of.download(out_frame_idx, (size_t)total, s);
          LFBA_CUDA(cudaStreamSynchronize(s));
        }
      }
    }  // device buffers are released (stream-ordered) before the stream goes away
    return rc;
  } catch (const CudaError& e) {
    return e.code;
  } catch (const std::exception&) {
    return LFBA_CUDA_ERROR;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// N4: CameraCalibration::initPlenopticParameters (src/CameraCalibration.cpp:456-498) — fL_init = fPH_init * pixelSize,
// then the linear model bL = v B + bL0 over all (frame, feature) pairs, bL = fL Z / (Z - fL) from the camera-frame depth
// of the feature's 3-D point; rows with v < 2 or bL < 0 are zeroed (:483-488). The reference solves the N x 2 system by
// a thin Jacobi SVD; for a full-rank two-column system the minimiser is the centred regression, computed here in two
// fixed-order device passes (means, then centred second moments): no squared condition number, no atomics.
// ------------------------------------------------------------------------------------------------------------------
namespace lfba {
namespace {
__device__ __forceinline__ bool init_row(const double* views, const double* points, const double* vd, const int32_t* fi,
                                         const int32_t* pi, int64_t k, double fL, double& v, double& b) {
  double fe[kFrameStride];
  frame_entry(views + 6 * (size_t)fi[k], fe);
  double pc[3];
  track_point(fe, points + 3 * (size_t)pi[k], pc);
  v = vd[k];
  b = (fL * pc[2]) / (pc[2] - fL);
  return !(v < 2.0 || b < 0.0);
}
// pass 0: n, sum v, sum b over the valid rows; pass 1: sum (v - vm)^2, sum (v - vm)(b - bm). One partial per CTA.
__global__ void __launch_bounds__(256) k_init_sums(const double* views, const double* points, const double* vd,
                                                   const int32_t* fi, const int32_t* pi, int64_t n, double fL, int pass,
                                                   double vm, double bm, double* part) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    double v, b;
    if (!init_row(views, points, vd, fi, pi, k, fL, v, b)) continue;
    if (pass == 0) {
      a0 += 1.0;
      a1 += v;
      a2 += b;
    } else {
      a0 += (v - vm) * (v - vm);
      a1 += (v - vm) * (b - bm);
    }
  }
  __shared__ double red[8][3];
  double vals[3] = {a0, a1, a2};
  for (int q = 0; q < 3; ++q) {
    double x = vals[q];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][q] = x;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double x = 0.0;
    for (int w = 0; w < 8; ++w) x += red[w][threadIdx.x];
    part[3 * blockIdx.x + threadIdx.x] = x;
  }
}
}  // namespace
}  // namespace lfba

extern "C" int lfba_init_plenoptic(double fph_init, double pixel_size_totfoc, int64_t n_pairs, const double* vdepth,
                                   const int32_t* frame_idx, const int32_t* point_idx, int32_t n_frames, const double* views6F,
                                   int32_t n_points, const double* points3P, double* fL_init, double* B_init, double* bL0_init,
                                   int32_t device) {
  if (n_pairs <= 0 || !vdepth || !frame_idx || !point_idx || !views6F || !points3P || !fL_init || !B_init || !bL0_init ||
      n_frames <= 0 || n_points <= 0)
    return LFBA_INVALID_ARGUMENT;
  for (int64_t k = 0; k < n_pairs; ++k)
    if (frame_idx[k] < 0 || frame_idx[k] >= n_frames || point_idx[k] < 0 || point_idx[k] >= n_points) return LFBA_INVALID_ARGUMENT;
  try {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return LFBA_NO_DEVICE;
    if (device >= 0) LFBA_CUDA(cudaSetDevice(device));
    cudaStream_t s;
    LFBA_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    struct SG { cudaStream_t s; ~SG() { cudaStreamSynchronize(s); cudaStreamDestroy(s); } } sg{s};
    alloc_stream() = s;
    const double fL = fph_init * pixel_size_totfoc;  // :460
    int rc = LFBA_OK;
    {
      const size_t n = (size_t)n_pairs;
      const int grid = (int)std::min<size_t>(592, (n + 255) / 256);
      DevBuf<double> vd(n), vw((size_t)6 * n_frames), pt((size_t)3 * n_points), part((size_t)3 * grid);
      DevBuf<int32_t> fi(n), pi(n);
      vd.upload(vdepth, n, s);
      fi.upload(frame_idx, n, s);
      pi.upload(point_idx, n, s);
      vw.upload(views6F, (size_t)6 * n_frames, s);
      pt.upload(points3P, (size_t)3 * n_points, s);
      std::vector<double> h((size_t)3 * grid);
      auto pass = [&](int which, double vm, double bm, double out[3]) {
        k_init_sums<<<grid, 256, 0, s>>>(vw.p, pt.p, vd.p, fi.p, pi.p, n_pairs, fL, which, vm, bm, part.p);
        LFBA_CUDA(cudaGetLastError());
        part.download(h.data(), h.size(), s);
        LFBA_CUDA(cudaStreamSynchronize(s));
        out[0] = out[1] = out[2] = 0.0;
        for (int b = 0; b < grid; ++b)  // fixed order
          for (int q = 0; q < 3; ++q) out[q] += h[3 * (size_t)b + q];
      };
      double m[3], c[3];
      pass(0, 0.0, 0.0, m);
      if (!(m[0] >= 2.0)) {
        rc = LFBA_FAILURE;  // fewer than two valid rows: the reference's SVD would return a rank-deficient minimum-norm answer
      } else {
        const double vm = m[1] / m[0], bm = m[2] / m[0];
        pass(1, vm, bm, c);
        if (!(c[0] > 0.0)) {
          rc = LFBA_FAILURE;
        } else {
          *fL_init = fL;
          *B_init = c[1] / c[0];
          *bL0_init = bm - *B_init * vm;
        }
      }
    }
    return rc;
  } catch (const CudaError& e) {
    return e.code;
  } catch (const std::exception&) {
    return LFBA_CUDA_ERROR;
  }
}

// The web itself (host), for inspection and tests: lines [n][3] = (ex, ey, base-line length), grouped by length.
extern "C" int lfba_epipolar_web(float lens_diameter, float rotation, int32_t rotation_on_grid, int32_t* n_lines,
                                 int32_t* n_groups, double* lines3, int32_t* group_begin) {
  if (!n_lines || !n_groups) return LFBA_INVALID_ARGUMENT;
  std::vector<WebLine> lines;
  std::vector<int32_t> gb;
  build_web(lens_diameter, rotation, rotation_on_grid != 0, lines, gb);
  *n_lines = (int32_t)lines.size();
  *n_groups = (int32_t)gb.size() - 1;
  if (lines3) std::memcpy(lines3, &lines[0].ex, lines.size() * 3 * sizeof(double));
  if (group_begin) std::memcpy(group_begin, gb.data(), gb.size() * sizeof(int32_t));
  return LFBA_OK;
}
