// lfba_eval.cu — eval-only kernel: residuals and Jacobians of every reprojection block, materialised in
// Ceres' block layout (2x17 | 2x6 | 2x3 row-major per observation), in the caller's observation order.
//
// This is what the reference computes per residual block through ceres::AutoDiffCostFunction
// (src/BundleAdjustment/BundleAdjustment.h:199-222) before Ceres stores it in its block-sparse Jacobian
// (416 B per observation). The LM loop never uses this kernel (its Jacobian stays in registers); it serves
// lfba_eval(): calcReprojectionError (src/CameraCalibration.cpp:1026-1103), parity tests, and the
// "M residual+Jacobian evals/s" metric with the Jacobian written to HBM.
// Algorithmic HBM bytes per observation: read 28 (double2 + 3 x int32), write 16 + 16*(17 + 6 + 3) = 432.
#include "lfba_device.cuh"
#include "lfba_kernels.h"

namespace lfba {

__device__ __forceinline__ double atomic_max_double(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a, assumed;
  do {
    assumed = old;
    if (__longlong_as_double((long long)assumed) >= v) break;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
  } while (assumed != old);
  return __longlong_as_double((long long)old);
}

__global__ void k_tables_for(Dev d, int which) {
  __shared__ CamModel cm;
  if (threadIdx.x == 0) cam_model_init(cm, d.camera[which], d.config, d.spx, d.spy, d.scale, d.opt.loss_a);
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < d.NL) lens_entry(cm, d.lens_xy[2 * i], d.lens_xy[2 * i + 1], d.lens + (size_t)i * kLensStride);
  if (i < d.F) frame_entry(d.views[which] + 6 * i, d.frames[which] + (size_t)i * kFrameStride);
}

template <int NC, int NRAD>
__global__ void __launch_bounds__(128) k_eval_only(Dev d, EvalIn in, EvalOut out, int which) {
  __shared__ CamModel cm;
  __shared__ double sred[4 * 6];
  if (threadIdx.x == 0) cam_model_init(cm, d.camera[which], d.config, d.spx, d.spy, d.scale, d.opt.loss_a);
  __syncthreads();
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  double ex2 = 0.0, ey2 = 0.0, mx = 0.0, my = 0.0, inl = 0.0, cost = 0.0;
  if (i < d.N) {
    const double2 o = in.obs[i];
    const int p = in.point_idx[i], f = in.frame_idx[i];
    const double* fe = d.frames[which] + (size_t)f * kFrameStride;
    const double* X = d.points[which] + 3 * (size_t)p;
    const double* le = d.lens + (size_t)in.lens_id[i] * kLensStride;
    double e[kLensStride];
#pragma unroll
    for (int k = 0; k < kLensStride; ++k) e[k] = le[k];
    double Pc[3];
    track_point(fe, X, Pc);
    TrackCtx tc;
    track_setup(cm, Pc, tc);
    double r[2], G[6], J[2 * NC];
    obs_eval<NC, NRAD>(cm, tc, e, o.x, o.y, r, G, J);
    out.residuals[2 * i] = r[0];
    out.residuals[2 * i + 1] = r[1];
    if (out.jac_camera) {
      double* jc = out.jac_camera + 34 * i;
#pragma unroll
      for (int row = 0; row < 2; ++row)
#pragma unroll
        for (int c = 0; c < 17; ++c) jc[17 * row + c] = c < NC ? J[NC * row + c] : 0.0;
    }
    if (out.jac_view) {
      double* jv = out.jac_view + 12 * i;
      double m[9];
      mat3_vec(fe + 9, X, m + 0);
      mat3_vec(fe + 18, X, m + 3);
      mat3_vec(fe + 27, X, m + 6);
#pragma unroll
      for (int row = 0; row < 2; ++row)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          jv[6 * row + k] =
              d.refine_poses ? G[3 * row] * m[3 * k] + G[3 * row + 1] * m[3 * k + 1] + G[3 * row + 2] * m[3 * k + 2] : 0.0;
          jv[6 * row + 3 + k] = d.refine_poses ? G[3 * row + k] : 0.0;
        }
    }
    if (out.jac_point) {
      double* jp = out.jac_point + 6 * i;
#pragma unroll
      for (int row = 0; row < 2; ++row)
#pragma unroll
        for (int k = 0; k < 3; ++k)
          jp[3 * row + k] =
              d.refine_points ? G[3 * row] * fe[k] + G[3 * row + 1] * fe[3 + k] + G[3 * row + 2] * fe[6 + k] : 0.0;
    }
    const double s = r[0] * r[0] + r[1] * r[1];
    double rho;
    robust_scale(cm, s, rho);
    cost = 0.5 * rho;
    ex2 = r[0] * r[0];
    ey2 = r[1] * r[1];
    mx = fabs(r[0]);
    my = fabs(r[1]);
    inl = s <= out.inlier_thr2 ? 1.0 : 0.0;
  }
  // CTA reduction of the statistics, then one atomic per CTA
  double vals[6] = {ex2, ey2, mx, my, inl, cost};
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int v = 0; v < 6; ++v) {
    double x = vals[v];
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) {
      const double y = __shfl_xor_sync(0xffffffffu, x, o2);
      x = (v == 2 || v == 3) ? fmax(x, y) : x + y;
    }
    if (lane == 0) sred[warp * 6 + v] = x;
  }
  __syncthreads();
  if (threadIdx.x < 6 && out.stats) {
    const int v = threadIdx.x;
    double x = sred[v];
    for (int w = 1; w < 4; ++w) x = (v == 2 || v == 3) ? fmax(x, sred[w * 6 + v]) : x + sred[w * 6 + v];
    if (v == 2 || v == 3) atomic_max_double(out.stats + v, x);
    else atomicAdd(out.stats + v, x);
  }
}

void launch_tables_for(const Dev& d, int which, cudaStream_t s) {
  const int n = d.NL > d.F ? d.NL : d.F;
  if (n > 0) k_tables_for<<<(n + 127) / 128, 128, 0, s>>>(d, which);
}

void launch_eval_only(const Dev& d, const EvalIn& in, const EvalOut& out, int which, cudaStream_t s) {
  if (d.N == 0) return;
  const unsigned grid = (unsigned)((d.N + 127) / 128);
  const int nrad = (int)(d.config & 3u), tang = (d.config & 0x4u) ? 1 : 0;
  switch (nrad * 2 + tang) {
    case 0: k_eval_only<5, 0><<<grid, 128, 0, s>>>(d, in, out, which); break;
    case 1: k_eval_only<7, 0><<<grid, 128, 0, s>>>(d, in, out, which); break;
    case 2: k_eval_only<6, 1><<<grid, 128, 0, s>>>(d, in, out, which); break;
    case 3: k_eval_only<8, 1><<<grid, 128, 0, s>>>(d, in, out, which); break;
    case 4: k_eval_only<7, 2><<<grid, 128, 0, s>>>(d, in, out, which); break;
    default: k_eval_only<9, 2><<<grid, 128, 0, s>>>(d, in, out, which); break;
  }
}

}  // namespace lfba
