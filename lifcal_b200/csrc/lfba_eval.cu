// lfba_eval.cu — eval-only kernel: residuals and Jacobians of every reprojection block, materialised in
// Ceres' block layout (2x17 | 2x6 | 2x3 row-major per observation), in the caller's observation order.
//
// This is what the reference computes per residual block through ceres::AutoDiffCostFunction
// (src/BundleAdjustment/BundleAdjustment.h:199-222) before Ceres stores it in its block-sparse Jacobian
// (416 B per observation). The LM loop never uses this kernel (its Jacobian stays in registers); it serves
// lfba_eval(): calcReprojectionError (src/CameraCalibration.cpp:1026-1103), parity tests, and the
// "M residual+Jacobian evals/s" metric with the Jacobian written to HBM.
// Algorithmic HBM bytes per observation: read 28 (double2 + 3 x int32), write 16 + 16*(17 + 6 + 3) = 432.
// HBM-bound: the outputs are staged per CTA in shared memory and streamed out coalesced (see stream_out).
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

// Output staging: every thread produces 54 doubles (r 2, J_camera 34, J_view 12, J_point 6) for ITS observation, i.e.
// at a 272 / 96 / 48 / 16-byte stride in HBM. Written directly that is one 32-byte sector per 8-byte store (measured:
// 25% of the HBM roofline). The CTA therefore stages its 128 observations in shared memory (row stride padded to an
// odd number of doubles: conflict-free for the strided writes AND for the linear read-back) and then streams each
// output array out with fully coalesced 8-byte stores: the 128 rows of a CTA are one contiguous range of each array.
constexpr int kEvalBlock = 128;
constexpr int kStrideC = 35, kStrideV = 13, kStrideP = 7, kStrideR = 3;  // padded rows of 34 / 12 / 6 / 2 doubles
constexpr int kEvalSmemDoubles = kEvalBlock * (kStrideC + kStrideV + kStrideP + kStrideR);

template <int W, int STRIDE>
__device__ __forceinline__ void stream_out(const double* __restrict__ sm, double* __restrict__ dst, int rows) {
  // dst[row * W + c] = sm[row * STRIDE + c] for the CTA's `rows` observations, as 16-byte streaming stores: consecutive
  // threads -> consecutive double2 (512 B per warp instruction). W is even, so a pair never straddles two rows; dst is
  // 16-byte aligned because a CTA starts at a multiple of 128 observations.
  static_assert(W % 2 == 0, "row width must be even");
  const int total2 = rows * (W / 2);
  double2* __restrict__ dst2 = reinterpret_cast<double2*>(dst);
  for (int j = threadIdx.x; j < total2; j += kEvalBlock) {
    const int row = j / (W / 2), c = 2 * (j - row * (W / 2));
    const double* p = sm + row * STRIDE + c;
    __stcs(dst2 + j, make_double2(p[0], p[1]));  // written once, never re-read by this kernel
  }
}

template <int NC, int NRAD>
__global__ void __launch_bounds__(kEvalBlock, 4) k_eval_only(Dev d, EvalIn in, EvalOut out, int which) {
  __shared__ CamModel cm;
  __shared__ double sred[4 * 6];
  extern __shared__ double stage[];
  double* sC = stage;
  double* sV = sC + kEvalBlock * kStrideC;
  double* sP = sV + kEvalBlock * kStrideV;
  double* sR = sP + kEvalBlock * kStrideP;
  if (threadIdx.x == 0) cam_model_init(cm, d.camera[which], d.config, d.spx, d.spy, d.scale, d.opt.loss_a);
  __syncthreads();
  const int64_t i0 = blockIdx.x * (int64_t)kEvalBlock;
  const int64_t i = i0 + threadIdx.x;
  double ex2 = 0.0, ey2 = 0.0, mx = 0.0, my = 0.0, inl = 0.0, cost = 0.0;
  if (i < d.N) {
    const double2 o = __ldcs(in.obs + i);
    const int p = __ldcs(in.point_idx + i), f = __ldcs(in.frame_idx + i);
    const double* fe = d.frames[which] + (size_t)f * kFrameStride;
    const double* X = d.points[which] + 3 * (size_t)p;
    const double2* le = reinterpret_cast<const double2*>(d.lens + (size_t)__ldcs(in.lens_id + i) * kLensStride);
    double e[kLensStride];
#pragma unroll
    for (int k = 0; k < kLensStride / 2; ++k) {
      const double2 v2 = __ldg(le + k);
      e[2 * k] = v2.x;
      e[2 * k + 1] = v2.y;
    }
    double Pc[3];
    track_point(fe, X, Pc);
    TrackCtx tc;
    track_setup(cm, Pc, tc);
    double r[2], G[6], J[2 * NC];
    obs_eval<NC, NRAD>(cm, tc, e, o.x, o.y, r, G, J);
    sR[threadIdx.x * kStrideR] = r[0];
    sR[threadIdx.x * kStrideR + 1] = r[1];
    if (out.jac_camera) {
      double* jc = sC + threadIdx.x * kStrideC;
#pragma unroll
      for (int row = 0; row < 2; ++row)
#pragma unroll
        for (int c = 0; c < 17; ++c) jc[17 * row + c] = c < NC ? J[NC * row + c] : 0.0;
    }
    if (out.jac_view) {
      double* jv = sV + threadIdx.x * kStrideV;
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
      double* jp = sP + threadIdx.x * kStrideP;
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
  __syncthreads();  // also orders the staging writes before the read-back
  const int rows = (int)(d.N - i0 < (int64_t)kEvalBlock ? d.N - i0 : (int64_t)kEvalBlock);
  stream_out<2, kStrideR>(sR, out.residuals + 2 * i0, rows);
  if (out.jac_camera) stream_out<34, kStrideC>(sC, out.jac_camera + 34 * i0, rows);
  if (out.jac_view) stream_out<12, kStrideV>(sV, out.jac_view + 12 * i0, rows);
  if (out.jac_point) stream_out<6, kStrideP>(sP, out.jac_point + 6 * i0, rows);
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
  const unsigned grid = (unsigned)((d.N + kEvalBlock - 1) / kEvalBlock);
  const size_t smem = (size_t)kEvalSmemDoubles * sizeof(double);
  static const bool prepared = [] {
    const int b = kEvalSmemDoubles * (int)sizeof(double);
    cudaFuncSetAttribute(k_eval_only<5, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, b);
    cudaFuncSetAttribute(k_eval_only<7, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, b);
    cudaFuncSetAttribute(k_eval_only<6, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, b);
    cudaFuncSetAttribute(k_eval_only<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, b);
    cudaFuncSetAttribute(k_eval_only<7, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, b);
    cudaFuncSetAttribute(k_eval_only<9, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, b);
    return true;
  }();
  (void)prepared;
  const int nrad = (int)(d.config & 3u), tang = (d.config & 0x4u) ? 1 : 0;
  switch (nrad * 2 + tang) {
    case 0: k_eval_only<5, 0><<<grid, kEvalBlock, smem, s>>>(d, in, out, which); break;
    case 1: k_eval_only<7, 0><<<grid, kEvalBlock, smem, s>>>(d, in, out, which); break;
    case 2: k_eval_only<6, 1><<<grid, kEvalBlock, smem, s>>>(d, in, out, which); break;
    case 3: k_eval_only<8, 1><<<grid, kEvalBlock, smem, s>>>(d, in, out, which); break;
    case 4: k_eval_only<7, 2><<<grid, kEvalBlock, smem, s>>>(d, in, out, which); break;
    default: k_eval_only<9, 2><<<grid, kEvalBlock, smem, s>>>(d, in, out, which); break;
  }
}

}  // namespace lfba
