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
constexpr int kStrideV = 13, kStrideP = 7, kStrideR = 3;  // padded rows of 12 / 6 / 2 doubles
// the camera block is staged without its 17 - NC structurally zero columns (they are written as zeros on the way out):
// 2 NC doubles per observation, padded to an odd stride; at least 17 so that a warp's 32 rows hold its 4 KB lens scratch
template <int NC>
constexpr int stride_c() { return 2 * NC + 1 < 17 ? 17 : 2 * NC + 1; }
template <int NC>
constexpr int eval_smem_doubles() { return kEvalBlock * (stride_c<NC>() + kStrideV + kStrideP + kStrideR); }

template <int W, int STRIDE>
__device__ __forceinline__ void stream_out(const double* __restrict__ sm, double* __restrict__ dst, int rows) {
  // dst[row * W + c] = sm[row * STRIDE + c] for the CTA's `rows` observations; consecutive threads -> consecutive doubles:
  // conflict-free 8-byte shared-memory reads (the rows are padded by one double) and 256 contiguous bytes per warp store.
  // (16-byte stores were measured no faster: the odd row stride makes their shared-memory side 2-way conflicted.)
  const int total = rows * W;
  for (int j = threadIdx.x; j < total; j += kEvalBlock) {
    const int row = j / W, c = j - row * W;
    __stcs(dst + j, sm[row * STRIDE + c]);  // streaming store: written once, never re-read by this kernel
  }
}

// camera block: Ceres' 2 x 17 row-major layout from the staged 2 x NC live columns
template <int NC, int STRIDE>
__device__ __forceinline__ void stream_out_camera(const double* __restrict__ sm, double* __restrict__ dst, int rows) {
  const int total = rows * 34;
  for (int j = threadIdx.x; j < total; j += kEvalBlock) {
    const int row = j / 34, c = j - row * 34;
    const int rr = c >= 17 ? 1 : 0, cc = c - 17 * rr;
    __stcs(dst + j, cc < NC ? sm[row * STRIDE + rr * NC + cc] : 0.0);
  }
}

template <int NC, int NRAD>
__global__ void __launch_bounds__(kEvalBlock, 4) k_eval_only(Dev d, EvalIn in, EvalOut out, int which) {
  __shared__ CamModel cm;
  __shared__ double sred[4 * 6];
  extern __shared__ double stage[];
  double* sC = stage;
  constexpr int kStrideC = stride_c<NC>();
  double* sV = sC + kEvalBlock * kStrideC;
  double* sP = sV + kEvalBlock * kStrideV;
  double* sR = sP + kEvalBlock * kStrideP;
  if (threadIdx.x == 0) cam_model_init(cm, d.camera[which], d.config, d.spx, d.spy, d.scale, d.opt.loss_a);
  __syncthreads();
  const int64_t i0 = blockIdx.x * (int64_t)kEvalBlock;
  const int64_t i = i0 + threadIdx.x;
  double ex2 = 0.0, ey2 = 0.0, mx = 0.0, my = 0.0, inl = 0.0, cost = 0.0;
  // Cooperative gather of the 128-byte lens-table entries of the warp's 32 observations: 8 consecutive lanes fetch the
  // 8 consecutive 16-byte chunks of ONE entry (4 lines per warp instruction instead of 32), the chunks are transposed
  // to their owner through an XOR-swizzled shared-memory scratch (conflict-free both ways). The scratch lives in the
  // warp's own part of the output staging area, which it only fills afterwards.
  double ecoop[kLensStride];
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lid = i < d.N ? __ldcs(in.lens_id + i) : -1;
    double2* scratch = reinterpret_cast<double2*>(sC + warp * 32 * kStrideC);  // 32 entries x 8 chunks
    const double2* lens2 = reinterpret_cast<const double2*>(d.lens);
    const int chunk = lane & 7, lane8 = lane & ~7;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int src = lane8 + k;
      const int lk = __shfl_sync(0xffffffffu, lid, src);
      if (lk >= 0) scratch[src * 8 + (chunk ^ k)] = __ldg(lens2 + (size_t)lk * 8 + chunk);
    }
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const double2 v2 = scratch[lane * 8 + (c ^ (lane & 7))];
      ecoop[2 * c] = v2.x;
      ecoop[2 * c + 1] = v2.y;
    }
    __syncwarp();
  }
  if (i < d.N) {
    const double2 o = __ldcs(in.obs + i);
    const int p = __ldcs(in.point_idx + i), f = __ldcs(in.frame_idx + i);
    const double* fe = d.frames[which] + (size_t)f * kFrameStride;
    const double* X = d.points[which] + 3 * (size_t)p;
    double e[kLensStride];
#pragma unroll
    for (int k = 0; k < kLensStride; ++k) e[k] = ecoop[k];
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
      for (int k = 0; k < 2 * NC; ++k) jc[k] = J[k];
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
  if (out.jac_camera) stream_out_camera<NC, kStrideC>(sC, out.jac_camera + 34 * i0, rows);
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

// per-device opt-in to > 48 KB of dynamic shared memory: call with that device current (Solver::create does)
void prepare_eval_only_kernels() {
  cudaFuncSetAttribute(k_eval_only<5, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, eval_smem_doubles<5>() * (int)sizeof(double));
  cudaFuncSetAttribute(k_eval_only<7, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, eval_smem_doubles<7>() * (int)sizeof(double));
  cudaFuncSetAttribute(k_eval_only<6, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, eval_smem_doubles<6>() * (int)sizeof(double));
  cudaFuncSetAttribute(k_eval_only<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, eval_smem_doubles<8>() * (int)sizeof(double));
  cudaFuncSetAttribute(k_eval_only<7, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, eval_smem_doubles<7>() * (int)sizeof(double));
  cudaFuncSetAttribute(k_eval_only<9, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, eval_smem_doubles<9>() * (int)sizeof(double));
}

void launch_tables_for(const Dev& d, int which, cudaStream_t s) {
  const int n = d.NL > d.F ? d.NL : d.F;
  if (n > 0) k_tables_for<<<(n + 127) / 128, 128, 0, s>>>(d, which);
}

void launch_eval_only(const Dev& d, const EvalIn& in, const EvalOut& out, int which, cudaStream_t s) {
  if (d.N == 0) return;
  const unsigned grid = (unsigned)((d.N + kEvalBlock - 1) / kEvalBlock);
  const int nrad = (int)(d.config & 3u), tang = (d.config & 0x4u) ? 1 : 0;
  switch (nrad * 2 + tang) {
    case 0: k_eval_only<5, 0><<<grid, kEvalBlock, eval_smem_doubles<5>() * sizeof(double), s>>>(d, in, out, which); break;
    case 1: k_eval_only<7, 0><<<grid, kEvalBlock, eval_smem_doubles<7>() * sizeof(double), s>>>(d, in, out, which); break;
    case 2: k_eval_only<6, 1><<<grid, kEvalBlock, eval_smem_doubles<6>() * sizeof(double), s>>>(d, in, out, which); break;
    case 3: k_eval_only<8, 1><<<grid, kEvalBlock, eval_smem_doubles<8>() * sizeof(double), s>>>(d, in, out, which); break;
    case 4: k_eval_only<7, 2><<<grid, kEvalBlock, eval_smem_doubles<7>() * sizeof(double), s>>>(d, in, out, which); break;
    default: k_eval_only<9, 2><<<grid, kEvalBlock, eval_smem_doubles<9>() * sizeof(double), s>>>(d, in, out, which); break;
  }
}

}  // namespace lfba
