// lfba_eval.cu — eval-only kernel: residuals and Jacobians of every reprojection block, materialised in
// Ceres' block layout (2x17 | 2x6 | 2x3 row-major per observation), in the caller's observation order.
//
// This is what the reference computes per residual block through ceres::AutoDiffCostFunction
// (src/BundleAdjustment/BundleAdjustment.h:199-222) before Ceres stores it in its block-sparse Jacobian
// (416 B per observation). The LM loop never uses this kernel (its Jacobian stays in registers); it serves
// lfba_eval(): calcReprojectionError (src/CameraCalibration.cpp:1026-1103), parity tests, and the
// "M residual+Jacobian evals/s" metric with the Jacobian written to HBM.
// Algorithmic HBM bytes per observation: read 28 (double2 + 3 x int32), write 16 + 16 (WC + 6 + 3), WC = 17 (Ceres layout)
// or NC (live columns). HBM-bound: the outputs are staged per CTA in shared memory and leave through the TMA engine.
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

// Output staging: every thread produces 2 + 2 WC + 12 + 6 doubles (r, J_camera, J_view, J_point) for ITS observation, i.e.
// at a 16 / 16 WC / 96 / 48-byte stride in HBM. Written directly that is one 32-byte sector per 8-byte store (measured: 25% of
// the HBM roofline). The CTA therefore stages its 128 observations in shared memory in EXACTLY the layout of the output
// arrays (unpadded: the 128 rows of a CTA are one contiguous range of each array) and one thread hands the four tiles to
// the TMA engine: cp.async.bulk.global.shared::cta (SASS UBLKCP) streams each tile out at full line granularity while the
// SM's load/store pipe is already free for the next CTA. (The predecessor read the staging area back with 8-byte LDS/STG
// pairs: 7.4 G warp instructions and 243 M bank conflicts per launch at the 1M x 1000 scene, L1TEX wavefront pipe 57%.)
// WC = 17 (Ceres' block layout, 8 of 17 columns structurally zero: lfba_eval) or NC (live columns only: the layout SURVEY.md
// 8(d) counts, 344 B per observation at NC = 9).
constexpr int kEvalBlock = 128;
template <int NC, bool COMPACT>
constexpr int cam_width() { return COMPACT ? NC : 17; }
template <int NC, bool COMPACT>
constexpr int eval_smem_doubles() {
  // camera tile first; it also hosts the lens-gather scratch (32 entries x 16 doubles per warp = 4 x 512 doubles)
  return kEvalBlock * (2 * cam_width<NC, COMPACT>() + 12 + 6 + 2) < 4 * 512 + kEvalBlock * 20
             ? 4 * 512 + kEvalBlock * 20
             : kEvalBlock * (2 * cam_width<NC, COMPACT>() + 12 + 6 + 2);
}

__device__ __forceinline__ void bulk_store(double* gdst, const double* ssrc, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes)
               : "memory");
}

template <int NC, int NRAD, bool COMPACT>
__global__ void __launch_bounds__(kEvalBlock, 4) k_eval_only(Dev d, EvalIn in, EvalOut out, int which) {
  __shared__ CamModel cm;
  __shared__ double sred[4 * 6];
  extern __shared__ __align__(128) double stage[];
  constexpr int WC = cam_width<NC, COMPACT>();
  double* sC = stage;                       // [128][2 WC]
  double* sV = sC + kEvalBlock * 2 * WC;    // [128][12]
  double* sP = sV + kEvalBlock * 12;        // [128][6]
  double* sR = sP + kEvalBlock * 6;         // [128][2]
  if (threadIdx.x == 0) cam_model_init(cm, d.camera[which], d.config, d.spx, d.spy, d.scale, d.opt.loss_a);
  __syncthreads();
  const int64_t i0 = blockIdx.x * (int64_t)kEvalBlock;
  const int64_t i = i0 + threadIdx.x;
  double ex2 = 0.0, ey2 = 0.0, mx = 0.0, my = 0.0, inl = 0.0, cost = 0.0;
  // Cooperative gather of the 128-byte lens-table entries of the warp's 32 observations: 8 consecutive lanes fetch the
  // 8 consecutive 16-byte chunks of ONE entry (4 lines per warp instruction instead of 32), the chunks are transposed
  // to their owner through an XOR-swizzled shared-memory scratch (conflict-free both ways). The scratch lives at the
  // start of the staging area, which is only filled afterwards (behind a CTA barrier).
  double ecoop[kLensStride];
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lid = i < d.N ? __ldcs(in.lens_id + i) : -1;
    double2* scratch = reinterpret_cast<double2*>(stage + warp * 512);  // 32 entries x 8 chunks
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
  }
  __syncthreads();  // every warp is done with its scratch: the staging tiles may overwrite it
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
    *reinterpret_cast<double2*>(sR + threadIdx.x * 2) = make_double2(r[0], r[1]);
    if (out.jac_camera) {
      double2* jc = reinterpret_cast<double2*>(sC + threadIdx.x * 2 * WC);  // rows of 2 WC doubles: 16-byte aligned
      double row2[2 * WC];
#pragma unroll
      for (int rr = 0; rr < 2; ++rr)
#pragma unroll
        for (int c = 0; c < WC; ++c) row2[rr * WC + c] = c < NC ? J[rr * NC + c] : 0.0;
#pragma unroll
      for (int k = 0; k < WC; ++k) jc[k] = make_double2(row2[2 * k], row2[2 * k + 1]);
    }
    if (out.jac_view) {
      double jv[12];
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
      double2* dst = reinterpret_cast<double2*>(sV + threadIdx.x * 12);
#pragma unroll
      for (int k = 0; k < 6; ++k) dst[k] = make_double2(jv[2 * k], jv[2 * k + 1]);
    }
    if (out.jac_point) {
      double jp[6];
#pragma unroll
      for (int row = 0; row < 2; ++row)
#pragma unroll
        for (int k = 0; k < 3; ++k)
          jp[3 * row + k] =
              d.refine_points ? G[3 * row] * fe[k] + G[3 * row + 1] * fe[3 + k] + G[3 * row + 2] * fe[6 + k] : 0.0;
      double2* dst = reinterpret_cast<double2*>(sP + threadIdx.x * 6);
#pragma unroll
      for (int k = 0; k < 3; ++k) dst[k] = make_double2(jp[2 * k], jp[2 * k + 1]);
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
  // the staging writes (generic proxy) must be visible to the TMA engine (async proxy) before it reads them
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const unsigned rows = (unsigned)(d.N - i0 < (int64_t)kEvalBlock ? d.N - i0 : (int64_t)kEvalBlock);
  if (threadIdx.x == 0) {
    bulk_store(out.residuals + 2 * i0, sR, rows * 2 * 8);
    if (out.jac_camera) bulk_store(out.jac_camera + (size_t)2 * WC * i0, sC, rows * 2 * WC * 8);
    if (out.jac_view) bulk_store(out.jac_view + 12 * i0, sV, rows * 12 * 8);
    if (out.jac_point) bulk_store(out.jac_point + 6 * i0, sP, rows * 6 * 8);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  if (threadIdx.x < 6 && out.stats) {
    const int v = threadIdx.x;
    double x = sred[v];
    for (int w = 1; w < 4; ++w) x = (v == 2 || v == 3) ? fmax(x, sred[w * 6 + v]) : x + sred[w * 6 + v];
    if (v == 2 || v == 3) atomic_max_double(out.stats + v, x);
    else atomicAdd(out.stats + v, x);
  }
  // the CTA's shared memory must stay allocated until the engine has read it
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// per-device opt-in to > 48 KB of dynamic shared memory: call with that device current (Solver::create does)
template <int NC, int NRAD>
static void prepare_eval_nc() {
  cudaFuncSetAttribute(k_eval_only<NC, NRAD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       eval_smem_doubles<NC, false>() * (int)sizeof(double));
  cudaFuncSetAttribute(k_eval_only<NC, NRAD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       eval_smem_doubles<NC, true>() * (int)sizeof(double));
}
void prepare_eval_only_kernels() {
  prepare_eval_nc<5, 0>();
  prepare_eval_nc<7, 0>();
  prepare_eval_nc<6, 1>();
  prepare_eval_nc<8, 1>();
  prepare_eval_nc<7, 2>();
  prepare_eval_nc<9, 2>();
}

void launch_tables_for(const Dev& d, int which, cudaStream_t s) {
  const int n = d.NL > d.F ? d.NL : d.F;
  if (n > 0) k_tables_for<<<(n + 127) / 128, 128, 0, s>>>(d, which);
}

template <int NC, int NRAD>
static void launch_eval_nc(const Dev& d, const EvalIn& in, const EvalOut& out, int which, unsigned grid, cudaStream_t s) {
  if (out.compact_camera)
    k_eval_only<NC, NRAD, true><<<grid, kEvalBlock, eval_smem_doubles<NC, true>() * sizeof(double), s>>>(d, in, out, which);
  else
    k_eval_only<NC, NRAD, false><<<grid, kEvalBlock, eval_smem_doubles<NC, false>() * sizeof(double), s>>>(d, in, out, which);
}

void launch_eval_only(const Dev& d, const EvalIn& in, const EvalOut& out, int which, cudaStream_t s) {
  if (d.N == 0) return;
  const unsigned grid = (unsigned)((d.N + kEvalBlock - 1) / kEvalBlock);
  const int nrad = (int)(d.config & 3u), tang = (d.config & 0x4u) ? 1 : 0;
  switch (nrad * 2 + tang) {
    case 0: launch_eval_nc<5, 0>(d, in, out, which, grid, s); break;
    case 1: launch_eval_nc<7, 0>(d, in, out, which, grid, s); break;
    case 2: launch_eval_nc<6, 1>(d, in, out, which, grid, s); break;
    case 3: launch_eval_nc<8, 1>(d, in, out, which, grid, s); break;
    case 4: launch_eval_nc<7, 2>(d, in, out, which, grid, s); break;
    default: launch_eval_nc<9, 2>(d, in, out, which, grid, s); break;
  }
}

}  // namespace lfba
