// lfba_gram2.cu — k_eval_gram2: the fused evaluation pass of the LM loop (default).
//
// What it replaces in the reference: one ceres::Problem::Evaluate over all reprojection blocks
// (AutoDiffCostFunction<OurCostFunctionBundle,2,17,6,3> + CauchyLoss/Corrector, src/BundleAdjustment/BundleAdjustment.h:
// 120-222, src/CameraCalibration.cpp:871-913) fused with the part of SchurEliminator that multiplies Jacobian blocks
// (E^T E, E^T F, F^T F per residual block). The Jacobian never leaves registers.
//
// Per observation a lane evaluates the residual and NC two-component FEATURES (lfba_math.cuh, obs_features9) and adds
// their weighted Gram matrix (NC (NC + 1) / 2 + NC = 54 running sums for NC = 9). When a track (point, frame) ends its L
// lanes combine the sums with xor-shuffles and expand them ONCE into the track record (A, b, C) and the camera block
// (Hcc, gc). Differences to k_eval_gram (lfba_kernels.cu):
//   * NC instead of NC + 1 features (f2 = M q is a per-track combination of f0, f1, f3): 22 fewer DFMA per observation;
//   * Cauchy weighting without square roots: the Gram is accumulated as (w f_a) . f_b with w = rho' = 1 / (1 + s / b)
//     (Corrector scales r and J by sqrt(rho'); only products of two scaled quantities are ever used), and the cost
//     0.5 b sum log(1 + s / b) is taken as ONE log of the running product per lane and track;
//   * the dependent load chain lens_id -> lens-table entry is taken off the critical path: observation and lens id are
//     loaded two steps ahead into registers, the 128-byte lens entry of the next step is prefetched into L1
//     (prefetch.global.L1) one step ahead, so the eight 16-byte gathers of the current step hit L1.
// FP64-pipe bound; algorithmic HBM traffic 20 B per observation + (9 + 3 NC) * 8 B per track.
#include "lfba_device.cuh"
#include "lfba_kernels.h"

namespace lfba {

// prefetch.global.L1 brings in the 32-byte SECTOR that holds the address (measured: one prefetch per 128-byte entry
// lifted the L1 hit rate of the gather from 16% to only 52%), so a lens-table entry takes four.
__device__ __forceinline__ void prefetch_l1(const void* p) {
  asm volatile(
      "prefetch.global.L1 [%0];\n\t"
      "prefetch.global.L1 [%0+32];\n\t"
      "prefetch.global.L1 [%0+64];\n\t"
      "prefetch.global.L1 [%0+96];" ::"l"(p));
}

template <int NV>
__device__ __forceinline__ void block_reduce_store2(double* vals, double* out, double* smem /*[nwarps*NV]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    double s = vals[v];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) smem[warp * NV + v] = s;
  }
  __syncthreads();
  for (int v = threadIdx.x; v < NV; v += blockDim.x) {
    double s = 0.0;
    for (int w = 0; w < nw; ++w) s += smem[w * NV + v];
    out[v] = s;
  }
  __syncthreads();
}

template <int NC, int NRAD, int L>
__global__ void __launch_bounds__(128, 2) k_eval_gram2(Dev d) {
  const LmState* st = d.st;
  if (st->done || st->eval_skip) return;
  const int cand = 1 - st->cur;
  constexpr int NH = NC * (NC + 1) / 2;
  constexpr int NV = NH + NC + 1;
  constexpr int RS = 9 + 3 * NC;
  constexpr int NF9 = Feat9Dims<NC>::NF, NQ9 = Feat9Dims<NC>::NQ, NG9 = Feat9Dims<NC>::NG;
  constexpr int NQ = FeatDims<NC>::NQ, NG = FeatDims<NC>::NG;
  typedef GramMap<NC> GM;
  __shared__ CamModel cm;
  __shared__ double red[4 * NV];
  extern __shared__ double pers[];  // [NV][128]: per-thread camera-block totals (touched once per track)
  if (threadIdx.x == 0) cam_model_init(cm, d.camera[cand], d.config, d.spx, d.spy, d.scale, d.opt.loss_a);
#pragma unroll
  for (int v = 0; v < NV; ++v) pers[v * 128 + threadIdx.x] = 0.0;
  __syncthreads();

  const int lig = threadIdx.x % L;
  const int group = (blockIdx.x * blockDim.x + threadIdx.x) / L;
  const int ngroups = (gridDim.x * blockDim.x) / L;
  const int iters = (d.T + ngroups - 1) / ngroups;
  const double* __restrict__ frames = d.frames[cand];
  const double* __restrict__ points = d.points[cand];
  const double* __restrict__ lens = d.lens;
  double* __restrict__ recs = d.rec[cand];
  const bool robust = cm.robust != 0;
  const double loss_c = cm.loss_c, half_b = 0.5 * cm.loss_b;
  double cost = 0.0;

  auto load_track = [&](int slot, int& tt, int& b, int& e2, bool& ok) {
    ok = slot < d.T;
    tt = ok ? (d.eval_order ? d.eval_order[slot] : slot) : 0;
    b = ok ? d.trk_begin[tt] : 0;
    e2 = ok ? d.trk_begin[tt + 1] : 0;
  };
  int t, ob, oe, nt, nob, noe;
  bool valid, nvalid;
  load_track(group, t, ob, oe, valid);
  load_track(group + ngroups, nt, nob, noe, nvalid);
  // pipeline registers: current (c), next (n), next-next (nn)
  double2 o_c = make_double2(0.0, 0.0), o_n = o_c, o_nn = o_c;
  int lid_c = 0, lid_n = 0, lid_nn = 0;
  bool v_c = false, v_n = false, have_n = true, v_nn = false, have_nn = true;
  {  // prologue: step 0 current, step +1 as next
    const int i0 = ob + lig;
    v_c = valid && i0 < oe;
    if (v_c) {
      o_c = d.obs[i0];
      lid_c = d.lens_id[i0];
    }
    const int ns0 = valid ? (oe - ob + L - 1) / L : 0;
    int i1;
    if (1 < ns0) { i1 = ob + lig + L; v_n = i1 < oe; }
    else { i1 = nob + lig; v_n = nvalid && i1 < noe; }
    if (v_n) {
      o_n = d.obs[i1];
      lid_n = d.lens_id[i1];
    }
  }

  for (int it = 0; it < iters; ++it) {
    double g[NG9];
#pragma unroll
    for (int v = 0; v < NG9; ++v) g[v] = 0.0;
    double prod = 1.0;
    TrackCtx tc;
    const int nsteps = valid ? (oe - ob + L - 1) / L : 0;
    const int nsteps_next = nvalid ? (noe - nob + L - 1) / L : 0;
    if (valid) {
      const int p = d.trk_point[t], f = d.trk_frame[t];
      double Pc[3];
      track_point(frames + (size_t)f * kFrameStride, points + 3 * (size_t)p, Pc);
      track_setup(cm, Pc, tc);
    }
    for (int m = 0; m < nsteps; ++m) {
      if (v_n && !have_n) {  // rare: the look-ahead could not see this item (track of <= L observations)
        const bool same = m + 1 < nsteps;
        const int i1 = same ? ob + lig + L * (m + 1) : nob + lig;
        v_n = same ? (i1 < oe) : (nvalid && i1 < noe);
        have_n = true;
        if (v_n) {
          o_n = d.obs[i1];
          lid_n = d.lens_id[i1];
        }
      }
      // next step's lens entry -> L1 (its lens id was loaded one step ago)
      if (v_n) prefetch_l1(lens + (size_t)lid_n * kLensStride);
      // observation + lens id two steps ahead -> registers
      {
        int i2 = 0;
        have_nn = true;
        if (m + 2 < nsteps) { i2 = ob + lig + L * (m + 2); v_nn = i2 < oe; }
        else if (m + 1 < nsteps) { i2 = nob + lig; v_nn = nvalid && i2 < noe; }
        else if (nsteps_next > 1) { i2 = nob + lig + L; v_nn = nvalid && i2 < noe; }
        else { v_nn = true; have_nn = false; }  // belongs to the round after next: fetched when it becomes "next"
        if (v_nn && have_nn) {
          o_nn = d.obs[i2];
          lid_nn = d.lens_id[i2];
        }
      }
      if (v_c) {
        const double2* lp = reinterpret_cast<const double2*>(lens + (size_t)lid_c * kLensStride);
        double e[kLensStride];
#pragma unroll
        for (int k = 0; k < kLensStride / 2; ++k) {
          const double2 v2 = __ldg(lp + k);
          e[2 * k] = v2.x;
          e[2 * k + 1] = v2.y;
        }
        double r[2], F[2 * NF9];
        obs_features9<NC, NRAD>(cm, tc, e, o_c.x, o_c.y, r, F);
        const double s = r[0] * r[0] + r[1] * r[1];
        double w = 1.0;
        if (robust) {
          const double sum = 1.0 + s * loss_c;
          w = 1.0 / sum;  // rho'
          prod *= sum;
          if (prod > 1e200) {  // keep the running product finite whatever the residuals are
            cost += half_b * log(prod);
            prod = 1.0;
          }
        } else {
          cost += 0.5 * s;
        }
        int q = 0;
#pragma unroll
        for (int a = 0; a < NF9; ++a) {
          const double wx = w * F[a], wy = w * F[NF9 + a];  // row a of the weighted features, live for this row only
#pragma unroll
          for (int b = 0; b <= a; ++b) {
            g[q] = fma(wx, F[b], fma(wy, F[NF9 + b], g[q]));
            ++q;
          }
          g[NQ9 + a] = fma(wx, r[0], fma(wy, r[1], g[NQ9 + a]));
        }
      }
      // rotate
      o_c = o_n;
      lid_c = lid_n;
      v_c = v_n;
      o_n = o_nn;
      lid_n = lid_nn;
      v_n = v_nn;
      have_n = have_nn;
    }
    if (robust && valid) cost += half_b * log(prod);
    if (L > 1) {
#pragma unroll
      for (int v = 0; v < NG9; ++v)
#pragma unroll
        for (int o = L / 2; o > 0; o >>= 1) g[v] += __shfl_xor_sync(0xffffffffu, g[v], o);
    }
    if (valid) {
      // rebuild the sums that involve f2, then expand into the track record and the camera block; the entries are
      // split over the L lanes
      double go[NG];
      gram9_expand<NC>(tc, tc.a1 * cm.gamma, g, go);
      const double* h = go + NQ;
      double* dst = recs + (size_t)t * RS;
      int v = 0;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = i; j < 3; ++j) {
          if ((v % L) == lig) dst[v] = GM::gg(tc, go, i, j);
          ++v;
        }
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        if ((v % L) == lig) dst[v] = (i == 2 ? -tc.g1 : tc.g1) * h[i];
        ++v;
      }
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          if ((v % L) == lig) dst[v] = GM::gcam(tc, go, i, c);
          ++v;
        }
      int hh = 0;
#pragma unroll
      for (int c1 = 0; c1 < NC; ++c1)
#pragma unroll
        for (int c2 = 0; c2 <= c1; ++c2) {
          if ((hh % L) == lig) pers[hh * 128 + threadIdx.x] += GM::cc(tc, go, c1, c2);
          ++hh;
        }
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        if (((NH + c) % L) == lig) {
          double gcv;
          if (c < 3) {
            double a, b;
            GM::geo(tc, c, a, b);
            gcv = a * h[3] + b * h[2];
          } else {
            gcv = h[c + 1];
          }
          pers[(NH + c) * 128 + threadIdx.x] += gcv;
        }
      }
    }
    // next round: the prefetched track becomes current; fetch the one after
    t = nt;
    ob = nob;
    oe = noe;
    valid = nvalid;
    load_track(group + (it + 2) * ngroups, nt, nob, noe, nvalid);
    if (!valid) {  // an invalid round has no steps: the "next" item of the dead round must not leak into a live one
      v_c = false;
      v_n = false;
    }
  }
  double acc[NV];
#pragma unroll
  for (int v = 0; v < NV - 1; ++v) acc[v] = pers[v * 128 + threadIdx.x];
  acc[NV - 1] = cost;
  block_reduce_store2<NV>(acc, d.part_eval + (size_t)blockIdx.x * 64, red);
}

template <int NC, int NRAD>
static void launch_gram2_nc(const Dev& d, int L, int grid, cudaStream_t s) {
  constexpr int NV = NC * (NC + 1) / 2 + NC + 1;
  const size_t smem = (size_t)NV * 128 * sizeof(double);
  switch (L) {
    case 1: k_eval_gram2<NC, NRAD, 1><<<grid, 128, smem, s>>>(d); break;
    case 2: k_eval_gram2<NC, NRAD, 2><<<grid, 128, smem, s>>>(d); break;
    case 4: k_eval_gram2<NC, NRAD, 4><<<grid, 128, smem, s>>>(d); break;
    case 8: k_eval_gram2<NC, NRAD, 8><<<grid, 128, smem, s>>>(d); break;
    default: k_eval_gram2<NC, NRAD, 16><<<grid, 128, smem, s>>>(d); break;
  }
}

template <int NC, int NRAD>
static void prepare_gram2_nc() {
  constexpr int NV = NC * (NC + 1) / 2 + NC + 1;
  const int smem = NV * 128 * (int)sizeof(double);
  cudaFuncSetAttribute(k_eval_gram2<NC, NRAD, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_eval_gram2<NC, NRAD, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_eval_gram2<NC, NRAD, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_eval_gram2<NC, NRAD, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_eval_gram2<NC, NRAD, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
}

void prepare_gram2_kernels() {
  prepare_gram2_nc<5, 0>();
  prepare_gram2_nc<7, 0>();
  prepare_gram2_nc<6, 1>();
  prepare_gram2_nc<8, 1>();
  prepare_gram2_nc<7, 2>();
  prepare_gram2_nc<9, 2>();
}

void launch_eval_gram2(const Dev& d, int L, cudaStream_t s) {
  const int nrad = (int)(d.config & 3u), tang = (d.config & 0x4u) ? 1 : 0;
  switch (nrad * 2 + tang) {
    case 0: launch_gram2_nc<5, 0>(d, L, d.grid_eval, s); break;
    case 1: launch_gram2_nc<7, 0>(d, L, d.grid_eval, s); break;
    case 2: launch_gram2_nc<6, 1>(d, L, d.grid_eval, s); break;
    case 3: launch_gram2_nc<8, 1>(d, L, d.grid_eval, s); break;
    case 4: launch_gram2_nc<7, 2>(d, L, d.grid_eval, s); break;
    default: launch_gram2_nc<9, 2>(d, L, d.grid_eval, s); break;
  }
}

}  // namespace lfba
