// lfba_rows.cu — k_eval_rows: the fused evaluation pass of the LM loop over the PACKED evaluation stream.
//
// What it replaces in the reference: one ceres::Problem::Evaluate over all reprojection blocks
// (AutoDiffCostFunction<OurCostFunctionBundle,2,17,6,3> + CauchyLoss/Corrector, src/BundleAdjustment/BundleAdjustment.h:
// 120-222, src/CameraCalibration.cpp:871-913) fused with the Jacobian-block products of SchurEliminator (E^T E, E^T F,
// F^T F per residual block). The Jacobian never leaves registers.
//
// Arithmetic: NC two-component FEATURES per observation (lfba_math.cuh, obs_features9: every Jacobian column of an
// observation is a per-track combination of them), weighted Gram matrix (54 running sums for NC = 9; Cauchy weight
// w = rho' applied as (w f_a).f_b, no square roots; cost = one log of the running product per lane and track), expanded
// once per track into the track record (A, b, C) and the camera block (Hcc, gc) — tests/test_device_math_cpu.py pins that
// expansion against the Jacobian block products.
//
// Memory system — why the kernel looks like this. ncu on its predecessors (one gather per lane) showed the L1TEX data pipe at 76% with the FP64
// pipe at 31%: every observation gathers its 128-byte lens-table entry as 8 x LDG.128, and with one observation per lane
// each of those warp instructions touches 32 different lines = 32 wavefronts (256 per warp step). Here
//   * the observations are pre-arranged at set-up in the order the kernel consumes them (lfba_setup.cuh, build_stream):
//     a warp streams rows of 32 entries, fully coalesced (512 B + 128 B per row), two rows ahead, with cp.async into a
//     three-slot shared-memory ring;
//   * the lens entries of the NEXT row are gathered cooperatively with cp.async (LDGSTS): the lens ids are shuffled so
//     that 8 consecutive lanes fetch the 8 consecutive 16-byte chunks of ONE entry (4 lines per warp instruction instead
//     of 32) straight into a padded, double-buffered shared-memory tile, from which each lane reads its own entry with
//     conflict-free LDS.128. About 100 instead of 280 L1TEX wavefronts per warp step, and the dependent chain
//     lens id -> lens entry is one row ahead of its use.
// Rounds (32 / L length-adjacent tracks, one per L-lane group) are split evenly by cost (rows + a fixed part per round)
// over the warps of the grid.
#include <cuda_pipeline.h>

#include <mutex>

#include "lfba_device.cuh"
#include "lfba_kernels.h"

namespace lfba {

// Camera model of the candidate parameters in CONSTANT memory: k_tables computes it on the device, launch_eval_rows copies
// it here device-to-device on the solver's stream right before the kernel. Every use becomes a constant-bank operand of
// the FP64 instruction itself instead of a shared-memory load with ~30 cycles of exposed latency (two warps per
// scheduler cannot hide those) and the registers that cached the hot fields are free again.
// The symbol is per DEVICE, solvers are per stream: launch_eval_rows chains every (copy, kernel) pair on a device behind
// the previous evaluation kernel of that device with an event, so two solver handles (or two host threads) on one
// device cannot overwrite each other's model between copy and use.
__constant__ CamModel c_cam;

// 16-byte cp.async that ALLOCATES IN L1 (the __pipeline_memcpy_async form is .cg = L2 only): the evaluation order keeps
// tracks of the same image neighbourhood together, so most lens-table lines a warp gathers were fetched a moment ago.
__device__ __forceinline__ void cp_async_ca16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

constexpr int kRing = 4;     // rows of the packed stream in flight per warp (slot = row % kRing; a power of two)
__device__ __forceinline__ int ring_slot(int row) { return (int)((unsigned)row & (unsigned)(kRing - 1)); }  // row >= 0
constexpr int kLensRow = 33;  // double2 per chunk row of the shared lens tile: 32 lanes + 1 pad (bank spread)

template <int NV>
__device__ __forceinline__ void block_reduce_store_rows(double* vals, double* out, double* smem /*[nwarps*NV]*/) {
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

constexpr int kFlushRows = 8;  // rows between two checks of the running product of the Cauchy sums

// 1 / x for finite x >= 1 (the Cauchy sum 1 + s / a^2): hardware seed (2^-20 or better) + two Newton steps, no range
// checks and no slow path, i.e. no branch inside the evaluation step. Within 1 ulp of the correctly rounded quotient.
__device__ __forceinline__ double rcp_ge1(double x) {
  double y;
  asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));  // volatile: stays where it is called (not sunk into a branch)
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}

// Where entry v of the camera block [Hcc lower triangle row-major | gc] waits at the end of k_eval_rows (one lane per
// track): entries that involve fL, bL0, B (a per-track coefficient) in compact order (c1, c2 <= min(c1, 2)) row by row,
// then gc(0..2), at slots 0 ..; the lens-column block (c1, c2 >= 3), then gc(3 ..), at slots 32 ...
template <int NC>
__device__ __forceinline__ int cam_entry_slot(int v) {
  constexpr int NH = NC * (NC + 1) / 2, NLC = NC - 3;
  if (v >= NH) {
    const int c = v - NH;
    return c < 3 ? 6 + 3 * NLC + c : 32 + NLC * (NLC + 1) / 2 + (c - 3);
  }
  int c1 = 0;
  while ((c1 + 1) * (c1 + 2) / 2 <= v) ++c1;
  const int c2 = v - c1 * (c1 + 1) / 2;
  if (c2 >= 3) return 32 + (c1 - 3) * (c1 - 2) / 2 + (c2 - 3);
  return c1 < 3 ? c1 * (c1 + 1) / 2 + c2 : 6 + 3 * (c1 - 3) + c2;
}

// The rounds are split over the warps of the grid by COST, not by row count: a round costs its rows plus a fixed part
// (track set-up, expansion of the Gram sums, record store, butterfly) worth about kRoundCost rows (ncu: 18% of the stall
// samples on 125k rounds against 82% on 3.38M rows at cfg4). Tracks are sorted by length, so an even split by rows
// gave the warps at the short end of the order several times the rounds of those at the long end, and the kernel
// waited for them (14% of the warp slots idle). The weight is Dev::round_cost. First round r whose cost prefix
// step_base[r] + kRoundCost r >= target:
__device__ __forceinline__ int round_lower_bound(const int32_t* __restrict__ step_base, int R, int64_t target, int kRoundCost) {
  int lo = 0, hi = R;  // answer in [0, R]
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if ((int64_t)step_base[mid] + (int64_t)kRoundCost * mid < target) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// FLAGS: -1 = the model flags are read at run time (L > 1: small problems); else bit 0 = micro-lens adjustment, bit 1 =
// robust loss, compiled in (L = 1: the evaluation step becomes one basic block).
template <int NC, int NRAD, int L, int FLAGS>
__global__ void __launch_bounds__(128, 2) k_eval_rows(Dev d) {
  const LmState* st = d.st;
  if (st->done || st->eval_skip) return;
  const int cand = 1 - st->cur;
  constexpr int G = 32 / L;
  constexpr int NH = NC * (NC + 1) / 2;
  constexpr int NV = NH + NC + 1;
  constexpr int NVL = L == 1 ? 0 : (NV + L - 1) / L;  // L > 1: camera-block totals a lane owns (entries v with v % L == lane % L)
  constexpr int RS = rec_stride(NC);  // padded to a multiple of 4: whole 32-byte sectors
  constexpr int NF9 = Feat9Dims<NC>::NF, NQ9 = Feat9Dims<NC>::NQ, NG9 = Feat9Dims<NC>::NG;
  constexpr int NQ = FeatDims<NC>::NQ, NG = FeatDims<NC>::NG;
  typedef GramMap<NC> GM;
  const CamModel& cm = c_cam;
  __shared__ double red[4 * NV];
  extern __shared__ double dyn[];
  double* pers = dyn;                                              // [NVL][128]
  double2* tiles = reinterpret_cast<double2*>(dyn + NVL * 128);    // [4 warps][2][8][kLensRow]
  double2* ring_o = tiles + 4 * 2 * 8 * kLensRow;                  // [4 warps][kRing][32] observations of rows s .. s+3
  int32_t* ring_l = reinterpret_cast<int32_t*>(ring_o + 4 * kRing * 32);  // [4 warps][kRing][32] their lens ids
  double* tc_save = reinterpret_cast<double*>(ring_l + 4 * kRing * 32);    // [10][128] per-thread track coefficients
#pragma unroll
  for (int v = 0; v < NVL; ++v) pers[v * 128 + threadIdx.x] = 0.0;
  for (int i = threadIdx.x; i < 4 * 2 * 8 * kLensRow; i += 128) tiles[i] = make_double2(0.0, 0.0);  // padding lanes read it
  for (int i = threadIdx.x; i < 4 * kRing * 32; i += 128) ring_l[i] = -1;  // no row yet
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lig = lane % L, grp = lane / L;
  double2* tile = tiles + warp * (2 * 8 * kLensRow);
  double2* my_o = ring_o + warp * (kRing * 32) + lane;
  int32_t* my_l = ring_l + warp * (kRing * 32) + lane;
  const double* __restrict__ frames = d.frames[cand];
  const double* __restrict__ points = d.points[cand];
  const double2* __restrict__ lens2 = reinterpret_cast<const double2*>(d.lens);
  const double2* __restrict__ s_obs = d.s_obs;
  const int32_t* __restrict__ s_lid = d.s_lid;
  const int32_t* __restrict__ step_base = d.step_base;
  double* __restrict__ recs = d.rec[cand];
  const bool robust = FLAGS >= 0 ? (FLAGS & 2) != 0 : cm.robust != 0;
  const bool ml_adjust = FLAGS >= 0 ? (FLAGS & 1) != 0 : cm.ml_adjust != 0;
  constexpr bool any_dist = NC > 5;
  const double loss_c = cm.loss_c, half_b = 0.5 * cm.loss_b;
  double cost = 0.0;
  double camacc = 0.0;  // L == 1: this lane's compact entry of the camera block (see the end of a round)

  // this warp's rounds: an even split of the cost, aligned to round boundaries
  const int R = d.n_rounds;
  const int64_t W = (int64_t)gridDim.x * 4, wg = (int64_t)blockIdx.x * 4 + warp;
  const int64_t total_cost = (int64_t)d.n_rows + (int64_t)d.round_cost * R;
  const int r_begin = round_lower_bound(step_base, R, total_cost * wg / W, d.round_cost);
  const int r_end = round_lower_bound(step_base, R, total_cost * (wg + 1) / W, d.round_cost);
  int row = R > 0 ? step_base[r_begin] : 0;
  const int row_end = R > 0 ? step_base[r_end] : 0;

  // cooperative gather of the lens entries of one row: 8 consecutive lanes fetch the 8 chunks of one entry. The lens ids
  // of the 8 lanes come straight from the ring slot (two broadcast LDS.128); a slot holds -1 until its row has landed.
  const int chunk = lane & 7, lane8 = lane & ~7;
  auto gather_row = [&](const int32_t* slot_ids /* the warp's 32 ids of that row */, double2* buf) {
    const int4 a = *reinterpret_cast<const int4*>(slot_ids + lane8), b = *reinterpret_cast<const int4*>(slot_ids + lane8 + 4);
    const int ids[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (ids[k] >= 0) cp_async_ca16(buf + chunk * kLensRow + lane8 + k, lens2 + (size_t)ids[k] * 8 + chunk);
  };
  int32_t* warp_l = ring_l + warp * (kRing * 32);

  // Software pipeline, all through cp.async (no register rotation: the compiler turns rotated load targets into moves
  // right behind the loads, which puts the full memory latency back on the critical path):
  //   ring slot row % 3 : observation + lens id of the row           (fetched two rows ahead)
  //   tile half row % 2 : the lens entries of the row's 32 lanes     (gathered one row ahead, needs that row's lens ids)
  auto fetch_row = [&](int rw) {  // coalesced: 512 B + 128 B per warp
    if (rw < row_end) {
      __pipeline_memcpy_async(my_o + ring_slot(rw) * 32, s_obs + (size_t)rw * 32 + lane, 16);
      __pipeline_memcpy_async(my_l + ring_slot(rw) * 32, s_lid + (size_t)rw * 32 + lane, 4);
    }
  };
  // inside the row loop (row_end >= 1): no branch, the rows behind the warp's last one re-fetch that one
  auto fetch_row_clamped = [&](int rw) {
    const int src = min(rw, row_end - 1);
    __pipeline_memcpy_async(my_o + ring_slot(rw) * 32, s_obs + (size_t)src * 32 + lane, 16);
    __pipeline_memcpy_async(my_l + ring_slot(rw) * 32, s_lid + (size_t)src * 32 + lane, 4);
  };
  // Copy groups alternate [gather of row x + 1] [stream rows x + 3] per step; "all but the newest group" at the top of
  // step x + 1 is then: the lens entries of row x + 1 (one step old, L1 / L2) and the stream up to row x + 2 (two steps
  // old: DRAM latency has two rows of arithmetic to hide behind).
  fetch_row(row);
  fetch_row(row + 1);
  fetch_row(row + 2);
  __pipeline_commit();
  __pipeline_wait_prior(0);
  __syncwarp();
  gather_row(warp_l + ring_slot(row) * 32, tile + (row & 1) * (8 * kLensRow));
  __pipeline_commit();
  __pipeline_commit();  // (empty: keeps the alternation)

  // Running sums of the round's track: Gram of the NC features + their products with r. With one lane per track the
  // block of the lens columns (features 3 ..: cx, cy, distortion) enters the camera block Hcc / gc as it is — no
  // per-track coefficient — so those (NC-3)(NC-2)/2 + (NC-3) sums simply keep running over ALL tracks of the lane and
  // are reduced once at the end of the kernel; only the other half is reset, expanded and reduced per round.
  constexpr bool kPersist = L == 1;
  constexpr int NLC = NC - 3;                         // lens columns
  constexpr int NPER = NLC * (NLC + 1) / 2 + NLC;     // persistent sums
  constexpr int NCMP = NV - 1 - NPER;                 // camera-block entries reduced per round (<= 27)
  static_assert(NCMP <= 32, "one butterfly chunk");
  double g[NG9];
#pragma unroll
  for (int v = 0; v < NG9; ++v) g[v] = 0.0;
  // (track, point, frame) of this lane's track in the next round: two coalesced loads, one round ahead of their use
  int t_n = 0;
  int2 pf_n = make_int2(0, 0);
  if (r_begin < r_end && r_begin * G + grp < d.T) {
    t_n = d.eval_order[r_begin * G + grp];
    pf_n = d.eval_pf[r_begin * G + grp];
  }
  for (int r = r_begin; r < r_end; ++r) {
    const int nsteps = step_base[r + 1] - step_base[r];
    const int pos = r * G + grp;
    const bool valid = pos < d.T;
    const int t = t_n;
    const int2 pf = pf_n;
    {
      const int pos_n = min(pos + G, d.T - 1);  // (the last round's look-ahead re-reads a valid position)
      if (d.T > 0) {
        t_n = d.eval_order[pos_n];
        pf_n = d.eval_pf[pos_n];
      }
    }
    {
      int q = 0;
#pragma unroll
      for (int a = 0; a < NF9; ++a) {
#pragma unroll
        for (int b = 0; b <= a; ++b) {
          if (!(kPersist && b >= 3)) g[q] = 0.0;
          ++q;
        }
        if (!(kPersist && a >= 3)) g[NQ9 + a] = 0.0;
      }
    }
    double prod = 1.0;
    // Only (wp, kl) of the track context are read per observation; the ten coefficients of the expansion wait in shared
    // memory until the end of the track instead of occupying twenty registers through the row loop.
    double t_wpx, t_wpy, t_kl;
    {
      TrackCtx tc0;
      double Pc[3];
      track_point(frames + (size_t)pf.y * kFrameStride, points + 3 * (size_t)pf.x, Pc);
      track_setup(cm, Pc, tc0);
      t_wpx = tc0.wpx;
      t_wpy = tc0.wpy;
      t_kl = tc0.kl;
      double* sv = tc_save + threadIdx.x;
      sv[0 * 128] = tc0.Px; sv[1 * 128] = tc0.Py; sv[2 * 128] = tc0.a1; sv[3 * 128] = tc0.g1;
      sv[4 * 128] = tc0.af; sv[5 * 128] = tc0.bf; sv[6 * 128] = tc0.ab; sv[7 * 128] = tc0.bb;
      sv[8 * 128] = tc0.aB; sv[9 * 128] = tc0.bB;
    }
    // One row: wait for its copies, read it from shared memory, queue the copies of the rows behind it, and turn it into
    // (features, residual, weight). Straight-line: a padding lane (lens id -1) reads a stale but finite tile entry (the
    // tile starts zeroed) and gets weight 0.
    auto row_features = [&](double* F, double* rr, double& w) {
      __pipeline_wait_prior(1);  // this lane's copies for rows row (lens) and row + 1 (observation) have landed ...
      __syncwarp();              // ... and so have everybody else's; nobody still reads what is refilled next
      // Read everything this step needs from shared memory BEFORE queueing the next copies: the load/store unit is in
      // order, an LDS issued behind the ten LDGSTS below would wait for all of them (measured: 10% of the kernel on
      // the first use of lid_c).
      const int lid_c = my_l[ring_slot(row) * 32];
      const double2 o_c = my_o[ring_slot(row) * 32];
      double e[kLensStride];
      {
        const double2* lp = tile + (row & 1) * (8 * kLensRow) + lane;
#pragma unroll
        for (int k = 0; k < kLensStride / 2; ++k) {
          const double2 v2 = lp[k * kLensRow];
          e[2 * k] = v2.x;
          e[2 * k + 1] = v2.y;
        }
      }
      gather_row(warp_l + ring_slot(row + 1) * 32, tile + ((row + 1) & 1) * (8 * kLensRow));
      __pipeline_commit();
      fetch_row_clamped(row + 3);
      __pipeline_commit();
      // The copies must be ISSUED here, a whole row of arithmetic ahead of their wait. Nothing below depends on them, so
      // inside one basic block the assembler's scheduler is free to sink them to the end of the step (measured: 42% of all
      // stall samples on the wait at the top); a warp-level barrier is a point it does not move memory operations across.
      __syncwarp();
      TrackCtx tc;
      tc.wpx = t_wpx;
      tc.wpy = t_wpy;
      tc.kl = t_kl;
      obs_features9<NC, NRAD>(cm, tc, e, o_c.x, o_c.y, rr, F, ml_adjust, any_dist);
      const bool live = lid_c >= 0;
      const double s = rr[0] * rr[0] + rr[1] * rr[1];
      if (robust) {
        const double sum = live ? fma(s, loss_c, 1.0) : 1.0;
        prod *= sum;
        const double w0 = rcp_ge1(sum);  // rho'
        w = live ? w0 : 0.0;
      } else {
        w = live ? 1.0 : 0.0;
        cost = fma(0.5 * w, s, cost);
      }
      ++row;
    };
    auto gram = [&](const double* F, const double* rr, double w) {
      int q = 0;
#pragma unroll
      for (int a = 0; a < NF9; ++a) {
        const double wx = w * F[a], wy = w * F[NF9 + a];  // row a of the weighted features, live for this row only
#pragma unroll
        for (int b = 0; b <= a; ++b) {
          g[q] = fma(wx, F[b], fma(wy, F[NF9 + b], g[q]));
          ++q;
        }
        g[NQ9 + a] = fma(wx, rr[0], fma(wy, rr[1], g[NQ9 + a]));
      }
    };
    // Software pipeline over the rows of the round: the Gram update of row m (54 independent chains) and the feature
    // chain of row m + 1 (long and narrow) sit in ONE basic block, so each hides the other's latency.
    double Fc[2 * NF9], rc[2], wc;
    row_features(Fc, rc, wc);
    for (int m = 1; m < nsteps; ++m) {
      double Fn[2 * NF9], rn[2], wn;
      row_features(Fn, rn, wn);
      gram(Fc, rc, wc);
#pragma unroll
      for (int k = 0; k < 2 * NF9; ++k) Fc[k] = Fn[k];
      rc[0] = rn[0];
      rc[1] = rn[1];
      wc = wn;
      // keep the running product finite: kFlushRows factors of at most 1 + |r|^2 / a^2 each
      if (robust && (m % kFlushRows) == 0 && prod > 1e100) {
        cost += half_b * log(prod);
        prod = 1.0;
      }
    }
    gram(Fc, rc, wc);
    TrackCtx tc;
    {
      const double* sv = tc_save + threadIdx.x;
      tc.Px = sv[0 * 128]; tc.Py = sv[1 * 128]; tc.a1 = sv[2 * 128]; tc.g1 = sv[3 * 128];
      tc.af = sv[4 * 128]; tc.bf = sv[5 * 128]; tc.ab = sv[6 * 128]; tc.bb = sv[7 * 128];
      tc.aB = sv[8 * 128]; tc.bB = sv[9 * 128];
    }
    if (robust) cost += half_b * log(prod);
    if (L > 1) {
#pragma unroll
      for (int v = 0; v < NG9; ++v)
#pragma unroll
        for (int o = L / 2; o > 0; o >>= 1) g[v] += __shfl_xor_sync(0xffffffffu, g[v], o);
    }
    double cv[L == 1 ? NCMP : 1];  // L == 1: camera-block contribution of this lane's track (zero for an idle lane)
    if constexpr (L == 1) {
#pragma unroll
      for (int v = 0; v < NCMP; ++v) cv[v] = 0.0;
    }
    if (valid) {
      // rebuild the sums that involve f2, then expand into the track record and the camera block; the entries are
      // split over the L lanes of the group
      double go[NG];
      gram9_expand<NC>(tc, tc.a1 * cm.gamma, g, go);
      const double* h = go + NQ;
      double* dst = recs + (size_t)t * RS;
      if (L == 1 && RS % 2 == 0) {
        // one lane owns the whole record: 256-bit (or 128-bit) stores instead of 8-byte ones, which reached L2 as 4.6 GB of
        // partially filled sectors for 1.15 GB of records (ncu l1tex__m_l1tex2xbar_write_bytes); measured -10% kernel time
        double rec_[RS];
#pragma unroll
        for (int k = 9 + 3 * NC; k < RS; ++k) rec_[k] = 0.0;  // padding
        int v = 0;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = i; j < 3; ++j) rec_[v++] = GM::gg(tc, go, i, j);
#pragma unroll
        for (int i = 0; i < 3; ++i) rec_[v++] = (i == 2 ? -tc.g1 : tc.g1) * h[i];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int c = 0; c < NC; ++c) rec_[v++] = GM::gcam(tc, go, i, c);
        if (RS % 4 == 0) {  // records are 32-byte aligned: one full sector per store
#pragma unroll
          for (int k = 0; k < RS / 4; ++k) stg256(dst + 4 * k, rec_[4 * k], rec_[4 * k + 1], rec_[4 * k + 2], rec_[4 * k + 3]);
        } else {
          double2* dst2 = reinterpret_cast<double2*>(dst);
#pragma unroll
          for (int k = 0; k < RS / 2; ++k) dst2[k] = make_double2(rec_[2 * k], rec_[2 * k + 1]);
        }
      } else {
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
      }
      if constexpr (L > 1) {
        int hh = 0;
#pragma unroll
        for (int c1 = 0; c1 < NC; ++c1)
#pragma unroll
          for (int c2 = 0; c2 <= c1; ++c2) {
            if ((hh % L) == lig) pers[(hh / L) * 128 + threadIdx.x] += GM::cc(tc, go, c1, c2);
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
            pers[((NH + c) / L) * 128 + threadIdx.x] += gcv;
          }
        }
      } else {
        // compact order (cam_entry_slot below): (c1, c2 <= min(c1, 2)) row by row, then gc of fL, bL0, B
        int j = 0;
#pragma unroll
        for (int c1 = 0; c1 < NC; ++c1)
#pragma unroll
          for (int c2 = 0; c2 <= (c1 < 2 ? c1 : 2); ++c2) cv[j++] = GM::cc(tc, go, c1, c2);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          double a, b;
          GM::geo(tc, c, a, b);
          cv[j++] = a * h[3] + b * h[2];
        }
      }
    }
    if constexpr (L == 1) {
      // Camera block (Hcc, gc), the entries with a per-track coefficient: the 32 tracks of the round are summed by a
      // reduce-scatter butterfly over the warp; lane l ends up with compact entry brev5(l), which it keeps in a register
      // for the whole kernel. (This replaced a 66 KB per-thread shared-memory accumulator: the freed space is L1 for the
      // lens gather.) Fixed order: deterministic.
      double v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = i < NCMP ? cv[i] : 0.0;
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const int half = 16 >> k;
        const bool up = (lane >> k) & 1;
#pragma unroll
        for (int i = 0; i < half; ++i) {
          const double send = up ? v[i] : v[i + half];
          const double keep = up ? v[i + half] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1 << k);
        }
      }
      camacc += v[0];
    }
  }
  __pipeline_wait_prior(0);
  if constexpr (L == 1) {
    // wsum[warp][slot]: slots 0 .. NCMP-1 the compact entries, 32 .. 32+NPER-1 the persistent ones, 63 the cost
    __shared__ double wsum[4][64];
    const int br = ((lane & 1) << 4) | ((lane & 2) << 2) | (lane & 4) | ((lane & 8) >> 2) | ((lane & 16) >> 4);
    wsum[warp][br] = camacc;
    {
      int k = 0;
#pragma unroll
      for (int part = 0; part < 2; ++part)
#pragma unroll
        for (int a = 3; a < NF9; ++a)
#pragma unroll
          for (int b = 3; b <= (part == 0 ? a : 3); ++b) {  // part 0: Gram (a, b >= 3) row by row; part 1: h(a)
            double x = part == 0 ? g[a * (a + 1) / 2 + b] : g[NQ9 + a];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if (lane == 0) wsum[warp][32 + k] = x;
            ++k;
          }
    }
    double c = cost;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) wsum[warp][63] = c;
    __syncthreads();
    for (int v = threadIdx.x; v < NV; v += blockDim.x) {
      const int slot = v == NV - 1 ? 63 : cam_entry_slot<NC>(v);
      double s_ = 0.0;
      for (int w = 0; w < 4; ++w) s_ += wsum[w][slot];
      d.part_eval[(size_t)blockIdx.x * 64 + v] = s_;
    }
  } else {
    double acc[NV];
#pragma unroll
    for (int v = 0; v < NV - 1; ++v) acc[v] = (v % L) == lig ? pers[(v / L) * 128 + threadIdx.x] : 0.0;
    acc[NV - 1] = cost;
    block_reduce_store_rows<NV>(acc, d.part_eval + (size_t)blockIdx.x * 64, red);
  }
}

template <int NC, int L>
constexpr int rows_smem_bytes() {
  return (L == 1 ? 0 : ((NC * (NC + 1) / 2 + NC + 1) + L - 1) / L) * 128 * (int)sizeof(double) + 4 * 2 * 8 * kLensRow * (int)sizeof(double2) +
         4 * kRing * 32 * ((int)sizeof(double2) + (int)sizeof(int32_t)) + 10 * 128 * (int)sizeof(double);
}

template <int NC, int NRAD>
static void launch_rows_nc(const Dev& d, int L, int grid, cudaStream_t s) {
  switch (L) {
    case 1:
      switch (((d.config & 0x800u) ? 1 : 0) | ((d.config & 0x200u) ? 2 : 0)) {
        case 0: k_eval_rows<NC, NRAD, 1, 0><<<grid, 128, rows_smem_bytes<NC, 1>(), s>>>(d); break;
        case 1: k_eval_rows<NC, NRAD, 1, 1><<<grid, 128, rows_smem_bytes<NC, 1>(), s>>>(d); break;
        case 2: k_eval_rows<NC, NRAD, 1, 2><<<grid, 128, rows_smem_bytes<NC, 1>(), s>>>(d); break;
        default: k_eval_rows<NC, NRAD, 1, 3><<<grid, 128, rows_smem_bytes<NC, 1>(), s>>>(d); break;
      }
      break;
    case 2: k_eval_rows<NC, NRAD, 2, -1><<<grid, 128, rows_smem_bytes<NC, 2>(), s>>>(d); break;
    case 4: k_eval_rows<NC, NRAD, 4, -1><<<grid, 128, rows_smem_bytes<NC, 4>(), s>>>(d); break;
    case 8: k_eval_rows<NC, NRAD, 8, -1><<<grid, 128, rows_smem_bytes<NC, 8>(), s>>>(d); break;
    default: k_eval_rows<NC, NRAD, 16, -1><<<grid, 128, rows_smem_bytes<NC, 16>(), s>>>(d); break;
  }
}

template <int NC, int NRAD>
static void prepare_rows_nc() {
  cudaFuncSetAttribute(k_eval_rows<NC, NRAD, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, rows_smem_bytes<NC, 1>());
  cudaFuncSetAttribute(k_eval_rows<NC, NRAD, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, rows_smem_bytes<NC, 1>());
  cudaFuncSetAttribute(k_eval_rows<NC, NRAD, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, rows_smem_bytes<NC, 1>());
  cudaFuncSetAttribute(k_eval_rows<NC, NRAD, 1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, rows_smem_bytes<NC, 1>());
  cudaFuncSetAttribute(k_eval_rows<NC, NRAD, 2, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, rows_smem_bytes<NC, 2>());
  cudaFuncSetAttribute(k_eval_rows<NC, NRAD, 4, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, rows_smem_bytes<NC, 4>());
  cudaFuncSetAttribute(k_eval_rows<NC, NRAD, 8, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, rows_smem_bytes<NC, 8>());
  cudaFuncSetAttribute(k_eval_rows<NC, NRAD, 16, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, rows_smem_bytes<NC, 16>());
}

void prepare_rows_kernels() {
  prepare_rows_nc<5, 0>();
  prepare_rows_nc<7, 0>();
  prepare_rows_nc<6, 1>();
  prepare_rows_nc<8, 1>();
  prepare_rows_nc<7, 2>();
  prepare_rows_nc<9, 2>();
}

namespace {
std::mutex g_cam_mutex;
cudaEvent_t g_cam_event[64] = {};  // per device: end of the last evaluation kernel that read c_cam
}  // namespace

void launch_eval_rows(const Dev& d, int L, cudaStream_t s) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(g_cam_mutex);
  cudaEvent_t& ev = g_cam_event[dev & 63];
  if (ev) cudaStreamWaitEvent(s, ev, 0);
  else cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
  struct Rec {
    cudaEvent_t e;
    cudaStream_t s;
    ~Rec() { cudaEventRecord(e, s); }
  } rec_at_exit{ev, s};
  cudaMemcpyToSymbolAsync(c_cam, d.cm_buf, sizeof(CamModel), 0, cudaMemcpyDeviceToDevice, s);
  const int nrad = (int)(d.config & 3u), tang = (d.config & 0x4u) ? 1 : 0;
  switch (nrad * 2 + tang) {
    case 0: launch_rows_nc<5, 0>(d, L, d.grid_eval, s); break;
    case 1: launch_rows_nc<7, 0>(d, L, d.grid_eval, s); break;
    case 2: launch_rows_nc<6, 1>(d, L, d.grid_eval, s); break;
    case 3: launch_rows_nc<8, 1>(d, L, d.grid_eval, s); break;
    case 4: launch_rows_nc<7, 2>(d, L, d.grid_eval, s); break;
    default: launch_rows_nc<9, 2>(d, L, d.grid_eval, s); break;
  }
}

}  // namespace lfba
