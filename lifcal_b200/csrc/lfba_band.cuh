// lfba_band.cuh — right-looking Cholesky sweep over the 6x6 pose pivots of a banded-arrowhead window that lives in
// shared memory; shared by k_chol_banded (lfba_chol.cu: whole system / separator system) and k_part_forward
// (lfba_chol_part.cu: partition interiors).
//
// What it replaces in the reference: the dense Eigen LLT of Ceres' DenseSchurComplementSolver, on the
// structure the reduced camera system really has with windowed visibility.
//
// The chain of pivots is the serial fraction of the multi-GPU LM round, so the step is built for LATENCY (round 1: about
// 5000 cycles per pivot, shared-memory-bandwidth bound in the trailing update and serialised by index arithmetic):
//   * the window is stored in SLOT space, lower triangle by slot index: frame f owns rows 6 (f mod (bw+1)) .. +5, border
//     rows follow. No per-step row map, no integer division on the critical path; entry (a, b) lives at [max][min].
//   * step = factor (one thread, registers) | panel (one thread per row, result also into a compact 16-byte-aligned
//     X array) | trailing update + refill of the freed pivot slot with the entering frame. Three CTA barriers.
//   * trailing update: a warp owns rows, lanes own columns; a lane keeps ITS columns' X vectors in registers for the whole
//     step and the row's X vector arrives as three broadcast 128-bit loads: about 0.3 shared-memory wavefronts per
//     updated entry instead of 0.9.
//   * the entering frame is prefetched from HBM into registers at the top of the step (in flight during the factorisation)
//     and every entry of the freed slot region is (re)written by exactly one thread: value or zero, no separate clearing.
#pragma once
#include <cstdio>

#include "lfba_device.cuh"

namespace lfba {

struct BandArgs {
  double* S;               // skyline storage of the system being factorised (L is written back in place)
  const int64_t* row_off;  // row offsets of S
  int np6;                 // first border row of S
  int bw;                  // band, in frames
  int fa, fb, hi;          // frames fa..hi enter the window; fa..fb-1 are pivots
  int ns, sp0;             // rows of the previous separator riding along as extra border rows (0 or 6 bw), its first frame
  int nb;                  // border rows of S, the rhs row included
  double* Lsep;            // [F][6 bw][6] panel rows of the previous separator (ns > 0)
  double* dinv_out;        // 1 / L_cc of the pose pivots, indexed 6 f + c (shared or global memory)
  int prof;                // LFBA_DEBUG: block 0 prints the cycles spent per phase of the step
};

constexpr int kBandPref = 5;  // prefetch registers per thread: 6 W <= kBandPref * 224 (seven of the eight warps prefetch)

// Skyline offsets in closed form (no dependent index loads on the latency chain). The solver lays the reduced system out
// as: pose row r = 6 f + i starts at column 6 max(0, f - bw); border rows (coupled points, camera, rhs) start at column 0.
// So frame f's six rows hold 36 min(f, bw) + 21 entries, and everything before frame f is a polynomial in f.
struct SkyMap {
  int bw, np6;
  __host__ __device__ long long frame_off(int f) const {
    return f <= bw ? 18ll * f * (f - 1) + 21ll * f
                   : 18ll * bw * (bw - 1) + 21ll * bw + (long long)(f - bw) * (36 * bw + 21);
  }
  __host__ __device__ int c0(int f) const { return 6 * (f > bw ? f - bw : 0); }
  __host__ __device__ long long row(int f, int i) const {  // offset of entry (6 f + i, c0(f))
    const int len0 = 6 * f - c0(f) + 1;
    return frame_off(f) + (long long)i * len0 + (i * (i - 1)) / 2;
  }
  __host__ __device__ long long border_row(int b) const {  // offset of entry (np6 + b, 0)
    return frame_off(np6 / 6) + (long long)b * (np6 + 1) + ((long long)b * (b - 1)) / 2;
  }
};

__device__ __forceinline__ long long band_sky_row(const BandArgs& g, int f, int i, int& c0) {
  const SkyMap m{g.bw, g.np6};
  c0 = m.c0(f);
  return m.row(f, i);
}
__device__ __forceinline__ long long band_border_row(const BandArgs& g, int b) {
  const SkyMap m{g.bw, g.np6};
  return m.border_row(b);
}

// shared memory needed by band_sweep: A [W][LDW], X [W][6], L_kk (21) + 1/diag (6), padded
__host__ __device__ inline size_t band_smem_doubles(int W) {
  const size_t LDW = (size_t)W | 1;
  size_t a = (size_t)W * LDW;
  a += a & 1;  // X must be 16-byte aligned
  return a + (size_t)W * 6 + 56;
}

// Initial window: frames fa .. fa + bw (those that exist) and their border columns.
__device__ inline void band_load_initial(const BandArgs& g, double* A) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int bw1 = g.bw + 1, NBAND = 6 * bw1, W = NBAND + g.ns + g.nb, LDW = W | 1;
  for (int e = tid; e < W * LDW; e += nt) A[e] = 0.0;
  __syncthreads();
  for (int f = g.fa; f <= min(g.fa + g.bw, g.hi); ++f) {
    int c0;
    const long long r0 = band_sky_row(g, f, 0, c0);
    const int len0 = 6 * f - c0 + 1, n_e = 6 * len0 + 15, sf = 6 * (f % bw1);
    for (int e = tid; e < n_e; e += nt) {
      int i = 0, rs = 0;
      while (i < 5 && e >= rs + len0 + i) { rs += len0 + i; ++i; }
      const int c = c0 + (e - rs), fc = c / 6, jj = c - 6 * fc;
      const double v = g.S[r0 + e];
      if (fc >= g.fa) {
        const int a = sf + i, b = 6 * (fc % bw1) + jj;
        A[max(a, b) * LDW + min(a, b)] = v;
      } else if (g.ns > 0 && fc >= g.sp0) {
        A[(NBAND + 6 * (fc - g.sp0) + jj) * LDW + sf + i] = v;
      }
    }
    for (int e = tid; e < 6 * g.nb; e += nt) {
      const int b = e / 6, i = e - 6 * b;
      A[(NBAND + g.ns + b) * LDW + sf + i] = g.S[band_border_row(g, b) + 6 * f + i];
    }
  }
  __syncthreads();
}

// 1 / sqrt(p) for a normal, positive p without the library's range checks (no branches on the chain): hardware seed
// (MUFU.RSQ64H, about 2^-22) and two Newton steps in double precision — full double accuracy for the pivots seen here.
__device__ __forceinline__ double band_rsqrt(double p) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(p));
  const double hp = 0.5 * p;
  r = r * fma(-hp, r * r, 1.5);
  r = r * fma(-hp, r * r, 1.5);
  return r;
}

// 6x6 Cholesky in registers (one thread); a = lower triangle, row-major packed. Writes L (21) and 1/diag (6) to out[27].
__device__ __forceinline__ bool band_factor6(double* a, double* out, double* dinv_g) {
  bool bad = false;
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    const double piv = a[c * (c + 1) / 2 + c];
    const bool okp = piv > 0.0;
    bad |= !okp;
    const double r = band_rsqrt(okp ? piv : 1.0);
    a[c * (c + 1) / 2 + c] = okp ? piv * r : 1.0;
    out[21 + c] = okp ? r : 0.0;
    dinv_g[c] = okp ? r : 0.0;
#pragma unroll
    for (int i = c + 1; i < 6; ++i) a[i * (i + 1) / 2 + c] *= okp ? r : 0.0;
#pragma unroll
    for (int i = c + 1; i < 6; ++i)
#pragma unroll
      for (int j = c + 1; j <= i; ++j)
        a[i * (i + 1) / 2 + j] = fma(-a[i * (i + 1) / 2 + c], a[j * (j + 1) / 2 + c], a[i * (i + 1) / 2 + j]);
  }
#pragma unroll
  for (int e = 0; e < 21; ++e) out[e] = a[e];
  return bad;
}

// The sweep. NPASS = ceil(W / 32). 256 threads: warp 7 is the factorisation warp (its lane 0 runs the 6x6 Cholesky of the
// NEXT pivot — look-ahead — while warps 0..6 do the trailing update of the current one), warps 0..6 own the panel rows,
// the trailing update, the write-back of L and the refill of the freed slot. Two CTA barriers per pivot.
// On return the window holds the Schur complement on the frames fb..hi and the border rows (slot space, [max][min]).
template <int NPASS>
__device__ inline void band_sweep(const BandArgs& g, double* A, int* s_fail) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int bw = g.bw, bw1 = bw + 1, NBAND = 6 * bw1, Bn = g.ns + g.nb, W = NBAND + Bn, LDW = W | 1;
  size_t aoff = (size_t)W * LDW;
  aoff += aoff & 1;
  double* Xs = A + aoff;              // [W][6], 16-byte aligned rows of 48 bytes
  double* Lsm = Xs + (size_t)W * 6;   // 2 x 28: L_kk lower (21) + 1 / diag (6) of the current / next pivot
  const int FT = nt - 32;             // the factorising thread (lane 0 of the last warp)
  const int nww = (nt >> 5) - 1;      // working warps
  const SkyMap sky{bw, g.np6};

  // this thread's entries of the pivot-slot region (row r of the window, column i of the pivot block): static decode
  const int npf = nt - 32;
  int er[kBandPref], ei[kBandPref];
#pragma unroll
  for (int q = 0; q < kBandPref; ++q) {
    const int e = tid + q * npf;
    er[q] = (tid < npf && e < 6 * W) ? e / 6 : -1;
    ei[q] = e - 6 * (e / 6);
  }
  const int prow = tid < W ? tid : -1;  // this thread's panel row (W <= 160: warps 0..4)
  const int prow_js = prow >= 0 && prow < NBAND ? prow / 6 : -1, prow_i = prow >= 0 ? prow - 6 * (prow / 6) : 0;

  long long pc[4] = {0, 0, 0, 0}, tl = 0;
  const bool prof = g.prof && blockIdx.x == 0 && (tid == FT || tid == 0);
  if (prof) tl = clock64();
#define BAND_TICK(i) if (prof) { const long long tn_ = clock64(); pc[i] += tn_ - tl; tl = tn_; }

  // global loads of the region entries for the frame that enters when pivot k retires (issued one step ahead)
  double pv[kBandPref];
  auto prefetch = [&](int k, int kmod) {
    const int fn = k + bw1;
    const bool enter = fn <= g.hi && k < g.fb;
    const long long row0 = enter ? sky.row(fn, 0) : 0;
    const int c0n = sky.c0(fn), len0 = 6 * fn - c0n + 1;
#pragma unroll
    for (int q = 0; q < kBandPref; ++q) {
      pv[q] = 0.0;
      const int r = er[q], i = ei[q];
      if (!enter || r < 0) continue;
      const long long rowi = row0 + (long long)i * len0 + (i * (i - 1)) / 2;
      if (r < NBAND) {
        const int js = r / 6, jj = r - 6 * js;
        if (js == kmod) {
          if (jj <= i) pv[q] = g.S[rowi + (6 * fn + jj - c0n)];
        } else {
          int dd = js - kmod;
          dd += dd < 0 ? bw1 : 0;
          const int fg = k + dd;  // frame living in that slot: k+1 .. k+bw
          if (fg <= g.hi) pv[q] = g.S[rowi + (6 * fg + jj - c0n)];
        }
      } else if (r >= NBAND + g.ns) {
        pv[q] = g.S[sky.border_row(r - NBAND - g.ns) + 6 * fn + i];
      }
    }
  };

  int kmod = g.fa % bw1;
  if (warp < nww) prefetch(g.fa, kmod);
  if (tid == FT && g.fa < g.fb) {  // first pivot: nothing to look ahead from
    const int s = 6 * kmod;
    double a[21];
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) a[i * (i + 1) / 2 + j] = A[(s + i) * LDW + s + j];
    if (band_factor6(a, Lsm, g.dinv_out + 6 * g.fa)) *s_fail = 1;
  }
  __syncthreads();
  for (int k = g.fa; k < g.fb; ++k) {
    const int s = 6 * kmod;
    const int kmod1 = kmod + 1 == bw1 ? 0 : kmod + 1, s1 = 6 * kmod1;
    const double* Lc = Lsm + 28 * ((k - g.fa) & 1);
    double* Ln = Lsm + 28 * (((k - g.fa) & 1) ^ 1);
    const bool ahead = k + 1 < g.fb;
    // ---- panel: X_r = A(r, pivot columns) L_kk^-T for every other row of the window (empty rows give zeros) ----
    double x[6] = {0, 0, 0, 0, 0, 0};
    const bool prow_on = prow >= 0 && (prow < s || prow >= s + 6);
    if (prow_on) {
      const int r = prow;
      if (r > s) {
#pragma unroll
        for (int c = 0; c < 6; ++c) x[c] = A[r * LDW + s + c];
      } else {
#pragma unroll
        for (int c = 0; c < 6; ++c) x[c] = A[(s + c) * LDW + r];
      }
      double l[21], di[6];
#pragma unroll
      for (int e = 0; e < 21; ++e) l[e] = Lc[e];
#pragma unroll
      for (int c = 0; c < 6; ++c) di[c] = Lc[21 + c];
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        double acc = x[c];
#pragma unroll
        for (int j = 0; j < c; ++j) acc = fma(-x[j], l[c * (c + 1) / 2 + j], acc);
        x[c] = acc * di[c];
      }
      double2* xd = reinterpret_cast<double2*>(Xs + (size_t)r * 6);
      xd[0] = make_double2(x[0], x[1]);
      xd[1] = make_double2(x[2], x[3]);
      xd[2] = make_double2(x[4], x[5]);
    } else if (tid >= nt - 21) {  // L_kk itself goes back to HBM (lanes 11..31 of the factorisation warp)
      const int e = tid - (nt - 21);
      int i = 0, j = e;
      while (j > i) { j -= i + 1; ++i; }
      g.S[sky.row(k, i) + (6 * k + j - sky.c0(k))] = Lc[e];
    }
    BAND_TICK(0)
    __syncthreads();
    BAND_TICK(1)
    if (warp == nww) {
      // ---- look-ahead: the next pivot block gets its update from this step and is factorised right away ----
      if (tid == FT && ahead) {
        double a[21], xr[36];
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
          for (int j = 0; j <= i; ++j) a[i * (i + 1) / 2 + j] = A[(s1 + i) * LDW + s1 + j];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const double2* xd = reinterpret_cast<const double2*>(Xs + (size_t)(s1 + i) * 6);
          const double2 v0 = xd[0], v1 = xd[1], v2 = xd[2];
          xr[6 * i] = v0.x; xr[6 * i + 1] = v0.y; xr[6 * i + 2] = v1.x; xr[6 * i + 3] = v1.y; xr[6 * i + 4] = v2.x; xr[6 * i + 5] = v2.y;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
          for (int j = 0; j <= i; ++j) {
            const double d0 = fma(xr[6 * i], xr[6 * j], fma(xr[6 * i + 1], xr[6 * j + 1], xr[6 * i + 2] * xr[6 * j + 2]));
            const double d1 = fma(xr[6 * i + 3], xr[6 * j + 3], fma(xr[6 * i + 4], xr[6 * j + 4], xr[6 * i + 5] * xr[6 * j + 5]));
            a[i * (i + 1) / 2 + j] -= d0 + d1;
          }
        if (band_factor6(a, Ln, g.dinv_out + 6 * (k + 1))) *s_fail = 1;
      }
    } else {
      // ---- trailing update of every entry outside the pivot slot (and outside the block the look-ahead owns):
      //      A(a, b) -= X_a . X_b, a >= b in slot order; rows in groups of four per warp for instruction-level parallelism
      double xb[NPASS][6];
#pragma unroll
      for (int p = 0; p < NPASS; ++p) {
        const int b = lane + 32 * p;
        if (b < W) {
          const double2* xd = reinterpret_cast<const double2*>(Xs + (size_t)b * 6);
          const double2 v0 = xd[0], v1 = xd[1], v2 = xd[2];
          xb[p][0] = v0.x; xb[p][1] = v0.y; xb[p][2] = v1.x; xb[p][3] = v1.y; xb[p][4] = v2.x; xb[p][5] = v2.y;
        } else {
#pragma unroll
          for (int c = 0; c < 6; ++c) xb[p][c] = 0.0;
        }
      }
      for (int a0 = warp; a0 < W; a0 += 4 * nww) {
        double xa[4][6], acc[4][NPASS];
        bool on[4][NPASS];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int a = a0 + u * nww;
          const bool row_on = a < W && (a < s || a >= s + 6);
          const bool in_next = ahead && a >= s1 && a < s1 + 6;
          const double2* xd = reinterpret_cast<const double2*>(Xs + (size_t)(row_on ? a : 0) * 6);  // broadcast loads
          const double2 v0 = xd[0], v1 = xd[1], v2 = xd[2];
          xa[u][0] = v0.x; xa[u][1] = v0.y; xa[u][2] = v1.x; xa[u][3] = v1.y; xa[u][4] = v2.x; xa[u][5] = v2.y;
#pragma unroll
          for (int p = 0; p < NPASS; ++p) {
            const int b = lane + 32 * p;
            on[u][p] = row_on && b <= a && (b < s || b >= s + 6) && !(in_next && b >= s1);
            acc[u][p] = on[u][p] ? A[(size_t)a * LDW + b] : 0.0;
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int p = 0; p < NPASS; ++p) {
            const double d0 = fma(xa[u][0], xb[p][0], fma(xa[u][1], xb[p][1], xa[u][2] * xb[p][2]));
            const double d1 = fma(xa[u][3], xb[p][3], fma(xa[u][4], xb[p][4], xa[u][5] * xb[p][5]));
            acc[u][p] -= d0 + d1;
          }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int p = 0; p < NPASS; ++p)
            if (on[u][p]) A[(size_t)(a0 + u * nww) * LDW + lane + 32 * p] = acc[u][p];
      }
      // ---- column block k of L goes back to HBM (backward substitution reads it there) ----
      if (prow_on) {
        const int r = prow;
        double* dst = nullptr;
        if (r < NBAND) {
          int dd = prow_js - kmod;
          dd += dd < 0 ? bw1 : 0;
          const int fg = k + dd;
          if (fg <= g.hi) dst = g.S + sky.row(fg, prow_i) + (6 * k - sky.c0(fg));
        } else if (r < NBAND + g.ns) {
          dst = g.Lsep + ((size_t)k * 6 * bw + (r - NBAND)) * 6;
        } else {
          dst = g.S + sky.border_row(r - NBAND - g.ns) + 6 * k;
        }
        if (dst) {
#pragma unroll
          for (int c = 0; c < 6; ++c) dst[c] = x[c];
        }
      }
      // ---- the freed slot takes frame k + bw + 1 (or zeros): every entry of the region is written by its one owner ----
#pragma unroll
      for (int q = 0; q < kBandPref; ++q) {
        const int r = er[q], i = ei[q];
        if (r < 0) continue;
        if (r >= s && r < s + 6) {
          const int jj = r - s;
          if (jj <= i) A[(s + i) * LDW + s + jj] = pv[q];
        } else {
          const int b = s + i;
          A[max(r, b) * LDW + min(r, b)] = pv[q];
        }
      }
      prefetch(k + 1, kmod1);  // in flight during the whole next step
    }
    BAND_TICK(2)
    __syncthreads();
    BAND_TICK(3)
    kmod = kmod1;
  }
  if (prof)
    printf("[lfba dbg] band sweep W=%d pivots=%d thread %d: cycles panel %lld | wait %lld | %s %lld | wait %lld\n", W, g.fb - g.fa,
           tid, pc[0], pc[1], tid == FT ? "look-ahead factor" : "trailing+store+refill+prefetch", pc[2], pc[3]);
#undef BAND_TICK
}

#define LFBA_BAND_DISPATCH(W_, CALL)                              \
  switch (((W_) + 31) / 32) {                                     \
    case 1: { constexpr int NPASS = 1; CALL; } break;             \
    case 2: { constexpr int NPASS = 2; CALL; } break;             \
    case 3: { constexpr int NPASS = 3; CALL; } break;             \
    case 4: { constexpr int NPASS = 4; CALL; } break;             \
    default: { constexpr int NPASS = 5; CALL; } break;            \
  }

}  // namespace lfba
