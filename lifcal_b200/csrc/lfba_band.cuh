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

constexpr int kBandThreads = 512;  // CTA size of the sweep kernels: 15 working warps + the factorisation warp
constexpr int kBandPref = 2;       // prefetch registers per thread: 6 W <= kBandPref * 480

__device__ __forceinline__ long long band_sky_row(const BandArgs& g, int f, int i, int& c0) {
  const SkyMap m{g.bw, g.np6};
  c0 = m.c0(f);
  return m.row(f, i);
}
__device__ __forceinline__ long long band_border_row(const BandArgs& g, int b) {
  const SkyMap m{g.bw, g.np6};
  return m.border_row(b);
}


// Geometry of the window in shared memory: rows padded to a multiple of 8 (tensor tiles), odd leading dimension.
struct BandGeom {
  int W, W8, LDW, NBAND;
  __host__ __device__ explicit BandGeom(int bw, int border_rows) {
    NBAND = 6 * (bw + 1);
    W = NBAND + border_rows;
    W8 = (W + 7) & ~7;
    LDW = W8 | 1;
  }
  __host__ __device__ size_t x_offset() const {  // X [W8][6] behind A, 16-byte aligned
    size_t a = (size_t)W8 * LDW;
    return a + (a & 1);
  }
  __host__ __device__ size_t doubles() const {  // + L slots (2 x 28), snapshot (24), tile list (2 bytes each), private X (36)
    const size_t nT = W8 >> 3, n_tiles = nT * (nT + 1) / 2;
    return x_offset() + (size_t)W8 * 6 + 2 * 28 + 24 + ((n_tiles + 3) >> 2) + 36 + 24 + 8;
  }
};

// Initial window: frames fa .. fa + bw (those that exist) and their border columns.
__device__ inline void band_load_initial(const BandArgs& g, double* A) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const BandGeom geo(g.bw, g.ns + g.nb);
  const int bw1 = g.bw + 1, NBAND = geo.NBAND, LDW = geo.LDW;
  for (int e = tid; e < (int)(geo.x_offset() + (size_t)geo.W8 * 6); e += nt) A[e] = 0.0;  // window and X (padding rows stay zero)
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

// D (8x8) += A (8x4) * B (4x8) on the FP64 tensor path (DMMA): one warp instruction for 256 fused multiply-adds. The
// trailing update of a pivot is X X^T with K = 6: per 8x8 tile of the window two of these replace 64 x 6 scalar DFMAs and,
// more importantly, about 25 instructions per entry of predicate / address / load / store overhead — the sweep was bound by
// instruction issue (ncu: 8900 warp instructions per pivot, 46% issue utilisation with 16 warps), not by arithmetic.
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// The sweep. kBandThreads threads: the last warp is the factorisation warp (its lanes update and its lane 0 factorises
// the NEXT pivot block — look-ahead — while the other warps do the trailing update of the current one); the other warps
// own the panel rows, the tensor-tile trailing update, the write-back of L and the refill of the freed slot.
// Three CTA barriers per pivot. On return the window holds the Schur complement on the frames fb..hi and the border rows
// (slot space, [max][min]).
__device__ inline void band_sweep(const BandArgs& g, double* A, int* s_fail) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int bw = g.bw, bw1 = bw + 1;
  const BandGeom geo(bw, g.ns + g.nb);
  const int NBAND = geo.NBAND, W = geo.W, W8 = geo.W8, LDW = geo.LDW;
  double* Xs = A + geo.x_offset();      // [W8][6], rows of 48 bytes, 16-byte aligned; rows >= W stay zero
  double* Lsm = Xs + (size_t)W8 * 6;    // 2 x 28: L_kk lower (21) + 1 / diag (6) of the current / next pivot
  double* snap = Lsm + 2 * 28;          // 21: the next pivot block as it was before this step's update
  const int FT = nt - 32;               // the factorising thread (lane 0 of the last warp)
  const int nww = (nt >> 5) - 1;        // working warps
  const SkyMap sky{bw, g.np6};
  const int nT = W8 >> 3, n_tiles = nT * (nT + 1) / 2;

  // this thread's entries of the pivot-slot region (row r of the window, column i of the pivot block): static decode
  const int npf = nt - 32;
  int er[kBandPref], ei[kBandPref];
  long long estat[kBandPref];  // static part of the entry's skyline offset
#pragma unroll
  for (int q = 0; q < kBandPref; ++q) {
    const int e = tid + q * npf;
    const bool live = tid < npf && e < 6 * W;
    const int r = live ? e / 6 : -1, i = e - 6 * (e / 6);
    er[q] = (r >= NBAND && r < NBAND + g.ns) ? -2 : r;  // previous-separator rows: always zero for an entering frame
    ei[q] = i;
    if (r >= NBAND + g.ns) estat[q] = sky.border_row(r - NBAND - g.ns) + i;       // + 6 fn
    else estat[q] = (long long)i * (6 * bw + 1) + (i * (i - 1)) / 2;              // + row0(fn) + column (entering frames have a full band)
  }
  const int prow = tid < W ? tid : -1;  // this thread's panel row (W <= 160: warps 0..4)
  const int prow_js = prow >= 0 && prow < NBAND ? prow / 6 : -1, prow_i = prow >= 0 ? prow - 6 * (prow / 6) : 0;

  // global loads of the region entries for the frame that enters when pivot k retires (issued one step ahead)
  double pv[kBandPref];
  long long row0 = sky.frame_off(g.fa + bw1);  // skyline offset of the entering frame's first row; frames > bw: constant stride
  const int fstride = 36 * bw + 21;
  auto prefetch = [&](int k, int kmod) {
    const int fn = k + bw1;
    const bool enter = fn <= g.hi && k < g.fb;
#pragma unroll
    for (int q = 0; q < kBandPref; ++q) {
      pv[q] = 0.0;
      const int r = er[q], i = ei[q];
      if (!enter || r < 0) continue;
      if (r < NBAND) {
        const int js = r / 6, jj = r - 6 * js;
        int dd = js - kmod;
        dd += dd < 0 ? bw1 : 0;        // 0: the entering frame's own block; 1..bw: the frame k + dd living in that slot
        if (dd == 0) {
          if (jj <= i) pv[q] = g.S[row0 + estat[q] + (6 * bw + jj)];
        } else if (k + dd <= g.hi) {
          pv[q] = g.S[row0 + estat[q] + (6 * (dd - 1) + jj)];
        }
      } else {
        pv[q] = g.S[estat[q] + 6 * fn];
      }
    }
  };

  // tile list of the trailing update (lower triangle of 8x8 tiles), decoded once
  unsigned short* tile_ij = reinterpret_cast<unsigned short*>(snap + 24);  // [n_tiles] (I << 8) | J, behind the snapshot
  for (int t = tid; t < n_tiles; t += nt) {
    int I = 0, J = t;
    while (J > I) { J -= I + 1; ++I; }
    tile_ij[t] = (unsigned short)((I << 8) | J);
  }
  double* fx = snap + 24 + ((n_tiles + 3) >> 2);  // 36 + 24: the factorisation warp's own X rows of the next slot, and its copy of the snapshot

  int kmod = g.fa % bw1;
  if (warp < nww) prefetch(g.fa, kmod);
  if (warp == nww && g.fa < g.fb) {  // first pivot: nothing to look ahead from
    const int s = 6 * kmod;
    if (lane == 0) {
      double a[21];
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) a[i * (i + 1) / 2 + j] = A[(s + i) * LDW + s + j];
      if (band_factor6(a, Lsm, g.dinv_out + 6 * g.fa)) *s_fail = 1;
    }
    __syncwarp();
    if (lane < 21) {
      int i = 0, j = lane;
      while (j > i) { j -= i + 1; ++i; }
      g.S[sky.row(g.fa, i) + (6 * g.fa + j - sky.c0(g.fa))] = Lsm[lane];
      if (g.fa + 1 < g.fb) {  // the second pivot block as it is now: input of the first look-ahead
        const int s1 = 6 * (kmod + 1 == bw1 ? 0 : kmod + 1);
        snap[lane] = A[(s1 + i) * LDW + s1 + j];
      }
    }
  }
  __syncthreads();
  // Per pivot k the two groups run side by side and meet at the end of the step:
  //   working warps   panel(k) with L_k | bar A | tensor-tile trailing update, L column block k -> HBM | bar B |
  //                   refill of slot k, snapshot of pivot block k+2, prefetch
  //   last warp       its own X rows of slot k+1 (6 rows), block (k+1) = snapshot - X X^T, Cholesky -> L_{k+1}
  // so the 6x6 factorisation chain (the latency that cannot be parallelised) overlaps the panel AND the trailing update
  // of the same step instead of following the panel. bar A also counts the last warp (arrive only): it has read the
  // pivot-column entries of slot k+1 by then, which the refill behind bar B overwrites.
  for (int k = g.fa; k < g.fb; ++k) {
    const int s = 6 * kmod;
    const int kmod1 = kmod + 1 == bw1 ? 0 : kmod + 1, s1 = 6 * kmod1;
    const int kmod2 = kmod1 + 1 == bw1 ? 0 : kmod1 + 1, s2 = 6 * kmod2;
    const double* Lc = Lsm + 28 * ((k - g.fa) & 1);
    double* Ln = Lsm + 28 * (((k - g.fa) & 1) ^ 1);
    const bool ahead = k + 1 < g.fb;
    if (warp == nww) {
      if (ahead && lane < 6) {
        const int r = s1 + lane;
        double x[6];
        if (r > s) {
#pragma unroll
          for (int c = 0; c < 6; ++c) x[c] = A[r * LDW + s + c];
        } else {
#pragma unroll
          for (int c = 0; c < 6; ++c) x[c] = A[(s + c) * LDW + r];
        }
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          double acc = x[c];
#pragma unroll
          for (int j = 0; j < c; ++j) acc = fma(-x[j], Lc[c * (c + 1) / 2 + j], acc);
          x[c] = acc * Lc[21 + c];
        }
#pragma unroll
        for (int c = 0; c < 6; ++c) fx[6 * lane + c] = x[c];
      }
      if (ahead && lane < 21) fx[36 + lane] = snap[lane];  // private copy: the workers rewrite the snapshot behind bar B
      __syncwarp();
      asm volatile("bar.arrive 1, %0;" ::"r"(nt) : "memory");
      if (ahead) {
        double* scratch = Ln;  // the next pivot's L slot doubles as the exchange buffer (overwritten by the factorisation)
        if (lane < 21) {
          int i = 0, j = lane;
          while (j > i) { j -= i + 1; ++i; }
          const double* xi = fx + 6 * i;
          const double* xj = fx + 6 * j;
          const double d0 = fma(xi[0], xj[0], fma(xi[1], xj[1], xi[2] * xj[2]));
          const double d1 = fma(xi[3], xj[3], fma(xi[4], xj[4], xi[5] * xj[5]));
          scratch[lane] = fx[36 + lane] - (d0 + d1);
        }
        __syncwarp();
        if (lane == 0) {
          double a[21];
#pragma unroll
          for (int e = 0; e < 21; ++e) a[e] = scratch[e];
          if (band_factor6(a, Ln, g.dinv_out + 6 * (k + 1))) *s_fail = 1;
        }
        __syncwarp();
        if (lane < 21) {  // L_{k+1,k+1} itself goes back to HBM
          int i = 0, j = lane;
          while (j > i) { j -= i + 1; ++i; }
          g.S[sky.row(k + 1, i) + (6 * (k + 1) + j - sky.c0(k + 1))] = Ln[lane];
        }
      }
    } else {
      // ---- panel: X_r = A(r, pivot columns) L_kk^-T for every other row of the window (empty rows give zeros);
      //      the pivot's own rows get X = 0 so that the tile update below needs no exclusions ----
      double x[6] = {0, 0, 0, 0, 0, 0};
      const bool prow_on = prow >= 0 && (prow < s || prow >= s + 6);
      if (prow >= 0) {
        const int r = prow;
        if (prow_on) {
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
        }
        double2* xd = reinterpret_cast<double2*>(Xs + (size_t)r * 6);
        xd[0] = make_double2(x[0], x[1]);
        xd[1] = make_double2(x[2], x[3]);
        xd[2] = make_double2(x[4], x[5]);
      }
      asm volatile("bar.sync 1, %0;" ::"r"(nt) : "memory");
      // ---- trailing update, A(a, b) -= X_a . X_b for a >= b in slot order, as 8x8 tensor tiles (two m8n8k4 DMMAs each;
      //      K = 6 padded to 8 with zeros). Diagonal tiles also write their (unused) upper half; rows of the pivot slot and
      //      the padding rows have X = 0 ----
      const int lr = lane >> 2, lc = lane & 3;
      for (int t = warp; t < n_tiles; t += nww) {
        const int ij = tile_ij[t], I = ij >> 8, J = ij & 255;
        const double* xi = Xs + (size_t)(8 * I + lr) * 6;
        const double* xj = Xs + (size_t)(8 * J + lr) * 6;
        const double a0 = -xi[lc], a1 = lc < 2 ? -xi[4 + lc] : 0.0;
        const double b0 = xj[lc], b1 = lc < 2 ? xj[4 + lc] : 0.0;
        double* cp = A + (size_t)(8 * I + lr) * LDW + 8 * J + 2 * lc;
        double c0 = cp[0], c1 = cp[1];
        dmma884(c0, c1, a0, b0);
        dmma884(c0, c1, a1, b1);
        cp[0] = c0;
        cp[1] = c1;
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
      asm volatile("bar.sync 2, %0;" ::"r"(nt - 32) : "memory");
      // ---- the freed slot takes frame k + bw + 1 (or zeros): every entry of the region is written by its one owner ----
#pragma unroll
      for (int q = 0; q < kBandPref; ++q) {
        const int r = er[q] == -2 ? (tid + q * npf) / 6 : er[q], i = ei[q];
        if (r < 0) continue;
        if (r >= s && r < s + 6) {
          const int jj = r - s;
          if (jj <= i) A[(s + i) * LDW + s + jj] = pv[q];
        } else {
          const int b = s + i;
          A[max(r, b) * LDW + min(r, b)] = pv[q];
        }
      }
      if (tid < 21 && k + 2 < g.fb) {  // pivot block k + 2 after this step's update: input of the next look-ahead
        int i = 0, j = tid;
        while (j > i) { j -= i + 1; ++i; }
        snap[tid] = A[(s2 + i) * LDW + s2 + j];
      }
      row0 += fstride;
      prefetch(k + 1, kmod1);  // in flight during the whole next step
    }
    __syncthreads();
    kmod = kmod1;
  }
}

}  // namespace lfba
