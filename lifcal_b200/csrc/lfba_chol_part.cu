// lfba_chol_part.cu — partitioned (nested-dissection) Cholesky of the banded-arrowhead reduced camera system.
//
// Replaces Ceres' DenseSchurComplementSolver::SolveReducedLinearSystem (dense Eigen LLT) for scenes with
// windowed visibility, like lfba_chol.cu's k_chol_banded — whose chain of F sequential 6x6 pivots (about 2.6 us each on
// B200) is the serial fraction of the multi-GPU iteration (2.6 ms at F = 1000). Here the frames are cut into P
// partitions; the last bw frames of every partition but the last are a SEPARATOR. Ordering the unknowns
// [interior_0 .. interior_{P-1} | separators | border] gives the same solution (any elimination order does) with
//   phase 1  P CTAs, one per partition: right-looking factorisation of the interior frames. The rows of the previous
//            separator behave like additional border rows (their fill slides along with the window), so a partition
//            is again a banded-arrowhead problem, with border = [previous separator | coupled points | camera | rhs].
//            What is left in the window at the end — the Schur complement on (this separator, previous separator,
//            border) — is written out as a small dense block.
//   phase 2  one CTA: sum the P blocks (fixed order) into the reduced system over [separators | border], itself banded
//            (block bandwidth 2 bw - 1) with the same border, and solve it with k_chol_banded.
//   phase 3  P CTAs: backward substitution of the interiors.
// The chain shrinks from F pivots to F / P + (P - 1) bw. Results differ from the sequential order only in rounding.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "lfba_device.cuh"
#include "lfba_kernels.h"

namespace lfba {

namespace {

struct PartDev {
  int P, bw, F, npiv;          // partitions, band (frames), frames, border unknowns (without the rhs row)
  const int* bounds;           // [P + 1] first frame of each partition
  double* Lsep;                // [F][6 bw][6] L entries (previous-separator rows, pivot block columns)
  double* X;                   // [P][M * M] extracted Schur complement blocks, M = 12 bw + npiv + 1
  double* dinv;                // [6 F] 1 / L_cc
  int M;
};

__device__ __forceinline__ int pslot(int f, int bw1) { return 6 * (f % bw1); }

struct Sky {
  const Dev& d;
  int bw;
  __device__ __forceinline__ int c0(int f) const { return 6 * max(0, f - bw); }
  __device__ __forceinline__ long long row(int f, int i) const {
    const int len0 = 6 * f - c0(f) + 1;
    return d.row_off[6 * f] + (long long)i * len0 + (i * (i - 1)) / 2;
  }
};

constexpr int kPref = 6;  // prefetch registers per thread for the entering frame (checked on the host)

// ---------------------------------------------------------------------------------------------------------------
// phase 1
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_part_forward(Dev d, PartDev pd) {
  LmState* st = d.st;
  if (linear_phase_idle(st) || !st->solve_ok) return;
  extern __shared__ double sm[];
  const int j = blockIdx.x, P = pd.P, bw = pd.bw, bw1 = bw + 1, F = pd.F;
  const int fa = pd.bounds[j];
  const bool last = j == P - 1;
  const int hi = last ? F - 1 : pd.bounds[j + 1] - 1;   // last frame that enters this partition's window
  const int fb = last ? F : pd.bounds[j + 1] - bw;      // interior = [fa, fb)
  const int ns = j > 0 ? 6 * bw : 0;                    // rows of the previous separator
  const int sp0 = fa - bw;                              // its first frame
  const int nb = pd.npiv + 1;                           // global border rows incl. the rhs row
  const int B = ns + nb;
  const int NBAND = 6 * bw1, W = NBAND + B, LDW = W | 1;
  double* A = sm;
  int* lrow = reinterpret_cast<int*>(A + (size_t)W * LDW);
  __shared__ double dinv[6];
  __shared__ int s_fail;
  const int tid = threadIdx.x, nt = blockDim.x;
  Sky sky{d, bw};
  const int64_t* row_off = d.row_off;
  const int seg_len = 36 * bw + 21;                     // skyline entries of one full-band frame (6 rows)

  for (int e = tid; e < W * LDW; e += nt) A[e] = 0.0;
  if (tid == 0) s_fail = 0;
  __syncthreads();

  // element e of frame f's skyline rows -> (value, position in the window); lidx < 0: nothing to store
  auto band_elem = [&](int f, int e, int& lidx) -> double {
    const int cz = sky.c0(f), len0 = 6 * f - cz + 1;
    int i = 0, rs = 0;
    while (i < 5 && e >= rs + len0 + i) { rs += len0 + i; ++i; }
    const int cc = e - rs;
    if (cc >= len0 + i) { lidx = -1; return 0.0; }
    const int c = cz + cc, fc = c / 6;
    if (fc >= fa) lidx = (pslot(f, bw1) + i) * LDW + pslot(fc, bw1) + (c - 6 * fc);
    else if (ns > 0 && fc >= sp0) lidx = (NBAND + 6 * (fc - sp0) + (c - 6 * fc)) * LDW + pslot(f, bw1) + i;  // transposed
    else lidx = -1;
    return d.S[row_off[6 * f] + e];
  };
  auto load_frame_sync = [&](int f) {
    const int cz = sky.c0(f), n_e = 6 * (6 * f - cz + 1) + 15;
    // previous-separator rows: zero where this frame is not coupled (the slot is reused)
    for (int e = tid; e < ns * 6; e += nt) A[(NBAND + e / 6) * LDW + pslot(f, bw1) + e % 6] = 0.0;
    __syncthreads();
    for (int e = tid; e < n_e; e += nt) {
      int li;
      const double v = band_elem(f, e, li);
      if (li >= 0) A[li] = v;
    }
    for (int e = tid; e < 6 * nb; e += nt)
      A[(NBAND + ns + e / 6) * LDW + pslot(f, bw1) + e % 6] = d.S[row_off[d.np6 + e / 6] + 6 * f + e % 6];
  };
  for (int f = fa; f <= min(fa + bw, hi); ++f) load_frame_sync(f);
  __syncthreads();

  const int tx = tid & 15, ty = tid >> 4;
  for (int k = fa; k < fb; ++k) {
    const int s = pslot(k, bw1);
    const int fn = k + bw1;
    const bool enter = fn <= hi;
    // ---- prefetch the entering frame into registers: in flight during the whole step ----
    double pv[kPref], pbv = 0.0;
    int pl[kPref];
#pragma unroll
    for (int q = 0; q < kPref; ++q) {
      pl[q] = -1;
      pv[q] = 0.0;
      const int e = tid + q * nt;
      if (enter && e < seg_len) pv[q] = band_elem(fn, e, pl[q]);
    }
    if (enter && tid < 6 * nb) pbv = d.S[row_off[d.np6 + tid / 6] + 6 * fn + tid % 6];
    // ---- 6x6 Cholesky of the pivot block (one thread, registers) ----
    if (tid == 0) {
      double a[21];
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int c = 0; c <= i; ++c) a[i * (i + 1) / 2 + c] = A[(s + i) * LDW + s + c];
      bool bad = false;
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        const double piv = a[c * (c + 1) / 2 + c];
        const bool okp = piv > 0.0;
        bad |= !okp;
        const double r = okp ? rsqrt(piv) : 0.0;
        a[c * (c + 1) / 2 + c] = okp ? piv * r : 1.0;
        dinv[c] = r;
        pd.dinv[6 * k + c] = r;
#pragma unroll
        for (int i = c + 1; i < 6; ++i) a[i * (i + 1) / 2 + c] *= r;
#pragma unroll
        for (int i = c + 1; i < 6; ++i)
#pragma unroll
          for (int c2 = c + 1; c2 <= i; ++c2)
            a[i * (i + 1) / 2 + c2] = fma(-a[i * (i + 1) / 2 + c], a[c2 * (c2 + 1) / 2 + c], a[i * (i + 1) / 2 + c2]);
      }
      if (bad) s_fail = 1;
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int c = 0; c <= i; ++c) A[(s + i) * LDW + s + c] = a[i * (i + 1) / 2 + c];
    }
    const int nbf = min(bw, hi - k);
    const int mrows = 6 * nbf + B;
    for (int i = tid; i < mrows; i += nt)
      lrow[i] = i < 6 * nbf ? pslot(k + 1 + i / 6, bw1) + i % 6 : NBAND + (i - 6 * nbf);
    __syncthreads();
    // ---- panel: X = A L_kk^-T for the band rows, the previous-separator rows and the border rows ----
    for (int i = tid; i < mrows; i += nt) {
      const int lr = lrow[i];
      double* dst;
      if (i < 6 * nbf) {
        const int f = k + 1 + i / 6;
        dst = d.S + sky.row(f, i % 6) + (6 * k - sky.c0(f));
      } else if (i < 6 * nbf + ns) {
        dst = pd.Lsep + ((size_t)k * 6 * bw + (i - 6 * nbf)) * 6;
      } else {
        dst = d.S + row_off[d.np6 + (i - 6 * nbf - ns)] + 6 * k;
      }
      double x[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        double a = A[lr * LDW + s + c];
#pragma unroll
        for (int c2 = 0; c2 < c; ++c2) a -= x[c2] * A[(s + c) * LDW + s + c2];
        x[c] = a * dinv[c];
      }
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        A[lr * LDW + s + c] = x[c];
        dst[c] = x[c];
      }
    }
    if (tid >= 224 && tid < 245) {  // L_kk to HBM
      int i = 0, c = tid - 224;
      while (c > i) { c -= i + 1; ++i; }
      d.S[sky.row(k, i) + (6 * k + c - sky.c0(k))] = A[(s + i) * LDW + s + c];
    }
    __syncthreads();
    // ---- trailing update of the window ----
    for (int i = ty; i < mrows; i += 16) {
      const int li = lrow[i];
      const double* xi = A + li * LDW + s;
      const double x0 = xi[0], x1 = xi[1], x2 = xi[2], x3 = xi[3], x4 = xi[4], x5 = xi[5];
      for (int c = tx; c <= i; c += 16) {
        const int lj = lrow[c];
        const double* xj = A + lj * LDW + s;
        A[li * LDW + lj] -= x0 * xj[0] + x1 * xj[1] + x2 * xj[2] + x3 * xj[3] + x4 * xj[4] + x5 * xj[5];
      }
    }
    __syncthreads();
    // ---- slide: frame fn takes the slot of frame k ----
    if (enter) {
      for (int e = tid; e < ns * 6; e += nt) A[(NBAND + e / 6) * LDW + s + e % 6] = 0.0;
      __syncthreads();
#pragma unroll
      for (int q = 0; q < kPref; ++q)
        if (pl[q] >= 0) A[pl[q]] = pv[q];
      if (tid < 6 * nb) A[(NBAND + ns + tid / 6) * LDW + s + tid % 6] = pbv;
      for (int e = tid + nt; e < 6 * nb; e += nt)
        A[(NBAND + ns + e / 6) * LDW + s + e % 6] = d.S[row_off[d.np6 + e / 6] + 6 * fn + e % 6];
      __syncthreads();
    }
  }
  if (tid == 0 && s_fail) st->solve_ok = 0;
  // ---- what is left: Schur complement on [this separator | previous separator | border] ----
  const int nsep = last ? 0 : 6 * bw;
  const int M = nsep + B;
  double* X = pd.X + (size_t)j * pd.M * pd.M;
  auto aidx = [&](int r) -> int {  // extraction index -> window index
    if (r < nsep) return pslot(fb + r / 6, bw1) + r % 6;
    return NBAND + (r - nsep);
  };
  for (int e = tid; e < M * M; e += nt) {
    const int r1 = e / M, r2 = e % M;
    if (r2 <= r1) X[(size_t)r1 * pd.M + r2] = A[aidx(r1) * LDW + aidx(r2)];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// phase 2a: reduced system over [separators | border | rhs] in skyline form (d2)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_part_assemble(Dev d, PartDev pd, Dev d2, long long s2_len) {
  LmState* st = d.st;
  if (linear_phase_idle(st) || !st->solve_ok) return;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int P = pd.P, bw = pd.bw, npiv = pd.npiv, nb = npiv + 1;
  const int nsf = 6 * bw * (P - 1);  // separator unknowns
  for (long long e = tid; e < s2_len; e += nt) d2.S[e] = 0.0;
  __syncthreads();
  auto at2 = [&](int r, int c) -> double* { return d2.S + d2.row_off[r] + (c - d2.row_c0[r]); };
  // original border block (coupled points + camera + the border part of the rhs row)
  for (int e = tid; e < nb * nb; e += nt) {
    const int b1 = e / nb, b2 = e % nb;
    if (b2 <= b1 && !(b1 == npiv && b2 == npiv)) *at2(nsf + b1, nsf + b2) = d.S[d.row_off[d.np6 + b1] + d.np6 + b2];
  }
  __syncthreads();
  for (int j = 0; j < P; ++j) {  // fixed order: deterministic sums
    const bool last = j == P - 1;
    const int nsep = last ? 0 : 6 * bw, ns = j > 0 ? 6 * bw : 0;
    const int M = nsep + ns + nb;
    const double* X = pd.X + (size_t)j * pd.M * pd.M;
    auto ridx = [&](int r) -> int {  // extraction index -> index in the reduced system
      if (r < nsep) return 6 * bw * j + r;
      if (r < nsep + ns) return 6 * bw * (j - 1) + (r - nsep);
      return nsf + (r - nsep - ns);
    };
    for (int e = tid; e < M * M; e += nt) {
      const int r1 = e / M, r2 = e % M;
      if (r2 > r1) continue;
      const int g1 = ridx(r1), g2 = ridx(r2);
      if (g1 == nsf + npiv && g2 == nsf + npiv) continue;
      *at2(max(g1, g2), min(g1, g2)) += X[(size_t)r1 * pd.M + r2];
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// phase 3: backward substitution of the interiors
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_part_backward(Dev d, PartDev pd, const double* __restrict__ y2) {
  LmState* st = d.st;
  if (linear_phase_idle(st) || !st->solve_ok) return;
  extern __shared__ double sm[];
  const int j = blockIdx.x, P = pd.P, bw = pd.bw, F = pd.F, npiv = pd.npiv;
  const int fa = pd.bounds[j];
  const bool last = j == P - 1;
  const int hi = last ? F - 1 : pd.bounds[j + 1] - 1;
  const int fb = last ? F : pd.bounds[j + 1] - bw;
  const int ns = j > 0 ? 6 * bw : 0;
  const int nsf = 6 * bw * (P - 1);
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  double* ys = sm;                          // [6 (hi - fa + 1)] y of this partition's frames (interior + separator)
  double* ysp = ys + 6 * (hi - fa + 1);     // [ns] y of the previous separator
  double* yb = ysp + 6 * bw;                // [npiv] y of the border
  __shared__ double tvec[6];
  __shared__ double Lkk[21];
  Sky sky{d, bw};
  const int64_t* row_off = d.row_off;
  // known parts of the solution (phase 2); every CTA also publishes what it owns into d.y
  if (!last)
    for (int e = tid; e < 6 * bw; e += nt) {
      const double v = y2[6 * bw * j + e];
      ys[6 * (fb - fa) + e] = v;
      d.y[6 * fb + e] = v;
    }
  for (int e = tid; e < ns; e += nt) ysp[e] = y2[6 * bw * (j - 1) + e];
  for (int e = tid; e < npiv; e += nt) {
    const double v = y2[nsf + e];
    yb[e] = v;
    if (j == 0) d.y[d.np6 + e] = v;
  }
  __syncthreads();
  const double* zrow = d.S + row_off[d.np6 + npiv];
  constexpr int per = 8;  // rows per lane: mrows <= 32 * per is checked on the host
  double lv[per];
  double lkk = 0.0, zk = 0.0;
  auto fetch_col = [&](int k) {
    const int nbf = min(bw, hi - k);
    const int mrows = 6 * nbf + ns + npiv;
    if (warp < 6) {
#pragma unroll
      for (int q = 0; q < per; ++q) {
        const int i = lane + 32 * q;
        lv[q] = 0.0;
        if (i < mrows) {
          if (i < 6 * nbf) {
            const int f = k + 1 + i / 6;
            lv[q] = d.S[sky.row(f, i % 6) + (6 * k + warp - sky.c0(f))];
          } else if (i < 6 * nbf + ns) {
            lv[q] = pd.Lsep[((size_t)k * 6 * bw + (i - 6 * nbf)) * 6 + warp];
          } else {
            lv[q] = d.S[row_off[d.np6 + (i - 6 * nbf - ns)] + 6 * k + warp];
          }
        }
      }
      if (lane == 0) zk = zrow[6 * k + warp];
    } else if (warp == 6 && lane < 21) {
      int i = 0, c = lane;
      while (c > i) { c -= i + 1; ++i; }
      lkk = d.S[sky.row(k, i) + (6 * k + c - sky.c0(k))];
    }
  };
  if (fb > fa) fetch_col(fb - 1);
  for (int k = fb - 1; k >= fa; --k) {
    const int nbf = min(bw, hi - k);
    const int mrows = 6 * nbf + ns + npiv;
    if (warp < 6) {
      double acc = 0.0;
#pragma unroll
      for (int q = 0; q < per; ++q) {
        const int i = lane + 32 * q;
        if (i < mrows) {
          const double yv = i < 6 * nbf ? ys[6 * (k + 1 - fa) + i] : (i < 6 * nbf + ns ? ysp[i - 6 * nbf] : yb[i - 6 * nbf - ns]);
          acc += lv[q] * yv;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) tvec[warp] = zk - acc;
    } else if (warp == 6 && lane < 21) {
      Lkk[lane] = lkk;
    }
    if (k > fa) fetch_col(k - 1);
    __syncthreads();
    if (tid == 0) {
      double t[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) t[c] = tvec[c];
#pragma unroll
      for (int c = 5; c >= 0; --c) {
        const double yc = t[c] * pd.dinv[6 * k + c];
        ys[6 * (k - fa) + c] = yc;
#pragma unroll
        for (int c2 = 0; c2 < c; ++c2) t[c2] = fma(-Lkk[c * (c + 1) / 2 + c2], yc, t[c2]);
      }
    }
    __syncthreads();
  }
  for (int e = tid; e < 6 * (fb - fa); e += nt) d.y[6 * fa + e] = ys[e];
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
struct PartPlan {
  bool active = false;
  PartDev pd{};
  Dev d2{};
  int bw2 = 0, nb = 0;
  long long s2_len = 0;
  size_t smem_fwd = 0, smem_bwd = 0, smem_red = 0;
  int *bounds = nullptr, *row_c02 = nullptr;
  long long* row_off2 = nullptr;
  double *Lsep = nullptr, *X = nullptr, *dinv = nullptr, *S2 = nullptr, *y2 = nullptr;
};

size_t banded_smem_bytes(const Dev& d, int bw, int nb);  // lfba_chol.cu
void launch_chol_banded(const Dev& d, int bw, int nb, size_t smem, cudaStream_t s);  // lfba_chol.cu

void part_plan_destroy(PartPlan* p) {
  if (!p) return;
  cudaFree(p->bounds);
  cudaFree(p->row_c02);
  cudaFree(p->row_off2);
  cudaFree(p->Lsep);
  cudaFree(p->X);
  cudaFree(p->dinv);
  cudaFree(p->S2);
  cudaFree(p->y2);
  delete p;
}

// Decides whether the partitioned path applies (enough frames per partition, shared memory) and builds its buffers.
PartPlan* part_plan_create(const Dev& d, int bw, cudaStream_t s) {
  PartPlan* p = new PartPlan();
  const int F = d.np6 / 6, npiv = d.n - d.np6, nb = npiv + 1;
  if (d.np6 == 0 || bw < 1) return p;
  if (const char* e = std::getenv("LFBA_CHOL_PARTS")) {
    if (std::atoi(e) <= 1) return p;
  }
  // chain length ~ F / P + 1.5 (P - 1) bw  ->  P ~ sqrt(F / (1.5 bw)); every interior needs at least bw + 1 frames
  int P = (int)std::lround(std::sqrt((double)F / (1.5 * bw)));
  if (const char* e = std::getenv("LFBA_CHOL_PARTS")) P = std::atoi(e);
  P = std::min(P, F / (2 * bw + 2));
  P = std::min(P, 64);
  if (P < 3) return p;
  if (36 * bw + 21 > kPref * 256) return p;        // prefetch registers of k_part_forward
  if (6 * bw + 6 * bw + npiv > 32 * 8) return p;   // backward-substitution lanes
  std::vector<int> hb((size_t)P + 1);
  for (int j = 0; j <= P; ++j) hb[j] = (int)((long long)F * j / P);
  const int B = 6 * bw + nb, NBAND = 6 * (bw + 1), W = NBAND + B, LDW = W | 1;
  p->smem_fwd = (size_t)W * LDW * sizeof(double) + (size_t)W * sizeof(int) + 16;
  int max_len = 0;
  for (int j = 0; j < P; ++j) max_len = std::max(max_len, hb[j + 1] - hb[j]);
  p->smem_bwd = (size_t)(6 * max_len + 6 * bw + npiv + 8) * sizeof(double);
  if (p->smem_fwd > 200 * 1024 || p->smem_bwd > 200 * 1024) return p;
  // reduced system [separators | border | rhs]
  const int F2 = bw * (P - 1), bw2 = std::min(2 * bw - 1, F2 - 1), n2 = 6 * F2 + npiv, n2_aug = n2 + 1;
  std::vector<int> c0((size_t)n2_aug);
  std::vector<long long> off((size_t)n2_aug + 1);
  for (int r = 0; r < n2_aug; ++r) c0[r] = r < 6 * F2 ? 6 * std::max(0, r / 6 - bw2) : 0;
  off[0] = 0;
  for (int r = 0; r < n2_aug; ++r) off[r + 1] = off[r] + (r - c0[r] + 1);
  p->s2_len = off[n2_aug];
  p->bw2 = bw2;
  p->nb = nb;
  Dev d2 = d;
  d2.np6 = 6 * F2;
  d2.n = n2;
  p->smem_red = banded_smem_bytes(d2, bw2, nb);
  if (p->smem_red == 0) return p;
  const int M = 12 * bw + nb;
  cudaMalloc(&p->bounds, (P + 1) * sizeof(int));
  cudaMalloc(&p->row_c02, n2_aug * sizeof(int));
  cudaMalloc(&p->row_off2, (n2_aug + 1) * sizeof(long long));
  cudaMalloc(&p->Lsep, (size_t)F * 6 * bw * 6 * sizeof(double));
  cudaMalloc(&p->X, (size_t)P * M * M * sizeof(double));
  cudaMalloc(&p->dinv, (size_t)6 * F * sizeof(double));
  cudaMalloc(&p->S2, (size_t)p->s2_len * sizeof(double));
  cudaMalloc(&p->y2, (size_t)n2 * sizeof(double));
  if (!p->bounds || !p->row_c02 || !p->row_off2 || !p->Lsep || !p->X || !p->dinv || !p->S2 || !p->y2) return p;
  cudaMemcpyAsync(p->bounds, hb.data(), (P + 1) * sizeof(int), cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(p->row_c02, c0.data(), n2_aug * sizeof(int), cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(p->row_off2, off.data(), (n2_aug + 1) * sizeof(long long), cudaMemcpyHostToDevice, s);
  cudaMemsetAsync(p->Lsep, 0, (size_t)F * 6 * bw * 6 * sizeof(double), s);
  cudaMemsetAsync(p->X, 0, (size_t)P * M * M * sizeof(double), s);
  cudaStreamSynchronize(s);  // the host vectors go out of scope
  d2.S = p->S2;
  d2.row_off = reinterpret_cast<const int64_t*>(p->row_off2);
  d2.row_c0 = p->row_c02;
  d2.y = p->y2;
  p->d2 = d2;
  p->pd = PartDev{P, bw, F, npiv, p->bounds, p->Lsep, p->X, p->dinv, M};
  cudaFuncSetAttribute(k_part_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_fwd);
  cudaFuncSetAttribute(k_part_backward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_bwd);
  p->active = true;
  return p;
}

bool part_plan_active(const PartPlan* p) { return p && p->active; }

// the three phases on stream s; returns the number of kernels launched
int launch_part_solve(const Dev& d, PartPlan* p, cudaStream_t s) {
  k_part_forward<<<p->pd.P, 256, p->smem_fwd, s>>>(d, p->pd);
  Dev d2 = p->d2;
  d2.st = d.st;
  k_part_assemble<<<1, 256, 0, s>>>(d, p->pd, d2, p->s2_len);
  launch_chol_banded(d2, p->bw2, p->nb, p->smem_red, s);
  k_part_backward<<<p->pd.P, 256, p->smem_bwd, s>>>(d, p->pd, p->y2);
  return 4;
}

}  // namespace lfba
