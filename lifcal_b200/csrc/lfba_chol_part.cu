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

#include "lfba_band.cuh"
#include "lfba_device.cuh"
#include "lfba_kernels.h"
#include "lfba_setup.cuh"

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

// ---------------------------------------------------------------------------------------------------------------
// phase 1
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBandThreads) k_part_forward(Dev d, PartDev pd) {
  LmState* st = d.st;
  if (linear_phase_idle(st) || !st->solve_ok) return;
  extern __shared__ __align__(16) double sm[];
  const int j = blockIdx.x, P = pd.P, bw = pd.bw, bw1 = bw + 1, F = pd.F;
  const int fa = pd.bounds[j];
  const bool last = j == P - 1;
  const int hi = last ? F - 1 : pd.bounds[j + 1] - 1;   // last frame that enters this partition's window
  const int fb = last ? F : pd.bounds[j + 1] - bw;      // interior = [fa, fb)
  const int ns = j > 0 ? 6 * bw : 0;                    // rows of the previous separator
  const int nb = pd.npiv + 1;                           // global border rows incl. the rhs row
  const int B = ns + nb;
  const BandGeom geo(bw, B);
  const int NBAND = geo.NBAND, LDW = geo.LDW;
  double* A = sm;
  __shared__ int s_fail;
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) s_fail = 0;
  BandArgs g{d.S, d.row_off, d.np6, bw, fa, fb, hi, ns, fa - bw, nb, pd.Lsep, pd.dinv, d.debug};
  band_load_initial(g, A);
  band_sweep(g, A, &s_fail);
  if (tid == 0 && s_fail) st->solve_ok = 0;
  // ---- what is left: Schur complement on [this separator | previous separator | border] (slot space, [max][min]) ----
  const int nsep = last ? 0 : 6 * bw;
  const int M = nsep + B;
  double* X = pd.X + (size_t)j * pd.M * pd.M;
  auto aidx = [&](int r) -> int {  // extraction index -> window index
    if (r < nsep) return pslot(fb + r / 6, bw1) + r % 6;
    return NBAND + (r - nsep);
  };
  for (int e = tid; e < M * M; e += nt) {
    const int r1 = e / M, r2 = e % M;
    if (r2 <= r1) {
      const int a1 = aidx(r1), a2 = aidx(r2);
      X[(size_t)r1 * pd.M + r2] = A[max(a1, a2) * LDW + min(a1, a2)];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// phase 2a: reduced system over [separators | border | rhs] in skyline form (d2)
// ---------------------------------------------------------------------------------------------------------------
// One thread per entry of the reduced skyline S2: it gathers its contributions (original border block; the extracted
// blocks of the at most two partitions that touch a separator entry, of all P partitions for a border entry) in
// partition order: deterministic, no atomics, no barriers.
__global__ void __launch_bounds__(256) k_part_assemble(Dev d, PartDev pd, Dev d2, long long s2_len) {
  LmState* st = d.st;
  if (linear_phase_idle(st) || !st->solve_ok) return;
  const int P = pd.P, bw = pd.bw, npiv = pd.npiv, nb = npiv + 1;
  const int sw = 6 * bw;             // separator width
  const int nsf = sw * (P - 1);      // separator unknowns
  const int n2_aug = nsf + nb;
  const SkyMap sky{bw, d.np6};
  for (int r = blockIdx.x; r < n2_aug; r += gridDim.x) {
    const int c0 = d2.row_c0[r];
    const long long ro = d2.row_off[r];
    for (int c = c0 + threadIdx.x; c <= r; c += blockDim.x) {
      if (r == nsf + npiv && c == nsf + npiv) continue;  // the rhs row has no diagonal entry
      double v = 0.0;
      auto xat = [&](int j, int r1, int r2) -> double {  // lower-triangular read of partition j's block
        const double* X = pd.X + (size_t)j * pd.M * pd.M;
        return r1 >= r2 ? X[(size_t)r1 * pd.M + r2] : X[(size_t)r2 * pd.M + r1];
      };
      // extraction index of reduced index q inside partition j's block, or -1
      auto xidx = [&](int j, int q) -> int {
        const bool last = j == P - 1;
        const int nsep = last ? 0 : sw, ns = j > 0 ? sw : 0;
        if (q >= nsf) return nsep + ns + (q - nsf);                           // border
        const int sj = q / sw, o = q - sw * sj;                               // separator sj
        if (!last && sj == j) return o;                                       // this partition's own separator
        if (j > 0 && sj == j - 1) return nsep + o;                            // the previous separator
        return -1;
      };
      if (r >= nsf && c >= nsf) {  // border x border: original block + every partition
        v = d.S[sky.border_row(r - nsf) + d.np6 + (c - nsf)];
        for (int j = 0; j < P; ++j) v += xat(j, xidx(j, r), xidx(j, c));
      } else {
        // at least one separator index: only the partitions adjacent to that separator contribute
        const int sc = c / sw;  // c < nsf here (c <= r and not both border)
        for (int j = sc; j <= min(sc + 1, P - 1); ++j) {
          const int i1 = xidx(j, r), i2 = xidx(j, c);
          if (i1 >= 0 && i2 >= 0) v += xat(j, i1, i2);
        }
      }
      d2.S[ro + (c - c0)] = v;
    }
  }
  (void)s2_len;
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
  __shared__ double Lkk[27];
  const SkyMap sky{bw, d.np6};
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
  const double* zrow = d.S + sky.border_row(npiv);
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
            lv[q] = d.S[sky.border_row(i - 6 * nbf - ns) + 6 * k + warp];
          }
        }
      }
      if (lane == 0) zk = zrow[6 * k + warp];
    } else if (warp == 6 && lane < 21) {
      int i = 0, c = lane;
      while (c > i) { c -= i + 1; ++i; }
      lkk = d.S[sky.row(k, i) + (6 * k + c - sky.c0(k))];
    } else if (warp == 6 && lane < 27) {
      lkk = pd.dinv[6 * k + (lane - 21)];  // 1 / L_cc: fetched one step ahead with the rest, not inside the solve chain
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
    } else if (warp == 6 && lane < 27) {
      Lkk[lane] = lkk;  // 21 entries of L_kk, then the six 1 / L_cc
    }
    if (k > fa) fetch_col(k - 1);
    __syncthreads();
    if (tid == 0) {
      double t[6], dv[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        t[c] = tvec[c];
        dv[c] = Lkk[21 + c];
      }
#pragma unroll
      for (int c = 5; c >= 0; --c) {
        const double yc = t[c] * dv[c];
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
  DevBuf<int> bounds, row_c02;
  DevBuf<long long> row_off2;
  DevBuf<double> Lsep, X, dinv, S2, y2;  // stream-ordered, from the library's block cache (no cudaMalloc / cudaFree per solve)
};

size_t banded_smem_bytes(const Dev& d, int bw, int nb);  // lfba_chol.cu
void launch_chol_banded(const Dev& d, int bw, int nb, size_t smem, cudaStream_t s);  // lfba_chol.cu

void part_plan_destroy(PartPlan* p) { delete p; }

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
  const BandGeom geo_f(bw, 6 * bw + nb);
  if (6 * geo_f.W > kBandPref * 480 || geo_f.W > 160) return p;  // prefetch registers / panel rows of band_sweep
  if (6 * bw + 6 * bw + npiv > 32 * 8) return p;   // backward-substitution lanes
  std::vector<int> hb((size_t)P + 1);
  for (int j = 0; j <= P; ++j) hb[j] = (int)((long long)F * j / P);
  const int B = 6 * bw + nb, NBAND = 6 * (bw + 1), W = NBAND + B, LDW = W | 1;
  p->smem_fwd = geo_f.doubles() * sizeof(double);
  (void)LDW;
  (void)W;
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
  d2.band = bw2;
  d2.n = n2;
  p->smem_red = banded_smem_bytes(d2, bw2, nb);
  if (p->smem_red == 0) return p;
  const int M = 12 * bw + nb;
  p->bounds.alloc((size_t)P + 1);  // alloc_stream() == s (Solver::create_finish)
  p->row_c02.alloc((size_t)n2_aug);
  p->row_off2.alloc((size_t)n2_aug + 1);
  p->Lsep.alloc((size_t)F * 6 * bw * 6);
  p->X.alloc((size_t)P * M * M);
  p->dinv.alloc((size_t)6 * F);
  p->S2.alloc((size_t)p->s2_len);
  p->y2.alloc((size_t)n2);
  p->bounds.upload(hb.data(), (size_t)P + 1, s);
  p->row_c02.upload(c0.data(), (size_t)n2_aug, s);
  p->row_off2.upload(off.data(), (size_t)n2_aug + 1, s);
  p->Lsep.zero(s);
  p->X.zero(s);
  cudaStreamSynchronize(s);  // the host vectors go out of scope
  d2.S = p->S2.p;
  d2.row_off = reinterpret_cast<const int64_t*>(p->row_off2.p);
  d2.row_c0 = p->row_c02.p;
  d2.y = p->y2.p;
  p->d2 = d2;
  p->pd = PartDev{P, bw, F, npiv, p->bounds.p, p->Lsep.p, p->X.p, p->dinv.p, M};
  cudaFuncSetAttribute(k_part_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_fwd);
  cudaFuncSetAttribute(k_part_backward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_bwd);
  p->active = true;
  return p;
}

bool part_plan_active(const PartPlan* p) { return p && p->active; }

// the three phases on stream s; returns the number of kernels launched
int launch_part_solve(const Dev& d, PartPlan* p, cudaStream_t s) {
  k_part_forward<<<p->pd.P, kBandThreads, p->smem_fwd, s>>>(d, p->pd);
  Dev d2 = p->d2;
  d2.st = d.st;
  k_part_assemble<<<std::min(148, d2.n + 1), 64, 0, s>>>(d, p->pd, d2, p->s2_len);
  launch_chol_banded(d2, p->bw2, p->nb, p->smem_red, s);
  k_part_backward<<<p->pd.P, 256, p->smem_bwd, s>>>(d, p->pd, p->y2.p);
  return 4;
}

}  // namespace lfba
