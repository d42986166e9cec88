// lfba_chol.cu — FP64 Cholesky of the reduced camera system on its skyline (envelope) storage.
//
// Replaces Ceres' DenseSchurComplementSolver::SolveReducedLinearSystem (Eigen LLT of the dense
// (17+6F)^2 matrix). The reduced system is ordered [poses | coupled points | camera | rhs],
// so with windowed visibility it is block-banded with a dense border (an arrowhead): Cholesky fills only
// inside the envelope, which is exactly what is stored. The right-hand side is kept as one extra matrix row:
// the factorisation leaves L^-1 g in it (forward substitution for free).
//
// Tiled left-looking factorisation, 64x64 tiles, one launch per tile column, one CTA per structurally
// non-zero tile of the column: A_ik -= sum_j L_ij L_kj^T (FP64 FMA, 4x4 register micro-tiles), POTRF of the
// diagonal tile (recomputed by every CTA of the column: no inter-CTA dependency inside a launch), TRSM.
// tcgen05 has no FP64 kind and the tiles are tiny, so this is plain DFMA work; the chain of n pivots is the
// latency that matters, not the flops.
#include "lfba_device.cuh"
#include "lfba_kernels.h"

namespace lfba {

constexpr int TS = kTile;      // 64
constexpr int LD = TS + 1;     // padded leading dimension in shared memory

__device__ __forceinline__ double sky_load(const Dev& d, int r, int c, int n_aug) {
  if (r >= n_aug || c > r) return 0.0;
  if (r == d.n && c == d.n) return 0.0;
  const int c0 = d.row_c0[r];
  if (c < c0) return 0.0;
  return d.S[d.row_off[r] + (c - c0)];
}
__device__ __forceinline__ void sky_store(const Dev& d, int r, int c, int n_aug, double v) {
  if (r >= n_aug || c > r) return;
  if (r == d.n && c == d.n) return;
  const int c0 = d.row_c0[r];
  if (c < c0) return;
  d.S[d.row_off[r] + (c - c0)] = v;
}

// load tile (ti, tj) into smem [TS][LD]
__device__ __forceinline__ void load_tile(const Dev& d, int ti, int tj, int n_aug, double* sm) {
  for (int e = threadIdx.x; e < TS * TS; e += blockDim.x) {
    const int r = e / TS, c = e % TS;
    sm[r * LD + c] = sky_load(d, ti * TS + r, tj * TS + c, n_aug);
  }
}

// One tile column k. grid.x = number of tile rows i >= k; blockIdx.x = i - k.
__global__ void __launch_bounds__(256) k_chol_column(Dev d, int k, const int* __restrict__ tile_first) {
  LmState* st = d.st;
  if (st->done || !st->solve_ok) return;
  const int i = k + blockIdx.x;
  const int n_aug = d.n + 1;
  if (tile_first[i] > k) return;  // structurally zero tile
  extern __shared__ double sm[];
  double* sA = sm;                 // L_ij  / later A_ik
  double* sB = sm + TS * LD;       // L_kj
  double* sD = sm + 2 * TS * LD;   // diagonal tile
  __shared__ int s_fail;
  if (threadIdx.x == 0) s_fail = 0;

  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;  // 16x16 threads, 4x4 micro-tile each
  double acc[16], accd[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) acc[e] = accd[e] = 0.0;

  const int jk = tile_first[k];
  const int ji = tile_first[i] > jk ? tile_first[i] : jk;
  for (int j = jk; j < k; ++j) {
    __syncthreads();
    load_tile(d, k, j, n_aug, sB);
    const bool both = (j >= ji) && (i != k);
    if (both) load_tile(d, i, j, n_aug, sA);
    __syncthreads();
#pragma unroll 4
    for (int t = 0; t < TS; ++t) {
      double b[4], bk[4], a[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        b[q] = sB[(tx * 4 + q) * LD + t];   // column index of the output (row of L_kj)
        bk[q] = sB[(ty * 4 + q) * LD + t];  // row index of the diagonal output
      }
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) accd[p * 4 + q] += bk[p] * b[q];
      if (both) {
#pragma unroll
        for (int p = 0; p < 4; ++p) a[p] = sA[(ty * 4 + p) * LD + t];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[p * 4 + q] += a[p] * b[q];
      }
    }
  }
  __syncthreads();
  // D = A_kk - accd ; A = A_ik - acc
  load_tile(d, k, k, n_aug, sD);
  if (i != k) load_tile(d, i, k, n_aug, sA);
  __syncthreads();
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = ty * 4 + p, c = tx * 4 + q;
      sD[r * LD + c] -= accd[p * 4 + q];
      if (i != k) sA[r * LD + c] -= acc[p * 4 + q];
    }
  __syncthreads();

  // POTRF of the diagonal tile (lower). Columns >= n (the rhs row and padding) are not pivots.
  const int nreal = min(TS, d.n - k * TS);
  for (int c = 0; c < nreal; ++c) {
    __syncthreads();  // trailing update of the previous column is complete
    const double piv = sD[c * LD + c];
    if (!(piv > 0.0)) {
      if (threadIdx.x == 0) s_fail = 1;
    }
    const double dinv = piv > 0.0 ? rsqrt(piv) : 0.0;
    __syncthreads();
    if (threadIdx.x < TS) {
      const int r = threadIdx.x;
      if (r == c) sD[c * LD + c] = piv > 0.0 ? sqrt(piv) : 1.0;
      else if (r > c) sD[r * LD + c] *= dinv;
    }
    __syncthreads();
    // trailing update of the lower triangle: (r, c2), c < c2 <= r
    const int m = TS - 1 - c;  // rows c+1..TS-1
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
      const int r = c + 1 + e / m, c2 = c + 1 + e % m;
      if (c2 <= r) sD[r * LD + c2] -= sD[r * LD + c] * sD[c2 * LD + c];
    }
  }
  __syncthreads();
  if (i == k) {
    for (int e = threadIdx.x; e < TS * TS; e += blockDim.x) {
      const int r = e / TS, c = e % TS;
      if (c <= r && c < nreal) sky_store(d, k * TS + r, k * TS + c, n_aug, sD[r * LD + c]);
    }
    if (threadIdx.x == 0 && s_fail) st->solve_ok = 0;
    return;
  }
  // TRSM: X L_kk^T = A  (column by column)
  for (int c = 0; c < nreal; ++c) {
    const double dinv = 1.0 / sD[c * LD + c];
    if (threadIdx.x < TS) sA[threadIdx.x * LD + c] *= dinv;
    __syncthreads();
    const int m = TS - 1 - c;
    for (int e = threadIdx.x; e < TS * m; e += blockDim.x) {
      const int r = e / m, c2 = c + 1 + e % m;
      sA[r * LD + c2] -= sA[r * LD + c] * sD[c2 * LD + c];
    }
    __syncthreads();
  }
  for (int e = threadIdx.x; e < TS * TS; e += blockDim.x) {
    const int r = e / TS, c = e % TS;
    if (c < nreal) sky_store(d, i * TS + r, k * TS + c, n_aug, sA[r * LD + c]);
  }
}

// Backward substitution L^T y = z with z = the factorised rhs row. One CTA; right-looking over the rows of L
// so that every access is a contiguous skyline row.
__global__ void __launch_bounds__(256) k_backsolve(Dev d) {
  LmState* st = d.st;
  if (st->done || !st->solve_ok) return;
  const int n = d.n;
  double* y = d.y;
  const double* zrow = d.S + d.row_off[n];
  for (int j = threadIdx.x; j < n; j += blockDim.x) y[j] = zrow[j];
  __shared__ double yc;
  __syncthreads();
  for (int c = n - 1; c >= 0; --c) {
    const int c0 = d.row_c0[c];
    const double* Lr = d.S + d.row_off[c];
    if (threadIdx.x == 0) {
      const double v = y[c] / Lr[c - c0];
      y[c] = v;
      yc = v;
    }
    __syncthreads();
    const double v = yc;
    for (int j = c0 + threadIdx.x; j < c; j += blockDim.x) y[j] -= Lr[j - c0] * v;
    __syncthreads();
  }
}

// per-device opt-in to > 48 KB of dynamic shared memory (call once per device with that device current)
void prepare_device_kernels() {
  cudaFuncSetAttribute(k_chol_column, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(3 * TS * LD * sizeof(double)));
}

int launch_reduced_solve(const Dev& d, int n_tiles, const int* d_tile_first, cudaStream_t s) {
  const size_t smem = 3 * TS * LD * sizeof(double);
  for (int k = 0; k < n_tiles; ++k) k_chol_column<<<n_tiles - k, 256, smem, s>>>(d, k, d_tile_first);
  k_backsolve<<<1, 256, 0, s>>>(d);
  return n_tiles + 1;
}

// ------------------------------------------------------------------------------------------------
// FP64 FMA throughput micro-benchmark (the roofline denominator for the FP64-pipe-bound fused kernel;
// MEASURED_PEAKS.json has no FP64 figure).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters) {
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
  const double b = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 123.456) out[0] = s;
}

double measure_fp64_tflops(cudaStream_t s) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  double* out = nullptr;
  cudaMalloc(&out, 8);
  const int iters = 4096, grid = sms * 8, block = 256;
  k_dfma_peak<<<grid, block, 0, s>>>(out, 64);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0, s);
    k_dfma_peak<<<grid, block, 0, s>>>(out, iters);
    cudaEventRecord(e1, s);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 64.0 * (double)iters * (double)grid * (double)block;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  return best;
}

}  // namespace lfba
