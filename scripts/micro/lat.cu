// Dependent-issue latencies that bound the reduced-system pivot chain (one warp, one lane active): DFMA, DMUL, rsqrt(double),
// 1/x, sqrt, shared-memory load, __syncthreads with 256 threads. nvcc -arch=sm_100a -O3 -o lat lat.cu && ./lat
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(double* out, long long* cyc, double seed) {
  __shared__ double sm[64];
  if (threadIdx.x < 64) sm[threadIdx.x] = seed + threadIdx.x;
  __syncthreads();
  double x = seed, y = 1.0000001;
  const int N = 2048;
  long long t0 = clock64();
  if (OP == 6) {
    for (int i = 0; i < N; ++i) __syncthreads();
  } else if (threadIdx.x == 0) {
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
      if (OP == 0) x = fma(x, y, 1e-9);
      if (OP == 1) x = x * y;
      if (OP == 2) x = rsqrt(x) + 1.5;
      if (OP == 3) x = 1.0 / x + 0.5;
      if (OP == 4) x = sqrt(x) + 2.0;
      if (OP == 5) x = sm[((int)x) & 63];
      if (OP == 7) x = (double)rsqrtf((float)x) + 1.5;
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[0] = x; cyc[0] = (t1 - t0); }
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
  const char* names[] = {"DFMA dependent", "DMUL dependent", "rsqrt(double)+add", "1/x+add", "sqrt+add", "LDS dependent (incl. F2I)", "__syncthreads (256 thr)", "rsqrtf via float + add"};
  for (int op = 0; op < 8; ++op) {
    for (int rep = 0; rep < 2; ++rep) {
      switch (op) {
        case 0: k<0><<<1, 256>>>(out, cyc, 1.0); break;
        case 1: k<1><<<1, 256>>>(out, cyc, 1.0); break;
        case 2: k<2><<<1, 256>>>(out, cyc, 1.0); break;
        case 3: k<3><<<1, 256>>>(out, cyc, 1.0); break;
        case 4: k<4><<<1, 256>>>(out, cyc, 1.0); break;
        case 5: k<5><<<1, 256>>>(out, cyc, 1.0); break;
        case 6: k<6><<<1, 256>>>(out, cyc, 1.0); break;
        case 7: k<7><<<1, 256>>>(out, cyc, 1.0); break;
      }
      cudaDeviceSynchronize();
    }
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s %8.1f cycles/op\n", names[op], c / 2048.0);
  }
  return 0;
}
