// Micro-benchmark: FP64 FMA throughput of one B200 as a function of resident warps per SM sub-partition and of the
// number of independent DFMA chains per thread (ILP). Answers: can 2 warps per scheduler (255-register kernels) keep
// the FP64 pipe busy, and how many independent chains does that take?
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, int iters) {
  double a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
  const double b = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 64 / ILP; ++u)
#pragma unroll
      for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], b, c);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i];
  if (s == 123.456) out[0] = s;
}
template <int ILP>
double run(int sms, int warps_per_sm, double* out) {
  const int iters = 2048, block = 32 * warps_per_sm >= 128 ? 128 : 32 * warps_per_sm;
  const int blocks_per_sm = (32 * warps_per_sm) / block;
  const int grid = sms * blocks_per_sm;
  k<ILP><<<grid, block>>>(out, 16);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0); k<ILP><<<grid, block>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double tf = 2.0 * 64.0 * iters * (double)grid * block / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  return best;
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out; cudaMalloc(&out, 8);
  printf("warps/SM  ILP1   ILP2   ILP4   ILP8   ILP16  ILP32 (TFLOP/s)\n");
  for (int w : {4, 8, 12, 16, 32, 64}) {
    printf("%7d  %6.2f %6.2f %6.2f %6.2f %6.2f %6.2f\n", w, run<1>(sms, w, out), run<2>(sms, w, out), run<4>(sms, w, out),
           run<8>(sms, w, out), run<16>(sms, w, out), run<32>(sms, w, out));
  }
  return 0;
}
