// micro-benchmark: CUB radix sort throughput on this GPU (setup path of lfba)
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdint>
__global__ void fill(uint64_t* k, int* v, int n, int P, int F) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t z = (uint64_t)i * 0x9e3779b97f4a7c15ull; z ^= z >> 29; z *= 0xbf58476d1ce4e5b9ull; z ^= z >> 32;
  k[i] = ((z % P) << 32) | ((z >> 40) % F);
  v[i] = i;
}
__global__ void to32(const uint64_t* k, uint32_t* k32, int n, int F) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) k32[i] = (uint32_t)((k[i] >> 32) * F + (k[i] & 0xffffffffu));
}
template <class K> float run(K* kin, K* kout, int* vin, int* vout, int n, int endbit, void* tmp, size_t bytes) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cub::DeviceRadixSort::SortPairs(tmp, bytes, kin, kout, vin, vout, n, 0, endbit);
  cudaEventRecord(a);
  cub::DeviceRadixSort::SortPairs(tmp, bytes, kin, kout, vin, vout, n, 0, endbit);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
  const int n = 112000000, P = 1000000, F = 1000;
  uint64_t *k, *ko; int *v, *vo; uint32_t *k32, *k32o;
  cudaMalloc(&k, 8ull * n); cudaMalloc(&ko, 8ull * n); cudaMalloc(&v, 4ull * n); cudaMalloc(&vo, 4ull * n);
  cudaMalloc(&k32, 4ull * n); cudaMalloc(&k32o, 4ull * n);
  fill<<<(n + 255) / 256, 256>>>(k, v, n, P, F);
  to32<<<(n + 255) / 256, 256>>>(k, k32, n, F);
  size_t bytes = 0, b2 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, k, ko, v, vo, n, 0, 64);
  cub::DeviceRadixSort::SortPairs(nullptr, b2, k32, k32o, v, vo, n, 0, 32);
  if (b2 > bytes) bytes = b2;
  void* tmp; cudaMalloc(&tmp, bytes);
  printf("temp bytes %zu\n", bytes);
  printf("64-bit keys, 52 bits: %.2f ms\n", run(k, ko, v, vo, n, 52, tmp, bytes));
  printf("64-bit keys, 64 bits: %.2f ms\n", run(k, ko, v, vo, n, 64, tmp, bytes));
  printf("32-bit keys, 30 bits: %.2f ms\n", run(k32, k32o, v, vo, n, 30, tmp, bytes));
  printf("32-bit keys, 32 bits: %.2f ms\n", run(k32, k32o, v, vo, n, 32, tmp, bytes));
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
