// Which host-side call makes cudaMallocAsync slow again on the next solve? (default pool, release threshold = max)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  cudaSetDevice(0);
  cudaMemPool_t pool; cudaDeviceGetDefaultMemPool(&pool, 0);
  unsigned long long thr = ~0ull; cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  const size_t sizes[] = {900u << 20, 900u << 20, 900u << 20, 900u << 20, 450u << 20, 450u << 20, 1800u << 20, 450u << 20, 1150u << 20, 1150u << 20, 1150u << 20};
  char* host = nullptr; cudaMallocHost(&host, 900u << 20);
  for (int it = 0; it < 5; ++it) {
    cudaStream_t s, cs; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking);
    cudaEvent_t e; cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    std::vector<void*> p;
    double t0 = now();
    for (size_t sz : sizes) { void* q = nullptr; cudaMallocAsync(&q, sz, s); p.push_back(q); }
    double t1 = now();
    if (mode & 1) {  // cross-stream use: copy stream writes into buffers allocated on s
      cudaEventRecord(e, s); cudaStreamWaitEvent(cs, e, 0);
      cudaMemcpyAsync(p[0], host, 900u << 20, cudaMemcpyHostToDevice, cs);
      cudaEventRecord(e, cs); cudaStreamWaitEvent(s, e, 0);
    }
    void* small[8] = {};
    if (mode & 2) for (auto& q : small) cudaMalloc(&q, 1 << 20);
    cudaStreamSynchronize(s);
    double t2 = now();
    for (void* q : p) cudaFreeAsync(q, s);
    cudaStreamSynchronize(s);
    if (mode & 2) for (auto& q : small) cudaFree(q);
    if (mode & 4) { int* h; cudaMallocHost(&h, 16); cudaFreeHost(h); }
    if (mode & 8) cudaDeviceSynchronize();
    cudaStreamSynchronize(cs);
    cudaEventDestroy(e); cudaStreamDestroy(s); cudaStreamDestroy(cs);
    double t3 = now();
    unsigned long long res = 0; cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &res);
    printf("mode %d iter %d: malloc calls %.2f ms, use+sync %.2f ms, free etc %.2f ms, pool reserved %.2f GB\n", mode, it, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), res / 1e9);
  }
  return 0;
}
