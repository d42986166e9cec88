// lfba_setup.cuh — one-time indexing of a rank's observations on the device.
//
// Replaces the O(N) host loop of src/CameraCalibration.cpp:859-914 that allocates three heap objects per
// observation (functor, AutoDiffCostFunction, CauchyLoss) and Ceres' Program/ordering preprocessing:
//   sort by (point, frame)  ->  tracks, point->track CSR, frame->track CSR, co-visible frame-pair CSR,
//   distinct micro-lens table, length-sorted evaluation order.
// CUB device primitives (radix sort / scan / run-length encode) are used here and only here: this is
// set-up plumbing executed once per solve, not the per-iteration hot path.
#pragma once
#include <cub/cub.cuh>
#include <cuda_runtime.h>

#include <cstdint>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "lfba_device.cuh"

namespace lfba {

struct CudaError : std::runtime_error {
  int code;
  CudaError(const std::string& m, int c) : std::runtime_error(m), code(c) {}
};
#define LFBA_CUDA(call)                                                                              \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess)                                                                           \
      throw CudaError(std::string(#call) + ": " + cudaGetErrorString(e_) + " at " + __FILE__ + ":" + \
                          std::to_string(__LINE__),                                                  \
                      e_ == cudaErrorMemoryAllocation ? LFBA_OUT_OF_MEMORY : LFBA_CUDA_ERROR);       \
  } while (0)

// Stream-ordered allocations from the device's default memory pool (cudaMallocAsync): with the pool's release
// threshold raised (Solver::create) the memory of a finished solve is reused by the next one, so repeated
// drop-in calls do not pay cudaMalloc/cudaFree of several GB each time. Every buffer of a solver is allocated,
// used and freed on that solver's one stream.
inline cudaStream_t& alloc_stream() {
  static thread_local cudaStream_t s = nullptr;
  return s;
}

// Large blocks (>= 1 MiB) are additionally cached by this library, per device and by size: a drop-in caller solves the
// same problem shape again and again, and even with the pool's memory retained cudaMallocAsync was measured at 12-90 ms
// per solve for the multi-GB blocks of the 1M x 1000 scene once the pool is fragmented (it remaps physical pages to make
// the virtual range contiguous); an exact-size hit here costs a mutex and, when the block was last used on another
// stream, one cudaStreamWaitEvent. lfba_trim_cache() (or an allocation failure) returns everything to the pool.
constexpr size_t kCacheMaxBytes = 32ull << 30;
struct BlockCache {
  struct Block {
    void* p;
    size_t bytes;
    int device;
    cudaStream_t stream;
    cudaEvent_t ev;
  };
  std::mutex mu;
  std::vector<Block> blocks;
  static BlockCache& get() {
    static BlockCache c;
    return c;
  }
  void* take(size_t bytes, int device, cudaStream_t s, size_t* got) {
    std::lock_guard<std::mutex> lock(mu);
    int best = -1;
    for (int i = 0; i < (int)blocks.size(); ++i) {
      const Block& b = blocks[i];
      if (b.device != device || b.bytes < bytes || b.bytes > bytes + bytes / 8 + (1u << 20)) continue;
      if (best < 0 || b.bytes < blocks[best].bytes) best = i;
    }
    if (best < 0) return nullptr;
    Block b = blocks[best];
    blocks.erase(blocks.begin() + best);
    if (b.stream != s) cudaStreamWaitEvent(s, b.ev, 0);
    cudaEventDestroy(b.ev);
    *got = b.bytes;
    return b.p;
  }
  void give(void* p, size_t bytes, int device, cudaStream_t s) {
    Block b{p, bytes, device, s, nullptr};
    if (cudaEventCreateWithFlags(&b.ev, cudaEventDisableTiming) != cudaSuccess || cudaEventRecord(b.ev, s) != cudaSuccess) {
      if (b.ev) cudaEventDestroy(b.ev);
      cudaFreeAsync(p, s);
      return;
    }
    std::lock_guard<std::mutex> lock(mu);
    blocks.push_back(b);
    // bounded: beyond 32 GB per process the oldest blocks go back to the pool (ordered behind their last use)
    size_t total = 0;
    for (const Block& q : blocks) total += q.bytes;
    while (total > kCacheMaxBytes && blocks.size() > 1) {
      Block o = blocks.front();
      blocks.erase(blocks.begin());
      total -= o.bytes;
      int cur = 0;
      cudaGetDevice(&cur);
      if (o.device == cur) {
        cudaStreamWaitEvent(s, o.ev, 0);
        cudaFreeAsync(o.p, s);
      } else {
        cudaSetDevice(o.device);
        cudaEventSynchronize(o.ev);
        cudaFree(o.p);
        cudaSetDevice(cur);
      }
      cudaEventDestroy(o.ev);
    }
  }
  void trim() {  // back to the pool (stream-ordered behind each block's last use)
    std::lock_guard<std::mutex> lock(mu);
    int cur = 0;
    cudaGetDevice(&cur);
    for (Block& b : blocks) {
      cudaSetDevice(b.device);
      cudaEventSynchronize(b.ev);
      cudaEventDestroy(b.ev);
      cudaFree(b.p);
    }
    blocks.clear();
    cudaSetDevice(cur);
  }
};
constexpr size_t kCacheMinBytes = 1u << 20;

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  size_t block_bytes = 0;  // actual size of the block behind p (>= n * sizeof(T) when it came from the cache)
  int device = 0;
  cudaStream_t st = nullptr;
  DevBuf() {}
  explicit DevBuf(size_t count) { alloc(count); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), block_bytes(o.block_bytes), device(o.device), st(o.st) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; block_bytes = o.block_bytes; device = o.device; st = o.st; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void alloc(size_t count) {
    release();
    n = count;
    st = alloc_stream();
    if (!count) return;
    const size_t bytes = count * sizeof(T);
    cudaGetDevice(&device);
    if (bytes >= kCacheMinBytes) {
      BlockCache& c = BlockCache::get();
      size_t got = 0;
      if (void* q = c.take(bytes, device, st, &got)) {
        p = static_cast<T*>(q);
        block_bytes = got;
        return;
      }
    }
    cudaError_t e = cudaMallocAsync(&p, bytes, st);
    if (e == cudaErrorMemoryAllocation) {  // give the cached blocks back and try once more
      cudaGetLastError();
      BlockCache::get().trim();
      e = cudaMallocAsync(&p, bytes, st);
    }
    if (e != cudaSuccess) {
      p = nullptr;
      n = 0;
      LFBA_CUDA(e);
    }
    block_bytes = bytes;
  }
  void release() {
    if (p) {
      if (block_bytes >= kCacheMinBytes) BlockCache::get().give(p, block_bytes, device, st);
      else cudaFreeAsync(p, st);
    }
    p = nullptr;
    n = 0;
    block_bytes = 0;
  }
  void zero(cudaStream_t s) { if (n) LFBA_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
  void upload(const T* h, size_t count, cudaStream_t s) {
    if (count) LFBA_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void download(T* h, size_t count, cudaStream_t s) const {
    if (count) LFBA_CUDA(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
  }
};

// Device-resident index of one rank's observations.
struct ProblemIndex {
  int64_t N = 0;
  int T = 0, P = 0, F = 0, NL = 0, npairs = 0, bandwidth = 0;
  int64_t n_pair_items = 0;
  DevBuf<double2> obs;        // sorted by (point, frame)   (empty when the input already was: see obs_sorted)
  DevBuf<int32_t> lens_id;    // sorted order
  const double2* obs_sorted = nullptr;    // -> obs or obs_in
  const int32_t* lens_id_sorted = nullptr;  // -> lens_id or lens_id_in
  bool presorted = false;
  int64_t h2d_bytes = 0;  // observation arrays uploaded (40 B per observation)
  double h2d_ms = 0;      // device time of that upload on its copy stream (CUDA events)
  DevBuf<int32_t> perm;       // sorted position -> input position
  DevBuf<int32_t> trk_point, trk_frame, trk_begin, pt_trk_begin, frm_begin, frm_trk;
  DevBuf<int32_t> pair_begin, pair_f1, pair_f2, pair_t1, pair_t2;
  DevBuf<int32_t> eval_order;
  DevBuf<int2> eval_pf;       // (point, frame) of track eval_order[pos]
  // packed evaluation stream (build_stream): observations re-laid out in the order the fused evaluation kernel consumes
  // them. A ROUND is 32 / L length-adjacent tracks (one per L-lane group of a warp); its rows are 32 entries each:
  // entry (row, lane) = observation  lane % L + L * step  of the track of group lane / L, or padding (lens id -1).
  DevBuf<double2> s_obs;      // [n_rows * 32]
  DevBuf<int32_t> s_lid;      // [n_rows * 32]
  DevBuf<int32_t> step_base;  // [n_rounds + 1] first row of each round
  int n_rounds = 0, n_rows = 0, stream_L = 0;
  DevBuf<double> lens_xy;
  // input-order copies for the eval-only API
  DevBuf<double2> obs_in;
  DevBuf<int32_t> lens_id_in, point_in, frame_in;
  std::vector<int32_t> h_frame_count;  // tracks per frame (host copy)
  std::vector<int32_t> h_pair_f1, h_pair_f2;
};

void build_index(const lfba_problem& pb, ProblemIndex& ix, cudaStream_t s, int64_t* launches);
void build_stream(ProblemIndex& ix, int L, cudaStream_t s, int64_t* launches);

}  // namespace lfba
