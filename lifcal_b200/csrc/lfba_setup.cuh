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

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaStream_t st = nullptr;
  DevBuf() {}
  explicit DevBuf(size_t count) { alloc(count); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), st(o.st) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; st = o.st; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void alloc(size_t count) {
    release();
    n = count;
    st = alloc_stream();
    if (count) LFBA_CUDA(cudaMallocAsync(&p, count * sizeof(T), st));
  }
  void release() {
    if (p) cudaFreeAsync(p, st);
    p = nullptr;
    n = 0;
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
  DevBuf<int32_t> perm;       // sorted position -> input position
  DevBuf<int32_t> trk_point, trk_frame, trk_begin, pt_trk_begin, frm_begin, frm_trk;
  DevBuf<int32_t> pair_begin, pair_f1, pair_f2, pair_t1, pair_t2;
  DevBuf<int32_t> eval_order;
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
