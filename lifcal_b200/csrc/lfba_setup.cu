// lfba_setup.cu — device-side indexing of the observations (see lfba_setup.cuh).
#include "lfba_setup.cuh"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace lfba {

namespace {

struct Temp {  // CUB scratch
  DevBuf<unsigned char> buf;
  void* p = nullptr;
  size_t bytes = 0;
  void reserve(size_t b) {
    if (b > bytes) {
      buf.alloc(b);
      p = buf.p;
      bytes = b;
    }
  }
};

template <class K, class V>
void sort_pairs(Temp& tmp, const K* kin, K* kout, const V* vin, V* vout, size_t n, cudaStream_t s, int end_bit,
                bool descending = false) {
  if (n == 0) return;
  size_t bytes = 0;
  if (descending) {
    cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, kin, kout, vin, vout, (int)n, 0, end_bit, s);
    tmp.reserve(bytes);
    LFBA_CUDA(cub::DeviceRadixSort::SortPairsDescending(tmp.p, bytes, kin, kout, vin, vout, (int)n, 0, end_bit, s));
  } else {
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, (int)n, 0, end_bit, s);
    tmp.reserve(bytes);
    LFBA_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, kin, kout, vin, vout, (int)n, 0, end_bit, s));
  }
}
template <class T>
void exclusive_sum(Temp& tmp, const T* in, T* out, size_t n, cudaStream_t s) {
  if (n == 0) return;
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)n, s);
  tmp.reserve(bytes);
  LFBA_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in, out, (int)n, s));
}
int bits_for(uint64_t v) {
  int b = 1;
  while (b < 64 && (v >> b) != 0) ++b;
  return b;
}

__global__ void k_make_keys(const int32_t* pt, const int32_t* fr, uint64_t* keys, int32_t* vals, int64_t n, int P, int F,
                            int* bad) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int p = pt[i], f = fr[i];
  if (p < 0 || p >= P || f < 0 || f >= F) {  // index validation happens here, not in an O(N) host loop
    *bad = 1;
    keys[i] = 0;
    vals[i] = (int32_t)i;
    return;
  }
  keys[i] = ((uint64_t)(uint32_t)p << 32) | (uint32_t)f;
  vals[i] = (int32_t)i;
}
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
__global__ void k_gather_obs(const int32_t* perm, const double* ox, const double* oy, const double* mx,
                             const double* my, double2* obs, uint64_t* ka, uint64_t* kb, uint64_t* kh, int32_t* idx, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t j = perm[i];
  obs[i] = make_double2(ox[j], oy[j]);
  const uint64_t a = (uint64_t)__double_as_longlong(my[j]), b = (uint64_t)__double_as_longlong(mx[j]);
  ka[i] = a;
  kb[i] = b;
  if (kh) kh[i] = mix64(a + 0x9e3779b97f4a7c15ull) ^ mix64(b * 0xd1342543de82ef95ull + 1);
  idx[i] = (int32_t)i;
}
__global__ void k_copy_in(const double* ox, const double* oy, double2* obs_in, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) obs_in[i] = make_double2(ox[i], oy[i]);
}
__global__ void k_head_flags(const uint64_t* keys, int32_t* flags, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}
__global__ void k_head_flags2(const uint64_t* ka, const uint64_t* kb, int32_t* flags, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) flags[i] = (i == 0 || ka[i] != ka[i - 1] || kb[i] != kb[i - 1]) ? 1 : 0;
}
// heads of runs in hash order: a new lens starts where the hash OR the actual centre differs from the predecessor.
// (A hash collision between two different centres can only split one lens into two table entries, never merge two.)
__global__ void k_head_flags_hash(const uint64_t* kh, const uint64_t* ka, const uint64_t* kb, const int32_t* idx,
                                  int32_t* flags, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i == 0) { flags[i] = 1; return; }
  const int32_t a = idx[i], b = idx[i - 1];
  flags[i] = (kh[i] != kh[i - 1] || ka[a] != ka[b] || kb[a] != kb[b]) ? 1 : 0;
}
__global__ void k_fill_lens_hash(const uint64_t* ka, const uint64_t* kb, const int32_t* flags, const int32_t* lid,
                                 const int32_t* idx, int32_t* lens_id, double* lens_xy, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int l = lid[i] + flags[i] - 1;
  const int32_t src = idx[i];
  lens_id[src] = l;
  if (flags[i]) {
    lens_xy[2 * (size_t)l] = __longlong_as_double((long long)kb[src]);
    lens_xy[2 * (size_t)l + 1] = __longlong_as_double((long long)ka[src]);
  }
}
__global__ void k_fill_tracks(const uint64_t* keys, const int32_t* flags, const int32_t* tid, int32_t* trk_begin,
                              int32_t* trk_point, int32_t* trk_frame, int32_t* pt_count, int32_t* fr_count,
                              int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n || !flags[i]) return;
  const int t = tid[i];
  const int p = (int)(keys[i] >> 32), f = (int)(keys[i] & 0xffffffffu);
  trk_begin[t] = (int32_t)i;
  trk_point[t] = p;
  trk_frame[t] = f;
  atomicAdd(pt_count + p, 1);
  atomicAdd(fr_count + f, 1);
}
__global__ void k_iota(int32_t* a, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = i;
}
// Sort key of the evaluation order: track length (so that the tracks of a warp round have the same number of steps),
// then the image cell of the track's first micro lens (64 x 64 px cells, row-major): tracks that are neighbours in the
// order gather largely the SAME lens-table entries, which the L1-allocating cp.async of k_eval_rows turns into hits.
__global__ void k_eval_pf(const int32_t* __restrict__ eval_order, const int32_t* __restrict__ trk_point,
                          const int32_t* __restrict__ trk_frame, int2* __restrict__ out, int T) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T) return;
  const int t = eval_order[i];
  out[i] = make_int2(trk_point[t], trk_frame[t]);
}
__global__ void k_track_len(const int32_t* trk_begin, const int32_t* __restrict__ lens_id, const double* __restrict__ lens_xy,
                            int32_t* len, int T) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const int b = trk_begin[t], n = trk_begin[t + 1] - b;
  int cell = 0;
  if (n > 0 && lens_id && lens_xy) {
    const int l = lens_id[b];
    const int cx = min(63, max(0, (int)(lens_xy[2 * l] * (1.0 / 64.0))));
    const int cy = min(63, max(0, (int)(lens_xy[2 * l + 1] * (1.0 / 64.0))));
    cell = cy * 64 + ((cy & 1) ? 63 - cx : cx);  // boustrophedon: the end of a cell row is next to the start of the next
  }
  len[t] = (min(n, (1 << 19) - 1) << 12) | cell;
}
__global__ void k_pair_counts(const int32_t* pt_trk_begin, int32_t* cnt, int P) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int w = pt_trk_begin[p + 1] - pt_trk_begin[p];
  cnt[p] = w * (w - 1) / 2;
}
__global__ void k_fill_pairs(const int32_t* pt_trk_begin, const int32_t* trk_frame, const int32_t* off,
                             uint64_t* keys, uint64_t* vals, int* bw, int P) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int tb = pt_trk_begin[p], te = pt_trk_begin[p + 1];
  int o = off[p];
  int mybw = 0;
  for (int t1 = tb + 1; t1 < te; ++t1)
    for (int t2 = tb; t2 < t1; ++t2) {  // tracks of a point are sorted by frame: f1 > f2
      const int f1 = trk_frame[t1], f2 = trk_frame[t2];
      keys[o] = ((uint64_t)(uint32_t)f1 << 32) | (uint32_t)f2;
      vals[o] = ((uint64_t)(uint32_t)t1 << 32) | (uint32_t)t2;
      mybw = max(mybw, f1 - f2);
      ++o;
    }
  if (mybw > 0) atomicMax(bw, mybw);
}
__global__ void k_unpack_pairs(const uint64_t* vals, int32_t* t1, int32_t* t2, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  t1[i] = (int32_t)(vals[i] >> 32);
  t2[i] = (int32_t)(vals[i] & 0xffffffffu);
}
__global__ void k_unpack_pair_keys(const uint64_t* keys, int32_t* f1, int32_t* f2, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  f1[i] = (int32_t)(keys[i] >> 32);
  f2[i] = (int32_t)(keys[i] & 0xffffffffu);
}
__global__ void k_fill_lens(const uint64_t* ka, const uint64_t* kb, const int32_t* flags, const int32_t* lid,
                            const int32_t* idx, int32_t* lens_id, double* lens_xy, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int l = lid[i] + flags[i] - 1;  // lid is the EXCLUSIVE scan of the head flags
  lens_id[idx[i]] = l;
  if (flags[i]) {
    lens_xy[2 * (size_t)l] = __longlong_as_double((long long)kb[i]);
    lens_xy[2 * (size_t)l + 1] = __longlong_as_double((long long)ka[i]);
  }
}
__global__ void k_gather_u64(const uint64_t* src, const int32_t* idx, uint64_t* dst, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}
__global__ void k_scatter_i32(const int32_t* perm, const int32_t* src, int32_t* dst, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) dst[perm[i]] = src[i];
}

// ---- distinct micro-lens centres through an exact open-addressing hash table (two-word keys, CAS per word) ----
constexpr uint64_t kEmpty = 0xffffffffffffffffull;
__global__ void k_lens_insert(const double* mx, const double* my, unsigned long long* keyA, unsigned long long* keyB,
                              uint32_t mask, int32_t* slot_of, int* overflow, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long a = (unsigned long long)__double_as_longlong(mx[i]);
  const unsigned long long b = (unsigned long long)__double_as_longlong(my[i]);
  uint32_t h = (uint32_t)(mix64(a + 0x9e3779b97f4a7c15ull) ^ mix64(b * 0xd1342543de82ef95ull + 1)) & mask;
  for (int probe = 0; probe < 4096; ++probe) {
    unsigned long long ua = keyA[h];
    if (ua == kEmpty) {
      const unsigned long long old = atomicCAS(keyA + h, kEmpty, a);
      ua = old == kEmpty ? a : old;
    }
    if (ua == a) {
      unsigned long long ub = keyB[h];
      if (ub == kEmpty) {
        const unsigned long long old = atomicCAS(keyB + h, kEmpty, b);
        ub = old == kEmpty ? b : old;
      }
      if (ub == b) {
        slot_of[i] = (int32_t)h;
        return;
      }
    }
    h = (h + 1) & mask;
  }
  slot_of[i] = 0;
  *overflow = 1;
}
__global__ void k_lens_mark(const unsigned long long* keyA, const unsigned long long* keyB, int32_t* occ, int cap) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h < cap) occ[h] = (keyA[h] != kEmpty && keyB[h] != kEmpty) ? 1 : 0;
}
__global__ void k_lens_table(const unsigned long long* keyA, const unsigned long long* keyB, const int32_t* occ,
                             const int32_t* id, double* lens_xy, int cap) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h < cap && occ[h]) {
    lens_xy[2 * (size_t)id[h]] = __longlong_as_double((long long)keyA[h]);
    lens_xy[2 * (size_t)id[h] + 1] = __longlong_as_double((long long)keyB[h]);
  }
}
__global__ void k_lens_assign(const int32_t* slot_of, const int32_t* id, int32_t* lens_id_in, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) lens_id_in[i] = id[slot_of[i]];
}
__global__ void k_check_sorted(const int32_t* pt, const int32_t* fr, int* unsorted, int* bad, int P, int F, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int p = pt[i], f = fr[i];
  if (p < 0 || p >= P || f < 0 || f >= F) {
    *bad = 1;
    return;
  }
  if (i > 0) {
    const int pp = pt[i - 1], pf = fr[i - 1];
    if (pp > p || (pp == p && pf > f)) *unsorted = 1;
  }
}
__global__ void k_keys_from_sorted_input(const int32_t* pt, const int32_t* fr, uint64_t* keys, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) keys[i] = ((uint64_t)(uint32_t)pt[i] << 32) | (uint32_t)fr[i];
}
__global__ void k_gather_sorted(const int32_t* perm, const double2* obs_in, const int32_t* lens_in, double2* obs,
                                int32_t* lens, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t j = perm[i];
  obs[i] = obs_in[j];
  lens[i] = lens_in[j];
}

inline unsigned grid_for(int64_t n, int b = 256) { return (unsigned)((n + b - 1) / b); }

}  // namespace

void build_index(const lfba_problem& pb, ProblemIndex& ix, cudaStream_t s, int64_t* launches) {
  const int64_t N = pb.n_obs;
  const int P = pb.n_points, F = pb.n_frames;
  if (N >= (int64_t)2147483647) throw CudaError("more than 2^31-1 observations per rank", LFBA_INVALID_ARGUMENT);
  ix.N = N;
  ix.P = P;
  ix.F = F;
  Temp tmp;
  int64_t nl = 0;
  bool pending_copy_in = false;
  const bool dbg = std::getenv("LFBA_DEBUG") != nullptr;
  double tp = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
  auto phase = [&](const char* name) {
    if (!dbg) return;
    cudaStreamSynchronize(s);
    const double t = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    std::fprintf(stderr, "[lfba dbg]   index %-26s %8.2f ms\n", name, 1e3 * (t - tp));
    tp = t;
  };

  // raw input (input order); the ox/oy/mx/my staging buffers live only until the indexed copies exist.
  // The six host arrays go up on a SEPARATE copy stream in the order the indexing consumes them — lens centres first,
  // indices next, observed points last — one cudaMemcpyAsync per array with an event behind each group, so that the lens
  // table, the sortedness check and the track tables are built while the remaining arrays are still crossing PCIe
  // (4.5 GB at the 1M x 1000 scene: the upload IS the end-to-end time, everything else hides behind it).
  auto hnow = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double th0 = hnow();
  DevBuf<double> ox(N), oy(N), mx(N), my(N);
  ix.point_in.alloc(N);
  ix.frame_in.alloc(N);
  ix.obs_in.alloc(N);
  ix.lens_id_in.alloc(N);
  ix.pt_trk_begin.alloc((size_t)P + 1);
  ix.frm_begin.alloc((size_t)F + 1);
  DevBuf<int32_t> pt_count((size_t)P + 1), fr_count((size_t)F + 1);
  pt_count.zero(s);
  fr_count.zero(s);
  struct CopyLane {  // copy stream + events, released on every exit path
    cudaStream_t cs = nullptr;
    cudaEvent_t ready = nullptr, e_ml = nullptr, e_idx = nullptr, e_obs = nullptr, t0 = nullptr, t1 = nullptr;
    ~CopyLane() {
      if (cs) cudaStreamSynchronize(cs);
      for (cudaEvent_t e : {ready, e_ml, e_idx, e_obs, t0, t1})
        if (e) cudaEventDestroy(e);
      if (cs) cudaStreamDestroy(cs);
    }
  } lane;
  const double th1 = hnow();
  LFBA_CUDA(cudaStreamCreateWithFlags(&lane.cs, cudaStreamNonBlocking));
  for (cudaEvent_t* e : {&lane.ready, &lane.e_ml, &lane.e_idx, &lane.e_obs})
    LFBA_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  LFBA_CUDA(cudaEventCreate(&lane.t0));
  LFBA_CUDA(cudaEventCreate(&lane.t1));
  LFBA_CUDA(cudaEventRecord(lane.ready, s));  // the stream-ordered allocations above exist from here on
  LFBA_CUDA(cudaStreamWaitEvent(lane.cs, lane.ready, 0));
  LFBA_CUDA(cudaEventRecord(lane.t0, lane.cs));
  const double th2 = hnow();
  mx.upload(pb.ml_x, N, lane.cs);
  my.upload(pb.ml_y, N, lane.cs);
  LFBA_CUDA(cudaEventRecord(lane.e_ml, lane.cs));
  ix.point_in.upload(pb.point_idx, N, lane.cs);
  ix.frame_in.upload(pb.frame_idx, N, lane.cs);
  LFBA_CUDA(cudaEventRecord(lane.e_idx, lane.cs));
  ox.upload(pb.obs_x, N, lane.cs);
  oy.upload(pb.obs_y, N, lane.cs);
  LFBA_CUDA(cudaEventRecord(lane.e_obs, lane.cs));
  LFBA_CUDA(cudaEventRecord(lane.t1, lane.cs));
  ix.h2d_bytes = 40 * N;
  if (dbg)
    std::fprintf(stderr, "[lfba dbg]   host: allocations %.2f ms, copy stream + events %.2f ms, six memcpy enqueues %.2f ms\n",
                 1e3 * (th1 - th0), 1e3 * (th2 - th1), 1e3 * (hnow() - th2));
  phase("alloc+H2D enqueue");
  if (N > 0) {
    // ---- distinct lenses: exact hash table on the (mx, my) bit patterns ----
    int cap = 1 << 16;
    while (cap < (1 << 24) && (int64_t)cap < 4 * std::min<int64_t>(N, (int64_t)1 << 22)) cap <<= 1;
    DevBuf<unsigned long long> keyA(cap), keyB(cap);
    DevBuf<int32_t> slot_of(N), occ((size_t)cap + 1), oid((size_t)cap + 1);
    DevBuf<int> flags2(3);
    flags2.zero(s);
    LFBA_CUDA(cudaMemsetAsync(keyA.p, 0xff, (size_t)cap * 8, s));
    LFBA_CUDA(cudaMemsetAsync(keyB.p, 0xff, (size_t)cap * 8, s));
    occ.zero(s);
    LFBA_CUDA(cudaStreamWaitEvent(s, lane.e_ml, 0));
    k_lens_insert<<<grid_for(N), 256, 0, s>>>(mx.p, my.p, keyA.p, keyB.p, (uint32_t)(cap - 1), slot_of.p, flags2.p + 2, N);
    k_lens_mark<<<grid_for(cap), 256, 0, s>>>(keyA.p, keyB.p, occ.p, cap);
    exclusive_sum(tmp, occ.p, oid.p, (size_t)cap + 1, s);
    // ---- index validation + "is the input already sorted by (point, frame)?" in one pass ----
    LFBA_CUDA(cudaStreamWaitEvent(s, lane.e_idx, 0));
    k_check_sorted<<<grid_for(N), 256, 0, s>>>(ix.point_in.p, ix.frame_in.p, flags2.p, flags2.p + 1, P, F, N);
    int h_flags[3] = {0, 0, 0};
    int32_t h_nl = 0;
    flags2.download(h_flags, 3, s);
    LFBA_CUDA(cudaMemcpyAsync(&h_nl, oid.p + cap, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    LFBA_CUDA(cudaStreamSynchronize(s));
    if (h_flags[1]) throw CudaError("observation index out of range", LFBA_INVALID_ARGUMENT);
    if (h_flags[2]) throw CudaError("more than ~4M distinct micro-lens centres: lens table overflow", LFBA_INVALID_ARGUMENT);
    ix.NL = h_nl;
    ix.lens_xy.alloc((size_t)2 * std::max(1, ix.NL));
    k_lens_table<<<grid_for(cap), 256, 0, s>>>(keyA.p, keyB.p, occ.p, oid.p, ix.lens_xy.p, cap);
    k_lens_assign<<<grid_for(N), 256, 0, s>>>(slot_of.p, oid.p, ix.lens_id_in.p, N);
    nl += 8;
    phase("lenses (hash) + sortedness");
    // ---- order by (point, frame) ----
    DevBuf<uint64_t> k1(N);
    bool copied_in = false;
    ix.presorted = h_flags[0] == 0;
    if (ix.presorted) {
      // the caller's order is already (point, frame)-major: no sort, no copy — the sorted view IS the input view
      k_keys_from_sorted_input<<<grid_for(N), 256, 0, s>>>(ix.point_in.p, ix.frame_in.p, k1.p, N);
      ix.obs_sorted = ix.obs_in.p;
      ix.lens_id_sorted = ix.lens_id_in.p;
      nl += 1;
    } else {
      // the gather below reads the observed points: they are the last arrays to arrive
      LFBA_CUDA(cudaStreamWaitEvent(s, lane.e_obs, 0));
      k_copy_in<<<grid_for(N), 256, 0, s>>>(ox.p, oy.p, ix.obs_in.p, N);
      copied_in = true;
      // radix sort is stable, so observations keep their input order inside a track
      DevBuf<uint64_t> k0(N);
      DevBuf<int32_t> v0(N);
      DevBuf<int> bad(1);
      bad.zero(s);
      ix.perm.alloc(N);
      ix.obs.alloc(N);
      ix.lens_id.alloc(N);
      k_make_keys<<<grid_for(N), 256, 0, s>>>(ix.point_in.p, ix.frame_in.p, k0.p, v0.p, N, P, F, bad.p);
      sort_pairs(tmp, k0.p, k1.p, v0.p, ix.perm.p, (size_t)N, s, 32 + bits_for((uint64_t)(P > 0 ? P - 1 : 0)));
      k_gather_sorted<<<grid_for(N), 256, 0, s>>>(ix.perm.p, ix.obs_in.p, ix.lens_id_in.p, ix.obs.p, ix.lens_id.p, N);
      ix.obs_sorted = ix.obs.p;
      ix.lens_id_sorted = ix.lens_id.p;
      nl += 4;
    }
    phase("order (point,frame)");
    // ---- tracks
    DevBuf<int32_t> flags(N), tid(N);
    k_head_flags<<<grid_for(N), 256, 0, s>>>(k1.p, flags.p, N);
    exclusive_sum(tmp, flags.p, tid.p, (size_t)N, s);
    int32_t last_tid = 0, last_flag = 0;
    LFBA_CUDA(cudaMemcpyAsync(&last_tid, tid.p + (N - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    LFBA_CUDA(cudaMemcpyAsync(&last_flag, flags.p + (N - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    LFBA_CUDA(cudaStreamSynchronize(s));
    const int T = last_tid + last_flag;
    ix.T = T;
    ix.trk_begin.alloc((size_t)T + 1);
    ix.trk_point.alloc(T);
    ix.trk_frame.alloc(T);
    k_fill_tracks<<<grid_for(N), 256, 0, s>>>(k1.p, flags.p, tid.p, ix.trk_begin.p, ix.trk_point.p, ix.trk_frame.p,
                                              pt_count.p, fr_count.p, N);
    const int32_t n32 = (int32_t)N;
    LFBA_CUDA(cudaMemcpyAsync(ix.trk_begin.p + T, &n32, sizeof(int32_t), cudaMemcpyHostToDevice, s));
    nl += 3;
    phase("tracks");
    // CSR / orders / pairs below do not touch the observed points; they are interleaved into obs_in at the very end
    // (see the end of this function), when their upload — the last one — has landed.
    pending_copy_in = !copied_in;
  } else {
    ix.T = 0;
    ix.NL = 0;
    ix.trk_begin.alloc(1);
    ix.trk_begin.zero(s);
    ix.lens_xy.alloc(2);
    ix.obs_sorted = ix.obs_in.p;
    ix.lens_id_sorted = ix.lens_id_in.p;
  }
  const int T = ix.T;
  // ---- CSR: point -> tracks (tracks are already grouped by point), frame -> tracks
  exclusive_sum(tmp, pt_count.p, ix.pt_trk_begin.p, (size_t)P + 1, s);
  exclusive_sum(tmp, fr_count.p, ix.frm_begin.p, (size_t)F + 1, s);
  ix.h_frame_count.assign((size_t)F + 1, 0);
  fr_count.download(ix.h_frame_count.data(), (size_t)F + 1, s);
  ix.frm_trk.alloc(T);
  ix.eval_order.alloc(T);
  ix.eval_pf.alloc(T);
  if (T > 0) {
    DevBuf<int32_t> iota(T), keys_out(T), len(T);
    k_iota<<<grid_for(T), 256, 0, s>>>(iota.p, T);
    sort_pairs(tmp, ix.trk_frame.p, keys_out.p, iota.p, ix.frm_trk.p, (size_t)T, s, bits_for((uint64_t)(F > 0 ? F - 1 : 0)));
    k_track_len<<<grid_for(T), 256, 0, s>>>(ix.trk_begin.p, ix.lens_id_sorted, ix.lens_xy.p, len.p, T);
    sort_pairs(tmp, len.p, keys_out.p, iota.p, ix.eval_order.p, (size_t)T, s, 32, true);
    k_eval_pf<<<grid_for(T), 256, 0, s>>>(ix.eval_order.p, ix.trk_point.p, ix.trk_frame.p, ix.eval_pf.p, T);
    nl += 7;
  }
  phase("csr+orders");
  // ---- co-visible frame pairs
  ix.npairs = 0;
  ix.bandwidth = 0;
  ix.n_pair_items = 0;
  if (T > 0 && P > 0) {
    DevBuf<int32_t> pcnt((size_t)P + 1), poff((size_t)P + 1);
    pcnt.zero(s);
    k_pair_counts<<<grid_for(P), 256, 0, s>>>(ix.pt_trk_begin.p, pcnt.p, P);
    exclusive_sum(tmp, pcnt.p, poff.p, (size_t)P + 1, s);
    int32_t total = 0;
    LFBA_CUDA(cudaMemcpyAsync(&total, poff.p + P, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    LFBA_CUDA(cudaStreamSynchronize(s));
    if (total < 0) throw CudaError("too many co-visible track pairs", LFBA_INVALID_ARGUMENT);
    ix.n_pair_items = total;
    if (total > 0) {
      DevBuf<uint64_t> pk(total), pv(total), pk2(total), pv2(total);
      DevBuf<int> bw(1);
      bw.zero(s);
      k_fill_pairs<<<grid_for(P), 256, 0, s>>>(ix.pt_trk_begin.p, ix.trk_frame.p, poff.p, pk.p, pv.p, bw.p, P);
      sort_pairs(tmp, pk.p, pk2.p, pv.p, pv2.p, (size_t)total, s, 32 + bits_for((uint64_t)(F > 0 ? F - 1 : 0)));
      ix.pair_t1.alloc(total);
      ix.pair_t2.alloc(total);
      k_unpack_pairs<<<grid_for(total), 256, 0, s>>>(pv2.p, ix.pair_t1.p, ix.pair_t2.p, total);
      // run-length encode the sorted keys -> distinct (f1, f2) and their item ranges
      DevBuf<uint64_t> uniq(total);
      DevBuf<int32_t> counts((size_t)total + 1), nruns(1);
      size_t bytes = 0;
      cub::DeviceRunLengthEncode::Encode(nullptr, bytes, pk2.p, uniq.p, counts.p, nruns.p, total, s);
      tmp.reserve(bytes);
      LFBA_CUDA(cub::DeviceRunLengthEncode::Encode(tmp.p, bytes, pk2.p, uniq.p, counts.p, nruns.p, total, s));
      int32_t nr = 0;
      LFBA_CUDA(cudaMemcpyAsync(&nr, nruns.p, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
      LFBA_CUDA(cudaMemcpyAsync(&ix.bandwidth, bw.p, sizeof(int), cudaMemcpyDeviceToHost, s));
      LFBA_CUDA(cudaStreamSynchronize(s));
      ix.npairs = nr;
      ix.pair_begin.alloc((size_t)nr + 1);
      ix.pair_f1.alloc(nr);
      ix.pair_f2.alloc(nr);
      LFBA_CUDA(cudaMemsetAsync(counts.p + nr, 0, sizeof(int32_t), s));
      exclusive_sum(tmp, counts.p, ix.pair_begin.p, (size_t)nr + 1, s);
      k_unpack_pair_keys<<<grid_for(nr), 256, 0, s>>>(uniq.p, ix.pair_f1.p, ix.pair_f2.p, nr);
      ix.h_pair_f1.resize(nr);
      ix.h_pair_f2.resize(nr);
      ix.pair_f1.download(ix.h_pair_f1.data(), nr, s);
      ix.pair_f2.download(ix.h_pair_f2.data(), nr, s);
      nl += 8;
    }
  }
  if (N > 0) {
    if (pending_copy_in) {
      LFBA_CUDA(cudaStreamWaitEvent(s, lane.e_obs, 0));
      k_copy_in<<<grid_for(N), 256, 0, s>>>(ox.p, oy.p, ix.obs_in.p, N);
      nl += 1;
    }
    // the staging buffers are freed stream-ordered on s: s must be behind the copy stream first
    LFBA_CUDA(cudaStreamWaitEvent(s, lane.e_obs, 0));
  }
  LFBA_CUDA(cudaStreamSynchronize(s));
  LFBA_CUDA(cudaStreamSynchronize(lane.cs));
  {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, lane.t0, lane.t1) == cudaSuccess) ix.h2d_ms = ms;
  }
  phase("pairs + last upload");
  if (launches) *launches += nl;
}

// ---- packed evaluation stream ----------------------------------------------------------------------------------
namespace {
__global__ void k_round_steps(const int32_t* eval_order, const int32_t* trk_begin, int32_t* steps, int G, int L, int R) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > R) return;
  if (r == R) {
    steps[r] = 0;
    return;
  }
  const int t = eval_order[(size_t)r * G];  // tracks are in descending length order: the first is the longest
  steps[r] = (trk_begin[t + 1] - trk_begin[t] + L - 1) / L;
}
// round of every row: thread r marks the rows [step_base[r], step_base[r + 1]) (a handful each)
__global__ void k_row_rounds(const int32_t* __restrict__ step_base, int32_t* __restrict__ row_round, int R) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  for (int row = step_base[r]; row < step_base[r + 1]; ++row) row_round[row] = r;
}
__global__ void k_fill_stream(const int32_t* __restrict__ eval_order, const int32_t* __restrict__ trk_begin,
                              const int32_t* __restrict__ step_base, const int32_t* __restrict__ row_round,
                              const double2* __restrict__ obs, const int32_t* __restrict__ lens_id,
                              double2* __restrict__ s_obs, int32_t* __restrict__ s_lid, int T, int G, int L,
                              int64_t n_entries) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= n_entries) return;
  const int row = (int)(e >> 5), lane = (int)(e & 31);
  const int r = row_round[row];
  const int m = row - step_base[r];
  const int pos = r * G + lane / L;
  double2 o = make_double2(0.0, 0.0);
  int lid = -1;
  if (pos < T) {
    const int t = eval_order[pos];
    const int i = trk_begin[t] + lane % L + L * m;
    if (i < trk_begin[t + 1]) {
      o = obs[i];
      lid = lens_id[i];
    }
  }
  s_obs[e] = o;
  s_lid[e] = lid;
}
}  // namespace

void build_stream(ProblemIndex& ix, int L, cudaStream_t s, int64_t* launches) {
  ix.stream_L = L;
  ix.n_rounds = 0;
  ix.n_rows = 0;
  if (ix.T <= 0) {
    ix.step_base.alloc(1);
    ix.step_base.zero(s);
    return;
  }
  const int G = 32 / L;
  const int R = (ix.T + G - 1) / G;
  Temp tmp;
  DevBuf<int32_t> steps((size_t)R + 1);
  ix.step_base.alloc((size_t)R + 1);
  k_round_steps<<<grid_for(R + 1), 256, 0, s>>>(ix.eval_order.p, ix.trk_begin.p, steps.p, G, L, R);
  exclusive_sum(tmp, steps.p, ix.step_base.p, (size_t)R + 1, s);
  int32_t rows = 0;
  LFBA_CUDA(cudaMemcpyAsync(&rows, ix.step_base.p + R, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  LFBA_CUDA(cudaStreamSynchronize(s));
  if (rows < 0) throw CudaError("evaluation stream too long", LFBA_INVALID_ARGUMENT);
  ix.n_rounds = R;
  ix.n_rows = rows;
  const int64_t n_entries = (int64_t)rows * 32;
  ix.s_obs.alloc((size_t)n_entries);
  ix.s_lid.alloc((size_t)n_entries);
  DevBuf<int32_t> row_round((size_t)std::max(1, rows));
  k_row_rounds<<<grid_for(R), 256, 0, s>>>(ix.step_base.p, row_round.p, R);
  k_fill_stream<<<grid_for(n_entries), 256, 0, s>>>(ix.eval_order.p, ix.trk_begin.p, ix.step_base.p, row_round.p,
                                                      ix.obs_sorted, ix.lens_id_sorted, ix.s_obs.p, ix.s_lid.p, ix.T, G, L,
                                                      n_entries);
  if (launches) *launches += 4;
}

}  // namespace lfba
