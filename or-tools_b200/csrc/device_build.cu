// device_build.cu -- builds the two SELL-32 images of K on the GPU from the
// caller's CSC arrays: the device counterpart of the ShardedQuadraticProgram
// constructor (sharded_quadratic_program.cc:79-107: keep K, build the explicit
// transpose, set up the sharders). The host only uploads the raw CSC arrays and
// reads back a handful of sizes.
//
// Everything is deterministic (no order-dependent atomics): the row-major copy
// comes from a stable LSD radix sort of the entries by row (warp-match ranking),
// so the entries of a row stay in ascending column order exactly like Eigen's
// transpose assignment and like the host builder (sell_builder.cc), which
// produces the same layout and is kept for A/B checks (PDLP_B200_HOST_BUILD=1).
//
// Layout rules (see sell_builder.cc / DESIGN.md section 3): rows longer than
// split_len first, cut into virtual slots; the rest stably sorted by descending
// length inside windows of `sigma` rows; slices of 32 slots, slot-major.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <chrono>
#include <vector>

#include "device_ops.h"

namespace pdlp_b200 {

#define CUDA_OK(expr)                                                                                   \
  do {                                                                                                  \
    cudaError_t e__ = (expr);                                                                           \
    if (e__ != cudaSuccess) {                                                                           \
      char buf__[512];                                                                                  \
      std::snprintf(buf__, sizeof(buf__), "CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), __FILE__, \
                    __LINE__, cudaGetErrorString(e__));                                                 \
      throw std::runtime_error(buf__);                                                                  \
    }                                                                                                   \
  } while (0)

namespace build_kernels {

constexpr int kT = 256;          // generic block size
constexpr int kScanT = 1024;     // scan block: 1024 threads x 4 items
constexpr int kScanItems = 4;
constexpr int kSortT = 1024;     // radix-sort block
constexpr int kSortTiles = 8;    // sub-tiles of 1024 items per block
constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;

__device__ __forceinline__ int64_t gtid() { return static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; }

// ---- exclusive scan of int64 ------------------------------------------------
// Block-level inclusive scan of one value per thread (1024 threads).
__device__ __forceinline__ int64_t block_inclusive_scan(int64_t v, int64_t* total) {
  __shared__ int64_t warp_sums[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int64_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  if (lane == 31) warp_sums[warp] = v;
  __syncthreads();
  if (warp == 0) {
    int64_t w = lane < (blockDim.x >> 5) ? warp_sums[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    warp_sums[lane] = w;
  }
  __syncthreads();
  const int64_t offset = warp > 0 ? warp_sums[warp - 1] : 0;
  *total = warp_sums[(blockDim.x >> 5) - 1];
  __syncthreads();
  return v + offset;
}

// out[i] = sum of in[0..i) inside each block of 4096 items; block totals to sums.
__global__ void __launch_bounds__(kScanT) k_scan_blocks(const int64_t* __restrict__ in, int64_t* __restrict__ out, int64_t n, int64_t* __restrict__ sums) {
  const int64_t base = (static_cast<int64_t>(blockIdx.x) * kScanT + threadIdx.x) * kScanItems;
  int64_t v[kScanItems];
  int64_t local = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    v[k] = base + k < n ? in[base + k] : 0;
    local += v[k];
  }
  int64_t total;
  const int64_t incl = block_inclusive_scan(local, &total);
  int64_t run = incl - local;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (base + k < n) out[base + k] = run;
    run += v[k];
  }
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}
// Exclusive scan of the block totals in place (one block); grand total to *total_out.
__global__ void __launch_bounds__(kScanT) k_scan_sums(int64_t* sums, int64_t count, int64_t* total_out) {
  int64_t carry = 0;
  for (int64_t base = 0; base < count; base += kScanT) {
    const int64_t i = base + threadIdx.x;
    const int64_t v = i < count ? sums[i] : 0;
    int64_t total;
    const int64_t incl = block_inclusive_scan(v, &total);
    if (i < count) sums[i] = carry + incl - v;
    carry += total;
  }
  if (threadIdx.x == 0) *total_out = carry;
}
__global__ void __launch_bounds__(kScanT) k_scan_add(int64_t* __restrict__ out, int64_t n, const int64_t* __restrict__ sums) {
  const int64_t base = (static_cast<int64_t>(blockIdx.x) * kScanT + threadIdx.x) * kScanItems;
  const int64_t add = sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kScanItems; ++k)
    if (base + k < n) out[base + k] += add;
}

// ---- CSC pass -----------------------------------------------------------------
// col_len[c] = entries of column c inside the row block; flags bad input.
__global__ void __launch_bounds__(kT) k_count_columns(int64_t n, const int64_t* __restrict__ cs, const int64_t* __restrict__ ri, int64_t m_full, int64_t r0,
                                                      int64_t r1, int64_t* __restrict__ col_len, int* __restrict__ error) {
  const int64_t c = gtid();
  if (c >= n) return;
  const int64_t b = cs[c], e = cs[c + 1];
  if (e < b) { *error = 1; col_len[c] = 0; return; }
  int64_t cnt = 0;
  for (int64_t k = b; k < e; ++k) {
    const int64_t r = ri[k];
    if (r < 0 || r >= m_full) { *error = 2; continue; }
    cnt += (r >= r0 && r < r1);
  }
  col_len[c] = cnt;
}
// Compacted entries of the row block in CSC order: local row, column, value.
__global__ void __launch_bounds__(kT) k_compact_columns(int64_t n, const int64_t* __restrict__ cs, const int64_t* __restrict__ ri, const double* __restrict__ va,
                                                        int64_t m_full, int64_t r0, int64_t r1, const int64_t* __restrict__ kt_start, int32_t* __restrict__ key,
                                                        int32_t* __restrict__ ecol, double* __restrict__ cval) {
  const int64_t c = gtid();
  if (c >= n) return;
  int64_t o = kt_start[c];
  for (int64_t k = cs[c]; k < cs[c + 1]; ++k) {
    const int64_t r = ri[k];
    if (r < 0 || r >= m_full || r < r0 || r >= r1) continue;
    key[o] = static_cast<int32_t>(r - r0);
    ecol[o] = static_cast<int32_t>(c);
    cval[o] = va[k];
    ++o;
  }
}
__global__ void __launch_bounds__(kT) k_histogram_rows(int64_t count, const int32_t* __restrict__ key, int32_t* __restrict__ row_len) {
  const int64_t e = gtid();
  if (e < count) atomicAdd(row_len + key[e], 1);  // integer counts: order-independent
}
__global__ void __launch_bounds__(kT) k_widen(int64_t n, const int32_t* __restrict__ in, int64_t* __restrict__ out) {
  const int64_t i = gtid();
  if (i < n) out[i] = in[i];
}

// ---- positions ------------------------------------------------------------------
__global__ void __launch_bounds__(kT) k_flag_split(int64_t rows, const int64_t* __restrict__ len, int64_t split_len, int64_t* __restrict__ flag) {
  const int64_t i = gtid();
  if (i < rows) flag[i] = len[i] > split_len ? 1 : 0;
}
// split rows go to positions [0, num_split) in index order; the others are
// listed (still in index order) in `rest`.
__global__ void __launch_bounds__(kT) k_partition_rows(int64_t rows, const int64_t* __restrict__ len, int64_t split_len, const int64_t* __restrict__ split_rank,
                                                       int32_t* __restrict__ row_of_pos, int32_t* __restrict__ rest) {
  const int64_t i = gtid();
  if (i >= rows) return;
  const int64_t sr = split_rank[i];
  if (len[i] > split_len) row_of_pos[sr] = static_cast<int32_t>(i);
  else rest[i - sr] = static_cast<int32_t>(i);
}
// One block per window of `sigma` (<= 4096) entries of `rest`: stable sort by
// descending length = bitonic sort of the unique composite keys
// ((split_len - len) << 12) | index_in_window.
__global__ void __launch_bounds__(1024) k_window_sort(int64_t rest_count, int sigma, const int32_t* __restrict__ rest, const int64_t* __restrict__ len,
                                                      int64_t split_len, int64_t num_split, int32_t* __restrict__ row_of_pos) {
  __shared__ unsigned int keys[4096];
  const int64_t w0 = static_cast<int64_t>(blockIdx.x) * sigma;
  const int cnt = static_cast<int>(min(static_cast<int64_t>(sigma), rest_count - w0));
  for (int k = threadIdx.x; k < 4096; k += blockDim.x) {
    unsigned int key = 0xFFFFFFFFu;
    if (k < cnt) key = (static_cast<unsigned int>(split_len - len[rest[w0 + k]]) << 12) | static_cast<unsigned int>(k);
    keys[k] = key;
  }
  __syncthreads();
  for (int size = 2; size <= 4096; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < 2048; t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool up = (lo & size) == 0;
        const unsigned int a = keys[lo], b = keys[hi];
        if ((a > b) == up) { keys[lo] = b; keys[hi] = a; }
      }
      __syncthreads();
    }
  }
  for (int k = threadIdx.x; k < cnt; k += blockDim.x) row_of_pos[num_split + w0 + k] = rest[w0 + (keys[k] & 0xFFFu)];
}
__global__ void __launch_bounds__(kT) k_invert(int64_t rows, const int32_t* __restrict__ row_of_pos, int32_t* __restrict__ pos_of_row) {
  const int64_t p = gtid();
  if (p < rows) pos_of_row[row_of_pos[p]] = static_cast<int32_t>(p);
}

// ---- slots ------------------------------------------------------------------------
__global__ void __launch_bounds__(kT) k_virtual_counts(int64_t num_split, const int32_t* __restrict__ row_of_pos, const int64_t* __restrict__ len, int64_t split_len,
                                                       int64_t* __restrict__ nv) {
  const int64_t i = gtid();
  if (i < num_split) nv[i] = (len[row_of_pos[i]] + split_len - 1) / split_len;
}
// Virtual slots of a split row are filled in TEAMS: the slots of the row that share a 32-slot slice
// take the elements of their common range round-robin (member k of a team of t gets elements
// e0 + k, e0 + k + t, ...), so that at every step of the row loop the lanes of the team read
// CONSECUTIVE entries of the row -- for a row whose columns are contiguous (transportation /
// multicommodity structure) the gathers of x coalesce into two 128-byte lines per warp instead of
// 32 sectors. slot_off = first element, slot_stride = t.
__global__ void __launch_bounds__(kT) k_virtual_slots(int64_t num_virtual, int64_t num_split, const int64_t* __restrict__ split_first64, const int32_t* __restrict__ row_of_pos,
                                                      const int64_t* __restrict__ len, int64_t split_len, int32_t* __restrict__ slot_len,
                                                      int32_t* __restrict__ slot_row, int64_t* __restrict__ slot_off, int32_t* __restrict__ slot_stride,
                                                      int32_t* __restrict__ virt_pos, int teams) {
  const int64_t v = gtid();
  if (v >= num_virtual) return;
  int64_t lo = 0, hi = num_split;  // last i with split_first[i] <= v
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (split_first64[mid] <= v) lo = mid; else hi = mid;
  }
  const int32_t row = row_of_pos[lo];
  const int64_t first = split_first64[lo], last = split_first64[lo + 1];  // the row's slots [first, last)
  // its team inside this slice (teams == 0, PDLP_B200_TEAM_SLOTS=0: every slot alone, entries [k T, (k + 1) T) of the row)
  const int64_t ts = teams ? max(first, (v >> 5) << 5) : v, te = teams ? min(last, ((v >> 5) + 1) << 5) : v + 1;
  const int64_t t = te - ts, k = v - ts;
  const int64_t e0 = (ts - first) * split_len, e1 = min(len[row], (te - first) * split_len);
  const int64_t team_len = e1 - e0;
  slot_len[v] = team_len > k ? static_cast<int32_t>((team_len - k + t - 1) / t) : 0;
  slot_row[v] = row;
  slot_off[v] = e0 + k;
  slot_stride[v] = static_cast<int32_t>(t);
  virt_pos[v] = static_cast<int32_t>(lo);
}
__global__ void __launch_bounds__(kT) k_plain_slots(int64_t rows, int64_t num_split, int64_t nvp, const int32_t* __restrict__ row_of_pos, const int64_t* __restrict__ len,
                                                    int32_t* __restrict__ slot_len, int32_t* __restrict__ slot_row, int64_t* __restrict__ slot_off) {
  const int64_t p = num_split + gtid();
  if (p >= rows) return;
  const int64_t slot = nvp + (p - num_split);
  const int32_t row = row_of_pos[p];
  slot_len[slot] = static_cast<int32_t>(len[row]);
  slot_row[slot] = row;
  slot_off[slot] = 0;
}
__global__ void __launch_bounds__(kT) k_fill_i32(int64_t n, int32_t* p, int32_t v) {
  const int64_t i = gtid();
  if (i < n) p[i] = v;
}
__global__ void __launch_bounds__(kT) k_slice_widths(int64_t num_slices, const int32_t* __restrict__ slot_len, int64_t* __restrict__ width32) {
  const int64_t s = gtid();
  if (s >= num_slices) return;
  int32_t w = 0;
  for (int l = 0; l < 32; ++l) w = max(w, slot_len[s * 32 + l]);
  width32[s] = static_cast<int64_t>(w) * 32;
}
__global__ void __launch_bounds__(kT) k_narrow_i32(int64_t n, const int64_t* __restrict__ in, int32_t* __restrict__ out) {
  const int64_t i = gtid();
  if (i < n) out[i] = static_cast<int32_t>(in[i]);
}

// ---- stable LSD radix sort of (key, payload) pairs --------------------------------
__global__ void __launch_bounds__(kSortT) k_radix_hist(int64_t count, const int32_t* __restrict__ key, int shift, int num_blocks, int64_t* __restrict__ hist) {
  __shared__ int h[kRadix];
  for (int d = threadIdx.x; d < kRadix; d += kSortT) h[d] = 0;
  __syncthreads();
  const int64_t start = static_cast<int64_t>(blockIdx.x) * kSortT * kSortTiles;
  for (int t = 0; t < kSortTiles; ++t) {
    const int64_t i = start + static_cast<int64_t>(t) * kSortT + threadIdx.x;
    if (i < count) atomicAdd(&h[(key[i] >> shift) & (kRadix - 1)], 1);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < kRadix; d += kSortT) hist[static_cast<int64_t>(d) * num_blocks + blockIdx.x] = h[d];
}
__global__ void __launch_bounds__(kSortT) k_radix_scatter(int64_t count, const int32_t* __restrict__ key_in, const int32_t* __restrict__ pay_in, int shift,
                                                          int num_blocks, const int64_t* __restrict__ offsets, int32_t* __restrict__ key_out,
                                                          int32_t* __restrict__ pay_out) {
  __shared__ int64_t run[kRadix];                  // next output index of each digit for this block
  __shared__ unsigned short wcnt[32][kRadix];     // per-warp digit counts of the current sub-tile
  __shared__ unsigned short wpre[32][kRadix];     // exclusive prefix over warps
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int d = threadIdx.x; d < kRadix; d += kSortT) run[d] = offsets[static_cast<int64_t>(d) * num_blocks + blockIdx.x];
  const int64_t start = static_cast<int64_t>(blockIdx.x) * kSortT * kSortTiles;
  for (int t = 0; t < kSortTiles; ++t) {
    for (int k = threadIdx.x; k < 32 * kRadix; k += kSortT) (&wcnt[0][0])[k] = 0;
    __syncthreads();
    const int64_t i = start + static_cast<int64_t>(t) * kSortT + threadIdx.x;
    const bool valid = i < count;
    int k = 0, p = 0;
    if (valid) { k = key_in[i]; p = pay_in[i]; }
    const int d = valid ? ((k >> shift) & (kRadix - 1)) : (kRadix + lane);  // invalid lanes form singleton groups
    const unsigned mask = __match_any_sync(0xffffffffu, d);
    const int rank = __popc(mask & ((1u << lane) - 1u));
    if (valid && rank == 0) wcnt[warp][d] = static_cast<unsigned short>(__popc(mask));
    __syncthreads();
    if (threadIdx.x < kRadix) {
      unsigned int acc = 0;
      for (int w = 0; w < 32; ++w) {
        wpre[w][threadIdx.x] = static_cast<unsigned short>(acc);
        acc += wcnt[w][threadIdx.x];
      }
      wcnt[0][threadIdx.x] = static_cast<unsigned short>(acc);  // total of the digit in this sub-tile
    }
    __syncthreads();
    if (valid) {
      const int64_t dst = run[d] + wpre[warp][d] + rank;
      key_out[dst] = k;
      pay_out[dst] = p;
    }
    __syncthreads();
    if (threadIdx.x < kRadix) run[threadIdx.x] += wcnt[0][threadIdx.x];
    __syncthreads();
  }
}
__global__ void __launch_bounds__(kT) k_iota(int64_t n, int32_t* p) {
  const int64_t i = gtid();
  if (i < n) p[i] = static_cast<int32_t>(i);
}

// ---- fill ---------------------------------------------------------------------------
// One thread per slot. `order` maps (row start + k) to an entry index (identity
// for the column copy, the sorted permutation for the row copy); `idx_of_entry`
// is the other orientation's original index of an entry and `other_pos` its
// position map (nullptr = keep original indices).
__global__ void __launch_bounds__(kT) k_fill_sell(int64_t num_slots, const int64_t* __restrict__ slice_ptr, const int32_t* __restrict__ slot_len,
                                                  const int32_t* __restrict__ slot_row, const int64_t* __restrict__ slot_off, const int32_t* __restrict__ slot_stride,
                                                  const int64_t* __restrict__ row_start, const int32_t* __restrict__ order,
                                                  const int32_t* __restrict__ idx_of_entry, const double* __restrict__ cval,
                                                  const int32_t* __restrict__ other_pos, int32_t* __restrict__ col, double* __restrict__ val) {
  const int64_t slot = gtid();
  if (slot >= num_slots) return;
  const int32_t row = slot_row[slot];
  if (row < 0) return;
  const int64_t base = slice_ptr[slot >> 5] + (slot & 31);
  const int64_t src0 = row_start[row] + slot_off[slot];
  const int64_t stride = slot_stride[slot];
  const int len = slot_len[slot];
  for (int j = 0; j < len; ++j) {
    const int64_t e = order != nullptr ? order[src0 + j * stride] : src0 + j * stride;
    const int32_t other = idx_of_entry[e];
    col[base + static_cast<int64_t>(j) * 32] = other_pos != nullptr ? other_pos[other] : other;
    val[base + static_cast<int64_t>(j) * 32] = cval[e];
  }
}

// values of the column copy back into the caller's CSC order
__global__ void __launch_bounds__(kT) k_unfill_values(int64_t num_slots, const int64_t* __restrict__ slice_ptr, const int32_t* __restrict__ slot_len,
                                                      const int32_t* __restrict__ slot_row, const int64_t* __restrict__ slot_off, const int32_t* __restrict__ slot_stride,
                                                      const int64_t* __restrict__ row_start, const double* __restrict__ val, double* __restrict__ out) {
  const int64_t slot = gtid();
  if (slot >= num_slots) return;
  const int32_t row = slot_row[slot];
  if (row < 0) return;
  const int64_t base = slice_ptr[slot >> 5] + (slot & 31);
  const int64_t dst0 = row_start[row] + slot_off[slot];
  const int64_t stride = slot_stride[slot];
  const int len = slot_len[slot];
  for (int j = 0; j < len; ++j) out[dst0 + j * stride] = val[base + static_cast<int64_t>(j) * 32];
}

}  // namespace build_kernels

using namespace build_kernels;

namespace {

inline int Blk(int64_t n, int t = kT) { return static_cast<int>(std::max<int64_t>(1, (n + t - 1) / t)); }

// Persistent arrays of the images: from the stream-ordered pool of the device (see
// Device::Device) on the stream of the running build; released with cudaFree.
thread_local cudaStream_t g_build_stream = nullptr;
struct BuildStreamScope {
  explicit BuildStreamScope(cudaStream_t s) { g_build_stream = s; }
  ~BuildStreamScope() { g_build_stream = nullptr; }
};
template <class T>
T* DevAlloc(int64_t count) {
  T* p = nullptr;
  const size_t bytes = sizeof(T) * static_cast<size_t>(std::max<int64_t>(count, 1) + 8);
  if (g_build_stream != nullptr) CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&p), bytes, g_build_stream));
  else CUDA_OK(cudaMalloc(&p, bytes));
  return p;
}

// RAII pool of temporaries freed when the build ends. They come from the
// device's stream-ordered memory pool, whose release threshold is raised so
// that the memory stays with the process: the ~40 temporaries of a build then
// cost no cudaMalloc / cudaFree round trips (which also synchronise the
// device) after the first build, and the builds of later solves start warm.
struct Temps {
  cudaStream_t stream;
  std::vector<void*> ptrs;
  explicit Temps(cudaStream_t s) : stream(s) {
    static bool pool_configured[64] = {false};
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64 && !pool_configured[dev]) {
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t keep = ~uint64_t{0};
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
      pool_configured[dev] = true;
    }
    cudaGetLastError();
  }
  template <class T>
  T* get(int64_t count) {
    T* p = nullptr;
    CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&p), sizeof(T) * static_cast<size_t>(std::max<int64_t>(count, 1) + 8), stream));
    ptrs.push_back(p);
    return p;
  }
  ~Temps() { for (void* p : ptrs) cudaFreeAsync(p, stream); }
};

struct Scanner {
  cudaStream_t stream;
  int64_t* sums = nullptr;
  int64_t* total_dev = nullptr;
  int64_t* total_host = nullptr;  // pinned
  int64_t sums_cap = 0;
  // (stream-ordered allocations: cudaMalloc / cudaFree / cudaMallocHost synchronise the device and cost
  // 0.1-1 ms each; the pinned word for the totals is allocated once per host thread and kept)
  explicit Scanner(cudaStream_t s) : stream(s) {
    CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&total_dev), sizeof(int64_t) * 8, stream));
    thread_local int64_t* pinned = nullptr;
    if (pinned == nullptr) CUDA_OK(cudaMallocHost(&pinned, sizeof(int64_t) * 8));
    total_host = pinned;
  }
  ~Scanner() { if (sums != nullptr) cudaFreeAsync(sums, stream); cudaFreeAsync(total_dev, stream); }
  // out[i] = sum in[0..i); returns the total (synchronises).
  int64_t Exclusive(const int64_t* in, int64_t* out, int64_t n, int64_t* launches) {
    if (n <= 0) return 0;
    const int64_t per_block = static_cast<int64_t>(kScanT) * kScanItems;
    const int64_t nb = (n + per_block - 1) / per_block;
    if (nb > sums_cap) {
      if (sums != nullptr) cudaFreeAsync(sums, stream);
      CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&sums), sizeof(int64_t) * (nb + 8), stream));
      sums_cap = nb;
    }
    k_scan_blocks<<<static_cast<int>(nb), kScanT, 0, stream>>>(in, out, n, sums);
    k_scan_sums<<<1, kScanT, 0, stream>>>(sums, nb, total_dev);
    k_scan_add<<<static_cast<int>(nb), kScanT, 0, stream>>>(out, n, sums);
    *launches += 3;
    CUDA_OK(cudaMemcpyAsync(total_host, total_dev, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
    CUDA_OK(cudaStreamSynchronize(stream));
    return *total_host;
  }
};

int EnvIntB(const char* name, int dflt) {
  const char* v = std::getenv(name);
  if (v == nullptr || *v == 0) return dflt;
  return std::atoi(v);
}

int32_t ChooseSplitLenDev(int64_t nnz) {  // same rule as sell_builder.cc
  const int forced = EnvIntB("PDLP_B200_SPLIT_LEN", 0);
  if (forced > 0) return forced;
  int64_t t = 64;
  while (t * 262144 < nnz) t *= 2;
  return static_cast<int32_t>(t);
}

// Positions + slots + slice pointers of one orientation. `len` = entries per
// logical row (int64, device). Allocates the persistent arrays of `out`.
struct Orientation {
  int32_t* row_of_pos = nullptr;  // persistent
  int32_t* pos_of_row = nullptr;  // temp (needed by the other orientation's fill)
  int32_t* slot_row = nullptr;    // persistent only for the column copy
  int64_t* slot_off = nullptr;
  int32_t* slot_stride = nullptr;  // 1, or the team size of a virtual slot (k_virtual_slots)
};

void BuildOrientation(cudaStream_t stream, Scanner& scan, Temps& tmp, int64_t rows, int64_t gathered_len, const int64_t* len, int32_t split_len, int sigma,
                      SellDev* out, Orientation* o, int64_t* launches) {
  out->num_rows = rows;
  out->num_cols = gathered_len;
  o->row_of_pos = DevAlloc<int32_t>(rows);
  o->pos_of_row = tmp.get<int32_t>(rows);
  int64_t* flag = tmp.get<int64_t>(rows);
  int64_t* split_rank = tmp.get<int64_t>(rows);
  int32_t* rest = tmp.get<int32_t>(rows);
  int64_t num_split = 0;
  if (rows > 0) {
    k_flag_split<<<Blk(rows), kT, 0, stream>>>(rows, len, split_len, flag);
    *launches += 1;
    num_split = scan.Exclusive(flag, split_rank, rows, launches);
    k_partition_rows<<<Blk(rows), kT, 0, stream>>>(rows, len, split_len, split_rank, o->row_of_pos, rest);
    const int64_t rest_count = rows - num_split;
    if (rest_count > 0) {
      const int64_t windows = (rest_count + sigma - 1) / sigma;
      k_window_sort<<<static_cast<int>(windows), 1024, 0, stream>>>(rest_count, sigma, rest, len, split_len, num_split, o->row_of_pos);
    }
    k_invert<<<Blk(rows), kT, 0, stream>>>(rows, o->row_of_pos, o->pos_of_row);
    *launches += 3;
  }
  out->num_split = num_split;
  // virtual slots of the split rows
  int64_t num_virtual = 0;
  int64_t* split_first64 = tmp.get<int64_t>(num_split + 1);
  if (num_split > 0) {
    int64_t* nv = tmp.get<int64_t>(num_split);
    k_virtual_counts<<<Blk(num_split), kT, 0, stream>>>(num_split, o->row_of_pos, len, split_len, nv);
    *launches += 1;
    num_virtual = scan.Exclusive(nv, split_first64, num_split, launches);
  }
  CUDA_OK(cudaMemcpyAsync(split_first64 + num_split, &num_virtual, sizeof(int64_t), cudaMemcpyHostToDevice, stream));
  CUDA_OK(cudaStreamSynchronize(stream));
  const int64_t nvp = (num_virtual + 31) / 32 * 32;
  const int64_t num_slots = (nvp + (rows - num_split) + 31) / 32 * 32;
  if (num_slots >= (int64_t{1} << 31)) throw std::runtime_error("too many rows for int32 slot indices");
  out->num_virtual_padded = nvp;
  out->num_slots = num_slots;
  out->slot_len = DevAlloc<int32_t>(num_slots);
  out->virt_pos = DevAlloc<int32_t>(nvp);
  out->split_first = DevAlloc<int32_t>(num_split + 1);
  out->virt_partial = DevAlloc<double>(nvp + 32);
  o->slot_row = DevAlloc<int32_t>(num_slots);
  o->slot_off = DevAlloc<int64_t>(num_slots);
  o->slot_stride = DevAlloc<int32_t>(num_slots);
  CUDA_OK(cudaMemsetAsync(out->slot_len, 0, sizeof(int32_t) * std::max<int64_t>(num_slots, 1), stream));
  CUDA_OK(cudaMemsetAsync(o->slot_off, 0, sizeof(int64_t) * std::max<int64_t>(num_slots, 1), stream));
  if (num_slots > 0) k_fill_i32<<<Blk(num_slots), kT, 0, stream>>>(num_slots, o->slot_row, -1);
  if (num_slots > 0) k_fill_i32<<<Blk(num_slots), kT, 0, stream>>>(num_slots, o->slot_stride, 1);
  if (nvp > 0) k_fill_i32<<<Blk(nvp), kT, 0, stream>>>(nvp, out->virt_pos, -1);
  k_narrow_i32<<<Blk(num_split + 1), kT, 0, stream>>>(num_split + 1, split_first64, out->split_first);
  if (num_virtual > 0)
    k_virtual_slots<<<Blk(num_virtual), kT, 0, stream>>>(num_virtual, num_split, split_first64, o->row_of_pos, len, split_len, out->slot_len, o->slot_row, o->slot_off,
                                                         o->slot_stride, out->virt_pos, EnvIntB("PDLP_B200_TEAM_SLOTS", 1) != 0 ? 1 : 0);
  if (rows - num_split > 0)
    k_plain_slots<<<Blk(rows - num_split), kT, 0, stream>>>(rows, num_split, nvp, o->row_of_pos, len, out->slot_len, o->slot_row, o->slot_off);
  *launches += 5;
  // slice pointers
  const int64_t num_slices = num_slots / 32;
  out->slice_ptr = DevAlloc<int64_t>(num_slices + 1);
  int64_t padded = 0;
  if (num_slices > 0) {
    int64_t* width32 = tmp.get<int64_t>(num_slices);
    k_slice_widths<<<Blk(num_slices), kT, 0, stream>>>(num_slices, out->slot_len, width32);
    *launches += 1;
    padded = scan.Exclusive(width32, out->slice_ptr, num_slices, launches);
  }
  CUDA_OK(cudaMemcpyAsync(out->slice_ptr + num_slices, &padded, sizeof(int64_t), cudaMemcpyHostToDevice, stream));
  CUDA_OK(cudaStreamSynchronize(stream));
  out->padded_nnz = padded;
  out->col = DevAlloc<int32_t>(padded);
  out->val = DevAlloc<double>(padded);
  CUDA_OK(cudaMemsetAsync(out->col, 0, sizeof(int32_t) * std::max<int64_t>(padded, 1), stream));
  CUDA_OK(cudaMemsetAsync(out->val, 0, sizeof(double) * std::max<int64_t>(padded, 1), stream));
}

}  // namespace

void Device::BuildSellPair(const PdlpProblemView& v, int64_t row_begin, int64_t row_end, int sigma, bool natural_primal_order, SellDev* rows_out,
                           SellDev* cols_out, int32_t** dual_perm, int32_t** primal_perm, DeviceBuildInfo* info) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int64_t n = v.num_variables, m_full = v.num_constraints;
  if (row_begin < 0 || row_end > m_full || row_begin > row_end) throw std::runtime_error("bad row range");
  if (n >= (int64_t{1} << 31) - 64 || m_full >= (int64_t{1} << 31) - 64) throw std::runtime_error("dimension exceeds int32 index range");
  sigma = std::min(4096, std::max(32, EnvIntB("PDLP_B200_SIGMA", sigma)));
  const int64_t m = row_end - row_begin;
  const int64_t nnz_full = n > 0 ? v.col_starts[n] : 0;
  const auto t_build_start = std::chrono::steady_clock::now();
  BuildStreamScope build_scope(stream);
  Temps tmp(stream);
  Scanner scan(stream);
  // ---- upload the raw CSC arrays
  int64_t* cs = tmp.get<int64_t>(n + 1);
  int64_t* ri = tmp.get<int64_t>(nnz_full);
  double* va = tmp.get<double>(nnz_full);
  if (n > 0) CUDA_OK(cudaMemcpyAsync(cs, v.col_starts, sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice, stream));
  else CUDA_OK(cudaMemsetAsync(cs, 0, sizeof(int64_t), stream));
  if (nnz_full > 0) {
    CUDA_OK(cudaMemcpyAsync(ri, v.row_indices, sizeof(int64_t) * nnz_full, cudaMemcpyHostToDevice, stream));
    CUDA_OK(cudaMemcpyAsync(va, v.values, sizeof(double) * nnz_full, cudaMemcpyHostToDevice, stream));
  }
  int* error = tmp.get<int>(1);
  CUDA_OK(cudaMemsetAsync(error, 0, sizeof(int), stream));
  const char* trace_env = std::getenv("PDLP_B200_TRACE");
  const bool trace = trace_env != nullptr && trace_env[0] == '1';
  const auto t_build0 = std::chrono::steady_clock::now();
  if (trace) {
    CUDA_OK(cudaStreamSynchronize(stream));
    std::fprintf(stderr, "[pdlp_b200 trace] SELL build: host->device copy of the CSC arrays (%.0f MB) %.4f s\n", (16.0 * nnz_full + 8.0 * (n + 1)) / 1e6,
                 std::chrono::duration<double>(std::chrono::steady_clock::now() - t_build_start).count());
  }
  // ---- column lengths and the compacted column-major entries of the row block
  int64_t* col_len = tmp.get<int64_t>(n);
  int64_t* kt_start = DevAlloc<int64_t>(n + 1);  // persistent (value download)
  int64_t nnz = 0;
  if (n > 0) {
    k_count_columns<<<Blk(n), kT, 0, stream>>>(n, cs, ri, m_full, row_begin, row_end, col_len, error);
    launches_ += 1;
    nnz = scan.Exclusive(col_len, kt_start, n, &launches_);
  }
  CUDA_OK(cudaMemcpyAsync(kt_start + n, &nnz, sizeof(int64_t), cudaMemcpyHostToDevice, stream));
  int herr = 0;
  CUDA_OK(cudaMemcpyAsync(&herr, error, sizeof(int), cudaMemcpyDeviceToHost, stream));
  CUDA_OK(cudaStreamSynchronize(stream));
  if (herr == 1) { cudaFree(kt_start); throw std::runtime_error("col_starts is not monotone"); }
  if (herr == 2) { cudaFree(kt_start); throw std::runtime_error("row index out of range"); }
  int32_t* key = tmp.get<int32_t>(nnz);
  int32_t* ecol = tmp.get<int32_t>(nnz);
  double* cval = tmp.get<double>(nnz);
  if (n > 0) {
    k_compact_columns<<<Blk(n), kT, 0, stream>>>(n, cs, ri, va, m_full, row_begin, row_end, kt_start, key, ecol, cval);
    launches_ += 1;
  }
  // ---- row lengths
  int32_t* row_len32 = tmp.get<int32_t>(m);
  int64_t* row_len = tmp.get<int64_t>(m);
  int64_t* row_start = tmp.get<int64_t>(m + 1);
  CUDA_OK(cudaMemsetAsync(row_len32, 0, sizeof(int32_t) * std::max<int64_t>(m, 1), stream));
  if (nnz > 0) k_histogram_rows<<<Blk(nnz), kT, 0, stream>>>(nnz, key, row_len32);
  if (m > 0) {
    k_widen<<<Blk(m), kT, 0, stream>>>(m, row_len32, row_len);
    scan.Exclusive(row_len, row_start, m, &launches_);
  }
  launches_ += 2;
  // ---- positions, slots, slice pointers of both orientations
  const int32_t split_len = ChooseSplitLenDev(nnz);
  Orientation orow, ocol;
  BuildOrientation(stream, scan, tmp, m, n, row_len, split_len, sigma, rows_out, &orow, &launches_);
  BuildOrientation(stream, scan, tmp, n, m, col_len, split_len, sigma, cols_out, &ocol, &launches_);
  // ---- row-major order of the entries: stable LSD radix sort by local row
  int32_t* order = nullptr;
  if (nnz > 0) {
    int32_t* k0 = key;
    int32_t* p0 = tmp.get<int32_t>(nnz);
    int32_t* k1 = tmp.get<int32_t>(nnz);
    int32_t* p1 = tmp.get<int32_t>(nnz);
    k_iota<<<Blk(nnz), kT, 0, stream>>>(nnz, p0);
    launches_ += 1;
    int bits = 1;
    while ((int64_t{1} << bits) < std::max<int64_t>(m, 2)) ++bits;
    const int64_t per_block = static_cast<int64_t>(kSortT) * kSortTiles;
    const int nb = static_cast<int>((nnz + per_block - 1) / per_block);
    int64_t* hist = tmp.get<int64_t>(static_cast<int64_t>(kRadix) * nb);
    int64_t* offs = tmp.get<int64_t>(static_cast<int64_t>(kRadix) * nb);
    // the sort consumes a copy of the keys: `key` itself is still needed by the fill
    int32_t* kcopy = tmp.get<int32_t>(nnz);
    CUDA_OK(cudaMemcpyAsync(kcopy, key, sizeof(int32_t) * nnz, cudaMemcpyDeviceToDevice, stream));
    k0 = kcopy;
    for (int shift = 0; shift < bits; shift += kRadixBits) {
      k_radix_hist<<<nb, kSortT, 0, stream>>>(nnz, k0, shift, nb, hist);
      launches_ += 1;
      scan.Exclusive(hist, offs, static_cast<int64_t>(kRadix) * nb, &launches_);
      k_radix_scatter<<<nb, kSortT, 0, stream>>>(nnz, k0, p0, shift, nb, offs, k1, p1);
      launches_ += 1;
      std::swap(k0, k1);
      std::swap(p0, p1);
    }
    order = p0;
  }
  // ---- fill both images
  if (rows_out->num_slots > 0)
    k_fill_sell<<<Blk(rows_out->num_slots), kT, 0, stream>>>(rows_out->num_slots, rows_out->slice_ptr, rows_out->slot_len, orow.slot_row, orow.slot_off, orow.slot_stride, row_start, order,
                                                             ecol, cval, natural_primal_order ? nullptr : ocol.pos_of_row, rows_out->col, rows_out->val);
  if (cols_out->num_slots > 0)
    k_fill_sell<<<Blk(cols_out->num_slots), kT, 0, stream>>>(cols_out->num_slots, cols_out->slice_ptr, cols_out->slot_len, ocol.slot_row, ocol.slot_off, ocol.slot_stride, kt_start, nullptr,
                                                             key, cval, orow.pos_of_row, cols_out->col, cols_out->val);
  launches_ += 2;
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaStreamSynchronize(stream));
  if (trace)
    std::fprintf(stderr, "[pdlp_b200 trace] SELL build: device passes %.4f s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t_build0).count());
  cudaFreeAsync(orow.slot_row, stream);
  cudaFreeAsync(orow.slot_off, stream);
  cudaFreeAsync(orow.slot_stride, stream);
  *dual_perm = orow.row_of_pos;
  *primal_perm = ocol.row_of_pos;
  info->n = n;
  info->m = m;
  info->nnz = nnz;
  info->col_slot_row = ocol.slot_row;
  info->col_slot_off = ocol.slot_off;
  info->col_slot_stride = ocol.slot_stride;
  info->col_start = kt_start;
}

// SELL-32 image of (K[:, col_begin:col_end])^T: one logical row per column of the
// slice, all constraint rows kept. Stored indices are row_pos_global[row]: the
// position of the row in the box-wide dual order of a row-sharded solve (block
// offset of the owning rank + position inside its row image), so that the image
// gathers from the all-gathered dual vector. *row_of_pos_out: column (relative
// to col_begin) at each position of the image (device, owned by the caller).
void Device::BuildColumnSliceImage(const PdlpProblemView& v, int64_t col_begin, int64_t col_end, const int32_t* row_pos_global, int sigma,
                                   SellDev* out, int32_t** row_of_pos_out) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int64_t n_all = v.num_variables, m_full = v.num_constraints;
  if (col_begin < 0 || col_end > n_all || col_begin > col_end) throw std::runtime_error("bad column range");
  sigma = std::min(4096, std::max(32, EnvIntB("PDLP_B200_SIGMA", sigma)));
  const int64_t n = col_end - col_begin;
  const int64_t first = n_all > 0 ? v.col_starts[col_begin] : 0;
  const int64_t nnz = n_all > 0 ? v.col_starts[col_end] - first : 0;
  BuildStreamScope build_scope(stream);
  Temps tmp(stream);
  Scanner scan(stream);
  std::vector<int64_t> rel(n + 1, 0);
  for (int64_t c = 0; c <= n && n_all > 0; ++c) rel[c] = v.col_starts[col_begin + c] - first;
  int64_t* cs = tmp.get<int64_t>(n + 1);
  int64_t* ri = tmp.get<int64_t>(nnz);
  double* va = tmp.get<double>(nnz);
  CUDA_OK(cudaMemcpyAsync(cs, rel.data(), sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice, stream));
  if (nnz > 0) {
    CUDA_OK(cudaMemcpyAsync(ri, v.row_indices + first, sizeof(int64_t) * nnz, cudaMemcpyHostToDevice, stream));
    CUDA_OK(cudaMemcpyAsync(va, v.values + first, sizeof(double) * nnz, cudaMemcpyHostToDevice, stream));
  }
  int* error = tmp.get<int>(1);
  CUDA_OK(cudaMemsetAsync(error, 0, sizeof(int), stream));
  int64_t* col_len = tmp.get<int64_t>(n);
  int64_t* kt_start = tmp.get<int64_t>(n + 1);
  int64_t total = 0;
  if (n > 0) {
    k_count_columns<<<Blk(n), kT, 0, stream>>>(n, cs, ri, m_full, 0, m_full, col_len, error);
    launches_ += 1;
    total = scan.Exclusive(col_len, kt_start, n, &launches_);
  }
  CUDA_OK(cudaMemcpyAsync(kt_start + n, &total, sizeof(int64_t), cudaMemcpyHostToDevice, stream));
  int herr = 0;
  CUDA_OK(cudaMemcpyAsync(&herr, error, sizeof(int), cudaMemcpyDeviceToHost, stream));
  CUDA_OK(cudaStreamSynchronize(stream));
  if (herr != 0) throw std::runtime_error("bad CSC input (column slice)");
  int32_t* key = tmp.get<int32_t>(total);
  int32_t* ecol = tmp.get<int32_t>(total);
  double* cval = tmp.get<double>(total);
  if (n > 0) {
    k_compact_columns<<<Blk(n), kT, 0, stream>>>(n, cs, ri, va, m_full, 0, m_full, kt_start, key, ecol, cval);
    launches_ += 1;
  }
  Orientation ocol;
  BuildOrientation(stream, scan, tmp, n, m_full, col_len, ChooseSplitLenDev(total), sigma, out, &ocol, &launches_);
  if (out->num_slots > 0) {
    k_fill_sell<<<Blk(out->num_slots), kT, 0, stream>>>(out->num_slots, out->slice_ptr, out->slot_len, ocol.slot_row, ocol.slot_off, ocol.slot_stride, kt_start, nullptr, key, cval,
                                                        row_pos_global, out->col, out->val);
    launches_ += 1;
  }
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaStreamSynchronize(stream));
  cudaFreeAsync(ocol.slot_row, stream);
  cudaFreeAsync(ocol.slot_off, stream);
  cudaFreeAsync(ocol.slot_stride, stream);
  *row_of_pos_out = ocol.row_of_pos;
}

void Device::FreeBuildInfo(DeviceBuildInfo& info) {
  cudaFree(info.col_slot_row);
  cudaFree(info.col_slot_off);
  cudaFree(info.col_slot_stride);
  cudaFree(info.col_start);
  info = DeviceBuildInfo();
}

void Device::DownloadValuesCscFromSell(const SellDev& cols, const DeviceBuildInfo& info, double* values_host) {
  if (info.nnz <= 0) return;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  double* out = DevAlloc<double>(info.nnz);
  k_unfill_values<<<Blk(cols.num_slots), kT, 0, stream>>>(cols.num_slots, cols.slice_ptr, cols.slot_len, info.col_slot_row, info.col_slot_off, info.col_slot_stride, info.col_start, cols.val, out);
  ++launches_;
  CUDA_OK(cudaMemcpyAsync(values_host, out, sizeof(double) * info.nnz, cudaMemcpyDeviceToHost, stream));
  CUDA_OK(cudaStreamSynchronize(stream));
  cudaFree(out);
}

}  // namespace pdlp_b200
