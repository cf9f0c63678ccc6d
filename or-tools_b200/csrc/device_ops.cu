// device_ops.cu -- hand-written sm_100a kernels of the PDLP hot path.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo --extended-lambda
//
// Kernel families (DESIGN.md has the roofline of each):
//   k_sell<MODE,NS>       SELL-32 thread-per-row sparse product with a fused
//                         epilogue functor and fused block reductions. MODE
//                         selects dot / max|.| / sum-of-squares accumulation
//                         (SpMV pair, Ruiz LInf norms, L2 norms).
//   k_primal_step         x' = proj(x - tau (c - K^T y)), x~ = 2x' - x, ||dx||^2,
//                         deferred primal average (pdhg.cc:1834-1901).
//   k_sell + DualEpi      K x~ fused with the dual update, ||dy||^2, deferred
//                         dual average (pdhg.cc:1902-1931).
//   k_sell + KtyEpi       K^T y' fused with the nonlinearity dot product
//                         (pdhg.cc:2588-2592).
//   k_step_decide         fixed-order final reductions + the adaptive step-size
//                         rule / accept test on the device (pdhg.cc:2595-2640).
//   k_reduce<NS,NM>       generic deterministic multi-reduction (KKT norms,
//                         stats, distances).
//   k_tr_*                trust-region threshold search by radix-16 bisection
//                         on the bit pattern of the critical step sizes.
// All reductions are warp-shuffle + shared-memory block reductions written to
// per-block slots and combined in a fixed order: no floating-point atomics, so
// results are run-to-run deterministic.
#include <cuda_runtime.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <type_traits>

#include "comm.h"
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

namespace kernels {

constexpr int kThreads = 256;
constexpr int kMaxReduceBlocks = 148 * 8;
constexpr double kInfD = __builtin_huge_val();

// Programmatic dependent launch (sm_90+): the kernels of the step loop are
// launched with programmaticStreamSerialization, so the blocks of kernel k+1
// are scheduled while the last wave of kernel k drains. Every such kernel first
// lets its own dependents go (pdl_trigger) and then waits until the preceding
// grid has completed and its writes are visible (pdl_wait) before it touches
// global memory. Both are no-ops for a kernel launched the ordinary way.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block reduction of NS sums followed by NM maxes; thread 0 stores them to out.
template <int NS, int NM, int BT = kThreads>
__device__ __forceinline__ void block_reduce_store(const double* s, const double* m, double* out) {
  __shared__ double sh[(NS + NM > 0 ? NS + NM : 1)][BT / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NS; ++k) {
    const double v = warp_sum(s[k]);
    if (lane == 0) sh[k][warp] = v;
  }
#pragma unroll
  for (int k = 0; k < NM; ++k) {
    const double v = warp_max(m[k]);
    if (lane == 0) sh[NS + k][warp] = v;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      double v = lane < BT / 32 ? sh[k][lane] : 0.0;
      v = warp_sum(v);
      if (lane == 0) out[k] = v;
    }
#pragma unroll
    for (int k = 0; k < NM; ++k) {
      double v = lane < BT / 32 ? sh[NS + k][lane] : -kInfD;
      v = warp_max(v);
      if (lane == 0) out[NS + k] = v;
    }
  }
}

template <int NS, int NM, class F>
__global__ void __launch_bounds__(kThreads) k_reduce(int64_t n, F f, double* partials) {
  double s[NS > 0 ? NS : 1], m[NM > 0 ? NM : 1];
#pragma unroll
  for (int k = 0; k < NS; ++k) s[k] = 0.0;
#pragma unroll
  for (int k = 0; k < NM; ++k) m[k] = -kInfD;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * kThreads) f(i, s, m);
  block_reduce_store<NS, NM>(s, m, partials + static_cast<int64_t>(blockIdx.x) * (NS + NM));
}

// Fixed-order combination of per-block partials (one block).
template <int NS, int NM>
__global__ void __launch_bounds__(kThreads) k_reduce_final(int nblocks, const double* partials, double* out) {
  double s[NS > 0 ? NS : 1], m[NM > 0 ? NM : 1];
#pragma unroll
  for (int k = 0; k < NS; ++k) s[k] = 0.0;
#pragma unroll
  for (int k = 0; k < NM; ++k) m[k] = -kInfD;
  for (int b = threadIdx.x; b < nblocks; b += kThreads) {
    const double* p = partials + static_cast<int64_t>(b) * (NS + NM);
#pragma unroll
    for (int k = 0; k < NS; ++k) s[k] += p[k];
#pragma unroll
    for (int k = 0; k < NM; ++k) m[k] = fmax(m[k], p[NS + k]);
  }
  block_reduce_store<NS, NM>(s, m, out);
}

template <class F>
__global__ void __launch_bounds__(kThreads) k_for(int64_t n, F f) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
  if (i < n) f(i);
}

// ------------------------------------------------------------------ SELL ---
enum { kDot = 0, kMaxAbs = 1, kSumSq = 2 };

template <int MODE>
__device__ __forceinline__ double combine(double acc, double v, double xv) {
  if (MODE == kDot) return acc + v * xv;
  if (MODE == kMaxAbs) return fmax(acc, fabs(v * xv));
  const double t = v * xv;
  return acc + t * t;
}

// One thread per slot; lanes of a warp read consecutive addresses of the
// slice (coalesced 256 B value / 128 B index requests, streamed with
// evict-first); x is gathered through L2.
// CG: the gathered vector changes while the kernel runs (the persistent step loop of a row-sharded
// solve, k_peer_loop): gathers go to L2 (ld.global.cg), never through the non-coherent L1 path.
template <bool CG>
__device__ __forceinline__ double gather_ld(const double* p) { return CG ? __ldcg(p) : __ldg(p); }
template <int MODE, int V, bool CG = false>
__device__ __forceinline__ double sell_row(const SellDev& a, int64_t slot, const double* __restrict__ x) {
  const int64_t base = a.slice_ptr[slot >> 5] + (slot & 31);
  const int n = a.slot_len[slot];
  const double* __restrict__ val = a.val + base;
  const int32_t* __restrict__ col = a.col + base;
  double acc = 0.0;
  int j = 0;
  if (V == 0) {
    for (; j + 4 <= n; j += 4) {
      double v[4];
      int c[4];
      double xv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        v[u] = __ldcs(val + static_cast<int64_t>(j + u) * 32);
        c[u] = __ldcs(col + static_cast<int64_t>(j + u) * 32);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) xv[u] = gather_ld<CG>(x + c[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u) acc = combine<MODE>(acc, v[u], xv[u]);
    }
  } else {
    // Software-pipelined: the value / index loads of chunk k+1 are issued
    // right after the gathers of chunk k, so the DRAM latency of the streams
    // overlaps the L2 latency of the gathers and the gather queue never drains
    // while a warp waits for its next indices.
    constexpr int U = (V == 1 || V == 8) ? 4 : 8;
    if (n >= U) {
      double v[U];
      int c[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        c[u] = __ldcs(col + static_cast<int64_t>(u) * 32);
        v[u] = __ldcs(val + static_cast<int64_t>(u) * 32);
      }
      for (;;) {
        double xv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) xv[u] = gather_ld<CG>(x + c[u]);
        j += U;
        const bool more = j + U <= n;
        double vn[U];
        int cn[U];
        if (more) {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            cn[u] = __ldcs(col + static_cast<int64_t>(j + u) * 32);
            vn[u] = __ldcs(val + static_cast<int64_t>(j + u) * 32);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc = combine<MODE>(acc, v[u], xv[u]);
        if (!more) break;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          v[u] = vn[u];
          c[u] = cn[u];
        }
      }
    }
  }
  for (; j < n; ++j) acc = combine<MODE>(acc, __ldcs(val + static_cast<int64_t>(j) * 32), gather_ld<CG>(x + __ldcs(col + static_cast<int64_t>(j) * 32)));
  return acc;
}

// The row loop of k_peer_loop: there a phase ends when its LAST slice does, so the latency of one
// slice counts, not only the throughput. Eight slots per trip, the tail trip masked instead of a
// serial remainder, software-pipelined like sell_row<., 1>: a 20-entry row is 3 dependent gather
// round trips instead of 5-6. Same accumulation order as sell_row (j = 0 .. n-1). Gathers at L2
// (the gathered vector changes inside the launch).
__device__ __forceinline__ double sell_row8_cg(const SellDev& a, int64_t slot, const double* x) {
  const int64_t base = a.slice_ptr[slot >> 5] + (slot & 31);
  const int n = a.slot_len[slot];
  const double* __restrict__ val = a.val + base;
  const int32_t* __restrict__ col = a.col + base;
  constexpr int U = 8;
  double acc = 0.0;
  double v[U];
  int c[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const bool ok = u < n;
    c[u] = ok ? __ldcs(col + static_cast<int64_t>(u) * 32) : -1;
    v[u] = ok ? __ldcs(val + static_cast<int64_t>(u) * 32) : 0.0;
  }
  for (int j = 0; j < n; j += U) {
    double xv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) xv[u] = c[u] >= 0 ? __ldcg(x + c[u]) : 0.0;
    double vn[U];
    int cn[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool ok = j + U + u < n;
      cn[u] = ok ? __ldcs(col + static_cast<int64_t>(j + U + u) * 32) : -1;
      vn[u] = ok ? __ldcs(val + static_cast<int64_t>(j + U + u) * 32) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (c[u] >= 0) acc = acc + v[u] * xv[u];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      v[u] = vn[u];
      c[u] = cn[u];
    }
  }
  return acc;
}

// Selects one of three kernel-parameter pointers without dynamic indexing of
// the parameter array (which would force a copy of it to local memory).
template <class T>
__device__ __forceinline__ T* pick3(T* const (&p)[3], int i) { return i == 0 ? p[0] : (i == 1 ? p[1] : p[2]); }

// ---- TMA (bulk asynchronous copy) staging of the value / index streams --------
// The SELL streams are contiguous per warp, so they do not need the LSU at all:
// one elected lane issues 1-D cp.async.bulk copies global -> shared (SASS UBLKCP)
// that complete on an mbarrier (SYNCS); only the gathers of x stay on the
// LSU -> L1 -> L2 request path, which is the unit that bounds the SpMV.
namespace tma {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      "  .reg .pred p;\n"
      "  mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "  selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ unsigned long long policy_evict_first() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// bytes: multiple of 16; dst / src 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar, unsigned long long policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
               : "memory");
}
}  // namespace tma

// The gathered vector is either fixed at launch or, inside the device-resident
// step loop, one of three buffers selected by the state's candidate index.
struct GatherSrc {
  const double* p[3];
  const StepState* st;
};

// The step decision runs at the HEAD of the last kernel of an attempt. All three sums it needs
// are known before K^T y' is even computed: ||dx||^2 from the primal kernel, and ||dy||^2 and the
// nonlinearity dx . K^T dy = (K dx) . dy from the dual kernel, because K dx = (K x~ - K x) / 2 is
// available row by row once K x of the current iterate is kept (StepPtrs::kx). So block 0 of the
// K^T y' kernel adds the partials and takes the decision while the other blocks already stream the
// matrix: the decision is off the critical path and an attempt is three launches. The state is
// double-buffered: the kernels of an attempt read slot `in`, the decision writes slot `out`, the
// next attempt reads that one.
struct DecideHead {
  const StepState* in = nullptr;  // nullptr: no decision in this launch
  StepState* out = nullptr;
  const double* pp = nullptr;     // ||dx||^2 partials of the primal-step kernel
  int np = 0;
  const double* pd = nullptr;     // {||dy||^2, (K dx) . dy} partials of the dual kernel, [block][2]
  int nd = 0;
  // Row-sharded solves: already reduced {||dx||^2, ||dy||^2, (K dx) . dy} triples, one per rank,
  // at scal[k * scal_stride + 0..2], added in rank order (replaces pd, and pp when scal_has_dx2).
  const double* scal = nullptr;
  int scal_count = 0, scal_stride = 0, scal_has_dx2 = 0;
};
template <int BT>
__device__ void run_decide_head(const DecideHead& h);

// Epi: struct Ctx; __device__ Ctx begin() const -- once per thread: resolves the rotating
//      buffers and step scalars from the device state;
//      struct Pre; __device__ Pre prefetch(const Ctx&, int64_t pos) const  -- issues the
//      epilogue's own loads before the gather loop so that they overlap it;
//      __device__ void operator()(const Ctx&, int64_t pos, double acc, double* red, const Pre&) const
// Work queue of a persistent launch (the step loop's SpMV pair): the grid is sized to what is
// resident at once, every block takes tiles (= what a block of an ordinary launch would do, same
// slot -> thread map, same per-tile partial sums: results do not depend on the schedule) from an
// atomic counter until none is left, so no SM idles through a last partial wave. The counter is
// zeroed by the OTHER kernel of the pair (which runs strictly before / after this one).
struct TileQueue {
  unsigned int* counter = nullptr;  // nullptr: ordinary launch, one tile per block
  unsigned int* reset = nullptr;    // the other kernel's counter, zeroed by this launch
  int num_tiles = 0;
};

template <int MODE, int NS, class Epi, int BT, int V>
__global__ void __launch_bounds__(BT, ((V == 2 ? 768 : V == 1 ? 1024 : V == 8 ? 896 : 1280) / BT)) k_sell(SellDev a, GatherSrc gs, Epi epi, double* partials, const int32_t* halt, int chunks,
                                                                                              DecideHead head, TileQueue q) {
  pdl_trigger();
  pdl_wait();
  if (q.reset != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *q.reset = 0u;
  if (halt != nullptr && *halt != 0) return;
  if (head.in != nullptr && blockIdx.x == 0) {
    // the launch has one extra block, the FIRST one scheduled, and it only takes the step decision
    // (see DecideHead): apart from the row loop, so that its registers do not count against the loop's
    run_decide_head<BT>(head);
    return;
  }
  const int64_t first = static_cast<int64_t>(blockIdx.x) - (head.in != nullptr ? 1 : 0);
  const int64_t workers = static_cast<int64_t>(gridDim.x) - (head.in != nullptr ? 1 : 0);
  const double* __restrict__ x = gs.st != nullptr ? pick3(gs.p, gs.st->cand) : gs.p[0];
  const typename Epi::Ctx ctx = epi.begin();
  __shared__ int64_t s_next;
  // `chunks` consecutive groups of BT slots per tile: the slot -> thread map is fixed, so the
  // partial sums stay deterministic.
  for (int64_t blk = first;;) {
    if (q.counter != nullptr && threadIdx.x == 0) s_next = workers + atomicAdd(q.counter, 1u);  // (used after this tile: latency hidden)
    double red[NS > 0 ? NS : 1];
#pragma unroll
    for (int k = 0; k < NS; ++k) red[k] = 0.0;
    for (int c = 0; c < chunks; ++c) {
      const int64_t slot = (blk * chunks + c) * BT + threadIdx.x;
      if (slot >= a.num_slots) break;
      const int64_t pos = a.num_split + (slot - a.num_virtual_padded);
      const bool own_row = slot >= a.num_virtual_padded && pos < a.num_rows;
      typename Epi::Pre pre;
      if (own_row) pre = epi.prefetch(ctx, pos);
      const double acc = sell_row<MODE, V>(a, slot, x);
      if (slot < a.num_virtual_padded) {
        a.virt_partial[slot] = acc;
      } else if (own_row) {
        epi(ctx, pos, acc, red, pre);
      }
    }
    if (NS > 0) block_reduce_store<NS, 0, BT>(red, nullptr, partials + blk * NS);
    if (q.counter == nullptr) break;
    __syncthreads();
    blk = s_next;
    __syncthreads();  // (every thread has read s_next before thread 0 overwrites it)
    if (blk >= q.num_tiles) break;
  }
}

// The same product with the value / index streams staged through shared memory
// by the TMA engine (1-D cp.async.bulk + mbarrier), NST stages of U slots per
// warp. A warp owns `chunks` CONSECUTIVE slices, so its part of val[] / col[] is
// one contiguous range that lane 0 streams in pieces of U x 32 elements, each
// piece completing on the stage's mbarrier; the lanes read their element of every
// slot from shared memory (conflict-free) and only the gathers of x use the
// LSU -> L2 request path. Slot -> thread map and reduction order are fixed, so
// results are run-to-run deterministic (they differ in the last bits from
// k_sell, whose blocks interleave the slices differently).
template <int MODE, int NS, class Epi, int BT, int U, int NST>
__global__ void __launch_bounds__(BT) k_sell_tma(SellDev a, GatherSrc gs, Epi epi, double* partials, const int32_t* halt, int chunks, DecideHead head) {
  constexpr int NW = BT / 32;
  __shared__ alignas(128) double s_val[NW][NST][U * 32];
  __shared__ alignas(128) int32_t s_col[NW][NST][U * 32];
  __shared__ alignas(8) unsigned long long s_bar[NW][NST];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();
  if (lane == 0) {
#pragma unroll
    for (int st = 0; st < NST; ++st) tma::mbar_init(&s_bar[warp][st], 1);
    tma::fence_barrier_init();
  }
  __syncwarp();
  pdl_wait();
  if (halt != nullptr && *halt != 0) return;  // (no copy has been issued yet)
  if (head.in != nullptr && blockIdx.x == 0) {
    run_decide_head<BT>(head);
    return;
  }
  const int64_t blk = static_cast<int64_t>(blockIdx.x) - (head.in != nullptr ? 1 : 0);
  const double* __restrict__ x = gs.st != nullptr ? pick3(gs.p, gs.st->cand) : gs.p[0];
  double red[NS > 0 ? NS : 1];
#pragma unroll
  for (int k = 0; k < NS; ++k) red[k] = 0.0;
  const typename Epi::Ctx ctx = epi.begin();
  const int64_t num_slices = a.num_slots >> 5;
  const int64_t slice0 = (blk * NW + warp) * chunks;
  const int nsl = static_cast<int>(max(static_cast<int64_t>(0), min(static_cast<int64_t>(chunks), num_slices - slice0)));
  if (nsl > 0) {
    const int64_t e0 = a.slice_ptr[slice0];
    const int R = static_cast<int>((a.slice_ptr[slice0 + nsl] - e0) >> 5);  // 32-element rows of this warp's stream
    const int Q = (R + U - 1) / U;                                         // pieces
    const double* __restrict__ gval = a.val + e0;
    const int32_t* __restrict__ gcol = a.col + e0;
    const unsigned long long policy = tma::policy_evict_first();
    auto issue = [&](int q) {  // lane 0 only
      const int rows = min(U, R - q * U);
      const int st = q % NST;
      tma::mbar_expect_tx(&s_bar[warp][st], static_cast<uint32_t>(rows) * 384u);
      tma::bulk_g2s(&s_val[warp][st][0], gval + static_cast<int64_t>(q) * (U * 32), static_cast<uint32_t>(rows) * 256u, &s_bar[warp][st], policy);
      tma::bulk_g2s(&s_col[warp][st][0], gcol + static_cast<int64_t>(q) * (U * 32), static_cast<uint32_t>(rows) * 128u, &s_bar[warp][st], policy);
    };
    if (lane == 0) {
      for (int q = 0; q < min(NST, Q); ++q) issue(q);
    }
    int r = 0;  // rows of the stream consumed so far
    int64_t e_prev = e0;
    for (int c = 0; c < nsl; ++c) {
      const int64_t slot = (slice0 + c) * 32 + lane;
      const int64_t e_next = a.slice_ptr[slice0 + c + 1];
      const int W = static_cast<int>((e_next - e_prev) >> 5);
      e_prev = e_next;
      const int n = a.slot_len[slot];
      const int64_t pos = a.num_split + (slot - a.num_virtual_padded);
      const bool own_row = slot >= a.num_virtual_padded && pos < a.num_rows;
      typename Epi::Pre pre;
      if (own_row) pre = epi.prefetch(ctx, pos);
      double acc = 0.0;
      int j = 0;
      while (j < W) {
        const int q = r / U, o = r % U;
        const int st = q % NST;
        if (o == 0) tma::mbar_wait(&s_bar[warp][st], static_cast<uint32_t>((q / NST) & 1));
        const int avail = min(U - o, W - j);
        const double* sv = &s_val[warp][st][o * 32 + lane];
        const int32_t* sc = &s_col[warp][st][o * 32 + lane];
        const int mine = min(avail, n - j);  // slots of this piece that belong to this lane's row
        double xv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) xv[u] = u < mine ? __ldg(x + sc[u * 32]) : 0.0;
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (u < mine) acc = combine<MODE>(acc, sv[u * 32], xv[u]);
        r += avail;
        j += avail;
        if ((r % U) == 0 || r == R) {  // piece q fully consumed: refill its stage
          __syncwarp();
          if (lane == 0 && q + NST < Q) {
            tma::fence_proxy_async();
            issue(q + NST);
          }
        }
      }
      if (slot < a.num_virtual_padded) {
        a.virt_partial[slot] = acc;
      } else if (own_row) {
        epi(ctx, pos, acc, red, pre);
      }
    }
  }
  if (NS > 0) block_reduce_store<NS, 0, BT>(red, nullptr, partials + blk * NS);
}

// Split rows: one warp per row combines the partials of its virtual slots in
// slot order and then runs the same epilogue.
template <int MODE, int NS, class Epi>
__global__ void __launch_bounds__(kThreads) k_sell_fixup(SellDev a, Epi epi, double* partials, const int32_t* halt) {
  pdl_trigger();
  pdl_wait();
  if (halt != nullptr && *halt != 0) return;
  double red[NS > 0 ? NS : 1];
#pragma unroll
  for (int k = 0; k < NS; ++k) red[k] = 0.0;
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row < a.num_split) {
    const int b = a.split_first[row], e = a.split_first[row + 1];
    double acc = 0.0;
    for (int k = b + lane; k < e; k += 32) acc = (MODE == kMaxAbs) ? fmax(acc, a.virt_partial[k]) : acc + a.virt_partial[k];
    acc = (MODE == kMaxAbs) ? warp_max(acc) : warp_sum(acc);
    if (lane == 0) {
      const typename Epi::Ctx ctx = epi.begin();
      epi(ctx, row, acc, red, epi.prefetch(ctx, row));
    }
  }
  if (NS > 0) block_reduce_store<NS, 0>(red, nullptr, partials + static_cast<int64_t>(blockIdx.x) * NS);
}

// --------------------------------------------------------- PDHG step -------
struct StepPtrs {
  int64_t n, m;
  double* x[3];
  double* y[3];
  double* kty[3];
  double* kx[3];   // K x of the iterates (dual length): K x' = (K x~ + K x) / 2 comes out of the dual kernel
  double* x_tilde;
  double* avg_x;
  double* avg_y;
  // Row-sharded peer solves only (else nullptr): K x and K^T y of the average, maintained with the
  // averages themselves (the products are linear in the iterate, and K x / K^T y of every iterate
  // exist) -- the restart test then needs no product of the average, i.e. no all-reduce of n doubles.
  double* avg_kx;
  double* avg_kty;
  const double *c, *q, *lv, *uv, *lc, *uc;
  StepState* state;  // the slot the kernels of this attempt read
};

// Peer-memory exchange of the row-sharded step loop: every rank's
// arena (comm.h PeerArena, mapped everywhere over NVLink) holds, in doubles,
//   [xt_off, +n)       x~ in the caller's column order; slice C_h is stored by rank h
//   [partial_off, +n)  this rank's (K[R_g,:])^T y' partial; peers pull their slice
//   [y_off, +m_global) y' in the box-wide dual order (block of rank h at its row_begin); the
//                      all-gather exchange stores it from the dual epilogue
//   [scal_off, +4 G)   {||dx||^2, ||dy||^2, dx.(K^T y' - K^T y)} partials of rank h at 4 h
//   [flags_off, +8 k)  barrier k: epoch last signalled by rank h at 8 k + h (u64)
//   [epoch_off, +4)    this rank's own epoch counters (u64)
//   [cand_off, +G seg) trust-region finish: rank h's bracket candidates (keys | a | b | count) at cand_off + h * PeerLayout::kTrCandSegment
//   [tr_off, +2*48*8)  trust-region round vector (kTrCols doubles) of rank h for round parity p at kTrCols (G p + h)
struct PeerPtrs {
  int world, rank;
  double* base[kMaxPeers];
  int64_t xt_off, partial_off, y_off, scal_off, flags_off, epoch_off, tr_off, cand_off;
  int tr_barrier;      // barrier index of the trust-region solve using these pointers (3, or 4 for the second of a concurrent pair)
  int64_t row_begin;   // this rank's first position in the box-wide dual order
  long long timeout_cycles;  // a peer that does not arrive within this many SM clocks halts the loop
  int64_t begin, end;  // this rank's slice of the primal vector (begin is even)
};
__device__ __forceinline__ double* peer_base(const PeerPtrs& pp, int h) {
  double* r = pp.base[0];
#pragma unroll
  for (int k = 1; k < kMaxPeers; ++k) r = h == k ? pp.base[k] : r;
  return r;
}

// Cross-GPU barrier `which` for the threads of ONE warp (lanes < world take
// part): lane h raises this rank's flag in rank h's arena, then waits for
// rank h's flag in the local arena. Epochs only grow and every rank runs the
// same kernel sequence, so a rank is never more than one epoch ahead. A peer
// that does not arrive within PDLP_B200_PEER_TIMEOUT_S (default 20 s) halts the
// loop with kHaltPeerTimeout instead of hanging the device.
__device__ __forceinline__ void peer_barrier(const PeerPtrs& pp, int which, int32_t* halt_flag) {
  const int lane = threadIdx.x & 31;
  double* local = peer_base(pp, pp.rank);
  unsigned long long* epoch = reinterpret_cast<unsigned long long*>(local + pp.epoch_off) + which;
  const unsigned long long e = *reinterpret_cast<volatile unsigned long long*>(epoch) + 1ull;
  __threadfence_system();
  if (lane < pp.world) {
    volatile unsigned long long* dst = reinterpret_cast<unsigned long long*>(peer_base(pp, lane) + pp.flags_off) + which * 8 + pp.rank;
    *dst = e;
    volatile unsigned long long* src = reinterpret_cast<unsigned long long*>(local + pp.flags_off) + which * 8 + lane;
    const long long t0 = clock64();
    while (*src < e) {
      if (clock64() - t0 > pp.timeout_cycles) {
        *halt_flag = kHaltPeerTimeout;
        break;
      }
    }
  }
  __syncwarp();
  __threadfence_system();
  if (lane == 0) *reinterpret_cast<volatile unsigned long long*>(epoch) = e;
}
__global__ void k_peer_barrier(PeerPtrs pp, int which, StepState* st) {
  pdl_trigger();
  pdl_wait();
  if (st->halt != 0) return;
  peer_barrier(pp, which, &st->halt);
}

// Primal half step, two elements per thread (128-bit loads/stores). PEER: this
// rank updates only its slice [begin, end) and stores x~ straight into every
// rank's arena (the all-gather of x~ fused into the producing kernel).
template <bool PEER>
__global__ void __launch_bounds__(kThreads) k_primal_step(StepPtrs b, PeerPtrs peer, double* partials) {
  pdl_trigger();
  pdl_wait();
  const StepState* st = b.state;
  if (st->halt != 0) return;
  const double* __restrict__ xc = pick3(b.x, st->cur);
  double* __restrict__ xn = pick3(b.x, st->cand);
  const double* __restrict__ kty = pick3(b.kty, st->cur);
  const double tau = st->step_size / st->primal_weight;
  const double ratio = st->pending_ratio;
  const bool has_q = b.q != nullptr;
  double s = 0.0;
  const int64_t i0 = (PEER ? peer.begin : 0) + (static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x) * 2;
  const int64_t iend = PEER ? peer.end : b.n;
  if (i0 + 1 < iend) {
    const double2 x2 = *reinterpret_cast<const double2*>(xc + i0);
    const double2 k2 = *reinterpret_cast<const double2*>(kty + i0);
    const double2 c2 = *reinterpret_cast<const double2*>(b.c + i0);
    const double2 l2 = *reinterpret_cast<const double2*>(b.lv + i0);
    const double2 u2 = *reinterpret_cast<const double2*>(b.uv + i0);
    double t0 = x2.x - tau * (c2.x - k2.x), t1 = x2.y - tau * (c2.y - k2.y);
    if (has_q) {
      const double2 q2 = *reinterpret_cast<const double2*>(b.q + i0);
      t0 = t0 / (tau * q2.x + 1.0);
      t1 = t1 / (tau * q2.y + 1.0);
    }
    double2 nx;
    nx.x = fmax(fmin(t0, u2.x), l2.x);
    nx.y = fmax(fmin(t1, u2.y), l2.y);
    const double d0 = nx.x - x2.x, d1 = nx.y - x2.y;
    *reinterpret_cast<double2*>(xn + i0) = nx;
    double2 xt;
    xt.x = nx.x + d0;
    xt.y = nx.y + d1;
    if (PEER) {
#pragma unroll
      for (int h = 0; h < kMaxPeers; ++h)
        if (h < peer.world) *reinterpret_cast<double2*>(peer.base[h] + peer.xt_off + i0) = xt;
    } else {
      *reinterpret_cast<double2*>(b.x_tilde + i0) = xt;
    }
    s = d0 * d0 + d1 * d1;
    if (ratio > 0.0) {
      double2 av = *reinterpret_cast<const double2*>(b.avg_x + i0);
      av.x += ratio * (x2.x - av.x);
      av.y += ratio * (x2.y - av.y);
      *reinterpret_cast<double2*>(b.avg_x + i0) = av;
      if (PEER && b.avg_kty != nullptr) {
        double2 ak = *reinterpret_cast<const double2*>(b.avg_kty + i0);
        ak.x += ratio * (k2.x - ak.x);
        ak.y += ratio * (k2.y - ak.y);
        *reinterpret_cast<double2*>(b.avg_kty + i0) = ak;
      }
    }
  } else if (i0 < iend) {
    const double x = xc[i0];
    double t = x - tau * (b.c[i0] - kty[i0]);
    if (has_q) t = t / (tau * b.q[i0] + 1.0);
    const double nx = fmax(fmin(t, b.uv[i0]), b.lv[i0]);
    const double d = nx - x;
    xn[i0] = nx;
    if (PEER) {
#pragma unroll
      for (int h = 0; h < kMaxPeers; ++h)
        if (h < peer.world) peer.base[h][peer.xt_off + i0] = nx + d;
    } else {
      b.x_tilde[i0] = nx + d;
    }
    s = d * d;
    if (ratio > 0.0) {
      b.avg_x[i0] += ratio * (x - b.avg_x[i0]);
      if (PEER && b.avg_kty != nullptr) b.avg_kty[i0] += ratio * (kty[i0] - b.avg_kty[i0]);
    }
  }
  block_reduce_store<1, 0>(&s, nullptr, partials + blockIdx.x);
  if (blockIdx.x == 0 && threadIdx.x == 0 && st->rule == PDLP_ADAPTIVE_LINESEARCH_RULE) {
    // the two pow() of this attempt's step-size decision, while the SpMV pair runs (see StepState)
    const double total = static_cast<double>(st->num_rejected_steps + st->inner_iterations + st->iterations_completed + 1);
    b.state->pow_reduction = pow(total + 1.0, -st->reduction_exponent);
    b.state->pow_growth = pow(total + 1.0, -st->growth_exponent);
    b.state->pow_total = total;
  }
}

// PUSH: the all-gather of y' fused into the producing epilogue -- every new
// dual value is also stored into every rank's arena at its box-wide position.
// AVGK: also maintains K x of the average (StepPtrs::avg_kx; peer exchange only). CG: the row-wise vectors
// may have been written by another SM earlier in the same launch (k_peer_loop hands slices out
// dynamically): they are read at L2 (ld.global.cg), never through this SM's L1.
template <bool CG>
__device__ __forceinline__ double row_ld(const double* p) { return CG ? __ldcg(p) : *p; }
template <bool PUSH, bool AVGK = false, bool CG = false>
struct DualEpiT {  // pdhg.cc:1912-1930 with theta = 1; nonlinearity of pdhg.cc:2588-2592 on the row side
  StepPtrs b;
  PeerPtrs peer;
  struct Ctx { const double* yc; double* yn; const double* kxc; double* kxn; double sigma, ratio; };
  struct Pre { double yc, lc, uc, avg; };
  __device__ __forceinline__ Ctx begin() const {
    const StepState* st = b.state;
    Ctx c;
    c.yc = pick3(b.y, st->cur);
    c.yn = pick3(b.y, st->cand);
    c.kxc = pick3(b.kx, st->cur);
    c.kxn = pick3(b.kx, st->cand);
    c.sigma = st->step_size * st->primal_weight;
    c.ratio = st->pending_ratio;
    return c;
  }
  __device__ __forceinline__ Pre prefetch(const Ctx& c, int64_t pos) const {
    Pre p;
    p.yc = row_ld<CG>(c.yc + pos);
    p.lc = __ldg(b.lc + pos);
    p.uc = __ldg(b.uc + pos);
    p.avg = c.ratio > 0.0 ? row_ld<CG>(b.avg_y + pos) : 0.0;
    return p;
  }
  // kxt = (K x~)_pos with x~ = 2 x' - x, so K x' = (kxt + K x) / 2 and K (x' - x) = (kxt - K x) / 2
  __device__ __forceinline__ void operator()(const Ctx& c, int64_t pos, double kxt, double* red, const Pre& p) const {
    const double kxc = row_ld<CG>(c.kxc + pos);  // (loaded here, not prefetched: the row loop is at its register limit)
    if (c.ratio > 0.0) {
      b.avg_y[pos] = p.avg + c.ratio * (p.yc - p.avg);
      if (AVGK && b.avg_kx != nullptr) {
        const double ak = row_ld<CG>(b.avg_kx + pos);
        b.avg_kx[pos] = ak + c.ratio * (kxc - ak);
      }
    }
    const double t = p.yc - c.sigma * kxt;
    const double yn = fmax(fmin(0.0, t + c.sigma * p.uc), t + c.sigma * p.lc);
    c.yn[pos] = yn;
    c.kxn[pos] = 0.5 * (kxt + kxc);
    if (PUSH) {
#pragma unroll
      for (int h = 0; h < kMaxPeers; ++h)
        if (h < peer.world) peer.base[h][peer.y_off + peer.row_begin + pos] = yn;
    }
    const double d = yn - p.yc;
    red[0] += d * d;
    red[1] += (0.5 * (kxt - kxc)) * d;  // (K dx)_pos dy_pos: summed over the rows it is dx . K^T dy
  }
};
struct DualEpi : DualEpiT<false> {};
inline DualEpi MakeDualEpi(const StepPtrs& p) {
  DualEpi e;
  e.b = p;
  std::memset(&e.peer, 0, sizeof(e.peer));
  return e;
}

struct KtyEpi {  // K^T y' of the candidate (pdhg.cc:1949-1959); the nonlinearity is computed on the row side
  StepPtrs b;
  struct Ctx { double* kty_cand; };
  struct Pre {};
  __device__ __forceinline__ Ctx begin() const { return Ctx{pick3(b.kty, b.state->cand)}; }
  __device__ __forceinline__ Pre prefetch(const Ctx&, int64_t) const { return Pre(); }
  __device__ __forceinline__ void operator()(const Ctx& c, int64_t pos, double kty_next, double*, const Pre&) const { c.kty_cand[pos] = kty_next; }
};

// All-gather exchange: K^T y' for this rank's column slice only. Position p of the
// slice image is column begin + perm[p] of the (replicated, column-ordered) primal side.
struct KtyEpiSlice {
  StepPtrs b;
  const int32_t* perm;
  int64_t col0;
  struct Ctx { double* kty_cand; };
  struct Pre { int32_t idx; };
  __device__ __forceinline__ Ctx begin() const { return Ctx{pick3(b.kty, b.state->cand) + col0}; }
  __device__ __forceinline__ Pre prefetch(const Ctx&, int64_t pos) const { return Pre{__ldg(perm + pos)}; }
  __device__ __forceinline__ void operator()(const Ctx& c, int64_t, double kty_next, double*, const Pre& p) const { c.kty_cand[p.idx] = kty_next; }
};

constexpr int kDecideThreads = 1024;
// Fixed-order sum of p[0..count): thread t adds p[t], p[t+1024], ... (independent
// loads, unrolled so they are all in flight together), then a shuffle tree.
__device__ __forceinline__ double block_sum_range(const double* __restrict__ p, int count) {
  __shared__ double sh[kDecideThreads / 32];
  double s = 0.0;
#pragma unroll 8
  for (int i = threadIdx.x; i < count; i += kDecideThreads) s += p[i];
  s = warp_sum(s);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < 32) {
    t = threadIdx.x < kDecideThreads / 32 ? sh[threadIdx.x] : 0.0;
    t = warp_sum(t);
  }
  return t;  // valid in warp 0
}

// Accept test and step-size update (pdhg.cc:2574-2640, 2651-2674) from the three
// reduced scalars; one thread. Reads the state slot of this attempt, writes the other one.
// The two powers of the adaptive rule depend on the attempt count only, so a caller may
// compute them (and load the state) while the partial sums are still in flight.
__device__ __forceinline__ double decide_total(const StepState& s) {
  return static_cast<double>(s.num_rejected_steps + s.inner_iterations + s.iterations_completed + 1);
}
__device__ __forceinline__ void decide_powers(const StepState& s, double& pow_reduction, double& pow_growth) {
  pow_reduction = pow_growth = 0.0;
  if (s.rule != PDLP_ADAPTIVE_LINESEARCH_RULE) return;
  const double total = decide_total(s);
  if (s.pow_total == total) {  // left by the primal-step kernel of this attempt
    pow_reduction = s.pow_reduction;
    pow_growth = s.pow_growth;
  } else {
    pow_reduction = pow(total + 1.0, -s.reduction_exponent);
    pow_growth = pow(total + 1.0, -s.growth_exponent);
  }
}
// (One load of the whole state into registers by the caller, one store at the end: the decision
// is a chain of dependent scalar updates and must not pay an L2 round trip per field.)
__device__ void decide_update(StepState* st_dev, StepState s, double pow_reduction, double pow_growth, double dx2, double dy2, double dot) {
  StepState* st = &s;
  const double eta = st->step_size, omega = st->primal_weight;
  const double movement = (0.5 * omega * dx2) + (0.5 / omega) * dy2;
  const double nonlinearity = -dot;
  st->last_dx2 = dx2;
  st->last_dy2 = dy2;
  st->last_nonlinearity = nonlinearity;
  st->last_movement = movement;
  st->attempts += 1;
  const bool adaptive = st->rule == PDLP_ADAPTIVE_LINESEARCH_RULE;
  if (movement == 0.0 || movement > 1.0e100) {
    // kForceNumericalTermination: the loop is left without incrementing
    // inner_iterations, then num_rejected_steps_ += inner_iterations - 1.
    st->halt = movement == 0.0 ? kHaltZeroMovement : kHaltDivergent;
    st->pending_ratio = st->pending_ratio_dual = st->pending_ratio0 = 0.0;
    if (adaptive) st->num_rejected_steps += st->inner_iterations - 1;
    st->inner_iterations = 0;
    *st_dev = s;
    return;
  }
  bool accepted = true;
  double new_eta = eta;
  if (adaptive) {
    const double limit = nonlinearity > 0 ? movement / nonlinearity : kInfD;
    accepted = eta <= limit;
    const double first = isinf(limit) ? limit : (1.0 - pow_reduction) * limit;
    const double second = (1.0 + pow_growth) * eta;
    new_eta = fmin(first, second);
  }
  if (accepted) {
    const int old_prev = st->prev;
    st->prev = st->cur;
    st->cur = st->cand;
    st->cand = old_prev;
    // ShardedWeightedAverage::Add(next, weight = step size used), deferred.
    if (eta > 0.0) {
      st->pending_ratio = eta / (st->avg_weight_sum + eta);
      st->avg_weight_sum += eta;
    } else {
      st->pending_ratio = 0.0;
    }
    st->pending_ratio_dual = st->pending_ratio;
    st->pending_ratio0 = 0.0;
    st->avg_num_terms += 1;
    st->num_rejected_steps += st->inner_iterations;
    st->inner_iterations = 0;
    st->iterations_completed += 1;
    st->step_size = new_eta;
    const double kkt = static_cast<double>(st->iterations_completed) + static_cast<double>(st->num_rejected_steps);
    if (st->iterations_completed >= st->k_stop || kkt >= st->kkt_pass_limit) st->halt = kHaltCheckpoint;
  } else {
    st->pending_ratio = st->pending_ratio_dual = st->pending_ratio0 = 0.0;
    st->step_size = new_eta;
    st->inner_iterations += 1;
    if (st->inner_iterations >= 60) {
      st->halt = kHaltInnerLimit;
      st->num_rejected_steps += st->inner_iterations - 1;
      st->inner_iterations = 0;
    }
  }
  *st_dev = s;
}

// Fixed-order sums of the step partials by one block of BT threads: out = {sum pp[0..np),
// sum pd[2 b], sum pd[2 b + 1]}, valid in thread 0. Thread t adds elements t, t + BT, ...; then a
// shuffle tree and a fixed-order sum over the warps: the order depends on the sizes only.
template <int BT, bool VOLATILE_L2>
__device__ __forceinline__ void step_partial_sums(const double* pp, int np, const double* pd, int nd, double out[3]) {
  __shared__ double sh3[3][BT / 32];
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll 8
  for (int i = threadIdx.x; i < np; i += BT) s0 += VOLATILE_L2 ? __ldcg(pp + i) : pp[i];
#pragma unroll 8
  for (int i = threadIdx.x; i < nd; i += BT) {
    const double2 v = VOLATILE_L2 ? __ldcg(reinterpret_cast<const double2*>(pd) + i) : reinterpret_cast<const double2*>(pd)[i];
    s1 += v.x;
    s2 += v.y;
  }
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if ((threadIdx.x & 31) == 0) {
    sh3[0][threadIdx.x >> 5] = s0;
    sh3[1][threadIdx.x >> 5] = s1;
    sh3[2][threadIdx.x >> 5] = s2;
  }
  __syncthreads();
  out[0] = out[1] = out[2] = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k)
      for (int w = 0; w < BT / 32; ++w) out[k] += sh3[k][w];
  }
}

// Head of the last kernel of an attempt (see DecideHead): all threads of block 0 call it, after
// pdl_wait. The state is loaded first and used last, so its round trip overlaps the partial loads.
template <int BT>
__device__ void run_decide_head(const DecideHead& h) {
  StepState loaded;
  if (threadIdx.x == 0) loaded = *h.in;
  double sums[3];
  step_partial_sums<BT, false>(h.scal != nullptr && h.scal_has_dx2 ? nullptr : h.pp, h.scal != nullptr && h.scal_has_dx2 ? 0 : h.np,
                               h.scal != nullptr ? nullptr : h.pd, h.scal != nullptr ? 0 : h.nd, sums);
  if (threadIdx.x != 0) return;
  if (h.scal != nullptr) {
    const volatile double* sc = h.scal;
    double t0 = 0.0, t1 = 0.0, t2 = 0.0;
    for (int k = 0; k < h.scal_count; ++k) {
      if (h.scal_has_dx2) t0 += sc[k * h.scal_stride + 0];
      t1 += sc[k * h.scal_stride + 1];
      t2 += sc[k * h.scal_stride + 2];
    }
    if (h.scal_has_dx2) sums[0] = t0;
    sums[1] = t1;
    sums[2] = t2;
  }
  double pr, pg;
  decide_powers(loaded, pr, pg);
  decide_update(h.out, loaded, pr, pg, sums[0], sums[1], sums[2]);
}

// The decision as a kernel of its own: problems without variables (no K^T y' kernel to carry it).
__global__ void __launch_bounds__(kThreads) k_decide_only(DecideHead head) {
  pdl_trigger();
  pdl_wait();
  if (head.in->halt != 0) return;
  run_decide_head<kThreads>(head);
}

// ---- row-sharded variants of the step ----------------------------
// NCCL exchange: the primal side is replicated, the dual side is this rank's row block. The local
// {||dy||^2, (K dx) . dy} go to exchange[n], exchange[n + 1]; the local K^T y' partial (scattered to
// column order) to exchange[0..n); one all-reduce of n + 2 doubles completes all of them; then
// k_kty_finish stores K^T y' and its block 0 takes the decision.
__global__ void __launch_bounds__(kDecideThreads) k_sum_to_slot(const StepState* st, const double* pd, int nd, double* slot) {
  pdl_trigger();
  pdl_wait();
  if (st->halt != 0) return;
  double sums[3];
  step_partial_sums<kDecideThreads, false>(nullptr, 0, pd, nd, sums);
  if (threadIdx.x == 0) {
    slot[0] = sums[1];
    slot[1] = sums[2];
  }
}
__global__ void __launch_bounds__(kThreads) k_kty_finish(StepPtrs b, const double* __restrict__ reduced, DecideHead head) {
  pdl_trigger();
  pdl_wait();
  const StepState* st = b.state;
  if (st->halt != 0) return;
  if (blockIdx.x == 0) run_decide_head<kThreads>(head);
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
  if (i < b.n) pick3(b.kty, st->cand)[i] = reduced[i];
}

// Peer exchange: the reduce-scatter of the K^T y' partials fused into the
// consumer. This rank pulls its slice of every rank's partial out of the peer
// arenas (128-bit loads over NVLink), adds them in rank order (the same order
// on every rank) and stores K^T y'; block 0 takes the decision first.
__global__ void __launch_bounds__(kThreads) k_kty_finish_peer(StepPtrs b, PeerPtrs peer, DecideHead head) {
  pdl_trigger();
  pdl_wait();
  const StepState* st = b.state;
  if (st->halt != 0) return;
  if (blockIdx.x == 0) run_decide_head<kThreads>(head);
  const int64_t i0 = peer.begin + (static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x) * 2;
  double* __restrict__ kc = pick3(b.kty, st->cand);
  if (i0 + 1 < peer.end) {
    double2 v = make_double2(0.0, 0.0);
#pragma unroll
    for (int h = 0; h < kMaxPeers; ++h) {
      if (h < peer.world) {
        const double2 t = *reinterpret_cast<const double2*>(peer.base[h] + peer.partial_off + i0);
        v.x += t.x;
        v.y += t.y;
      }
    }
    *reinterpret_cast<double2*>(kc + i0) = v;
  } else if (i0 < peer.end) {
    double v = 0.0;
#pragma unroll
    for (int h = 0; h < kMaxPeers; ++h)
      if (h < peer.world) v += peer.base[h][peer.partial_off + i0];
    kc[i0] = v;
  }
}

// Peer exchange, between the dual kernel and the K^T y' side: this rank's fixed-order sums
// {||dx||^2, ||dy||^2, (K dx) . dy} are stored into every rank's arena and the ranks meet at barrier
// `which` -- the same barrier that publishes y' (all-gather exchange) or the K^T y' partials
// (reduce-scatter exchange). Afterwards every rank's decision head adds the G triples in rank
// order and takes the identical decision: no barrier of its own for the decision.
__global__ void __launch_bounds__(kDecideThreads) k_sum_push_barrier(StepState* st, PeerPtrs peer, int which, const double* pp, int np, const double* pd, int nd) {
  pdl_trigger();
  pdl_wait();
  if (st->halt != 0) return;
  double sums[3];
  step_partial_sums<kDecideThreads, false>(pp, np, pd, nd, sums);
  __shared__ double sh[3];
  if (threadIdx.x == 0) { sh[0] = sums[0]; sh[1] = sums[1]; sh[2] = sums[2]; }
  __syncthreads();
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  if (lane < peer.world) {
    volatile double* dst = peer_base(peer, lane) + peer.scal_off + 4 * peer.rank;
    dst[0] = sh[0];
    dst[1] = sh[1];
    dst[2] = sh[2];
  }
  peer_barrier(peer, which, &st->halt);
}

// [pbegin, pend): the part of the primal average this rank maintains (all of
// it unless the peer exchange slices the primal update).
__global__ void __launch_bounds__(kThreads) k_flush_average(StepPtrs b, int64_t total, int64_t pbegin, int64_t pend) {
  StepState* st = b.state;
  const double r0 = st->pending_ratio0, r1 = st->pending_ratio, rd = st->pending_ratio_dual;
  if (!(r0 > 0.0 || r1 > 0.0 || rd > 0.0)) return;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
  if (i < b.n) {
    if (i >= pbegin && i < pend) {
      double av = b.avg_x[i];
      if (r0 > 0.0) av += r0 * (pick3(b.x, st->prev)[i] - av);
      if (r1 > 0.0) av += r1 * (pick3(b.x, st->cur)[i] - av);
      b.avg_x[i] = av;
      if (b.avg_kty != nullptr && r1 > 0.0) b.avg_kty[i] += r1 * (pick3(b.kty, st->cur)[i] - b.avg_kty[i]);  // (r0 is Malitsky-Pock only: not maintained there)
    }
  } else if (i < total) {
    const int64_t j = i - b.n;
    if (rd > 0.0) {
      b.avg_y[j] += rd * (pick3(b.y, st->cur)[j] - b.avg_y[j]);
      if (b.avg_kx != nullptr) b.avg_kx[j] += rd * (pick3(b.kx, st->cur)[j] - b.avg_kx[j]);
    }
  }
}
__global__ void k_clear_pending(StepState* st) { st->pending_ratio = st->pending_ratio_dual = st->pending_ratio0 = 0.0; }

// ---- the row-sharded step loop as ONE persistent cooperative launch ----------
// At 1/8 of a problem per rank the five launches and two barrier kernels of an attempt cost more
// than the work in them (C2 on 8 GPUs: K x~ 31 us for 12 us of work). k_peer_loop runs a whole
// chunk of attempts in one launch, one block per SM: the phases of an attempt are separated by
// grid barriers (a ticket; the block that arrives last also meets the other ranks -- barrier A /
// B of DESIGN.md 5 -- before it releases the grid), every rank takes the identical decision from
// the G triples, and the loop leaves when the decision sets `halt` (checkpoint reached, numerical
// halt) -- a rejected step costs no host round trip.
//   P   primal slice step, x~ stored into every arena, ||dx||^2 per block       | grid + peer barrier A
//   D   K[R_g,:] x~ + dual update (+ y' stored into every arena, MODE 0)         | MODE 1: grid barrier
//   T1  (MODE 1, reduce-scatter) (K[R_g,:])^T y' partial into the own arena
//       the last block adds the per-block {||dx||^2, ||dy||^2, (K dx).dy} in block order and
//       stores the triple into every arena                                      | grid + peer barrier B
//   T   MODE 0: (K[:,C_g])^T y' from the all-gathered y'; MODE 1: pull slice C_g of every rank's
//       partial (128-bit peer loads), add in rank order. One warp first takes the step decision
//       and writes the other state slot.                                        | grid barrier
// The elementwise phases (P, the pull of T) are mapped to threads statically. The slices of the
// row loops (D, T1, T of MODE 0) are handed to warps DYNAMICALLY, one atomic ticket per slice
// (fetched one slice ahead): the SMs of a B200 do not run a request-bound row loop at the same
// speed (GPCs of 16 / 18 / 20 SMs share a crossbar port; a static map left the fastest SM idle
// for 25 % of a phase). Determinism does not depend on the schedule: every slice leaves its
// {||dy||^2, (K dx).dy} in its own slot, the slots are added per block over fixed ranges after the
// next grid barrier, the blocks in block order. Data another SM or GPU produced inside the launch
// (x~, y', K^T y, the row-wise vectors, the partials, the state) is read with ld.global.cg /
// volatile loads only: the L1 of an SM is not coherent across the phases of one kernel.
struct PeerLoopArgs {
  StepPtrs p;        // p.state: slot 0 of the two state slots
  PeerPtrs peer;
  SellDev rows;      // K[R_g,:]
  SellDev cols;      // MODE 0: (K[:,C_g])^T, indices in the box-wide dual order; MODE 1: (K[R_g,:])^T
  const int32_t* perm;   // MODE 0: column (relative to col0) of a position of the slice image; MODE 1: column of a position
  int64_t col0;
  double* block_partials;  // [blocks][4]: ||dx||^2, ||dy||^2, (K dx).dy
  double* slice_partials;  // [slices of rows][2]: ||dy||^2, (K dx).dy of a slice
  unsigned int* sync;      // [0] arrivals [1] generation [2] error (a peer never arrived) [4..6] slice tickets of D / T1 / T; zero before the launch
  int first_slot, max_attempts;
  unsigned long long* trace;  // nullptr, or [0] attempts, [1..6] summed ns of the phases (block 0)
};
enum { kLoopSyncPlain = 0, kLoopSyncPeer = 1, kLoopSyncSumsPeer = 2 };

// Grid barrier number `bar` of the launch (all threads of all blocks). The block that arrives last
// optionally (kLoopSyncSumsPeer) adds the partials in a fixed order -- the per-block ones in block
// order and, when nslices > 0, the per-slice ones of the dual phase in slice order -- and stores
// the triple into every arena, then meets the other ranks at peer barrier `which` before it
// releases the grid. Returns true when a peer never arrived.
template <int BT>
__device__ __forceinline__ bool loop_grid_sync(const PeerLoopArgs& g, unsigned& bar, int kind, int which, int64_t nslices, int* s_flags, double* s_tot) {
  // s_flags: [0] error [1] this block arrived last
  __syncthreads();
  if (threadIdx.x == 0) {
    // (cumulative: the block's stores -- local and into the peers' arenas -- ordered by the barrier above)
    if (kind == kLoopSyncPlain) __threadfence(); else __threadfence_system();
    const unsigned ticket = atomicAdd(g.sync, 1u);
    s_flags[1] = ticket == gridDim.x * (bar + 1u) - 1u ? 1 : 0;
  }
  __syncthreads();
  if (s_flags[1] != 0) {
    __threadfence();
    if (kind == kLoopSyncSumsPeer) {
      double r[3] = {0.0, 0.0, 0.0};
      for (int k = threadIdx.x; k < static_cast<int>(gridDim.x); k += BT) {
        r[0] += __ldcg(g.block_partials + 4 * k + 0);
        if (nslices == 0) {
          r[1] += __ldcg(g.block_partials + 4 * k + 1);
          r[2] += __ldcg(g.block_partials + 4 * k + 2);
        }
      }
      for (int64_t k = threadIdx.x; k < nslices; k += BT) {
        const double2 v = __ldcg(reinterpret_cast<const double2*>(g.slice_partials) + k);
        r[1] += v.x;
        r[2] += v.y;
      }
      block_reduce_store<3, 0, BT>(r, nullptr, s_tot);
      __syncthreads();
    }
    if (threadIdx.x < 32) {
      const int lane = threadIdx.x;
      if (kind == kLoopSyncSumsPeer && lane < g.peer.world) {
        volatile double* dst = peer_base(g.peer, lane) + g.peer.scal_off + 4 * g.peer.rank;
        dst[0] = s_tot[0];
        dst[1] = s_tot[1];
        dst[2] = s_tot[2];
      }
      if (kind != kLoopSyncPlain) peer_barrier(g.peer, which, reinterpret_cast<int32_t*>(g.sync + 2));
      if (lane == 0) {
        __threadfence();
        *reinterpret_cast<volatile unsigned int*>(g.sync + 1) = bar + 1u;
      }
    }
  } else if (threadIdx.x == 0) {
    const long long t0 = clock64();
    while (*reinterpret_cast<volatile unsigned int*>(g.sync + 1) < bar + 1u) {
      if (clock64() - t0 > 2 * g.peer.timeout_cycles) {  // (the releasing block itself gives up after timeout_cycles)
        *reinterpret_cast<volatile int*>(g.sync + 2) = kHaltPeerTimeout;
        break;
      }
    }
    __threadfence();
  }
  if (threadIdx.x == 0) s_flags[0] = *reinterpret_cast<volatile int*>(g.sync + 2);
  __syncthreads();
  ++bar;
  return s_flags[0] != 0;
}

// The step decision of k_peer_loop (one thread; run_decide_head of the multi-launch path): the G
// triples in rank order, the same on every rank. Not inlined: its registers (two pow()) must not
// count against the row loops.
__device__ __noinline__ void loop_decide(StepState* st_out, const StepState* st_smem, const double* scal, int world) {
  const volatile double* sc = scal;
  double dx2 = 0.0, dy2 = 0.0, dot = 0.0;
  for (int k = 0; k < world; ++k) {
    dx2 += sc[4 * k + 0];
    dy2 += sc[4 * k + 1];
    dot += sc[4 * k + 2];
  }
  StepState loaded = *st_smem;
  loaded.pow_total = -1.0;  // (nobody precomputed the powers of this attempt)
  double pr, pg;
  decide_powers(loaded, pr, pg);
  decide_update(st_out, loaded, pr, pg, dx2, dy2, dot);
}

__device__ __forceinline__ unsigned long long loop_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

template <int MODE, int BT>
__global__ void __launch_bounds__(BT, 1) k_peer_loop(PeerLoopArgs g) {
  constexpr int kWords = static_cast<int>(sizeof(StepState) / sizeof(double));
  static_assert(sizeof(StepState) % sizeof(double) == 0 && kWords <= BT, "the state is copied as doubles");
  __shared__ StepState s_st;
  __shared__ int s_flags[2];
  __shared__ double s_tot[3];
  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t nthreads = static_cast<int64_t>(gridDim.x) * BT;
  const int64_t gthread = static_cast<int64_t>(blockIdx.x) * BT + tid;
  const int64_t nwarps = nthreads >> 5, gwarp = gthread >> 5;
  const StepPtrs& b = g.p;
  const PeerPtrs& peer = g.peer;
  const bool has_q = b.q != nullptr;
  const bool tracing = g.trace != nullptr && blockIdx.x == 0 && tid == 0;
  const bool decider = gwarp == nwarps - 1;
  unsigned bar = 0;
  if (tid == 0) s_flags[0] = 0;
  for (int it = 0; it < g.max_attempts; ++it) {
    StepState* st_in = b.state + ((g.first_slot + it) & 1);
    StepState* st_out = b.state + ((g.first_slot + it + 1) & 1);
    __syncthreads();  // (everybody is done with the previous attempt's copy)
    if (tid < kWords) reinterpret_cast<double*>(&s_st)[tid] = __ldcg(reinterpret_cast<const double*>(st_in) + tid);
    __syncthreads();
    if (s_st.halt != 0) break;  // (the same state on every block and rank)
    const int cur = s_st.cur, cand = s_st.cand;
    const double ratio = s_st.pending_ratio;
    unsigned long long t0 = 0;
    if (tracing) t0 = loop_now();

    // ---- P: primal slice step (k_primal_step<PEER>) ------------------------------------
    {
      const double* __restrict__ xc = pick3(b.x, cur);
      double* __restrict__ xn = pick3(b.x, cand);
      const double* kty = pick3(b.kty, cur);
      const double tau = s_st.step_size / s_st.primal_weight;
      double s = 0.0;
      for (int64_t i0 = peer.begin + 2 * gthread; i0 < peer.end; i0 += 2 * nthreads) {
        if (i0 + 1 < peer.end) {
          const double2 x2 = *reinterpret_cast<const double2*>(xc + i0);
          const double2 k2 = __ldcg(reinterpret_cast<const double2*>(kty + i0));
          const double2 c2 = *reinterpret_cast<const double2*>(b.c + i0);
          const double2 l2 = *reinterpret_cast<const double2*>(b.lv + i0);
          const double2 u2 = *reinterpret_cast<const double2*>(b.uv + i0);
          double t0v = x2.x - tau * (c2.x - k2.x), t1v = x2.y - tau * (c2.y - k2.y);
          if (has_q) {
            const double2 q2 = *reinterpret_cast<const double2*>(b.q + i0);
            t0v = t0v / (tau * q2.x + 1.0);
            t1v = t1v / (tau * q2.y + 1.0);
          }
          double2 nx;
          nx.x = fmax(fmin(t0v, u2.x), l2.x);
          nx.y = fmax(fmin(t1v, u2.y), l2.y);
          const double d0 = nx.x - x2.x, d1 = nx.y - x2.y;
          *reinterpret_cast<double2*>(xn + i0) = nx;
          double2 xt;
          xt.x = nx.x + d0;
          xt.y = nx.y + d1;
#pragma unroll
          for (int h = 0; h < kMaxPeers; ++h)
            if (h < peer.world) *reinterpret_cast<double2*>(peer.base[h] + peer.xt_off + i0) = xt;
          s += d0 * d0 + d1 * d1;
          if (ratio > 0.0) {
            double2 av = *reinterpret_cast<const double2*>(b.avg_x + i0);
            av.x += ratio * (x2.x - av.x);
            av.y += ratio * (x2.y - av.y);
            *reinterpret_cast<double2*>(b.avg_x + i0) = av;
            if (b.avg_kty != nullptr) {
              double2 ak = *reinterpret_cast<const double2*>(b.avg_kty + i0);
              ak.x += ratio * (k2.x - ak.x);
              ak.y += ratio * (k2.y - ak.y);
              *reinterpret_cast<double2*>(b.avg_kty + i0) = ak;
            }
          }
        } else {
          const double x = xc[i0];
          const double kv = __ldcg(kty + i0);
          double t = x - tau * (b.c[i0] - kv);
          if (has_q) t = t / (tau * b.q[i0] + 1.0);
          const double nx = fmax(fmin(t, b.uv[i0]), b.lv[i0]);
          const double d = nx - x;
          xn[i0] = nx;
#pragma unroll
          for (int h = 0; h < kMaxPeers; ++h)
            if (h < peer.world) peer.base[h][peer.xt_off + i0] = nx + d;
          s += d * d;
          if (ratio > 0.0) {
            b.avg_x[i0] += ratio * (x - b.avg_x[i0]);
            if (b.avg_kty != nullptr) b.avg_kty[i0] += ratio * (kv - b.avg_kty[i0]);
          }
        }
      }
      block_reduce_store<1, 0, BT>(&s, nullptr, g.block_partials + 4 * blockIdx.x);
    }
    unsigned long long t1 = 0;
    if (tracing) t1 = loop_now();
    if (loop_grid_sync<BT>(g, bar, kLoopSyncPeer, 0, 0, s_flags, s_tot)) break;
    unsigned long long t2 = 0;
    if (tracing) t2 = loop_now();

    // ---- D: K[R_g,:] x~ and the dual update (k_sell + DualEpi) --------------------------
    if (blockIdx.x == 0 && tid == 0) { g.sync[5] = 0u; g.sync[6] = 0u; }  // (tickets of T1 / T: nobody draws them before the next grid barrier)
    {
      typedef DualEpiT<MODE == 0, true, true> Epi;
      Epi epi;
      epi.b = b;
      epi.peer = peer;
      typename Epi::Ctx ctx;
      ctx.yc = pick3(b.y, cur);
      ctx.yn = pick3(b.y, cand);
      ctx.kxc = pick3(b.kx, cur);
      ctx.kxn = pick3(b.kx, cand);
      ctx.sigma = s_st.step_size * s_st.primal_weight;
      ctx.ratio = ratio;
      const double* xt = peer.base[peer.rank] + peer.xt_off;
      const SellDev& a = g.rows;
      const unsigned nsl = static_cast<unsigned>(a.num_slots >> 5);
      unsigned next = 0;
      if (lane == 0) next = atomicAdd(g.sync + 4, 1u);
      next = __shfl_sync(0xffffffffu, next, 0);
      while (next < nsl) {
        const unsigned sl = next;
        if (lane == 0) next = atomicAdd(g.sync + 4, 1u);  // (the next ticket travels while this slice is worked on)
        const int64_t slot = (static_cast<int64_t>(sl) << 5) + lane;
        const int64_t pos = a.num_split + (slot - a.num_virtual_padded);  // (no split rows: the host takes the multi-launch path for those)
        const bool own_row = slot >= a.num_virtual_padded && pos < a.num_rows;
        typename Epi::Pre pre;
        if (own_row) pre = epi.prefetch(ctx, pos);
        const double acc = sell_row8_cg(a, slot, xt);
        double red[2] = {0.0, 0.0};
        if (own_row) epi(ctx, pos, acc, red, pre);
        const double r0 = warp_sum(red[0]), r1 = warp_sum(red[1]);
        if (lane == 0) *reinterpret_cast<double2*>(g.slice_partials + 2 * static_cast<int64_t>(sl)) = make_double2(r0, r1);
        next = __shfl_sync(0xffffffffu, next, 0);
      }
    }
    unsigned long long t3 = 0;
    if (tracing) t3 = loop_now();
    if (MODE == 1) {
      if (loop_grid_sync<BT>(g, bar, kLoopSyncPlain, 0, 0, s_flags, s_tot)) break;
      // the slots of a fixed range of slices per block, in a fixed order
      if (blockIdx.x == 0 && tid == 0) g.sync[4] = 0u;
      const int64_t nsl = g.rows.num_slots >> 5;
      const int64_t per = (nsl + gridDim.x - 1) / gridDim.x;
      const int64_t sb = min(nsl, per * blockIdx.x), se = min(nsl, sb + per);
      double red[2] = {0.0, 0.0};
      for (int64_t k = sb + tid; k < se; k += BT) {
        const double2 v = __ldcg(reinterpret_cast<const double2*>(g.slice_partials) + k);
        red[0] += v.x;
        red[1] += v.y;
      }
      block_reduce_store<2, 0, BT>(red, nullptr, g.block_partials + 4 * blockIdx.x + 1);
    }
    if (MODE == 1) {
      // ---- T1: the local partial (K[R_g,:])^T y' into the own arena, column order --------
      const double* yn = pick3(b.y, cand);
      double* partial = peer.base[peer.rank] + peer.partial_off;
      const SellDev& a = g.cols;
      const unsigned nsl = static_cast<unsigned>(a.num_slots >> 5);
      unsigned next = 0;
      if (lane == 0) next = atomicAdd(g.sync + 5, 1u);
      next = __shfl_sync(0xffffffffu, next, 0);
      while (next < nsl) {
        const unsigned sl = next;
        if (lane == 0) next = atomicAdd(g.sync + 5, 1u);
        const int64_t slot = (static_cast<int64_t>(sl) << 5) + lane;
        const int64_t pos = a.num_split + (slot - a.num_virtual_padded);
        const bool own_row = slot >= a.num_virtual_padded && pos < a.num_rows;
        int32_t dst = 0;
        if (own_row) dst = __ldg(g.perm + pos);
        const double acc = sell_row8_cg(a, slot, yn);
        if (own_row) partial[dst] = acc;
        next = __shfl_sync(0xffffffffu, next, 0);
      }
    }
    unsigned long long t4 = 0;
    if (tracing) t4 = loop_now();
    if (loop_grid_sync<BT>(g, bar, kLoopSyncSumsPeer, 1, MODE == 0 ? (g.rows.num_slots >> 5) : 0, s_flags, s_tot)) break;
    unsigned long long t5 = 0;
    if (tracing) t5 = loop_now();

    if (MODE == 0 && blockIdx.x == 0 && tid == 0) g.sync[4] = 0u;  // (tickets of D: every draw of this attempt is behind the barrier above)
    // ---- the decision (one warp; run_decide_head of the multi-launch path) --------------
    if (decider && lane == 0) loop_decide(st_out, &s_st, peer.base[peer.rank] + peer.scal_off, peer.world);
    // ---- T: K^T y' of the candidate for this rank's column slice -------------------------
    double* kc = pick3(b.kty, cand);
    if (MODE == 0) {
      const double* yall = peer.base[peer.rank] + peer.y_off;
      const SellDev& a = g.cols;
      const unsigned nsl = static_cast<unsigned>(a.num_slots >> 5);
      unsigned next = 0;
      if (lane == 0) next = atomicAdd(g.sync + 6, 1u);
      next = __shfl_sync(0xffffffffu, next, 0);
      while (next < nsl) {
        const unsigned sl = next;
        if (lane == 0) next = atomicAdd(g.sync + 6, 1u);
        const int64_t slot = (static_cast<int64_t>(sl) << 5) + lane;
        const int64_t pos = a.num_split + (slot - a.num_virtual_padded);
        const bool own_row = slot >= a.num_virtual_padded && pos < a.num_rows;
        int32_t dst = 0;
        if (own_row) dst = __ldg(g.perm + pos);
        const double acc = sell_row8_cg(a, slot, yall);
        if (own_row) kc[g.col0 + dst] = acc;
        next = __shfl_sync(0xffffffffu, next, 0);
      }
    } else {
      for (int64_t i0 = peer.begin + 2 * gthread; i0 < peer.end; i0 += 2 * nthreads) {
        if (i0 + 1 < peer.end) {
          double2 v = make_double2(0.0, 0.0);
#pragma unroll
          for (int h = 0; h < kMaxPeers; ++h) {
            if (h < peer.world) {
              const double2 t = __ldcg(reinterpret_cast<const double2*>(peer.base[h] + peer.partial_off + i0));
              v.x += t.x;
              v.y += t.y;
            }
          }
          *reinterpret_cast<double2*>(kc + i0) = v;
        } else {
          double v = 0.0;
#pragma unroll
          for (int h = 0; h < kMaxPeers; ++h)
            if (h < peer.world) v += __ldcg(peer.base[h] + peer.partial_off + i0);
          kc[i0] = v;
        }
      }
    }
    unsigned long long t6 = 0;
    if (tracing) t6 = loop_now();
    if (loop_grid_sync<BT>(g, bar, kLoopSyncPlain, 0, 0, s_flags, s_tot)) break;
    if (tracing) {
      const unsigned long long t7 = loop_now();
      g.trace[0] += 1ull;
      g.trace[1] += t1 - t0;  // P
      g.trace[2] += t2 - t1;  // grid + peer barrier A
      g.trace[3] += t3 - t2;  // D
      g.trace[4] += t4 - t3;  // grid barrier + block sums (+ T1, MODE 1)
      g.trace[5] += t5 - t4;  // sums + grid + peer barrier B
      g.trace[6] += t6 - t5;  // decision + T
      g.trace[7] += t7 - t6;  // closing grid barrier
    }
  }
  if (s_flags[0] != 0 && blockIdx.x == 0 && tid == 0) {  // a peer never arrived: both slots say so, whatever the host reads
    b.state[0].halt = kHaltPeerTimeout;
    b.state[1].halt = kHaltPeerTimeout;
  }
}

// ---- Malitsky-Pock rule on the device (pdhg.cc:2463-2556) ---------------------------
// One attempt = one inner iteration of the line search; the kernels of an attempt read state slot
// `in`, k_mp_decide writes the other slot. x' and K x' are computed by the first attempt of an
// iteration only (mp_skip_primal marks the retries).
__global__ void __launch_bounds__(kThreads) k_mp_primal(StepPtrs b, double* partials) {
  pdl_trigger();
  pdl_wait();
  const StepState* st = b.state;
  if (st->halt != 0 || st->mp_skip_primal != 0) return;
  const double* __restrict__ xc = pick3(b.x, st->cur);
  const double* __restrict__ xp = pick3(b.x, st->prev);
  double* __restrict__ xn = pick3(b.x, st->cand);
  const double* __restrict__ kty = pick3(b.kty, st->cur);
  const double tau = st->step_size / st->primal_weight;
  const double r0 = st->pending_ratio0, r1 = st->pending_ratio;
  double s = 0.0;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
  if (i < b.n) {
    const double x = xc[i];
    double t = x - tau * (b.c[i] - kty[i]);
    if (b.q != nullptr) t = t / (tau * b.q[i] + 1.0);
    const double nx = fmax(fmin(t, b.uv[i]), b.lv[i]);
    xn[i] = nx;
    const double d = nx - x;
    s = d * d;
    if (r0 > 0.0 || r1 > 0.0) {  // deferred ShardedWeightedAverage::Add of the accepted step (and of its starting point)
      double av = b.avg_x[i];
      if (r0 > 0.0) av += r0 * (xp[i] - av);
      if (r1 > 0.0) av += r1 * (x - av);
      b.avg_x[i] = av;
    }
  }
  block_reduce_store<1, 0>(&s, nullptr, partials + blockIdx.x);
}
struct KxStoreEpi {  // K x' of the candidate into kx[cand]
  StepPtrs b;
  struct Ctx { double* kx_cand; };
  struct Pre {};
  __device__ __forceinline__ Ctx begin() const { return Ctx{pick3(b.kx, b.state->cand)}; }
  __device__ __forceinline__ Pre prefetch(const Ctx&, int64_t) const { return Pre(); }
  __device__ __forceinline__ void operator()(const Ctx& c, int64_t pos, double v, double*, const Pre&) const { c.kx_cand[pos] = v; }
};
__global__ void __launch_bounds__(kThreads) k_mp_dual(StepPtrs b, double* partials) {  // pdhg.cc:1905-1910, 1923-1928 with the trial step
  pdl_trigger();
  pdl_wait();
  const StepState* st = b.state;
  if (st->halt != 0) return;
  const double* __restrict__ yc = pick3(b.y, st->cur);
  double* __restrict__ yn = pick3(b.y, st->cand);
  const double* __restrict__ kxc = pick3(b.kx, st->cur);
  const double* __restrict__ kxn = pick3(b.kx, st->cand);
  const double omega = st->primal_weight;
  const double tau = st->step_size / omega, new_tau = st->mp_new_tau;
  const double theta = new_tau / tau;
  const double sigma = (omega * omega) * new_tau;
  const double rd = st->pending_ratio_dual;
  double s = 0.0;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
  if (i < b.m) {
    const double y = yc[i];
    const double t = y - sigma * (-theta * kxc[i] + (theta + 1) * kxn[i]);
    const double ny = fmax(fmin(0.0, t + sigma * b.uc[i]), t + sigma * b.lc[i]);
    yn[i] = ny;
    const double d = ny - y;
    s = d * d;
    if (rd > 0.0) b.avg_y[i] += rd * (y - b.avg_y[i]);
  }
  block_reduce_store<1, 0>(&s, nullptr, partials + blockIdx.x);
}
struct KtyDiffEpi {  // K^T y' of the trial into kty[cand], ||K^T y - K^T y'||^2
  StepPtrs b;
  struct Ctx { const double* kty_cur; double* kty_cand; };
  struct Pre { double kty; };
  __device__ __forceinline__ Ctx begin() const { return Ctx{pick3(b.kty, b.state->cur), pick3(b.kty, b.state->cand)}; }
  __device__ __forceinline__ Pre prefetch(const Ctx& c, int64_t pos) const { return Pre{c.kty_cur[pos]}; }
  __device__ __forceinline__ void operator()(const Ctx& c, int64_t pos, double v, double* red, const Pre& p) const {
    c.kty_cand[pos] = v;
    const double d = p.kty - v;
    red[0] += d * d;
  }
};
__global__ void __launch_bounds__(kDecideThreads) k_mp_decide(const StepState* in, StepState* out, const double* pp, int np, const double* pd, int nd, const double* pt, int nt) {
  pdl_trigger();
  pdl_wait();
  StepState loaded;
  if (threadIdx.x == 0) loaded = *in;
  __shared__ double sh3[3][kDecideThreads / 32];
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll 4
  for (int i = threadIdx.x; i < np; i += kDecideThreads) s0 += pp[i];
#pragma unroll 4
  for (int i = threadIdx.x; i < nd; i += kDecideThreads) s1 += pd[i];
#pragma unroll 8
  for (int i = threadIdx.x; i < nt; i += kDecideThreads) s2 += pt[i];
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if ((threadIdx.x & 31) == 0) {
    sh3[0][threadIdx.x >> 5] = s0;
    sh3[1][threadIdx.x >> 5] = s1;
    sh3[2][threadIdx.x >> 5] = s2;
  }
  __syncthreads();
  if (threadIdx.x != 0 || loaded.halt != 0) return;
  double dx2 = 0.0, dy2 = 0.0, dk2 = 0.0;
  for (int w = 0; w < kDecideThreads / 32; ++w) { dx2 += sh3[0][w]; dy2 += sh3[1][w]; dk2 += sh3[2][w]; }
  StepState s = loaded;
  s.attempts += 1;
  const double omega = s.primal_weight;
  const double tau = s.step_size / omega, new_tau = s.mp_new_tau;
  const double theta = new_tau / tau;
  if (omega * new_tau * sqrt(dk2) <= s.mp_contraction * sqrt(dy2)) {
    s.step_size = new_tau * omega;
    s.mp_ratio = theta;
    // averaging (pdhg.cc:2514-2532): the starting point enters an empty primal average first
    s.pending_ratio0 = 0.0;
    if (!(s.avg_weight_sum_primal > 0.0)) {
      const double w0 = new_tau * theta;
      if (w0 > 0.0) {
        s.pending_ratio0 = 1.0;
        s.avg_weight_sum_primal = w0;
      }
      s.avg_num_terms_primal += 1;
    }
    if (new_tau > 0.0) {
      s.pending_ratio = new_tau / (s.avg_weight_sum_primal + new_tau);
      s.avg_weight_sum_primal += new_tau;
      s.pending_ratio_dual = new_tau / (s.avg_weight_sum + new_tau);
      s.avg_weight_sum += new_tau;
    } else {
      s.pending_ratio = s.pending_ratio_dual = 0.0;
    }
    s.avg_num_terms_primal += 1;
    s.avg_num_terms += 1;
    const int old_prev = s.prev;
    s.prev = s.cur;
    s.cur = s.cand;
    s.cand = old_prev;
    const double movement = (0.5 * omega * dx2) + (0.5 / omega) * dy2;
    s.last_dx2 = dx2;
    s.last_dy2 = dy2;
    s.last_movement = movement;
    s.last_nonlinearity = 0.0;
    s.num_rejected_steps += s.inner_iterations;
    s.inner_iterations = 0;
    if (movement == 0.0 || movement > 1.0e100) {
      s.halt = movement == 0.0 ? kHaltZeroMovement : kHaltDivergent;  // (the host counts this iteration, as for the other rules)
    } else {
      s.iterations_completed += 1;
      s.mp_new_tau = new_tau * (1.0 + s.mp_interpolation * (sqrt(1.0 + theta) - 1.0));
      const double kkt = static_cast<double>(s.iterations_completed) + 0.5 * static_cast<double>(s.num_rejected_steps);
      if (s.iterations_completed >= s.k_stop || kkt >= s.kkt_pass_limit) s.halt = kHaltCheckpoint;
    }
    s.mp_skip_primal = s.halt != 0 ? 1 : 0;
  } else {
    s.mp_new_tau = s.mp_downscaling * new_tau;
    s.inner_iterations += 1;
    s.pending_ratio = s.pending_ratio_dual = s.pending_ratio0 = 0.0;
    s.mp_skip_primal = 1;
    if (s.inner_iterations >= 60) {
      s.halt = kHaltInnerLimit;
      s.num_rejected_steps += s.inner_iterations;
      s.inner_iterations = 0;
    }
  }
  *out = s;
}

// ------------------------------------------------------ trust region -------
__device__ __forceinline__ unsigned long long crit_key(double crit) {
  // crit >= 0 (or +inf); canonicalise -0.0 and tiny negatives from roundoff.
  const double c = crit > 0.0 ? crit : 0.0;
  return static_cast<unsigned long long>(__double_as_longlong(c));
}
constexpr unsigned long long kMaxKey = 0x7FEFFFFFFFFFFFFFull;  // DBL_MAX

struct TrSearchState {
  unsigned long long lo;
  int done, pad;   // no key with a nonzero contribution is left inside the bracket
  int shift, width_shift;  // k_tr_solve: digit position of the next pass; the bracket is (lo, lo + 2^width_shift]
  double fixed_radius_sq, variable_coef;
  double radius_sq;
  double max_abs_objective;
  double step_size;
  double frozen_a, frozen_b;     // k_tr_solve: sums of the elements already compacted away (below / above the bracket)
  double lag_primal, lag_dual;   // k_tr_solve<JOINT>: Lagrangian value parts (sou.cc:446-527)
  double radius;
  double dist_primal_sq, dist_dual_sq;  // k_tr_solve<JOINT>: ||x - x0||^2, ||y - y0||^2 (when asked for)
};

// Joint problem elements, trust_region.cc:115-162 / 538-607.
struct JointElem {
  const double *x, *y, *kx, *kty, *c, *q, *lv, *uv, *lc, *uc;
  double primal_weight;
  int64_t n, m;   // joint index i < n is primal element pbeg + i; i >= n is dual element i - n
  int64_t pbeg;   // first primal element this rank counts (a slice of the replicated primal side on a row-sharded solve)
  __device__ __forceinline__ double primal_gradient(int64_t i) const {
    return q != nullptr ? (c[i] + q[i] * x[i] - kty[i]) : (c[i] - kty[i]);
  }
  __device__ __forceinline__ double subgradient_coefficient(int64_t j) const {  // sou.cc:476-500
    const double dual = y[j], lb = lc[j], ub = uc[j], pp = kx[j];
    if (dual < 0.0) return ub;
    if (dual > 0.0) return lb;
    const bool lf = isfinite(lb), uf = isfinite(ub);
    if (lf && uf) return pp < lb ? lb : (pp > ub ? ub : pp);
    if (lf) return lb;
    if (uf) return ub;
    return 0.0;
  }
  // k_tr_solve<JOINT> round 0: Lagrangian value parts (sou.cc:446-527) and squared distances to
  // (x0, y0) from what get_k loaded (arithmetic only).
  __device__ __forceinline__ void joint_extras(int64_t i, double obj, double center, double qd, double x0v, double coef, double& e0, double& e1,
                                               double& d0, double& d1) const {
    const double d = center - x0v;
    if (i < n) {
      const double op = qd * center;  // (q x; qd is 0 for an LP)
      e0 += center * (obj - 0.5 * op);
      d0 += d * d;
    } else {
      e1 += coef * center;
      d1 += d * d;
    }
  }
  __device__ __forceinline__ double weighted_distance_sq(double d0, double d1) const { return (0.5 * primal_weight) * d0 + (0.5 / primal_weight) * d1; }
  // Branch-free accessors for a segment of one kind (KIND 0: primal elements i < n, 1: dual
  // elements): the sweeps of k_tr_solve walk the two segments of a block's range separately so
  // that the loads of several elements can be in flight together.
  __device__ __forceinline__ int64_t split() const { return n; }
  template <int KIND>
  __device__ __forceinline__ void get_k(int64_t i, double& obj, double& lb, double& ub, double& center, double& w, double& qd, double& x0v, double& coef,
                                        const double* x0, const double* y0) const {
    if (KIND == 0) {
      const int64_t ip = pbeg + i;
      center = x[ip];
      qd = q != nullptr ? q[ip] : 0.0;
      obj = q != nullptr ? (c[ip] + qd * center - kty[ip]) : (c[ip] - kty[ip]);
      lb = lv[ip];
      ub = uv[ip];
      w = 0.5 * primal_weight;
      x0v = x0 != nullptr ? x0[ip] : center;
      coef = 0.0;
    } else {
      const int64_t j = i - n;
      const double lcv = lc[j], ucv = uc[j];
      center = y[j];
      coef = subgradient_coefficient(j);
      obj = -(coef - kx[j]);
      lb = isfinite(ucv) ? -kInfD : 0.0;
      ub = isfinite(lcv) ? kInfD : 0.0;
      w = 0.5 / primal_weight;
      qd = 0.0;
      x0v = y0 != nullptr ? y0[j] : center;
    }
  }
  __device__ __forceinline__ void get(int64_t i, double& obj, double& lb, double& ub, double& center, double& w, double& qd) const {
    if (i < n) {
      const int64_t ip = pbeg + i;
      obj = primal_gradient(ip);
      lb = lv[ip];
      ub = uv[ip];
      center = x[ip];
      w = 0.5 * primal_weight;
      qd = q != nullptr ? q[ip] : 0.0;
    } else {
      const int64_t j = i - n;
      obj = -(subgradient_coefficient(j) - kx[j]);
      lb = isfinite(uc[j]) ? -kInfD : 0.0;
      ub = isfinite(lc[j]) ? kInfD : 0.0;
      center = y[j];
      w = 0.5 / primal_weight;
      qd = 0.0;
    }
  }
};
struct VectorElem {
  const double *obj_, *lb_, *ub_, *center_, *w_, *q_;
  int64_t n;  // (unused; k_tr_solve<JOINT> is never instantiated for explicit vectors)
  __device__ __forceinline__ void joint_extras(int64_t, double, double, double, double, double, double&, double&, double&, double&) const {}
  __device__ __forceinline__ double weighted_distance_sq(double, double) const { return 0.0; }
  __device__ __forceinline__ void get(int64_t i, double& obj, double& lb, double& ub, double& center, double& w, double& qd) const {
    obj = obj_[i]; lb = lb_[i]; ub = ub_[i]; center = center_[i]; w = w_[i];
    qd = q_ != nullptr ? q_[i] : 0.0;
  }
  __device__ __forceinline__ int64_t split() const { return INT64_MAX; }  // one kind only
  template <int KIND>
  __device__ __forceinline__ void get_k(int64_t i, double& obj, double& lb, double& ub, double& center, double& w, double& qd, double& x0v, double& coef,
                                        const double*, const double*) const {
    get(i, obj, lb, ub, center, w, qd);
    x0v = center;
    coef = 0.0;
  }
};

// crit (as key), a = w dist^2 (radius^2 if fixed at its bound), b = obj^2 / w.
template <class Elem>
__global__ void __launch_bounds__(kThreads) k_tr_prepare(int64_t total, int64_t first, Elem el, unsigned long long* keys, double* a, double* bcoef, double* partials) {
  double m[1] = {-kInfD};
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * kThreads) {
    if (i < first) {  // replicated (primal) element, counted on rank 0 only
      keys[i] = 0ull;
      a[i] = 0.0;
      bcoef[i] = 0.0;
      continue;
    }
    double obj, lb, ub, center, w, qd;
    el.get(i, obj, lb, ub, center, w, qd);
    double crit, dist = 0.0;
    if (obj == 0.0) {
      crit = kInfD;
    } else {
      dist = (obj > 0.0 ? lb : ub) - center;  // DistanceAtCriticalStepSize
      crit = -w * dist / obj;                  // CriticalStepSize
    }
    keys[i] = crit_key(crit);
    a[i] = w * dist * dist;
    bcoef[i] = obj * obj / w;
    m[0] = fmax(m[0], fabs(obj));
  }
  block_reduce_store<0, 1>(nullptr, m, partials + blockIdx.x);
}

// One radix-16 pass: per candidate threshold c_j = lo + j * 2^shift (j = 0..15)
// accumulate A_j = sum_{key <= c_j} a and B_j = sum_{key > c_j} b via 17 bins.
// Four elements per thread and trip are loaded up front (12 independent loads in
// flight) because the 34 accumulators keep the occupancy low.
constexpr int kTrUnroll = 4;
__global__ void __launch_bounds__(kThreads) k_tr_pass(int64_t total, const unsigned long long* __restrict__ keys, const double* __restrict__ a,
                                                      const double* __restrict__ bcoef, const TrSearchState* st, int shift, double* partials) {
  if (st->done != 0) return;
  const unsigned long long lo = st->lo;
  double ba[17], bb[17];
#pragma unroll
  for (int j = 0; j < 17; ++j) { ba[j] = 0.0; bb[j] = 0.0; }
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads * kTrUnroll;
  for (int64_t base = static_cast<int64_t>(blockIdx.x) * kThreads * kTrUnroll + threadIdx.x; base < total; base += stride) {
    unsigned long long key[kTrUnroll];
    double av[kTrUnroll], bv[kTrUnroll];
#pragma unroll
    for (int u = 0; u < kTrUnroll; ++u) {
      const int64_t i = base + static_cast<int64_t>(u) * kThreads;
      const bool ok = i < total;
      key[u] = ok ? __ldcs(keys + i) : 0ull;
      av[u] = ok ? __ldcs(a + i) : 0.0;
      bv[u] = ok ? __ldcs(bcoef + i) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < kTrUnroll; ++u) {
      int j0 = 0;
      if (key[u] > lo) {
        const unsigned long long d = (key[u] - lo - 1ull) >> shift;
        j0 = d >= 15ull ? 16 : static_cast<int>(d) + 1;
      }
#pragma unroll
      for (int j = 0; j < 17; ++j) {
        const bool p = (j0 == j);
        ba[j] += p ? av[u] : 0.0;
        bb[j] += p ? bv[u] : 0.0;
      }
    }
  }
  double s[34];
#pragma unroll
  for (int j = 0; j < 17; ++j) { s[j] = ba[j]; s[17 + j] = bb[j]; }
  block_reduce_store<34, 0>(s, nullptr, partials + static_cast<int64_t>(blockIdx.x) * 34);
}

// The 34 bin totals: 17 warps sum two columns each over the blocks in a fixed
// lane-strided order (k_tr_totals); row-sharded solves all-reduce the totals;
// one thread then picks the bracket (k_tr_pick).
__device__ __forceinline__ void tr_pick(const double* tot, TrSearchState* st_dev, int shift);
// pick_here: single-GPU launches finish the pass in the same kernel.
__global__ void __launch_bounds__(17 * 32) k_tr_totals(int nblocks, const double* __restrict__ partials, double* __restrict__ tot, TrSearchState* st,
                                                       int shift, int pick_here) {
  if (st->done != 0) return;
  __shared__ double sh_tot[34];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int colx = warp + 17 * h;
    double s = 0.0;
#pragma unroll 4
    for (int b = lane; b < nblocks; b += 32) s += partials[static_cast<int64_t>(b) * 34 + colx];
    s = warp_sum(s);
    if (lane == 0) { tot[colx] = s; sh_tot[colx] = s; }
  }
  if (!pick_here) return;
  __syncthreads();
  if (threadIdx.x == 0) tr_pick(sh_tot, st, shift);
}
__global__ void k_tr_pick(const double* __restrict__ tot, TrSearchState* st, int shift) {
  if (st->done != 0) return;
  tr_pick(tot, st, shift);
}
__device__ __forceinline__ void tr_pick(const double* tot, TrSearchState* st, int shift) {
  const unsigned long long lo = st->lo;
  // A_j = sum of bins 0..j of a; B_j = sum of bins j+1..16 of b.
  double A = 0.0;
  double suffix[18];
  suffix[17] = 0.0;
  for (int j = 16; j >= 0; --j) suffix[j] = suffix[j + 1] + tot[17 + j];
  int best = 0;
  double bestA = tot[0], bestB = suffix[1];
  for (int j = 0; j < 16; ++j) {
    A += tot[j];
    if (j == 0) continue;
    const unsigned long long cj = lo + (static_cast<unsigned long long>(j) << shift);
    if (cj > kMaxKey || cj < lo) break;
    const double t = __longlong_as_double(static_cast<long long>(cj));
    const double B = suffix[j + 1];
    const double f = A + (B > 0.0 ? t * t * B : 0.0);
    if (f <= st->radius_sq) { best = j; bestA = A; bestB = B; } else break;
  }
  st->lo = lo + (static_cast<unsigned long long>(best) << shift);
  st->fixed_radius_sq = bestA;
  st->variable_coef = bestB;
  // The next bracket is bin best+1. If nothing in it contributes to either sum,
  // A and B can no longer change and the closed form of k_tr_finish is exact.
  if (tot[best + 1] == 0.0 && tot[17 + best + 1] == 0.0) st->done = 1;
}

__device__ __forceinline__ double projected_value(double center, double obj, double w, double lb, double ub, double step) {
  const double full = center - step * obj / w;  // trust_region.h:223-228
  return fmin(fmax(full, lb), ub);
}

// ---- the whole trust-region solve in ONE persistent cooperative launch -----------
// k_tr_solve does what k_tr_prepare + k_tr_init + the radix passes + k_tr_finish (+ the reductions
// around them) did in separate launches:
//   round 0   every block walks its CONTIGUOUS range of the joint problem once: critical step
//             sizes (as keys), a, b to scratch; max |objective|, the smallest positive and the
//             largest finite critical step size, (JOINT) the Lagrangian value parts and the
//             distance to the last restart point (= the radius).
//   passes    radix-16 on the key bits, but STARTING at the bits where the smallest and the largest
//             key differ (the top passes of a search from bit 60 split nothing). Every pass filters
//             the block's list to the current bracket and compacts it in place (stable,
//             block-local), so a pass costs as much as the bracket is populated; elements that left
//             the bracket live on in the running sums frozen_a / frozen_b. Once the bracket holds
//             at most kTrFinishCap elements the block that closes the round finishes the remaining
//             passes alone (no more grid rounds).
//   final     (JOINT) objective deltas at the solution (trust_region.cc:929-967).
// A round ends with a ticket: the block that arrives last adds the per-block vectors in a fixed
// order (independent of which block it is), exchanges them with the other ranks of a row-sharded
// solve through the peer arenas, takes the bracket decision, publishes the state and bumps a
// generation counter the other blocks wait for. One launch, one device->host copy.
constexpr int kTrBlocksPerSm = 2;
constexpr int kTrMaxBlocks = 320;    // grid cap
constexpr int kTrCols = 42;          // doubles per round vector: 34 bins + extras
constexpr int kTrFinishCap = 4096;   // bracket population below which one block finishes the search
enum { kTrLagP = 34, kTrLagD = 35, kTrDistP = 36, kTrDistD = 37, kTrCount = 38, kTrFirstMax = 39, kTrMaxObj = 39, kTrMaxCrit = 40, kTrNegMinCrit = 41 };
struct TrSolveArgs {
  int64_t total;
  unsigned long long* keys;
  double *a, *b;             // scratch [total] each
  TrSearchState* st;         // result (device)
  double* partials;          // [blocks][kTrCols]
  unsigned int* sync;        // [0] arrivals [1] generation; zero before the launch
  double radius;             // < 0 (JOINT only): the weighted distance to (x0, y0), computed in round 0
  const double *x0, *y0;
  double* out;               // JOINT: {lagrangian primal part, dual part, primal delta, dual delta, radius, ||x-x0||^2, ||y-y0||^2}
  int32_t* peer_error;
  unsigned long long* cand_keys;  // [kTrFinishCap] each: the bracket's elements, gathered by the finishing block
  double *cand_a, *cand_b;
  unsigned long long* trace;  // PDLP_B200_TRACE=1: [0] count, then globaltimer (ns) stamps of block 0
};
__device__ __forceinline__ void tr_stamp(const TrSolveArgs& g) {
  if (g.trace == nullptr || blockIdx.x != 0 || threadIdx.x != 0) return;
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  const unsigned long long k = g.trace[0];
  if (k < 30ull) { g.trace[1 + k] = t; g.trace[0] = k + 1ull; }
}

__device__ __forceinline__ void tr_pick_compact(const double* tot, TrSearchState* st, int shift) {
  const unsigned long long lo = st->lo;
  // A_j = frozen_a + bins 0..j of a; B_j = frozen_b + bins j+1..16 of b.
  double suffix[18];
  suffix[17] = st->frozen_b;
  for (int j = 16; j >= 0; --j) suffix[j] = suffix[j + 1] + tot[17 + j];
  double A = st->frozen_a;
  int best = 0;
  double bestA = A + tot[0], bestB = suffix[1];
  for (int j = 0; j < 16; ++j) {
    A += tot[j];
    if (j == 0) continue;
    const unsigned long long cj = lo + (static_cast<unsigned long long>(j) << shift);
    if (cj > kMaxKey || cj < lo) break;
    const double t = __longlong_as_double(static_cast<long long>(cj));
    const double B = suffix[j + 1];
    const double f = A + (B > 0.0 ? t * t * B : 0.0);
    if (f <= st->radius_sq) { best = j; bestA = A; bestB = B; } else break;
  }
  st->lo = lo + (static_cast<unsigned long long>(best) << shift);
  st->fixed_radius_sq = bestA;
  st->variable_coef = bestB;
  st->frozen_a = bestA;               // bins 0..best leave the list below the bracket
  st->frozen_b = suffix[best + 2];    // bins best+2.. leave it above
  st->width_shift = shift;            // the next bracket is (lo, lo + 2^shift]
  st->shift = shift >= 4 ? shift - 4 : 0;
  if ((tot[best + 1] == 0.0 && tot[17 + best + 1] == 0.0) || shift == 0) st->done = 1;
}

// Column sums of the per-thread bins of this block (warp w adds columns w, w + 8, ... over the 256
// threads: lane-strided, then a shuffle tree); columns >= kTrFirstMax are maxima.
__device__ __forceinline__ void tr_block_totals(const double* bins, double* dst, int K) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int j = warp; j < K; j += kThreads / 32) {
    double v;
    if (j >= kTrFirstMax) {
      v = -kInfD;
#pragma unroll
      for (int t = lane; t < kThreads; t += 32) v = fmax(v, bins[j * kThreads + t]);
      v = warp_max(v);
    } else {
      v = 0.0;
#pragma unroll
      for (int t = lane; t < kThreads; t += 32) v += bins[j * kThreads + t];
      v = warp_sum(v);
    }
    if (lane == 0) dst[j] = v;
  }
}

// End of a round. In: bins[c][t] holds thread t's contribution to column c < K. Out (in the block
// that arrived last, returns true there): tot[0..K) over all blocks and ranks. All threads of all
// blocks call it; `round` counts the calls.
template <bool PEER>
__device__ bool tr_round(const TrSolveArgs& g, const PeerPtrs& peer, double* bins, double* tot, double (*part)[48], int* s_last, int K, unsigned round) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nb = gridDim.x;
  __syncthreads();
  tr_block_totals(bins, g.partials + static_cast<int64_t>(blockIdx.x) * kTrCols, K);
  __syncthreads();
  if (tid == 0) {
    __threadfence();  // (cumulative: the whole block's stores, ordered by the barrier, are visible before the ticket)
    const unsigned int ticket = atomicAdd(g.sync, 1u);
    *s_last = ticket == static_cast<unsigned int>(nb) * (round + 1u) - 1u ? 1 : 0;
  }
  __syncthreads();
  if (*s_last == 0) return false;
  __threadfence();
  {  // fixed-order sum over the blocks: five interleaved row groups per column (240 threads), all
     // loads of a thread in flight together, groups combined in order
    constexpr int kGroups = 5, kBatch = 16;
    const int colx = tid % 48, grp = tid / 48;
    if (grp < kGroups && colx < K) {
      const bool is_max = colx >= kTrFirstMax;
      double v = is_max ? -kInfD : 0.0;
      for (int b0 = grp; b0 < nb; b0 += kGroups * kBatch) {
        double term[kBatch];
#pragma unroll
        for (int k = 0; k < kBatch; ++k) {
          const int b = b0 + kGroups * k;
          term[k] = b < nb ? __ldcg(g.partials + static_cast<int64_t>(b) * kTrCols + colx) : (is_max ? -kInfD : 0.0);
        }
#pragma unroll
        for (int k = 0; k < kBatch; ++k) v = is_max ? fmax(v, term[k]) : v + term[k];
      }
      part[grp][colx] = v;
    }
    __syncthreads();
    if (tid < K) {
      double v = part[0][tid];
#pragma unroll
      for (int k = 1; k < kGroups; ++k) v = tid >= kTrFirstMax ? fmax(v, part[k][tid]) : v + part[k][tid];
      tot[tid] = v;
    }
    __syncthreads();
  }
  if (PEER) {
    const int64_t slot = peer.tr_off + static_cast<int64_t>(round & 1u) * kTrCols * peer.world;
    if (warp == 0) {
      for (int h = 0; h < peer.world; ++h) {
        volatile double* dst = peer_base(peer, h) + slot + kTrCols * peer.rank;
        if (lane < K) dst[lane] = tot[lane];
        if (32 + lane < K) dst[32 + lane] = tot[32 + lane];
      }
      peer_barrier(peer, peer.tr_barrier, g.peer_error);
    }
    __syncthreads();
    if (tid < K) {
      const volatile double* src = peer_base(peer, peer.rank) + slot;
      double v = tid >= kTrFirstMax ? -kInfD : 0.0;
      for (int h = 0; h < peer.world; ++h) v = tid >= kTrFirstMax ? fmax(v, src[kTrCols * h + tid]) : v + src[kTrCols * h + tid];
      tot[tid] = v;
    }
    __syncthreads();
  }
  return true;
}

// Publishes the state decided by the last block / waits for it in the others; afterwards the
// shared copy `s` of every block equals the global state.
__device__ __forceinline__ void tr_publish_or_wait(const TrSolveArgs& g, TrSearchState* s, bool last, unsigned round) {
  constexpr int kWords = static_cast<int>(sizeof(TrSearchState) / sizeof(double));
  if (last) {
    __syncthreads();
    if (threadIdx.x < kWords) reinterpret_cast<double*>(g.st)[threadIdx.x] = reinterpret_cast<const double*>(s)[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      *reinterpret_cast<volatile unsigned int*>(g.sync + 1) = round + 1u;
    }
  } else {
    if (threadIdx.x == 0) {
      while (*reinterpret_cast<volatile unsigned int*>(g.sync + 1) < round + 1u) {}
      __threadfence();
    }
    __syncthreads();
    if (threadIdx.x < kWords) reinterpret_cast<double*>(s)[threadIdx.x] = __ldcg(reinterpret_cast<const double*>(g.st) + threadIdx.x);
  }
  __syncthreads();
}

// bin (1..16) of a key inside the bracket (lo, lo + 16 << shift]
__device__ __forceinline__ int tr_bin(unsigned long long key, unsigned long long lo, int shift) {
  const unsigned long long d = (key - lo - 1ull) >> shift;
  return d >= 15ull ? 16 : static_cast<int>(d) + 1;
}

template <class Elem, bool PEER, bool JOINT>
__global__ void __launch_bounds__(kThreads, kTrBlocksPerSm) k_tr_solve(TrSolveArgs g, Elem el, PeerPtrs peer) {
  extern __shared__ double bins[];  // [kTrCols][kThreads]
  __shared__ double tot[kTrCols];
  __shared__ double part[5][48];
  __shared__ TrSearchState s;
  __shared__ int s_last;
  __shared__ int s_wtot[kTrUnroll][kThreads / 32];
  static_assert(sizeof(TrSearchState) % sizeof(double) == 0 && sizeof(TrSearchState) / sizeof(double) <= kThreads, "state is copied as doubles");
  static_assert(kTrFinishCap == PeerLayout::kTrCandCap, "arena segment size");
  static_assert(kTrCols <= 48 && 5 * 48 <= kThreads, "layout of the cross-block sum");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = gridDim.x;
  constexpr int64_t kTile = static_cast<int64_t>(kThreads) * kTrUnroll;
  const int64_t per = ((g.total + nb - 1) / nb + kTile - 1) / kTile * kTile;
  const int64_t beg = min(g.total, static_cast<int64_t>(blockIdx.x) * per), end = min(g.total, beg + per);
  unsigned round = 0;
  tr_stamp(g);

  // ---- round 0: prepare ---------------------------------------------------------------
  {
    double mx = 0.0, maxc = -kInfD, negminc = -kInfD, e0 = 0.0, e1 = 0.0, d0 = 0.0, d1 = 0.0;
    // kTrUnroll elements per thread and trip, every load issued before the first store (the sweep
    // is bound by memory latency at two 256-thread blocks per SM); the primal and the dual segment
    // of the range are walked separately so that the accessors are branch-free
    auto sweep = [&](auto kind, int64_t sb, int64_t se) {
      constexpr int KIND = decltype(kind)::value;
      for (int64_t base = sb + tid; base < se; base += kTile) {
        double obj[kTrUnroll], lb[kTrUnroll], ub[kTrUnroll], center[kTrUnroll], w[kTrUnroll], qd[kTrUnroll];
        double x0v[kTrUnroll], coef[kTrUnroll];
#pragma unroll
        for (int u = 0; u < kTrUnroll; ++u) {
          const int64_t i = min(base + static_cast<int64_t>(u) * kThreads, se - 1);  // (clamped: a valid, discarded load)
          el.template get_k<KIND>(i, obj[u], lb[u], ub[u], center[u], w[u], qd[u], x0v[u], coef[u], g.x0, g.y0);
        }
#pragma unroll
        for (int u = 0; u < kTrUnroll; ++u) {
          const int64_t i = base + static_cast<int64_t>(u) * kThreads;
          if (i >= se) continue;
          double crit, dist = 0.0;
          if (obj[u] == 0.0) {
            crit = kInfD;
          } else {
            dist = (obj[u] > 0.0 ? lb[u] : ub[u]) - center[u];  // DistanceAtCriticalStepSize
            crit = -w[u] * dist / obj[u];                        // CriticalStepSize
          }
          const unsigned long long key = crit_key(crit);
          g.keys[i] = key;
          g.a[i] = w[u] * dist * dist;
          g.b[i] = obj[u] * obj[u] / w[u];
          mx = fmax(mx, fabs(obj[u]));
          if (key > 0ull && key <= kMaxKey) {  // positive and finite
            const double c = __longlong_as_double(static_cast<long long>(key));
            maxc = fmax(maxc, c);
            negminc = fmax(negminc, -c);
          }
          if (JOINT) el.joint_extras(i, obj[u], center[u], qd[u], x0v[u], coef[u], e0, e1, d0, d1);
        }
      }
    };
    const int64_t cut = el.split();
    sweep(std::integral_constant<int, 0>(), beg, min(end, cut));
    sweep(std::integral_constant<int, 1>(), max(beg, min(end, cut)), end);
    bins[kTrLagP * kThreads + tid] = e0;
    bins[kTrLagD * kThreads + tid] = e1;
    bins[kTrDistP * kThreads + tid] = d0;
    bins[kTrDistD * kThreads + tid] = d1;
    bins[kTrCount * kThreads + tid] = 0.0;
    bins[kTrMaxObj * kThreads + tid] = mx;
    bins[kTrMaxCrit * kThreads + tid] = maxc;
    bins[kTrNegMinCrit * kThreads + tid] = negminc;
#pragma unroll
    for (int j = 0; j < 34; ++j) bins[j * kThreads + tid] = 0.0;
  }
  tr_stamp(g);
  bool last = tr_round<PEER>(g, peer, bins, tot, part, &s_last, kTrCols, round);
  if (last && tid == 0) {
    double radius = g.radius;
    if (JOINT && radius < 0.0) radius = sqrt(el.weighted_distance_sq(tot[kTrDistP], tot[kTrDistD]));
    s.done = 0;
    s.pad = 0;
    s.fixed_radius_sq = 0.0;
    s.variable_coef = 0.0;
    s.radius_sq = radius * radius;
    s.max_abs_objective = fmax(0.0, tot[kTrMaxObj]);
    s.step_size = 0.0;
    s.frozen_a = 0.0;
    s.frozen_b = 0.0;
    s.lag_primal = tot[kTrLagP];
    s.lag_dual = tot[kTrLagD];
    s.radius = radius;
    s.dist_primal_sq = tot[kTrDistP];
    s.dist_dual_sq = tot[kTrDistD];
    // first bracket: just below the smallest positive key up to the largest finite one, in 16 bins
    s.lo = 0ull;
    s.shift = 60;
    if (tot[kTrMaxCrit] > 0.0) {
      const unsigned long long maxkey = static_cast<unsigned long long>(__double_as_longlong(tot[kTrMaxCrit]));
      const unsigned long long minkey = static_cast<unsigned long long>(__double_as_longlong(-tot[kTrNegMinCrit]));
      s.lo = minkey - 1ull;
      const unsigned long long span = maxkey - s.lo - 1ull;  // largest (key - lo - 1)
      s.shift = span < 16ull ? 0 : 64 - __clzll(static_cast<long long>(span)) - 4;
    }
    s.width_shift = 64;  // no upper end yet
    if (PEER && *reinterpret_cast<volatile int32_t*>(g.peer_error) != 0) s.done = 1;  // a peer never arrived
  }
  tr_publish_or_wait(g, &s, last, round);
  ++round;
  tr_stamp(g);

  // ---- passes: filter to the bracket, compact in place, bin ------------------------------
  int64_t len = end - beg;
  bool first_pass = true;
  while (s.done == 0) {
    const int shift = s.shift;
    const unsigned long long lo = s.lo;
    const bool open_end = s.width_shift >= 64;
    const unsigned long long width = open_end ? ~0ull : (1ull << s.width_shift);
#pragma unroll
    for (int j = 0; j < 34; ++j) bins[j * kThreads + tid] = 0.0;
    if (first_pass) {
      // Every positive key is inside the first bracket: a plain streaming pass (no compaction, no
      // barriers); keys <= lo (zero keys) are the part below every bracket, bin 0.
      for (int64_t base = beg + tid; base < end; base += kTile) {
        unsigned long long key[kTrUnroll];
        double av[kTrUnroll], bv[kTrUnroll];
#pragma unroll
        for (int u = 0; u < kTrUnroll; ++u) {
          const int64_t i = base + static_cast<int64_t>(u) * kThreads;
          const bool ok = i < end;
          key[u] = ok ? g.keys[i] : 0ull;
          av[u] = ok ? g.a[i] : 0.0;
          bv[u] = ok ? g.b[i] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < kTrUnroll; ++u) {
          const int j0 = key[u] > lo ? tr_bin(key[u], lo, shift) : 0;
          bins[j0 * kThreads + tid] += av[u];
          if (j0 > 0) bins[(17 + j0) * kThreads + tid] += bv[u];
        }
      }
    } else {
      int64_t wpos = 0;
      for (int64_t tbase = 0; tbase < len; tbase += kTile) {
        unsigned long long key[kTrUnroll];
        double av[kTrUnroll], bv[kTrUnroll];
        bool surv[kTrUnroll];
        int rank_in_warp[kTrUnroll];
#pragma unroll
        for (int u = 0; u < kTrUnroll; ++u) {
          const int64_t idx = tbase + static_cast<int64_t>(u) * kThreads + tid;
          const bool ok = idx < len;
          key[u] = ok ? g.keys[beg + idx] : 0ull;
          surv[u] = ok && key[u] > lo && key[u] - lo <= width;
          av[u] = surv[u] ? g.a[beg + idx] : 0.0;
          bv[u] = surv[u] ? g.b[beg + idx] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < kTrUnroll; ++u) {
          const unsigned int ballot = __ballot_sync(0xffffffffu, surv[u]);
          rank_in_warp[u] = __popc(ballot & ((1u << lane) - 1u));
          if (lane == 0) s_wtot[u][warp] = __popc(ballot);
          if (surv[u]) {
            const int j0 = tr_bin(key[u], lo, shift);
            bins[j0 * kThreads + tid] += av[u];
            bins[(17 + j0) * kThreads + tid] += bv[u];
          }
        }
        __syncthreads();  // every load of this tile is done before any of its survivors is written
        int tile_total = 0;
        int my_off[kTrUnroll];
#pragma unroll
        for (int u = 0; u < kTrUnroll; ++u) {
#pragma unroll
          for (int w = 0; w < kThreads / 32; ++w) {
            if (w == warp) my_off[u] = tile_total;
            tile_total += s_wtot[u][w];
          }
        }
#pragma unroll
        for (int u = 0; u < kTrUnroll; ++u) {
          if (surv[u]) {
            const int64_t dst = beg + wpos + my_off[u] + rank_in_warp[u];
            g.keys[dst] = key[u];
            g.a[dst] = av[u];
            g.b[dst] = bv[u];
          }
        }
        wpos += tile_total;
        __syncthreads();  // s_wtot is reused; the compacted prefix is complete before the next tile's loads
      }
      len = wpos;
    }
    // (after the first pass the list is still the whole range: its population is unknown, report "large")
    bins[kTrCount * kThreads + tid] = tid == 0 ? (first_pass ? 1.0e18 : static_cast<double>(len)) : 0.0;
    first_pass = false;
    last = tr_round<PEER>(g, peer, bins, tot, part, &s_last, kTrCount + 1, round);
    if (last) {
      if (tid == 0) {
        tr_pick_compact(tot, &s, shift);
        if (PEER && *reinterpret_cast<volatile int32_t*>(g.peer_error) != 0) s.done = 1;
      }
      __syncthreads();
      // Few elements left in the bracket: this block finishes the remaining passes alone. It first
      // gathers the compacted lists of all blocks (their lengths are column kTrCount of the
      // partials) into one contiguous buffer -- every load independent of the others -- and then
      // runs the passes over that buffer.
      // Row-sharded solves: every rank stores its candidates into every rank's arena, the ranks
      // meet once more, and then each of them runs the same passes over the same concatenation (in
      // rank order): identical decisions, no further exchange.
      if (s.done == 0 && tot[kTrCount] <= static_cast<double>(kTrFinishCap)) {
        __shared__ int s_start[kTrMaxBlocks + 1];
        for (int b = tid; b < nb; b += kThreads) s_start[b + 1] = static_cast<int>(__ldcg(g.partials + static_cast<int64_t>(b) * kTrCols + kTrCount));
        if (tid == 0) s_start[0] = 0;
        __syncthreads();
        if (warp == 0) {  // inclusive scan of the lengths: lane l owns a run of consecutive blocks
          const int chunk = (nb + 31) / 32;
          int local = 0;
          for (int k = 0; k < chunk; ++k) {
            const int b = lane * chunk + k;
            if (b < nb) local += s_start[b + 1];
          }
          int incl = local;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
          }
          int run = incl - local;
          for (int k = 0; k < chunk; ++k) {
            const int b = lane * chunk + k;
            if (b < nb) {
              run += s_start[b + 1];
              s_start[b + 1] = run;
            }
          }
        }
        __syncthreads();
        const int ncand = s_start[nb];
        constexpr int64_t kSeg = PeerLayout::kTrCandSegment;
        for (int idx = tid; idx < ncand; idx += kThreads) {
          int lo_b = 0, hi_b = nb;  // the block whose list holds candidate idx: s_start[b] <= idx < s_start[b + 1]
          while (hi_b - lo_b > 1) {
            const int mid = (lo_b + hi_b) >> 1;
            if (s_start[mid] <= idx) lo_b = mid; else hi_b = mid;
          }
          const int64_t src = min(g.total, static_cast<int64_t>(lo_b) * per) + (idx - s_start[lo_b]);
          const unsigned long long key = __ldcg(g.keys + src);
          const double av = __ldcg(g.a + src), bv = __ldcg(g.b + src);
          if (PEER) {
            for (int h = 0; h < peer.world; ++h) {
              double* seg = peer_base(peer, h) + peer.cand_off + kSeg * peer.rank;
              reinterpret_cast<unsigned long long*>(seg)[idx] = key;
              seg[kTrFinishCap + idx] = av;
              seg[2 * kTrFinishCap + idx] = bv;
            }
          } else {
            g.cand_keys[idx] = key;
            g.cand_a[idx] = av;
            g.cand_b[idx] = bv;
          }
        }
        int nseg = 1;
        if (PEER) {
          if (tid == 0)
            for (int h = 0; h < peer.world; ++h) *reinterpret_cast<volatile long long*>(peer_base(peer, h) + peer.cand_off + kSeg * peer.rank + 3 * kTrFinishCap) = ncand;
          __syncthreads();  // all stores of the block are issued before the (system-scope) fence of the barrier
          if (warp == 0) peer_barrier(peer, peer.tr_barrier, g.peer_error);
          nseg = peer.world;
        }
        __syncthreads();
        if (!(PEER && *reinterpret_cast<volatile int32_t*>(g.peer_error) != 0)) {
          while (s.done == 0) {
            const int sh = s.shift;
            const unsigned long long l2 = s.lo;
            const unsigned long long w2 = 1ull << s.width_shift;
#pragma unroll
            for (int j = 0; j < 34; ++j) bins[j * kThreads + tid] = 0.0;
            for (int h = 0; h < nseg; ++h) {  // rank order, then list order: the same on every rank
              const double* seg = PEER ? peer_base(peer, peer.rank) + peer.cand_off + kSeg * h : nullptr;
              const unsigned long long* ck = PEER ? reinterpret_cast<const unsigned long long*>(seg) : g.cand_keys;
              const double* ca = PEER ? seg + kTrFinishCap : g.cand_a;
              const double* cb = PEER ? seg + 2 * kTrFinishCap : g.cand_b;
              const int cnt = PEER ? static_cast<int>(*reinterpret_cast<const volatile long long*>(seg + 3 * kTrFinishCap)) : ncand;
              for (int base = tid; base < cnt; base += kThreads * kTrUnroll) {
                unsigned long long key[kTrUnroll];
                double av[kTrUnroll], bv[kTrUnroll];
#pragma unroll
                for (int u = 0; u < kTrUnroll; ++u) {
                  const int i = base + u * kThreads;
                  key[u] = i < cnt ? __ldcg(ck + i) : 0ull;
                  av[u] = i < cnt ? __ldcg(ca + i) : 0.0;
                  bv[u] = i < cnt ? __ldcg(cb + i) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < kTrUnroll; ++u) {
                  if (key[u] > l2 && key[u] - l2 <= w2) {
                    const int j0 = tr_bin(key[u], l2, sh);
                    bins[j0 * kThreads + tid] += av[u];
                    bins[(17 + j0) * kThreads + tid] += bv[u];
                  }
                }
              }
            }
            __syncthreads();
            tr_block_totals(bins, tot, 34);
            __syncthreads();
            if (tid == 0) tr_pick_compact(tot, &s, sh);
            __syncthreads();
          }
        } else if (tid == 0) {
          s.done = 1;  // a peer never arrived
        }
        __syncthreads();
      }
    }
    tr_publish_or_wait(g, &s, last, round);
    ++round;
    tr_stamp(g);
  }

  // trust_region.cc:345-365, 429-444 (every block computes the same value from the same state)
  double step = 0.0;
  if (!(s.radius_sq == 0.0 || !(s.max_abs_objective > 0.0))) step = s.variable_coef > 0.0 ? sqrt((s.radius_sq - s.fixed_radius_sq) / s.variable_coef) : DBL_MAX;
  if (!JOINT) {
    if (blockIdx.x == 0 && tid == 0) {
      s.step_size = step;
      *g.st = s;
    }
    return;
  }
  // ---- final: objective deltas at the solution (trust_region.cc:929-967) ---------------
  {
    double f0 = 0.0, f1 = 0.0;
    auto sweep = [&](auto kind, int64_t sb, int64_t se) {
      constexpr int KIND = decltype(kind)::value;
      for (int64_t base = sb + tid; base < se; base += kTile) {
        double obj[kTrUnroll], lb[kTrUnroll], ub[kTrUnroll], center[kTrUnroll], w[kTrUnroll], qd[kTrUnroll], x0v[kTrUnroll], coef[kTrUnroll];
#pragma unroll
        for (int u = 0; u < kTrUnroll; ++u) {
          const int64_t i = min(base + static_cast<int64_t>(u) * kThreads, se - 1);
          el.template get_k<KIND>(i, obj[u], lb[u], ub[u], center[u], w[u], qd[u], x0v[u], coef[u], nullptr, nullptr);
        }
#pragma unroll
        for (int u = 0; u < kTrUnroll; ++u) {
          const int64_t i = base + static_cast<int64_t>(u) * kThreads;
          if (i >= se) continue;
          const double sol = projected_value(center[u], obj[u], w[u], lb[u], ub[u], step);
          if (KIND == 0) f0 += obj[u] * (sol - center[u]);
          else f1 += (-obj[u]) * (sol - center[u]);
        }
      }
    };
    const int64_t cut = el.split();
    sweep(std::integral_constant<int, 0>(), beg, min(end, cut));
    sweep(std::integral_constant<int, 1>(), max(beg, min(end, cut)), end);
    __syncthreads();
    bins[0 * kThreads + tid] = f0;
    bins[1 * kThreads + tid] = f1;
  }
  tr_stamp(g);
  last = tr_round<PEER>(g, peer, bins, tot, part, &s_last, 2, round);
  tr_stamp(g);
  if (last && tid == 0) {
    s.step_size = step;
    *g.st = s;
    g.out[0] = s.lag_primal;
    g.out[1] = s.lag_dual;
    g.out[2] = tot[0];
    g.out[3] = tot[1];
    g.out[4] = s.radius;
    g.out[5] = s.dist_primal_sq;
    g.out[6] = s.dist_dual_sq;
  }
}

__global__ void k_tr_init(TrSearchState* st, double radius, const double* maxabs_partials, int nblocks) {
  double mx = 0.0;
  for (int b = threadIdx.x; b < nblocks; b += 32) mx = fmax(mx, maxabs_partials[b]);
  mx = warp_max(mx);
  if (threadIdx.x != 0) return;
  st->lo = 0ull;
  st->done = 0;
  st->pad = 0;
  st->shift = 60;
  st->width_shift = 64;
  st->fixed_radius_sq = 0.0;
  st->variable_coef = 0.0;
  st->radius_sq = radius * radius;
  st->max_abs_objective = mx;
  st->step_size = 0.0;
  st->frozen_a = 0.0;
  st->frozen_b = 0.0;
  st->lag_primal = 0.0;
  st->lag_dual = 0.0;
  st->radius = radius;
  st->dist_primal_sq = 0.0;
  st->dist_dual_sq = 0.0;
}
__global__ void k_tr_finish(TrSearchState* st, double radius) {
  // trust_region.cc:345-365, 429-444
  if (radius == 0.0 || !(st->max_abs_objective > 0.0)) { st->step_size = 0.0; return; }
  st->step_size = st->variable_coef > 0.0 ? sqrt((st->radius_sq - st->fixed_radius_sq) / st->variable_coef) : DBL_MAX;
}


}  // namespace kernels

using namespace kernels;

// ===========================================================================
// Device
// ===========================================================================
int Device::DeviceCount() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

Device::Device(int cuda_device) : device_(cuda_device) {
  int n = DeviceCount();
  if (n <= 0) throw std::runtime_error("no usable CUDA device");
  if (cuda_device < 0 || cuda_device >= n) throw std::runtime_error("CUDA device index out of range");
  CUDA_OK(cudaSetDevice(cuda_device));
  cudaStream_t s;
  CUDA_OK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  stream_ = s;
  {
    // Vectors, matrix images and scratch come from the device's stream-ordered pool with a raised
    // release threshold: after the first solve of a process a new solve maps no fresh device
    // memory (cudaMalloc / cudaFree of hundreds of MB are the slow driver calls of a solve).
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, cuda_device) == cudaSuccess) {
      uint64_t keep = ~uint64_t{0};
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
  }
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, cuda_device));
  num_sms_ = prop.multiProcessorCount;
  CUDA_OK(cudaMalloc(&partials_, sizeof(double) * kMaxReduceBlocks * 40));
  CUDA_OK(cudaMalloc(&tr_peer_error_, 64));
  CUDA_OK(cudaMemset(tr_peer_error_, 0, 64));
  CUDA_OK(cudaMalloc(&tile_counters_, 64));
  CUDA_OK(cudaMemset(tile_counters_, 0, 64));
  CUDA_OK(cudaMalloc(&loop_sync_, 256));
  CUDA_OK(cudaMemset(loop_sync_, 0, 256));
  CUDA_OK(cudaMallocHost(&loop_trace_host_, 8 * sizeof(unsigned long long)));
  CUDA_OK(cudaMalloc(&results_, sizeof(double) * 64));
  CUDA_OK(cudaMallocHost(&host_results_, sizeof(double) * 64));
}

Device::~Device() {
  cudaSetDevice(device_);
  cudaFree(partials_);
  cudaFree(tr_peer_error_);
  cudaFree(tile_counters_);
  cudaFree(loop_sync_);
  cudaFreeHost(loop_trace_host_);
  cudaFree(results_);
  cudaFreeHost(host_results_);
  cudaFree(tr_scratch_);
  cudaFree(step_partials_);
  if (const char* t = std::getenv("PDLP_B200_TRACE"); t != nullptr && t[0] == '1' && detail_samples_ > 0) {
    std::fprintf(stderr, "[pdlp_b200 trace] step sub-phases (us, %lld samples):", static_cast<long long>(detail_samples_));
    for (int k = 0; k < 7; ++k) std::fprintf(stderr, " %.1f", 1000.0 * detail_ms_[k] / detail_samples_);
    std::fprintf(stderr, "\n");
  }
  for (void* e : timing_events_) cudaEventDestroy(static_cast<cudaEvent_t>(e));
  for (auto& pr : timeline_ev_) for (void* e : pr) if (e != nullptr) cudaEventDestroy(static_cast<cudaEvent_t>(e));
  for (void* e : pair_ev_) if (e != nullptr) cudaEventDestroy(static_cast<cudaEvent_t>(e));
  if (stream2_ != nullptr) cudaStreamDestroy(static_cast<cudaStream_t>(stream2_));
  if (stream_ != nullptr) cudaStreamDestroy(static_cast<cudaStream_t>(stream_));
}

#define STREAM static_cast<cudaStream_t>(stream_)
#define LAUNCHED() do { ++launches_; CUDA_OK(cudaGetLastError()); } while (0)

void Device::Sync() { CUDA_OK(cudaStreamSynchronize(STREAM)); }
double* Device::AllocF64(int64_t n) {
  double* p = nullptr;
  CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&p), sizeof(double) * static_cast<size_t>(std::max<int64_t>(n, 1) + 2), STREAM));
  return p;
}
void Device::Free(void* p) { if (p != nullptr) cudaFree(p); }  // (cudaFree also returns pool allocations to their pool)
void Device::Upload(double* dst, const double* src, int64_t n) {
  if (n > 0) { CUDA_OK(cudaMemcpyAsync(dst, src, sizeof(double) * n, cudaMemcpyHostToDevice, STREAM)); Sync(); }
}
void Device::Download(double* dst, const double* src, int64_t n) {
  if (n > 0) { CUDA_OK(cudaMemcpyAsync(dst, src, sizeof(double) * n, cudaMemcpyDeviceToHost, STREAM)); Sync(); }
}
void Device::CopyD2D(double* dst, const double* src, int64_t n) {
  if (n > 0) CUDA_OK(cudaMemcpyAsync(dst, src, sizeof(double) * n, cudaMemcpyDeviceToDevice, STREAM));
}
static inline int Blocks(int64_t n) { return static_cast<int>(std::max<int64_t>(1, (n + kThreads - 1) / kThreads)); }
static inline int ReduceBlocks(int64_t n) { return static_cast<int>(std::min<int64_t>(kMaxReduceBlocks, std::max<int64_t>(1, (n + kThreads * 4 - 1) / (kThreads * 4)))); }

void Device::Fill(double* dst, double value, int64_t n) {
  if (n <= 0) return;
  k_for<<<Blocks(n), kThreads, 0, STREAM>>>(n, [=] __device__(int64_t i) { dst[i] = value; });
  LAUNCHED();
}
int32_t* Device::UploadI32(const std::vector<int32_t>& v) {
  int32_t* p = nullptr;
  CUDA_OK(cudaMalloc(&p, sizeof(int32_t) * (v.size() + 4)));
  if (!v.empty()) CUDA_OK(cudaMemcpy(p, v.data(), sizeof(int32_t) * v.size(), cudaMemcpyHostToDevice));
  return p;
}
void Device::UploadPermuted(double* dst, const double* src_host, const int32_t* row_of_pos, int64_t n) {
  if (n <= 0) return;
  double* tmp = AllocF64(n);
  Upload(tmp, src_host, n);
  k_for<<<Blocks(n), kThreads, 0, STREAM>>>(n, [=] __device__(int64_t p) { dst[p] = tmp[row_of_pos[p]]; });
  LAUNCHED();
  Sync();
  Free(tmp);
}
void Device::DownloadPermuted(double* dst_host, const double* src, const int32_t* row_of_pos, int64_t n) {
  if (n <= 0) return;
  double* tmp = AllocF64(n);
  k_for<<<Blocks(n), kThreads, 0, STREAM>>>(n, [=] __device__(int64_t p) { tmp[row_of_pos[p]] = src[p]; });
  LAUNCHED();
  Download(dst_host, tmp, n);
  Free(tmp);
}

void Device::ScatterInto(double* dst, const double* src, const int32_t* row_of_pos, int64_t n) {
  if (n <= 0) return;
  k_for<<<Blocks(n), kThreads, 0, STREAM>>>(n, [=] __device__(int64_t p) { dst[row_of_pos[p]] = src[p]; });
  LAUNCHED();
}

SellDev Device::UploadSell(const SellHost& h) {
  SellDev d;
  d.num_rows = h.num_rows; d.num_cols = h.num_cols; d.num_split = h.num_split;
  d.num_virtual_padded = h.num_virtual_padded; d.num_slots = h.num_slots; d.padded_nnz = h.padded_nnz;
  auto up = [&](void** dst, const void* src, size_t bytes) {
    CUDA_OK(cudaMalloc(dst, bytes + 64));
    if (bytes > 0) CUDA_OK(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
  };
  up(reinterpret_cast<void**>(&d.slice_ptr), h.slice_ptr.data(), h.slice_ptr.size() * sizeof(int64_t));
  up(reinterpret_cast<void**>(&d.slot_len), h.slot_len.data(), h.slot_len.size() * sizeof(int32_t));
  up(reinterpret_cast<void**>(&d.col), h.col.data(), h.col.size() * sizeof(int32_t));
  up(reinterpret_cast<void**>(&d.val), h.val.data(), h.val.size() * sizeof(double));
  up(reinterpret_cast<void**>(&d.split_first), h.split_first.data(), h.split_first.size() * sizeof(int32_t));
  up(reinterpret_cast<void**>(&d.virt_pos), h.virt_pos.data(), h.virt_pos.size() * sizeof(int32_t));
  CUDA_OK(cudaMalloc(&d.virt_partial, sizeof(double) * (h.num_virtual_padded + 32)));
  return d;
}
void Device::FreeSell(SellDev& s) {
  cudaFree(s.slice_ptr); cudaFree(s.slot_len); cudaFree(s.col); cudaFree(s.val);
  cudaFree(s.split_first); cudaFree(s.virt_pos); cudaFree(s.virt_partial);
  s = SellDev();
}
void Device::DownloadSellValues(const SellDev& s, std::vector<double>& out) {
  out.resize(s.padded_nnz);
  if (s.padded_nnz > 0) CUDA_OK(cudaMemcpy(out.data(), s.val, sizeof(double) * s.padded_nnz, cudaMemcpyDeviceToHost));
}

// ---- generic SELL launch --------------------------------------------------
namespace kernels {
// Launch shape of k_sell, tunable for A/B runs: block size (PDLP_B200_SELL_THREADS),
// row-loop variant (PDLP_B200_SELL_VARIANT) and slot groups per block (PDLP_B200_SELL_CHUNKS).
int SellThreads() {
  static const int bt = [] {
    const char* v = std::getenv("PDLP_B200_SELL_THREADS");
    const int t = (v != nullptr && *v != 0) ? std::atoi(v) : 128;
    return t == 256 ? t : 128;
  }();
  return bt;
}
int SellVariant() {
  static const int v = [] {
    const char* e = std::getenv("PDLP_B200_SELL_VARIANT");
    return (e != nullptr && *e != 0) ? std::atoi(e) : 1;
  }();
  return v;
}
int SellCarveout() {  // PDLP_B200_SELL_CARVEOUT: preferred shared-memory carve-out (percent) of the TMA-staged kernels; -1 = driver default
  static const int v = [] {
    const char* e = std::getenv("PDLP_B200_SELL_CARVEOUT");
    return (e != nullptr && *e != 0) ? std::atoi(e) : -1;
  }();
  return v;
}
int SellChunks() {
  static const int v = [] {
    const char* e = std::getenv("PDLP_B200_SELL_CHUNKS");
    const int c = (e != nullptr && *e != 0) ? std::atoi(e) : 2;
    return std::max(1, std::min(64, c));
  }();
  return v;
}
// Tiles of a k_sell launch: one per `chunks` groups of SellThreads() slots (an ordinary launch has one block per tile).
int SellGridFor(const SellDev& a, int chunks) {
  const int64_t per_block = static_cast<int64_t>(SellThreads()) * chunks;
  return static_cast<int>(std::max<int64_t>(1, (a.num_slots + per_block - 1) / per_block));
}
int SellGrid(const SellDev& a) { return SellGridFor(a, SellChunks()); }
// Step loop only: persistent SpMV launches with a tile queue (PDLP_B200_SELL_PERSIST=1) and their groups per tile.
// Off by default: measured equal to the ordinary launches (220.3 vs 220.5 us per iteration on C2,
// profiles/r02s_ab_persistent.txt) -- the hardware block scheduler plus programmatic dependent launch
// already keep the SMs fed through the last wave; the kernels are bound by the gather request rate.
bool SellPersist() {
  static const bool v = [] { const char* e = std::getenv("PDLP_B200_SELL_PERSIST"); return e != nullptr && e[0] == '1'; }();
  return v;
}
int SellPersistChunks() {
  static const int v = [] {
    const char* e = std::getenv("PDLP_B200_SELL_PERSIST_CHUNKS");
    const int c = (e != nullptr && *e != 0) ? std::atoi(e) : 1;
    return std::max(1, std::min(64, c));
  }();
  return v;
}
int SmCount() {
  static const int v = [] {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
  }();
  return v;
}
// Launch with (pdl) or without the programmatic-dependent-launch attribute.
template <class... KArgs, class... Args>
void launch_k(bool pdl, void (*kernel)(KArgs...), int grid, int block, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
}
// PDLP_B200_PEER_LOOP: 0 the row-sharded step loop as separate launches per attempt, 2 always k_peer_loop, 1 (default) by size.
int PeerLoop() {
  static const int v = [] { const char* e = std::getenv("PDLP_B200_PEER_LOOP"); return (e != nullptr && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1; }();
  return v;
}
bool StepPdl() {
  static const bool v = [] { const char* e = std::getenv("PDLP_B200_PDL"); return !(e != nullptr && e[0] == '0'); }();
  return v;
}

template <int MODE, int NS, class Epi>
void launch_sell(cudaStream_t stream, const SellDev& a, GatherSrc x, Epi epi, double* partials, const int32_t* halt, int64_t* launches,
                 int* main_blocks, int* fix_blocks, bool pdl = false, DecideHead head = DecideHead(), TileQueue queue = TileQueue(), int chunks = 0) {
  if (chunks <= 0) chunks = SellChunks();
  const int nb = SellGridFor(a, chunks);
  queue.num_tiles = nb;
  // the dual kernel (two reductions, a 6-operand epilogue) is 4 % faster with 72 registers at 7 blocks
  // per SM than squeezed into 64 at 8; the store-only kernels prefer the 8 blocks (profiles/r02m_ab.txt)
  const int variant = (SellVariant() == 1 && NS == 2) ? 8 : SellVariant();
  const DecideHead& main_tail = head;  // (an extra block of the main kernel takes the step decision when asked to)
  // variants: 0 plain, 1 / 2 software-pipelined register staging (4 / 8 slots; 8 = 1 with 72 registers, 7 blocks per SM), 3.. TMA staging
  // (slots per stage x stages) 3: 8 x 2, 4: 4 x 2, 5: 2 x 2, 6: 2 x 4, 7: 4 x 3. Shared memory used for
  // staging is taken from the L1 that tracks the outstanding gather misses, so the stages are small
  // and the carve-out is pinned to what the resident blocks need.
  const int grid = nb + (head.in != nullptr ? 1 : 0);  // (+ the block that only takes the step decision)
#define PDLP_SELL_LAUNCH(BT, V)                                                                                                       \
  do {                                                                                                                                \
    int g__ = grid;                                                                                                                   \
    if (queue.counter != nullptr) {  /* persistent: as many workers as are resident at once (+ the decision block) */                 \
      int per_sm__ = 0;                                                                                                               \
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm__, k_sell<MODE, NS, Epi, BT, V>, BT, 0);                                  \
      g__ = std::min(nb, std::max(1, per_sm__) * SmCount()) + (head.in != nullptr ? 1 : 0);                                           \
    }                                                                                                                                 \
    launch_k(pdl, k_sell<MODE, NS, Epi, BT, V>, g__, BT, stream, a, x, epi, partials, halt, chunks, main_tail, queue);                \
  } while (0)
#define PDLP_SELL_LAUNCH_T(BT, U, NST)                                                                                              \
  do {                                                                                                                              \
    static bool carved = false;                                                                                                     \
    if (!carved) {                                                                                                                  \
      carved = true;                                                                                                                \
      if (const int pct = SellCarveout(); pct >= 0)                                                                                 \
        cudaFuncSetAttribute(k_sell_tma<MODE, NS, Epi, BT, U, NST>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);          \
    }                                                                                                                               \
    launch_k(pdl, k_sell_tma<MODE, NS, Epi, BT, U, NST>, grid, BT, stream, a, x, epi, partials, halt, chunks, main_tail);           \
  } while (0)
#define PDLP_SELL_LAUNCH_V(BT)                                 \
  do {                                                         \
    if (variant == 3) PDLP_SELL_LAUNCH_T(128, 8, 2);           \
    else if (variant == 4) PDLP_SELL_LAUNCH_T(128, 4, 2);      \
    else if (variant == 5) PDLP_SELL_LAUNCH_T(128, 2, 2);      \
    else if (variant == 6) PDLP_SELL_LAUNCH_T(128, 2, 4);      \
    else if (variant == 7) PDLP_SELL_LAUNCH_T(128, 4, 3);      \
    else if (variant == 1) PDLP_SELL_LAUNCH(BT, 1);            \
    else if (variant == 8) PDLP_SELL_LAUNCH(BT, 8);            \
    else if (variant == 2) PDLP_SELL_LAUNCH(BT, 2);            \
    else PDLP_SELL_LAUNCH(BT, 0);                              \
  } while (0)
  switch (SellThreads()) {
    case 256:  // (static shared memory: 8 warps only fit the small stagings)
      if (variant >= 3) PDLP_SELL_LAUNCH_T(256, 2, 2);
      else PDLP_SELL_LAUNCH_V(256);
      break;
    default: PDLP_SELL_LAUNCH_V(128); break;
  }
#undef PDLP_SELL_LAUNCH_V
#undef PDLP_SELL_LAUNCH_T
#undef PDLP_SELL_LAUNCH
  ++*launches;
  int nf = 0;
  if (a.num_split > 0) {
    nf = static_cast<int>((a.num_split * 32 + kThreads - 1) / kThreads);
    launch_k(pdl, k_sell_fixup<MODE, NS, Epi>, nf, kThreads, stream, a, epi, partials != nullptr ? partials + static_cast<int64_t>(nb) * NS : nullptr, halt);
    ++*launches;
  }
  if (main_blocks != nullptr) *main_blocks = nb;
  if (fix_blocks != nullptr) *fix_blocks = nf;
}
struct NoCtx {};
struct StoreEpi {
  double* out;
  typedef NoCtx Ctx;
  struct Pre {};
  __device__ __forceinline__ Ctx begin() const { return Ctx(); }
  __device__ __forceinline__ Pre prefetch(const Ctx&, int64_t) const { return Pre(); }
  __device__ __forceinline__ void operator()(const Ctx&, int64_t pos, double acc, double*, const Pre&) const { out[pos] = acc; }
};
struct ScatterEpi {  // out[perm[pos]] = acc
  double* out;
  const int32_t* perm;
  typedef NoCtx Ctx;
  struct Pre { int32_t dst; };
  __device__ __forceinline__ Ctx begin() const { return Ctx(); }
  __device__ __forceinline__ Pre prefetch(const Ctx&, int64_t pos) const { return Pre{__ldg(perm + pos)}; }
  __device__ __forceinline__ void operator()(const Ctx&, int64_t, double acc, double*, const Pre& p) const { out[p.dst] = acc; }
};
struct NormEpi {
  double* out;
  const double* own;
  int l2;
  typedef NoCtx Ctx;
  struct Pre { double own; };
  __device__ __forceinline__ Ctx begin() const { return Ctx(); }
  __device__ __forceinline__ Pre prefetch(const Ctx&, int64_t pos) const { return Pre{own[pos]}; }
  __device__ __forceinline__ void operator()(const Ctx&, int64_t pos, double acc, double*, const Pre& p) const { out[pos] = (l2 ? sqrt(acc) : acc) * fabs(p.own); }
};
}  // namespace kernels

void Device::SpMV(const SellDev& a, const double* x, double* out) {
  if (a.num_rows <= 0) return;
  launch_sell<kDot, 0>(STREAM, a, GatherSrc{{x, nullptr, nullptr}, nullptr}, StoreEpi{out}, nullptr, nullptr, &launches_, nullptr, nullptr);
  CUDA_OK(cudaGetLastError());
}
void Device::SpMVScatter(const SellDev& a, const double* x, const int32_t* perm, double* out) {
  if (a.num_rows <= 0) return;
  launch_sell<kDot, 0>(STREAM, a, GatherSrc{{x, nullptr, nullptr}, nullptr}, ScatterEpi{out, perm}, nullptr, nullptr, &launches_, nullptr, nullptr);
  CUDA_OK(cudaGetLastError());
}
void Device::RowNormRawScatter(const SellDev& a, int norm, const double* other_scale, const int32_t* perm, double* out) {
  if (a.num_rows <= 0) return;
  if (norm == 0) launch_sell<kMaxAbs, 0>(STREAM, a, GatherSrc{{other_scale, nullptr, nullptr}, nullptr}, ScatterEpi{out, perm}, nullptr, nullptr, &launches_, nullptr, nullptr);
  else launch_sell<kSumSq, 0>(STREAM, a, GatherSrc{{other_scale, nullptr, nullptr}, nullptr}, ScatterEpi{out, perm}, nullptr, nullptr, &launches_, nullptr, nullptr);
  CUDA_OK(cudaGetLastError());
}
void Device::FinishRowNorm(double* out, int norm, const double* own_scale, int64_t n) {
  if (n <= 0) return;
  k_for<<<Blocks(n), kThreads, 0, STREAM>>>(n, [=] __device__(int64_t i) { out[i] = (norm ? sqrt(out[i]) : out[i]) * fabs(own_scale[i]); });
  LAUNCHED();
}
bool Device::count_primal() const { return comm_ == nullptr || comm_->rank() == 0; }
double Device::RootValue(double v) {
  if (comm_ == nullptr) return v;
  host_results_[0] = comm_->rank() == 0 ? v : 0.0;
  CUDA_OK(cudaMemcpyAsync(results_, host_results_, sizeof(double), cudaMemcpyHostToDevice, STREAM));
  comm_->AllReduceSum(results_, results_, 1, stream_);
  CUDA_OK(cudaMemcpyAsync(host_results_, results_, sizeof(double), cudaMemcpyDeviceToHost, STREAM));
  Sync();
  return host_results_[0];
}
double Device::MaxOverRanks(double v) {
  if (comm_ == nullptr) return v;
  host_results_[0] = v;
  CUDA_OK(cudaMemcpyAsync(results_, host_results_, sizeof(double), cudaMemcpyHostToDevice, STREAM));
  comm_->AllReduceMax(results_, results_, 1, stream_);
  CUDA_OK(cudaMemcpyAsync(host_results_, results_, sizeof(double), cudaMemcpyDeviceToHost, STREAM));
  Sync();
  return host_results_[0];
}
void Device::AllReduceSumVec(double* buf, int64_t n) { if (comm_ != nullptr) comm_->AllReduceSum(buf, buf, n, stream_); }
void Device::AllReduceMaxVec(double* buf, int64_t n) { if (comm_ != nullptr) comm_->AllReduceMax(buf, buf, n, stream_); }

void Device::ScaledRowNorm(const SellDev& a, int norm, const double* other_scale, const double* own_scale, double* out) {
  if (a.num_rows <= 0) return;
  if (norm == 0) launch_sell<kMaxAbs, 0>(STREAM, a, GatherSrc{{other_scale, nullptr, nullptr}, nullptr}, NormEpi{out, own_scale, 0}, nullptr, nullptr, &launches_, nullptr, nullptr);
  else launch_sell<kSumSq, 0>(STREAM, a, GatherSrc{{other_scale, nullptr, nullptr}, nullptr}, NormEpi{out, own_scale, 1}, nullptr, nullptr, &launches_, nullptr, nullptr);
  CUDA_OK(cudaGetLastError());
}
void Device::ScaleMatrix(SellDev& a, const double* own_scale, const double* other_scale, const int32_t* own_perm) {
  if (a.num_slots <= 0) return;
  const SellDev s = a;
  k_for<<<Blocks(s.num_slots), kThreads, 0, STREAM>>>(s.num_slots, [=] __device__(int64_t slot) {
    const int n = s.slot_len[slot];
    if (n == 0) return;
    const int64_t pos = slot < s.num_virtual_padded ? s.virt_pos[slot] : s.num_split + (slot - s.num_virtual_padded);
    const double r = own_scale[own_perm != nullptr ? own_perm[pos] : pos];
    const int64_t base = s.slice_ptr[slot >> 5] + (slot & 31);
    for (int j = 0; j < n; ++j) {
      const int64_t k = base + static_cast<int64_t>(j) * 32;
      s.val[k] *= r * other_scale[s.col[k]];
    }
  });
  LAUNCHED();
}

#define ELEMENTWISE(n, ...)                                                                   \
  do {                                                                                        \
    if ((n) > 0) {                                                                            \
      k_for<<<Blocks(n), kThreads, 0, STREAM>>>((n), [=] __device__(int64_t i) __VA_ARGS__);  \
      LAUNCHED();                                                                             \
    }                                                                                         \
  } while (0)

void Device::WriteGlobalRowPositions(double* out_full, const int32_t* row_of_pos, int64_t row_begin, int64_t m) {
  ELEMENTWISE(m, { out_full[row_begin + row_of_pos[i]] = static_cast<double>(row_begin + i); });
}
void Device::DoublesToI32(int32_t* dst, const double* src, int64_t n) { ELEMENTWISE(n, { dst[i] = static_cast<int32_t>(src[i]); }); }
void Device::DivideBySqrt(double* vec, const double* divisor, int64_t n) { ELEMENTWISE(n, { if (divisor[i] != 0) vec[i] /= sqrt(divisor[i]); }); }
void Device::Mul(double* dst, const double* a, int64_t n) { ELEMENTWISE(n, { dst[i] = dst[i] * a[i]; }); }
void Device::Div(double* dst, const double* a, int64_t n) { ELEMENTWISE(n, { dst[i] = dst[i] / a[i]; }); }
void Device::MulSq(double* dst, const double* a, int64_t n) { ELEMENTWISE(n, { dst[i] = dst[i] * (a[i] * a[i]); }); }
void Device::Axpy(double* dst, double s, const double* a, int64_t n) { ELEMENTWISE(n, { dst[i] += s * a[i]; }); }
void Device::Sub(double* dst, const double* a, const double* b, int64_t n) { ELEMENTWISE(n, { dst[i] = a[i] - b[i]; }); }
void Device::ReplaceLargeWithInf(double* v, double threshold, int64_t n) {
  ELEMENTWISE(n, { if (v[i] <= -threshold) v[i] = -kInfD; if (v[i] >= threshold) v[i] = kInfD; });
}
void Device::MapFiniteValuesToZero(double* dst, const double* src, int64_t n) { ELEMENTWISE(n, { dst[i] = isfinite(src[i]) ? 0.0 : src[i]; }); }
void Device::DualTrustRegionProblem(const double* g, const double* lc, const double* uc, double* objective, double* lb, double* ub, int64_t m) {
  ELEMENTWISE(m, {
    objective[i] = -g[i];
    lb[i] = isfinite(uc[i]) ? -kInfD : 0.0;
    ub[i] = isfinite(lc[i]) ? kInfD : 0.0;
  });
}
void Device::ClampPrimal(double* x, const double* lb, const double* ub, bool feas, int64_t n) {
  ELEMENTWISE(n, {
    double u = ub[i], l = lb[i];
    if (feas) { u = isfinite(u) ? 0.0 : u; l = isfinite(l) ? 0.0 : l; }
    x[i] = fmax(fmin(x[i], u), l);
  });
}
void Device::ClampDual(double* y, const double* lc, const double* uc, int64_t m) {
  ELEMENTWISE(m, {
    double v = y[i];
    if (!isfinite(uc[i])) v = fmax(v, 0.0);
    if (!isfinite(lc[i])) v = fmin(v, 0.0);
    y[i] = v;
  });
}
void Device::WeightedAverageAdd(double* avg, const double* v, double ratio, int64_t n) { ELEMENTWISE(n, { avg[i] += ratio * (v[i] - avg[i]); }); }
void Device::PrimalGradient(const double* x, const double* kty, const double* c, const double* q, bool zero_objective, double* out, int64_t n) {
  ELEMENTWISE(n, {
    if (zero_objective) out[i] = -kty[i];
    else out[i] = (q != nullptr ? q[i] * x[i] : 0.0) + (c[i] - kty[i]);
  });
}
void Device::PrimalStep(const double* x, const double* kty, const double* c, const double* q, const double* lv, const double* uv, double tau,
                        double* x_next, int64_t n) {
  ELEMENTWISE(n, {
    double t = x[i] - tau * (c[i] - kty[i]);
    if (q != nullptr) t = t / (tau * q[i] + 1.0);
    x_next[i] = fmax(fmin(t, uv[i]), lv[i]);
  });
}
void Device::DualStepFromProducts(const double* y, const double* kx_cur, const double* kx_next, const double* lc, const double* uc, double sigma,
                                  double theta, double* y_next, int64_t m) {
  ELEMENTWISE(m, {  // pdhg.cc:1905-1910, 1923-1928
    const double t = y[i] - sigma * (-theta * kx_cur[i] + (theta + 1) * kx_next[i]);
    y_next[i] = fmax(fmin(0.0, t + sigma * uc[i]), t + sigma * lc[i]);
  });
}

// ---- reductions -------------------------------------------------------------
// Where the next reduction leaves its results. Outside a batch: the start of results_, copied to
// the host and synchronised at once (host_results_[0..)). Inside a batch (BeginBatch .. EndBatch)
// every reduction gets its own range, nothing is copied until EndBatch makes ONE device->host copy
// and ONE synchronisation for all of them; the *Launch functions return the offset to read at.
double* Device::ReduceTarget(int count) {
  if (!batch_active_) { last_result_off_ = 0; return results_; }
  if (batch_off_ + count > 64) throw std::runtime_error("reduction batch overflows the result buffer");
  last_result_off_ = batch_off_;
  batch_off_ += (count + 1) / 2 * 2;
  return results_ + last_result_off_;
}
void Device::BeginBatch() {
  batch_active_ = true;
  batch_off_ = 0;
}
void Device::EndBatch() {
  batch_active_ = false;
  if (batch_off_ > 0) CUDA_OK(cudaMemcpyAsync(host_results_, results_, sizeof(double) * batch_off_, cudaMemcpyDeviceToHost, STREAM));
  Sync();
}
// `sharded`: the reduced elements are row-sharded across ranks (dual side), so
// the sums / maxes are completed by an all-reduce; replicated (primal side)
// reductions are identical on every rank and need none. Joint reductions count
// their primal part on rank 0 only (count_primal()) and are sharded.
#define REDUCE_S(sharded, NS, NM, n, ...)                                                                                         \
  do {                                                                                                                            \
    const int nb__ = ReduceBlocks(n);                                                                                             \
    k_reduce<NS, NM><<<nb__, kThreads, 0, STREAM>>>((n), [=] __device__(int64_t i, double* s, double* m) __VA_ARGS__, partials_); \
    LAUNCHED();                                                                                                            \
    double* res__ = ReduceTarget((NS) + (NM));                                                                             \
    k_reduce_final<NS, NM><<<1, kThreads, 0, STREAM>>>(nb__, partials_, res__);                                            \
    LAUNCHED();                                                                                                            \
    if ((sharded) && comm_ != nullptr) {                                                                                   \
      if ((NS) > 0) comm_->AllReduceSum(res__, res__, (NS), stream_);                                                      \
      if ((NM) > 0) comm_->AllReduceMax(res__ + (NS), res__ + (NS), (NM), stream_);                                        \
    }                                                                                                                      \
    if (!batch_active_) {                                                                                                  \
      CUDA_OK(cudaMemcpyAsync(host_results_, results_, sizeof(double) * ((NS) + (NM)), cudaMemcpyDeviceToHost, STREAM));   \
      Sync();                                                                                                              \
    }                                                                                                                      \
  } while (0)
#define REDUCE(NS, NM, n, ...) REDUCE_S(false, NS, NM, n, __VA_ARGS__)
// Reduction over replicated primal-length vectors: on a row-sharded solve every
// rank reduces its slice [PrimalSliceBegin, PrimalSliceEnd) and the results are
// all-reduced (same value everywhere); otherwise the whole range.
#define REDUCE_P(NS, NM, n, ...)                                                                                                  \
  do {                                                                                                                            \
    const int64_t off__ = PrimalSliceBegin(n), len__ = PrimalSliceEnd(n) - off__;                                                 \
    const int nb__ = ReduceBlocks(len__);                                                                                         \
    k_reduce<NS, NM><<<nb__, kThreads, 0, STREAM>>>(len__, [=] __device__(int64_t i__, double* s, double* m) { const int64_t i = i__ + off__; __VA_ARGS__ }, partials_); \
    LAUNCHED();                                                                                                            \
    double* res__ = ReduceTarget((NS) + (NM));                                                                             \
    k_reduce_final<NS, NM><<<1, kThreads, 0, STREAM>>>(nb__, partials_, res__);                                            \
    LAUNCHED();                                                                                                            \
    if (PrimalSliced(n)) {                                                                                                 \
      if ((NS) > 0) comm_->AllReduceSum(res__, res__, (NS), stream_);                                                      \
      if ((NM) > 0) comm_->AllReduceMax(res__ + (NS), res__ + (NS), (NM), stream_);                                        \
    }                                                                                                                      \
    if (!batch_active_) {                                                                                                  \
      CUDA_OK(cudaMemcpyAsync(host_results_, results_, sizeof(double) * ((NS) + (NM)), cudaMemcpyDeviceToHost, STREAM));   \
      Sync();                                                                                                              \
    }                                                                                                                      \
  } while (0)

double Device::Dot(const double* a, const double* b, int64_t n, bool sharded) { REDUCE_S(sharded, 1, 0, n, { s[0] += a[i] * b[i]; }); return host_results_[0]; }
double Device::SumSq(const double* a, int64_t n, bool sharded) { REDUCE_S(sharded, 1, 0, n, { s[0] += a[i] * a[i]; }); return host_results_[0]; }
double Device::SumSqDiff(const double* a, const double* b, int64_t n, bool sharded) { REDUCE_S(sharded, 1, 0, n, { const double d = a[i] - b[i]; s[0] += d * d; }); return host_results_[0]; }
double Device::LInf(const double* a, int64_t n, bool sharded) { REDUCE_S(sharded, 0, 1, n, { m[0] = fmax(m[0], fabs(a[i])); }); return std::max(0.0, host_results_[0]); }
double Device::L1(const double* a, int64_t n, bool sharded) { REDUCE_S(sharded, 1, 0, n, { s[0] += fabs(a[i]); }); return host_results_[0]; }
double Device::ScaledLInf(const double* a, const double* sc, int64_t n, bool sharded) { REDUCE_S(sharded, 0, 1, n, { m[0] = fmax(m[0], fabs(a[i] * sc[i])); }); return std::max(0.0, host_results_[0]); }
double Device::ScaledSumSq(const double* a, const double* sc, int64_t n, bool sharded) { REDUCE_S(sharded, 1, 0, n, { const double t = a[i] * sc[i]; s[0] += t * t; }); return host_results_[0]; }
void Device::DistancesSq(const double* x, const double* x0, int64_t n, const double* y, const double* y0, int64_t mm, double out[2]) {
  const int64_t pbeg = PrimalSliceBegin(n), plen = PrimalSliceEnd(n) - pbeg;
  const int64_t total = plen + mm;
  REDUCE_S(true, 2, 0, total, {
    if (i < plen) { const double d = x[pbeg + i] - x0[pbeg + i]; s[0] += d * d; }
    else { const int64_t j = i - plen; const double d = y[j] - y0[j]; s[1] += d * d; }
  });
  out[0] = host_results_[0];
  out[1] = host_results_[1];
}

namespace kernels {
__device__ __forceinline__ void info_add(double v, double* s, double* m) {  // VectorInfoAccumulator::Add, sou.cc:130-143
  if (isinf(v)) { s[1] += 1.0; }
  else if (v == 0) { s[2] += 1.0; }
  else {
    const double a = fabs(v);  // NaN lands here and poisons sum / sumsq like the reference
    s[0] += 1.0; s[3] += a; s[4] += a * a;
    m[0] = fmax(m[0], a); m[1] = fmax(m[1], -a);
  }
}
__device__ __forceinline__ double combine_bounds(double v1, double v2) {  // sou.cc:83-92
  double mx = 0.0;
  if (fabs(v1) < kInfD) mx = fabs(v1);
  if (fabs(v2) < kInfD) mx = fmax(mx, fabs(v2));
  return mx;
}
}  // namespace kernels

static VectorInfoDev InfoFromHost(const double* r) {
  VectorInfoDev v;
  v.num_finite_nonzero = r[0]; v.num_infinite = r[1]; v.num_zero = r[2]; v.sum = r[3]; v.sumsq = r[4];
  v.largest = r[5]; v.smallest = -r[6];
  return v;
}
VectorInfoDev Device::VectorInfo(const double* v, int64_t n, bool sharded) { REDUCE_S(sharded, 5, 2, n, { info_add(v[i], s, m); }); return InfoFromHost(host_results_); }
VectorInfoDev Device::CombinedBoundsInfo(const double* a, const double* b, int64_t n, bool sharded) { REDUCE_S(sharded, 5, 2, n, { info_add(combine_bounds(a[i], b[i]), s, m); }); return InfoFromHost(host_results_); }
VectorInfoDev Device::GapInfo(const double* lb, const double* ub, int64_t n) { REDUCE(5, 2, n, { info_add(ub[i] - lb[i], s, m); }); return InfoFromHost(host_results_); }
VectorInfoDev Device::MatrixInfo(const SellDev& a) {
  const SellDev sd = a;
  REDUCE_S(true, 5, 2, sd.num_slots, {
    const int n = sd.slot_len[i];
    const int64_t base = sd.slice_ptr[i >> 5] + (i & 31);
    for (int j = 0; j < n; ++j) info_add(sd.val[base + static_cast<int64_t>(j) * 32], s, m);
  });
  return InfoFromHost(host_results_);
}
bool Device::BoundsValid(const double* lb, const double* ub, int64_t n, bool sharded) {  // sou.cc:701-723
  REDUCE_S(sharded, 1, 0, n, { if (!(lb[i] <= ub[i] && lb[i] < kInfD && ub[i] > -kInfD)) s[0] += 1.0; });
  return host_results_[0] == 0.0;
}
bool Device::AllNonNegative(const double* v, int64_t n) {
  REDUCE(1, 0, n, { if (!(v[i] >= 0.0)) s[0] += 1.0; });
  return host_results_[0] == 0.0;
}

MSideStats Device::DualSideStats(const double* y, const double* kx, const double* lc, const double* uc, const double* dr, double cw_offset,
                                 bool homogeneous, int64_t mm) {
  return ReadDualSideStats(DualSideStatsLaunch(y, kx, lc, uc, dr, cw_offset, homogeneous, mm));
}
int Device::DualSideStatsLaunch(const double* y, const double* kx, const double* lc, const double* uc, const double* dr, double cw_offset,
                                bool homogeneous, int64_t mm) {
  REDUCE_S(true, 3, 3, mm, {  // iteration_stats.cc:66-134, 328-350
    const double rs = dr != nullptr ? dr[i] : 1.0;
    const double ub = (homogeneous && isfinite(uc[i])) ? 0.0 : uc[i];
    const double lb = (homogeneous && isfinite(lc[i])) ? 0.0 : lc[i];
    const double v = kx[i];
    double scaled_residual = 0.0, residual_bound = 0.0;
    if (v > ub) { scaled_residual = v - ub; residual_bound = ub; }
    else if (v < lb) { scaled_residual = lb - v; residual_bound = lb; }
    const double residual = scaled_residual / rs;
    m[0] = fmax(m[0], residual);
    s[0] += residual * residual;
    if (residual > 0.0) m[1] = fmax(m[1], residual / (cw_offset + fabs(residual_bound / rs)));
    const double yi = y[i];
    if (yi > 0.0) s[1] += lc[i] * yi; else if (yi < 0.0) s[1] += uc[i] * yi;
    const double ys = yi * rs;
    m[2] = fmax(m[2], fabs(ys));
    s[2] += ys * ys;
  });
  return last_result_off_;
}
MSideStats Device::ReadDualSideStats(int off) const {
  const double* h = host_results_ + off;
  MSideStats r;
  r.sumsq_residual = h[0]; r.bounds_term = h[1]; r.sumsq_scaled = h[2];
  r.linf_residual = std::max(0.0, h[3]); r.cw_residual = std::max(0.0, h[4]); r.linf_scaled = std::max(0.0, h[5]);
  return r;
}

NSideStats Device::PrimalSideStats(const double* x, const double* xb, const double* kty, const double* c, const double* q, const double* lv,
                                   const double* uv, const double* dc, double cw_offset, bool zero_objective, bool handle_as_residuals, int64_t n) {
  return ReadPrimalSideStats(PrimalSideStatsLaunch(x, xb, kty, c, q, lv, uv, dc, cw_offset, zero_objective, handle_as_residuals, n));
}
int Device::PrimalSideStatsLaunch(const double* x, const double* xb, const double* kty, const double* c, const double* q, const double* lv,
                                  const double* uv, const double* dc, double cw_offset, bool zero_objective, bool handle_as_residuals, int64_t n) {
  REDUCE_P(6, 4, n, {  // iteration_stats.cc:189-270, 273-323
    const double cs = dc != nullptr ? dc[i] : 1.0;
    const double xi = x[i];
    const double qx = q != nullptr ? q[i] * xi : 0.0;
    const double g = zero_objective ? -kty[i] : qx + (c[i] - kty[i]);
    if (g != 0.0) {
      const double ub = uv[i], lb = lv[i], xv = xb[i];
      const double bound_for_rc = g > 0.0 ? lb : ub;
      s[1] += bound_for_rc * g;
      const bool lfin = handle_as_residuals ? (fabs(xv - lb) <= fabs(xv)) : isfinite(lb);
      const bool ufin = handle_as_residuals ? (fabs(xv - ub) <= fabs(xv)) : isfinite(ub);
      const double elb = lfin ? lb : -kInfD, eub = ufin ? ub : kInfD;
      const double primary = g >= 0.0 ? elb : eub, secondary = g >= 0.0 ? eub : elb;
      const double bound = isfinite(primary) ? primary : (isfinite(secondary) ? secondary : 0.0);
      s[0] += bound * g;
      const double eff = g > 0.0 ? elb : eub;
      if (isinf(eff)) {
        const double residual = fabs(g) / cs;
        m[0] = fmax(m[0], residual);
        s[2] += residual * residual;
        if (residual > 0.0) m[1] = fmax(m[1], residual / (cw_offset + fabs(c[i] / cs)));
      }
    }
    s[3] += c[i] * xi;
    s[4] += qx * xi;
    const double xs = xi * cs;
    m[2] = fmax(m[2], fabs(xs));
    s[5] += xs * xs;
    m[3] = fmax(m[3], fabs(qx));
  });
  return last_result_off_;
}
NSideStats Device::ReadPrimalSideStats(int off) const {
  const double* h = host_results_ + off;
  NSideStats r;
  r.correction = h[0]; r.full_correction = h[1]; r.sumsq_residual = h[2];
  r.objective_dot = h[3]; r.quadratic = h[4]; r.sumsq_scaled = h[5];
  r.linf_residual = std::max(0.0, h[6]); r.cw_residual = std::max(0.0, h[7]);
  r.linf_scaled = std::max(0.0, h[8]); r.linf_qx = std::max(0.0, h[9]);
  return r;
}

double Device::LagrangianPrimalGradient(const double* x, const double* kty, const double* c, const double* q, double* grad, int64_t n) {
  REDUCE(1, 0, n, {  // sou.cc:446-474
    if (q == nullptr) {
      const double g = c[i] - kty[i];
      grad[i] = g;
      s[0] += x[i] * g;
    } else {
      const double op = q[i] * x[i];
      const double g = c[i] + op - kty[i];
      grad[i] = g;
      s[0] += x[i] * (g - 0.5 * op);
    }
  });
  return host_results_[0];
}
double Device::LagrangianDualGradient(const double* y, const double* kx, const double* lc, const double* uc, double* grad, int64_t mm) {
  JointElem el{nullptr, y, kx, nullptr, nullptr, nullptr, nullptr, nullptr, lc, uc, 1.0, 0, mm, 0};
  REDUCE_S(true, 1, 0, mm, {  // sou.cc:502-527
    const double coef = el.subgradient_coefficient(i);
    s[0] += coef * y[i];
    grad[i] = coef - kx[i];
  });
  return host_results_[0];
}
void Device::ActiveSetPrimal(const double* x, const double* x0, const double* lv, const double* uv, int64_t n, int64_t out[2]) {
  ReadCounts(ActiveSetPrimalLaunch(x, x0, lv, uv, n), out);
}
void Device::ReadCounts(int off, int64_t out[2]) const {
  out[0] = static_cast<int64_t>(host_results_[off]);
  out[1] = static_cast<int64_t>(host_results_[off + 1]);
}
int Device::ActiveSetPrimalLaunch(const double* x, const double* x0, const double* lv, const double* uv, int64_t n) {
  REDUCE_P(2, 0, n, {
    const bool a = x[i] > lv[i] && x[i] < uv[i];
    const bool b = x0[i] > lv[i] && x0[i] < uv[i];
    s[0] += a ? 1.0 : 0.0;
    s[1] += (a != b) ? 1.0 : 0.0;
  });
  return last_result_off_;
}
void Device::ActiveSetDual(const double* y, const double* y0, const double* lc, const double* uc, int64_t mm, int64_t out[2]) {
  ReadCounts(ActiveSetDualLaunch(y, y0, lc, uc, mm), out);
}
int Device::ActiveSetDualLaunch(const double* y, const double* y0, const double* lc, const double* uc, int64_t mm) {
  REDUCE_S(true, 2, 0, mm, {
    const bool free_row = lc[i] == -kInfD && uc[i] == kInfD;
    const bool a = y[i] != 0.0 || free_row;
    const bool b = y0[i] != 0.0 || free_row;
    s[0] += a ? 1.0 : 0.0;
    s[1] += (a != b) ? 1.0 : 0.0;
  });
  return last_result_off_;
}

namespace kernels {
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
// Counter-based standard normal (Box-Muller on two hashed uniforms). The
// reference's values depend on its shard count and are unpinned by its tests
// (solvers.proto:416-422), so any fixed Gaussian stream is conforming.
__device__ __forceinline__ double gaussian(uint32_t seed, uint32_t stream_id, int64_t i) {
  const uint32_t lo = static_cast<uint32_t>(i), hi = static_cast<uint32_t>(static_cast<uint64_t>(i) >> 32);
  const uint32_t h1 = mix32(lo ^ mix32(hi + 0x9E3779B9u * (seed + 1u)) ^ (stream_id * 0x85EBCA6Bu));
  const uint32_t h2 = mix32(h1 ^ 0xC2B2AE35u ^ seed);
  const double u1 = (static_cast<double>(h1) + 1.0) * (1.0 / 4294967297.0);
  const double u2 = static_cast<double>(h2) * (1.0 / 4294967296.0);
  return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}
}  // namespace kernels
double Device::RandomProjection(const double* v, int64_t n, uint32_t seed, uint32_t stream_id, bool sharded, int64_t index_offset) {
  REDUCE_S(sharded, 2, 0, n, { const double z = gaussian(seed, stream_id, i + index_offset); s[0] += z * v[i]; s[1] += z * z; });
  return host_results_[0] / std::sqrt(host_results_[1]);
}

// ---- trust region -----------------------------------------------------------
double* Device::TrScratch(int64_t doubles) {
  if (doubles > tr_scratch_size_) {
    cudaFree(tr_scratch_);
    tr_scratch_ = nullptr;
    CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&tr_scratch_), sizeof(double) * static_cast<size_t>(doubles + 64), STREAM));
    tr_scratch_size_ = doubles;
  }
  return tr_scratch_;
}

namespace kernels {
static bool TrLegacy();
static int TrSms();
static long long PeerTimeoutCycles() {
  static const long long v = [] {
    const char* e = std::getenv("PDLP_B200_PEER_TIMEOUT_S");
    const double seconds = (e != nullptr && *e != 0) ? std::max(0.1, std::atof(e)) : 20.0;
    return static_cast<long long>(seconds * 1.9e9);  // SM clocks (~1.9 GHz)
  }();
  return v;
}
static PeerPtrs MakeTrPeerPtrs(const PeerArena* arena, int64_t n, int64_t m_global, int set = 0) {
  PeerPtrs pp;
  std::memset(&pp, 0, sizeof(pp));
  if (arena == nullptr) return pp;
  pp.world = arena->world;
  pp.rank = arena->rank;
  for (int h = 0; h < kMaxPeers; ++h) pp.base[h] = static_cast<double*>(arena->base[h]);
  const PeerLayout l = PeerLayout::For(n, m_global, pp.world);
  pp.xt_off = l.xt_off; pp.partial_off = l.partial_off; pp.y_off = l.y_off; pp.scal_off = l.scal_off; pp.flags_off = l.flags_off; pp.epoch_off = l.epoch_off;
  pp.tr_off = set == 0 ? l.tr_off : l.tr2_off;
  pp.cand_off = set == 0 ? l.cand_off : l.cand2_off;
  pp.tr_barrier = set == 0 ? 3 : 4;
  pp.timeout_cycles = PeerTimeoutCycles();
  return pp;
}
// Persistent solve (k_tr_solve). Returns false when it does not apply (PDLP_B200_TR_LEGACY=1, or a
// row-sharded solve without peer arenas) and the caller must take the multi-launch path.
// JOINT: x0 / y0 / out as in TrSolveArgs.
template <class Elem, bool JOINT>
bool tr_solve_persistent(cudaStream_t stream, Comm* comm, const PeerPtrs& peer, int32_t* peer_error, int64_t total, Elem el, double radius,
                         const double* x0, const double* y0, double* scratch, double* partials, TrSearchState* st, double* out, int64_t* launches,
                         int blocks_per_sm = kTrBlocksPerSm) {
  const bool use_peer = comm != nullptr && peer.world > 1;
  if (TrLegacy() || (comm != nullptr && !use_peer)) return false;
  constexpr int64_t kTile = static_cast<int64_t>(kThreads) * kTrUnroll;
  const int nbp = static_cast<int>(std::min<int64_t>(std::min<int64_t>(kTrMaxBlocks, static_cast<int64_t>(TrSms()) * blocks_per_sm), std::max<int64_t>(1, (total + kTile - 1) / kTile)));
  TrSolveArgs g;
  g.total = total;
  g.keys = reinterpret_cast<unsigned long long*>(scratch);
  g.a = scratch + total;
  g.b = scratch + 2 * total;
  g.st = st;
  g.partials = partials;
  g.sync = reinterpret_cast<unsigned int*>(scratch + 3 * total + 32);
  g.radius = radius;
  g.x0 = x0;
  g.y0 = y0;
  g.out = out;
  g.peer_error = peer_error;
  g.cand_keys = reinterpret_cast<unsigned long long*>(scratch + 3 * total + 128);
  g.cand_a = scratch + 3 * total + 128 + kTrFinishCap;
  g.cand_b = scratch + 3 * total + 128 + 2 * kTrFinishCap;
  static const bool trace = [] { const char* t = std::getenv("PDLP_B200_TRACE"); return t != nullptr && t[0] == '1'; }();
  static int traced_calls = 0;
  g.trace = nullptr;
  if (trace && traced_calls < 8) {
    g.trace = reinterpret_cast<unsigned long long*>(scratch + 3 * total + 34);  // (after the two sync words; scratch has 64 spare doubles... the stamps need 31)
    CUDA_OK(cudaMemsetAsync(g.trace, 0, 31 * sizeof(unsigned long long), stream));
  }
  CUDA_OK(cudaMemsetAsync(g.sync, 0, 2 * sizeof(unsigned int), stream));
  const size_t smem = sizeof(double) * kTrCols * kThreads;
  PeerPtrs pp = peer;
  void* args[] = {&g, &el, &pp};
  const void* fn = use_peer ? reinterpret_cast<const void*>(k_tr_solve<Elem, true, JOINT>) : reinterpret_cast<const void*>(k_tr_solve<Elem, false, JOINT>);
  // (per device, so set on every launch: a process may drive several devices)
  cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(nbp), dim3(kThreads), args, smem, stream);
  if (e != cudaSuccess) throw std::runtime_error(std::string("cooperative launch of k_tr_solve failed: ") + cudaGetErrorString(e));
  *launches += 1;
  if (g.trace != nullptr) {
    ++traced_calls;
    unsigned long long h[31];
    CUDA_OK(cudaMemcpyAsync(h, g.trace, sizeof(h), cudaMemcpyDeviceToHost, stream));
    CUDA_OK(cudaStreamSynchronize(stream));
    std::fprintf(stderr, "[pdlp_b200 trace] k_tr_solve (%lld elements, %d blocks) phase us of block 0 [start | prepare | round 0 | passes ... | final sweep | final round]:", static_cast<long long>(total), nbp);
    for (unsigned long long k = 1; k < h[0] && k < 30; ++k) std::fprintf(stderr, " %.1f", (h[1 + k] - h[k]) / 1000.0);
    std::fprintf(stderr, "\n");
  }
  return true;
}

// Multi-launch threshold search (PDLP_B200_TR_LEGACY=1 and row-sharded solves whose ranks share no
// peer arenas: the bin totals are all-reduced by NCCL between the passes). Leaves the step size in
// st->step_size (device).
template <class Elem>
void tr_search_multilaunch(cudaStream_t stream, Comm* comm, int64_t total, Elem el, double radius, double* scratch, double* partials,
                           TrSearchState* st, double* totals, int64_t* launches) {
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(scratch);
  double* a = scratch + total;
  double* b = scratch + 2 * total;
  const int nb = static_cast<int>(std::min<int64_t>(148 * 2, std::max<int64_t>(1, (total + kThreads * kTrUnroll - 1) / (kThreads * kTrUnroll))));
  const int nb_prep = static_cast<int>(std::min<int64_t>(kMaxReduceBlocks, std::max<int64_t>(1, (total + kThreads * 2 - 1) / (kThreads * 2))));
  k_tr_prepare<Elem><<<nb_prep, kThreads, 0, stream>>>(total, 0, el, keys, a, b, partials);
  k_tr_init<<<1, 32, 0, stream>>>(st, radius, partials, nb_prep);
  *launches += 2;
  if (comm != nullptr) {
    double* mx = reinterpret_cast<double*>(st) + offsetof(TrSearchState, max_abs_objective) / sizeof(double);
    comm->AllReduceMax(mx, mx, 1, stream);
  }
  for (int shift = 60; shift >= 0; shift -= 4) {
    k_tr_pass<<<nb, kThreads, 0, stream>>>(total, keys, a, b, st, shift, partials);
    k_tr_totals<<<1, 17 * 32, 0, stream>>>(nb, partials, totals, st, shift, comm == nullptr ? 1 : 0);
    *launches += 2;
    if (comm != nullptr) {
      comm->AllReduceSum(totals, totals, 34, stream);
      k_tr_pick<<<1, 1, 0, stream>>>(totals, st, shift);
      *launches += 1;
    }
  }
  k_tr_finish<<<1, 1, 0, stream>>>(st, radius);
  *launches += 1;
}
static bool TrLegacy() {
  static const bool v = [] { const char* e = std::getenv("PDLP_B200_TR_LEGACY"); return e != nullptr && e[0] == '1'; }();
  return v;
}
static int TrSms() {
  int dev = 0, n = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n;
}
}  // namespace kernels

// The restart test needs the bounds at the average AND at the current iterate (pdhg.cc:2109-2170):
// two independent joint problems. One GPU: two persistent launches on two streams, one block per
// SM each (both co-resident), so the latency-bound rounds of one solve overlap the other's; one
// host synchronisation for both. Row-sharded peer solves: the second solve exchanges through its own
// arena segments and barrier. Not used with the diagonal solver or the NCCL-only exchange: returns
// false and the caller solves them in turn.
bool Device::LocalizedLagrangianBoundsPair(const double* const x[2], const double* const y[2], const double* const kx[2], const double* const kty[2],
                                           const double* c, const double* q, const double* lv, const double* uv, const double* lc, const double* uc,
                                           double primal_weight, int64_t n, int64_t mm, const double* x0, const double* y0, double out[2][3],
                                           double extra_out[2][3]) {
  static const bool enabled = [] { const char* e = std::getenv("PDLP_B200_TR_PAIR"); return !(e != nullptr && e[0] == '0'); }();
  const bool use_peer = comm_ != nullptr && peer_arena_ != nullptr && peer_arena_->world > 1;
  if (!enabled || (comm_ != nullptr && !use_peer) || TrLegacy() || x0 == nullptr || y0 == nullptr) return false;
  // (row-sharded: every rank counts its own slice of the replicated primal side; the two solves exchange
  // through their own arena segments and barrier, PeerLayout::tr_off / tr2_off)
  const int64_t pbeg = PrimalSliceBegin(n), plen = PrimalSliceEnd(n) - pbeg;
  const int64_t total = plen + mm;
  const int64_t per = ((3 * total + 128 + 3 * kTrFinishCap) + 63) / 64 * 64;
  double* scratch = TrScratch(2 * per);
  if (stream2_ == nullptr) {
    cudaStream_t s2;
    CUDA_OK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    stream2_ = s2;
    for (void*& e : pair_ev_) { cudaEvent_t ev; CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); e = ev; }
  }
  cudaStream_t s2 = static_cast<cudaStream_t>(stream2_);
  CUDA_OK(cudaEventRecord(static_cast<cudaEvent_t>(pair_ev_[0]), STREAM));   // (the second stream starts after what the first has queued so far)
  CUDA_OK(cudaStreamWaitEvent(s2, static_cast<cudaEvent_t>(pair_ev_[0]), 0));
  int32_t* err = reinterpret_cast<int32_t*>(host_results_ + 40);
  err[0] = err[1] = 0;
  for (int k = 0; k < 2; ++k) {
    const JointElem el{x[k], y[k], kx[k], kty[k], c, q, lv, uv, lc, uc, primal_weight, plen, mm, pbeg};
    double* sc = scratch + k * per;
    const PeerPtrs tr_peer = MakeTrPeerPtrs(use_peer ? peer_arena_ : nullptr, peer_arena_n_, peer_arena_m_, k);
    cudaStream_t sk = k == 0 ? STREAM : s2;
    if (!tr_solve_persistent<JointElem, true>(sk, use_peer ? comm_ : nullptr, tr_peer, tr_peer_error_ + k, total, el, -1.0, x0, y0, sc, partials_ + k * (kMaxReduceBlocks * 20),
                                              reinterpret_cast<TrSearchState*>(sc + 3 * total), results_ + 16 * k, &launches_, 1))
      throw std::runtime_error("persistent trust-region launch refused");
    CUDA_OK(cudaMemcpyAsync(host_results_ + 16 * k, results_ + 16 * k, sizeof(double) * 7, cudaMemcpyDeviceToHost, sk));
    if (use_peer) CUDA_OK(cudaMemcpyAsync(err + k, tr_peer_error_ + k, sizeof(int32_t), cudaMemcpyDeviceToHost, sk));
  }
  CUDA_OK(cudaEventRecord(static_cast<cudaEvent_t>(pair_ev_[1]), s2));
  CUDA_OK(cudaStreamWaitEvent(STREAM, static_cast<cudaEvent_t>(pair_ev_[1]), 0));
  Sync();
  if (err[0] != 0 || err[1] != 0) throw std::runtime_error("peer-memory exchange timed out in the trust-region search: a rank did not arrive");
  for (int k = 0; k < 2; ++k) {
    const double* r = host_results_ + 16 * k;
    const double lagrangian = r[0] + r[1];
    out[k][0] = lagrangian;
    out[k][1] = lagrangian + r[2];
    out[k][2] = lagrangian + r[3];
    extra_out[k][0] = r[4];
    extra_out[k][1] = r[5];
    extra_out[k][2] = r[6];
  }
  return true;
}

void Device::LocalizedLagrangianBounds(const double* x, const double* y, const double* kx, const double* kty, const double* c, const double* q,
                                       const double* lv, const double* uv, const double* lc, const double* uc, double primal_weight, double radius,
                                       bool use_diagonal_solver, double diagonal_tol, int64_t n, int64_t mm, double out[3], const double* x0,
                                       const double* y0, double* extra_out) {
  // The primal side is replicated on a row-sharded solve: every rank counts its own slice of it.
  const int64_t pbeg = PrimalSliceBegin(n), plen = PrimalSliceEnd(n) - pbeg;
  const int64_t total = plen + mm;
  const JointElem el{x, y, kx, kty, c, q, lv, uv, lc, uc, primal_weight, plen, mm, pbeg};
  double* scratch = TrScratch(3 * total + 128 + 3 * kTrFinishCap);
  TrSearchState* st = reinterpret_cast<TrSearchState*>(scratch + 3 * total);
  const bool radius_from_distance = radius < 0.0 && x0 != nullptr && y0 != nullptr;

  if (!use_diagonal_solver) {
    // One persistent launch: Lagrangian value, radius (when asked for), threshold search, objective
    // deltas (trust_region.cc:886-1016); one device->host copy.
    const PeerPtrs tr_peer = MakeTrPeerPtrs(peer_arena_, peer_arena_n_, peer_arena_m_);
    if (tr_solve_persistent<JointElem, true>(STREAM, comm_, tr_peer, tr_peer_error_, total, el, radius, radius_from_distance ? x0 : nullptr,
                                             radius_from_distance ? y0 : nullptr, scratch, partials_, st, results_, &launches_)) {
      CUDA_OK(cudaGetLastError());
      CUDA_OK(cudaMemcpyAsync(host_results_, results_, sizeof(double) * 7, cudaMemcpyDeviceToHost, STREAM));
      int32_t* err = reinterpret_cast<int32_t*>(host_results_ + 8);
      *err = 0;
      if (comm_ != nullptr && peer_arena_ != nullptr) CUDA_OK(cudaMemcpyAsync(err, tr_peer_error_, sizeof(int32_t), cudaMemcpyDeviceToHost, STREAM));
      Sync();
      if (*err != 0) throw std::runtime_error("peer-memory exchange timed out in the trust-region search: a rank did not arrive");
      const double lagrangian = host_results_[0] + host_results_[1];
      out[0] = lagrangian;
      out[1] = lagrangian + host_results_[2];
      out[2] = lagrangian + host_results_[3];
      if (extra_out != nullptr) {
        extra_out[0] = host_results_[4];
        extra_out[1] = radius_from_distance ? host_results_[5] : -1.0;
        extra_out[2] = radius_from_distance ? host_results_[6] : -1.0;
      }
      return;
    }
  }
  double dist[2] = {-1.0, -1.0};
  if (radius_from_distance) {
    DistancesSq(x, x0, n, y, y0, mm, dist);
    radius = std::sqrt((0.5 * primal_weight) * dist[0] + (0.5 / primal_weight) * dist[1]);
  }
  if (extra_out != nullptr) {
    extra_out[0] = radius;
    extra_out[1] = dist[0];
    extra_out[2] = dist[1];
  }
  // Lagrangian value = primal part + dual part (sou.cc:446-527).
  REDUCE_S(true, 2, 0, total, {
    if (i < plen) {
      const int64_t ip = pbeg + i;
      const double g = el.primal_gradient(ip);
      const double op = q != nullptr ? q[ip] * x[ip] : 0.0;
      s[0] += x[ip] * (g - 0.5 * op);
    } else {
      const int64_t j = i - plen;
      s[1] += el.subgradient_coefficient(j) * y[j];
    }
  });
  const double lagrangian = host_results_[0] + host_results_[1];

  if (!use_diagonal_solver) {
    tr_search_multilaunch(STREAM, comm_, total, el, radius, scratch, partials_, st, scratch + 3 * total + 16, &launches_);
    CUDA_OK(cudaGetLastError());
    // objective deltas at the solution (trust_region.cc:929-967)
    const TrSearchState* cst = st;
    REDUCE_S(true, 2, 0, total, {
        double obj, lb, ub, center, w, qd;
      el.get(i, obj, lb, ub, center, w, qd);
      const double sol = projected_value(center, obj, w, lb, ub, cst->step_size);
      if (i < plen) s[0] += obj * (sol - center);
      else s[1] += (-obj) * (sol - center);
    });
    out[0] = lagrangian;
    out[1] = lagrangian + host_results_[0];
    out[2] = lagrangian + host_results_[1];

    return;
  }
  // Diagonal solver: bisection on the scaling factor (trust_region.cc:611-753).
  double scaling = 0.0;
  if (radius != 0.0) {
    // FindScalingFactor: first double the bracket, then bisect.
    double lo = 0.0, hi = 1.0;
    bool bracketing = true;
    for (;;) {
      const double sf = bracketing ? hi : (lo + hi) / 2.0;
      if (!bracketing && !((hi - lo) >= diagonal_tol * std::max(1.0, lo))) break;
      REDUCE_S(true, 1, 0, total, {
            double obj, lb, ub, center, w, qd;
        el.get(i, obj, lb, ub, center, w, qd);
        const double sw = sqrt(w);
        const double v = fmin(fmax((-obj / sw) / (qd / w + sf), sw * (lb - center)), sw * (ub - center));
        s[0] += v * v;
      });
      const double norm = std::sqrt(host_results_[0]);
      if (bracketing) {
        if (norm >= radius) { lo = hi; hi *= 2; } else { bracketing = false; }
      } else {
        if (norm <= radius) hi = sf; else lo = sf;
      }
    }
    scaling = (hi + lo) / 2.0;
  }
  const bool zero_radius = radius == 0.0;
  REDUCE_S(true, 2, 0, total, {
    double obj, lb, ub, center, w, qd;
    el.get(i, obj, lb, ub, center, w, qd);
    double diff = 0.0;
    if (!zero_radius) {
      const double sw = sqrt(w);
      const double v = fmin(fmax((-obj / sw) / (qd / w + scaling), sw * (lb - center)), sw * (ub - center));
      const double sol = center + sqrt(1 / w) * v;
      diff = sol - center;
    }
    if (i < plen) s[0] += obj * diff + 0.5 * qd * diff * diff;
    else s[1] += (-obj) * diff;
  });
  out[0] = lagrangian;
  out[1] = lagrangian + host_results_[0];
  out[2] = lagrangian + host_results_[1];
}

void Device::SolveTrustRegion(const double* obj, const double* lb, const double* ub, const double* center, const double* w, double radius,
                              int64_t n, double* solution, double* step_size, double* objective_value) {
  const VectorElem el{obj, lb, ub, center, w, nullptr};
  double* scratch = TrScratch(3 * n + 128 + 3 * kTrFinishCap);
  TrSearchState* st = reinterpret_cast<TrSearchState*>(scratch + 3 * n);
  if (!tr_solve_persistent<VectorElem, false>(STREAM, nullptr, PeerPtrs{}, nullptr, n, el, radius, nullptr, nullptr, scratch, partials_, st, nullptr, &launches_))
    tr_search_multilaunch(STREAM, nullptr, n, el, radius, scratch, partials_, st, scratch + 3 * n + 16, &launches_);
  CUDA_OK(cudaGetLastError());
  const TrSearchState* cst = st;
  REDUCE(1, 0, n, {
    const double sol = projected_value(center[i], obj[i], w[i], lb[i], ub[i], cst->step_size);
    solution[i] = sol;
    s[0] += obj[i] * (sol - center[i]);
  });
  *objective_value = host_results_[0];
  TrSearchState hs;
  CUDA_OK(cudaMemcpy(&hs, st, sizeof(hs), cudaMemcpyDeviceToHost));
  *step_size = hs.step_size;
}

void Device::SolveDiagonalTrustRegion(const double* obj, const double* qdiag, const double* lb, const double* ub, const double* center,
                                      const double* w, double radius, double tol, int64_t n, double* solution, double* step_size,
                                      double* objective_value) {
  if (radius == 0.0) {
    CopyD2D(solution, center, n);
    Sync();
    *step_size = 0.0;
    *objective_value = 0.0;
    return;
  }
  double lo = 0.0, hi = 1.0;
  bool bracketing = true;
  for (;;) {
    const double sf = bracketing ? hi : (lo + hi) / 2.0;
    if (!bracketing && !((hi - lo) >= tol * std::max(1.0, lo))) break;
    REDUCE(1, 0, n, {
      const double sw = sqrt(w[i]);
      const double v = fmin(fmax((-obj[i] / sw) / (qdiag[i] / w[i] + sf), sw * (lb[i] - center[i])), sw * (ub[i] - center[i]));
      s[0] += v * v;
    });
    const double norm = std::sqrt(host_results_[0]);
    if (bracketing) {
      if (norm >= radius) { lo = hi; hi *= 2; } else { bracketing = false; }
    } else {
      if (norm <= radius) hi = sf; else lo = sf;
    }
  }
  const double scaling = (hi + lo) / 2.0;
  REDUCE(1, 0, n, {
    const double sw = sqrt(w[i]);
    const double v = fmin(fmax((-obj[i] / sw) / (qdiag[i] / w[i] + scaling), sw * (lb[i] - center[i])), sw * (ub[i] - center[i]));
    const double sol = center[i] + sqrt(1 / w[i]) * v;
    solution[i] = sol;
    const double d = sol - center[i];
    s[0] += 0.5 * d * qdiag[i] * d + d * obj[i];
  });
  *step_size = scaling;
  *objective_value = host_results_[0];
}

// ---- PDHG step ----------------------------------------------------------------
StepState* Device::AllocState() {  // two slots (see DecideHead)
  StepState* p = nullptr;
  CUDA_OK(cudaMalloc(&p, 2 * sizeof(StepState)));
  CUDA_OK(cudaMemset(p, 0, 2 * sizeof(StepState)));
  return p;
}
void Device::UploadState(StepState* dev, const StepState& host) {
  // (pageable source: the runtime stages the bytes before it returns, and the stream orders the
  // copy before the step kernels -- no host synchronisation needed)
  CUDA_OK(cudaMemcpyAsync(dev, &host, sizeof(StepState), cudaMemcpyHostToDevice, STREAM));
}
void Device::DownloadState(StepState& host, const StepState* dev) {
  CUDA_OK(cudaMemcpyAsync(&host, dev, sizeof(StepState), cudaMemcpyDeviceToHost, STREAM));
  Sync();
}
int Device::DownloadLatestState(StepState& host, const StepState* slots, int preferred_slot) {
  StepState both[2];
  CUDA_OK(cudaMemcpyAsync(both, slots, 2 * sizeof(StepState), cudaMemcpyDeviceToHost, STREAM));
  Sync();
  // every executed decision increments `attempts` while it moves the state to the other slot
  int slot = preferred_slot;
  if (both[1 - preferred_slot].attempts > both[preferred_slot].attempts) slot = 1 - preferred_slot;
  host = both[slot];
  return slot;
}

static StepPtrs MakePtrs(const Device::StepBuffers& b) {
  StepPtrs p;
  p.n = b.n; p.m = b.m;
  for (int k = 0; k < 3; ++k) { p.x[k] = b.x[k]; p.y[k] = b.y[k]; p.kty[k] = b.kty[k]; p.kx[k] = b.kx[k]; }
  p.x_tilde = b.x_tilde; p.avg_x = b.avg_x; p.avg_y = b.avg_y;
  p.avg_kx = b.arena != nullptr ? b.avg_kx : nullptr;   // (maintained by the peer-exchange kernels only)
  p.avg_kty = b.arena != nullptr ? b.avg_kty : nullptr;
  p.c = b.c; p.q = b.q; p.lv = b.lv; p.uv = b.uv; p.lc = b.lc; p.uc = b.uc;
  p.state = b.state;
  return p;
}

static PeerPtrs MakePeerPtrs(const Device::StepBuffers& b) {
  PeerPtrs pp;
  std::memset(&pp, 0, sizeof(pp));
  if (b.arena == nullptr) return pp;
  pp.world = b.arena->world;
  pp.rank = b.arena->rank;
  for (int h = 0; h < kMaxPeers; ++h) pp.base[h] = static_cast<double*>(b.arena->base[h]);
  const PeerLayout l = PeerLayout::For(b.n, b.m_global, pp.world);
  pp.xt_off = l.xt_off; pp.partial_off = l.partial_off; pp.y_off = l.y_off; pp.scal_off = l.scal_off; pp.flags_off = l.flags_off; pp.epoch_off = l.epoch_off; pp.tr_off = l.tr_off; pp.cand_off = l.cand_off;
  pp.tr_barrier = 3;
  pp.row_begin = b.row_begin;
  pp.timeout_cycles = PeerTimeoutCycles();
  pp.begin = b.slice_begin;
  pp.end = b.slice_end;
  return pp;
}

void Device::EnqueueSteps(const StepBuffers& b, const SellDev& rows, const SellDev& cols, int count, int first_slot) {
  StepPtrs p = MakePtrs(b);
  const bool use_peer = b.arena != nullptr;
  const bool pdl = StepPdl();
  const PeerPtrs peer = MakePeerPtrs(b);
  const int64_t primal_work = use_peer ? b.slice_end - b.slice_begin : b.n;
  const int np = static_cast<int>(std::max<int64_t>(1, ((primal_work + 1) / 2 + kThreads - 1) / kThreads));
  // one GPU: the SpMV pair runs as persistent launches with a tile queue (no partial last wave)
  const bool persist = SellPersist() && !use_peer && comm_ == nullptr && SellVariant() < 3;
  const int pair_chunks = persist ? SellPersistChunks() : SellChunks();
  TileQueue q_dual, q_kty;
  if (persist) {
    q_dual.counter = tile_counters_;
    q_dual.reset = tile_counters_ + 1;
    q_kty.counter = tile_counters_ + 1;
    q_kty.reset = tile_counters_;
    CUDA_OK(cudaMemsetAsync(tile_counters_, 0, 2 * sizeof(unsigned int), STREAM));
  }
  const int nd_main = SellGridFor(rows, pair_chunks);
  const int nd_fix = rows.num_split > 0 ? static_cast<int>((rows.num_split * 32 + kThreads - 1) / kThreads) : 0;
  const int nd = b.m > 0 ? nd_main + nd_fix : 0;
  const int64_t need = std::max<int64_t>(static_cast<int64_t>(np) + 2 * static_cast<int64_t>(nd_main + nd_fix) + 8, 4 * static_cast<int64_t>(num_sms_) + 8 + 2 * (rows.num_slots / 32) + 8);
  if (need > step_partials_size_) {
    cudaFree(step_partials_);
    step_partials_ = nullptr;
    CUDA_OK(cudaMalloc(&step_partials_, sizeof(double) * need));
    CUDA_OK(cudaMemset(step_partials_, 0, sizeof(double) * need));
    step_partials_size_ = need;
  }
  double* pd = step_partials_;              // [nd][2] (16-byte aligned: read as double2)
  double* pp = pd + 2 * (nd_main + nd_fix);  // [np]
  constexpr int kMaxSamples = 64;
  timing_attempt_idx_.clear();
  timing_peer_ = use_peer;
  auto ev = [&](int slot, int k) { CUDA_OK(cudaEventRecord(static_cast<cudaEvent_t>(timing_events_[slot * kEvPerSlot + k]), STREAM)); };
  // k_peer_loop wins where an attempt is launch- and barrier-bound (C2 on 8 GPUs: 5905 against 5396 it/s);
  // with more work per rank the separate launches are level or ahead (C2 / C4 on 2 GPUs: 4338 against 4252,
  // 906 against 864 it/s): their row loops are the tuned ones and a phase of the loop ends with the tail of
  // its last slice. Chosen by the size of the larger image; PDLP_B200_PEER_LOOP=0 / 2 forces the launches /
  // the loop. Split rows need their fix-up kernel.
  const SellDev& kty_image = b.cols_slice != nullptr ? *b.cols_slice : cols;
  constexpr int kLoopThreads = 512;
  constexpr int64_t kPeerLoopMaxSlices = 16384;
  const bool loop_small = std::max(rows.num_slots, kty_image.num_slots) / 32 <= kPeerLoopMaxSlices;
  if (use_peer && (PeerLoop() == 2 || (PeerLoop() == 1 && loop_small)) && b.m > 0 && b.n > 0 && rows.num_split == 0 && kty_image.num_split == 0) {
    // the whole chunk of attempts as one persistent cooperative launch (k_peer_loop); it leaves when the
    // decision halts, so rejected steps need no second pass (the cap only bounds a state that never halts)
    PeerLoopArgs g;
    g.p = p;
    g.p.state = b.state;
    g.peer = peer;
    g.rows = rows;
    g.cols = b.cols_slice != nullptr ? *b.cols_slice : cols;
    g.perm = b.cols_slice != nullptr ? b.slice_perm : b.primal_scatter;
    g.col0 = b.slice_begin;
    g.block_partials = step_partials_;
    g.slice_partials = step_partials_ + 4 * static_cast<int64_t>(num_sms_) + 8;
    g.sync = loop_sync_;
    g.first_slot = first_slot;
    g.max_attempts = count + 64;
    static const bool trace_env = [] { const char* t = std::getenv("PDLP_B200_TRACE"); return t != nullptr && t[0] == '1'; }();
    g.trace = (step_timing_ || trace_env) ? reinterpret_cast<unsigned long long*>(loop_sync_ + 16) : nullptr;
    CUDA_OK(cudaMemsetAsync(loop_sync_, 0, 128, STREAM));
    const void* fn = b.cols_slice != nullptr ? reinterpret_cast<const void*>(k_peer_loop<0, kLoopThreads>) : reinterpret_cast<const void*>(k_peer_loop<1, kLoopThreads>);
    void* args[] = {&g};
    const cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(num_sms_), dim3(kLoopThreads), args, 0, STREAM);
    if (e != cudaSuccess) throw std::runtime_error(std::string("cooperative launch of k_peer_loop failed: ") + cudaGetErrorString(e));
    ++launches_;
    timing_attempt_idx_.clear();
    timing_peer_ = true;
    loop_traced_ = g.trace != nullptr;
    // (rides on the synchronisation of the state download that follows every batch)
    if (loop_traced_) CUDA_OK(cudaMemcpyAsync(loop_trace_host_, loop_sync_ + 16, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, STREAM));
    return;
  }
  for (int it = 0; it < count; ++it) {
    // the kernels of this attempt read state slot `in`; its decision writes the other slot
    StepState* st_in = b.state + ((first_slot + it) & 1);
    StepState* st_out = b.state + ((first_slot + it + 1) & 1);
    p.state = st_in;
    const int32_t* halt = &st_in->halt;
    DecideHead head;
    head.in = st_in;
    head.out = st_out;
    head.pp = pp;
    head.np = b.n > 0 ? np : 0;
    head.pd = pd;
    head.nd = nd;
    int slot = -1;
    if (step_timing_ && it % step_timing_stride_ == 0 && static_cast<int>(timing_attempt_idx_.size()) < kMaxSamples) {
      slot = static_cast<int>(timing_attempt_idx_.size());
      while (static_cast<int>(timing_events_.size()) < (slot + 1) * kEvPerSlot) {
        cudaEvent_t e;
        CUDA_OK(cudaEventCreate(&e));
        timing_events_.push_back(e);
      }
      timing_attempt_idx_.push_back(it);
      ev(slot, 0);
    }
    if (use_peer) {
      // the decision adds the G triples {||dx||^2, ||dy||^2, (K dx) . dy} that k_sum_push_barrier stored into this rank's arena
      head.scal = peer.base[peer.rank] + peer.scal_off;
      head.scal_count = peer.world;
      head.scal_stride = 4;
      head.scal_has_dx2 = 1;
    }
    if (use_peer && b.cols_slice != nullptr) {
      // all-gather exchange; sub-phases (events 0..7): primal slice + x~ stores | barrier | K x~ + dual + y' stores | - | sums + barrier | K^T y' slice (+ decision) | -
      launch_k(pdl, k_primal_step<true>, np, kThreads, STREAM, p, peer, pp);
      if (slot >= 0) ev(slot, 1);
      launch_k(pdl, k_peer_barrier, 1, 32, STREAM, peer, 0, st_in);
      launches_ += 2;
      if (slot >= 0) ev(slot, 2);
      if (b.m > 0) {
        DualEpiT<true, true> de;
        de.b = p;
        de.peer = peer;
        launch_sell<kDot, 2>(STREAM, rows, GatherSrc{{peer.base[peer.rank] + peer.xt_off, nullptr, nullptr}, nullptr}, de, pd, halt, &launches_, nullptr, nullptr, pdl);
      }
      if (slot >= 0) { ev(slot, 3); ev(slot, 4); }
      launch_k(pdl, k_sum_push_barrier, 1, kDecideThreads, STREAM, st_in, peer, 1, pp, np, pd, nd);
      ++launches_;
      if (slot >= 0) ev(slot, 5);
      launch_sell<kDot, 0>(STREAM, *b.cols_slice, GatherSrc{{peer.base[peer.rank] + peer.y_off, nullptr, nullptr}, nullptr}, KtyEpiSlice{p, b.slice_perm, b.slice_begin}, nullptr, halt,
                           &launches_, nullptr, nullptr, pdl, head);
      if (slot >= 0) { ev(slot, 6); ev(slot, 7); }
      continue;
    }
    if (use_peer) {
      // reduce-scatter exchange; sub-phases (events 0..7): primal slice + x~ stores | barrier | K x~ + dual | K^T y' partial | sums + barrier | slice pull (+ decision) | -
      launch_k(pdl, k_primal_step<true>, np, kThreads, STREAM, p, peer, pp);
      if (slot >= 0) ev(slot, 1);
      launch_k(pdl, k_peer_barrier, 1, 32, STREAM, peer, 0, st_in);
      launches_ += 2;
      if (slot >= 0) ev(slot, 2);
      if (b.m > 0) {
        DualEpiT<false, true> de;
        de.b = p;
        std::memset(&de.peer, 0, sizeof(de.peer));
        launch_sell<kDot, 2>(STREAM, rows, GatherSrc{{peer.base[peer.rank] + peer.xt_off, nullptr, nullptr}, nullptr}, de, pd, halt, &launches_, nullptr, nullptr, pdl);
      }
      if (slot >= 0) ev(slot, 3);
      launch_sell<kDot, 0>(STREAM, cols, GatherSrc{{b.y[0], b.y[1], b.y[2]}, st_in}, ScatterEpi{peer.base[peer.rank] + peer.partial_off, b.primal_scatter}, nullptr, halt, &launches_, nullptr, nullptr, pdl);
      if (slot >= 0) ev(slot, 4);
      launch_k(pdl, k_sum_push_barrier, 1, kDecideThreads, STREAM, st_in, peer, 1, pp, np, pd, nd);
      if (slot >= 0) ev(slot, 5);
      launch_k(pdl, k_kty_finish_peer, np, kThreads, STREAM, p, peer, head);
      launches_ += 2;
      if (slot >= 0) { ev(slot, 6); ev(slot, 7); }
      continue;
    }
    launch_k(pdl, k_primal_step<false>, np, kThreads, STREAM, p, peer, pp);
    ++launches_;
    if (slot >= 0) ev(slot, 1);
    if (b.m > 0) {
      launch_sell<kDot, 2>(STREAM, rows, GatherSrc{{b.x_tilde, nullptr, nullptr}, nullptr}, MakeDualEpi(p), pd, halt, &launches_, nullptr, nullptr, pdl, DecideHead(), q_dual, pair_chunks);
    }
    if (slot >= 0) ev(slot, 2);
    if (comm_ != nullptr) {
      // NCCL exchange: local {||dy||^2, (K dx) . dy} -> exchange[n], exchange[n + 1]; local K^T y' partial -> exchange[0..n) in column order
      k_sum_to_slot<<<1, kDecideThreads, 0, STREAM>>>(st_in, pd, nd, b.exchange + b.n);
      ++launches_;
      launch_sell<kDot, 0>(STREAM, cols, GatherSrc{{b.y[0], b.y[1], b.y[2]}, st_in}, ScatterEpi{b.exchange, b.primal_scatter}, nullptr, halt, &launches_, nullptr, nullptr, pdl);
      comm_->AllReduceSum(b.exchange, b.exchange, b.n + 2, stream_);
      head.scal = b.exchange + b.n - 1;  // (so that entries 1, 2 of the "triple" are exchange[n], exchange[n + 1]; ||dx||^2 comes from pp: the primal side is replicated)
      head.scal_count = 1;
      head.scal_stride = 0;
      head.scal_has_dx2 = 0;
      const int nk = Blocks(b.n);
      k_kty_finish<<<nk, kThreads, 0, STREAM>>>(p, b.exchange, head);
      ++launches_;
      if (slot >= 0) ev(slot, 3);
    } else if (b.n > 0) {
      // three launches per attempt: block 0 of the K^T y' kernel takes the decision while the others stream the matrix
      launch_sell<kDot, 0>(STREAM, cols, GatherSrc{{b.y[0], b.y[1], b.y[2]}, st_in}, KtyEpi{p}, nullptr, halt, &launches_, nullptr, nullptr, pdl, head, q_kty, pair_chunks);
      if (slot >= 0) ev(slot, 3);
    } else {
      launch_k(pdl, k_decide_only, 1, kThreads, STREAM, head);
      ++launches_;
      if (slot >= 0) ev(slot, 3);
    }
    if (slot >= 0) ev(slot, 4);
  }
  CUDA_OK(cudaGetLastError());
}

void Device::EnqueueMalitskyPockSteps(const StepBuffers& b, const SellDev& rows, const SellDev& cols, int count, int first_slot) {
  StepPtrs p = MakePtrs(b);
  const bool pdl = StepPdl();
  const int np = Blocks(b.n), nd = Blocks(b.m);
  const int nt_main = SellGrid(cols);
  const int nt_fix = cols.num_split > 0 ? static_cast<int>((cols.num_split * 32 + kThreads - 1) / kThreads) : 0;
  const int64_t need = static_cast<int64_t>(np) + nd + nt_main + nt_fix + 8;
  if (need > step_partials_size_) {
    cudaFree(step_partials_);
    step_partials_ = nullptr;
    CUDA_OK(cudaMalloc(&step_partials_, sizeof(double) * need));
    CUDA_OK(cudaMemset(step_partials_, 0, sizeof(double) * need));
    step_partials_size_ = need;
  }
  double* pp = step_partials_;
  double* pd = pp + np;
  double* pt = pd + nd;
  timing_attempt_idx_.clear();
  for (int it = 0; it < count; ++it) {
    StepState* st_in = b.state + ((first_slot + it) & 1);
    StepState* st_out = b.state + ((first_slot + it + 1) & 1);
    p.state = st_in;
    if (b.n > 0) {
      launch_k(pdl, k_mp_primal, np, kThreads, STREAM, p, pp);
      ++launches_;
      if (b.m > 0)  // K x' of the candidate, first attempt of an iteration only
        launch_sell<kDot, 0>(STREAM, rows, GatherSrc{{b.x[0], b.x[1], b.x[2]}, st_in}, KxStoreEpi{p}, nullptr, &st_in->mp_skip_primal, &launches_, nullptr, nullptr, pdl);
    }
    if (b.m > 0) {
      launch_k(pdl, k_mp_dual, nd, kThreads, STREAM, p, pd);
      ++launches_;
    }
    if (b.n > 0) launch_sell<kDot, 1>(STREAM, cols, GatherSrc{{b.y[0], b.y[1], b.y[2]}, st_in}, KtyDiffEpi{p}, pt, &st_in->halt, &launches_, nullptr, nullptr, pdl);
    launch_k(pdl, k_mp_decide, 1, kDecideThreads, STREAM, static_cast<const StepState*>(st_in), st_out, static_cast<const double*>(pp), b.n > 0 ? np : 0,
             static_cast<const double*>(pd), b.m > 0 ? nd : 0, static_cast<const double*>(pt), b.n > 0 ? nt_main + nt_fix : 0);
    ++launches_;
  }
  CUDA_OK(cudaGetLastError());
}

void Device::EnableStepTiming(bool on, int stride) {
  step_timing_ = on;
  step_timing_stride_ = std::max(1, stride);
  timing_attempt_idx_.clear();
}
void Device::CollectStepTimings(int64_t executed_attempts) {
  if (loop_traced_) {
    // k_peer_loop: block 0 summed the phase times of every attempt (globaltimer, ns):
    // [0] attempts, P | barrier A | D | (T1) | sums + barrier B | decision + T | closing barrier
    loop_traced_ = false;
    const unsigned long long* h = loop_trace_host_;  // copied behind the launch, complete since the state download synchronised
    static const int kLoopClass[7] = {0, 0, 1, 2, 2, 2, 3};
    if (h[0] > 0) {
      for (int k = 0; k < 7; ++k) {
        const double ms = static_cast<double>(h[1 + k]) * 1e-6;
        step_timings_.ms[kLoopClass[k]] += ms;
        detail_ms_[k] += ms;
      }
      for (int k = 0; k < 4; ++k) step_timings_.samples[k] += static_cast<int64_t>(h[0]);
      detail_samples_ += static_cast<int64_t>(h[0]);
    }
    return;
  }
  // sub-phase -> public kernel class (device_ops.h): peer exchange has 7 sub-phases, otherwise the 4 classes themselves
  static const int kPeerClass[7] = {0, 0, 1, 2, 2, 2, 3};
  const int phases = timing_peer_ ? 7 : 4;
  for (size_t s = 0; s < timing_attempt_idx_.size(); ++s) {
    if (timing_attempt_idx_[s] >= executed_attempts) continue;
    for (int k = 0; k < phases; ++k) {
      float ms = 0.f;
      CUDA_OK(cudaEventElapsedTime(&ms, static_cast<cudaEvent_t>(timing_events_[s * kEvPerSlot + k]), static_cast<cudaEvent_t>(timing_events_[s * kEvPerSlot + k + 1])));
      const int cls = timing_peer_ ? kPeerClass[k] : k;
      step_timings_.ms[cls] += ms;
      detail_ms_[k] += ms;
    }
    for (int k = 0; k < 4; ++k) step_timings_.samples[k] += 1;
    detail_samples_ += 1;
  }
  timing_attempt_idx_.clear();
}
void Device::TimelineStart(int id) {
  for (int k = 0; k < 2; ++k)
    if (timeline_ev_[id][k] == nullptr) { cudaEvent_t e; CUDA_OK(cudaEventCreate(&e)); timeline_ev_[id][k] = e; }
  CUDA_OK(cudaEventRecord(static_cast<cudaEvent_t>(timeline_ev_[id][0]), STREAM));
}
void Device::TimelineStop(int id) {
  CUDA_OK(cudaEventRecord(static_cast<cudaEvent_t>(timeline_ev_[id][1]), STREAM));
  timeline_pending_[id] = true;
}
double Device::TimelineCollectMs(int id) {
  if (!timeline_pending_[id]) return 0.0;
  timeline_pending_[id] = false;
  CUDA_OK(cudaEventSynchronize(static_cast<cudaEvent_t>(timeline_ev_[id][1])));
  float ms = 0.f;
  CUDA_OK(cudaEventElapsedTime(&ms, static_cast<cudaEvent_t>(timeline_ev_[id][0]), static_cast<cudaEvent_t>(timeline_ev_[id][1])));
  return ms;
}
double Device::TimelineStopMs(int id) {
  CUDA_OK(cudaEventRecord(static_cast<cudaEvent_t>(timeline_ev_[id][1]), STREAM));
  CUDA_OK(cudaEventSynchronize(static_cast<cudaEvent_t>(timeline_ev_[id][1])));
  float ms = 0.f;
  CUDA_OK(cudaEventElapsedTime(&ms, static_cast<cudaEvent_t>(timeline_ev_[id][0]), static_cast<cudaEvent_t>(timeline_ev_[id][1])));
  return ms;
}

void Device::GatherPrimalSlices(const StepBuffers& b, int cur, int prev) {
  if (b.arena == nullptr || comm_ == nullptr) return;
  // prev < 0: the previous iterate is only read by the iterate difference, which is materialised
  // lazily (it then asks for that one vector with cur < 0)
  comm_->GroupStart();
  if (cur >= 0)
    for (double* v : {b.x[cur], b.kty[cur], b.avg_x, b.avg_kty})
      if (v != nullptr) comm_->AllGatherInPlace(v, b.slice_stride, stream_);
  if (prev >= 0) comm_->AllGatherInPlace(b.x[prev], b.slice_stride, stream_);
  comm_->GroupEnd();
}

void Device::FlushAverages(const StepBuffers& b, int slot) {
  StepPtrs p = MakePtrs(b);
  p.state = b.state + slot;
  const int64_t total = b.n + b.m;
  if (total > 0) {
    k_flush_average<<<Blocks(total), kThreads, 0, STREAM>>>(p, total, b.arena != nullptr ? b.slice_begin : 0, b.arena != nullptr ? b.slice_end : b.n);
    ++launches_;
  }
  k_clear_pending<<<1, 1, 0, STREAM>>>(b.state + slot);
  LAUNCHED();
}

}  // namespace pdlp_b200
