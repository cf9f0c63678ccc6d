// Standalone probe (not part of the product): random 8-byte gather throughput
// from (a) global memory tables of several sizes (L1 / L2 resident), (b) shared
// memory, (c) distributed shared memory across a thread-block cluster.
// Decides whether a column-panel SpMV with x staged in (D)SMEM can beat the
// L1TEX-wavefront-bound gather of the plain SELL kernel.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// global gather: each thread does `per` gathers from table[0..tsize) with precomputed indices (coalesced index stream)
template <int VEC>
__global__ void k_gl(const int* __restrict__ idx, const double* __restrict__ table, size_t n, double* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  double acc = 0;
  for (; i + 7 * stride < n; i += 8 * stride) {
    int j[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) j[u] = idx[i + u * stride];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (VEC == 1) acc += table[j[u]];
      else { double2 v = *reinterpret_cast<const double2*>(table + (j[u] & ~1)); acc += v.x + v.y; }
    }
  }
  if (acc == 1.2345e300) out[0] = acc;
}

// shared-memory gather: table of tsize doubles in smem per CTA, indices streamed from global
__global__ void k_sm(const int* __restrict__ idx, const double* __restrict__ table, int tsize, size_t n, double* out) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < tsize; i += blockDim.x) sm[i] = table[i];
  __syncthreads();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  double acc = 0;
  for (; i + 7 * stride < n; i += 8 * stride) {
    int j[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) j[u] = idx[i + u * stride];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += sm[j[u]];
  }
  if (acc == 1.2345e300) out[0] = acc;
}

// DSMEM gather: cluster of CS CTAs, each holds tsize doubles; index -> (rank = idx / tsize, off = idx % tsize)
__global__ void k_dsm(const int* __restrict__ idx, const double* __restrict__ table, int tsize, size_t n, double* out) {
  extern __shared__ double sm[];
  cg::cluster_group cluster = cg::this_cluster();
  unsigned cs = cluster.num_blocks(), rank = cluster.block_rank();
  for (int i = threadIdx.x; i < tsize; i += blockDim.x) sm[i] = table[(size_t)rank * tsize + i];
  cluster.sync();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  double acc = 0;
  for (; i + 7 * stride < n; i += 8 * stride) {
    int j[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) j[u] = idx[i + u * stride];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      unsigned r = (unsigned)j[u] / (unsigned)tsize; int off = j[u] - r * tsize;
      const double* p = cluster.map_shared_rank(sm, r);
      acc += p[off];
    }
  }
  cluster.sync();
  if (acc == 1.2345e300) out[0] = acc;
  (void)cs;
}

template <class F> float timeit(F f, int reps = 10) {
  for (int i = 0; i < 2; ++i) f();
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  CK(cudaDeviceSynchronize()); cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b);
  CK(cudaGetLastError());
  return ms / reps;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("device %s SMs=%d clock %d kHz\n", p.name, sms, p.clockRate);
  const size_t N = 40000000;  // gathers per launch
  std::vector<int> h(N);
  double* table; CK(cudaMalloc(&table, (size_t)64 << 20)); CK(cudaMemset(table, 0, (size_t)64 << 20));
  int* idx; CK(cudaMalloc(&idx, N * 4)); double* out; CK(cudaMalloc(&out, 8));
  auto fill = [&](uint32_t range) { uint64_t s = 88172645463325252ull; for (size_t i = 0; i < N; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int)((s >> 20) % range); } CK(cudaMemcpy(idx, h.data(), N * 4, cudaMemcpyHostToDevice)); };
  // (a) global tables
  for (uint32_t range : {4096u, 16384u, 1u << 20, 2u << 20, 8u << 20}) {
    fill(range);
    float ms = timeit([&] { k_gl<1><<<sms * 8, 512>>>(idx, table, N, out); });
    float ms2 = timeit([&] { k_gl<2><<<sms * 8, 512>>>(idx, table, N, out); });
    printf("global table %8u doubles (%6.1f MB): 8B %.3f ms %.1f Gg/s (%.2f /clk/SM @1.965GHz) | 16B %.3f ms %.1f Gg/s\n", range, range * 8e-6, ms, N / ms * 1e-6, N / ms * 1e-6 / sms / 1.965, ms2, N / ms2 * 1e-6);
  }
  // (b) shared memory
  for (int tsize : {8192, 24576}) {
    fill(tsize);
    CK(cudaFuncSetAttribute(k_sm, cudaFuncAttributeMaxDynamicSharedMemorySize, tsize * 8));
    float ms = timeit([&] { k_sm<<<sms, 1024, tsize * 8>>>(idx, table, tsize, N, out); });
    printf("smem table %6d doubles: %.3f ms %.1f Gg/s (%.2f /clk/SM)\n", tsize, ms, N / ms * 1e-6, N / ms * 1e-6 / sms / 1.965);
  }
  // (c) DSMEM
  for (int cs : {2, 4, 8, 16}) {
    int tsize = 24576;  // 192 KB per CTA
    fill((uint32_t)tsize * cs);
    CK(cudaFuncSetAttribute(k_dsm, cudaFuncAttributeMaxDynamicSharedMemorySize, tsize * 8));
    if (cs > 8) CK(cudaFuncSetAttribute(k_dsm, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {}; cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1; cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = tsize * 8;
    int nclusters = 0; cfg.gridDim = dim3(cs);
    cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, k_dsm, &cfg);
    if (e != cudaSuccess || nclusters == 0) { printf("cluster %d: not launchable (%s)\n", cs, cudaGetErrorString(e)); cudaGetLastError(); continue; }
    cfg.gridDim = dim3(nclusters * cs);
    const int* cidx = idx; const double* ctab = table; size_t n = N;
    float ms = timeit([&] { CK(cudaLaunchKernelEx(&cfg, k_dsm, cidx, ctab, tsize, n, out)); });
    printf("dsmem cluster=%2d (%d clusters, table %.1f MB): %.3f ms %.1f Gg/s (%.2f /clk/SM over %d SMs)\n", cs, nclusters, tsize * cs * 8e-6, ms, N / ms * 1e-6, N / ms * 1e-6 / (nclusters * cs) / 1.965, nclusters * cs);
  }
  return 0;
}
