// Standalone probe (not part of the product): measures candidate fp64 SpMV
// layouts on a synthetic C2-shaped matrix (1M x 2M, 20 nnz/row)
// so the device layout of K / K^T can be chosen from data.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo tools/spmv_probe.cu -o gpurun_out/spmv_probe
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <numeric>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static inline uint64_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }
static inline double rnd01() { return (rnd() >> 11) * (1.0 / 9007199254740992.0); }

struct Csr { int rows, cols; std::vector<int> ptr, idx; std::vector<double> val; };

static Csr transpose(const Csr& a) {
  Csr t; t.rows = a.cols; t.cols = a.rows; t.ptr.assign(t.rows + 1, 0);
  for (int c : a.idx) t.ptr[c + 1]++;
  for (int i = 0; i < t.rows; ++i) t.ptr[i + 1] += t.ptr[i];
  t.idx.resize(a.idx.size()); t.val.resize(a.val.size());
  std::vector<int> pos(t.ptr.begin(), t.ptr.end() - 1);
  for (int r = 0; r < a.rows; ++r)
    for (int k = a.ptr[r]; k < a.ptr[r + 1]; ++k) { int p = pos[a.idx[k]]++; t.idx[p] = r; t.val[p] = a.val[k]; }
  return t;
}

// ---------------- kernels ----------------
__global__ void k_stream(const double* __restrict__ v, const int* __restrict__ c, size_t nnz, double* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; size_t stride = (size_t)gridDim.x * blockDim.x;
  double acc = 0;
  for (; i < nnz; i += stride) acc += v[i] * (double)c[i];
  if (acc == 1.2345e300) out[0] = acc;
}
__global__ void k_gather(const int* __restrict__ c, const double* __restrict__ x, size_t nnz, double* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; size_t stride = (size_t)gridDim.x * blockDim.x;
  double acc = 0;
  for (; i < nnz; i += stride) acc += x[c[i]];
  if (acc == 1.2345e300) out[0] = acc;
}

template <int W>
__global__ void k_csr_vec(const int* __restrict__ ptr, const int* __restrict__ idx, const double* __restrict__ val,
                          const double* __restrict__ x, double* __restrict__ y, int rows) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int row = t / W, lane = t % W;
  if (row >= rows) return;
  int b = ptr[row], e = ptr[row + 1];
  double acc = 0;
  for (int k = b + lane; k < e; k += W) acc += val[k] * x[idx[k]];
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o, W);
  if (lane == 0) y[row] = acc;
}

// SELL-32: slot-major. slice_ptr[s] = offset of slice s; element j of lane l at slice_ptr[s] + j*32 + l.
template <int U>
__global__ void __launch_bounds__(256) k_sell32(const int64_t* __restrict__ slice_ptr, const int* __restrict__ len,
                         const int* __restrict__ idx, const double* __restrict__ val,
                         const int* __restrict__ perm, const double* __restrict__ x, double* __restrict__ y, int nslots) {
  int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= nslots) return;
  int64_t base = slice_ptr[slot >> 5] + (slot & 31);
  int n = len[slot];
  double acc = 0;
  int j = 0;
  for (; j + U <= n; j += U) {
    double v[U]; int c[U]; double xv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { v[u] = val[base + (int64_t)(j + u) * 32]; c[u] = idx[base + (int64_t)(j + u) * 32]; }
#pragma unroll
    for (int u = 0; u < U; ++u) xv[u] = x[c[u]];
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u] * xv[u];
  }
  for (; j < n; ++j) acc += val[base + (int64_t)j * 32] * x[idx[base + (int64_t)j * 32]];
  int out = perm ? perm[slot] : slot;
  if (out >= 0) y[out] = acc;
}

// SELL-32 with 64-bit packed (index in low bits not used) variant omitted.

struct Sell { int nslots; std::vector<int64_t> slice_ptr; std::vector<int> len, perm, idx; std::vector<double> val; size_t padded; };
static Sell to_sell(const Csr& a, int sigma) {
  Sell s; int rows = a.rows; int nsl = (rows + 31) / 32; s.nslots = nsl * 32;
  s.perm.assign(s.nslots, -1); s.len.assign(s.nslots, 0);
  std::vector<int> order(rows); std::iota(order.begin(), order.end(), 0);
  if (sigma > 1) for (int w = 0; w < rows; w += sigma) {
    int e = std::min(rows, w + sigma);
    std::stable_sort(order.begin() + w, order.begin() + e, [&](int p, int q) { return (a.ptr[p + 1] - a.ptr[p]) > (a.ptr[q + 1] - a.ptr[q]); });
  }
  s.slice_ptr.assign(nsl + 1, 0);
  for (int sl = 0; sl < nsl; ++sl) {
    int w = 0;
    for (int l = 0; l < 32; ++l) { int slot = sl * 32 + l; if (slot < rows) { int r = order[slot]; s.perm[slot] = r; s.len[slot] = a.ptr[r + 1] - a.ptr[r]; w = std::max(w, s.len[slot]); } }
    s.slice_ptr[sl + 1] = s.slice_ptr[sl] + (int64_t)w * 32;
  }
  s.padded = s.slice_ptr[nsl]; s.idx.assign(s.padded, 0); s.val.assign(s.padded, 0.0);
  for (int slot = 0; slot < rows; ++slot) {
    int r = s.perm[slot]; int64_t base = s.slice_ptr[slot >> 5] + (slot & 31);
    for (int j = 0; j < s.len[slot]; ++j) { s.idx[base + (int64_t)j * 32] = a.idx[a.ptr[r] + j]; s.val[base + (int64_t)j * 32] = a.val[a.ptr[r] + j]; }
  }
  return s;
}

template <class T> T* up(const std::vector<T>& v) { T* d; CK(cudaMalloc(&d, v.size() * sizeof(T) + 16)); CK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice)); return d; }

template <class F> float timeit(F f, int reps = 20) {
  for (int i = 0; i < 3; ++i) f();
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  CK(cudaDeviceSynchronize()); cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b);
  CK(cudaGetLastError());
  return ms / reps;
}

static double check(const Csr& a, const std::vector<double>& x, const double* dy) {
  std::vector<double> y(a.rows); CK(cudaMemcpy(y.data(), dy, a.rows * 8, cudaMemcpyDeviceToHost));
  double maxrel = 0;
  for (int r = 0; r < a.rows; r += 997) { double s = 0, sa = 0; for (int k = a.ptr[r]; k < a.ptr[r + 1]; ++k) { s += a.val[k] * x[a.idx[k]]; sa += fabs(a.val[k] * x[a.idx[k]]); } if (sa > 0) maxrel = std::max(maxrel, fabs(s - y[r]) / sa); }
  return maxrel;
}

static void run_orientation(const char* name, const Csr& a) {
  size_t nnz = a.val.size();
  printf("== %s: rows=%d cols=%d nnz=%zu\n", name, a.rows, a.cols, nnz);
  std::vector<double> x(a.cols); for (auto& v : x) v = rnd01() - 0.5;
  int* dptr = up(a.ptr); int* didx = up(a.idx); double* dval = up(a.val); double* dx = up(x);
  double* dy; CK(cudaMalloc(&dy, (size_t)(a.rows + 64) * 8)); double* dout; CK(cudaMalloc(&dout, 8));
  double alg = 12.0 * nnz + 4.0 * a.rows + 8.0 * a.rows + 8.0 * a.cols;
  float ms;
  ms = timeit([&] { k_stream<<<148 * 16, 256>>>(dval, didx, nnz, dout); });
  printf("stream(vals+idx)        %8.3f ms  %8.1f GB/s\n", ms, 12.0 * nnz / ms * 1e-6);
  ms = timeit([&] { k_gather<<<148 * 16, 256>>>(didx, dx, nnz, dout); });
  printf("gather(idx+x[idx])      %8.3f ms  %8.1f GB/s(idx only) %8.1f Ggather/s\n", ms, 4.0 * nnz / ms * 1e-6, nnz / ms * 1e-6);
#define RUNVEC(W) { int threads = 256; long long tot = (long long)a.rows * W; int blocks = (int)((tot + threads - 1) / threads); \
    ms = timeit([&] { k_csr_vec<W><<<blocks, threads>>>(dptr, didx, dval, dx, dy, a.rows); }); \
    printf("csr_vec<%2d>             %8.3f ms  %8.1f GB/s alg  err=%.2e\n", W, ms, alg / ms * 1e-6, check(a, x, dy)); }
  RUNVEC(1) RUNVEC(2) RUNVEC(4) RUNVEC(8) RUNVEC(16) RUNVEC(32)
  for (int sigma : {1, 256, 4096}) {
    Sell s = to_sell(a, sigma);
    int64_t* dsp = up(s.slice_ptr); int* dlen = up(s.len); int* dperm = up(s.perm); int* dsi = up(s.idx); double* dsv = up(s.val);
    int blocks = (s.nslots + 255) / 256;
    double algs = 12.0 * s.padded + 8.0 * (s.nslots / 32) + 8.0 * a.rows + 8.0 * a.rows + 8.0 * a.cols;
    ms = timeit([&] { k_sell32<1><<<blocks, 256>>>(dsp, dlen, dsi, dsv, dperm, dx, dy, s.nslots); });
    printf("sell32 sigma=%-5d U=1   %8.3f ms  %8.1f GB/s alg(csr bytes) pad=%.3f err=%.2e\n", sigma, ms, alg / ms * 1e-6, (double)s.padded / nnz, check(a, x, dy));
    ms = timeit([&] { k_sell32<2><<<blocks, 256>>>(dsp, dlen, dsi, dsv, dperm, dx, dy, s.nslots); });
    printf("sell32 sigma=%-5d U=2   %8.3f ms  %8.1f GB/s\n", sigma, ms, alg / ms * 1e-6);
    ms = timeit([&] { k_sell32<4><<<blocks, 256>>>(dsp, dlen, dsi, dsv, dperm, dx, dy, s.nslots); });
    printf("sell32 sigma=%-5d U=4   %8.3f ms  %8.1f GB/s  (sell bytes %.1f MB)\n", sigma, ms, alg / ms * 1e-6, algs * 1e-6);
    // slot-ordered output (vectors kept permuted): no perm indirection
    ms = timeit([&] { k_sell32<4><<<blocks, 256>>>(dsp, dlen, dsi, dsv, nullptr, dx, dy, s.nslots); });
    printf("sell32 sigma=%-5d U=4 noperm %8.3f ms  %8.1f GB/s\n", sigma, ms, alg / ms * 1e-6);
    cudaFree(dsp); cudaFree(dlen); cudaFree(dperm); cudaFree(dsi); cudaFree(dsv);
  }
  cudaFree(dptr); cudaFree(didx); cudaFree(dval); cudaFree(dx); cudaFree(dy); cudaFree(dout);
}

int main(int argc, char** argv) {
  int m = argc > 1 ? atoi(argv[1]) : 1000000, n = argc > 2 ? atoi(argv[2]) : 2000000, per = argc > 3 ? atoi(argv[3]) : 20;
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s SMs=%d L2=%d MB\n", p.name, p.multiProcessorCount, p.l2CacheSize >> 20);
  Csr a; a.rows = m; a.cols = n; a.ptr.resize(m + 1); a.idx.resize((size_t)m * per); a.val.resize((size_t)m * per);
  for (int r = 0; r < m; ++r) {
    a.ptr[r] = r * per; int* c = &a.idx[(size_t)r * per];
    for (int j = 0; j < per; ++j) c[j] = (int)(rnd() % (uint64_t)n);
    std::sort(c, c + per);
    for (int j = 0; j < per; ++j) a.val[(size_t)r * per + j] = rnd01() * 2 - 1;
  }
  a.ptr[m] = m * per;
  // plain copy bandwidth reference
  { size_t N = 1ull << 28; double *s, *d; CK(cudaMalloc(&s, N * 8)); CK(cudaMalloc(&d, N * 8)); CK(cudaMemset(s, 1, N * 8));
    float ms = timeit([&] { cudaMemcpyAsync(d, s, N * 8, cudaMemcpyDeviceToDevice); }, 10);
    printf("memcpy d2d 2GiB: %.3f ms  %.1f GB/s (read+write)\n", ms, 2.0 * N * 8 / ms * 1e-6); cudaFree(s); cudaFree(d); }
  run_orientation("K rows (uniform 20/row)", a);
  Csr t = transpose(a);
  run_orientation("K^T rows (Poisson ~10/row)", t);
  return 0;
}
