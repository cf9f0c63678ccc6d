// comm.cc -- see comm.h.
#include "comm.h"

#include <cuda_runtime_api.h>
#include <dlfcn.h>

#include <cstring>
#include <mutex>
#include <stdexcept>
#include <vector>

namespace pdlp_b200 {

namespace {
// The few NCCL declarations used here (nccl.h of NCCL 2.x; ABI-stable values).
struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
constexpr int kNcclSum = 0, kNcclMax = 2, kNcclFloat64 = 8;
using GetUniqueIdFn = int (*)(NcclUniqueId*);
using CommInitRankFn = int (*)(NcclComm*, int, NcclUniqueId, int);
using CommDestroyFn = int (*)(NcclComm);
using AllReduceFn = int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
using AllGatherFn = int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t);
using GetErrorStringFn = const char* (*)(int);
constexpr int kNcclInt8 = 0;
}  // namespace

struct Comm::Api {
  void* handle = nullptr;
  GetUniqueIdFn get_unique_id = nullptr;
  CommInitRankFn comm_init_rank = nullptr;
  CommDestroyFn comm_destroy = nullptr;
  AllReduceFn all_reduce = nullptr;
  AllGatherFn all_gather = nullptr;
  GetErrorStringFn get_error_string = nullptr;
  typedef int (*CommGetAsyncErrorFn)(NcclComm, int*);
  typedef int (*CommAbortFn)(NcclComm);
  CommGetAsyncErrorFn comm_get_async_error = nullptr;  // optional (NCCL >= 2.4)
  CommAbortFn comm_abort = nullptr;
  typedef int (*GroupFn)();
  GroupFn group_start = nullptr, group_end = nullptr;  // optional: several collectives in one launch

  static Api* Load(const char* path) {
    static std::mutex mu;
    static Api* api = nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (api != nullptr) return api;
    const char* names[] = {path, "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
      if (n == nullptr || *n == 0) continue;
      h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (h != nullptr) break;
    }
    if (h == nullptr) throw std::runtime_error(std::string("cannot load NCCL: ") + dlerror());
    Api* a = new Api;
    a->handle = h;
    a->get_unique_id = reinterpret_cast<GetUniqueIdFn>(dlsym(h, "ncclGetUniqueId"));
    a->comm_init_rank = reinterpret_cast<CommInitRankFn>(dlsym(h, "ncclCommInitRank"));
    a->comm_destroy = reinterpret_cast<CommDestroyFn>(dlsym(h, "ncclCommDestroy"));
    a->all_reduce = reinterpret_cast<AllReduceFn>(dlsym(h, "ncclAllReduce"));
    a->all_gather = reinterpret_cast<AllGatherFn>(dlsym(h, "ncclAllGather"));
    a->get_error_string = reinterpret_cast<GetErrorStringFn>(dlsym(h, "ncclGetErrorString"));
    a->comm_get_async_error = reinterpret_cast<CommGetAsyncErrorFn>(dlsym(h, "ncclCommGetAsyncError"));
    a->comm_abort = reinterpret_cast<CommAbortFn>(dlsym(h, "ncclCommAbort"));
    a->group_start = reinterpret_cast<GroupFn>(dlsym(h, "ncclGroupStart"));
    a->group_end = reinterpret_cast<GroupFn>(dlsym(h, "ncclGroupEnd"));
    if (!a->get_unique_id || !a->comm_init_rank || !a->comm_destroy || !a->all_reduce || !a->all_gather) {
      delete a;
      throw std::runtime_error("the NCCL library lacks a required symbol");
    }
    api = a;
    return api;
  }
  void Check(int rc, const char* what) const {
    if (rc == 0) return;
    throw std::runtime_error(std::string("NCCL error in ") + what + ": " + (get_error_string ? get_error_string(rc) : "?"));
  }
};

void Comm::UniqueId(const char* nccl_library_path, uint8_t out_id[128]) {
  Api* api = Api::Load(nccl_library_path);
  NcclUniqueId id;
  api->Check(api->get_unique_id(&id), "ncclGetUniqueId");
  std::memcpy(out_id, id.internal, 128);
}

Comm::Comm(const char* nccl_library_path, int rank, int world_size, int cuda_device, const uint8_t unique_id[128])
    : api_(Api::Load(nccl_library_path)), rank_(rank), world_(world_size), device_(cuda_device) {
  if (world_size < 1 || rank < 0 || rank >= world_size) throw std::runtime_error("bad rank / world size");
  if (cudaSetDevice(cuda_device) != cudaSuccess) throw std::runtime_error("cudaSetDevice failed");
  NcclUniqueId id;
  std::memcpy(id.internal, unique_id, 128);
  NcclComm c = nullptr;
  api_->Check(api_->comm_init_rank(&c, world_size, id, rank), "ncclCommInitRank");
  comm_ = c;
}

Comm::~Comm() {
  if (cached_arena_ != nullptr) {
    try { DestroyPeerArena(cached_arena_, nullptr); } catch (...) {}
  }
  if (comm_ != nullptr) api_->comm_destroy(comm_);
}

// Failure detection: an error NCCL noticed asynchronously (a peer process died, a
// link went down) is turned into a CommError; the communicator is aborted first so that the other
// ranks' pending collectives fail instead of hanging.
void Comm::CheckAsyncError() {
  if (comm_ == nullptr || api_->comm_get_async_error == nullptr) return;
  int async = 0;
  const int rc = api_->comm_get_async_error(comm_, &async);
  if (rc == 0 && async == 0) return;
  const int code = rc != 0 ? rc : async;
  const std::string what = std::string("NCCL asynchronous error: ") + (api_->get_error_string ? api_->get_error_string(code) : "?");
  if (api_->comm_abort != nullptr) {
    api_->comm_abort(comm_);
    comm_ = nullptr;
  }
  throw CommError(what);
}

void Comm::AllReduceSum(const double* send, double* recv, int64_t count, void* stream) {
  if (count <= 0) return;
  api_->Check(api_->all_reduce(send, recv, static_cast<size_t>(count), kNcclFloat64, kNcclSum, comm_, static_cast<cudaStream_t>(stream)), "ncclAllReduce(sum)");
  ++collectives_;
}
void Comm::AllReduceMax(const double* send, double* recv, int64_t count, void* stream) {
  if (count <= 0) return;
  api_->Check(api_->all_reduce(send, recv, static_cast<size_t>(count), kNcclFloat64, kNcclMax, comm_, static_cast<cudaStream_t>(stream)), "ncclAllReduce(max)");
  ++collectives_;
}

void Comm::AllGatherBytes(const void* send, void* recv, int64_t bytes_per_rank, void* stream) {
  if (bytes_per_rank <= 0) return;
  api_->Check(api_->all_gather(send, recv, static_cast<size_t>(bytes_per_rank), kNcclInt8, comm_, static_cast<cudaStream_t>(stream)), "ncclAllGather");
  ++collectives_;
}
void Comm::GroupStart() { if (api_->group_start != nullptr && api_->group_end != nullptr) api_->Check(api_->group_start(), "ncclGroupStart"); }
void Comm::GroupEnd() { if (api_->group_start != nullptr && api_->group_end != nullptr) api_->Check(api_->group_end(), "ncclGroupEnd"); }
void Comm::AllGatherInPlace(double* buf, int64_t count_per_rank, void* stream) {
  if (count_per_rank <= 0) return;
  api_->Check(api_->all_gather(buf + static_cast<int64_t>(rank_) * count_per_rank, buf, static_cast<size_t>(count_per_rank), kNcclFloat64, comm_,
                               static_cast<cudaStream_t>(stream)), "ncclAllGather(in place)");
  ++collectives_;
}

namespace {
// min over ranks of a 0/1 flag, through the communicator (device scratch of 8 bytes)
bool AllAgree(Comm* c, bool ok, double* scratch, cudaStream_t s) {
  const double v = ok ? 0.0 : 1.0;
  if (cudaMemcpyAsync(scratch, &v, sizeof(double), cudaMemcpyHostToDevice, s) != cudaSuccess) return false;
  c->AllReduceMax(scratch, scratch, 1, s);
  double r = 1.0;
  if (cudaMemcpyAsync(&r, scratch, sizeof(double), cudaMemcpyDeviceToHost, s) != cudaSuccess) return false;
  if (cudaStreamSynchronize(s) != cudaSuccess) return false;
  return r == 0.0;
}
}  // namespace

PeerArena* Comm::CreatePeerArena(int64_t bytes, void* stream_v) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_v);
  if (world_ > kMaxPeers) return nullptr;
  PeerArena* a = new PeerArena;
  a->world = world_;
  a->rank = rank_;
  a->bytes = bytes;
  char* staging = nullptr;  // [world * 64] handles + 8 bytes scratch
  bool ok = cudaMalloc(reinterpret_cast<void**>(&staging), static_cast<size_t>(world_) * sizeof(cudaIpcMemHandle_t) + 64) == cudaSuccess;
  double* scratch = reinterpret_cast<double*>(staging + static_cast<size_t>(world_) * sizeof(cudaIpcMemHandle_t));
  void* local = nullptr;
  cudaIpcMemHandle_t mine;
  std::memset(&mine, 0, sizeof(mine));
  ok = ok && cudaMalloc(&local, static_cast<size_t>(bytes)) == cudaSuccess;
  ok = ok && cudaMemsetAsync(local, 0, static_cast<size_t>(bytes), s) == cudaSuccess;
  ok = ok && cudaIpcGetMemHandle(&mine, local) == cudaSuccess;
  if (staging == nullptr) {  // cannot even talk: every rank must still take part in the collectives below
    cudaGetLastError();
    throw std::runtime_error("cudaMalloc failed while creating the peer arena");
  }
  std::vector<cudaIpcMemHandle_t> all(world_);
  cudaMemcpyAsync(staging + static_cast<size_t>(rank_) * sizeof(mine), &mine, sizeof(mine), cudaMemcpyHostToDevice, s);
  AllGatherBytes(staging + static_cast<size_t>(rank_) * sizeof(mine), staging, sizeof(mine), s);
  cudaMemcpyAsync(all.data(), staging, static_cast<size_t>(world_) * sizeof(mine), cudaMemcpyDeviceToHost, s);
  cudaStreamSynchronize(s);
  ok = AllAgree(this, ok, scratch, s);  // all allocations + memsets done everywhere
  if (ok) {
    a->base[rank_] = local;
    for (int h = 0; h < world_ && ok; ++h) {
      if (h == rank_) continue;
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[h], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; cudaGetLastError(); break; }
      a->base[h] = p;
    }
    ok = AllAgree(this, ok, scratch, s);
  }
  if (!ok) {
    for (int h = 0; h < world_; ++h)
      if (h != rank_ && a->base[h] != nullptr) cudaIpcCloseMemHandle(a->base[h]);
    AllAgree(this, true, scratch, s);  // nobody frees before everybody has unmapped
    if (local != nullptr) cudaFree(local);
    cudaFree(staging);
    cudaGetLastError();
    delete a;
    return nullptr;
  }
  cudaFree(staging);
  return a;
}

PeerArena* Comm::AcquirePeerArena(int64_t bytes, void* stream_v) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_v);
  if (arena_unavailable_) return nullptr;
  if (cached_arena_busy_) return CreatePeerArena(bytes, s);
  if (cached_arena_ != nullptr && cached_arena_->bytes >= bytes) {
    // same size request on every rank (the layout depends only on n and the world size)
    cudaMemsetAsync(cached_arena_->base[rank_], 0, static_cast<size_t>(cached_arena_->bytes), s);
    double* scratch = nullptr;
    if (cudaMalloc(reinterpret_cast<void**>(&scratch), 64) != cudaSuccess) throw std::runtime_error("cudaMalloc failed");
    const bool ok = AllAgree(this, true, scratch, s);  // every arena is clean before anybody signals into it
    cudaFree(scratch);
    if (!ok) throw std::runtime_error("peer arena reset failed");
    cached_arena_busy_ = true;
    return cached_arena_;
  }
  if (cached_arena_ != nullptr) {
    DestroyPeerArena(cached_arena_, s);
    cached_arena_ = nullptr;
  }
  cached_arena_ = CreatePeerArena(bytes + bytes / 4, s);
  if (cached_arena_ == nullptr) arena_unavailable_ = true;
  cached_arena_busy_ = cached_arena_ != nullptr;
  return cached_arena_;
}

void Comm::ReleasePeerArena(PeerArena* a, void* stream_v) {
  if (a == nullptr) return;
  if (a == cached_arena_) {
    cudaStreamSynchronize(static_cast<cudaStream_t>(stream_v));
    cached_arena_busy_ = false;
    return;
  }
  DestroyPeerArena(a, stream_v);
}

void Comm::DestroyPeerArena(PeerArena* a, void* stream_v) {
  if (a == nullptr) return;
  cudaStream_t s = static_cast<cudaStream_t>(stream_v);
  cudaStreamSynchronize(s);
  for (int h = 0; h < a->world; ++h)
    if (h != a->rank && a->base[h] != nullptr) cudaIpcCloseMemHandle(a->base[h]);
  // nobody frees before everybody has unmapped
  double* scratch = nullptr;
  if (cudaMalloc(reinterpret_cast<void**>(&scratch), 64) == cudaSuccess) {
    try { AllAgree(this, true, scratch, s); } catch (...) {}
    cudaFree(scratch);
  }
  cudaFree(a->base[a->rank]);
  delete a;
}

}  // namespace pdlp_b200
