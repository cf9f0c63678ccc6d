// comm.cc -- see comm.h.
#include "comm.h"

#include <cuda_runtime_api.h>
#include <dlfcn.h>

#include <cstring>
#include <mutex>
#include <stdexcept>

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
using GetErrorStringFn = const char* (*)(int);
}  // namespace

struct Comm::Api {
  void* handle = nullptr;
  GetUniqueIdFn get_unique_id = nullptr;
  CommInitRankFn comm_init_rank = nullptr;
  CommDestroyFn comm_destroy = nullptr;
  AllReduceFn all_reduce = nullptr;
  GetErrorStringFn get_error_string = nullptr;

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
    a->get_error_string = reinterpret_cast<GetErrorStringFn>(dlsym(h, "ncclGetErrorString"));
    if (!a->get_unique_id || !a->comm_init_rank || !a->comm_destroy || !a->all_reduce) {
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
  if (comm_ != nullptr) api_->comm_destroy(comm_);
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

}  // namespace pdlp_b200
