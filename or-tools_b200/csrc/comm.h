// comm.h -- the exchange step of the row-sharded solve: a thin
// C++ face over NCCL (NVLink 5 / NVSwitch), bound at run time with dlopen so
// the library needs no NCCL headers or link-time dependency (the caller passes
// the path of the NCCL shared object, e.g. the one bundled with PyTorch).
// One process per GPU; every call is enqueued on the caller's CUDA stream.
#ifndef PDLP_B200_COMM_H_
#define PDLP_B200_COMM_H_

#include <cstdint>
#include <stdexcept>
#include <string>

namespace pdlp_b200 {

// A failure of the communicator itself (as opposed to a bad argument): the solve ends with
// TERMINATION_REASON_OTHER on the rank that sees it.
struct CommError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// One allocation per rank that every rank of the box maps into its own address
// space (CUDA IPC over NVLink / NVSwitch peer memory): the fused exchange
// kernels of the step loop store x~ slices straight into every peer's arena
// and pull the K^T y' partial slices out of them, so the per-step exchange
// needs no collective library call.
constexpr int kMaxPeers = 8;
struct PeerArena {
  int world = 1, rank = 0;
  int64_t bytes = 0;
  void* base[kMaxPeers] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // base[rank] is the local allocation
};

class Comm {
 public:
  // ncclGetUniqueId on rank 0; the caller broadcasts the 128 bytes.
  static void UniqueId(const char* nccl_library_path, uint8_t out_id[128]);
  // ncclCommInitRank. cudaSetDevice(cuda_device) must be current / is set here.
  Comm(const char* nccl_library_path, int rank, int world_size, int cuda_device, const uint8_t unique_id[128]);
  ~Comm();
  Comm(const Comm&) = delete;
  Comm& operator=(const Comm&) = delete;

  int rank() const { return rank_; }
  int world_size() const { return world_; }
  int cuda_device() const { return device_; }
  // fp64 all-reduce on `stream` (in place when send == recv).
  void AllReduceSum(const double* send, double* recv, int64_t count, void* stream);
  void AllReduceMax(const double* send, double* recv, int64_t count, void* stream);
  // Every rank contributes `bytes_per_rank` bytes (device memory); recv holds world * bytes_per_rank.
  void AllGatherBytes(const void* send, void* recv, int64_t bytes_per_rank, void* stream);
  // In-place all-gather of equal fp64 slices: rank r's slice is buf[r * count_per_rank ...).
  void AllGatherInPlace(double* buf, int64_t count_per_rank, void* stream);
  // ncclGroupStart / ncclGroupEnd around several collectives: one launch instead of one each.
  void GroupStart();
  void GroupEnd();
  // Collective: allocates `bytes` (zero-filled) on every rank and maps all of
  // them everywhere. Returns nullptr (on every rank) when peer mapping is not
  // possible on this box; the caller then keeps the NCCL exchange.
  PeerArena* CreatePeerArena(int64_t bytes, void* stream);
  void DestroyPeerArena(PeerArena* arena, void* stream);
  // Collective: a zero-filled arena of at least `bytes`, owned by the
  // communicator and reused across solves (mapping peer memory costs ~1 s, far
  // more than a solve); nullptr when peer mapping is not possible.
  // While one holder has it, a second request gets a freshly mapped arena of
  // its own. Every holder hands its arena back with ReleasePeerArena.
  PeerArena* AcquirePeerArena(int64_t bytes, void* stream);
  void ReleasePeerArena(PeerArena* arena, void* stream);
  int64_t collectives() const { return collectives_; }
  // ncclCommGetAsyncError: throws CommError (after ncclCommAbort) if NCCL has seen an
  // asynchronous failure; polled by the solver at every checkpoint of a row-sharded solve.
  void CheckAsyncError();

 private:
  struct Api;
  Api* api_ = nullptr;
  void* comm_ = nullptr;
  int rank_ = 0, world_ = 1, device_ = 0;
  int64_t collectives_ = 0;
  PeerArena* cached_arena_ = nullptr;
  bool arena_unavailable_ = false;
  bool cached_arena_busy_ = false;
};

}  // namespace pdlp_b200

#endif  // PDLP_B200_COMM_H_
