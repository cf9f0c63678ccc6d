// comm.h -- the exchange step of the row-sharded solve: a thin
// C++ face over NCCL (NVLink 5 / NVSwitch), bound at run time with dlopen so
// the library needs no NCCL headers or link-time dependency (the caller passes
// the path of the NCCL shared object, e.g. the one bundled with PyTorch).
// One process per GPU; every call is enqueued on the caller's CUDA stream.
#ifndef PDLP_B200_COMM_H_
#define PDLP_B200_COMM_H_

#include <cstdint>
#include <string>

namespace pdlp_b200 {

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
  int64_t collectives() const { return collectives_; }

 private:
  struct Api;
  Api* api_ = nullptr;
  void* comm_ = nullptr;
  int rank_ = 0, world_ = 1, device_ = 0;
  int64_t collectives_ = 0;
};

}  // namespace pdlp_b200

#endif  // PDLP_B200_COMM_H_
