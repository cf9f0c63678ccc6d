// solver.h -- host C++ driver of the device PDHG solve.
#ifndef PDLP_B200_SOLVER_H_
#define PDLP_B200_SOLVER_H_

#include <cstdlib>
#include <cstring>
#include <functional>
#include <new>
#include <memory>
#include <optional>
#include <string>
#include <vector>

#include "../../include/pdlp_b200.h"

namespace pdlp_b200 {

class Comm;  // comm.h

// FeasibilityPolishingDetails (solve_log.proto:371-383)
struct PolishingDetailsCpp {
  int polishing_phase_type = 0;
  int main_iteration_count = 0;
  PdlpParams params{};
  int termination_reason = 0;
  int iteration_count = 0;
  double solve_time_sec = 0;
  PdlpIterationStats solution_stats{};
  int solution_type = 0;
  std::vector<PdlpIterationStats> iteration_stats;
};

struct SolveLogCpp {
  std::optional<std::string> instance_name;
  int termination_reason = PDLP_TERMINATION_REASON_UNSPECIFIED;
  std::optional<std::string> termination_string;
  int iteration_count = 0;
  double solve_time_sec = 0, preprocessing_time_sec = 0;
  int solution_type = PDLP_POINT_TYPE_UNSPECIFIED;
  bool has_solution_stats = false;
  PdlpIterationStats solution_stats{};
  bool has_original_stats = false, has_preprocessed_stats = false;
  PdlpQuadraticProgramStats original_stats{}, preprocessed_stats{};
  std::vector<PdlpIterationStats> iteration_stats;
  PdlpParams params{};
  std::vector<PolishingDetailsCpp> feasibility_polishing_details;
  int64_t gpu_kernel_launches = 0;
  double device_iteration_time_sec = 0;
};

// A malloc-backed array of doubles whose storage can be handed to the C ABI result without a copy
// (PdlpResult's arrays are released with free()); resize leaves the elements uninitialised (they
// are the destination of a device->host copy).
class HostVec {
 public:
  HostVec() = default;
  HostVec(const HostVec& o) { assign(o.p_, o.n_); }
  HostVec(HostVec&& o) noexcept : p_(o.p_), n_(o.n_) { o.p_ = nullptr; o.n_ = 0; }
  HostVec& operator=(const HostVec& o) { if (this != &o) assign(o.p_, o.n_); return *this; }
  HostVec& operator=(HostVec&& o) noexcept { if (this != &o) { std::free(p_); p_ = o.p_; n_ = o.n_; o.p_ = nullptr; o.n_ = 0; } return *this; }
  ~HostVec() { std::free(p_); }
  void resize(size_t n) {
    std::free(p_);
    p_ = n > 0 ? static_cast<double*>(std::malloc(n * sizeof(double))) : nullptr;
    if (n > 0 && p_ == nullptr) throw std::bad_alloc();
    n_ = n;
  }
  double* data() { return p_; }
  const double* data() const { return p_; }
  size_t size() const { return n_; }
  bool empty() const { return n_ == 0; }
  double* release() { double* p = p_; p_ = nullptr; n_ = 0; return p; }  // the caller frees it with free()

 private:
  void assign(const double* p, size_t n) {
    resize(n);
    if (n > 0) std::memcpy(p_, p, n * sizeof(double));
  }
  double* p_ = nullptr;
  size_t n_ = 0;
};

struct SolverResultCpp {
  HostVec primal_solution, dual_solution, reduced_costs;
  SolveLogCpp solve_log;
};

struct InitialSolution {
  std::vector<double> primal, dual;
};

struct Logger {
  PdlpMessageCallback cb = nullptr;
  void* user = nullptr;
  void Log(const std::string& s) const;
};

using StatsCallback = std::function<void(const PdlpIterationCallbackInfo&)>;

// params.cc
void SetDefaultParams(PdlpParams* p);
std::string ValidateParams(const PdlpParams& p);  // "" if valid, else the reference's message
// "" if the CSC view is well-formed (dimensions, col_starts[0] == 0, monotone col_starts,
// col_starts[n] == num_nonzeros, 0 <= row < m); every host-side walker of the arrays calls it first.
std::string ValidateView(const PdlpProblemView& v);

// PrimalDualHybridGradient (pdhg.cc:3107-3152) on the CUDA device. Throws
// std::runtime_error only for CUDA / device failures; every solver-level
// outcome (including invalid input) is reported through the returned log.
SolverResultCpp PrimalDualHybridGradient(const PdlpProblemView& view, const PdlpParams& params,
                                         std::optional<InitialSolution> initial_solution,
                                         const volatile int32_t* interrupt_solve, const Logger& logger,
                                         StatsCallback callback, int cuda_device, Comm* comm = nullptr);

// A resumable solve whose problem and iterates stay resident in HBM between
// calls (C ABI: pdlp_b200_session_*). Advance() runs the same loop as
// PrimalDualHybridGradient; stopping and resuming does not change the iterates.
class SolveSession {
 public:
  static std::unique_ptr<SolveSession> Create(const PdlpProblemView& view, const PdlpParams& params,
                                              std::optional<InitialSolution> initial_solution, const Logger& logger,
                                              StatsCallback callback, int cuda_device, Comm* comm = nullptr);
  ~SolveSession();
  // Runs until `target_iterations` iterations are completed or the solve
  // terminates; returns true once terminated.
  bool Advance(int target_iterations, const volatile int32_t* interrupt_solve);
  void EnableTiming(bool on, int stride);
  void Status(PdlpSessionStatus* out) const;
  // The SolverResult; if the solve has not terminated it is stopped as if
  // interrupted by the user.
  SolverResultCpp Finish();

 private:
  SolveSession();
  struct Impl;
  std::unique_ptr<Impl> impl_;
};

}  // namespace pdlp_b200

#endif  // PDLP_B200_SOLVER_H_
