// solver.h -- host C++ driver of the device PDHG solve.
#ifndef PDLP_B200_SOLVER_H_
#define PDLP_B200_SOLVER_H_

#include <functional>
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

struct SolverResultCpp {
  std::vector<double> primal_solution, dual_solution, reduced_costs;
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
