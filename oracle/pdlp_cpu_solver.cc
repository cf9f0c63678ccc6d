// oracle/pdlp_cpu_solver.cc
//
// TEST INFRASTRUCTURE ONLY -- NOT PART OF THE PRODUCT (see pdlp_cpu_core.h).
// CPU restatement of the PDLP driver: PrimalDualHybridGradient(),
// PreprocessSolver and Solver of ortools/pdlp/primal_dual_hybrid_gradient.cc,
// multithreaded with the same shard-parallel structure as the reference, plus
// a C ABI (prefix pdlp_oracle_) with the same POD structs as the product so the
// parity tests can call both sides identically. Doubles as the timed CPU
// baseline in bench.py (kind "port"): the reference itself is unbuildable in
// this image.
//
// Not restated (host feature outside the hot path): glop presolve
// -> TERMINATION_REASON_INVALID_PARAMETER. Feasibility polishing is restated
// (pdhg.cc:2676-3015) and pinned by the reference's FeasibilityPolishing tests.
#include <chrono>
#include <cstdarg>
#include <memory>

#include "pdlp_cpu_core.h"

namespace pdlp_oracle {
namespace {

using Clock = std::chrono::steady_clock;
struct WallTimer {
  Clock::time_point start = Clock::now();
  double accumulated = 0;
  bool running = true;
  void Start() { start = Clock::now(); accumulated = 0; running = true; }
  void Stop() { if (running) { accumulated += std::chrono::duration<double>(Clock::now() - start).count(); running = false; } }
  void Resume() { if (!running) { start = Clock::now(); running = true; } }
  double Get() const { return accumulated + (running ? std::chrono::duration<double>(Clock::now() - start).count() : 0.0); }
};

struct Logger {
  PdlpMessageCallback cb = nullptr;
  void* user = nullptr;
  void Log(const std::string& s) const {
    if (cb != nullptr) cb(s.c_str(), user);
    else { std::fputs(s.c_str(), stdout); std::fputc('\n', stdout); }
  }
};

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
};
struct SolverResultCpp {
  Vec primal_solution, dual_solution, reduced_costs;
  SolveLogCpp solve_log;
};
struct InitialSolution { Vec primal, dual; };

using StatsCallback = std::function<void(const PdlpIterationCallbackInfo&)>;

std::string Fmt(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  std::vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  return buf;
}

// pdhg.cc:776-785
SolverResultCpp ErrorSolverResult(int reason, const std::string& message, const Logger& logger) {
  SolverResultCpp r;
  r.solve_log.termination_reason = reason;
  r.solve_log.termination_string = message;
  logger.Log("The solver did not run because of invalid input: " + message);
  return r;
}

// pdhg.cc:100-125
int NumThreads(int num_threads, int num_shards, const QuadraticProgram& qp, const Logger& logger) {
  int capped = num_threads;
  if (num_shards > 0) capped = std::min(capped, num_shards);
  const int64_t limit = std::max<int64_t>(qp.variable_lower_bounds.size(), qp.constraint_lower_bounds.size());
  capped = static_cast<int>(std::min<int64_t>(capped, limit));
  capped = std::max(capped, 1);
  if (capped != num_threads)
    logger.Log(Fmt("WARNING: Reducing num_threads from %d to %d because additional threads would be useless.", num_threads, capped));
  return capped;
}
int NumShards(int num_threads, int num_shards) {
  if (num_shards > 0) return num_shards;
  return num_threads == 1 ? 1 : 4 * num_threads;
}

const PdlpConvergenceInformation* GetConvergenceInformation(const PdlpIterationStats& s, int type) {
  for (int i = 0; i < s.num_convergence_information; ++i)
    if (s.convergence_information[i].candidate_type == type) return &s.convergence_information[i];
  return nullptr;
}

// pdhg.cc:127-318 (log table).
void LogIterationStatsHeader(int verbosity, const Logger& logger) {
  const std::string work = verbosity >= 3 ? Fmt("%6s %8s %6s", "iter#", "kkt_pass", "time") : Fmt("%6s %6s", "iter#", "time");
  const std::string conv = verbosity >= 3
      ? Fmt("%12s %12s %12s | %12s %12s %12s | %12s %12s | %12s %12s", "rel_prim_res", "rel_dual_res", "rel_gap", "prim_resid",
            "dual_resid", "obj_gap", "prim_obj", "dual_obj", "prim_var_l2", "dual_var_l2")
      : Fmt("%10s %10s %10s | %10s %10s", "rel_p_res", "rel_d_res", "rel_gap", "prim_obj", "dual_obj");
  logger.Log(std::string(verbosity >= 4 ? "I " : "") + work + " | " + conv);
}
void LogIterationStats(int verbosity, const PdlpIterationStats& st, const PdlpTerminationCriteria& tc, const PdlpBoundNorms& bn,
                       int preferred, const Logger& logger) {
  const std::string iter = verbosity >= 3 ? Fmt("%6d %8.1f %6.1f", st.iteration_number, st.cumulative_kkt_matrix_passes, st.cumulative_time_sec)
                                          : Fmt("%6d %6.1f", st.iteration_number, st.cumulative_time_sec);
  const PdlpConvergenceInformation* ci = GetConvergenceInformation(st, preferred);
  if (ci == nullptr && st.num_convergence_information > 0) ci = &st.convergence_information[0];
  if (ci == nullptr) { logger.Log(std::string(verbosity >= 4 ? "? " : "") + iter); return; }
  const char* tag = "";
  if (verbosity >= 4) {
    switch (ci->candidate_type) {
      case PDLP_POINT_TYPE_CURRENT_ITERATE: tag = "C "; break;
      case PDLP_POINT_TYPE_AVERAGE_ITERATE: tag = "A "; break;
      case PDLP_POINT_TYPE_ITERATE_DIFFERENCE: tag = "D "; break;
      default: tag = "? ";
    }
  }
  const RelativeConvergenceInformation rel = ComputeRelativeResiduals(EffectiveOptimalityCriteria(tc), *ci, bn);
  double rp, rd, ap, ad;
  switch (tc.optimality_norm) {
    case PDLP_OPTIMALITY_NORM_L_INF: rp = rel.relative_l_inf_primal_residual; rd = rel.relative_l_inf_dual_residual; ap = ci->l_inf_primal_residual; ad = ci->l_inf_dual_residual; break;
    case PDLP_OPTIMALITY_NORM_L_INF_COMPONENTWISE: rp = ci->l_inf_componentwise_primal_residual; rd = ci->l_inf_componentwise_dual_residual; ap = ci->l_inf_primal_residual; ad = ci->l_inf_dual_residual; break;
    default: rp = rel.relative_l2_primal_residual; rd = rel.relative_l2_dual_residual; ap = ci->l2_primal_residual; ad = ci->l2_dual_residual;
  }
  const std::string conv = verbosity >= 3
      ? Fmt("%#12.6g %#12.6g %#12.6g | %#12.6g %#12.6g %#12.6g | %#12.6g %#12.6g | %#12.6g %#12.6g", rp, rd, rel.relative_optimality_gap, ap, ad,
            ci->primal_objective - ci->dual_objective, ci->primal_objective, ci->dual_objective, ci->l2_primal_variable, ci->l2_dual_variable)
      : Fmt("%#10.4g %#10.4g %#10.4g | %#10.4g %#10.4g", rp, rd, rel.relative_optimality_gap, ci->primal_objective, ci->dual_objective);
  logger.Log(std::string(tag) + iter + " | " + conv);
}

// pdhg.cc:791-983
std::optional<SolverResultCpp> CheckProblemStats(const PdlpQuadraticProgramStats& s, double objective_offset, bool check_small, const Logger& logger) {
  const double kBig = 1e50, kSmall = 1e-50, kRange = 1e20;
  auto err = [&](const std::string& m) { return ErrorSolverResult(PDLP_TERMINATION_REASON_INVALID_PROBLEM, m, logger); };
  if (std::isnan(s.constraint_matrix_l2_norm)) return err("Constraint matrix has a NAN.");
  if (s.constraint_matrix_abs_max > kBig) return err("Constraint matrix has a non-zero with absolute value " + FormatDouble(s.constraint_matrix_abs_max) + " which exceeds limit of " + FormatDouble(kBig) + ".");
  if (s.constraint_matrix_abs_max > kRange * s.constraint_matrix_abs_min)
    logger.Log("WARNING: Constraint matrix has largest absolute value " + FormatDouble(s.constraint_matrix_abs_max) + " and smallest non-zero absolute value " + FormatDouble(s.constraint_matrix_abs_min) + " performance may suffer.");
  if (s.constraint_matrix_col_min_l_inf_norm > 0 && s.constraint_matrix_col_min_l_inf_norm < kSmall)
    return err("Constraint matrix has a column with Linf norm " + FormatDouble(s.constraint_matrix_col_min_l_inf_norm) + " which is less than limit of " + FormatDouble(kSmall) + ".");
  if (s.constraint_matrix_row_min_l_inf_norm > 0 && s.constraint_matrix_row_min_l_inf_norm < kSmall)
    return err("Constraint matrix has a row with Linf norm " + FormatDouble(s.constraint_matrix_row_min_l_inf_norm) + " which is less than limit of " + FormatDouble(kSmall) + ".");
  if (std::isnan(s.combined_bounds_l2_norm)) return err("Constraint bounds vector has a NAN.");
  if (s.combined_bounds_max > kBig) return err("Combined constraint bounds vector has a non-zero with absolute value " + FormatDouble(s.combined_bounds_max) + " which exceeds limit of " + FormatDouble(kBig) + ".");
  if (check_small && s.combined_bounds_min > 0 && s.combined_bounds_min < kSmall)
    return err("Combined constraint bounds vector has a non-zero with absolute value " + FormatDouble(s.combined_bounds_min) + " which is less than the limit of " + FormatDouble(kSmall) + ".");
  if (s.combined_bounds_max > kRange * s.combined_bounds_min) logger.Log("WARNING: Combined constraint bounds vector has a large dynamic range; performance may suffer.");
  if (std::isnan(s.combined_variable_bounds_l2_norm)) return err("Variable bounds vector has a NAN.");
  if (s.combined_variable_bounds_max > kBig) return err("Combined variable bounds vector has a non-zero with absolute value " + FormatDouble(s.combined_variable_bounds_max) + " which exceeds limit of " + FormatDouble(kBig) + ".");
  if (check_small && s.combined_variable_bounds_min > 0 && s.combined_variable_bounds_min < kSmall)
    return err("Combined variable bounds vector has a non-zero with absolute value " + FormatDouble(s.combined_variable_bounds_min) + " which is less than the limit of " + FormatDouble(kSmall) + ".");
  if (s.combined_variable_bounds_max > kRange * s.combined_variable_bounds_min) logger.Log("WARNING: Combined variable bounds vector has a large dynamic range; performance may suffer.");
  if (s.variable_bound_gaps_max > kRange * s.variable_bound_gaps_min) logger.Log("WARNING: Variable bound gap vector has a large dynamic range; performance may suffer.");
  if (std::isnan(objective_offset)) return err("Objective offset is NAN.");
  if (std::abs(objective_offset) > kBig) return err("Objective offset " + FormatDouble(objective_offset) + " has absolute value which exceeds limit of " + FormatDouble(kBig) + ".");
  if (std::isnan(s.objective_vector_l2_norm)) return err("Objective vector has a NAN.");
  if (s.objective_vector_abs_max > kBig) return err("Objective vector has a non-zero with absolute value " + FormatDouble(s.objective_vector_abs_max) + " which exceeds limit of " + FormatDouble(kBig) + ".");
  if (check_small && s.objective_vector_abs_min > 0 && s.objective_vector_abs_min < kSmall)
    return err("Objective vector has a non-zero with absolute value " + FormatDouble(s.objective_vector_abs_min) + " which is less than the limit of " + FormatDouble(kSmall) + ".");
  if (s.objective_vector_abs_max > kRange * s.objective_vector_abs_min) logger.Log("WARNING: Objective vector has a large dynamic range; performance may suffer.");
  if (std::isnan(s.objective_matrix_l2_norm)) return err("Objective matrix has a NAN.");
  if (s.objective_matrix_abs_max > kBig) return err("Objective matrix has a non-zero with absolute value " + FormatDouble(s.objective_matrix_abs_max) + " which exceeds limit of " + FormatDouble(kBig) + ".");
  if (s.objective_matrix_abs_max > kRange * s.objective_matrix_abs_min) logger.Log("WARNING: Objective matrix has a large dynamic range; performance may suffer.");
  return std::nullopt;
}

// pdhg.cc:985-1037
std::optional<SolverResultCpp> CheckInitialSolution(const ShardedQp& sqp, const InitialSolution& init, const Logger& logger) {
  const double kBig = 1e50;
  auto err = [&](const std::string& m) { return ErrorSolverResult(PDLP_TERMINATION_REASON_INVALID_INITIAL_SOLUTION, m, logger); };
  if (static_cast<int64_t>(init.primal.size()) != sqp.PrimalSize())
    return err(Fmt("Initial primal solution has size %lld which differs from problem primal size %lld", (long long)init.primal.size(), (long long)sqp.PrimalSize()));
  if (std::isnan(Norm(init.primal, sqp.PrimalSharder()))) return err("Initial primal solution has a NAN.");
  if (const double n = LInfNorm(init.primal, sqp.PrimalSharder()); n > kBig)
    return err("Initial primal solution has an entry with absolute value " + FormatDouble(n) + " which exceeds limit of " + FormatDouble(kBig));
  if (static_cast<int64_t>(init.dual.size()) != sqp.DualSize())
    return err(Fmt("Initial dual solution has size %lld which differs from problem dual size %lld", (long long)init.dual.size(), (long long)sqp.DualSize()));
  if (std::isnan(Norm(init.dual, sqp.DualSharder()))) return err("Initial dual solution has a NAN.");
  if (const double n = LInfNorm(init.dual, sqp.DualSharder()); n > kBig)
    return err("Initial dual solution has an entry with absolute value " + FormatDouble(n) + " which exceeds limit of " + FormatDouble(kBig));
  return std::nullopt;
}

// SetActiveSetInformation, pdhg.cc:1476-1545.
void SetActiveSetInformation(const ShardedQp& sqp, const Vec& primal, const Vec& dual, const Vec& primal_start, const Vec& dual_start, PdlpPointMetadata& md) {
  const QuadraticProgram& qp = sqp.Qp();
  auto p_active = [&](const Vec& v, int64_t i) { return v[i] > qp.variable_lower_bounds[i] && v[i] < qp.variable_upper_bounds[i]; };
  auto d_active = [&](const Vec& v, int64_t i) { return v[i] != 0.0 || (qp.constraint_lower_bounds[i] == -kInf && qp.constraint_upper_bounds[i] == kInf); };
  md.active_primal_variable_count = static_cast<int64_t>(sqp.PrimalSharder().SumOverShards([&](int, int64_t b, int64_t e) { double c = 0; for (int64_t i = b; i < e; ++i) c += p_active(primal, i); return c; }));
  md.active_primal_variable_change = static_cast<int64_t>(sqp.PrimalSharder().SumOverShards([&](int, int64_t b, int64_t e) { double c = 0; for (int64_t i = b; i < e; ++i) c += (p_active(primal, i) != p_active(primal_start, i)); return c; }));
  md.active_dual_variable_count = static_cast<int64_t>(sqp.DualSharder().SumOverShards([&](int, int64_t b, int64_t e) { double c = 0; for (int64_t i = b; i < e; ++i) c += d_active(dual, i); return c; }));
  md.active_dual_variable_change = static_cast<int64_t>(sqp.DualSharder().SumOverShards([&](int, int64_t b, int64_t e) { double c = 0; for (int64_t i = b; i < e; ++i) c += (d_active(dual, i) != d_active(dual_start, i)); return c; }));
  md.has_active_set_information = 1;
}
// RandomProjection, iteration_stats.cc:355-380 (per-shard std::mt19937 seeded
// from a seed generator; the reference draws with absl::Gaussian -- values are
// unpinned by the reference's tests, only the counts are).
double RandomProjection(const Vec& v, const Sharder& sharder, std::mt19937& seed_gen) {
  std::vector<uint32_t> seeds(sharder.NumShards());
  for (auto& s : seeds) s = static_cast<uint32_t>(seed_gen());
  Vec dot(sharder.NumShards(), 0.0), nsq(sharder.NumShards(), 0.0);
  sharder.ForEachShard([&](int s, int64_t b, int64_t e) {
    std::mt19937 gen(seeds[s]);
    std::normal_distribution<double> g(0.0, 1.0);
    double d = 0, q = 0;
    for (int64_t i = b; i < e; ++i) { const double z = g(gen); d += z * v[i]; q += z * z; }
    dot[s] = d; nsq[s] = q;
  });
  double d = 0, q = 0;
  for (int s = 0; s < sharder.NumShards(); ++s) { d += dot[s]; q += nsq[s]; }
  return d / std::sqrt(q);
}

// ---- feasibility polishing helpers (pdhg.cc:2298-2358, 2684-2700, 2867-2886) ------------
PdlpIterationStats AddWorkStats(PdlpIterationStats stats, const PdlpIterationStats& more) {
  stats.iteration_number += more.iteration_number;
  stats.cumulative_kkt_matrix_passes += more.cumulative_kkt_matrix_passes;
  stats.cumulative_rejected_steps += more.cumulative_rejected_steps;
  stats.cumulative_time_sec += more.cumulative_time_sec;
  return stats;
}
PdlpIterationStats WorkFromFeasibilityPolishing(const SolveLogCpp& log) {
  PdlpIterationStats result;
  std::memset(&result, 0, sizeof(result));
  for (const PolishingDetailsCpp& d : log.feasibility_polishing_details) result = AddWorkStats(result, d.solution_stats);
  return result;
}
bool TerminationReasonIsInterrupted(int reason) { return reason == PDLP_TERMINATION_REASON_INTERRUPTED_BY_USER; }
bool TerminationReasonIsWorkLimitNotInterrupted(int reason) {
  return reason == PDLP_TERMINATION_REASON_ITERATION_LIMIT || reason == PDLP_TERMINATION_REASON_TIME_LIMIT ||
         reason == PDLP_TERMINATION_REASON_KKT_MATRIX_PASS_LIMIT;
}
bool TerminationReasonIsWorkLimit(int reason) { return TerminationReasonIsWorkLimitNotInterrupted(reason) || TerminationReasonIsInterrupted(reason); }
bool DoFeasibilityPolishingAfterLimitsReached(const PdlpParams& params, int reason) {
  if (TerminationReasonIsWorkLimitNotInterrupted(reason)) return params.apply_feasibility_polishing_after_limits_reached != 0;
  if (TerminationReasonIsInterrupted(reason)) return params.apply_feasibility_polishing_if_solver_is_interrupted != 0;
  return false;
}
PolishingDetailsCpp BuildFeasibilityPolishingDetails(int phase_type, int iteration_count, const PdlpParams& params, const SolveLogCpp& log) {
  PolishingDetailsCpp d;
  d.polishing_phase_type = phase_type;
  d.main_iteration_count = iteration_count;
  d.params = params;
  d.termination_reason = log.termination_reason;
  d.iteration_count = log.iteration_count;
  d.solve_time_sec = log.solve_time_sec;
  d.solution_stats = log.solution_stats;
  d.solution_type = log.solution_type;
  d.iteration_stats = log.iteration_stats;
  return d;
}
PdlpTerminationCriteria ReduceWorkLimitsByPreviousWork(PdlpTerminationCriteria criteria, int iteration_limit, const PdlpIterationStats& previous_work,
                                                       bool apply_after_limits_reached) {
  if (apply_after_limits_reached) {
    criteria.iteration_limit = iteration_limit;
    criteria.kkt_matrix_pass_limit = kInf;
    criteria.time_sec_limit = kInf;
  } else {
    criteria.iteration_limit = std::max(0, std::min(iteration_limit, criteria.iteration_limit - previous_work.iteration_number));
    criteria.kkt_matrix_pass_limit = std::max(0.0, criteria.kkt_matrix_pass_limit - previous_work.cumulative_kkt_matrix_passes);
    criteria.time_sec_limit = std::max(0.0, criteria.time_sec_limit - previous_work.cumulative_time_sec);
  }
  return criteria;
}
// The detailed criteria of a polishing phase: `base` with the listed tolerances made infinite.
void SetDetailedCriteria(PdlpTerminationCriteria& c, const DetailedCriteria& d) {
  c.optimality_criteria_case = PDLP_DETAILED_OPTIMALITY_CRITERIA;
  c.eps_optimal_primal_residual_absolute = d.primal_abs;
  c.eps_optimal_primal_residual_relative = d.primal_rel;
  c.eps_optimal_dual_residual_absolute = d.dual_abs;
  c.eps_optimal_dual_residual_relative = d.dual_rel;
  c.eps_optimal_objective_gap_absolute = d.gap_abs;
  c.eps_optimal_objective_gap_relative = d.gap_rel;
  c.has_eps_optimal_absolute = 0;
  c.has_eps_optimal_relative = 0;
}
Vec MapFiniteValuesToZero(const Vec& in) {
  Vec out(in.size());
  for (size_t i = 0; i < in.size(); ++i) out[i] = std::isfinite(in[i]) ? 0.0 : in[i];
  return out;
}
SolverResultCpp ConstructSolverResult(Vec primal, Vec dual, const PdlpIterationStats& stats, int reason, int output_type, SolveLogCpp log) {  // pdhg.cc:329-342
  log.iteration_count = stats.iteration_number;
  log.termination_reason = reason;
  log.solution_type = output_type;
  log.solve_time_sec = stats.cumulative_time_sec;
  log.solution_stats = stats;
  log.has_solution_stats = true;
  SolverResultCpp r;
  r.primal_solution = std::move(primal);
  r.dual_solution = std::move(dual);
  r.solve_log = std::move(log);
  return r;
}

class Solver;

// ---------------------------------------------------------------------------
// PreprocessSolver (pdhg.cc:345-535)
// ---------------------------------------------------------------------------
class PreprocessSolver {
 public:
  PreprocessSolver(QuadraticProgram qp, const PdlpParams& params, const Logger* logger)
      : num_threads_(NumThreads(params.num_threads, params.num_shards, qp, *logger)),
        num_shards_(NumShards(num_threads_, params.num_shards)),
        sharded_qp_(std::move(qp), num_threads_, num_shards_),
        logger_(*logger) {}

  SolverResultCpp PreprocessAndSolve(const PdlpParams& params, std::optional<InitialSolution> initial_solution,
                                     const volatile int32_t* interrupt_solve, StatsCallback callback);

  std::optional<TerminationReasonAndPointType> UpdateIterationStatsAndCheckTermination(
      const PdlpParams& params, bool force_numerical_termination, const Vec& primal_current, const Vec& dual_current,
      const Vec* primal_average, const Vec* dual_average, const Vec* primal_delta, const Vec* dual_delta,
      const Vec& last_primal_start, const Vec& last_dual_start, const volatile int32_t* interrupt_solve, int iteration_type,
      const PdlpIterationStats& full_stats, PdlpIterationStats& stats);

  void ComputeConvergenceAndInfeasibilityFromWorkingSolution(const PdlpParams& params, const Vec& working_primal, const Vec& working_dual,
                                                             int candidate_type, PdlpConvergenceInformation* conv,
                                                             PdlpInfeasibilityInformation* infeas) const;
  SolverResultCpp ConstructOriginalSolverResult(const PdlpParams& params, SolverResultCpp result) const;

  const ShardedQp& ShardedWorkingQp() const { return sharded_qp_; }
  // pdhg.cc:420-443: exchange bounds / objective with the working problem (feasibility polishing)
  void SwapVariableBounds(Vec& lower, Vec& upper) {
    std::swap(sharded_qp_.MutableQp().variable_lower_bounds, lower);
    std::swap(sharded_qp_.MutableQp().variable_upper_bounds, upper);
  }
  void SwapConstraintBounds(Vec& lower, Vec& upper) {
    std::swap(sharded_qp_.MutableQp().constraint_lower_bounds, lower);
    std::swap(sharded_qp_.MutableQp().constraint_upper_bounds, upper);
  }
  void SwapObjectiveVector(Vec& objective) { std::swap(sharded_qp_.MutableQp().objective_vector, objective); }
  const PdlpBoundNorms& OriginalBoundNorms() const { return original_bound_norms_; }
  const Logger& GetLogger() const { return logger_; }

 private:
  void AddPointMetadata(const PdlpParams& params, const Vec& primal, const Vec& dual, int point_type, const Vec& last_primal_start,
                        const Vec& last_dual_start, PdlpIterationStats& stats) const;
  void LogQuadraticProgramStats(const PdlpQuadraticProgramStats& s) const;

  const int num_threads_;
  const int num_shards_;
  PdlpBoundNorms original_bound_norms_{};
  ShardedQp sharded_qp_;
  Vec col_scaling_vec_, row_scaling_vec_;
  int log_counter_ = 0;
  double time_of_last_log_ = -kInf;
  WallTimer log_clock_;
  const Logger& logger_;
  StatsCallback iteration_stats_callback_;
};

// ---------------------------------------------------------------------------
// Solver (pdhg.cc:538-764)
// ---------------------------------------------------------------------------
enum class InnerStepOutcome { kSuccessful, kForceNumericalTermination };

class Solver {
 public:
  Solver(const PdlpParams& params, Vec starting_primal, Vec starting_dual, double initial_step_size, double initial_primal_weight,
         PreprocessSolver* preprocess_solver)
      : params_(params),
        current_primal_solution_(std::move(starting_primal)),
        current_dual_solution_(std::move(starting_dual)),
        primal_average_(&preprocess_solver->ShardedWorkingQp().PrimalSharder()),
        dual_average_(&preprocess_solver->ShardedWorkingQp().DualSharder()),
        step_size_(initial_step_size),
        primal_weight_(initial_primal_weight),
        preprocess_solver_(preprocess_solver) {}

  SolverResultCpp Solve(int iteration_type, const volatile int32_t* interrupt_solve, SolveLogCpp solve_log);

 private:
  struct NextSolutionAndDelta { Vec value, delta; };
  static constexpr double kDivergentMovement = 1.0e100;

  const QuadraticProgram& WorkingQp() const { return ShardedWorkingQp().Qp(); }
  const ShardedQp& ShardedWorkingQp() const { return preprocess_solver_->ShardedWorkingQp(); }

  // pdhg.cc:1834-1880
  NextSolutionAndDelta ComputeNextPrimalSolution(double primal_step_size) const {
    const int64_t n = ShardedWorkingQp().PrimalSize();
    NextSolutionAndDelta r{Vec(n), Vec(n)};
    const QuadraticProgram& qp = WorkingQp();
    ShardedWorkingQp().PrimalSharder().ForEachShard([&](int, int64_t b, int64_t e) {
      if (!IsLinearProgram(qp)) {
        const Vec& q = *qp.objective_matrix;
        for (int64_t i = b; i < e; ++i) {
          const double scaling = primal_step_size * q[i] + 1.0;
          const double t = (current_primal_solution_[i] - primal_step_size * (qp.objective_vector[i] - current_dual_product_[i])) / scaling;
          r.value[i] = std::max(std::min(t, qp.variable_upper_bounds[i]), qp.variable_lower_bounds[i]);
        }
      } else {
        for (int64_t i = b; i < e; ++i) {
          const double t = current_primal_solution_[i] - primal_step_size * (qp.objective_vector[i] - current_dual_product_[i]);
          r.value[i] = std::max(std::min(t, qp.variable_upper_bounds[i]), qp.variable_lower_bounds[i]);
        }
      }
      for (int64_t i = b; i < e; ++i) r.delta[i] = r.value[i] - current_primal_solution_[i];
    });
    return r;
  }
  // pdhg.cc:1882-1933
  NextSolutionAndDelta ComputeNextDualSolution(double dual_step_size, double extrapolation_factor, const NextSolutionAndDelta& next_primal,
                                               const Vec* next_primal_product = nullptr) const {
    const int64_t m = ShardedWorkingQp().DualSize();
    NextSolutionAndDelta r{Vec(m), Vec(m)};
    const QuadraticProgram& qp = WorkingQp();
    Vec extrapolated;
    if (next_primal_product == nullptr) {
      extrapolated.resize(ShardedWorkingQp().PrimalSize());
      ShardedWorkingQp().PrimalSharder().ForEachShard([&](int, int64_t b, int64_t e) {
        for (int64_t i = b; i < e; ++i) extrapolated[i] = next_primal.value[i] + extrapolation_factor * next_primal.delta[i];
      });
    }
    const SparseCsc& kt = ShardedWorkingQp().TransposedConstraintMatrix();
    ShardedWorkingQp().TransposedConstraintMatrixSharder().ForEachShard([&](int, int64_t b, int64_t e) {
      for (int64_t i = b; i < e; ++i) {
        double temp;
        if (next_primal_product != nullptr) {
          temp = current_dual_solution_[i] -
                 dual_step_size * (-extrapolation_factor * (*current_primal_product_)[i] + (extrapolation_factor + 1) * (*next_primal_product)[i]);
        } else {
          double kx = 0.0;
          for (int64_t k = kt.starts[i]; k < kt.starts[i + 1]; ++k) kx += kt.value[k] * extrapolated[kt.index[k]];
          temp = current_dual_solution_[i] - dual_step_size * kx;
        }
        r.value[i] = std::max(std::min(0.0, temp + dual_step_size * qp.constraint_upper_bounds[i]), temp + dual_step_size * qp.constraint_lower_bounds[i]);
        r.delta[i] = r.value[i] - current_dual_solution_[i];
      }
    });
    return r;
  }
  // pdhg.cc:1935-1959
  std::pair<double, double> ComputeMovementTerms(const Vec& dp, const Vec& dd) const {
    return {SquaredNorm(dp, ShardedWorkingQp().PrimalSharder()), SquaredNorm(dd, ShardedWorkingQp().DualSharder())};
  }
  double ComputeMovement(const Vec& dp, const Vec& dd) const {
    const auto [p, d] = ComputeMovementTerms(dp, dd);
    return (0.5 * primal_weight_ * p) + (0.5 / primal_weight_) * d;
  }
  double ComputeNonlinearity(const Vec& delta_primal, const Vec& next_dual_product) const {
    return ShardedWorkingQp().PrimalSharder().SumOverShards([&](int, int64_t b, int64_t e) {
      double s = 0.0;
      for (int64_t i = b; i < e; ++i) s += delta_primal[i] * (next_dual_product[i] - current_dual_product_[i]);
      return -s;
    });
  }
  // pdhg.cc:1961-1974
  void SetCurrentPrimalAndDualProducts() {
    if (params_.linesearch_rule == PDLP_MALITSKY_POCK_LINESEARCH_RULE) {
      current_primal_product_ = TransposedMatrixVectorProduct(ShardedWorkingQp().TransposedConstraintMatrix(), current_primal_solution_,
                                                              ShardedWorkingQp().TransposedConstraintMatrixSharder());
    } else {
      current_primal_product_.reset();
    }
    current_dual_product_ = TransposedMatrixVectorProduct(WorkingQp().constraint_matrix, current_dual_solution_, ShardedWorkingQp().ConstraintMatrixSharder());
  }
  // pdhg.cc:1976-1996
  PdlpIterationStats CreateSimpleIterationStats(int restart_used) const {
    PdlpIterationStats s;
    std::memset(&s, 0, sizeof(s));
    const double per_rejected = params_.linesearch_rule == PDLP_MALITSKY_POCK_LINESEARCH_RULE ? 0.5 : 1.0;
    s.iteration_number = iterations_completed_;
    s.cumulative_rejected_steps = num_rejected_steps_;
    s.cumulative_kkt_matrix_passes = iterations_completed_ + per_rejected * num_rejected_steps_;
    s.cumulative_time_sec = preprocessing_time_sec_ + timer_.Get();
    s.restart_used = restart_used;
    s.step_size = step_size_;
    s.primal_weight = primal_weight_;
    return s;
  }
  // pdhg.cc:1998-2007
  double DistanceTraveledFromLastStart(const Vec& primal, const Vec& dual) const {
    return std::sqrt((0.5 * primal_weight_) * SquaredDistance(primal, last_primal_start_point_, ShardedWorkingQp().PrimalSharder()) +
                     (0.5 / primal_weight_) * SquaredDistance(dual, last_dual_start_point_, ShardedWorkingQp().DualSharder()));
  }
  // pdhg.cc:2009-2039
  LocalizedLagrangianBounds ComputeLocalizedBoundsAtCurrent() const {
    const double dist = DistanceTraveledFromLastStart(current_primal_solution_, current_dual_solution_);
    return ComputeLocalizedLagrangianBounds(ShardedWorkingQp(), current_primal_solution_, current_dual_solution_, PrimalDualNorm::kEuclideanNorm,
                                            primal_weight_, dist, current_primal_product_.has_value() ? &*current_primal_product_ : nullptr,
                                            &current_dual_product_, params_.use_diagonal_qp_trust_region_solver != 0,
                                            params_.diagonal_qp_trust_region_solver_tolerance);
  }
  LocalizedLagrangianBounds ComputeLocalizedBoundsAtAverage() const {
    const Vec ap = PrimalAverage(), ad = DualAverage();
    const double dist = DistanceTraveledFromLastStart(ap, ad);
    return ComputeLocalizedLagrangianBounds(ShardedWorkingQp(), ap, ad, PrimalDualNorm::kEuclideanNorm, primal_weight_, dist, nullptr, nullptr,
                                            params_.use_diagonal_qp_trust_region_solver != 0, params_.diagonal_qp_trust_region_solver_tolerance);
  }
  // pdhg.cc:2041-2072
  static bool AverageHasBetterPotential(const LocalizedLagrangianBounds& avg, const LocalizedLagrangianBounds& cur) {
    return BoundGap(avg) / Sq(avg.radius) < BoundGap(cur) / Sq(cur.radius);
  }
  static double NormalizedGap(const LocalizedLagrangianBounds& b) { return BoundGap(b) / b.radius; }
  bool ShouldDoAdaptiveRestartHeuristic(double candidate_normalized_gap) const {
    const double ratio = candidate_normalized_gap / normalized_gap_at_last_restart_;
    if (ratio < params_.sufficient_reduction_for_restart) return true;
    if (ratio < params_.necessary_reduction_for_restart && candidate_normalized_gap > normalized_gap_at_last_trial_) return true;
    return false;
  }
  // pdhg.cc:2074-2107
  int DetermineDistanceBasedRestartChoice() const {
    if (primal_average_.NumTerms() == 0) return PDLP_RESTART_CHOICE_NO_RESTART;
    if (distance_based_restart_info_.length_of_last_restart_period == 0) return PDLP_RESTART_CHOICE_RESTART_TO_AVERAGE;
    const int period = primal_average_.NumTerms();
    const double dist_avg = DistanceTraveledFromLastStart(primal_average_.ComputeAverage(), dual_average_.ComputeAverage());
    const double dist_last = distance_based_restart_info_.distance_moved_last_restart_period;
    if ((dist_avg / period) < params_.sufficient_reduction_for_restart * (dist_last / distance_based_restart_info_.length_of_last_restart_period)) {
      if (AverageHasBetterPotential(ComputeLocalizedBoundsAtAverage(), ComputeLocalizedBoundsAtCurrent())) return PDLP_RESTART_CHOICE_RESTART_TO_AVERAGE;
      return PDLP_RESTART_CHOICE_WEIGHTED_AVERAGE_RESET;
    }
    return PDLP_RESTART_CHOICE_NO_RESTART;
  }
  // pdhg.cc:2109-2170
  int ChooseRestartToApply(bool is_major_iteration) {
    if (!primal_average_.HasNonzeroWeight() && !dual_average_.HasNonzeroWeight()) return PDLP_RESTART_CHOICE_NO_RESTART;
    const int restart_length = primal_average_.NumTerms();
    if (restart_length >= iterations_completed_ / 2 && params_.restart_strategy == PDLP_ADAPTIVE_HEURISTIC) {
      if (AverageHasBetterPotential(ComputeLocalizedBoundsAtAverage(), ComputeLocalizedBoundsAtCurrent())) return PDLP_RESTART_CHOICE_RESTART_TO_AVERAGE;
      return PDLP_RESTART_CHOICE_WEIGHTED_AVERAGE_RESET;
    }
    if (!is_major_iteration) return PDLP_RESTART_CHOICE_NO_RESTART;
    switch (params_.restart_strategy) {
      case PDLP_NO_RESTARTS: return PDLP_RESTART_CHOICE_WEIGHTED_AVERAGE_RESET;
      case PDLP_EVERY_MAJOR_ITERATION: return PDLP_RESTART_CHOICE_RESTART_TO_AVERAGE;
      case PDLP_ADAPTIVE_HEURISTIC: {
        const LocalizedLagrangianBounds avg = ComputeLocalizedBoundsAtAverage();
        const LocalizedLagrangianBounds cur = ComputeLocalizedBoundsAtCurrent();
        double normalized_gap; int choice;
        if (AverageHasBetterPotential(avg, cur)) { normalized_gap = NormalizedGap(avg); choice = PDLP_RESTART_CHOICE_RESTART_TO_AVERAGE; }
        else { normalized_gap = NormalizedGap(cur); choice = PDLP_RESTART_CHOICE_WEIGHTED_AVERAGE_RESET; }
        if (ShouldDoAdaptiveRestartHeuristic(normalized_gap)) return choice;
        normalized_gap_at_last_trial_ = normalized_gap;
        return PDLP_RESTART_CHOICE_NO_RESTART;
      }
      case PDLP_ADAPTIVE_DISTANCE_BASED: return DetermineDistanceBasedRestartChoice();
      default: return PDLP_RESTART_CHOICE_UNSPECIFIED;
    }
  }
  // pdhg.cc:2172-2186
  Vec PrimalAverage() const { return primal_average_.HasNonzeroWeight() ? primal_average_.ComputeAverage() : current_primal_solution_; }
  Vec DualAverage() const { return dual_average_.HasNonzeroWeight() ? dual_average_.ComputeAverage() : current_dual_solution_; }
  // pdhg.cc:2188-2214
  double ComputeNewPrimalWeight() const {
    const double pd = Distance(current_primal_solution_, last_primal_start_point_, ShardedWorkingQp().PrimalSharder());
    const double dd = Distance(current_dual_solution_, last_dual_start_point_, ShardedWorkingQp().DualSharder());
    constexpr double kNonzeroTol = 1.0e-10;
    if (pd <= kNonzeroTol || pd >= 1.0 / kNonzeroTol || dd <= kNonzeroTol || dd >= 1.0 / kNonzeroTol) return primal_weight_;
    const double smoothing = params_.primal_weight_update_smoothing;
    const double unsmoothed = dd / pd;
    const double w = std::exp(smoothing * std::log(unsmoothed) + (1.0 - smoothing) * std::log(primal_weight_));
    if (params_.verbosity_level >= 4) preprocess_solver_->GetLogger().Log(Fmt("New computed primal weight is %g at iteration %d", w, iterations_completed_));
    return w;
  }
  // pdhg.cc:2216-2244 + 329-342
  SolverResultCpp PickSolutionAndConstructSolverResult(Vec primal, Vec dual, const PdlpIterationStats& stats, int reason, int output_type, SolveLogCpp log) const {
    switch (output_type) {
      case PDLP_POINT_TYPE_CURRENT_ITERATE: primal = current_primal_solution_; dual = current_dual_solution_; break;
      case PDLP_POINT_TYPE_ITERATE_DIFFERENCE: primal = current_primal_delta_; dual = current_dual_delta_; break;
      case PDLP_POINT_TYPE_AVERAGE_ITERATE:
      case PDLP_POINT_TYPE_PRESOLVER_SOLUTION: break;
      default: output_type = PDLP_POINT_TYPE_AVERAGE_ITERATE; break;
    }
    log.iteration_count = stats.iteration_number;
    log.termination_reason = reason;
    log.solution_type = output_type;
    log.solve_time_sec = stats.cumulative_time_sec;
    log.solution_stats = stats;
    log.has_solution_stats = true;
    SolverResultCpp r;
    r.primal_solution = std::move(primal);
    r.dual_solution = std::move(dual);
    r.solve_log = std::move(log);
    return r;
  }
  // pdhg.cc:2246-2296
  void ApplyRestartChoice(int restart) {
    switch (restart) {
      case PDLP_RESTART_CHOICE_UNSPECIFIED:
      case PDLP_RESTART_CHOICE_NO_RESTART: return;
      case PDLP_RESTART_CHOICE_WEIGHTED_AVERAGE_RESET:
        if (params_.verbosity_level >= 4) preprocess_solver_->GetLogger().Log(Fmt("Restarted to current on iteration %d after %d iterations", iterations_completed_, primal_average_.NumTerms()));
        break;
      case PDLP_RESTART_CHOICE_RESTART_TO_AVERAGE:
        if (params_.verbosity_level >= 4) preprocess_solver_->GetLogger().Log(Fmt("Restarted to average on iteration %d after %d iterations", iterations_completed_, primal_average_.NumTerms()));
        current_primal_solution_ = primal_average_.ComputeAverage();
        current_dual_solution_ = dual_average_.ComputeAverage();
        SetCurrentPrimalAndDualProducts();
        break;
    }
    primal_weight_ = ComputeNewPrimalWeight();
    ratio_last_two_step_sizes_ = 1;
    if (params_.restart_strategy == PDLP_ADAPTIVE_HEURISTIC) {
      const LocalizedLagrangianBounds b = ComputeLocalizedBoundsAtCurrent();
      normalized_gap_at_last_restart_ = BoundGap(b) / b.radius;
      normalized_gap_at_last_trial_ = kInf;
    } else if (params_.restart_strategy == PDLP_ADAPTIVE_DISTANCE_BASED) {
      distance_based_restart_info_ = {DistanceTraveledFromLastStart(current_primal_solution_, current_dual_solution_), primal_average_.NumTerms()};
    }
    primal_average_.Clear();
    dual_average_.Clear();
    last_primal_start_point_ = current_primal_solution_;
    last_dual_start_point_ = current_dual_solution_;
  }
  // pdhg.cc:2360-2435
  std::optional<SolverResultCpp> MajorIterationAndTerminationCheck(int iteration_type, bool force_numerical_termination,
                                                                  const volatile int32_t* interrupt_solve,
                                                                  const PdlpIterationStats& work_from_feasibility_polishing, SolveLogCpp& solve_log) {
    const int cycle = iterations_completed_ % params_.major_iteration_frequency;
    const bool is_major = cycle == 0 && iterations_completed_ > 0;
    const int restart = force_numerical_termination ? PDLP_RESTART_CHOICE_NO_RESTART : ChooseRestartToApply(is_major);
    PdlpIterationStats stats = CreateSimpleIterationStats(restart);
    const PdlpIterationStats full_work_stats = AddWorkStats(stats, work_from_feasibility_polishing);
    const auto simple = CheckSimpleTerminationCriteria(params_.termination_criteria, full_work_stats, interrupt_solve);
    const bool check_termination = cycle % params_.termination_check_frequency == 0 || simple.has_value() || force_numerical_termination;
    if (check_termination) {
      Vec primal_average = PrimalAverage();
      Vec dual_average = DualAverage();
      const auto maybe = preprocess_solver_->UpdateIterationStatsAndCheckTermination(
          params_, force_numerical_termination, current_primal_solution_, current_dual_solution_,
          primal_average_.HasNonzeroWeight() ? &primal_average : nullptr, dual_average_.HasNonzeroWeight() ? &dual_average : nullptr,
          !current_primal_delta_.empty() ? &current_primal_delta_ : nullptr, !current_dual_delta_.empty() ? &current_dual_delta_ : nullptr,
          last_primal_start_point_, last_dual_start_point_, interrupt_solve, iteration_type, full_work_stats, stats);
      if (params_.record_iteration_stats) solve_log.iteration_stats.push_back(stats);
      if (maybe.has_value()) {
        if (iteration_type == PDLP_ITERATION_TYPE_NORMAL && DoFeasibilityPolishingAfterLimitsReached(params_, maybe->reason)) {
          auto feasibility_result = TryFeasibilityPolishing(iterations_completed_ / kFeasibilityIterationFraction, interrupt_solve, solve_log);
          if (feasibility_result.has_value()) return feasibility_result;
        }
        const PdlpIterationStats terminating_full_stats = AddWorkStats(stats, work_from_feasibility_polishing);
        return PickSolutionAndConstructSolverResult(std::move(primal_average), std::move(dual_average), terminating_full_stats, maybe->reason, maybe->type,
                                                    std::move(solve_log));
      }
    } else if (params_.record_iteration_stats) {
      solve_log.iteration_stats.push_back(stats);
    }
    ApplyRestartChoice(restart);
    return std::nullopt;
  }
  // ---- feasibility polishing (pdhg.cc:2676-3015) -------------------------------------
  static constexpr int kFeasibilityIterationFraction = 8;
  PdlpIterationStats TotalWorkSoFar(const SolveLogCpp& solve_log) const {
    return AddWorkStats(CreateSimpleIterationStats(PDLP_RESTART_CHOICE_NO_RESTART), WorkFromFeasibilityPolishing(solve_log));
  }
  std::optional<SolverResultCpp> TryFeasibilityPolishing(int iteration_limit, const volatile int32_t* interrupt_solve, SolveLogCpp& solve_log);
  SolverResultCpp TryPrimalPolishing(Vec starting_primal, int iteration_limit, const volatile int32_t* interrupt_solve, SolveLogCpp& solve_log);
  SolverResultCpp TryDualPolishing(Vec starting_dual, int iteration_limit, const volatile int32_t* interrupt_solve, SolveLogCpp& solve_log);
  // pdhg.cc:2437-2442
  void ResetAverageToCurrent() {
    primal_average_.Clear(); dual_average_.Clear();
    primal_average_.Add(current_primal_solution_, 1.0);
    dual_average_.Add(current_dual_solution_, 1.0);
  }
  void LogNumericalTermination(const Vec& dp, const Vec& dd) const {
    if (params_.verbosity_level >= 2) {
      const auto [p, d] = ComputeMovementTerms(dp, dd);
      preprocess_solver_->GetLogger().Log(Fmt("Forced numerical termination at iteration %d with primal delta squared norm %g dual delta squared norm %g primal weight %g", iterations_completed_, p, d, primal_weight_));
    }
  }
  void LogInnerIterationLimitHit() const { preprocess_solver_->GetLogger().Log(Fmt("WARNING: Inner iteration limit reached at iteration %d", iterations_completed_)); }

  // pdhg.cc:2463-2556
  InnerStepOutcome TakeMalitskyPockStep() {
    InnerStepOutcome outcome = InnerStepOutcome::kSuccessful;
    const double primal_step_size = step_size_ / primal_weight_;
    NextSolutionAndDelta next_primal = ComputeNextPrimalSolution(primal_step_size);
    const double dilating = 1 + (params_.malitsky_pock_step_size_interpolation * (std::sqrt(1 + ratio_last_two_step_sizes_) - 1));
    double new_primal_step_size = primal_step_size * dilating;
    const double downscaling = params_.malitsky_pock_step_size_downscaling_factor;
    const double contraction = params_.malitsky_pock_linesearch_contraction_factor;
    const double dual_weight = primal_weight_ * primal_weight_;
    int inner_iterations = 0;
    Vec next_primal_product = TransposedMatrixVectorProduct(ShardedWorkingQp().TransposedConstraintMatrix(), next_primal.value,
                                                            ShardedWorkingQp().TransposedConstraintMatrixSharder());
    for (bool accepted = false; !accepted; ++inner_iterations) {
      if (inner_iterations >= 60) {
        LogInnerIterationLimitHit();
        ResetAverageToCurrent();
        outcome = InnerStepOutcome::kForceNumericalTermination;
        break;
      }
      const double new_ratio = new_primal_step_size / primal_step_size;
      NextSolutionAndDelta next_dual = ComputeNextDualSolution(dual_weight * new_primal_step_size, new_ratio, next_primal, &next_primal_product);
      Vec next_dual_product = TransposedMatrixVectorProduct(WorkingQp().constraint_matrix, next_dual.value, ShardedWorkingQp().ConstraintMatrixSharder());
      const double delta_dual_norm = Norm(next_dual.delta, ShardedWorkingQp().DualSharder());
      const double delta_dual_prod_norm = Distance(current_dual_product_, next_dual_product, ShardedWorkingQp().PrimalSharder());
      if (primal_weight_ * new_primal_step_size * delta_dual_prod_norm <= contraction * delta_dual_norm) {
        step_size_ = new_primal_step_size * primal_weight_;
        ratio_last_two_step_sizes_ = new_ratio;
        if (!primal_average_.HasNonzeroWeight()) primal_average_.Add(current_primal_solution_, new_primal_step_size * new_ratio);
        current_primal_solution_ = std::move(next_primal.value);
        current_dual_solution_ = std::move(next_dual.value);
        current_dual_product_ = std::move(next_dual_product);
        current_primal_product_ = std::move(next_primal_product);
        primal_average_.Add(current_primal_solution_, new_primal_step_size);
        dual_average_.Add(current_dual_solution_, new_primal_step_size);
        const double movement = ComputeMovement(next_primal.delta, next_dual.delta);
        if (movement == 0.0) {
          LogNumericalTermination(next_primal.delta, next_dual.delta);
          ResetAverageToCurrent();
          outcome = InnerStepOutcome::kForceNumericalTermination;
        } else if (movement > kDivergentMovement) {
          LogNumericalTermination(next_primal.delta, next_dual.delta);
          outcome = InnerStepOutcome::kForceNumericalTermination;
        }
        current_primal_delta_ = std::move(next_primal.delta);
        current_dual_delta_ = std::move(next_dual.delta);
        break;
      } else {
        new_primal_step_size = downscaling * new_primal_step_size;
      }
    }
    num_rejected_steps_ += inner_iterations;
    return outcome;
  }
  // pdhg.cc:2558-2642
  InnerStepOutcome TakeAdaptiveStep() {
    InnerStepOutcome outcome = InnerStepOutcome::kSuccessful;
    int inner_iterations = 0;
    for (bool accepted = false; !accepted; ++inner_iterations) {
      if (inner_iterations >= 60) {
        LogInnerIterationLimitHit();
        ResetAverageToCurrent();
        outcome = InnerStepOutcome::kForceNumericalTermination;
        break;
      }
      const double primal_step_size = step_size_ / primal_weight_;
      const double dual_step_size = step_size_ * primal_weight_;
      NextSolutionAndDelta next_primal = ComputeNextPrimalSolution(primal_step_size);
      NextSolutionAndDelta next_dual = ComputeNextDualSolution(dual_step_size, 1.0, next_primal);
      const double movement = ComputeMovement(next_primal.delta, next_dual.delta);
      if (movement == 0.0) {
        LogNumericalTermination(next_primal.delta, next_dual.delta);
        ResetAverageToCurrent();
        outcome = InnerStepOutcome::kForceNumericalTermination;
        break;
      } else if (movement > kDivergentMovement) {
        LogNumericalTermination(next_primal.delta, next_dual.delta);
        outcome = InnerStepOutcome::kForceNumericalTermination;
        break;
      }
      Vec next_dual_product = TransposedMatrixVectorProduct(WorkingQp().constraint_matrix, next_dual.value, ShardedWorkingQp().ConstraintMatrixSharder());
      const double nonlinearity = ComputeNonlinearity(next_primal.delta, next_dual_product);
      const double step_size_limit = nonlinearity > 0 ? movement / nonlinearity : kInf;
      if (step_size_ <= step_size_limit) {
        current_primal_solution_ = std::move(next_primal.value);
        current_dual_solution_ = std::move(next_dual.value);
        current_dual_product_ = std::move(next_dual_product);
        current_primal_product_.reset();
        current_primal_delta_ = std::move(next_primal.delta);
        current_dual_delta_ = std::move(next_dual.delta);
        primal_average_.Add(current_primal_solution_, step_size_);
        dual_average_.Add(current_dual_solution_, step_size_);
        accepted = true;
      }
      const double total_steps_attempted = num_rejected_steps_ + inner_iterations + iterations_completed_ + 1;
      const double first_term = std::isinf(step_size_limit)
                                    ? step_size_limit
                                    : (1 - std::pow(total_steps_attempted + 1.0, -params_.adaptive_step_size_reduction_exponent)) * step_size_limit;
      const double second_term = (1 + std::pow(total_steps_attempted + 1.0, -params_.adaptive_step_size_growth_exponent)) * step_size_;
      step_size_ = std::min(first_term, second_term);
    }
    num_rejected_steps_ += inner_iterations - 1;
    return outcome;
  }
  // pdhg.cc:2644-2675
  InnerStepOutcome TakeConstantSizeStep() {
    const double primal_step_size = step_size_ / primal_weight_;
    const double dual_step_size = step_size_ * primal_weight_;
    NextSolutionAndDelta next_primal = ComputeNextPrimalSolution(primal_step_size);
    NextSolutionAndDelta next_dual = ComputeNextDualSolution(dual_step_size, 1.0, next_primal);
    const double movement = ComputeMovement(next_primal.delta, next_dual.delta);
    if (movement == 0.0) {
      LogNumericalTermination(next_primal.delta, next_dual.delta);
      ResetAverageToCurrent();
      return InnerStepOutcome::kForceNumericalTermination;
    } else if (movement > kDivergentMovement) {
      LogNumericalTermination(next_primal.delta, next_dual.delta);
      return InnerStepOutcome::kForceNumericalTermination;
    }
    Vec next_dual_product = TransposedMatrixVectorProduct(WorkingQp().constraint_matrix, next_dual.value, ShardedWorkingQp().ConstraintMatrixSharder());
    current_primal_solution_ = std::move(next_primal.value);
    current_dual_solution_ = std::move(next_dual.value);
    current_dual_product_ = std::move(next_dual_product);
    current_primal_product_.reset();
    current_primal_delta_ = std::move(next_primal.delta);
    current_dual_delta_ = std::move(next_dual.delta);
    primal_average_.Add(current_primal_solution_, step_size_);
    dual_average_.Add(current_dual_solution_, step_size_);
    return InnerStepOutcome::kSuccessful;
  }

  const PdlpParams params_;
  Vec current_primal_solution_, current_dual_solution_, current_primal_delta_, current_dual_delta_;
  WeightedAverage primal_average_, dual_average_;
  double step_size_, primal_weight_;
  PreprocessSolver* preprocess_solver_;
  double ratio_last_two_step_sizes_ = 1;
  double normalized_gap_at_last_trial_ = kInf, normalized_gap_at_last_restart_ = kInf;
  double preprocessing_time_sec_ = 0;
  WallTimer timer_;
  int iterations_completed_ = 0, num_rejected_steps_ = 0;
  std::optional<Vec> current_primal_product_;
  Vec current_dual_product_;
  Vec last_primal_start_point_, last_dual_start_point_;
  struct { double distance_moved_last_restart_period = kInf; int length_of_last_restart_period = 1; } distance_based_restart_info_;
};

// pdhg.cc:3017-3092
SolverResultCpp Solver::Solve(int iteration_type, const volatile int32_t* interrupt_solve, SolveLogCpp solve_log) {
  preprocessing_time_sec_ = solve_log.preprocessing_time_sec;
  timer_.Start();
  last_primal_start_point_ = current_primal_solution_;
  last_dual_start_point_ = current_dual_solution_;
  ratio_last_two_step_sizes_ = 1;
  SetCurrentPrimalAndDualProducts();
  bool force_numerical_termination = false;
  int next_feasibility_polishing_iteration = 100;
  num_rejected_steps_ = 0;
  PdlpIterationStats work_from_feasibility_polishing = WorkFromFeasibilityPolishing(solve_log);
  for (iterations_completed_ = 0;; ++iterations_completed_) {
    auto maybe = MajorIterationAndTerminationCheck(iteration_type, force_numerical_termination, interrupt_solve, work_from_feasibility_polishing, solve_log);
    if (maybe.has_value()) return std::move(*maybe);
    if (params_.use_feasibility_polishing && iteration_type == PDLP_ITERATION_TYPE_NORMAL && iterations_completed_ >= next_feasibility_polishing_iteration) {
      auto feasibility_result = TryFeasibilityPolishing(iterations_completed_ / kFeasibilityIterationFraction, interrupt_solve, solve_log);
      if (feasibility_result.has_value()) return std::move(*feasibility_result);
      next_feasibility_polishing_iteration *= 2;
      work_from_feasibility_polishing = WorkFromFeasibilityPolishing(solve_log);
    }
    InnerStepOutcome outcome;
    switch (params_.linesearch_rule) {
      case PDLP_MALITSKY_POCK_LINESEARCH_RULE: outcome = TakeMalitskyPockStep(); break;
      case PDLP_CONSTANT_STEP_SIZE_RULE: outcome = TakeConstantSizeStep(); break;
      default: outcome = TakeAdaptiveStep(); break;
    }
    if (outcome == InnerStepOutcome::kForceNumericalTermination) force_numerical_termination = true;
  }
}

// pdhg.cc:2702-2865
std::optional<SolverResultCpp> Solver::TryFeasibilityPolishing(int iteration_limit, const volatile int32_t* interrupt_solve, SolveLogCpp& solve_log) {
  const Logger& logger = preprocess_solver_->GetLogger();
  const DetailedCriteria optimality_criteria = EffectiveOptimalityCriteria(params_.termination_criteria);
  Vec average_primal = PrimalAverage();
  Vec average_dual = DualAverage();
  PdlpConvergenceInformation first_convergence_info;
  preprocess_solver_->ComputeConvergenceAndInfeasibilityFromWorkingSolution(params_, average_primal, average_dual, PDLP_POINT_TYPE_AVERAGE_ITERATE,
                                                                            &first_convergence_info, nullptr);
  // The objective gap is usually increased by polishing: do not start while it is still too large.
  if (!ObjectiveGapMet(optimality_criteria, first_convergence_info)) {
    const auto simple = CheckSimpleTerminationCriteria(params_.termination_criteria, TotalWorkSoFar(solve_log), interrupt_solve);
    if (!(simple.has_value() && DoFeasibilityPolishingAfterLimitsReached(params_, simple->reason))) {
      if (params_.verbosity_level >= 2) logger.Log("Skipping feasibility polishing because the objective gap is too large.");
      return std::nullopt;
    }
  }
  if (params_.verbosity_level >= 2) logger.Log("Starting primal feasibility polishing");
  SolverResultCpp primal_result = TryPrimalPolishing(std::move(average_primal), iteration_limit, interrupt_solve, solve_log);
  if (params_.verbosity_level >= 2) logger.Log(Fmt("Primal feasibility polishing termination reason: %d", primal_result.solve_log.termination_reason));
  if (TerminationReasonIsWorkLimit(primal_result.solve_log.termination_reason)) {
    const auto simple = CheckSimpleTerminationCriteria(params_.termination_criteria, TotalWorkSoFar(solve_log), interrupt_solve);
    if (!(simple.has_value() && DoFeasibilityPolishingAfterLimitsReached(params_, simple->reason))) return std::nullopt;
  } else if (primal_result.solve_log.termination_reason != PDLP_TERMINATION_REASON_OPTIMAL) {
    logger.Log(Fmt("WARNING: Primal feasibility polishing terminated with error %d", primal_result.solve_log.termination_reason));
    return std::nullopt;
  }
  if (params_.verbosity_level >= 2) logger.Log("Starting dual feasibility polishing");
  SolverResultCpp dual_result = TryDualPolishing(std::move(average_dual), iteration_limit, interrupt_solve, solve_log);
  if (params_.verbosity_level >= 2) logger.Log(Fmt("Dual feasibility polishing termination reason: %d", dual_result.solve_log.termination_reason));
  PdlpIterationStats full_stats = TotalWorkSoFar(solve_log);
  const auto simple = CheckSimpleTerminationCriteria(params_.termination_criteria, full_stats, interrupt_solve);
  auto add_polished_convergence = [&] {
    preprocess_solver_->ComputeConvergenceAndInfeasibilityFromWorkingSolution(
        params_, primal_result.primal_solution, dual_result.dual_solution, PDLP_POINT_TYPE_FEASIBILITY_POLISHING_SOLUTION,
        &full_stats.convergence_information[full_stats.num_convergence_information], nullptr);
    full_stats.num_convergence_information += 1;
  };
  if (TerminationReasonIsWorkLimit(dual_result.solve_log.termination_reason)) {
    if (simple.has_value() && DoFeasibilityPolishingAfterLimitsReached(params_, simple->reason)) {
      add_polished_convergence();
      return ConstructSolverResult(std::move(primal_result.primal_solution), std::move(dual_result.dual_solution), full_stats, simple->reason,
                                   PDLP_POINT_TYPE_FEASIBILITY_POLISHING_SOLUTION, solve_log);
    }
    return std::nullopt;
  } else if (dual_result.solve_log.termination_reason != PDLP_TERMINATION_REASON_OPTIMAL) {
    logger.Log(Fmt("WARNING: Dual feasibility polishing terminated with error %d", dual_result.solve_log.termination_reason));
    return std::nullopt;
  }
  add_polished_convergence();
  if (params_.verbosity_level >= 2) {
    logger.Log("solution stats for polished solution:");
    LogIterationStatsHeader(params_.verbosity_level, logger);
    LogIterationStats(params_.verbosity_level, full_stats, params_.termination_criteria, preprocess_solver_->OriginalBoundNorms(),
                      PDLP_POINT_TYPE_FEASIBILITY_POLISHING_SOLUTION, logger);
  }
  const auto earned = CheckIterateTerminationCriteria(params_.termination_criteria, full_stats, preprocess_solver_->OriginalBoundNorms(),
                                                      /*force_numerical_termination=*/false);
  if (earned.has_value() || (simple.has_value() && DoFeasibilityPolishingAfterLimitsReached(params_, simple->reason))) {
    return ConstructSolverResult(std::move(primal_result.primal_solution), std::move(dual_result.dual_solution), full_stats,
                                 earned.has_value() ? earned->reason : simple->reason, PDLP_POINT_TYPE_FEASIBILITY_POLISHING_SOLUTION, solve_log);
  }
  return std::nullopt;
}

// pdhg.cc:2888-2938
SolverResultCpp Solver::TryPrimalPolishing(Vec starting_primal, int iteration_limit, const volatile int32_t* interrupt_solve, SolveLogCpp& solve_log) {
  PdlpParams phase_params = params_;
  phase_params.termination_criteria = ReduceWorkLimitsByPreviousWork(params_.termination_criteria, iteration_limit, TotalWorkSoFar(solve_log),
                                                                     params_.apply_feasibility_polishing_after_limits_reached != 0);
  if (params_.apply_feasibility_polishing_if_solver_is_interrupted) interrupt_solve = nullptr;
  Vec objective(ShardedWorkingQp().PrimalSize(), 0.0);  // holds the original objective after the swap
  preprocess_solver_->SwapObjectiveVector(objective);
  DetailedCriteria criteria = EffectiveOptimalityCriteria(params_.termination_criteria);
  criteria.dual_abs = criteria.dual_rel = kInf;
  criteria.gap_abs = criteria.gap_rel = kInf;
  SetDetailedCriteria(phase_params.termination_criteria, criteria);
  Vec starting_dual(ShardedWorkingQp().DualSize(), 0.0);
  Solver primal_solver(phase_params, std::move(starting_primal), std::move(starting_dual), step_size_, primal_weight_, preprocess_solver_);
  SolveLogCpp phase_log;
  timer_.Stop();  // the time inside the phase is recorded by its own timer
  SolverResultCpp result = primal_solver.Solve(PDLP_ITERATION_TYPE_PRIMAL_FEASIBILITY, interrupt_solve, phase_log);
  timer_.Resume();
  preprocess_solver_->SwapObjectiveVector(objective);
  solve_log.feasibility_polishing_details.push_back(
      BuildFeasibilityPolishingDetails(PDLP_POLISHING_PHASE_TYPE_PRIMAL_FEASIBILITY, iterations_completed_, phase_params, result.solve_log));
  return result;
}

// pdhg.cc:2951-3015
SolverResultCpp Solver::TryDualPolishing(Vec starting_dual, int iteration_limit, const volatile int32_t* interrupt_solve, SolveLogCpp& solve_log) {
  PdlpParams phase_params = params_;
  phase_params.termination_criteria = ReduceWorkLimitsByPreviousWork(params_.termination_criteria, iteration_limit, TotalWorkSoFar(solve_log),
                                                                     params_.apply_feasibility_polishing_after_limits_reached != 0);
  if (params_.apply_feasibility_polishing_if_solver_is_interrupted) interrupt_solve = nullptr;
  // homogeneous bounds now, the original ones after the swap
  Vec constraint_lower = MapFiniteValuesToZero(WorkingQp().constraint_lower_bounds);
  Vec constraint_upper = MapFiniteValuesToZero(WorkingQp().constraint_upper_bounds);
  Vec variable_lower = MapFiniteValuesToZero(WorkingQp().variable_lower_bounds);
  Vec variable_upper = MapFiniteValuesToZero(WorkingQp().variable_upper_bounds);
  preprocess_solver_->SwapConstraintBounds(constraint_lower, constraint_upper);
  preprocess_solver_->SwapVariableBounds(variable_lower, variable_upper);
  DetailedCriteria criteria = EffectiveOptimalityCriteria(params_.termination_criteria);
  criteria.primal_abs = criteria.primal_rel = kInf;
  criteria.gap_abs = criteria.gap_rel = kInf;
  SetDetailedCriteria(phase_params.termination_criteria, criteria);
  Vec starting_primal(ShardedWorkingQp().PrimalSize(), 0.0);
  Solver dual_solver(phase_params, std::move(starting_primal), std::move(starting_dual), step_size_, primal_weight_, preprocess_solver_);
  SolveLogCpp phase_log;
  timer_.Stop();
  SolverResultCpp result = dual_solver.Solve(PDLP_ITERATION_TYPE_DUAL_FEASIBILITY, interrupt_solve, phase_log);
  timer_.Resume();
  preprocess_solver_->SwapConstraintBounds(constraint_lower, constraint_upper);
  preprocess_solver_->SwapVariableBounds(variable_lower, variable_upper);
  solve_log.feasibility_polishing_details.push_back(
      BuildFeasibilityPolishingDetails(PDLP_POLISHING_PHASE_TYPE_DUAL_FEASIBILITY, iterations_completed_, phase_params, result.solve_log));
  return result;
}

void PreprocessSolver::LogQuadraticProgramStats(const PdlpQuadraticProgramStats& s) const {
  logger_.Log(Fmt("There are %lld variables, %lld constraints, and %lld constraint matrix nonzeros.", (long long)s.num_variables, (long long)s.num_constraints, (long long)s.constraint_matrix_num_nonzeros));
  if (s.constraint_matrix_num_nonzeros > 0) {
    logger_.Log(Fmt("Absolute values of nonzero constraint matrix elements: largest=%f, smallest=%f, avg=%f", s.constraint_matrix_abs_max, s.constraint_matrix_abs_min, s.constraint_matrix_abs_avg));
    logger_.Log(Fmt("Constraint matrix, infinity norm: max(row & col)=%f, min_col=%f, min_row=%f", s.constraint_matrix_abs_max, s.constraint_matrix_col_min_l_inf_norm, s.constraint_matrix_row_min_l_inf_norm));
    logger_.Log(Fmt("Constraint bounds statistics (max absolute value per row): largest=%f, smallest=%f, avg=%f, l2_norm=%f", s.combined_bounds_max, s.combined_bounds_min, s.combined_bounds_avg, s.combined_bounds_l2_norm));
  }
  if (!IsLinearProgram(sharded_qp_.Qp())) {
    logger_.Log(Fmt("There are %lld nonzero diagonal coefficients in the objective matrix.", (long long)s.objective_matrix_num_nonzeros));
    logger_.Log(Fmt("Absolute values of nonzero objective matrix elements: largest=%f, smallest=%f, avg=%f", s.objective_matrix_abs_max, s.objective_matrix_abs_min, s.objective_matrix_abs_avg));
  }
  logger_.Log(Fmt("Absolute values of objective vector elements: largest=%f, smallest=%f, avg=%f, l2_norm=%f", s.objective_vector_abs_max, s.objective_vector_abs_min, s.objective_vector_abs_avg, s.objective_vector_l2_norm));
  logger_.Log(Fmt("Gaps between variable upper and lower bounds: #finite=%lld of %lld, largest=%f, smallest=%f, avg=%f", (long long)s.variable_bound_gaps_num_finite, (long long)s.num_variables, s.variable_bound_gaps_max, s.variable_bound_gaps_min, s.variable_bound_gaps_avg));
}

// pdhg.cc:1039-1221
SolverResultCpp PreprocessSolver::PreprocessAndSolve(const PdlpParams& params, std::optional<InitialSolution> initial_solution,
                                                     const volatile int32_t* interrupt_solve, StatsCallback callback) {
  WallTimer timer;
  timer.Start();
  SolveLogCpp solve_log;
  if (params.verbosity_level >= 1) logger_.Log("Solving with PDLP parameters: (PdlpParams POD)");
  if (sharded_qp_.Qp().problem_name.has_value()) solve_log.instance_name = *sharded_qp_.Qp().problem_name;
  solve_log.params = params;
  sharded_qp_.ReplaceLargeConstraintBoundsWithInfinity(params.infinite_constraint_bound_threshold);
  if (!HasValidBounds(sharded_qp_)) {
    return ErrorSolverResult(PDLP_TERMINATION_REASON_INVALID_PROBLEM,
                             "The input problem has invalid bounds (after replacing large constraint bounds with infinity): some variable or "
                             "constraint has lower_bound > upper_bound, lower_bound == inf, or upper_bound == -inf.", logger_);
  }
  if (sharded_qp_.Qp().objective_matrix.has_value()) {
    const Vec& q = *sharded_qp_.Qp().objective_matrix;
    const bool convex = sharded_qp_.PrimalSharder().TrueForAllShards([&](int, int64_t b, int64_t e) { for (int64_t i = b; i < e; ++i) if (!(q[i] >= 0.0)) return false; return true; });
    if (!convex) return ErrorSolverResult(PDLP_TERMINATION_REASON_INVALID_PROBLEM, "The objective is not convex (i.e., the objective matrix contains negative or NAN entries).", logger_);
  }
  solve_log.original_stats = ComputeStats(sharded_qp_);
  solve_log.has_original_stats = true;
  if (auto r = CheckProblemStats(solve_log.original_stats, sharded_qp_.Qp().objective_offset, params.presolve_use_glop != 0, logger_); r.has_value()) return std::move(*r);
  if (initial_solution.has_value()) {
    if (auto r = CheckInitialSolution(sharded_qp_, *initial_solution, logger_); r.has_value()) return std::move(*r);
  }
  original_bound_norms_ = BoundNormsFromProblemStats(solve_log.original_stats);
  if (params.verbosity_level >= 1) { logger_.Log("Problem stats before rescaling:"); LogQuadraticProgramStats(solve_log.original_stats); }
  iteration_stats_callback_ = std::move(callback);

  Vec starting_primal, starting_dual;
  if (initial_solution.has_value()) { starting_primal = std::move(initial_solution->primal); starting_dual = std::move(initial_solution->dual); }
  else { SetZero(sharded_qp_.PrimalSharder(), starting_primal); SetZero(sharded_qp_.DualSharder(), starting_dual); }
  ProjectToPrimalVariableBounds(sharded_qp_, starting_primal);
  ProjectToDualVariableBounds(sharded_qp_, starting_dual);

  // ComputeAndApplyRescaling, pdhg.cc:1325-1339
  ScalingVectors scaling = ApplyRescaling(params.l_inf_ruiz_iterations, params.l2_norm_rescaling != 0, sharded_qp_);
  row_scaling_vec_ = std::move(scaling.row_scaling_vec);
  col_scaling_vec_ = std::move(scaling.col_scaling_vec);
  CoefficientWiseQuotientInPlace(col_scaling_vec_, sharded_qp_.PrimalSharder(), starting_primal);
  CoefficientWiseQuotientInPlace(row_scaling_vec_, sharded_qp_.DualSharder(), starting_dual);

  solve_log.preprocessed_stats = ComputeStats(sharded_qp_);
  solve_log.has_preprocessed_stats = true;
  if (params.verbosity_level >= 1) { logger_.Log("Problem stats after rescaling:"); LogQuadraticProgramStats(solve_log.preprocessed_stats); }

  double step_size = 0.0;
  if (params.linesearch_rule == PDLP_CONSTANT_STEP_SIZE_RULE) {
    std::mt19937 random(1);
    const auto lip = EstimateMaximumSingularValueOfConstraintMatrix(sharded_qp_, std::nullopt, std::nullopt, 0.2, 0.0005, random);
    const double upper = lip.singular_value / (1.0 - lip.estimated_relative_error);
    step_size = upper > 0.0 ? 1.0 / upper : 1.0;
  } else {
    step_size = 1.0 / std::max(1.0e-20, solve_log.preprocessed_stats.constraint_matrix_abs_max);
  }
  step_size *= params.initial_step_size_scaling;

  // InitialPrimalWeight, pdhg.cc:1401-1419
  double primal_weight = 1.0;
  if (params.has_initial_primal_weight) primal_weight = params.initial_primal_weight;
  else if (solve_log.preprocessed_stats.objective_vector_l2_norm > 0.0 && solve_log.preprocessed_stats.combined_bounds_l2_norm > 0.0)
    primal_weight = solve_log.preprocessed_stats.objective_vector_l2_norm / solve_log.preprocessed_stats.combined_bounds_l2_norm;

  Solver solver(params, starting_primal, starting_dual, step_size, primal_weight, this);
  solve_log.preprocessing_time_sec = timer.Get();
  SolverResultCpp result = solver.Solve(PDLP_ITERATION_TYPE_NORMAL, interrupt_solve, std::move(solve_log));
  return ConstructOriginalSolverResult(params, std::move(result));
}

// pdhg.cc:1547-1565
void PreprocessSolver::AddPointMetadata(const PdlpParams& params, const Vec& primal, const Vec& dual, int point_type, const Vec& last_primal_start,
                                        const Vec& last_dual_start, PdlpIterationStats& stats) const {
  PdlpPointMetadata md;
  std::memset(&md, 0, sizeof(md));
  md.point_type = point_type;
  const int ns = std::min<int>(params.num_random_projection_seeds, PDLP_MAX_RANDOM_PROJECTION_SEEDS);
  md.num_random_projections = ns;
  for (int k = 0; k < ns; ++k) {
    std::mt19937 seed_gen(static_cast<uint32_t>(params.random_projection_seeds[k]));
    md.random_primal_projections[k] = RandomProjection(primal, sharded_qp_.PrimalSharder(), seed_gen);
    md.random_dual_projections[k] = RandomProjection(dual, sharded_qp_.DualSharder(), seed_gen);
  }
  if (point_type != PDLP_POINT_TYPE_ITERATE_DIFFERENCE) SetActiveSetInformation(sharded_qp_, primal, dual, last_primal_start, last_dual_start, md);
  stats.point_metadata[stats.num_point_metadata++] = md;
}

// pdhg.cc:1567-1653
std::optional<TerminationReasonAndPointType> PreprocessSolver::UpdateIterationStatsAndCheckTermination(
    const PdlpParams& params, bool force_numerical_termination, const Vec& primal_current, const Vec& dual_current, const Vec* primal_average,
    const Vec* dual_average, const Vec* primal_delta, const Vec* dual_delta, const Vec& last_primal_start, const Vec& last_dual_start,
    const volatile int32_t* interrupt_solve, int iteration_type, const PdlpIterationStats& full_stats, PdlpIterationStats& stats) {
  ComputeConvergenceAndInfeasibilityFromWorkingSolution(params, primal_current, dual_current, PDLP_POINT_TYPE_CURRENT_ITERATE,
                                                        &stats.convergence_information[stats.num_convergence_information],
                                                        &stats.infeasibility_information[stats.num_infeasibility_information]);
  stats.num_convergence_information++; stats.num_infeasibility_information++;
  AddPointMetadata(params, primal_current, dual_current, PDLP_POINT_TYPE_CURRENT_ITERATE, last_primal_start, last_dual_start, stats);
  if (primal_average != nullptr && dual_average != nullptr) {
    ComputeConvergenceAndInfeasibilityFromWorkingSolution(params, *primal_average, *dual_average, PDLP_POINT_TYPE_AVERAGE_ITERATE,
                                                          &stats.convergence_information[stats.num_convergence_information],
                                                          &stats.infeasibility_information[stats.num_infeasibility_information]);
    stats.num_convergence_information++; stats.num_infeasibility_information++;
    AddPointMetadata(params, *primal_average, *dual_average, PDLP_POINT_TYPE_AVERAGE_ITERATE, last_primal_start, last_dual_start, stats);
  }
  if (primal_delta != nullptr && dual_delta != nullptr) {
    ComputeConvergenceAndInfeasibilityFromWorkingSolution(params, *primal_delta, *dual_delta, PDLP_POINT_TYPE_ITERATE_DIFFERENCE, nullptr,
                                                          &stats.infeasibility_information[stats.num_infeasibility_information]);
    stats.num_infeasibility_information++;
    AddPointMetadata(params, *primal_delta, *dual_delta, PDLP_POINT_TYPE_ITERATE_DIFFERENCE, last_primal_start, last_dual_start, stats);
  }
  constexpr int kLogEvery = 15;
  const double now = log_clock_.Get();
  if (params.verbosity_level >= 2 && (params.log_interval_seconds == 0.0 || now - time_of_last_log_ >= params.log_interval_seconds)) {
    if (log_counter_ == 0) LogIterationStatsHeader(params.verbosity_level, logger_);
    LogIterationStats(params.verbosity_level, stats, params.termination_criteria, original_bound_norms_, PDLP_POINT_TYPE_AVERAGE_ITERATE, logger_);
    if (params.verbosity_level >= 4 && GetConvergenceInformation(stats, PDLP_POINT_TYPE_AVERAGE_ITERATE) != nullptr)
      LogIterationStats(params.verbosity_level, stats, params.termination_criteria, original_bound_norms_, PDLP_POINT_TYPE_CURRENT_ITERATE, logger_);
    time_of_last_log_ = now;
    if (++log_counter_ >= kLogEvery) log_counter_ = 0;
  }
  if (iteration_stats_callback_) {
    PdlpIterationCallbackInfo info{iteration_type, &params.termination_criteria, &stats, original_bound_norms_};
    iteration_stats_callback_(info);
  }
  if (const auto t = CheckIterateTerminationCriteria(params.termination_criteria, stats, original_bound_norms_, force_numerical_termination); t.has_value()) return t;
  return CheckSimpleTerminationCriteria(params.termination_criteria, full_stats, interrupt_solve);
}

// pdhg.cc:1655-1724 (no-presolve branch)
void PreprocessSolver::ComputeConvergenceAndInfeasibilityFromWorkingSolution(const PdlpParams& params, const Vec& working_primal, const Vec& working_dual,
                                                                             int candidate_type, PdlpConvergenceInformation* conv,
                                                                             PdlpInfeasibilityInformation* infeas) const {
  const DetailedCriteria oc = EffectiveOptimalityCriteria(params.termination_criteria);
  const double primal_ratio = EpsilonRatio(oc.primal_abs, oc.primal_rel);
  const double dual_ratio = EpsilonRatio(oc.dual_abs, oc.dual_rel);
  const bool har = params.handle_some_primal_gradients_on_finite_bounds_as_residuals != 0;
  if (conv != nullptr)
    *conv = ComputeConvergenceInformation(har, sharded_qp_, col_scaling_vec_, row_scaling_vec_, working_primal, working_dual, primal_ratio, dual_ratio, candidate_type);
  if (infeas != nullptr) {
    Vec primal_copy = working_primal;
    ProjectToPrimalVariableBounds(sharded_qp_, primal_copy, /*use_feasibility_bounds=*/true);
    if (candidate_type == PDLP_POINT_TYPE_ITERATE_DIFFERENCE) {
      Vec dual_copy = working_dual;
      ProjectToDualVariableBounds(sharded_qp_, dual_copy);
      *infeas = ComputeInfeasibilityInformation(har, sharded_qp_, col_scaling_vec_, row_scaling_vec_, primal_copy, dual_copy, working_primal, candidate_type);
    } else {
      *infeas = ComputeInfeasibilityInformation(har, sharded_qp_, col_scaling_vec_, row_scaling_vec_, primal_copy, working_dual, working_primal, candidate_type);
    }
  }
}

// pdhg.cc:1728-1818 (no-presolve branch)
SolverResultCpp PreprocessSolver::ConstructOriginalSolverResult(const PdlpParams& params, SolverResultCpp result) const {
  const bool use_zero_primal_objective = result.solve_log.termination_reason == PDLP_TERMINATION_REASON_PRIMAL_INFEASIBLE;
  if (result.solve_log.termination_reason == PDLP_TERMINATION_REASON_DUAL_INFEASIBLE) ProjectToPrimalVariableBounds(sharded_qp_, result.primal_solution, true);
  if (result.solve_log.termination_reason == PDLP_TERMINATION_REASON_PRIMAL_INFEASIBLE) ProjectToDualVariableBounds(sharded_qp_, result.dual_solution);
  result.reduced_costs = ReducedCosts(sharded_qp_, result.primal_solution, result.dual_solution, use_zero_primal_objective);
  CoefficientWiseProductInPlace(col_scaling_vec_, sharded_qp_.PrimalSharder(), result.primal_solution);
  CoefficientWiseProductInPlace(row_scaling_vec_, sharded_qp_.DualSharder(), result.dual_solution);
  CoefficientWiseQuotientInPlace(col_scaling_vec_, sharded_qp_.PrimalSharder(), result.reduced_costs);
  if (iteration_stats_callback_) {
    const int termination_type = result.solve_log.solution_type == PDLP_POINT_TYPE_FEASIBILITY_POLISHING_SOLUTION
                                     ? PDLP_ITERATION_TYPE_FEASIBILITY_POLISHING_TERMINATION
                                     : (result.solve_log.solution_type == PDLP_POINT_TYPE_PRESOLVER_SOLUTION ? PDLP_ITERATION_TYPE_PRESOLVE_TERMINATION
                                                                                                            : PDLP_ITERATION_TYPE_NORMAL_TERMINATION);
    PdlpIterationCallbackInfo info{termination_type, &params.termination_criteria, &result.solve_log.solution_stats, original_bound_norms_};
    iteration_stats_callback_(info);
  }
  if (params.verbosity_level >= 1) {
    logger_.Log(Fmt("Termination reason: %d", result.solve_log.termination_reason));
    logger_.Log(Fmt("Solution point type: %d", result.solve_log.solution_type));
    logger_.Log("Final solution stats:");
    LogIterationStatsHeader(params.verbosity_level, logger_);
    LogIterationStats(params.verbosity_level, result.solve_log.solution_stats, params.termination_criteria, original_bound_norms_, result.solve_log.solution_type, logger_);
    const PdlpConvergenceInformation* ci = GetConvergenceInformation(result.solve_log.solution_stats, result.solve_log.solution_type);
    if (ci != nullptr && std::isfinite(ci->corrected_dual_objective)) logger_.Log(Fmt("Dual objective after infeasibility correction: %g", ci->corrected_dual_objective));
  }
  return result;
}

// quadratic_program.cc:38-97 on the lengths carried by the view.
std::string ValidateDimensions(const PdlpProblemView& v) {
  auto sz = [](int64_t given, int64_t dflt) { return given < 0 ? dflt : given; };
  const int64_t n = v.num_variables, m = v.num_constraints;
  const int64_t var_lb = sz(v.variable_lower_bounds_size, n), var_ub = sz(v.variable_upper_bounds_size, n), obj = sz(v.objective_vector_size, n);
  const int64_t con_lb = sz(v.constraint_lower_bounds_size, m), con_ub = sz(v.constraint_upper_bounds_size, m);
  if (var_lb != var_ub) return Fmt("Inconsistent dimensions: variable lower bound vector has size %lld while variable upper bound vector has size %lld", (long long)var_lb, (long long)var_ub);
  if (var_lb != obj) return Fmt("Inconsistent dimensions: variable lower bound vector has size %lld while objective vector has size %lld", (long long)var_lb, (long long)obj);
  if (var_lb != n) return Fmt("Inconsistent dimensions: variable lower bound vector has size %lld while constraint matrix has %lld columns", (long long)var_lb, (long long)n);
  if (v.objective_matrix_diagonal != nullptr && var_lb != sz(v.objective_matrix_size, n))
    return Fmt("Inconsistent dimensions: variable lower bound vector has size %lld while objective matrix has %lld rows", (long long)var_lb, (long long)sz(v.objective_matrix_size, n));
  if (con_lb != con_ub) return Fmt("Inconsistent dimensions: constraint lower bound vector has size %lld while constraint upper bound vector has size %lld", (long long)con_lb, (long long)con_ub);
  if (con_lb != m) return Fmt("Inconsistent dimensions: constraint lower bound vector has size %lld while constraint matrix has %lld rows ", (long long)con_lb, (long long)m);
  return "";
}

QuadraticProgram QpFromView(const PdlpProblemView& v) {
  QuadraticProgram qp;
  const int64_t n = v.num_variables, m = v.num_constraints, nnz = v.num_nonzeros;
  qp.objective_vector.assign(v.objective_vector, v.objective_vector + n);
  if (v.objective_matrix_diagonal != nullptr) qp.objective_matrix = Vec(v.objective_matrix_diagonal, v.objective_matrix_diagonal + n);
  qp.constraint_matrix.rows = m; qp.constraint_matrix.cols = n;
  qp.constraint_matrix.starts.assign(v.col_starts, v.col_starts + n + 1);
  qp.constraint_matrix.index.assign(v.row_indices, v.row_indices + nnz);
  qp.constraint_matrix.value.assign(v.values, v.values + nnz);
  qp.constraint_lower_bounds.assign(v.constraint_lower_bounds, v.constraint_lower_bounds + m);
  qp.constraint_upper_bounds.assign(v.constraint_upper_bounds, v.constraint_upper_bounds + m);
  qp.variable_lower_bounds.assign(v.variable_lower_bounds, v.variable_lower_bounds + n);
  qp.variable_upper_bounds.assign(v.variable_upper_bounds, v.variable_upper_bounds + n);
  qp.objective_offset = v.objective_offset;
  qp.objective_scaling_factor = v.objective_scaling_factor;
  if (v.problem_name != nullptr) qp.problem_name = std::string(v.problem_name);
  return qp;
}

// pdhg.cc:3107-3152
SolverResultCpp PrimalDualHybridGradient(const PdlpProblemView& view, const PdlpParams& params, std::optional<InitialSolution> initial_solution,
                                         const volatile int32_t* interrupt_solve, const Logger& logger, StatsCallback callback) {
  const std::string perr = ValidateParams(params);
  if (!perr.empty()) return ErrorSolverResult(PDLP_TERMINATION_REASON_INVALID_PARAMETER, "INVALID_ARGUMENT: " + perr, logger);
  const std::string derr = ValidateDimensions(view);
  if (!derr.empty()) return ErrorSolverResult(PDLP_TERMINATION_REASON_INVALID_PROBLEM, "INVALID_ARGUMENT: " + derr, logger);
  if (view.objective_scaling_factor == 0) return ErrorSolverResult(PDLP_TERMINATION_REASON_INVALID_PROBLEM, "The objective scaling factor cannot be zero.", logger);
  if (params.use_feasibility_polishing && view.objective_matrix_diagonal != nullptr)
    return ErrorSolverResult(PDLP_TERMINATION_REASON_INVALID_PARAMETER, "use_feasibility_polishing is only implemented for linear programs.", logger);
  if (params.presolve_use_glop)
    return ErrorSolverResult(PDLP_TERMINATION_REASON_INVALID_PARAMETER, "presolve_options.use_glop is a host-side feature that this build does not provide.", logger);
  if (params.num_random_projection_seeds > PDLP_MAX_RANDOM_PROJECTION_SEEDS)
    return ErrorSolverResult(PDLP_TERMINATION_REASON_INVALID_PARAMETER, "at most 8 random_projection_seeds are supported.", logger);
  PreprocessSolver solver(QpFromView(view), params, &logger);
  return solver.PreprocessAndSolve(params, std::move(initial_solution), interrupt_solve, std::move(callback));
}

char* DupString(const std::string& s) {
  char* p = static_cast<char*>(std::malloc(s.size() + 1));
  std::memcpy(p, s.c_str(), s.size() + 1);
  return p;
}
double* DupVec(const Vec& v) {
  if (v.empty()) return nullptr;
  double* p = static_cast<double*>(std::malloc(v.size() * sizeof(double)));
  std::memcpy(p, v.data(), v.size() * sizeof(double));
  return p;
}
void FillResult(SolverResultCpp&& r, PdlpResult* out) {
  std::memset(out, 0, sizeof(*out));
  out->primal_size = static_cast<int64_t>(r.primal_solution.size());
  out->dual_size = static_cast<int64_t>(r.dual_solution.size());
  out->primal_solution = DupVec(r.primal_solution);
  out->dual_solution = DupVec(r.dual_solution);
  out->reduced_costs = DupVec(r.reduced_costs);
  const SolveLogCpp& l = r.solve_log;
  out->instance_name = l.instance_name ? DupString(*l.instance_name) : nullptr;
  out->termination_reason = l.termination_reason;
  out->termination_string = l.termination_string ? DupString(*l.termination_string) : nullptr;
  out->iteration_count = l.iteration_count;
  out->solve_time_sec = l.solve_time_sec;
  out->preprocessing_time_sec = l.preprocessing_time_sec;
  out->solution_type = l.solution_type;
  out->has_solution_stats = l.has_solution_stats;
  out->solution_stats = l.solution_stats;
  out->has_original_problem_stats = l.has_original_stats;
  out->has_preprocessed_problem_stats = l.has_preprocessed_stats;
  out->original_problem_stats = l.original_stats;
  out->preprocessed_problem_stats = l.preprocessed_stats;
  out->num_iteration_stats = static_cast<int64_t>(l.iteration_stats.size());
  if (!l.iteration_stats.empty()) {
    out->iteration_stats = static_cast<PdlpIterationStats*>(std::malloc(l.iteration_stats.size() * sizeof(PdlpIterationStats)));
    std::memcpy(out->iteration_stats, l.iteration_stats.data(), l.iteration_stats.size() * sizeof(PdlpIterationStats));
  }
  out->params = l.params;
  out->num_feasibility_polishing_details = static_cast<int64_t>(l.feasibility_polishing_details.size());
  if (!l.feasibility_polishing_details.empty()) {
    out->feasibility_polishing_details =
        static_cast<PdlpFeasibilityPolishingDetails*>(std::calloc(l.feasibility_polishing_details.size(), sizeof(PdlpFeasibilityPolishingDetails)));
    for (size_t k = 0; k < l.feasibility_polishing_details.size(); ++k) {
      const PolishingDetailsCpp& d = l.feasibility_polishing_details[k];
      PdlpFeasibilityPolishingDetails& o = out->feasibility_polishing_details[k];
      o.polishing_phase_type = d.polishing_phase_type;
      o.main_iteration_count = d.main_iteration_count;
      o.params = d.params;
      o.termination_reason = d.termination_reason;
      o.iteration_count = d.iteration_count;
      o.solve_time_sec = d.solve_time_sec;
      o.solution_stats = d.solution_stats;
      o.solution_type = d.solution_type;
      o.num_iteration_stats = static_cast<int64_t>(d.iteration_stats.size());
      if (!d.iteration_stats.empty()) {
        o.iteration_stats = static_cast<PdlpIterationStats*>(std::malloc(d.iteration_stats.size() * sizeof(PdlpIterationStats)));
        std::memcpy(o.iteration_stats, d.iteration_stats.data(), d.iteration_stats.size() * sizeof(PdlpIterationStats));
      }
    }
  }
}

}  // namespace

// Handle for the kernel-level entry points: a ShardedQuadraticProgram.
struct OracleProblem {
  std::unique_ptr<ShardedQp> sqp;
};

}  // namespace pdlp_oracle

using namespace pdlp_oracle;

static const Vec& OnesOr(const double* p, int64_t n, Vec& store) {
  if (p != nullptr) store.assign(p, p + n); else store.assign(n, 1.0);
  return store;
}

extern "C" {

void pdlp_oracle_params_set_defaults(PdlpParams* params) { SetDefaultParams(params); }

int32_t pdlp_oracle_params_validate(const PdlpParams* params, char* message, int64_t capacity) {
  const std::string e = ValidateParams(*params);
  if (message != nullptr && capacity > 0) { std::strncpy(message, e.c_str(), capacity - 1); message[capacity - 1] = 0; }
  return e.empty() ? 1 : 0;
}

int32_t pdlp_oracle_primal_dual_hybrid_gradient(const PdlpProblemView* qp, const PdlpParams* params, const double* initial_primal,
                                                int64_t initial_primal_size, const double* initial_dual, int64_t initial_dual_size,
                                                const volatile int32_t* interrupt_solve, PdlpMessageCallback message_callback,
                                                PdlpIterationStatsCallback stats_callback, void* user_data, PdlpResult* result) {
  Logger logger{message_callback, user_data};
  std::optional<InitialSolution> init;
  if (initial_primal != nullptr || initial_dual != nullptr) {
    init.emplace();
    if (initial_primal != nullptr) init->primal.assign(initial_primal, initial_primal + initial_primal_size);
    if (initial_dual != nullptr) init->dual.assign(initial_dual, initial_dual + initial_dual_size);
  }
  StatsCallback cb;
  if (stats_callback != nullptr) cb = [=](const PdlpIterationCallbackInfo& info) { stats_callback(&info, user_data); };
  FillResult(PrimalDualHybridGradient(*qp, *params, std::move(init), interrupt_solve, logger, std::move(cb)), result);
  return PDLP_B200_STATUS_OK;
}

void pdlp_oracle_result_free(PdlpResult* r) {
  if (r == nullptr) return;
  std::free(r->primal_solution); std::free(r->dual_solution); std::free(r->reduced_costs);
  std::free(r->instance_name); std::free(r->termination_string); std::free(r->iteration_stats);
  for (int64_t k = 0; k < r->num_feasibility_polishing_details; ++k) std::free(r->feasibility_polishing_details[k].iteration_stats);
  std::free(r->feasibility_polishing_details);
  std::memset(r, 0, sizeof(*r));
}

// ---- kernel-level entry points (same shapes as include/pdlp_b200.h) -------
int32_t pdlp_oracle_problem_create(const PdlpProblemView* qp, int32_t num_threads, int32_t num_shards, OracleProblem** out) {
  auto* p = new OracleProblem;
  const int shards = num_shards > 0 ? num_shards : (num_threads == 1 ? 1 : 4 * num_threads);
  p->sqp.reset(new ShardedQp(QpFromView(*qp), std::max(1, num_threads), shards));
  *out = p;
  return 0;
}
void pdlp_oracle_problem_destroy(OracleProblem* p) { delete p; }

// Sharder::ShardStartsForTesting for the four sharders: which = 0 constraint
// matrix, 1 transposed matrix, 2 primal, 3 dual. Returns number of entries.
int64_t pdlp_oracle_shard_starts(OracleProblem* p, int32_t which, int64_t* out, int64_t capacity) {
  const Sharder* s = which == 0 ? &p->sqp->ConstraintMatrixSharder() : which == 1 ? &p->sqp->TransposedConstraintMatrixSharder()
                   : which == 2 ? &p->sqp->PrimalSharder() : &p->sqp->DualSharder();
  const auto& st = s->starts();
  for (int64_t i = 0; i < static_cast<int64_t>(st.size()) && i < capacity; ++i) out[i] = st[i];
  return static_cast<int64_t>(st.size());
}
// Generic Sharder constructors for sharder_test.cc goldens: masses==NULL =>
// unit-mass constructor. Returns number of starts written.
int64_t pdlp_oracle_sharder_starts(int64_t num_elements, int32_t num_shards, const int64_t* masses, int64_t* out, int64_t capacity) {
  Sharder s = masses != nullptr ? Sharder(num_elements, num_shards, nullptr, [&](int64_t i) { return masses[i]; }) : Sharder(num_elements, num_shards, nullptr);
  const auto& st = s.starts();
  for (int64_t i = 0; i < static_cast<int64_t>(st.size()) && i < capacity; ++i) out[i] = st[i];
  return static_cast<int64_t>(st.size());
}

int32_t pdlp_oracle_transposed_matrix_vector_product(OracleProblem* p, const double* y, double* out) {
  const Vec v(y, y + p->sqp->DualSize());
  const Vec r = TransposedMatrixVectorProduct(p->sqp->Qp().constraint_matrix, v, p->sqp->ConstraintMatrixSharder());
  std::copy(r.begin(), r.end(), out);
  return 0;
}
int32_t pdlp_oracle_matrix_vector_product(OracleProblem* p, const double* x, double* out) {
  const Vec v(x, x + p->sqp->PrimalSize());
  const Vec r = TransposedMatrixVectorProduct(p->sqp->TransposedConstraintMatrix(), v, p->sqp->TransposedConstraintMatrixSharder());
  std::copy(r.begin(), r.end(), out);
  return 0;
}
int32_t pdlp_oracle_apply_rescaling(OracleProblem* p, int32_t ruiz, int32_t l2, double* row_scaling, double* col_scaling) {
  const ScalingVectors sv = ApplyRescaling(ruiz, l2 != 0, *p->sqp);
  std::copy(sv.row_scaling_vec.begin(), sv.row_scaling_vec.end(), row_scaling);
  std::copy(sv.col_scaling_vec.begin(), sv.col_scaling_vec.end(), col_scaling);
  return 0;
}
int32_t pdlp_oracle_scaling_iterations(OracleProblem* p, int32_t norm, int32_t iters, double* row_scaling, double* col_scaling) {
  Vec r(row_scaling, row_scaling + p->sqp->DualSize()), c(col_scaling, col_scaling + p->sqp->PrimalSize());
  ApplyScalingIterationsForNorm(*p->sqp, iters, norm == 0 ? ScalingNorm::kLInf : ScalingNorm::kL2, r, c);
  std::copy(r.begin(), r.end(), row_scaling);
  std::copy(c.begin(), c.end(), col_scaling);
  return 0;
}
int32_t pdlp_oracle_scaled_col_norm(OracleProblem* p, int32_t norm, const double* row_scaling, const double* col_scaling, double* out) {
  const Vec r(row_scaling, row_scaling + p->sqp->DualSize()), c(col_scaling, col_scaling + p->sqp->PrimalSize());
  const Vec o = norm == 0 ? ScaledColLInfNorm(p->sqp->Qp().constraint_matrix, r, c, p->sqp->ConstraintMatrixSharder())
                          : ScaledColL2Norm(p->sqp->Qp().constraint_matrix, r, c, p->sqp->ConstraintMatrixSharder());
  std::copy(o.begin(), o.end(), out);
  return 0;
}
int32_t pdlp_oracle_scaled_row_norm(OracleProblem* p, int32_t norm, const double* row_scaling, const double* col_scaling, double* out) {
  const Vec r(row_scaling, row_scaling + p->sqp->DualSize()), c(col_scaling, col_scaling + p->sqp->PrimalSize());
  const Vec o = norm == 0 ? ScaledColLInfNorm(p->sqp->TransposedConstraintMatrix(), c, r, p->sqp->TransposedConstraintMatrixSharder())
                          : ScaledColL2Norm(p->sqp->TransposedConstraintMatrix(), c, r, p->sqp->TransposedConstraintMatrixSharder());
  std::copy(o.begin(), o.end(), out);
  return 0;
}
int32_t pdlp_oracle_rescale_quadratic_program(OracleProblem* p, const double* col_scaling, const double* row_scaling) {
  const Vec c(col_scaling, col_scaling + p->sqp->PrimalSize()), r(row_scaling, row_scaling + p->sqp->DualSize());
  p->sqp->RescaleQuadraticProgram(c, r);
  return 0;
}
int32_t pdlp_oracle_problem_download(OracleProblem* p, double* values, double* objective_vector, double* objective_matrix_diagonal,
                                     double* clb, double* cub, double* vlb, double* vub) {
  const QuadraticProgram& qp = p->sqp->Qp();
  if (values) std::copy(qp.constraint_matrix.value.begin(), qp.constraint_matrix.value.end(), values);
  if (objective_vector) std::copy(qp.objective_vector.begin(), qp.objective_vector.end(), objective_vector);
  if (objective_matrix_diagonal && qp.objective_matrix) std::copy(qp.objective_matrix->begin(), qp.objective_matrix->end(), objective_matrix_diagonal);
  if (clb) std::copy(qp.constraint_lower_bounds.begin(), qp.constraint_lower_bounds.end(), clb);
  if (cub) std::copy(qp.constraint_upper_bounds.begin(), qp.constraint_upper_bounds.end(), cub);
  if (vlb) std::copy(qp.variable_lower_bounds.begin(), qp.variable_lower_bounds.end(), vlb);
  if (vub) std::copy(qp.variable_upper_bounds.begin(), qp.variable_upper_bounds.end(), vub);
  return 0;
}
// Values of the stored transpose in its own (row-major of K) order, for
// sharded_quadratic_program_test.cc goldens.
int32_t pdlp_oracle_transposed_values(OracleProblem* p, double* values) {
  const auto& v = p->sqp->TransposedConstraintMatrix().value;
  std::copy(v.begin(), v.end(), values);
  return 0;
}
int32_t pdlp_oracle_replace_large_constraint_bounds_with_infinity(OracleProblem* p, double threshold) {
  p->sqp->ReplaceLargeConstraintBoundsWithInfinity(threshold);
  return 0;
}
int32_t pdlp_oracle_has_valid_bounds(OracleProblem* p) { return HasValidBounds(*p->sqp) ? 1 : 0; }
int32_t pdlp_oracle_compute_stats(OracleProblem* p, PdlpQuadraticProgramStats* out) { *out = ComputeStats(*p->sqp); return 0; }
int32_t pdlp_oracle_project_to_primal_variable_bounds(OracleProblem* p, double* primal, int32_t use_feasibility_bounds) {
  Vec v(primal, primal + p->sqp->PrimalSize());
  ProjectToPrimalVariableBounds(*p->sqp, v, use_feasibility_bounds != 0);
  std::copy(v.begin(), v.end(), primal);
  return 0;
}
int32_t pdlp_oracle_project_to_dual_variable_bounds(OracleProblem* p, double* dual) {
  Vec v(dual, dual + p->sqp->DualSize());
  ProjectToDualVariableBounds(*p->sqp, v);
  std::copy(v.begin(), v.end(), dual);
  return 0;
}
int32_t pdlp_oracle_compute_primal_gradient(OracleProblem* p, const double* primal, const double* dual_product, double* gradient, double* value) {
  const int64_t n = p->sqp->PrimalSize();
  const LagrangianPart r = ComputePrimalGradient(*p->sqp, Vec(primal, primal + n), Vec(dual_product, dual_product + n));
  std::copy(r.gradient.begin(), r.gradient.end(), gradient);
  *value = r.value;
  return 0;
}
int32_t pdlp_oracle_compute_dual_gradient(OracleProblem* p, const double* dual, const double* primal_product, double* gradient, double* value) {
  const int64_t m = p->sqp->DualSize();
  const LagrangianPart r = ComputeDualGradient(*p->sqp, Vec(dual, dual + m), Vec(primal_product, primal_product + m));
  std::copy(r.gradient.begin(), r.gradient.end(), gradient);
  *value = r.value;
  return 0;
}
int32_t pdlp_oracle_compute_convergence_information(OracleProblem* p, const PdlpParams* params, const double* col_scaling, const double* row_scaling,
                                                    const double* primal, const double* dual, double cw_primal_offset, double cw_dual_offset,
                                                    int32_t candidate_type, PdlpConvergenceInformation* out) {
  const int64_t n = p->sqp->PrimalSize(), m = p->sqp->DualSize();
  Vec cs, rs;
  *out = ComputeConvergenceInformation(params->handle_some_primal_gradients_on_finite_bounds_as_residuals != 0, *p->sqp, OnesOr(col_scaling, n, cs),
                                       OnesOr(row_scaling, m, rs), Vec(primal, primal + n), Vec(dual, dual + m), cw_primal_offset, cw_dual_offset, candidate_type);
  return 0;
}
int32_t pdlp_oracle_compute_infeasibility_information(OracleProblem* p, const PdlpParams* params, const double* col_scaling, const double* row_scaling,
                                                      const double* primal_ray, const double* dual_ray, const double* primal_for_residual_tests,
                                                      int32_t candidate_type, PdlpInfeasibilityInformation* out) {
  const int64_t n = p->sqp->PrimalSize(), m = p->sqp->DualSize();
  Vec cs, rs;
  *out = ComputeInfeasibilityInformation(params->handle_some_primal_gradients_on_finite_bounds_as_residuals != 0, *p->sqp, OnesOr(col_scaling, n, cs),
                                         OnesOr(row_scaling, m, rs), Vec(primal_ray, primal_ray + n), Vec(dual_ray, dual_ray + m),
                                         Vec(primal_for_residual_tests, primal_for_residual_tests + n), candidate_type);
  return 0;
}
int32_t pdlp_oracle_reduced_costs(OracleProblem* p, const PdlpParams*, const double* primal, const double* dual, int32_t use_zero_primal_objective, double* out) {
  const int64_t n = p->sqp->PrimalSize(), m = p->sqp->DualSize();
  const Vec r = ReducedCosts(*p->sqp, Vec(primal, primal + n), Vec(dual, dual + m), use_zero_primal_objective != 0);
  std::copy(r.begin(), r.end(), out);
  return 0;
}
// norm_kind: 0 = Euclidean (what the solver uses), 1 = max norm.
int32_t pdlp_oracle_compute_localized_lagrangian_bounds(OracleProblem* p, const double* primal, const double* dual, double primal_weight, double radius,
                                                        const double* primal_product, const double* dual_product, int32_t use_diagonal_qp_solver,
                                                        double diagonal_tol, int32_t norm_kind, double out[4]) {
  const int64_t n = p->sqp->PrimalSize(), m = p->sqp->DualSize();
  Vec pp, dp;
  if (primal_product) pp.assign(primal_product, primal_product + m);
  if (dual_product) dp.assign(dual_product, dual_product + n);
  const LocalizedLagrangianBounds b = ComputeLocalizedLagrangianBounds(
      *p->sqp, Vec(primal, primal + n), Vec(dual, dual + m), norm_kind == 0 ? PrimalDualNorm::kEuclideanNorm : PrimalDualNorm::kMaxNorm, primal_weight,
      radius, primal_product ? &pp : nullptr, dual_product ? &dp : nullptr, use_diagonal_qp_solver != 0, diagonal_tol);
  out[0] = b.lagrangian_value; out[1] = b.lower_bound; out[2] = b.upper_bound; out[3] = b.radius;
  return 0;
}
int32_t pdlp_oracle_solve_trust_region(int32_t num_threads, int32_t num_shards, int64_t size, const double* objective, const double* lb, const double* ub,
                                       const double* center, const double* weights, double target_radius, double* solution, double* step_size,
                                       double* objective_value) {
  std::unique_ptr<ThreadPool> pool(num_threads > 1 ? new ThreadPool(num_threads) : nullptr);
  const Sharder sharder(size, std::max(1, num_shards), pool.get());
  const TrustRegionResult r = SolveTrustRegion(Vec(objective, objective + size), Vec(lb, lb + size), Vec(ub, ub + size), Vec(center, center + size),
                                               Vec(weights, weights + size), target_radius, sharder);
  std::copy(r.solution.begin(), r.solution.end(), solution);
  *step_size = r.solution_step_size; *objective_value = r.objective_value;
  return 0;
}
int32_t pdlp_oracle_solve_diagonal_trust_region(int32_t num_threads, int32_t num_shards, int64_t size, const double* objective, const double* qdiag,
                                                const double* lb, const double* ub, const double* center, const double* weights, double target_radius,
                                                double tol, double* solution, double* step_size, double* objective_value) {
  std::unique_ptr<ThreadPool> pool(num_threads > 1 ? new ThreadPool(num_threads) : nullptr);
  const Sharder sharder(size, std::max(1, num_shards), pool.get());
  const TrustRegionResult r = SolveDiagonalTrustRegion(Vec(objective, objective + size), Vec(qdiag, qdiag + size), Vec(lb, lb + size), Vec(ub, ub + size),
                                                       Vec(center, center + size), Vec(weights, weights + size), target_radius, sharder, tol);
  std::copy(r.solution.begin(), r.solution.end(), solution);
  *step_size = r.solution_step_size; *objective_value = r.objective_value;
  return 0;
}
int32_t pdlp_oracle_weighted_average(int32_t num_shards, int64_t size, int64_t count, const double* datapoints, const double* weights, double* out_average,
                                     double* out_sum_weights, int32_t* out_num_terms) {
  const Sharder sharder(size, std::max(1, num_shards), nullptr);
  WeightedAverage avg(&sharder);
  for (int64_t k = 0; k < count; ++k) avg.Add(Vec(datapoints + k * size, datapoints + (k + 1) * size), weights[k]);
  const Vec a = avg.ComputeAverage();
  std::copy(a.begin(), a.end(), out_average);
  if (out_sum_weights) *out_sum_weights = avg.Weight();
  if (out_num_terms) *out_num_terms = avg.NumTerms();
  return 0;
}
int32_t pdlp_oracle_vector_reduce(int32_t num_shards, int32_t op, int64_t size, const double* a, const double* b, double* out) {
  const Sharder sharder(size, std::max(1, num_shards), nullptr);
  const Vec va(a, a + size);
  Vec vb;
  if (b != nullptr) vb.assign(b, b + size);
  switch (op) {
    case PDLP_VECOP_DOT: *out = Dot(va, vb, sharder); break;
    case PDLP_VECOP_LINF_NORM: *out = LInfNorm(va, sharder); break;
    case PDLP_VECOP_L1_NORM: *out = L1Norm(va, sharder); break;
    case PDLP_VECOP_SQUARED_NORM: *out = SquaredNorm(va, sharder); break;
    case PDLP_VECOP_NORM: *out = Norm(va, sharder); break;
    case PDLP_VECOP_SQUARED_DISTANCE: *out = SquaredDistance(va, vb, sharder); break;
    case PDLP_VECOP_DISTANCE: *out = Distance(va, vb, sharder); break;
    case PDLP_VECOP_SCALED_LINF_NORM: *out = ScaledLInfNorm(va, vb, sharder); break;
    case PDLP_VECOP_SCALED_SQUARED_NORM: *out = ScaledSquaredNorm(va, vb, sharder); break;
    case PDLP_VECOP_SCALED_NORM: *out = ScaledNorm(va, vb, sharder); break;
    default: return PDLP_B200_STATUS_BAD_ARGUMENT;
  }
  return 0;
}
// AddScaledVector / CoefficientWise{Product,Quotient}InPlace (sharder.cc:196-228):
// op 0: dest += scale*a ; 1: dest *= a ; 2: dest /= a.
int32_t pdlp_oracle_vector_update(int32_t num_shards, int32_t op, int64_t size, double scale, const double* a, double* dest) {
  const Sharder sharder(size, std::max(1, num_shards), nullptr);
  Vec d(dest, dest + size);
  const Vec va(a, a + size);
  if (op == 0) AddScaledVector(scale, va, sharder, d);
  else if (op == 1) CoefficientWiseProductInPlace(va, sharder, d);
  else CoefficientWiseQuotientInPlace(va, sharder, d);
  std::copy(d.begin(), d.end(), dest);
  return 0;
}
// EstimateMaximumSingularValueOfConstraintMatrix (sou.cc:676-699); primal/dual may be NULL.
int32_t pdlp_oracle_estimate_max_singular_value(OracleProblem* p, const double* primal, const double* dual, double desired_relative_error,
                                                double failure_probability, uint32_t seed, double* singular_value, int32_t* num_iterations) {
  std::mt19937 gen(seed);
  std::optional<Vec> ps, ds;
  if (primal) ps = Vec(primal, primal + p->sqp->PrimalSize());
  if (dual) ds = Vec(dual, dual + p->sqp->DualSize());
  const auto r = EstimateMaximumSingularValueOfConstraintMatrix(*p->sqp, ps, ds, desired_relative_error, failure_probability, gen);
  *singular_value = r.singular_value; *num_iterations = r.num_iterations;
  return 0;
}
// Termination predicates on caller-supplied stats (termination_test.cc goldens).
int32_t pdlp_oracle_check_simple_termination_criteria(const PdlpTerminationCriteria* c, const PdlpIterationStats* stats, const volatile int32_t* interrupt,
                                                      int32_t* reason, int32_t* type) {
  const auto r = CheckSimpleTerminationCriteria(*c, *stats, interrupt);
  if (!r) return 0;
  *reason = r->reason; *type = r->type;
  return 1;
}
int32_t pdlp_oracle_check_iterate_termination_criteria(const PdlpTerminationCriteria* c, const PdlpIterationStats* stats, const PdlpBoundNorms* bn,
                                                       int32_t force_numerical_termination, int32_t* reason, int32_t* type) {
  const auto r = CheckIterateTerminationCriteria(*c, *stats, *bn, force_numerical_termination != 0);
  if (!r) return 0;
  *reason = r->reason; *type = r->type;
  return 1;
}
int32_t pdlp_oracle_compute_relative_residuals(const PdlpTerminationCriteria* c, const PdlpConvergenceInformation* s, const PdlpBoundNorms* bn, double out[5]) {
  const RelativeConvergenceInformation r = ComputeRelativeResiduals(EffectiveOptimalityCriteria(*c), *s, *bn);
  out[0] = r.relative_l_inf_primal_residual; out[1] = r.relative_l2_primal_residual; out[2] = r.relative_l_inf_dual_residual;
  out[3] = r.relative_l2_dual_residual; out[4] = r.relative_optimality_gap;
  return 0;
}
int32_t pdlp_oracle_optimality_criteria_met(const PdlpTerminationCriteria* c, const PdlpConvergenceInformation* s, const PdlpBoundNorms* bn,
                                            int32_t* objective_gap_met) {
  const DetailedCriteria oc = EffectiveOptimalityCriteria(*c);
  if (objective_gap_met != nullptr) *objective_gap_met = ObjectiveGapMet(oc, *s) ? 1 : 0;
  return OptimalityCriteriaMet(oc, *s, c->optimality_norm, *bn) ? 1 : 0;
}
void pdlp_oracle_effective_optimality_criteria(const PdlpTerminationCriteria* c, double out[6]) {
  const DetailedCriteria oc = EffectiveOptimalityCriteria(*c);
  out[0] = oc.primal_abs; out[1] = oc.primal_rel; out[2] = oc.dual_abs; out[3] = oc.dual_rel; out[4] = oc.gap_abs; out[5] = oc.gap_rel;
}
void pdlp_oracle_bound_norms_from_problem_stats(const PdlpQuadraticProgramStats* s, PdlpBoundNorms* out) { *out = BoundNormsFromProblemStats(*s); }

const char* pdlp_oracle_version(void) { return "pdlp-oracle 0.1 (CPU restatement of or-tools 9.15 ortools/pdlp)"; }

}  // extern "C"
