// solver.cc -- host driver of the device-resident PDHG solve.
//
// Mirrors PrimalDualHybridGradient() / PreprocessSolver / Solver of
// ortools/pdlp/primal_dual_hybrid_gradient.cc, restructured for the GPU:
//  * every vector stays in HBM; the host only sees reduced scalars;
//  * the adaptive / constant step loop (pdhg.cc:2558-2675) runs on the device:
//    accept test and step-size rule are evaluated by k_step_decide, so a
//    rejected step costs no host round trip. The host enqueues attempts up to
//    the next *checkpoint* -- the next iteration at which the reference's
//    MajorIterationAndTerminationCheck (pdhg.cc:2360-2435) would do anything
//    (termination check, major iteration, artificial restart, iteration or
//    KKT-pass limit) -- and replays that function there.
//  * the Malitsky-Pock rule (pdhg.cc:2463-2556) is host-driven per inner step.
#include "solver.h"

#include "comm.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <stdexcept>

#include "device_problem.h"

namespace pdlp_b200 {

void Logger::Log(const std::string& s) const {
  if (cb != nullptr) cb(s.c_str(), user);
  else { std::fputs(s.c_str(), stdout); std::fputc('\n', stdout); }
}

namespace {

constexpr double kInf = std::numeric_limits<double>::infinity();
inline double Sq(double v) { return v * v; }

std::string Fmt(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  std::vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  return buf;
}
std::string G(double v) { return Fmt("%g", v); }

struct WallTimer {
  using Clock = std::chrono::steady_clock;
  Clock::time_point start = Clock::now();
  double accumulated = 0;
  bool running = true;
  void Start() { start = Clock::now(); accumulated = 0; running = true; }
  void Stop() { if (running) { accumulated += std::chrono::duration<double>(Clock::now() - start).count(); running = false; } }
  void Resume() { if (!running) { start = Clock::now(); running = true; } }
  double Get() const { return accumulated + (running ? std::chrono::duration<double>(Clock::now() - start).count() : 0.0); }
};

SolverResultCpp ErrorSolverResult(int reason, const std::string& message, const Logger& logger) {  // pdhg.cc:776-785
  SolverResultCpp r;
  r.solve_log.termination_reason = reason;
  r.solve_log.termination_string = message;
  logger.Log("The solver did not run because of invalid input: " + message);
  return r;
}

// ---- termination.cc (scalar logic, host) -----------------------------------
struct DetailedCriteria { double primal_abs, primal_rel, dual_abs, dual_rel, gap_abs, gap_rel; };
DetailedCriteria EffectiveOptimalityCriteria(const PdlpTerminationCriteria& c) {  // termination.cc:126-159
  if (c.optimality_criteria_case == PDLP_DETAILED_OPTIMALITY_CRITERIA)
    return {c.eps_optimal_primal_residual_absolute, c.eps_optimal_primal_residual_relative, c.eps_optimal_dual_residual_absolute,
            c.eps_optimal_dual_residual_relative, c.eps_optimal_objective_gap_absolute, c.eps_optimal_objective_gap_relative};
  const bool simple = c.optimality_criteria_case == PDLP_SIMPLE_OPTIMALITY_CRITERIA;
  const double a = simple ? c.simple_eps_optimal_absolute : c.eps_optimal_absolute;
  const double r = simple ? c.simple_eps_optimal_relative : c.eps_optimal_relative;
  return {a, r, a, r, a, r};
}
double EpsilonRatio(double a, double r) { return a == r ? 1.0 : a / r; }  // termination.cc:230-237
bool ObjectiveGapMet(const DetailedCriteria& oc, const PdlpConvergenceInformation& s) {  // termination.cc:26-41
  if (std::isinf(oc.gap_abs) || std::isinf(oc.gap_rel)) return true;
  const double abs_obj = std::abs(s.primal_objective) + std::abs(s.dual_objective);
  const double gap = std::abs(s.primal_objective - s.dual_objective);
  return std::isfinite(abs_obj) && gap <= oc.gap_abs + oc.gap_rel * abs_obj;
}
bool OptimalityCriteriaMet(const DetailedCriteria& oc, const PdlpConvergenceInformation& s, int norm, const PdlpBoundNorms& bn) {  // :43-97
  double perr, pbase, derr, dbase, pabs = oc.primal_abs, dabs = oc.dual_abs;
  if (norm == PDLP_OPTIMALITY_NORM_L_INF) {
    perr = s.l_inf_primal_residual; pbase = bn.l_inf_norm_constraint_bounds; derr = s.l_inf_dual_residual; dbase = bn.l_inf_norm_primal_linear_objective;
  } else if (norm == PDLP_OPTIMALITY_NORM_L_INF_COMPONENTWISE) {
    perr = s.l_inf_componentwise_primal_residual; pbase = 1.0; pabs = 0.0; derr = s.l_inf_componentwise_dual_residual; dbase = 1.0; dabs = 0.0;
  } else {
    perr = s.l2_primal_residual; pbase = bn.l2_norm_constraint_bounds; derr = s.l2_dual_residual; dbase = bn.l2_norm_primal_linear_objective;
  }
  const bool p_ok = std::isinf(oc.primal_abs) || std::isinf(oc.primal_rel) || perr <= pabs + oc.primal_rel * pbase;
  const bool d_ok = std::isinf(oc.dual_abs) || std::isinf(oc.dual_rel) || derr <= dabs + oc.dual_rel * dbase;
  return p_ok && d_ok && ObjectiveGapMet(oc, s);
}
struct ReasonAndType { int reason, type; };
std::optional<ReasonAndType> CheckSimpleTerminationCriteria(const PdlpTerminationCriteria& c, const PdlpIterationStats& st,
                                                            bool interrupted) {  // termination.cc:161-184
  if (st.iteration_number >= c.iteration_limit) return ReasonAndType{PDLP_TERMINATION_REASON_ITERATION_LIMIT, PDLP_POINT_TYPE_NONE};
  if (st.cumulative_kkt_matrix_passes >= c.kkt_matrix_pass_limit) return ReasonAndType{PDLP_TERMINATION_REASON_KKT_MATRIX_PASS_LIMIT, PDLP_POINT_TYPE_NONE};
  if (st.cumulative_time_sec >= c.time_sec_limit) return ReasonAndType{PDLP_TERMINATION_REASON_TIME_LIMIT, PDLP_POINT_TYPE_NONE};
  if (interrupted) return ReasonAndType{PDLP_TERMINATION_REASON_INTERRUPTED_BY_USER, PDLP_POINT_TYPE_NONE};
  return std::nullopt;
}
std::optional<ReasonAndType> CheckIterateTerminationCriteria(const PdlpTerminationCriteria& c, const PdlpIterationStats& st,
                                                             const PdlpBoundNorms& bn, bool force_numerical) {  // termination.cc:186-219
  const DetailedCriteria oc = EffectiveOptimalityCriteria(c);
  for (int i = 0; i < st.num_convergence_information; ++i)
    if (OptimalityCriteriaMet(oc, st.convergence_information[i], c.optimality_norm, bn))
      return ReasonAndType{PDLP_TERMINATION_REASON_OPTIMAL, st.convergence_information[i].candidate_type};
  for (int i = 0; i < st.num_infeasibility_information; ++i) {
    const PdlpInfeasibilityInformation& s = st.infeasibility_information[i];
    if (s.dual_ray_objective > 0.0 && s.max_dual_ray_infeasibility / s.dual_ray_objective <= c.eps_primal_infeasible)
      return ReasonAndType{PDLP_TERMINATION_REASON_PRIMAL_INFEASIBLE, s.candidate_type};
    if (s.primal_ray_linear_objective < 0.0 && s.max_primal_ray_infeasibility / -s.primal_ray_linear_objective <= c.eps_dual_infeasible &&
        s.primal_ray_quadratic_norm / -s.primal_ray_linear_objective <= c.eps_dual_infeasible)
      return ReasonAndType{PDLP_TERMINATION_REASON_DUAL_INFEASIBLE, s.candidate_type};
  }
  if (force_numerical) return ReasonAndType{PDLP_TERMINATION_REASON_NUMERICAL_ERROR, PDLP_POINT_TYPE_NONE};
  return std::nullopt;
}
PdlpBoundNorms BoundNormsFromProblemStats(const PdlpQuadraticProgramStats& s) {  // termination.cc:221-228
  return {s.objective_vector_l2_norm, s.combined_bounds_l2_norm, s.objective_vector_abs_max, s.combined_bounds_max};
}

struct RelativeResiduals { double l_inf_primal, l2_primal, l_inf_dual, l2_dual, gap; };
RelativeResiduals ComputeRelativeResiduals(const DetailedCriteria& oc, const PdlpConvergenceInformation& s, const PdlpBoundNorms& bn) {  // termination.cc:239-271
  const double rp = EpsilonRatio(oc.primal_abs, oc.primal_rel), rd = EpsilonRatio(oc.dual_abs, oc.dual_rel), rg = EpsilonRatio(oc.gap_abs, oc.gap_rel);
  RelativeResiduals r;
  r.l_inf_primal = s.l_inf_primal_residual / (rp + bn.l_inf_norm_constraint_bounds);
  r.l2_primal = s.l2_primal_residual / (rp + bn.l2_norm_constraint_bounds);
  r.l_inf_dual = s.l_inf_dual_residual / (rd + bn.l_inf_norm_primal_linear_objective);
  r.l2_dual = s.l2_dual_residual / (rd + bn.l2_norm_primal_linear_objective);
  const double abs_obj = std::abs(s.primal_objective) + std::abs(s.dual_objective);  // summed first, like termination.cc:266-268
  r.gap = (s.primal_objective - s.dual_objective) / (rg + abs_obj);
  return r;
}

const PdlpConvergenceInformation* GetConvergenceInformation(const PdlpIterationStats& s, int type) {
  for (int i = 0; i < s.num_convergence_information; ++i)
    if (s.convergence_information[i].candidate_type == type) return &s.convergence_information[i];
  return nullptr;
}

// ---- log table (pdhg.cc:127-318) --------------------------------------------
void LogIterationStatsHeader(int verbosity, const Logger& logger) {
  const std::string work = verbosity >= 3 ? Fmt("%6s %8s %6s", "iter#", "kkt_pass", "time") : Fmt("%6s %6s", "iter#", "time");
  const std::string conv = verbosity >= 3
      ? Fmt("%12s %12s %12s | %12s %12s %12s | %12s %12s | %12s %12s", "rel_prim_res", "rel_dual_res", "rel_gap", "prim_resid", "dual_resid",
            "obj_gap", "prim_obj", "dual_obj", "prim_var_l2", "dual_var_l2")
      : Fmt("%10s %10s %10s | %10s %10s", "rel_p_res", "rel_d_res", "rel_gap", "prim_obj", "dual_obj");
  logger.Log(std::string(verbosity >= 4 ? "I " : "") + work + " | " + conv);
}
void LogIterationStats(int verbosity, const PdlpIterationStats& st, const PdlpTerminationCriteria& tc, const PdlpBoundNorms& bn, int preferred,
                       const Logger& logger) {
  const std::string iter = verbosity >= 3 ? Fmt("%6d %8.1f %6.1f", st.iteration_number, st.cumulative_kkt_matrix_passes, st.cumulative_time_sec)
                                          : Fmt("%6d %6.1f", st.iteration_number, st.cumulative_time_sec);
  const PdlpConvergenceInformation* ci = GetConvergenceInformation(st, preferred);
  if (ci == nullptr && st.num_convergence_information > 0) ci = &st.convergence_information[0];
  if (ci == nullptr) { logger.Log(std::string(verbosity >= 4 ? "? " : "") + iter); return; }
  const char* tag = "";
  if (verbosity >= 4)
    tag = ci->candidate_type == PDLP_POINT_TYPE_CURRENT_ITERATE ? "C " : ci->candidate_type == PDLP_POINT_TYPE_AVERAGE_ITERATE ? "A "
        : ci->candidate_type == PDLP_POINT_TYPE_ITERATE_DIFFERENCE ? "D " : "? ";
  const RelativeResiduals rr = ComputeRelativeResiduals(EffectiveOptimalityCriteria(tc), *ci, bn);
  const double rel_gap = rr.gap;
  double relp, reld, absp, absd;
  if (tc.optimality_norm == PDLP_OPTIMALITY_NORM_L_INF) {
    relp = rr.l_inf_primal; reld = rr.l_inf_dual;
    absp = ci->l_inf_primal_residual; absd = ci->l_inf_dual_residual;
  } else if (tc.optimality_norm == PDLP_OPTIMALITY_NORM_L_INF_COMPONENTWISE) {
    relp = ci->l_inf_componentwise_primal_residual; reld = ci->l_inf_componentwise_dual_residual; absp = ci->l_inf_primal_residual; absd = ci->l_inf_dual_residual;
  } else {
    relp = rr.l2_primal; reld = rr.l2_dual;
    absp = ci->l2_primal_residual; absd = ci->l2_dual_residual;
  }
  const std::string conv = verbosity >= 3
      ? Fmt("%#12.6g %#12.6g %#12.6g | %#12.6g %#12.6g %#12.6g | %#12.6g %#12.6g | %#12.6g %#12.6g", relp, reld, rel_gap, absp, absd,
            ci->primal_objective - ci->dual_objective, ci->primal_objective, ci->dual_objective, ci->l2_primal_variable, ci->l2_dual_variable)
      : Fmt("%#10.4g %#10.4g %#10.4g | %#10.4g %#10.4g", relp, reld, rel_gap, ci->primal_objective, ci->dual_objective);
  logger.Log(std::string(tag) + iter + " | " + conv);
}

// pdhg.cc:791-983
std::optional<SolverResultCpp> CheckProblemStats(const PdlpQuadraticProgramStats& s, double objective_offset, bool check_small, const Logger& logger) {
  const double kBig = 1e50, kSmall = 1e-50, kRange = 1e20;
  auto err = [&](const std::string& m) { return ErrorSolverResult(PDLP_TERMINATION_REASON_INVALID_PROBLEM, m, logger); };
  auto warn_range = [&](const char* what, double mx, double mn) {
    if (mx > kRange * mn) logger.Log(std::string("WARNING: ") + what + " has largest absolute value " + G(mx) + " and smallest non-zero absolute value " + G(mn) + "; performance may suffer.");
  };
  if (std::isnan(s.constraint_matrix_l2_norm)) return err("Constraint matrix has a NAN.");
  if (s.constraint_matrix_abs_max > kBig) return err("Constraint matrix has a non-zero with absolute value " + G(s.constraint_matrix_abs_max) + " which exceeds limit of " + G(kBig) + ".");
  warn_range("Constraint matrix", s.constraint_matrix_abs_max, s.constraint_matrix_abs_min);
  if (s.constraint_matrix_col_min_l_inf_norm > 0 && s.constraint_matrix_col_min_l_inf_norm < kSmall)
    return err("Constraint matrix has a column with Linf norm " + G(s.constraint_matrix_col_min_l_inf_norm) + " which is less than limit of " + G(kSmall) + ".");
  if (s.constraint_matrix_row_min_l_inf_norm > 0 && s.constraint_matrix_row_min_l_inf_norm < kSmall)
    return err("Constraint matrix has a row with Linf norm " + G(s.constraint_matrix_row_min_l_inf_norm) + " which is less than limit of " + G(kSmall) + ".");
  if (std::isnan(s.combined_bounds_l2_norm)) return err("Constraint bounds vector has a NAN.");
  if (s.combined_bounds_max > kBig) return err("Combined constraint bounds vector has a non-zero with absolute value " + G(s.combined_bounds_max) + " which exceeds limit of " + G(kBig) + ".");
  if (check_small && s.combined_bounds_min > 0 && s.combined_bounds_min < kSmall)
    return err("Combined constraint bounds vector has a non-zero with absolute value " + G(s.combined_bounds_min) + " which is less than the limit of " + G(kSmall) + ".");
  warn_range("Combined constraint bounds vector", s.combined_bounds_max, s.combined_bounds_min);
  if (std::isnan(s.combined_variable_bounds_l2_norm)) return err("Variable bounds vector has a NAN.");
  if (s.combined_variable_bounds_max > kBig) return err("Combined variable bounds vector has a non-zero with absolute value " + G(s.combined_variable_bounds_max) + " which exceeds limit of " + G(kBig) + ".");
  if (check_small && s.combined_variable_bounds_min > 0 && s.combined_variable_bounds_min < kSmall)
    return err("Combined variable bounds vector has a non-zero with absolute value " + G(s.combined_variable_bounds_min) + " which is less than the limit of " + G(kSmall) + ".");
  warn_range("Combined variable bounds vector", s.combined_variable_bounds_max, s.combined_variable_bounds_min);
  warn_range("Variable bound gap vector", s.variable_bound_gaps_max, s.variable_bound_gaps_min);
  if (std::isnan(objective_offset)) return err("Objective offset is NAN.");
  if (std::abs(objective_offset) > kBig) return err("Objective offset " + G(objective_offset) + " has absolute value which exceeds limit of " + G(kBig) + ".");
  if (std::isnan(s.objective_vector_l2_norm)) return err("Objective vector has a NAN.");
  if (s.objective_vector_abs_max > kBig) return err("Objective vector has a non-zero with absolute value " + G(s.objective_vector_abs_max) + " which exceeds limit of " + G(kBig) + ".");
  if (check_small && s.objective_vector_abs_min > 0 && s.objective_vector_abs_min < kSmall)
    return err("Objective vector has a non-zero with absolute value " + G(s.objective_vector_abs_min) + " which is less than the limit of " + G(kSmall) + ".");
  warn_range("Objective vector", s.objective_vector_abs_max, s.objective_vector_abs_min);
  if (std::isnan(s.objective_matrix_l2_norm)) return err("Objective matrix has a NAN.");
  if (s.objective_matrix_abs_max > kBig) return err("Objective matrix has a non-zero with absolute value " + G(s.objective_matrix_abs_max) + " which exceeds limit of " + G(kBig) + ".");
  warn_range("Objective matrix", s.objective_matrix_abs_max, s.objective_matrix_abs_min);
  return std::nullopt;
}

std::string ValidateDimensions(const PdlpProblemView& v) {  // quadratic_program.cc:38-97
  auto sz = [](int64_t given, int64_t dflt) { return given < 0 ? dflt : given; };
  const long long n = v.num_variables, m = v.num_constraints;
  const long long var_lb = sz(v.variable_lower_bounds_size, n), var_ub = sz(v.variable_upper_bounds_size, n), obj = sz(v.objective_vector_size, n);
  const long long con_lb = sz(v.constraint_lower_bounds_size, m), con_ub = sz(v.constraint_upper_bounds_size, m);
  if (var_lb != var_ub) return Fmt("Inconsistent dimensions: variable lower bound vector has size %lld while variable upper bound vector has size %lld", var_lb, var_ub);
  if (var_lb != obj) return Fmt("Inconsistent dimensions: variable lower bound vector has size %lld while objective vector has size %lld", var_lb, obj);
  if (var_lb != n) return Fmt("Inconsistent dimensions: variable lower bound vector has size %lld while constraint matrix has %lld columns", var_lb, n);
  if (v.objective_matrix_diagonal != nullptr && var_lb != sz(v.objective_matrix_size, n))
    return Fmt("Inconsistent dimensions: variable lower bound vector has size %lld while objective matrix has %lld rows", var_lb, (long long)sz(v.objective_matrix_size, n));
  if (con_lb != con_ub) return Fmt("Inconsistent dimensions: constraint lower bound vector has size %lld while constraint upper bound vector has size %lld", con_lb, con_ub);
  if (con_lb != m) return Fmt("Inconsistent dimensions: constraint lower bound vector has size %lld while constraint matrix has %lld rows ", con_lb, m);
  return "";

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

struct LocalizedBounds { double lagrangian_value, lower_bound, upper_bound, radius; };
inline double BoundGap(const LocalizedBounds& b) { return b.upper_bound - b.lower_bound; }

// ---------------------------------------------------------------------------
// The device solve. One object plays both PreprocessSolver (original-problem
// bookkeeping, scaling vectors, termination checks) and Solver (iterate state).
// ---------------------------------------------------------------------------
class DeviceSolve {
 public:
  DeviceSolve(DeviceProblem& p, const PdlpParams& params, const Logger& logger, StatsCallback cb)
      : P(p), D(p.dev()), params_(params), logger_(logger), callback_(std::move(cb)) {}
  ~DeviceSolve() {
    if (const char* t = std::getenv("PDLP_B200_TRACE"); t != nullptr && t[0] == '1') {
      std::fprintf(stderr, "[pdlp_b200 trace] host wall seconds: restart-choice %.4f termination-check %.4f apply-restart %.4f step-loop %.4f flush+gather %.4f deltas %.4f\n",
                   phase_s_[0], phase_s_[1], phase_s_[2], phase_s_[3], phase_s_[4], phase_s_[5]);
    }
    for (int k = 0; k < 3; ++k) { D.Free(buf_.x[k]); D.Free(buf_.y[k]); D.Free(buf_.kty[k]); D.Free(buf_.kx[k]); }
    for (double* v : {buf_.x_tilde, buf_.avg_x, buf_.avg_y, buf_.avg_kx, buf_.avg_kty, x0_, y0_, delta_x_, delta_y_, pc_kx_avg_, pc_kty_avg_, polish_x_, polish_y_}) D.Free(v);
    if (!nested_) { D.Free(dc_); D.Free(dr_); }
    D.Free(buf_.state);
  }

  SolverResultCpp PreprocessAndSolve(std::optional<InitialSolution> initial_solution, const volatile int32_t* interrupt_solve, const std::string* name);
  // Resumable form (sessions): Prepare = PreprocessSolver work up to the first
  // iteration; Advance = the loop of Solver::Solve (pdhg.cc:3042-3091) until
  // `target_iterations` are completed or a termination criterion fires.
  std::optional<SolverResultCpp> Prepare(std::optional<InitialSolution> initial_solution, const std::string* name);
  std::optional<SolverResultCpp> Advance(int target_iterations, const volatile int32_t* interrupt_solve);
  SolverResultCpp ConstructOriginalSolverResult(SolverResultCpp result);
  void FillStatus(PdlpSessionStatus* out) const;
  void EnableTiming(bool on, int stride) { D.EnableStepTiming(on, stride); }
  void ForceRecheck() { check_done_ = false; }

 private:
  enum class Outcome { kSuccessful, kForceNumericalTermination };

  // ---- state helpers --------------------------------------------------------
  double* X() const { return buf_.x[hs_.cur]; }
  double* Y() const { return buf_.y[hs_.cur]; }
  double* Kty() const { return buf_.kty[hs_.cur]; }
  double* Kx() const { return buf_.kx[hs_.cur]; }  // K x of the current iterate (kept by the dual kernel: K x' = (K x~ + K x) / 2)
  bool PrimalAvgHasWeight() const { return avg_x_weight_ > 0.0; }
  bool DualAvgHasWeight() const { return avg_y_weight_ > 0.0; }
  const double* PrimalAverage() const { return PrimalAvgHasWeight() ? buf_.avg_x : X(); }  // pdhg.cc:2172-2186
  const double* DualAverage() const { return DualAvgHasWeight() ? buf_.avg_y : Y(); }
  void PushState() { D.UploadState(buf_.state, hs_); state_slot_ = 0; }  // (slot 0 of the two device slots)
  void ClearAverages() {
    InvalidateProducts(false, true);
    D.Fill(buf_.avg_x, 0.0, P.n());
    D.Fill(buf_.avg_y, 0.0, P.m());
    if (buf_.avg_kx != nullptr) {
      D.Fill(buf_.avg_kx, 0.0, P.m());
      D.Fill(buf_.avg_kty, 0.0, P.n());
      // valid from an empty average on, as long as every term enters through the peer-exchange step kernels
      avg_products_ok_ = params_.linesearch_rule != PDLP_MALITSKY_POCK_LINESEARCH_RULE;
    }
    avg_x_weight_ = avg_y_weight_ = 0.0;
    avg_x_terms_ = avg_y_terms_ = 0;
    hs_.avg_weight_sum = 0.0;
    hs_.avg_num_terms = 0;
    hs_.pending_ratio = 0.0;
  }
  void AverageAdd(bool primal, const double* v, double weight) {  // ShardedWeightedAverage::Add, sou.cc:54-66
    double& w = primal ? avg_x_weight_ : avg_y_weight_;
    int& terms = primal ? avg_x_terms_ : avg_y_terms_;
    if (weight > 0.0) {
      InvalidateProducts(false, true);
      avg_products_ok_ = false;  // (a term added from the host: the maintained products of the average no longer match)
      D.WeightedAverageAdd(primal ? buf_.avg_x : buf_.avg_y, v, weight / (w + weight), primal ? P.n() : P.m());
      w += weight;
    }
    ++terms;
  }
  void ResetAverageToCurrent() {  // pdhg.cc:2437-2442
    ClearAverages();
    AverageAdd(true, X(), 1.0);
    AverageAdd(false, Y(), 1.0);
    if (buf_.avg_kx != nullptr && params_.linesearch_rule != PDLP_MALITSKY_POCK_LINESEARCH_RULE) {  // the average IS the current iterate
      D.CopyD2D(buf_.avg_kx, Kx(), P.m());
      D.CopyD2D(buf_.avg_kty, Kty(), P.n());
      avg_products_ok_ = true;
    }
    hs_.avg_weight_sum = 1.0;
    hs_.avg_num_terms = 1;
  }
  void SetCurrentPrimalAndDualProducts() {  // pdhg.cc:1961-1974
    P.Kx(X(), Kx());  // (the reference keeps K x only under Malitsky-Pock; here the dual kernel maintains it for every rule)
    P.KTy(Y(), Kty());
  }
  double DistanceTraveledFromLastStart(const double* x, const double* y) {  // pdhg.cc:1998-2007
    double d[2];
    D.DistancesSq(x, x0_, P.n(), y, y0_, P.m(), d);
    return std::sqrt((0.5 * hs_.primal_weight) * d[0] + (0.5 / hs_.primal_weight) * d[1]);
  }
  LocalizedBounds BoundsAt(const double* x, const double* y, const double* kx, const double* kty) {
    // radius = DistanceTraveledFromLastStart(x, y) (pdhg.cc:1998-2007), computed inside the same launch
    double out[4], dist[2];
    P.ComputeLocalizedLagrangianBounds(x, y, hs_.primal_weight, /*radius=*/-1.0, kx, kty, params_.use_diagonal_qp_trust_region_solver != 0,
                                       params_.diagonal_qp_trust_region_solver_tolerance, out, x0_, y0_, dist);
    // ||x - x0||^2, ||y - y0||^2 come out of the same launch: the primal weight update of a restart reuses them
    if (x == X() && y == Y()) { dist_cur_[0] = dist[0]; dist_cur_[1] = dist[1]; dist_cur_ok_ = dist[0] >= 0.0; }
    else if (x == buf_.avg_x && y == buf_.avg_y) { dist_avg_[0] = dist[0]; dist_avg_[1] = dist[1]; dist_avg_ok_ = dist[0] >= 0.0; }
    return {out[0], out[1], out[2], out[3]};
  }
  LocalizedBounds ComputeLocalizedBoundsAtCurrent() {  // pdhg.cc:2009-2021
    return BoundsAt(X(), Y(), CachedKx(X()), Kty());
  }
  LocalizedBounds ComputeLocalizedBoundsAtAverage() {  // :2023-2039
    const double* ax = PrimalAverage();
    const double* ay = DualAverage();
    return BoundsAt(ax, ay, CachedKx(ax), CachedKty(ay));
  }
  // Both points of the restart test (pdhg.cc:2109-2170 evaluates them back to back): concurrently on one GPU.
  void ComputeLocalizedBoundsAtAverageAndCurrent(LocalizedBounds& avg, LocalizedBounds& cur) {
    const double* ax = PrimalAverage();
    const double* ay = DualAverage();
    const double* const xs[2] = {ax, X()};
    const double* const ys[2] = {ay, Y()};
    const double* const kxs[2] = {CachedKx(ax), CachedKx(X())};
    const double* const ktys[2] = {CachedKty(ay), Kty()};
    double out[2][4], dist[2][2];
    if (kxs[0] != nullptr && ktys[0] != nullptr &&
        P.ComputeLocalizedLagrangianBoundsPair(xs, ys, kxs, ktys, hs_.primal_weight, params_.use_diagonal_qp_trust_region_solver != 0, x0_, y0_, out, dist)) {
      avg = {out[0][0], out[0][1], out[0][2], out[0][3]};
      cur = {out[1][0], out[1][1], out[1][2], out[1][3]};
      if (ax == buf_.avg_x && ay == buf_.avg_y) { dist_avg_[0] = dist[0][0]; dist_avg_[1] = dist[0][1]; dist_avg_ok_ = dist[0][0] >= 0.0; }
      dist_cur_[0] = dist[1][0]; dist_cur_[1] = dist[1][1]; dist_cur_ok_ = dist[1][0] >= 0.0;
      return;
    }
    avg = ComputeLocalizedBoundsAtAverage();
    cur = ComputeLocalizedBoundsAtCurrent();
  }
  static bool AverageHasBetterPotential(const LocalizedBounds& avg, const LocalizedBounds& cur) {  // :2041-2048
    return BoundGap(avg) / Sq(avg.radius) < BoundGap(cur) / Sq(cur.radius);
  }
  bool ShouldDoAdaptiveRestartHeuristic(double gap) const {  // :2058-2072
    const double ratio = gap / normalized_gap_at_last_restart_;
    if (ratio < params_.sufficient_reduction_for_restart) return true;
    return ratio < params_.necessary_reduction_for_restart && gap > normalized_gap_at_last_trial_;
  }
  int DetermineDistanceBasedRestartChoice();
  int ChooseRestartToApply(bool is_major_iteration);
  double ComputeNewPrimalWeight();
  void ApplyRestartChoice(int restart);
  PdlpIterationStats CreateSimpleIterationStats(int restart_used) const;
  std::optional<SolverResultCpp> MajorIterationAndTerminationCheck(bool force_numerical, const volatile int32_t* interrupt, SolveLogCpp& log);
  std::optional<ReasonAndType> UpdateIterationStatsAndCheckTermination(bool force_numerical, bool interrupted,
                                                                       const PdlpIterationStats& full_stats, PdlpIterationStats& stats);
  // The interrupt flag as every rank must see it (rank 0 decides on a row-sharded solve).
  bool Interrupted(const volatile int32_t* interrupt) {
    const bool local = interrupt != nullptr && *interrupt != 0;
    if (!P.sharded()) return local;
    if (!interrupt_polled_) return false;  // no rank was given a flag (agreed on in Advance): nothing to exchange
    return D.RootValue(local ? 1.0 : 0.0) != 0.0;
  }
  void ConvergenceAndInfeasibility(const double* x, const double* y, const double* kty_or_null, int type, PdlpConvergenceInformation* conv,
                                   PdlpInfeasibilityInformation* infeas);
  void AddPointMetadata(const double* x, const double* y, int type, PdlpIterationStats& stats);
  SolverResultCpp PickSolutionAndConstructSolverResult(const double* avg_x, const double* avg_y, const PdlpIterationStats& stats, int reason,
                                                       int output_type, SolveLogCpp log);
  void MaterializeDeltas();
  // The iterate difference is only read by termination checks and overwritten by restarts to the
  // average: it is materialised when one of the two happens, not after every chunk of steps.
  bool delta_pending_ = false;
  bool prev_slices_gathered_ = true;  // x[prev] is whole on every rank
  void EnsureDeltas() { if (delta_pending_) { MaterializeDeltas(); delta_pending_ = false; } }
  int NextCheckpoint(int k) const;
  // Iterations between two polls of the time limit / interrupt flag (the reference polls every
  // iteration): about 50 ms of measured iteration time, at least one and at most 32 iterations.
  int PollStride() const {
    if (iterations_completed_ <= 0 || !(device_time_sec_ > 0.0)) return 32;
    const double seconds_per_iteration = device_time_sec_ / iterations_completed_;
    return static_cast<int>(std::max(1.0, std::min(32.0, 0.05 / seconds_per_iteration)));
  }
  Outcome RunDeviceSteps(int k, const volatile int32_t* interrupt);
  Outcome TakeMalitskyPockStep();
  void LogQuadraticProgramStats(const PdlpQuadraticProgramStats& s) const;

  DeviceProblem& P;
  Device& D;
  const PdlpParams params_;
  const Logger& logger_;
  StatsCallback callback_;
  PdlpBoundNorms original_bound_norms_{};
  double *dc_ = nullptr, *dr_ = nullptr;  // col / row scaling vectors
  Device::StepBuffers buf_;
  StepState hs_{};
  double *x0_ = nullptr, *y0_ = nullptr;            // last restart point
  int state_slot_ = 0;  // which of the two device state slots is current
  // Products of the average that a major iteration looks at more than once (restart test,
  // termination check, restart): K x / K^T y. Valid only until the averages change. (K x and
  // K^T y of the current iterate are iterate buffers of the step loop.)
  double *pc_kx_avg_ = nullptr, *pc_kty_avg_ = nullptr;
  bool pc_kx_avg_ok_ = false, pc_kty_avg_ok_ = false;
  // buf_.avg_kx / buf_.avg_kty (row-sharded peer exchange) hold K x / K^T y of the current averages
  bool avg_products_ok_ = false;
  // squared distances of the current / average point to the last restart point, as left by BoundsAt
  double dist_cur_[2] = {0, 0}, dist_avg_[2] = {0, 0};
  bool dist_cur_ok_ = false, dist_avg_ok_ = false;
  void InvalidateProducts(bool current, bool average) {
    if (current) dist_cur_ok_ = false;
    if (average) pc_kx_avg_ok_ = pc_kty_avg_ok_ = dist_avg_ok_ = false;
  }
  const double* CachedKx(const double* x) {  // nullptr: not one of the cached points
    if (x == X()) return Kx();
    if (x == buf_.avg_x) {
      if (avg_products_ok_ && buf_.avg_kx != nullptr) return buf_.avg_kx;
      if (!pc_kx_avg_ok_) { P.Kx(x, pc_kx_avg_); pc_kx_avg_ok_ = true; }
      return pc_kx_avg_;
    }
    return nullptr;
  }
  const double* CachedKty(const double* y) {
    if (y == Y()) return Kty();
    if (y == buf_.avg_y) {
      if (avg_products_ok_ && buf_.avg_kty != nullptr) return buf_.avg_kty;
      if (!pc_kty_avg_ok_) { P.KTy(y, pc_kty_avg_); pc_kty_avg_ok_ = true; }
      return pc_kty_avg_;
    }
    return nullptr;
  }
  double *delta_x_ = nullptr, *delta_y_ = nullptr;   // materialised iterate difference
  bool have_delta_ = false;
  double avg_x_weight_ = 0, avg_y_weight_ = 0;
  int avg_x_terms_ = 0, avg_y_terms_ = 0;
  double ratio_last_two_step_sizes_ = 1;
  double normalized_gap_at_last_trial_ = kInf, normalized_gap_at_last_restart_ = kInf;
  struct { double distance_moved_last_restart_period = kInf; int length_of_last_restart_period = 1; } distance_info_;
  double preprocessing_time_sec_ = 0;
  WallTimer timer_;
  WallTimer log_clock_;
  double time_of_last_log_ = -kInf;
  int log_counter_ = 0;
  int iterations_completed_ = 0;
  int num_rejected_steps_ = 0;
  double device_time_sec_ = 0;
  // session state
  SolveLogCpp solve_log_;
  bool check_done_ = false;
  bool force_numerical_ = false;
  int target_stop_ = std::numeric_limits<int>::max();
  double device_step_ms_ = 0, device_total_ms_ = 0;
  bool interrupt_polled_ = false;  // some rank was given an interrupt flag
  // ---- feasibility polishing (pdhg.cc:2676-3015) ----------------------------
  static constexpr int kFeasibilityIterationFraction = 8;
  int iteration_type_ = PDLP_ITERATION_TYPE_NORMAL;  // what this object's loop is: the main solve or a polishing phase
  bool nested_ = false;                               // a polishing phase: shares the parent's scaling vectors
  int next_feasibility_polishing_iteration_ = 100;
  PdlpIterationStats work_from_feasibility_polishing_{};
  double *polish_x_ = nullptr, *polish_y_ = nullptr;  // working-space result of the primal / dual phase
  void AllocateIterates();
  void InitNested(const DeviceSolve& parent, int iteration_type, const double* x_start, const double* y_start);
  PdlpIterationStats TotalWorkSoFar() const {
    return AddWorkStats(CreateSimpleIterationStats(PDLP_RESTART_CHOICE_NO_RESTART), WorkFromFeasibilityPolishing(solve_log_));
  }
  std::optional<SolverResultCpp> TryFeasibilityPolishing(int iteration_limit, const volatile int32_t* interrupt);
  SolveLogCpp RunPolishingPhase(bool primal_phase, const double* start, int iteration_limit, const volatile int32_t* interrupt);
  SolverResultCpp ConstructSolverResultFromPolished(const PdlpIterationStats& stats, int reason);
  double phase_s_[6] = {0, 0, 0, 0, 0, 0};  // PDLP_B200_TRACE=1: host wall time per phase
};

int DeviceSolve::DetermineDistanceBasedRestartChoice() {  // pdhg.cc:2074-2107
  if (avg_x_terms_ == 0) return PDLP_RESTART_CHOICE_NO_RESTART;
  if (distance_info_.length_of_last_restart_period == 0) return PDLP_RESTART_CHOICE_RESTART_TO_AVERAGE;
  const int period = avg_x_terms_;
  const double dist_avg = DistanceTraveledFromLastStart(buf_.avg_x, buf_.avg_y);
  if ((dist_avg / period) < params_.sufficient_reduction_for_restart *
                                (distance_info_.distance_moved_last_restart_period / distance_info_.length_of_last_restart_period)) {
    LocalizedBounds avg, cur;
    ComputeLocalizedBoundsAtAverageAndCurrent(avg, cur);
    return AverageHasBetterPotential(avg, cur) ? PDLP_RESTART_CHOICE_RESTART_TO_AVERAGE : PDLP_RESTART_CHOICE_WEIGHTED_AVERAGE_RESET;
  }
  return PDLP_RESTART_CHOICE_NO_RESTART;
}

int DeviceSolve::ChooseRestartToApply(bool is_major) {  // pdhg.cc:2109-2170
  if (!PrimalAvgHasWeight() && !DualAvgHasWeight()) return PDLP_RESTART_CHOICE_NO_RESTART;
  const int restart_length = avg_x_terms_;
  if (restart_length >= iterations_completed_ / 2 && params_.restart_strategy == PDLP_ADAPTIVE_HEURISTIC) {
    LocalizedBounds avg, cur;
    ComputeLocalizedBoundsAtAverageAndCurrent(avg, cur);
    return AverageHasBetterPotential(avg, cur) ? PDLP_RESTART_CHOICE_RESTART_TO_AVERAGE : PDLP_RESTART_CHOICE_WEIGHTED_AVERAGE_RESET;
  }
  if (!is_major) return PDLP_RESTART_CHOICE_NO_RESTART;
  switch (params_.restart_strategy) {
    case PDLP_NO_RESTARTS: return PDLP_RESTART_CHOICE_WEIGHTED_AVERAGE_RESET;
    case PDLP_EVERY_MAJOR_ITERATION: return PDLP_RESTART_CHOICE_RESTART_TO_AVERAGE;
    case PDLP_ADAPTIVE_HEURISTIC: {
      LocalizedBounds avg, cur;
      ComputeLocalizedBoundsAtAverageAndCurrent(avg, cur);
      double gap;
      int choice;
      if (AverageHasBetterPotential(avg, cur)) { gap = BoundGap(avg) / avg.radius; choice = PDLP_RESTART_CHOICE_RESTART_TO_AVERAGE; }
      else { gap = BoundGap(cur) / cur.radius; choice = PDLP_RESTART_CHOICE_WEIGHTED_AVERAGE_RESET; }
      if (ShouldDoAdaptiveRestartHeuristic(gap)) return choice;
      normalized_gap_at_last_trial_ = gap;
      return PDLP_RESTART_CHOICE_NO_RESTART;
    }
    case PDLP_ADAPTIVE_DISTANCE_BASED: return DetermineDistanceBasedRestartChoice();
    default: return PDLP_RESTART_CHOICE_UNSPECIFIED;
  }
}

double DeviceSolve::ComputeNewPrimalWeight() {  // pdhg.cc:2188-2214
  double d[2];
  if (dist_cur_ok_) { d[0] = dist_cur_[0]; d[1] = dist_cur_[1]; }
  else D.DistancesSq(X(), x0_, P.n(), Y(), y0_, P.m(), d);
  const double primal_distance = std::sqrt(d[0]), dual_distance = std::sqrt(d[1]);
  constexpr double kNonzeroTol = 1.0e-10;
  if (primal_distance <= kNonzeroTol || primal_distance >= 1.0 / kNonzeroTol || dual_distance <= kNonzeroTol || dual_distance >= 1.0 / kNonzeroTol)
    return hs_.primal_weight;
  const double s = params_.primal_weight_update_smoothing;
  const double w = std::exp(s * std::log(dual_distance / primal_distance) + (1.0 - s) * std::log(hs_.primal_weight));
  if (params_.verbosity_level >= 4) logger_.Log(Fmt("New computed primal weight is %g at iteration %d", w, iterations_completed_));
  return w;
}

void DeviceSolve::ApplyRestartChoice(int restart) {  // pdhg.cc:2246-2296
  switch (restart) {
    case PDLP_RESTART_CHOICE_UNSPECIFIED:
    case PDLP_RESTART_CHOICE_NO_RESTART: return;
    case PDLP_RESTART_CHOICE_WEIGHTED_AVERAGE_RESET:
      if (params_.verbosity_level >= 4) logger_.Log(Fmt("Restarted to current on iteration %d after %d iterations", iterations_completed_, avg_x_terms_));
      break;
    case PDLP_RESTART_CHOICE_RESTART_TO_AVERAGE:
      if (params_.verbosity_level >= 4) logger_.Log(Fmt("Restarted to average on iteration %d after %d iterations", iterations_completed_, avg_x_terms_));
      EnsureDeltas();  // x[cur] - x[prev] of the last accepted step, before x[cur] is overwritten
      D.CopyD2D(X(), buf_.avg_x, P.n());
      D.CopyD2D(Y(), buf_.avg_y, P.m());
      dist_cur_[0] = dist_avg_[0]; dist_cur_[1] = dist_avg_[1]; dist_cur_ok_ = dist_avg_ok_;
      // the new current iterate is the average: its products are the average's
      if (avg_products_ok_ && buf_.avg_kty != nullptr) { D.CopyD2D(Kty(), buf_.avg_kty, P.n()); D.CopyD2D(Kx(), buf_.avg_kx, P.m()); }
      else {
        if (pc_kty_avg_ok_) D.CopyD2D(Kty(), pc_kty_avg_, P.n()); else P.KTy(Y(), Kty());
        if (pc_kx_avg_ok_) D.CopyD2D(Kx(), pc_kx_avg_, P.m()); else P.Kx(X(), Kx());
      }

      break;
  }
  hs_.primal_weight = ComputeNewPrimalWeight();
  ratio_last_two_step_sizes_ = 1;
  if (params_.restart_strategy == PDLP_ADAPTIVE_HEURISTIC) {
    const LocalizedBounds b = ComputeLocalizedBoundsAtCurrent();
    normalized_gap_at_last_restart_ = BoundGap(b) / b.radius;
    normalized_gap_at_last_trial_ = kInf;
  } else if (params_.restart_strategy == PDLP_ADAPTIVE_DISTANCE_BASED) {
    distance_info_.distance_moved_last_restart_period = DistanceTraveledFromLastStart(X(), Y());
    distance_info_.length_of_last_restart_period = avg_x_terms_;
  }
  ClearAverages();
  D.CopyD2D(x0_, X(), P.n());
  D.CopyD2D(y0_, Y(), P.m());
  dist_cur_ok_ = dist_avg_ok_ = false;  // the restart point moved
}

PdlpIterationStats DeviceSolve::CreateSimpleIterationStats(int restart_used) const {  // pdhg.cc:1976-1996
  PdlpIterationStats s;
  std::memset(&s, 0, sizeof(s));
  const double per_rejected = params_.linesearch_rule == PDLP_MALITSKY_POCK_LINESEARCH_RULE ? 0.5 : 1.0;
  s.iteration_number = iterations_completed_;
  s.cumulative_rejected_steps = num_rejected_steps_;
  s.cumulative_kkt_matrix_passes = iterations_completed_ + per_rejected * num_rejected_steps_;
  s.cumulative_time_sec = preprocessing_time_sec_ + timer_.Get();
  s.restart_used = restart_used;
  s.step_size = hs_.step_size;
  s.primal_weight = hs_.primal_weight;
  return s;
}

// pdhg.cc:1655-1724 (no presolve)
void DeviceSolve::ConvergenceAndInfeasibility(const double* x, const double* y, const double* kty_or_null, int type, PdlpConvergenceInformation* conv,
                                              PdlpInfeasibilityInformation* infeas) {
  const DetailedCriteria oc = EffectiveOptimalityCriteria(params_.termination_criteria);
  const bool har = params_.handle_some_primal_gradients_on_finite_bounds_as_residuals != 0;
  if (kty_or_null == nullptr) kty_or_null = CachedKty(y);
  if (conv != nullptr)
    *conv = P.ComputeConvergenceInformation(har, dc_, dr_, x, y, kty_or_null, EpsilonRatio(oc.primal_abs, oc.primal_rel),
                                            EpsilonRatio(oc.dual_abs, oc.dual_rel), type, CachedKx(x));
  if (infeas != nullptr) {
    double* primal_copy = P.tmp_n(3);
    D.CopyD2D(primal_copy, x, P.n());
    D.ClampPrimal(primal_copy, P.lv(), P.uv(), /*feasibility_bounds=*/true, P.n());
    if (type == PDLP_POINT_TYPE_ITERATE_DIFFERENCE) {
      double* dual_copy = P.tmp_m(3);
      D.CopyD2D(dual_copy, y, P.m());
      D.ClampDual(dual_copy, P.lc(), P.uc(), P.m());
      *infeas = P.ComputeInfeasibilityInformation(har, dc_, dr_, primal_copy, dual_copy, x, nullptr, type);
    } else {
      // the dual ray is y itself: reuse K^T y when the caller has it
      const double* kty = kty_or_null;
      if (kty == nullptr && conv != nullptr) kty = P.tmp_n(0);  // left there by ComputeConvergenceInformation
      *infeas = P.ComputeInfeasibilityInformation(har, dc_, dr_, primal_copy, y, x, kty, type);
    }
  }
}

void DeviceSolve::AddPointMetadata(const double* x, const double* y, int type, PdlpIterationStats& stats) {  // pdhg.cc:1547-1565
  PdlpPointMetadata md;
  std::memset(&md, 0, sizeof(md));
  md.point_type = type;
  md.num_random_projections = std::min<int>(params_.num_random_projection_seeds, PDLP_MAX_RANDOM_PROJECTION_SEEDS);
  for (int k = 0; k < md.num_random_projections; ++k) {
    const uint32_t seed = static_cast<uint32_t>(params_.random_projection_seeds[k]);
    md.random_primal_projections[k] = D.RandomProjection(x, P.n(), seed, 0);
    md.random_dual_projections[k] = D.RandomProjection(y, P.m(), seed, 1, P.sharded(), P.row_begin());
  }
  if (type != PDLP_POINT_TYPE_ITERATE_DIFFERENCE) {  // SetActiveSetInformation, pdhg.cc:1476-1545
    int64_t pc[2], dcnt[2];
    D.BeginBatch();  // both counts in one host round trip
    const int op = D.ActiveSetPrimalLaunch(x, x0_, P.lv(), P.uv(), P.n());
    const int od = D.ActiveSetDualLaunch(y, y0_, P.lc(), P.uc(), P.m());
    D.EndBatch();
    D.ReadCounts(op, pc);
    D.ReadCounts(od, dcnt);
    md.has_active_set_information = 1;
    md.active_primal_variable_count = pc[0];
    md.active_primal_variable_change = pc[1];
    md.active_dual_variable_count = dcnt[0];
    md.active_dual_variable_change = dcnt[1];
  }
  stats.point_metadata[stats.num_point_metadata++] = md;
}

void DeviceSolve::MaterializeDeltas() {
  if (!prev_slices_gathered_) {  // (row-sharded peer exchange: inside the loop every rank advances only its slice of x)
    D.GatherPrimalSlices(buf_, -1, hs_.prev);
    prev_slices_gathered_ = true;
  }
  // current_primal_delta_ / current_dual_delta_ of the last accepted step
  // (pdhg.cc:2604-2605) are x[cur]-x[prev] and y[cur]-y[prev]: the third
  // buffer keeps `prev` intact across rejected candidates. Called right after
  // the steps of a chunk so that a later restart (which overwrites x[cur])
  // cannot disturb them.
  D.Sub(delta_x_, buf_.x[hs_.cur], buf_.x[hs_.prev], P.n());
  D.Sub(delta_y_, buf_.y[hs_.cur], buf_.y[hs_.prev], P.m());
  have_delta_ = true;
}

// pdhg.cc:1567-1653
std::optional<ReasonAndType> DeviceSolve::UpdateIterationStatsAndCheckTermination(bool force_numerical, bool interrupted,
                                                                                  const PdlpIterationStats& full_stats, PdlpIterationStats& stats) {
  EnsureDeltas();
  ConvergenceAndInfeasibility(X(), Y(), Kty(), PDLP_POINT_TYPE_CURRENT_ITERATE, &stats.convergence_information[stats.num_convergence_information],
                              &stats.infeasibility_information[stats.num_infeasibility_information]);
  stats.num_convergence_information++;
  stats.num_infeasibility_information++;
  AddPointMetadata(X(), Y(), PDLP_POINT_TYPE_CURRENT_ITERATE, stats);
  if (PrimalAvgHasWeight() && DualAvgHasWeight()) {
    ConvergenceAndInfeasibility(buf_.avg_x, buf_.avg_y, nullptr, PDLP_POINT_TYPE_AVERAGE_ITERATE,
                                &stats.convergence_information[stats.num_convergence_information],
                                &stats.infeasibility_information[stats.num_infeasibility_information]);
    stats.num_convergence_information++;
    stats.num_infeasibility_information++;
    AddPointMetadata(buf_.avg_x, buf_.avg_y, PDLP_POINT_TYPE_AVERAGE_ITERATE, stats);
  }
  if (have_delta_) {
    ConvergenceAndInfeasibility(delta_x_, delta_y_, nullptr, PDLP_POINT_TYPE_ITERATE_DIFFERENCE, nullptr,
                                &stats.infeasibility_information[stats.num_infeasibility_information]);
    stats.num_infeasibility_information++;
    AddPointMetadata(delta_x_, delta_y_, PDLP_POINT_TYPE_ITERATE_DIFFERENCE, stats);
  }
  constexpr int kLogEvery = 15;
  const double now = log_clock_.Get();
  if (params_.verbosity_level >= 2 && (params_.log_interval_seconds == 0.0 || now - time_of_last_log_ >= params_.log_interval_seconds)) {
    if (log_counter_ == 0) LogIterationStatsHeader(params_.verbosity_level, logger_);
    LogIterationStats(params_.verbosity_level, stats, params_.termination_criteria, original_bound_norms_, PDLP_POINT_TYPE_AVERAGE_ITERATE, logger_);
    if (params_.verbosity_level >= 4 && GetConvergenceInformation(stats, PDLP_POINT_TYPE_AVERAGE_ITERATE) != nullptr)
      LogIterationStats(params_.verbosity_level, stats, params_.termination_criteria, original_bound_norms_, PDLP_POINT_TYPE_CURRENT_ITERATE, logger_);
    time_of_last_log_ = now;
    if (++log_counter_ >= kLogEvery) log_counter_ = 0;
  }
  if (callback_) {
    PdlpIterationCallbackInfo info{iteration_type_, &params_.termination_criteria, &stats, original_bound_norms_};
    callback_(info);
  }
  if (const auto t = CheckIterateTerminationCriteria(params_.termination_criteria, stats, original_bound_norms_, force_numerical); t.has_value()) return t;
  return CheckSimpleTerminationCriteria(params_.termination_criteria, full_stats, interrupted);
}

// pdhg.cc:2216-2244 + 329-342. Downloads the chosen point (still scaled).
SolverResultCpp DeviceSolve::PickSolutionAndConstructSolverResult(const double* avg_x, const double* avg_y, const PdlpIterationStats& stats, int reason,
                                                                  int output_type, SolveLogCpp log) {
  const double *px = avg_x, *py = avg_y;
  switch (output_type) {
    case PDLP_POINT_TYPE_CURRENT_ITERATE: px = X(); py = Y(); break;
    case PDLP_POINT_TYPE_ITERATE_DIFFERENCE: px = delta_x_; py = delta_y_; break;
    case PDLP_POINT_TYPE_AVERAGE_ITERATE:
    case PDLP_POINT_TYPE_PRESOLVER_SOLUTION: break;
    default: output_type = PDLP_POINT_TYPE_AVERAGE_ITERATE; break;
  }
  // keep the chosen point on the device in tmp vectors for the unscaling step
  D.CopyD2D(P.tmp_n(1), px, P.n());
  D.CopyD2D(P.tmp_m(1), py, P.m());
  log.iteration_count = stats.iteration_number;
  log.termination_reason = reason;
  log.solution_type = output_type;
  log.solve_time_sec = stats.cumulative_time_sec;
  log.solution_stats = stats;
  log.has_solution_stats = true;
  SolverResultCpp r;
  r.solve_log = std::move(log);
  return r;
}

// pdhg.cc:1728-1818 (no presolve); the chosen scaled point is in tmp_n(1)/tmp_m(1).
SolverResultCpp DeviceSolve::ConstructOriginalSolverResult(SolverResultCpp result) {
  double* x = P.tmp_n(1);
  double* y = P.tmp_m(1);
  const int reason = result.solve_log.termination_reason;
  const bool use_zero_primal_objective = reason == PDLP_TERMINATION_REASON_PRIMAL_INFEASIBLE;
  if (reason == PDLP_TERMINATION_REASON_DUAL_INFEASIBLE) D.ClampPrimal(x, P.lv(), P.uv(), true, P.n());
  if (reason == PDLP_TERMINATION_REASON_PRIMAL_INFEASIBLE) D.ClampDual(y, P.lc(), P.uc(), P.m());
  double* rc = P.tmp_n(2);
  P.ReducedCosts(x, y, use_zero_primal_objective, rc);
  D.Mul(x, dc_, P.n());
  D.Mul(y, dr_, P.m());
  D.Div(rc, dc_, P.n());
  result.primal_solution.resize(P.n());
  result.dual_solution.resize(P.m_global());
  result.reduced_costs.resize(P.n());
  P.DownloadPrimal(result.primal_solution.data(), x);
  P.DownloadDual(result.dual_solution.data(), y);
  P.DownloadPrimal(result.reduced_costs.data(), rc);
  if (callback_) {
    const int termination_type = result.solve_log.solution_type == PDLP_POINT_TYPE_FEASIBILITY_POLISHING_SOLUTION
                                     ? PDLP_ITERATION_TYPE_FEASIBILITY_POLISHING_TERMINATION
                                     : (result.solve_log.solution_type == PDLP_POINT_TYPE_PRESOLVER_SOLUTION ? PDLP_ITERATION_TYPE_PRESOLVE_TERMINATION
                                                                                                            : PDLP_ITERATION_TYPE_NORMAL_TERMINATION);
    PdlpIterationCallbackInfo info{termination_type, &params_.termination_criteria, &result.solve_log.solution_stats, original_bound_norms_};
    callback_(info);
  }
  if (params_.verbosity_level >= 1) {
    logger_.Log(Fmt("Termination reason: %d", result.solve_log.termination_reason));
    logger_.Log(Fmt("Solution point type: %d", result.solve_log.solution_type));
    logger_.Log("Final solution stats:");
    LogIterationStatsHeader(params_.verbosity_level, logger_);
    LogIterationStats(params_.verbosity_level, result.solve_log.solution_stats, params_.termination_criteria, original_bound_norms_,
                      result.solve_log.solution_type, logger_);
    const PdlpConvergenceInformation* ci = GetConvergenceInformation(result.solve_log.solution_stats, result.solve_log.solution_type);
    if (ci != nullptr && std::isfinite(ci->corrected_dual_objective)) logger_.Log(Fmt("Dual objective after infeasibility correction: %g", ci->corrected_dual_objective));
  }
  return result;
}

// pdhg.cc:2360-2435
std::optional<SolverResultCpp> DeviceSolve::MajorIterationAndTerminationCheck(bool force_numerical, const volatile int32_t* interrupt, SolveLogCpp& log) {
  const int cycle = iterations_completed_ % params_.major_iteration_frequency;
  const bool is_major = cycle == 0 && iterations_completed_ > 0;
  WallTimer phase;
  const int restart = force_numerical ? PDLP_RESTART_CHOICE_NO_RESTART : ChooseRestartToApply(is_major);
  phase_s_[0] += phase.Get();
  PdlpIterationStats stats = CreateSimpleIterationStats(restart);
  // one clock decides the time limit (a host round trip through the communicator: only when there is a limit)
  if (P.sharded() && std::isfinite(params_.termination_criteria.time_sec_limit)) stats.cumulative_time_sec = D.RootValue(stats.cumulative_time_sec);
  const PdlpIterationStats full_work_stats = AddWorkStats(stats, work_from_feasibility_polishing_);
  const bool interrupted = Interrupted(interrupt);
  const auto simple = CheckSimpleTerminationCriteria(params_.termination_criteria, full_work_stats, interrupted);
  const bool check_termination = cycle % params_.termination_check_frequency == 0 || simple.has_value() || force_numerical;
  if (check_termination) {
    const double* avg_x = PrimalAverage();
    const double* avg_y = DualAverage();
    phase.Start();
    const auto maybe = UpdateIterationStatsAndCheckTermination(force_numerical, interrupted, full_work_stats, stats);
    phase_s_[1] += phase.Get();
    if (params_.record_iteration_stats) log.iteration_stats.push_back(stats);
    if (maybe.has_value()) {
      if (iteration_type_ == PDLP_ITERATION_TYPE_NORMAL && DoFeasibilityPolishingAfterLimitsReached(params_, maybe->reason)) {
        auto feasibility_result = TryFeasibilityPolishing(iterations_completed_ / kFeasibilityIterationFraction, interrupt);
        if (feasibility_result.has_value()) return feasibility_result;
      }
      const PdlpIterationStats terminating_full_stats = AddWorkStats(stats, work_from_feasibility_polishing_);
      return PickSolutionAndConstructSolverResult(avg_x, avg_y, terminating_full_stats, maybe->reason, maybe->type, std::move(log));
    }
  } else if (params_.record_iteration_stats) {
    log.iteration_stats.push_back(stats);
  }
  phase.Start();
  ApplyRestartChoice(restart);
  phase_s_[2] += phase.Get();
  return std::nullopt;
}

// pdhg.cc:329-342 for the polished pair: leaves it where ConstructOriginalSolverResult reads it.
SolverResultCpp DeviceSolve::ConstructSolverResultFromPolished(const PdlpIterationStats& stats, int reason) {
  D.CopyD2D(P.tmp_n(1), polish_x_, P.n());
  D.CopyD2D(P.tmp_m(1), polish_y_, P.m());
  SolveLogCpp log = solve_log_;
  log.iteration_count = stats.iteration_number;
  log.termination_reason = reason;
  log.solution_type = PDLP_POINT_TYPE_FEASIBILITY_POLISHING_SOLUTION;
  log.solve_time_sec = stats.cumulative_time_sec;
  log.solution_stats = stats;
  log.has_solution_stats = true;
  SolverResultCpp r;
  r.solve_log = std::move(log);
  return r;
}

// TryPrimalPolishing / TryDualPolishing (pdhg.cc:2888-3015). The primal phase solves the
// working problem with a zero objective from (average primal, 0); the dual phase the
// problem with homogeneous bounds (finite -> 0) from (0, average dual). Only the pointers
// of the device problem are exchanged; the result (working space) is kept in polish_x_ /
// polish_y_. Returns the phase's solve log; appends the details to solve_log_.
SolveLogCpp DeviceSolve::RunPolishingPhase(bool primal_phase, const double* start, int iteration_limit, const volatile int32_t* interrupt) {
  PdlpParams phase_params = params_;
  phase_params.termination_criteria = ReduceWorkLimitsByPreviousWork(params_.termination_criteria, iteration_limit, TotalWorkSoFar(),
                                                                     params_.apply_feasibility_polishing_after_limits_reached != 0);
  if (params_.apply_feasibility_polishing_if_solver_is_interrupted) interrupt = nullptr;
  DetailedCriteria criteria = EffectiveOptimalityCriteria(params_.termination_criteria);
  if (primal_phase) criteria.dual_abs = criteria.dual_rel = kInf; else criteria.primal_abs = criteria.primal_rel = kInf;
  criteria.gap_abs = criteria.gap_rel = kInf;
  SetDetailedCriteria(phase_params.termination_criteria, criteria);
  const int64_t n = P.n(), m = P.m();
  if (polish_x_ == nullptr) { polish_x_ = P.NewPrimal(); polish_y_ = P.NewDual(); }
  double *swap_c = nullptr, *swap_lv = nullptr, *swap_uv = nullptr, *swap_lc = nullptr, *swap_uc = nullptr;
  if (primal_phase) {
    swap_c = P.NewPrimal();
    D.Fill(swap_c, 0.0, n);
    P.SwapObjectiveVector(&swap_c);
  } else {
    swap_lv = P.NewPrimal(); swap_uv = P.NewPrimal(); swap_lc = P.NewDual(); swap_uc = P.NewDual();
    D.MapFiniteValuesToZero(swap_lv, P.lv(), n);
    D.MapFiniteValuesToZero(swap_uv, P.uv(), n);
    D.MapFiniteValuesToZero(swap_lc, P.lc(), m);
    D.MapFiniteValuesToZero(swap_uc, P.uc(), m);
    P.SwapVariableBounds(&swap_lv, &swap_uv);
    P.SwapConstraintBounds(&swap_lc, &swap_uc);
  }
  SolveLogCpp phase_log;
  timer_.Stop();  // the time inside the phase is recorded by its own timer
  {
    DeviceSolve phase(P, phase_params, logger_, callback_);
    phase.InitNested(*this, primal_phase ? PDLP_ITERATION_TYPE_PRIMAL_FEASIBILITY : PDLP_ITERATION_TYPE_DUAL_FEASIBILITY, primal_phase ? start : nullptr,
                     primal_phase ? nullptr : start);
    auto result = phase.Advance(std::numeric_limits<int>::max(), interrupt);
    phase_log = std::move(result->solve_log);
    // the phase's chosen point is in tmp_n(1) / tmp_m(1) (PickSolutionAndConstructSolverResult)
    if (primal_phase) D.CopyD2D(polish_x_, P.tmp_n(1), n); else D.CopyD2D(polish_y_, P.tmp_m(1), m);
    device_time_sec_ += phase.device_time_sec_;
    device_step_ms_ += phase.device_step_ms_;
  }
  timer_.Resume();
  if (primal_phase) {
    P.SwapObjectiveVector(&swap_c);
    D.Free(swap_c);
  } else {
    P.SwapVariableBounds(&swap_lv, &swap_uv);
    P.SwapConstraintBounds(&swap_lc, &swap_uc);
    for (double* v : {swap_lv, swap_uv, swap_lc, swap_uc}) D.Free(v);
  }
  InvalidateProducts(true, true);  // scratch vectors were reused by the phase
  PolishingDetailsCpp d;  // BuildFeasibilityPolishingDetails, pdhg.cc:2684-2700
  d.polishing_phase_type = primal_phase ? PDLP_POLISHING_PHASE_TYPE_PRIMAL_FEASIBILITY : PDLP_POLISHING_PHASE_TYPE_DUAL_FEASIBILITY;
  d.main_iteration_count = iterations_completed_;
  d.params = phase_params;
  d.termination_reason = phase_log.termination_reason;
  d.iteration_count = phase_log.iteration_count;
  d.solve_time_sec = phase_log.solve_time_sec;
  d.solution_stats = phase_log.solution_stats;
  d.solution_type = phase_log.solution_type;
  d.iteration_stats = phase_log.iteration_stats;
  solve_log_.feasibility_polishing_details.push_back(std::move(d));
  return phase_log;
}

// pdhg.cc:2702-2865
std::optional<SolverResultCpp> DeviceSolve::TryFeasibilityPolishing(int iteration_limit, const volatile int32_t* interrupt) {
  const DetailedCriteria optimality_criteria = EffectiveOptimalityCriteria(params_.termination_criteria);
  // copies: the averages must survive the phases (they reuse nothing of this object, but a
  // restart-to-average of the main loop later must still see them)
  const double* average_primal = PrimalAverage();
  const double* average_dual = DualAverage();
  PdlpConvergenceInformation first_convergence_info;
  ConvergenceAndInfeasibility(average_primal, average_dual, nullptr, PDLP_POINT_TYPE_AVERAGE_ITERATE, &first_convergence_info, nullptr);
  auto simple_now = [&] { return CheckSimpleTerminationCriteria(params_.termination_criteria, TotalWorkSoFar(), Interrupted(interrupt)); };
  // The objective gap is usually increased by polishing: do not start while it is still too large.
  if (!ObjectiveGapMet(optimality_criteria, first_convergence_info)) {
    const auto simple = simple_now();
    if (!(simple.has_value() && DoFeasibilityPolishingAfterLimitsReached(params_, simple->reason))) {
      if (params_.verbosity_level >= 2) logger_.Log("Skipping feasibility polishing because the objective gap is too large.");
      return std::nullopt;
    }
  }
  if (params_.verbosity_level >= 2) logger_.Log("Starting primal feasibility polishing");
  const SolveLogCpp primal_log = RunPolishingPhase(true, average_primal, iteration_limit, interrupt);
  if (params_.verbosity_level >= 2) logger_.Log(Fmt("Primal feasibility polishing termination reason: %d", primal_log.termination_reason));
  if (TerminationReasonIsWorkLimit(primal_log.termination_reason)) {
    const auto simple = simple_now();
    if (!(simple.has_value() && DoFeasibilityPolishingAfterLimitsReached(params_, simple->reason))) return std::nullopt;
  } else if (primal_log.termination_reason != PDLP_TERMINATION_REASON_OPTIMAL) {
    logger_.Log(Fmt("WARNING: Primal feasibility polishing terminated with error %d", primal_log.termination_reason));
    return std::nullopt;
  }
  if (params_.verbosity_level >= 2) logger_.Log("Starting dual feasibility polishing");
  const SolveLogCpp dual_log = RunPolishingPhase(false, average_dual, iteration_limit, interrupt);
  if (params_.verbosity_level >= 2) logger_.Log(Fmt("Dual feasibility polishing termination reason: %d", dual_log.termination_reason));
  PdlpIterationStats full_stats = TotalWorkSoFar();
  const auto simple = CheckSimpleTerminationCriteria(params_.termination_criteria, full_stats, Interrupted(interrupt));
  auto add_polished_convergence = [&] {
    ConvergenceAndInfeasibility(polish_x_, polish_y_, nullptr, PDLP_POINT_TYPE_FEASIBILITY_POLISHING_SOLUTION,
                                &full_stats.convergence_information[full_stats.num_convergence_information], nullptr);
    full_stats.num_convergence_information += 1;
  };
  if (TerminationReasonIsWorkLimit(dual_log.termination_reason)) {
    if (simple.has_value() && DoFeasibilityPolishingAfterLimitsReached(params_, simple->reason)) {
      add_polished_convergence();
      return ConstructSolverResultFromPolished(full_stats, simple->reason);
    }
    return std::nullopt;
  } else if (dual_log.termination_reason != PDLP_TERMINATION_REASON_OPTIMAL) {
    logger_.Log(Fmt("WARNING: Dual feasibility polishing terminated with error %d", dual_log.termination_reason));
    return std::nullopt;
  }
  add_polished_convergence();
  if (params_.verbosity_level >= 2) {
    logger_.Log("solution stats for polished solution:");
    LogIterationStatsHeader(params_.verbosity_level, logger_);
    LogIterationStats(params_.verbosity_level, full_stats, params_.termination_criteria, original_bound_norms_, PDLP_POINT_TYPE_FEASIBILITY_POLISHING_SOLUTION,
                      logger_);
  }
  const auto earned = CheckIterateTerminationCriteria(params_.termination_criteria, full_stats, original_bound_norms_, /*force_numerical=*/false);
  if (earned.has_value() || (simple.has_value() && DoFeasibilityPolishingAfterLimitsReached(params_, simple->reason)))
    return ConstructSolverResultFromPolished(full_stats, earned.has_value() ? earned->reason : simple->reason);
  return std::nullopt;
}

// Smallest k' > k at which MajorIterationAndTerminationCheck does something.
int DeviceSolve::NextCheckpoint(int k) const {
  const PdlpTerminationCriteria& tc = params_.termination_criteria;
  if (params_.record_iteration_stats) return k + 1;
  const int major = params_.major_iteration_frequency, check = params_.termination_check_frequency;
  int best = std::numeric_limits<int>::max();
  auto consider = [&](int64_t v) { if (v > k && v < best) best = static_cast<int>(std::min<int64_t>(v, std::numeric_limits<int>::max())); };
  consider((static_cast<int64_t>(k) / major + 1) * major);  // next major iteration (also a termination check)
  {  // next termination check inside the current major cycle
    const int cyc = k % major;
    const int next_cyc = (cyc / check + 1) * check;
    if (next_cyc < major) consider(static_cast<int64_t>(k) - cyc + next_cyc);
  }
  consider(static_cast<int64_t>(tc.iteration_limit) - work_from_feasibility_polishing_.iteration_number);
  consider(target_stop_);
  if (params_.use_feasibility_polishing && iteration_type_ == PDLP_ITERATION_TYPE_NORMAL) consider(next_feasibility_polishing_iteration_);
  if (params_.restart_strategy == PDLP_ADAPTIVE_HEURISTIC) {
    // artificial restart (pdhg.cc:2120-2130): first k' with terms + (k'-k) >= k'/2
    // (under Malitsky-Pock the first accepted step after a restart adds TWO terms to the primal average)
    const int terms = avg_x_terms_ + (params_.linesearch_rule == PDLP_MALITSKY_POCK_LINESEARCH_RULE && !PrimalAvgHasWeight() ? 1 : 0);
    for (int64_t j = 1;; ++j) {
      if (terms + j >= (static_cast<int64_t>(k) + j) / 2) { consider(static_cast<int64_t>(k) + j); break; }
      if (k + j >= best) break;
    }
  }
  // time limit / interrupt are polled at checkpoints: keep them close.
  if (std::isfinite(tc.time_sec_limit)) consider(static_cast<int64_t>(k) + PollStride());
  return best;
}

DeviceSolve::Outcome DeviceSolve::RunDeviceSteps(int k, const volatile int32_t* interrupt) {
  InvalidateProducts(true, true);
  hs_.iterations_completed = k;
  hs_.num_rejected_steps = num_rejected_steps_;
  hs_.inner_iterations = 0;
  hs_.halt = kHaltNone;
  int k_stop = NextCheckpoint(k);
  if (interrupt_polled_) k_stop = std::min(k_stop, k + PollStride());
  hs_.k_stop = k_stop;
  hs_.kkt_pass_limit = params_.termination_criteria.kkt_matrix_pass_limit - work_from_feasibility_polishing_.cumulative_kkt_matrix_passes;
  const bool mp = params_.linesearch_rule == PDLP_MALITSKY_POCK_LINESEARCH_RULE;
  hs_.avg_weight_sum = mp ? avg_y_weight_ : avg_x_weight_;
  hs_.avg_num_terms = mp ? avg_y_terms_ : avg_x_terms_;
  hs_.avg_weight_sum_primal = avg_x_weight_;
  hs_.avg_num_terms_primal = avg_x_terms_;
  hs_.pending_ratio = hs_.pending_ratio_dual = hs_.pending_ratio0 = 0.0;
  if (mp) {  // the trial step of the first attempt (pdhg.cc:2473-2480); later ones are set by the device decision
    hs_.mp_ratio = ratio_last_two_step_sizes_;
    hs_.mp_interpolation = params_.malitsky_pock_step_size_interpolation;
    hs_.mp_downscaling = params_.malitsky_pock_step_size_downscaling_factor;
    hs_.mp_contraction = params_.malitsky_pock_linesearch_contraction_factor;
    hs_.mp_new_tau = (hs_.step_size / hs_.primal_weight) * (1.0 + hs_.mp_interpolation * (std::sqrt(1.0 + hs_.mp_ratio) - 1.0));
    hs_.mp_skip_primal = 0;
  }
  hs_.pow_total = -1.0;  // (nothing cached for this attempt count yet)
  PushState();
  WallTimer t;
  device_step_ms_ += D.TimelineCollectMs(1);  // (the previous chunk's; its events are about to be reused)
  D.TimelineStart(1);
  for (;;) {
    const int remaining = std::max(1, hs_.k_stop - hs_.iterations_completed);
    const int64_t attempts_before = hs_.attempts;
    // exactly the attempts that reach the checkpoint if every step is accepted; rejected
    // steps (rare) are made up by another pass of this loop
    if (mp) D.EnqueueMalitskyPockSteps(buf_, P.rows(), P.cols(), std::min(remaining, 4096), state_slot_);
    else D.EnqueueSteps(buf_, P.rows(), P.cols(), std::min(remaining, 4096), state_slot_);
    state_slot_ = D.DownloadLatestState(hs_, buf_.state, state_slot_);
    if (P.sharded() && D.comm() != nullptr) D.comm()->CheckAsyncError();
    D.CollectStepTimings(hs_.attempts - attempts_before);
    if (hs_.halt != kHaltNone) break;
  }
  phase_s_[3] += t.Get();
  WallTimer phase;
  D.FlushAverages(buf_, state_slot_);
  hs_.pending_ratio = 0.0;  // (all FlushAverages changes in the device state)
  if (hs_.halt == kHaltPeerTimeout) throw std::runtime_error("peer-memory exchange timed out: a rank of the row-sharded solve did not arrive");
  D.GatherPrimalSlices(buf_, hs_.cur, -1);
  prev_slices_gathered_ = false;
  D.TimelineStop(1);  // collected lazily: no host synchronisation here
  device_time_sec_ += t.Get();
  phase_s_[4] += phase.Get();
  phase.Start();
  if (hs_.iterations_completed > k) delta_pending_ = true;
  phase_s_[5] += phase.Get();
  avg_y_weight_ = hs_.avg_weight_sum;
  avg_y_terms_ = hs_.avg_num_terms;
  avg_x_weight_ = mp ? hs_.avg_weight_sum_primal : hs_.avg_weight_sum;
  avg_x_terms_ = mp ? hs_.avg_num_terms_primal : hs_.avg_num_terms;
  if (mp) ratio_last_two_step_sizes_ = hs_.mp_ratio;
  num_rejected_steps_ = hs_.num_rejected_steps;
  iterations_completed_ = hs_.iterations_completed;
  if (hs_.halt == kHaltCheckpoint) return Outcome::kSuccessful;
  // kForceNumericalTermination (pdhg.cc:2562-2587, 2653-2662)
  if (params_.verbosity_level >= 2 && hs_.halt != kHaltInnerLimit)
    logger_.Log(Fmt("Forced numerical termination at iteration %d with primal delta squared norm %g dual delta squared norm %g primal weight %g",
                    iterations_completed_, hs_.last_dx2, hs_.last_dy2, hs_.primal_weight));
  if (hs_.halt == kHaltInnerLimit) logger_.Log(Fmt("WARNING: Inner iteration limit reached at iteration %d", iterations_completed_));
  if (hs_.halt == kHaltZeroMovement || hs_.halt == kHaltInnerLimit) ResetAverageToCurrent();
  // the reference's outer loop still increments iterations_completed_
  iterations_completed_ += 1;
  return Outcome::kForceNumericalTermination;
}

// pdhg.cc:2463-2556, host-driven (one sync per inner step).
DeviceSolve::Outcome DeviceSolve::TakeMalitskyPockStep() {
  InvalidateProducts(true, true);
  Outcome outcome = Outcome::kSuccessful;
  const int64_t n = P.n(), m = P.m();
  const double omega = hs_.primal_weight;
  const double primal_step_size = hs_.step_size / omega;
  double* x_next = buf_.x[hs_.cand];
  double* y_next = buf_.y[hs_.cand];
  double* kty_next = buf_.kty[hs_.cand];
  D.PrimalStep(X(), Kty(), P.c(), P.q(), P.lv(), P.uv(), primal_step_size, x_next, n);
  const double dilating = 1 + (params_.malitsky_pock_step_size_interpolation * (std::sqrt(1 + ratio_last_two_step_sizes_) - 1));
  double new_primal_step_size = primal_step_size * dilating;
  const double downscaling = params_.malitsky_pock_step_size_downscaling_factor;
  const double contraction = params_.malitsky_pock_linesearch_contraction_factor;
  const double dual_weight = omega * omega;
  int inner_iterations = 0;
  double* kx_next = buf_.kx[hs_.cand];
  P.Kx(x_next, kx_next);
  for (bool accepted = false; !accepted; ++inner_iterations) {
    if (inner_iterations >= 60) {
      logger_.Log(Fmt("WARNING: Inner iteration limit reached at iteration %d", iterations_completed_));
      ResetAverageToCurrent();
      outcome = Outcome::kForceNumericalTermination;
      break;
    }
    const double new_ratio = new_primal_step_size / primal_step_size;
    D.DualStepFromProducts(Y(), Kx(), kx_next, P.lc(), P.uc(), dual_weight * new_primal_step_size, new_ratio, y_next, m);
    P.KTy(y_next, kty_next);
    const double delta_dual_norm = std::sqrt(D.SumSqDiff(y_next, Y(), m, P.sharded()));
    const double delta_dual_prod_norm = std::sqrt(D.SumSqDiff(Kty(), kty_next, n));
    if (omega * new_primal_step_size * delta_dual_prod_norm <= contraction * delta_dual_norm) {
      hs_.step_size = new_primal_step_size * omega;
      ratio_last_two_step_sizes_ = new_ratio;
      if (!PrimalAvgHasWeight()) AverageAdd(true, X(), new_primal_step_size * new_ratio);
      const double dx2 = D.SumSqDiff(x_next, X(), n);
      const double dy2 = delta_dual_norm * delta_dual_norm;
      const int old_prev = hs_.prev;
      hs_.prev = hs_.cur;
      hs_.cur = hs_.cand;
      hs_.cand = old_prev;
      MaterializeDeltas();
      AverageAdd(true, X(), new_primal_step_size);
      AverageAdd(false, Y(), new_primal_step_size);
      const double movement = (0.5 * omega * dx2) + (0.5 / omega) * dy2;
      if (movement == 0.0 || movement > 1.0e100) {
        if (params_.verbosity_level >= 2)
          logger_.Log(Fmt("Forced numerical termination at iteration %d with primal delta squared norm %g dual delta squared norm %g primal weight %g",
                          iterations_completed_, dx2, dy2, omega));
        if (movement == 0.0) ResetAverageToCurrent();
        outcome = Outcome::kForceNumericalTermination;
      }
      break;
    } else {
      new_primal_step_size = downscaling * new_primal_step_size;
    }
  }
  num_rejected_steps_ += inner_iterations;
  iterations_completed_ += 1;
  return outcome;
}

// pdhg.cc:3017-3092 (loop part; the start-up part is at the end of Prepare)
std::optional<SolverResultCpp> DeviceSolve::Advance(int target_iterations, const volatile int32_t* interrupt_solve) {
  target_stop_ = target_iterations;
  interrupt_polled_ = D.MaxOverRanks(interrupt_solve != nullptr ? 1.0 : 0.0) != 0.0;
  // (row-sharded Malitsky-Pock solves are still driven from the host, one synchronisation per inner step)
  const bool device_loop = params_.linesearch_rule != PDLP_MALITSKY_POCK_LINESEARCH_RULE || !P.sharded();
  if (!nested_) D.TimelineStart(0);
  std::optional<SolverResultCpp> done;
  for (;;) {
    if (!check_done_) {
      auto maybe = MajorIterationAndTerminationCheck(force_numerical_, interrupt_solve, solve_log_);
      check_done_ = true;
      if (maybe.has_value()) {
        maybe->solve_log.gpu_kernel_launches = D.launches();
        maybe->solve_log.device_iteration_time_sec = device_time_sec_;
        done = std::move(maybe);
        break;
      }
    }
    // pdhg.cc:3056-3070: polishing attempts at iterations 100, 200, 400, ... of the main solve
    if (params_.use_feasibility_polishing && iteration_type_ == PDLP_ITERATION_TYPE_NORMAL &&
        iterations_completed_ >= next_feasibility_polishing_iteration_) {
      auto feasibility_result = TryFeasibilityPolishing(iterations_completed_ / kFeasibilityIterationFraction, interrupt_solve);
      if (feasibility_result.has_value()) {
        feasibility_result->solve_log.gpu_kernel_launches = D.launches();
        feasibility_result->solve_log.device_iteration_time_sec = device_time_sec_;
        done = std::move(feasibility_result);
        break;
      }
      next_feasibility_polishing_iteration_ *= 2;
      work_from_feasibility_polishing_ = WorkFromFeasibilityPolishing(solve_log_);
    }
    if (iterations_completed_ >= target_iterations) break;
    const Outcome outcome = device_loop ? RunDeviceSteps(iterations_completed_, interrupt_solve) : TakeMalitskyPockStep();
    check_done_ = false;
    if (outcome == Outcome::kForceNumericalTermination) force_numerical_ = true;
  }
  device_step_ms_ += D.TimelineCollectMs(1);
  if (!nested_) device_total_ms_ += D.TimelineStopMs(0);
  return done;
}

void DeviceSolve::FillStatus(PdlpSessionStatus* out) const {
  std::memset(out, 0, sizeof(*out));
  out->iterations_completed = iterations_completed_;
  out->num_rejected_steps = num_rejected_steps_;
  out->step_size = hs_.step_size;
  out->primal_weight = hs_.primal_weight;
  out->gpu_kernel_launches = D.launches();
  out->device_step_ms = device_step_ms_;
  out->device_total_ms = device_total_ms_;
  const Device::StepTimings& t = D.step_timings();
  const double n = static_cast<double>(P.n()), m = static_cast<double>(P.m());
  const double qn = P.is_lp() ? 0.0 : 1.0;
  for (int k = 0; k < 4; ++k) { out->kernel_ms[k] = t.ms[k]; out->kernel_samples[k] = t.samples[k]; }
  // Algorithmic bytes per launch (DESIGN.md "Roofline accounting"):
  //  primal step: reads x, c, K^T y, l_v, u_v (+Q), writes x', x~; deferred average R/W.
  //  (peer exchange of a row-sharded solve: a rank advances only its slice of the primal vectors)
  const double n_step = (P.sharded() && P.arena() != nullptr) ? static_cast<double>(P.slice_end() - P.slice_begin()) : n;
  out->kernel_algorithmic_bytes[0] = 8.0 * (9.0 + qn) * n_step;
  //  K x~ + dual epilogue: K-by-rows (8 B value + 4 B index per nonzero, 4 B row offset),
  //  gathers x~ once, reads y, l_c, u_c, K x, writes y', K x'; deferred dual average R/W.
  out->kernel_algorithmic_bytes[1] = 12.0 * static_cast<double>(P.nnz()) + 4.0 * (m + 1) + 8.0 * n + 8.0 * 8.0 * m;
  //  K^T y': K-by-columns, gathers y' once, writes K^T y' (the nonlinearity is a by-product of the dual kernel).
  out->kernel_algorithmic_bytes[2] = 12.0 * static_cast<double>(P.nnz()) + 4.0 * (n + 1) + 8.0 * m + 8.0 * n;
  out->kernel_algorithmic_bytes[3] = 0.0;
}

void DeviceSolve::LogQuadraticProgramStats(const PdlpQuadraticProgramStats& s) const {  // pdhg.cc:1341-1399
  logger_.Log(Fmt("There are %lld variables, %lld constraints, and %lld constraint matrix nonzeros.", (long long)s.num_variables,
                  (long long)s.num_constraints, (long long)s.constraint_matrix_num_nonzeros));
  if (s.constraint_matrix_num_nonzeros > 0) {
    logger_.Log(Fmt("Absolute values of nonzero constraint matrix elements: largest=%f, smallest=%f, avg=%f", s.constraint_matrix_abs_max,
                    s.constraint_matrix_abs_min, s.constraint_matrix_abs_avg));
    logger_.Log(Fmt("Constraint matrix, infinity norm: max(row & col)=%f, min_col=%f, min_row=%f", s.constraint_matrix_abs_max,
                    s.constraint_matrix_col_min_l_inf_norm, s.constraint_matrix_row_min_l_inf_norm));
    logger_.Log(Fmt("Constraint bounds statistics (max absolute value per row): largest=%f, smallest=%f, avg=%f, l2_norm=%f", s.combined_bounds_max,
                    s.combined_bounds_min, s.combined_bounds_avg, s.combined_bounds_l2_norm));
  }
  if (!P.is_lp()) {
    logger_.Log(Fmt("There are %lld nonzero diagonal coefficients in the objective matrix.", (long long)s.objective_matrix_num_nonzeros));
    logger_.Log(Fmt("Absolute values of nonzero objective matrix elements: largest=%f, smallest=%f, avg=%f", s.objective_matrix_abs_max,
                    s.objective_matrix_abs_min, s.objective_matrix_abs_avg));
  }
  logger_.Log(Fmt("Absolute values of objective vector elements: largest=%f, smallest=%f, avg=%f, l2_norm=%f", s.objective_vector_abs_max,
                  s.objective_vector_abs_min, s.objective_vector_abs_avg, s.objective_vector_l2_norm));
  logger_.Log(Fmt("Gaps between variable upper and lower bounds: #finite=%lld of %lld, largest=%f, smallest=%f, avg=%f",
                  (long long)s.variable_bound_gaps_num_finite, (long long)s.num_variables, s.variable_bound_gaps_max, s.variable_bound_gaps_min,
                  s.variable_bound_gaps_avg));
}

SolverResultCpp DeviceSolve::PreprocessAndSolve(std::optional<InitialSolution> initial_solution, const volatile int32_t* interrupt_solve,
                                                const std::string* name) {
  const char* tr = std::getenv("PDLP_B200_TRACE");
  const bool trace = tr != nullptr && tr[0] == '1';
  WallTimer t;
  if (auto err = Prepare(std::move(initial_solution), name); err.has_value()) return std::move(*err);
  if (trace) D.Sync();
  const double t_prepare = t.Get();
  auto result = Advance(std::numeric_limits<int>::max(), interrupt_solve);
  const double t_advance = t.Get();
  SolverResultCpp out = ConstructOriginalSolverResult(std::move(*result));
  if (trace)
    std::fprintf(stderr, "[pdlp_b200 trace] solve wall seconds: prepare (checks, stats, rescaling) %.4f, loop %.4f, result (unscale, reduced costs, download) %.4f\n",
                 t_prepare, t_advance - t_prepare, t.Get() - t_advance);
  return out;
}

// pdhg.cc:1039-1221
void DeviceSolve::AllocateIterates() {
  const int64_t n = P.n(), m = P.m();
  buf_.n = n;
  buf_.m = m;
  for (int k = 0; k < 3; ++k) { buf_.x[k] = P.NewPrimal(); buf_.y[k] = P.NewDual(); buf_.kty[k] = P.NewPrimal(); buf_.kx[k] = P.NewDual(); }
  buf_.x_tilde = P.NewPrimal();
  buf_.avg_x = P.NewPrimal();
  buf_.avg_y = P.NewDual();
  x0_ = P.NewPrimal();
  y0_ = P.NewDual();
  delta_x_ = P.NewPrimal();
  pc_kx_avg_ = P.NewDual();
  pc_kty_avg_ = P.NewPrimal();
  delta_y_ = P.NewDual();
  if (P.sharded() && P.arena() != nullptr) {  // peer exchange: the step kernels maintain K x / K^T y of the average (device_ops.h)
    buf_.avg_kx = P.NewDual();
    buf_.avg_kty = P.NewPrimal();
  }
  buf_.c = P.c(); buf_.q = P.q(); buf_.lv = P.lv(); buf_.uv = P.uv(); buf_.lc = P.lc(); buf_.uc = P.uc();
  buf_.state = D.AllocState();
  buf_.exchange = P.exchange();
  buf_.primal_scatter = P.sharded() ? P.primal_scatter() : nullptr;
  buf_.arena = P.arena();
  buf_.slice_begin = P.slice_begin();
  buf_.slice_end = P.slice_end();
  buf_.slice_stride = P.slice_stride();
  buf_.m_global = P.m_global();
  buf_.row_begin = P.row_begin();
  buf_.cols_slice = P.cols_slice();
  buf_.slice_perm = P.slice_perm();
  std::memset(&hs_, 0, sizeof(hs_));
  hs_.cur = 0; hs_.prev = 1; hs_.cand = 2;

}

// A polishing phase (the nested `Solver` of pdhg.cc:2913-2921 / 2992-2999): same working
// problem, scaling vectors and bound norms as the parent, its own iterates; starts from
// (x_start, y_start) (nullptr = zero) with the parent's step size and primal weight and
// performs the start of Solver::Solve (pdhg.cc:3017-3041).
void DeviceSolve::InitNested(const DeviceSolve& parent, int iteration_type, const double* x_start, const double* y_start) {
  nested_ = true;
  iteration_type_ = iteration_type;
  dc_ = parent.dc_;
  dr_ = parent.dr_;
  original_bound_norms_ = parent.original_bound_norms_;
  AllocateIterates();
  if (x_start != nullptr) D.CopyD2D(X(), x_start, P.n()); else D.Fill(X(), 0.0, P.n());
  if (y_start != nullptr) D.CopyD2D(Y(), y_start, P.m()); else D.Fill(Y(), 0.0, P.m());
  hs_.step_size = parent.hs_.step_size;
  hs_.primal_weight = parent.hs_.primal_weight;
  hs_.rule = params_.linesearch_rule;
  hs_.reduction_exponent = params_.adaptive_step_size_reduction_exponent;
  hs_.growth_exponent = params_.adaptive_step_size_growth_exponent;
  ClearAverages();
  solve_log_ = SolveLogCpp();
  preprocessing_time_sec_ = 0;
  timer_.Start();
  D.CopyD2D(x0_, X(), P.n());
  D.CopyD2D(y0_, Y(), P.m());
  ratio_last_two_step_sizes_ = 1;
  SetCurrentPrimalAndDualProducts();
  force_numerical_ = false;
  check_done_ = false;
  num_rejected_steps_ = 0;
  iterations_completed_ = 0;
}

std::optional<SolverResultCpp> DeviceSolve::Prepare(std::optional<InitialSolution> initial_solution, const std::string* name) {
  WallTimer timer;
  SolveLogCpp solve_log;
  if (params_.verbosity_level >= 1) logger_.Log("Solving with PDLP parameters: (PdlpParams POD)");
  if (name != nullptr) solve_log.instance_name = *name;
  solve_log.params = params_;
  const int64_t n = P.n(), m = P.m();
  P.ReplaceLargeConstraintBoundsWithInfinity(params_.infinite_constraint_bound_threshold);
  if (!P.HasValidBounds())
    return ErrorSolverResult(PDLP_TERMINATION_REASON_INVALID_PROBLEM,
                             "The input problem has invalid bounds (after replacing large constraint bounds with infinity): some variable or "
                             "constraint has lower_bound > upper_bound, lower_bound == inf, or upper_bound == -inf.", logger_);
  if (!P.ObjectiveMatrixIsNonNegative())
    return ErrorSolverResult(PDLP_TERMINATION_REASON_INVALID_PROBLEM,
                             "The objective is not convex (i.e., the objective matrix contains negative or NAN entries).", logger_);
  const double t_checks = timer.Get();
  solve_log.original_stats = P.ComputeStats();
  solve_log.has_original_stats = true;
  const double t_stats0 = timer.Get();
  if (auto r = CheckProblemStats(solve_log.original_stats, P.objective_offset(), params_.presolve_use_glop != 0, logger_); r.has_value()) return std::move(*r);

  AllocateIterates();

  if (initial_solution.has_value()) {  // CheckInitialSolution, pdhg.cc:985-1037
    const double kBig = 1e50;
    auto err = [&](const std::string& msg) { return ErrorSolverResult(PDLP_TERMINATION_REASON_INVALID_INITIAL_SOLUTION, msg, logger_); };
    if (static_cast<int64_t>(initial_solution->primal.size()) != n)
      return err(Fmt("Initial primal solution has size %lld which differs from problem primal size %lld", (long long)initial_solution->primal.size(), (long long)n));
    P.UploadPrimal(X(), initial_solution->primal.data());
    if (std::isnan(std::sqrt(D.SumSq(X(), n)))) return err("Initial primal solution has a NAN.");
    if (const double v = D.LInf(X(), n); v > kBig) return err("Initial primal solution has an entry with absolute value " + G(v) + " which exceeds limit of " + G(kBig));
    if (static_cast<int64_t>(initial_solution->dual.size()) != P.m_global())
      return err(Fmt("Initial dual solution has size %lld which differs from problem dual size %lld", (long long)initial_solution->dual.size(), (long long)P.m_global()));
    P.UploadDual(Y(), initial_solution->dual.data());
    if (std::isnan(std::sqrt(D.SumSq(Y(), m, P.sharded())))) return err("Initial dual solution has a NAN.");
    if (const double v = D.LInf(Y(), m, P.sharded()); v > kBig) return err("Initial dual solution has an entry with absolute value " + G(v) + " which exceeds limit of " + G(kBig));
  } else {
    D.Fill(X(), 0.0, n);
    D.Fill(Y(), 0.0, m);
  }
  original_bound_norms_ = BoundNormsFromProblemStats(solve_log.original_stats);
  if (params_.verbosity_level >= 1) { logger_.Log("Problem stats before rescaling:"); LogQuadraticProgramStats(solve_log.original_stats); }

  D.ClampPrimal(X(), P.lv(), P.uv(), false, n);  // ProjectToPrimalVariableBounds, pdhg.cc:1164
  D.ClampDual(Y(), P.lc(), P.uc(), m);

  // ComputeAndApplyRescaling, pdhg.cc:1325-1339
  const double t_before_rescaling = timer.Get();
  P.ApplyRescaling(params_.l_inf_ruiz_iterations, params_.l2_norm_rescaling != 0, &dr_, &dc_);
  D.Div(X(), dc_, n);
  D.Div(Y(), dr_, m);
  const double t_rescaled = timer.Get();
  solve_log.preprocessed_stats = P.ComputeStats();
  solve_log.has_preprocessed_stats = true;
  if (const char* tr = std::getenv("PDLP_B200_TRACE"); tr != nullptr && tr[0] == '1')
    std::fprintf(stderr, "[pdlp_b200 trace] prepare wall seconds: bound checks %.4f, stats %.4f, allocate + project %.4f, rescaling (enqueue) %.4f, stats after rescaling %.4f\n",
                 t_checks, t_stats0 - t_checks, t_before_rescaling - t_stats0, t_rescaled - t_before_rescaling, timer.Get() - t_rescaled);
  if (params_.verbosity_level >= 1) { logger_.Log("Problem stats after rescaling:"); LogQuadraticProgramStats(solve_log.preprocessed_stats); }

  double step_size;
  if (params_.linesearch_rule == PDLP_CONSTANT_STEP_SIZE_RULE) {
    // EstimateMaximumSingularValueOfConstraintMatrix (sou.cc:559-699): power
    // iteration on K^T K with a counter-based Gaussian start vector.
    double* v = P.tmp_n(0);
    double* w = P.tmp_m(0);
    double* nv = P.tmp_n(1);
    std::vector<double> start(n);
    {
      // deterministic N(0,1) start (the reference seeds std::mt19937(1); its
      // stream is unpinned by its tests beyond the 0.2 relative error target)
      uint64_t s = 0x9E3779B97F4A7C15ull;
      for (int64_t i = 0; i < n; ++i) {
        auto next = [&] { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (s >> 11) * (1.0 / 9007199254740992.0); };
        const double u1 = std::max(next(), 1e-300), u2 = next();
        start[i] = std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
      }
    }
    D.Upload(v, start.data(), n);
    auto normalize = [&](double* vec) { const double nrm = std::sqrt(D.SumSq(vec, n)); if (nrm != 0.0) { D.Fill(P.tmp_n(2), 1.0 / nrm, n); D.Mul(vec, P.tmp_n(2), n); } };
    normalize(v);
    const double desired_relative_error = 0.2, failure_probability = 0.0005;
    const double epsilon = 1.0 - Sq(1.0 - desired_relative_error);
    auto failure = [&](int k) {
      if (k < 2 || epsilon <= 0.0) return 1.0;
      return std::min(0.824, 0.354 / std::sqrt(epsilon * (k - 1))) * std::sqrt(static_cast<double>(n)) * std::pow(1.0 - epsilon, k - 0.5);
    };
    double eigenvalue = 0.0;
    int iters = 0;
    while (failure(iters) > failure_probability) {
      P.Kx(v, w);
      P.KTy(w, nv);
      eigenvalue = D.Dot(v, nv, n);
      D.CopyD2D(v, nv, n);
      ++iters;
      normalize(v);
    }
    const double upper = std::sqrt(eigenvalue) / (1.0 - desired_relative_error);
    step_size = upper > 0.0 ? 1.0 / upper : 1.0;
  } else {
    step_size = 1.0 / std::max(1.0e-20, solve_log.preprocessed_stats.constraint_matrix_abs_max);  // pdhg.cc:1203-1207
  }
  step_size *= params_.initial_step_size_scaling;
  double primal_weight = 1.0;  // InitialPrimalWeight, pdhg.cc:1401-1419
  if (params_.has_initial_primal_weight) primal_weight = params_.initial_primal_weight;
  else if (solve_log.preprocessed_stats.objective_vector_l2_norm > 0.0 && solve_log.preprocessed_stats.combined_bounds_l2_norm > 0.0)
    primal_weight = solve_log.preprocessed_stats.objective_vector_l2_norm / solve_log.preprocessed_stats.combined_bounds_l2_norm;
  hs_.step_size = step_size;
  hs_.primal_weight = primal_weight;
  hs_.rule = params_.linesearch_rule;
  hs_.reduction_exponent = params_.adaptive_step_size_reduction_exponent;
  hs_.growth_exponent = params_.adaptive_step_size_growth_exponent;
  ClearAverages();
  solve_log.preprocessing_time_sec = timer.Get();
  // start of Solver::Solve, pdhg.cc:3017-3041
  solve_log_ = std::move(solve_log);
  preprocessing_time_sec_ = solve_log_.preprocessing_time_sec;
  timer_.Start();
  D.CopyD2D(x0_, X(), P.n());
  D.CopyD2D(y0_, Y(), P.m());
  ratio_last_two_step_sizes_ = 1;
  SetCurrentPrimalAndDualProducts();
  force_numerical_ = false;
  check_done_ = false;
  num_rejected_steps_ = 0;
  iterations_completed_ = 0;
  return std::nullopt;
}

}  // namespace

namespace {
// pdhg.cc:3120-3147
std::optional<SolverResultCpp> ValidateInputs(const PdlpProblemView& view, const PdlpParams& params, const Logger& logger) {
  const std::string perr = ValidateParams(params);
  if (!perr.empty()) return ErrorSolverResult(PDLP_TERMINATION_REASON_INVALID_PARAMETER, "INVALID_ARGUMENT: " + perr, logger);
  const std::string derr = ValidateDimensions(view);
  if (!derr.empty()) return ErrorSolverResult(PDLP_TERMINATION_REASON_INVALID_PROBLEM, "INVALID_ARGUMENT: " + derr, logger);
  if (view.num_variables > 0 && view.col_starts != nullptr && view.col_starts[view.num_variables] != view.num_nonzeros)
    return ErrorSolverResult(PDLP_TERMINATION_REASON_INVALID_PROBLEM, "INVALID_ARGUMENT: col_starts[num_variables] differs from num_nonzeros", logger);
  if (view.objective_scaling_factor == 0) return ErrorSolverResult(PDLP_TERMINATION_REASON_INVALID_PROBLEM, "The objective scaling factor cannot be zero.", logger);
  if (params.use_feasibility_polishing && view.objective_matrix_diagonal != nullptr)
    return ErrorSolverResult(PDLP_TERMINATION_REASON_INVALID_PARAMETER, "use_feasibility_polishing is only implemented for linear programs.", logger);
  if (params.presolve_use_glop)
    return ErrorSolverResult(PDLP_TERMINATION_REASON_INVALID_PARAMETER,
                             "presolve_options.use_glop (glop presolve) is a host-side feature outside this library's scope.", logger);
  if (params.num_random_projection_seeds > PDLP_MAX_RANDOM_PROJECTION_SEEDS)
    return ErrorSolverResult(PDLP_TERMINATION_REASON_INVALID_PARAMETER, "at most 8 random_projection_seeds are supported.", logger);
  return std::nullopt;
}
}  // namespace

SolverResultCpp PrimalDualHybridGradient(const PdlpProblemView& view, const PdlpParams& params, std::optional<InitialSolution> initial_solution,
                                         const volatile int32_t* interrupt_solve, const Logger& logger, StatsCallback callback, int cuda_device,
                                         Comm* comm) {
  const char* tr = std::getenv("PDLP_B200_TRACE");
  const bool trace = tr != nullptr && tr[0] == '1';
  WallTimer t;
  if (auto err = ValidateInputs(view, params, logger); err.has_value()) return std::move(*err);
  const double t_validate = t.Get();
  SolverResultCpp result;
  double t_problem = 0, t_solve = 0;
  try {
    DeviceProblem problem(view, cuda_device, comm);
    t_problem = t.Get();
    DeviceSolve solve(problem, params, logger, std::move(callback));
    const std::string name = view.problem_name != nullptr ? std::string(view.problem_name) : std::string();
    result = solve.PreprocessAndSolve(std::move(initial_solution), interrupt_solve, view.problem_name != nullptr ? &name : nullptr);
    t_solve = t.Get();
  } catch (const CommError& e) {
    // the communicator failed under the solve (ncclCommGetAsyncError): no usable point, the
    // reason is not one of the problem's (pdhg.cc has no analogue; SolveLog's catch-all applies)
    return ErrorSolverResult(PDLP_TERMINATION_REASON_OTHER, std::string("The row-sharded solve was aborted: ") + e.what(), logger);
  }
  if (trace)
    std::fprintf(stderr, "[pdlp_b200 trace] entry point wall seconds: validate %.4f, device problem (upload + SELL build) %.4f, preprocess + solve + result %.4f, teardown %.4f\n",
                 t_validate, t_problem - t_validate, t_solve - t_problem, t.Get() - t_solve);
  return result;
}

// ---- sessions ---------------------------------------------------------------
struct SolveSession::Impl {
  Logger logger;
  std::unique_ptr<DeviceProblem> problem;
  std::unique_ptr<DeviceSolve> solve;
  std::optional<SolverResultCpp> finished;  // scaled-point result or input error
  std::optional<SolverResultCpp> original;  // the result in the caller's space (built by the first Finish)
  bool input_error = false;
};

SolveSession::SolveSession() : impl_(new Impl) {}
SolveSession::~SolveSession() = default;

std::unique_ptr<SolveSession> SolveSession::Create(const PdlpProblemView& view, const PdlpParams& params, std::optional<InitialSolution> initial_solution,
                                                   const Logger& logger, StatsCallback callback, int cuda_device, Comm* comm) {
  std::unique_ptr<SolveSession> s(new SolveSession);
  Impl& im = *s->impl_;
  im.logger = logger;
  if (auto err = ValidateInputs(view, params, im.logger); err.has_value()) {
    im.finished = std::move(*err);
    im.input_error = true;
    return s;
  }
  im.problem.reset(new DeviceProblem(view, cuda_device, comm));
  im.solve.reset(new DeviceSolve(*im.problem, params, im.logger, std::move(callback)));
  const std::string name = view.problem_name != nullptr ? std::string(view.problem_name) : std::string();
  if (auto err = im.solve->Prepare(std::move(initial_solution), view.problem_name != nullptr ? &name : nullptr); err.has_value()) {
    im.finished = std::move(*err);
    im.input_error = true;
  }
  return s;
}

bool SolveSession::Advance(int target_iterations, const volatile int32_t* interrupt_solve) {
  Impl& im = *impl_;
  if (im.finished.has_value()) return true;
  auto r = im.solve->Advance(target_iterations, interrupt_solve);
  if (r.has_value()) im.finished = std::move(*r);
  return im.finished.has_value();
}

void SolveSession::EnableTiming(bool on, int stride) {
  if (impl_->solve) impl_->solve->EnableTiming(on, stride);
}

void SolveSession::Status(PdlpSessionStatus* out) const {
  std::memset(out, 0, sizeof(*out));
  if (impl_->solve) impl_->solve->FillStatus(out);
  out->terminated = impl_->finished.has_value() ? 1 : 0;
  out->termination_reason = impl_->finished.has_value() ? impl_->finished->solve_log.termination_reason : PDLP_TERMINATION_REASON_UNSPECIFIED;
}

SolverResultCpp SolveSession::Finish() {
  Impl& im = *impl_;
  if (!im.finished.has_value()) {
    // Stop like an interrupted solve (pdhg.cc:364-368): the next termination
    // check sees the flag and returns the average iterate.
    const int32_t stop = 1;
    im.solve->ForceRecheck();
    Advance(std::numeric_limits<int>::max(), &stop);
  }
  if (im.input_error) return *im.finished;
  // Unscaling works in place on the device copy of the chosen point and fires the termination
  // callback: done once, later calls return the same result.
  if (!im.original.has_value()) im.original = im.solve->ConstructOriginalSolverResult(*im.finished);
  return *im.original;
}

// ---- termination.h as host-only C entry points (no device needed): the scalar predicates the
// solve loop uses, exported so that the reference's termination_test known answers run against
// this library's own code.
extern "C" {
int32_t pdlp_b200_check_simple_termination_criteria(const PdlpTerminationCriteria* criteria, const PdlpIterationStats* stats,
                                                    const volatile int32_t* interrupt_solve, int32_t* reason, int32_t* point_type) {
  const auto r = CheckSimpleTerminationCriteria(*criteria, *stats, interrupt_solve != nullptr && *interrupt_solve != 0);
  if (!r.has_value()) return 0;
  *reason = r->reason;
  *point_type = r->type;
  return 1;
}
int32_t pdlp_b200_check_iterate_termination_criteria(const PdlpTerminationCriteria* criteria, const PdlpIterationStats* stats,
                                                     const PdlpBoundNorms* bound_norms, int32_t force_numerical_termination, int32_t* reason,
                                                     int32_t* point_type) {
  const auto r = CheckIterateTerminationCriteria(*criteria, *stats, *bound_norms, force_numerical_termination != 0);
  if (!r.has_value()) return 0;
  *reason = r->reason;
  *point_type = r->type;
  return 1;
}
int32_t pdlp_b200_optimality_criteria_met(const PdlpTerminationCriteria* criteria, const PdlpConvergenceInformation* stats,
                                          const PdlpBoundNorms* bound_norms, int32_t* objective_gap_met) {
  const DetailedCriteria oc = EffectiveOptimalityCriteria(*criteria);
  if (objective_gap_met != nullptr) *objective_gap_met = ObjectiveGapMet(oc, *stats) ? 1 : 0;
  return OptimalityCriteriaMet(oc, *stats, criteria->optimality_norm, *bound_norms) ? 1 : 0;
}
void pdlp_b200_effective_optimality_criteria(const PdlpTerminationCriteria* criteria, double out[6]) {
  const DetailedCriteria oc = EffectiveOptimalityCriteria(*criteria);
  out[0] = oc.primal_abs; out[1] = oc.primal_rel; out[2] = oc.dual_abs; out[3] = oc.dual_rel; out[4] = oc.gap_abs; out[5] = oc.gap_rel;
}
void pdlp_b200_compute_relative_residuals(const PdlpTerminationCriteria* criteria, const PdlpConvergenceInformation* stats,
                                          const PdlpBoundNorms* bound_norms, double out[5]) {
  const RelativeResiduals r = ComputeRelativeResiduals(EffectiveOptimalityCriteria(*criteria), *stats, *bound_norms);
  out[0] = r.l_inf_primal; out[1] = r.l2_primal; out[2] = r.l_inf_dual; out[3] = r.l2_dual; out[4] = r.gap;
}
void pdlp_b200_bound_norms_from_problem_stats(const PdlpQuadraticProgramStats* stats, PdlpBoundNorms* out) { *out = BoundNormsFromProblemStats(*stats); }
}  // extern "C"

}  // namespace pdlp_b200
