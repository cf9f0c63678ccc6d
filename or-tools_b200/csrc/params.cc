// params.cc -- proto defaults and ValidatePrimalDualHybridGradientParams for the
// POD mirror of PrimalDualHybridGradientParams (solvers.proto:66-497,
// solvers_proto_validation.cc:33-298). Error strings follow the reference.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>

#include "solver.h"

namespace pdlp_b200 {
namespace {

constexpr double kTiny = 1.0e-50, kHuge = 1.0e50;

std::string G(double v) {
  char b[64];
  std::snprintf(b, sizeof(b), "%g", v);
  return b;
}

struct Check {  // accumulates the first error only
  std::string error;
  bool ok() const { return error.empty(); }
  void Fail(const std::string& m) { if (ok()) error = m; }
  void NonNegative(double v, const char* name) {
    if (!ok()) return;
    if (std::isnan(v)) Fail(std::string(name) + " is NAN");
    else if (v < 0) Fail(std::string(name) + " must be non-negative");
  }
  void NotNan(double v, const char* name) { if (ok() && std::isnan(v)) Fail(std::string(name) + " is NAN"); }
};

std::string ValidateCriteria(const PdlpTerminationCriteria& c) {  // :45-118
  Check k;
  if (c.optimality_norm != PDLP_OPTIMALITY_NORM_L_INF && c.optimality_norm != PDLP_OPTIMALITY_NORM_L2 &&
      c.optimality_norm != PDLP_OPTIMALITY_NORM_L_INF_COMPONENTWISE)
    return "invalid value for optimality_norm";
  if (c.optimality_criteria_case != PDLP_OPTIMALITY_CRITERIA_NOT_SET) {
    if (c.has_eps_optimal_absolute)
      return "eps_optimal_absolute should not be set if detailed_optimality_criteria or simple_optimality_criteria is used";
    if (c.has_eps_optimal_relative)
      return "eps_optimal_relative should not be set if detailed_optimality_criteria or simple_optimality_criteria is used";
  }
  switch (c.optimality_criteria_case) {
    case PDLP_DETAILED_OPTIMALITY_CRITERIA:
      k.NonNegative(c.eps_optimal_primal_residual_absolute, "detailed_optimality_criteria.eps_optimal_primal_residual_absolute");
      k.NonNegative(c.eps_optimal_primal_residual_relative, "detailed_optimality_criteria.eps_optimal_primal_residual_relative");
      k.NonNegative(c.eps_optimal_dual_residual_absolute, "detailed_optimality_criteria.eps_optimal_dual_residual_absolute");
      k.NonNegative(c.eps_optimal_dual_residual_relative, "detailed_optimality_criteria.eps_optimal_dual_residual_relative");
      k.NonNegative(c.eps_optimal_objective_gap_absolute, "detailed_optimality_criteria.eps_optimal_objective_gap_absolute");
      k.NonNegative(c.eps_optimal_objective_gap_relative, "detailed_optimality_criteria.eps_optimal_objective_gap_relative");
      break;
    case PDLP_SIMPLE_OPTIMALITY_CRITERIA:
      k.NonNegative(c.simple_eps_optimal_absolute, "simple_optimality_criteria.eps_optimal_absolute");
      k.NonNegative(c.simple_eps_optimal_relative, "simple_optimality_criteria.eps_optimal_relative");
      break;
    default:
      k.NonNegative(c.eps_optimal_absolute, "eps_optimal_absolute");
      k.NonNegative(c.eps_optimal_relative, "eps_optimal_relative");
  }
  k.NonNegative(c.eps_primal_infeasible, "eps_primal_infeasible");
  k.NonNegative(c.eps_dual_infeasible, "eps_dual_infeasible");
  k.NonNegative(c.time_sec_limit, "time_sec_limit");
  if (k.ok() && c.iteration_limit < 0) k.Fail("iteration_limit must be non-negative");
  k.NonNegative(c.kkt_matrix_pass_limit, "kkt_matrix_pass_limit");
  return k.error;
}

}  // namespace

void SetDefaultParams(PdlpParams* p) {
  std::memset(p, 0, sizeof(*p));
  const double inf = std::numeric_limits<double>::infinity();
  PdlpTerminationCriteria& t = p->termination_criteria;
  t.optimality_norm = PDLP_OPTIMALITY_NORM_L2;
  t.simple_eps_optimal_absolute = t.simple_eps_optimal_relative = 1e-6;
  t.eps_optimal_primal_residual_absolute = t.eps_optimal_primal_residual_relative = 1e-6;
  t.eps_optimal_dual_residual_absolute = t.eps_optimal_dual_residual_relative = 1e-6;
  t.eps_optimal_objective_gap_absolute = t.eps_optimal_objective_gap_relative = 1e-6;
  t.eps_optimal_absolute = t.eps_optimal_relative = 1e-6;
  t.eps_primal_infeasible = t.eps_dual_infeasible = 1e-8;
  t.time_sec_limit = inf;
  t.iteration_limit = std::numeric_limits<int32_t>::max();
  t.kkt_matrix_pass_limit = inf;
  p->num_threads = 1;
  p->scheduler_type = PDLP_SCHEDULER_TYPE_GOOGLE_THREADPOOL;
  p->major_iteration_frequency = 64;
  p->termination_check_frequency = 64;
  p->restart_strategy = PDLP_ADAPTIVE_HEURISTIC;
  p->primal_weight_update_smoothing = 0.5;
  p->l_inf_ruiz_iterations = 5;
  p->l2_norm_rescaling = 1;
  p->sufficient_reduction_for_restart = 0.1;
  p->necessary_reduction_for_restart = 0.9;
  p->linesearch_rule = PDLP_ADAPTIVE_LINESEARCH_RULE;
  p->adaptive_step_size_reduction_exponent = 0.3;
  p->adaptive_step_size_growth_exponent = 0.6;
  p->malitsky_pock_step_size_downscaling_factor = 0.7;
  p->malitsky_pock_linesearch_contraction_factor = 0.99;
  p->malitsky_pock_step_size_interpolation = 1.0;
  p->initial_step_size_scaling = 1.0;
  p->infinite_constraint_bound_threshold = inf;
  p->handle_some_primal_gradients_on_finite_bounds_as_residuals = 1;
  p->diagonal_qp_trust_region_solver_tolerance = 1e-8;
}

std::string ValidateParams(const PdlpParams& p) {  // :171-298
  {
    const std::string e = ValidateCriteria(p.termination_criteria);
    if (!e.empty()) return e + "; termination_criteria invalid";
  }
  Check k;
  auto range_msg = [](const char* name, const char* rest) { return std::string(name) + rest; };
  if (p.num_threads <= 0) return "num_threads must be positive";
  if (p.verbosity_level < 0) return "verbosity_level must be non-negative";
  if (p.log_interval_seconds < 0.0) return "log_interval_seconds must be non-negative";
  k.NotNan(p.log_interval_seconds, "log_interval_seconds");
  if (!k.ok()) return k.error;
  if (p.major_iteration_frequency <= 0) return "major_iteration_frequency must be positive";
  if (p.termination_check_frequency <= 0) return "termination_check_frequency must be positive";
  if (p.restart_strategy < PDLP_NO_RESTARTS || p.restart_strategy > PDLP_ADAPTIVE_DISTANCE_BASED) return "invalid restart_strategy";
  k.NotNan(p.primal_weight_update_smoothing, "primal_weight_update_smoothing");
  if (!k.ok()) return k.error;
  if (p.primal_weight_update_smoothing < 0 || p.primal_weight_update_smoothing > 1)
    return range_msg("primal_weight_update_smoothing", " must be between 0 and 1 inclusive");
  k.NotNan(p.initial_primal_weight, "initial_primal_weight");
  if (!k.ok()) return k.error;
  if (p.has_initial_primal_weight && (p.initial_primal_weight <= kTiny || p.initial_primal_weight >= kHuge))
    return "initial_primal_weight must be between " + G(kTiny) + " and " + G(kHuge) + " if specified";
  if (p.l_inf_ruiz_iterations < 0) return "l_inf_ruiz_iterations must be non-negative";
  if (p.l_inf_ruiz_iterations > 100) return "l_inf_ruiz_iterations must be at most 100";
  k.NotNan(p.sufficient_reduction_for_restart, "sufficient_reduction_for_restart");
  if (!k.ok()) return k.error;
  if (p.sufficient_reduction_for_restart <= 0 || p.sufficient_reduction_for_restart >= 1)
    return range_msg("sufficient_reduction_for_restart", " must be between 0 and 1 exclusive");
  k.NotNan(p.necessary_reduction_for_restart, "necessary_reduction_for_restart");
  if (!k.ok()) return k.error;
  if (p.necessary_reduction_for_restart < p.sufficient_reduction_for_restart || p.necessary_reduction_for_restart >= 1)
    return "necessary_reduction_for_restart must be in the interval [sufficient_reduction_for_restart, 1)";
  if (p.linesearch_rule != PDLP_ADAPTIVE_LINESEARCH_RULE && p.linesearch_rule != PDLP_MALITSKY_POCK_LINESEARCH_RULE &&
      p.linesearch_rule != PDLP_CONSTANT_STEP_SIZE_RULE)
    return "invalid linesearch_rule";
  {  // :120-139
    Check a;
    a.NotNan(p.adaptive_step_size_reduction_exponent, "step_size_reduction_exponent");
    if (a.ok() && (p.adaptive_step_size_reduction_exponent < 0.1 || p.adaptive_step_size_reduction_exponent > 1.0))
      a.Fail("step_size_reduction_exponent must be between 0.1 and 1.0 inclusive");
    a.NotNan(p.adaptive_step_size_growth_exponent, "step_size_growth_exponent");
    if (a.ok() && (p.adaptive_step_size_growth_exponent < 0.1 || p.adaptive_step_size_growth_exponent > 1.0))
      a.Fail("step_size_growth_exponent must be between 0.1 and 1.0 inclusive");
    if (!a.ok()) return a.error + "; adaptive_linesearch_parameters invalid";
  }
  {  // :141-169
    Check a;
    a.NotNan(p.malitsky_pock_step_size_downscaling_factor, "step_size_downscaling_factor");
    if (a.ok() && (p.malitsky_pock_step_size_downscaling_factor <= kTiny || p.malitsky_pock_step_size_downscaling_factor >= 1))
      a.Fail("step_size_downscaling_factor must be between " + G(kTiny) + " and 1 exclusive");
    a.NotNan(p.malitsky_pock_linesearch_contraction_factor, "linesearch_contraction_factor");
    if (a.ok() && (p.malitsky_pock_linesearch_contraction_factor <= 0 || p.malitsky_pock_linesearch_contraction_factor >= 1))
      a.Fail("linesearch_contraction_factor must be between 0 and 1 exclusive");
    a.NotNan(p.malitsky_pock_step_size_interpolation, "step_size_interpolation");
    if (a.ok() && (p.malitsky_pock_step_size_interpolation < 0 || p.malitsky_pock_step_size_interpolation >= kHuge))
      a.Fail("step_size_interpolation must be non-negative and less than " + G(kHuge));
    if (!a.ok()) return a.error + "; malitsky_pock_parameters invalid";
  }
  k.NotNan(p.initial_step_size_scaling, "initial_step_size_scaling");
  if (!k.ok()) return k.error;
  if (p.initial_step_size_scaling <= kTiny || p.initial_step_size_scaling >= kHuge)
    return "initial_step_size_scaling must be between " + G(kTiny) + " and " + G(kHuge);
  k.NotNan(p.infinite_constraint_bound_threshold, "infinite_constraint_bound_threshold");
  if (!k.ok()) return k.error;
  if (p.infinite_constraint_bound_threshold <= 0.0) return "infinite_constraint_bound_threshold must be positive";
  k.NotNan(p.diagonal_qp_trust_region_solver_tolerance, "diagonal_qp_trust_region_solver_tolerance");
  if (!k.ok()) return k.error;
  if (p.diagonal_qp_trust_region_solver_tolerance < 10 * std::numeric_limits<double>::epsilon())
    return "diagonal_qp_trust_region_solver_tolerance must be at least " + G(10 * std::numeric_limits<double>::epsilon());
  if (p.use_feasibility_polishing && p.handle_some_primal_gradients_on_finite_bounds_as_residuals)
    return "use_feasibility_polishing requires !handle_some_primal_gradients_on_finite_bounds_as_residuals";
  if (p.use_feasibility_polishing && p.presolve_use_glop) return "use_feasibility_polishing and glop presolve can not be used together.";
  return "";
}

// Structure of the caller's CSC view, shared by every entry point that walks the arrays on the
// host (the format writers; the device path checks the same things while it builds the images):
// consistent vector sizes, non-null arrays, col_starts[0] == 0, monotone col_starts,
// col_starts[n] == num_nonzeros and 0 <= row < m.
std::string ValidateView(const PdlpProblemView& v) {
  const int64_t n = v.num_variables, m = v.num_constraints;
  if (n < 0 || m < 0 || v.num_nonzeros < 0) return "negative dimension";
  auto size_ok = [](int64_t given, int64_t want) { return given < 0 || given == want; };
  if (!size_ok(v.variable_lower_bounds_size, n) || !size_ok(v.variable_upper_bounds_size, n) || !size_ok(v.objective_vector_size, n) ||
      !size_ok(v.constraint_lower_bounds_size, m) || !size_ok(v.constraint_upper_bounds_size, m) ||
      (v.objective_matrix_diagonal != nullptr && !size_ok(v.objective_matrix_size, n)))
    return "Inconsistent dimensions: a vector of the problem view does not match the constraint matrix";
  if (n > 0 && (v.variable_lower_bounds == nullptr || v.variable_upper_bounds == nullptr || v.objective_vector == nullptr)) return "null vector in the problem view";
  if (m > 0 && (v.constraint_lower_bounds == nullptr || v.constraint_upper_bounds == nullptr)) return "null vector in the problem view";
  if (v.col_starts == nullptr) return n == 0 && v.num_nonzeros == 0 ? "" : "col_starts is null";
  if (v.col_starts[0] != 0) return "col_starts[0] is not 0";
  for (int64_t c = 0; c < n; ++c)
    if (v.col_starts[c + 1] < v.col_starts[c]) return "col_starts is not monotone";
  if (v.col_starts[n] != v.num_nonzeros)
    return "col_starts[n] = " + std::to_string(static_cast<long long>(v.col_starts[n])) + " differs from num_nonzeros = " + std::to_string(static_cast<long long>(v.num_nonzeros));
  if (v.num_nonzeros > 0 && (v.row_indices == nullptr || v.values == nullptr)) return "null matrix array in the problem view";
  for (int64_t k = 0; k < v.num_nonzeros; ++k)
    if (v.row_indices[k] < 0 || v.row_indices[k] >= m) return "row index out of range";
  return "";
}

}  // namespace pdlp_b200
