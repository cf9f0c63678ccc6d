// pdlp_b200.hpp -- header-only C++17 face of the C ABI (pdlp_b200.h) with the names and
// semantics of the reference's direct API, for callers that hold plain vectors instead of
// Eigen / protobuf objects:
//
//   ortools/pdlp/quadratic_program.h:61-151          -> pdlp_b200::QuadraticProgram
//   ortools/pdlp/solvers.proto:238-497               -> pdlp_b200::PrimalDualHybridGradientParams (the POD, defaults set)
//   ortools/pdlp/primal_dual_hybrid_gradient.h:36-169-> pdlp_b200::SolverResult, PrimalDualHybridGradient(...)
//   ortools/pdlp/iteration_stats.h (GetConvergenceInformation), *_Name() of the generated enums
//
// examples/solve_simple_lp.cc is ortools/pdlp/samples/simple_pdlp_program.cc written
// against this header. Link with -lpdlp_b200 (or-tools_b200/lib). There is no CPU fallback:
// without a usable GPU the result carries TERMINATION_REASON_OTHER and an explanatory string.
#ifndef PDLP_B200_HPP_
#define PDLP_B200_HPP_

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <functional>
#include <limits>
#include <optional>
#include <string>
#include <tuple>
#include <vector>

#include "pdlp_b200.h"

namespace pdlp_b200 {

constexpr double kInfinity = std::numeric_limits<double>::infinity();

// (row, column, value), like Eigen::Triplet<double, int64_t>.
struct Triplet {
  int64_t row, col;
  double value;
};

// Adds up runs of consecutive triplets with the same (row, column) in place, keeping the order
// (CombineRepeatedTripletsInPlace, quadratic_program.h:200-208).
inline void CombineRepeatedTripletsInPlace(std::vector<Triplet>& triplets) {
  size_t kept = 0;
  for (size_t k = 0; k < triplets.size(); ++k) {
    if (kept > 0 && triplets[kept - 1].row == triplets[k].row && triplets[kept - 1].col == triplets[k].col) {
      triplets[kept - 1].value += triplets[k].value;
    } else {
      triplets[kept++] = triplets[k];
    }
  }
  triplets.resize(kept);
}

// min 1/2 x'Qx + c'x  s.t.  l_c <= Kx <= u_c,  l_v <= x <= u_v, Q diagonal (quadratic_program.h:61-151).
// K is held in compressed column form with int64 indices, like the reference's
// Eigen::SparseMatrix<double, ColMajor, int64_t>.
struct QuadraticProgram {
  QuadraticProgram() = default;
  QuadraticProgram(int64_t num_variables, int64_t num_constraints) { ResizeAndInitialize(num_variables, num_constraints); }
  void ResizeAndInitialize(int64_t num_variables, int64_t num_constraints) {  // quadratic_program.h:92-107
    num_variables_ = num_variables;
    num_constraints_ = num_constraints;
    objective_vector.assign(num_variables, 0.0);
    objective_matrix_diagonal.reset();
    col_starts.assign(num_variables + 1, 0);
    row_indices.clear();
    values.clear();
    constraint_lower_bounds.assign(num_constraints, -kInfinity);
    constraint_upper_bounds.assign(num_constraints, kInfinity);
    variable_lower_bounds.assign(num_variables, -kInfinity);
    variable_upper_bounds.assign(num_variables, kInfinity);
    problem_name.reset();
    variable_names.reset();
    constraint_names.reset();
    objective_offset = 0.0;
    objective_scaling_factor = 1.0;
  }
  // constraint_matrix.setFromTriplets(): duplicates are summed, rows sorted inside a column.
  void SetConstraintMatrixFromTriplets(std::vector<Triplet> t) {
    std::sort(t.begin(), t.end(), [](const Triplet& a, const Triplet& b) { return std::tie(a.col, a.row) < std::tie(b.col, b.row); });
    col_starts.assign(num_variables_ + 1, 0);
    row_indices.clear();
    values.clear();
    for (size_t k = 0; k < t.size(); ++k) {
      if (k > 0 && t[k].col == t[k - 1].col && t[k].row == t[k - 1].row) {
        values.back() += t[k].value;
        continue;
      }
      row_indices.push_back(t[k].row);
      values.push_back(t[k].value);
      ++col_starts[t[k].col + 1];
    }
    for (int64_t c = 0; c < num_variables_; ++c) col_starts[c + 1] += col_starts[c];
  }
  int64_t num_variables() const { return num_variables_; }
  int64_t num_constraints() const { return num_constraints_; }
  double ApplyObjectiveScalingAndOffset(double objective) const { return objective_scaling_factor * (objective + objective_offset); }

  std::vector<double> objective_vector;
  std::optional<std::vector<double>> objective_matrix_diagonal;  // nullopt: a linear program
  std::vector<int64_t> col_starts, row_indices;                  // K, compressed columns
  std::vector<double> values;
  std::vector<double> constraint_lower_bounds, constraint_upper_bounds, variable_lower_bounds, variable_upper_bounds;
  std::optional<std::string> problem_name;
  std::optional<std::vector<std::string>> variable_names, constraint_names;  // quadratic_program.h:146-147
  double objective_offset = 0.0, objective_scaling_factor = 1.0;

  PdlpProblemView View() const {
    PdlpProblemView v{};
    v.num_variables = num_variables_;
    v.num_constraints = num_constraints_;
    v.num_nonzeros = static_cast<int64_t>(values.size());
    v.col_starts = col_starts.data();
    v.row_indices = row_indices.data();
    v.values = values.data();
    v.objective_vector = objective_vector.data();
    v.objective_matrix_diagonal = objective_matrix_diagonal ? objective_matrix_diagonal->data() : nullptr;
    v.constraint_lower_bounds = constraint_lower_bounds.data();
    v.constraint_upper_bounds = constraint_upper_bounds.data();
    v.variable_lower_bounds = variable_lower_bounds.data();
    v.variable_upper_bounds = variable_upper_bounds.data();
    v.objective_offset = objective_offset;
    v.objective_scaling_factor = objective_scaling_factor;
    v.problem_name = problem_name ? problem_name->c_str() : nullptr;
    v.objective_vector_size = static_cast<int64_t>(objective_vector.size());
    v.objective_matrix_size = objective_matrix_diagonal ? static_cast<int64_t>(objective_matrix_diagonal->size()) : -1;
    v.constraint_lower_bounds_size = static_cast<int64_t>(constraint_lower_bounds.size());
    v.constraint_upper_bounds_size = static_cast<int64_t>(constraint_upper_bounds.size());
    v.variable_lower_bounds_size = static_cast<int64_t>(variable_lower_bounds.size());
    v.variable_upper_bounds_size = static_cast<int64_t>(variable_upper_bounds.size());
    return v;
  }

 private:
  int64_t num_variables_ = 0, num_constraints_ = 0;
};

// quadratic_program.h:153-158
inline bool IsLinearProgram(const QuadraticProgram& qp) { return !qp.objective_matrix_diagonal.has_value(); }

// ValidateQuadraticProgramDimensions (quadratic_program.cc:38-97): "" if the vector and matrix sizes
// agree, else the description of the first inconsistency (the reference's InvalidArgumentError text).
inline std::string ValidateQuadraticProgramDimensions(const QuadraticProgram& qp) {
  const int64_t var_lb_size = static_cast<int64_t>(qp.variable_lower_bounds.size());
  const int64_t con_lb_size = static_cast<int64_t>(qp.constraint_lower_bounds.size());
  auto mismatch = [](const char* a, int64_t a_size, const char* b, int64_t b_size, const char* unit) {
    return std::string("Inconsistent dimensions: ") + a + " has size " + std::to_string(a_size) + " while " + b + " has " +
           (unit[0] != '\0' ? std::to_string(b_size) + " " + unit : "size " + std::to_string(b_size));
  };
  if (var_lb_size != static_cast<int64_t>(qp.variable_upper_bounds.size()))
    return mismatch("variable lower bound vector", var_lb_size, "variable upper bound vector", static_cast<int64_t>(qp.variable_upper_bounds.size()), "");
  if (var_lb_size != static_cast<int64_t>(qp.objective_vector.size()))
    return mismatch("variable lower bound vector", var_lb_size, "objective vector", static_cast<int64_t>(qp.objective_vector.size()), "");
  if (var_lb_size != qp.num_variables() || var_lb_size + 1 != static_cast<int64_t>(qp.col_starts.size()))
    return mismatch("variable lower bound vector", var_lb_size, "constraint matrix", static_cast<int64_t>(qp.col_starts.size()) - 1, "columns");
  if (qp.objective_matrix_diagonal && var_lb_size != static_cast<int64_t>(qp.objective_matrix_diagonal->size()))
    return mismatch("variable lower bound vector", var_lb_size, "objective matrix", static_cast<int64_t>(qp.objective_matrix_diagonal->size()), "rows");
  if (con_lb_size != static_cast<int64_t>(qp.constraint_upper_bounds.size()))
    return mismatch("constraint lower bound vector", con_lb_size, "constraint upper bound vector", static_cast<int64_t>(qp.constraint_upper_bounds.size()), "");
  if (con_lb_size != qp.num_constraints())
    return mismatch("constraint lower bound vector", con_lb_size, "constraint matrix", qp.num_constraints(), "rows");
  if (qp.variable_names && var_lb_size != static_cast<int64_t>(qp.variable_names->size()))
    return mismatch("variable lower bound vector", var_lb_size, "variable names", static_cast<int64_t>(qp.variable_names->size()), "");
  if (qp.constraint_names && con_lb_size != static_cast<int64_t>(qp.constraint_names->size()))
    return mismatch("constraint lower bound vector", con_lb_size, "constraint names", static_cast<int64_t>(qp.constraint_names->size()), "");
  return std::string();
}

// The parameter POD with the proto defaults (solvers.proto:238-497); fields are set directly,
// e.g. params.termination_criteria.simple_eps_optimal_relative = 1e-6 after choosing
// params.termination_criteria.optimality_criteria_case = PDLP_SIMPLE_OPTIMALITY_CRITERIA.
struct PrimalDualHybridGradientParams : PdlpParams {
  PrimalDualHybridGradientParams() { pdlp_b200_params_set_defaults(this); }
  void SetSimpleOptimalityCriteria(double eps_optimal_absolute, double eps_optimal_relative) {
    termination_criteria.optimality_criteria_case = PDLP_SIMPLE_OPTIMALITY_CRITERIA;
    termination_criteria.simple_eps_optimal_absolute = eps_optimal_absolute;
    termination_criteria.simple_eps_optimal_relative = eps_optimal_relative;
  }
};

struct PrimalAndDualSolution {  // primal_dual_hybrid_gradient.h:31-34
  std::vector<double> primal_solution, dual_solution;
};

struct FeasibilityPolishingDetails {  // solve_log.proto:371-383
  int polishing_phase_type = 0, main_iteration_count = 0, termination_reason = 0, iteration_count = 0, solution_type = 0;
  double solve_time_sec = 0;
  PdlpParams params{};
  PdlpIterationStats solution_stats{};
  std::vector<PdlpIterationStats> iteration_stats;
};

struct SolveLog {  // solve_log.proto:385-459
  std::optional<std::string> instance_name, termination_string;
  int termination_reason = PDLP_TERMINATION_REASON_UNSPECIFIED;
  int iteration_count = 0;
  double solve_time_sec = 0, preprocessing_time_sec = 0;
  int solution_type = PDLP_POINT_TYPE_UNSPECIFIED;
  std::optional<PdlpIterationStats> solution_stats;
  std::optional<PdlpQuadraticProgramStats> original_problem_stats, preprocessed_problem_stats;
  std::vector<PdlpIterationStats> iteration_stats;
  PdlpParams params{};
  std::vector<FeasibilityPolishingDetails> feasibility_polishing_details;
  int64_t gpu_kernel_launches = 0;
};

struct SolverResult {  // primal_dual_hybrid_gradient.h:36-71: vectors are for the original, unscaled problem
  std::vector<double> primal_solution, dual_solution, reduced_costs;
  SolveLog solve_log;
};

using IterationStatsCallback = std::function<void(const PdlpIterationCallbackInfo&)>;

inline const char* TerminationReason_Name(int reason) {
  static const char* const kNames[] = {"TERMINATION_REASON_UNSPECIFIED", "TERMINATION_REASON_OPTIMAL", "TERMINATION_REASON_PRIMAL_INFEASIBLE",
                                       "TERMINATION_REASON_DUAL_INFEASIBLE", "TERMINATION_REASON_TIME_LIMIT", "TERMINATION_REASON_ITERATION_LIMIT",
                                       "TERMINATION_REASON_NUMERICAL_ERROR", "TERMINATION_REASON_OTHER", "TERMINATION_REASON_KKT_MATRIX_PASS_LIMIT",
                                       "TERMINATION_REASON_INVALID_PROBLEM", "TERMINATION_REASON_INVALID_PARAMETER",
                                       "TERMINATION_REASON_PRIMAL_OR_DUAL_INFEASIBLE", "TERMINATION_REASON_INTERRUPTED_BY_USER",
                                       "TERMINATION_REASON_INVALID_INITIAL_SOLUTION"};
  return reason >= 0 && reason < 14 ? kNames[reason] : "?";
}
inline const char* PointType_Name(int type) {
  static const char* const kNames[] = {"POINT_TYPE_UNSPECIFIED", "POINT_TYPE_CURRENT_ITERATE", "POINT_TYPE_ITERATE_DIFFERENCE", "POINT_TYPE_AVERAGE_ITERATE",
                                       "POINT_TYPE_NONE", "POINT_TYPE_PRESOLVER_SOLUTION", "POINT_TYPE_FEASIBILITY_POLISHING_SOLUTION"};
  return type >= 0 && type < 7 ? kNames[type] : "?";
}
// iteration_stats.cc:597-606
inline std::optional<PdlpConvergenceInformation> GetConvergenceInformation(const PdlpIterationStats& stats, int candidate_type) {
  for (int i = 0; i < stats.num_convergence_information; ++i)
    if (stats.convergence_information[i].candidate_type == candidate_type) return stats.convergence_information[i];
  return std::nullopt;
}
// iteration_stats.cc:605-625
inline std::optional<PdlpInfeasibilityInformation> GetInfeasibilityInformation(const PdlpIterationStats& stats, int candidate_type) {
  for (int i = 0; i < stats.num_infeasibility_information; ++i)
    if (stats.infeasibility_information[i].candidate_type == candidate_type) return stats.infeasibility_information[i];
  return std::nullopt;
}
inline std::optional<PdlpPointMetadata> GetPointMetadata(const PdlpIterationStats& stats, int point_type) {
  for (int i = 0; i < stats.num_point_metadata; ++i)
    if (stats.point_metadata[i].point_type == point_type) return stats.point_metadata[i];
  return std::nullopt;
}

// PrimalDualHybridGradient(qp, params, initial_solution, interrupt_solve, message_callback,
// iteration_stats_callback), primal_dual_hybrid_gradient.h:151-169. Blocking; callbacks run on
// the calling thread; `interrupt_solve` is polled at the termination checks (through the
// iteration callback a std::atomic<bool> is mirrored into the int32 flag of the C ABI).
inline SolverResult PrimalDualHybridGradient(const QuadraticProgram& qp, const PrimalDualHybridGradientParams& params,
                                             const std::optional<PrimalAndDualSolution>& initial_solution = std::nullopt,
                                             const std::atomic<bool>* interrupt_solve = nullptr,
                                             std::function<void(const std::string&)> message_callback = nullptr,
                                             IterationStatsCallback iteration_stats_callback = nullptr) {
  struct Ctx {
    std::function<void(const std::string&)>* msg;
    IterationStatsCallback* stats;
    const std::atomic<bool>* stop;
    volatile int32_t flag;
  } ctx{&message_callback, &iteration_stats_callback, interrupt_solve, 0};
  if (interrupt_solve != nullptr && interrupt_solve->load()) ctx.flag = 1;
  auto on_message = [](const char* m, void* u) {
    Ctx* c = static_cast<Ctx*>(u);
    if (*c->msg) (*c->msg)(m);
  };
  auto on_stats = [](const PdlpIterationCallbackInfo* info, void* u) {
    Ctx* c = static_cast<Ctx*>(u);
    if (*c->stats) (*c->stats)(*info);
    if (c->stop != nullptr && c->stop->load()) c->flag = 1;
  };
  const PdlpProblemView view = qp.View();
  PdlpResult r{};
  const bool has_init = initial_solution.has_value();
  const int32_t rc = pdlp_b200_primal_dual_hybrid_gradient(
      &view, &params, has_init ? initial_solution->primal_solution.data() : nullptr,
      has_init ? static_cast<int64_t>(initial_solution->primal_solution.size()) : 0, has_init ? initial_solution->dual_solution.data() : nullptr,
      has_init ? static_cast<int64_t>(initial_solution->dual_solution.size()) : 0, interrupt_solve != nullptr ? &ctx.flag : nullptr,
      message_callback ? +on_message : nullptr, (iteration_stats_callback || interrupt_solve != nullptr) ? +on_stats : nullptr, &ctx, &r);
  SolverResult out;
  SolveLog& log = out.solve_log;
  if (rc == PDLP_B200_STATUS_NO_DEVICE) {
    log.termination_reason = PDLP_TERMINATION_REASON_OTHER;
    log.termination_string = "no usable CUDA device (libpdlp_b200 has no CPU fallback)";
    return out;
  }
  if (r.primal_solution != nullptr) out.primal_solution.assign(r.primal_solution, r.primal_solution + r.primal_size);
  if (r.dual_solution != nullptr) out.dual_solution.assign(r.dual_solution, r.dual_solution + r.dual_size);
  if (r.reduced_costs != nullptr) out.reduced_costs.assign(r.reduced_costs, r.reduced_costs + r.primal_size);
  if (r.instance_name != nullptr) log.instance_name = r.instance_name;
  if (r.termination_string != nullptr) log.termination_string = r.termination_string;
  log.termination_reason = r.termination_reason;
  log.iteration_count = r.iteration_count;
  log.solve_time_sec = r.solve_time_sec;
  log.preprocessing_time_sec = r.preprocessing_time_sec;
  log.solution_type = r.solution_type;
  if (r.has_solution_stats) log.solution_stats = r.solution_stats;
  if (r.has_original_problem_stats) log.original_problem_stats = r.original_problem_stats;
  if (r.has_preprocessed_problem_stats) log.preprocessed_problem_stats = r.preprocessed_problem_stats;
  if (r.iteration_stats != nullptr) log.iteration_stats.assign(r.iteration_stats, r.iteration_stats + r.num_iteration_stats);
  log.params = r.params;
  for (int64_t k = 0; k < r.num_feasibility_polishing_details; ++k) {
    const PdlpFeasibilityPolishingDetails& d = r.feasibility_polishing_details[k];
    FeasibilityPolishingDetails o;
    o.polishing_phase_type = d.polishing_phase_type;
    o.main_iteration_count = d.main_iteration_count;
    o.params = d.params;
    o.termination_reason = d.termination_reason;
    o.iteration_count = d.iteration_count;
    o.solve_time_sec = d.solve_time_sec;
    o.solution_stats = d.solution_stats;
    o.solution_type = d.solution_type;
    if (d.iteration_stats != nullptr) o.iteration_stats.assign(d.iteration_stats, d.iteration_stats + d.num_iteration_stats);
    log.feasibility_polishing_details.push_back(std::move(o));
  }
  log.gpu_kernel_launches = r.gpu_kernel_launches;
  pdlp_b200_result_free(&r);
  return out;
}

}  // namespace pdlp_b200

#endif  // PDLP_B200_HPP_
