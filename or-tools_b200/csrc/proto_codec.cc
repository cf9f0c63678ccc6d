// proto_codec.cc -- see proto_codec.h.
#include "proto_codec.h"

#include <algorithm>
#include <cctype>
#include <cerrno>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>

namespace pdlp_b200 {
namespace proto {

// ---------------------------------------------------------------------------
// wire primitives
// ---------------------------------------------------------------------------
void Writer::Varint(uint64_t v) {
  while (v >= 0x80) {
    out_.push_back(static_cast<char>((v & 0x7F) | 0x80));
    v >>= 7;
  }
  out_.push_back(static_cast<char>(v));
}

void Writer::RawDouble(double v) {
  uint64_t bits;
  std::memcpy(&bits, &v, 8);
  char b[8];
  for (int i = 0; i < 8; ++i) b[i] = static_cast<char>((bits >> (8 * i)) & 0xFF);  // little endian
  out_.append(b, 8);
}

void Writer::Double(int field, double v) {
  Tag(field, kFixed64);
  RawDouble(v);
}

void Writer::Bytes(int field, std::string_view v) {
  Tag(field, kLengthDelimited);
  Varint(v.size());
  out_.append(v.data(), v.size());
}

void Writer::PackedDoubles(int field, const double* v, int64_t n) {
  if (n <= 0) return;
  Tag(field, kLengthDelimited);
  Varint(static_cast<uint64_t>(n) * 8);
  for (int64_t i = 0; i < n; ++i) RawDouble(v[i]);
}

void Writer::PackedInts(int field, const int32_t* v, int64_t n) {
  if (n <= 0) return;
  Writer body;
  for (int64_t i = 0; i < n; ++i) body.Varint(static_cast<uint64_t>(static_cast<int64_t>(v[i])));
  Bytes(field, body.out());
}

double WireField::AsDouble() const {
  if (type == kFixed32) {
    uint32_t b = static_cast<uint32_t>(value);
    float f;
    std::memcpy(&f, &b, 4);
    return f;
  }
  double d;
  std::memcpy(&d, &value, 8);
  return d;
}

bool Reader::Varint(uint64_t* v) {
  uint64_t r = 0;
  for (int shift = 0; shift < 70; shift += 7) {
    if (p_ >= end_) return ok_ = false;
    const uint8_t b = static_cast<uint8_t>(*p_++);
    if (shift < 64) r |= static_cast<uint64_t>(b & 0x7F) << shift;
    if (!(b & 0x80)) {
      *v = r;
      return true;
    }
  }
  return ok_ = false;
}

bool Reader::Next(WireField* f) {
  if (!ok_ || p_ >= end_) return false;
  uint64_t key;
  if (!Varint(&key)) return false;
  f->number = static_cast<int>(key >> 3);
  f->type = static_cast<WireType>(key & 7);
  f->value = 0;
  f->bytes = {};
  if (f->number <= 0) return ok_ = false;
  switch (f->type) {
    case kVarint:
      return Varint(&f->value);
    case kFixed64: {
      if (end_ - p_ < 8) return ok_ = false;
      uint64_t b = 0;
      for (int i = 0; i < 8; ++i) b |= static_cast<uint64_t>(static_cast<uint8_t>(p_[i])) << (8 * i);
      f->value = b;
      p_ += 8;
      return true;
    }
    case kFixed32: {
      if (end_ - p_ < 4) return ok_ = false;
      uint64_t b = 0;
      for (int i = 0; i < 4; ++i) b |= static_cast<uint64_t>(static_cast<uint8_t>(p_[i])) << (8 * i);
      f->value = b;
      p_ += 4;
      return true;
    }
    case kLengthDelimited: {
      uint64_t n;
      if (!Varint(&n)) return false;
      if (n > static_cast<uint64_t>(end_ - p_)) return ok_ = false;
      f->bytes = std::string_view(p_, n);
      p_ += n;
      return true;
    }
    default:
      return ok_ = false;  // groups are not used by these messages
  }
}

bool AppendDoubles(const WireField& f, std::vector<double>* out) {
  if (f.type == kFixed64) {
    out->push_back(f.AsDouble());
    return true;
  }
  if (f.type != kLengthDelimited || f.bytes.size() % 8 != 0) return false;
  const size_t n = f.bytes.size() / 8, at = out->size();
  out->resize(at + n);
  if (n) std::memcpy(out->data() + at, f.bytes.data(), n * 8);  // little-endian host (x86-64 / aarch64)
  return true;
}

namespace {
// Calls emit(value) for each of the varints stored back to back in `bytes`.
template <typename Emit>
bool ForEachPackedVarint(std::string_view bytes, Emit emit) {
  size_t at = 0;
  while (at < bytes.size()) {
    uint64_t acc = 0;
    int shift = 0;
    bool done = false;
    while (at < bytes.size() && shift < 70) {
      const uint8_t b = static_cast<uint8_t>(bytes[at++]);
      if (shift < 64) acc |= static_cast<uint64_t>(b & 0x7F) << shift;
      shift += 7;
      if (!(b & 0x80)) {
        done = true;
        break;
      }
    }
    if (!done) return false;
    emit(acc);
  }
  return true;
}
}  // namespace

bool AppendInt32s(const WireField& f, std::vector<int32_t>* out) {
  if (f.type == kVarint) {
    out->push_back(f.AsInt32());
    return true;
  }
  if (f.type != kLengthDelimited) return false;
  return ForEachPackedVarint(f.bytes, [&](uint64_t v) { out->push_back(static_cast<int32_t>(static_cast<int64_t>(v))); });
}

// ---------------------------------------------------------------------------
// schema tables
// ---------------------------------------------------------------------------
const char* EnumDef::NameOf(int number) const {
  for (const auto& v : values)
    if (v.number == number) return v.name;
  return nullptr;
}
bool EnumDef::NumberOf(std::string_view n, int* number) const {
  for (const auto& v : values)
    if (n == v.name) {
      *number = v.number;
      return true;
    }
  return false;
}
const FieldDef* Schema::ByName(std::string_view n) const {
  for (const auto& f : fields)
    if (n == f.name) return &f;
  return nullptr;
}
const FieldDef* Schema::ByNumber(int number) const {
  for (const auto& f : fields)
    if (f.number == number) return &f;
  return nullptr;
}

namespace {
using FT = FieldType;
FieldDef D(const char* n, int tag) { return {n, tag, FT::kDouble}; }
FieldDef I32(const char* n, int tag) { return {n, tag, FT::kInt32}; }
FieldDef I64(const char* n, int tag) { return {n, tag, FT::kInt64}; }
FieldDef B(const char* n, int tag) { return {n, tag, FT::kBool}; }
FieldDef S(const char* n, int tag) { return {n, tag, FT::kString}; }
FieldDef E(const char* n, int tag, const EnumDef* e) { return {n, tag, FT::kEnum, false, false, nullptr, e}; }
FieldDef M(const char* n, int tag, const Schema* s, bool repeated = false) { return {n, tag, FT::kMessage, repeated, false, s, nullptr}; }
FieldDef Rep(FieldDef f, bool packed) {
  f.repeated = true;
  f.packed = packed;
  return f;
}
FieldDef OneOf(FieldDef f, int id) {
  f.oneof = id;
  return f;
}

const EnumDef& OptimalityNormEnum() {  // solvers.proto:24-41
  static const EnumDef e{"OptimalityNorm", {{"OPTIMALITY_NORM_UNSPECIFIED", 0}, {"OPTIMALITY_NORM_L_INF", 1}, {"OPTIMALITY_NORM_L2", 2},
                                            {"OPTIMALITY_NORM_L_INF_COMPONENTWISE", 3}}};
  return e;
}
const EnumDef& SchedulerTypeEnum() {  // solvers.proto:44-51
  static const EnumDef e{"SchedulerType", {{"SCHEDULER_TYPE_UNSPECIFIED", 0}, {"SCHEDULER_TYPE_GOOGLE_THREADPOOL", 1},
                                           {"SCHEDULER_TYPE_EIGEN_THREADPOOL", 3}}};
  return e;
}
const EnumDef& RestartStrategyEnum() {  // solvers.proto:239-277
  static const EnumDef e{"RestartStrategy", {{"RESTART_STRATEGY_UNSPECIFIED", 0}, {"NO_RESTARTS", 1}, {"EVERY_MAJOR_ITERATION", 2},
                                             {"ADAPTIVE_HEURISTIC", 3}, {"ADAPTIVE_DISTANCE_BASED", 4}}};
  return e;
}
const EnumDef& LinesearchRuleEnum() {
  static const EnumDef e{"LinesearchRule", {{"LINESEARCH_RULE_UNSPECIFIED", 0}, {"ADAPTIVE_LINESEARCH_RULE", 1},
                                            {"MALITSKY_POCK_LINESEARCH_RULE", 2}, {"CONSTANT_STEP_SIZE_RULE", 3}}};
  return e;
}
const EnumDef& RestartChoiceEnum() {  // solve_log.proto:105-117
  static const EnumDef e{"RestartChoice", {{"RESTART_CHOICE_UNSPECIFIED", 0}, {"RESTART_CHOICE_NO_RESTART", 1},
                                           {"RESTART_CHOICE_WEIGHTED_AVERAGE_RESET", 2}, {"RESTART_CHOICE_RESTART_TO_AVERAGE", 3}}};
  return e;
}
const EnumDef& PointTypeEnum() {  // solve_log.proto:121-135
  static const EnumDef e{"PointType", {{"POINT_TYPE_UNSPECIFIED", 0}, {"POINT_TYPE_CURRENT_ITERATE", 1}, {"POINT_TYPE_ITERATE_DIFFERENCE", 2},
                                       {"POINT_TYPE_AVERAGE_ITERATE", 3}, {"POINT_TYPE_NONE", 4}, {"POINT_TYPE_PRESOLVER_SOLUTION", 5},
                                       {"POINT_TYPE_FEASIBILITY_POLISHING_SOLUTION", 6}}};
  return e;
}
const EnumDef& TerminationReasonEnum() {  // solve_log.proto:336-360
  static const EnumDef e{"TerminationReason",
                         {{"TERMINATION_REASON_UNSPECIFIED", 0}, {"TERMINATION_REASON_OPTIMAL", 1}, {"TERMINATION_REASON_PRIMAL_INFEASIBLE", 2},
                          {"TERMINATION_REASON_DUAL_INFEASIBLE", 3}, {"TERMINATION_REASON_TIME_LIMIT", 4}, {"TERMINATION_REASON_ITERATION_LIMIT", 5},
                          {"TERMINATION_REASON_NUMERICAL_ERROR", 6}, {"TERMINATION_REASON_OTHER", 7}, {"TERMINATION_REASON_KKT_MATRIX_PASS_LIMIT", 8},
                          {"TERMINATION_REASON_INVALID_PROBLEM", 9}, {"TERMINATION_REASON_INVALID_PARAMETER", 10},
                          {"TERMINATION_REASON_PRIMAL_OR_DUAL_INFEASIBLE", 11}, {"TERMINATION_REASON_INTERRUPTED_BY_USER", 12},
                          {"TERMINATION_REASON_INVALID_INITIAL_SOLUTION", 13}}};
  return e;
}
const EnumDef& PolishingPhaseTypeEnum() {  // solve_log.proto:362-369
  static const EnumDef e{"PolishingPhaseType", {{"POLISHING_PHASE_TYPE_UNSPECIFIED", 0}, {"POLISHING_PHASE_TYPE_PRIMAL_FEASIBILITY", 1},
                                                {"POLISHING_PHASE_TYPE_DUAL_FEASIBILITY", 2}}};
  return e;
}
const EnumDef& MPSolverResponseStatusEnum() {  // linear_solver.proto:519-586
  static const EnumDef e{"MPSolverResponseStatus",
                         {{"MPSOLVER_OPTIMAL", 0}, {"MPSOLVER_FEASIBLE", 1}, {"MPSOLVER_INFEASIBLE", 2}, {"MPSOLVER_UNBOUNDED", 3},
                          {"MPSOLVER_ABNORMAL", 4}, {"MPSOLVER_NOT_SOLVED", 6}, {"MPSOLVER_MODEL_IS_VALID", 97},
                          {"MPSOLVER_CANCELLED_BY_USER", 98}, {"MPSOLVER_UNKNOWN_STATUS", 99}, {"MPSOLVER_MODEL_INVALID", 5},
                          {"MPSOLVER_MODEL_INVALID_SOLUTION_HINT", 84}, {"MPSOLVER_MODEL_INVALID_SOLVER_PARAMETERS", 85},
                          {"MPSOLVER_SOLVER_TYPE_UNAVAILABLE", 7}, {"MPSOLVER_INCOMPATIBLE_OPTIONS", 113}}};
  return e;
}
const EnumDef& SolverTypeEnum() {  // MPModelRequest.SolverType, linear_solver.proto:456-490
  static const EnumDef e{"SolverType",
                         {{"CLP_LINEAR_PROGRAMMING", 0}, {"GLOP_LINEAR_PROGRAMMING", 2}, {"GLPK_LINEAR_PROGRAMMING", 1},
                          {"GUROBI_LINEAR_PROGRAMMING", 6}, {"XPRESS_LINEAR_PROGRAMMING", 101}, {"CPLEX_LINEAR_PROGRAMMING", 10},
                          {"HIGHS_LINEAR_PROGRAMMING", 15}, {"SCIP_MIXED_INTEGER_PROGRAMMING", 3}, {"GLPK_MIXED_INTEGER_PROGRAMMING", 4},
                          {"CBC_MIXED_INTEGER_PROGRAMMING", 5}, {"GUROBI_MIXED_INTEGER_PROGRAMMING", 7},
                          {"XPRESS_MIXED_INTEGER_PROGRAMMING", 102}, {"CPLEX_MIXED_INTEGER_PROGRAMMING", 11},
                          {"HIGHS_MIXED_INTEGER_PROGRAMMING", 16}, {"BOP_INTEGER_PROGRAMMING", 12}, {"SAT_INTEGER_PROGRAMMING", 14},
                          {"PDLP_LINEAR_PROGRAMMING", 8}, {"KNAPSACK_MIXED_INTEGER_PROGRAMMING", 13}}};
  return e;
}

const Schema& SimpleCriteriaSchema() {  // solvers.proto:113-120
  static const Schema s{"SimpleOptimalityCriteria", {D("eps_optimal_absolute", 1), D("eps_optimal_relative", 2)}};
  return s;
}
const Schema& DetailedCriteriaSchema() {  // solvers.proto:122-160
  static const Schema s{"DetailedOptimalityCriteria",
                        {D("eps_optimal_primal_residual_absolute", 1), D("eps_optimal_primal_residual_relative", 2),
                         D("eps_optimal_dual_residual_absolute", 3), D("eps_optimal_dual_residual_relative", 4),
                         D("eps_optimal_objective_gap_absolute", 5), D("eps_optimal_objective_gap_relative", 6)}};
  return s;
}
const Schema& AdaptiveLinesearchSchema() {  // solvers.proto:189-204
  static const Schema s{"AdaptiveLinesearchParams", {D("step_size_reduction_exponent", 1), D("step_size_growth_exponent", 2)}};
  return s;
}
const Schema& MalitskyPockSchema() {  // solvers.proto:206-226
  static const Schema s{"MalitskyPockParams",
                        {D("step_size_downscaling_factor", 1), D("linesearch_contraction_factor", 2), D("step_size_interpolation", 3)}};
  return s;
}
const Schema& PresolveOptionsSchema() {  // solvers.proto:366-384; glop_parameters (tag 2) is carried as opaque bytes
  static const Schema s{"PresolveOptions", {B("use_glop", 1), {"glop_parameters", 2, FT::kBytes}}};
  return s;
}
const Schema& QuadraticProgramStatsSchema() {  // solve_log.proto:28-102
  static const Schema s{"QuadraticProgramStats",
                        {I64("num_variables", 1), I64("num_constraints", 2), D("constraint_matrix_col_min_l_inf_norm", 3),
                         D("constraint_matrix_row_min_l_inf_norm", 4), I64("constraint_matrix_num_nonzeros", 5), D("constraint_matrix_abs_max", 6),
                         D("constraint_matrix_abs_min", 7), D("constraint_matrix_abs_avg", 8), D("constraint_matrix_l2_norm", 25),
                         D("combined_bounds_max", 9), D("combined_bounds_min", 10), D("combined_bounds_avg", 11), D("combined_bounds_l2_norm", 24),
                         D("combined_variable_bounds_max", 28), D("combined_variable_bounds_min", 29), D("combined_variable_bounds_avg", 30),
                         D("combined_variable_bounds_l2_norm", 31), I64("variable_bound_gaps_num_finite", 12), D("variable_bound_gaps_max", 13),
                         D("variable_bound_gaps_min", 14), D("variable_bound_gaps_avg", 15), D("variable_bound_gaps_l2_norm", 26),
                         D("objective_vector_abs_max", 16), D("objective_vector_abs_min", 17), D("objective_vector_abs_avg", 18),
                         D("objective_vector_l2_norm", 23), I64("objective_matrix_num_nonzeros", 19), D("objective_matrix_abs_max", 20),
                         D("objective_matrix_abs_min", 21), D("objective_matrix_abs_avg", 22), D("objective_matrix_l2_norm", 27)}};
  return s;
}
const Schema& ConvergenceInformationSchema() {  // solve_log.proto:139-205
  static const Schema s{"ConvergenceInformation",
                        {E("candidate_type", 1, &PointTypeEnum()), D("primal_objective", 2), D("dual_objective", 3), D("corrected_dual_objective", 4),
                         D("l_inf_primal_residual", 5), D("l2_primal_residual", 6), D("l_inf_componentwise_primal_residual", 24),
                         D("l_inf_dual_residual", 7), D("l2_dual_residual", 8), D("l_inf_componentwise_dual_residual", 25),
                         D("l_inf_primal_variable", 14), D("l2_primal_variable", 15), D("l_inf_dual_variable", 16), D("l2_dual_variable", 17)}};
  return s;
}
const Schema& InfeasibilityInformationSchema() {  // solve_log.proto:209-249
  static const Schema s{"InfeasibilityInformation",
                        {D("max_primal_ray_infeasibility", 1), D("primal_ray_linear_objective", 2), D("primal_ray_quadratic_norm", 3),
                         D("max_dual_ray_infeasibility", 4), D("dual_ray_objective", 5), E("candidate_type", 6, &PointTypeEnum())}};
  return s;
}
const Schema& PointMetadataSchema() {  // solve_log.proto:251-274
  static const Schema s{"PointMetadata",
                        {E("point_type", 1, &PointTypeEnum()), Rep(D("random_primal_projections", 2), true), Rep(D("random_dual_projections", 3), true),
                         I64("active_primal_variable_count", 4), I64("active_dual_variable_count", 5), I64("active_primal_variable_change", 6),
                         I64("active_dual_variable_change", 7)}};
  return s;
}
const Schema& FeasibilityPolishingDetailsSchema() {  // solve_log.proto:371-383
  static const Schema s{"FeasibilityPolishingDetails",
                        {E("polishing_phase_type", 1, &PolishingPhaseTypeEnum()), I32("main_iteration_count", 2), M("params", 3, &ParamsSchema()),
                         E("termination_reason", 4, &TerminationReasonEnum()), I32("iteration_count", 5), D("solve_time_sec", 6),
                         M("solution_stats", 7, &IterationStatsSchema()), E("solution_type", 8, &PointTypeEnum()),
                         M("iteration_stats", 9, &IterationStatsSchema(), true)}};
  return s;
}
const Schema& MPVariableSchema() {  // linear_solver.proto:49-73
  static const Schema s{"MPVariableProto",
                        {D("lower_bound", 1), D("upper_bound", 2), D("objective_coefficient", 3), B("is_integer", 4), S("name", 5),
                         I32("branching_priority", 6)}};
  return s;
}
const Schema& MPConstraintSchema() {  // linear_solver.proto:80-107
  static const Schema s{"MPConstraintProto",
                        {Rep(I32("var_index", 6), true), Rep(D("coefficient", 7), true), D("lower_bound", 2), D("upper_bound", 3), S("name", 4),
                         B("is_lazy", 5)}};
  return s;
}
const Schema& MPGeneralConstraintSchema() {  // only its presence matters (quadratic_program.cc:101-103)
  static const Schema s{"MPGeneralConstraintProto", {S("name", 1)}};
  return s;
}
const Schema& MPQuadraticObjectiveSchema() {  // linear_solver.proto:216-233
  static const Schema s{"MPQuadraticObjective", {Rep(I32("qvar1_index", 1), false), Rep(I32("qvar2_index", 2), false), Rep(D("coefficient", 3), false)}};
  return s;
}
}  // namespace

const Schema& TerminationCriteriaSchema() {  // solvers.proto:66-187
  static const Schema s{"TerminationCriteria",
                        {E("optimality_norm", 1, &OptimalityNormEnum()), OneOf(M("simple_optimality_criteria", 9, &SimpleCriteriaSchema()), 1),
                         OneOf(M("detailed_optimality_criteria", 10, &DetailedCriteriaSchema()), 1), D("eps_optimal_absolute", 2), D("eps_optimal_relative", 3),
                         D("eps_primal_infeasible", 4), D("eps_dual_infeasible", 5), D("time_sec_limit", 6), I32("iteration_limit", 7),
                         D("kkt_matrix_pass_limit", 8)}};
  return s;
}

const Schema& ParamsSchema() {  // solvers.proto:238-497
  static const Schema s{"PrimalDualHybridGradientParams",
                        {M("termination_criteria", 1, &TerminationCriteriaSchema()), I32("num_threads", 2), I32("num_shards", 27),
                         E("scheduler_type", 32, &SchedulerTypeEnum()), B("record_iteration_stats", 3), I32("verbosity_level", 26),
                         D("log_interval_seconds", 31), I32("major_iteration_frequency", 4), I32("termination_check_frequency", 5),
                         E("restart_strategy", 6, &RestartStrategyEnum()), D("primal_weight_update_smoothing", 7), D("initial_primal_weight", 8),
                         M("presolve_options", 16, &PresolveOptionsSchema()), I32("l_inf_ruiz_iterations", 9), B("l2_norm_rescaling", 10),
                         D("sufficient_reduction_for_restart", 11), D("necessary_reduction_for_restart", 17),
                         E("linesearch_rule", 12, &LinesearchRuleEnum()), M("adaptive_linesearch_parameters", 18, &AdaptiveLinesearchSchema()),
                         M("malitsky_pock_parameters", 19, &MalitskyPockSchema()), D("initial_step_size_scaling", 25),
                         Rep(I32("random_projection_seeds", 28), true), D("infinite_constraint_bound_threshold", 22),
                         B("handle_some_primal_gradients_on_finite_bounds_as_residuals", 29), B("use_diagonal_qp_trust_region_solver", 23),
                         D("diagonal_qp_trust_region_solver_tolerance", 24), B("use_feasibility_polishing", 30),
                         B("apply_feasibility_polishing_after_limits_reached", 33), B("apply_feasibility_polishing_if_solver_is_interrupted", 34)}};
  return s;
}

const Schema& IterationStatsSchema() {  // solve_log.proto:281-334
  static const Schema s{"IterationStats",
                        {I32("iteration_number", 1), M("convergence_information", 2, &ConvergenceInformationSchema(), true),
                         M("infeasibility_information", 3, &InfeasibilityInformationSchema(), true), M("point_metadata", 11, &PointMetadataSchema(), true),
                         D("cumulative_kkt_matrix_passes", 4), I32("cumulative_rejected_steps", 5), D("cumulative_time_sec", 6),
                         E("restart_used", 7, &RestartChoiceEnum()), D("step_size", 8), D("primal_weight", 9)}};
  return s;
}

const Schema& SolveLogSchema() {  // solve_log.proto:385-459
  static const Schema s{"SolveLog",
                        {S("instance_name", 1), M("params", 14, &ParamsSchema()), E("termination_reason", 3, &TerminationReasonEnum()),
                         S("termination_string", 4), I32("iteration_count", 5), D("preprocessing_time_sec", 13), D("solve_time_sec", 6),
                         M("solution_stats", 8, &IterationStatsSchema()), E("solution_type", 10, &PointTypeEnum()),
                         M("iteration_stats", 7, &IterationStatsSchema(), true), M("original_problem_stats", 11, &QuadraticProgramStatsSchema()),
                         M("preprocessed_problem_stats", 12, &QuadraticProgramStatsSchema()),
                         M("feasibility_polishing_details", 15, &FeasibilityPolishingDetailsSchema(), true)}};
  return s;
}

const Schema& MPModelSchema() {  // linear_solver.proto:263-317
  static const Schema s{"MPModelProto",
                        {M("variable", 3, &MPVariableSchema(), true), M("constraint", 4, &MPConstraintSchema(), true),
                         M("general_constraint", 7, &MPGeneralConstraintSchema(), true), B("maximize", 1), D("objective_offset", 2),
                         M("quadratic_objective", 8, &MPQuadraticObjectiveSchema()), S("name", 5)}};
  return s;
}

const Schema& MPModelRequestSchema() {  // linear_solver.proto:444-516
  static const Schema s{"MPModelRequest",
                        {M("model", 1, &MPModelSchema()), E("solver_type", 2, &SolverTypeEnum()), D("solver_time_limit_seconds", 3),
                         B("enable_internal_solver_output", 4), S("solver_specific_parameters", 5),
                         B("ignore_solver_specific_parameters_failure", 9)}};
  return s;
}

const Schema& MPSolutionResponseSchema() {  // linear_solver.proto:600-672
  static const Schema s{"MPSolutionResponse",
                        {E("status", 1, &MPSolverResponseStatusEnum()), S("status_str", 7), D("objective_value", 2), D("best_objective_bound", 5),
                         Rep(D("variable_value", 3), true), Rep(D("dual_value", 4), true), Rep(D("reduced_cost", 6), true),
                         {"solver_specific_info", 11, FT::kBytes}}};
  return s;
}

// ---------------------------------------------------------------------------
// number / string formatting
// ---------------------------------------------------------------------------
std::string RoundTripDouble(double v) {
  if (std::isnan(v)) return "nan";
  if (std::isinf(v)) return v > 0 ? "inf" : "-inf";
  char buf[40];
  const std::to_chars_result r = std::to_chars(buf, buf + sizeof buf, v);  // the shortest text that reads back to v
  return std::string(buf, r.ptr);
}

namespace {

std::string EscapeBytes(std::string_view s) {  // CEscape, as TextFormat prints strings
  std::string out;
  for (unsigned char c : s) {
    switch (c) {
      case '\n': out += "\\n"; break;
      case '\r': out += "\\r"; break;
      case '\t': out += "\\t"; break;
      case '"': out += "\\\""; break;
      case '\'': out += "\\'"; break;
      case '\\': out += "\\\\"; break;
      default:
        if (c < 0x20 || c >= 0x7F) {
          char b[8];
          std::snprintf(b, sizeof b, "\\%03o", c);
          out += b;
        } else {
          out.push_back(static_cast<char>(c));
        }
    }
  }
  return out;
}

std::string JsonString(std::string_view s) {
  std::string out = "\"";
  for (unsigned char c : s) {
    switch (c) {
      case '"': out += "\\\""; break;
      case '\\': out += "\\\\"; break;
      case '\n': out += "\\n"; break;
      case '\r': out += "\\r"; break;
      case '\t': out += "\\t"; break;
      default:
        if (c < 0x20) {
          char b[8];
          std::snprintf(b, sizeof b, "\\u%04x", c);
          out += b;
        } else {
          out.push_back(static_cast<char>(c));
        }
    }
  }
  out.push_back('"');
  return out;
}

const char kB64[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
std::string Base64(std::string_view s) {
  std::string out;
  size_t i = 0;
  for (; i + 2 < s.size(); i += 3) {
    const uint32_t v = (static_cast<uint8_t>(s[i]) << 16) | (static_cast<uint8_t>(s[i + 1]) << 8) | static_cast<uint8_t>(s[i + 2]);
    out += {kB64[v >> 18], kB64[(v >> 12) & 63], kB64[(v >> 6) & 63], kB64[v & 63]};
  }
  if (i + 1 == s.size()) {
    const uint32_t v = static_cast<uint8_t>(s[i]) << 16;
    out += {kB64[v >> 18], kB64[(v >> 12) & 63], '=', '='};
  } else if (i + 2 == s.size()) {
    const uint32_t v = (static_cast<uint8_t>(s[i]) << 16) | (static_cast<uint8_t>(s[i + 1]) << 8);
    out += {kB64[v >> 18], kB64[(v >> 12) & 63], kB64[(v >> 6) & 63], '='};
  }
  return out;
}
bool Base64Decode(std::string_view s, std::string* out) {
  uint32_t acc = 0;
  int bits = 0;
  for (char c : s) {
    if (c == '=') break;
    int v;
    if (c >= 'A' && c <= 'Z') v = c - 'A';
    else if (c >= 'a' && c <= 'z') v = c - 'a' + 26;
    else if (c >= '0' && c <= '9') v = c - '0' + 52;
    else if (c == '+' || c == '-') v = 62;
    else if (c == '/' || c == '_') v = 63;
    else return false;
    acc = (acc << 6) | static_cast<uint32_t>(v);
    bits += 6;
    if (bits >= 8) {
      bits -= 8;
      out->push_back(static_cast<char>((acc >> bits) & 0xFF));
    }
  }
  return true;
}

std::string CamelCase(const char* name) {  // proto3 JSON name
  std::string out;
  bool up = false;
  for (const char* p = name; *p; ++p) {
    if (*p == '_') {
      up = true;
    } else if (up) {
      out.push_back(static_cast<char>(std::toupper(static_cast<unsigned char>(*p))));
      up = false;
    } else {
      out.push_back(*p);
    }
  }
  return out;
}

// One scalar value read from text / JSON, written to the wire.
void WriteScalar(Writer* w, const FieldDef& f, double d, int64_t i, std::string_view s) {
  switch (f.type) {
    case FT::kDouble: w->Double(f.number, d); break;
    case FT::kInt32: case FT::kInt64: case FT::kEnum: case FT::kBool: w->Int(f.number, i); break;
    case FT::kString: case FT::kBytes: w->Bytes(f.number, s); break;
    case FT::kMessage: break;
  }
}

bool ParseDoubleToken(std::string tok, double* out) {
  if (tok.empty()) return false;
  std::string low;
  for (char c : tok) low.push_back(static_cast<char>(std::tolower(static_cast<unsigned char>(c))));
  bool neg = false;
  std::string body = low;
  if (body[0] == '-' || body[0] == '+') {
    neg = body[0] == '-';
    body = body.substr(1);
  }
  if (body == "inf" || body == "infinity") {
    *out = neg ? -std::numeric_limits<double>::infinity() : std::numeric_limits<double>::infinity();
    return true;
  }
  if (body == "nan") {
    *out = std::numeric_limits<double>::quiet_NaN();
    return true;
  }
  if (!body.empty() && body.back() == 'f') low.pop_back();  // 1.5f
  if (low.empty()) return false;
  char* end = nullptr;
  *out = std::strtod(low.c_str(), &end);
  return end != nullptr && *end == '\0' && (std::isdigit(static_cast<unsigned char>(body[0])) || body[0] == '.');
}

bool ParseIntToken(const std::string& tok, int64_t* out) {
  if (tok.empty()) return false;
  const char* s = tok.c_str();
  bool neg = false;
  if (*s == '-') {
    neg = true;
    ++s;
  }
  if (!std::isdigit(static_cast<unsigned char>(*s))) return false;
  char* end = nullptr;
  errno = 0;
  const unsigned long long v = std::strtoull(s, &end, 0);  // decimal, 0x hex, 0 octal
  if (errno != 0 || end == nullptr || *end != '\0') return false;
  if (neg) {
    if (v > static_cast<unsigned long long>(std::numeric_limits<int64_t>::max()) + 1ull) return false;
    *out = static_cast<int64_t>(0ull - v);
  } else {
    if (v > static_cast<unsigned long long>(std::numeric_limits<int64_t>::max())) return false;
    *out = static_cast<int64_t>(v);
  }
  return true;
}

bool InInt32(int64_t v) { return v >= std::numeric_limits<int32_t>::min() && v <= std::numeric_limits<int32_t>::max(); }

// ---------------------------------------------------------------------------
// text format parser
// ---------------------------------------------------------------------------
class TextParser {
 public:
  TextParser(std::string_view text, std::string* error, bool allow_singular_overwrites)
      : p_(text.data()), begin_(text.data()), end_(text.data() + text.size()), error_(error), strict_(!allow_singular_overwrites) {}

  bool ParseMessage(const Schema& schema, char terminator, Writer* w) {
    std::vector<const FieldDef*> seen;  // non-repeated fields of this message given so far (Parse policy only)
    for (;;) {
      Skip();
      if (p_ >= end_) {
        if (terminator == 0) return true;
        return Fail(std::string("expected '") + terminator + "' before the end of the text");
      }
      if (terminator != 0 && *p_ == terminator) {
        ++p_;
        return true;
      }
      if (*p_ == '[') return Fail("extensions and Any are not supported");
      std::string name = Identifier();
      if (name.empty()) return Fail("expected a field name");
      const FieldDef* f = schema.ByName(name);
      if (f == nullptr) return Fail("message " + std::string(schema.name) + " has no field named \"" + name + "\"");
      if (strict_ && !f->repeated) {
        for (const FieldDef* g : seen) {
          if (g == f) return Fail("non-repeated field \"" + name + "\" is specified multiple times");
          if (f->oneof != 0 && g->oneof == f->oneof)
            return Fail("field \"" + name + "\" is specified along with field \"" + g->name + "\", another member of the same oneof");
        }
        seen.push_back(f);
      }
      Skip();
      bool colon = false;
      if (p_ < end_ && *p_ == ':') {
        colon = true;
        ++p_;
        Skip();
      }
      if (f->type == FT::kMessage) {
        if (p_ < end_ && *p_ == '[') {
          if (!f->repeated) return Fail("field " + name + " is not repeated");
          ++p_;
          Skip();
          if (p_ < end_ && *p_ == ']') {
            ++p_;
          } else {
            for (;;) {
              if (!SubMessage(*f, w)) return false;
              Skip();
              if (p_ < end_ && *p_ == ',') {
                ++p_;
                Skip();
                continue;
              }
              if (p_ < end_ && *p_ == ']') {
                ++p_;
                break;
              }
              return Fail("expected ',' or ']'");
            }
          }
        } else if (!SubMessage(*f, w)) {
          return false;
        }
      } else {
        if (!colon) return Fail("expected ':' after " + name);
        if (p_ < end_ && *p_ == '[') {
          if (!f->repeated) return Fail("field " + name + " is not repeated");
          ++p_;
          Skip();
          if (p_ < end_ && *p_ == ']') {
            ++p_;
          } else {
            for (;;) {
              if (!Scalar(*f, w)) return false;
              Skip();
              if (p_ < end_ && *p_ == ',') {
                ++p_;
                Skip();
                continue;
              }
              if (p_ < end_ && *p_ == ']') {
                ++p_;
                break;
              }
              return Fail("expected ',' or ']'");
            }
          }
        } else if (!Scalar(*f, w)) {
          return false;
        }
      }
      Skip();
      if (p_ < end_ && (*p_ == ',' || *p_ == ';')) ++p_;
    }
  }

 private:
  bool Fail(const std::string& what) {
    if (error_ != nullptr && error_->empty()) {
      int line = 1, col = 1;
      for (const char* q = begin_; q < p_ && q < end_; ++q) {
        if (*q == '\n') {
          ++line;
          col = 1;
        } else {
          ++col;
        }
      }
      *error_ = std::to_string(line) + ":" + std::to_string(col) + ": " + what;
    }
    return false;
  }
  void Skip() {
    while (p_ < end_) {
      if (std::isspace(static_cast<unsigned char>(*p_))) {
        ++p_;
      } else if (*p_ == '#') {
        while (p_ < end_ && *p_ != '\n') ++p_;
      } else {
        break;
      }
    }
  }
  std::string Identifier() {
    const char* s = p_;
    while (p_ < end_ && (std::isalnum(static_cast<unsigned char>(*p_)) || *p_ == '_')) ++p_;
    return std::string(s, p_ - s);
  }
  std::string NumberOrWord() {
    const char* s = p_;
    while (p_ < end_ && (std::isalnum(static_cast<unsigned char>(*p_)) || *p_ == '_' || *p_ == '.' || *p_ == '+' || *p_ == '-')) ++p_;
    return std::string(s, p_ - s);
  }
  bool SubMessage(const FieldDef& f, Writer* w) {
    Skip();
    if (p_ >= end_ || (*p_ != '{' && *p_ != '<')) return Fail(std::string("expected '{' after ") + f.name);
    const char close = *p_ == '{' ? '}' : '>';
    ++p_;
    Writer sub;
    if (!ParseMessage(*f.message, close, &sub)) return false;
    w->Bytes(f.number, sub.out());
    return true;
  }
  bool QuotedString(std::string* out) {
    bool any = false;
    for (;;) {
      Skip();
      if (p_ >= end_ || (*p_ != '"' && *p_ != '\'')) break;
      const char q = *p_++;
      any = true;
      while (p_ < end_ && *p_ != q) {
        char c = *p_++;
        if (c == '\n') return Fail("newline inside a string");
        if (c != '\\') {
          out->push_back(c);
          continue;
        }
        if (p_ >= end_) return Fail("unterminated string");
        c = *p_++;
        switch (c) {
          case 'n': out->push_back('\n'); break;
          case 'r': out->push_back('\r'); break;
          case 't': out->push_back('\t'); break;
          case 'a': out->push_back('\a'); break;
          case 'b': out->push_back('\b'); break;
          case 'f': out->push_back('\f'); break;
          case 'v': out->push_back('\v'); break;
          case '\\': case '\'': case '"': case '?': out->push_back(c); break;
          case 'x': case 'X': {
            int v = 0, n = 0;
            while (p_ < end_ && n < 2 && std::isxdigit(static_cast<unsigned char>(*p_))) {
              const char h = *p_++;
              v = v * 16 + (std::isdigit(static_cast<unsigned char>(h)) ? h - '0' : std::tolower(h) - 'a' + 10);
              ++n;
            }
            if (n == 0) return Fail("bad \\x escape");
            out->push_back(static_cast<char>(v));
            break;
          }
          default:
            if (c >= '0' && c <= '7') {
              int v = c - '0', n = 1;
              while (p_ < end_ && n < 3 && *p_ >= '0' && *p_ <= '7') {
                v = v * 8 + (*p_++ - '0');
                ++n;
              }
              out->push_back(static_cast<char>(v));
            } else {
              return Fail("unknown escape in string");
            }
        }
      }
      if (p_ >= end_) return Fail("unterminated string");
      ++p_;
    }
    if (!any) return Fail("expected a quoted string");
    return true;
  }
  bool Scalar(const FieldDef& f, Writer* w) {
    Skip();
    if (f.type == FT::kString || f.type == FT::kBytes) {
      std::string s;
      if (!QuotedString(&s)) return false;
      WriteScalar(w, f, 0, 0, s);
      return true;
    }
    const std::string tok = NumberOrWord();
    if (tok.empty()) return Fail(std::string("expected a value for ") + f.name);
    switch (f.type) {
      case FT::kDouble: {
        double d;
        if (!ParseDoubleToken(tok, &d)) return Fail("expected a number for " + std::string(f.name) + ", got \"" + tok + "\"");
        WriteScalar(w, f, d, 0, {});
        return true;
      }
      case FT::kInt32: case FT::kInt64: {
        int64_t v;
        if (!ParseIntToken(tok, &v) || (f.type == FT::kInt32 && !InInt32(v)))
          return Fail("expected an integer for " + std::string(f.name) + ", got \"" + tok + "\"");
        WriteScalar(w, f, 0, v, {});
        return true;
      }
      case FT::kBool: {
        int64_t v;
        if (tok == "true" || tok == "True" || tok == "t" || tok == "1") v = 1;
        else if (tok == "false" || tok == "False" || tok == "f" || tok == "0") v = 0;
        else return Fail("expected true or false for " + std::string(f.name) + ", got \"" + tok + "\"");
        WriteScalar(w, f, 0, v, {});
        return true;
      }
      case FT::kEnum: {
        int number;
        int64_t v;
        if (f.enumeration->NumberOf(tok, &number)) v = number;
        else if (!ParseIntToken(tok, &v) || !InInt32(v) || f.enumeration->NameOf(static_cast<int>(v)) == nullptr)
          return Fail("unknown value \"" + tok + "\" for enum " + std::string(f.enumeration->name));
        WriteScalar(w, f, 0, v, {});
        return true;
      }
      default:
        return Fail("unsupported field type");
    }
  }

  const char* p_;
  const char* begin_;
  const char* end_;
  std::string* error_;
  bool strict_;
};

// ---------------------------------------------------------------------------
// wire -> text / JSON
// ---------------------------------------------------------------------------
struct Decoded {
  const FieldDef* def;
  WireField wire;
};

bool DecodeKnown(const Schema& schema, std::string_view bytes, std::vector<Decoded>* out) {
  Reader r(bytes);
  WireField f;
  while (r.Next(&f)) {
    const FieldDef* def = schema.ByNumber(f.number);
    if (def == nullptr) continue;
    if (def->oneof != 0)  // setting a member of a oneof clears the member that was set before it
      out->erase(std::remove_if(out->begin(), out->end(), [&](const Decoded& d) { return d.def->oneof == def->oneof && d.def != def; }), out->end());
    out->push_back({def, f});
  }
  if (!r.ok()) return false;
  std::stable_sort(out->begin(), out->end(), [](const Decoded& a, const Decoded& b) { return a.def->number < b.def->number; });
  return true;
}

// Expands one wire field into scalar texts (packed fields give several).
template <typename Emit>
bool ForEachScalar(const Decoded& d, bool json, Emit emit) {
  const FieldDef& f = *d.def;
  auto dbl = [&](double v) -> std::string {
    if (!json) return RoundTripDouble(v);
    if (std::isnan(v)) return "\"NaN\"";
    if (std::isinf(v)) return v > 0 ? "\"Infinity\"" : "\"-Infinity\"";
    if (v == 0.0 && std::signbit(v)) return "-0.0";  // "-0" is an integer to most JSON readers and loses the sign
    return RoundTripDouble(v);
  };
  auto integer = [&](int64_t v) -> std::string {
    switch (f.type) {
      case FT::kBool: return v ? "true" : "false";
      case FT::kEnum: {
        const char* n = f.enumeration->NameOf(static_cast<int>(v));
        if (n == nullptr) return std::to_string(v);
        return json ? JsonString(n) : std::string(n);
      }
      case FT::kInt64: return json ? "\"" + std::to_string(v) + "\"" : std::to_string(v);
      case FT::kInt32: return std::to_string(static_cast<int32_t>(v));
      default: return std::to_string(v);
    }
  };
  switch (f.type) {
    case FT::kDouble: {
      std::vector<double> v;
      if (!AppendDoubles(d.wire, &v)) return false;
      for (double x : v) emit(dbl(x));
      return true;
    }
    case FT::kInt32: case FT::kEnum: case FT::kBool: case FT::kInt64: {
      if (d.wire.type == kVarint) {
        emit(integer(d.wire.AsInt64()));
        return true;
      }
      if (d.wire.type != kLengthDelimited) return false;
      return ForEachPackedVarint(d.wire.bytes, [&](uint64_t v) { emit(integer(static_cast<int64_t>(v))); });
    }
    case FT::kString:
      if (d.wire.type != kLengthDelimited) return false;
      emit(json ? JsonString(d.wire.bytes) : "\"" + EscapeBytes(d.wire.bytes) + "\"");
      return true;
    case FT::kBytes:
      if (d.wire.type != kLengthDelimited) return false;
      emit(json ? "\"" + Base64(d.wire.bytes) + "\"" : "\"" + EscapeBytes(d.wire.bytes) + "\"");
      return true;
    case FT::kMessage:
      return false;
  }
  return false;
}

bool TextOut(const Schema& schema, std::string_view bytes, int indent, std::string* out) {
  std::vector<Decoded> fields;
  if (!DecodeKnown(schema, bytes, &fields)) return false;
  const std::string pad(static_cast<size_t>(indent) * 2, ' ');
  for (const Decoded& d : fields) {
    if (d.def->type == FT::kMessage) {
      if (d.wire.type != kLengthDelimited) return false;
      *out += pad + d.def->name + " {\n";
      if (!TextOut(*d.def->message, d.wire.bytes, indent + 1, out)) return false;
      *out += pad + "}\n";
    } else if (!ForEachScalar(d, false, [&](const std::string& v) { *out += pad + d.def->name + ": " + v + "\n"; })) {
      return false;
    }
  }
  return true;
}

bool JsonOut(const Schema& schema, std::string_view bytes, int indent, std::string* out) {
  std::vector<Decoded> fields;
  if (!DecodeKnown(schema, bytes, &fields)) return false;
  if (fields.empty()) {
    *out += "{}";
    return true;
  }
  const std::string pad(static_cast<size_t>(indent + 1), ' ');
  *out += "{\n";
  size_t i = 0;
  bool first = true;
  while (i < fields.size()) {
    const FieldDef* def = fields[i].def;
    size_t j = i;
    while (j < fields.size() && fields[j].def == def) ++j;
    // a non-repeated field seen several times: the last value wins (scalars)
    std::vector<std::string> values;
    for (size_t k = def->repeated ? i : j - 1; k < j; ++k) {
      if (def->type == FT::kMessage) {
        if (fields[k].wire.type != kLengthDelimited) return false;
        std::string sub;
        if (!JsonOut(*def->message, fields[k].wire.bytes, indent + (def->repeated ? 2 : 1), &sub)) return false;
        values.push_back(sub);
      } else if (!ForEachScalar(fields[k], true, [&](const std::string& v) { values.push_back(v); })) {
        return false;
      }
    }
    if (values.empty() && !def->repeated) {  // e.g. a scalar field that arrived as an empty packed run
      i = j;
      continue;
    }
    if (!first) *out += ",\n";
    first = false;
    *out += pad + JsonString(CamelCase(def->name)) + ": ";
    if (def->repeated) {
      const std::string pad2(static_cast<size_t>(indent + 2), ' ');
      *out += "[\n";
      for (size_t k = 0; k < values.size(); ++k) *out += pad2 + values[k] + (k + 1 < values.size() ? ",\n" : "\n");
      *out += pad + "]";
    } else {
      *out += values.back();
    }
    i = j;
  }
  *out += "\n" + std::string(static_cast<size_t>(indent), ' ') + "}";
  return true;
}

// ---------------------------------------------------------------------------
// JSON parser
// ---------------------------------------------------------------------------
class JsonParser {
 public:
  JsonParser(std::string_view text, std::string* error) : p_(text.data()), begin_(text.data()), end_(text.data() + text.size()), error_(error) {}

  bool ParseObject(const Schema& schema, Writer* w) {
    Skip();
    if (p_ >= end_ || *p_ != '{') return Fail("expected '{'");
    ++p_;
    Skip();
    if (p_ < end_ && *p_ == '}') {
      ++p_;
      return true;
    }
    for (;;) {
      Skip();
      std::string key;
      if (!String(&key)) return false;
      Skip();
      if (p_ >= end_ || *p_ != ':') return Fail("expected ':'");
      ++p_;
      const FieldDef* f = schema.ByName(key);
      if (f == nullptr)
        for (const auto& cand : schema.fields)
          if (CamelCase(cand.name) == key) f = &cand;
      if (f == nullptr) return Fail("message " + std::string(schema.name) + " has no field named \"" + key + "\"");
      Skip();
      if (p_ + 4 <= end_ && std::strncmp(p_, "null", 4) == 0) {
        p_ += 4;
      } else if (f->repeated) {
        if (p_ >= end_ || *p_ != '[') return Fail("expected '[' for repeated field " + key);
        ++p_;
        Skip();
        if (p_ < end_ && *p_ == ']') {
          ++p_;
        } else {
          for (;;) {
            if (!Value(*f, w)) return false;
            Skip();
            if (p_ < end_ && *p_ == ',') {
              ++p_;
              continue;
            }
            if (p_ < end_ && *p_ == ']') {
              ++p_;
              break;
            }
            return Fail("expected ',' or ']'");
          }
        }
      } else if (!Value(*f, w)) {
        return false;
      }
      Skip();
      if (p_ < end_ && *p_ == ',') {
        ++p_;
        continue;
      }
      if (p_ < end_ && *p_ == '}') {
        ++p_;
        return true;
      }
      return Fail("expected ',' or '}'");
    }
  }
  bool AtEnd() {
    Skip();
    return p_ >= end_;
  }

 private:
  bool Fail(const std::string& what) {
    if (error_ != nullptr && error_->empty()) *error_ = "JSON offset " + std::to_string(p_ - begin_) + ": " + what;
    return false;
  }
  void Skip() {
    while (p_ < end_ && std::isspace(static_cast<unsigned char>(*p_))) ++p_;
  }
  bool Hex4(unsigned* v) {
    if (end_ - p_ < 4) return false;
    unsigned r = 0;
    for (int k = 0; k < 4; ++k) {
      const char h = p_[k];
      if (!std::isxdigit(static_cast<unsigned char>(h))) return false;
      r = r * 16 + static_cast<unsigned>(std::isdigit(static_cast<unsigned char>(h)) ? h - '0' : std::tolower(static_cast<unsigned char>(h)) - 'a' + 10);
    }
    p_ += 4;
    *v = r;
    return true;
  }
  bool String(std::string* out) {
    if (p_ >= end_ || *p_ != '"') return Fail("expected a string");
    ++p_;
    while (p_ < end_ && *p_ != '"') {
      char c = *p_++;
      if (c != '\\') {
        out->push_back(c);
        continue;
      }
      if (p_ >= end_) return Fail("unterminated string");
      c = *p_++;
      switch (c) {
        case 'n': out->push_back('\n'); break;
        case 'r': out->push_back('\r'); break;
        case 't': out->push_back('\t'); break;
        case 'b': out->push_back('\b'); break;
        case 'f': out->push_back('\f'); break;
        case '/': case '\\': case '"': out->push_back(c); break;
        case 'u': {
          unsigned v = 0;
          if (!Hex4(&v)) return Fail("bad \\u escape");
          if (v >= 0xD800 && v <= 0xDBFF && end_ - p_ >= 6 && p_[0] == '\\' && p_[1] == 'u') {  // surrogate pair
            const char* save = p_;
            p_ += 2;
            unsigned lo = 0;
            if (Hex4(&lo) && lo >= 0xDC00 && lo <= 0xDFFF) v = 0x10000 + ((v - 0xD800) << 10) + (lo - 0xDC00);
            else p_ = save;
          }
          if (v < 0x80) {
            out->push_back(static_cast<char>(v));
          } else if (v < 0x800) {
            out->push_back(static_cast<char>(0xC0 | (v >> 6)));
            out->push_back(static_cast<char>(0x80 | (v & 0x3F)));
          } else if (v < 0x10000) {
            out->push_back(static_cast<char>(0xE0 | (v >> 12)));
            out->push_back(static_cast<char>(0x80 | ((v >> 6) & 0x3F)));
            out->push_back(static_cast<char>(0x80 | (v & 0x3F)));
          } else {
            out->push_back(static_cast<char>(0xF0 | (v >> 18)));
            out->push_back(static_cast<char>(0x80 | ((v >> 12) & 0x3F)));
            out->push_back(static_cast<char>(0x80 | ((v >> 6) & 0x3F)));
            out->push_back(static_cast<char>(0x80 | (v & 0x3F)));
          }
          break;
        }
        default: return Fail("unknown escape");
      }
    }
    if (p_ >= end_) return Fail("unterminated string");
    ++p_;
    return true;
  }
  bool Value(const FieldDef& f, Writer* w) {
    Skip();
    if (f.type == FT::kMessage) {
      Writer sub;
      if (!ParseObject(*f.message, &sub)) return false;
      w->Bytes(f.number, sub.out());
      return true;
    }
    std::string tok;
    bool quoted = false;
    if (p_ < end_ && *p_ == '"') {
      quoted = true;
      if (!String(&tok)) return false;
    } else {
      const char* s = p_;
      while (p_ < end_ && (std::isalnum(static_cast<unsigned char>(*p_)) || *p_ == '.' || *p_ == '+' || *p_ == '-')) ++p_;
      tok.assign(s, p_ - s);
      if (tok.empty()) return Fail(std::string("expected a value for ") + f.name);
    }
    switch (f.type) {
      case FT::kString:
        if (!quoted) return Fail(std::string("expected a string for ") + f.name);
        WriteScalar(w, f, 0, 0, tok);
        return true;
      case FT::kBytes: {
        std::string raw;
        if (!quoted || !Base64Decode(tok, &raw)) return Fail(std::string("expected base64 for ") + f.name);
        WriteScalar(w, f, 0, 0, raw);
        return true;
      }
      case FT::kDouble: {
        double d;
        if (!ParseDoubleToken(tok, &d)) return Fail("expected a number for " + std::string(f.name));
        WriteScalar(w, f, d, 0, {});
        return true;
      }
      case FT::kInt32: case FT::kInt64: {
        int64_t v;
        double d;
        if (!ParseIntToken(tok, &v)) {  // 1.0 / 1e3 are legal JSON integers
          if (!ParseDoubleToken(tok, &d) || d != std::floor(d) || std::fabs(d) > 9.2e18) return Fail("expected an integer for " + std::string(f.name));
          v = static_cast<int64_t>(d);
        }
        if (f.type == FT::kInt32 && !InInt32(v)) return Fail("integer out of range for " + std::string(f.name));
        WriteScalar(w, f, 0, v, {});
        return true;
      }
      case FT::kBool:
        if (tok == "true") WriteScalar(w, f, 0, 1, {});
        else if (tok == "false") WriteScalar(w, f, 0, 0, {});
        else return Fail("expected true or false for " + std::string(f.name));
        return true;
      case FT::kEnum: {
        int number;
        int64_t v;
        if (f.enumeration->NumberOf(tok, &number)) v = number;
        else if (!ParseIntToken(tok, &v) || !InInt32(v)) return Fail("unknown enum value \"" + tok + "\"");
        WriteScalar(w, f, 0, v, {});
        return true;
      }
      default:
        return Fail("unsupported field type");
    }
  }

  const char* p_;
  const char* begin_;
  const char* end_;
  std::string* error_;
};

// Rewrites a message the way protobuf serialises it: fields in tag order, a
// non-repeated scalar keeps its last value, the occurrences of a non-repeated
// message are merged, packed repeated fields become one packed run. Unknown
// fields are dropped.
bool Canonical(const Schema& schema, std::string_view bytes, Writer* w) {
  std::vector<Decoded> fields;
  if (!DecodeKnown(schema, bytes, &fields)) return false;
  size_t i = 0;
  while (i < fields.size()) {
    const FieldDef* def = fields[i].def;
    size_t j = i;
    while (j < fields.size() && fields[j].def == def) ++j;
    if (def->type == FT::kMessage) {
      if (def->repeated) {
        for (size_t k = i; k < j; ++k) {
          if (fields[k].wire.type != kLengthDelimited) return false;
          Writer sub;
          if (!Canonical(*def->message, fields[k].wire.bytes, &sub)) return false;
          w->Bytes(def->number, sub.out());
        }
      } else {
        std::string merged;
        for (size_t k = i; k < j; ++k) {
          if (fields[k].wire.type != kLengthDelimited) return false;
          merged.append(fields[k].wire.bytes);
        }
        Writer sub;
        if (!Canonical(*def->message, merged, &sub)) return false;
        w->Bytes(def->number, sub.out());
      }
    } else if (def->type == FT::kString || def->type == FT::kBytes) {
      for (size_t k = def->repeated ? i : j - 1; k < j; ++k) {
        if (fields[k].wire.type != kLengthDelimited) return false;
        w->Bytes(def->number, fields[k].wire.bytes);
      }
    } else if (def->type == FT::kDouble) {
      std::vector<double> v;
      for (size_t k = def->repeated ? i : j - 1; k < j; ++k)
        if (!AppendDoubles(fields[k].wire, &v)) return false;
      if (def->repeated && def->packed) w->PackedDoubles(def->number, v.data(), static_cast<int64_t>(v.size()));
      else for (double x : v) w->Double(def->number, x);
    } else {
      std::vector<int64_t> v;
      for (size_t k = def->repeated ? i : j - 1; k < j; ++k) {
        if (fields[k].wire.type == kVarint) v.push_back(fields[k].wire.AsInt64());
        else if (fields[k].wire.type != kLengthDelimited ||
                 !ForEachPackedVarint(fields[k].wire.bytes, [&](uint64_t x) { v.push_back(static_cast<int64_t>(x)); }))
          return false;
      }
      if (def->repeated && def->packed) {
        if (!v.empty()) {
          Writer body;
          for (int64_t x : v) body.Varint(static_cast<uint64_t>(x));
          w->Bytes(def->number, body.out());
        }
      } else {
        for (int64_t x : v) w->Int(def->number, x);
      }
    }
    i = j;
  }
  return true;
}

}  // namespace

bool TextToWire(const Schema& schema, std::string_view text, std::string* out, std::string* error, bool allow_singular_overwrites) {
  std::string local;
  std::string* err = error != nullptr ? error : &local;
  err->clear();
  Writer w;
  TextParser parser(text, err, allow_singular_overwrites);
  if (!parser.ParseMessage(schema, 0, &w)) return false;
  Writer canonical;
  if (!Canonical(schema, w.out(), &canonical)) return false;
  out->append(canonical.out());
  return true;
}

bool WireToText(const Schema& schema, std::string_view bytes, std::string* out) { return TextOut(schema, bytes, 0, out); }

bool WireToJson(const Schema& schema, std::string_view bytes, std::string* out) {
  if (!JsonOut(schema, bytes, 0, out)) return false;
  out->push_back('\n');
  return true;
}

bool JsonToWire(const Schema& schema, std::string_view json, std::string* out, std::string* error) {
  std::string local;
  std::string* err = error != nullptr ? error : &local;
  err->clear();
  Writer w;
  JsonParser parser(json, err);
  if (!parser.ParseObject(schema, &w)) return false;
  if (!parser.AtEnd()) {
    *err = "trailing characters after the JSON object";
    return false;
  }
  Writer canonical;
  if (!Canonical(schema, w.out(), &canonical)) return false;
  out->append(canonical.out());
  return true;
}

}  // namespace proto
}  // namespace pdlp_b200
