// formats.cc -- the data formats either side of the PDLP path (C ABI:
// include/pdlp_b200_io.h): parameter / log protos, MPModelProto <->
// QuadraticProgram, the MPS reader / writer and PdlpSolveProto. Host-only code
// on top of the proto2 codec of proto_codec.cc; the solve itself goes through
// pdlp_b200_primal_dual_hybrid_gradient (no CPU fallback).
#include <dlfcn.h>
#include <zlib.h>

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/pdlp_b200_io.h"
#include "proto_codec.h"
#include "solver.h"

namespace pdlp_b200 {
namespace {

using proto::Reader;
using proto::WireField;
using proto::Writer;
constexpr double kInf = std::numeric_limits<double>::infinity();

void SetError(char* error, int64_t capacity, const std::string& message) {
  if (error == nullptr || capacity <= 0) return;
  const size_t n = std::min<size_t>(message.size(), static_cast<size_t>(capacity - 1));
  std::memcpy(error, message.data(), n);
  error[n] = '\0';
}

int32_t BadArgument(char* error, int64_t capacity, const std::string& message) {
  SetError(error, capacity, message);
  return PDLP_B200_STATUS_BAD_ARGUMENT;
}

bool EndsWith(const std::string& s, const char* suffix) {
  const size_t n = std::strlen(suffix);
  return s.size() >= n && s.compare(s.size() - n, n, suffix) == 0;
}

int32_t ToBlob(const std::string& bytes, PdlpBlob* out) {
  out->data = static_cast<uint8_t*>(std::malloc(bytes.size() + 1));
  if (out->data == nullptr) return PDLP_B200_STATUS_BAD_ARGUMENT;
  if (!bytes.empty()) std::memcpy(out->data, bytes.data(), bytes.size());
  out->data[bytes.size()] = 0;
  out->size = static_cast<int64_t>(bytes.size());
  return PDLP_B200_STATUS_OK;
}

// ---------------------------------------------------------------------------
// PrimalDualHybridGradientParams <-> wire (solvers.proto:66-497)
// ---------------------------------------------------------------------------
bool IsDouble(const WireField& f) { return f.type == proto::kFixed64; }
bool IsVarint(const WireField& f) { return f.type == proto::kVarint; }
bool IsBytes(const WireField& f) { return f.type == proto::kLengthDelimited; }

bool MergeSimpleCriteria(std::string_view bytes, PdlpTerminationCriteria* tc) {
  Reader r(bytes);
  WireField f;
  while (r.Next(&f)) {
    if (!IsDouble(f)) continue;
    if (f.number == 1) tc->simple_eps_optimal_absolute = f.AsDouble();
    if (f.number == 2) tc->simple_eps_optimal_relative = f.AsDouble();
  }
  return r.ok();
}

bool MergeDetailedCriteria(std::string_view bytes, PdlpTerminationCriteria* tc) {
  Reader r(bytes);
  WireField f;
  double* slots[] = {nullptr, &tc->eps_optimal_primal_residual_absolute, &tc->eps_optimal_primal_residual_relative,
                     &tc->eps_optimal_dual_residual_absolute, &tc->eps_optimal_dual_residual_relative,
                     &tc->eps_optimal_objective_gap_absolute, &tc->eps_optimal_objective_gap_relative};
  while (r.Next(&f))
    if (IsDouble(f) && f.number >= 1 && f.number <= 6) *slots[f.number] = f.AsDouble();
  return r.ok();
}

// Both members of the optimality_criteria oneof back to their defaults.
void ResetOptimalityCriteria(PdlpTerminationCriteria* tc) {
  tc->simple_eps_optimal_absolute = tc->simple_eps_optimal_relative = 1.0e-6;
  tc->eps_optimal_primal_residual_absolute = tc->eps_optimal_primal_residual_relative = 1.0e-6;
  tc->eps_optimal_dual_residual_absolute = tc->eps_optimal_dual_residual_relative = 1.0e-6;
  tc->eps_optimal_objective_gap_absolute = tc->eps_optimal_objective_gap_relative = 1.0e-6;
}

bool MergeTerminationCriteria(std::string_view bytes, PdlpTerminationCriteria* tc) {
  Reader r(bytes);
  WireField f;
  while (r.Next(&f)) {
    switch (f.number) {
      case 1: if (IsVarint(f)) tc->optimality_norm = f.AsInt32(); break;
      case 9:
        if (!IsBytes(f)) break;
        if (tc->optimality_criteria_case != PDLP_SIMPLE_OPTIMALITY_CRITERIA) {  // setting a oneof member clears the other
          ResetOptimalityCriteria(tc);
          tc->optimality_criteria_case = PDLP_SIMPLE_OPTIMALITY_CRITERIA;
        }
        if (!MergeSimpleCriteria(f.bytes, tc)) return false;
        break;
      case 10:
        if (!IsBytes(f)) break;
        if (tc->optimality_criteria_case != PDLP_DETAILED_OPTIMALITY_CRITERIA) {
          ResetOptimalityCriteria(tc);
          tc->optimality_criteria_case = PDLP_DETAILED_OPTIMALITY_CRITERIA;
        }
        if (!MergeDetailedCriteria(f.bytes, tc)) return false;
        break;
      case 2: if (IsDouble(f)) { tc->eps_optimal_absolute = f.AsDouble(); tc->has_eps_optimal_absolute = 1; } break;
      case 3: if (IsDouble(f)) { tc->eps_optimal_relative = f.AsDouble(); tc->has_eps_optimal_relative = 1; } break;
      case 4: if (IsDouble(f)) tc->eps_primal_infeasible = f.AsDouble(); break;
      case 5: if (IsDouble(f)) tc->eps_dual_infeasible = f.AsDouble(); break;
      case 6: if (IsDouble(f)) tc->time_sec_limit = f.AsDouble(); break;
      case 7: if (IsVarint(f)) tc->iteration_limit = f.AsInt32(); break;
      case 8: if (IsDouble(f)) tc->kkt_matrix_pass_limit = f.AsDouble(); break;
      default: break;
    }
  }
  return r.ok();
}

bool MergeParams(std::string_view bytes, PdlpParams* p, std::string* error) {
  Reader r(bytes);
  WireField f;
  std::vector<int32_t> seeds;
  bool seeds_seen = false;
  auto sub_double = [&](std::string_view sub, int n, double* const* slots) {
    Reader s(sub);
    WireField g;
    while (s.Next(&g))
      if (IsDouble(g) && g.number >= 1 && g.number <= n) *slots[g.number - 1] = g.AsDouble();
    return s.ok();
  };
  while (r.Next(&f)) {
    switch (f.number) {
      case 1: if (IsBytes(f) && !MergeTerminationCriteria(f.bytes, &p->termination_criteria)) { *error = "malformed termination_criteria"; return false; } break;
      case 2: if (IsVarint(f)) p->num_threads = f.AsInt32(); break;
      case 27: if (IsVarint(f)) p->num_shards = f.AsInt32(); break;
      case 32: if (IsVarint(f)) p->scheduler_type = f.AsInt32(); break;
      case 3: if (IsVarint(f)) p->record_iteration_stats = f.AsBool(); break;
      case 26: if (IsVarint(f)) p->verbosity_level = f.AsInt32(); break;
      case 31: if (IsDouble(f)) p->log_interval_seconds = f.AsDouble(); break;
      case 4: if (IsVarint(f)) p->major_iteration_frequency = f.AsInt32(); break;
      case 5: if (IsVarint(f)) p->termination_check_frequency = f.AsInt32(); break;
      case 6: if (IsVarint(f)) p->restart_strategy = f.AsInt32(); break;
      case 7: if (IsDouble(f)) p->primal_weight_update_smoothing = f.AsDouble(); break;
      case 8: if (IsDouble(f)) { p->initial_primal_weight = f.AsDouble(); p->has_initial_primal_weight = 1; } break;
      case 16: {
        if (!IsBytes(f)) break;
        Reader s(f.bytes);
        WireField g;
        while (s.Next(&g))
          if (g.number == 1 && IsVarint(g)) p->presolve_use_glop = g.AsBool();  // glop_parameters (tag 2): host-only presolve, ignored
        if (!s.ok()) { *error = "malformed presolve_options"; return false; }
        break;
      }
      case 9: if (IsVarint(f)) p->l_inf_ruiz_iterations = f.AsInt32(); break;
      case 10: if (IsVarint(f)) p->l2_norm_rescaling = f.AsBool(); break;
      case 11: if (IsDouble(f)) p->sufficient_reduction_for_restart = f.AsDouble(); break;
      case 17: if (IsDouble(f)) p->necessary_reduction_for_restart = f.AsDouble(); break;
      case 12: if (IsVarint(f)) p->linesearch_rule = f.AsInt32(); break;
      case 18: {
        double* slots[] = {&p->adaptive_step_size_reduction_exponent, &p->adaptive_step_size_growth_exponent};
        if (IsBytes(f) && !sub_double(f.bytes, 2, slots)) { *error = "malformed adaptive_linesearch_parameters"; return false; }
        break;
      }
      case 19: {
        double* slots[] = {&p->malitsky_pock_step_size_downscaling_factor, &p->malitsky_pock_linesearch_contraction_factor,
                           &p->malitsky_pock_step_size_interpolation};
        if (IsBytes(f) && !sub_double(f.bytes, 3, slots)) { *error = "malformed malitsky_pock_parameters"; return false; }
        break;
      }
      case 25: if (IsDouble(f)) p->initial_step_size_scaling = f.AsDouble(); break;
      case 28:
        if (!proto::AppendInt32s(f, &seeds)) { *error = "malformed random_projection_seeds"; return false; }
        seeds_seen = true;
        break;
      case 22: if (IsDouble(f)) p->infinite_constraint_bound_threshold = f.AsDouble(); break;
      case 29: if (IsVarint(f)) p->handle_some_primal_gradients_on_finite_bounds_as_residuals = f.AsBool(); break;
      case 23: if (IsVarint(f)) p->use_diagonal_qp_trust_region_solver = f.AsBool(); break;
      case 24: if (IsDouble(f)) p->diagonal_qp_trust_region_solver_tolerance = f.AsDouble(); break;
      case 30: if (IsVarint(f)) p->use_feasibility_polishing = f.AsBool(); break;
      case 33: if (IsVarint(f)) p->apply_feasibility_polishing_after_limits_reached = f.AsBool(); break;
      case 34: if (IsVarint(f)) p->apply_feasibility_polishing_if_solver_is_interrupted = f.AsBool(); break;
      default: break;  // unknown fields are skipped, like protobuf
    }
  }
  if (!r.ok()) {
    *error = "malformed PrimalDualHybridGradientParams bytes";
    return false;
  }
  if (seeds_seen) {  // repeated fields are appended by a merge
    const int have = std::max(0, std::min<int>(p->num_random_projection_seeds, PDLP_MAX_RANDOM_PROJECTION_SEEDS));
    const int64_t total = have + static_cast<int64_t>(seeds.size());
    for (int64_t k = have; k < std::min<int64_t>(total, PDLP_MAX_RANDOM_PROJECTION_SEEDS); ++k) p->random_projection_seeds[k] = seeds[k - have];
    p->num_random_projection_seeds = static_cast<int32_t>(std::min<int64_t>(total, std::numeric_limits<int32_t>::max()));  // > 8 is rejected by the solve
  }
  return true;
}

// `differs`: NaN differs from every default.
bool Differs(double v, double d) { return !(v == d); }

void TerminationCriteriaToWire(const PdlpTerminationCriteria& tc, Writer* w) {
  if (tc.optimality_norm != PDLP_OPTIMALITY_NORM_L2) w->Int(1, tc.optimality_norm);
  if (tc.has_eps_optimal_absolute) w->Double(2, tc.eps_optimal_absolute);
  if (tc.has_eps_optimal_relative) w->Double(3, tc.eps_optimal_relative);
  if (Differs(tc.eps_primal_infeasible, 1.0e-8)) w->Double(4, tc.eps_primal_infeasible);
  if (Differs(tc.eps_dual_infeasible, 1.0e-8)) w->Double(5, tc.eps_dual_infeasible);
  if (Differs(tc.time_sec_limit, kInf)) w->Double(6, tc.time_sec_limit);
  if (tc.iteration_limit != std::numeric_limits<int32_t>::max()) w->Int(7, tc.iteration_limit);
  if (Differs(tc.kkt_matrix_pass_limit, kInf)) w->Double(8, tc.kkt_matrix_pass_limit);
  if (tc.optimality_criteria_case == PDLP_SIMPLE_OPTIMALITY_CRITERIA) {
    Writer s;
    if (Differs(tc.simple_eps_optimal_absolute, 1.0e-6)) s.Double(1, tc.simple_eps_optimal_absolute);
    if (Differs(tc.simple_eps_optimal_relative, 1.0e-6)) s.Double(2, tc.simple_eps_optimal_relative);
    w->Bytes(9, s.out());
  } else if (tc.optimality_criteria_case == PDLP_DETAILED_OPTIMALITY_CRITERIA) {
    Writer s;
    const double v[] = {tc.eps_optimal_primal_residual_absolute, tc.eps_optimal_primal_residual_relative, tc.eps_optimal_dual_residual_absolute,
                        tc.eps_optimal_dual_residual_relative, tc.eps_optimal_objective_gap_absolute, tc.eps_optimal_objective_gap_relative};
    for (int k = 0; k < 6; ++k)
      if (Differs(v[k], 1.0e-6)) s.Double(k + 1, v[k]);
    w->Bytes(10, s.out());
  }
}

void ParamsToWire(const PdlpParams& p, Writer* w) {  // fields in tag order, like protobuf serializes
  {
    Writer tc;
    TerminationCriteriaToWire(p.termination_criteria, &tc);
    if (!tc.out().empty()) w->Bytes(1, tc.out());
  }
  if (p.num_threads != 1) w->Int(2, p.num_threads);
  if (p.record_iteration_stats) w->Bool(3, true);
  if (p.major_iteration_frequency != 64) w->Int(4, p.major_iteration_frequency);
  if (p.termination_check_frequency != 64) w->Int(5, p.termination_check_frequency);
  if (p.restart_strategy != PDLP_ADAPTIVE_HEURISTIC) w->Int(6, p.restart_strategy);
  if (Differs(p.primal_weight_update_smoothing, 0.5)) w->Double(7, p.primal_weight_update_smoothing);
  if (p.has_initial_primal_weight) w->Double(8, p.initial_primal_weight);
  if (p.l_inf_ruiz_iterations != 5) w->Int(9, p.l_inf_ruiz_iterations);
  if (!p.l2_norm_rescaling) w->Bool(10, false);
  if (Differs(p.sufficient_reduction_for_restart, 0.1)) w->Double(11, p.sufficient_reduction_for_restart);
  if (p.linesearch_rule != PDLP_ADAPTIVE_LINESEARCH_RULE) w->Int(12, p.linesearch_rule);
  if (p.presolve_use_glop) {
    Writer s;
    s.Bool(1, true);
    w->Bytes(16, s.out());
  }
  if (Differs(p.necessary_reduction_for_restart, 0.9)) w->Double(17, p.necessary_reduction_for_restart);
  {
    Writer s;
    if (Differs(p.adaptive_step_size_reduction_exponent, 0.3)) s.Double(1, p.adaptive_step_size_reduction_exponent);
    if (Differs(p.adaptive_step_size_growth_exponent, 0.6)) s.Double(2, p.adaptive_step_size_growth_exponent);
    if (!s.out().empty()) w->Bytes(18, s.out());
  }
  {
    Writer s;
    if (Differs(p.malitsky_pock_step_size_downscaling_factor, 0.7)) s.Double(1, p.malitsky_pock_step_size_downscaling_factor);
    if (Differs(p.malitsky_pock_linesearch_contraction_factor, 0.99)) s.Double(2, p.malitsky_pock_linesearch_contraction_factor);
    if (Differs(p.malitsky_pock_step_size_interpolation, 1.0)) s.Double(3, p.malitsky_pock_step_size_interpolation);
    if (!s.out().empty()) w->Bytes(19, s.out());
  }
  if (Differs(p.infinite_constraint_bound_threshold, kInf)) w->Double(22, p.infinite_constraint_bound_threshold);
  if (p.use_diagonal_qp_trust_region_solver) w->Bool(23, true);
  if (Differs(p.diagonal_qp_trust_region_solver_tolerance, 1.0e-8)) w->Double(24, p.diagonal_qp_trust_region_solver_tolerance);
  if (Differs(p.initial_step_size_scaling, 1.0)) w->Double(25, p.initial_step_size_scaling);
  if (p.verbosity_level != 0) w->Int(26, p.verbosity_level);
  if (p.num_shards != 0) w->Int(27, p.num_shards);
  w->PackedInts(28, p.random_projection_seeds, std::max(0, std::min<int>(p.num_random_projection_seeds, PDLP_MAX_RANDOM_PROJECTION_SEEDS)));
  if (!p.handle_some_primal_gradients_on_finite_bounds_as_residuals) w->Bool(29, false);
  if (p.use_feasibility_polishing) w->Bool(30, true);
  if (Differs(p.log_interval_seconds, 0.0)) w->Double(31, p.log_interval_seconds);
  if (p.scheduler_type != PDLP_SCHEDULER_TYPE_GOOGLE_THREADPOOL) w->Int(32, p.scheduler_type);
  if (p.apply_feasibility_polishing_after_limits_reached) w->Bool(33, true);
  if (p.apply_feasibility_polishing_if_solver_is_interrupted) w->Bool(34, true);
}

// ---------------------------------------------------------------------------
// SolveLog -> wire (solve_log.proto:28-459)
// ---------------------------------------------------------------------------
void StatsToWire(const PdlpQuadraticProgramStats& s, Writer* w) {  // tag order
  w->Int(1, s.num_variables);
  w->Int(2, s.num_constraints);
  w->Double(3, s.constraint_matrix_col_min_l_inf_norm);
  w->Double(4, s.constraint_matrix_row_min_l_inf_norm);
  w->Int(5, s.constraint_matrix_num_nonzeros);
  w->Double(6, s.constraint_matrix_abs_max);
  w->Double(7, s.constraint_matrix_abs_min);
  w->Double(8, s.constraint_matrix_abs_avg);
  w->Double(9, s.combined_bounds_max);
  w->Double(10, s.combined_bounds_min);
  w->Double(11, s.combined_bounds_avg);
  w->Int(12, s.variable_bound_gaps_num_finite);
  w->Double(13, s.variable_bound_gaps_max);
  w->Double(14, s.variable_bound_gaps_min);
  w->Double(15, s.variable_bound_gaps_avg);
  w->Double(16, s.objective_vector_abs_max);
  w->Double(17, s.objective_vector_abs_min);
  w->Double(18, s.objective_vector_abs_avg);
  w->Int(19, s.objective_matrix_num_nonzeros);
  w->Double(20, s.objective_matrix_abs_max);
  w->Double(21, s.objective_matrix_abs_min);
  w->Double(22, s.objective_matrix_abs_avg);
  w->Double(23, s.objective_vector_l2_norm);
  w->Double(24, s.combined_bounds_l2_norm);
  w->Double(25, s.constraint_matrix_l2_norm);
  w->Double(26, s.variable_bound_gaps_l2_norm);
  w->Double(27, s.objective_matrix_l2_norm);
  w->Double(28, s.combined_variable_bounds_max);
  w->Double(29, s.combined_variable_bounds_min);
  w->Double(30, s.combined_variable_bounds_avg);
  w->Double(31, s.combined_variable_bounds_l2_norm);
}

void IterationStatsToWire(const PdlpIterationStats& s, Writer* w) {
  w->Int(1, s.iteration_number);
  for (int k = 0; k < s.num_convergence_information; ++k) {
    const PdlpConvergenceInformation& c = s.convergence_information[k];
    Writer m;
    m.Int(1, c.candidate_type);
    m.Double(2, c.primal_objective);
    m.Double(3, c.dual_objective);
    m.Double(4, c.corrected_dual_objective);
    m.Double(5, c.l_inf_primal_residual);
    m.Double(6, c.l2_primal_residual);
    m.Double(7, c.l_inf_dual_residual);
    m.Double(8, c.l2_dual_residual);
    m.Double(14, c.l_inf_primal_variable);
    m.Double(15, c.l2_primal_variable);
    m.Double(16, c.l_inf_dual_variable);
    m.Double(17, c.l2_dual_variable);
    m.Double(24, c.l_inf_componentwise_primal_residual);
    m.Double(25, c.l_inf_componentwise_dual_residual);
    w->Bytes(2, m.out());
  }
  for (int k = 0; k < s.num_infeasibility_information; ++k) {
    const PdlpInfeasibilityInformation& c = s.infeasibility_information[k];
    Writer m;
    m.Double(1, c.max_primal_ray_infeasibility);
    m.Double(2, c.primal_ray_linear_objective);
    m.Double(3, c.primal_ray_quadratic_norm);
    m.Double(4, c.max_dual_ray_infeasibility);
    m.Double(5, c.dual_ray_objective);
    m.Int(6, c.candidate_type);
    w->Bytes(3, m.out());
  }
  w->Double(4, s.cumulative_kkt_matrix_passes);
  w->Int(5, s.cumulative_rejected_steps);
  w->Double(6, s.cumulative_time_sec);
  w->Int(7, s.restart_used);
  w->Double(8, s.step_size);
  w->Double(9, s.primal_weight);
  for (int k = 0; k < s.num_point_metadata; ++k) {
    const PdlpPointMetadata& p = s.point_metadata[k];
    Writer m;
    m.Int(1, p.point_type);
    const int np = std::max(0, std::min<int>(p.num_random_projections, PDLP_MAX_RANDOM_PROJECTION_SEEDS));
    m.PackedDoubles(2, p.random_primal_projections, np);
    m.PackedDoubles(3, p.random_dual_projections, np);
    if (p.has_active_set_information) {
      m.Int(4, p.active_primal_variable_count);
      m.Int(5, p.active_dual_variable_count);
      m.Int(6, p.active_primal_variable_change);
      m.Int(7, p.active_dual_variable_change);
    }
    w->Bytes(11, m.out());
  }
}

bool AllZero(const void* p, size_t n) {
  const unsigned char* b = static_cast<const unsigned char*>(p);
  for (size_t i = 0; i < n; ++i)
    if (b[i] != 0) return false;
  return true;
}

void SolveLogToWire(const PdlpResult& r, Writer* w) {  // tag order
  if (r.instance_name != nullptr && r.instance_name[0] != '\0') w->Bytes(1, r.instance_name);
  w->Int(3, r.termination_reason);
  if (r.termination_string != nullptr && r.termination_string[0] != '\0') w->Bytes(4, r.termination_string);
  w->Int(5, r.iteration_count);
  w->Double(6, r.solve_time_sec);
  for (int64_t k = 0; k < r.num_iteration_stats; ++k) {
    Writer m;
    IterationStatsToWire(r.iteration_stats[k], &m);
    w->Bytes(7, m.out());
  }
  if (r.has_solution_stats) {
    Writer m;
    IterationStatsToWire(r.solution_stats, &m);
    w->Bytes(8, m.out());
  }
  w->Int(10, r.solution_type);
  if (r.has_original_problem_stats) {
    Writer m;
    StatsToWire(r.original_problem_stats, &m);
    w->Bytes(11, m.out());
  }
  if (r.has_preprocessed_problem_stats) {
    Writer m;
    StatsToWire(r.preprocessed_problem_stats, &m);
    w->Bytes(12, m.out());
  }
  w->Double(13, r.preprocessing_time_sec);
  if (!AllZero(&r.params, sizeof r.params)) {  // an error result carries no parameters (ErrorSolverResult, pdhg.cc:329-342)
    Writer m;
    ParamsToWire(r.params, &m);
    w->Bytes(14, m.out());  // primal_dual_hybrid_gradient.cc:1053 echoes the parameters of a solve that ran
  }
  for (int64_t k = 0; k < r.num_feasibility_polishing_details; ++k) {
    const PdlpFeasibilityPolishingDetails& d = r.feasibility_polishing_details[k];
    Writer m;
    m.Int(1, d.polishing_phase_type);
    m.Int(2, d.main_iteration_count);
    {
      Writer pm;
      ParamsToWire(d.params, &pm);
      m.Bytes(3, pm.out());
    }
    m.Int(4, d.termination_reason);
    m.Int(5, d.iteration_count);
    m.Double(6, d.solve_time_sec);
    {
      Writer sm;
      IterationStatsToWire(d.solution_stats, &sm);
      m.Bytes(7, sm.out());
    }
    m.Int(8, d.solution_type);
    for (int64_t i = 0; i < d.num_iteration_stats; ++i) {
      Writer sm;
      IterationStatsToWire(d.iteration_stats[i], &sm);
      m.Bytes(9, sm.out());
    }
    w->Bytes(15, m.out());
  }
}

bool Encode(const proto::Schema& schema, const std::string& wire, int32_t format, std::string* out) {
  switch (format) {
    case PDLP_FORMAT_BINARY: *out = wire; return true;
    case PDLP_FORMAT_TEXT: return proto::WireToText(schema, wire, out);
    case PDLP_FORMAT_JSON: return proto::WireToJson(schema, wire, out);
    default: return false;
  }
}

// ---------------------------------------------------------------------------
// files
// ---------------------------------------------------------------------------
// bzip2 input (quadratic_program_io.cc:50-68 reads .mps.bz2 through the MPS reader's file
// layer). The image carries libbz2.so.1.0 but not bzlib.h, so the three decompression entry
// points are bound at run time against a restatement of the public stream record.
struct BzStream {
  char* next_in;
  unsigned int avail_in, total_in_lo32, total_in_hi32;
  char* next_out;
  unsigned int avail_out, total_out_lo32, total_out_hi32;
  void* state;
  void* (*bzalloc)(void*, int, int);
  void (*bzfree)(void*, void*);
  void* opaque;
};
struct BzApi {
  int (*init)(BzStream*, int, int) = nullptr;
  int (*run)(BzStream*) = nullptr;
  int (*end)(BzStream*) = nullptr;
  bool ok = false;
};
const BzApi& Bz() {
  static const BzApi api = [] {
    BzApi a;
    void* h = nullptr;
    for (const char* name : {"libbz2.so.1.0", "libbz2.so.1", "libbz2.so"}) {
      h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (h != nullptr) break;
    }
    if (h == nullptr) return a;
    a.init = reinterpret_cast<int (*)(BzStream*, int, int)>(dlsym(h, "BZ2_bzDecompressInit"));
    a.run = reinterpret_cast<int (*)(BzStream*)>(dlsym(h, "BZ2_bzDecompress"));
    a.end = reinterpret_cast<int (*)(BzStream*)>(dlsym(h, "BZ2_bzDecompressEnd"));
    a.ok = a.init != nullptr && a.run != nullptr && a.end != nullptr;
    return a;
  }();
  return api;
}

bool ReadBzip2File(const std::string& path, std::string* out, std::string* error) {
  const BzApi& bz = Bz();
  if (!bz.ok) {
    *error = "cannot read " + path + ": libbz2 was not found on this machine; decompress the file first";
    return false;
  }
  std::FILE* f = std::fopen(path.c_str(), "rb");
  if (f == nullptr) {
    *error = "cannot open " + path;
    return false;
  }
  std::string packed;
  std::vector<char> buf(1 << 20);
  for (size_t n; (n = std::fread(buf.data(), 1, buf.size(), f)) > 0;) packed.append(buf.data(), n);
  const bool read_ok = std::ferror(f) == 0;
  std::fclose(f);
  if (!read_ok) {
    *error = "error while reading " + path;
    return false;
  }
  // a .bz2 file is one or more streams back to back (bzip2 concatenation); an empty file is empty
  size_t at = 0;
  while (at < packed.size()) {
    BzStream s{};
    if (bz.init(&s, 0, 0) != 0) {
      *error = "error while reading " + path + ": bzip2 decoder could not start";
      return false;
    }
    int rc = 0;  // BZ_OK
    while (rc == 0) {
      const size_t in = std::min<size_t>(packed.size() - at, 1u << 30);
      s.next_in = packed.data() + at;
      s.avail_in = static_cast<unsigned int>(in);
      s.next_out = buf.data();
      s.avail_out = static_cast<unsigned int>(buf.size());
      rc = bz.run(&s);
      at += in - s.avail_in;
      const size_t produced = buf.size() - s.avail_out;
      out->append(buf.data(), produced);
      if (rc == 0 && in == s.avail_in && produced == 0) rc = -7;  // no progress: the stream is cut short (BZ_UNEXPECTED_EOF)
    }
    bz.end(&s);
    if (rc != 4) {  // BZ_STREAM_END
      *error = "error while reading " + path + ": not a valid bzip2 stream";
      return false;
    }
  }
  return true;
}

bool ReadFile(const std::string& path, std::string* out, std::string* error) {
  if (EndsWith(path, ".bz2")) return ReadBzip2File(path, out, error);
  gzFile f = gzopen(path.c_str(), "rb");  // transparent for files that are not gzipped
  if (f == nullptr) {
    *error = "cannot open " + path;
    return false;
  }
  gzbuffer(f, 1 << 20);
  std::vector<char> buf(1 << 20);
  for (;;) {
    const int n = gzread(f, buf.data(), static_cast<unsigned>(buf.size()));
    if (n < 0) {
      *error = "error while reading " + path;
      gzclose(f);
      return false;
    }
    if (n == 0) break;
    out->append(buf.data(), static_cast<size_t>(n));
  }
  gzclose(f);
  return true;
}

bool WriteFile(const std::string& path, const std::string& data, std::string* error) {
  std::FILE* f = std::fopen(path.c_str(), "wb");
  if (f == nullptr) {
    *error = "cannot open " + path + " for writing";
    return false;
  }
  const bool ok = data.empty() || std::fwrite(data.data(), 1, data.size(), f) == data.size();
  if (std::fclose(f) != 0 || !ok) {
    *error = "error while writing " + path;
    return false;
  }
  return true;
}

}  // namespace
}  // namespace pdlp_b200

// ---------------------------------------------------------------------------
// models
// ---------------------------------------------------------------------------
struct PdlpModel {
  std::vector<int64_t> col_starts, row_indices;
  std::vector<double> values, objective, objective_matrix, lc, uc, lv, uv;
  bool has_objective_matrix = false;
  double objective_offset = 0.0, objective_scaling_factor = 1.0;
  bool has_names = false;
  std::string name;
  std::vector<std::string> variable_names, constraint_names;
  PdlpProblemView view{};

  void Finish() {
    const int64_t n = static_cast<int64_t>(lv.size()), m = static_cast<int64_t>(lc.size());
    view = PdlpProblemView{};
    view.num_variables = n;
    view.num_constraints = m;
    view.num_nonzeros = static_cast<int64_t>(values.size());
    view.col_starts = col_starts.data();
    view.row_indices = row_indices.data();
    view.values = values.data();
    view.objective_vector = objective.data();
    view.objective_matrix_diagonal = has_objective_matrix ? objective_matrix.data() : nullptr;
    view.constraint_lower_bounds = lc.data();
    view.constraint_upper_bounds = uc.data();
    view.variable_lower_bounds = lv.data();
    view.variable_upper_bounds = uv.data();
    view.objective_offset = objective_offset;
    view.objective_scaling_factor = objective_scaling_factor;
    view.problem_name = has_names ? name.c_str() : nullptr;
    view.objective_vector_size = view.objective_matrix_size = -1;
    view.constraint_lower_bounds_size = view.constraint_upper_bounds_size = -1;
    view.variable_lower_bounds_size = view.variable_upper_bounds_size = -1;
  }
};

namespace pdlp_b200 {
namespace {

struct Triplet {
  int64_t row, col;
  double value;
};

// Column-compressed K from triplets: rows sorted within a column, repeated
// (row, col) entries summed (SetEigenMatrixFromTriplets, quadratic_program.cc:338-358).
void BuildCsc(int64_t num_cols, std::vector<Triplet>& entries, PdlpModel* model) {
  std::stable_sort(entries.begin(), entries.end(), [](const Triplet& a, const Triplet& b) { return a.col != b.col ? a.col < b.col : a.row < b.row; });
  model->col_starts.assign(static_cast<size_t>(num_cols) + 1, 0);
  model->row_indices.clear();
  model->values.clear();
  model->row_indices.reserve(entries.size());
  model->values.reserve(entries.size());
  int64_t last_row = -1, last_col = -1;
  for (const Triplet& t : entries) {
    if (t.col == last_col && t.row == last_row) {
      model->values.back() += t.value;
      continue;
    }
    model->row_indices.push_back(t.row);
    model->values.push_back(t.value);
    ++model->col_starts[static_cast<size_t>(t.col) + 1];
    last_row = t.row;
    last_col = t.col;
  }
  for (int64_t j = 0; j < num_cols; ++j) model->col_starts[j + 1] += model->col_starts[j];
}

// ---------------------------------------------------------------------------
// QpFromMpModelProto (quadratic_program.cc:98-211)
// ---------------------------------------------------------------------------
bool ModelFromMpModelBytes(std::string_view bytes, bool relax_integer_variables, bool include_names, PdlpModel* model, std::string* error) {
  Reader r(bytes);
  WireField f;
  bool maximize = false;
  double offset = 0.0;
  std::vector<Triplet> entries;
  std::vector<int32_t> q1, q2;
  std::vector<double> qc;
  bool has_general = false;
  std::vector<int32_t> idx;
  std::vector<double> coef;
  // first pass: variables (constraints may come before them on the wire)
  int64_t num_variables = 0;
  {
    Reader count(bytes);
    while (count.Next(&f))
      if (f.number == 3 && IsBytes(f)) ++num_variables;
    if (!count.ok()) {
      *error = "malformed MPModelProto bytes";
      return false;
    }
  }
  model->lv.reserve(num_variables);
  while (r.Next(&f)) {
    switch (f.number) {
      case 1: if (IsVarint(f)) maximize = f.AsBool(); break;
      case 2: if (IsDouble(f)) offset = f.AsDouble(); break;
      case 5: if (IsBytes(f)) model->name = std::string(f.bytes); break;
      case 7: if (IsBytes(f)) has_general = true; break;
      case 3: {
        if (!IsBytes(f)) break;
        double lo = -kInf, hi = kInf, c = 0.0;
        bool is_integer = false;
        std::string name;
        Reader s(f.bytes);
        WireField g;
        while (s.Next(&g)) {
          if (g.number == 1 && IsDouble(g)) lo = g.AsDouble();
          else if (g.number == 2 && IsDouble(g)) hi = g.AsDouble();
          else if (g.number == 3 && IsDouble(g)) c = g.AsDouble();
          else if (g.number == 4 && IsVarint(g)) is_integer = g.AsBool();
          else if (g.number == 5 && IsBytes(g) && include_names) name = std::string(g.bytes);
        }
        if (!s.ok()) { *error = "malformed MPVariableProto"; return false; }
        if (is_integer && !relax_integer_variables) {
          *error = "Integer variable encountered with relax_integer_variables == false";
          return false;
        }
        model->lv.push_back(lo);
        model->uv.push_back(hi);
        model->objective.push_back(c);
        if (include_names) model->variable_names.push_back(std::move(name));
        break;
      }
      case 4: {
        if (!IsBytes(f)) break;
        double lo = -kInf, hi = kInf;
        std::string name;
        idx.clear();
        coef.clear();
        Reader s(f.bytes);
        WireField g;
        while (s.Next(&g)) {
          if (g.number == 6) { if (!proto::AppendInt32s(g, &idx)) { *error = "malformed var_index"; return false; } }
          else if (g.number == 7) { if (!proto::AppendDoubles(g, &coef)) { *error = "malformed coefficient"; return false; } }
          else if (g.number == 2 && IsDouble(g)) lo = g.AsDouble();
          else if (g.number == 3 && IsDouble(g)) hi = g.AsDouble();
          else if (g.number == 4 && IsBytes(g) && include_names) name = std::string(g.bytes);
        }
        if (!s.ok()) { *error = "malformed MPConstraintProto"; return false; }
        const int64_t i = static_cast<int64_t>(model->lc.size());
        if (idx.size() != coef.size()) {
          *error = std::to_string(i) + "th constraint has " + std::to_string(coef.size()) + " coefficients, expected " + std::to_string(idx.size());
          return false;
        }
        for (size_t j = 0; j < idx.size(); ++j) {
          if (idx[j] < 0 || idx[j] >= num_variables) {
            *error = "Variable index of " + std::to_string(i) + "th constraint's " + std::to_string(j) + "th nonzero is " + std::to_string(idx[j]) +
                     " which is not in the allowed range [0, " + std::to_string(num_variables) + ")";
            return false;
          }
          entries.push_back({i, idx[j], coef[j]});
        }
        model->lc.push_back(lo);
        model->uc.push_back(hi);
        if (include_names) model->constraint_names.push_back(std::move(name));
        break;
      }
      case 8: {
        if (!IsBytes(f)) break;
        Reader s(f.bytes);
        WireField g;
        while (s.Next(&g)) {
          bool ok = true;
          if (g.number == 1) ok = proto::AppendInt32s(g, &q1);
          else if (g.number == 2) ok = proto::AppendInt32s(g, &q2);
          else if (g.number == 3) ok = proto::AppendDoubles(g, &qc);
          if (!ok) { *error = "malformed MPQuadraticObjective"; return false; }
        }
        if (!s.ok()) { *error = "malformed MPQuadraticObjective"; return false; }
        break;
      }
      default: break;
    }
  }
  if (!r.ok()) {
    *error = "malformed MPModelProto bytes";
    return false;
  }
  if (has_general) {
    *error = "General constraints are not supported.";
    return false;
  }
  const int64_t n = num_variables;
  BuildCsc(n, entries, model);
  if (q1.size() != q2.size() || q1.size() != qc.size()) {
    *error = "The quadratic objective has " + std::to_string(q1.size()) + " qvar1_indices, " + std::to_string(q2.size()) + " qvar2_indices, and " +
             std::to_string(qc.size()) + " coefficients, expected equal numbers.";
    return false;
  }
  if (!q1.empty()) {
    model->has_objective_matrix = true;
    model->objective_matrix.assign(static_cast<size_t>(n), 0.0);
  }
  for (size_t k = 0; k < q1.size(); ++k) {
    const int64_t a = q1[k], b = q2[k];
    if (a < 0 || b < 0 || a >= n || b >= n) {
      *error = "The quadratic objective's " + std::to_string(k) + "th nonzero has indices " + std::to_string(a) + " and " + std::to_string(b) +
               ", which are not both in the expected range [0, " + std::to_string(n) + ")";
      return false;
    }
    if (a != b) {
      *error = "The quadratic objective's " + std::to_string(k) + "th nonzero has off-diagonal element at (" + std::to_string(a) + ", " +
               std::to_string(b) + "). Only diagonal objective matrices are supported.";
      return false;
    }
    model->objective_matrix[static_cast<size_t>(a)] = 2.0 * qc[k];  // QuadraticProgram has an implicit 1/2 in front of the quadratic term
  }
  model->objective_offset = offset;
  if (maximize) {  // quadratic_program.cc:200-208
    model->objective_offset *= -1;
    for (double& v : model->objective) v *= -1;
    for (double& v : model->objective_matrix) v *= -1;
    model->objective_scaling_factor = -1.0;
  }
  model->has_names = include_names;
  if (!include_names) model->name.clear();
  model->Finish();
  return true;
}

// ---------------------------------------------------------------------------
// QpToMpModelProto (quadratic_program.cc:236-315)
// ---------------------------------------------------------------------------
bool QpToMpModelWire(const PdlpProblemView& qp, const char* const* variable_names, const char* const* constraint_names, std::string* out,
                     std::string* error) {
  const int64_t n = qp.num_variables, m = qp.num_constraints;
  const int64_t kMax = std::numeric_limits<int32_t>::max();
  if (n > kMax) { *error = "Too many variables (" + std::to_string(n) + ") to index with an int32_t."; return false; }
  if (m > kMax) { *error = "Too many constraints (" + std::to_string(m) + ") to index with an int32_t."; return false; }
  const double s = qp.objective_scaling_factor;
  if (s == 0) { *error = "objective_scaling_factor cannot be zero."; return false; }
  Writer w;
  w.Bool(1, s < 0);
  w.Double(2, s * qp.objective_offset);
  for (int64_t j = 0; j < n; ++j) {
    Writer v;
    v.Double(1, qp.variable_lower_bounds[j]);
    v.Double(2, qp.variable_upper_bounds[j]);
    v.Double(3, s * qp.objective_vector[j]);
    if (variable_names != nullptr && variable_names[j] != nullptr && variable_names[j][0] != '\0') v.Bytes(5, variable_names[j]);
    w.Bytes(3, v.out());
  }
  // rows of K from its columns
  std::vector<int64_t> row_starts(static_cast<size_t>(m) + 1, 0);
  const int64_t nnz = n > 0 ? qp.col_starts[n] : 0;
  for (int64_t p = 0; p < nnz; ++p) ++row_starts[static_cast<size_t>(qp.row_indices[p]) + 1];
  for (int64_t i = 0; i < m; ++i) row_starts[i + 1] += row_starts[i];
  std::vector<int32_t> cols(static_cast<size_t>(nnz));
  std::vector<double> vals(static_cast<size_t>(nnz));
  {
    std::vector<int64_t> at(row_starts.begin(), row_starts.end() - 1);
    for (int64_t j = 0; j < n; ++j)
      for (int64_t p = qp.col_starts[j]; p < qp.col_starts[j + 1]; ++p) {
        const int64_t q = at[static_cast<size_t>(qp.row_indices[p])]++;
        cols[static_cast<size_t>(q)] = static_cast<int32_t>(j);
        vals[static_cast<size_t>(q)] = qp.values[p];
      }
  }
  for (int64_t i = 0; i < m; ++i) {
    Writer c;
    c.Double(2, qp.constraint_lower_bounds[i]);
    c.Double(3, qp.constraint_upper_bounds[i]);
    if (constraint_names != nullptr && constraint_names[i] != nullptr && constraint_names[i][0] != '\0') c.Bytes(4, constraint_names[i]);
    const int64_t b = row_starts[i], len = row_starts[i + 1] - b;
    c.PackedInts(6, cols.data() + b, len);
    c.PackedDoubles(7, vals.data() + b, len);
    w.Bytes(4, c.out());
  }
  if (qp.problem_name != nullptr && qp.problem_name[0] != '\0') w.Bytes(5, qp.problem_name);
  if (qp.objective_matrix_diagonal != nullptr) {
    Writer q;
    bool any = false;
    for (int64_t j = 0; j < n; ++j)
      if (qp.objective_matrix_diagonal[j] != 0.0) { q.Int(1, j); any = true; }
    for (int64_t j = 0; j < n; ++j)
      if (qp.objective_matrix_diagonal[j] != 0.0) q.Int(2, j);
    for (int64_t j = 0; j < n; ++j)
      if (qp.objective_matrix_diagonal[j] != 0.0) q.Double(3, s * qp.objective_matrix_diagonal[j] / 2.0);  // undo the implicit 1/2
    if (any) w.Bytes(8, q.out());
  }
  *out = std::move(w.out());
  return true;
}

// ---------------------------------------------------------------------------
// MPS reader: ortools/lp_data/mps_reader_template.h:90-260 as PDLP uses it
// (quadratic_program_io.cc:364-407): rows and columns numbered in order of
// first appearance, integrality dropped, the negated right-hand side of the
// objective row is the objective offset, maximisation becomes minimisation.
// ---------------------------------------------------------------------------
using Fields = std::vector<std::string_view>;

bool IsSpace(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f'; }

void SplitFields(std::string_view line, Fields* out) {
  out->clear();
  size_t i = 0;
  const size_t n = line.size();
  while (i < n) {
    while (i < n && IsSpace(line[i])) ++i;
    size_t j = i;
    while (j < n && !IsSpace(line[j])) ++j;
    if (j > i) out->push_back(line.substr(i, j - i));
    i = j;
  }
}

std::string_view Strip(std::string_view s) {
  size_t a = 0, b = s.size();
  while (a < b && IsSpace(s[a])) ++a;
  while (b > a && IsSpace(s[b - 1])) --b;
  return s.substr(a, b - a);
}

// Fixed-format fields: columns 2-3, 5-12, 15-22, 25-36, 40-47, 50-61 (1-based).
void FixedFields(std::string_view line, Fields* out) {
  static const int kCols[6][2] = {{1, 3}, {4, 12}, {14, 22}, {24, 36}, {39, 47}, {49, 61}};
  out->clear();
  for (const auto& c : kCols) {
    if (static_cast<int>(line.size()) <= c[0]) continue;
    out->push_back(Strip(line.substr(static_cast<size_t>(c[0]), static_cast<size_t>(c[1] - c[0]))));
  }
  while (!out->empty() && out->back().empty()) out->pop_back();
}

// Case-insensitive comparison with an upper-case literal.
bool IsWord(std::string_view s, std::string_view upper) {
  if (s.size() != upper.size()) return false;
  for (size_t i = 0; i < s.size(); ++i)
    if (std::toupper(static_cast<unsigned char>(s[i])) != upper[i]) return false;
  return true;
}

// Names -> dense indices in order of first appearance: open addressing over
// (hash, index) pairs, so a lookup touches one slot run and one name (the
// chained std::unordered_map cost ~400 ns per entry of a large COLUMNS section).
class NameIndex {
 public:
  int64_t Find(std::string_view s) const {
    if (slots_.empty()) return -1;
    const uint64_t h = Hash(s);
    for (size_t at = h & mask_;; at = (at + 1) & mask_) {
      const Slot& slot = slots_[at];
      if (slot.hash == 0) return -1;
      if (slot.hash == h && names_[static_cast<size_t>(slot.index)] == s) return slot.index;
    }
  }
  // Appends a name that is not present yet; returns its index.
  int64_t Insert(std::string_view s) {
    if ((names_.size() + 1) * 2 > slots_.size()) Grow();
    const int64_t index = static_cast<int64_t>(names_.size());
    names_.emplace_back(s);
    Place(Hash(s), index);
    return index;
  }
  size_t size() const { return names_.size(); }
  std::vector<std::string>& names() { return names_; }

 private:
  struct Slot {
    uint64_t hash = 0;  // 0 = empty
    int64_t index = 0;
  };
  static uint64_t Hash(std::string_view s) {  // FNV-1a, never 0
    uint64_t h = 1469598103934665603ull;
    for (unsigned char c : s) h = (h ^ c) * 1099511628211ull;
    return h | 1ull;
  }
  void Place(uint64_t h, int64_t index) {
    size_t at = h & mask_;
    while (slots_[at].hash != 0) at = (at + 1) & mask_;
    slots_[at].hash = h;
    slots_[at].index = index;
  }
  void Grow() {
    const size_t size = slots_.empty() ? 1024 : slots_.size() * 2;
    slots_.assign(size, Slot{});
    mask_ = size - 1;
    for (size_t i = 0; i < names_.size(); ++i) Place(Hash(names_[i]), static_cast<int64_t>(i));
  }
  std::vector<std::string> names_;
  std::vector<Slot> slots_;
  size_t mask_ = 0;
};

struct MpsParser {
  enum Section { kNone, kName, kObjsense, kRows, kColumns, kRhs, kRanges, kBounds, kOther };

  bool include_names;
  PdlpModel* model;
  std::string error;

  Section section = kNone;
  bool maximize = false;
  NameIndex rows, cols;  // constraint rows / columns in order of first appearance
  bool has_objective_row = false;
  std::string objective_row;
  std::unordered_set<std::string> ignored_rows;  // extra N rows
  std::vector<char> binary_by_default;
  std::vector<Triplet> entries;
  double offset = 0.0;
  bool in_integer_block = false;
  int lineno = 0;
  // scratch reused from line to line
  Fields f, g;
  std::string key;
  std::string last_col_name;
  int64_t last_col = -1;

  bool Fail(const std::string& what) {
    if (error.empty()) error = "line " + std::to_string(lineno) + ": " + what;
    return false;
  }
  static std::string Quoted(std::string_view s) { return "'" + std::string(s) + "'"; }

  bool Number(std::string_view tok, double* out) {
    // fast path: plain decimal / scientific notation
    if (!tok.empty()) {
      const char* b = tok.data();
      const char* e = b + tok.size();
      auto r = std::from_chars(*b == '+' && tok.size() > 1 && tok[1] != '-' && tok[1] != '+' ? b + 1 : b, e, *out);
      if (r.ec == std::errc() && r.ptr == e && !std::isnan(*out) && !std::isinf(*out)) return true;
    }
    // slow path: Fortran exponents (1.5D0), inf / infinity, out-of-range values
    std::string t(tok), low;
    for (char c : t) low.push_back(static_cast<char>(std::tolower(static_cast<unsigned char>(c))));
    if (low.find("inf") == std::string::npos)
      for (char& c : t)
        if (c == 'D' || c == 'd') c = 'e';
    char* end = nullptr;
    const double v = std::strtod(t.c_str(), &end);
    if (t.empty() || end == nullptr || *end != '\0' || end == t.c_str()) return Fail("cannot parse number " + Quoted(tok));
    if (std::isnan(v)) return Fail("NaN value");
    if (low.find('x') != std::string::npos) return Fail("cannot parse number " + Quoted(tok));  // strtod accepts hex floats; the reader does not
    *out = v;
    return true;
  }
  bool IsObjectiveRow(std::string_view row) const { return has_objective_row && row == objective_row; }
  bool IsIgnoredRow(std::string_view row) {
    if (ignored_rows.empty()) return false;
    key.assign(row);
    return ignored_rows.count(key) != 0;
  }
  // Index of a constraint row, -1 if unknown.
  int64_t FindRow(std::string_view row) const { return rows.Find(row); }
  int64_t FindCol(std::string_view name, bool* is_new = nullptr) {
    if (is_new != nullptr) *is_new = false;
    if (last_col >= 0 && name == last_col_name) return last_col;  // consecutive lines of a column
    int64_t j = cols.Find(name);
    if (j < 0) {
      j = cols.Insert(name);
      model->objective.push_back(0.0);
      model->lv.push_back(0.0);
      model->uv.push_back(kInf);
      binary_by_default.push_back(0);
      if (is_new != nullptr) *is_new = true;
    }
    last_col = j;
    last_col_name.assign(name);
    return j;
  }
  bool SetRhs(std::string_view row, double value) {
    if (IsObjectiveRow(row)) {
      offset = -value;  // minus the right-hand side of the objective row
      return true;
    }
    if (IsIgnoredRow(row)) return true;
    const int64_t i = FindRow(row);
    if (i < 0) return Fail("unknown row " + Quoted(row));
    double& lo = model->lc[static_cast<size_t>(i)];
    double& hi = model->uc[static_cast<size_t>(i)];
    if (lo != -kInf) lo = value;
    if (hi != kInf) hi = value;
    return true;
  }
  bool SetRange(std::string_view row, double value) {
    if (IsObjectiveRow(row) || IsIgnoredRow(row)) return true;
    const int64_t i = FindRow(row);
    if (i < 0) return Fail("unknown row " + Quoted(row));
    double lo = model->lc[static_cast<size_t>(i)], hi = model->uc[static_cast<size_t>(i)];
    if (lo == hi) {
      if (value < 0.0) lo += value;
      else hi += value;
    }
    if (lo == -kInf) lo = hi - std::fabs(value);
    if (hi == kInf) hi = lo + std::fabs(value);
    model->lc[static_cast<size_t>(i)] = lo;
    model->uc[static_cast<size_t>(i)] = hi;
    return true;
  }
  void SetSense(std::string_view word) { maximize = IsWord(word, "MAX") || IsWord(word, "MAXIMIZE"); }

  bool Header(std::string_view line, bool* done) {
    SplitFields(line, &f);
    std::string name;
    for (char c : f[0]) name.push_back(static_cast<char>(std::toupper(static_cast<unsigned char>(c))));
    if (name == "NAME") {
      section = kName;
      model->name.clear();
      for (size_t k = 1; k < f.size(); ++k) {
        if (k > 1) model->name += " ";
        model->name.append(f[k]);
      }
    } else if (name == "OBJSENSE" || name == "OBJSENCE") {
      section = kObjsense;
      if (f.size() > 1) SetSense(f[1]);
    } else if (name == "OBJSENSEMAX") {
      section = kOther;
      maximize = true;
    } else if (name == "ROWS" || name == "LAZYCONS" || name == "USERCUTS") {
      section = kRows;
    } else if (name == "COLUMNS") {
      section = kColumns;
    } else if (name == "RHS") {
      section = kRhs;
    } else if (name == "RANGES") {
      section = kRanges;
    } else if (name == "BOUNDS") {
      section = kBounds;
    } else if (name == "ENDATA") {
      section = kOther;
      *done = true;
    } else if (name == "QUADOBJ" || name == "QMATRIX" || name == "QSECTION") {
      return Fail("quadratic objective sections are not supported by the linear-program reader");
    } else if (name == "INDICATORS" || name == "SOS") {
      return Fail("section " + name + " is not supported");
    } else {
      return Fail("unknown section " + Quoted(f[0]));
    }
    return true;
  }

  // Returns false on error; *done is set at ENDATA.
  bool Line(std::string_view line, bool* done) {
    ++lineno;
    while (!line.empty() && (line.back() == '\r' || line.back() == '\n')) line.remove_suffix(1);
    const std::string_view stripped = Strip(line);
    if (stripped.empty() || stripped[0] == '*') return true;
    if (!IsSpace(line[0])) return Header(line, done);
    SplitFields(line, &f);
    switch (section) {
      case kObjsense:
        SetSense(f[0]);
        return true;
      case kRows: {
        if (f.size() != 2) FixedFields(line, &f);
        if (f.size() != 2) return Fail("expected <type> <row name>");
        const std::string_view t = f[0], name = f[1];
        if (IsWord(t, "N")) {
          if (!has_objective_row) {
            has_objective_row = true;
            objective_row.assign(name);
          } else {
            ignored_rows.insert(std::string(name));
          }
          return true;
        }
        const bool e = IsWord(t, "E"), l = IsWord(t, "L"), gq = IsWord(t, "G");
        if (!e && !l && !gq) return Fail("unknown row type " + Quoted(t));
        if (rows.Find(name) >= 0) return Fail("duplicate row " + Quoted(name));
        rows.Insert(name);
        model->lc.push_back(l ? -kInf : 0.0);
        model->uc.push_back(gq ? kInf : 0.0);
        return true;
      }
      case kColumns: {
        if (f.size() >= 3 && IsWord(f[1], "'MARKER'")) {
          std::string upper;
          for (char c : f[2]) upper.push_back(static_cast<char>(std::toupper(static_cast<unsigned char>(c))));
          in_integer_block = upper.find("INTORG") != std::string::npos;
          return true;
        }
        if (f.size() != 3 && f.size() != 5) {
          FixedFields(line, &f);
          if (!f.empty()) f.erase(f.begin());
        }
        if (f.size() != 3 && f.size() != 5) return Fail("expected <column> <row> <value> [<row> <value>]");
        bool is_new;
        const int64_t j = FindCol(f[0], &is_new);
        if (is_new && in_integer_block) {  // integer by marker, no bound yet: [0, 1]
          binary_by_default[static_cast<size_t>(j)] = 1;
          model->uv[static_cast<size_t>(j)] = 1.0;
        }
        for (size_t k = 1; k + 1 < f.size(); k += 2) {
          double value;
          if (!Number(f[k + 1], &value)) return false;
          const std::string_view row = f[k];
          if (IsObjectiveRow(row)) {
            model->objective[static_cast<size_t>(j)] = value;
          } else if (IsIgnoredRow(row)) {
            continue;
          } else {
            const int64_t i = FindRow(row);
            if (i < 0) return Fail("unknown row " + Quoted(row));
            entries.push_back({i, j, value});
          }
        }
        return true;
      }
      case kRhs:
      case kRanges: {
        g.assign(f.begin(), f.end());
        if (g.size() % 2 == 0) g.insert(g.begin(), std::string_view());  // the set name may be missing
        if (g.size() != 3 && g.size() != 5) {
          FixedFields(line, &g);
          if (!g.empty()) g.erase(g.begin());
        }
        for (size_t k = 1; k + 1 < g.size(); k += 2) {
          double value;
          if (!Number(g[k + 1], &value)) return false;
          if (!(section == kRhs ? SetRhs(g[k], value) : SetRange(g[k], value))) return false;
        }
        return true;
      }
      case kBounds: {
        const std::string_view kind = f[0];
        const bool lo_k = IsWord(kind, "LO"), li_k = IsWord(kind, "LI"), up_k = IsWord(kind, "UP"), ui_k = IsWord(kind, "UI");
        const bool fx_k = IsWord(kind, "FX"), sc_k = IsWord(kind, "SC");
        const bool needs_value = lo_k || up_k || fx_k || li_k || ui_k || sc_k;
        std::string_view column;
        double value = 0.0;
        if (needs_value) {  // ' <type> <set name> <column> <value>'; the set name may be missing
          if (f.size() == 4) {
            column = f[2];
            if (!Number(f[3], &value)) return false;
          } else if (f.size() == 3) {
            column = f[1];
            if (!Number(f[2], &value)) return false;
          } else {
            FixedFields(line, &g);
            if (g.size() < 4) return Fail("malformed bound");
            column = g[2];
            if (!Number(g[3], &value)) return false;
          }
        } else if (f.size() >= 3) {
          column = f[2];
        } else if (f.size() == 2) {
          column = f[1];
        } else {
          return Fail("malformed bound");
        }
        const int64_t j = FindCol(column);
        double lo = model->lv[static_cast<size_t>(j)], hi = model->uv[static_cast<size_t>(j)];
        if (binary_by_default[static_cast<size_t>(j)]) {
          lo = 0.0;
          hi = kInf;
        }
        if (lo_k || li_k) {
          lo = value;
          if (li_k && lo == 0.0) hi = kInf;
        } else if (up_k || ui_k) {
          hi = value;
        } else if (fx_k) {
          lo = hi = value;
        } else if (IsWord(kind, "FR")) {
          lo = -kInf;
          hi = kInf;
        } else if (IsWord(kind, "MI")) {
          lo = -kInf;
        } else if (IsWord(kind, "PL")) {
          hi = kInf;
        } else if (IsWord(kind, "BV")) {
          lo = 0.0;
          hi = 1.0;
        } else if (sc_k) {
          return Fail("semi-continuous variables are not supported");
        } else {
          return Fail("unknown bound type " + Quoted(kind));
        }
        binary_by_default[static_cast<size_t>(j)] = 0;
        model->lv[static_cast<size_t>(j)] = lo;
        model->uv[static_cast<size_t>(j)] = hi;
        return true;
      }
      case kName:
        return true;
      default:
        return Fail("data outside of a section");
    }
  }

  bool Finish() {
    const int64_t n = static_cast<int64_t>(cols.size());
    BuildCsc(n, entries, model);
    model->objective_offset = offset;
    if (maximize) {  // quadratic_program_io.cc:259-266
      model->objective_scaling_factor = -1.0;
      model->objective_offset *= -1;
      for (double& v : model->objective) v *= -1;
    }
    model->has_names = include_names;
    if (include_names) {
      model->variable_names = std::move(cols.names());
      model->constraint_names = std::move(rows.names());
    } else {
      model->name.clear();
    }
    model->Finish();
    return true;
  }
};

bool ModelFromMpsText(std::string_view text, bool include_names, PdlpModel* model, std::string* error) {
  MpsParser parser{include_names, model};
  size_t at = 0;
  bool done = false;
  while (at < text.size() && !done) {
    const char* nl = static_cast<const char*>(std::memchr(text.data() + at, '\n', text.size() - at));
    const size_t end = nl == nullptr ? text.size() : static_cast<size_t>(nl - text.data());
    if (!parser.Line(text.substr(at, end - at), &done)) {
      *error = parser.error;
      return false;
    }
    at = end + 1;
  }
  return parser.Finish();
}

// ---------------------------------------------------------------------------
// WriteLinearProgramToMps (quadratic_program_io.cc:80-93), free format
// ---------------------------------------------------------------------------
bool LinearProgramToMps(const PdlpProblemView& qp, const char* const* variable_names, const char* const* constraint_names, std::string* out,
                        std::string* error) {
  if (qp.objective_matrix_diagonal != nullptr) {
    for (int64_t j = 0; j < qp.num_variables; ++j)
      if (qp.objective_matrix_diagonal[j] != 0.0) {
        *error = "'linear_program' has a quadratic objective";
        return false;
      }
  }
  const int64_t n = qp.num_variables, m = qp.num_constraints;
  const double s = qp.objective_scaling_factor;
  std::vector<std::string> row_names(static_cast<size_t>(m));
  for (int64_t i = 0; i < m; ++i)
    row_names[static_cast<size_t>(i)] =
        constraint_names != nullptr && constraint_names[i] != nullptr && constraint_names[i][0] ? std::string(constraint_names[i]) : "R" + std::to_string(i);
  auto col_name = [&](int64_t j) { return variable_names != nullptr && variable_names[j] != nullptr && variable_names[j][0] ? std::string(variable_names[j]) : "C" + std::to_string(j); };
  std::string& o = *out;
  o.reserve(static_cast<size_t>(64 + 24 * (m + n) + 48 * (n > 0 ? qp.col_starts[n] : 0)));
  auto number = [&](double v) {
    if (std::isinf(v)) {
      o += v > 0 ? "inf" : "-inf";
      return;
    }
    char buf[40];
    const std::to_chars_result r = std::to_chars(buf, buf + sizeof buf, v);
    o.append(buf, r.ptr);
  };
  auto entry = [&](const char* lead, const std::string& a, const std::string& b, double v) {  // "<lead><a> <b> <v>\n"
    o += lead;
    o += a;
    o += ' ';
    o += b;
    o += ' ';
    number(v);
    o += '\n';
  };
  static const std::string kCost = "COST", kRhs = "RHS", kRng = "RNG", kBnd = "BND";
  o += "NAME ";
  o += qp.problem_name != nullptr ? qp.problem_name : "";
  o += '\n';
  if (s < 0) o += "OBJSENSE\n    MAX\n";
  o += "ROWS\n N COST\n";
  std::vector<char> kinds(static_cast<size_t>(m));
  const double* lc = qp.constraint_lower_bounds;
  const double* uc = qp.constraint_upper_bounds;
  for (int64_t i = 0; i < m; ++i) {
    char t;
    if (lc[i] == uc[i]) t = 'E';
    else if (lc[i] == -kInf && uc[i] == kInf) { *error = "free constraint row " + std::to_string(i) + " cannot be written to MPS"; return false; }
    else if (lc[i] == -kInf) t = 'L';
    else t = 'G';  // ranged rows: G with a RANGES entry
    kinds[static_cast<size_t>(i)] = t;
    o += ' ';
    o += t;
    o += ' ';
    o += row_names[static_cast<size_t>(i)];
    o += '\n';
  }
  o += "COLUMNS\n";
  for (int64_t j = 0; j < n; ++j) {
    bool wrote = false;
    const std::string cn = col_name(j);
    if (qp.objective_vector[j] != 0.0) {
      entry("    ", cn, kCost, s * qp.objective_vector[j]);
      wrote = true;
    }
    for (int64_t p = qp.col_starts[j]; p < qp.col_starts[j + 1]; ++p) {
      entry("    ", cn, row_names[static_cast<size_t>(qp.row_indices[p])], qp.values[p]);
      wrote = true;
    }
    if (!wrote) entry("    ", cn, kCost, 0.0);
  }
  o += "RHS\n";
  if (qp.objective_offset != 0.0) entry("    ", kRhs, kCost, -s * qp.objective_offset);
  for (int64_t i = 0; i < m; ++i) {
    const double rhs = kinds[static_cast<size_t>(i)] == 'L' ? uc[i] : lc[i];
    if (rhs != 0.0) entry("    ", kRhs, row_names[static_cast<size_t>(i)], rhs);
  }
  bool any_range = false;
  for (int64_t i = 0; i < m; ++i)
    if (kinds[static_cast<size_t>(i)] == 'G' && uc[i] != kInf) {
      if (!any_range) o += "RANGES\n";
      any_range = true;
      entry("    ", kRng, row_names[static_cast<size_t>(i)], uc[i] - lc[i]);
    }
  o += "BOUNDS\n";
  const double* lv = qp.variable_lower_bounds;
  const double* uv = qp.variable_upper_bounds;
  for (int64_t j = 0; j < n; ++j) {
    const std::string cn = col_name(j);
    if (lv[j] == -kInf && uv[j] == kInf) {
      o += " FR BND " + cn + "\n";
    } else if (lv[j] == uv[j]) {
      entry(" FX ", kBnd, cn, lv[j]);
    } else {
      if (lv[j] == -kInf) o += " MI BND " + cn + "\n";
      else if (lv[j] != 0.0) entry(" LO ", kBnd, cn, lv[j]);
      if (uv[j] != kInf) entry(" UP ", kBnd, cn, uv[j]);
    }
  }
  o += "ENDATA\n";
  return true;
}

// ---------------------------------------------------------------------------
// ReadQuadraticProgramOrDie (quadratic_program_io.cc:50-68)
// ---------------------------------------------------------------------------
bool ReadModel(const std::string& path, bool include_names, PdlpModel* model, std::string* error) {
  std::string base = path;
  if (EndsWith(base, ".gz")) base.resize(base.size() - 3);
  else if (EndsWith(base, ".bz2")) base.resize(base.size() - 4);
  const bool mps = EndsWith(base, ".mps");
  const bool pb = EndsWith(base, ".pb"), textproto = EndsWith(base, ".textproto"), json = EndsWith(base, ".json");
  if (!mps && !pb && !textproto && !json) {
    *error = "Invalid filename suffix in " + path + ". Valid suffixes are .mps, .mps.gz, .pb, .textproto, .json, and .json.gz";
    return false;
  }
  std::string data;
  if (!ReadFile(path, &data, error)) return false;
  if (mps) return ModelFromMpsText(data, include_names, model, error);
  std::string wire;
  if (textproto) {
    if (!proto::TextToWire(proto::MPModelSchema(), data, &wire, error)) return false;
  } else if (json) {
    if (!proto::JsonToWire(proto::MPModelSchema(), data, &wire, error)) return false;
  } else {
    wire = std::move(data);
  }
  return ModelFromMpModelBytes(wire, /*relax_integer_variables=*/true, include_names, model, error);
}

// ---------------------------------------------------------------------------
// PdlpSolveProto (pdlp_proto_solver.cc:36-130)
// ---------------------------------------------------------------------------
enum { kMpSolverOptimal = 0, kMpSolverInfeasible = 2, kMpSolverAbnormal = 4, kMpSolverModelInvalid = 5, kMpSolverNotSolved = 6,
       kMpSolverModelInvalidSolverParameters = 85, kMpSolverCancelledByUser = 98 };

int32_t SolveProto(std::string_view request, bool relax_integer_variables, const volatile int32_t* interrupt_solve, std::string* response) {
  Reader r(request);
  WireField f;
  std::string_view model_bytes;
  bool has_model = false, has_time_limit = false, verbose = false;
  double time_limit = 0.0;
  std::string params_text;
  while (r.Next(&f)) {
    if (f.number == 1 && IsBytes(f)) { model_bytes = f.bytes; has_model = true; }
    else if (f.number == 3 && IsDouble(f)) { time_limit = f.AsDouble(); has_time_limit = true; }
    else if (f.number == 4 && IsVarint(f)) verbose = f.AsBool();
    else if (f.number == 5 && IsBytes(f)) params_text = std::string(f.bytes);
  }
  Writer w;
  auto finish = [&] { *response = std::move(w.out()); return PDLP_B200_STATUS_OK; };
  if (!r.ok()) {
    w.Int(1, kMpSolverModelInvalid);
    w.Bytes(7, "The request is not a valid MPModelRequest.");
    return finish();
  }
  PdlpParams params;
  SetDefaultParams(&params);
  params.verbosity_level = verbose ? 3 : 0;
  {
    std::string wire, error;
    if (!proto::TextToWire(proto::ParamsSchema(), params_text, &wire, &error, /*allow_singular_overwrites=*/true) || !MergeParams(wire, &params, &error)) {
      w.Int(1, kMpSolverModelInvalidSolverParameters);
      return finish();
    }
  }
  if (interrupt_solve != nullptr && *interrupt_solve != 0) {
    w.Int(1, kMpSolverNotSolved);
    return finish();
  }
  if (has_time_limit) params.termination_criteria.time_sec_limit = time_limit;
  if (!has_model) {
    w.Int(1, kMpSolverModelInvalid);
    w.Bytes(7, "The request has no model.");
    return finish();
  }
  PdlpModel model;
  {
    std::string error;
    if (!ModelFromMpModelBytes(model_bytes, relax_integer_variables, /*include_names=*/false, &model, &error)) {
      w.Int(1, kMpSolverModelInvalid);  // the reference returns the InvalidArgument status of QpFromMpModelProto to its caller
      w.Bytes(7, error);
      return finish();
    }
  }
  const double scaling = model.objective_scaling_factor;
  PdlpResult result;
  std::memset(&result, 0, sizeof result);
  const int32_t rc = pdlp_b200_primal_dual_hybrid_gradient(&model.view, &params, nullptr, 0, nullptr, 0, interrupt_solve, nullptr, nullptr, nullptr, &result);
  if (rc != PDLP_B200_STATUS_OK) {
    pdlp_b200_result_free(&result);  // (a failed solve may already own a termination string)
    return rc;
  }
  int status = kMpSolverNotSolved;
  switch (result.termination_reason) {
    case PDLP_TERMINATION_REASON_OPTIMAL: status = kMpSolverOptimal; break;
    case PDLP_TERMINATION_REASON_NUMERICAL_ERROR: status = kMpSolverAbnormal; break;
    case PDLP_TERMINATION_REASON_PRIMAL_INFEASIBLE: status = kMpSolverInfeasible; break;
    case PDLP_TERMINATION_REASON_INTERRUPTED_BY_USER: status = kMpSolverCancelledByUser; break;
    default: break;
  }
  w.Int(1, status);
  if (result.has_solution_stats)  // GetConvergenceInformation, iteration_stats.cc:597-606
    for (int k = 0; k < result.solution_stats.num_convergence_information; ++k)
      if (result.solution_stats.convergence_information[k].candidate_type == result.solution_type) {
        w.Double(2, result.solution_stats.convergence_information[k].primal_objective);
        break;
      }
  w.PackedDoubles(3, result.primal_solution, result.primal_size);
  {  // maximisation was turned into minimisation: duals and reduced costs change sign back
    std::vector<double> tmp(static_cast<size_t>(std::max(result.dual_size, result.primal_size)));
    for (int64_t i = 0; i < result.dual_size; ++i) tmp[static_cast<size_t>(i)] = scaling * result.dual_solution[i];
    w.PackedDoubles(4, tmp.data(), result.dual_size);
    for (int64_t i = 0; i < result.primal_size; ++i) tmp[static_cast<size_t>(i)] = scaling * result.reduced_costs[i];
    w.PackedDoubles(6, tmp.data(), result.primal_size);
  }
  if (result.termination_string != nullptr && result.termination_string[0] != '\0') w.Bytes(7, result.termination_string);
  {
    Writer log;
    SolveLogToWire(result, &log);
    w.Bytes(11, log.out());
  }
  pdlp_b200_result_free(&result);
  return finish();
}

const proto::Schema* SchemaByName(const std::string& name) {
  if (name == "PrimalDualHybridGradientParams") return &proto::ParamsSchema();
  if (name == "TerminationCriteria") return &proto::TerminationCriteriaSchema();
  if (name == "SolveLog") return &proto::SolveLogSchema();
  if (name == "IterationStats") return &proto::IterationStatsSchema();
  if (name == "MPModelProto") return &proto::MPModelSchema();
  if (name == "MPModelRequest") return &proto::MPModelRequestSchema();
  if (name == "MPSolutionResponse") return &proto::MPSolutionResponseSchema();
  return nullptr;
}

}  // namespace
}  // namespace pdlp_b200

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
using namespace pdlp_b200;  // NOLINT

namespace {
// No exception may cross the C boundary: allocation failures and the like
// become PDLP_B200_STATUS_BAD_ARGUMENT with the message.
template <class F>
int32_t HostGuard(char* error, int64_t error_capacity, F body) {
  try {
    return body();
  } catch (const std::exception& e) {
    return BadArgument(error, error_capacity, e.what());
  } catch (...) {
    return BadArgument(error, error_capacity, "unknown failure");
  }
}
}  // namespace

extern "C" {

void pdlp_b200_blob_free(PdlpBlob* blob) {
  if (blob == nullptr) return;
  std::free(blob->data);
  blob->data = nullptr;
  blob->size = 0;
}

int32_t pdlp_b200_params_merge_bytes(const uint8_t* data, int64_t size, PdlpParams* params, char* error, int64_t error_capacity) {
  return HostGuard(error, error_capacity, [&]() -> int32_t {
    if (params == nullptr || size < 0 || (data == nullptr && size > 0)) return BadArgument(error, error_capacity, "null argument");
    std::string err;
    PdlpParams merged = *params;  // all or nothing
    if (!MergeParams(std::string_view(reinterpret_cast<const char*>(data), static_cast<size_t>(size)), &merged, &err))
      return BadArgument(error, error_capacity, err);
    *params = merged;
    return PDLP_B200_STATUS_OK;
  });
}

int32_t pdlp_b200_params_parse_bytes(const uint8_t* data, int64_t size, PdlpParams* params, char* error, int64_t error_capacity) {
  return HostGuard(error, error_capacity, [&]() -> int32_t {
    if (params == nullptr) return BadArgument(error, error_capacity, "null argument");
    PdlpParams p;
    SetDefaultParams(&p);
    const int32_t rc = pdlp_b200_params_merge_bytes(data, size, &p, error, error_capacity);
    if (rc == PDLP_B200_STATUS_OK) *params = p;
    return rc;
  });
}

int32_t pdlp_b200_params_merge_text(const char* text, PdlpParams* params, char* error, int64_t error_capacity) {
  return HostGuard(error, error_capacity, [&]() -> int32_t {
    if (params == nullptr || text == nullptr) return BadArgument(error, error_capacity, "null argument");
    std::string wire, err;
    if (!proto::TextToWire(proto::ParamsSchema(), text, &wire, &err, /*allow_singular_overwrites=*/true)) return BadArgument(error, error_capacity, err);
    return pdlp_b200_params_merge_bytes(reinterpret_cast<const uint8_t*>(wire.data()), static_cast<int64_t>(wire.size()), params, error, error_capacity);
  });
}

int32_t pdlp_b200_params_parse_text(const char* text, PdlpParams* params, char* error, int64_t error_capacity) {
  return HostGuard(error, error_capacity, [&]() -> int32_t {
    if (params == nullptr || text == nullptr) return BadArgument(error, error_capacity, "null argument");
    std::string wire, err;
    if (!proto::TextToWire(proto::ParamsSchema(), text, &wire, &err, /*allow_singular_overwrites=*/false)) return BadArgument(error, error_capacity, err);
    return pdlp_b200_params_parse_bytes(reinterpret_cast<const uint8_t*>(wire.data()), static_cast<int64_t>(wire.size()), params, error, error_capacity);
  });
}

int32_t pdlp_b200_params_serialize(const PdlpParams* params, int32_t format, PdlpBlob* out) {
  return HostGuard(nullptr, 0, [&]() -> int32_t {
    if (params == nullptr || out == nullptr) return PDLP_B200_STATUS_BAD_ARGUMENT;
    Writer w;
    ParamsToWire(*params, &w);
    std::string encoded;
    if (!Encode(proto::ParamsSchema(), w.out(), format, &encoded)) return PDLP_B200_STATUS_BAD_ARGUMENT;
    return ToBlob(encoded, out);
  });
}

int32_t pdlp_b200_solve_log_serialize(const PdlpResult* result, int32_t format, PdlpBlob* out) {
  return HostGuard(nullptr, 0, [&]() -> int32_t {
    if (result == nullptr || out == nullptr) return PDLP_B200_STATUS_BAD_ARGUMENT;
    Writer w;
    SolveLogToWire(*result, &w);
    std::string encoded;
    if (!Encode(proto::SolveLogSchema(), w.out(), format, &encoded)) return PDLP_B200_STATUS_BAD_ARGUMENT;
    return ToBlob(encoded, out);
  });
}

int32_t pdlp_b200_write_solve_log(const PdlpResult* result, const char* path, char* error, int64_t error_capacity) {
  return HostGuard(error, error_capacity, [&]() -> int32_t {
    if (result == nullptr || path == nullptr) return BadArgument(error, error_capacity, "null argument");
    const std::string p = path;
    int32_t format;
    if (EndsWith(p, ".textproto")) format = PDLP_FORMAT_TEXT;
    else if (EndsWith(p, ".pb")) format = PDLP_FORMAT_BINARY;
    else if (EndsWith(p, ".json")) format = PDLP_FORMAT_JSON;
    else return BadArgument(error, error_capacity, "Unrecognized file extension for --solve_log_file: " + p + ". Expected .textproto, .pb, or .json");
    Writer w;
    SolveLogToWire(*result, &w);
    std::string encoded, err;
    if (!Encode(proto::SolveLogSchema(), w.out(), format, &encoded)) return BadArgument(error, error_capacity, "cannot encode the solve log");
    if (!WriteFile(p, encoded, &err)) return BadArgument(error, error_capacity, err);
    return PDLP_B200_STATUS_OK;
  });
}

int32_t pdlp_b200_read_quadratic_program(const char* path, int32_t include_names, PdlpModel** out_model, char* error, int64_t error_capacity) {
  return HostGuard(error, error_capacity, [&]() -> int32_t {
    if (path == nullptr || out_model == nullptr) return BadArgument(error, error_capacity, "null argument");
    auto model = std::make_unique<PdlpModel>();
    std::string err;
    if (!ReadModel(path, include_names != 0, model.get(), &err)) return BadArgument(error, error_capacity, err);
    *out_model = model.release();
    return PDLP_B200_STATUS_OK;
  });
}

int32_t pdlp_b200_model_from_mps_text(const char* text, int64_t size, int32_t include_names, PdlpModel** out_model, char* error,
                                      int64_t error_capacity) {
  return HostGuard(error, error_capacity, [&]() -> int32_t {
    if (text == nullptr || size < 0 || out_model == nullptr) return BadArgument(error, error_capacity, "null argument");
    auto model = std::make_unique<PdlpModel>();
    std::string err;
    if (!ModelFromMpsText(std::string_view(text, static_cast<size_t>(size)), include_names != 0, model.get(), &err))
      return BadArgument(error, error_capacity, err);
    *out_model = model.release();
    return PDLP_B200_STATUS_OK;
  });
}

int32_t pdlp_b200_model_from_mp_model_proto(const uint8_t* data, int64_t size, int32_t relax_integer_variables, int32_t include_names,
                                            PdlpModel** out_model, char* error, int64_t error_capacity) {
  return HostGuard(error, error_capacity, [&]() -> int32_t {
    if (size < 0 || (data == nullptr && size > 0) || out_model == nullptr) return BadArgument(error, error_capacity, "null argument");
    auto model = std::make_unique<PdlpModel>();
    std::string err;
    if (!ModelFromMpModelBytes(std::string_view(reinterpret_cast<const char*>(data), static_cast<size_t>(size)), relax_integer_variables != 0,
                               include_names != 0, model.get(), &err))
      return BadArgument(error, error_capacity, err);
    *out_model = model.release();
    return PDLP_B200_STATUS_OK;
  });
}

const PdlpProblemView* pdlp_b200_model_view(const PdlpModel* model) { return model == nullptr ? nullptr : &model->view; }

const char* pdlp_b200_model_variable_name(const PdlpModel* model, int64_t index) {
  if (model == nullptr || !model->has_names || index < 0 || index >= static_cast<int64_t>(model->variable_names.size())) return nullptr;
  return model->variable_names[static_cast<size_t>(index)].c_str();
}

const char* pdlp_b200_model_constraint_name(const PdlpModel* model, int64_t index) {
  if (model == nullptr || !model->has_names || index < 0 || index >= static_cast<int64_t>(model->constraint_names.size())) return nullptr;
  return model->constraint_names[static_cast<size_t>(index)].c_str();
}

void pdlp_b200_model_free(PdlpModel* model) { delete model; }

int32_t pdlp_b200_qp_to_mp_model_proto(const PdlpProblemView* qp, const char* const* variable_names, const char* const* constraint_names,
                                       PdlpBlob* out, char* error, int64_t error_capacity) {
  return HostGuard(error, error_capacity, [&]() -> int32_t {
    if (qp == nullptr || out == nullptr) return BadArgument(error, error_capacity, "null argument");
    if (const std::string bad = ValidateView(*qp); !bad.empty()) return BadArgument(error, error_capacity, bad);
    std::string wire, err;
    if (!QpToMpModelWire(*qp, variable_names, constraint_names, &wire, &err)) return BadArgument(error, error_capacity, err);
    return ToBlob(wire, out);
  });
}

int32_t pdlp_b200_write_linear_program_to_mps(const PdlpProblemView* qp, const char* const* variable_names, const char* const* constraint_names,
                                              const char* path, char* error, int64_t error_capacity) {
  return HostGuard(error, error_capacity, [&]() -> int32_t {
    if (qp == nullptr || path == nullptr) return BadArgument(error, error_capacity, "null argument");
    if (const std::string bad = ValidateView(*qp); !bad.empty()) return BadArgument(error, error_capacity, bad);
    std::string text, err;
    if (!LinearProgramToMps(*qp, variable_names, constraint_names, &text, &err) || !WriteFile(path, text, &err))
      return BadArgument(error, error_capacity, err);
    return PDLP_B200_STATUS_OK;
  });
}

int32_t pdlp_b200_write_quadratic_program_to_mp_model_proto(const PdlpProblemView* qp, const char* const* variable_names,
                                                            const char* const* constraint_names, const char* path, char* error,
                                                            int64_t error_capacity) {
  return HostGuard(error, error_capacity, [&]() -> int32_t {
    if (qp == nullptr || path == nullptr) return BadArgument(error, error_capacity, "null argument");
    if (const std::string bad = ValidateView(*qp); !bad.empty()) return BadArgument(error, error_capacity, bad);
    std::string wire, err;
    if (!QpToMpModelWire(*qp, variable_names, constraint_names, &wire, &err) || !WriteFile(path, wire, &err))
      return BadArgument(error, error_capacity, err);
    return PDLP_B200_STATUS_OK;
  });
}

int32_t pdlp_b200_solve_proto(const uint8_t* request, int64_t request_size, int32_t relax_integer_variables,
                              const volatile int32_t* interrupt_solve, PdlpBlob* response) {
  return HostGuard(nullptr, 0, [&]() -> int32_t {
    if (request_size < 0 || (request == nullptr && request_size > 0) || response == nullptr) return PDLP_B200_STATUS_BAD_ARGUMENT;
    std::string out;
    const int32_t rc = SolveProto(std::string_view(reinterpret_cast<const char*>(request), static_cast<size_t>(request_size)),
                                  relax_integer_variables != 0, interrupt_solve, &out);
    if (rc != PDLP_B200_STATUS_OK) return rc;
    return ToBlob(out, response);
  });
}

int32_t pdlp_b200_proto_convert(const char* message, int32_t from_format, const uint8_t* data, int64_t size, int32_t to_format, PdlpBlob* out,
                                char* error, int64_t error_capacity) {
  return HostGuard(error, error_capacity, [&]() -> int32_t {
    if (message == nullptr || size < 0 || (data == nullptr && size > 0) || out == nullptr) return BadArgument(error, error_capacity, "null argument");
    const proto::Schema* schema = SchemaByName(message);
    if (schema == nullptr) return BadArgument(error, error_capacity, std::string("unknown message type ") + message);
    const std::string_view in(reinterpret_cast<const char*>(data), static_cast<size_t>(size));
    std::string wire, err;
    switch (from_format) {
      case PDLP_FORMAT_BINARY: wire = std::string(in); break;
      case PDLP_FORMAT_TEXT: if (!proto::TextToWire(*schema, in, &wire, &err)) return BadArgument(error, error_capacity, err); break;
      case PDLP_FORMAT_JSON: if (!proto::JsonToWire(*schema, in, &wire, &err)) return BadArgument(error, error_capacity, err); break;
      default: return BadArgument(error, error_capacity, "unknown input format");
    }
    std::string encoded;
    if (!Encode(*schema, wire, to_format, &encoded)) return BadArgument(error, error_capacity, "malformed message bytes or unknown output format");
    return ToBlob(encoded, out);
  });
}

}  // extern "C"
