// pdlp_b200_io.hpp -- header-only C++17 face of pdlp_b200_io.h, next to pdlp_b200.hpp: the
// reference's file / proto helpers on plain C++ types (strings of bytes instead of protobuf
// objects; errors as std::optional + message instead of absl::Status).
//
//   ortools/pdlp/quadratic_program_io.h:28-59   -> ReadQuadraticProgram[OrDie], WriteLinearProgramToMps, ...
//   ortools/pdlp/quadratic_program.h:160-186    -> QpFromMpModelProto, QpToMpModelProto (serialized MPModelProto)
//   solvers.proto / solve_log.proto text & wire -> ParamsFromText, SerializeParams, SerializeSolveLog
//   proto_solver/pdlp_proto_solver.h            -> PdlpSolveProto (serialized MPModelRequest -> MPSolutionResponse)
#ifndef PDLP_B200_IO_HPP_
#define PDLP_B200_IO_HPP_

#include <cstdio>
#include <cstdlib>
#include <optional>
#include <string>
#include <vector>

#include "pdlp_b200.hpp"
#include "pdlp_b200_io.h"

namespace pdlp_b200 {

enum class ProtoFormat { kBinary = PDLP_FORMAT_BINARY, kText = PDLP_FORMAT_TEXT, kJson = PDLP_FORMAT_JSON };

namespace io_internal {
inline std::string TakeBlob(PdlpBlob* blob) {
  std::string out(reinterpret_cast<const char*>(blob->data), static_cast<size_t>(blob->size));
  pdlp_b200_blob_free(blob);
  return out;
}
inline void SetError(std::string* error, const char* text) {
  if (error != nullptr) *error = text;
}
inline QuadraticProgram FromModel(PdlpModel* model, bool include_names) {
  const PdlpProblemView& v = *pdlp_b200_model_view(model);
  QuadraticProgram qp(v.num_variables, v.num_constraints);
  qp.col_starts.assign(v.col_starts, v.col_starts + v.num_variables + 1);
  qp.row_indices.assign(v.row_indices, v.row_indices + v.num_nonzeros);
  qp.values.assign(v.values, v.values + v.num_nonzeros);
  qp.objective_vector.assign(v.objective_vector, v.objective_vector + v.num_variables);
  if (v.objective_matrix_diagonal != nullptr)
    qp.objective_matrix_diagonal = std::vector<double>(v.objective_matrix_diagonal, v.objective_matrix_diagonal + v.num_variables);
  qp.constraint_lower_bounds.assign(v.constraint_lower_bounds, v.constraint_lower_bounds + v.num_constraints);
  qp.constraint_upper_bounds.assign(v.constraint_upper_bounds, v.constraint_upper_bounds + v.num_constraints);
  qp.variable_lower_bounds.assign(v.variable_lower_bounds, v.variable_lower_bounds + v.num_variables);
  qp.variable_upper_bounds.assign(v.variable_upper_bounds, v.variable_upper_bounds + v.num_variables);
  qp.objective_offset = v.objective_offset;
  qp.objective_scaling_factor = v.objective_scaling_factor;
  if (include_names) {
    qp.problem_name = v.problem_name != nullptr ? v.problem_name : "";
    qp.variable_names.emplace();
    qp.constraint_names.emplace();
    for (int64_t j = 0; j < v.num_variables; ++j) {
      const char* n = pdlp_b200_model_variable_name(model, j);
      qp.variable_names->push_back(n != nullptr ? n : "");
    }
    for (int64_t i = 0; i < v.num_constraints; ++i) {
      const char* n = pdlp_b200_model_constraint_name(model, i);
      qp.constraint_names->push_back(n != nullptr ? n : "");
    }
  }
  pdlp_b200_model_free(model);
  return qp;
}
struct NameArrays {  // the optional name vectors as the char* arrays of the C ABI
  std::vector<const char*> variables, constraints;
  explicit NameArrays(const QuadraticProgram& qp) {
    if (qp.variable_names && static_cast<int64_t>(qp.variable_names->size()) == qp.num_variables())
      for (const std::string& s : *qp.variable_names) variables.push_back(s.c_str());
    if (qp.constraint_names && static_cast<int64_t>(qp.constraint_names->size()) == qp.num_constraints())
      for (const std::string& s : *qp.constraint_names) constraints.push_back(s.c_str());
  }
  const char* const* v() const { return variables.empty() ? nullptr : variables.data(); }
  const char* const* c() const { return constraints.empty() ? nullptr : constraints.data(); }
};
}  // namespace io_internal

// ---- parameters ---------------------------------------------------------------------------------
// Text-format PrimalDualHybridGradientParams merged onto `params` (TextFormat::Merge, what
// pdlp_solve.cc:86 and pdlp_proto_solver.cc:47 do). False + *error on a parse error.
inline bool MergeParamsFromText(const std::string& text, PrimalDualHybridGradientParams* params, std::string* error = nullptr) {
  char buf[1024] = "";
  if (pdlp_b200_params_merge_text(text.c_str(), params, buf, sizeof buf) == PDLP_B200_STATUS_OK) return true;
  io_internal::SetError(error, buf);
  return false;
}
inline std::optional<PrimalDualHybridGradientParams> ParamsFromText(const std::string& text, std::string* error = nullptr) {
  PrimalDualHybridGradientParams p;
  char buf[1024] = "";
  if (pdlp_b200_params_parse_text(text.c_str(), &p, buf, sizeof buf) == PDLP_B200_STATUS_OK) return p;
  io_internal::SetError(error, buf);
  return std::nullopt;
}
inline std::optional<PrimalDualHybridGradientParams> ParamsFromBytes(const std::string& bytes, std::string* error = nullptr) {
  PrimalDualHybridGradientParams p;
  char buf[1024] = "";
  if (pdlp_b200_params_parse_bytes(reinterpret_cast<const uint8_t*>(bytes.data()), static_cast<int64_t>(bytes.size()), &p, buf, sizeof buf) ==
      PDLP_B200_STATUS_OK)
    return p;
  io_internal::SetError(error, buf);
  return std::nullopt;
}
inline std::string SerializeParams(const PdlpParams& params, ProtoFormat format = ProtoFormat::kBinary) {
  PdlpBlob blob{};
  if (pdlp_b200_params_serialize(&params, static_cast<int32_t>(format), &blob) != PDLP_B200_STATUS_OK) return {};
  return io_internal::TakeBlob(&blob);
}

// ---- SolveLog -------------------------------------------------------------------------------------
// The SolveLog message of `log` (what pdlp_proto_solver.cc:127 puts into solver_specific_info and
// pdlp_solve.cc writes to --solve_log_file) as bytes, text or JSON.
inline std::string SerializeSolveLog(const SolveLog& log, ProtoFormat format = ProtoFormat::kBinary) {
  PdlpResult r{};
  r.instance_name = log.instance_name ? const_cast<char*>(log.instance_name->c_str()) : nullptr;
  r.termination_string = log.termination_string ? const_cast<char*>(log.termination_string->c_str()) : nullptr;
  r.termination_reason = log.termination_reason;
  r.iteration_count = log.iteration_count;
  r.solve_time_sec = log.solve_time_sec;
  r.preprocessing_time_sec = log.preprocessing_time_sec;
  r.solution_type = log.solution_type;
  if (log.solution_stats) {
    r.has_solution_stats = 1;
    r.solution_stats = *log.solution_stats;
  }
  if (log.original_problem_stats) {
    r.has_original_problem_stats = 1;
    r.original_problem_stats = *log.original_problem_stats;
  }
  if (log.preprocessed_problem_stats) {
    r.has_preprocessed_problem_stats = 1;
    r.preprocessed_problem_stats = *log.preprocessed_problem_stats;
  }
  r.num_iteration_stats = static_cast<int64_t>(log.iteration_stats.size());
  r.iteration_stats = const_cast<PdlpIterationStats*>(log.iteration_stats.data());
  r.params = log.params;
  std::vector<PdlpFeasibilityPolishingDetails> details;
  for (const FeasibilityPolishingDetails& d : log.feasibility_polishing_details) {
    PdlpFeasibilityPolishingDetails o{};
    o.polishing_phase_type = d.polishing_phase_type;
    o.main_iteration_count = d.main_iteration_count;
    o.params = d.params;
    o.termination_reason = d.termination_reason;
    o.iteration_count = d.iteration_count;
    o.solve_time_sec = d.solve_time_sec;
    o.solution_stats = d.solution_stats;
    o.solution_type = d.solution_type;
    o.num_iteration_stats = static_cast<int64_t>(d.iteration_stats.size());
    o.iteration_stats = const_cast<PdlpIterationStats*>(d.iteration_stats.data());
    details.push_back(o);
  }
  r.num_feasibility_polishing_details = static_cast<int64_t>(details.size());
  r.feasibility_polishing_details = details.data();
  PdlpBlob blob{};
  if (pdlp_b200_solve_log_serialize(&r, static_cast<int32_t>(format), &blob) != PDLP_B200_STATUS_OK) return {};
  return io_internal::TakeBlob(&blob);
}

// ---- problems -------------------------------------------------------------------------------------
// quadratic_program_io.h:28-40. nullopt + *error instead of dying.
inline std::optional<QuadraticProgram> ReadQuadraticProgram(const std::string& filename, bool include_names = false, std::string* error = nullptr) {
  PdlpModel* model = nullptr;
  char buf[1024] = "";
  if (pdlp_b200_read_quadratic_program(filename.c_str(), include_names ? 1 : 0, &model, buf, sizeof buf) != PDLP_B200_STATUS_OK) {
    io_internal::SetError(error, buf);
    return std::nullopt;
  }
  return io_internal::FromModel(model, include_names);
}
inline QuadraticProgram ReadQuadraticProgramOrDie(const std::string& filename, bool include_names = false) {
  std::string error;
  std::optional<QuadraticProgram> qp = ReadQuadraticProgram(filename, include_names, &error);
  if (!qp) {
    std::fprintf(stderr, "%s\n", error.c_str());
    std::abort();
  }
  return *std::move(qp);
}
// quadratic_program.h:160-170 on serialized MPModelProto bytes.
inline std::optional<QuadraticProgram> QpFromMpModelProto(const std::string& serialized_model, bool relax_integer_variables, bool include_names = false,
                                                          std::string* error = nullptr) {
  PdlpModel* model = nullptr;
  char buf[1024] = "";
  if (pdlp_b200_model_from_mp_model_proto(reinterpret_cast<const uint8_t*>(serialized_model.data()), static_cast<int64_t>(serialized_model.size()),
                                          relax_integer_variables ? 1 : 0, include_names ? 1 : 0, &model, buf, sizeof buf) != PDLP_B200_STATUS_OK) {
    io_internal::SetError(error, buf);
    return std::nullopt;
  }
  return io_internal::FromModel(model, include_names);
}
// quadratic_program.h:182-186: the serialized MPModelProto of `qp`.
inline std::optional<std::string> QpToMpModelProto(const QuadraticProgram& qp, std::string* error = nullptr) {
  const PdlpProblemView view = qp.View();
  const io_internal::NameArrays names(qp);
  PdlpBlob blob{};
  char buf[1024] = "";
  if (pdlp_b200_qp_to_mp_model_proto(&view, names.v(), names.c(), &blob, buf, sizeof buf) != PDLP_B200_STATUS_OK) {
    io_internal::SetError(error, buf);
    return std::nullopt;
  }
  return io_internal::TakeBlob(&blob);
}
// quadratic_program_io.h:47-59
inline bool WriteLinearProgramToMps(const QuadraticProgram& linear_program, const std::string& mps_file, std::string* error = nullptr) {
  const PdlpProblemView view = linear_program.View();
  const io_internal::NameArrays names(linear_program);
  char buf[1024] = "";
  if (pdlp_b200_write_linear_program_to_mps(&view, names.v(), names.c(), mps_file.c_str(), buf, sizeof buf) == PDLP_B200_STATUS_OK) return true;
  io_internal::SetError(error, buf);
  return false;
}
inline bool WriteQuadraticProgramToMPModelProto(const QuadraticProgram& quadratic_program, const std::string& mpmodel_proto_file,
                                                std::string* error = nullptr) {
  const PdlpProblemView view = quadratic_program.View();
  const io_internal::NameArrays names(quadratic_program);
  char buf[1024] = "";
  if (pdlp_b200_write_quadratic_program_to_mp_model_proto(&view, names.v(), names.c(), mpmodel_proto_file.c_str(), buf, sizeof buf) ==
      PDLP_B200_STATUS_OK)
    return true;
  io_internal::SetError(error, buf);
  return false;
}

// ---- PdlpSolveProto (pdlp_proto_solver.h) -----------------------------------------------------------
// Serialized MPModelRequest -> serialized MPSolutionResponse. nullopt when the solve could not run
// on a device (no CPU fallback); invalid parameters / models are statuses inside the response.
inline std::optional<std::string> PdlpSolveProto(const std::string& serialized_request, bool relax_integer_variables = false,
                                                 const volatile int32_t* interrupt_solve = nullptr) {
  PdlpBlob blob{};
  if (pdlp_b200_solve_proto(reinterpret_cast<const uint8_t*>(serialized_request.data()), static_cast<int64_t>(serialized_request.size()),
                            relax_integer_variables ? 1 : 0, interrupt_solve, &blob) != PDLP_B200_STATUS_OK)
    return std::nullopt;
  return io_internal::TakeBlob(&blob);
}

}  // namespace pdlp_b200

#endif  // PDLP_B200_IO_HPP_
