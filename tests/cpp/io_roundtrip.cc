// io_roundtrip.cc -- exercises include/pdlp_b200_io.hpp (host-only entry points) from C++; driven by
// tests/test_cpp_example.py, which checks the lines printed here. argv[1] = a scratch directory.
#include <cstdio>
#include <string>

#include "pdlp_b200_io.hpp"

using namespace pdlp_b200;  // NOLINT

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  const std::string dir = argv[1];
  std::string error;
  // parameters: text -> POD -> text
  PrimalDualHybridGradientParams params;
  params.verbosity_level = 2;
  if (!MergeParamsFromText("termination_criteria { simple_optimality_criteria { eps_optimal_relative: 1e-4 } iteration_limit: 100 } verbosity_level: 0",
                           &params, &error))
    return 3;
  std::printf("verbosity=%d limit=%d case=%d eps=%g\n", params.verbosity_level, params.termination_criteria.iteration_limit,
              params.termination_criteria.optimality_criteria_case, params.termination_criteria.simple_eps_optimal_relative);
  std::printf("params-text-begin\n%sparams-text-end\n", SerializeParams(params, ProtoFormat::kText).c_str());
  if (ParamsFromText("no_such_field: 1", &error)) return 4;
  std::printf("error: %s\n", error.c_str());
  const auto reparsed = ParamsFromBytes(SerializeParams(params));
  if (!reparsed || reparsed->termination_criteria.iteration_limit != 100) return 5;
  // a model: build -> MPS -> read -> MPModelProto bytes -> read
  QuadraticProgram lp(2, 1);  // min x + 2y  s.t.  x + y >= 1, 0 <= x, y <= 4
  lp.objective_vector = {1.0, 2.0};
  lp.SetConstraintMatrixFromTriplets({{0, 0, 1.0}, {0, 1, 1.0}});
  lp.constraint_lower_bounds = {1.0};
  lp.variable_lower_bounds = {0.0, 0.0};
  lp.variable_upper_bounds = {4.0, 4.0};
  lp.problem_name = "tiny";
  lp.variable_names = std::vector<std::string>{"x", "y"};
  lp.constraint_names = std::vector<std::string>{"cover"};
  if (!WriteLinearProgramToMps(lp, dir + "/tiny.mps", &error)) return 6;
  const auto from_mps = ReadQuadraticProgram(dir + "/tiny.mps", /*include_names=*/true, &error);
  if (!from_mps) return 7;
  std::printf("mps: n=%lld m=%lld nnz=%zu name=%s var1=%s con0=%s c=[%g, %g] lc=%g uv=%g\n", static_cast<long long>(from_mps->num_variables()),
              static_cast<long long>(from_mps->num_constraints()), from_mps->values.size(), from_mps->problem_name->c_str(),
              (*from_mps->variable_names)[1].c_str(), (*from_mps->constraint_names)[0].c_str(), from_mps->objective_vector[0],
              from_mps->objective_vector[1], from_mps->constraint_lower_bounds[0], from_mps->variable_upper_bounds[1]);
  const auto bytes = QpToMpModelProto(lp, &error);
  if (!bytes) return 8;
  const auto from_proto = QpFromMpModelProto(*bytes, /*relax_integer_variables=*/false, /*include_names=*/true, &error);
  if (!from_proto || from_proto->values != lp.values || from_proto->col_starts != lp.col_starts || *from_proto->variable_names != *lp.variable_names) return 9;
  if (!WriteQuadraticProgramToMPModelProto(lp, dir + "/tiny.pb", &error)) return 10;
  const auto from_pb = ReadQuadraticProgram(dir + "/tiny.pb");
  if (!from_pb || from_pb->objective_vector != lp.objective_vector) return 11;
  if (ReadQuadraticProgram(dir + "/tiny.lp", false, &error)) return 12;
  std::printf("error: %s\n", error.c_str());
  // a solve log
  SolveLog log;
  log.instance_name = "tiny";
  log.termination_reason = PDLP_TERMINATION_REASON_OPTIMAL;
  log.iteration_count = 12;
  log.params = params;
  std::printf("log-json-begin\n%slog-json-end\n", SerializeSolveLog(log, ProtoFormat::kJson).c_str());
  // the proto solver refuses a request without a model before it needs a device
  const auto response = PdlpSolveProto("");
  std::printf("response bytes=%zu\n", response ? response->size() : 0);
  return 0;
}
