// pdlp_solve_cli.cc -- command-line front end of the B200 PDLP library: reads a
// model file, solves it on the GPU and writes the log / solution. The
// counterpart of the reference's examples/cpp/pdlp_solve.cc with the same four
// flags and file conventions, written against the C ABI only
// (include/pdlp_b200.h, include/pdlp_b200_io.h): no protobuf, no absl.
//
//   pdlp_solve --input=model.mps[.gz] | model.pb | model.textproto | model.json[.gz]
//              [--params='termination_criteria { simple_optimality_criteria { eps_optimal_relative: 1e-4 } }']
//              [--solve_log_file=log.textproto|log.pb|log.json] [--sol_file=out.sol]
//
// Exit status: 0 after a solve (whatever its termination reason), 1 for bad
// flags / unreadable input / no usable CUDA device (there is no CPU fallback).
#include <csignal>
#include <cstdio>
#include <cstring>
#include <string>

#include "pdlp_b200_io.h"

namespace {

volatile int32_t g_interrupted = 0;
void OnSigint(int) { g_interrupted = 1; }  // polled by the solver like the reference's std::atomic<bool>

void LogLine(const char* message, void*) { std::fprintf(stderr, "%s\n", message); }

const char* TerminationReasonName(int32_t reason) {
  static const char* kNames[] = {"TERMINATION_REASON_UNSPECIFIED", "TERMINATION_REASON_OPTIMAL", "TERMINATION_REASON_PRIMAL_INFEASIBLE",
                                 "TERMINATION_REASON_DUAL_INFEASIBLE", "TERMINATION_REASON_TIME_LIMIT", "TERMINATION_REASON_ITERATION_LIMIT",
                                 "TERMINATION_REASON_NUMERICAL_ERROR", "TERMINATION_REASON_OTHER", "TERMINATION_REASON_KKT_MATRIX_PASS_LIMIT",
                                 "TERMINATION_REASON_INVALID_PROBLEM", "TERMINATION_REASON_INVALID_PARAMETER",
                                 "TERMINATION_REASON_PRIMAL_OR_DUAL_INFEASIBLE", "TERMINATION_REASON_INTERRUPTED_BY_USER",
                                 "TERMINATION_REASON_INVALID_INITIAL_SOLUTION"};
  return reason >= 0 && reason < 14 ? kNames[reason] : "?";
}

// --flag=value or --flag value; returns false if argv[*i] is not `name`.
bool Flag(int argc, char** argv, int* i, const char* name, std::string* value) {
  const std::string arg = argv[*i];
  const std::string dashed = std::string("--") + name;
  if (arg.rfind(dashed + "=", 0) == 0) {
    *value = arg.substr(dashed.size() + 1);
    return true;
  }
  if (arg == dashed && *i + 1 < argc) {
    *value = argv[++*i];
    return true;
  }
  return false;
}

// Shortest text that reads back to the same double (RoundTripDoubleFormat).
std::string Number(double v) {
  char buf[40];
  for (int prec = 15; prec <= 17; ++prec) {
    std::snprintf(buf, sizeof buf, "%.*g", prec, v);
    double back = 0;
    if (std::sscanf(buf, "%lf", &back) == 1 && back == v) break;
  }
  return buf;
}

}  // namespace

int main(int argc, char** argv) {
  std::string input, params_text, solve_log_file, sol_file;
  for (int i = 1; i < argc; ++i) {
    if (Flag(argc, argv, &i, "input", &input) || Flag(argc, argv, &i, "params", &params_text) ||
        Flag(argc, argv, &i, "solve_log_file", &solve_log_file) || Flag(argc, argv, &i, "sol_file", &sol_file))
      continue;
    std::fprintf(stderr, "unknown argument %s\nusage: %s --input=FILE [--params=TEXT] [--solve_log_file=FILE] [--sol_file=FILE]\n", argv[i], argv[0]);
    return 1;
  }
  if (input.empty()) {
    std::fprintf(stderr, "--input is required\n");
    return 1;
  }
  char error[1024] = "";
  // Print iteration statistics by default; verbosity_level in --params overrides it.
  PdlpParams params;
  pdlp_b200_params_set_defaults(&params);
  params.verbosity_level = 2;
  if (pdlp_b200_params_merge_text(params_text.c_str(), &params, error, sizeof error) != PDLP_B200_STATUS_OK) {
    std::fprintf(stderr, "Error parsing --params: %s\n", error);
    return 1;
  }
  if (!solve_log_file.empty()) {  // fail before the solve, not after it
    const auto ends = [&](const char* s) { const size_t n = std::strlen(s); return solve_log_file.size() >= n && solve_log_file.compare(solve_log_file.size() - n, n, s) == 0; };
    if (!ends(".textproto") && !ends(".pb") && !ends(".json")) {
      std::fprintf(stderr, "Unrecognized file extension for --solve_log_file: %s. Expected .textproto, .pb, or .json\n", solve_log_file.c_str());
      return 1;
    }
  }
  // Integrality constraints are dropped by the reader.
  PdlpModel* model = nullptr;
  if (pdlp_b200_read_quadratic_program(input.c_str(), /*include_names=*/1, &model, error, sizeof error) != PDLP_B200_STATUS_OK) {
    std::fprintf(stderr, "%s\n", error);
    return 1;
  }
  std::signal(SIGINT, OnSigint);  // ^C interrupts the solve, the results so far are still written
  PdlpResult result;
  std::memset(&result, 0, sizeof result);
  const int32_t rc = pdlp_b200_primal_dual_hybrid_gradient(pdlp_b200_model_view(model), &params, nullptr, 0, nullptr, 0, &g_interrupted, LogLine,
                                                           nullptr, nullptr, &result);
  if (rc != PDLP_B200_STATUS_OK) {
    std::fprintf(stderr, "%s\n", rc == PDLP_B200_STATUS_NO_DEVICE ? "no usable CUDA device (this library has no CPU fallback)"
                                                                  : "the solve failed with a CUDA / NCCL error");
    pdlp_b200_model_free(model);
    return 1;
  }
  std::fprintf(stderr, "Termination reason: %s, %d iterations, %.3f s\n", TerminationReasonName(result.termination_reason), result.iteration_count,
               result.solve_time_sec);
  int exit_code = 0;
  if (!solve_log_file.empty()) {
    std::fprintf(stderr, "Writing SolveLog to '%s'.\n", solve_log_file.c_str());
    if (pdlp_b200_write_solve_log(&result, solve_log_file.c_str(), error, sizeof error) != PDLP_B200_STATUS_OK) {
      std::fprintf(stderr, "%s\n", error);
      exit_code = 1;
    }
  }
  // The primal solution in Miplib .sol format, if the log has convergence
  // information for the returned point.
  const PdlpConvergenceInformation* info = nullptr;
  if (result.has_solution_stats)
    for (int k = 0; k < result.solution_stats.num_convergence_information; ++k)
      if (result.solution_stats.convergence_information[k].candidate_type == result.solution_type) {
        info = &result.solution_stats.convergence_information[k];
        break;
      }
  if (!sol_file.empty() && info != nullptr) {
    std::fprintf(stderr, "Writing .sol solution to '%s'.\n", sol_file.c_str());
    std::FILE* f = std::fopen(sol_file.c_str(), "w");
    if (f == nullptr) {
      std::fprintf(stderr, "cannot open %s for writing\n", sol_file.c_str());
      exit_code = 1;
    } else {
      std::fprintf(f, "=obj= %s\n", Number(info->primal_objective).c_str());
      for (int64_t j = 0; j < result.primal_size; ++j) {
        const char* name = pdlp_b200_model_variable_name(model, j);
        if (name != nullptr && name[0] != '\0') std::fprintf(f, "%s %s\n", name, Number(result.primal_solution[j]).c_str());
        else std::fprintf(f, "var%lld %s\n", static_cast<long long>(j), Number(result.primal_solution[j]).c_str());
      }
      if (std::fclose(f) != 0) exit_code = 1;
    }
  }
  pdlp_b200_result_free(&result);
  pdlp_b200_model_free(model);
  return exit_code;
}
