// A four-variable LP solved on the GPU through include/pdlp_b200.hpp. It plays the role of the
// reference's direct-API sample (ortools/pdlp/samples/simple_pdlp_program.cc): same problem and
// parameter choices, same report, so the two outputs can be compared line by line.
//
//   g++ -std=c++17 -Iinclude examples/solve_simple_lp.cc -Lor-tools_b200/lib -lpdlp_b200
//       -Wl,-rpath,$PWD/or-tools_b200/lib -o solve_simple_lp && ./solve_simple_lp
#include <iostream>
#include <optional>
#include <vector>

#include "pdlp_b200.hpp"

namespace pdlp = ::pdlp_b200;
using pdlp::kInfinity;

// minimise c'x - 14 with c = (5.5, -2, -1, 1) subject to l_c <= K x <= u_c, l_v <= x <= u_v where
//       | 2  1  1    2 |        l_c = (12, -inf, -4, -1)     l_v = (-inf, -2, -inf, 2.5)
//   K = | 1  0  1    0 |        u_c = (12,    7, inf,  1)     u_v = ( inf, inf,    6, 3.5)
//       | 4  0  0    0 |
//       | 0  0  1.5 -1 |
// (the TestLp of the reference's test_util.cc; optimum x = (-1, 8, 1, 2.5), objective -34).
pdlp::QuadraticProgram FourVariableLp() {
  pdlp::QuadraticProgram lp(4, 4);
  lp.constraint_lower_bounds = {12, -kInfinity, -4, -1};
  lp.constraint_upper_bounds = {12, 7, kInfinity, 1};
  lp.variable_lower_bounds = {-kInfinity, -2, -kInfinity, 2.5};
  lp.variable_upper_bounds = {kInfinity, kInfinity, 6, 3.5};
  lp.SetConstraintMatrixFromTriplets({{0, 0, 2}, {0, 1, 1}, {0, 2, 1}, {0, 3, 2}, {1, 0, 1}, {1, 2, 1}, {2, 0, 4}, {3, 2, 1.5}, {3, 3, -1}});
  lp.objective_vector = {5.5, -2, -1, 1};
  lp.objective_offset = -14;
  return lp;
}

static void Print(const char* title, const std::vector<double>& v) {
  std::cout << title << '\n';
  for (double x : v) std::cout << x << '\n';
}

int main() {
  pdlp::PrimalDualHybridGradientParams params;
  // The knobs callers usually touch, set to their default values.
  params.SetSimpleOptimalityCriteria(/*eps_optimal_absolute=*/1.0e-6, /*eps_optimal_relative=*/1.0e-6);
  params.termination_criteria.time_sec_limit = kInfinity;
  params.num_threads = 1;
  params.verbosity_level = 0;
  params.presolve_use_glop = 0;

  const pdlp::SolverResult result = pdlp::PrimalDualHybridGradient(FourVariableLp(), params);
  const pdlp::SolveLog& solve_log = result.solve_log;

  if (solve_log.termination_reason == PDLP_TERMINATION_REASON_OPTIMAL) {
    std::cout << "Solve successful" << '\n';
  } else {
    std::cout << "Solve not successful. Status: " << pdlp::TerminationReason_Name(solve_log.termination_reason) << '\n';
    if (solve_log.termination_string) std::cout << *solve_log.termination_string << '\n';
  }
  // The three vectors are filled whatever the outcome; what they mean depends on the
  // termination reason (primal_dual_hybrid_gradient.h:36-71).
  Print("Primal solution:", result.primal_solution);
  Print("Dual solution:", result.dual_solution);
  Print("Reduced costs:", result.reduced_costs);

  const int solution_type = solve_log.solution_type;
  std::cout << "Solution type: " << pdlp::PointType_Name(solution_type) << '\n';
  if (solve_log.solution_stats) {
    if (const auto ci = pdlp::GetConvergenceInformation(*solve_log.solution_stats, solution_type); ci.has_value()) {
      std::cout << "Primal objective: " << ci->primal_objective << '\n';
      std::cout << "Dual objective: " << ci->dual_objective << '\n';
    }
  }
  std::cout << "Iterations: " << solve_log.iteration_count << '\n';
  std::cout << "Solve time (sec): " << solve_log.solve_time_sec << '\n';
  return solve_log.termination_reason == PDLP_TERMINATION_REASON_OPTIMAL ? 0 : 1;
}
