"""A four-variable LP through the Python face of the library, written the way a user of
``ortools.pdlp.python.pdlp`` writes it (the counterpart of ``ortools/pdlp/samples/simple_pdlp_program.py``):
only the imports differ. Parameters are a real ``PrimalDualHybridGradientParams`` protobuf message,
built at import time from descriptor tables (there is no generated ``solvers_pb2`` here).

    python examples/solve_simple_lp.py

Needs a CUDA device: without one the call raises (the library has no CPU fallback).
"""
import os
import sys

import numpy as np
import scipy.sparse

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ortools_b200 import pdlp, pdlp_proto  # noqa: E402


def small_lp():
    """min 5.5 x0 - 2 x1 - x2 + x3 - 14  subject to
         2 x0 + x1 + x2 + 2 x3 = 12,   x0 + x2 <= 7,   4 x0 >= -4,   -1 <= 1.5 x2 - x3 <= 1,
         x0 free,  x1 >= -2,  x2 <= 6,  2.5 <= x3 <= 3.5."""
    lp = pdlp.QuadraticProgram()
    lp.objective_offset = -14
    lp.objective_vector = [5.5, -2, -1, 1]
    lp.constraint_lower_bounds = [12, -np.inf, -4, -1]
    lp.constraint_upper_bounds = [12, 7, np.inf, 1]
    lp.variable_lower_bounds = [-np.inf, -2, -np.inf, 2.5]
    lp.variable_upper_bounds = [np.inf, np.inf, 6, 3.5]
    lp.constraint_matrix = scipy.sparse.csc_matrix(np.array([[2, 1, 1, 2], [1, 0, 1, 0], [4, 0, 0, 0], [0, 0, 1.5, -1]], dtype=float))
    return lp


def main():
    params = pdlp_proto.PrimalDualHybridGradientParamsProto()
    criteria = params.termination_criteria.simple_optimality_criteria
    criteria.eps_optimal_relative = 1.0e-6
    criteria.eps_optimal_absolute = 1.0e-6
    params.termination_criteria.time_sec_limit = np.inf
    params.verbosity_level = 0

    result = pdlp.primal_dual_hybrid_gradient(small_lp(), params)
    log = result.solve_log
    if log.termination_reason == pdlp.TerminationReason.TERMINATION_REASON_OPTIMAL:
        print("Solve successful")
    else:
        print("Solve not successful. Status:", pdlp.TerminationReason.Name(log.termination_reason))
    # The vectors are always returned; what they mean depends on the termination reason.
    print("Primal solution:", result.primal_solution)
    print("Dual solution:", result.dual_solution)
    print("Reduced costs:", result.reduced_costs)
    print("Solution type:", pdlp.PointType.Name(log.solution_type))
    info = pdlp.get_convergence_information(log.solution_stats, log.solution_type)
    if info is not None:
        print("Primal objective:", info.primal_objective)
        print("Dual objective:", info.dual_objective)
    print("Iterations:", log.iteration_count)
    print("Solve time (sec):", log.solve_time_sec)


if __name__ == "__main__":
    main()
