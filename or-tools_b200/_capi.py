"""ctypes mirror of include/pdlp_b200.h (POD structs + function prototypes).

The struct layouts are the interface of the C-ABI boundary (entry points with the
prefix ``pdlp_b200_``); a library that exports the same entry points under another
prefix can be bound with the same prototypes (``bind(lib, prefix)``).
"""
import ctypes as C

import numpy as np

MAX_SEEDS = 8

c_double_p = C.POINTER(C.c_double)
c_int64_p = C.POINTER(C.c_int64)


class PdlpTerminationCriteria(C.Structure):
    _fields_ = [
        ("optimality_norm", C.c_int32),
        ("optimality_criteria_case", C.c_int32),
        ("simple_eps_optimal_absolute", C.c_double),
        ("simple_eps_optimal_relative", C.c_double),
        ("eps_optimal_primal_residual_absolute", C.c_double),
        ("eps_optimal_primal_residual_relative", C.c_double),
        ("eps_optimal_dual_residual_absolute", C.c_double),
        ("eps_optimal_dual_residual_relative", C.c_double),
        ("eps_optimal_objective_gap_absolute", C.c_double),
        ("eps_optimal_objective_gap_relative", C.c_double),
        ("has_eps_optimal_absolute", C.c_int32),
        ("has_eps_optimal_relative", C.c_int32),
        ("eps_optimal_absolute", C.c_double),
        ("eps_optimal_relative", C.c_double),
        ("eps_primal_infeasible", C.c_double),
        ("eps_dual_infeasible", C.c_double),
        ("time_sec_limit", C.c_double),
        ("iteration_limit", C.c_int32),
        ("kkt_matrix_pass_limit", C.c_double),
    ]


class PdlpParams(C.Structure):
    _fields_ = [
        ("termination_criteria", PdlpTerminationCriteria),
        ("num_threads", C.c_int32),
        ("num_shards", C.c_int32),
        ("scheduler_type", C.c_int32),
        ("record_iteration_stats", C.c_int32),
        ("verbosity_level", C.c_int32),
        ("log_interval_seconds", C.c_double),
        ("major_iteration_frequency", C.c_int32),
        ("termination_check_frequency", C.c_int32),
        ("restart_strategy", C.c_int32),
        ("primal_weight_update_smoothing", C.c_double),
        ("has_initial_primal_weight", C.c_int32),
        ("initial_primal_weight", C.c_double),
        ("l_inf_ruiz_iterations", C.c_int32),
        ("l2_norm_rescaling", C.c_int32),
        ("sufficient_reduction_for_restart", C.c_double),
        ("necessary_reduction_for_restart", C.c_double),
        ("linesearch_rule", C.c_int32),
        ("adaptive_step_size_reduction_exponent", C.c_double),
        ("adaptive_step_size_growth_exponent", C.c_double),
        ("malitsky_pock_step_size_downscaling_factor", C.c_double),
        ("malitsky_pock_linesearch_contraction_factor", C.c_double),
        ("malitsky_pock_step_size_interpolation", C.c_double),
        ("initial_step_size_scaling", C.c_double),
        ("infinite_constraint_bound_threshold", C.c_double),
        ("handle_some_primal_gradients_on_finite_bounds_as_residuals", C.c_int32),
        ("use_diagonal_qp_trust_region_solver", C.c_int32),
        ("diagonal_qp_trust_region_solver_tolerance", C.c_double),
        ("num_random_projection_seeds", C.c_int32),
        ("random_projection_seeds", C.c_int32 * MAX_SEEDS),
        ("presolve_use_glop", C.c_int32),
        ("use_feasibility_polishing", C.c_int32),
        ("apply_feasibility_polishing_after_limits_reached", C.c_int32),
        ("apply_feasibility_polishing_if_solver_is_interrupted", C.c_int32),
    ]


class PdlpProblemView(C.Structure):
    _fields_ = [
        ("num_variables", C.c_int64),
        ("num_constraints", C.c_int64),
        ("num_nonzeros", C.c_int64),
        ("col_starts", c_int64_p),
        ("row_indices", c_int64_p),
        ("values", c_double_p),
        ("objective_vector", c_double_p),
        ("objective_matrix_diagonal", c_double_p),
        ("constraint_lower_bounds", c_double_p),
        ("constraint_upper_bounds", c_double_p),
        ("variable_lower_bounds", c_double_p),
        ("variable_upper_bounds", c_double_p),
        ("objective_offset", C.c_double),
        ("objective_scaling_factor", C.c_double),
        ("problem_name", C.c_char_p),
        ("objective_vector_size", C.c_int64),
        ("objective_matrix_size", C.c_int64),
        ("constraint_lower_bounds_size", C.c_int64),
        ("constraint_upper_bounds_size", C.c_int64),
        ("variable_lower_bounds_size", C.c_int64),
        ("variable_upper_bounds_size", C.c_int64),
    ]


_STATS_FIELDS = [
    ("num_variables", C.c_int64), ("num_constraints", C.c_int64),
    ("constraint_matrix_col_min_l_inf_norm", C.c_double), ("constraint_matrix_row_min_l_inf_norm", C.c_double),
    ("constraint_matrix_num_nonzeros", C.c_int64),
    ("constraint_matrix_abs_max", C.c_double), ("constraint_matrix_abs_min", C.c_double),
    ("constraint_matrix_abs_avg", C.c_double), ("constraint_matrix_l2_norm", C.c_double),
    ("combined_bounds_max", C.c_double), ("combined_bounds_min", C.c_double),
    ("combined_bounds_avg", C.c_double), ("combined_bounds_l2_norm", C.c_double),
    ("combined_variable_bounds_max", C.c_double), ("combined_variable_bounds_min", C.c_double),
    ("combined_variable_bounds_avg", C.c_double), ("combined_variable_bounds_l2_norm", C.c_double),
    ("variable_bound_gaps_num_finite", C.c_int64),
    ("variable_bound_gaps_max", C.c_double), ("variable_bound_gaps_min", C.c_double),
    ("variable_bound_gaps_avg", C.c_double), ("variable_bound_gaps_l2_norm", C.c_double),
    ("objective_vector_abs_max", C.c_double), ("objective_vector_abs_min", C.c_double),
    ("objective_vector_abs_avg", C.c_double), ("objective_vector_l2_norm", C.c_double),
    ("objective_matrix_num_nonzeros", C.c_int64),
    ("objective_matrix_abs_max", C.c_double), ("objective_matrix_abs_min", C.c_double),
    ("objective_matrix_abs_avg", C.c_double), ("objective_matrix_l2_norm", C.c_double),
]


class PdlpQuadraticProgramStats(C.Structure):
    _fields_ = _STATS_FIELDS


class PdlpConvergenceInformation(C.Structure):
    _fields_ = [
        ("candidate_type", C.c_int32),
        ("primal_objective", C.c_double), ("dual_objective", C.c_double), ("corrected_dual_objective", C.c_double),
        ("l_inf_primal_residual", C.c_double), ("l2_primal_residual", C.c_double),
        ("l_inf_componentwise_primal_residual", C.c_double),
        ("l_inf_dual_residual", C.c_double), ("l2_dual_residual", C.c_double),
        ("l_inf_componentwise_dual_residual", C.c_double),
        ("l_inf_primal_variable", C.c_double), ("l2_primal_variable", C.c_double),
        ("l_inf_dual_variable", C.c_double), ("l2_dual_variable", C.c_double),
    ]


class PdlpInfeasibilityInformation(C.Structure):
    _fields_ = [
        ("candidate_type", C.c_int32),
        ("max_primal_ray_infeasibility", C.c_double), ("primal_ray_linear_objective", C.c_double),
        ("primal_ray_quadratic_norm", C.c_double),
        ("max_dual_ray_infeasibility", C.c_double), ("dual_ray_objective", C.c_double),
    ]


class PdlpPointMetadata(C.Structure):
    _fields_ = [
        ("point_type", C.c_int32),
        ("num_random_projections", C.c_int32),
        ("random_primal_projections", C.c_double * MAX_SEEDS),
        ("random_dual_projections", C.c_double * MAX_SEEDS),
        ("has_active_set_information", C.c_int32),
        ("active_primal_variable_count", C.c_int64), ("active_dual_variable_count", C.c_int64),
        ("active_primal_variable_change", C.c_int64), ("active_dual_variable_change", C.c_int64),
    ]


class PdlpIterationStats(C.Structure):
    _fields_ = [
        ("iteration_number", C.c_int32),
        ("num_convergence_information", C.c_int32),
        ("convergence_information", PdlpConvergenceInformation * 3),
        ("num_infeasibility_information", C.c_int32),
        ("infeasibility_information", PdlpInfeasibilityInformation * 3),
        ("num_point_metadata", C.c_int32),
        ("point_metadata", PdlpPointMetadata * 3),
        ("cumulative_kkt_matrix_passes", C.c_double),
        ("cumulative_rejected_steps", C.c_int32),
        ("cumulative_time_sec", C.c_double),
        ("restart_used", C.c_int32),
        ("step_size", C.c_double),
        ("primal_weight", C.c_double),
    ]


class PdlpBoundNorms(C.Structure):
    _fields_ = [
        ("l2_norm_primal_linear_objective", C.c_double), ("l2_norm_constraint_bounds", C.c_double),
        ("l_inf_norm_primal_linear_objective", C.c_double), ("l_inf_norm_constraint_bounds", C.c_double),
    ]


class PdlpIterationCallbackInfo(C.Structure):
    _fields_ = [
        ("iteration_type", C.c_int32),
        ("termination_criteria", C.POINTER(PdlpTerminationCriteria)),
        ("iteration_stats", C.POINTER(PdlpIterationStats)),
        ("bound_norms", PdlpBoundNorms),
    ]


class PdlpFeasibilityPolishingDetails(C.Structure):
    _fields_ = [
        ("polishing_phase_type", C.c_int32), ("main_iteration_count", C.c_int32),
        ("params", PdlpParams),
        ("termination_reason", C.c_int32), ("iteration_count", C.c_int32),
        ("solve_time_sec", C.c_double),
        ("solution_stats", PdlpIterationStats),
        ("solution_type", C.c_int32),
        ("num_iteration_stats", C.c_int64),
        ("iteration_stats", C.POINTER(PdlpIterationStats)),
    ]


class PdlpResult(C.Structure):
    _fields_ = [
        ("primal_size", C.c_int64), ("dual_size", C.c_int64),
        ("primal_solution", c_double_p), ("dual_solution", c_double_p), ("reduced_costs", c_double_p),
        ("instance_name", C.c_void_p),
        ("termination_reason", C.c_int32),
        ("termination_string", C.c_void_p),
        ("iteration_count", C.c_int32),
        ("solve_time_sec", C.c_double),
        ("preprocessing_time_sec", C.c_double),
        ("solution_type", C.c_int32),
        ("has_solution_stats", C.c_int32),
        ("solution_stats", PdlpIterationStats),
        ("has_original_problem_stats", C.c_int32), ("has_preprocessed_problem_stats", C.c_int32),
        ("original_problem_stats", PdlpQuadraticProgramStats),
        ("preprocessed_problem_stats", PdlpQuadraticProgramStats),
        ("num_iteration_stats", C.c_int64),
        ("iteration_stats", C.POINTER(PdlpIterationStats)),
        ("params", PdlpParams),
        ("num_feasibility_polishing_details", C.c_int64),
        ("feasibility_polishing_details", C.POINTER(PdlpFeasibilityPolishingDetails)),
        ("gpu_kernel_launches", C.c_int64),
        ("device_iteration_time_sec", C.c_double),
    ]


class PdlpSessionStatus(C.Structure):
    _fields_ = [
        ("terminated", C.c_int32), ("termination_reason", C.c_int32),
        ("iterations_completed", C.c_int32), ("num_rejected_steps", C.c_int32),
        ("step_size", C.c_double), ("primal_weight", C.c_double),
        ("gpu_kernel_launches", C.c_int64),
        ("device_step_ms", C.c_double), ("device_total_ms", C.c_double),
        ("kernel_ms", C.c_double * 4), ("kernel_samples", C.c_int64 * 4),
        ("kernel_algorithmic_bytes", C.c_double * 4),
    ]


MESSAGE_CALLBACK = C.CFUNCTYPE(None, C.c_char_p, C.c_void_p)
STATS_CALLBACK = C.CFUNCTYPE(None, C.POINTER(PdlpIterationCallbackInfo), C.c_void_p)


def struct_to_dict(s):
    """Recursively converts a ctypes Structure into plain Python objects."""
    out = {}
    for name, typ in s._fields_:
        v = getattr(s, name)
        if isinstance(v, C.Structure):
            out[name] = struct_to_dict(v)
        elif isinstance(v, C.Array):
            if issubclass(v._type_, C.Structure):
                out[name] = [struct_to_dict(e) for e in v]
            else:
                out[name] = list(v)
        elif isinstance(v, (int, float, bytes)) or v is None:
            out[name] = v
        else:
            out[name] = v
    return out


def as_f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def ptr_f64(a):
    return a.ctypes.data_as(c_double_p) if a is not None else None


def ptr_i64(a):
    return a.ctypes.data_as(c_int64_p)
