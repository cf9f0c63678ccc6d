"""ValidatePrimalDualHybridGradientParams and its sub-validators (solvers_proto_validation.cc:33-207)
with the known answers of ``ortools/pdlp/solvers_proto_validation_test.cc:34-706``. The C ABI has one
validator for the whole parameter message, so the cases the reference runs on a bare
TerminationCriteria / AdaptiveLinesearchParams / MalitskyPockParams are wrapped in the field of
PrimalDualHybridGradientParams that holds them. Parameters are written as the reference writes them
(text-format protos); both the CPU restatement and the product's own host code are checked (no GPU)."""
import pytest

from ortools_b200 import pdlp, pdlp_proto


@pytest.fixture(params=["oracle", "product"])
def lib(request):
    if request.param == "oracle":
        from oracle import pdlp_oracle
        return pdlp_oracle.backend()
    return pdlp.backend()


def check_invalid(lib, text, substring):
    ok, message = lib.validate_params(pdlp_proto.params_from_text(text))
    assert not ok, f'with parameters "{text}"'
    assert substring in message, f'with parameters "{text}": {message!r}'


def check_valid(lib, text):
    ok, message = lib.validate_params(pdlp_proto.params_from_text(text))
    assert ok, f'with parameters "{text}": {message!r}'


def tc(text):
    return "termination_criteria { " + text + " }"


# --- ValidateTerminationCriteria (:34-235) ---------------------------------------------------------
def test_defaults_are_valid(lib):  # :34-38, :236-240, :280-284, :333-337
    check_valid(lib, "")
    check_valid(lib, tc(""))
    check_valid(lib, "adaptive_linesearch_parameters { }")
    check_valid(lib, "malitsky_pock_parameters { }")


def test_bad_optimality_norm(lib):  # :40-46
    check_invalid(lib, tc("optimality_norm: OPTIMALITY_NORM_UNSPECIFIED"), "optimality_norm")


@pytest.mark.parametrize("field", ["eps_optimal_absolute", "eps_optimal_relative"])  # :76-88
@pytest.mark.parametrize("value", ["-1.0", "nan"])
def test_bad_deprecated_eps_optimal(lib, field, value):
    check_invalid(lib, tc(f"{field}: {value}"), field)


@pytest.mark.parametrize("field", ["eps_optimal_absolute", "eps_optimal_relative"])  # :90-106
@pytest.mark.parametrize("value", ["-1.0", "nan"])
def test_bad_simple_eps_optimal(lib, field, value):
    check_invalid(lib, tc(f"simple_optimality_criteria {{ {field}: {value} }}"), f"simple_optimality_criteria.{field}")


@pytest.mark.parametrize("field", [  # :108-136
    "eps_optimal_primal_residual_absolute", "eps_optimal_primal_residual_relative",
    "eps_optimal_dual_residual_absolute", "eps_optimal_dual_residual_relative",
    "eps_optimal_objective_gap_absolute", "eps_optimal_objective_gap_relative"])
@pytest.mark.parametrize("value", ["-1.0", "nan"])
def test_bad_detailed_eps_optimal(lib, field, value):
    check_invalid(lib, tc(f"detailed_optimality_criteria {{ {field}: {value} }}"), f"detailed_optimality_criteria.{field}")


@pytest.mark.parametrize("deprecated", ["eps_optimal_absolute", "eps_optimal_relative"])  # :138-172
@pytest.mark.parametrize("oneof", ["simple_optimality_criteria", "detailed_optimality_criteria"])
def test_deprecated_eps_with_oneof_criteria(lib, deprecated, oneof):
    check_invalid(lib, tc(f"{deprecated}: 1.0 {oneof} {{ }}"), oneof)


@pytest.mark.parametrize("field", ["eps_primal_infeasible", "eps_dual_infeasible", "time_sec_limit",  # :174-232
                                   "kkt_matrix_pass_limit"])
@pytest.mark.parametrize("value", ["-1.0", "nan"])
def test_bad_nonnegative_termination_fields(lib, field, value):
    check_invalid(lib, tc(f"{field}: {value}"), field)


def test_bad_iteration_limit(lib):  # :214-220
    check_invalid(lib, tc("iteration_limit: -1"), "iteration_limit")


# --- ValidateAdaptiveLinesearchParams (:242-278) ---------------------------------------------------
@pytest.mark.parametrize("field", ["step_size_reduction_exponent", "step_size_growth_exponent"])
@pytest.mark.parametrize("value", ["0.0", "2.0", "nan"])
def test_bad_adaptive_linesearch_exponents(lib, field, value):
    check_invalid(lib, f"adaptive_linesearch_parameters {{ {field}: {value} }}", field)


# --- ValidateMalitskyPockParams (:286-331) ---------------------------------------------------------
@pytest.mark.parametrize("value", ["0.0", "1.0", "nan", "1.0e-300"])
def test_bad_downscaling_factor(lib, value):  # :286-310
    check_invalid(lib, f"malitsky_pock_parameters {{ step_size_downscaling_factor: {value} }}", "step_size_downscaling_factor")


@pytest.mark.parametrize("value", ["0.0", "1.0", "nan"])
def test_bad_contraction_factor(lib, value):  # :312-331
    check_invalid(lib, f"malitsky_pock_parameters {{ linesearch_contraction_factor: {value} }}", "linesearch_contraction_factor")


@pytest.mark.parametrize("value", ["-1.0", "nan", "1.0e300"])
def test_bad_step_size_interpolation(lib, value):  # :333-352
    check_invalid(lib, f"malitsky_pock_parameters {{ step_size_interpolation: {value} }}", "step_size_interpolation")


# --- ValidatePrimalDualHybridGradientParams (:354-706) ---------------------------------------------
@pytest.mark.parametrize("text,substring", [
    (tc("eps_dual_infeasible: -1.0"), "eps_dual_infeasible"),                                    # :360-366
    ("num_threads: 0", "num_threads"),                                                           # :368-374
    ("verbosity_level: -1", "verbosity_level"),                                                  # :376-382
    ("log_interval_seconds: -1.0", "log_interval_seconds"),                                      # :384-399
    ("log_interval_seconds: nan", "log_interval_seconds"),
    ("major_iteration_frequency: 0", "major_iteration_frequency"),                               # :401-407
    ("termination_check_frequency: 0", "termination_check_frequency"),                           # :409-415
    ("restart_strategy: RESTART_STRATEGY_UNSPECIFIED", "restart_strategy"),                      # :417-424
    ("primal_weight_update_smoothing: 1.1", "primal_weight_update_smoothing"),                   # :426-451
    ("primal_weight_update_smoothing: -0.1", "primal_weight_update_smoothing"),
    ("primal_weight_update_smoothing: nan", "primal_weight_update_smoothing"),
    ("initial_primal_weight: -1.0", "initial_primal_weight"),                                    # :453-483
    ("initial_primal_weight: nan", "initial_primal_weight"),
    ("initial_primal_weight: 1.0e-300", "initial_primal_weight"),
    ("initial_primal_weight: 1.0e300", "initial_primal_weight"),
    ("l_inf_ruiz_iterations: -1", "l_inf_ruiz_iterations"),                                      # :485-498
    ("l_inf_ruiz_iterations: 1000", "l_inf_ruiz_iterations"),
    ("sufficient_reduction_for_restart: 1.0", "sufficient_reduction_for_restart"),               # :500-525
    ("sufficient_reduction_for_restart: 0.0", "sufficient_reduction_for_restart"),
    ("sufficient_reduction_for_restart: nan", "sufficient_reduction_for_restart"),
    ("necessary_reduction_for_restart: 1.0", "necessary_reduction_for_restart"),                 # :527-553
    ("sufficient_reduction_for_restart: 0.5 necessary_reduction_for_restart: 0.4", "necessary_reduction_for_restart"),
    ("necessary_reduction_for_restart: nan", "necessary_reduction_for_restart"),
    ("linesearch_rule: LINESEARCH_RULE_UNSPECIFIED", "linesearch_rule"),                         # :555-562
    ("adaptive_linesearch_parameters { step_size_reduction_exponent: -1.0 }", "step_size_reduction_exponent"),  # :564-571
    ("malitsky_pock_parameters { linesearch_contraction_factor: -1.0 }", "linesearch_contraction_factor"),      # :573-580
    ("initial_step_size_scaling: -1.0", "initial_step_size_scaling"),                            # :582-614
    ("initial_step_size_scaling: nan", "initial_step_size_scaling"),
    ("initial_step_size_scaling: 1.0e-300", "initial_step_size_scaling"),
    ("initial_step_size_scaling: 1.0e300", "initial_step_size_scaling"),
    ("infinite_constraint_bound_threshold: -1.0", "infinite_constraint_bound_threshold"),        # :616-634
    ("infinite_constraint_bound_threshold: nan", "infinite_constraint_bound_threshold"),
    ("diagonal_qp_trust_region_solver_tolerance: -1.0", "diagonal_qp_trust_region_solver_tolerance"),  # :636-664
    ("diagonal_qp_trust_region_solver_tolerance: nan", "diagonal_qp_trust_region_solver_tolerance"),
    ("diagonal_qp_trust_region_solver_tolerance: 2.220446049250313e-16", "diagonal_qp_trust_region_solver_tolerance"),
    ("use_feasibility_polishing: true handle_some_primal_gradients_on_finite_bounds_as_residuals: true "
     "presolve_options { use_glop: false }", "use_feasibility_polishing"),                       # :674-683
    ("use_feasibility_polishing: true handle_some_primal_gradients_on_finite_bounds_as_residuals: false "
     "presolve_options { use_glop: true }", "use_feasibility_polishing"),                        # :685-695
])
def test_bad_solver_parameters(lib, text, substring):
    check_invalid(lib, text, substring)


def test_feasibility_polishing_valid_options(lib):  # :666-673
    check_valid(lib, "use_feasibility_polishing: true handle_some_primal_gradients_on_finite_bounds_as_residuals: false "
                     "presolve_options { use_glop: false }")


def test_oracle_and_product_messages_agree():
    """The two validators report the same first error (the order the reference checks in)."""
    from oracle import pdlp_oracle
    a, b = pdlp_oracle.backend(), pdlp.backend()
    for text in ["", "num_threads: 0 verbosity_level: -1", tc("iteration_limit: -1 time_sec_limit: -1"),
                 "major_iteration_frequency: 0 termination_check_frequency: 0",
                 "malitsky_pock_parameters { step_size_interpolation: -1 linesearch_contraction_factor: 2 }",
                 "initial_primal_weight: 1e300 initial_step_size_scaling: 1e300"]:
        p = pdlp_proto.params_from_text(text)
        assert a.validate_params(p) == b.validate_params(p), text
