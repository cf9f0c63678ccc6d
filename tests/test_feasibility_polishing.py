"""Feasibility polishing (primal_dual_hybrid_gradient.cc:2676-3015).
Transcribed from ``primal_dual_hybrid_gradient_test.cc:1400-2120`` (every FeasibilityPolishing*Test):
the two small LPs need ~2500 iterations without polishing and solve at the first polishing
attempt (iteration 100) with it. Run against the CPU restatement (not gpu) and the CUDA path."""
import ctypes

import numpy as np
import pytest
import scipy.sparse as sp

import fixtures as fx
from ortools_b200 import pdlp

TR = pdlp.TerminationReason
INF = float("inf")


def polishing_params():  # FeasibilityPolishingTest, :1669-1690
    p = pdlp.PrimalDualHybridGradientParams()
    p.linesearch_rule = p.CONSTANT_STEP_SIZE_RULE
    d = p.termination_criteria.detailed_optimality_criteria
    d.eps_optimal_primal_residual_absolute = 1.0e-6
    d.eps_optimal_primal_residual_relative = 1.0e-6
    d.eps_optimal_dual_residual_absolute = 1.0e-6
    d.eps_optimal_dual_residual_relative = 1.0e-6
    d.eps_optimal_objective_gap_absolute = 1.0e-2
    d.eps_optimal_objective_gap_relative = 1.0e-2
    p.termination_criteria.iteration_limit = 500
    p.handle_some_primal_gradients_on_finite_bounds_as_residuals = False
    p.primal_weight_update_smoothing = 0.0
    p.use_feasibility_polishing = True
    return p


def primal_lp():  # min x_1 + 1.001 x_2 s.t. x_1 + x_2 = 1, x >= -1 (:1696-1712)
    qp = pdlp.QuadraticProgram(2, 1)
    qp.objective_vector = np.array([1.0, 1.001])
    qp.variable_lower_bounds = np.array([-1.0, -1.0])
    qp.variable_upper_bounds = np.array([INF, INF])
    qp.constraint_lower_bounds = np.array([1.0])
    qp.constraint_upper_bounds = np.array([1.0])
    qp.constraint_matrix = sp.csc_matrix(np.array([[1.0, 1.0]]))
    return qp


def dual_lp():  # min -y s.t. y <= 1, y <= 1.001, y free (:1719-1735)
    qp = pdlp.QuadraticProgram(1, 2)
    qp.objective_vector = np.array([-1.0])
    qp.variable_lower_bounds = np.array([-INF])
    qp.variable_upper_bounds = np.array([INF])
    qp.constraint_lower_bounds = np.array([-INF, -INF])
    qp.constraint_upper_bounds = np.array([1.0, 1.001])
    qp.constraint_matrix = sp.csc_matrix(np.array([[1.0], [1.0]]))
    return qp


def test_polishing_is_rejected_for_a_qp(backend):  # :1400-1408
    p = pdlp.PrimalDualHybridGradientParams()
    p.use_feasibility_polishing = True
    p.handle_some_primal_gradients_on_finite_bounds_as_residuals = False
    out = backend.primal_dual_hybrid_gradient(fx.test_diagonal_qp1(), p)
    assert out.solve_log.termination_reason == TR.TERMINATION_REASON_INVALID_PARAMETER


@pytest.mark.parametrize("make_lp", [primal_lp, dual_lp])
def test_feasibility_polishing_solves_faster(backend, make_lp):  # :1737-1763
    p = polishing_params()
    p.use_feasibility_polishing = False
    base = backend.primal_dual_hybrid_gradient(make_lp(), p)
    assert base.solve_log.termination_reason == TR.TERMINATION_REASON_ITERATION_LIMIT
    p.use_feasibility_polishing = True
    out = backend.primal_dual_hybrid_gradient(make_lp(), p)
    assert out.solve_log.termination_reason == TR.TERMINATION_REASON_OPTIMAL
    assert out.solve_log.solution_type == pdlp.PointType.POINT_TYPE_FEASIBILITY_POLISHING_SOLUTION


def test_feasibility_polishing_finds_valid_solution(backend):  # :1765-1788
    out = backend.primal_dual_hybrid_gradient(primal_lp(), polishing_params())
    assert out.solve_log.termination_reason == TR.TERMINATION_REASON_OPTIMAL
    ci = [c for c in out.solve_log.solution_stats.convergence_information if c.candidate_type == out.solve_log.solution_type][0]
    assert ci.primal_objective == pytest.approx(1.0, abs=1e-2) and ci.dual_objective == pytest.approx(1.0, abs=1e-2)
    x, y = out.primal_solution, out.dual_solution
    assert x[0] + x[1] == pytest.approx(1.0, abs=1e-6) and x[0] >= 0.0 and x[1] >= 0.0
    assert y[0] <= 1.0 and y[0] >= 1.0 - 1e-2
    np.testing.assert_allclose(out.reduced_costs, [1.0 - y[0], 1.001 - y[0]], atol=1e-12)


@pytest.mark.parametrize("limit,after_limits,expected", [
    (50, False, TR.TERMINATION_REASON_ITERATION_LIMIT),   # :1790-1800
    (50, True, TR.TERMINATION_REASON_OPTIMAL),            # :1802-1809
    (101, False, TR.TERMINATION_REASON_ITERATION_LIMIT),  # :1811-1821 (the limit stops primal polishing)
    (101, True, TR.TERMINATION_REASON_OPTIMAL),           # :1823-1830
])
def test_polishing_and_the_iteration_limit(backend, limit, after_limits, expected):
    p = polishing_params()
    p.apply_feasibility_polishing_after_limits_reached = after_limits
    p.termination_criteria.iteration_limit = limit
    out = backend.primal_dual_hybrid_gradient(primal_lp(), p)
    assert out.solve_log.termination_reason == expected


def test_polishing_stops_after_continuing_after_iteration_limit_when_not_optimal(backend):  # :1832-1851
    p = polishing_params()
    p.apply_feasibility_polishing_after_limits_reached = True
    p.termination_criteria.iteration_limit = 101
    d = p.termination_criteria.detailed_optimality_criteria
    d.eps_optimal_primal_residual_absolute = 1.0e-16
    d.eps_optimal_primal_residual_relative = 0.0
    d.eps_optimal_dual_residual_absolute = 1.0e-16
    d.eps_optimal_dual_residual_relative = 0.0
    d.eps_optimal_objective_gap_absolute = 1.0e-16
    d.eps_optimal_objective_gap_relative = 0.0
    out = backend.primal_dual_hybrid_gradient(primal_lp(), p)
    assert out.solve_log.termination_reason == TR.TERMINATION_REASON_ITERATION_LIMIT
    assert out.solve_log.iteration_count <= 124   # 100 main + at most 12 + 12 polishing iterations


@pytest.mark.parametrize("trigger,if_interrupted,expected", [
    ("iteration", False, TR.TERMINATION_REASON_INTERRUPTED_BY_USER),  # :1853-1870
    ("iteration", True, TR.TERMINATION_REASON_OPTIMAL),               # :1872-1890
    ("phase", False, TR.TERMINATION_REASON_INTERRUPTED_BY_USER),      # :1892-1910
    ("phase", True, TR.TERMINATION_REASON_OPTIMAL),                   # :1912-1930
])
def test_polishing_and_the_interrupt_flag(backend, trigger, if_interrupted, expected):
    p = polishing_params()
    p.apply_feasibility_polishing_if_solver_is_interrupted = if_interrupted
    flag = ctypes.c_int32(0)

    def cb(info):
        if trigger == "iteration" and info.iteration_stats.iteration_number >= 50:
            flag.value = 1
        if trigger == "phase" and info.iteration_type == pdlp.IterationType.PRIMAL_FEASIBILITY:
            flag.value = 1

    out = backend.primal_dual_hybrid_gradient(primal_lp(), p, interrupt_solve=flag, iteration_stats_callback=cb)
    assert out.solve_log.termination_reason == expected


def test_feasibility_polishing_details_in_log(backend):  # :1932-1968
    out = backend.primal_dual_hybrid_gradient(primal_lp(), polishing_params())
    assert out.solve_log.termination_reason == TR.TERMINATION_REASON_OPTIMAL
    primal_optimal = dual_optimal = 0
    for phase in out.solve_log.feasibility_polishing_details:
        assert phase.polishing_phase_type in (pdlp.PolishingPhaseType.POLISHING_PHASE_TYPE_PRIMAL_FEASIBILITY,
                                              pdlp.PolishingPhaseType.POLISHING_PHASE_TYPE_DUAL_FEASIBILITY)
        if phase.termination_reason == TR.TERMINATION_REASON_OPTIMAL:
            if phase.polishing_phase_type == pdlp.PolishingPhaseType.POLISHING_PHASE_TYPE_PRIMAL_FEASIBILITY:
                primal_optimal += 1
            else:
                dual_optimal += 1
        assert phase.solution_stats.iteration_number == phase.iteration_count
    assert primal_optimal >= 1 and dual_optimal >= 1
    # and through the SolveLog proto
    from ortools_b200 import pdlp_proto
    msg = pdlp_proto.solve_log_to_proto(out.solve_log)
    assert len(msg.feasibility_polishing_details) == len(out.solve_log.feasibility_polishing_details) >= 2


def _phases(out):
    P = pdlp.PolishingPhaseType
    for phase in out.solve_log.feasibility_polishing_details:
        assert phase.polishing_phase_type in (P.POLISHING_PHASE_TYPE_PRIMAL_FEASIBILITY, P.POLISHING_PHASE_TYPE_DUAL_FEASIBILITY)
        yield phase, phase.polishing_phase_type == P.POLISHING_PHASE_TYPE_PRIMAL_FEASIBILITY


def test_feasibility_polishing_uses_infinite_tolerances(backend):  # :1962-2001
    """Each phase's recorded parameters carry detailed criteria in which everything except the
    residual the phase works on is switched off with infinite tolerances (pdhg.cc:2833-2861)."""
    out = backend.primal_dual_hybrid_gradient(primal_lp(), polishing_params())
    seen = 0
    for phase, primal in _phases(out):
        c = phase.params.termination_criteria   # the POD mirror is flat: the oneof case is the field's tag
        assert c.optimality_criteria_case == 10  # has_detailed_optimality_criteria (solvers.proto: tag 10)
        assert c.eps_optimal_objective_gap_absolute == INF and c.eps_optimal_objective_gap_relative == INF
        if primal:
            assert c.eps_optimal_dual_residual_absolute == INF and c.eps_optimal_dual_residual_relative == INF
            assert c.eps_optimal_primal_residual_absolute == 1.0e-6 and c.eps_optimal_primal_residual_relative == 1.0e-6
        else:
            assert c.eps_optimal_primal_residual_absolute == INF and c.eps_optimal_primal_residual_relative == INF
            assert c.eps_optimal_dual_residual_absolute == 1.0e-6 and c.eps_optimal_dual_residual_relative == 1.0e-6
        seen += 1
    assert seen >= 2


def test_feasibility_polishing_work_stats_are_correct(backend):  # :2003-2042
    p = polishing_params()
    p.record_iteration_stats = True
    out = backend.primal_dual_hybrid_gradient(primal_lp(), p)
    assert out.solve_log.termination_reason == TR.TERMINATION_REASON_OPTIMAL
    polishing_iterations, polishing_time = 0, 0.0
    for phase, _ in _phases(out):
        assert phase.solution_stats.iteration_number == phase.iteration_count
        assert 0 <= phase.iteration_count <= phase.main_iteration_count // 4
        polishing_iterations += phase.iteration_count
        polishing_time += phase.solve_time_sec
    assert len(out.solve_log.iteration_stats) > 0
    last = out.solve_log.iteration_stats[-1]
    assert polishing_iterations >= 1 and last.iteration_number >= 1
    # iteration_count includes main and polishing iterations; the main loop's iteration_number does not
    assert last.iteration_number + polishing_iterations == out.solve_log.iteration_count
    assert polishing_time > 0.0 and last.cumulative_time_sec > 0.0
    assert polishing_time + last.cumulative_time_sec <= out.solve_log.solve_time_sec


def test_feasibility_objectives_are_zero(backend):  # :2044-2078
    p = polishing_params()
    p.record_iteration_stats = True
    out = backend.primal_dual_hybrid_gradient(primal_lp(), p)
    primal_infos = dual_infos = 0
    for phase, primal in _phases(out):
        for stats in phase.iteration_stats:
            for ci in stats.convergence_information:
                if primal:
                    primal_infos += 1
                    assert ci.primal_objective == 0.0
                else:
                    dual_infos += 1
                    assert ci.dual_objective == 0.0
    assert primal_infos >= 1 and dual_infos >= 1


def test_calls_callback_for_all_three_phases(backend):  # :2080-2100
    counts, order = {}, []

    def cb(info):
        counts[info.iteration_type] = counts.get(info.iteration_type, 0) + 1
        order.append(info.iteration_type)

    backend.primal_dual_hybrid_gradient(primal_lp(), polishing_params(), iteration_stats_callback=cb)
    T = pdlp.IterationType
    assert order[-1] == T.FEASIBILITY_POLISHING_TERMINATION
    assert sorted(counts) == sorted([T.NORMAL, T.PRIMAL_FEASIBILITY, T.DUAL_FEASIBILITY, T.FEASIBILITY_POLISHING_TERMINATION])
    assert all(v >= 1 for v in counts.values())


def test_iteration_limit_includes_feasibility_phases(backend):  # :2102-2120
    p = polishing_params()
    p.termination_criteria.iteration_limit = 300
    p.record_iteration_stats = True
    out = backend.primal_dual_hybrid_gradient(fx.tiny_lp(), p)
    assert out.solve_log.termination_reason == TR.TERMINATION_REASON_OPTIMAL
    polishing_iterations = sum(phase.iteration_count for phase, _ in _phases(out))
    assert polishing_iterations >= 12   # at least the first, failed, primal polishing attempt
    assert len(out.solve_log.iteration_stats) > 0
    assert out.solve_log.iteration_stats[-1].iteration_number + polishing_iterations == out.solve_log.iteration_count
