"""Full-solve known-answer tests, transcribed from the reference's
ortools/pdlp/primal_dual_hybrid_gradient_test.cc (line ranges cited per test).

Each test runs against the CPU oracle (pins the oracle, no GPU needed) and,
with ``-m gpu``, against the CUDA product through its C ABI.
"""
import ctypes
import itertools
import math

import numpy as np
import pytest

import fixtures as fx
from ortools_b200 import pdlp

INF = float("inf")
TR = pdlp.TerminationReason
PT = pdlp.PointType
RC = pdlp.RestartChoice
P = pdlp.PrimalDualHybridGradientParams


def create_solver_params(iteration_limit, eps_optimal_absolute, enable_scaling, num_threads, use_iteration_limit,
                         use_malitsky_pock, use_diag_tr):
    # primal_dual_hybrid_gradient_test.cc:68-105
    p = P()
    if not enable_scaling:
        p.l2_norm_rescaling = False
        p.l_inf_ruiz_iterations = 0
    if use_malitsky_pock:
        p.linesearch_rule = P.MALITSKY_POCK_LINESEARCH_RULE
    p.termination_criteria.simple_optimality_criteria.eps_optimal_relative = 0.0
    if use_iteration_limit:
        p.termination_criteria.iteration_limit = iteration_limit
        p.termination_criteria.simple_optimality_criteria.eps_optimal_absolute = 0.0
    else:
        p.termination_criteria.simple_optimality_criteria.eps_optimal_absolute = eps_optimal_absolute
    if use_diag_tr:
        p.use_diagonal_qp_trust_region_solver = True
        p.diagonal_qp_trust_region_solver_tolerance = 1.0e-8
    p.num_threads = num_threads
    p.termination_criteria.kkt_matrix_pass_limit = 1000000.0
    return p


def verify_termination(params, out, use_iteration_limit):
    # :118-148
    log = out.solve_log
    if use_iteration_limit:
        assert log.termination_reason in (TR.TERMINATION_REASON_ITERATION_LIMIT, TR.TERMINATION_REASON_NUMERICAL_ERROR,
                                          TR.TERMINATION_REASON_OPTIMAL)
        if log.termination_reason == TR.TERMINATION_REASON_ITERATION_LIMIT:
            assert log.iteration_count == params.termination_criteria.iteration_limit
        else:
            assert log.iteration_count <= params.termination_criteria.iteration_limit
    else:
        assert log.termination_reason == TR.TERMINATION_REASON_OPTIMAL
        assert log.iteration_count <= params.termination_criteria.iteration_limit


def convergence_info(log, point_type=None):
    t = log.solution_type if point_type is None else point_type
    for ci in log.solution_stats.convergence_information:
        if ci.candidate_type == t:
            return ci
    return None


def verify_objectives(out, value, tol):
    ci = convergence_info(out.solve_log)
    assert ci is not None
    assert ci.primal_objective == pytest.approx(value, abs=tol)
    assert ci.dual_objective == pytest.approx(value, abs=tol)


# {scaling} x {1,4 threads} x {iteration-limit vs eps} x {adaptive, M-P}   (:239-252)
LP_GRID = list(itertools.product([False, True], [1, 4], [False, True], [False, True]))
LP_IDS = ["%s_%dT_%s_%s" % ("Scaling" if s else "NoScaling", t, "IterLimit" if i else "Eps", "MP" if mp else "Adaptive")
          for s, t, i, mp in LP_GRID]


def lp_params(cfg, iteration_limit, eps):
    s, t, i, mp = cfg
    return create_solver_params(iteration_limit, eps, s, t, i, mp, False)


@pytest.mark.parametrize("cfg", LP_GRID, ids=LP_IDS)
def test_lp_unbounded_variables(backend, cfg):
    # :266-285
    params = lp_params(cfg, 980, 1.0e-7)
    params.major_iteration_frequency = 100
    out = backend.primal_dual_hybrid_gradient(fx.test_lp(), params)
    verify_termination(params, out, cfg[2])
    verify_objectives(out, -34.0, 1.0e-6)
    assert out.primal_solution == pytest.approx([-1, 8, 1, 2.5], abs=1e-4)
    assert out.dual_solution == pytest.approx([-2, 0, 2.375, 2.0 / 3], abs=1e-4)
    assert out.solve_log.original_problem_stats.num_variables == 4
    assert out.solve_log.preprocessed_problem_stats.num_constraints <= 4


@pytest.mark.parametrize("cfg", LP_GRID, ids=LP_IDS)
def test_lp_tiny(backend, cfg):
    # :287-307
    params = lp_params(cfg, 300, 1.0e-5)
    params.major_iteration_frequency = 60
    out = backend.primal_dual_hybrid_gradient(fx.tiny_lp(), params)
    verify_termination(params, out, cfg[2])
    verify_objectives(out, -1.0, 1.0e-4)
    assert out.primal_solution == pytest.approx([1, 0, 6, 2], abs=1e-4)
    assert out.dual_solution == pytest.approx([0.5, 4.0, 0.0], abs=1e-4)
    assert out.reduced_costs == pytest.approx([0.0, 1.5, -3.5, 0.0], abs=1e-4)


@pytest.mark.parametrize("cfg", LP_GRID, ids=LP_IDS)
def test_lp_correlation_clustering_one(backend, cfg):
    # :309-338
    params = lp_params(cfg, 9, 1.0e-10)
    params.major_iteration_frequency = 2
    out = backend.primal_dual_hybrid_gradient(fx.correlation_clustering_lp(), params)
    verify_termination(params, out, cfg[2])
    verify_objectives(out, 1.0, 1.0e-14)
    assert out.primal_solution == pytest.approx([1, 1, 0, 1, 0, 0], abs=1e-14)
    y = out.dual_solution
    assert y.size == 3 and all(v >= 0 for v in y) and y[0] + y[1] >= 1 - 1e-14
    assert convergence_info(out.solve_log).corrected_dual_objective == pytest.approx(1.0, abs=1e-14)


@pytest.mark.parametrize("cfg", LP_GRID, ids=LP_IDS)
def test_lp_correlation_clustering_star(backend, cfg):
    # :340-359
    params = lp_params(cfg, 45, 1.0e-6)
    params.major_iteration_frequency = 5
    out = backend.primal_dual_hybrid_gradient(fx.correlation_clustering_star_lp(), params)
    verify_termination(params, out, cfg[2])
    verify_objectives(out, 1.5, 1.0e-6)
    assert out.primal_solution == pytest.approx([0.5, 0.5, 0.5, 0, 0, 0], abs=1e-6)
    assert out.dual_solution == pytest.approx([0.5, 0.5, 0.5], abs=1e-6)


@pytest.mark.parametrize("cfg", LP_GRID, ids=LP_IDS)
def test_lp_inactive_two_sided_constraint(backend, cfg):
    # :365-383
    params = lp_params(cfg, 500, 1.0e-8)
    params.major_iteration_frequency = 60
    qp = fx.test_lp()
    qp.constraint_lower_bounds[1] = -10
    out = backend.primal_dual_hybrid_gradient(qp, params)
    verify_termination(params, out, cfg[2])
    verify_objectives(out, -34.0, 1.0e-6)
    assert out.primal_solution == pytest.approx([-1, 8, 1, 2.5], abs=1e-7)
    assert out.dual_solution == pytest.approx([-2.0, 0.0, 2.375, 2.0 / 3], abs=1e-7)


@pytest.mark.parametrize("cfg", LP_GRID, ids=LP_IDS)
def test_lp_infeasible_primal(backend, cfg):
    # :385-418
    params = lp_params(cfg, 2000, 1.0e-6)
    params.major_iteration_frequency = 5
    params.termination_criteria.eps_primal_infeasible = 1.0e-6
    out = backend.primal_dual_hybrid_gradient(fx.small_primal_infeasible_lp(), params)
    assert out.solve_log.termination_reason == TR.TERMINATION_REASON_PRIMAL_INFEASIBLE
    dual = out.dual_solution
    assert dual[0] / dual[1] == pytest.approx(1, abs=1e-6) and dual[1] < 0.0
    assert out.reduced_costs == pytest.approx([dual[1] - dual[0], dual[0] - dual[1]], abs=1e-6)
    assert out.solve_log.iteration_count <= 2000


@pytest.mark.parametrize("cfg", LP_GRID, ids=LP_IDS)
def test_lp_infeasible_dual(backend, cfg):
    # :420-437
    params = lp_params(cfg, 500, 1.0e-6)
    params.major_iteration_frequency = 5
    out = backend.primal_dual_hybrid_gradient(fx.small_dual_infeasible_lp(), params)
    assert out.solve_log.termination_reason == TR.TERMINATION_REASON_DUAL_INFEASIBLE
    x = out.primal_solution
    assert x[0] / x[1] == pytest.approx(1, abs=1e-6) and x[1] > 0.0
    assert out.solve_log.iteration_count <= 500


@pytest.mark.parametrize("cfg", LP_GRID, ids=LP_IDS)
def test_lp_infeasible_primal_with_reduced_costs(backend, cfg):
    # :439-467
    lp = fx._qp([[1.0]], [2], [INF], [0], [1], [1.0])
    params = lp_params(cfg, 100, 1.0e-6)
    params.major_iteration_frequency = 5
    params.termination_criteria.eps_primal_infeasible = 1.0e-6
    out = backend.primal_dual_hybrid_gradient(lp, params)
    assert out.solve_log.termination_reason == TR.TERMINATION_REASON_PRIMAL_INFEASIBLE
    assert out.dual_solution[0] > 0.0 and out.reduced_costs[0] == -out.dual_solution[0]
    assert out.solve_log.iteration_count < 100


@pytest.mark.parametrize("cfg", LP_GRID, ids=LP_IDS)
def test_lp_infeasible_primal_dual(backend, cfg):
    # :469-487
    params = lp_params(cfg, 600, 1.0e-6)
    params.restart_strategy = P.NO_RESTARTS
    params.major_iteration_frequency = 5
    out = backend.primal_dual_hybrid_gradient(fx.small_primal_dual_infeasible_lp(), params)
    assert out.solve_log.termination_reason in (TR.TERMINATION_REASON_DUAL_INFEASIBLE, TR.TERMINATION_REASON_PRIMAL_INFEASIBLE)
    assert out.solve_log.iteration_count <= 600


@pytest.mark.parametrize("cfg", LP_GRID, ids=LP_IDS)
def test_lp_without_constraints_or_variables(backend, cfg):
    # :573-650
    params = lp_params(cfg, 2, 1.0e-6)
    qp = fx._qp(np.zeros((0, 3)), [], [], [-1, -INF, -2], [INF, 4, 10], [1, -1, 2])
    out = backend.primal_dual_hybrid_gradient(qp, params)
    verify_termination(params, out, cfg[2])
    verify_objectives(out, -9.0, 1.0e-6)
    assert out.primal_solution == pytest.approx([-1, 4, -2], abs=1e-6) and out.dual_solution.size == 0
    assert out.solve_log.preprocessed_problem_stats.num_constraints == 0
    qp = pdlp.QuadraticProgram(0, 3)
    qp.constraint_lower_bounds = np.array([-1, -INF, -2.0])
    qp.constraint_upper_bounds = np.array([INF, 4, 10.0])
    out = backend.primal_dual_hybrid_gradient(qp, params)
    verify_termination(params, out, cfg[2])
    verify_objectives(out, 0.0, 1.0e-6)
    assert out.primal_solution.size == 0 and out.dual_solution == pytest.approx([0, 0, 0], abs=1e-6)
    # :616-631 only fixed variable
    params0 = lp_params(cfg, 2, 0.0)
    qp = fx._qp(np.zeros((0, 1)), [], [], [1], [1], [1])
    out = backend.primal_dual_hybrid_gradient(qp, params0)
    assert out.solve_log.termination_reason == TR.TERMINATION_REASON_OPTIMAL
    assert out.primal_solution == pytest.approx([1.0], abs=1e-6) and out.solve_log.iteration_count <= 2
    # :633-650 infeasible LP without variables
    qp = pdlp.QuadraticProgram(0, 1)
    qp.constraint_lower_bounds = np.array([-1.0])
    qp.constraint_upper_bounds = np.array([-1.0])
    out = backend.primal_dual_hybrid_gradient(qp, params)
    assert out.solve_log.termination_reason == TR.TERMINATION_REASON_PRIMAL_INFEASIBLE
    assert out.dual_solution[0] < 0.0


QP_GRID = list(itertools.product([False, True], [1, 4], [False, True], [False, True], [False, True]))
QP_IDS = ["%s_%dT_%s_%s_%s" % ("Scaling" if s else "NoScaling", t, "IterLimit" if i else "Eps", "MP" if mp else "Adaptive",
                               "TRDiag" if d else "TRNoDiag") for s, t, i, mp, d in QP_GRID]


@pytest.mark.parametrize("cfg", QP_GRID, ids=QP_IDS)
def test_diagonal_qps(backend, cfg):
    # :489-541
    s, t, i, mp, d = cfg
    for qp_fn, limit, major, obj, x, y, rc in (
            (fx.test_diagonal_qp1, 96, 12, 6.0, [1.0, 0.0], [-1.0], [4.0, 0.0]),
            (fx.test_diagonal_qp2, 240, 12, -5.0, [3.0, 1.0], [0.0], [0.0, 0.0]),
            (fx.test_diagonal_qp3, 300, 15, 2.0, [2.0, 0.0, 1.0], [-1.0, 1.0], [0, 0, 0])):
        params = create_solver_params(limit, 1.0e-6, s, t, i, mp, d)
        params.major_iteration_frequency = major
        out = backend.primal_dual_hybrid_gradient(qp_fn(), params)
        verify_termination(params, out, i)
        verify_objectives(out, obj, 1.0e-6)
        assert out.primal_solution == pytest.approx(x, abs=1e-6)
        assert out.dual_solution == pytest.approx(y, abs=1e-6)
        assert out.reduced_costs == pytest.approx(rc, abs=1e-6)


@pytest.mark.parametrize("cfg", QP_GRID, ids=QP_IDS)
def test_qp_warm_start(backend, cfg):
    # :545-571
    s, t, i, mp, d = cfg
    params = create_solver_params(35, 1.0e-6, s, t, i, mp, d)
    params.major_iteration_frequency = 5
    params.primal_weight_update_smoothing = 0.0
    init = pdlp.PrimalAndDualSolution([0.999, 0.001], [-0.999])
    out = backend.primal_dual_hybrid_gradient(fx.test_diagonal_qp1(), params, init)
    verify_termination(params, out, i)
    verify_objectives(out, 6.0, 1.0e-6)
    assert out.primal_solution == pytest.approx([1.0, 0.0], abs=1e-6)
    assert out.dual_solution == pytest.approx([-1.0], abs=1e-6)
    assert out.reduced_costs == pytest.approx([4.0, 0.0], abs=1e-6)


def params_with_no_limits():
    # :652-664
    p = P()
    p.termination_criteria.simple_optimality_criteria.eps_optimal_relative = 0.0
    p.termination_criteria.simple_optimality_criteria.eps_optimal_absolute = 0.0
    p.record_iteration_stats = True
    return p


@pytest.mark.parametrize("strategy,expected", [(P.NO_RESTARTS, RC.RESTART_CHOICE_WEIGHTED_AVERAGE_RESET),
                                               (P.EVERY_MAJOR_ITERATION, RC.RESTART_CHOICE_RESTART_TO_AVERAGE)])
def test_restart_bookkeeping(backend, strategy, expected):
    # :666-733
    major, limit = 17, 100
    p = params_with_no_limits()
    p.termination_criteria.iteration_limit = limit
    p.termination_check_frequency = 1
    p.major_iteration_frequency = major
    p.restart_strategy = strategy
    out = backend.primal_dual_hybrid_gradient(fx.test_lp(), p)
    assert out.solve_log.iteration_count == limit
    assert len(out.solve_log.iteration_stats) == limit + 1
    for i, st in enumerate(out.solve_log.iteration_stats):
        assert st.iteration_number == i
        if i == 0 or i % major != 0:
            assert st.restart_used == RC.RESTART_CHOICE_NO_RESTART, i
        else:
            assert st.restart_used == expected, i


def test_solve_log_name_and_params(backend):
    # :735-760
    p = P()
    p.termination_criteria.iteration_limit = 1
    qp = fx.test_lp()
    qp.problem_name = "Test LP"
    assert backend.primal_dual_hybrid_gradient(qp, p).solve_log.instance_name == "Test LP"
    out = backend.primal_dual_hybrid_gradient(fx.test_lp(), p)
    assert out.solve_log.instance_name is None
    assert out.solve_log.params.termination_criteria.iteration_limit == 1


def test_adaptive_distance_based_restarts(backend):
    # :762-835
    p = P()
    p.major_iteration_frequency = 16
    p.termination_criteria.iteration_limit = 128
    p.restart_strategy = P.ADAPTIVE_DISTANCE_BASED
    p.necessary_reduction_for_restart = 0.99
    out = backend.primal_dual_hybrid_gradient(fx.test_lp(), p)
    assert out.solve_log.termination_reason == TR.TERMINATION_REASON_OPTIMAL
    assert out.primal_solution == pytest.approx([-1, 8, 1, 2.5], abs=1e-4)
    assert out.dual_solution == pytest.approx([-2, 0, 2.375, 2.0 / 3], abs=1e-4)
    ci = convergence_info(out.solve_log)
    assert ci.primal_objective == pytest.approx(-34.0, abs=1e-4) and ci.dual_objective == pytest.approx(-34.0, abs=1e-4)
    assert backend.primal_dual_hybrid_gradient(fx.test_diagonal_qp1(), p).solve_log.termination_reason == TR.TERMINATION_REASON_OPTIMAL
    major, limit = 13, 100
    p = params_with_no_limits()
    p.termination_criteria.iteration_limit = limit
    p.termination_check_frequency = 1
    p.major_iteration_frequency = major
    p.restart_strategy = P.ADAPTIVE_DISTANCE_BASED
    p.necessary_reduction_for_restart = 0.75
    out = backend.primal_dual_hybrid_gradient(fx.test_lp(), p)
    assert out.solve_log.iteration_count == limit and len(out.solve_log.iteration_stats) == limit + 1
    for i, st in enumerate(out.solve_log.iteration_stats):
        assert st.iteration_number == i
        if i == 0:
            assert st.restart_used == RC.RESTART_CHOICE_NO_RESTART
        elif i == major:
            assert st.restart_used in (RC.RESTART_CHOICE_RESTART_TO_AVERAGE, RC.RESTART_CHOICE_WEIGHTED_AVERAGE_RESET)
        elif i % major != 0:
            assert st.restart_used == RC.RESTART_CHOICE_NO_RESTART


def test_primal_weight_frozen(backend):
    # :837-856
    p = params_with_no_limits()
    p.termination_criteria.iteration_limit = 100
    p.major_iteration_frequency = 17
    p.restart_strategy = P.EVERY_MAJOR_ITERATION
    p.initial_primal_weight = 1.5
    p.primal_weight_update_smoothing = 0.0
    out = backend.primal_dual_hybrid_gradient(fx.test_lp(), p)
    assert out.solve_log.iteration_stats and all(st.primal_weight == 1.5 for st in out.solve_log.iteration_stats)


def test_constant_step_size_and_scaling(backend):
    # :858-898
    p = params_with_no_limits()
    p.termination_criteria.iteration_limit = 100
    p.termination_check_frequency = 1
    p.linesearch_rule = P.CONSTANT_STEP_SIZE_RULE
    out = backend.primal_dual_hybrid_gradient(fx.test_lp(), p)
    steps = [st.step_size for st in out.solve_log.iteration_stats]
    assert steps and all(s == steps[0] for s in steps)
    p.termination_criteria.iteration_limit = 1
    unscaled = backend.primal_dual_hybrid_gradient(fx.test_lp(), p).solve_log.iteration_stats[0].step_size
    p.initial_step_size_scaling = 0.5
    scaled = backend.primal_dual_hybrid_gradient(fx.test_lp(), p).solve_log.iteration_stats[0].step_size
    assert scaled == unscaled * 0.5


def test_kkt_matrix_pass_termination(backend):
    # :901-916
    p = params_with_no_limits()
    p.termination_criteria.kkt_matrix_pass_limit = 13
    p.linesearch_rule = P.CONSTANT_STEP_SIZE_RULE
    out = backend.primal_dual_hybrid_gradient(fx.test_lp(), p)
    assert out.solve_log.termination_reason == TR.TERMINATION_REASON_KKT_MATRIX_PASS_LIMIT
    assert out.solve_log.solution_stats.cumulative_kkt_matrix_passes == 13


def test_stats_recorded_each_iteration(backend):
    # :918-1017
    p = params_with_no_limits()
    p.termination_check_frequency = 1
    p.termination_criteria.iteration_limit = 100
    p.major_iteration_frequency = 17
    p.random_projection_seeds = [1, 2]
    out = backend.primal_dual_hybrid_gradient(fx.test_lp(), p)
    assert len(out.solve_log.iteration_stats) == 101
    for st in out.solve_log.iteration_stats:
        types_c = {c.candidate_type for c in st.convergence_information}
        types_i = {c.candidate_type for c in st.infeasibility_information}
        types_m = {c.point_type for c in st.point_metadata}
        assert PT.POINT_TYPE_CURRENT_ITERATE in types_c & types_i & types_m
        if st.iteration_number > 0:
            assert PT.POINT_TYPE_AVERAGE_ITERATE in types_c & types_i & types_m
            assert PT.POINT_TYPE_ITERATE_DIFFERENCE in types_i & types_m
        for md in st.point_metadata:
            assert len(md.random_primal_projections) == 2 and len(md.random_dual_projections) == 2
    p.record_iteration_stats = False
    assert len(backend.primal_dual_hybrid_gradient(fx.test_lp(), p).solve_log.iteration_stats) == 0
    p.record_iteration_stats = True
    p.random_projection_seeds = []
    out = backend.primal_dual_hybrid_gradient(fx.test_lp(), p)
    for st in out.solve_log.iteration_stats:
        assert st.point_metadata
        assert all(len(md.random_primal_projections) == 0 for md in st.point_metadata)


def test_projects_initial_point(backend):
    # :1019-1047
    p = params_with_no_limits()
    p.termination_criteria.iteration_limit = 5
    out = backend.primal_dual_hybrid_gradient(fx.small_initialization_lp(), p)
    assert len(out.solve_log.iteration_stats) == 6 and out.primal_solution[0] > 0.0
    out = backend.primal_dual_hybrid_gradient(fx.small_initialization_lp(), p, pdlp.PrimalAndDualSolution([1.0, 0.0], [-1.0, -1.0]))
    assert len(out.solve_log.iteration_stats) == 6 and out.dual_solution[0] <= 0.0


def _reason(backend, qp, params=None, init=None):
    return backend.primal_dual_hybrid_gradient(qp, params or P(), init).solve_log.termination_reason


def test_input_validation(backend):
    # :1049-1408
    NAN = float("nan")
    assert _reason(backend, fx.small_invalid_problem_lp()) == TR.TERMINATION_REASON_INVALID_PROBLEM
    qp = fx.test_diagonal_qp1(); qp.objective_matrix[0] = -1.0
    assert _reason(backend, qp) == TR.TERMINATION_REASON_INVALID_PROBLEM
    qp = fx.tiny_lp(); qp.objective_vector = np.zeros(0)
    assert _reason(backend, qp) == TR.TERMINATION_REASON_INVALID_PROBLEM
    p = P(); p.num_threads = 0
    assert _reason(backend, fx.tiny_lp(), p) == TR.TERMINATION_REASON_INVALID_PARAMETER

    def with_k(i, j, v):
        q = fx.test_lp(); K = q.constraint_matrix.toarray(); K[i, j] = v
        import scipy.sparse as sp
        q.constraint_matrix = sp.csc_matrix(K); return q
    for q in (with_k(0, 0, NAN), with_k(0, 0, 1e60), with_k(0, 1, 1e-60), with_k(2, 0, 1e-60)):
        assert _reason(backend, q) == TR.TERMINATION_REASON_INVALID_PROBLEM

    def with_attr(name, idx, v, base=fx.test_lp):
        q = base(); getattr(q, name)[idx] = v; return q
    for q in (with_attr("constraint_upper_bounds", 1, NAN), with_attr("constraint_upper_bounds", 1, 1e60),
              with_attr("constraint_lower_bounds", 2, -1e60), with_attr("variable_lower_bounds", 3, NAN),
              with_attr("variable_lower_bounds", 3, -1e60), with_attr("objective_vector", 3, NAN),
              with_attr("objective_vector", 3, -1e60),
              with_attr("objective_matrix", 0, NAN, fx.test_diagonal_qp1), with_attr("objective_matrix", 0, 1e60, fx.test_diagonal_qp1)):
        assert _reason(backend, q) == TR.TERMINATION_REASON_INVALID_PROBLEM
    # excessively small values are fine without presolve (:1152-1171, 1201-1211, 1259-1267)
    for q in (with_attr("constraint_upper_bounds", 1, 1e-60), with_attr("constraint_lower_bounds", 2, 1e-60),
              with_attr("objective_vector", 3, 1e-60)):
        assert _reason(backend, q) == TR.TERMINATION_REASON_OPTIMAL
    q = fx.test_lp(); q.variable_lower_bounds[1] = 0.0; q.variable_upper_bounds[1] = 1e-60
    assert _reason(backend, q) == TR.TERMINATION_REASON_OPTIMAL
    q = fx.test_lp(); q.objective_offset = NAN
    assert _reason(backend, q) == TR.TERMINATION_REASON_INVALID_PROBLEM
    q = fx.test_lp(); q.objective_offset = -1e60
    assert _reason(backend, q) == TR.TERMINATION_REASON_INVALID_PROBLEM
    q = fx.test_lp(); q.objective_scaling_factor = 0.0
    assert _reason(backend, q) == TR.TERMINATION_REASON_INVALID_PROBLEM
    ones = [1.0, 1.0, 1.0, 1.0]
    for x0, y0 in (([1.0, NAN, 1.0, 1.0], ones), ([1.0, 1e100, 1.0, 1.0], ones), ([1.0, 1.0, 1.0], ones),
                   (ones, [1.0, NAN, 1.0, 1.0]), (ones, [1.0, 1e100, 1.0, 1.0]), (ones, [1.0, 1.0, 1.0])):
        assert _reason(backend, fx.test_lp(), None, pdlp.PrimalAndDualSolution(x0, y0)) == TR.TERMINATION_REASON_INVALID_INITIAL_SOLUTION
    p = P(); p.use_feasibility_polishing = True
    assert _reason(backend, fx.test_diagonal_qp1(), p) == TR.TERMINATION_REASON_INVALID_PARAMETER


def test_checks_termination_at_correct_frequency(backend):
    # :1410-1437
    p = params_with_no_limits()
    p.termination_criteria.iteration_limit = 16
    p.termination_check_frequency = 2
    p.major_iteration_frequency = 5
    out = backend.primal_dual_hybrid_gradient(fx.test_lp(), p)
    assert out.solve_log.iteration_count == 16
    checked = [st.iteration_number for st in out.solve_log.iteration_stats if st.convergence_information]
    assert checked == [0, 2, 4, 5, 7, 9, 10, 12, 14, 15, 16]


def test_calls_callback(backend):
    # :1439-1462
    p = params_with_no_limits()
    p.termination_criteria.iteration_limit = 16
    p.termination_check_frequency = 5
    p.major_iteration_frequency = 5
    calls = []
    out = backend.primal_dual_hybrid_gradient(fx.test_lp(), p, iteration_stats_callback=lambda info: calls.append(info))
    assert out.solve_log.iteration_count == 16
    assert len(calls) == 6
    assert [c.iteration_stats.iteration_number for c in calls] == [0, 5, 10, 15, 16, 16]
    assert calls[-1].iteration_type == pdlp.IterationType.NORMAL_TERMINATION
    assert calls[0].bound_norms.l2_norm_constraint_bounds == pytest.approx(math.sqrt(210.0))


def test_warm_start_at_optimum(backend):
    # :1472-1504
    p = params_with_no_limits()
    p.termination_criteria.simple_optimality_criteria.eps_optimal_absolute = 1.0e-10
    sol = pdlp.PrimalAndDualSolution([1.0, 0.0, 6.0, 2.0], [0.5, 4.0, 0.0])
    out = backend.primal_dual_hybrid_gradient(fx.tiny_lp(), p, sol)
    assert out.primal_solution == pytest.approx([1, 0, 6, 2], abs=1e-10)
    assert out.dual_solution == pytest.approx([0.5, 4.0, 0.0], abs=1e-10)
    assert out.solve_log.iteration_count <= 0 and out.solve_log.termination_reason == TR.TERMINATION_REASON_OPTIMAL
    p = params_with_no_limits()
    p.l2_norm_rescaling = False
    p.l_inf_ruiz_iterations = 0
    out = backend.primal_dual_hybrid_gradient(fx.tiny_lp(), p, sol)
    assert list(out.primal_solution) == [1, 0, 6, 2] and list(out.dual_solution) == [0.5, 4.0, 0.0]
    assert out.solve_log.iteration_count <= 1 and out.solve_log.termination_reason == TR.TERMINATION_REASON_OPTIMAL


def test_empty_qp(backend):
    # :1506-1517
    out = backend.primal_dual_hybrid_gradient(pdlp.QuadraticProgram(0, 0), params_with_no_limits())
    assert out.primal_solution.size == 0 and out.dual_solution.size == 0
    assert out.solve_log.iteration_count == 0 and out.solve_log.termination_reason == TR.TERMINATION_REASON_OPTIMAL


def test_interrupts(backend):
    # :1519-1573
    p = P()
    p.termination_criteria.simple_optimality_criteria.eps_optimal_absolute = 0.0
    p.termination_criteria.simple_optimality_criteria.eps_optimal_relative = 0.0
    flag = ctypes.c_int32(1)
    assert backend.primal_dual_hybrid_gradient(fx.test_lp(), p, interrupt_solve=flag).solve_log.termination_reason == TR.TERMINATION_REASON_INTERRUPTED_BY_USER
    flag = ctypes.c_int32(0)

    def cb(info):
        if info.iteration_stats.iteration_number >= 10:
            flag.value = 1
    out = backend.primal_dual_hybrid_gradient(fx.test_lp(), p, interrupt_solve=flag, iteration_stats_callback=cb)
    assert out.solve_log.termination_reason == TR.TERMINATION_REASON_INTERRUPTED_BY_USER and out.solve_log.iteration_count >= 10
    p.termination_criteria.kkt_matrix_pass_limit = 1
    flag = ctypes.c_int32(0)
    assert backend.primal_dual_hybrid_gradient(fx.test_lp(), p, interrupt_solve=flag).solve_log.termination_reason == TR.TERMINATION_REASON_KKT_MATRIX_PASS_LIMIT


def test_huge_threads_and_shards(backend):
    # :1575-1598
    for field in ("num_threads", "num_shards"):
        p = params_with_no_limits()
        p.termination_criteria.iteration_limit = 10
        setattr(p, field, 1000000000)
        assert backend.primal_dual_hybrid_gradient(fx.test_lp(), p).solve_log.termination_reason == TR.TERMINATION_REASON_ITERATION_LIMIT


def test_detailed_termination_criteria(backend):
    # :1600-1640
    p = create_solver_params(300, 1.0e-5, True, 4, False, False, False)
    p.major_iteration_frequency = 60
    p.termination_criteria.ClearField("simple_optimality_criteria")
    object.__setattr__(p.termination_criteria, "_oneof", None)
    d = p.termination_criteria.detailed_optimality_criteria
    d.eps_optimal_primal_residual_absolute = 1.0e-5
    d.eps_optimal_primal_residual_relative = 0.0
    d.eps_optimal_dual_residual_absolute = 1.0e-5
    d.eps_optimal_dual_residual_relative = 0.0
    d.eps_optimal_objective_gap_absolute = 1.0e-5
    d.eps_optimal_objective_gap_relative = 0.0
    out = backend.primal_dual_hybrid_gradient(fx.tiny_lp(), p)
    assert out.solve_log.termination_reason == TR.TERMINATION_REASON_OPTIMAL
    verify_objectives(out, -1.0, 1.0e-4)
    assert out.primal_solution == pytest.approx([1, 0, 6, 2], abs=1e-4)


@pytest.mark.parametrize("verbosity", [0, 1, 2, 3, 4])
def test_verbosity_levels(backend, verbosity):
    # :2125-2153: logging goes to the message callback; level 0 is silent
    p = P()
    p.verbosity_level = verbosity
    p.termination_criteria.iteration_limit = 200
    lines = []
    out = backend.primal_dual_hybrid_gradient(fx.test_lp(), p, message_callback=lines.append)
    assert out.solve_log.termination_reason == TR.TERMINATION_REASON_OPTIMAL
    assert (len(lines) == 0) == (verbosity == 0)
