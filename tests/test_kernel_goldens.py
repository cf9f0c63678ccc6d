"""Known-answer tests of the per-kernel operators of the PDLP hot path.

Every case is transcribed from the reference's own tests (file:line cited per
test) and runs against BOTH sides through the same C-ABI-shaped interface:
the CPU oracle (pins the oracle; runs without a GPU) and the CUDA product
(``-m gpu``; calls libpdlp_b200.so).
"""
import math

import numpy as np
import pytest

import fixtures as fx
from ortools_b200 import pdlp

INF = float("inf")
TOL = 1e-12  # fp64 reduction-order tolerance for values that are exact in the reference


def mk(be, qp, **kw):
    return be.problem(qp, num_threads=kw.get("num_threads", 2), num_shards=kw.get("num_shards", 2))


def approx(v, tol=TOL):
    return pytest.approx(v, rel=tol, abs=tol)


# ---------------------------------------------------------------- SpMV ------
def test_transposed_matrix_vector_product_small(backend):
    # sharder_test.cc:254-261
    p = mk(backend, fx.matrix_only_qp(fx.sharder_test_matrix()), num_shards=3)
    assert p.transposed_matrix_vector_product([1, 2, 3]) == approx([6.0, -0.5, 6.0, 19.0])
    # the stored transpose: K x
    assert p.matrix_vector_product([1, 2, 3, 4]) == approx([6.0, 18.0, 19.0])


def _power_law_matrix(size, seed=48709241):
    # Same construction as sharder_test.cc:62-82 (col i has ~size/(i+1) nonzeros);
    # numpy RNG instead of absl::Uniform(std::mt19937).
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    rows, cols, vals = [], [], []
    for col in range(size):
        row = -1
        while row < size:
            row += int(rng.integers(1, col + 2))
            if row < size:
                rows.append(row); cols.append(col); vals.append(rng.uniform(1, 10))
    return sp.csc_matrix((vals, (rows, cols)), shape=(size, size))


@pytest.mark.parametrize("size", [10, 1000, 20000])
def test_large_mat_vec(backend, size):
    # sharder_test.cc:432-449: threaded product vs direct, ||diff||_2 <= 1e-8
    K = _power_law_matrix(size)
    qp = fx.matrix_only_qp(np.zeros((1, 1)))
    qp.resize_and_initialize(size, size)
    qp.constraint_matrix = K
    p = backend.problem(qp, num_threads=5, num_shards=15)
    rng = np.random.default_rng(1)
    y = rng.uniform(-1, 1, size)
    x = rng.uniform(-1, 1, size)
    assert np.linalg.norm(K.T @ y - p.transposed_matrix_vector_product(y)) <= 1e-8
    assert np.linalg.norm(K @ x - p.matrix_vector_product(x)) <= 1e-8
    # north_star per-kernel gate: 1e-12 relative (to sum |a_ij x_j|)
    absK = abs(K)
    assert np.max(np.abs(K.T @ y - p.transposed_matrix_vector_product(y)) / (absK.T @ np.abs(y) + 1e-300)) <= 1e-12
    assert np.max(np.abs(K @ x - p.matrix_vector_product(x)) / (absK @ np.abs(x) + 1e-300)) <= 1e-12


def test_empty_and_ragged_matrix(backend):
    # rows / columns without entries, and a matrix with zero rows
    K = np.array([[0, 0, 0, 0], [1, 0, 0, 2], [0, 0, 0, 0.0]])
    p = mk(backend, fx.matrix_only_qp(K))
    assert p.matrix_vector_product([1, 1, 1, 1]) == approx([0.0, 3.0, 0.0])
    assert p.transposed_matrix_vector_product([5, 2, 7]) == approx([2.0, 0.0, 0.0, 4.0])
    p0 = mk(backend, fx.lp_without_constraints())
    assert p0.matrix_vector_product([1, 2]).size == 0
    assert p0.transposed_matrix_vector_product([]) == approx([0.0, 0.0])


# ------------------------------------------------------------ vector ops ----
def test_vector_reductions(backend):
    # sharder_test.cc:320-408
    V = backend.vector_reduce
    assert V(0, [1, 2, 3], [4, 5, 6]) == approx(32.0, 1e-13)        # Dot
    assert V(1, [-1, 2, -3]) == 3                                     # LInfNorm
    assert V(2, [-1, 2, -3]) == 6                                     # L1Norm
    assert V(3, [1, 2, 3]) == approx(14.0, 1e-13)                     # SquaredNorm
    assert V(4, [1, 2, 3]) == approx(math.sqrt(14.0), 1e-13)          # Norm
    assert V(5, [1, 1, 1], [1, 2, 3]) == approx(5.0, 1e-13)           # SquaredDistance
    assert V(6, [1, 1, 1], [1, 2, 3]) == approx(math.sqrt(5.0), 1e-13)
    assert V(7, [-1, 2, -3], [4, 6, 1]) == 12                         # ScaledLInfNorm
    assert V(8, [-1, 2, -3], [4, 6, 1]) == 169                        # ScaledSquaredNorm
    assert V(9, [-1, 2, -3], [4, 6, 1]) == 13
    assert V(1, []) == 0 and V(2, []) == 0                            # empty vectors


@pytest.mark.parametrize("size", [10, 1000, 100000])
def test_large_vector_squared_norm(backend, size):
    # sharder_test.cc:451-461
    v = np.random.default_rng(size).uniform(-1, 1, size)
    assert abs(backend.vector_reduce(3, v) - float(v @ v)) <= size * 1e-14


def test_scaled_col_norms(backend):
    # sharder_test.cc:410-430
    p = mk(backend, fx.matrix_only_qp(fx.sharder_test_matrix()), num_shards=3)
    r, c = [1, -2, 1], [1, 2, -1, -1]
    assert p.scaled_col_norm(0, r, c) == approx([7, 1, 6, 5])
    assert p.scaled_col_norm(1, r, c) == approx([math.sqrt(54), 1.0, 6.0, math.sqrt(41)])
    # the same through the stored transpose (row norms of K)
    K = fx.sharder_test_matrix() * np.array(r)[:, None] * np.array(c)[None, :]
    assert p.scaled_row_norm(0, r, c) == approx(np.abs(K).max(axis=1))
    assert p.scaled_row_norm(1, r, c) == approx(np.sqrt((K * K).sum(axis=1)))


# ------------------------------------------------------- weighted average ---
def test_weighted_average(backend):
    # sharded_optimization_utils_test.cc:42-60
    avg, w, n = backend.weighted_average(np.array([[4.0, 1.0], [1.0, 7.0]]), [1.0, 2.0])
    assert avg == approx([2.0, 5.0]) and w == 3.0 and n == 2
    # zero weights are counted as terms but carry no weight (:62-102)
    avg, w, n = backend.weighted_average(np.array([[4.0, 1.0], [1.0, 7.0]]), [0.0, 0.0])
    assert list(avg) == [0.0, 0.0] and w == 0.0 and n == 2


def test_weighted_average_has_no_roundoff_on_constants(backend):
    # sharded_optimization_utils_test.cc:104-116: bit-identical
    data = np.tile(np.array([1.0 / 3.0]), (3, 1))
    avg, _, _ = backend.weighted_average(data, [341.45, 1.4134, 7.23])
    assert avg[0] == 1.0 / 3.0


# -------------------------------------------------------------- ComputeStats -
def test_compute_stats_test_lp(backend):
    # sharded_optimization_utils_test.cc:134-165
    s = mk(backend, fx.test_lp()).compute_stats()
    assert (s.num_variables, s.num_constraints, s.constraint_matrix_num_nonzeros) == (4, 4, 9)
    assert s.constraint_matrix_col_min_l_inf_norm == approx(1.0)
    assert s.constraint_matrix_row_min_l_inf_norm == approx(1.0)
    assert s.constraint_matrix_abs_max == approx(4.0) and s.constraint_matrix_abs_min == approx(1.0)
    assert s.constraint_matrix_abs_avg == approx(14.5 / 9.0)
    assert s.constraint_matrix_l2_norm == approx(math.sqrt(31.25))
    assert s.objective_vector_abs_max == approx(5.5) and s.objective_vector_abs_min == approx(1.0)
    assert s.objective_vector_abs_avg == approx(2.375)
    assert s.objective_vector_l2_norm == approx(math.sqrt(36.25))
    assert s.objective_matrix_num_nonzeros == 0 and s.objective_matrix_abs_max == 0.0
    assert s.objective_matrix_abs_min == 0.0 and math.isnan(s.objective_matrix_abs_avg)
    assert s.objective_matrix_l2_norm == 0.0
    assert s.variable_bound_gaps_num_finite == 1
    assert (s.variable_bound_gaps_max, s.variable_bound_gaps_min, s.variable_bound_gaps_avg, s.variable_bound_gaps_l2_norm) == (1.0, 1.0, 1.0, 1.0)
    assert s.combined_bounds_max == approx(12.0) and s.combined_bounds_min == approx(1.0)
    assert s.combined_bounds_avg == approx(6.0) and s.combined_bounds_l2_norm == approx(math.sqrt(210.0))


def test_compute_stats_tiny_lp(backend):
    # sharded_optimization_utils_test.cc:167-198
    s = mk(backend, fx.tiny_lp()).compute_stats()
    assert (s.num_variables, s.num_constraints, s.constraint_matrix_num_nonzeros) == (4, 3, 8)
    assert s.constraint_matrix_abs_max == approx(2.0) and s.constraint_matrix_abs_avg == approx(1.25)
    assert s.constraint_matrix_l2_norm == approx(math.sqrt(14.0))
    assert s.objective_vector_abs_max == approx(5.0) and s.objective_vector_abs_avg == approx(2.25)
    assert s.objective_vector_l2_norm == approx(math.sqrt(31.0))
    assert s.variable_bound_gaps_num_finite == 4
    assert s.variable_bound_gaps_max == approx(6.0) and s.variable_bound_gaps_min == approx(2.0)
    assert s.variable_bound_gaps_avg == approx(3.75) and s.variable_bound_gaps_l2_norm == approx(math.sqrt(65.0))
    assert s.combined_bounds_avg == approx(20.0 / 3.0) and s.combined_bounds_l2_norm == approx(math.sqrt(194.0))


def test_compute_stats_qp(backend):
    # sharded_optimization_utils_test.cc:200-245
    s = mk(backend, fx.test_diagonal_qp1()).compute_stats()
    assert s.constraint_matrix_l2_norm == approx(math.sqrt(2.0))
    assert s.objective_matrix_num_nonzeros == 2 and s.objective_matrix_abs_max == approx(4.0)
    assert s.objective_matrix_abs_min == approx(1.0) and s.objective_matrix_abs_avg == approx(2.5)
    assert s.objective_matrix_l2_norm == approx(math.sqrt(17.0))
    assert s.variable_bound_gaps_avg == approx(3.5) and s.variable_bound_gaps_l2_norm == approx(math.sqrt(37.0))
    assert s.combined_bounds_max == approx(1.0) and s.combined_bounds_l2_norm == approx(1.0)
    qp = fx.test_diagonal_qp1()
    qp.set_objective_matrix_diagonal([2.0, 0.0])
    s = mk(backend, qp).compute_stats()
    assert s.objective_matrix_num_nonzeros == 1 and s.objective_matrix_abs_max == approx(2.0)
    assert s.objective_matrix_abs_min == approx(2.0) and s.objective_matrix_abs_avg == approx(1.0)
    assert s.objective_matrix_l2_norm == approx(2.0)


def test_compute_stats_degenerate(backend):
    # sharded_optimization_utils_test.cc:247-308
    s = mk(backend, fx.small_invalid_problem_lp()).compute_stats()
    assert s.variable_bound_gaps_num_finite == 0 and s.variable_bound_gaps_max == 0.0
    assert s.variable_bound_gaps_min == 0.0 and math.isnan(s.variable_bound_gaps_avg)
    assert s.variable_bound_gaps_l2_norm == 0.0
    s = mk(backend, fx.lp_without_constraints()).compute_stats()
    assert s.constraint_matrix_num_nonzeros == 0 and s.constraint_matrix_abs_max == 0.0
    assert s.constraint_matrix_abs_min == 0.0 and math.isnan(s.constraint_matrix_abs_avg)
    assert s.constraint_matrix_l2_norm == 0.0 and s.constraint_matrix_col_min_l_inf_norm == 0.0
    assert s.constraint_matrix_row_min_l_inf_norm == 0.0
    assert s.combined_bounds_max == 0.0 and s.combined_bounds_min == 0.0 and math.isnan(s.combined_bounds_avg)
    s = mk(backend, pdlp.QuadraticProgram(0, 0)).compute_stats()
    assert s.num_variables == 0 and s.num_constraints == 0
    assert math.isnan(s.objective_vector_abs_avg) and s.objective_vector_l2_norm == 0.0
    assert math.isnan(s.variable_bound_gaps_avg) and math.isnan(s.combined_bounds_avg)


# ---------------------------------------------------------------- rescaling -
def test_linf_ruiz_one_iteration(backend):
    # sharded_optimization_utils_test.cc:314-322
    r, c = mk(backend, fx.test_lp()).scaling_iterations(0, 1, [1, 2, 1, 3], [0, 1, 2, -1])
    assert r == approx([1 / math.sqrt(2), 1.0, 1.0, 1.0])
    assert c == approx([0.0, 1.0, 2.0 / 3.0, -1.0 / math.sqrt(3.0)])


def test_l2_rescaling_one_iteration(backend):
    # sharded_optimization_utils_test.cc:328-355
    r, c = mk(backend, fx.test_lp()).scaling_iterations(1, 1, [1, 2, 1, 3], [0, 1, 2, -1])
    assert r == approx([1.0 / 3.0 ** 0.5, 1.0, 1.0, 3.0 / 90.0 ** 0.25])
    assert c == approx([0.0, 1.0, 2.0 / 101 ** 0.25, -1.0 / 13.0 ** 0.25])
    r, c = mk(backend, fx.matrix_only_qp([[2.0, 3.0]])).scaling_iterations(1, 1, [1.0], [1.0, 1.0])
    assert r == approx([1.0 / 13.0 ** 0.25])
    assert c == approx([1.0 / math.sqrt(2.0), 1.0 / math.sqrt(3.0)])


def test_linf_ruiz_convergence(backend):
    # sharded_optimization_utils_test.cc:359-373
    p = mk(backend, fx.test_lp())
    r, c = p.scaling_iterations(0, 20, [1, 1, 1, 1], [1, 1, 1, 1])
    assert p.scaled_col_norm(0, r, c) == pytest.approx([1, 1, 1, 1], abs=1e-4)
    assert p.scaled_row_norm(0, r, c) == pytest.approx([1, 1, 1, 1], abs=1e-4)


def test_apply_rescaling_test_lp(backend):
    # sharded_optimization_utils_test.cc:383-398
    p = mk(backend, fx.test_lp())
    r, c = p.apply_rescaling(1, True)
    assert r == pytest.approx([1 / math.sqrt(2.0 * 1.5275), 1 / math.sqrt(1.0 * 0.9574), 1 / math.sqrt(4.0), 1 / math.sqrt(1.5 * 1.1547)], abs=1e-4)
    assert c == pytest.approx([1 / math.sqrt(4.0 * 1.3229), 1 / math.sqrt(0.7071), 1 / math.sqrt(1.5 * 1.4142), 1 / math.sqrt(2.0 * 1.1547)], abs=1e-4)
    # and the QP was rescaled in place: K_ij * r_i * c_j etc.
    d = p.download()
    qp = fx.test_lp()
    K = qp.constraint_matrix.toarray() * r[:, None] * c[None, :]
    import scipy.sparse as sp
    assert d["values"] == approx(sp.csc_matrix(K).data)
    assert d["objective_vector"] == approx(qp.objective_vector * c)
    assert d["constraint_lower_bounds"] == approx(qp.constraint_lower_bounds * r)
    assert d["variable_upper_bounds"] == approx(qp.variable_upper_bounds / c)


def test_rescale_quadratic_program(backend):
    # sharded_quadratic_program_test.cc:125-154
    p = mk(backend, fx.test_diagonal_qp1())
    p.rescale_quadratic_program([1.0, 0.5], [0.5])
    d = p.download()
    assert d["constraint_lower_bounds"][0] == -INF and d["constraint_upper_bounds"] == approx([0.5])
    assert d["variable_lower_bounds"] == approx([1, -4]) and d["variable_upper_bounds"] == approx([2, 8])
    assert d["objective_vector"] == approx([-1, -0.5])
    assert d["values"] == approx([0.5, 0.25])
    assert d["objective_matrix_diagonal"] == approx([4, 0.25])
    # both orientations were rescaled: K x and K^T y agree with the scaled matrix
    assert p.matrix_vector_product([1, 1]) == approx([0.75])
    assert p.transposed_matrix_vector_product([2]) == approx([1.0, 0.5])


# ------------------------------------------------------------- gradients ----
def test_primal_and_dual_gradient_lp(backend):
    # sharded_optimization_utils_test.cc:400-436
    p = mk(backend, fx.test_lp())
    x, y = [0.0, 0.0, 0.0, 3.0], [-1.0, 0.0, 1.0, 1.0]
    g, v = p.compute_primal_gradient(x, p.transposed_matrix_vector_product(y))
    assert g == approx([3.5, -1.0, -1.5, 4.0]) and v == approx(12.0)
    g, v = p.compute_dual_gradient(y, p.matrix_vector_product(x))
    assert g == approx([6.0, 7.0, -4.0, 2.0]) and v == approx(-17.0)


def test_dual_gradient_two_sided(backend):
    # sharded_optimization_utils_test.cc:438-462
    qp = fx.test_lp()
    qp.constraint_lower_bounds[0] = 4
    qp.constraint_lower_bounds[1] = 5
    qp.constraint_upper_bounds[2] = -1
    p = mk(backend, qp)
    g, v = p.compute_dual_gradient([0.0, 0.0, 0.0, -1.0], p.matrix_vector_product([0.0, 0.0, 0.0, 3.0]))
    assert g == approx([0.0, 5.0, -1.0, 4.0]) and v == approx(-1.0)


def test_gradients_qp(backend):
    # sharded_optimization_utils_test.cc:509-544
    p = mk(backend, fx.test_diagonal_qp1())
    g, v = p.compute_primal_gradient([1.0, 2.0], p.transposed_matrix_vector_product([-2.0]))
    assert g == approx([5.0, 3.0]) and v == approx(7.0)
    g, v = p.compute_dual_gradient([-2.0], p.matrix_vector_product([1.0, 2.0]))
    assert g == approx([-2.0]) and v == approx(-2.0)


# ------------------------------------------------------------ projections ---
def test_projections(backend):
    # sharded_optimization_utils_test.cc:626-648
    p = mk(backend, fx.test_lp())
    assert list(p.project_to_primal_variable_bounds([-3, -3, 5, 5])) == [-3, -2, 5, 3.5]
    assert list(p.project_to_primal_variable_bounds([-3, -3, 5, 5], use_feasibility_bounds=True)) == [-3, 0, 0, 0]
    assert list(p.project_to_dual_variable_bounds([1, 1, -1, -1])) == [1, 0, 0, -1]


# -------------------------------------------------- convergence information --
def _check_scaled_and_unscaled(be, qp_fn, x, y, off_p, off_d, expected):
    """iteration_stats_test.cc:50-99: the same KKT quantities must come out
    unscaled, and after an arbitrary diagonal rescale of the problem."""
    x, y = np.array(x, dtype=float), np.array(y, dtype=float)

    def check(ci, exp):
        for k, v in exp.items():
            got = getattr(ci, k)
            if math.isinf(v):
                assert got == v, k
            else:
                assert got == pytest.approx(v, rel=1e-9, abs=1e-9), k

    p = be.problem(qp_fn(), num_threads=2, num_shards=10)
    check(p.compute_convergence_information(None, None, None, x, y, off_p, off_d), expected)
    col = np.where(x != 0, np.abs(x), 1.0)
    row = np.where(y != 0, np.abs(y), 1.0)
    p.rescale_quadratic_program(col, row)
    check(p.compute_convergence_information(None, col, row, x / col, y / row, off_p, off_d), expected)
    check(p.compute_convergence_information(None, None, None, x / col, y / row, off_p, off_d),
          {"primal_objective": expected["primal_objective"], "dual_objective": expected["dual_objective"],
           "l_inf_primal_variable": 1.0, "l_inf_dual_variable": 1.0})


def test_convergence_information_at_optimum(backend):
    # iteration_stats_test.cc:139-161
    _check_scaled_and_unscaled(backend, fx.test_lp, [-1.0, 8.0, 1.0, 2.5], [-2.0, 0.0, 2.375, 2.0 / 3], 1.0, 1.0, dict(
        primal_objective=-34.0, dual_objective=-34.0, corrected_dual_objective=-34.0,
        l_inf_primal_residual=0.0, l2_primal_residual=0.0, l_inf_componentwise_primal_residual=0.0,
        l_inf_dual_residual=0.0, l2_dual_residual=0.0, l_inf_componentwise_dual_residual=0.0,
        l_inf_primal_variable=8.0, l2_primal_variable=8.5, l_inf_dual_variable=2.375, l2_dual_variable=3.1756998353818715))


def test_convergence_information_primal_residual(backend):
    # iteration_stats_test.cc:163-188
    _check_scaled_and_unscaled(backend, fx.test_lp, [-1.0, 8.0, 1.0, 3.5], [-2.0, 0.0, 2.375, 2.0 / 3], 1.0, 1.0, dict(
        primal_objective=-33.0, dual_objective=-34.0, corrected_dual_objective=-34.0,
        l_inf_primal_residual=2.0, l2_primal_residual=2.2360679774997896, l_inf_componentwise_primal_residual=0.5,
        l_inf_dual_residual=0.0, l2_dual_residual=0.0, l_inf_componentwise_dual_residual=0.0,
        l_inf_primal_variable=8.0, l2_primal_variable=8.8459030064770662, l_inf_dual_variable=2.375, l2_dual_variable=3.1756998353818715))


def test_convergence_information_dual_residual(backend):
    # iteration_stats_test.cc:190-215
    _check_scaled_and_unscaled(backend, fx.test_lp, [-1.0, 8.0, 1.0, 2.5], [-2.0, -1.0, 2.375, 2.0 / 3], 1.0, 1.0, dict(
        primal_objective=-34.0, dual_objective=-41.0, corrected_dual_objective=-INF,
        l_inf_primal_residual=0.0, l2_primal_residual=0.0, l_inf_componentwise_primal_residual=0.0,
        l_inf_dual_residual=1.0, l2_dual_residual=1.4142135623730950, l_inf_componentwise_dual_residual=0.5,
        l_inf_primal_variable=8.0, l2_primal_variable=8.5, l_inf_dual_variable=2.375, l2_dual_variable=3.3294247918288294))


def test_convergence_information_both_residuals(backend):
    # iteration_stats_test.cc:217-250 (different componentwise offsets)
    _check_scaled_and_unscaled(backend, fx.test_lp, [-1.0, 8.0, 1.0, 3.5], [-2.0, -1.0, 2.375, 2.0 / 3], 3.0, 1.0, dict(
        primal_objective=-33.0, dual_objective=-41.0, corrected_dual_objective=-INF,
        l_inf_primal_residual=2.0, l2_primal_residual=2.2360679774997896, l_inf_componentwise_primal_residual=0.25,
        l_inf_dual_residual=1.0, l2_dual_residual=1.4142135623730950, l_inf_componentwise_dual_residual=0.5,
        l_inf_primal_variable=8.0, l2_primal_variable=8.8459030064770662, l_inf_dual_variable=2.375, l2_dual_variable=3.3294247918288294))


def test_convergence_information_qp_at_optimum(backend):
    # iteration_stats_test.cc:252-271
    _check_scaled_and_unscaled(backend, fx.test_diagonal_qp1, [1.0, 0.0], [-1.0], 1.0, 1.0, dict(
        primal_objective=6.0, dual_objective=6.0, corrected_dual_objective=6.0,
        l_inf_primal_residual=0.0, l2_primal_residual=0.0, l_inf_dual_residual=0.0, l2_dual_residual=0.0,
        l_inf_primal_variable=1.0, l2_primal_variable=1.0, l_inf_dual_variable=1.0, l2_dual_variable=1.0))


def _params(handle):
    p = pdlp.PrimalDualHybridGradientParams()
    p.handle_some_primal_gradients_on_finite_bounds_as_residuals = handle
    return p


def test_gap_residuals_zero_primal(backend):
    # iteration_stats_test.cc:273-323
    p = backend.problem(fx.test_lp(), num_threads=2, num_shards=10)
    x, y = np.zeros(4), [1.0, 0.0, 0.0, -1.0]
    ci = p.compute_convergence_information(_params(True), None, None, x, y)
    assert ci.dual_objective == approx(-3.0) and ci.corrected_dual_objective == -INF
    assert ci.l_inf_dual_residual == approx(3.5) and ci.l2_dual_residual == approx(5.0497524691810389)
    ci = p.compute_convergence_information(_params(False), None, None, x, y)
    assert ci.dual_objective == approx(-7.0) and ci.corrected_dual_objective == -INF
    assert ci.l_inf_dual_residual == approx(3.5) and ci.l2_dual_residual == approx(4.6097722286464436)


def test_gap_residuals_nonzero_primal(backend):
    # iteration_stats_test.cc:325-373
    p = backend.problem(fx.test_lp(), num_threads=2, num_shards=10)
    x, y = [0.0, 0.0, 4.0, 3.0], [1.0, 0.0, 0.0, -1.0]
    ci = p.compute_convergence_information(_params(True), None, None, x, y)
    assert ci.dual_objective == approx(-13.0) and ci.l2_dual_residual == approx(4.6097722286464436)
    ci = p.compute_convergence_information(_params(False), None, None, x, y)
    assert ci.dual_objective == approx(-7.0) and ci.l2_dual_residual == approx(4.6097722286464436)


def test_convergence_information_simple_qp(backend):
    # iteration_stats_test.cc:375-418
    p = backend.problem(fx.test_diagonal_qp1(), num_threads=2, num_shards=10)
    ci = p.compute_convergence_information(_params(True), None, None, [1.0, 2.0], [0.0])
    assert ci.dual_objective == approx(8.0) and ci.corrected_dual_objective == approx(2.0)
    assert ci.l_inf_dual_residual == approx(1.0) and ci.l2_dual_residual == approx(1.0)
    ci = p.compute_convergence_information(_params(False), None, None, [1.0, 2.0], [0.0])
    assert ci.dual_objective == approx(2.0) and ci.corrected_dual_objective == approx(2.0)
    assert ci.l_inf_dual_residual == 0.0 and ci.l2_dual_residual == 0.0


def test_corrected_dual_objective(backend):
    # iteration_stats_test.cc:488-574
    p = backend.problem(fx.test_lp(), num_threads=2, num_shards=10)
    ci = p.compute_convergence_information(None, None, None, [0, 0, 6, 2.5], [-2, 0, 2.375, 1])
    assert ci.dual_objective == approx(-36.5) and ci.corrected_dual_objective == approx(-36.5)
    ci = p.compute_convergence_information(_params(True), None, None, [0, 0, 2, 2.5], [-2, 0, 2.375, 1])
    assert ci.dual_objective == approx(-33.5) and ci.corrected_dual_objective == approx(-36.5)
    assert ci.l_inf_dual_residual == approx(0.5) and ci.l2_dual_residual == approx(0.5)
    assert ci.l_inf_componentwise_dual_residual == approx(0.25)
    ci = p.compute_convergence_information(_params(False), None, None, [0, 0, 2, 2.5], [-2, 0, 2.375, 1])
    assert ci.dual_objective == approx(-36.5) and ci.corrected_dual_objective == approx(-36.5)
    assert ci.l_inf_dual_residual == 0.0 and ci.l_inf_componentwise_dual_residual == 0.0
    q = backend.problem(fx.test_diagonal_qp1(), num_threads=2, num_shards=10)
    assert q.compute_convergence_information(None, None, None, [-2.0, 2.0], [-3.0]).corrected_dual_objective == approx(-28.0)


# ------------------------------------------------ infeasibility information --
def _check_infeasibility(be, qp, primal_ray, dual_ray, x_res, expected):
    # iteration_stats_test.cc:101-137 (unscaled, then rescaled by the rays)
    primal_ray, dual_ray, x_res = (np.array(v, dtype=float) for v in (primal_ray, dual_ray, x_res))

    def check(info):
        for k, v in expected.items():
            got = getattr(info, k)
            if math.isinf(v):
                assert got == v, k
            else:
                assert got == pytest.approx(v, rel=1e-9, abs=1e-12), k

    p = be.problem(qp, num_threads=2, num_shards=2)
    check(p.compute_infeasibility_information(None, None, None, primal_ray, dual_ray, x_res))
    col = np.where(primal_ray != 0, np.abs(primal_ray), 1.0)
    row = np.where(dual_ray != 0, np.abs(dual_ray), 1.0)
    p.rescale_quadratic_program(col, row)
    check(p.compute_infeasibility_information(None, col, row, primal_ray / col, dual_ray / row, x_res / col))


def test_infeasibility_information(backend):
    # iteration_stats_test.cc:420-486
    _check_infeasibility(backend, fx.small_primal_infeasible_lp(), [0.0, 0.0], [-1.0, -1.0], [0.0, 0.0], dict(
        max_primal_ray_infeasibility=0, primal_ray_linear_objective=0, primal_ray_quadratic_norm=0,
        max_dual_ray_infeasibility=0, dual_ray_objective=1))
    _check_infeasibility(backend, fx.small_primal_infeasible_lp(), [2.0, 1.0], [-1.0, -3.0], [2.0, 1.0], dict(
        max_primal_ray_infeasibility=0.5, primal_ray_linear_objective=1.5, primal_ray_quadratic_norm=0,
        max_dual_ray_infeasibility=0.66666666666666663, dual_ray_objective=1.6666666666666667))
    _check_infeasibility(backend, fx.small_primal_infeasible_lp(), [0.0, 0.0], [1.0, 1.0], [0.0, 0.0], dict(
        max_dual_ray_infeasibility=0.0, dual_ray_objective=-INF))
    lp = fx._qp([[1.0]], [2], [INF], [0], [1], [1.0])
    _check_infeasibility(backend, lp, [0.0], [1.0], [1.0], dict(max_dual_ray_infeasibility=0.0, dual_ray_objective=1.0))


def test_reduced_costs(backend):
    # iteration_stats_test.cc:607-649
    p = backend.problem(fx.test_lp(), num_threads=2, num_shards=10)
    x, y = [0.0, -2.0, 6.0, 3.5], [1.0, 0.0, 0.0, -2.0]
    assert p.reduced_costs(None, x, y) == approx([3.5, -3.0, 1.0, -3.0])
    assert p.reduced_costs(None, x, y, use_zero_primal_objective=True) == approx([-2.0, -1.0, 2.0, -4.0])
    q = backend.problem(fx.test_diagonal_qp1(), num_threads=2, num_shards=10)
    assert q.reduced_costs(None, [1.0, 2.0], [0.0]) == approx([3.0, 1.0])
    assert list(q.reduced_costs(None, [1.0, 2.0], [0.0], use_zero_primal_objective=True)) == [0.0, 0.0]


# ------------------------------------------------------------ trust region --
TR_CASES = [
    # (objective, lb, ub, center, weights, radius, expected solution, expected value)  trust_region_test.cc:54-339
    ([1.0, 1.0], [-INF, -INF], [INF, INF], [2.0, -5.0], [1, 1], math.sqrt(2.0), [1.0, -6.0], -2.0),
    ([1.0, -1.0, 1.0], [2.0, -INF, -INF], [INF, INF, INF], [2.0, -5.0, 1.0], [1, 1, 1], math.sqrt(2.0), [2.0, -4.0, 0.0], -2.0),
    ([1.0, -1.0], [2.0, -INF], [INF, -5.0], [2.0, -5.0], [1, 1], 1.0, [2.0, -5.0], 0.0),
    ([1.0, -1.0, 1.0], [2.0, -INF, 0.5], [INF, -5.0, INF], [2.0, -5.0, 1.0], [1, 1, 1], 1.0, [2.0, -5.0, 0.5], -0.5),
    ([1.0, -1.0, 1.0], [2.0, -INF, 0.5], [INF, -5.0, INF], [2.0, -5.0, 1.0], [1, 1, 1], 0.0, [2.0, -5.0, 1.0], 0.0),
    ([1.0, -1.0, 1.0], [2.0, -INF, 0.5], [INF, -5.0, INF], [2.0, -5.0, 1.0], [1, 1, 1], INF, [2.0, -5.0, 0.5], -0.5),
    ([2.0, 1.0], [1.0, 0.0], [INF, INF], [2.0, 1.0], [1, 1], math.sqrt(1.25), [1.0, 0.5], -2.5),
    ([0.0], [-INF], [INF], [2.0], [1], 1.0, [2.0], 0.0),
    # weighted, trust_region_test.cc:350-501
    ([1.0, 2.0], [-INF, -INF], [INF, INF], [2.0, -5.0], [1.0, 2.0], math.sqrt(3.0), [1.0, -6.0], -3.0),
    ([0.5, -2.0, 3.0], [2.0, -INF, -INF], [INF, INF, INF], [2.0, -5.0, 1.0], [0.5, 2.0, 3.0], math.sqrt(5.0), [2.0, -4.0, 0.0], -5.0),
    ([1.0, 2.0], [1.0, 0.0], [INF, INF], [2.0, 1.0], [0.5, 2.0], 1.0, [1.0, 0.5], -2.0),
    ([1000.0, 2.0], [1.0, 0.0], [INF, INF], [2.0, 1.0], [500.0, 2.0], math.sqrt(500.5), [1.0, 0.5], -1001.0),
]


@pytest.mark.parametrize("case", TR_CASES)
def test_solve_trust_region(backend, case):
    obj, lb, ub, center, w, radius, sol, val = case
    r = backend.solve_trust_region(obj, lb, ub, center, w, radius)
    # the reference demands EigenArrayEq / DOUBLE_EQ for the linear-time solver,
    # except DoubleNear(1e-13) on the two "hits bounds" weighted cases (:440-501)
    tol = 1e-13 if w[0] in (0.5, 500.0) and len(w) == 2 else 4e-16
    assert r.solution == pytest.approx(sol, rel=tol, abs=tol)
    assert r.objective_value == pytest.approx(val, rel=9e-16, abs=4e-16)


@pytest.mark.parametrize("case", TR_CASES)
def test_solve_diagonal_trust_region_linear(backend, case):
    obj, lb, ub, center, w, radius, sol, val = case
    if math.isinf(radius):
        tol = 1e-6
    r = backend.solve_diagonal_trust_region(obj, np.zeros(len(obj)), lb, ub, center, w, radius, 1e-8)
    assert r.solution == pytest.approx(sol, abs=1e-6)
    assert r.objective_value == pytest.approx(val, abs=2e-6 * max(1.0, abs(val)))


def test_solve_diagonal_trust_region_qp(backend):
    # trust_region_test.cc:965-1000: OneDimQp data, primal weight 1
    r = backend.solve_diagonal_trust_region([2, -1], [2.0, 0.0], [-INF, -INF], [INF, INF], [0, -1], [0.5, 0.5], 0.5, 1e-6)
    assert r.solution == pytest.approx([-0.5, -0.5], abs=1e-6)
    assert r.solution_step_size == pytest.approx(4.0, rel=1e-6)
    assert r.objective_value == pytest.approx(-1.25, abs=1e-6)


def test_trust_region_random_against_bruteforce(backend):
    """Property: the solution lies on the ball (or everything is at a bound) and
    no feasible point on the path has a lower objective."""
    rng = np.random.default_rng(7)
    n = 2000
    obj = rng.normal(size=n) * (rng.uniform(size=n) < 0.8)
    center = rng.normal(size=n)
    lb = np.where(rng.uniform(size=n) < 0.5, center - rng.exponential(size=n), -INF)
    ub = np.where(rng.uniform(size=n) < 0.5, center + rng.exponential(size=n), INF)
    w = rng.uniform(0.1, 3.0, size=n)
    r = backend.solve_trust_region(obj, lb, ub, center, w, 5.0)
    d = r.solution - center
    assert np.all(r.solution >= lb) and np.all(r.solution <= ub)
    assert math.sqrt(float(np.sum(w * d * d))) == pytest.approx(5.0, rel=1e-9)
    assert r.objective_value == pytest.approx(float(obj @ d), rel=1e-9)
    proj = np.clip(center - r.solution_step_size * obj / w, lb, ub)
    assert r.solution == pytest.approx(proj, rel=1e-12, abs=1e-12)


# ---------------------------------------------- localized Lagrangian bounds --
def test_localized_bounds_zero_gap_at_optimal(backend):
    # trust_region_test.cc:617-636
    p = mk(backend, fx.test_lp())
    for diag in (False, True):
        b = p.compute_localized_lagrangian_bounds([-1.0, 8.0, 1.0, 2.5], [-2.0, 0.0, 2.375, 2.0 / 3.0], 1.0, 1.0,
                                                  use_diagonal_qp_trust_region_solver=diag,
                                                  diagonal_qp_trust_region_solver_tolerance=1e-2)
        assert b.radius == 1.0 and b.lagrangian_value == approx(-20.0)
        assert b.lower_bound == pytest.approx(-20.0, abs=1e-9) and b.upper_bound == pytest.approx(-20.0, abs=1e-9)


def test_localized_bounds_optimal_in_range(backend):
    # trust_region_test.cc:640-673 and :760-794 (cached products)
    p = mk(backend, fx.test_lp())
    dist = math.sqrt(0.5 * (1.0 + 64.0 + 1.0 + 0.25) + 0.5 * (4.0 + 2.375 ** 2 + 4.0 / 9.0))
    for kw in ({}, {"primal_product": [6.0, 0.0, 0.0, -3.0], "dual_product": np.zeros(4)}):
        for diag in (False, True):
            b = p.compute_localized_lagrangian_bounds([0.0, 0.0, 0.0, 3.0], np.zeros(4), 1.0, dist,
                                                      use_diagonal_qp_trust_region_solver=diag,
                                                      diagonal_qp_trust_region_solver_tolerance=1e-6, **kw)
            assert b.lagrangian_value == approx(3.0) and b.lower_bound <= -20.0 and b.upper_bound >= -20.0


def test_localized_bounds_closed_form(backend):
    # trust_region_test.cc:675-728 (Euclidean branch)
    p = mk(backend, fx.test_lp())
    for diag in (False, True):
        b = p.compute_localized_lagrangian_bounds([0.0, 0.0, 0.0, 3.0], np.zeros(4), 1.0, 0.1,
                                                  use_diagonal_qp_trust_region_solver=diag,
                                                  diagonal_qp_trust_region_solver_tolerance=1e-6)
        assert b.lagrangian_value == approx(3.0)
        assert b.lower_bound == pytest.approx(3.0 - 0.1 * math.sqrt(2) * 36.25 / math.sqrt(76.25), abs=1e-6)
        assert b.upper_bound == pytest.approx(3.0 + 0.1 * math.sqrt(2) * 40 / math.sqrt(76.25), abs=1e-6)


def test_localized_bounds_one_dim(backend):
    # trust_region_test.cc:884-963
    p = mk(backend, fx.one_dim_lp())
    b = p.compute_localized_lagrangian_bounds([0.0], [-1.0], 1.0, 1.0 / math.sqrt(2.0))
    assert b.lagrangian_value == approx(-1.0)
    assert b.lower_bound == approx(-1.0 - 4.0 / math.sqrt(5)) and b.upper_bound == approx(-1.0 + 1.0 / math.sqrt(5))
    b = p.compute_localized_lagrangian_bounds([0.0], [-1.0], 100.0, 1.0 / math.sqrt(2.0))
    assert b.upper_bound - b.lower_bound == approx(10.00199980003999)
    b = p.compute_localized_lagrangian_bounds([0.0], [-1.0], 100.0, 1.0 / math.sqrt(2.0),
                                              use_diagonal_qp_trust_region_solver=True,
                                              diagonal_qp_trust_region_solver_tolerance=1e-8)
    assert b.upper_bound - b.lower_bound == pytest.approx(10.00199980003999, rel=1e-7)


def test_localized_bounds_one_dim_qp_diagonal_solver(backend):
    # trust_region_test.cc:1073-1092: lb = -1.75, ub = -0.5
    p = mk(backend, fx.one_dim_qp())
    b = p.compute_localized_lagrangian_bounds([0.0], [-1.0], 1.0, 0.5, use_diagonal_qp_trust_region_solver=True,
                                              diagonal_qp_trust_region_solver_tolerance=1e-6)
    assert b.lagrangian_value == approx(-1.0)
    assert b.lower_bound == pytest.approx(-1.75, abs=1e-5) and b.upper_bound == pytest.approx(-0.5, abs=1e-5)


def test_localized_bounds_max_norm(backend):
    # PrimalDualNorm::kMaxNorm (trust_region.cc:855-884): trust_region_test.cc:617-636 (zero gap at the
    # optimum), :675-712 (closed form), :884-963 (one-dimensional LP, primal weights 1 and 100)
    p = mk(backend, fx.test_lp())
    b = p.compute_localized_lagrangian_bounds([-1.0, 8.0, 1.0, 2.5], [-2.0, 0.0, 2.375, 2.0 / 3.0], 1.0, 1.0, max_norm=True)
    assert b.lagrangian_value == approx(-20.0)
    assert b.lower_bound == pytest.approx(-20.0, abs=1e-9) and b.upper_bound == pytest.approx(-20.0, abs=1e-9)
    b = p.compute_localized_lagrangian_bounds([0.0, 0.0, 0.0, 3.0], np.zeros(4), 1.0, 0.1, max_norm=True)
    assert b.lagrangian_value == approx(3.0)
    assert b.lower_bound == pytest.approx(3.0 - 0.1 * math.sqrt(2) * math.sqrt(36.25), abs=1e-6)
    assert b.upper_bound == pytest.approx(3.0 + 0.1 * math.sqrt(2) * math.sqrt(40.0), abs=1e-6)
    q = mk(backend, fx.one_dim_lp())
    b = q.compute_localized_lagrangian_bounds([0.0], [-1.0], 1.0, 1.0 / math.sqrt(2.0), max_norm=True)
    assert b.lagrangian_value == approx(-1.0) and b.lower_bound == approx(-3.0) and b.upper_bound == approx(0.0)
    b = q.compute_localized_lagrangian_bounds([0.0], [-1.0], 100.0, 1.0 / math.sqrt(2.0), max_norm=True)
    assert b.lower_bound == approx(-1.2) and b.upper_bound == approx(9.0)
