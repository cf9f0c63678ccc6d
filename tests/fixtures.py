"""Fixture LPs/QPs with known optima, transcribed from the reference's
ortools/pdlp/test_util.cc:35-259 (optima documented in test_util.h:33-202) and
trust_region_test.cc:799-829."""
import numpy as np
import scipy.sparse as sp

from ortools_b200 import pdlp

INF = float("inf")


def _qp(K, lc, uc, lv, uv, c, offset=0.0, q=None):
    K = np.asarray(K, dtype=np.float64)
    m, n = K.shape if K.ndim == 2 else (0, len(c))
    qp = pdlp.QuadraticProgram(n, m)
    qp.constraint_matrix = sp.csc_matrix(K.reshape(m, n))
    qp.constraint_lower_bounds = np.array(lc, dtype=np.float64)
    qp.constraint_upper_bounds = np.array(uc, dtype=np.float64)
    qp.variable_lower_bounds = np.array(lv, dtype=np.float64)
    qp.variable_upper_bounds = np.array(uv, dtype=np.float64)
    qp.objective_vector = np.array(c, dtype=np.float64)
    qp.objective_offset = offset
    if q is not None:
        qp.set_objective_matrix_diagonal(q)
    return qp


def test_lp():  # test_util.cc:35-49; optimum x=[-1,8,1,2.5], y=[-2,0,2.375,2/3], obj -34
    return _qp([[2, 1, 1, 2], [1, 0, 1, 0], [4, 0, 0, 0], [0, 0, 1.5, -1]],
               [12, -INF, -4, -1], [12, 7, INF, 1], [-INF, -2, -INF, 2.5], [INF, INF, 6, 3.5],
               [5.5, -2, -1, 1], -14)


def tiny_lp():  # test_util.cc:69-87; x=[1,0,6,2], y=[0.5,4,0], rc=[0,1.5,-3.5,0], obj -1
    return _qp([[2, 1, 1, 2], [1, 0, 1, 0], [0, 0, 1, -1]], [12, 7, 1], [12, INF, INF],
               [0, 0, 0, 0], [2, 4, 6, 3], [5, 2, 1, 1], -14)


def correlation_clustering_lp():  # test_util.cc:89-108; x=[1,1,0,1,0,0], obj 1
    K = np.zeros((3, 6))
    K[0, 1], K[0, 2], K[0, 5] = -1, 1, -1
    K[1, 3], K[1, 4], K[1, 5] = -1, 1, -1
    K[2, 0], K[2, 1], K[2, 3] = -1, -1, 1
    return _qp(K, [-1] * 3, [INF] * 3, [0] * 6, [1] * 6, [-1, -1, 1, -1, 1, -1], 4)


def correlation_clustering_star_lp():  # test_util.cc:110-129; x=[.5,.5,.5,0,0,0], y=[.5]*3, obj 1.5
    K = np.zeros((3, 6))
    K[0, 0], K[0, 1], K[0, 3] = -1, -1, 1
    K[1, 0], K[1, 2], K[1, 4] = -1, -1, 1
    K[2, 1], K[2, 2], K[2, 5] = -1, -1, 1
    return _qp(K, [-1] * 3, [INF] * 3, [0] * 6, [1] * 6, [-1, -1, -1, 1, 1, 1], 3)


def test_diagonal_qp1():  # test_util.cc:131-146; x=[1,0], y=[-1], rc=[4,0], obj 6
    return _qp([[1, 1]], [-INF], [1], [1, -2], [2, 4], [-1, -1], 5, q=[4.0, 1.0])


def test_diagonal_qp2():  # test_util.cc:148-163; x=[3,1], y=[0], obj -5
    return _qp([[1, -1]], [2], [2], [0, 0], [INF, INF], [-3, -1], 0, q=[1.0, 1.0])


def test_diagonal_qp3():  # test_util.cc:165-180; x=[2,0,1], y=[-1,1], obj 2
    return _qp([[1, 0, -1], [2, 0, 0]], [1, 4], [1, 4], [0, 0, 0], [INF] * 3, [1, 0, -1], 0, q=[0.0, 1.0, 2.0])


def small_invalid_problem_lp():  # test_util.cc:182-193
    return _qp([[1, -1]], [2], [1], [0, 0], [INF, INF], [1, 1])


def small_inconsistent_variable_bounds_lp():  # test_util.cc:195-206
    return _qp([[1, -1]], [-INF], [1], [2, 0], [1, INF], [1, 1])


def small_primal_infeasible_lp():  # test_util.cc:208-222
    return _qp([[1, -1], [-1, 1]], [-INF, -INF], [1, -2], [0, 0], [INF, INF], [1, 1])


def small_dual_infeasible_lp():  # test_util.cc:224-229
    qp = small_primal_infeasible_lp()
    qp.constraint_upper_bounds[1] = 2.0
    qp.objective_vector = -qp.objective_vector
    return qp


def small_primal_dual_infeasible_lp():  # test_util.cc:231-235
    qp = small_primal_infeasible_lp()
    qp.objective_vector = -qp.objective_vector
    return qp


def small_initialization_lp():  # test_util.cc:237-250
    return _qp([[1, 1], [1, 2]], [-INF, -INF], [2, 2], [0.5, 0.5], [2, 2], [-4, 0])


def lp_without_constraints():  # test_util.cc:252-258
    return _qp(np.zeros((0, 2)), [], [], [0, -INF], [INF, 0], [4, 0])


def one_dim_lp():  # trust_region_test.cc:799-811
    return _qp([[1]], [0], [1], [-INF], [INF], [1.0])


def one_dim_qp():  # trust_region_test.cc:813-829
    return _qp([[1]], [0], [1], [-INF], [INF], [1.0], q=[2.0])


def sharder_test_matrix():  # sharder_test.cc:45-60 (3x4)
    return np.array([[7, -0.5, 0, 0], [1, 0, 3, 2], [-1, 0, 0, 5.0]])


def matrix_only_qp(K):
    K = np.asarray(K, dtype=np.float64)
    m, n = K.shape
    return _qp(K, [-INF] * m, [INF] * m, [-INF] * n, [INF] * n, [0.0] * n)
