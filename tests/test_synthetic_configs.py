"""Parity on the synthetic benchmark configurations (BASELINE.json configs 1-4 =
SURVEY.md 8d C2..C5) at sizes the CPU oracle finishes in seconds, plus
size-independent properties at full C2 size. north_star's three checks:

  * per-kernel results within 1e-12 relative (fp64),
  * after a fixed iteration count with restarts disabled the iterates agree
    within 1e-9 relative,
  * full solves reach the same termination status with objectives and relative
    KKT error within eps_optimal.
"""
import numpy as np
import pytest

from ortools_b200 import pdlp, synthetic

TR = pdlp.TerminationReason
P = pdlp.PrimalDualHybridGradientParams
SMALL = {"c2": 0.004, "c3": 0.002, "c4": 0.001, "c5": 0.003}


def small_problem(name):
    return synthetic.CONFIGS[name](scale=SMALL[name])


def eps_params(eps, iteration_limit=100000):
    p = P()
    p.termination_criteria.simple_optimality_criteria.eps_optimal_absolute = eps
    p.termination_criteria.simple_optimality_criteria.eps_optimal_relative = eps
    p.termination_criteria.iteration_limit = iteration_limit
    return p


def chosen(log):
    return [c for c in log.solution_stats.convergence_information if c.candidate_type == log.solution_type][0]


def rel_kkt(qp, res, eps):
    """Relative KKT error of the returned point, recomputed in numpy
    (termination.cc:43-97, L2 norm)."""
    k = qp.constraint_matrix
    x, y = res.primal_solution, res.dual_solution
    kx = k @ x
    viol = np.maximum(kx - qp.constraint_upper_bounds, 0) + np.maximum(qp.constraint_lower_bounds - kx, 0)
    bc = np.where(np.isfinite(qp.constraint_lower_bounds), np.abs(qp.constraint_lower_bounds), 0)
    bc = np.maximum(bc, np.where(np.isfinite(qp.constraint_upper_bounds), np.abs(qp.constraint_upper_bounds), 0))
    return np.linalg.norm(viol) / (1.0 + np.linalg.norm(bc))


# ---------------------------------------------------------------- generators / oracle
@pytest.mark.parametrize("name", ["c2", "c4"])
def test_oracle_recovers_planted_objective(name, oracle_backend):
    qp, info = small_problem(name)
    res = oracle_backend.primal_dual_hybrid_gradient(qp, eps_params(1e-8))
    assert res.solve_log.termination_reason == TR.TERMINATION_REASON_OPTIMAL
    ci = chosen(res.solve_log)
    assert ci.primal_objective == pytest.approx(info["objective"], rel=1e-6, abs=1e-6)
    assert ci.dual_objective == pytest.approx(info["objective"], rel=1e-6, abs=1e-6)


def test_generators_have_the_documented_shape():
    qp, _ = small_problem("c2")
    k = qp.constraint_matrix.tocsr()
    assert k.shape[1] == 2 * k.shape[0] and np.all(np.diff(k.indptr) == 20)
    qp, _ = small_problem("c3")
    k = qp.constraint_matrix
    assert np.all(np.diff(k.indptr) == 4)  # every arc: supply, demand, two bundles
    lens = np.diff(k.tocsr().indptr)
    assert lens.std() / lens.mean() > 1.0  # skewed rows
    qp, _ = small_problem("c5")
    assert qp.objective_matrix is not None and qp.constraint_matrix.shape[1] == 5 * qp.constraint_matrix.shape[0]


# ---------------------------------------------------------------- per-kernel parity
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c2", "c3", "c4", "c5"])
def test_spmv_pair_parity(name, b200_backend, oracle_backend):
    qp, _ = small_problem(name)
    k = qp.constraint_matrix
    rng = np.random.default_rng(1)
    x, y = rng.normal(size=k.shape[1]), rng.normal(size=k.shape[0])
    dev, ref = b200_backend.problem(qp), oracle_backend.problem(qp)
    a, b = dev.matrix_vector_product(x), ref.matrix_vector_product(x)
    assert np.max(np.abs(a - b) / (abs(k) @ np.abs(x) + 1e-300)) < 1e-12
    a, b = dev.transposed_matrix_vector_product(y), ref.transposed_matrix_vector_product(y)
    assert np.max(np.abs(a - b) / (abs(k).T @ np.abs(y) + 1e-300)) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c2", "c3", "c5"])
def test_rescaling_and_stats_parity(name, b200_backend, oracle_backend):
    qp, _ = small_problem(name)
    dev, ref = b200_backend.problem(qp), oracle_backend.problem(qp)
    r1, c1 = dev.apply_rescaling(5, True)
    r2, c2 = ref.apply_rescaling(5, True)
    np.testing.assert_allclose(r1, r2, rtol=1e-12)
    np.testing.assert_allclose(c1, c2, rtol=1e-12)
    s1, s2 = dev.compute_stats(), ref.compute_stats()
    for f in ("constraint_matrix_abs_max", "constraint_matrix_abs_min", "constraint_matrix_l2_norm", "objective_vector_l2_norm",
              "combined_bounds_l2_norm", "constraint_matrix_col_min_l_inf_norm", "constraint_matrix_row_min_l_inf_norm"):
        assert getattr(s1, f) == pytest.approx(getattr(s2, f), rel=1e-12), f


# ---------------------------------------------------------------- fixed-iteration iterates
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c2", "c3", "c5"])
def test_fixed_iteration_iterates_agree(name, b200_backend, oracle_backend):
    # "restarts disabled": NO_RESTARTS with the primal weight frozen
    # (primal_dual_hybrid_gradient_test.cc:837-856)
    qp, _ = small_problem(name)
    p = P()
    p.restart_strategy = P.NO_RESTARTS
    p.primal_weight_update_smoothing = 0.0
    p.termination_criteria.simple_optimality_criteria.eps_optimal_absolute = 0.0
    p.termination_criteria.simple_optimality_criteria.eps_optimal_relative = 0.0
    p.termination_criteria.iteration_limit = 48
    got = b200_backend.primal_dual_hybrid_gradient(qp, p)
    ref = oracle_backend.primal_dual_hybrid_gradient(qp, p)
    assert got.solve_log.termination_reason == ref.solve_log.termination_reason == TR.TERMINATION_REASON_ITERATION_LIMIT
    assert got.solve_log.iteration_count == ref.solve_log.iteration_count == 48
    assert got.solve_log.solution_stats.cumulative_rejected_steps == ref.solve_log.solution_stats.cumulative_rejected_steps
    for a, b in ((got.primal_solution, ref.primal_solution), (got.dual_solution, ref.dual_solution)):
        assert np.linalg.norm(a - b) <= 1e-9 * max(1.0, np.linalg.norm(b))
    assert got.solve_log.solution_stats.step_size == pytest.approx(ref.solve_log.solution_stats.step_size, rel=1e-9)


# ---------------------------------------------------------------- full solves
@pytest.mark.gpu
@pytest.mark.parametrize("eps", [1e-4, 1e-8])
@pytest.mark.parametrize("name", ["c2", "c3", "c4", "c5"])
def test_full_solve_matches_reference(name, eps, b200_backend, oracle_backend):
    qp, info = small_problem(name)
    p = eps_params(eps)
    got = b200_backend.primal_dual_hybrid_gradient(qp, p)
    ref = oracle_backend.primal_dual_hybrid_gradient(qp, p)
    assert got.solve_log.termination_reason == ref.solve_log.termination_reason == TR.TERMINATION_REASON_OPTIMAL
    g, r = chosen(got.solve_log), chosen(ref.solve_log)
    scale = 1.0 + abs(r.primal_objective)
    assert abs(g.primal_objective - r.primal_objective) <= 4 * eps * scale
    assert abs(g.dual_objective - r.dual_objective) <= 4 * eps * scale
    # relative KKT error of both within eps_optimal
    for ci in (g, r):
        assert abs(ci.primal_objective - ci.dual_objective) <= eps * (1 + abs(ci.primal_objective) + abs(ci.dual_objective))
    assert rel_kkt(qp, got, eps) <= 2 * eps
    if "objective" in info:
        assert g.primal_objective == pytest.approx(info["objective"], abs=10 * eps * scale)
    assert got.solve_log.gpu_kernel_launches > 0


# ---------------------------------------------------------------- resident sessions
@pytest.mark.gpu
def test_session_chunks_equal_one_solve(b200_backend):
    qp, _ = small_problem("c2")
    p = eps_params(1e-6)
    one = b200_backend.primal_dual_hybrid_gradient(qp, p)
    s = b200_backend.session(qp, p)
    st = None
    for target in (10, 64, 65, 130, 10**9):
        st = s.advance(target)
        if not st.terminated:
            assert st.iterations_completed == target
    assert st.terminated and st.termination_reason == TR.TERMINATION_REASON_OPTIMAL
    res = s.finish()
    s.close()
    assert res.solve_log.iteration_count == one.solve_log.iteration_count
    np.testing.assert_array_equal(res.primal_solution, one.primal_solution)
    np.testing.assert_array_equal(res.dual_solution, one.dual_solution)


@pytest.mark.gpu
def test_session_finish_before_termination_behaves_like_interrupt(b200_backend):
    qp, _ = small_problem("c2")
    s = b200_backend.session(qp, eps_params(0.0))
    s.enable_timing(True, 2)
    st = s.advance(100)
    assert not st.terminated and st.iterations_completed == 100
    assert st.kernel_samples[1] > 0 and st.kernel_ms[1] > 0 and st.device_total_ms > 0
    res = s.finish()
    assert res.solve_log.termination_reason == TR.TERMINATION_REASON_INTERRUPTED_BY_USER
    assert res.solve_log.iteration_count == 100 and res.primal_solution.size == qp.constraint_matrix.shape[1]


# ---------------------------------------------------------------- full size, size-independent properties
@pytest.mark.gpu
def test_c2_full_size_properties(b200_backend):
    qp, info = synthetic.c2(scale=1.0)
    k = qp.constraint_matrix
    dev = b200_backend.problem(qp)
    rng = np.random.default_rng(7)
    x1, x2, y = rng.normal(size=k.shape[1]), rng.normal(size=k.shape[1]), rng.normal(size=k.shape[0])
    # adjoint identity <K x, y> = <x, K^T y> and linearity, relative to sum |a||x||y|
    kx1, kx2 = dev.matrix_vector_product(x1), dev.matrix_vector_product(x2)
    kty = dev.transposed_matrix_vector_product(y)
    scale = float(np.abs(y) @ (abs(k) @ np.abs(x1)))
    assert abs(kx1 @ y - x1 @ kty) <= 1e-12 * scale
    k12 = dev.matrix_vector_product(x1 + 2.0 * x2)
    assert np.max(np.abs(k12 - (kx1 + 2.0 * kx2)) / (abs(k) @ (np.abs(x1) + 2 * np.abs(x2)))) < 1e-12
    # against scipy on the host (fp64, different summation order)
    assert np.max(np.abs(kx1 - k @ x1) / (abs(k) @ np.abs(x1))) < 1e-12
    dev.close()
    # full solve: the planted optimal objective is recovered at eps 1e-4
    res = b200_backend.primal_dual_hybrid_gradient(qp, eps_params(1e-4))
    assert res.solve_log.termination_reason == TR.TERMINATION_REASON_OPTIMAL
    ci = chosen(res.solve_log)
    assert ci.primal_objective == pytest.approx(info["objective"], rel=5e-4)
    assert ci.dual_objective == pytest.approx(info["objective"], rel=5e-4)


def fixed_iteration_params(iterations):
    p = P()
    p.restart_strategy = P.NO_RESTARTS
    p.primal_weight_update_smoothing = 0.0
    p.termination_criteria.simple_optimality_criteria.eps_optimal_absolute = 0.0
    p.termination_criteria.simple_optimality_criteria.eps_optimal_relative = 0.0
    p.termination_criteria.iteration_limit = iterations
    p.num_threads = 16
    return p


@pytest.mark.gpu
def test_c3_full_size_meets_oracle(b200_backend, oracle_backend):
    """C3 at FULL size (8 022 x 10 M, 40 M nonzeros; 45 rows of more than 10^5 nonzeros, the longest
    979 143): the split-row path (virtual slots + k_sell_fixup) against the CPU oracle -- the SpMV
    pair within 1e-12 and 16 fixed PDHG iterations (restarts disabled) within 1e-9."""
    qp, _ = synthetic.c3(scale=1.0)
    k = qp.constraint_matrix
    lens = np.diff(k.tocsr().indptr)
    assert lens.max() > 500_000 and (lens > 100_000).sum() >= 10
    rng = np.random.default_rng(11)
    x, y = rng.normal(size=k.shape[1]), rng.normal(size=k.shape[0])
    dev, ref = b200_backend.problem(qp), oracle_backend.problem(qp)
    a, b = dev.matrix_vector_product(x), ref.matrix_vector_product(x)
    assert np.max(np.abs(a - b) / (abs(k) @ np.abs(x) + 1e-300)) < 1e-12
    a, b = dev.transposed_matrix_vector_product(y), ref.transposed_matrix_vector_product(y)
    assert np.max(np.abs(a - b) / (abs(k).T @ np.abs(y) + 1e-300)) < 1e-12
    dev.close()
    ref.close()
    p = fixed_iteration_params(16)
    got = b200_backend.primal_dual_hybrid_gradient(qp, p)
    want = oracle_backend.primal_dual_hybrid_gradient(qp, p)
    assert got.solve_log.termination_reason == want.solve_log.termination_reason == TR.TERMINATION_REASON_ITERATION_LIMIT
    assert got.solve_log.iteration_count == want.solve_log.iteration_count == 16
    assert got.solve_log.solution_stats.cumulative_rejected_steps == want.solve_log.solution_stats.cumulative_rejected_steps
    for u, v in ((got.primal_solution, want.primal_solution), (got.dual_solution, want.dual_solution)):
        assert np.linalg.norm(u - v) <= 1e-9 * max(1.0, np.linalg.norm(v))
    assert got.solve_log.solution_stats.step_size == pytest.approx(want.solve_log.solution_stats.step_size, rel=1e-9)
