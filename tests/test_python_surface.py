"""The reference's own Python test, ``ortools/pdlp/python/pdlp_test.py:23-250``, against this
package's module with the same calls (``solvers_pb2`` / ``linear_solver_pb2`` are the runtime
message classes of ``pdlp_proto`` / ``mp_model``; the solver takes the protobuf parameters
directly, like the pybind wrapper)."""
import numpy as np
import pytest
import scipy.sparse

from ortools_b200 import mp_model, pdlp, pdlp_proto

TR = pdlp.TerminationReason


def small_proto_lp():  # pdlp_test.py:26-49: min -2y s.t. x + y <= 1, x, y >= 0
    m = mp_model.MPModelProto(maximize=False, objective_offset=0.0)
    m.variable.add(lower_bound=0, upper_bound=np.inf, objective_coefficient=0, name="x")
    m.variable.add(lower_bound=0, upper_bound=np.inf, objective_coefficient=-2, name="y")
    m.constraint.add(var_index=[0, 1], coefficient=[1, 1], lower_bound=-np.inf, upper_bound=1)
    return m


def small_proto_qp():  # pdlp_test.py:52-76: min 2 x*x s.t. x + y <= 1, x, y >= 0
    m = mp_model.MPModelProto(maximize=False, objective_offset=0.0)
    m.variable.add(lower_bound=0, upper_bound=np.inf, objective_coefficient=0, name="x")
    m.variable.add(lower_bound=0, upper_bound=np.inf, objective_coefficient=0, name="y")
    m.constraint.add(var_index=[0, 1], coefficient=[1, 1], lower_bound=-np.inf, upper_bound=1)
    m.quadratic_objective.qvar1_index.append(0)
    m.quadratic_objective.qvar2_index.append(0)
    m.quadratic_objective.coefficient.append(2)
    return m


def test_validate_quadratic_program_dimensions_for_empty_qp():  # :81-85
    qp = pdlp.QuadraticProgram()
    qp.resize_and_initialize(3, 2)
    pdlp.validate_quadratic_program_dimensions(qp)
    assert pdlp.is_linear_program(qp)


def test_converts_from_tiny_mpmodel_lp_and_qp():  # :86-98
    qp = pdlp.qp_from_mpmodel_proto(small_proto_lp(), relax_integer_variables=False)
    pdlp.validate_quadratic_program_dimensions(qp)
    assert pdlp.is_linear_program(qp)
    assert sorted(qp.objective_vector) == [-2, 0]
    qp = pdlp.qp_from_mpmodel_proto(small_proto_qp(), relax_integer_variables=False)
    pdlp.validate_quadratic_program_dimensions(qp)
    assert not pdlp.is_linear_program(qp)
    assert list(qp.objective_vector) == [0, 0]


def _build(objective, diagonal=None):
    qp = pdlp.QuadraticProgram()
    qp.objective_vector = objective
    qp.constraint_matrix = scipy.sparse.csr_matrix(np.array([[1.0, 1.0]]))
    if diagonal is not None:
        qp.set_objective_matrix_diagonal(diagonal)
    qp.constraint_lower_bounds = [-np.inf]
    qp.constraint_upper_bounds = [1.0]
    qp.variable_lower_bounds = [0.0, 0.0]
    qp.variable_upper_bounds = [np.inf, np.inf]
    qp.variable_names = ["x", "y"]
    return qp


def test_build_lp_and_qp():  # :100-127
    assert pdlp.qp_to_mpmodel_proto(_build([0, -2])) == small_proto_lp()
    assert pdlp.qp_to_mpmodel_proto(_build([0, 0], [4.0])) == small_proto_qp()


def tiny_lp():  # :130-157, optimum x = [1, 0, 6, 2], y = [0.5, 4, 0], reduced costs [0, 1.5, -3.5, 0]
    qp = pdlp.QuadraticProgram()
    qp.objective_offset = -14
    qp.objective_vector = [5, 2, 1, 1]
    qp.constraint_lower_bounds = [12, 7, 1]
    qp.constraint_upper_bounds = [12, np.inf, np.inf]
    qp.variable_lower_bounds = np.zeros(4)
    qp.variable_upper_bounds = [2, 4, 6, 3]
    qp.constraint_matrix = scipy.sparse.csr_matrix(np.array([[2, 1, 1, 2], [1, 0, 1, 0], [0, 0, 1, -1]]))
    return qp


def small_lp():  # :160-188, optimum x = [-1, 8, 1, 2.5], y = [-2, 0, 2.375, 2/3]
    qp = pdlp.QuadraticProgram()
    qp.objective_offset = -14
    qp.objective_vector = [5.5, -2, -1, 1]
    qp.constraint_lower_bounds = [12, -np.inf, -4, -1]
    qp.constraint_upper_bounds = [12, 7, np.inf, 1]
    qp.variable_lower_bounds = [-np.inf, -2, -np.inf, 2.5]
    qp.variable_upper_bounds = [np.inf, np.inf, 6, 3.5]
    qp.constraint_matrix = scipy.sparse.csr_matrix(np.array([[2, 1, 1, 2], [1, 0, 1, 0], [4, 0, 0, 0], [0, 0, 1.5, -1]]))
    return qp


def _tight_params():
    params = pdlp_proto.PrimalDualHybridGradientParamsProto()   # solvers_pb2.PrimalDualHybridGradientParams()
    params.termination_criteria.simple_optimality_criteria.eps_optimal_relative = 0.0
    params.termination_criteria.simple_optimality_criteria.eps_optimal_absolute = 1.0e-10
    return params


def test_iteration_limit(backend):  # :193-202
    params = pdlp_proto.PrimalDualHybridGradientParamsProto()
    params.termination_criteria.iteration_limit = 1
    params.termination_check_frequency = 1
    result = backend.primal_dual_hybrid_gradient(tiny_lp(), params)
    assert result.solve_log.iteration_count <= 1
    assert result.solve_log.termination_reason == TR.TERMINATION_REASON_ITERATION_LIMIT


def test_solution(backend):  # :204-216 (assertSequenceAlmostEqual: 7 places)
    result = backend.primal_dual_hybrid_gradient(tiny_lp(), _tight_params())
    assert result.solve_log.termination_reason == TR.TERMINATION_REASON_OPTIMAL
    np.testing.assert_allclose(result.primal_solution, [1.0, 0.0, 6.0, 2.0], atol=5e-8)
    np.testing.assert_allclose(result.dual_solution, [0.5, 4.0, 0.0], atol=5e-8)
    np.testing.assert_allclose(result.reduced_costs, [0.0, 1.5, -3.5, 0.0], atol=5e-8)


def test_solution_2(backend):  # :218-229
    result = backend.primal_dual_hybrid_gradient(small_lp(), _tight_params())
    assert result.solve_log.termination_reason == TR.TERMINATION_REASON_OPTIMAL
    np.testing.assert_allclose(result.primal_solution, [-1, 8, 1, 2.5], atol=5e-8)
    np.testing.assert_allclose(result.dual_solution, [-2, 0, 2.375, 2 / 3], atol=5e-8)


def test_starting_point(backend):  # :231-250
    params = _tight_params()
    params.l_inf_ruiz_iterations = 0
    params.l2_norm_rescaling = False
    start = pdlp.PrimalAndDualSolution()
    start.primal_solution = [1.0, 0.0, 6.0, 2.0]
    start.dual_solution = [0.5, 4.0, 0.0]
    result = backend.primal_dual_hybrid_gradient(tiny_lp(), params, initial_solution=start)
    assert result.solve_log.termination_reason == TR.TERMINATION_REASON_OPTIMAL
    assert result.solve_log.iteration_count == 0


def test_read_quadratic_program_or_die(tmp_path):  # quadratic_program_io.h:28-40 through the Python surface
    import fixtures
    import bz2
    from ortools_b200 import qp_io
    lp = fixtures.test_lp()
    lp.problem_name, lp.variable_names, lp.constraint_names = "lp", ["a", "b", "c", "d"], ["r0", "r1", "r2", "r3"]
    path = str(tmp_path / "lp.mps")
    qp_io.write_linear_program_to_mps(lp, path)
    got = pdlp.read_quadratic_program_or_die(path, include_names=True)          # the library's C++ reader
    np.testing.assert_array_equal(got.constraint_matrix.toarray(), lp.constraint_matrix.toarray())
    np.testing.assert_array_equal(got.objective_vector, lp.objective_vector)
    assert got.variable_names == lp.variable_names and got.constraint_names == lp.constraint_names and got.problem_name == "lp"
    raw = open(path, "rb").read()
    with bz2.open(path + ".bz2", "wb") as f:                                      # bzip2: the C++ reader too (libbz2 bound at run time)
        f.write(raw)
    got = pdlp.read_quadratic_program_or_die(path + ".bz2", include_names=True)
    np.testing.assert_array_equal(got.constraint_matrix.toarray(), lp.constraint_matrix.toarray())
    np.testing.assert_array_equal(got.objective_vector, lp.objective_vector)
    assert got.variable_names == lp.variable_names
    cut = raw.index(b"COLUMNS")                                                   # two streams back to back decompress to their concatenation
    open(str(tmp_path / "two.mps.bz2"), "wb").write(bz2.compress(raw[:cut]) + bz2.compress(raw[cut:]))
    np.testing.assert_array_equal(pdlp.read_quadratic_program_or_die(str(tmp_path / "two.mps.bz2")).objective_vector, lp.objective_vector)
    open(str(tmp_path / "cut.mps.bz2"), "wb").write(bz2.compress(raw * 50)[:-20])  # a stream cut short is an error, not a shorter model
    with pytest.raises(ValueError, match="bzip2"):
        pdlp.read_quadratic_program_or_die(str(tmp_path / "cut.mps.bz2"))
    open(str(tmp_path / "bad.mps.bz2"), "wb").write(raw)                          # not bzip2 at all
    with pytest.raises(ValueError, match="bzip2"):
        pdlp.read_quadratic_program_or_die(str(tmp_path / "bad.mps.bz2"))
    with pytest.raises(ValueError, match="Invalid filename suffix"):
        pdlp.read_quadratic_program_or_die(str(tmp_path / "lp.txt"))


def test_get_entries_of_iteration_stats_by_point_type():  # iteration_stats_test.cc:651-732
    from google.protobuf import text_format
    stats = text_format.Parse("""
        convergence_information { candidate_type: POINT_TYPE_CURRENT_ITERATE primal_objective: 1.0 }
        convergence_information { candidate_type: POINT_TYPE_AVERAGE_ITERATE primal_objective: 2.0 }
        infeasibility_information { candidate_type: POINT_TYPE_CURRENT_ITERATE primal_ray_linear_objective: 1.0 }
        infeasibility_information { candidate_type: POINT_TYPE_AVERAGE_ITERATE primal_ray_linear_objective: 2.0 }
        point_metadata { point_type: POINT_TYPE_CURRENT_ITERATE active_primal_variable_count: 1 }
        point_metadata { point_type: POINT_TYPE_AVERAGE_ITERATE active_primal_variable_count: 2 }""", pdlp_proto.IterationStatsProto())
    PT = pdlp.PointType
    assert pdlp.get_convergence_information(stats, PT.POINT_TYPE_AVERAGE_ITERATE).primal_objective == 2.0
    assert pdlp.get_convergence_information(stats, PT.POINT_TYPE_CURRENT_ITERATE).primal_objective == 1.0
    assert pdlp.get_convergence_information(stats, PT.POINT_TYPE_ITERATE_DIFFERENCE) is None
    assert pdlp.get_infeasibility_information(stats, PT.POINT_TYPE_AVERAGE_ITERATE).primal_ray_linear_objective == 2.0
    assert pdlp.get_infeasibility_information(stats, PT.POINT_TYPE_CURRENT_ITERATE).primal_ray_linear_objective == 1.0
    assert pdlp.get_infeasibility_information(stats, PT.POINT_TYPE_ITERATE_DIFFERENCE) is None
    assert pdlp.get_point_metadata(stats, PT.POINT_TYPE_AVERAGE_ITERATE).active_primal_variable_count == 2
    assert pdlp.get_point_metadata(stats, PT.POINT_TYPE_CURRENT_ITERATE).active_primal_variable_count == 1
    assert pdlp.get_point_metadata(stats, PT.POINT_TYPE_ITERATE_DIFFERENCE) is None
    assert pdlp.get_convergence_information(None, PT.POINT_TYPE_CURRENT_ITERATE) is None


def test_quadratic_program_to_string():  # quadratic_program_test.cc:477-556
    import fixtures
    lp_body = ("c0: 12 <= + 2 x0 + 1 x1 + 1 x2 + 2 x3 <= 12\n"
               "c1: + 1 x0 + 1 x2 <= 7\n"
               "c2: -4 <= + 4 x0\n"
               "c3: -1 <= + 1.5 x2 + -1 x3 <= 1\n"
               "Bounds\n"
               "x0 free\n"
               "x1 >= -2\n"
               "x2 <= 6\n"
               "2.5 <= x3 <= 3.5\n")
    qp = fixtures.test_lp()
    assert pdlp.to_string(qp) == "minimize 1 * (-14 + 5.5 x0 + -2 x1 + -1 x2 + 1 x3)\n" + lp_body
    qp.objective_scaling_factor = -1
    assert pdlp.to_string(qp) == "maximize -1 * (-14 + 5.5 x0 + -2 x1 + -1 x2 + 1 x3)\n" + lp_body
    cut = pdlp.to_string(fixtures.test_lp(), 100)
    assert len(cut) == 100 and cut == ("minimize 1 * (-14 + 5.5 x0 + -2 x1 + -1 x2 + 1 x3)\n"
                                       "c0: 12 <= + 2 x0 + 1 x1 + 1 x2 + 2 x3 <= 12\n"
                                       "c...\n")
    qp = fixtures.test_diagonal_qp1()
    assert pdlp.to_string(qp) == ("minimize 1 * (5 + -1 x0 + -1 x1 + 1/2 * ( + 4 x0^2 + 1 x1^2))\n"
                                  "c0: + 1 x0 + 1 x1 <= 1\n"
                                  "Bounds\n"
                                  "1 <= x0 <= 2\n"
                                  "-2 <= x1 <= 4\n")
    qp.problem_name, qp.variable_names, qp.constraint_names = "test", ["x", "y"], ["total"]
    assert pdlp.to_string(qp) == ("test:\n"
                                  "minimize 1 * (5 + -1 x + -1 y + 1/2 * ( + 4 x^2 + 1 y^2))\n"
                                  "total: + 1 x + 1 y <= 1\n"
                                  "Bounds\n"
                                  "1 <= x <= 2\n"
                                  "-2 <= y <= 4\n")
    qp = fixtures.test_lp()
    qp.variable_lower_bounds = qp.variable_lower_bounds[:3]
    assert pdlp.to_string(qp).startswith("Quadratic program with inconsistent dimensions: ")


@pytest.mark.parametrize("field,size", [("constraint_lower_bounds", 10), ("constraint_upper_bounds", 10), ("objective_vector", 10),
                                        ("variable_lower_bounds", 10), ("variable_upper_bounds", 10), ("objective_matrix", 10),
                                        ("variable_names", 1), ("constraint_names", 1), ("constraint_matrix", (10, 2)), ("constraint_matrix", (3, 10))])
def test_validate_quadratic_program_dimensions_inconsistent(field, size):  # quadratic_program_test.cc:83-162
    qp = pdlp.QuadraticProgram(2, 3)
    pdlp.validate_quadratic_program_dimensions(qp)
    if field == "constraint_matrix":
        qp.constraint_matrix = scipy.sparse.csc_matrix(size)
    elif field.endswith("_names"):
        setattr(qp, field, ["n"] * size)
    else:
        setattr(qp, field, np.zeros(size))
    with pytest.raises(ValueError, match="Inconsistent dimensions"):
        pdlp.validate_quadratic_program_dimensions(qp)
    assert pdlp.to_string(qp).startswith("Quadratic program with inconsistent dimensions: ")


def _run_python_example():
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return subprocess.run([sys.executable, os.path.join(root, "examples", "solve_simple_lp.py")], capture_output=True, text=True, timeout=300)


def test_python_example_fails_loudly_without_a_gpu():
    if pdlp.backend().device_count() > 0:
        pytest.skip("a CUDA device is present")
    p = _run_python_example()
    assert p.returncode != 0 and "no usable CUDA device" in p.stderr


@pytest.mark.gpu
def test_python_example_solves_the_sample_lp(b200_backend):  # samples/simple_pdlp_program.py
    p = _run_python_example()
    assert p.returncode == 0, p.stdout + p.stderr
    assert "Solve successful" in p.stdout
    objective = [line for line in p.stdout.splitlines() if line.startswith("Primal objective:")]
    assert objective and float(objective[0].split(":")[1]) == pytest.approx(-34.0, abs=1e-4)
