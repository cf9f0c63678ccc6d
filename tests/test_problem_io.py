"""MPModelProto <-> QuadraticProgram, the MPSolver-style proto solver, MPS reading and the
command-line front end. Known answers transcribed from
``ortools/pdlp/quadratic_program_test.cc`` (TestLpProto / TestQpProto, :166-222, :262-283) and
``ortools/lp_data/mps_reader_template.h`` (the worked example :38-75 and the RANGES / BOUNDS
tables :185-240). The solves on CPU go through the oracle (checker only); the GPU tests run
the same through the CUDA library."""
import json
import os

import numpy as np
import pytest
from google.protobuf import text_format

import fixtures
from ortools_b200 import mp_model, pdlp, pdlp_proto, pdlp_solve, qp_io

INF = float("inf")
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "pdlp_proto_tags.json")

TEST_LP_PROTO = """
  variable { lower_bound: -inf upper_bound: inf objective_coefficient: 5.5 }
  variable { lower_bound: -2 upper_bound: inf objective_coefficient: -2 }
  variable { lower_bound: -inf upper_bound: 6 objective_coefficient: -1 }
  variable { lower_bound: 2.5 upper_bound: 3.5 objective_coefficient: 1 }
  constraint { lower_bound: 12 upper_bound: 12 var_index: [0, 1, 2, 3] coefficient: [2, 1, 1, 2] }
  constraint { lower_bound: -inf upper_bound: 7 var_index: [0, 2] coefficient: [1, 1] }
  constraint { lower_bound: -4 upper_bound: inf var_index: [0] coefficient: [4] }
  constraint { lower_bound: -1 upper_bound: 1 var_index: [2, 3] coefficient: [1.5, -1] }
  objective_offset: -14
"""
TEST_QP_PROTO = """
  variable { lower_bound: -1 upper_bound: 2 objective_coefficient: 3 }
  variable { lower_bound: -2 upper_bound: 3 objective_coefficient: 0 }
  constraint { lower_bound: -inf upper_bound: 42 var_index: [0, 1] coefficient: [1, 1] }
  objective_offset: -4
  quadratic_objective { qvar1_index: [0, 1] qvar2_index: [0, 1] coefficient: [1, 1] }
"""


def _proto(text, maximize=False):
    p = text_format.Parse(text, mp_model.MPModelProto())
    p.maximize = maximize
    return p


def test_linear_solver_schema_subset_matches_the_reference_tags():
    golden = json.load(open(GOLDEN))
    for name, fields in golden["linear_solver_messages"].items():
        d = mp_model._pool.FindMessageTypeByName("operations_research." + name)
        for f in d.fields:
            assert f.name in fields, (name, f.name)
            assert f.number == fields[f.name][0], (name, f.name)
            assert pdlp_proto._is_repeated(f) == (fields[f.name][1] == "repeated"), (name, f.name)
    st = mp_model._pool.FindEnumTypeByName("operations_research.MPSolverResponseStatus")
    assert {v.name: v.number for v in st.values} == golden["linear_solver_enums"]["MPSolverResponseStatus"]
    ty = mp_model._pool.FindEnumTypeByName("operations_research.MPModelRequest.SolverType")
    assert {v.name: v.number for v in ty.values} == golden["linear_solver_enums"]["MPModelRequest.SolverType"]


@pytest.mark.parametrize("maximize", [False, True])
def test_lp_from_mp_model_proto(maximize):  # quadratic_program_test.cc:227-234 + VerifyTestLp
    lp = mp_model.qp_from_mp_model_proto(_proto(TEST_LP_PROTO, maximize), relax_integer_variables=False)
    sign = -1.0 if maximize else 1.0
    assert sign * lp.objective_offset == -14
    assert lp.objective_scaling_factor == sign
    np.testing.assert_array_equal(sign * lp.objective_vector, [5.5, -2, -1, 1])
    np.testing.assert_array_equal(lp.constraint_lower_bounds, [12, -INF, -4, -1])
    np.testing.assert_array_equal(lp.constraint_upper_bounds, [12, 7, INF, 1])
    np.testing.assert_array_equal(lp.variable_lower_bounds, [-INF, -2, -INF, 2.5])
    np.testing.assert_array_equal(lp.variable_upper_bounds, [INF, INF, 6, 3.5])
    np.testing.assert_array_equal(lp.constraint_matrix.toarray(), [[2, 1, 1, 2], [1, 0, 1, 0], [4, 0, 0, 0], [0, 0, 1.5, -1]])
    assert lp.objective_matrix is None


@pytest.mark.parametrize("maximize", [False, True])
def test_lp_and_qp_round_trip(maximize):  # quadratic_program_test.cc:248-255, 350-364
    for text in (TEST_LP_PROTO, TEST_QP_PROTO):
        proto = _proto(text, maximize)
        qp = mp_model.qp_from_mp_model_proto(proto, relax_integer_variables=False)
        assert mp_model.qp_to_mp_model_proto(qp) == proto


@pytest.mark.parametrize("maximize", [False, True])
def test_qp_from_mp_model_proto(maximize):  # quadratic_program_test.cc:284-303
    qp = mp_model.qp_from_mp_model_proto(_proto(TEST_QP_PROTO, maximize), relax_integer_variables=False)
    sign = -1.0 if maximize else 1.0
    np.testing.assert_array_equal(qp.constraint_lower_bounds, [-INF])
    np.testing.assert_array_equal(qp.constraint_upper_bounds, [42])
    np.testing.assert_array_equal(qp.variable_lower_bounds, [-1, -2])
    np.testing.assert_array_equal(qp.variable_upper_bounds, [2, 3])
    assert sign * qp.objective_offset == -4 and qp.objective_scaling_factor == sign
    np.testing.assert_array_equal(sign * qp.objective_vector, [3, 0])
    np.testing.assert_array_equal(sign * qp.objective_matrix, [2, 2])


def test_conversion_errors_and_names():  # quadratic_program_test.cc:305-326, 366-464
    off_diagonal = _proto(TEST_QP_PROTO)
    del off_diagonal.quadratic_objective.qvar1_index[:], off_diagonal.quadratic_objective.qvar2_index[:], off_diagonal.quadratic_objective.coefficient[:]
    off_diagonal.quadratic_objective.qvar1_index.append(0)
    off_diagonal.quadratic_objective.qvar2_index.append(1)
    off_diagonal.quadratic_objective.coefficient.append(1)
    with pytest.raises(mp_model.InvalidArgument):
        mp_model.qp_from_mp_model_proto(off_diagonal, False)
    integer = text_format.Parse("""variable { lower_bound: -1 upper_bound: 2 objective_coefficient: 1 }
        variable { lower_bound: -2 upper_bound: 3 objective_coefficient: 2 is_integer: true }
        constraint { lower_bound: -inf upper_bound: 1 var_index: [0, 1] coefficient: [1, 1] }""", mp_model.MPModelProto())
    with pytest.raises(mp_model.InvalidArgument):
        mp_model.qp_from_mp_model_proto(integer, relax_integer_variables=False)
    lp = mp_model.qp_from_mp_model_proto(integer, relax_integer_variables=True)
    np.testing.assert_array_equal(lp.objective_vector, [1, 2])
    assert lp.objective_matrix is None and lp.objective_offset == 0
    empty = mp_model.qp_from_mp_model_proto(mp_model.MPModelProto(), False)
    assert empty.constraint_matrix.shape == (0, 0) and empty.objective_scaling_factor == 1 and empty.objective_matrix is None
    named = text_format.Parse("""name: "problem"
        variable { name: "x_0" lower_bound: -1 upper_bound: 2 objective_coefficient: 1 }
        variable { name: "x_1" lower_bound: -2 upper_bound: 3 objective_coefficient: 2 }
        constraint { name: "c_0" lower_bound: -inf upper_bound: 1 var_index: [0, 1] coefficient: [1, 1] }""", mp_model.MPModelProto())
    without = mp_model.qp_from_mp_model_proto(named, True, include_names=False)
    assert without.problem_name is None and without.variable_names is None and without.constraint_names is None
    with_names = mp_model.qp_from_mp_model_proto(named, True, include_names=True)
    assert with_names.problem_name == "problem" and with_names.variable_names == ["x_0", "x_1"] and with_names.constraint_names == ["c_0"]
    with pytest.raises(mp_model.InvalidArgument):  # quadratic_program_test.cc:336-348
        mp_model.can_fit_in_mp_model_proto(lp, largest_ok_size=1)


# the worked example of mps_reader_template.h:38-75
EXAMPLE_MPS = """NAME          TESTEQ
ROWS
 N  COST
 L  LIM1
 G  LIM2
 E  MYEQN
COLUMNS
    XONE      COST         1   LIM1         1
    XONE      LIM2         1
    YTWO      COST         2   LIM1         1
    YTWO      MYEQN       -1
    ZTHREE    COST         3   MYEQN        1
RHS
    RHS1      COST       -10
    RHS1      LIM1         4   LIM2         1
    RHS1      MYEQN        7
BOUNDS
 UP BND1      XONE         4
 LO BND1      YTWO        -1
 UP BND1      YTWO         1
ENDATA
"""


def test_mps_example_is_read_like_the_reference_reader():
    qp = qp_io.parse_mps(EXAMPLE_MPS.splitlines(), include_names=True)
    assert qp.problem_name == "TESTEQ"
    assert qp.variable_names == ["XONE", "YTWO", "ZTHREE"] and qp.constraint_names == ["LIM1", "LIM2", "MYEQN"]
    np.testing.assert_array_equal(qp.objective_vector, [1, 2, 3])
    assert qp.objective_offset == 10          # minus the RHS of the objective row
    np.testing.assert_array_equal(qp.constraint_matrix.toarray(), [[1, 1, 0], [1, 0, 0], [0, -1, 1]])
    np.testing.assert_array_equal(qp.constraint_lower_bounds, [-INF, 1, 7])
    np.testing.assert_array_equal(qp.constraint_upper_bounds, [4, INF, 7])
    np.testing.assert_array_equal(qp.variable_lower_bounds, [0, -1, 0])
    np.testing.assert_array_equal(qp.variable_upper_bounds, [4, 1, INF])


def test_mps_ranges_bounds_markers_and_objsense():
    text = """NAME RANGED
OBJSENSE
    MAX
ROWS
 N OBJ
 G G1
 L L1
 E E1
 E E2
 N IGNORED
COLUMNS
    MARKER 'MARKER' 'INTORG'
    I1 OBJ 1 G1 1
    I2 OBJ 1 L1 1
    MARKER 'MARKER' 'INTEND'
    X OBJ 2 E1 1
    X E2 1
    Y E1 1
    Y IGNORED 5
    F G1 1
    M L1 1
RHS
    RHS G1 1 L1 10
    RHS E1 5 E2 5
RANGES
    RNG G1 -3 L1 4
    RNG E1 2 E2 -2
BOUNDS
 UP BND I2 7
 FR BND F
 MI BND M
 FX BND Y 3
 BV BND X
ENDATA
"""
    qp = qp_io.parse_mps(text.splitlines(), include_names=True)
    assert qp.variable_names == ["I1", "I2", "X", "Y", "F", "M"]
    # RANGES table of mps_reader_template.h:195-202
    np.testing.assert_array_equal(qp.constraint_lower_bounds, [1, 6, 5, 3])
    np.testing.assert_array_equal(qp.constraint_upper_bounds, [4, 10, 7, 5])
    # integer-by-marker without bounds: [0, 1]; with an explicit bound: default [0, inf) then the bound
    np.testing.assert_array_equal(qp.variable_lower_bounds, [0, 0, 0, 3, -INF, -INF])
    np.testing.assert_array_equal(qp.variable_upper_bounds, [1, 7, 1, 3, INF, INF])
    # maximisation becomes minimisation of the negated objective
    assert qp.objective_scaling_factor == -1
    np.testing.assert_array_equal(qp.objective_vector, [-1, -1, -2, 0, 0, 0])


def test_mps_write_read_round_trip_and_suffix_dispatch(tmp_path):
    lp = fixtures.test_lp()
    lp.problem_name = "test_lp"
    path = str(tmp_path / "lp.mps")
    qp_io.write_linear_program_to_mps(lp, path)
    back = qp_io.read_quadratic_program(path)
    np.testing.assert_array_equal(back.constraint_matrix.toarray(), lp.constraint_matrix.toarray())
    for f in ("objective_vector", "constraint_lower_bounds", "constraint_upper_bounds", "variable_lower_bounds", "variable_upper_bounds"):
        np.testing.assert_array_equal(getattr(back, f), getattr(lp, f))
    assert back.objective_offset == lp.objective_offset
    import gzip
    gz = str(tmp_path / "lp.mps.gz")
    with gzip.open(gz, "wb") as f:
        f.write(open(path, "rb").read())
    np.testing.assert_array_equal(qp_io.read_quadratic_program(gz).objective_vector, lp.objective_vector)
    pb = str(tmp_path / "qp.pb")
    qp_io.write_quadratic_program_to_mp_model_proto(fixtures.test_diagonal_qp1(), pb)
    q = qp_io.read_quadratic_program(pb)
    np.testing.assert_array_equal(q.objective_matrix, fixtures.test_diagonal_qp1().objective_matrix)
    tp = str(tmp_path / "lp.textproto")
    open(tp, "w").write(TEST_LP_PROTO)
    assert qp_io.read_quadratic_program(tp).constraint_matrix.shape == (4, 4)
    with pytest.raises(ValueError):
        qp_io.read_quadratic_program(str(tmp_path / "lp.txt"))
    with pytest.raises(ValueError):
        qp_io.write_linear_program_to_mps(fixtures.test_diagonal_qp1(), str(tmp_path / "qp.mps"))


def _oracle():
    from oracle import pdlp_oracle
    return pdlp_oracle.backend()


def _request(maximize):
    req = mp_model.MPModelRequestProto()
    req.model.CopyFrom(mp_model.qp_to_mp_model_proto(fixtures.tiny_lp()))
    if maximize:  # same optimum: maximise the negated objective
        req.model.maximize = True
        req.model.objective_offset = -req.model.objective_offset
        for v in req.model.variable:
            v.objective_coefficient = -v.objective_coefficient
    req.solver_type = mp_model.SolverType.PDLP_LINEAR_PROGRAMMING
    req.solver_specific_parameters = "termination_criteria { simple_optimality_criteria { eps_optimal_absolute: 1e-8 eps_optimal_relative: 1e-8 } }"
    return req


def _check_tiny_lp_response(resp, maximize):
    # TinyLp optimum (test_util.h): x = [1, 0, 6, 2], y = [0.5, 4, 0], reduced costs [0, 1.5, -3.5, 0], objective -1
    sign = -1.0 if maximize else 1.0
    assert resp.status == mp_model.MPSolverResponseStatus.MPSOLVER_OPTIMAL
    assert resp.objective_value == pytest.approx(sign * -1.0, abs=1e-6)
    np.testing.assert_allclose(resp.variable_value, [1, 0, 6, 2], atol=1e-5)
    np.testing.assert_allclose(resp.dual_value, sign * np.array([0.5, 4, 0]), atol=1e-5)   # dual sign flip, pdlp_proto_solver.cc:113-121
    np.testing.assert_allclose(resp.reduced_cost, sign * np.array([0, 1.5, -3.5, 0]), atol=1e-5)
    log = pdlp_proto.SolveLogProto()
    log.ParseFromString(resp.solver_specific_info)
    assert log.termination_reason == pdlp.TerminationReason.TERMINATION_REASON_OPTIMAL


@pytest.mark.parametrize("maximize", [False, True])
def test_proto_solver_maps_status_values_and_dual_signs(maximize):
    _check_tiny_lp_response(mp_model.pdlp_solve_proto(_request(maximize), backend=_oracle()), maximize)


def test_proto_solver_rejects_bad_parameters_and_applies_the_time_limit():
    req = _request(False)
    req.solver_specific_parameters = "no_such_field: 1"
    assert mp_model.pdlp_solve_proto(req, backend=_oracle()).status == mp_model.MPSolverResponseStatus.MPSOLVER_MODEL_INVALID_SOLVER_PARAMETERS
    req = _request(False)
    req.solver_time_limit_seconds = 0.0
    resp = mp_model.pdlp_solve_proto(req, backend=_oracle())
    assert resp.status == mp_model.MPSolverResponseStatus.MPSOLVER_NOT_SOLVED
    log = pdlp_proto.SolveLogProto()
    log.ParseFromString(resp.solver_specific_info)
    assert log.termination_reason == pdlp.TerminationReason.TERMINATION_REASON_TIME_LIMIT
    infeasible = mp_model.MPModelRequestProto()
    infeasible.model.CopyFrom(mp_model.qp_to_mp_model_proto(fixtures.small_primal_infeasible_lp()))
    assert mp_model.pdlp_solve_proto(infeasible, backend=_oracle()).status == mp_model.MPSolverResponseStatus.MPSOLVER_INFEASIBLE


def _run_cli(tmp_path, backend):
    mps = str(tmp_path / "tiny.mps")
    lp = fixtures.tiny_lp()
    lp.variable_names = ["a", "b", "c", "d"]
    qp_io.write_linear_program_to_mps(lp, mps)
    log_file, sol_file = str(tmp_path / "log.textproto"), str(tmp_path / "tiny.sol")
    messages = open(str(tmp_path / "messages.txt"), "w")
    result = pdlp_solve.solve(mps, "termination_criteria { simple_optimality_criteria { eps_optimal_absolute: 1e-8 eps_optimal_relative: 1e-8 } }",
                              log_file, sol_file, backend=backend, out=messages)
    messages.close()
    assert result.solve_log.termination_reason == pdlp.TerminationReason.TERMINATION_REASON_OPTIMAL
    log = text_format.Parse(open(log_file).read(), pdlp_proto.SolveLogProto())
    assert log.termination_reason == pdlp.TerminationReason.TERMINATION_REASON_OPTIMAL and log.params.verbosity_level == 2
    sol = open(sol_file).read().splitlines()
    assert sol[0].startswith("=obj= ") and float(sol[0].split()[1]) == pytest.approx(-1.0, abs=1e-6)
    assert [l.split()[0] for l in sol[1:]] == ["a", "b", "c", "d"]
    np.testing.assert_allclose([float(l.split()[1]) for l in sol[1:]], [1, 0, 6, 2], atol=1e-5)
    assert os.path.getsize(str(tmp_path / "messages.txt")) > 0   # verbosity 2 prints iteration statistics


def test_command_line_solve_writes_log_and_sol(tmp_path):
    _run_cli(tmp_path, _oracle())


def test_command_line_solve_is_interruptible_without_callbacks(tmp_path):
    """^C during a solve that prints nothing (verbosity 0: no callback ever hands control back to the
    interpreter) still reaches interrupt_solve, like the reference binary's signal handler
    (pdlp_solve.cc:100-108): the solve runs on a worker thread while the main thread takes signals."""
    import signal
    import threading
    import time
    import scipy.sparse as sps
    rng = np.random.default_rng(5)
    m, n = 400, 800
    lp = pdlp.QuadraticProgram(n, m)
    lp.constraint_matrix = sps.random(m, n, density=0.02, random_state=7, format="csc", data_rvs=lambda k: rng.normal(size=k))
    lp.objective_vector = rng.normal(size=n)
    lp.constraint_lower_bounds, lp.constraint_upper_bounds = np.full(m, -1.0), np.full(m, 1.0)
    lp.variable_lower_bounds, lp.variable_upper_bounds = np.full(n, -1.0), np.full(n, 1.0)
    mps = str(tmp_path / "long.mps")
    qp_io.write_linear_program_to_mps(lp, mps)
    params = ("verbosity_level: 0 termination_criteria { iteration_limit: 2000000000 simple_optimality_criteria "
              "{ eps_optimal_absolute: 0 eps_optimal_relative: 0 } }")
    timer = threading.Timer(1.0, lambda: os.kill(os.getpid(), signal.SIGINT))
    original = signal.getsignal(signal.SIGINT)
    early = []
    before = lambda *_: early.append(1)   # (a signal that beat solve()'s own handler must not abort the test session)
    signal.signal(signal.SIGINT, before)
    start = time.time()
    timer.start()
    try:
        result = pdlp_solve.solve(mps, params, backend=_oracle(), out=open(os.devnull, "w"))
        assert signal.getsignal(signal.SIGINT) is before   # the previous handler is back
    finally:
        timer.cancel()
        signal.signal(signal.SIGINT, original)
    assert not early
    assert result.solve_log.termination_reason == pdlp.TerminationReason.TERMINATION_REASON_INTERRUPTED_BY_USER
    assert time.time() - start < 20.0 and result.solve_log.iteration_count > 0


@pytest.mark.gpu
@pytest.mark.parametrize("maximize", [False, True])
def test_proto_solver_on_the_gpu(b200_backend, maximize):
    _check_tiny_lp_response(mp_model.pdlp_solve_proto(_request(maximize)), maximize)


@pytest.mark.gpu
def test_command_line_solve_on_the_gpu(tmp_path, b200_backend):
    _run_cli(tmp_path, None)
