"""The native format layer (include/pdlp_b200_io.h; csrc/proto_codec.cc, csrc/formats.cc): the proto2
wire / text / JSON codec, parameter and SolveLog adapters, MPModelProto and MPS conversions and
PdlpSolveProto in C++ -- checked against the ``google.protobuf`` runtime (whose schemas are pinned to
the reference's tags by tests/golden/pdlp_proto_tags.json) and against the Python adapters, which
carry the reference's known answers (test_proto_io.py, test_problem_io.py). Host-only except the
two GPU tests at the end."""
import ctypes as C
import gzip
import math

import numpy as np
import pytest
from google.protobuf import json_format, text_format

import fixtures
from ortools_b200 import _capi as capi
from ortools_b200 import mp_model, native_io, pdlp, pdlp_proto, qp_io
from test_problem_io import EXAMPLE_MPS, TEST_LP_PROTO, TEST_QP_PROTO, _check_tiny_lp_response, _request

INF = float("inf")


def pod_fields(pod, prefix=""):
    out = {}
    for name, _ in pod._fields_:
        v = getattr(pod, name)
        if isinstance(v, C.Structure):
            out.update(pod_fields(v, prefix + name + "."))
        elif isinstance(v, C.Array):
            out[prefix + name] = list(v)
        else:
            out[prefix + name] = v
    return out


def same_pod(a, b):
    fa, fb = pod_fields(a), pod_fields(b)
    n = fa["num_random_projection_seeds"]
    for d in (fa, fb):  # entries past the count are not meaningful
        d["random_projection_seeds"] = d["random_projection_seeds"][: min(n, capi.MAX_SEEDS)]
    assert fa.keys() == fb.keys()
    for k in fa:
        x, y = fa[k], fb[k]
        if isinstance(x, float) and math.isnan(x):
            assert math.isnan(y), k
        else:
            assert x == y, (k, x, y)


PARAM_TEXTS = [
    "",
    "termination_criteria { simple_optimality_criteria { eps_optimal_absolute: 1e-4 eps_optimal_relative: 1.0e-4 } }",
    "termination_criteria { detailed_optimality_criteria { eps_optimal_primal_residual_absolute: 1e-3 eps_optimal_objective_gap_relative: 0 } }",
    "termination_criteria { simple_optimality_criteria { } }",
    "termination_criteria { eps_optimal_absolute: 1.0e-6 eps_optimal_relative: nan optimality_norm: OPTIMALITY_NORM_L_INF_COMPONENTWISE }",
    "termination_criteria { time_sec_limit: 12.5 iteration_limit: 1000 kkt_matrix_pass_limit: inf eps_primal_infeasible: -inf eps_dual_infeasible: 1e-300 }",
    "termination_criteria: { iteration_limit: 0x10 }  # hexadecimal, colon before a message",
    "termination_criteria < iteration_limit: 7 >; num_threads: 4, num_shards: 16",
    "num_threads: 8 num_shards: 32 scheduler_type: SCHEDULER_TYPE_EIGEN_THREADPOOL record_iteration_stats: true verbosity_level: 4 log_interval_seconds: 1.5",
    "major_iteration_frequency: 128 termination_check_frequency: 32 restart_strategy: ADAPTIVE_DISTANCE_BASED primal_weight_update_smoothing: 0.25",
    "initial_primal_weight: 2.5 l_inf_ruiz_iterations: 0 l2_norm_rescaling: false sufficient_reduction_for_restart: 0.2 necessary_reduction_for_restart: 0.8",
    "initial_primal_weight: 0",
    "linesearch_rule: MALITSKY_POCK_LINESEARCH_RULE malitsky_pock_parameters { step_size_downscaling_factor: 0.5 linesearch_contraction_factor: 0.9 step_size_interpolation: 2 }",
    "linesearch_rule: 3 adaptive_linesearch_parameters { step_size_reduction_exponent: 0.5 step_size_growth_exponent: 0.75 } initial_step_size_scaling: 4",
    "random_projection_seeds: 1 random_projection_seeds: 2 random_projection_seeds: -3",
    "random_projection_seeds: [4, 5, 6]",
    "infinite_constraint_bound_threshold: 1e20 handle_some_primal_gradients_on_finite_bounds_as_residuals: false",
    "use_diagonal_qp_trust_region_solver: true diagonal_qp_trust_region_solver_tolerance: 1e-6",
    "use_feasibility_polishing: true apply_feasibility_polishing_after_limits_reached: true apply_feasibility_polishing_if_solver_is_interrupted: true",
    "presolve_options { use_glop: true }",
    "restart_strategy: NO_RESTARTS\nprimal_weight_update_smoothing: 0.0\n# a comment line\nrecord_iteration_stats: True",
]


@pytest.mark.parametrize("text", PARAM_TEXTS)
def test_params_text_matches_the_protobuf_runtime(text):
    want = pdlp_proto.params_from_text(text).to_pod()
    same_pod(native_io.params_from_text(text), want)
    # ... and so does the binary encoding of the same message
    blob = text_format.Parse(text, pdlp_proto.PrimalDualHybridGradientParamsProto()).SerializeToString()
    same_pod(native_io.params_from_bytes(blob), want)


@pytest.mark.parametrize("text", PARAM_TEXTS)
def test_params_round_trip_through_every_encoding(text):
    pod = pdlp_proto.params_from_text(text).to_pod()
    blob = native_io.params_serialize(pod, native_io.BINARY)
    msg = pdlp_proto.PrimalDualHybridGradientParamsProto()
    msg.ParseFromString(blob)                                   # the runtime accepts our bytes
    same_pod(pdlp_proto.params_from_proto(msg).to_pod(), pod)
    same_pod(native_io.params_from_bytes(blob), pod)
    as_text = native_io.params_serialize(pod, native_io.TEXT).decode()
    assert text_format.Parse(as_text, pdlp_proto.PrimalDualHybridGradientParamsProto()) == msg or "nan" in as_text
    same_pod(native_io.params_from_text(as_text), pod)
    as_json = native_io.params_serialize(pod, native_io.JSON).decode()
    assert json_format.Parse(as_json, pdlp_proto.PrimalDualHybridGradientParamsProto()) == msg or "NaN" in as_json
    assert native_io.convert("PrimalDualHybridGradientParams", as_json, native_io.JSON, native_io.BINARY) == blob
    assert native_io.convert("PrimalDualHybridGradientParams", as_text, native_io.TEXT, native_io.BINARY) == blob


def test_params_serialization_is_canonical():
    """Fields come out in tag order with packed repeated fields, i.e. byte for byte what the protobuf
    runtime serialises for the same message."""
    for text in PARAM_TEXTS:
        pod = pdlp_proto.params_from_text(text).to_pod()
        blob = native_io.params_serialize(pod)
        msg = pdlp_proto.PrimalDualHybridGradientParamsProto()
        msg.ParseFromString(blob)
        assert blob == msg.SerializeToString(), text


def test_params_merge_and_unknown_fields():
    base = native_io.params_from_text("verbosity_level: 2 random_projection_seeds: 1")
    merged = native_io.params_from_text("termination_criteria { iteration_limit: 5 } random_projection_seeds: 2", onto=base)
    assert merged.verbosity_level == 2 and merged.termination_criteria.iteration_limit == 5
    assert merged.num_random_projection_seeds == 2 and list(merged.random_projection_seeds)[:2] == [1, 2]   # repeated fields append
    # glop_parameters (tag 2 of PresolveOptions, a message this path never reads) and a field
    # number nobody defines are skipped like protobuf skips unknown fields
    inner = b"\x08\x01" + b"\x12\x02\x08\x01"          # use_glop: true, glop_parameters { <tag 1>: 1 }
    blob = b"\x82\x01" + bytes([len(inner)]) + inner + b"\xf8\x07\x2a"   # presolve_options (16), unknown varint field 127
    got = native_io.params_from_bytes(blob)
    assert got.presolve_use_glop == 1 and got.num_threads == 1
    with pytest.raises(native_io.NativeIoError):
        native_io.params_from_bytes(b"\x0a\x05\x01")   # truncated length-delimited field
    # merging the other member of a oneof replaces the one that was set (and its values)
    simple = text_format.Parse("termination_criteria { simple_optimality_criteria { eps_optimal_absolute: 1 } }", pdlp_proto.PrimalDualHybridGradientParamsProto())
    detailed = text_format.Parse("termination_criteria { detailed_optimality_criteria { eps_optimal_dual_residual_relative: 2 } }", pdlp_proto.PrimalDualHybridGradientParamsProto())
    for first, second in ((simple, detailed), (detailed, simple)):
        both = pdlp_proto.PrimalDualHybridGradientParamsProto()
        both.ParseFromString(first.SerializeToString() + second.SerializeToString())
        same_pod(native_io.params_from_bytes(first.SerializeToString() + second.SerializeToString()), pdlp_proto.params_from_proto(both).to_pod())
        same_pod(native_io.params_from_bytes(second.SerializeToString(), onto=native_io.params_from_bytes(first.SerializeToString())),
                 pdlp_proto.params_from_proto(both).to_pod())


@pytest.mark.parametrize("text", [
    "no_such_field: 1", "num_threads 4", "num_threads: four", "num_threads: 1.5", "num_threads: 99999999999",
    "restart_strategy: SOMETIMES", "restart_strategy: 17", "termination_criteria { iteration_limit: 5", "termination_criteria: 5",
    "record_iteration_stats: maybe", "log_interval_seconds: abc", "random_projection_seeds: [1, 2", "num_threads: [1]",
    "termination_criteria { simple_optimality_criteria { eps_optimal_primal_residual_absolute: 1 } }",
    "termination_criteria { simple_optimality_criteria { } detailed_optimality_criteria { } }",
    "verbosity_level: 2 verbosity_level: 3",
    "termination_criteria { } termination_criteria { }",
])
def test_params_text_errors_agree_with_the_protobuf_runtime(text):
    with pytest.raises(text_format.ParseError):
        text_format.Parse(text, pdlp_proto.PrimalDualHybridGradientParamsProto())
    with pytest.raises(native_io.NativeIoError):
        native_io.params_from_text(text)


@pytest.mark.parametrize("text", [
    "verbosity_level: 2 verbosity_level: 3",
    "termination_criteria { iteration_limit: 4 } num_threads: 2 termination_criteria { time_sec_limit: 3 }",
    "termination_criteria { simple_optimality_criteria { eps_optimal_absolute: 1 } simple_optimality_criteria { eps_optimal_relative: 3 } }",
    "termination_criteria { simple_optimality_criteria { eps_optimal_absolute: 1 } detailed_optimality_criteria { eps_optimal_dual_residual_relative: 2 } }",
    "termination_criteria { detailed_optimality_criteria { eps_optimal_dual_residual_relative: 2 } simple_optimality_criteria { eps_optimal_relative: 3 } }",
])
def test_params_text_merge_policy(text):
    """TextFormat::Merge (what pdlp_solve.cc:86 and pdlp_proto_solver.cc:47 call): a field may be given
    again and the other member of a oneof replaces the first."""
    preset = pdlp_proto.PrimalDualHybridGradientParamsProto()
    preset.verbosity_level = 2                                   # pdlp_solve.cc:84
    want = pdlp_proto.params_from_proto(text_format.Merge(text, preset)).to_pod()
    same_pod(native_io.params_from_text(text, onto=native_io.params_from_text("verbosity_level: 2")), want)


def test_params_from_text_feeds_the_validator():
    be = pdlp.backend()
    ok, _ = be.validate_params(native_io.params_from_text("termination_criteria { simple_optimality_criteria { eps_optimal_absolute: 1e-4 } }"))
    assert ok
    ok, message = be.validate_params(native_io.params_from_text("termination_criteria { eps_optimal_absolute: 1 simple_optimality_criteria { } }"))
    assert not ok and "simple_optimality_criteria" in message


# --------------------------------------------------------------------------------------------------
# SolveLog
# --------------------------------------------------------------------------------------------------
def _oracle_result(qp, params):
    """A raw PdlpResult of a real solve (made by the CPU checker), serialised by the product's host code."""
    from oracle import pdlp_oracle
    blobs = {}

    def grab(res):
        for name, fmt in (("binary", native_io.BINARY), ("text", native_io.TEXT), ("json", native_io.JSON)):
            blobs[name] = native_io.solve_log_serialize(res, fmt)
    result = pdlp_oracle.backend().primal_dual_hybrid_gradient(qp, params, result_pod_consumer=grab)
    return result, blobs


def _strip_params(msg):
    msg.ClearField("params")
    for d in msg.feasibility_polishing_details:
        d.ClearField("params")
    return msg


@pytest.mark.parametrize("case", ["plain", "polishing", "invalid"])
def test_solve_log_matches_the_python_adapter(case):
    params = pdlp.PrimalDualHybridGradientParams()
    params.record_iteration_stats = True
    params.random_projection_seeds = [1, 2]
    params.major_iteration_frequency = params.termination_check_frequency = 8
    params.termination_criteria.simple_optimality_criteria.eps_optimal_absolute = 1e-7
    params.termination_criteria.simple_optimality_criteria.eps_optimal_relative = 1e-7
    qp = fixtures.test_lp()
    qp.problem_name = "test \"lp\"\n"
    if case == "polishing":  # the set-up of test_feasibility_polishing.py (primal_dual_hybrid_gradient_test.cc:1669-1712)
        from test_feasibility_polishing import polishing_params, primal_lp
        params, qp = polishing_params(), primal_lp()
        params.record_iteration_stats = True
        qp.problem_name = "polish"
    if case == "invalid":
        params.major_iteration_frequency = 0
    result, blobs = _oracle_result(qp, params)
    want = pdlp_proto.solve_log_to_proto(result.solve_log, params)
    got = pdlp_proto.SolveLogProto()
    got.ParseFromString(blobs["binary"])
    # SolveLog.params: the POD keeps values, not presence -- compare as parameter objects
    if case == "invalid":
        assert not got.HasField("params")                        # ErrorSolverResult carries no parameters
    else:
        same_pod(pdlp_proto.params_from_proto(got.params).to_pod(), params.to_pod())
    if case == "polishing":
        assert len(got.feasibility_polishing_details) >= 1
        for d in got.feasibility_polishing_details:
            assert d.HasField("params") and 0 < d.params.termination_criteria.iteration_limit <= 500   # the budget of the phase
    if case == "invalid":
        assert got.termination_reason == pdlp.TerminationReason.TERMINATION_REASON_INVALID_PARAMETER and got.termination_string
    assert _strip_params(got) == _strip_params(want)
    from_text = text_format.Parse(blobs["text"].decode(), pdlp_proto.SolveLogProto())
    from_json = json_format.Parse(blobs["json"].decode(), pdlp_proto.SolveLogProto())
    full = pdlp_proto.SolveLogProto()
    full.ParseFromString(blobs["binary"])
    assert from_text == full and from_json == full
    # byte for byte the runtime's own serialisation of that message
    assert blobs["binary"] == full.SerializeToString()
    # generic conversions: text -> binary -> JSON -> binary
    assert native_io.convert("SolveLog", blobs["text"], native_io.TEXT, native_io.BINARY) == blobs["binary"]
    assert native_io.convert("SolveLog", blobs["json"], native_io.JSON, native_io.BINARY) == blobs["binary"]


def test_write_solve_log_picks_the_format_from_the_suffix(tmp_path):
    from oracle import pdlp_oracle
    params = pdlp.PrimalDualHybridGradientParams()
    paths = {s: str(tmp_path / ("log" + s)) for s in (".textproto", ".pb", ".json", ".txt")}
    errors = []

    def write(res):
        for s, p in paths.items():
            try:
                native_io.write_solve_log(res, p)
            except native_io.NativeIoError as e:
                errors.append((s, str(e)))
    pdlp_oracle.backend().primal_dual_hybrid_gradient(fixtures.tiny_lp(), params, result_pod_consumer=write)
    assert [s for s, _ in errors] == [".txt"] and "Expected .textproto, .pb, or .json" in errors[0][1]   # pdlp_solve.cc:69-72
    a = text_format.Parse(open(paths[".textproto"]).read(), pdlp_proto.SolveLogProto())
    b = pdlp_proto.SolveLogProto()
    b.ParseFromString(open(paths[".pb"], "rb").read())
    c = json_format.Parse(open(paths[".json"]).read(), pdlp_proto.SolveLogProto())
    assert a == b == c and a.termination_reason == pdlp.TerminationReason.TERMINATION_REASON_OPTIMAL


# --------------------------------------------------------------------------------------------------
# MPModelProto <-> QuadraticProgram
# --------------------------------------------------------------------------------------------------
def same_qp(a, b, names=False):
    assert a.constraint_matrix.shape == b.constraint_matrix.shape
    ka, kb = a.constraint_matrix.tocsc(), b.constraint_matrix.tocsc()
    ka.sort_indices(), kb.sort_indices()
    np.testing.assert_array_equal(ka.indptr, kb.indptr)
    np.testing.assert_array_equal(ka.indices, kb.indices)
    np.testing.assert_array_equal(ka.data, kb.data)
    for f in ("objective_vector", "constraint_lower_bounds", "constraint_upper_bounds", "variable_lower_bounds", "variable_upper_bounds"):
        np.testing.assert_array_equal(getattr(a, f), getattr(b, f), err_msg=f)
    assert (a.objective_matrix is None) == (b.objective_matrix is None)
    if a.objective_matrix is not None:
        np.testing.assert_array_equal(a.objective_matrix, b.objective_matrix)
    assert a.objective_offset == b.objective_offset and a.objective_scaling_factor == b.objective_scaling_factor
    if names:
        assert (a.problem_name or "") == (b.problem_name or "")
        assert list(a.variable_names or []) == list(b.variable_names or [])
        assert list(a.constraint_names or []) == list(b.constraint_names or [])


@pytest.mark.parametrize("text", [TEST_LP_PROTO, TEST_QP_PROTO])
@pytest.mark.parametrize("maximize", [False, True])
def test_qp_from_mp_model_proto_bytes(text, maximize):  # quadratic_program_test.cc:227-303
    msg = text_format.Parse(text, mp_model.MPModelProto())
    msg.maximize = maximize
    want = mp_model.qp_from_mp_model_proto(msg, relax_integer_variables=False)
    same_qp(native_io.qp_from_mp_model_proto_bytes(msg.SerializeToString(), False), want)


def test_qp_to_mp_model_proto_bytes_equal_the_runtime_serialisation():  # quadratic_program_test.cc:236-255, 350-364
    for qp in (fixtures.test_lp(), fixtures.tiny_lp(), fixtures.test_diagonal_qp1(), fixtures.correlation_clustering_lp()):
        for scale in (1.0, -1.0):
            qp.objective_scaling_factor = scale
            want = mp_model.qp_to_mp_model_proto(qp)
            got = native_io.qp_to_mp_model_proto_bytes(qp)
            assert got == want.SerializeToString()
            same_qp(native_io.qp_from_mp_model_proto_bytes(got, False), qp)          # round trip
    named = fixtures.tiny_lp()
    named.problem_name, named.variable_names, named.constraint_names = "tiny", ["a", "b", "", "d"], ["r0", "r1", "r2"]
    assert native_io.qp_to_mp_model_proto_bytes(named) == mp_model.qp_to_mp_model_proto(named).SerializeToString()
    back = native_io.qp_from_mp_model_proto_bytes(native_io.qp_to_mp_model_proto_bytes(named), False, include_names=True)
    same_qp(back, named, names=True)


def test_mp_model_conversion_errors_carry_the_reference_messages():  # quadratic_program_test.cc:305-326, 366-464
    def both(msg, relax=False):
        with pytest.raises(mp_model.InvalidArgument) as py:
            mp_model.qp_from_mp_model_proto(msg, relax)
        with pytest.raises(native_io.NativeIoError) as cc:
            native_io.qp_from_mp_model_proto_bytes(msg.SerializeToString(), relax)
        assert str(py.value) == str(cc.value)
    off_diagonal = text_format.Parse(TEST_QP_PROTO, mp_model.MPModelProto())
    off_diagonal.quadratic_objective.qvar1_index.append(0)
    off_diagonal.quadratic_objective.qvar2_index.append(1)
    off_diagonal.quadratic_objective.coefficient.append(1)
    both(off_diagonal)
    integer = text_format.Parse("variable { is_integer: true } ", mp_model.MPModelProto())
    both(integer)
    assert native_io.qp_from_mp_model_proto_bytes(integer.SerializeToString(), True).constraint_matrix.shape == (0, 1)
    general = mp_model.MPModelProto()
    general.general_constraint.add()
    both(general)
    ragged = text_format.Parse("variable { } constraint { var_index: [0] coefficient: [1, 2] }", mp_model.MPModelProto())
    both(ragged)
    out_of_range = text_format.Parse("variable { } constraint { var_index: [1] coefficient: [1] }", mp_model.MPModelProto())
    both(out_of_range)
    bad_q = text_format.Parse("variable { } quadratic_objective { qvar1_index: [0] qvar2_index: [0] }", mp_model.MPModelProto())
    both(bad_q)
    q_range = text_format.Parse("variable { } quadratic_objective { qvar1_index: [3] qvar2_index: [3] coefficient: [1] }", mp_model.MPModelProto())
    both(q_range)
    empty = native_io.qp_from_mp_model_proto_bytes(b"", False)
    assert empty.constraint_matrix.shape == (0, 0) and empty.objective_scaling_factor == 1 and empty.objective_matrix is None
    zero_scale = fixtures.tiny_lp()
    zero_scale.objective_scaling_factor = 0.0
    with pytest.raises(native_io.NativeIoError, match="objective_scaling_factor cannot be zero"):
        native_io.qp_to_mp_model_proto_bytes(zero_scale)


def test_duplicate_entries_are_summed_and_constraints_may_precede_variables():
    msg = text_format.Parse("""variable { } variable { }
        constraint { var_index: [1, 0, 1] coefficient: [1, 2, 3] } constraint { var_index: [0] coefficient: [0] }""", mp_model.MPModelProto())
    want = mp_model.qp_from_mp_model_proto(msg, False)
    same_qp(native_io.qp_from_mp_model_proto_bytes(msg.SerializeToString(), False), want)
    # the wire order of fields is free: constraints first
    parts = [b"\x22" + bytes([len(c.SerializeToString())]) + c.SerializeToString() for c in msg.constraint]
    parts += [b"\x1a" + bytes([len(v.SerializeToString())]) + v.SerializeToString() for v in msg.variable]
    same_qp(native_io.qp_from_mp_model_proto_bytes(b"".join(parts), False), want)


# --------------------------------------------------------------------------------------------------
# MPS and file dispatch
# --------------------------------------------------------------------------------------------------
RANGED_MPS = """NAME RANGED
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
    RHS OBJ 2.5
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
def _fixed(f1="", f2="", f3="", f4="", f5="", f6=""):
    """One fixed-format data line: fields in columns 2-3, 5-12, 15-22, 25-36, 40-47, 50-61."""
    return (" %-2s %-8s  %-8s  %-12s   %-8s  %-12s" % (f1, f2, f3, f4, f5, f6)).rstrip() + "\n"


FIXED_FORMAT_MPS = (                      # names with spaces only parse by column position
    "NAME          FIXED FORMAT\n" "ROWS\n" + _fixed("N", "COST") + _fixed("L", "ROW ONE") + _fixed("G", "LIM2") + "COLUMNS\n"
    + _fixed("", "X 1", "COST", "1.5D0", "ROW ONE", "1.0") + _fixed("", "X 1", "LIM2", "1.0")
    + _fixed("", "Y", "COST", "2.0", "ROW ONE", "1.0e0") + "RHS\n" + _fixed("", "RHS", "ROW ONE", "4.0", "LIM2", "1.0")
    + "BOUNDS\n" + _fixed("UP", "BND", "X 1", "4.0") + "ENDATA\n")


@pytest.mark.parametrize("text", [EXAMPLE_MPS, RANGED_MPS, FIXED_FORMAT_MPS,
                                  "NAME\nROWS\n N obj\nCOLUMNS\nRHS\nENDATA\n",
                                  "* comment\nNAME x\nOBJSENSE MAX\nROWS\n N c\n E r\nCOLUMNS\n a c 1 r 1\n a r 2\nRHS\n r 3\nRANGES\n r 1\nBOUNDS\n LO a -1\n PL a\nENDATA\nignored after ENDATA\n"])
def test_mps_reader_matches_the_python_reader(text):
    want = qp_io.parse_mps(text.splitlines(), include_names=True)
    same_qp(native_io.qp_from_mps_text(text, include_names=True), want, names=True)
    without = native_io.qp_from_mps_text(text)
    assert without.problem_name is None and without.variable_names is None
    same_qp(without, want)


def test_mps_example_known_answers():  # mps_reader_template.h:38-75
    qp = native_io.qp_from_mps_text(EXAMPLE_MPS, include_names=True)
    assert qp.problem_name == "TESTEQ"
    assert qp.variable_names == ["XONE", "YTWO", "ZTHREE"] and qp.constraint_names == ["LIM1", "LIM2", "MYEQN"]
    np.testing.assert_array_equal(qp.objective_vector, [1, 2, 3])
    assert qp.objective_offset == 10
    np.testing.assert_array_equal(qp.constraint_matrix.toarray(), [[1, 1, 0], [1, 0, 0], [0, -1, 1]])
    np.testing.assert_array_equal(qp.constraint_lower_bounds, [-INF, 1, 7])
    np.testing.assert_array_equal(qp.constraint_upper_bounds, [4, INF, 7])
    np.testing.assert_array_equal(qp.variable_lower_bounds, [0, -1, 0])
    np.testing.assert_array_equal(qp.variable_upper_bounds, [4, 1, INF])


@pytest.mark.parametrize("text", [
    "NAME x\nROWS\n N c\n Q r\nENDATA\n", "NAME x\nROWS\n N c\n E r\n E r\nENDATA\n", "NAME x\nROWS\n N c\nCOLUMNS\n a nosuchrow 1\nENDATA\n",
    "NAME x\nROWS\n N c\n E r\nCOLUMNS\n a r one\nENDATA\n", "NAME x\nWHATEVER\nENDATA\n", "NAME x\nROWS\n N c\nCOLUMNS\n a c 1\nBOUNDS\n XX BND a 1\nENDATA\n",
    "NAME x\nROWS\n N c\nCOLUMNS\n a c 1\nQUADOBJ\n a a 1\nENDATA\n", "NAME x\nROWS\n N c\n E r\nCOLUMNS\n a r nan\nENDATA\n",
    " a b c\n",
])
def test_mps_errors_agree(text):
    with pytest.raises(qp_io.MpsError):
        qp_io.parse_mps(text.splitlines())
    with pytest.raises(native_io.NativeIoError, match="line "):
        native_io.qp_from_mps_text(text)


def test_files_round_trip_and_suffix_dispatch(tmp_path):  # quadratic_program_io.cc:50-101
    lp = fixtures.test_lp()
    lp.problem_name = "test_lp"
    lp.variable_names, lp.constraint_names = ["x0", "x1", "x2", "x3"], ["c0", "c1", "c2", "c3"]
    mps = str(tmp_path / "lp.mps")
    native_io.write_linear_program_to_mps(lp, mps)
    same_qp(native_io.read_quadratic_program(mps, include_names=True), lp, names=True)
    same_qp(qp_io.read_quadratic_program(mps, include_names=True), lp, names=True)     # the Python reader reads our MPS
    py_mps = str(tmp_path / "py.mps")
    qp_io.write_linear_program_to_mps(lp, py_mps)
    same_qp(native_io.read_quadratic_program(py_mps, include_names=True), lp, names=True)  # and we read its MPS
    for scale in (1.0, -1.0):                                                          # maximisation survives the trip
        lp.objective_scaling_factor = scale
        native_io.write_linear_program_to_mps(lp, mps)
        same_qp(native_io.read_quadratic_program(mps), lp)
    gz = str(tmp_path / "lp.mps.gz")
    with gzip.open(gz, "wb") as f:
        f.write(open(mps, "rb").read())
    same_qp(native_io.read_quadratic_program(gz), lp)
    qp = fixtures.test_diagonal_qp1()
    pb = str(tmp_path / "qp.pb")
    native_io.write_quadratic_program_to_mp_model_proto(qp, pb)
    assert open(pb, "rb").read() == mp_model.qp_to_mp_model_proto(qp).SerializeToString()
    same_qp(native_io.read_quadratic_program(pb), qp)
    with gzip.open(pb + ".gz", "wb") as f:
        f.write(open(pb, "rb").read())
    same_qp(native_io.read_quadratic_program(pb + ".gz"), qp)
    tp = str(tmp_path / "lp.textproto")
    open(tp, "w").write(TEST_LP_PROTO)
    same_qp(native_io.read_quadratic_program(tp), qp_io.read_quadratic_program(tp))
    js = str(tmp_path / "qp.json")
    open(js, "w").write(json_format.MessageToJson(mp_model.qp_to_mp_model_proto(qp)))
    same_qp(native_io.read_quadratic_program(js), qp)
    with gzip.open(js + ".gz", "wb") as f:
        f.write(open(js, "rb").read())
    same_qp(native_io.read_quadratic_program(js + ".gz"), qp)
    with pytest.raises(native_io.NativeIoError, match="Invalid filename suffix"):
        native_io.read_quadratic_program(str(tmp_path / "lp.txt"))
    with pytest.raises(native_io.NativeIoError, match="cannot open"):
        native_io.read_quadratic_program(str(tmp_path / "missing.mps"))
    with pytest.raises(native_io.NativeIoError, match="quadratic objective"):
        native_io.write_linear_program_to_mps(qp, str(tmp_path / "qp.mps"))


def test_larger_random_model_through_both_readers(tmp_path):
    rng = np.random.default_rng(7)
    import scipy.sparse as sp
    m, n = 300, 500
    k = sp.random(m, n, density=0.02, random_state=7, format="csc", data_rvs=lambda s: rng.normal(size=s))
    qp = pdlp.QuadraticProgram(n, m)
    qp.constraint_matrix = k
    qp.objective_vector = rng.normal(size=n)
    lo = rng.normal(size=m)
    kind = rng.integers(0, 4, size=m)
    qp.constraint_lower_bounds = np.where(kind == 1, -INF, lo)
    qp.constraint_upper_bounds = np.where(kind == 0, lo, np.where(kind == 2, INF, lo + 1.0))
    qp.variable_lower_bounds = np.where(rng.uniform(size=n) < 0.3, -INF, rng.normal(size=n) - 2)
    qp.variable_upper_bounds = np.where(rng.uniform(size=n) < 0.3, INF, rng.normal(size=n) + 2)
    qp.objective_offset = 3.25
    blob = native_io.qp_to_mp_model_proto_bytes(qp)
    assert blob == mp_model.qp_to_mp_model_proto(qp).SerializeToString()
    same_qp(native_io.qp_from_mp_model_proto_bytes(blob, False), qp)
    path = str(tmp_path / "random.mps")
    native_io.write_linear_program_to_mps(qp, path)
    got, want = native_io.read_quadratic_program(path), qp_io.read_quadratic_program(path)
    same_qp(got, want)
    # ranged rows are written as G + RANGES, so the upper bound comes back as lower + (upper - lower)
    np.testing.assert_allclose(got.constraint_upper_bounds, qp.constraint_upper_bounds, rtol=1e-15)
    np.testing.assert_array_equal(got.constraint_lower_bounds, qp.constraint_lower_bounds)
    np.testing.assert_array_equal(got.constraint_matrix.toarray(), qp.constraint_matrix.toarray())


# --------------------------------------------------------------------------------------------------
# PdlpSolveProto
# --------------------------------------------------------------------------------------------------
def _response(blob):
    resp = mp_model.MPSolutionResponseProto()
    resp.ParseFromString(blob)
    return resp


def test_solve_proto_rejections_need_no_device():  # pdlp_proto_solver.cc:47-66
    st = mp_model.MPSolverResponseStatus
    req = _request(False)
    req.solver_specific_parameters = "no_such_parameter: 1"
    assert _response(native_io.solve_proto(req.SerializeToString())).status == st.MPSOLVER_MODEL_INVALID_SOLVER_PARAMETERS
    flag = C.c_int32(1)
    assert _response(native_io.solve_proto(_request(False).SerializeToString(), interrupt_solve=flag)).status == st.MPSOLVER_NOT_SOLVED
    no_model = _request(False)
    no_model.ClearField("model")
    assert _response(native_io.solve_proto(no_model.SerializeToString())).status == st.MPSOLVER_MODEL_INVALID
    bad = _request(False)
    bad.model.general_constraint.add()
    resp = _response(native_io.solve_proto(bad.SerializeToString()))
    assert resp.status == st.MPSOLVER_MODEL_INVALID and "General constraints" in resp.status_str


def test_solve_proto_has_no_cpu_fallback():
    if pdlp.backend().device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        native_io.solve_proto(_request(False).SerializeToString())


@pytest.mark.gpu
@pytest.mark.parametrize("maximize", [False, True])
def test_solve_proto_on_the_gpu(b200_backend, maximize):
    req = _request(maximize)
    req.solver_time_limit_seconds = 60.0
    resp = _response(native_io.solve_proto(req.SerializeToString()))
    _check_tiny_lp_response(resp, maximize)
    log = pdlp_proto.SolveLogProto()
    log.ParseFromString(resp.solver_specific_info)
    assert log.params.termination_criteria.time_sec_limit == 60.0
    assert log.params.termination_criteria.simple_optimality_criteria.eps_optimal_absolute == 1e-8


@pytest.mark.gpu
def test_solve_log_of_a_device_solve(b200_backend):
    params = pdlp.PrimalDualHybridGradientParams()
    params.record_iteration_stats = True
    blobs = []
    result = b200_backend.primal_dual_hybrid_gradient(fixtures.test_lp(), params, result_pod_consumer=lambda res: blobs.append(native_io.solve_log_serialize(res)))
    got = pdlp_proto.SolveLogProto()
    got.ParseFromString(blobs[0])
    assert _strip_params(got) == _strip_params(pdlp_proto.solve_log_to_proto(result.solve_log, params))


# --------------------------------------------------------------------------------------------------
# the command-line front end (examples/pdlp_solve_cli.cc; counterpart of examples/cpp/pdlp_solve.cc)
# --------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def cli():
    import os
    import __graft_entry__
    if not os.path.exists(__graft_entry__.CLI):
        __graft_entry__.build()
    return __graft_entry__.CLI


def _run_cli(cli, *args):
    import subprocess
    return subprocess.run([cli, *args], capture_output=True, text=True, timeout=300)


def test_cli_rejects_bad_invocations_before_touching_the_device(cli, tmp_path):
    mps = str(tmp_path / "example.mps")
    open(mps, "w").write(EXAMPLE_MPS)
    p = _run_cli(cli)
    assert p.returncode == 1 and "--input is required" in p.stderr
    p = _run_cli(cli, "--input=" + mps, "--params=no_such_field: 1")
    assert p.returncode == 1 and "Error parsing --params" in p.stderr and "no_such_field" in p.stderr
    p = _run_cli(cli, "--input", mps, "--solve_log_file", str(tmp_path / "log.txt"))
    assert p.returncode == 1 and "Expected .textproto, .pb, or .json" in p.stderr
    p = _run_cli(cli, "--input=" + str(tmp_path / "missing.mps"))
    assert p.returncode == 1 and "cannot open" in p.stderr
    p = _run_cli(cli, "--input=" + str(tmp_path / "model.lp"))
    assert p.returncode == 1 and "Invalid filename suffix" in p.stderr
    p = _run_cli(cli, "--frobnicate")
    assert p.returncode == 1 and "usage:" in p.stderr


def test_cli_fails_loudly_without_a_gpu(cli, tmp_path):
    if pdlp.backend().device_count() > 0:
        pytest.skip("a CUDA device is present")
    mps = str(tmp_path / "example.mps")
    open(mps, "w").write(EXAMPLE_MPS)
    p = _run_cli(cli, "--input=" + mps)
    assert p.returncode == 1 and "no usable CUDA device" in p.stderr


@pytest.mark.gpu
def test_cli_solves_an_mps_file_on_the_gpu(cli, tmp_path, b200_backend):
    # the worked example of mps_reader_template.h:38-75: min x + 2y + 3z + 10, optimum at x = 1, y = -1, z = 6 (objective 27)
    mps = str(tmp_path / "example.mps")
    open(mps, "w").write(EXAMPLE_MPS)
    log, sol = str(tmp_path / "log.json"), str(tmp_path / "out.sol")
    p = _run_cli(cli, "--input=" + mps, "--params=termination_criteria { simple_optimality_criteria { eps_optimal_absolute: 1e-8 eps_optimal_relative: 1e-8 } } verbosity_level: 0",
                 "--solve_log_file=" + log, "--sol_file=" + sol)
    assert p.returncode == 0, p.stderr
    assert "TERMINATION_REASON_OPTIMAL" in p.stderr
    msg = json_format.Parse(open(log).read(), pdlp_proto.SolveLogProto())
    assert msg.termination_reason == pdlp.TerminationReason.TERMINATION_REASON_OPTIMAL and msg.instance_name == "TESTEQ"
    assert msg.params.verbosity_level == 0 and msg.params.termination_criteria.simple_optimality_criteria.eps_optimal_absolute == 1e-8
    lines = open(sol).read().split("\n")
    assert lines[0].startswith("=obj= ")
    want = pdlp.backend().primal_dual_hybrid_gradient(qp_io.parse_mps(EXAMPLE_MPS.splitlines()), pdlp_proto.params_from_text(
        "termination_criteria { simple_optimality_criteria { eps_optimal_absolute: 1e-8 eps_optimal_relative: 1e-8 } }"))
    got = {l.split()[0]: float(l.split()[1]) for l in lines[1:] if l}
    assert list(got) == ["XONE", "YTWO", "ZTHREE"]
    np.testing.assert_allclose(list(got.values()), want.primal_solution, rtol=0, atol=1e-9)   # the same library on the same inputs
    np.testing.assert_allclose(list(got.values()), [1, -1, 6], atol=1e-6)
    assert float(lines[0].split()[1]) == pytest.approx(27.0, abs=1e-6)


def test_format_layer_survives_mutation_fuzzing(tmp_path):
    """tools/fuzz_formats.cc under AddressSanitizer + UBSan: mutated parameter texts, JSON, MPS and
    serialized messages through every host-only entry point must all come back as status codes."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    seeds = []
    for name, blob in (("lp.pb", mp_model.qp_to_mp_model_proto(fixtures.test_lp()).SerializeToString()),
                       ("req.pb", _request(True).SerializeToString()),
                       ("params.pb", text_format.Parse(PARAM_TEXTS[2] + " random_projection_seeds: [1, 2]", pdlp_proto.PrimalDualHybridGradientParamsProto()).SerializeToString())):
        path = str(tmp_path / name)
        open(path, "wb").write(blob)
        seeds.append(path)
    p = subprocess.run([os.path.join(root, "tools", "fuzz_formats.sh"), "4000", *seeds], capture_output=True, text=True, timeout=600)
    if p.returncode != 0 and ("cannot find -lasan" in p.stderr or "libasan" in p.stderr and "No such file" in p.stderr):
        pytest.skip("no sanitizer runtime in this toolchain")
    assert p.returncode == 0 and "fuzz ok: 4000 inputs" in p.stdout, p.stdout[-2000:] + p.stderr[-4000:]


def test_writers_reject_a_malformed_view():
    """The format writers walk the caller's CSC arrays on the host: a view whose row indices, column starts
    or nonzero count are inconsistent is refused with BAD_ARGUMENT instead of being indexed."""
    import ctypes as C
    from ortools_b200 import _capi as capi
    qp = fixtures.test_lp()
    be = pdlp.backend()
    blob = native_io.PdlpBlob()
    err = C.create_string_buffer(256)
    fn = be.fn("qp_to_mp_model_proto")

    def call(mutate):
        view, keep = qp._to_view()
        mutate(view, keep)
        rc = fn(C.byref(view), None, None, C.byref(blob), err, C.c_int64(256))
        del keep
        return rc

    assert call(lambda v, k: None) == 0
    be.fn("blob_free", None)(C.byref(blob))

    def bad_row(v, k):
        k["row_indices"][1] = v.num_constraints + 3
    assert call(bad_row) != 0 and b"row index" in err.value

    def bad_nnz(v, k):
        v.num_nonzeros = v.num_nonzeros + 1
    assert call(bad_nnz) != 0 and b"num_nonzeros" in err.value

    def bad_starts(v, k):
        k["col_starts"][1] = k["col_starts"][2] + 1
    assert call(bad_starts) != 0 and b"monotone" in err.value
