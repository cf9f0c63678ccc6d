"""Wire / text-format adapters: the runtime-built schemas of
``ortools_b200.pdlp_proto`` against the tag tables extracted from the reference's own
``solvers.proto`` / ``solve_log.proto`` (tests/golden/pdlp_proto_tags.json, made by
tools/make_proto_tag_fixture.py), known-answer encodings, and a solve whose SolveLog
goes through bytes and text and back."""
import json
import math
import os

import numpy as np
import pytest

from ortools_b200 import pdlp, pdlp_proto
import fixtures

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "pdlp_proto_tags.json")
PKG = "operations_research.pdlp."
TYPE_NAMES = {1: "double", 5: "int32", 3: "int64", 8: "bool", 9: "string"}


def _descriptor(name):
    return pdlp_proto._pool.FindMessageTypeByName(PKG + name)


def test_every_reference_field_has_the_same_tag_type_label_and_default():
    golden = json.load(open(GOLDEN))
    for msg_name, fields in golden["messages"].items():
        d = _descriptor(msg_name)
        ours = {f.name: f for f in d.fields}
        for fname, (tag, label, ftype, default) in fields.items():
            if fname == "glop_parameters":  # host-only presolve options: opaque on this path
                assert fname not in ours
                continue
            assert fname in ours, (msg_name, fname)
            f = ours[fname]
            assert f.number == tag, (msg_name, fname)
            assert pdlp_proto._is_repeated(f) == (label == "repeated"), (msg_name, fname)
            if f.type in TYPE_NAMES:
                assert TYPE_NAMES[f.type] == ftype, (msg_name, fname)
            elif f.type == f.TYPE_ENUM:
                assert f.enum_type.name == ftype, (msg_name, fname)
            else:
                assert f.message_type.name == ftype, (msg_name, fname)
            if default is not None:
                assert f.has_default_value, (msg_name, fname)
                if f.type == f.TYPE_ENUM:
                    assert f.enum_type.values_by_number[f.default_value].name == default
                elif f.type == f.TYPE_BOOL:
                    assert f.default_value == (default == "true")
                else:
                    assert float(f.default_value) == float(default), (msg_name, fname)
        # nothing extra on our side
        assert set(ours) <= set(fields), (msg_name, set(ours) - set(fields))


def test_every_reference_enum_value_has_the_same_number():
    golden = json.load(open(GOLDEN))
    for enum_name, values in golden["enums"].items():
        e = pdlp_proto._pool.FindEnumTypeByName(PKG + enum_name)
        assert {v.name: v.number for v in e.values} == values, enum_name


def test_known_answer_encodings():
    p = pdlp_proto.PrimalDualHybridGradientParamsProto()
    p.termination_criteria.iteration_limit = 10
    assert p.SerializeToString() == bytes([0x0A, 0x02, 0x38, 0x0A])           # field 1 (len 2) { field 7 varint 10 }
    p = pdlp_proto.PrimalDualHybridGradientParamsProto()
    p.restart_strategy = pdlp.RestartStrategy.NO_RESTARTS
    p.verbosity_level = 3
    assert p.SerializeToString() == bytes([0x30, 0x01, 0xD0, 0x01, 0x03])     # field 6 varint 1; field 26 varint 3
    p = pdlp_proto.PrimalDualHybridGradientParamsProto()
    p.random_projection_seeds.extend([1, 2, 300])
    assert p.SerializeToString() == bytes([0xE2, 0x01, 0x04, 0x01, 0x02, 0xAC, 0x02])  # field 28 packed
    log = pdlp_proto.SolveLogProto()
    log.termination_reason = pdlp.TerminationReason.TERMINATION_REASON_OPTIMAL
    log.iteration_count = 576
    assert log.SerializeToString() == bytes([0x18, 0x01, 0x28, 0xC0, 0x04])   # field 3 varint 1; field 5 varint 576
    t = pdlp_proto.TerminationCriteriaProto()
    t.simple_optimality_criteria.eps_optimal_absolute = 1.0
    assert t.SerializeToString() == bytes([0x4A, 0x09, 0x09]) + np.float64(1.0).tobytes()  # field 9 { field 1 fixed64 }


def test_params_text_round_trip_keeps_presence_and_the_oneof():
    text = """
      termination_criteria {
        simple_optimality_criteria { eps_optimal_absolute: 1e-4 eps_optimal_relative: 1e-4 }
        iteration_limit: 2000
      }
      restart_strategy: ADAPTIVE_DISTANCE_BASED
      linesearch_rule: MALITSKY_POCK_LINESEARCH_RULE
      malitsky_pock_parameters { step_size_downscaling_factor: 0.5 }
      random_projection_seeds: 7 random_projection_seeds: 11
      l2_norm_rescaling: false
    """
    p = pdlp_proto.params_from_text(text)
    assert p.termination_criteria.WhichOneof("optimality_criteria") == "simple_optimality_criteria"
    assert p.termination_criteria.simple_optimality_criteria.eps_optimal_absolute == 1e-4
    assert p.termination_criteria.iteration_limit == 2000
    assert not p.termination_criteria.HasField("eps_optimal_absolute")     # deprecated fields stay unset
    assert p.restart_strategy == pdlp.RestartStrategy.ADAPTIVE_DISTANCE_BASED
    assert p.linesearch_rule == pdlp.LinesearchRule.MALITSKY_POCK_LINESEARCH_RULE
    assert p.malitsky_pock_parameters.step_size_downscaling_factor == 0.5
    assert p.malitsky_pock_parameters.linesearch_contraction_factor == 0.99  # default
    assert tuple(p.random_projection_seeds) == (7, 11)
    assert p.l2_norm_rescaling is False and p.HasField("l2_norm_rescaling")
    assert not p.HasField("initial_primal_weight")
    # proto -> object -> proto -> bytes -> object
    again = pdlp_proto.params_from_bytes(pdlp_proto.params_to_proto(p).SerializeToString())
    assert repr(again) == repr(p)
    # the POD the C ABI receives is identical either way
    a, b = p.to_pod(), again.to_pod()
    assert bytes(a) == bytes(b)


def test_detailed_criteria_and_invalid_mix_are_reported_like_the_reference():
    p = pdlp_proto.params_from_text("termination_criteria { detailed_optimality_criteria { eps_optimal_objective_gap_absolute: 1e-3 } }")
    assert p.termination_criteria.WhichOneof("optimality_criteria") == "detailed_optimality_criteria"
    # deprecated eps_optimal_* together with the oneof is INVALID_PARAMETER (solvers_proto_validation.cc:51-63)
    bad = pdlp_proto.params_from_text("termination_criteria { eps_optimal_absolute: 1e-3 simple_optimality_criteria { } }")
    res = _oracle().primal_dual_hybrid_gradient(fixtures.test_lp(), bad)
    assert res.solve_log.termination_reason == pdlp.TerminationReason.TERMINATION_REASON_INVALID_PARAMETER


def _oracle():
    from oracle import pdlp_oracle
    return pdlp_oracle.backend()


def test_solve_log_goes_through_bytes_and_text():
    params = pdlp_proto.params_from_text("""
      termination_criteria { simple_optimality_criteria { eps_optimal_absolute: 1e-8 eps_optimal_relative: 1e-8 } }
      record_iteration_stats: true
      major_iteration_frequency: 16 termination_check_frequency: 16
    """)
    qp = fixtures.test_lp()
    qp.problem_name = "test_lp"
    res = _oracle().primal_dual_hybrid_gradient(qp, params)
    assert res.solve_log.termination_reason == pdlp.TerminationReason.TERMINATION_REASON_OPTIMAL
    msg = pdlp_proto.solve_log_to_proto(res.solve_log, params)
    back = pdlp_proto.SolveLogProto()
    back.ParseFromString(msg.SerializeToString())
    assert back == msg
    assert back.termination_reason == pdlp.TerminationReason.TERMINATION_REASON_OPTIMAL
    assert back.iteration_count == res.solve_log.iteration_count
    assert back.solution_type == res.solve_log.solution_type
    assert len(back.iteration_stats) == len(res.solve_log.iteration_stats) > 0
    # what the consumers read (pdlp_proto_solver.cc:80-127): the ConvergenceInformation of the solution type
    ci = [c for c in back.solution_stats.convergence_information if c.candidate_type == back.solution_type]
    assert len(ci) == 1 and ci[0].primal_objective == pytest.approx(-34.0, abs=1e-6)   # primal_dual_hybrid_gradient_test.cc TestLp optimum
    assert back.original_problem_stats.num_variables == 4 and back.original_problem_stats.num_constraints == 4
    assert back.params.record_iteration_stats and back.params.major_iteration_frequency == 16
    assert back.params.termination_criteria.simple_optimality_criteria.eps_optimal_absolute == 1e-8
    # text format round trip
    from google.protobuf import text_format
    text = pdlp_proto.solve_log_to_text(res.solve_log, params)
    assert "termination_reason: TERMINATION_REASON_OPTIMAL" in text
    assert text_format.Parse(text, pdlp_proto.SolveLogProto()) == msg


def test_empty_statistics_keep_nan_and_infinity():
    s = pdlp_proto.QuadraticProgramStatsProto()
    s.constraint_matrix_abs_min = float("nan")
    s.variable_bound_gaps_max = float("inf")
    back = pdlp_proto.QuadraticProgramStatsProto()
    back.ParseFromString(s.SerializeToString())
    assert math.isnan(back.constraint_matrix_abs_min) and back.variable_bound_gaps_max == float("inf")
