"""Property tests of the native proto2 codec (csrc/proto_codec.cc) against the ``google.protobuf``
runtime: random messages of every schema the library knows go binary -> text / JSON -> binary through
``pdlp_b200_proto_convert`` and through the runtime, in both directions, and must come back equal."""
import math
import os

import pytest
from google.protobuf import json_format, text_format
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from ortools_b200 import mp_model, native_io, pdlp_proto

MESSAGES = {
    "PrimalDualHybridGradientParams": pdlp_proto.PrimalDualHybridGradientParamsProto,
    "TerminationCriteria": pdlp_proto.TerminationCriteriaProto,
    "SolveLog": pdlp_proto.SolveLogProto,
    "IterationStats": pdlp_proto.IterationStatsProto,
    "MPModelProto": mp_model.MPModelProto,
    "MPModelRequest": mp_model.MPModelRequestProto,
    "MPSolutionResponse": mp_model.MPSolutionResponseProto,
}
SKIP_FIELDS = {"general_constraint"}  # carried only as "present" by the native schema

doubles = st.one_of(st.floats(allow_nan=False), st.sampled_from([0.0, -0.0, 1.0, -1.0, 1e-6, 1e300, -1e-300, math.inf, -math.inf, 0.1, 1 / 3]))
texts = st.text(alphabet=st.characters(blacklist_categories=("Cs",)), max_size=12)


def scalar(fd):
    if fd.type == fd.TYPE_DOUBLE:
        return doubles
    if fd.type == fd.TYPE_INT32:
        return st.integers(-2**31, 2**31 - 1)
    if fd.type == fd.TYPE_INT64:
        return st.integers(-2**63, 2**63 - 1)
    if fd.type == fd.TYPE_BOOL:
        return st.booleans()
    if fd.type == fd.TYPE_STRING:
        return texts
    if fd.type == fd.TYPE_BYTES:
        return st.binary(max_size=12)
    if fd.type == fd.TYPE_ENUM:
        return st.sampled_from([v.number for v in fd.enum_type.values])
    raise AssertionError(fd.type)


@st.composite
def messages(draw, cls, depth=0):
    msg = cls()
    oneofs_taken = set()
    for fd in msg.DESCRIPTOR.fields:
        if fd.name in SKIP_FIELDS or not draw(st.booleans()):
            continue
        if fd.containing_oneof is not None:
            if fd.containing_oneof.name in oneofs_taken:
                continue
            oneofs_taken.add(fd.containing_oneof.name)
        repeated = pdlp_proto._is_repeated(fd)
        if fd.type == fd.TYPE_MESSAGE:
            if depth >= 3:
                continue
            sub_cls = type(getattr(msg, fd.name).add()) if repeated else type(getattr(msg, fd.name))
            if repeated:
                del getattr(msg, fd.name)[:]
                for _ in range(draw(st.integers(0, 2))):
                    getattr(msg, fd.name).add().CopyFrom(draw(messages(sub_cls, depth + 1)))
            else:
                getattr(msg, fd.name).CopyFrom(draw(messages(sub_cls, depth + 1)))
        elif repeated:
            getattr(msg, fd.name).extend(draw(st.lists(scalar(fd), max_size=3)))
        else:
            setattr(msg, fd.name, draw(scalar(fd)))
    return msg


@pytest.mark.parametrize("name", sorted(MESSAGES))
@settings(max_examples=int(os.environ.get("PDLP_B200_PROPERTY_EXAMPLES", "60")), deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])
@given(data=st.data())
def test_codec_round_trips_agree_with_the_protobuf_runtime(name, data):
    cls = MESSAGES[name]
    msg = data.draw(messages(cls))
    blob = msg.SerializeToString()
    # binary -> text / JSON by the native codec, read back by the runtime
    as_text = native_io.convert(name, blob, native_io.BINARY, native_io.TEXT)
    assert text_format.Parse(as_text, cls()) == msg, as_text
    as_json = native_io.convert(name, blob, native_io.BINARY, native_io.JSON)
    assert json_format.Parse(as_json, cls()) == msg, as_json
    # text / JSON written by the runtime, read by the native codec: canonical bytes of the same message
    for encoded, fmt in ((text_format.MessageToString(msg), native_io.TEXT), (json_format.MessageToJson(msg), native_io.JSON),
                         (json_format.MessageToJson(msg, preserving_proto_field_name=True), native_io.JSON), (as_text, native_io.TEXT), (as_json, native_io.JSON)):
        back = cls()
        back.ParseFromString(native_io.convert(name, encoded, fmt, native_io.BINARY))
        assert back == msg, encoded
    # binary -> binary is the identity on canonical input
    assert native_io.convert(name, blob, native_io.BINARY, native_io.BINARY) == blob


# --------------------------------------------------------------------------------------------------
# typed adapters: parameters, models, MPS
# --------------------------------------------------------------------------------------------------
EXAMPLES = int(os.environ.get("PDLP_B200_PROPERTY_EXAMPLES", "60"))


@settings(max_examples=EXAMPLES, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])
@given(msg=messages(pdlp_proto.PrimalDualHybridGradientParamsProto))
def test_parameter_pod_agrees_with_the_python_adapter(msg):
    from test_native_io import same_pod
    if len(msg.random_projection_seeds) > 8:
        return
    want = pdlp_proto.params_from_proto(msg).to_pod()
    got = native_io.params_from_bytes(msg.SerializeToString())
    same_pod(got, want)
    same_pod(native_io.params_from_text(text_format.MessageToString(msg)), want)
    again = pdlp_proto.PrimalDualHybridGradientParamsProto()
    again.ParseFromString(native_io.params_serialize(got))
    same_pod(pdlp_proto.params_from_proto(again).to_pod(), want)          # POD -> bytes -> POD is the identity


@st.composite
def models(draw):
    import numpy as np
    n = draw(st.integers(0, 6))
    m = draw(st.integers(0, 5))
    finite = st.floats(-1e6, 1e6, allow_nan=False)
    bound = st.one_of(finite, st.just(math.inf), st.just(-math.inf))
    msg = mp_model.MPModelProto()
    for j in range(n):
        v = msg.variable.add()
        v.lower_bound, v.upper_bound, v.objective_coefficient = draw(bound), draw(bound), draw(finite)
        if draw(st.booleans()):
            v.name = "v%d" % j
    for i in range(m):
        c = msg.constraint.add()
        c.lower_bound, c.upper_bound = draw(bound), draw(bound)
        if n > 0:
            idx = draw(st.lists(st.integers(0, n - 1), max_size=5))
            c.var_index.extend(idx)
            c.coefficient.extend(draw(st.lists(finite, min_size=len(idx), max_size=len(idx))))
        if draw(st.booleans()):
            c.name = "c%d" % i
    msg.maximize = draw(st.booleans())
    msg.objective_offset = draw(finite)
    if n > 0 and draw(st.booleans()):
        diag = draw(st.lists(st.integers(0, n - 1), max_size=3, unique=True))
        msg.quadratic_objective.qvar1_index.extend(diag)
        msg.quadratic_objective.qvar2_index.extend(diag)
        msg.quadratic_objective.coefficient.extend(draw(st.lists(finite, min_size=len(diag), max_size=len(diag))))
    if draw(st.booleans()):
        msg.name = draw(st.text(alphabet="abc xyz", max_size=6))
    return msg


@settings(max_examples=EXAMPLES, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])
@given(msg=models(), names=st.booleans())
def test_model_conversion_agrees_with_the_python_adapter(msg, names):
    from test_native_io import same_qp
    want = mp_model.qp_from_mp_model_proto(msg, relax_integer_variables=True, include_names=names)
    got = native_io.qp_from_mp_model_proto_bytes(msg.SerializeToString(), True, include_names=names)
    same_qp(got, want, names=names)
    assert native_io.qp_to_mp_model_proto_bytes(got) == mp_model.qp_to_mp_model_proto(want).SerializeToString()


@settings(max_examples=EXAMPLES, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large, HealthCheck.function_scoped_fixture])
@given(msg=models())
def test_mps_written_natively_is_read_alike_by_both_readers(msg, tmp_path):
    import numpy as np
    from ortools_b200 import qp_io
    from test_native_io import same_qp
    msg.ClearField("quadratic_objective")
    qp = mp_model.qp_from_mp_model_proto(msg, True)
    free = (qp.constraint_lower_bounds == -math.inf) & (qp.constraint_upper_bounds == math.inf)
    crossed = qp.constraint_lower_bounds > qp.constraint_upper_bounds      # not expressible as a ranged MPS row
    if free.any() or crossed.any() or np.isinf(qp.constraint_lower_bounds[qp.constraint_lower_bounds == qp.constraint_upper_bounds]).any():
        return
    if (qp.variable_lower_bounds > qp.variable_upper_bounds).any() or (qp.variable_lower_bounds == math.inf).any() or (qp.variable_upper_bounds == -math.inf).any():
        return
    path = str(tmp_path / "model.mps")
    native_io.write_linear_program_to_mps(qp, path)
    a, b = native_io.read_quadratic_program(path), qp_io.read_quadratic_program(path)
    same_qp(a, b)
    np.testing.assert_array_equal(a.constraint_matrix.toarray(), qp.constraint_matrix.toarray())
    np.testing.assert_array_equal(a.objective_vector, qp.objective_vector)
    np.testing.assert_array_equal(a.variable_lower_bounds, qp.variable_lower_bounds)
    np.testing.assert_array_equal(a.variable_upper_bounds, qp.variable_upper_bounds)
    np.testing.assert_array_equal(a.constraint_lower_bounds, qp.constraint_lower_bounds)
    np.testing.assert_allclose(a.constraint_upper_bounds, qp.constraint_upper_bounds, rtol=1e-15, atol=1e-9)   # ranged rows: lower + (upper - lower)
    assert a.objective_offset == qp.objective_offset and a.objective_scaling_factor == qp.objective_scaling_factor


# --------------------------------------------------------------------------------------------------
# differential test of the two MPS readers on structured random text (valid and invalid)
# --------------------------------------------------------------------------------------------------
NUMBERS = ["1", "-2.5", "1e3", "1E-2", "1D2", "2.5d-1", "+4", "inf", "-inf", "Infinity", "-INFINITY", "1e400", ".5", "5.", "0", "-0", "1e-320",
           "nan", "abc", "0x10", "1e", "--1", "1.2.3"]
ROW_NAMES = ["r0", "r1", "r2", "obj", "obj2", "nosuch"]
COL_NAMES = ["x", "y", "z"]


@st.composite
def mps_texts(draw):
    lines = ["NAME " + draw(st.sampled_from(["", "model", "two words"]))]
    if draw(st.booleans()):
        lines += draw(st.sampled_from([["OBJSENSE", "    MAX"], ["OBJSENSE MAXIMIZE"], ["OBJSENSE", " MIN"], ["OBJSENSEMAX"]]))
    lines.append("ROWS")
    lines.append(" N obj")
    for r in ["r0", "r1", "r2"]:
        if draw(st.booleans()):
            lines.append(" %s %s" % (draw(st.sampled_from(["E", "L", "G", "e", "g", "N", "Q"])), r))
    if draw(st.booleans()):
        lines.append(" N obj2")
    lines.append("COLUMNS")
    for _ in range(draw(st.integers(0, 6))):
        if draw(st.integers(0, 9)) == 0:
            lines.append("    M1 'MARKER' '%s'" % draw(st.sampled_from(["INTORG", "INTEND"])))
            continue
        pairs = draw(st.integers(1, 2))
        toks = [draw(st.sampled_from(COL_NAMES))]
        for _ in range(pairs):
            toks += [draw(st.sampled_from(ROW_NAMES)), draw(st.sampled_from(NUMBERS))]
        lines.append("    " + " ".join(toks))
    for section in ("RHS", "RANGES"):
        if draw(st.booleans()):
            lines.append(section)
            for _ in range(draw(st.integers(0, 3))):
                toks = [draw(st.sampled_from(ROW_NAMES)), draw(st.sampled_from(NUMBERS))]
                if draw(st.booleans()):
                    toks = ["SET"] + toks
                if draw(st.integers(0, 3)) == 0:
                    toks += [draw(st.sampled_from(ROW_NAMES)), draw(st.sampled_from(NUMBERS))]
                lines.append("    " + " ".join(toks))
    if draw(st.booleans()):
        lines.append("BOUNDS")
        for _ in range(draw(st.integers(0, 4))):
            kind = draw(st.sampled_from(["LO", "UP", "FX", "FR", "MI", "PL", "BV", "LI", "UI", "SC", "XX", "up"]))
            toks = [kind]
            if draw(st.booleans()):
                toks.append("BND")
            toks.append(draw(st.sampled_from(COL_NAMES + ["w"])))
            if kind.upper() in ("LO", "UP", "FX", "LI", "UI", "SC") or draw(st.integers(0, 4)) == 0:
                toks.append(draw(st.sampled_from(NUMBERS)))
            lines.append(" " + " ".join(toks))
    if draw(st.integers(0, 5)) > 0:
        lines.append("ENDATA")
    if draw(st.integers(0, 6)) == 0:
        lines.insert(draw(st.integers(0, len(lines))), draw(st.sampled_from(["* a comment", "", "WHATEVER", "    stray data", "QUADOBJ"])))
    return "\n".join(lines) + "\n"


@settings(max_examples=4 * EXAMPLES, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])
@given(text=mps_texts(), names=st.booleans())
def test_both_mps_readers_agree_on_random_text(text, names):
    from ortools_b200 import qp_io
    from test_native_io import same_qp
    want = error = None
    try:
        want = qp_io.parse_mps(text.splitlines(), include_names=names)
    except qp_io.MpsError as e:
        error = str(e)
    if error is not None:
        with pytest.raises(native_io.NativeIoError) as caught:
            native_io.qp_from_mps_text(text, include_names=names)
        assert str(caught.value).split(":")[0] == error.split(":")[0], (text, error, str(caught.value))      # the same line is blamed
    else:
        same_qp(native_io.qp_from_mps_text(text, include_names=names), want, names=names)
