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
