"""ctypes binding of the native format layer (include/pdlp_b200_io.h): parameter / log
protos, MPModelProto and MPS conversions and ``PdlpSolveProto`` implemented in C++ inside
libpdlp_b200.so (csrc/proto_codec.cc, csrc/formats.cc) -- no protobuf runtime involved.
Everything here except ``solve_proto`` is host-only.

The pure-Python adapters (``pdlp_proto``, ``mp_model``, ``qp_io``) do the same jobs on top of
the ``google.protobuf`` runtime; ``tests/test_native_io.py`` checks the two against each other.
"""
import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import _capi as capi
from . import pdlp

BINARY, TEXT, JSON = 0, 1, 2


class PdlpBlob(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_uint8)), ("size", C.c_int64)]


class NativeIoError(ValueError):
    """PDLP_B200_STATUS_BAD_ARGUMENT with the library's explanation."""


def _lib():
    return pdlp.backend()


def _take(blob):
    try:
        return C.string_at(blob.data, blob.size) if blob.size > 0 else b""
    finally:
        _lib().fn("blob_free", None)(C.byref(blob))


def _call(name, *args):
    err = C.create_string_buffer(2048)
    rc = _lib().fn(name)(*args, err, C.c_int64(len(err)))
    if rc == 3:
        raise NativeIoError(err.value.decode(errors="replace"))
    _lib()._check(rc, name)


# ---- parameters -------------------------------------------------------------------------------
def params_from_text(text, onto=None):
    """Text-format PrimalDualHybridGradientParams -> PdlpParams POD. `onto`: merge onto this POD
    (protobuf MergeFrom semantics) instead of the defaults."""
    if onto is None:
        p = capi.PdlpParams()
        _call("params_parse_text", text.encode(), C.byref(p))
    else:
        p = capi.PdlpParams.from_buffer_copy(onto)
        _call("params_merge_text", text.encode(), C.byref(p))
    return p


def params_from_bytes(blob, onto=None):
    buf = (C.c_uint8 * max(1, len(blob))).from_buffer_copy(blob or b"\0")
    if onto is None:
        p = capi.PdlpParams()
        _call("params_parse_bytes", buf, C.c_int64(len(blob)), C.byref(p))
    else:
        p = capi.PdlpParams.from_buffer_copy(onto)
        _call("params_merge_bytes", buf, C.c_int64(len(blob)), C.byref(p))
    return p


def params_serialize(params, fmt=BINARY):
    pod = pdlp.params_to_pod(params)
    out = PdlpBlob()
    _lib()._check(_lib().fn("params_serialize")(C.byref(pod), C.c_int32(fmt), C.byref(out)), "params_serialize")
    return _take(out)


# ---- solve log --------------------------------------------------------------------------------
def solve_log_serialize(result_pod, fmt=BINARY):
    """SolveLog of a raw PdlpResult (see Backend.primal_dual_hybrid_gradient(result_pod_consumer=...))."""
    out = PdlpBlob()
    _lib()._check(_lib().fn("solve_log_serialize")(C.byref(result_pod), C.c_int32(fmt), C.byref(out)), "solve_log_serialize")
    return _take(out)


def write_solve_log(result_pod, path):
    _call("write_solve_log", C.byref(result_pod), str(path).encode())


# ---- problems ---------------------------------------------------------------------------------
def _qp_from_model(handle, include_names):
    lib = _lib()
    try:
        view_fn = lib.fn("model_view", C.POINTER(capi.PdlpProblemView))
        v = view_fn(handle).contents
        n, m, nnz = v.num_variables, v.num_constraints, v.num_nonzeros

        def arr(p, count, dtype):
            return np.ctypeslib.as_array(p, shape=(count,)).astype(dtype, copy=True) if count > 0 else np.zeros(0, dtype=dtype)

        qp = pdlp.QuadraticProgram(n, m)
        indptr = arr(v.col_starts, n + 1, np.int64)
        qp.constraint_matrix = sp.csc_matrix((arr(v.values, nnz, np.float64), arr(v.row_indices, nnz, np.int64), indptr), shape=(m, n))
        qp.objective_vector = arr(v.objective_vector, n, np.float64)
        qp.objective_matrix = arr(v.objective_matrix_diagonal, n, np.float64) if v.objective_matrix_diagonal else None
        qp.constraint_lower_bounds = arr(v.constraint_lower_bounds, m, np.float64)
        qp.constraint_upper_bounds = arr(v.constraint_upper_bounds, m, np.float64)
        qp.variable_lower_bounds = arr(v.variable_lower_bounds, n, np.float64)
        qp.variable_upper_bounds = arr(v.variable_upper_bounds, n, np.float64)
        qp.objective_offset = v.objective_offset
        qp.objective_scaling_factor = v.objective_scaling_factor
        if include_names:
            qp.problem_name = v.problem_name.decode() if v.problem_name is not None else ""
            vn, cn = lib.fn("model_variable_name", C.c_char_p), lib.fn("model_constraint_name", C.c_char_p)
            qp.variable_names = [(vn(handle, C.c_int64(j)) or b"").decode() for j in range(n)]
            qp.constraint_names = [(cn(handle, C.c_int64(i)) or b"").decode() for i in range(m)]
        return qp
    finally:
        lib.fn("model_free", None)(handle)


def read_quadratic_program(path, include_names=False):
    """ReadQuadraticProgramOrDie in C++ (raises NativeIoError instead of dying)."""
    h = C.c_void_p()
    _call("read_quadratic_program", str(path).encode(), C.c_int32(int(include_names)), C.byref(h))
    return _qp_from_model(h, include_names)


def qp_from_mps_text(text, include_names=False):
    raw = text.encode() if isinstance(text, str) else bytes(text)
    h = C.c_void_p()
    _call("model_from_mps_text", raw, C.c_int64(len(raw)), C.c_int32(int(include_names)), C.byref(h))
    return _qp_from_model(h, include_names)


def qp_from_mp_model_proto_bytes(blob, relax_integer_variables, include_names=False):
    buf = (C.c_uint8 * max(1, len(blob))).from_buffer_copy(blob or b"\0")
    h = C.c_void_p()
    _call("model_from_mp_model_proto", buf, C.c_int64(len(blob)), C.c_int32(int(relax_integer_variables)), C.c_int32(int(include_names)), C.byref(h))
    return _qp_from_model(h, include_names)


def _names(names, count):
    if not names:
        return None, None
    keep = [(s or "").encode() for s in names]
    arr = (C.c_char_p * count)(*keep)
    return arr, keep


def qp_to_mp_model_proto_bytes(qp):
    view, keep = qp._to_view()
    n, m = view.num_variables, view.num_constraints
    vn, k1 = _names(getattr(qp, "variable_names", None), n)
    cn, k2 = _names(getattr(qp, "constraint_names", None), m)
    out = PdlpBlob()
    _call("qp_to_mp_model_proto", C.byref(view), vn, cn, C.byref(out))
    del keep, k1, k2
    return _take(out)


def write_linear_program_to_mps(qp, path):
    view, keep = qp._to_view()
    vn, k1 = _names(getattr(qp, "variable_names", None), view.num_variables)
    cn, k2 = _names(getattr(qp, "constraint_names", None), view.num_constraints)
    _call("write_linear_program_to_mps", C.byref(view), vn, cn, str(path).encode())
    del keep, k1, k2


def write_quadratic_program_to_mp_model_proto(qp, path):
    view, keep = qp._to_view()
    vn, k1 = _names(getattr(qp, "variable_names", None), view.num_variables)
    cn, k2 = _names(getattr(qp, "constraint_names", None), view.num_constraints)
    _call("write_quadratic_program_to_mp_model_proto", C.byref(view), vn, cn, str(path).encode())
    del keep, k1, k2


# ---- PdlpSolveProto ---------------------------------------------------------------------------
def solve_proto(request_bytes, relax_integer_variables=False, interrupt_solve=None):
    """Serialized MPModelRequest -> serialized MPSolutionResponse (needs a CUDA device unless the
    request is rejected before the solve)."""
    buf = (C.c_uint8 * max(1, len(request_bytes))).from_buffer_copy(request_bytes or b"\0")
    out = PdlpBlob()
    flag = None if interrupt_solve is None else C.byref(interrupt_solve)
    rc = _lib().fn("solve_proto")(buf, C.c_int64(len(request_bytes)), C.c_int32(int(relax_integer_variables)), flag, C.byref(out))
    _lib()._check(rc, "solve_proto")
    return _take(out)


def convert(message, data, from_format, to_format):
    """Re-encodes a message between binary / text / JSON (pdlp_b200_proto_convert)."""
    raw = data.encode() if isinstance(data, str) else bytes(data)
    buf = (C.c_uint8 * max(1, len(raw))).from_buffer_copy(raw or b"\0")
    out = PdlpBlob()
    _call("proto_convert", message.encode(), C.c_int32(from_format), buf, C.c_int64(len(raw)), C.c_int32(to_format), C.byref(out))
    got = _take(out)
    return got if to_format == BINARY else got.decode()
