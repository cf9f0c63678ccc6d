"""MPModelProto <-> QuadraticProgram and the MPSolver-style proto solver
(SURVEY.md 8f rank 3): the step either side of the hot path for callers that
hold an ``MPModelRequest`` (``ortools/linear_solver/proto_solver/pdlp_proto_solver.cc:36-130``).

The subset of ``ortools/linear_solver/linear_solver.proto`` this path reads and
writes is restated as runtime descriptors (package ``operations_research``,
same names and tags -- pinned by tests/golden/pdlp_proto_tags.json); fields the
path never touches stay in the unknown-field set of a parsed message and are
written back untouched.

    request = mp_model.MPModelRequestProto(); request.ParseFromString(blob)
    response = mp_model.pdlp_solve_proto(request)          # -> MPSolutionResponse
"""
import numpy as np
import scipy.sparse as sp
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory, text_format

from . import pdlp, pdlp_proto

_F = descriptor_pb2.FieldDescriptorProto
_PKG = "operations_research"
INF = float("inf")


class MPSolverResponseStatus:  # linear_solver.proto:519-586
    MPSOLVER_OPTIMAL = 0x0
    MPSOLVER_FEASIBLE = 0x1
    MPSOLVER_INFEASIBLE = 0x2
    MPSOLVER_UNBOUNDED = 0x3
    MPSOLVER_ABNORMAL = 0x4
    MPSOLVER_NOT_SOLVED = 0x6
    MPSOLVER_MODEL_IS_VALID = 0x61
    MPSOLVER_CANCELLED_BY_USER = 0x62
    MPSOLVER_UNKNOWN_STATUS = 0x63
    MPSOLVER_MODEL_INVALID = 0x5
    MPSOLVER_MODEL_INVALID_SOLUTION_HINT = 0x54
    MPSOLVER_MODEL_INVALID_SOLVER_PARAMETERS = 0x55
    MPSOLVER_SOLVER_TYPE_UNAVAILABLE = 0x7
    MPSOLVER_INCOMPATIBLE_OPTIONS = 0x71


class SolverType:  # MPModelRequest.SolverType, linear_solver.proto:456-490 (the values this path can be asked for)
    CLP_LINEAR_PROGRAMMING = 0
    GLPK_LINEAR_PROGRAMMING = 1
    GLOP_LINEAR_PROGRAMMING = 2
    SCIP_MIXED_INTEGER_PROGRAMMING = 3
    GLPK_MIXED_INTEGER_PROGRAMMING = 4
    CBC_MIXED_INTEGER_PROGRAMMING = 5
    GUROBI_LINEAR_PROGRAMMING = 6
    GUROBI_MIXED_INTEGER_PROGRAMMING = 7
    PDLP_LINEAR_PROGRAMMING = 8
    CPLEX_LINEAR_PROGRAMMING = 10
    CPLEX_MIXED_INTEGER_PROGRAMMING = 11
    BOP_INTEGER_PROGRAMMING = 12
    KNAPSACK_MIXED_INTEGER_PROGRAMMING = 13
    SAT_INTEGER_PROGRAMMING = 14
    HIGHS_LINEAR_PROGRAMMING = 15
    HIGHS_MIXED_INTEGER_PROGRAMMING = 16
    XPRESS_LINEAR_PROGRAMMING = 101
    XPRESS_MIXED_INTEGER_PROGRAMMING = 102


def _add(msg, name, number, ftype, default=None, repeated=False, packed=False):
    f = msg.field.add()
    f.name, f.number = name, number
    f.label = _F.LABEL_REPEATED if repeated else _F.LABEL_OPTIONAL
    simple = {"double": _F.TYPE_DOUBLE, "int32": _F.TYPE_INT32, "bool": _F.TYPE_BOOL, "string": _F.TYPE_STRING, "bytes": _F.TYPE_BYTES}
    if ftype in simple:
        f.type = simple[ftype]
    elif ftype.startswith("enum:"):
        f.type, f.type_name = _F.TYPE_ENUM, "." + ftype[5:]
    else:
        f.type, f.type_name = _F.TYPE_MESSAGE, "." + ftype
    if default is not None:
        f.default_value = default
    if packed:
        f.options.packed = True


def _file():
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "ortools/linear_solver/linear_solver.proto"
    fd.package = _PKG
    fd.syntax = "proto2"
    v = fd.message_type.add()                                   # linear_solver.proto:49-73
    v.name = "MPVariableProto"
    _add(v, "lower_bound", 1, "double", "-inf")
    _add(v, "upper_bound", 2, "double", "inf")
    _add(v, "objective_coefficient", 3, "double", "0")
    _add(v, "is_integer", 4, "bool", "false")
    _add(v, "name", 5, "string", "")
    _add(v, "branching_priority", 6, "int32", "0")
    c = fd.message_type.add()                                   # linear_solver.proto:80-107
    c.name = "MPConstraintProto"
    _add(c, "var_index", 6, "int32", repeated=True, packed=True)
    _add(c, "coefficient", 7, "double", repeated=True, packed=True)
    _add(c, "lower_bound", 2, "double", "-inf")
    _add(c, "upper_bound", 3, "double", "inf")
    _add(c, "name", 4, "string", "")
    _add(c, "is_lazy", 5, "bool", "false")
    g = fd.message_type.add()                                   # only its presence matters here (quadratic_program.cc:101-103)
    g.name = "MPGeneralConstraintProto"
    q = fd.message_type.add()                                   # linear_solver.proto:216-233
    q.name = "MPQuadraticObjective"
    _add(q, "qvar1_index", 1, "int32", repeated=True)
    _add(q, "qvar2_index", 2, "int32", repeated=True)
    _add(q, "coefficient", 3, "double", repeated=True)
    m = fd.message_type.add()                                   # linear_solver.proto:263-317
    m.name = "MPModelProto"
    _add(m, "variable", 3, _PKG + ".MPVariableProto", repeated=True)
    _add(m, "constraint", 4, _PKG + ".MPConstraintProto", repeated=True)
    _add(m, "general_constraint", 7, _PKG + ".MPGeneralConstraintProto", repeated=True)
    _add(m, "maximize", 1, "bool", "false")
    _add(m, "objective_offset", 2, "double", "0")
    _add(m, "quadratic_objective", 8, _PKG + ".MPQuadraticObjective")
    _add(m, "name", 5, "string", "")
    r = fd.message_type.add()                                   # linear_solver.proto:444-516
    r.name = "MPModelRequest"
    e = r.enum_type.add()
    e.name = "SolverType"
    for k, val in sorted(((k, val) for k, val in vars(SolverType).items() if not k.startswith("_")), key=lambda kv: kv[1]):
        ev = e.value.add()
        ev.name, ev.number = k, val
    _add(r, "model", 1, _PKG + ".MPModelProto")
    _add(r, "solver_type", 2, "enum:%s.MPModelRequest.SolverType" % _PKG, "GLOP_LINEAR_PROGRAMMING")
    _add(r, "solver_time_limit_seconds", 3, "double")
    _add(r, "enable_internal_solver_output", 4, "bool", "false")
    _add(r, "solver_specific_parameters", 5, "string")
    _add(r, "ignore_solver_specific_parameters_failure", 9, "bool", "false")
    st = fd.enum_type.add()
    st.name = "MPSolverResponseStatus"
    for k, val in sorted(((k, val) for k, val in vars(MPSolverResponseStatus).items() if not k.startswith("_")), key=lambda kv: kv[1]):
        ev = st.value.add()
        ev.name, ev.number = k, val
    s = fd.message_type.add()                                   # linear_solver.proto:600-672
    s.name = "MPSolutionResponse"
    _add(s, "status", 1, "enum:%s.MPSolverResponseStatus" % _PKG, "MPSOLVER_UNKNOWN_STATUS")
    _add(s, "status_str", 7, "string")
    _add(s, "objective_value", 2, "double")
    _add(s, "best_objective_bound", 5, "double")
    _add(s, "variable_value", 3, "double", repeated=True, packed=True)
    _add(s, "dual_value", 4, "double", repeated=True, packed=True)
    _add(s, "reduced_cost", 6, "double", repeated=True, packed=True)
    _add(s, "solver_specific_info", 11, "bytes")
    return fd


_pool = descriptor_pool.DescriptorPool()
_pool.Add(_file())


def _cls(name):
    return message_factory.GetMessageClass(_pool.FindMessageTypeByName(_PKG + "." + name))


MPModelProto = _cls("MPModelProto")
MPModelRequestProto = _cls("MPModelRequest")
MPSolutionResponseProto = _cls("MPSolutionResponse")


class InvalidArgument(ValueError):
    """absl::InvalidArgumentError of the reference's converters."""


def qp_from_mp_model_proto(proto, relax_integer_variables, include_names=False):
    """QpFromMpModelProto, quadratic_program.cc:98-211. Maximisation problems are
    turned into minimisation (objective negated, objective_scaling_factor = -1)."""
    if len(proto.general_constraint) > 0:
        raise InvalidArgument("General constraints are not supported.")
    n, m = len(proto.variable), len(proto.constraint)
    qp = pdlp.QuadraticProgram(n, m)
    if include_names:
        qp.problem_name = proto.name
        qp.variable_names = [v.name for v in proto.variable]
        qp.constraint_names = [c.name for c in proto.constraint]
    for i, var in enumerate(proto.variable):
        qp.variable_lower_bounds[i] = var.lower_bound
        qp.variable_upper_bounds[i] = var.upper_bound
        qp.objective_vector[i] = var.objective_coefficient
        if var.is_integer and not relax_integer_variables:
            raise InvalidArgument("Integer variable encountered with relax_integer_variables == false")
    rows, cols, vals = [], [], []
    for i, con in enumerate(proto.constraint):
        if len(con.var_index) != len(con.coefficient):
            raise InvalidArgument("%dth constraint has %d coefficients, expected %d" % (i, len(con.coefficient), len(con.var_index)))
        for j, vi in enumerate(con.var_index):
            if vi < 0 or vi >= n:
                raise InvalidArgument("Variable index of %dth constraint's %dth nonzero is %d which is not in the allowed range [0, %d)" % (i, j, vi, n))
        rows.extend([i] * len(con.var_index))
        cols.extend(con.var_index)
        vals.extend(con.coefficient)
        qp.constraint_lower_bounds[i] = con.lower_bound
        qp.constraint_upper_bounds[i] = con.upper_bound
    k = sp.csc_matrix((np.asarray(vals, dtype=np.float64), (np.asarray(rows, dtype=np.int64), np.asarray(cols, dtype=np.int64))), shape=(m, n))
    k.sum_duplicates()
    k.sort_indices()
    qp.constraint_matrix = k
    quad = proto.quadratic_objective
    if not (len(quad.qvar1_index) == len(quad.qvar2_index) == len(quad.coefficient)):
        raise InvalidArgument("The quadratic objective has %d qvar1_indices, %d qvar2_indices, and %d coefficients, expected equal numbers." % (
            len(quad.qvar1_index), len(quad.qvar2_index), len(quad.coefficient)))
    if len(quad.qvar1_index) > 0:
        qp.objective_matrix = np.zeros(n)
    for i, (i1, i2, coef) in enumerate(zip(quad.qvar1_index, quad.qvar2_index, quad.coefficient)):
        if i1 < 0 or i2 < 0 or i1 >= n or i2 >= n:
            raise InvalidArgument("The quadratic objective's %dth nonzero has indices %d and %d, which are not both in the expected range [0, %d)" % (i, i1, i2, n))
        if i1 != i2:
            raise InvalidArgument("The quadratic objective's %dth nonzero has off-diagonal element at (%d, %d). Only diagonal objective matrices are supported." % (i, i1, i2))
        qp.objective_matrix[i1] = 2 * coef  # QuadraticProgram has an implicit 1/2 in front of the quadratic term
    qp.objective_offset = proto.objective_offset
    if proto.maximize:
        qp.objective_offset *= -1
        qp.objective_vector *= -1
        if qp.objective_matrix is not None:
            qp.objective_matrix *= -1
        qp.objective_scaling_factor = -1.0
    return qp


def can_fit_in_mp_model_proto(qp, largest_ok_size=2**31 - 1):
    """CanFitInMpModelProto / TestableCanFitInMpModelProto, quadratic_program.cc:213-234."""
    n, m = len(qp.variable_lower_bounds), len(qp.constraint_lower_bounds)
    if n > largest_ok_size:
        raise InvalidArgument("Too many variables (%d) to index with an int32_t." % n)
    if m > largest_ok_size:
        raise InvalidArgument("Too many constraints (%d) to index with an int32_t." % m)


def qp_to_mp_model_proto(qp):
    """QpToMpModelProto, quadratic_program.cc:236-315."""
    can_fit_in_mp_model_proto(qp)
    if qp.objective_scaling_factor == 0:
        raise InvalidArgument("objective_scaling_factor cannot be zero.")
    proto = MPModelProto()
    if qp.problem_name:
        proto.name = qp.problem_name
    s = qp.objective_scaling_factor
    proto.objective_offset = s * qp.objective_offset
    proto.maximize = bool(s < 0)
    for i in range(len(qp.variable_lower_bounds)):
        var = proto.variable.add()
        var.lower_bound = float(qp.variable_lower_bounds[i])
        var.upper_bound = float(qp.variable_upper_bounds[i])
        var.objective_coefficient = float(s * qp.objective_vector[i])
        if qp.variable_names is not None and i < len(qp.variable_names) and qp.variable_names[i]:
            var.name = qp.variable_names[i]
    for i in range(len(qp.constraint_lower_bounds)):
        con = proto.constraint.add()
        con.lower_bound = float(qp.constraint_lower_bounds[i])
        con.upper_bound = float(qp.constraint_upper_bounds[i])
        if qp.constraint_names is not None and i < len(qp.constraint_names) and qp.constraint_names[i]:
            con.name = qp.constraint_names[i]
    k = sp.csc_matrix(qp.constraint_matrix)
    k.sort_indices()
    for col in range(k.shape[1]):
        for p in range(k.indptr[col], k.indptr[col + 1]):
            con = proto.constraint[int(k.indices[p])]
            con.var_index.append(col)
            con.coefficient.append(float(k.data[p]))
    if not pdlp.is_linear_program(qp):
        for i, d in enumerate(qp.objective_matrix):
            if d != 0.0:
                proto.quadratic_objective.qvar1_index.append(i)
                proto.quadratic_objective.qvar2_index.append(i)
                proto.quadratic_objective.coefficient.append(float(s * d / 2.0))  # undo the implicit 1/2
    return proto


def get_convergence_information(stats, candidate_type):
    """GetConvergenceInformation, iteration_stats.cc:597-606."""
    if stats is None:
        return None
    for c in stats.convergence_information:
        if c.candidate_type == candidate_type:
            return c
    return None


def pdlp_solve_proto(request, relax_integer_variables=False, interrupt_solve=None, backend=None):
    """PdlpSolveProto, pdlp_proto_solver.cc:36-130: MPModelRequest -> MPSolutionResponse.
    `backend`: the PDLP implementation to call (default: the CUDA library)."""
    params_msg = pdlp_proto.PrimalDualHybridGradientParamsProto()
    params_msg.verbosity_level = 3 if request.enable_internal_solver_output else 0
    response = MPSolutionResponseProto()
    try:
        text_format.Merge(request.solver_specific_parameters, params_msg)
    except text_format.ParseError:
        response.status = MPSolverResponseStatus.MPSOLVER_MODEL_INVALID_SOLVER_PARAMETERS
        return response
    if interrupt_solve is not None and bool(interrupt_solve[0] if hasattr(interrupt_solve, "__getitem__") else interrupt_solve):
        response.status = MPSolverResponseStatus.MPSOLVER_NOT_SOLVED
        return response
    if request.HasField("solver_time_limit_seconds"):
        params_msg.termination_criteria.time_sec_limit = request.solver_time_limit_seconds
    if not request.HasField("model"):
        response.status = MPSolverResponseStatus.MPSOLVER_MODEL_INVALID
        response.status_str = "The request has no model."
        return response
    qp = qp_from_mp_model_proto(request.model, relax_integer_variables)
    params = pdlp_proto.params_from_proto(params_msg)
    scaling = qp.objective_scaling_factor
    be = backend if backend is not None else pdlp.backend()
    result = be.primal_dual_hybrid_gradient(qp, params, interrupt_solve=interrupt_solve) if interrupt_solve is not None \
        else be.primal_dual_hybrid_gradient(qp, params)
    tr = pdlp.TerminationReason
    reason = result.solve_log.termination_reason
    response.status = {
        tr.TERMINATION_REASON_OPTIMAL: MPSolverResponseStatus.MPSOLVER_OPTIMAL,
        tr.TERMINATION_REASON_NUMERICAL_ERROR: MPSolverResponseStatus.MPSOLVER_ABNORMAL,
        tr.TERMINATION_REASON_PRIMAL_INFEASIBLE: MPSolverResponseStatus.MPSOLVER_INFEASIBLE,
        tr.TERMINATION_REASON_INTERRUPTED_BY_USER: MPSolverResponseStatus.MPSOLVER_CANCELLED_BY_USER,
    }.get(reason, MPSolverResponseStatus.MPSOLVER_NOT_SOLVED)
    if result.solve_log.termination_string:
        response.status_str = result.solve_log.termination_string
    ci = get_convergence_information(result.solve_log.solution_stats, result.solve_log.solution_type)
    if ci is not None:
        response.objective_value = ci.primal_objective
    response.variable_value.extend(float(v) for v in result.primal_solution)
    # maximisation was turned into minimisation: duals and reduced costs change sign back
    response.dual_value.extend(float(scaling * v) for v in result.dual_solution)
    response.reduced_cost.extend(float(scaling * v) for v in result.reduced_costs)
    response.solver_specific_info = pdlp_proto.solve_log_to_proto(result.solve_log, params).SerializeToString()
    return response
