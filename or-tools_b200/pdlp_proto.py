"""Wire / text-format adapters for the parameter and log messages of PDLP
(SURVEY.md 8f rank 1).

The reference's callers hand ``PrimalDualHybridGradientParams`` to the solver
as a protobuf (text on the command line -- ``examples/cpp/pdlp_solve.cc:64-79``
-- or binary inside ``MPModelRequest.solver_specific_parameters`` --
``ortools/linear_solver/proto_solver/pdlp_proto_solver.cc:47``) and read a
``SolveLog`` protobuf back (``pdlp_proto_solver.cc:127``). There is no
``protoc`` in this image, so the two schemas (``ortools/pdlp/solvers.proto``,
``ortools/pdlp/solve_log.proto``: same package, message / field / enum names,
tags, types and defaults) are restated here as descriptor tables and turned
into real protobuf message classes at import time with the ``google.protobuf``
runtime. Bytes and text produced here parse with the reference's generated
classes and vice versa; the known-answer encodings in
``tests/test_proto_io.py`` pin the tags.

    params = pdlp_proto.params_from_text('termination_criteria { simple_optimality_criteria { eps_optimal_absolute: 1e-4 } }')
    result = pdlp.primal_dual_hybrid_gradient(qp, params)
    blob = pdlp_proto.solve_log_to_proto(result.solve_log, params).SerializeToString()
"""
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory, text_format

from . import pdlp

_F = descriptor_pb2.FieldDescriptorProto
_PKG = "operations_research.pdlp"
_T = {"double": _F.TYPE_DOUBLE, "int32": _F.TYPE_INT32, "int64": _F.TYPE_INT64, "bool": _F.TYPE_BOOL, "string": _F.TYPE_STRING}


def _enum(container, name, cls):
    e = container.enum_type.add()
    e.name = name
    for k, v in sorted(((k, v) for k, v in vars(cls).items() if not k.startswith("_") and isinstance(v, int)), key=lambda kv: kv[1]):
        val = e.value.add()
        val.name, val.number = k, v
    return e


def _field(msg, name, number, ftype, default=None, repeated=False, packed=False, oneof=None):
    f = msg.field.add()
    f.name, f.number = name, number
    f.label = _F.LABEL_REPEATED if repeated else _F.LABEL_OPTIONAL
    if ftype in _T:
        f.type = _T[ftype]
    elif ftype.startswith("enum:"):
        f.type = _F.TYPE_ENUM
        f.type_name = "." + ftype[5:]
    else:
        f.type = _F.TYPE_MESSAGE
        f.type_name = "." + ftype
    if default is not None:
        if isinstance(default, bool):
            f.default_value = "true" if default else "false"
        elif isinstance(default, float):
            f.default_value = "inf" if default == float("inf") else repr(default)
        else:
            f.default_value = str(default)
    if packed:
        f.options.packed = True
    if oneof is not None:
        f.oneof_index = oneof
    return f


def _solvers_file():
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "ortools/pdlp/solvers.proto"
    fd.package = _PKG
    fd.syntax = "proto2"
    _enum(fd, "OptimalityNorm", pdlp.OptimalityNorm)      # solvers.proto:24-41
    _enum(fd, "SchedulerType", pdlp.SchedulerType)        # solvers.proto:44-51

    tc = fd.message_type.add()                            # solvers.proto:66-187
    tc.name = "TerminationCriteria"
    s = tc.nested_type.add()
    s.name = "SimpleOptimalityCriteria"
    _field(s, "eps_optimal_absolute", 1, "double", 1.0e-6)
    _field(s, "eps_optimal_relative", 2, "double", 1.0e-6)
    d = tc.nested_type.add()
    d.name = "DetailedOptimalityCriteria"
    for i, n in enumerate(["primal_residual_absolute", "primal_residual_relative", "dual_residual_absolute", "dual_residual_relative",
                           "objective_gap_absolute", "objective_gap_relative"]):
        _field(d, "eps_optimal_" + n, i + 1, "double", 1.0e-6)
    tc.oneof_decl.add().name = "optimality_criteria"
    _field(tc, "optimality_norm", 1, "enum:%s.OptimalityNorm" % _PKG, "OPTIMALITY_NORM_L2")
    _field(tc, "simple_optimality_criteria", 9, _PKG + ".TerminationCriteria.SimpleOptimalityCriteria", oneof=0)
    _field(tc, "detailed_optimality_criteria", 10, _PKG + ".TerminationCriteria.DetailedOptimalityCriteria", oneof=0)
    f = _field(tc, "eps_optimal_absolute", 2, "double", 1.0e-6)
    f.options.deprecated = True
    f = _field(tc, "eps_optimal_relative", 3, "double", 1.0e-6)
    f.options.deprecated = True
    _field(tc, "eps_primal_infeasible", 4, "double", 1.0e-8)
    _field(tc, "eps_dual_infeasible", 5, "double", 1.0e-8)
    _field(tc, "time_sec_limit", 6, "double", float("inf"))
    _field(tc, "iteration_limit", 7, "int32", 2147483647)
    _field(tc, "kkt_matrix_pass_limit", 8, "double", float("inf"))

    al = fd.message_type.add()                            # solvers.proto:189-204
    al.name = "AdaptiveLinesearchParams"
    _field(al, "step_size_reduction_exponent", 1, "double", 0.3)
    _field(al, "step_size_growth_exponent", 2, "double", 0.6)
    mp = fd.message_type.add()                            # solvers.proto:206-226
    mp.name = "MalitskyPockParams"
    _field(mp, "step_size_downscaling_factor", 1, "double", 0.7)
    _field(mp, "linesearch_contraction_factor", 2, "double", 0.99)
    _field(mp, "step_size_interpolation", 3, "double", 1.0)

    p = fd.message_type.add()                             # solvers.proto:238-497
    p.name = "PrimalDualHybridGradientParams"
    _enum(p, "RestartStrategy", pdlp.RestartStrategy)
    _enum(p, "LinesearchRule", pdlp.LinesearchRule)
    po = p.nested_type.add()                              # solvers.proto:366-384
    po.name = "PresolveOptions"
    _field(po, "use_glop", 1, "bool")
    # glop_parameters (tag 2, operations_research.glop.GlopParameters) is host-only
    # presolve configuration: carried as raw bytes in the unknown-field set.
    me = _PKG + ".PrimalDualHybridGradientParams"
    _field(p, "termination_criteria", 1, _PKG + ".TerminationCriteria")
    _field(p, "num_threads", 2, "int32", 1)
    _field(p, "num_shards", 27, "int32", 0)
    _field(p, "scheduler_type", 32, "enum:%s.SchedulerType" % _PKG, "SCHEDULER_TYPE_GOOGLE_THREADPOOL")
    _field(p, "record_iteration_stats", 3, "bool")
    _field(p, "verbosity_level", 26, "int32", 0)
    _field(p, "log_interval_seconds", 31, "double", 0.0)
    _field(p, "major_iteration_frequency", 4, "int32", 64)
    _field(p, "termination_check_frequency", 5, "int32", 64)
    _field(p, "restart_strategy", 6, "enum:%s.RestartStrategy" % me, "ADAPTIVE_HEURISTIC")
    _field(p, "primal_weight_update_smoothing", 7, "double", 0.5)
    _field(p, "initial_primal_weight", 8, "double")
    _field(p, "presolve_options", 16, me + ".PresolveOptions")
    _field(p, "l_inf_ruiz_iterations", 9, "int32", 5)
    _field(p, "l2_norm_rescaling", 10, "bool", True)
    _field(p, "sufficient_reduction_for_restart", 11, "double", 0.1)
    _field(p, "necessary_reduction_for_restart", 17, "double", 0.9)
    _field(p, "linesearch_rule", 12, "enum:%s.LinesearchRule" % me, "ADAPTIVE_LINESEARCH_RULE")
    _field(p, "adaptive_linesearch_parameters", 18, _PKG + ".AdaptiveLinesearchParams")
    _field(p, "malitsky_pock_parameters", 19, _PKG + ".MalitskyPockParams")
    _field(p, "initial_step_size_scaling", 25, "double", 1.0)
    _field(p, "random_projection_seeds", 28, "int32", repeated=True, packed=True)
    _field(p, "infinite_constraint_bound_threshold", 22, "double", float("inf"))
    _field(p, "handle_some_primal_gradients_on_finite_bounds_as_residuals", 29, "bool", True)
    _field(p, "use_diagonal_qp_trust_region_solver", 23, "bool", False)
    _field(p, "diagonal_qp_trust_region_solver_tolerance", 24, "double", 1.0e-8)
    _field(p, "use_feasibility_polishing", 30, "bool", False)
    _field(p, "apply_feasibility_polishing_after_limits_reached", 33, "bool", False)
    _field(p, "apply_feasibility_polishing_if_solver_is_interrupted", 34, "bool", False)
    for lo in (13, 14, 15, 20, 21):
        r = p.reserved_range.add()
        r.start, r.end = lo, lo + 1
    return fd


class _PolishingPhaseType:  # solve_log.proto:362-369
    POLISHING_PHASE_TYPE_UNSPECIFIED = 0
    POLISHING_PHASE_TYPE_PRIMAL_FEASIBILITY = 1
    POLISHING_PHASE_TYPE_DUAL_FEASIBILITY = 2


_STATS_TAGS = {  # solve_log.proto:28-102
    "num_variables": (1, "int64"), "num_constraints": (2, "int64"),
    "constraint_matrix_col_min_l_inf_norm": (3, "double"), "constraint_matrix_row_min_l_inf_norm": (4, "double"),
    "constraint_matrix_num_nonzeros": (5, "int64"),
    "constraint_matrix_abs_max": (6, "double"), "constraint_matrix_abs_min": (7, "double"), "constraint_matrix_abs_avg": (8, "double"),
    "constraint_matrix_l2_norm": (25, "double"),
    "combined_bounds_max": (9, "double"), "combined_bounds_min": (10, "double"), "combined_bounds_avg": (11, "double"),
    "combined_bounds_l2_norm": (24, "double"),
    "combined_variable_bounds_max": (28, "double"), "combined_variable_bounds_min": (29, "double"),
    "combined_variable_bounds_avg": (30, "double"), "combined_variable_bounds_l2_norm": (31, "double"),
    "variable_bound_gaps_num_finite": (12, "int64"),
    "variable_bound_gaps_max": (13, "double"), "variable_bound_gaps_min": (14, "double"), "variable_bound_gaps_avg": (15, "double"),
    "variable_bound_gaps_l2_norm": (26, "double"),
    "objective_vector_abs_max": (16, "double"), "objective_vector_abs_min": (17, "double"), "objective_vector_abs_avg": (18, "double"),
    "objective_vector_l2_norm": (23, "double"),
    "objective_matrix_num_nonzeros": (19, "int64"),
    "objective_matrix_abs_max": (20, "double"), "objective_matrix_abs_min": (21, "double"), "objective_matrix_abs_avg": (22, "double"),
    "objective_matrix_l2_norm": (27, "double"),
}
_CONVERGENCE_TAGS = {  # solve_log.proto:139-205
    "primal_objective": 2, "dual_objective": 3, "corrected_dual_objective": 4, "l_inf_primal_residual": 5, "l2_primal_residual": 6,
    "l_inf_componentwise_primal_residual": 24, "l_inf_dual_residual": 7, "l2_dual_residual": 8, "l_inf_componentwise_dual_residual": 25,
    "l_inf_primal_variable": 14, "l2_primal_variable": 15, "l_inf_dual_variable": 16, "l2_dual_variable": 17,
}
_INFEASIBILITY_TAGS = {  # solve_log.proto:209-249
    "max_primal_ray_infeasibility": 1, "primal_ray_linear_objective": 2, "primal_ray_quadratic_norm": 3,
    "max_dual_ray_infeasibility": 4, "dual_ray_objective": 5,
}


def _solve_log_file():
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "ortools/pdlp/solve_log.proto"
    fd.package = _PKG
    fd.syntax = "proto2"
    fd.dependency.append("ortools/pdlp/solvers.proto")
    qs = fd.message_type.add()
    qs.name = "QuadraticProgramStats"
    for name, (tag, t) in _STATS_TAGS.items():
        _field(qs, name, tag, t)
    _enum(fd, "RestartChoice", pdlp.RestartChoice)        # solve_log.proto:105-117
    _enum(fd, "PointType", pdlp.PointType)                # solve_log.proto:121-135
    ci = fd.message_type.add()
    ci.name = "ConvergenceInformation"
    _field(ci, "candidate_type", 1, "enum:%s.PointType" % _PKG)
    for name, tag in _CONVERGENCE_TAGS.items():
        _field(ci, name, tag, "double")
    ii = fd.message_type.add()
    ii.name = "InfeasibilityInformation"
    for name, tag in _INFEASIBILITY_TAGS.items():
        _field(ii, name, tag, "double")
    _field(ii, "candidate_type", 6, "enum:%s.PointType" % _PKG)
    pm = fd.message_type.add()                            # solve_log.proto:251-274
    pm.name = "PointMetadata"
    _field(pm, "point_type", 1, "enum:%s.PointType" % _PKG)
    _field(pm, "random_primal_projections", 2, "double", repeated=True, packed=True)
    _field(pm, "random_dual_projections", 3, "double", repeated=True, packed=True)
    _field(pm, "active_primal_variable_count", 4, "int64")
    _field(pm, "active_dual_variable_count", 5, "int64")
    _field(pm, "active_primal_variable_change", 6, "int64")
    _field(pm, "active_dual_variable_change", 7, "int64")
    it = fd.message_type.add()                            # solve_log.proto:281-334
    it.name = "IterationStats"
    _field(it, "iteration_number", 1, "int32")
    _field(it, "convergence_information", 2, _PKG + ".ConvergenceInformation", repeated=True)
    _field(it, "infeasibility_information", 3, _PKG + ".InfeasibilityInformation", repeated=True)
    _field(it, "point_metadata", 11, _PKG + ".PointMetadata", repeated=True)
    _field(it, "cumulative_kkt_matrix_passes", 4, "double")
    _field(it, "cumulative_rejected_steps", 5, "int32")
    _field(it, "cumulative_time_sec", 6, "double")
    _field(it, "restart_used", 7, "enum:%s.RestartChoice" % _PKG)
    _field(it, "step_size", 8, "double")
    _field(it, "primal_weight", 9, "double")
    _enum(fd, "TerminationReason", pdlp.TerminationReason)  # solve_log.proto:336-360
    _enum(fd, "PolishingPhaseType", _PolishingPhaseType)
    fp = fd.message_type.add()                            # solve_log.proto:371-383
    fp.name = "FeasibilityPolishingDetails"
    _field(fp, "polishing_phase_type", 1, "enum:%s.PolishingPhaseType" % _PKG)
    _field(fp, "main_iteration_count", 2, "int32")
    _field(fp, "params", 3, _PKG + ".PrimalDualHybridGradientParams")
    _field(fp, "termination_reason", 4, "enum:%s.TerminationReason" % _PKG)
    _field(fp, "iteration_count", 5, "int32")
    _field(fp, "solve_time_sec", 6, "double")
    _field(fp, "solution_stats", 7, _PKG + ".IterationStats")
    _field(fp, "solution_type", 8, "enum:%s.PointType" % _PKG)
    _field(fp, "iteration_stats", 9, _PKG + ".IterationStats", repeated=True)
    sl = fd.message_type.add()                            # solve_log.proto:385-459
    sl.name = "SolveLog"
    _field(sl, "instance_name", 1, "string")
    _field(sl, "params", 14, _PKG + ".PrimalDualHybridGradientParams")
    _field(sl, "termination_reason", 3, "enum:%s.TerminationReason" % _PKG)
    _field(sl, "termination_string", 4, "string")
    _field(sl, "iteration_count", 5, "int32")
    _field(sl, "preprocessing_time_sec", 13, "double")
    _field(sl, "solve_time_sec", 6, "double")
    _field(sl, "solution_stats", 8, _PKG + ".IterationStats")
    _field(sl, "solution_type", 10, "enum:%s.PointType" % _PKG)
    _field(sl, "iteration_stats", 7, _PKG + ".IterationStats", repeated=True)
    _field(sl, "original_problem_stats", 11, _PKG + ".QuadraticProgramStats")
    _field(sl, "preprocessed_problem_stats", 12, _PKG + ".QuadraticProgramStats")
    _field(sl, "feasibility_polishing_details", 15, _PKG + ".FeasibilityPolishingDetails", repeated=True)
    for lo in (2, 9):
        r = sl.reserved_range.add()
        r.start, r.end = lo, lo + 1
    return fd


_pool = descriptor_pool.DescriptorPool()
_pool.Add(_solvers_file())
_pool.Add(_solve_log_file())


def _cls(name):
    return message_factory.GetMessageClass(_pool.FindMessageTypeByName(_PKG + "." + name))


TerminationCriteriaProto = _cls("TerminationCriteria")
PrimalDualHybridGradientParamsProto = _cls("PrimalDualHybridGradientParams")
QuadraticProgramStatsProto = _cls("QuadraticProgramStats")
ConvergenceInformationProto = _cls("ConvergenceInformation")
InfeasibilityInformationProto = _cls("InfeasibilityInformation")
PointMetadataProto = _cls("PointMetadata")
IterationStatsProto = _cls("IterationStats")
SolveLogProto = _cls("SolveLog")


# --------------------------------------------------------------------------
# params: proto <-> the solver's parameter object
# --------------------------------------------------------------------------
def _is_repeated(fdesc):
    return fdesc.is_repeated if hasattr(fdesc, "is_repeated") else fdesc.label == fdesc.LABEL_REPEATED


def _copy_set_fields(src, dst):
    """Copies every field that is present in protobuf message `src` into the
    look-alike `dst` (pdlp._Message), recursing into sub-messages, so presence
    (HasField) carries over -- validation depends on it (solvers.proto:81-96)."""
    for fdesc, value in src.ListFields():
        if fdesc.type == fdesc.TYPE_MESSAGE:
            sub = getattr(dst, fdesc.name)
            _copy_set_fields(value, sub)
            if fdesc.containing_oneof is not None:
                dst._touch(fdesc.name)
        elif _is_repeated(fdesc):
            setattr(dst, fdesc.name, tuple(value))
        else:
            setattr(dst, fdesc.name, value)


def params_from_proto(msg):
    """PrimalDualHybridGradientParams protobuf -> pdlp.PrimalDualHybridGradientParams."""
    out = pdlp.PrimalDualHybridGradientParams()
    _copy_set_fields(msg, out)
    return out


def params_from_text(text):
    """Text-format PrimalDualHybridGradientParams (examples/cpp/pdlp_solve.cc --params)."""
    return params_from_proto(text_format.Parse(text, PrimalDualHybridGradientParamsProto()))


def params_from_bytes(blob):
    """Binary PrimalDualHybridGradientParams (MPModelRequest.solver_specific_parameters)."""
    msg = PrimalDualHybridGradientParamsProto()
    msg.ParseFromString(blob)
    return params_from_proto(msg)


def _fill_set_fields(src, dst):
    for name, value in src._set.items():
        if isinstance(value, (tuple, list)):
            getattr(dst, name).extend(int(v) for v in value)
        else:
            setattr(dst, name, value)
    for name in src._submessages:
        sub = getattr(src, name)
        if sub._any_set() or getattr(src, "_oneof", None) == name:
            getattr(dst, name).SetInParent()
            _fill_set_fields(sub, getattr(dst, name))


def params_to_proto(params):
    """pdlp.PrimalDualHybridGradientParams -> protobuf (only the fields that were set)."""
    msg = PrimalDualHybridGradientParamsProto()
    _fill_set_fields(params, msg)
    return msg


# --------------------------------------------------------------------------
# SolveLog -> proto
# --------------------------------------------------------------------------
def _stats_to_proto(s, out):
    for name in _STATS_TAGS:
        if hasattr(s, name):
            setattr(out, name, getattr(s, name))


def _iteration_stats_to_proto(s, out):
    out.iteration_number = int(s.iteration_number)
    for c in s.convergence_information:
        m = out.convergence_information.add()
        m.candidate_type = int(c.candidate_type)
        for name in _CONVERGENCE_TAGS:
            setattr(m, name, getattr(c, name))
    for c in s.infeasibility_information:
        m = out.infeasibility_information.add()
        m.candidate_type = int(c.candidate_type)
        for name in _INFEASIBILITY_TAGS:
            setattr(m, name, getattr(c, name))
    for p in s.point_metadata:
        m = out.point_metadata.add()
        m.point_type = int(p.point_type)
        m.random_primal_projections.extend(p.random_primal_projections)
        m.random_dual_projections.extend(p.random_dual_projections)
        if getattr(p, "has_active_set_information", 1):
            for name in ("active_primal_variable_count", "active_dual_variable_count", "active_primal_variable_change", "active_dual_variable_change"):
                setattr(m, name, int(getattr(p, name)))
    out.cumulative_kkt_matrix_passes = s.cumulative_kkt_matrix_passes
    out.cumulative_rejected_steps = int(s.cumulative_rejected_steps)
    out.cumulative_time_sec = s.cumulative_time_sec
    out.restart_used = int(s.restart_used)
    out.step_size = s.step_size
    out.primal_weight = s.primal_weight


def solve_log_to_proto(solve_log, params=None):
    """SolverResult.solve_log -> SolveLog protobuf (what pdlp_proto_solver.cc:127
    serialises into MPSolutionResponse.solver_specific_info and pdlp_solve.cc
    writes to --solve_log_file). `params`: the parameter object of the solve,
    echoed in SolveLog.params like primal_dual_hybrid_gradient.cc:1053."""
    out = SolveLogProto()
    if solve_log.instance_name:
        out.instance_name = solve_log.instance_name
    if params is not None:
        out.params.CopyFrom(params_to_proto(params))
    out.termination_reason = int(solve_log.termination_reason)
    if solve_log.termination_string:
        out.termination_string = solve_log.termination_string
    out.iteration_count = int(solve_log.iteration_count)
    out.preprocessing_time_sec = solve_log.preprocessing_time_sec
    out.solve_time_sec = solve_log.solve_time_sec
    if solve_log.solution_stats is not None:
        _iteration_stats_to_proto(solve_log.solution_stats, out.solution_stats)
    out.solution_type = int(solve_log.solution_type)
    for s in solve_log.iteration_stats:
        _iteration_stats_to_proto(s, out.iteration_stats.add())
    if solve_log.original_problem_stats is not None:
        _stats_to_proto(solve_log.original_problem_stats, out.original_problem_stats)
    if solve_log.preprocessed_problem_stats is not None:
        _stats_to_proto(solve_log.preprocessed_problem_stats, out.preprocessed_problem_stats)
    for d in getattr(solve_log, "feasibility_polishing_details", []):
        m = out.feasibility_polishing_details.add()
        m.polishing_phase_type = int(d.polishing_phase_type)
        m.main_iteration_count = int(d.main_iteration_count)
        m.termination_reason = int(d.termination_reason)
        m.iteration_count = int(d.iteration_count)
        m.solve_time_sec = d.solve_time_sec
        _iteration_stats_to_proto(d.solution_stats, m.solution_stats)
        m.solution_type = int(d.solution_type)
        for s in d.iteration_stats:
            _iteration_stats_to_proto(s, m.iteration_stats.add())
    return out


def solve_log_to_text(solve_log, params=None):
    return text_format.MessageToString(solve_log_to_proto(solve_log, params))
