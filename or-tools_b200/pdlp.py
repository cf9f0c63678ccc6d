"""Host-side mirror of the reference's Python surface for PDLP.

Mirrors ``ortools/pdlp/python/pdlp.cc:39-151`` (``QuadraticProgram``,
``PrimalAndDualSolution``, ``SolverResult``, ``primal_dual_hybrid_gradient``)
and the parameter / log messages of ``ortools/pdlp/solvers.proto`` and
``ortools/pdlp/solve_log.proto`` (same field names, same enum numbers, proto2
presence for the optional fields that the validation logic distinguishes).

Everything funnels into the C ABI declared in ``include/pdlp_b200.h``. There is
no CPU fallback: if ``libpdlp_b200.so`` is missing or no CUDA device is usable
the calls raise.
"""
import ctypes as C
import math
import os
import types

import numpy as np

from . import _capi as capi

INF = float("inf")


# --------------------------------------------------------------------------
# enums (numeric values of the protos)
# --------------------------------------------------------------------------
class _Enum:
    @classmethod
    def Name(cls, value):
        for k, v in vars(cls).items():
            if not k.startswith("_") and isinstance(v, int) and v == value:
                return k
        return str(value)


class OptimalityNorm(_Enum):  # solvers.proto:24-41
    OPTIMALITY_NORM_UNSPECIFIED = 0
    OPTIMALITY_NORM_L_INF = 1
    OPTIMALITY_NORM_L2 = 2
    OPTIMALITY_NORM_L_INF_COMPONENTWISE = 3


class SchedulerType(_Enum):  # solvers.proto:44-51
    SCHEDULER_TYPE_UNSPECIFIED = 0
    SCHEDULER_TYPE_GOOGLE_THREADPOOL = 1
    SCHEDULER_TYPE_EIGEN_THREADPOOL = 3


class RestartStrategy(_Enum):  # solvers.proto:239-277
    RESTART_STRATEGY_UNSPECIFIED = 0
    NO_RESTARTS = 1
    EVERY_MAJOR_ITERATION = 2
    ADAPTIVE_HEURISTIC = 3
    ADAPTIVE_DISTANCE_BASED = 4


class LinesearchRule(_Enum):
    LINESEARCH_RULE_UNSPECIFIED = 0
    ADAPTIVE_LINESEARCH_RULE = 1
    MALITSKY_POCK_LINESEARCH_RULE = 2
    CONSTANT_STEP_SIZE_RULE = 3


class RestartChoice(_Enum):  # solve_log.proto:105-117
    RESTART_CHOICE_UNSPECIFIED = 0
    RESTART_CHOICE_NO_RESTART = 1
    RESTART_CHOICE_WEIGHTED_AVERAGE_RESET = 2
    RESTART_CHOICE_RESTART_TO_AVERAGE = 3


class PointType(_Enum):  # solve_log.proto:121-135
    POINT_TYPE_UNSPECIFIED = 0
    POINT_TYPE_CURRENT_ITERATE = 1
    POINT_TYPE_ITERATE_DIFFERENCE = 2
    POINT_TYPE_AVERAGE_ITERATE = 3
    POINT_TYPE_NONE = 4
    POINT_TYPE_PRESOLVER_SOLUTION = 5
    POINT_TYPE_FEASIBILITY_POLISHING_SOLUTION = 6


class TerminationReason(_Enum):  # solve_log.proto:336-360
    TERMINATION_REASON_UNSPECIFIED = 0
    TERMINATION_REASON_OPTIMAL = 1
    TERMINATION_REASON_PRIMAL_INFEASIBLE = 2
    TERMINATION_REASON_DUAL_INFEASIBLE = 3
    TERMINATION_REASON_TIME_LIMIT = 4
    TERMINATION_REASON_ITERATION_LIMIT = 5
    TERMINATION_REASON_NUMERICAL_ERROR = 6
    TERMINATION_REASON_OTHER = 7
    TERMINATION_REASON_KKT_MATRIX_PASS_LIMIT = 8
    TERMINATION_REASON_INVALID_PROBLEM = 9
    TERMINATION_REASON_INVALID_PARAMETER = 10
    TERMINATION_REASON_PRIMAL_OR_DUAL_INFEASIBLE = 11
    TERMINATION_REASON_INTERRUPTED_BY_USER = 12
    TERMINATION_REASON_INVALID_INITIAL_SOLUTION = 13


class PolishingPhaseType(_Enum):  # solve_log.proto:362-369
    POLISHING_PHASE_TYPE_UNSPECIFIED = 0
    POLISHING_PHASE_TYPE_PRIMAL_FEASIBILITY = 1
    POLISHING_PHASE_TYPE_DUAL_FEASIBILITY = 2


class IterationType(_Enum):  # primal_dual_hybrid_gradient.h:75-88
    NORMAL = 0
    PRIMAL_FEASIBILITY = 1
    DUAL_FEASIBILITY = 2
    PRESOLVE_TERMINATION = 3
    NORMAL_TERMINATION = 4
    FEASIBILITY_POLISHING_TERMINATION = 5


# --------------------------------------------------------------------------
# parameter messages (proto2-like: defaults + presence)
# --------------------------------------------------------------------------
class _Message:
    """Tiny proto2 look-alike: scalar fields with defaults and HasField()."""

    _defaults = {}
    _submessages = {}

    def __init__(self, **kwargs):
        object.__setattr__(self, "_set", {})
        for name, cls in self._submessages.items():
            object.__setattr__(self, name, cls())
        for k, v in kwargs.items():
            setattr(self, k, v)

    def __setattr__(self, name, value):
        if name in self._submessages:
            if not isinstance(value, self._submessages[name]):
                raise TypeError(f"{name} must be a {self._submessages[name].__name__}")
            object.__setattr__(self, name, value)
            self._touch(name)
            return
        if name not in self._defaults:
            raise AttributeError(f"{type(self).__name__} has no field {name!r}")
        self._set[name] = value

    def __getattr__(self, name):
        d = type(self)._defaults
        if name in d:
            return self.__dict__["_set"].get(name, d[name])
        raise AttributeError(name)

    def _touch(self, name):
        pass

    def HasField(self, name):
        if name in self._submessages:
            return getattr(self, name)._any_set()
        return name in self._set

    def ClearField(self, name):
        if name in self._submessages:
            object.__setattr__(self, name, self._submessages[name]())
        else:
            self._set.pop(name, None)

    def _any_set(self):
        return bool(self._set) or any(getattr(self, n)._any_set() for n in self._submessages)

    def __repr__(self):
        parts = [f"{k}={v!r}" for k, v in self._set.items()]
        parts += [f"{n}={getattr(self, n)!r}" for n in self._submessages if getattr(self, n)._any_set()]
        return f"{type(self).__name__}({', '.join(parts)})"


class SimpleOptimalityCriteria(_Message):  # solvers.proto:113-120
    _defaults = {"eps_optimal_absolute": 1e-6, "eps_optimal_relative": 1e-6}


class DetailedOptimalityCriteria(_Message):  # solvers.proto:122-160
    _defaults = {
        "eps_optimal_primal_residual_absolute": 1e-6, "eps_optimal_primal_residual_relative": 1e-6,
        "eps_optimal_dual_residual_absolute": 1e-6, "eps_optimal_dual_residual_relative": 1e-6,
        "eps_optimal_objective_gap_absolute": 1e-6, "eps_optimal_objective_gap_relative": 1e-6,
    }


class TerminationCriteria(_Message):  # solvers.proto:66-187
    _defaults = {
        "optimality_norm": OptimalityNorm.OPTIMALITY_NORM_L2,
        "eps_optimal_absolute": 1e-6, "eps_optimal_relative": 1e-6,
        "eps_primal_infeasible": 1e-8, "eps_dual_infeasible": 1e-8,
        "time_sec_limit": INF, "iteration_limit": 2147483647, "kkt_matrix_pass_limit": INF,
    }
    _submessages = {"simple_optimality_criteria": SimpleOptimalityCriteria,
                    "detailed_optimality_criteria": DetailedOptimalityCriteria}

    def __init__(self, **kwargs):
        object.__setattr__(self, "_oneof", None)
        super().__init__(**kwargs)

    def _touch(self, name):
        object.__setattr__(self, "_oneof", name)

    def WhichOneof(self, _name="optimality_criteria"):
        # A sub-message counts as selected once assigned or once one of its
        # fields was written (python proto semantics).
        if self._oneof is not None:
            return self._oneof
        for n in ("detailed_optimality_criteria", "simple_optimality_criteria"):
            if getattr(self, n)._any_set():
                return n
        return None

    def to_pod(self):
        t = capi.PdlpTerminationCriteria()
        t.optimality_norm = int(self.optimality_norm)
        which = self.WhichOneof()
        t.optimality_criteria_case = {None: 0, "simple_optimality_criteria": 9, "detailed_optimality_criteria": 10}[which]
        s, d = self.simple_optimality_criteria, self.detailed_optimality_criteria
        t.simple_eps_optimal_absolute = s.eps_optimal_absolute
        t.simple_eps_optimal_relative = s.eps_optimal_relative
        for f in DetailedOptimalityCriteria._defaults:
            setattr(t, f, getattr(d, f))
        t.has_eps_optimal_absolute = int(self.HasField("eps_optimal_absolute"))
        t.has_eps_optimal_relative = int(self.HasField("eps_optimal_relative"))
        t.eps_optimal_absolute = self.eps_optimal_absolute
        t.eps_optimal_relative = self.eps_optimal_relative
        t.eps_primal_infeasible = self.eps_primal_infeasible
        t.eps_dual_infeasible = self.eps_dual_infeasible
        t.time_sec_limit = self.time_sec_limit
        t.iteration_limit = int(self.iteration_limit)
        t.kkt_matrix_pass_limit = self.kkt_matrix_pass_limit
        return t


class AdaptiveLinesearchParams(_Message):  # solvers.proto:189-204
    _defaults = {"step_size_reduction_exponent": 0.3, "step_size_growth_exponent": 0.6}


class MalitskyPockParams(_Message):  # solvers.proto:206-226
    _defaults = {"step_size_downscaling_factor": 0.7, "linesearch_contraction_factor": 0.99,
                 "step_size_interpolation": 1.0}


class PresolveOptions(_Message):  # solvers.proto:355-384
    _defaults = {"use_glop": False}


class PrimalDualHybridGradientParams(_Message):  # solvers.proto:238-497
    RestartStrategy = RestartStrategy
    LinesearchRule = LinesearchRule
    NO_RESTARTS = RestartStrategy.NO_RESTARTS
    EVERY_MAJOR_ITERATION = RestartStrategy.EVERY_MAJOR_ITERATION
    ADAPTIVE_HEURISTIC = RestartStrategy.ADAPTIVE_HEURISTIC
    ADAPTIVE_DISTANCE_BASED = RestartStrategy.ADAPTIVE_DISTANCE_BASED
    ADAPTIVE_LINESEARCH_RULE = LinesearchRule.ADAPTIVE_LINESEARCH_RULE
    MALITSKY_POCK_LINESEARCH_RULE = LinesearchRule.MALITSKY_POCK_LINESEARCH_RULE
    CONSTANT_STEP_SIZE_RULE = LinesearchRule.CONSTANT_STEP_SIZE_RULE

    _defaults = {
        "num_threads": 1, "num_shards": 0,
        "scheduler_type": SchedulerType.SCHEDULER_TYPE_GOOGLE_THREADPOOL,
        "record_iteration_stats": False, "verbosity_level": 0, "log_interval_seconds": 0.0,
        "major_iteration_frequency": 64, "termination_check_frequency": 64,
        "restart_strategy": RestartStrategy.ADAPTIVE_HEURISTIC,
        "primal_weight_update_smoothing": 0.5, "initial_primal_weight": 0.0,
        "l_inf_ruiz_iterations": 5, "l2_norm_rescaling": True,
        "sufficient_reduction_for_restart": 0.1, "necessary_reduction_for_restart": 0.9,
        "linesearch_rule": LinesearchRule.ADAPTIVE_LINESEARCH_RULE,
        "initial_step_size_scaling": 1.0, "infinite_constraint_bound_threshold": INF,
        "handle_some_primal_gradients_on_finite_bounds_as_residuals": True,
        "use_diagonal_qp_trust_region_solver": False,
        "diagonal_qp_trust_region_solver_tolerance": 1e-8,
        "random_projection_seeds": (),
        "use_feasibility_polishing": False,
        "apply_feasibility_polishing_after_limits_reached": False,
        "apply_feasibility_polishing_if_solver_is_interrupted": False,
    }
    _submessages = {"termination_criteria": TerminationCriteria,
                    "adaptive_linesearch_parameters": AdaptiveLinesearchParams,
                    "malitsky_pock_parameters": MalitskyPockParams,
                    "presolve_options": PresolveOptions}

    def to_pod(self):
        p = capi.PdlpParams()
        p.termination_criteria = self.termination_criteria.to_pod()
        for f in ("num_threads", "num_shards", "scheduler_type", "verbosity_level", "major_iteration_frequency",
                  "termination_check_frequency", "restart_strategy", "l_inf_ruiz_iterations", "linesearch_rule"):
            setattr(p, f, int(getattr(self, f)))
        for f in ("record_iteration_stats", "l2_norm_rescaling",
                  "handle_some_primal_gradients_on_finite_bounds_as_residuals",
                  "use_diagonal_qp_trust_region_solver", "use_feasibility_polishing",
                  "apply_feasibility_polishing_after_limits_reached",
                  "apply_feasibility_polishing_if_solver_is_interrupted"):
            setattr(p, f, int(bool(getattr(self, f))))
        for f in ("log_interval_seconds", "primal_weight_update_smoothing", "sufficient_reduction_for_restart",
                  "necessary_reduction_for_restart", "initial_step_size_scaling",
                  "infinite_constraint_bound_threshold", "diagonal_qp_trust_region_solver_tolerance"):
            setattr(p, f, float(getattr(self, f)))
        p.has_initial_primal_weight = int(self.HasField("initial_primal_weight"))
        p.initial_primal_weight = float(self.initial_primal_weight)
        a, mp = self.adaptive_linesearch_parameters, self.malitsky_pock_parameters
        p.adaptive_step_size_reduction_exponent = a.step_size_reduction_exponent
        p.adaptive_step_size_growth_exponent = a.step_size_growth_exponent
        p.malitsky_pock_step_size_downscaling_factor = mp.step_size_downscaling_factor
        p.malitsky_pock_linesearch_contraction_factor = mp.linesearch_contraction_factor
        p.malitsky_pock_step_size_interpolation = mp.step_size_interpolation
        seeds = list(self.random_projection_seeds)
        p.num_random_projection_seeds = len(seeds)
        for i, s in enumerate(seeds[: capi.MAX_SEEDS]):
            p.random_projection_seeds[i] = int(s)
        p.presolve_use_glop = int(bool(self.presolve_options.use_glop))
        return p


def params_to_pod(params):
    """The C-ABI parameter POD of `params`: a PrimalDualHybridGradientParams of this module, a
    protobuf PrimalDualHybridGradientParams message (what the reference's wrapper takes,
    python/pdlp.cc:143-150; see pdlp_proto), or an already built POD."""
    if hasattr(params, "to_pod"):
        return params.to_pod()
    if hasattr(params, "ListFields") and hasattr(params, "DESCRIPTOR"):  # a protobuf message
        from . import pdlp_proto
        return pdlp_proto.params_from_proto(params).to_pod()
    return params


# --------------------------------------------------------------------------
# QuadraticProgram (quadratic_program.h:61-151; python/pdlp.cc:48-85)
# --------------------------------------------------------------------------
class QuadraticProgram:
    def __init__(self, num_variables=0, num_constraints=0):
        self.resize_and_initialize(num_variables, num_constraints)

    def resize_and_initialize(self, num_variables, num_constraints):
        import scipy.sparse as sp

        self.objective_vector = np.zeros(num_variables)
        self.objective_matrix = None  # diagonal as a vector, or None for an LP
        self.constraint_matrix = sp.csc_matrix((num_constraints, num_variables), dtype=np.float64)
        self.constraint_lower_bounds = np.full(num_constraints, -INF)
        self.constraint_upper_bounds = np.full(num_constraints, INF)
        self.variable_lower_bounds = np.full(num_variables, -INF)
        self.variable_upper_bounds = np.full(num_variables, INF)
        self.problem_name = None
        self.variable_names = None
        self.constraint_names = None
        self.objective_offset = 0.0
        self.objective_scaling_factor = 1.0

    def set_objective_matrix_diagonal(self, diagonal):
        self.objective_matrix = np.asarray(diagonal, dtype=np.float64).copy()

    def clear_objective_matrix(self):
        self.objective_matrix = None

    def apply_objective_scaling_and_offset(self, objective):
        return self.objective_scaling_factor * (objective + self.objective_offset)

    # -- C view -----------------------------------------------------------
    def _to_view(self):
        """Returns (PdlpProblemView, keepalive) with K as int64 CSC arrays."""
        import scipy.sparse as sp

        k = self.constraint_matrix
        if not sp.issparse(k):
            k = sp.csc_matrix(np.asarray(k, dtype=np.float64))
        k = k.tocsc()
        if not k.has_sorted_indices:
            k = k.sorted_indices()
        m, n = k.shape
        keep = {
            "col_starts": np.ascontiguousarray(k.indptr, dtype=np.int64),
            "row_indices": np.ascontiguousarray(k.indices, dtype=np.int64),
            "values": capi.as_f64(k.data),
            "c": capi.as_f64(self.objective_vector),
            "q": None if self.objective_matrix is None else capi.as_f64(self.objective_matrix),
            "lc": capi.as_f64(self.constraint_lower_bounds), "uc": capi.as_f64(self.constraint_upper_bounds),
            "lv": capi.as_f64(self.variable_lower_bounds), "uv": capi.as_f64(self.variable_upper_bounds),
            "name": None if self.problem_name is None else str(self.problem_name).encode(),
        }
        v = capi.PdlpProblemView()
        v.num_variables, v.num_constraints, v.num_nonzeros = n, m, k.nnz
        v.col_starts = capi.ptr_i64(keep["col_starts"])
        v.row_indices = capi.ptr_i64(keep["row_indices"])
        v.values = capi.ptr_f64(keep["values"])
        v.objective_vector = capi.ptr_f64(keep["c"])
        v.objective_matrix_diagonal = capi.ptr_f64(keep["q"])
        v.constraint_lower_bounds = capi.ptr_f64(keep["lc"])
        v.constraint_upper_bounds = capi.ptr_f64(keep["uc"])
        v.variable_lower_bounds = capi.ptr_f64(keep["lv"])
        v.variable_upper_bounds = capi.ptr_f64(keep["uv"])
        v.objective_offset = float(self.objective_offset)
        v.objective_scaling_factor = float(self.objective_scaling_factor)
        v.problem_name = keep["name"]
        v.objective_vector_size = keep["c"].size
        v.objective_matrix_size = -1 if keep["q"] is None else keep["q"].size
        v.constraint_lower_bounds_size = keep["lc"].size
        v.constraint_upper_bounds_size = keep["uc"].size
        v.variable_lower_bounds_size = keep["lv"].size
        v.variable_upper_bounds_size = keep["uv"].size
        return v, keep


def is_linear_program(qp):
    return qp.objective_matrix is None


def validate_quadratic_program_dimensions(qp):
    """quadratic_program.cc:38-97; raises ValueError like the pybind wrapper."""
    m, n = qp.constraint_matrix.shape
    var_lb, con_lb = len(qp.variable_lower_bounds), len(qp.constraint_lower_bounds)
    if var_lb != len(qp.variable_upper_bounds):
        raise ValueError(f"Inconsistent dimensions: variable lower bound vector has size {var_lb} while variable upper bound vector has size {len(qp.variable_upper_bounds)}")
    if var_lb != len(qp.objective_vector):
        raise ValueError(f"Inconsistent dimensions: variable lower bound vector has size {var_lb} while objective vector has size {len(qp.objective_vector)}")
    if var_lb != n:
        raise ValueError(f"Inconsistent dimensions: variable lower bound vector has size {var_lb} while constraint matrix has {n} columns")
    if qp.objective_matrix is not None and var_lb != len(qp.objective_matrix):
        raise ValueError(f"Inconsistent dimensions: variable lower bound vector has size {var_lb} while objective matrix has {len(qp.objective_matrix)} rows")
    if con_lb != len(qp.constraint_upper_bounds):
        raise ValueError(f"Inconsistent dimensions: constraint lower bound vector has size {con_lb} while constraint upper bound vector has size {len(qp.constraint_upper_bounds)}")
    if con_lb != m:
        raise ValueError(f"Inconsistent dimensions: constraint lower bound vector has size {con_lb} while constraint matrix has {m} rows ")
    if qp.variable_names is not None and var_lb != len(qp.variable_names):
        raise ValueError(f"Inconsistent dimensions: variable lower bound vector has size {var_lb} while variable names has size {len(qp.variable_names)}")
    if qp.constraint_names is not None and con_lb != len(qp.constraint_names):
        raise ValueError(f"Inconsistent dimensions: constraint lower bound vector has size {con_lb} while constraint names has size {len(qp.constraint_names)}")


class PrimalAndDualSolution:
    def __init__(self, primal_solution=None, dual_solution=None):
        self.primal_solution = np.zeros(0) if primal_solution is None else np.asarray(primal_solution, dtype=np.float64)
        self.dual_solution = np.zeros(0) if dual_solution is None else np.asarray(dual_solution, dtype=np.float64)


class SolverResult:
    def __init__(self):
        self.primal_solution = np.zeros(0)
        self.dual_solution = np.zeros(0)
        self.reduced_costs = np.zeros(0)
        self.solve_log = None


def _ns(d):
    """dict -> attribute namespace, recursively (SolveLog / IterationStats)."""
    if isinstance(d, dict):
        return types.SimpleNamespace(**{k: _ns(v) for k, v in d.items()})
    if isinstance(d, list):
        return [_ns(v) for v in d]
    return d


def _iteration_stats_from_pod(s):
    d = capi.struct_to_dict(s)
    d["convergence_information"] = d["convergence_information"][: d.pop("num_convergence_information")]
    d["infeasibility_information"] = d["infeasibility_information"][: d.pop("num_infeasibility_information")]
    md = d["point_metadata"][: d.pop("num_point_metadata")]
    for m in md:
        k = m.pop("num_random_projections")
        m["random_primal_projections"] = m["random_primal_projections"][:k]
        m["random_dual_projections"] = m["random_dual_projections"][:k]
    d["point_metadata"] = md
    return _ns(d)


def _cstr(ptr):
    return None if not ptr else C.cast(ptr, C.c_char_p).value.decode(errors="replace")


# --------------------------------------------------------------------------
class _ResultOwner:
    """Frees a PdlpResult (result_free of the backend it came from) when the numpy views of its
    solution vectors are gone."""

    def __init__(self, backend, res):
        self._free = backend.fn("result_free", None)
        self._res = res

    def __del__(self):
        try:
            self._free(C.byref(self._res))
        except Exception:
            pass


# Backend: binds one shared library exporting the C ABI with a given prefix.
# --------------------------------------------------------------------------
class Backend:
    """Thin ctypes binding of the C ABI (include/pdlp_b200.h)."""

    def __init__(self, library_path, prefix):
        self.library_path = library_path
        self.prefix = prefix
        self.lib = C.CDLL(library_path)

    def fn(self, name, restype=C.c_int32):
        f = getattr(self.lib, self.prefix + name)
        f.restype = restype
        return f

    def default_params_pod(self):
        p = capi.PdlpParams()
        self.fn("params_set_defaults", None)(C.byref(p))
        return p

    def validate_params(self, params):
        pod = params_to_pod(params)
        buf = C.create_string_buffer(1024)
        ok = self.fn("params_validate")(C.byref(pod), buf, C.c_int64(1024))
        return bool(ok), buf.value.decode()

    def _check(self, rc, what):
        if rc != 0:
            names = {1: "no usable CUDA device (the library has no CPU fallback)", 2: "CUDA/NCCL error", 3: "bad argument"}
            raise RuntimeError(f"{self.prefix}{what} failed: {names.get(rc, rc)}")

    def _solve_entry(self):
        return self.fn("primal_dual_hybrid_gradient"), ()

    def primal_dual_hybrid_gradient(self, qp, params, initial_solution=None, interrupt_solve=None,
                                    message_callback=None, iteration_stats_callback=None, result_pod_consumer=None):
        """`result_pod_consumer(res)`: called with the raw PdlpResult before it is released
        (e.g. native_io.solve_log_serialize)."""
        view, keep = qp._to_view()
        pod = params_to_pod(params)
        x0 = y0 = None
        if initial_solution is not None:
            x0 = capi.as_f64(initial_solution.primal_solution)
            y0 = capi.as_f64(initial_solution.dual_solution)
        if interrupt_solve is None:
            flag_ptr = None
        else:  # a ctypes c_int32 shared with the caller (std::atomic<bool> stand-in)
            flag_ptr = C.byref(interrupt_solve)
        msg_cb = capi.MESSAGE_CALLBACK(lambda m, _u: message_callback(m.decode(errors="replace"))) if message_callback else capi.MESSAGE_CALLBACK()
        if iteration_stats_callback:
            def _cb(info_p, _u):
                info = info_p.contents
                iteration_stats_callback(types.SimpleNamespace(
                    iteration_type=info.iteration_type,
                    termination_criteria=_ns(capi.struct_to_dict(info.termination_criteria.contents)),
                    iteration_stats=_iteration_stats_from_pod(info.iteration_stats.contents),
                    bound_norms=_ns(capi.struct_to_dict(info.bound_norms))))
            st_cb = capi.STATS_CALLBACK(_cb)
        else:
            st_cb = capi.STATS_CALLBACK()
        res = capi.PdlpResult()
        entry, lead = self._solve_entry()
        rc = entry(*lead, C.byref(view), C.byref(pod),
                   capi.ptr_f64(x0), C.c_int64(0 if x0 is None else x0.size),
                   capi.ptr_f64(y0), C.c_int64(0 if y0 is None else y0.size),
                   flag_ptr, msg_cb, st_cb, None, C.byref(res))
        del keep
        try:
            self._check(rc, "primal_dual_hybrid_gradient")
            if result_pod_consumer is not None:
                result_pod_consumer(res)
        except BaseException:
            self.fn("result_free", None)(C.byref(res))
            raise
        return self._result_from_pod(res)  # (takes ownership of res)

    def _result_from_pod(self, res):
        """SolverResult from the POD. The three solution vectors are NOT copied: they are numpy
        views of the library's buffers, which are released (result_free) when the last of them is
        garbage-collected -- the caller of this method must not free `res` itself."""
        out = SolverResult()
        owner = _ResultOwner(self, res)

        def vec(p, n):
            if n <= 0 or not p:
                return np.zeros(0)
            buf = (C.c_double * n).from_address(C.addressof(p.contents))
            buf._owner = owner  # the buffer object is the array's base: it keeps the owner alive
            return np.frombuffer(buf, dtype=np.float64)
        out.primal_solution = vec(res.primal_solution, res.primal_size)
        out.dual_solution = vec(res.dual_solution, res.dual_size)
        out.reduced_costs = vec(res.reduced_costs, res.primal_size)
        log = types.SimpleNamespace()
        log.instance_name = _cstr(res.instance_name)
        log.termination_reason = res.termination_reason
        log.termination_string = _cstr(res.termination_string) or ""
        log.iteration_count = res.iteration_count
        log.solve_time_sec = res.solve_time_sec
        log.preprocessing_time_sec = res.preprocessing_time_sec
        log.solution_type = res.solution_type
        log.solution_stats = _iteration_stats_from_pod(res.solution_stats) if res.has_solution_stats else None
        log.original_problem_stats = _ns(capi.struct_to_dict(res.original_problem_stats)) if res.has_original_problem_stats else None
        log.preprocessed_problem_stats = _ns(capi.struct_to_dict(res.preprocessed_problem_stats)) if res.has_preprocessed_problem_stats else None
        log.iteration_stats = [_iteration_stats_from_pod(res.iteration_stats[i]) for i in range(res.num_iteration_stats)]
        log.params = _ns(capi.struct_to_dict(res.params))
        log.feasibility_polishing_details = []
        for k in range(res.num_feasibility_polishing_details):
            d = res.feasibility_polishing_details[k]
            log.feasibility_polishing_details.append(types.SimpleNamespace(
                polishing_phase_type=d.polishing_phase_type, main_iteration_count=d.main_iteration_count,
                params=_ns(capi.struct_to_dict(d.params)), termination_reason=d.termination_reason,
                iteration_count=d.iteration_count, solve_time_sec=d.solve_time_sec,
                solution_stats=_iteration_stats_from_pod(d.solution_stats), solution_type=d.solution_type,
                iteration_stats=[_iteration_stats_from_pod(d.iteration_stats[i]) for i in range(d.num_iteration_stats)]))
        log.gpu_kernel_launches = res.gpu_kernel_launches
        log.device_iteration_time_sec = res.device_iteration_time_sec
        out.solve_log = log
        return out

    # ---- kernel-level entry points ------------------------------------
    def problem(self, qp, **kwargs):
        return DeviceProblem(self, qp, **kwargs)

    # The two calls whose argument lists depend on the library behind the C ABI (a subclass
    # binding another library overrides them).
    def _problem_create(self, view, handle, cuda_device=0, num_threads=1, num_shards=0):
        return self.fn("problem_create")(C.byref(view), C.c_int32(cuda_device), C.byref(handle))

    def _localized_bounds(self, prob, args, max_norm, out):
        if max_norm:  # PrimalDualNorm::kMaxNorm has its own entry point
            prob._call("compute_localized_lagrangian_bounds_max_norm", *args[:6], out)
        else:
            prob._call("compute_localized_lagrangian_bounds", *args, out)


class DeviceProblem:
    """A QP resident on the backend (ShardedQuadraticProgram equivalent)."""

    def __init__(self, backend, qp, cuda_device=0, num_threads=1, num_shards=0):
        self.b = backend
        self.qp = qp
        view, self._keep = qp._to_view()
        self.m, self.n, self.nnz = view.num_constraints, view.num_variables, view.num_nonzeros
        self.h = C.c_void_p()
        rc = backend._problem_create(view, self.h, cuda_device=cuda_device, num_threads=num_threads, num_shards=num_shards)
        backend._check(rc, "problem_create")

    def close(self):
        if self.h:
            self.b.fn("problem_destroy", None)(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _call(self, name, *args):
        self.b._check(self.b.fn(name)(self.h, *args), name)

    def transposed_matrix_vector_product(self, y):
        y = capi.as_f64(y); out = np.empty(self.n)
        self._call("transposed_matrix_vector_product", capi.ptr_f64(y), capi.ptr_f64(out))
        return out

    def matrix_vector_product(self, x):
        x = capi.as_f64(x); out = np.empty(self.m)
        self._call("matrix_vector_product", capi.ptr_f64(x), capi.ptr_f64(out))
        return out

    def apply_rescaling(self, l_inf_ruiz_iterations, l2_norm_rescaling):
        r, c = np.empty(self.m), np.empty(self.n)
        self._call("apply_rescaling", C.c_int32(l_inf_ruiz_iterations), C.c_int32(int(l2_norm_rescaling)), capi.ptr_f64(r), capi.ptr_f64(c))
        return r, c

    def scaling_iterations(self, norm, num_iterations, row_scaling_vec, col_scaling_vec):
        r, c = capi.as_f64(row_scaling_vec).copy(), capi.as_f64(col_scaling_vec).copy()
        self._call("scaling_iterations", C.c_int32(norm), C.c_int32(num_iterations), capi.ptr_f64(r), capi.ptr_f64(c))
        return r, c

    def scaled_col_norm(self, norm, row_scaling_vec, col_scaling_vec):
        r, c = capi.as_f64(row_scaling_vec), capi.as_f64(col_scaling_vec); out = np.empty(self.n)
        self._call("scaled_col_norm", C.c_int32(norm), capi.ptr_f64(r), capi.ptr_f64(c), capi.ptr_f64(out))
        return out

    def scaled_row_norm(self, norm, row_scaling_vec, col_scaling_vec):
        r, c = capi.as_f64(row_scaling_vec), capi.as_f64(col_scaling_vec); out = np.empty(self.m)
        self._call("scaled_row_norm", C.c_int32(norm), capi.ptr_f64(r), capi.ptr_f64(c), capi.ptr_f64(out))
        return out

    def rescale_quadratic_program(self, col_scaling_vec, row_scaling_vec):
        c, r = capi.as_f64(col_scaling_vec), capi.as_f64(row_scaling_vec)
        self._call("rescale_quadratic_program", capi.ptr_f64(c), capi.ptr_f64(r))

    def download(self):
        out = {"values": np.empty(self.nnz), "objective_vector": np.empty(self.n),
               "objective_matrix_diagonal": None if self.qp.objective_matrix is None else np.empty(self.n),
               "constraint_lower_bounds": np.empty(self.m), "constraint_upper_bounds": np.empty(self.m),
               "variable_lower_bounds": np.empty(self.n), "variable_upper_bounds": np.empty(self.n)}
        self._call("problem_download", *[capi.ptr_f64(out[k]) for k in (
            "values", "objective_vector", "objective_matrix_diagonal", "constraint_lower_bounds",
            "constraint_upper_bounds", "variable_lower_bounds", "variable_upper_bounds")])
        return out

    def compute_stats(self):
        s = capi.PdlpQuadraticProgramStats()
        self._call("compute_stats", C.byref(s))
        return _ns(capi.struct_to_dict(s))

    def project_to_primal_variable_bounds(self, primal, use_feasibility_bounds=False):
        v = capi.as_f64(primal).copy()
        self._call("project_to_primal_variable_bounds", capi.ptr_f64(v), C.c_int32(int(use_feasibility_bounds)))
        return v

    def project_to_dual_variable_bounds(self, dual):
        v = capi.as_f64(dual).copy()
        self._call("project_to_dual_variable_bounds", capi.ptr_f64(v))
        return v

    def compute_primal_gradient(self, primal, dual_product):
        x, dp = capi.as_f64(primal), capi.as_f64(dual_product); g = np.empty(self.n); val = C.c_double()
        self._call("compute_primal_gradient", capi.ptr_f64(x), capi.ptr_f64(dp), capi.ptr_f64(g), C.byref(val))
        return g, val.value

    def compute_dual_gradient(self, dual, primal_product):
        y, pp = capi.as_f64(dual), capi.as_f64(primal_product); g = np.empty(self.m); val = C.c_double()
        self._call("compute_dual_gradient", capi.ptr_f64(y), capi.ptr_f64(pp), capi.ptr_f64(g), C.byref(val))
        return g, val.value

    def _pod(self, params):
        if params is None:
            params = PrimalDualHybridGradientParams()
        return params.to_pod() if hasattr(params, "to_pod") else params

    def compute_convergence_information(self, params, col_scaling_vec, row_scaling_vec, primal, dual,
                                        componentwise_primal_residual_offset=1.0,
                                        componentwise_dual_residual_offset=1.0,
                                        candidate_type=PointType.POINT_TYPE_CURRENT_ITERATE):
        pod = self._pod(params)
        cs = None if col_scaling_vec is None else capi.as_f64(col_scaling_vec)
        rs = None if row_scaling_vec is None else capi.as_f64(row_scaling_vec)
        x, y = capi.as_f64(primal), capi.as_f64(dual)
        out = capi.PdlpConvergenceInformation()
        self._call("compute_convergence_information", C.byref(pod), capi.ptr_f64(cs), capi.ptr_f64(rs), capi.ptr_f64(x), capi.ptr_f64(y),
                   C.c_double(componentwise_primal_residual_offset), C.c_double(componentwise_dual_residual_offset),
                   C.c_int32(candidate_type), C.byref(out))
        return _ns(capi.struct_to_dict(out))

    def compute_infeasibility_information(self, params, col_scaling_vec, row_scaling_vec, primal_ray, dual_ray,
                                          primal_solution_for_residual_tests,
                                          candidate_type=PointType.POINT_TYPE_CURRENT_ITERATE):
        pod = self._pod(params)
        cs = None if col_scaling_vec is None else capi.as_f64(col_scaling_vec)
        rs = None if row_scaling_vec is None else capi.as_f64(row_scaling_vec)
        x, y, xr = capi.as_f64(primal_ray), capi.as_f64(dual_ray), capi.as_f64(primal_solution_for_residual_tests)
        out = capi.PdlpInfeasibilityInformation()
        self._call("compute_infeasibility_information", C.byref(pod), capi.ptr_f64(cs), capi.ptr_f64(rs), capi.ptr_f64(x), capi.ptr_f64(y),
                   capi.ptr_f64(xr), C.c_int32(candidate_type), C.byref(out))
        return _ns(capi.struct_to_dict(out))

    def reduced_costs(self, params, primal, dual, use_zero_primal_objective=False):
        pod = self._pod(params)
        x, y = capi.as_f64(primal), capi.as_f64(dual); out = np.empty(self.n)
        self._call("reduced_costs", C.byref(pod), capi.ptr_f64(x), capi.ptr_f64(y), C.c_int32(int(use_zero_primal_objective)), capi.ptr_f64(out))
        return out

    def compute_localized_lagrangian_bounds(self, primal, dual, primal_weight, radius, primal_product=None,
                                            dual_product=None, use_diagonal_qp_trust_region_solver=False,
                                            diagonal_qp_trust_region_solver_tolerance=1e-8, max_norm=False):
        x, y = capi.as_f64(primal), capi.as_f64(dual)
        pp = None if primal_product is None else capi.as_f64(primal_product)
        dp = None if dual_product is None else capi.as_f64(dual_product)
        out = (C.c_double * 4)()
        args = [capi.ptr_f64(x), capi.ptr_f64(y), C.c_double(primal_weight), C.c_double(radius), capi.ptr_f64(pp), capi.ptr_f64(dp),
                C.c_int32(int(use_diagonal_qp_trust_region_solver)), C.c_double(diagonal_qp_trust_region_solver_tolerance)]
        self.b._localized_bounds(self, args, max_norm, out)
        return types.SimpleNamespace(lagrangian_value=out[0], lower_bound=out[1], upper_bound=out[2], radius=out[3])


# --------------------------------------------------------------------------
# The product backend
# --------------------------------------------------------------------------
_LIB_NAME = "libpdlp_b200.so"
_backend = None


def library_path():
    """The in-tree build; PDLP_B200_LIBRARY names another build of the same library (the sanitizer
    build of tools/asan_host.sh)."""
    override = os.environ.get("PDLP_B200_LIBRARY", "")
    return override if override else os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", _LIB_NAME)


class _ProductBackend(Backend):
    def __init__(self):
        path = library_path()
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
        super().__init__(path, "pdlp_b200_")

    def device_count(self):
        return int(self.fn("device_count")())

    def set_default_device(self, cuda_device):
        self._check(self.fn("set_default_device")(C.c_int32(cuda_device)), "set_default_device")

    def session(self, qp, params, initial_solution=None, cuda_device=0):
        return SolveSession(self, qp, params, initial_solution, cuda_device)

    def solve_trust_region(self, objective_vector, variable_lower_bounds, variable_upper_bounds, center_point,
                           norm_weights, target_radius, cuda_device=0):
        arrs = [capi.as_f64(a) for a in (objective_vector, variable_lower_bounds, variable_upper_bounds, center_point, norm_weights)]
        n = arrs[0].size
        sol = np.empty(n); step = C.c_double(); val = C.c_double()
        self._check(self.fn("solve_trust_region")(C.c_int32(cuda_device), C.c_int64(n), *[capi.ptr_f64(a) for a in arrs],
                                                  C.c_double(target_radius), capi.ptr_f64(sol), C.byref(step), C.byref(val)), "solve_trust_region")
        return types.SimpleNamespace(solution=sol, solution_step_size=step.value, objective_value=val.value)

    def solve_diagonal_trust_region(self, objective_vector, objective_matrix_diagonal, variable_lower_bounds,
                                    variable_upper_bounds, center_point, norm_weights, target_radius,
                                    solve_tolerance, cuda_device=0):
        arrs = [capi.as_f64(a) for a in (objective_vector, objective_matrix_diagonal, variable_lower_bounds, variable_upper_bounds, center_point, norm_weights)]
        n = arrs[0].size
        sol = np.empty(n); step = C.c_double(); val = C.c_double()
        self._check(self.fn("solve_diagonal_trust_region")(C.c_int32(cuda_device), C.c_int64(n), *[capi.ptr_f64(a) for a in arrs],
                                                           C.c_double(target_radius), C.c_double(solve_tolerance), capi.ptr_f64(sol),
                                                           C.byref(step), C.byref(val)), "solve_diagonal_trust_region")
        return types.SimpleNamespace(solution=sol, solution_step_size=step.value, objective_value=val.value)

    def weighted_average(self, datapoints, weights, cuda_device=0):
        d = capi.as_f64(datapoints); w = capi.as_f64(weights)
        count, size = d.shape
        out = np.empty(size); sw = C.c_double(); nt = C.c_int32()
        self._check(self.fn("weighted_average")(C.c_int32(cuda_device), C.c_int64(size), C.c_int64(count), capi.ptr_f64(d), capi.ptr_f64(w),
                                                capi.ptr_f64(out), C.byref(sw), C.byref(nt)), "weighted_average")
        return out, sw.value, nt.value

    def vector_reduce(self, op, a, b=None, cuda_device=0):
        a = capi.as_f64(a); bb = None if b is None else capi.as_f64(b); out = C.c_double()
        self._check(self.fn("vector_reduce")(C.c_int32(cuda_device), C.c_int32(op), C.c_int64(a.size), capi.ptr_f64(a), capi.ptr_f64(bb), C.byref(out)), "vector_reduce")
        return out.value


class SolveSession:
    """A solve whose problem and iterates stay resident in HBM between calls
    (C ABI ``pdlp_b200_session_*``): create = preprocessing, ``advance`` = the
    PDHG loop up to a target iteration count, ``finish`` = the SolverResult."""

    def __init__(self, backend, qp, params, initial_solution=None, cuda_device=0):
        self.b = backend
        view, keep = qp._to_view()
        pod = params_to_pod(params)
        x0 = y0 = None
        if initial_solution is not None:
            x0 = capi.as_f64(initial_solution.primal_solution)
            y0 = capi.as_f64(initial_solution.dual_solution)
        self.h = C.c_void_p()
        rc = backend.fn("session_create")(C.byref(view), C.byref(pod), capi.ptr_f64(x0), C.c_int64(0 if x0 is None else x0.size),
                                          capi.ptr_f64(y0), C.c_int64(0 if y0 is None else y0.size), capi.MESSAGE_CALLBACK(),
                                          capi.STATS_CALLBACK(), None, C.c_int32(cuda_device), C.byref(self.h))
        del keep
        backend._check(rc, "session_create")

    def enable_timing(self, enable=True, sample_stride=8):
        self.b._check(self.b.fn("session_enable_timing")(self.h, C.c_int32(int(enable)), C.c_int32(sample_stride)), "session_enable_timing")

    def advance(self, target_iterations, interrupt_solve=None):
        st = capi.PdlpSessionStatus()
        flag = None if interrupt_solve is None else C.byref(interrupt_solve)
        self.b._check(self.b.fn("session_advance")(self.h, C.c_int32(int(target_iterations)), flag, C.byref(st)), "session_advance")
        return _ns(capi.struct_to_dict(st))

    def status(self):
        st = capi.PdlpSessionStatus()
        self.b._check(self.b.fn("session_status")(self.h, C.byref(st)), "session_status")
        return _ns(capi.struct_to_dict(st))

    def finish(self):
        res = capi.PdlpResult()
        try:
            self.b._check(self.b.fn("session_finish")(self.h, C.byref(res)), "session_finish")
        except BaseException:
            self.b.fn("result_free", None)(C.byref(res))
            raise
        return self.b._result_from_pod(res)  # (takes ownership of res)

    def close(self):
        if self.h:
            self.b.fn("session_destroy", None)(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def backend():
    """The process-wide binding of libpdlp_b200.so (loaded on first use)."""
    global _backend
    if _backend is None:
        _backend = _ProductBackend()
    return _backend


def primal_dual_hybrid_gradient(qp, params, initial_solution=None, interrupt_solve=None,
                                message_callback=None, iteration_stats_callback=None):
    """``pdlp.primal_dual_hybrid_gradient`` (python/pdlp.cc:143-150).

    The two extra keyword arguments expose what the C++ API has and the pybind
    wrapper leaves as a TODO (interrupt flag: a ``ctypes.c_int32``; callbacks).
    """
    return backend().primal_dual_hybrid_gradient(qp, params, initial_solution, interrupt_solve,
                                                 message_callback, iteration_stats_callback)


# --------------------------------------------------------------------------
# quadratic_program.h / quadratic_program_io.h functions of the reference's wrapper
# (python/pdlp.cc:95-126); errors surface as ValueError like std::invalid_argument.
# --------------------------------------------------------------------------
def to_string(qp, max_size=1_000_000):
    """A readable rendering of `qp` cut to at most `max_size` characters (ToString,
    quadratic_program.h:188-192; for debugging, not LP format). Numbers print with six
    significant digits like absl::StrCat."""
    try:
        validate_quadratic_program_dimensions(qp)
    except ValueError as e:
        return "Quadratic program with inconsistent dimensions: %s" % e
    import scipy.sparse as sp
    n, m = len(qp.variable_lower_bounds), len(qp.constraint_lower_bounds)
    vname = (lambda j: qp.variable_names[j]) if qp.variable_names is not None else (lambda j: "x%d" % j)
    cname = (lambda i: qp.constraint_names[i]) if qp.constraint_names is not None else (lambda i: "c%d" % i)
    num = lambda v: "%g" % v
    pieces, size = [], 0

    def emit(text):
        nonlocal size
        pieces.append(text)
        size += len(text)
        return size < max_size

    def render():
        if qp.problem_name is not None:
            emit("%s:\n" % qp.problem_name)
        scale = qp.objective_scaling_factor
        emit("%s %s * (%s" % ("maximize" if scale < 0.0 else "minimize", num(scale), num(qp.objective_offset)))
        for j in np.flatnonzero(np.asarray(qp.objective_vector) != 0.0):
            if not emit(" + %s %s" % (num(qp.objective_vector[j]), vname(j))):
                break
        if qp.objective_matrix is not None:
            emit(" + 1/2 * (")
            for j in np.flatnonzero(np.asarray(qp.objective_matrix) != 0.0):
                if not emit(" + %s %s^2" % (num(qp.objective_matrix[j]), vname(j))):
                    break
            emit(")")
        emit(")\n")
        rows = sp.csr_matrix(qp.constraint_matrix)
        rows.sort_indices()
        for i in range(m):
            emit("%s:" % cname(i))
            if qp.constraint_lower_bounds[i] != -math.inf:
                emit(" %s <=" % num(qp.constraint_lower_bounds[i]))
            for p in range(rows.indptr[i], rows.indptr[i + 1]):
                if not emit(" + %s %s" % (num(rows.data[p]), vname(rows.indices[p]))):
                    break
            if qp.constraint_upper_bounds[i] != math.inf:
                emit(" <= %s" % num(qp.constraint_upper_bounds[i]))
            if not emit("\n"):
                return
        emit("Bounds\n")
        for j in range(n):
            lo, hi = qp.variable_lower_bounds[j], qp.variable_upper_bounds[j]
            if lo == -math.inf:
                line = "%s free\n" % vname(j) if hi == math.inf else "%s <= %s\n" % (vname(j), num(hi))
            else:
                line = "%s >= %s\n" % (vname(j), num(lo)) if hi == math.inf else "%s <= %s <= %s\n" % (num(lo), vname(j), num(hi))
            if not emit(line):
                return

    render()
    text = "".join(pieces)
    if len(text) > max_size:
        text = text[: max(0, max_size - 4)] + "...\n"
    return text


# iteration_stats.h:94-106: the entry of a repeated field of IterationStats for one point type, or None.
def get_convergence_information(stats, candidate_type):
    return next((c for c in (stats.convergence_information if stats is not None else []) if c.candidate_type == candidate_type), None)


def get_infeasibility_information(stats, candidate_type):
    return next((c for c in (stats.infeasibility_information if stats is not None else []) if c.candidate_type == candidate_type), None)


def get_point_metadata(stats, point_type):
    return next((c for c in (stats.point_metadata if stats is not None else []) if c.point_type == point_type), None)


def qp_from_mpmodel_proto(proto_str, relax_integer_variables, include_names=False):
    from . import mp_model
    return mp_model.qp_from_mp_model_proto(proto_str, relax_integer_variables, include_names)


def qp_to_mpmodel_proto(qp):
    from . import mp_model
    return mp_model.qp_to_mp_model_proto(qp)


def read_quadratic_program_or_die(filename, include_names=False):
    """ReadQuadraticProgramOrDie (quadratic_program_io.h:28-40; raises instead of dying). Read by the
    library's own C++ readers (pdlp_b200_read_quadratic_program), .gz and .bz2 included."""
    from . import native_io
    return native_io.read_quadratic_program(filename, include_names)
