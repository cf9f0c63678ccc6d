"""TEST INFRASTRUCTURE ONLY -- python binding of the CPU oracle.

Builds (g++) and loads oracle/libpdlp_oracle.so, the CPU restatement of the
reference PDLP (oracle/pdlp_cpu_core.h, oracle/pdlp_cpu_solver.cc), and exposes
it through the same ``Backend`` interface the product uses, so parity tests can
call both sides with identical code. Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference leg may import this module.
"""
import ctypes as C
import os
import subprocess
import types

import numpy as np

from ortools_b200 import pdlp
from ortools_b200 import _capi as capi

_DIR = os.path.dirname(os.path.abspath(__file__))
# (PDLP_ORACLE_LIBRARY: another build of the same sources, e.g. the ThreadSanitizer build of tools/tsan_oracle.sh)
LIB = os.environ.get("PDLP_ORACLE_LIBRARY") or os.path.join(_DIR, "libpdlp_oracle.so")
SOURCES = [os.path.join(_DIR, "pdlp_cpu_solver.cc"), os.path.join(_DIR, "pdlp_cpu_core.h"),
           os.path.join(_DIR, "..", "include", "pdlp_b200.h")]


def build(force=False):
    """Compiles the oracle if it is missing or older than its sources."""
    if os.environ.get("PDLP_ORACLE_LIBRARY"):
        return LIB
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in SOURCES if os.path.exists(s)):
        return LIB
    # x86-64-v3 (AVX2 + FMA), not -march=native: the library is built in one container and travels to the GPU
    # box with the snapshot, whose host CPU need not have every extension of the build machine's
    cmd = ["g++", "-O3", "-march=x86-64-v3", "-std=c++17", "-fPIC", "-shared", "-pthread", SOURCES[0], "-o", LIB]
    subprocess.check_call(cmd)
    return LIB


class OracleBackend(pdlp.Backend):
    def __init__(self):
        if not os.path.exists(LIB):
            build()
        super().__init__(LIB, "pdlp_oracle_")

    # the oracle's problem takes the reference's thread / shard counts instead of a CUDA device,
    # and one bounds entry point serves both norms
    def _problem_create(self, view, handle, cuda_device=0, num_threads=1, num_shards=0):
        return self.fn("problem_create")(C.byref(view), C.c_int32(num_threads), C.c_int32(num_shards), C.byref(handle))

    def _localized_bounds(self, prob, args, max_norm, out):
        prob._call("compute_localized_lagrangian_bounds", *args, C.c_int32(1 if max_norm else 0), out)

    def sharder_starts(self, num_elements, num_shards, masses=None):
        out = np.zeros(num_elements + 2, dtype=np.int64)
        mp = None if masses is None else np.ascontiguousarray(masses, dtype=np.int64)
        f = self.fn("sharder_starts", C.c_int64)
        k = f(C.c_int64(num_elements), C.c_int32(num_shards), None if mp is None else capi.ptr_i64(mp), capi.ptr_i64(out), C.c_int64(out.size))
        return out[:k].tolist()

    def solve_trust_region(self, objective_vector, variable_lower_bounds, variable_upper_bounds, center_point,
                           norm_weights, target_radius, num_threads=1, num_shards=1):
        arrs = [capi.as_f64(a) for a in (objective_vector, variable_lower_bounds, variable_upper_bounds, center_point, norm_weights)]
        n = arrs[0].size
        sol = np.empty(n); step = C.c_double(); val = C.c_double()
        self.fn("solve_trust_region")(C.c_int32(num_threads), C.c_int32(num_shards), C.c_int64(n), *[capi.ptr_f64(a) for a in arrs],
                                      C.c_double(target_radius), capi.ptr_f64(sol), C.byref(step), C.byref(val))
        return types.SimpleNamespace(solution=sol, solution_step_size=step.value, objective_value=val.value)

    def solve_diagonal_trust_region(self, objective_vector, objective_matrix_diagonal, variable_lower_bounds,
                                    variable_upper_bounds, center_point, norm_weights, target_radius,
                                    solve_tolerance, num_threads=1, num_shards=1):
        arrs = [capi.as_f64(a) for a in (objective_vector, objective_matrix_diagonal, variable_lower_bounds, variable_upper_bounds, center_point, norm_weights)]
        n = arrs[0].size
        sol = np.empty(n); step = C.c_double(); val = C.c_double()
        self.fn("solve_diagonal_trust_region")(C.c_int32(num_threads), C.c_int32(num_shards), C.c_int64(n), *[capi.ptr_f64(a) for a in arrs],
                                               C.c_double(target_radius), C.c_double(solve_tolerance), capi.ptr_f64(sol), C.byref(step), C.byref(val))
        return types.SimpleNamespace(solution=sol, solution_step_size=step.value, objective_value=val.value)

    def weighted_average(self, datapoints, weights, num_shards=1):
        d = capi.as_f64(datapoints); w = capi.as_f64(weights)
        count, size = d.shape
        out = np.empty(size); sw = C.c_double(); nt = C.c_int32()
        self.fn("weighted_average")(C.c_int32(num_shards), C.c_int64(size), C.c_int64(count), capi.ptr_f64(d), capi.ptr_f64(w),
                                    capi.ptr_f64(out), C.byref(sw), C.byref(nt))
        return out, sw.value, nt.value

    def vector_reduce(self, op, a, b=None, num_shards=1):
        a = capi.as_f64(a); bb = None if b is None else capi.as_f64(b); out = C.c_double()
        self.fn("vector_reduce")(C.c_int32(num_shards), C.c_int32(op), C.c_int64(a.size), capi.ptr_f64(a), capi.ptr_f64(bb), C.byref(out))
        return out.value

    def vector_update(self, op, scale, a, dest, num_shards=1):
        a = capi.as_f64(a); d = capi.as_f64(dest).copy()
        self.fn("vector_update")(C.c_int32(num_shards), C.c_int32(op), C.c_int64(a.size), C.c_double(scale), capi.ptr_f64(a), capi.ptr_f64(d))
        return d

    def check_simple_termination_criteria(self, criteria_pod, stats_pod, interrupt=None):
        reason, typ = C.c_int32(), C.c_int32()
        hit = self.fn("check_simple_termination_criteria")(C.byref(criteria_pod), C.byref(stats_pod), None if interrupt is None else C.byref(interrupt), C.byref(reason), C.byref(typ))
        return (reason.value, typ.value) if hit else None

    def check_iterate_termination_criteria(self, criteria_pod, stats_pod, bound_norms_pod, force_numerical_termination=False):
        reason, typ = C.c_int32(), C.c_int32()
        hit = self.fn("check_iterate_termination_criteria")(C.byref(criteria_pod), C.byref(stats_pod), C.byref(bound_norms_pod),
                                                            C.c_int32(int(force_numerical_termination)), C.byref(reason), C.byref(typ))
        return (reason.value, typ.value) if hit else None

    def compute_relative_residuals(self, criteria_pod, conv_pod, bound_norms_pod):
        out = (C.c_double * 5)()
        self.fn("compute_relative_residuals")(C.byref(criteria_pod), C.byref(conv_pod), C.byref(bound_norms_pod), out)
        return types.SimpleNamespace(relative_l_inf_primal_residual=out[0], relative_l2_primal_residual=out[1],
                                     relative_l_inf_dual_residual=out[2], relative_l2_dual_residual=out[3],
                                     relative_optimality_gap=out[4])


def problem_extras(prob):
    """Oracle-only accessors on a DeviceProblem created from OracleBackend."""
    b = prob.b

    def shard_starts(which):
        out = np.zeros(max(prob.n, prob.m) + 2, dtype=np.int64)
        k = b.fn("shard_starts", C.c_int64)(prob.h, C.c_int32(which), capi.ptr_i64(out), C.c_int64(out.size))
        return out[:k].tolist()

    def transposed_values():
        out = np.empty(prob.nnz)
        b.fn("transposed_values")(prob.h, capi.ptr_f64(out))
        return out

    def replace_large_bounds(threshold):
        b.fn("replace_large_constraint_bounds_with_infinity")(prob.h, C.c_double(threshold))

    def has_valid_bounds():
        return bool(b.fn("has_valid_bounds")(prob.h))

    def estimate_max_singular_value(primal=None, dual=None, desired_relative_error=0.01, failure_probability=0.001, seed=1):
        sv = C.c_double(); it = C.c_int32()
        p = None if primal is None else capi.as_f64(primal)
        d = None if dual is None else capi.as_f64(dual)
        b.fn("estimate_max_singular_value")(prob.h, capi.ptr_f64(p), capi.ptr_f64(d), C.c_double(desired_relative_error),
                                            C.c_double(failure_probability), C.c_uint32(seed), C.byref(sv), C.byref(it))
        return sv.value, it.value

    return types.SimpleNamespace(shard_starts=shard_starts, transposed_values=transposed_values,
                                 replace_large_bounds=replace_large_bounds, has_valid_bounds=has_valid_bounds,
                                 estimate_max_singular_value=estimate_max_singular_value)


_backend = None


def backend():
    global _backend
    if _backend is None:
        build()
        _backend = OracleBackend()
    return _backend
