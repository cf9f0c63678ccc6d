"""The drop-in boundary: libpdlp_b200.so loads, exports every symbol that
include/pdlp_b200.h declares, its POD layouts match the ctypes mirror, and the
parameter defaults / validation match the reference protos (no compute calls:
these run without a GPU)."""
import ctypes as C
import os
import re

import pytest

from ortools_b200 import _capi as capi
from ortools_b200 import pdlp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADERS = [os.path.join(ROOT, "include", "pdlp_b200.h"), os.path.join(ROOT, "include", "pdlp_b200_io.h")]


def declared_functions():
    names = set()
    for header in HEADERS:
        text = open(header).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names.update(re.findall(r"\b(pdlp_b200_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_header_declares_the_expected_entry_points():
    names = declared_functions()
    assert "pdlp_b200_primal_dual_hybrid_gradient" in names
    assert "pdlp_b200_session_create" in names and "pdlp_b200_session_advance" in names
    assert "pdlp_b200_solve_proto" in names and "pdlp_b200_read_quadratic_program" in names   # pdlp_b200_io.h
    assert len(names) >= 55


@pytest.mark.parametrize("name", declared_functions())
def test_library_exports_declared_symbol(name):
    lib = C.CDLL(pdlp.library_path())
    assert getattr(lib, name) is not None


def test_no_cpu_fallback_without_device():
    be = pdlp.backend()
    if be.device_count() > 0:
        pytest.skip("a CUDA device is present")
    qp = pdlp.QuadraticProgram(2, 1)
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        be.primal_dual_hybrid_gradient(qp, pdlp.PrimalDualHybridGradientParams())
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        be.problem(qp)
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        be.session(qp, pdlp.PrimalDualHybridGradientParams())


def test_pod_sizes_match_the_library():
    # The library reports sizeof() of every struct crossing the boundary.
    lib = C.CDLL(pdlp.library_path())
    lib.pdlp_b200_sizeof.restype = C.c_int64
    for idx, typ in enumerate([capi.PdlpTerminationCriteria, capi.PdlpParams, capi.PdlpProblemView, capi.PdlpQuadraticProgramStats,
                               capi.PdlpConvergenceInformation, capi.PdlpInfeasibilityInformation, capi.PdlpPointMetadata,
                               capi.PdlpIterationStats, capi.PdlpBoundNorms, capi.PdlpIterationCallbackInfo, capi.PdlpResult,
                               capi.PdlpSessionStatus]):
        assert lib.pdlp_b200_sizeof(C.c_int32(idx)) == C.sizeof(typ), typ.__name__


def test_param_defaults_match_solvers_proto():
    # solvers.proto:66-497 defaults
    p = pdlp.backend().default_params_pod()
    tc = p.termination_criteria
    assert tc.optimality_norm == pdlp.OptimalityNorm.OPTIMALITY_NORM_L2
    assert tc.eps_optimal_absolute == 1e-6 and tc.eps_optimal_relative == 1e-6
    assert tc.eps_primal_infeasible == 1e-8 and tc.eps_dual_infeasible == 1e-8
    assert tc.iteration_limit == 2**31 - 1 and tc.time_sec_limit == float("inf")
    assert p.major_iteration_frequency == 64 and p.termination_check_frequency == 64
    assert p.restart_strategy == pdlp.RestartStrategy.ADAPTIVE_HEURISTIC
    assert p.l_inf_ruiz_iterations == 5 and p.l2_norm_rescaling == 1
    assert p.sufficient_reduction_for_restart == 0.1 and p.necessary_reduction_for_restart == 0.9
    assert p.adaptive_step_size_reduction_exponent == 0.3 and p.adaptive_step_size_growth_exponent == 0.6
    assert p.primal_weight_update_smoothing == 0.5
    assert p.handle_some_primal_gradients_on_finite_bounds_as_residuals == 1


def test_param_validation_messages():
    be = pdlp.backend()
    p = pdlp.PrimalDualHybridGradientParams()
    ok, msg = be.validate_params(p)
    assert ok and msg == ""
    p.major_iteration_frequency = 0
    ok, msg = be.validate_params(p)
    assert not ok and "major_iteration_frequency" in msg


def test_trust_region_preconditions_are_statuses_not_crashes():
    """trust_region_test.cc:503-583 (TrustRegionDeathTest): the reference CHECK-fails on a negative radius or
    a non-positive norm weight; across the C ABI they are PDLP_B200_STATUS_BAD_ARGUMENT, decided on the
    host before any device work (so this runs without a GPU)."""
    be = pdlp.backend()
    ones, zeros = [1.0, 1.0], [0.0, 0.0]
    lo, hi = [-10.0, -10.0], [10.0, 10.0]
    for weights, radius in (([1.0, 0.0], 1.0), ([1.0, -2.0], 1.0), ([1.0, float("nan")], 1.0), (ones, -1.0), (ones, float("nan"))):
        with pytest.raises(RuntimeError, match="bad argument"):
            be.solve_trust_region(ones, lo, hi, zeros, weights, radius)
        with pytest.raises(RuntimeError, match="bad argument"):
            be.solve_diagonal_trust_region(ones, ones, lo, hi, zeros, weights, radius, 1e-8)


def test_public_headers_compile_as_c99_and_strict_cpp(tmp_path):
    """include/*.h are plain C (the boundary has no C++ or torch types); the .hpp faces are warning-free C++17."""
    import subprocess
    inc = os.path.join(ROOT, "include")
    for header in ("pdlp_b200.h", "pdlp_b200_io.h"):
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Wpedantic", "-Werror", "-I" + inc, "-fsyntax-only", "-x", "c", os.path.join(inc, header)])
    src = tmp_path / "faces.cc"
    src.write_text('#include "pdlp_b200.hpp"\n#include "pdlp_b200_io.hpp"\nint main() { return 0; }\n')
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wextra", "-Wpedantic", "-Wshadow", "-Werror", "-I" + inc, "-fsyntax-only", str(src)])
