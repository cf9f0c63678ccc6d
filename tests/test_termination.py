"""termination.h (termination.cc:26-271): the scalar termination logic, with the known answers of
``ortools/pdlp/termination_test.cc:121-1130``. Both the CPU restatement and THIS library's own
host code (exported as host-only C entry points, no device needed) are checked. Criteria and
statistics are written as the reference writes them -- text-format protos -- and turned into the
PODs of the C ABI."""
import ctypes as C
import math

import pytest
from google.protobuf import text_format

from ortools_b200 import _capi as capi
from ortools_b200 import pdlp, pdlp_proto

TR, PT, NORM = pdlp.TerminationReason, pdlp.PointType, pdlp.OptimalityNorm
INF = float("inf")
NORMS = [NORM.OPTIMALITY_NORM_L2, NORM.OPTIMALITY_NORM_L_INF, NORM.OPTIMALITY_NORM_L_INF_COMPONENTWISE]


@pytest.fixture(params=["oracle", "product"])
def lib(request):
    if request.param == "oracle":
        from oracle import pdlp_oracle
        return pdlp_oracle.backend()
    return pdlp.backend()   # the host-side entry points need no GPU


def criteria(text, norm=None):
    msg = text_format.Parse(text, pdlp_proto.TerminationCriteriaProto())
    if norm is not None:
        msg.optimality_norm = norm
    tc = pdlp.TerminationCriteria()
    pdlp_proto._copy_set_fields(msg, tc)
    return tc.to_pod()


def stats(text=""):
    msg = text_format.Parse(text, pdlp_proto.IterationStatsProto())
    s = capi.PdlpIterationStats()
    s.iteration_number = msg.iteration_number
    s.cumulative_kkt_matrix_passes = msg.cumulative_kkt_matrix_passes
    s.cumulative_time_sec = msg.cumulative_time_sec
    s.num_convergence_information = len(msg.convergence_information)
    for k, c in enumerate(msg.convergence_information):
        for f, _ in capi.PdlpConvergenceInformation._fields_:
            setattr(s.convergence_information[k], f, getattr(c, f))
    s.num_infeasibility_information = len(msg.infeasibility_information)
    for k, c in enumerate(msg.infeasibility_information):
        for f, _ in capi.PdlpInfeasibilityInformation._fields_:
            setattr(s.infeasibility_information[k], f, getattr(c, f))
    return s


def bound_norms(l2_obj, l2_bounds, linf_obj, linf_bounds):
    b = capi.PdlpBoundNorms()
    b.l2_norm_primal_linear_objective, b.l2_norm_constraint_bounds = l2_obj, l2_bounds
    b.l_inf_norm_primal_linear_objective, b.l_inf_norm_constraint_bounds = linf_obj, linf_bounds
    return b


def test_lp_bound_norms():  # termination_test.cc:35-41
    return bound_norms(math.sqrt(36.25), math.sqrt(210.0), 5.5, 12.0)


test_lp_bound_norms.__test__ = False
ZERO_NORMS = (0.0, 0.0, 0.0, 0.0)


def check_simple(lib, c, s, interrupt=None):
    reason, typ = C.c_int32(), C.c_int32()
    flag = None if interrupt is None else C.byref(C.c_int32(int(interrupt)))
    hit = lib.fn("check_simple_termination_criteria")(C.byref(c), C.byref(s), flag, C.byref(reason), C.byref(typ))
    return (reason.value, typ.value) if hit else None


def check_iterate(lib, c, s, bn=None, force=False):
    bn = test_lp_bound_norms() if bn is None else bn
    reason, typ = C.c_int32(), C.c_int32()
    hit = lib.fn("check_iterate_termination_criteria")(C.byref(c), C.byref(s), C.byref(bn), C.c_int32(int(force)), C.byref(reason), C.byref(typ))
    return (reason.value, typ.value) if hit else None


def criteria_met(lib, c, s, bn=None):
    bn = test_lp_bound_norms() if bn is None else bn
    gap = C.c_int32()
    met = lib.fn("optimality_criteria_met")(C.byref(c), C.byref(s.convergence_information[0]), C.byref(bn), C.byref(gap))
    return bool(met), bool(gap.value)


def effective(lib, c):
    out = (C.c_double * 6)()
    lib.fn("effective_optimality_criteria", None)(C.byref(c), out)
    return list(out)


def relative(lib, c, conv, bn=None):
    bn = test_lp_bound_norms() if bn is None else bn
    out = (C.c_double * 5)()
    lib.fn("compute_relative_residuals", None)(C.byref(c), C.byref(conv), C.byref(bn), out)
    return list(out)


SIMPLE = "time_sec_limit: 1.0 kkt_matrix_pass_limit: 2000 iteration_limit: 10"
ITERATE = """simple_optimality_criteria { eps_optimal_absolute: 1.0e-4 eps_optimal_relative: 1.0e-4 }
             eps_primal_infeasible: 1.0e-6 eps_dual_infeasible: 1.0e-6 time_sec_limit: 1.0 kkt_matrix_pass_limit: 2000 iteration_limit: 10"""
DETAILED_RELATIVE = """detailed_optimality_criteria {
    eps_optimal_primal_residual_absolute: 0.0 eps_optimal_primal_residual_relative: 1.0e-4
    eps_optimal_dual_residual_absolute: 0.0 eps_optimal_dual_residual_relative: 1.0e-4
    eps_optimal_objective_gap_absolute: 0.0 eps_optimal_objective_gap_relative: 1.0e-4 }"""
DETAILED_ABSOLUTE = """detailed_optimality_criteria {
    eps_optimal_primal_residual_absolute: 1.0e-4 eps_optimal_primal_residual_relative: 0.0
    eps_optimal_dual_residual_absolute: 1.0e-4 eps_optimal_dual_residual_relative: 0.0
    eps_optimal_objective_gap_absolute: 1.0e-4 eps_optimal_objective_gap_relative: 0.0 }"""


def conv(pobj, dobj, linf_p, linf_d, l2_p, l2_d, cw_p, cw_d, typed=True):
    return stats("""convergence_information { primal_objective: %r dual_objective: %r l_inf_primal_residual: %r l_inf_dual_residual: %r
        l2_primal_residual: %r l2_dual_residual: %r l_inf_componentwise_primal_residual: %r l_inf_componentwise_dual_residual: %r %s }""" % (
        pobj, dobj, linf_p, linf_d, l2_p, l2_d, cw_p, cw_d, "candidate_type: POINT_TYPE_CURRENT_ITERATE" if typed else ""))


OPTIMAL_CURRENT = (TR.TERMINATION_REASON_OPTIMAL, PT.POINT_TYPE_CURRENT_ITERATE)


def test_effective_optimality_criteria(lib):  # :121-182
    want = [1.0e-4, 2.0e-4, 1.0e-4, 2.0e-4, 1.0e-4, 2.0e-4]
    assert effective(lib, criteria("simple_optimality_criteria { eps_optimal_absolute: 1.0e-4 eps_optimal_relative: 2.0e-4 }")) == want
    assert effective(lib, criteria("eps_optimal_absolute: 1.0e-4 eps_optimal_relative: 2.0e-4")) == want   # deprecated input
    detailed = """detailed_optimality_criteria { eps_optimal_primal_residual_absolute: 1.0e-4 eps_optimal_primal_residual_relative: 2.0e-4
        eps_optimal_dual_residual_absolute: 3.0e-4 eps_optimal_dual_residual_relative: 4.0e-4
        eps_optimal_objective_gap_absolute: 5.0e-4 eps_optimal_objective_gap_relative: 6.0e-4 }"""
    assert effective(lib, criteria(detailed)) == [1.0e-4, 2.0e-4, 3.0e-4, 4.0e-4, 5.0e-4, 6.0e-4]


@pytest.mark.parametrize("norm", NORMS)
def test_detailed_relative_termination(lib, norm):  # :184-277
    c = criteria(DETAILED_RELATIVE, norm)
    near = conv(1.00019, 1.0, 11.0e-4, 5.4e-4, 14.0e-4, 6.0e-4, 9.0e-5, 9.0e-5)
    assert criteria_met(lib, c, near) == (True, True) and check_iterate(lib, c, near) == OPTIMAL_CURRENT
    gap = conv(1.00021, 1.0, 11.0e-4, 5.4e-4, 14.0e-4, 6.0e-4, 9.0e-5, 9.0e-5)
    assert criteria_met(lib, c, gap) == (False, False) and check_iterate(lib, c, gap) is None
    primal = conv(1.00019, 1.0, 13.0e-4, 5.4e-4, 15.0e-4, 6.0e-4, 1.1e-4, 9.0e-5)
    assert not criteria_met(lib, c, primal)[0] and check_iterate(lib, c, primal) is None
    dual = conv(1.00019, 1.0, 11.0e-4, 5.6e-4, 14.0e-4, 7.0e-4, 9.0e-5, 1.1e-4)
    assert not criteria_met(lib, c, dual)[0] and check_iterate(lib, c, dual) is None


@pytest.mark.parametrize("norm", NORMS)
def test_detailed_absolute_termination(lib, norm):  # :279-372
    c = criteria(DETAILED_ABSOLUTE, norm)
    near = conv(1.00009, 1.0, 9.0e-5, 9.0e-5, 9.0e-5, 9.0e-5, 0.0, 0.0)
    assert criteria_met(lib, c, near) == (True, True) and check_iterate(lib, c, near) == OPTIMAL_CURRENT
    gap = conv(1.00011, 1.0, 9.0e-5, 9.0e-5, 9.0e-5, 9.0e-5, 0.0, 0.0)
    assert criteria_met(lib, c, gap) == (False, False) and check_iterate(lib, c, gap) is None
    primal = conv(1.00009, 1.0, 11.0e-5, 9.0e-5, 11.0e-5, 9.0e-5, 1.0e-6, 0.0)
    assert not criteria_met(lib, c, primal)[0] and check_iterate(lib, c, primal) is None
    dual = conv(1.00009, 1.0, 9.0e-5, 11.0e-5, 9.0e-5, 11.0e-5, 0.0, 1.0e-6)
    assert not criteria_met(lib, c, dual)[0] and check_iterate(lib, c, dual) is None


def test_simple_termination(lib):  # :386-446
    c = criteria(SIMPLE)
    assert check_simple(lib, c, stats()) is None
    assert check_simple(lib, c, stats(), interrupt=True) == (TR.TERMINATION_REASON_INTERRUPTED_BY_USER, PT.POINT_TYPE_NONE)
    assert check_simple(lib, c, stats("cumulative_time_sec: 100.0")) == (TR.TERMINATION_REASON_TIME_LIMIT, PT.POINT_TYPE_NONE)
    assert check_simple(lib, c, stats("cumulative_kkt_matrix_passes: 2500")) == (TR.TERMINATION_REASON_KKT_MATRIX_PASS_LIMIT, PT.POINT_TYPE_NONE)
    assert check_simple(lib, c, stats("iteration_number: 20")) == (TR.TERMINATION_REASON_ITERATION_LIMIT, PT.POINT_TYPE_NONE)


@pytest.mark.parametrize("norm", NORMS)
def test_iterate_termination_without_optimality(lib, norm):  # :374-384, 392-397, 409-417, 448-542
    c = criteria(ITERATE, norm)
    assert check_iterate(lib, c, stats("convergence_information { primal_objective: 50.0 dual_objective: -50.0 }")) is None
    assert check_iterate(lib, c, stats()) is None
    assert check_iterate(lib, c, stats(), force=True) == (TR.TERMINATION_REASON_NUMERICAL_ERROR, PT.POINT_TYPE_NONE)
    ray = "infeasibility_information: { dual_ray_objective: %r max_dual_ray_infeasibility: %r %s }"
    assert check_iterate(lib, c, stats(ray % (1.0, 1.0e-16, "candidate_type: POINT_TYPE_ITERATE_DIFFERENCE"))) == (
        TR.TERMINATION_REASON_PRIMAL_INFEASIBLE, PT.POINT_TYPE_ITERATE_DIFFERENCE)
    assert check_iterate(lib, c, stats(ray % (1.0, 1.0e-5, ""))) is None     # ray too infeasible
    assert check_iterate(lib, c, stats(ray % (-1.0, 0.0, ""))) is None       # wrong sign
    assert check_iterate(lib, c, stats(ray % (0.0, 0.0, ""))) is None        # zero objective
    ray = "infeasibility_information: { primal_ray_linear_objective: %r max_primal_ray_infeasibility: %r %s }"
    assert check_iterate(lib, c, stats(ray % (-1.0, 1.0e-16, "candidate_type: POINT_TYPE_AVERAGE_ITERATE"))) == (
        TR.TERMINATION_REASON_DUAL_INFEASIBLE, PT.POINT_TYPE_AVERAGE_ITERATE)
    assert check_iterate(lib, c, stats(ray % (-1.0, 1.0e-5, ""))) is None
    assert check_iterate(lib, c, stats(ray % (1.0, 0.0, ""))) is None
    assert check_iterate(lib, c, stats(ray % (0.0, 0.0, ""))) is None


@pytest.mark.parametrize("norm", NORMS)
def test_iterate_termination_with_optimality(lib, norm):  # :544-778
    c = criteria(ITERATE, norm)
    exact = conv(1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0)
    assert check_iterate(lib, c, exact) == OPTIMAL_CURRENT
    assert check_iterate(lib, c, conv(1.00019, 1.0, 11.0e-4, 5.4e-4, 14.0e-4, 6.0e-4, 9.0e-5, 9.0e-5)) == OPTIMAL_CURRENT
    assert check_iterate(lib, c, exact, force=True) == OPTIMAL_CURRENT                     # optimal even with a numerical error
    zero = bound_norms(*ZERO_NORMS)
    ones = conv(1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0)
    for a, r, point in ((INF, 0.0, ones), (0.0, INF, conv(0.0, 0.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0)), (INF, INF, ones)):
        tol = criteria("simple_optimality_criteria { eps_optimal_absolute: %s eps_optimal_relative: %s } eps_primal_infeasible: 1.0e-6 "
                       "eps_dual_infeasible: 1.0e-6 time_sec_limit: 1.0 kkt_matrix_pass_limit: 2000 iteration_limit: 10" % (
                           "inf" if a == INF else a, "inf" if r == INF else r), norm)
        assert check_iterate(lib, tol, point, bn=zero) == OPTIMAL_CURRENT                  # infinite tolerances accept anything
    assert check_iterate(lib, c, conv(10.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, typed=False)) is None      # bad gap
    assert check_iterate(lib, c, conv(0.0, -INF, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, typed=False)) is None      # infinite gap
    assert check_iterate(lib, c, conv(1.0, 1.0, 1.0, 0.0, 1.0, 0.0, 1.0, 0.0, typed=False)) is None       # bad primal residual
    assert check_iterate(lib, c, conv(1.0, 1.0, 0.0, 1.0, 0.0, 1.0, 0.0, 1.0, typed=False)) is None       # bad dual residual
    zero_tol = criteria(ITERATE.replace("eps_optimal_absolute: 1.0e-4 eps_optimal_relative: 1.0e-4", "eps_optimal_absolute: 0.0 eps_optimal_relative: 0.0"), norm)
    assert check_iterate(lib, zero_tol, exact) == OPTIMAL_CURRENT                          # zero tolerance, zero error


def test_optimality_norms_differ(lib):  # :796-873: L2 accepts up to 14.49, L_inf up to 12, componentwise up to 1
    for residual, l2, linf, componentwise in ((0.5, True, True, True), (10.0, True, True, False), (13.0, True, False, False), (15.0, False, False, False)):
        s = stats("""convergence_information { primal_objective: 1.0 dual_objective: 1.0 l_inf_primal_residual: %r l2_primal_residual: %r
                     l_inf_componentwise_primal_residual: %r candidate_type: POINT_TYPE_CURRENT_ITERATE }""" % (residual, residual, residual))
        for norm, expected in ((NORM.OPTIMALITY_NORM_L_INF, linf), (NORM.OPTIMALITY_NORM_L2, l2), (NORM.OPTIMALITY_NORM_L_INF_COMPONENTWISE, componentwise)):
            got = check_iterate(lib, criteria("simple_optimality_criteria { eps_optimal_relative: 1.0 }", norm), s)
            assert got == (OPTIMAL_CURRENT if expected else None), (residual, norm)


def test_bound_norms_from_problem_stats(lib):  # :875-887
    q = capi.PdlpQuadraticProgramStats()
    q.objective_vector_l2_norm, q.combined_bounds_l2_norm, q.objective_vector_abs_max, q.combined_bounds_max = 4.0, 3.0, 1.0, 2.0
    out = capi.PdlpBoundNorms()
    lib.fn("bound_norms_from_problem_stats", None)(C.byref(q), C.byref(out))
    assert (out.l2_norm_primal_linear_objective, out.l2_norm_constraint_bounds, out.l_inf_norm_primal_linear_objective, out.l_inf_norm_constraint_bounds) == (4.0, 3.0, 1.0, 2.0)


def test_compute_relative_residuals(lib):  # :902-1130
    point = stats("""convergence_information { primal_objective: 10.0 dual_objective: 5.0 l_inf_primal_residual: 1.0 l2_primal_residual: 1.0
                     l_inf_dual_residual: 1.0 l2_dual_residual: 1.0 }""").convergence_information[0]

    def simple(a, r):
        return criteria("simple_optimality_criteria { eps_optimal_absolute: %s eps_optimal_relative: %s }" % (a, r))

    plain = [1.0 / 12.0, 1.0 / math.sqrt(210.0), 1.0 / 5.5, 1.0 / math.sqrt(36.25), 5.0 / 15.0]
    shifted = [1.0 / (1.0 + 12.0), 1.0 / (1.0 + math.sqrt(210.0)), 1.0 / (1.0 + 5.5), 1.0 / (1.0 + math.sqrt(36.25)), 5.0 / (1.0 + 15.0)]
    assert relative(lib, simple(0.0, 1.0e-6), point) == plain                 # zero absolute tolerance
    assert relative(lib, simple(1.0e-6, 0.0), point) == [0.0] * 5             # zero relative tolerance
    assert relative(lib, simple(1.0e-6, 1.0e-6), point) == shifted            # equal tolerances
    assert relative(lib, simple(0.0, 0.0), point) == shifted                  # both zero
    assert relative(lib, simple("inf", 1.0e-6), point) == [0.0] * 5           # infinite absolute tolerance
    assert relative(lib, simple(1.0e-6, "inf"), point) == plain               # infinite relative tolerance
    assert relative(lib, simple("inf", "inf"), point)[:4] == shifted[:4]      # both infinite
    detailed = criteria("""detailed_optimality_criteria { eps_optimal_primal_residual_absolute: 2.0e-6 eps_optimal_primal_residual_relative: 2.0e-4
        eps_optimal_dual_residual_absolute: 1.0e-3 eps_optimal_dual_residual_relative: 1.0e-4
        eps_optimal_objective_gap_absolute: 3.0e-8 eps_optimal_objective_gap_relative: 3.0e-7 }""")
    got = relative(lib, detailed, point)
    want = [1.0 / (0.01 + 12.0), 1.0 / (0.01 + math.sqrt(210.0)), 1.0 / (10.0 + 5.5), 1.0 / (10.0 + math.sqrt(36.25)), 5.0 / (0.1 + 15.0)]
    assert got == pytest.approx(want, rel=1e-15)
