"""Differential property tests of the host-side scalar logic: the CPU restatement (oracle) and this
library's own host code (exported as host-only C entry points) must agree on random inputs --
parameter validation (solvers_proto_validation.cc) and the termination predicates (termination.cc).
The known answers of the reference's tests for both are in test_params_validation.py and
test_termination.py; these tests widen them. No GPU involved."""
import ctypes as C
import math
import os

import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from ortools_b200 import _capi as capi
from ortools_b200 import pdlp, pdlp_proto
from test_proto_codec_properties import messages
from test_termination import check_iterate, check_simple, criteria_met, effective, relative

EXAMPLES = int(os.environ.get("PDLP_B200_PROPERTY_EXAMPLES", "60"))
SETTINGS = dict(max_examples=4 * EXAMPLES, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])
edge_doubles = st.one_of(st.floats(allow_nan=True), st.sampled_from([0.0, 1.0, -1.0, 1e-6, 1e-50, 1e-51, 1e50, 1e51, math.inf, -math.inf, math.nan, 0.5, 0.99, 1.0 - 1e-16]))


def libs():
    from oracle import pdlp_oracle
    return pdlp_oracle.backend(), pdlp.backend()


@st.composite
def parameter_pods(draw):
    """A random parameter message, then a few fields pushed to the edges the validator tests."""
    msg = draw(messages(pdlp_proto.PrimalDualHybridGradientParamsProto))
    del msg.random_projection_seeds[:]
    pod = pdlp_proto.params_from_proto(msg).to_pod()
    for _ in range(draw(st.integers(0, 3))):
        path = draw(st.sampled_from([
            "termination_criteria.eps_primal_infeasible", "termination_criteria.time_sec_limit", "termination_criteria.kkt_matrix_pass_limit",
            "termination_criteria.simple_eps_optimal_absolute", "termination_criteria.eps_optimal_objective_gap_relative",
            "termination_criteria.eps_optimal_absolute", "log_interval_seconds", "primal_weight_update_smoothing", "initial_primal_weight",
            "sufficient_reduction_for_restart", "necessary_reduction_for_restart", "adaptive_step_size_reduction_exponent",
            "adaptive_step_size_growth_exponent", "malitsky_pock_step_size_downscaling_factor", "malitsky_pock_linesearch_contraction_factor",
            "malitsky_pock_step_size_interpolation", "initial_step_size_scaling", "infinite_constraint_bound_threshold",
            "diagonal_qp_trust_region_solver_tolerance"]))
        target = pod
        *parents, leaf = path.split(".")
        for p in parents:
            target = getattr(target, p)
        setattr(target, leaf, draw(edge_doubles))
    for name in draw(st.lists(st.sampled_from(["num_threads", "num_shards", "verbosity_level", "major_iteration_frequency", "termination_check_frequency",
                                               "l_inf_ruiz_iterations", "restart_strategy", "linesearch_rule", "scheduler_type"]), max_size=2)):
        setattr(pod, name, draw(st.integers(-2, 120)))
    if draw(st.booleans()):
        pod.termination_criteria.optimality_norm = draw(st.integers(-1, 5))
    if draw(st.booleans()):
        pod.termination_criteria.iteration_limit = draw(st.integers(-3, 10))
    return pod


@settings(**SETTINGS)
@given(pod=parameter_pods())
def test_both_validators_agree(pod):
    oracle, product = libs()
    assert oracle.validate_params(pod) == product.validate_params(pod)


small = st.one_of(st.floats(0, 10, allow_nan=False), st.sampled_from([0.0, 1e-9, 1e-6, 1e-4, 1e-3, 1.0, math.inf, math.nan]))
signed = st.one_of(st.floats(-100, 100, allow_nan=False), st.sampled_from([0.0, 1.0, -1.0, math.inf, -math.inf, math.nan]))


@st.composite
def termination_inputs(draw):
    tc = pdlp.TerminationCriteria()
    pdlp_proto._copy_set_fields(draw(messages(pdlp_proto.TerminationCriteriaProto)), tc)
    c = tc.to_pod()
    c.optimality_norm = draw(st.sampled_from([1, 2, 3]))
    for name in ("simple_eps_optimal_absolute", "simple_eps_optimal_relative", "eps_optimal_primal_residual_absolute", "eps_optimal_primal_residual_relative",
                 "eps_optimal_dual_residual_absolute", "eps_optimal_dual_residual_relative", "eps_optimal_objective_gap_absolute",
                 "eps_optimal_objective_gap_relative", "eps_optimal_absolute", "eps_optimal_relative", "eps_primal_infeasible", "eps_dual_infeasible"):
        setattr(c, name, draw(small))
    c.time_sec_limit, c.kkt_matrix_pass_limit = draw(small), draw(st.one_of(small, st.just(1e6)))
    c.iteration_limit = draw(st.integers(0, 20))
    s = capi.PdlpIterationStats()
    s.iteration_number = draw(st.integers(0, 20))
    s.cumulative_kkt_matrix_passes, s.cumulative_time_sec = draw(small), draw(small)
    s.num_convergence_information = draw(st.integers(0, 3))
    for k in range(s.num_convergence_information):
        ci = s.convergence_information[k]
        ci.candidate_type = draw(st.sampled_from([0, 1, 3, 6]))
        for f, _ in capi.PdlpConvergenceInformation._fields_:
            if f != "candidate_type":
                setattr(ci, f, draw(signed if "objective" in f else small))
    s.num_infeasibility_information = draw(st.integers(0, 3))
    for k in range(s.num_infeasibility_information):
        ii = s.infeasibility_information[k]
        ii.candidate_type = draw(st.sampled_from([0, 1, 2, 3]))
        for f, _ in capi.PdlpInfeasibilityInformation._fields_:
            if f != "candidate_type":
                setattr(ii, f, draw(signed))
    bn = capi.PdlpBoundNorms()
    for f, _ in capi.PdlpBoundNorms._fields_:
        setattr(bn, f, draw(small))
    return c, s, bn, draw(st.booleans()), draw(st.sampled_from([None, False, True]))


def same(a, b):
    if isinstance(a, (list, tuple)) and isinstance(b, (list, tuple)):
        return len(a) == len(b) and all(same(x, y) for x, y in zip(a, b))
    if isinstance(a, float) and isinstance(b, float) and math.isnan(a) and math.isnan(b):
        return True
    return a == b


@settings(**SETTINGS)
@given(inputs=termination_inputs())
def test_termination_predicates_agree(inputs):
    c, s, bn, force, interrupt = inputs
    oracle, product = libs()
    assert same(effective(oracle, c), effective(product, c))
    assert check_simple(oracle, c, s, interrupt) == check_simple(product, c, s, interrupt)
    assert check_iterate(oracle, c, s, bn, force) == check_iterate(product, c, s, bn, force)
    if s.num_convergence_information > 0:
        assert criteria_met(oracle, c, s, bn) == criteria_met(product, c, s, bn)
        assert same(relative(oracle, c, s.convergence_information[0], bn), relative(product, c, s.convergence_information[0], bn))
