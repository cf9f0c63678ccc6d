#!/bin/bash
# compute-sanitizer passes over the CUDA path on small problems ("race
# detection"): memcheck, racecheck (shared-memory hazards in the block reductions, the trust-region
# bins, the radix sort of the device builder), initcheck and synccheck. Run on a GPU box:
#   gpurun --timeout 900 -- 'tools/sanitize.sh > gpurun_out/sanitize.log 2>&1'
# Each pass solves TestLp / TinyLp / a diagonal QP / a 2000 x 4000 random LP (adaptive and Malitsky-Pock
# rules; restarts, i.e. the persistent trust-region kernel, included) through the C ABI, builds the SELL pair on
# the device and runs the kernel-level entry points; the summaries go to profiles/ by hand.
set -u
cd "$(dirname "$0")/.."
SAN=${SAN:-/usr/local/cuda/bin/compute-sanitizer}
PY=${PY:-python}
SCRIPT='
import sys, numpy as np, scipy.sparse as sp
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import fixtures
from ortools_b200 import pdlp
be = pdlp.backend()
p = pdlp.PrimalDualHybridGradientParams()
p.termination_criteria.simple_optimality_criteria.eps_optimal_absolute = 1e-6
p.termination_criteria.simple_optimality_criteria.eps_optimal_relative = 1e-6
p.termination_criteria.iteration_limit = 400
for make in (fixtures.test_lp, fixtures.tiny_lp, fixtures.test_diagonal_qp1):
    r = be.primal_dual_hybrid_gradient(make(), p)
    print(make.__name__, pdlp.TerminationReason.Name(r.solve_log.termination_reason), r.solve_log.iteration_count, flush=True)
rng = np.random.default_rng(0)
m, n, per = 2000, 4000, 8
K = sp.csc_matrix((rng.normal(size=m * per), (np.repeat(np.arange(m), per), rng.integers(0, n, size=m * per))), shape=(m, n))
K.sum_duplicates()
qp = pdlp.QuadraticProgram(n, m)
qp.constraint_matrix = K
xs = rng.uniform(0, 1, n) * (rng.uniform(size=n) < 0.5)
qp.constraint_lower_bounds = K @ xs - rng.uniform(0, 1, m)
qp.constraint_upper_bounds = np.full(m, np.inf)
qp.variable_lower_bounds = np.zeros(n); qp.variable_upper_bounds = np.full(n, 10.0)
qp.objective_vector = rng.uniform(0, 1, n)
p.termination_criteria.iteration_limit = 192
r = be.primal_dual_hybrid_gradient(qp, p)
print("random", pdlp.TerminationReason.Name(r.solve_log.termination_reason), r.solve_log.iteration_count, flush=True)
p.linesearch_rule = p.MALITSKY_POCK_LINESEARCH_RULE   # the device-resident Malitsky-Pock loop
p.termination_criteria.iteration_limit = 96
r = be.primal_dual_hybrid_gradient(qp, p)
print("random, Malitsky-Pock", pdlp.TerminationReason.Name(r.solve_log.termination_reason), r.solve_log.iteration_count, flush=True)
p.linesearch_rule = p.ADAPTIVE_LINESEARCH_RULE
prob = be.problem(qp)
prob.matrix_vector_product(rng.normal(size=n)); prob.transposed_matrix_vector_product(rng.normal(size=m))
print(prob.compute_stats().constraint_matrix_num_nonzeros, flush=True)
'
status=0
for tool in memcheck racecheck initcheck synccheck; do
  echo "=== compute-sanitizer --tool $tool"
  timeout 600 "$SAN" --tool "$tool" --error-exitcode 9 --print-limit 20 "$PY" -c "$SCRIPT" 2>&1 | grep -v "^$" | tail -40
  rc=${PIPESTATUS[0]}
  echo "=== $tool exit code $rc"
  [ "$rc" -ne 0 ] && status=1
done
# the TMA-staged SELL variant (cp.async.bulk + mbarrier), memcheck + racecheck only
for tool in memcheck racecheck; do
  echo "=== compute-sanitizer --tool $tool (PDLP_B200_SELL_VARIANT=5, TMA staging)"
  PDLP_B200_SELL_VARIANT=5 timeout 600 "$SAN" --tool "$tool" --error-exitcode 9 --print-limit 20 "$PY" -c "$SCRIPT" 2>&1 | grep -v "^$" | tail -12
  rc=${PIPESTATUS[0]}
  echo "=== $tool (TMA variant) exit code $rc"
  [ "$rc" -ne 0 ] && status=1
done
# The row-sharded step loop (needs >= 2 GPUs; SANITIZE_MULTI=1): the 2-rank parity tests with the persistent k_peer_loop forced,
# child processes followed. memcheck / initcheck / synccheck see each rank's own accesses (also those that land in a peer's
# arena); no tool of the suite orders accesses BETWEEN processes -- that part is argued in DESIGN.md 5 and modelled under
# ThreadSanitizer by tests/test_peer_protocol_model.py. (Written after the GPU budget of round 2 was spent: not run yet.)
if [ "${SANITIZE_MULTI:-0}" = "1" ]; then
  for tool in memcheck initcheck synccheck; do
    echo "=== compute-sanitizer --tool $tool --target-processes all (2 ranks, PDLP_B200_PEER_LOOP=2)"
    PDLP_B200_PEER_LOOP=2 PDLP_B200_PEER_TIMEOUT_S=120 timeout 1500 "$SAN" --tool "$tool" --target-processes all --error-exitcode 9 --print-limit 20 \
      "$PY" -m pytest tests/test_distributed.py -m gpu -q -k "d-2 or s-2" 2>&1 | grep -v "^$" | tail -20
    rc=${PIPESTATUS[0]}
    echo "=== $tool (row-sharded loop) exit code $rc"
    [ "$rc" -ne 0 ] && status=1
  done
fi
exit $status
