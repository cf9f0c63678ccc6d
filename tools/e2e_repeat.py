#!/usr/bin/env python
"""Repeats the e2e leg of bench.py (full 1e-4 solve of C2 through the C ABI from page-locked host
buffers) in one process and prints every wall time: run-to-run variance of the headline number.
    PDLP_B200_TRACE=1 python tools/e2e_repeat.py [repeats]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ortools_b200 import pdlp, synthetic  # noqa: E402


def main():
    import torch
    repeats = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    be = pdlp.backend()
    qp, _ = synthetic.c2()
    params = bench.make_params(pdlp, 1e-4, iteration_limit=200000)
    keep = qp._to_view()
    qp._to_view = lambda: keep
    for k in range(repeats):
        if k == 2:
            pinned = bench.pin_host_arrays(keep[1])
            print("pinned %d arrays" % len(pinned), flush=True)
        torch.cuda.synchronize()
        t0 = time.time()
        res = be.primal_dual_hybrid_gradient(qp, params)
        torch.cuda.synchronize()
        dt = time.time() - t0
        print("solve %d: %.4f s  (%d iterations, %.0f it/s)" % (k, dt, res.solve_log.iteration_count, res.solve_log.iteration_count / dt), flush=True)
        del res


if __name__ == "__main__":
    main()
