"""A/B of pageable against page-locked caller buffers for the e2e call (bench.py pins them in place)."""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from ortools_b200 import pdlp, synthetic
import bench

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
qp, info = synthetic.c2(scale=scale)
params = bench.make_params(pdlp, 1e-4, iteration_limit=200)
be = pdlp.backend()
vk = qp._to_view()
qp._to_view = lambda: vk
rt = ctypes.CDLL("libcudart.so.12")
pinned = []
ref = None
for label in ("pageable", "pageable", "pinned", "pinned"):
    if label == "pinned" and not pinned:
        for arr in vk[1].values():
            if hasattr(arr, "ctypes") and arr.nbytes >= (1 << 16):
                if rt.cudaHostRegister(ctypes.c_void_p(arr.ctypes.data), ctypes.c_size_t(arr.nbytes), 0) == 0:
                    pinned.append(arr)
    t = time.time()
    r = be.primal_dual_hybrid_gradient(qp, params)
    dt = time.time() - t
    if ref is None:
        ref = r.primal_solution.copy()
    print("%s: %.4f s, %d iterations, %d arrays pinned, same result: %s" % (label, dt, r.solve_log.iteration_count, len(pinned),
                                                                          bool(np.array_equal(ref, r.primal_solution))), flush=True)
