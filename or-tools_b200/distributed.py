"""Row-sharded multi-GPU solve: one process per GPU, launched
with torchrun; ``torch.distributed`` is only the plumbing that hands the NCCL
unique id to every rank -- the exchange itself (peer-memory stores / loads fused
into the step kernels, or the NCCL all-reduce fallback) is issued
by libpdlp_b200.so on its own stream.
"""
import ctypes as C
import glob
import os

import numpy as np

from . import _capi as capi
from . import pdlp


def nccl_library_path():
    """The NCCL shared object PyTorch itself uses (so both share one copy)."""
    env = os.environ.get("PDLP_B200_NCCL_LIBRARY")
    if env:
        return env
    try:
        import nvidia.nccl as _n  # the wheel torch depends on
        base = list(_n.__path__)[0]
        hits = sorted(glob.glob(os.path.join(base, "lib", "libnccl.so*")))
        if hits:
            return hits[0]
    except Exception:
        pass
    return "libnccl.so.2"


def row_block(qp, rank, world_size):
    """[begin, end) of the constraint rows rank keeps (host-only)."""
    view, keep = qp._to_view()
    b, e = C.c_int64(), C.c_int64()
    rc = pdlp.backend().fn("row_block")(C.byref(view), C.c_int32(rank), C.c_int32(world_size), C.byref(b), C.byref(e))
    del keep
    pdlp.backend()._check(rc, "row_block")
    return b.value, e.value


class Context:
    """NCCL communicator owned by the library (pdlp_b200_distributed_init)."""

    def __init__(self, rank=None, world_size=None, cuda_device=None):
        import torch.distributed as dist

        self.b = pdlp.backend()
        self.rank = dist.get_rank() if rank is None else rank
        self.world_size = dist.get_world_size() if world_size is None else world_size
        self.cuda_device = int(os.environ.get("LOCAL_RANK", self.rank)) if cuda_device is None else cuda_device
        lib = nccl_library_path().encode()
        uid = (C.c_uint8 * 128)()
        if self.rank == 0:
            self.b._check(self.b.fn("nccl_unique_id")(lib, uid), "nccl_unique_id")
        box = [bytes(uid)]
        dist.broadcast_object_list(box, src=0)
        uid = (C.c_uint8 * 128).from_buffer_copy(box[0])
        self.h = C.c_void_p()
        self.b._check(self.b.fn("distributed_init")(lib, C.c_int32(self.rank), C.c_int32(self.world_size), C.c_int32(self.cuda_device),
                                                    uid, C.byref(self.h)), "distributed_init")

    def close(self):
        if self.h:
            self.b.fn("distributed_destroy", None)(self.h)
            self.h = C.c_void_p()

    def primal_dual_hybrid_gradient(self, qp, params, initial_solution=None):
        """Every rank passes the same problem and receives the same SolverResult."""
        view, keep = qp._to_view()
        pod = pdlp.params_to_pod(params)
        x0 = y0 = None
        if initial_solution is not None:
            x0 = capi.as_f64(initial_solution.primal_solution)
            y0 = capi.as_f64(initial_solution.dual_solution)
        res = capi.PdlpResult()
        rc = self.b.fn("primal_dual_hybrid_gradient_distributed")(
            self.h, C.byref(view), C.byref(pod), capi.ptr_f64(x0), C.c_int64(0 if x0 is None else x0.size),
            capi.ptr_f64(y0), C.c_int64(0 if y0 is None else y0.size), None, capi.MESSAGE_CALLBACK(), capi.STATS_CALLBACK(), None, C.byref(res))
        del keep
        try:
            self.b._check(rc, "primal_dual_hybrid_gradient_distributed")
        except BaseException:
            self.b.fn("result_free", None)(C.byref(res))
            raise
        return self.b._result_from_pod(res)  # (takes ownership of res)

    def session(self, qp, params):
        s = pdlp.SolveSession.__new__(pdlp.SolveSession)
        s.b = self.b
        view, keep = qp._to_view()
        pod = pdlp.params_to_pod(params)
        s.h = C.c_void_p()
        rc = self.b.fn("session_create_distributed")(self.h, C.byref(view), C.byref(pod), C.byref(s.h))
        del keep
        self.b._check(rc, "session_create_distributed")
        return s


_context = None


def context():
    global _context
    if _context is None:
        _context = Context()
    return _context


def session(qp, params, rank=None, world_size=None, cuda_device=None):
    global _context
    if _context is None:
        _context = Context(rank, world_size, cuda_device)
    return _context.session(qp, params)
