"""Row-sharded multi-GPU path.

CPU (not gpu): the row partition exported by the C ABI, and a world_size-2
``gloo`` run of the exchange protocol the CUDA path uses inside the step loop
(each rank advances a slice of the primal vector and its block of dual rows;
per PDHG step: all-gather of the x~ slices, then -- together -- the reduce-scatter
of the K^T y' partials added in rank order and three partial sums per rank added
in rank order; the nonlinearity is the row-side sum (K dx) . dy),
restated in numpy and compared with the unsharded iteration.

GPU (needs >= 2 devices): the real thing over NCCL against the 1-GPU solve.
"""
import os
import socket
import sys

import numpy as np
import pytest

from ortools_b200 import distributed, pdlp, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TR = pdlp.TerminationReason


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


# ------------------------------------------------------------------ row partition
@pytest.mark.parametrize("name,scale", [("c2", 0.002), ("c3", 0.002), ("c4", 0.0005)])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_row_blocks_are_a_balanced_contiguous_partition(name, scale, world):
    qp, _ = synthetic.CONFIGS[name](scale=scale)
    k = qp.constraint_matrix.tocsr()
    m = k.shape[0]
    blocks = [distributed.row_block(qp, r, world) for r in range(world)]
    assert blocks[0][0] == 0 and blocks[-1][1] == m
    for (b0, e0), (b1, e1) in zip(blocks, blocks[1:]):
        assert e0 == b1 and b0 <= e0
    mass = np.diff(k.indptr) + 1  # nnz + 1 per row, the balanced quantity
    per = [mass[b:e].sum() for b, e in blocks]
    assert max(per) <= mass.sum() / world + mass.max() + 1


def test_row_blocks_of_a_large_matrix_follow_the_mass_rule_exactly():
    """Above 2 M nonzeros the row histogram is threaded (relaxed atomic counts): the boundaries must
    still be exactly the first rows where nnz + rows reaches g / world of the total, and a bad row
    index must still be reported."""
    import ctypes as C
    import scipy.sparse as sp
    from ortools_b200 import pdlp
    rng = np.random.default_rng(3)
    m, n, nnz = 200_000, 300_000, 2_500_000
    rows = np.minimum((rng.pareto(1.5, size=nnz) * m / 20).astype(np.int64), m - 1)          # skewed row lengths
    k = sp.csc_matrix((np.ones(nnz), (rows, rng.integers(0, n, size=nnz))), shape=(m, n))
    qp = pdlp.QuadraticProgram(n, m)
    qp.constraint_matrix = k
    total = float(k.nnz + m)
    running = np.concatenate([[0], np.cumsum(np.diff(k.tocsr().indptr))]) + np.arange(m + 1)     # nnz + rows before row r
    for world in (1, 3, 8):
        want = [0] + [int(np.searchsorted(running, total * g / world, side="left")) for g in range(1, world)] + [m]
        got = [distributed.row_block(qp, r, world) for r in range(world)]
        assert got == list(zip(want[:-1], want[1:])), world
    view, keep = qp._to_view()
    keep["row_indices"][12345] = m + 7
    b, e = C.c_int64(), C.c_int64()
    assert pdlp.backend().fn("row_block")(C.byref(view), C.c_int32(0), C.c_int32(2), C.byref(b), C.byref(e)) != 0


# ------------------------------------------------------------------ exchange protocol on gloo
def pdhg_reference(k, c, lc, uc, lv, uv, iters, step, weight):
    """Unsharded adaptive PDHG iteration (pdhg.cc:1834-1959, 2558-2640), numpy."""
    n, m = k.shape[1], k.shape[0]
    x, y = np.zeros(n), np.zeros(m)
    kty = k.T @ y
    rejected, done, trace = 0, 0, []
    while done < iters:
        inner = 0
        while True:
            tau, sigma = step / weight, step * weight
            xn = np.clip(x - tau * (c - kty), lv, uv)
            xt = 2 * xn - x
            t = y - sigma * (k @ xt)
            yn = np.maximum(np.minimum(0.0, t + sigma * uc), t + sigma * lc)
            ktyn = k.T @ yn
            dx, dy = xn - x, yn - y
            movement = 0.5 * weight * (dx @ dx) + 0.5 / weight * (dy @ dy)
            nonlin = -(dx @ (ktyn - kty))
            limit = movement / nonlin if nonlin > 0 else np.inf
            total = rejected + inner + done + 1
            first = limit if np.isinf(limit) else (1 - (total + 1) ** -0.3) * limit
            second = (1 + (total + 1) ** -0.6) * step
            accepted = step <= limit
            step = min(first, second)
            trace.append(accepted)
            if accepted:
                x, y, kty = xn, yn, ktyn
                rejected += inner
                done += 1
                break
            inner += 1
    return x, y, trace


def test_maintained_average_products_match_products_of_the_average():
    """The peer-exchange step kernels keep K x and K^T y of the weighted average up to date with the
    averages themselves (StepPtrs::avg_kx / avg_kty): avg <- avg + r (v - avg) applied to the iterate
    and, with the same ratio r = eta / (weight so far + eta), to the iterate's products -- K x' kept by
    the dual kernel as (K x~ + K x) / 2, K^T y' from the K^T side. Restated in numpy over an adaptive
    PDHG run: the maintained products equal K avg_x / K^T avg_y (sou.cc:54-79 linearity) to 1e-12."""
    qp, _ = synthetic.c2(scale=0.001)
    k = qp.constraint_matrix.tocsr()
    n, m = k.shape[1], k.shape[0]
    c, lv, uv = qp.objective_vector, qp.variable_lower_bounds, qp.variable_upper_bounds
    lc, uc = qp.constraint_lower_bounds, qp.constraint_upper_bounds
    x, y = np.zeros(n), np.zeros(m)
    kx, kty = k @ x, k.T @ y
    avg_x, avg_y, avg_kx, avg_kty = np.zeros(n), np.zeros(m), np.zeros(m), np.zeros(n)
    wsum, step, weight, total = 0.0, 0.05, 1.0, 0
    for _ in range(60):
        while True:
            total += 1
            tau, sigma = step / weight, step * weight
            xn = np.clip(x - tau * (c - kty), lv, uv)
            kxt = k @ (2 * xn - x)
            t = y - sigma * kxt
            yn = np.maximum(np.minimum(0.0, t + sigma * uc), t + sigma * lc)
            kxn, ktyn = 0.5 * (kxt + kx), k.T @ yn      # what the dual kernel / the K^T side leave
            dx, dy = xn - x, yn - y
            movement = 0.5 * weight * (dx @ dx) + 0.5 / weight * (dy @ dy)
            nonlin = -((0.5 * (kxt - kx)) @ dy)           # (K dx) . dy, the row-side form
            limit = movement / nonlin if nonlin > 0 else np.inf
            used = step
            step = min(limit if np.isinf(limit) else (1 - (total + 1) ** -0.3) * limit, (1 + (total + 1) ** -0.6) * step)
            if used <= limit:
                break
        x, y, kx, kty = xn, yn, kxn, ktyn
        r = used / (wsum + used)
        wsum += used
        avg_x += r * (x - avg_x)
        avg_y += r * (y - avg_y)
        avg_kx += r * (kx - avg_kx)
        avg_kty += r * (kty - avg_kty)
    assert wsum > 0
    for got, want in ((avg_kx, k @ avg_x), (avg_kty, k.T @ avg_y)):
        assert np.linalg.norm(got - want) <= 1e-12 * max(1.0, np.linalg.norm(want))


@pytest.mark.parametrize("n,m,world", [(2_000_000, 1_000_000, 8), (4_000_000, 10_000_000, 8), (7, 3, 2), (1, 1, 1), (1001, 0, 3)])
def test_peer_arena_segments_are_disjoint_and_ordered(n, m, world):
    """PeerLayout::For (device_ops.h): every segment of the arena the fused exchange writes through starts
    where the previous one ends or later, the primal slices cover n, the flag / epoch blocks hold the
    barriers in use (0 / 1 step loop, 3 / 4 the two concurrent trust-region solves), and both sets of
    trust-region segments are whole."""
    import ctypes as C
    out = (C.c_int64 * 13)()
    assert pdlp.backend().fn("peer_arena_layout")(C.c_int64(n), C.c_int64(m), C.c_int32(world), out) == 0
    stride, n_pad, xt, partial, y, scal, flags, epoch, tr, cand, tr2, cand2, total = list(out)
    assert stride % 2 == 0 and stride * world == n_pad >= n
    assert primal_slice(n, 0, world)[2] == stride
    assert xt == 0 and partial >= xt + n_pad and y >= partial + n_pad and scal >= y + m
    assert flags >= scal + 4 * world and epoch >= flags + 8 * 5 and tr >= epoch + 5       # barriers 0..4, 8 ranks each
    seg = 3 * 4096 + 8
    assert cand >= tr + 2 * 42 * world and tr2 >= cand + seg * world and cand2 >= tr2 + 2 * 42 * world and total >= cand2 + seg * world
    assert pdlp.backend().fn("peer_arena_layout")(C.c_int64(n), C.c_int64(m), C.c_int32(9), out) != 0


def primal_slice(n, rank, world):
    """Slice of the primal vector rank advances inside the step loop (PeerLayout::For in device_ops.h)."""
    stride = 2 * ((n + 2 * world - 1) // (2 * world))
    b = min(n, stride * rank)
    return b, min(n, b + stride), stride


def _protocol_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    qp, _ = synthetic.c2(scale=0.0005)
    k = qp.constraint_matrix.tocsr()
    n = k.shape[1]
    b, e = distributed.row_block(qp, rank, world)
    kg = k[b:e]
    c0, c1, stride = primal_slice(n, rank, world)
    c, lv, uv = qp.objective_vector[c0:c1], qp.variable_lower_bounds[c0:c1], qp.variable_upper_bounds[c0:c1]
    lc, uc = qp.constraint_lower_bounds[b:e], qp.constraint_upper_bounds[b:e]
    x, y = np.zeros(c1 - c0), np.zeros(e - b)       # this rank's slice of x, block of y
    kty = np.zeros(c1 - c0)
    kx = np.zeros(e - b)                             # K x of the current iterate, kept by the dual kernel
    step, weight = 0.1, 1.0
    rejected, done, trace = 0, 0, []

    def all_gather_slices(v):                        # what the x~ stores into every peer arena amount to
        mine = torch.zeros(stride, dtype=torch.float64)
        mine[:v.size] = torch.from_numpy(v)
        parts = [torch.zeros(stride, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(parts, mine)
        return torch.cat(parts).numpy()[:n]

    while done < 25:
        inner = 0
        while True:
            tau, sigma = step / weight, step * weight
            xn = np.clip(x - tau * (c - kty), lv, uv)              # primal half step on the slice
            xt = all_gather_slices(2 * xn - x)                     # exchange 1: x~ slices -> whole x~ everywhere
            kxt = kg @ xt                                          # K x~ = 2 K x' - K x on the row block
            t = y - sigma * kxt                                    # row-local dual half step
            yn = np.maximum(np.minimum(0.0, t + sigma * uc), t + sigma * lc)
            kxn = 0.5 * (kxt + kx)                                 # K x'
            dx, dy = xn - x, yn - y
            # the nonlinearity dx . K^T dy = (K dx) . dy is known on the row side: all three sums
            # travel with exchange 2, the decision needs no exchange of its own
            scal = torch.tensor([dx @ dx, dy @ dy, (0.5 * (kxt - kx)) @ dy], dtype=torch.float64)
            partial = torch.from_numpy(kg.T @ yn)                  # [n] partial of K^T y'
            # exchange 2: reduce-scatter of the partials (each rank adds its slice of every
            # rank's partial in rank order); gloo has no reduce_scatter, so gather + fixed-order sum
            parts = [torch.zeros(n, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(parts, partial)
            per_rank = [torch.zeros(3, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(per_rank, scal)                        # ... and the three partial sums per rank,
            tot = np.zeros(3)                                      # added in rank order on every rank
            for h in range(world):
                tot = tot + per_rank[h].numpy()
            ktyn = np.zeros(c1 - c0)
            for h in range(world):
                ktyn = ktyn + parts[h].numpy()[c0:c1]
            movement = 0.5 * weight * tot[0] + 0.5 / weight * tot[1]
            nonlin = -tot[2]
            limit = movement / nonlin if nonlin > 0 else np.inf
            total = rejected + inner + done + 1
            first = limit if np.isinf(limit) else (1 - (total + 1) ** -0.3) * limit
            second = (1 + (total + 1) ** -0.6) * step
            accepted = step <= limit
            step = min(first, second)
            trace.append(accepted)
            if accepted:
                x, y, kty, kx = xn, yn, ktyn.copy(), kxn
                rejected += inner
                done += 1
                break
            inner += 1
    x_full = all_gather_slices(x)                                  # leaving the loop: the primal side is made whole again
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), x=x_full, y=y, b=b, e=e, trace=np.array(trace))
    dist.destroy_process_group()


def test_exchange_protocol_matches_unsharded_iteration_gloo(tmp_path):
    import torch.multiprocessing as mp

    world = 2
    mp.spawn(_protocol_worker, args=(world, free_port(), str(tmp_path)), nprocs=world, join=True)
    qp, _ = synthetic.c2(scale=0.0005)
    k = qp.constraint_matrix.tocsr()
    x, y, trace = pdhg_reference(k, qp.objective_vector, qp.constraint_lower_bounds, qp.constraint_upper_bounds,
                                 qp.variable_lower_bounds, qp.variable_upper_bounds, 25, 0.1, 1.0)
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    for p in parts:
        assert list(p["trace"]) == trace                      # identical accept / reject decisions on every rank
        np.testing.assert_allclose(p["x"], x, rtol=1e-12, atol=1e-13)
    np.testing.assert_array_equal(parts[0]["x"], parts[1]["x"])  # replicated primal side is bitwise identical
    ysh = np.concatenate([p["y"] for p in parts])
    np.testing.assert_allclose(ysh, y, rtol=1e-12, atol=1e-13)


def _protocol_worker_all_gather(rank, world, port, out_dir):
    """The all-gather exchange (`peer-d`): rank g keeps K[R_g, :] AND K[:, C_g]; per step the x~
    slices and the y' blocks are all-gathered, K^T y' of the slice is computed whole (no partial
    sums), three partial sums per rank are added in rank order."""
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    qp, _ = synthetic.c2(scale=0.0005)
    k = qp.constraint_matrix.tocsr()
    m, n = k.shape
    blocks = [distributed.row_block(qp, r, world) for r in range(world)]
    b, e = blocks[rank]
    c0, c1, stride = primal_slice(n, rank, world)
    k_rows = k[b:e]                       # row block, all columns
    k_cols_t = k.tocsc()[:, c0:c1].T.tocsr()  # (K[:, C_g])^T: slice columns over ALL rows
    c, lv, uv = qp.objective_vector[c0:c1], qp.variable_lower_bounds[c0:c1], qp.variable_upper_bounds[c0:c1]
    lc, uc = qp.constraint_lower_bounds[b:e], qp.constraint_upper_bounds[b:e]
    x, y, kty = np.zeros(c1 - c0), np.zeros(e - b), np.zeros(c1 - c0)
    kx = np.zeros(e - b)
    step, weight = 0.1, 1.0
    rejected, done, trace = 0, 0, []
    max_rows = max(e1 - b1 for b1, e1 in blocks)

    def all_gather(v, width, total, offsets):
        mine = torch.zeros(width, dtype=torch.float64)
        mine[:v.size] = torch.from_numpy(v)
        parts = [torch.zeros(width, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(parts, mine)
        out = np.zeros(total)
        for h in range(world):
            lo, hi = offsets[h]
            out[lo:hi] = parts[h].numpy()[:hi - lo]
        return out

    col_offsets = [primal_slice(n, h, world)[:2] for h in range(world)]
    while done < 25:
        inner = 0
        while True:
            tau, sigma = step / weight, step * weight
            xn = np.clip(x - tau * (c - kty), lv, uv)
            xt = all_gather(2 * xn - x, stride, n, col_offsets)            # exchange 1: x~
            kxt = k_rows @ xt
            t = y - sigma * kxt
            yn = np.maximum(np.minimum(0.0, t + sigma * uc), t + sigma * lc)
            kxn = 0.5 * (kxt + kx)
            dx, dy = xn - x, yn - y
            scal = torch.tensor([dx @ dx, dy @ dy, (0.5 * (kxt - kx)) @ dy], dtype=torch.float64)
            y_full = all_gather(yn, max_rows, m, blocks)                   # exchange 2: y' ...
            per_rank = [torch.zeros(3, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(per_rank, scal)                                # ... with the three sums
            ktyn = k_cols_t @ y_full                                       # whole products of the slice
            tot = np.zeros(3)
            for h in range(world):
                tot = tot + per_rank[h].numpy()
            movement = 0.5 * weight * tot[0] + 0.5 / weight * tot[1]
            nonlin = -tot[2]
            limit = movement / nonlin if nonlin > 0 else np.inf
            total = rejected + inner + done + 1
            first = limit if np.isinf(limit) else (1 - (total + 1) ** -0.3) * limit
            second = (1 + (total + 1) ** -0.6) * step
            accepted = step <= limit
            step = min(first, second)
            trace.append(accepted)
            if accepted:
                x, y, kty, kx = xn, yn, ktyn.copy(), kxn
                rejected += inner
                done += 1
                break
            inner += 1
    x_full = all_gather(x, stride, n, col_offsets)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), x=x_full, y=y, b=b, e=e, trace=np.array(trace))
    dist.destroy_process_group()


def test_all_gather_exchange_protocol_matches_unsharded_iteration_gloo(tmp_path):
    import torch.multiprocessing as mp

    world = 2
    mp.spawn(_protocol_worker_all_gather, args=(world, free_port(), str(tmp_path)), nprocs=world, join=True)
    qp, _ = synthetic.c2(scale=0.0005)
    k = qp.constraint_matrix.tocsr()
    x, y, trace = pdhg_reference(k, qp.objective_vector, qp.constraint_lower_bounds, qp.constraint_upper_bounds,
                                 qp.variable_lower_bounds, qp.variable_upper_bounds, 25, 0.1, 1.0)
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    for p in parts:
        assert list(p["trace"]) == trace
        np.testing.assert_allclose(p["x"], x, rtol=1e-12, atol=1e-13)
    np.testing.assert_array_equal(parts[0]["x"], parts[1]["x"])
    np.testing.assert_allclose(np.concatenate([p["y"] for p in parts]), y, rtol=1e-12, atol=1e-13)


# ------------------------------------------------------------------ the CUDA path over NCCL
def _polishing_params():
    p = pdlp.PrimalDualHybridGradientParams()
    p.use_feasibility_polishing = True
    p.handle_some_primal_gradients_on_finite_bounds_as_residuals = False
    d = p.termination_criteria.detailed_optimality_criteria
    for f in ("primal_residual", "dual_residual"):
        setattr(d, "eps_optimal_%s_absolute" % f, 1e-6)
        setattr(d, "eps_optimal_%s_relative" % f, 1e-6)
    d.eps_optimal_objective_gap_absolute = 0.2   # loose: the polishing phases start at iteration 400 and finish the solve
    d.eps_optimal_objective_gap_relative = 0.2
    p.termination_criteria.iteration_limit = 100000
    return p


def _nccl_worker(rank, world, port, out_dir, exchange):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    os.environ["LOCAL_RANK"] = str(rank)
    os.environ["PDLP_B200_EXCHANGE"] = exchange
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    ctx = distributed.Context(rank, world, rank)
    out = {}
    for name, scale in (("c2", 0.004), ("c3", 0.002), ("c5", 0.003)):
        qp, _ = synthetic.CONFIGS[name](scale=scale)
        p = pdlp.PrimalDualHybridGradientParams()
        p.termination_criteria.simple_optimality_criteria.eps_optimal_absolute = 1e-6
        p.termination_criteria.simple_optimality_criteria.eps_optimal_relative = 1e-6
        p.termination_criteria.iteration_limit = 100000
        res = ctx.primal_dual_hybrid_gradient(qp, p)
        ci = [c for c in res.solve_log.solution_stats.convergence_information if c.candidate_type == res.solve_log.solution_type][0]
        out[name + "_x"] = res.primal_solution
        out[name + "_y"] = res.dual_solution
        out[name + "_meta"] = np.array([res.solve_log.termination_reason, res.solve_log.iteration_count, ci.primal_objective, ci.dual_objective])
        # fixed iteration count, restarts disabled: iterates vs the 1-GPU run
        p = pdlp.PrimalDualHybridGradientParams()
        p.restart_strategy = p.NO_RESTARTS
        p.primal_weight_update_smoothing = 0.0
        p.termination_criteria.simple_optimality_criteria.eps_optimal_absolute = 0.0
        p.termination_criteria.simple_optimality_criteria.eps_optimal_relative = 0.0
        p.termination_criteria.iteration_limit = 48
        res = ctx.primal_dual_hybrid_gradient(qp, p)
        out[name + "_x48"] = res.primal_solution
        out[name + "_y48"] = res.dual_solution
        if name == "c2":  # feasibility polishing phases on the sharded problem
            res = ctx.primal_dual_hybrid_gradient(qp, _polishing_params())
            ci = [c for c in res.solve_log.solution_stats.convergence_information if c.candidate_type == res.solve_log.solution_type][0]
            out["polish_meta"] = np.array([res.solve_log.termination_reason, res.solve_log.solution_type, ci.primal_objective,
                                           len(res.solve_log.feasibility_polishing_details)])
    np.savez(os.path.join(out_dir, "nccl_rank%d.npz" % rank), **out)
    ctx.close()
    dist.destroy_process_group()


# peer-d: x~ and y' all-gathered through peer memory (column-slice image); peer-s: x~ all-gathered,
# K^T y' partials reduce-scattered through peer memory; nccl: one NCCL all-reduce per step.
@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("exchange", ["peer-d", "peer-s", "nccl"])
def test_row_sharded_solve_matches_single_gpu(tmp_path, b200_backend, exchange, world):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs (run with gpurun --gpus %d)" % (world, world))
    mp.spawn(_nccl_worker, args=(world, free_port(), str(tmp_path), exchange), nprocs=world, join=True)
    parts = [np.load(os.path.join(str(tmp_path), "nccl_rank%d.npz" % r)) for r in range(world)]
    for name, scale in (("c2", 0.004), ("c3", 0.002), ("c5", 0.003)):
        qp, _ = synthetic.CONFIGS[name](scale=scale)
        p = pdlp.PrimalDualHybridGradientParams()
        p.termination_criteria.simple_optimality_criteria.eps_optimal_absolute = 1e-6
        p.termination_criteria.simple_optimality_criteria.eps_optimal_relative = 1e-6
        p.termination_criteria.iteration_limit = 100000
        one = b200_backend.primal_dual_hybrid_gradient(qp, p)
        ci = [c for c in one.solve_log.solution_stats.convergence_information if c.candidate_type == one.solve_log.solution_type][0]
        for part in parts:
            reason, iters, pobj, dobj = part[name + "_meta"]
            assert int(reason) == one.solve_log.termination_reason == TR.TERMINATION_REASON_OPTIMAL
            assert pobj == pytest.approx(ci.primal_objective, rel=1e-5, abs=1e-5)
            assert dobj == pytest.approx(ci.dual_objective, rel=1e-5, abs=1e-5)
        # every rank returns the same full-length vectors
        for part in parts[1:]:
            np.testing.assert_array_equal(parts[0][name + "_x"], part[name + "_x"])
            np.testing.assert_array_equal(parts[0][name + "_y"], part[name + "_y"])
        p = pdlp.PrimalDualHybridGradientParams()
        p.restart_strategy = p.NO_RESTARTS
        p.primal_weight_update_smoothing = 0.0
        p.termination_criteria.simple_optimality_criteria.eps_optimal_absolute = 0.0
        p.termination_criteria.simple_optimality_criteria.eps_optimal_relative = 0.0
        p.termination_criteria.iteration_limit = 48
        one = b200_backend.primal_dual_hybrid_gradient(qp, p)
        for key, ref in (("_x48", one.primal_solution), ("_y48", one.dual_solution)):
            got = parts[0][name + key]
            assert np.linalg.norm(got - ref) <= 1e-9 * max(1.0, np.linalg.norm(ref)), (name, key)
    # feasibility polishing: same outcome as on one GPU
    qp, _ = synthetic.CONFIGS["c2"](scale=0.004)
    one = b200_backend.primal_dual_hybrid_gradient(qp, _polishing_params())
    ci = [c for c in one.solve_log.solution_stats.convergence_information if c.candidate_type == one.solve_log.solution_type][0]
    for part in parts:
        reason, sol_type, pobj, phases = part["polish_meta"]
        assert int(reason) == one.solve_log.termination_reason == TR.TERMINATION_REASON_OPTIMAL
        assert int(sol_type) == one.solve_log.solution_type
        assert pobj == pytest.approx(ci.primal_objective, rel=2e-3, abs=2e-3)
        assert int(phases) == len(one.solve_log.feasibility_polishing_details) >= 2
        assert int(sol_type) == pdlp.PointType.POINT_TYPE_FEASIBILITY_POLISHING_SOLUTION
