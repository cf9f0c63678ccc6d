"""Synthetic problem generators for the benchmark configurations of
BASELINE.json / SURVEY.md section 8(d). Every generator is deterministic in its
seed and takes a ``scale`` in (0, 1] that shrinks the instance for parity tests
(the shapes, density per row and structure are kept).

  C2  random sparse LP, 1M x 2M, 20 nnz per row, planted optimal pair
  C3  skewed transportation LP (~10M arcs, ~40M nnz, Zipf bundle rows)
  C4  the C2 generator but tall (10M x 4M, 200M nnz) for row sharding
  C5  diagonal QP: LASSO split, K = [A, -A, -I], Q = diag(0, 0, 1)
"""
import numpy as np
import scipy.sparse as sp

from . import pdlp

INF = float("inf")


def _rows_without_replacement(rng, m, n, per_row):
    """(m, per_row) column indices, distinct inside every row."""
    cols = rng.integers(0, n, size=(m, per_row), dtype=np.int64)
    for _ in range(100):
        s = np.sort(cols, axis=1)
        bad = np.nonzero((s[:, 1:] == s[:, :-1]).any(axis=1))[0]
        if bad.size == 0:
            break
        cols[bad] = rng.integers(0, n, size=(bad.size, per_row), dtype=np.int64)
    return cols


def random_sparse_lp(m=1_000_000, n=2_000_000, per_row=20, seed=20240902, name="c2_random_sparse_lp"):
    """C2 / C4 generator. Planted optimal pair (x*, y*):
    x*_j = U(0,1) * Bernoulli(0.5) with 0 <= x <= 10; y*_i ~ N(0,1) on a random
    half of the rows (equalities l_c = u_c = K x*), the other rows are
    inequalities l_c = K x* - U(0,1), u_c = +inf with y*_i = 0; c = K^T y* + r
    with r_j = U(0,1) where x*_j = 0 and 0 elsewhere. Returns (qp, info)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    per_row = min(per_row, n)
    cols = _rows_without_replacement(rng, m, n, per_row)
    vals = rng.standard_normal(size=(m, per_row))
    indptr = np.arange(0, m * per_row + 1, per_row, dtype=np.int64)
    k = sp.csr_matrix((vals.ravel(), cols.ravel(), indptr), shape=(m, n))
    del cols, vals
    xs = rng.uniform(0.0, 1.0, n) * (rng.uniform(size=n) < 0.5)
    eq = rng.uniform(size=m) < 0.5
    ys = np.where(eq, rng.standard_normal(m), 0.0)
    kx = k @ xs
    lc = np.where(eq, kx, kx - rng.uniform(0.0, 1.0, m))
    uc = np.where(eq, kx, INF)
    c = k.T @ ys + np.where(xs == 0.0, rng.uniform(0.0, 1.0, n), 0.0)
    qp = pdlp.QuadraticProgram(n, m)
    qp.constraint_matrix = k.tocsc()
    qp.constraint_matrix.sort_indices()
    qp.constraint_lower_bounds = lc
    qp.constraint_upper_bounds = uc
    qp.variable_lower_bounds = np.zeros(n)
    qp.variable_upper_bounds = np.full(n, 10.0)
    qp.objective_vector = c
    qp.problem_name = name
    return qp, {"x_star": xs, "y_star": ys, "objective": float(c @ xs)}


def c2(scale=1.0, seed=20240902):
    m = max(8, int(round(1_000_000 * scale)))
    return random_sparse_lp(m, 2 * m, 20, seed, "c2_random_sparse_lp")


def c4(scale=1.0, seed=20240904):
    m = max(10, int(round(10_000_000 * scale)))
    return random_sparse_lp(m, max(4, (4 * m) // 10), 20, seed, "c4_tall_lp")


def _zipf_lengths(rng, total, lo, hi, a=1.2):
    """Row lengths ~ Zipf(a) truncated to [lo, hi] whose sum is exactly total."""
    out = []
    acc = 0
    while acc < total:
        batch = rng.zipf(a, size=4096)
        batch = batch[(batch >= lo) & (batch <= hi)]
        for v in batch:
            v = int(min(v, total - acc))
            if v <= 0:
                break
            out.append(v)
            acc += v
            if acc >= total:
                break
    return np.asarray(out, dtype=np.int64)


def skewed_transportation_lp(num_sources=2000, num_sinks=5000, max_bundle=1_000_000, seed=20240903,
                             name="c3_skewed_transportation_lp"):
    """C3: transportation core (supply rows of length T, demand
    rows of length S) plus two bundle/capacity memberships per arc in rows whose
    lengths follow Zipf(1.2) truncated to [2, max_bundle]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    s_cnt, t_cnt = num_sources, num_sinks
    n = s_cnt * t_cnt
    arc = np.arange(n, dtype=np.int64)
    src, dst = arc // t_cnt, arc % t_cnt
    supply = rng.uniform(1.0, 2.0, s_cnt)
    demand = rng.uniform(1.0, 2.0, t_cnt)
    demand *= supply.sum() / demand.sum()  # balanced
    x_feas = supply[src] * demand[dst] / supply.sum()
    rows = [src, s_cnt + dst]
    cols = [arc, arc]
    base = s_cnt + t_cnt
    caps = []
    for _ in range(2):
        lens = _zipf_lengths(rng, n, 2, max_bundle)
        if lens[-1] < 2 and lens.size > 1:  # fold a trailing singleton into its neighbour
            lens[-2] += lens[-1]
            lens = lens[:-1]
        member = np.repeat(np.arange(lens.size, dtype=np.int64), lens)
        perm = rng.permutation(n)
        rows.append(base + member)
        cols.append(perm)
        caps.append(1.2 * np.bincount(member, weights=x_feas[perm], minlength=lens.size))
        base += lens.size
    m = base
    rows = np.concatenate(rows)
    cols = np.concatenate(cols)
    k = sp.csr_matrix((np.ones(rows.size), (rows, cols)), shape=(m, n)).tocsc()
    k.sort_indices()
    lc = np.concatenate([np.full(s_cnt, -INF), demand, np.full(m - s_cnt - t_cnt, -INF)])
    uc = np.concatenate([supply, np.full(t_cnt, INF), np.concatenate(caps)])
    qp = pdlp.QuadraticProgram(n, m)
    qp.constraint_matrix = k
    qp.constraint_lower_bounds = lc
    qp.constraint_upper_bounds = uc
    qp.variable_lower_bounds = np.zeros(n)
    qp.variable_upper_bounds = np.full(n, INF)
    qp.objective_vector = rng.uniform(1.0, 10.0, n)
    qp.problem_name = name
    return qp, {"x_feasible": x_feas}


def c3(scale=1.0, seed=20240903):
    f = np.sqrt(scale)
    return skewed_transportation_lp(max(4, int(round(2000 * f))), max(6, int(round(5000 * f))),
                                    max(4, int(1_000_000 * scale)), seed)


def lasso_qp(num_samples=1_000_000, num_features=2_000_000, per_row=24, lam=0.1, seed=20240905,
             name="c5_lasso_diagonal_qp"):
    """C5: min 0.5 ||z||^2 + lam 1^T (u + v) s.t. A u - A v - z = b,
    u, v >= 0; variables (u, v, z), Q = diag(0, 0, 1)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    p, q = num_samples, num_features
    per_row = min(per_row, q)
    cols = _rows_without_replacement(rng, p, q, per_row)
    vals = rng.standard_normal(size=(p, per_row)) / np.sqrt(per_row)
    indptr = np.arange(0, p * per_row + 1, per_row, dtype=np.int64)
    a = sp.csr_matrix((vals.ravel(), cols.ravel(), indptr), shape=(p, q))
    del cols, vals
    beta = rng.standard_normal(q) * (rng.uniform(size=q) < 0.01)
    b = a @ beta + 0.01 * rng.standard_normal(p)
    k = sp.hstack([a, -a, -sp.identity(p, format="csr")], format="csc")
    k.sort_indices()
    n = 2 * q + p
    qp = pdlp.QuadraticProgram(n, p)
    qp.constraint_matrix = k
    qp.constraint_lower_bounds = b.copy()
    qp.constraint_upper_bounds = b.copy()
    qp.variable_lower_bounds = np.concatenate([np.zeros(2 * q), np.full(p, -INF)])
    qp.variable_upper_bounds = np.full(n, INF)
    qp.objective_vector = np.concatenate([np.full(2 * q, lam), np.zeros(p)])
    qp.set_objective_matrix_diagonal(np.concatenate([np.zeros(2 * q), np.ones(p)]))
    qp.problem_name = name
    return qp, {"beta_star": beta}


def c5(scale=1.0, seed=20240905):
    p = max(8, int(round(1_000_000 * scale)))
    return lasso_qp(p, 2 * p, 24, 0.1, seed)


CONFIGS = {"c2": c2, "c3": c3, "c4": c4, "c5": c5}
