"""The device-side SELL-32 builder (csrc/device_build.cu) against the host
builder (csrc/sell_builder.cc): both must produce the same layout, hence
bitwise-identical products, for ragged / empty / split rows."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from ortools_b200 import pdlp, synthetic


def qp_from_matrix(k):
    m, n = k.shape
    qp = pdlp.QuadraticProgram(n, m)
    qp.constraint_matrix = sp.csc_matrix(k)
    qp.variable_lower_bounds = np.zeros(n)
    qp.variable_upper_bounds = np.ones(n)
    qp.constraint_lower_bounds = -np.ones(m)
    qp.constraint_upper_bounds = np.ones(m)
    qp.objective_vector = np.arange(n, dtype=float)
    return qp


def cases():
    rng = np.random.default_rng(5)
    out = {}
    out["ragged"] = qp_from_matrix(sp.random(700, 900, density=0.01, random_state=3, format="csc"))
    k = sp.random(300, 200, density=0.05, random_state=4, format="lil")
    k[5, :] = 0          # empty rows / columns
    k[:, 7] = 0
    k[11, :] = rng.normal(size=200)   # one dense row, one dense column
    k[:, 13] = rng.normal(size=(300, 1))
    out["dense_and_empty"] = qp_from_matrix(k.tocsc())
    out["empty_matrix"] = qp_from_matrix(sp.csc_matrix((4, 6)))
    out["c3"] = synthetic.c3(scale=0.002)[0]
    out["c5"] = synthetic.c5(scale=0.002)[0]
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ragged", "dense_and_empty", "empty_matrix", "c3", "c5"])
@pytest.mark.parametrize("split_len,sigma", [(0, 0), (4, 32), (64, 4096)])
def test_device_build_equals_host_build(name, split_len, sigma, b200_backend, monkeypatch):
    qp = cases()[name]
    k = qp.constraint_matrix
    if split_len:
        monkeypatch.setenv("PDLP_B200_SPLIT_LEN", str(split_len))
        monkeypatch.setenv("PDLP_B200_SIGMA", str(sigma))
    rng = np.random.default_rng(0)
    x, y = rng.normal(size=k.shape[1]), rng.normal(size=k.shape[0])
    monkeypatch.setenv("PDLP_B200_HOST_BUILD", "1")
    host = b200_backend.problem(qp)
    monkeypatch.setenv("PDLP_B200_HOST_BUILD", "0")
    dev = b200_backend.problem(qp)
    np.testing.assert_array_equal(dev.matrix_vector_product(x), host.matrix_vector_product(x))
    np.testing.assert_array_equal(dev.transposed_matrix_vector_product(y), host.transposed_matrix_vector_product(y))
    # values come back in the caller's CSC order, before and after rescaling
    kk = sp.csc_matrix(k)
    kk.sort_indices()
    np.testing.assert_array_equal(dev.download()["values"], kk.data)
    r1, c1 = dev.apply_rescaling(3, True)
    r2, c2 = host.apply_rescaling(3, True)
    np.testing.assert_array_equal(r1, r2)
    np.testing.assert_array_equal(c1, c2)
    np.testing.assert_array_equal(dev.download()["values"], host.download()["values"])
    if k.nnz:
        assert np.max(np.abs(dev.matrix_vector_product(x) - (sp.diags(r1) @ k @ sp.diags(c1)) @ x)) < 1e-10
