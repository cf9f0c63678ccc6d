"""The SELL-32 images as the host builder lays them out, checked without a GPU:
``pdlp_b200_host_sell_layout`` exposes the structure and the test rebuilds the caller's matrix from
it, entry for entry -- every nonzero exactly once, with its exact value, padding zero, permutations
consistent, split rows cut into virtual slots of at most split_len entries. (On the GPU,
test_device_build.py checks that the device builder produces the same products as this layout.)"""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

from ortools_b200 import _capi as capi
from ortools_b200 import pdlp, synthetic
from test_device_build import cases, qp_from_matrix


class PdlpSellLayout(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("num_rows", "num_cols", "num_split", "num_virtual", "num_virtual_padded", "num_slots", "padded_nnz")] + [
        ("split_len", C.c_int32), ("slice_ptr", C.POINTER(C.c_int64)), ("slot_len", C.POINTER(C.c_int32)), ("col", C.POINTER(C.c_int32)),
        ("val", C.POINTER(C.c_double)), ("split_first", C.POINTER(C.c_int32)), ("virt_pos", C.POINTER(C.c_int32)),
        ("row_of_pos", C.POINTER(C.c_int32)), ("pos_of_row", C.POINTER(C.c_int32))]


def layouts(qp, row_begin=0, row_end=None, sigma=4096, natural=False):
    be = pdlp.backend()
    view, keep = qp._to_view()
    rows, cols = PdlpSellLayout(), PdlpSellLayout()
    row_end = view.num_constraints if row_end is None else row_end
    rc = be.fn("host_sell_layout")(C.byref(view), C.c_int64(row_begin), C.c_int64(row_end), C.c_int32(sigma), C.c_int32(int(natural)), C.byref(rows), C.byref(cols))
    be._check(rc, "host_sell_layout")

    def take(lay):
        arr = lambda p, n, dt: np.ctypeslib.as_array(p, shape=(n,)).astype(dt, copy=True) if n > 0 else np.zeros(0, dtype=dt)
        out = {f: getattr(lay, f) for f, t in PdlpSellLayout._fields_ if t in (C.c_int64, C.c_int32)}
        out["slice_ptr"] = arr(lay.slice_ptr, lay.num_slots // 32 + 1, np.int64)
        out["slot_len"] = arr(lay.slot_len, lay.num_slots, np.int64)
        out["col"] = arr(lay.col, lay.padded_nnz, np.int64)
        out["val"] = arr(lay.val, lay.padded_nnz, np.float64)
        out["split_first"] = arr(lay.split_first, lay.num_split + 1, np.int64)
        out["virt_pos"] = arr(lay.virt_pos, lay.num_virtual_padded, np.int64)
        out["row_of_pos"] = arr(lay.row_of_pos, lay.num_rows, np.int64)
        out["pos_of_row"] = arr(lay.pos_of_row, lay.num_rows, np.int64)
        be.fn("sell_layout_free", None)(C.byref(lay))
        return out
    del keep
    return take(rows), take(cols)


def rebuild(image, other, plain_col=False):
    """(row, col, value) triplets stored by one image, in the caller's indices; also checks its structure.
    plain_col: `col` already holds the caller's index (the row image of a sharded solve)."""
    L = image
    assert L["num_slots"] % 32 == 0 and L["num_virtual_padded"] % 32 == 0 and L["num_virtual"] <= L["num_virtual_padded"]
    assert L["slice_ptr"][0] == 0 and L["slice_ptr"][-1] == L["padded_nnz"] and np.all(np.diff(L["slice_ptr"]) >= 0)
    assert sorted(L["row_of_pos"]) == list(range(L["num_rows"]))                       # a permutation ...
    np.testing.assert_array_equal(L["pos_of_row"][L["row_of_pos"]], np.arange(L["num_rows"]))   # ... and its inverse
    used = np.zeros(L["padded_nnz"], dtype=bool)
    trip = []
    seen_virtual = {}
    for s in range(L["num_slots"]):
        if s < L["num_virtual_padded"]:
            p = L["virt_pos"][s]
            if p < 0:
                assert L["slot_len"][s] == 0
                continue
            assert 0 <= p < L["num_split"]
            assert L["split_first"][p] <= s < L["split_first"][p + 1]                # the virtual slots of a split row are contiguous
            assert 0 < L["slot_len"][s] <= L["split_len"]
            seen_virtual[p] = seen_virtual.get(p, 0) + 1
        else:
            p = L["num_split"] + (s - L["num_virtual_padded"])
            if p >= L["num_rows"]:
                assert L["slot_len"][s] == 0
                continue
            assert L["slot_len"][s] <= max(L["split_len"], 0) or L["split_len"] <= 0   # longer rows were split
        base, lane = L["slice_ptr"][s // 32], s % 32
        width = (L["slice_ptr"][s // 32 + 1] - base) // 32
        assert L["slot_len"][s] <= width                                             # a slice is as wide as its longest slot
        idx = base + 32 * np.arange(L["slot_len"][s]) + lane
        assert not used[idx].any()
        used[idx] = True
        r = L["row_of_pos"][p]
        for k in idx:
            trip.append((r, L["col"][k] if plain_col else other["row_of_pos"][L["col"][k]], L["val"][k]))
    assert np.all(L["val"][~used] == 0.0) and np.all(L["col"][~used] == 0)            # padding is inert
    for p in range(L["num_split"]):
        assert seen_virtual.get(p, 0) == L["split_first"][p + 1] - L["split_first"][p] >= 2     # a split row has at least two pieces
    return trip


def check(qp, row_begin=0, row_end=None, **kw):
    k = sp.csc_matrix(qp.constraint_matrix)
    k.sort_indices()
    m = k.shape[0]
    row_end = m if row_end is None else row_end
    rows, cols = layouts(qp, row_begin, row_end, **kw)
    assert rows["num_rows"] == row_end - row_begin and cols["num_rows"] == k.shape[1]
    block = k[row_begin:row_end].tocoo()
    want = sorted(zip(block.row.tolist(), block.col.tolist(), block.data.tolist()))
    got_rows = sorted((int(r), int(c), float(v)) for r, c, v in rebuild(rows, cols, plain_col=kw.get("natural", False)))
    got_cols = sorted((int(r), int(c), float(v)) for c, r, v in rebuild(cols, rows))   # the K^T image stores (col, row)
    assert got_rows == want
    assert got_cols == want
    return rows, cols


@pytest.mark.parametrize("name", ["ragged", "dense_and_empty", "empty_matrix", "c3", "c5"])
@pytest.mark.parametrize("split_len,sigma", [(0, 4096), (4, 32), (64, 4096)])
def test_images_store_the_matrix_entry_for_entry(name, split_len, sigma, monkeypatch):
    if split_len:
        monkeypatch.setenv("PDLP_B200_SPLIT_LEN", str(split_len))
    rows, cols = check(cases()[name], sigma=sigma)
    if split_len == 4 and name in ("dense_and_empty", "c3"):
        assert rows["num_split"] > 0                                      # long rows are cut into virtual slots
        assert (cols["num_split"] > 0) == (name == "dense_and_empty")     # every column of C3 has exactly 4 entries


def test_window_sorting_removes_padding_and_keeps_rows_in_their_window():
    qp = synthetic.c2(scale=0.005)[0]
    rows, cols = check(qp)
    for image in (rows, cols):
        nnz = int(image["slot_len"].sum())
        assert image["padded_nnz"] <= 1.05 * nnz + 32 * 32                # almost no slice padding
        regular = image["row_of_pos"][image["num_split"]:]
        window = np.arange(regular.size) // 4096
        # a row only moves inside its window of 4096 consecutive rows (gathers stay local)
        order_in = np.sort(regular.reshape(-1)[: (regular.size // 4096) * 4096].reshape(-1, 4096), axis=1)
        assert np.all(np.diff(order_in.min(axis=1)) > 0) and window.size == regular.size


@pytest.mark.parametrize("world", [2, 3])
def test_row_blocks_of_a_sharded_solve_tile_the_matrix(world):
    from ortools_b200 import distributed
    qp = synthetic.c3(scale=0.002)[0]
    m = qp.constraint_matrix.shape[0]
    covered = 0
    for rank in range(world):
        b, e = distributed.row_block(qp, rank, world)
        check(qp, b, e, natural=True)     # the row image gathers x~ in the caller's column order (plain indices)
        covered += e - b
    assert covered == m


def test_bad_row_range_is_a_status():
    qp = qp_from_matrix(sp.random(10, 12, density=0.3, random_state=1, format="csc"))
    view, keep = qp._to_view()
    a, b = PdlpSellLayout(), PdlpSellLayout()
    assert pdlp.backend().fn("host_sell_layout")(C.byref(view), C.c_int64(3), C.c_int64(99), C.c_int32(4096), C.c_int32(0), C.byref(a), C.byref(b)) == 3
