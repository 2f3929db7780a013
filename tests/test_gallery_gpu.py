"""GPU: device-side input builders and conversions are bit-identical to the
reference pipeline (gallery -> DIA -> cusp::convert), as restated by the oracle."""
import numpy as np
import pytest
import torch

import cusp_autotuned_b200 as cusp
from cusp_autotuned_b200 import capi, convert, gallery
from cusp_autotuned_b200.partition import plane_partition
from helpers import upload
from oracle import oracle as O

pytestmark = pytest.mark.gpu
DTYPES = [(np.float32, torch.float32), (np.float64, torch.float64)]


def _eq(t, a):
    return np.array_equal(t.cpu().numpy(), a)


@pytest.mark.parametrize("ndt,tdt", DTYPES)
@pytest.mark.parametrize("stencil,dims", [(5, (2, 3)), (5, (17, 9)), (5, (1, 7)), (7, (2, 2, 2)), (7, (9, 5, 4)),
                                          (7, (3, 1, 6)), (7, (16, 16, 16))])
def test_poisson_builders_bit_identical(stencil, dims, ndt, tdt, dev):
    ref = O.poisson(stencil, dims, ndt, "dia")
    A = gallery.poisson("dia", stencil, dims, dtype=tdt)
    assert A.num_entries == ref["num_entries"] and A.pitch == ref["pitch"]
    assert _eq(A.diagonal_offsets, ref["diagonal_offsets"]) and _eq(A.values, ref["values"])
    e = O.convert(ref, "ell")
    E = gallery.poisson("ell", stencil, dims, dtype=tdt)
    assert E.num_cols_per_row == e["num_cols_per_row"] and E.pitch == e["pitch"]
    assert _eq(E.column_indices, e["column_indices"]) and _eq(E.values, e["values"])
    c = O.convert(ref, "csr")
    Cm = gallery.poisson("csr", stencil, dims, dtype=tdt)
    assert _eq(Cm.row_offsets, c["row_offsets"]) and _eq(Cm.column_indices, c["column_indices"])
    assert _eq(Cm.values, c["values"])
    o = O.convert(ref, "coo")
    Om = gallery.poisson("coo", stencil, dims, dtype=tdt)
    assert _eq(Om.row_indices, o["row_indices"]) and _eq(Om.column_indices, o["column_indices"])
    assert _eq(Om.values, o["values"])
    assert capi.poisson_num_entries(stencil, *(dims if stencil == 7 else (*dims, 1)), 0, ref["num_rows"]) == ref["num_entries"]


@pytest.mark.parametrize("world", [2, 3, 4])
def test_partitioned_builders_reassemble_the_global_operator(world, dev):
    """row blocks + window-relative columns == the rows of the global matrix"""
    dims = (6, 5, 8)
    full = O.poisson(7, dims, np.float64, "csr")
    x = np.random.default_rng(1).uniform(-1, 1, full["num_cols"])
    want = O.spmv(full, x)
    for fmt in ("dia", "ell", "csr"):
        got = []
        for rank in range(world):
            blk = plane_partition(dims, world, rank)
            A = gallery.poisson(fmt, 7, dims, dtype=torch.float64, row_begin=blk.row_begin, num_rows=blk.num_rows,
                                halo_lo=blk.halo_lo, halo_hi=blk.halo_hi)
            assert A.num_rows == blk.num_rows and A.num_cols == blk.window
            xw = torch.from_numpy(x[blk.col_shift: blk.col_shift + blk.window].copy()).to(dev)
            y = torch.zeros(blk.num_rows, dtype=torch.float64, device=dev)
            # threads_per_row=1: the CSR kernel that keeps the reference's summation order
            cusp.multiply(A, xw, y, cfg=capi.Cfg(threads_per_row=1))
            got.append(y.cpu().numpy())
        assert np.array_equal(np.concatenate(got), want), fmt


@pytest.mark.parametrize("ndt,tdt", DTYPES)
def test_device_conversions_bit_identical(ndt, tdt, dev):
    """COO -> CSR -> ELL / HYB with the reference's rules"""
    rng = np.random.default_rng(5)
    coo = O.gallery_random(6000, 6000, 90000, ndt, "coo")
    coo["values"] = rng.uniform(0.5, 1.5, coo["num_entries"]).astype(ndt)
    Ad = upload("coo", coo, dev)
    csr_d = convert.coo_to_csr(Ad)
    csr = O.convert(coo, "csr")
    assert _eq(csr_d.row_offsets, csr["row_offsets"])
    K = convert.optimal_entries_per_row(csr_d.row_offsets)
    assert K == O.optimal_entries_per_row(csr["row_offsets"])
    hyb_d = convert.csr_to_hyb(csr_d)
    hyb = O.convert(csr, "hyb")
    assert hyb_d.ell.num_cols_per_row == hyb["ell"]["num_cols_per_row"] and hyb_d.ell.pitch == hyb["ell"]["pitch"]
    assert _eq(hyb_d.ell.column_indices, hyb["ell"]["column_indices"]) and _eq(hyb_d.ell.values, hyb["ell"]["values"])
    assert _eq(hyb_d.coo.row_indices, hyb["coo"]["row_indices"])
    assert _eq(hyb_d.coo.column_indices, hyb["coo"]["column_indices"]) and _eq(hyb_d.coo.values, hyb["coo"]["values"])
    ell_d = convert.csr_to_ell(csr_d)
    ell = O.convert(csr, "ell")
    assert _eq(ell_d.column_indices, ell["column_indices"]) and _eq(ell_d.values, ell["values"])
    assert ell_d.num_entries == ell["num_entries"]


def test_rmat_generator_properties(dev):
    A = convert.rmat(12, 16, seed=42, values="ones")
    r = A.row_indices.cpu().numpy().astype(np.int64)
    c = A.column_indices.cpu().numpy().astype(np.int64)
    key = r * A.num_cols + c
    assert A.num_rows == 4096 and np.all(np.diff(key) > 0)  # sorted by (row, col), no duplicates
    assert 0.5 * 16 * 4096 < A.num_entries <= 16 * 4096
    B = convert.rmat(12, 16, seed=42, values="ones")
    assert torch.equal(A.row_indices, B.row_indices) and torch.equal(A.column_indices, B.column_indices)
    deg = np.bincount(r, minlength=4096)
    assert deg.max() > 20 * deg.mean()  # power-law hub rows
    # all-ones values: y = A 1 are exact integer row degrees
    x = torch.ones(4096, dtype=torch.float32, device=dev)
    y = torch.zeros(4096, dtype=torch.float32, device=dev)
    cusp.multiply(A, x, y)
    assert np.array_equal(y.cpu().numpy(), deg.astype(np.float32))
