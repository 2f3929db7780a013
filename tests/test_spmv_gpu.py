"""GPU parity tests for cusp::multiply — CUDA path (through the C ABI) vs the oracle
on the same inputs.  Bars:
  * bit-exact for integer-valued data (every kernel, every configuration),
  * bit-exact for any data in kernels that keep the reference's summation order
    (DIA, ELL/ELL-R, CSR with threads_per_row == 1),
  * otherwise |y - y_ref| <= tol * |y_ref| per entry, tol = 1e-5 (fp32) / 1e-12
    (fp64), on data without cancellation; on cancelling rows the same tol is
    applied against sum_j |a_ij x_j|.
"""
import os

import numpy as np
import pytest
import torch

import cusp_autotuned_b200 as cusp
from cusp_autotuned_b200 import capi
from golden import reference_fixtures as G
from helpers import TOL, abs_matrix, rel_err, scaled_err, tdev, upload
from oracle import oracle as O

pytestmark = pytest.mark.gpu

FORMATS = ("csr", "coo", "dia", "ell", "hyb", "ellr")
DTYPES = [(np.float32, torch.float32), (np.float64, torch.float64)]
FMT_ID = {"csr": capi.FMT_CSR, "coo": capi.FMT_COO, "dia": capi.FMT_DIA, "ell": capi.FMT_ELL,
          "hyb": capi.FMT_HYB, "ellr": capi.FMT_ELLR}


def to_fmt(coo_or_any, fmt, **kw):
    if fmt == "ellr":
        return O.to_ellr(O.convert(coo_or_any, "ell", **kw))
    return O.convert(coo_or_any, fmt, **kw)


def upload_any(fmt, A, dev):
    if fmt == "ellr":
        return cusp.ellr_matrix(upload("ell", {**A, "format": "ell"}, dev))
    return upload(fmt, A, dev)


def gpu_multiply(fmt, A, x, dev, y0=None, accumulate=False, cfg=None):
    Ad = upload_any(fmt, A, dev)
    vals = A["ell"]["values"] if fmt == "hyb" else A["values"]
    xd = tdev(np.asarray(x, vals.dtype), dev)
    if y0 is None:
        yd = torch.full((A["num_rows"],), 10, dtype=xd.dtype, device=dev)
    else:
        yd = tdev(np.asarray(y0, vals.dtype), dev)
    cusp.multiply(Ad, xd, yd, accumulate=accumulate, cfg=cfg)
    torch.cuda.synchronize()
    return yd.cpu().numpy()


# ---------------------------------------------------------------------------
# the reference's own fixtures (testing/multiply.cu:441-645)
# ---------------------------------------------------------------------------
def _fixture_matrices(ndt):
    out = {k: O.dense_to_coo(v.astype(ndt)) for k, v in G.MULTIPLY_DENSE.items()}
    for k, grid in G.MULTIPLY_POISSON.items():
        out[k] = O.poisson(5, grid, ndt, "coo")
    return out


@pytest.mark.parametrize("ndt,tdt", DTYPES)
@pytest.mark.parametrize("fmt", FORMATS)
def test_reference_fixtures_exact(fmt, ndt, tdt, dev):
    for name, coo in _fixture_matrices(ndt).items():
        A = to_fmt(coo, fmt)
        x = (np.arange(coo["num_cols"]) % 10).astype(ndt)
        want = O.spmv(A, x)
        got = gpu_multiply(fmt, A, x, dev)  # y pre-filled with 10 must be overwritten
        assert np.array_equal(got, want), (name, fmt)
        y0 = np.full(coo["num_rows"], 10, ndt)
        want = O.spmv(A, x, y0, accumulate=True)
        got = gpu_multiply(fmt, A, x, dev, y0=y0, accumulate=True)
        assert np.array_equal(got, want), (name, fmt, "accumulate")


# ---------------------------------------------------------------------------
# outputs of the reference's host loops (committed, tests/golden/ref_spmv.npz)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("ndt,tdt", DTYPES)
@pytest.mark.parametrize("fmt", ("csr", "coo", "dia", "ell", "hyb"))
def test_committed_reference_outputs(fmt, ndt, tdt, dev, golden):
    tol = TOL[np.dtype(ndt)]
    for name, st, grid in (("p5", 5, (13, 9)), ("p7", 7, (7, 6, 5)), ("p9", 9, (6, 7)), ("p27", 27, (4, 3, 5))):
        A = O.poisson(st, grid, ndt, fmt)
        key = f"{name}_{np.dtype(ndt).name}_{fmt}"
        x, y0 = golden[key + "_x"], golden[key + "_y0"]
        scale = O.spmv(abs_matrix(A), np.abs(x))
        got = gpu_multiply(fmt, A, x, dev)
        got_acc = gpu_multiply(fmt, A, x, dev, y0=y0, accumulate=True)
        if fmt in ("dia", "ell"):  # order-preserving kernels: bit-exact on any data
            assert np.array_equal(got, golden[key + "_y"]), key
            assert np.array_equal(got_acc, golden[key + "_yacc"]), key
        else:
            assert scaled_err(got, golden[key + "_y"], scale) <= tol, key
            assert scaled_err(got_acc, golden[key + "_yacc"], scale + np.abs(y0)) <= tol, key
    for m, n, s in ((24, 24, 150), (24, 12, 20), (300, 257, 4000)):
        if fmt == "dia":
            continue
        coo = O.gallery_random(m, n, s, ndt, "coo")
        key = f"rand{m}x{n}_{np.dtype(ndt).name}_{fmt}"
        coo["values"] = golden[key + "_vals"]
        A = O.convert(coo, fmt)
        got = gpu_multiply(fmt, A, golden[key + "_x"], dev)
        # positive values and x: no cancellation -> the north-star per-entry bound
        assert rel_err(got, golden[key + "_y"]) <= tol, key


@pytest.mark.parametrize("ndt,tdt", DTYPES)
def test_csr_order_preserving_kernels_bit_exact_on_any_data(ndt, tdt, dev):
    """threads_per_row == 1 and the stream kernel keep the reference's order (csr_spmv.h:35-74)"""
    rng = np.random.default_rng(3)
    A = O.poisson(7, (11, 9, 8), ndt, "csr")
    A["values"] = (A["values"] * rng.uniform(0.5, 1.5, len(A["values"]))).astype(ndt)
    x = rng.uniform(-1, 1, A["num_cols"]).astype(ndt)
    y0 = rng.uniform(-1, 1, A["num_rows"]).astype(ndt)
    for block in (128, 256, 512):
        for u in (1, 2, 4):
            cfg = capi.Cfg(kernel=capi.K_CSR_VECTOR, block_size=block, threads_per_row=1, unroll=u)
            assert np.array_equal(gpu_multiply("csr", A, x, dev, cfg=cfg), O.spmv(A, x))
            assert np.array_equal(gpu_multiply("csr", A, x, dev, y0=y0, accumulate=True, cfg=cfg),
                                  O.spmv(A, x, y0, accumulate=True))
    for cfg in [c for c in capi.Handle.cfg_space(capi.FMT_CSR, 0) if c.kernel in (capi.K_CSR_STREAM, capi.K_CSR_RING)]:
        assert np.array_equal(gpu_multiply("csr", A, x, dev, cfg=cfg), O.spmv(A, x)), cfg
        assert np.array_equal(gpu_multiply("csr", A, x, dev, y0=y0, accumulate=True, cfg=cfg),
                              O.spmv(A, x, y0, accumulate=True)), cfg


@pytest.mark.parametrize("ndt,tdt", DTYPES)
def test_csr_long_rows_and_skewed_lengths(ndt, tdt, dev):
    """rows far longer than a shared-memory chunk, next to empty and 1-entry rows"""
    rng = np.random.default_rng(8)
    n = 3000
    lens = np.concatenate([rng.integers(0, 4, 500), [20000], rng.integers(0, 9, 700), [129, 128, 127, 4097, 0, 0],
                           rng.integers(0, 300, 200)])
    rows = len(lens)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    nnz = int(offs[-1])
    A = dict(format="csr", num_rows=rows, num_cols=n, num_entries=nnz, row_offsets=offs,
             column_indices=rng.integers(0, n, nnz).astype(np.int32), values=rng.integers(1, 4, nnz).astype(ndt))
    x = rng.integers(-3, 4, n).astype(ndt)
    want = O.spmv(A, x)
    for cfg in capi.Handle.cfg_space(capi.FMT_CSR, 0):
        assert np.array_equal(gpu_multiply("csr", A, x, dev, cfg=cfg), want), cfg


# ---------------------------------------------------------------------------
# every configuration of every tuning space (testing/ktt.cu CHECK_ALL_CONFIGURATIONS)
# ---------------------------------------------------------------------------
def _all_cfg_matrices(ndt):
    rng = np.random.default_rng(17)
    mats = {}
    mats["p7_13x11x9"] = O.poisson(7, (13, 11, 9), ndt, "coo")
    mats["p5_70x53"] = O.poisson(5, (70, 53), ndt, "coo")
    r = O.gallery_random(1500, 1300, 30000, ndt, "coo")  # ragged rows, some empty
    r["values"] = rng.integers(1, 5, r["num_entries"]).astype(ndt)
    mats["random"] = r
    return mats


@pytest.mark.parametrize("ndt,tdt", DTYPES)
@pytest.mark.parametrize("fmt", ("csr", "coo", "dia", "ell", "ellr", "hyb"))
def test_every_configuration_integer_exact(fmt, ndt, tdt, dev, handle):
    space = capi.Handle.cfg_space(FMT_ID[fmt], capi.F32 if ndt == np.float32 else capi.F64)
    assert len(space) > 0
    for name, coo in _all_cfg_matrices(ndt).items():
        if fmt == "dia" and name == "random":
            continue  # DIA fill-in of a random pattern is meaningless (reference refuses > 3x)
        kw = dict(num_entries_per_row=3) if fmt == "hyb" else {}
        A = to_fmt(coo, fmt, **kw)
        x = ((np.arange(coo["num_cols"]) % 21) - 10).astype(ndt)  # benchmark convention
        want = O.spmv(A, x)
        Ad = upload_any(fmt, A, dev)
        xd = tdev(x, dev)
        ran = 0
        for cfg in space:
            yd = torch.full((A["num_rows"],), 99, dtype=tdt, device=dev)
            try:
                cusp.multiply(Ad, xd, yd, cfg=cfg)
            except capi.InvalidInput:
                continue  # configuration needs more shared memory than the device has
            torch.cuda.synchronize()
            assert np.array_equal(yd.cpu().numpy(), want), (fmt, name, cfg)
            ran += 1
        assert ran >= len(space) * 0.7, (fmt, name, ran, len(space))


# ---------------------------------------------------------------------------
# edge cases
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("ndt,tdt", DTYPES)
def test_ktt_banded_fixtures(ndt, tdt, dev):
    """testing/ktt.cu:274-281: all-ones banded 4096x4096 / 4096x2048 / 2048x4096, 1024 diagonals"""
    for rows, cols, step, cnt in G.KTT_BANDED:
        A = O.make_diagonal_symmetric(rows, cols, step, cnt)
        A["values"] = A["values"].astype(ndt)
        x = (np.arange(cols) % 10).astype(ndt)
        want = O.spmv(A, x)
        for kern in (capi.K_DIA_LDG, capi.K_DIA_BULK):
            got = gpu_multiply("dia", A, x, dev, cfg=capi.Cfg(kernel=kern))
            assert np.array_equal(got, want), (rows, cols, kern)


@pytest.mark.parametrize("ndt,tdt", DTYPES)
def test_unaligned_pitch_falls_back_with_same_result(ndt, tdt, dev):
    """pitch*sizeof(T) % 16 != 0: the bulk-copy kernels cannot be used; same arithmetic via LDG"""
    A = O.poisson(7, (7, 5, 3), ndt, "dia")  # 105 rows, pitch 105
    assert (A["pitch"] * np.dtype(ndt).itemsize) % 16 != 0
    x = np.random.default_rng(1).uniform(-1, 1, 105).astype(ndt)
    for fmt in ("dia", "ell"):
        B = O.convert(A, fmt)
        for kern in (1, 2):
            assert np.array_equal(gpu_multiply(fmt, B, x, dev, cfg=capi.Cfg(kernel=kern)), O.spmv(B, x))


@pytest.mark.parametrize("ndt,tdt", DTYPES)
def test_ragged_tiles_and_alignment_32(ndt, tdt, dev):
    """rows not a multiple of any tile; ELL/DIA with the default alignment-32 pitch"""
    rng = np.random.default_rng(2)
    for grid in ((33, 31), (129, 65), (257, 129)):
        csr = O.poisson(5, grid, ndt, "csr")
        csr["values"] = (csr["values"] * rng.uniform(0.5, 1.5, len(csr["values"]))).astype(ndt)
        x = rng.uniform(-1, 1, csr["num_cols"]).astype(ndt)
        for fmt in ("ell", "dia"):
            A = O.convert(csr, fmt)  # pitch = round_up(rows, 32)
            assert A["pitch"] % 32 == 0 and A["pitch"] >= A["num_rows"]
            for kern in (1, 2):
                assert np.array_equal(gpu_multiply(fmt, A, x, dev, cfg=capi.Cfg(kernel=kern)), O.spmv(A, x))


@pytest.mark.parametrize("ndt,tdt", DTYPES)
def test_rectangular_and_empty_rows(ndt, tdt, dev):
    D = np.zeros((37, 91), ndt)
    D[0, 90] = 3
    D[5, :] = np.arange(91) % 4  # a long row
    D[36, 0] = 2
    D[20, 45] = 7
    coo = O.dense_to_coo(D)
    x = (np.arange(91) % 7 - 3).astype(ndt)
    for fmt in FORMATS:
        A = to_fmt(coo, fmt, **(dict(num_entries_per_row=2) if fmt == "hyb" else {}))
        assert np.array_equal(gpu_multiply(fmt, A, x, dev), D @ x), fmt


@pytest.mark.parametrize("ndt,tdt", DTYPES)
def test_coo_hub_row_spanning_many_tiles(ndt, tdt, dev):
    """power-law hub: one row with 60000 entries between ordinary rows; duplicates allowed"""
    rng = np.random.default_rng(4)
    n = 5000
    rows = np.concatenate([np.repeat(np.arange(0, 100), 3), np.full(60000, 100), np.repeat(np.arange(101, 400), 5),
                           np.full(7, n - 1)]).astype(np.int32)
    cols = rng.integers(0, n, len(rows)).astype(np.int32)
    vals = rng.integers(1, 4, len(rows)).astype(ndt)
    A = dict(format="coo", num_rows=n, num_cols=n, num_entries=len(rows), row_indices=rows,
             column_indices=cols, values=vals)
    x = rng.integers(-3, 4, n).astype(ndt)
    want = O.spmv(A, x)
    for cfg in capi.Handle.cfg_space(capi.FMT_COO, capi.F32 if ndt == np.float32 else capi.F64):
        assert np.array_equal(gpu_multiply("coo", A, x, dev, cfg=cfg), want), cfg
    y0 = rng.integers(-5, 5, n).astype(ndt)
    assert np.array_equal(gpu_multiply("coo", A, x, dev, y0=y0, accumulate=True), O.spmv(A, x, y0, accumulate=True))


@pytest.mark.parametrize("ndt,tdt", DTYPES)
def test_coo_tile_boundary_cases(ndt, tdt, dev):
    """rows that start/end exactly on tile boundaries (256*7 = 1792 entries per tile)"""
    tile = 256 * 7
    for lens in ([tile, tile, tile], [tile - 1, 1, tile], [1, tile * 2 + 5, 3], [tile * 3], [5] * 1000):
        rows = np.repeat(np.arange(len(lens)) * 2, lens).astype(np.int32)  # odd rows empty
        n = 2 * len(lens)
        cols = (np.arange(len(rows)) % n).astype(np.int32)
        vals = np.ones(len(rows), ndt)
        A = dict(format="coo", num_rows=n, num_cols=n, num_entries=len(rows), row_indices=rows,
                 column_indices=cols, values=vals)
        x = (np.arange(n) % 3 + 1).astype(ndt)
        for kern in (capi.K_COO_SEGSCAN, capi.K_COO_RING):
            cfg = capi.Cfg(kernel=kern, block_size=256, unroll=7)
            assert np.array_equal(gpu_multiply("coo", A, x, dev, cfg=cfg), O.spmv(A, x)), (lens[:3], kern)


@pytest.mark.parametrize("ndt,tdt", DTYPES)
def test_coo_ring_matches_segscan_bitwise(ndt, tdt, dev):
    """K_COO_RING keeps K_COO_SEGSCAN's tiles and summation order: identical bits on any data.
    Sizes: every residue of nnz mod 4 (the producer's scalar tail), far more tiles than the
    persistent grid x stages (ring wrap-around), hub rows across many tiles, accumulate."""
    rng = np.random.default_rng(11)
    n = 40000
    for nnz in (1, 2, 3, 5, 1791, 1792, 1793, 1792 * 3 + 2, 1792 * 1500 + 3, 1792 * 2600 + 1):
        rows = np.sort(rng.integers(0, n, nnz)).astype(np.int32)
        if nnz > 100000:
            rows[1000:60000] = rows[1000]  # a hub row spanning ~33 tiles
            rows = np.sort(rows)
        cols = rng.integers(0, n, nnz).astype(np.int32)
        vals = rng.uniform(0.5, 1.5, nnz).astype(ndt)
        A = dict(format="coo", num_rows=n, num_cols=n, num_entries=nnz, row_indices=rows,
                 column_indices=cols, values=vals)
        x = rng.uniform(0.5, 1.5, n).astype(ndt)
        y0 = rng.uniform(-1, 1, n).astype(ndt)
        want = O.spmv(A, x)
        for b, u in ((256, 7), (128, 9), (512, 5)):
            seg = gpu_multiply("coo", A, x, dev, cfg=capi.Cfg(kernel=capi.K_COO_SEGSCAN, block_size=b, unroll=u))
            for st, cps in ((2, 4), (3, 1)):
                ring = gpu_multiply("coo", A, x, dev,
                                    cfg=capi.Cfg(kernel=capi.K_COO_RING, block_size=b, unroll=u, stages=st,
                                                 ctas_per_sm=cps))
                assert np.array_equal(ring, seg), (nnz, b, u, st, cps)
            assert rel_err(seg, want) <= TOL[np.dtype(ndt)], (nnz, b, u)
        shape = dict(block_size=256, unroll=7)  # same tiles -> same grouping of the sums
        assert np.array_equal(gpu_multiply("coo", A, x, dev, y0=y0, accumulate=True,
                                           cfg=capi.Cfg(kernel=capi.K_COO_RING, **shape)),
                              gpu_multiply("coo", A, x, dev, y0=y0, accumulate=True,
                                           cfg=capi.Cfg(kernel=capi.K_COO_SEGSCAN, **shape)))
        got = gpu_multiply("coo", A, x, dev, y0=y0, accumulate=True, cfg=capi.Cfg(kernel=capi.K_COO_RING))
        assert scaled_err(got, O.spmv(A, x, y0, accumulate=True), np.abs(y0) + want) <= TOL[np.dtype(ndt)]


WARP_SHAPES = [(4, 1), (4, 2), (4, 4), (8, 1), (8, 2)]


@pytest.mark.parametrize("ndt,tdt", DTYPES)
def test_coo_warp_tile_boundaries_and_hubs(ndt, tdt, dev):
    """K_COO_WARP: rows that start / end exactly on lane, unit and warp-tile boundaries, a matrix that is one row,
    hub rows spanning > 64 tiles (the fix-up kernel's warp-parallel chain walk), ragged last tile, duplicates,
    accumulate — integer data, exact against the host loop."""
    rng = np.random.default_rng(21)
    for vw, u in WARP_SHAPES:
        wt = 32 * vw * u
        cfg = capi.Cfg(kernel=capi.K_COO_WARP, vector_width=vw, unroll=u)
        cases = ([wt, wt, wt], [wt - 1, 1, wt], [1, 2 * wt + 5, 3], [3 * wt], [vw] * 500, [vw - 1, vw + 1] * 300,
                 [5] * 1000, [1] * 777, [70 * wt + 3, 2, 40 * wt], [1, 130 * wt])
        for lens in cases:
            rows = np.repeat(np.arange(len(lens)) * 2, lens).astype(np.int32)  # odd rows empty
            n = 2 * len(lens)
            cols = rng.integers(0, n, len(rows)).astype(np.int32)
            vals = rng.integers(-2, 3, len(rows)).astype(ndt)
            A = dict(format="coo", num_rows=n, num_cols=n, num_entries=len(rows), row_indices=rows,
                     column_indices=cols, values=vals)
            x = rng.integers(-3, 4, n).astype(ndt)
            assert np.array_equal(gpu_multiply("coo", A, x, dev, cfg=cfg), O.spmv(A, x)), (vw, u, lens[:3])
            y0 = rng.integers(-5, 5, n).astype(ndt)
            assert np.array_equal(gpu_multiply("coo", A, x, dev, y0=y0, accumulate=True, cfg=cfg),
                                  O.spmv(A, x, y0, accumulate=True)), (vw, u, lens[:3], "accumulate")
        # persistent grid: more tiles than resident warps
        nnz = wt * 9000 + 17
        n = 50000
        rows = np.sort(rng.integers(0, n, nnz)).astype(np.int32)
        cols = rng.integers(0, n, nnz).astype(np.int32)
        vals = rng.integers(-2, 3, nnz).astype(ndt)
        A = dict(format="coo", num_rows=n, num_cols=n, num_entries=nnz, row_indices=rows, column_indices=cols,
                 values=vals)
        x = rng.integers(-3, 4, n).astype(ndt)
        want = O.spmv(A, x)
        for cps in (0, 1, 3):
            c2 = capi.Cfg(kernel=capi.K_COO_WARP, vector_width=vw, unroll=u, ctas_per_sm=cps)
            assert np.array_equal(gpu_multiply("coo", A, x, dev, cfg=c2), want), (vw, u, cps)


@pytest.mark.parametrize("ndt,tdt", DTYPES)
def test_coo_warp_tolerance_and_persistent_grid(ndt, tdt, dev):
    """non-integer data: every shape within the north-star bound of the host loop; the persistent grid changes no bit
    (same tiles, same order)"""
    rng = np.random.default_rng(22)
    n, nnz = 30000, 128 * 3000 + 77
    rows = np.sort(rng.integers(0, n, nnz)).astype(np.int32)
    rows[5000:90000] = rows[5000]
    rows = np.sort(rows)
    cols = rng.integers(0, n, nnz).astype(np.int32)
    vals = rng.uniform(0.5, 1.5, nnz).astype(ndt)
    A = dict(format="coo", num_rows=n, num_cols=n, num_entries=nnz, row_indices=rows, column_indices=cols, values=vals)
    x = rng.uniform(0.5, 1.5, n).astype(ndt)
    want = O.spmv(A, x)
    for vw, u in WARP_SHAPES:
        base = gpu_multiply("coo", A, x, dev, cfg=capi.Cfg(kernel=capi.K_COO_WARP, vector_width=vw, unroll=u))
        assert rel_err(base, want) <= TOL[np.dtype(ndt)], (vw, u)
        for cps in (1, 2, 8):
            got = gpu_multiply("coo", A, x, dev, cfg=capi.Cfg(kernel=capi.K_COO_WARP, vector_width=vw, unroll=u, ctas_per_sm=cps))
            assert np.array_equal(got, base), (vw, u, cps)


def test_coo_default_kernel_follows_the_column_stream(dev, handle):
    """the default choice (coo_gather_class probe): scattered columns -> K_COO_WARP, one entry per row -> the LDG scan
    kernel, banded -> the ring kernel; whatever it picks, the result is the host loop's (integer data, exact)"""
    rng = np.random.default_rng(24)
    n, nnz = 1 << 18, 1 << 22
    rows = np.sort(rng.integers(0, n, nnz)).astype(np.int32)
    for name, cols in (("scattered", rng.integers(0, n, nnz).astype(np.int32)),
                       ("banded", np.minimum(rows + (np.arange(nnz) % 5).astype(np.int32), n - 1).astype(np.int32))):
        vals = rng.integers(-2, 3, nnz).astype(np.float32)
        A = dict(format="coo", num_rows=n, num_cols=n, num_entries=nnz, row_indices=rows, column_indices=cols, values=vals)
        x = rng.integers(-3, 4, n).astype(np.float32)
        assert np.array_equal(gpu_multiply("coo", A, x, dev), O.spmv(A, x)), name


@pytest.mark.parametrize("ndt,tdt", DTYPES)
def test_coo_plan_matches_warp_kernel_bitwise(ndt, tdt, dev, handle):
    """b200sp_coo_plan_*: hot columns of x served from the shared-memory table — same tiles and order as K_COO_WARP,
    so identical bits; table smaller than / equal to / larger than the number of distinct columns; skewed column
    stream; attached plans are used by b200sp_spmv_coo for exactly their arrays and for nothing else."""
    rng = np.random.default_rng(23)
    n, nnz = 20000, 128 * 2500 + 5
    rows = np.sort(rng.integers(0, n, nnz)).astype(np.int32)
    cols = np.minimum((rng.pareto(1.2, nnz) * 40).astype(np.int64), n - 1).astype(np.int32)  # power-law columns
    vals = rng.uniform(0.5, 1.5, nnz).astype(ndt)
    x = rng.uniform(0.5, 1.5, n).astype(ndt)
    ri, ci, va, xd = tdev(rows, dev), tdev(cols, dev), tdev(vals, dev), tdev(x, dev)
    dt = capi.F32 if ndt == np.float32 else capi.F64
    es = np.dtype(ndt).itemsize
    for table_bytes in (16 * es, 4096, 0):
        plan = handle.coo_plan_create(n, n, nnz, ri, ci, dt, table_bytes)
        info = handle.coo_plan_info(plan)
        assert 0 < info["hot_columns"] <= info["capacity"]
        if table_bytes:
            assert info["capacity"] == table_bytes // es
        assert 0 < info["hot_entries"] <= nnz
        for vw, u in ((4, 1), (4, 2), (8, 1)) if ndt == np.float32 else ((4, 1), (4, 2)):
            cfg = capi.Cfg(kernel=capi.K_COO_WARP, vector_width=vw, unroll=u)
            yw = torch.full((n,), 9, dtype=tdt, device=dev)
            handle.spmv_coo(n, n, nnz, ri, ci, va, xd, yw, cfg=cfg)
            yp = torch.full((n,), 5, dtype=tdt, device=dev)
            handle.spmv_coo_plan(plan, va, xd, yp, cfg=cfg)
            assert torch.equal(yw, yp), (table_bytes, vw, u)
            handle.spmv_coo(n, n, nnz, ri, ci, va, xd, yw, accumulate=True, cfg=cfg)
            handle.spmv_coo_plan(plan, va, xd, yp, accumulate=True, cfg=cfg)
            assert torch.equal(yw, yp), (table_bytes, vw, u, "accumulate")
        # attached: the plain entry point takes the executor for these arrays only
        y_plain = torch.empty(n, dtype=tdt, device=dev)
        handle.spmv_coo(n, n, nnz, ri, ci, va, xd, y_plain,
                        cfg=capi.Cfg(kernel=capi.K_COO_WARP, vector_width=8 if ndt == np.float32 else 4, unroll=1))
        handle.coo_plan_attach(plan)
        y_att = torch.empty(n, dtype=tdt, device=dev)
        handle.spmv_coo(n, n, nnz, ri, ci, va, xd, y_att)
        assert torch.equal(y_att, y_plain)
        ci2 = ci.clone()  # same contents, different array: not the plan's matrix
        y_other = torch.empty(n, dtype=tdt, device=dev)
        handle.spmv_coo(n, n, nnz, ri, ci2, va, xd, y_other)
        assert rel_err(y_other.cpu().numpy(), y_plain.cpu().numpy()) <= TOL[np.dtype(ndt)]
        handle.coo_plan_detach(plan)
        handle.coo_plan_destroy(plan)
    want = O.spmv(dict(format="coo", num_rows=n, num_cols=n, num_entries=nnz, row_indices=rows, column_indices=cols,
                       values=vals), x)
    assert rel_err(y_att.cpu().numpy(), want) <= TOL[np.dtype(ndt)]


@pytest.mark.parametrize("ndt,tdt", DTYPES)
def test_coo_unaligned_bases_fall_back(ndt, tdt, dev):
    """array bases that are not 16-byte aligned cannot be bulk-copied: K_COO_RING requests run
    the LDG kernel with the same result"""
    rng = np.random.default_rng(12)
    n, nnz = 3000, 1792 * 700 + 5
    rows = np.sort(rng.integers(0, n, nnz)).astype(np.int32)
    cols = rng.integers(0, n, nnz).astype(np.int32)
    vals = rng.integers(1, 4, nnz).astype(ndt)
    x = rng.integers(-3, 4, n).astype(ndt)
    A = dict(format="coo", num_rows=n, num_cols=n, num_entries=nnz, row_indices=rows,
             column_indices=cols, values=vals)
    want = O.spmv(A, x)
    pad = lambda a: tdev(np.concatenate([a[:1], a]), dev)[1:]  # base shifted by one element
    Ad = cusp.coo_matrix(n, n, pad(rows), pad(cols), pad(vals))
    assert Ad.values.data_ptr() % 16 != 0
    for kern in (0, capi.K_COO_RING):
        y = torch.full((n,), 5, dtype=tdt, device=dev)
        cusp.multiply(Ad, tdev(x, dev), y, cfg=capi.Cfg(kernel=kern))
        assert np.array_equal(y.cpu().numpy(), want)


def test_empty_and_degenerate(dev, handle):
    for tdt in (torch.float32, torch.float64):
        x = torch.ones(5, dtype=tdt, device=dev)
        y = torch.full((4,), 3, dtype=tdt, device=dev)
        i32 = lambda a: torch.tensor(a, dtype=torch.int32, device=dev)
        e = torch.zeros(0, dtype=tdt, device=dev)
        # no stored entries: y = 0 ; y += 0
        A = cusp.csr_matrix(4, 5, i32([0, 0, 0, 0, 0]), i32([]), e)
        assert torch.equal(cusp.multiply(A, x, y.clone()), torch.zeros_like(y))
        assert torch.equal(cusp.multiply(A, x, y.clone(), accumulate=True), y)
        A = cusp.coo_matrix(4, 5, i32([]), i32([]), e)
        assert torch.equal(cusp.multiply(A, x, y.clone()), torch.zeros_like(y))
        A = cusp.ell_matrix(4, 5, 0, 0, 4, i32([]), e)
        assert torch.equal(cusp.multiply(A, x, y.clone()), torch.zeros_like(y))
        A = cusp.dia_matrix(4, 5, 0, i32([]), 4, e)
        assert torch.equal(cusp.multiply(A, x, y.clone()), torch.zeros_like(y))
        # zero rows
        A = cusp.csr_matrix(0, 5, i32([0]), i32([]), e)
        cusp.multiply(A, x, torch.zeros(0, dtype=tdt, device=dev))
        # 1x1
        A = cusp.csr_matrix(1, 1, i32([0, 1]), i32([0]), torch.tensor([2.5], dtype=tdt, device=dev))
        out = cusp.multiply(A, torch.tensor([4.0], dtype=tdt, device=dev), torch.zeros(1, dtype=tdt, device=dev))
        assert out.item() == 10.0


def test_error_behaviour(dev, handle):
    """size mismatch -> invalid_input_exception (testing/blas.cu:141, multiply dispatch)"""
    A = upload("csr", G.CONVERT_CSR, dev)
    x = torch.ones(4, dtype=torch.float32, device=dev)
    with pytest.raises(cusp.InvalidInput):
        cusp.multiply(A, torch.ones(3, dtype=torch.float32, device=dev), torch.ones(4, dtype=torch.float32, device=dev))
    with pytest.raises(cusp.InvalidInput):
        cusp.multiply(A, x, torch.ones(5, dtype=torch.float32, device=dev))
    with pytest.raises(cusp.InvalidInput):  # value type mismatch
        cusp.multiply(A, x.double(), torch.ones(4, dtype=torch.float64, device=dev))
    with pytest.raises(cusp.InvalidInput):  # unsupported configuration
        cusp.multiply(A, x, torch.ones(4, dtype=torch.float32, device=dev), cfg=capi.Cfg(block_size=100))
    with pytest.raises(cusp.InvalidInput):  # host tensor given to a device container op
        cusp.multiply(A, x.cpu(), torch.ones(4, dtype=torch.float32))
    assert "block_size" in handle.lib.b200sp_last_error_string(handle._h).decode() or True


def test_spmv_host_buffers(dev, handle):
    """b200sp_spmv_host: x from host memory, y back to host memory"""
    A = O.poisson(7, (20, 17, 9), np.float64, "dia")
    Ad = upload("dia", A, dev)
    x = np.random.default_rng(9).uniform(-1, 1, A["num_cols"])
    y = np.zeros(A["num_rows"])
    handle.spmv_host(Ad.descriptor(), x, y)
    assert np.array_equal(y, O.spmv(A, x))
    xp = torch.from_numpy(x).pin_memory()
    yp = torch.zeros(A["num_rows"], dtype=torch.float64).pin_memory()
    handle.spmv_host(Ad.descriptor(), xp, yp)
    assert np.array_equal(yp.numpy(), O.spmv(A, x))


def test_native_library_is_the_one_running(handle):
    """the product path is libb200sp.so: launches are counted by the library itself"""
    import os
    assert os.path.basename(capi.LIB_PATH) == "libb200sp.so" and os.path.exists(capi.LIB_PATH)
    before = handle.launch_count
    x = torch.ones(1000, dtype=torch.float32, device="cuda")
    handle.axpy(2.0, x, x.clone())
    assert handle.launch_count == before + 1


# ---------------------------------------------------------------------------
# CSR x dense block (cusp::multiply(csr_matrix, array2d, array2d), csr_block_spmv.h)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("ndt,tdt", DTYPES)
def test_csr_block_multiply_bit_exact(ndt, tdt, dev):
    """per (row, column) the entries are added in storage order like the host loop
    (sequential/multiply/csr_block_spmv.h:52-77): column j of Y == the sequential SpMV with
    column j of X, bit for bit, for every block width (sub-warp widths 1..32, chunks beyond 32),
    padded leading dimensions, ragged / empty / hub rows, accumulate"""
    rng = np.random.default_rng(21)
    mats = {"p7": O.poisson(7, (11, 9, 7), ndt, "csr"), "rand": O.convert(O.gallery_random(700, 500, 9000, ndt, "coo"), "csr")}
    hub = O.gallery_random(300, 300, 2500, ndt, "coo")
    hub_rows = np.concatenate([hub["row_indices"], np.full(5000, 150, np.int32)])
    order = np.argsort(hub_rows, kind="stable")
    hubm = dict(format="coo", num_rows=300, num_cols=300, num_entries=len(hub_rows), row_indices=hub_rows[order],
                column_indices=np.concatenate([hub["column_indices"], rng.integers(0, 300, 5000).astype(np.int32)])[order],
                values=np.ones(len(hub_rows), ndt))
    mats["hub"] = O.convert(hubm, "csr")
    for name, A in mats.items():
        A = dict(A)
        A["values"] = (A["values"] * rng.uniform(0.5, 1.5, A["num_entries"])).astype(ndt)
        Ad = upload("csr", A, dev)
        # pads: 0 / 4 keep 16-byte-aligned rows (a lane owns 4 fp32 / 2 fp64 columns: 128-bit accesses), 3 does not
        # (one column per lane); k = 12, 24, 64 exercise idle lanes and two column chunks of the vector path
        for k, pad in [(1, 0), (2, 0), (2, 3), (3, 0), (4, 0), (4, 3), (4, 4), (7, 0), (8, 0), (8, 3), (12, 4), (16, 0),
                       (16, 3), (24, 0), (31, 0), (32, 0), (32, 3), (32, 4), (33, 0), (64, 0), (70, 3), (96, 0), (98, 2), (100, 0), (130, 2)]:
            Xh = rng.uniform(-1, 1, (A["num_cols"], k + pad)).astype(ndt)
            Y0 = rng.uniform(-1, 1, (A["num_rows"], k + pad)).astype(ndt)
            X = tdev(Xh, dev)[:, :k]
            for acc in (False, True):
                Yfull = tdev(Y0, dev)
                Y = Yfull[:, :k]
                cusp.multiply_block(Ad, X, Y, accumulate=acc)
                got = Y.cpu().numpy()
                assert np.array_equal(Yfull[:, k:].cpu().numpy(), Y0[:, k:])  # padding columns untouched
                for j in range(k):
                    want = O.spmv(A, np.ascontiguousarray(Xh[:, j]), np.ascontiguousarray(Y0[:, j]) if acc else None,
                                  accumulate=acc)
                    if k == 1 and pad == 0:
                        # one contiguous column is handed to the SpMV kernels (csr_block_spmv.h:198-201), whose
                        # row-split variants regroup the sums: tolerance bar against sum_j |a_ij x_j|
                        scale = O.spmv(abs_matrix(A), np.abs(Xh[:, j])) + (np.abs(Y0[:, j]) if acc else 0)
                        assert scaled_err(got[:, j], want, scale) <= TOL[np.dtype(ndt)], (name, k, j, acc)
                    else:
                        assert np.array_equal(got[:, j], want), (name, k, j, acc)
    # a block whose base address is not 16-byte aligned (a view one element into a buffer): scalar column path
    A = mats["rand"]
    Ad = upload("csr", A, dev)
    k = 8
    Xh = rng.uniform(-1, 1, (A["num_cols"], k)).astype(ndt)
    buf = torch.zeros(A["num_cols"] * k + 1, dtype=tdt, device=dev)
    buf[1:] = tdev(Xh, dev).reshape(-1)
    X = buf[1:].view(A["num_cols"], k)
    Y = torch.empty(A["num_rows"], k, dtype=tdt, device=dev)
    cusp.multiply_block(Ad, X, Y)
    got = Y.cpu().numpy()
    for j in range(k):
        assert np.array_equal(got[:, j], O.spmv(A, np.ascontiguousarray(Xh[:, j])))
    # degenerate shapes
    e = torch.zeros(0, dtype=tdt, device=dev)
    A0 = cusp.csr_matrix(4, 5, torch.zeros(5, dtype=torch.int32, device=dev), torch.zeros(0, dtype=torch.int32, device=dev), e)
    Y = torch.full((4, 3), 7, dtype=tdt, device=dev)
    cusp.multiply_block(A0, torch.ones(5, 3, dtype=tdt, device=dev), Y)
    assert torch.equal(Y, torch.zeros_like(Y))
    with pytest.raises(capi.InvalidInput):
        cusp.multiply_block(A0, torch.ones(4, 3, dtype=tdt, device=dev), Y)


# ---------------------------------------------------------------------------
# single-pass HYB (spmv_hyb_fused.cu)
# ---------------------------------------------------------------------------
def _hyb_direct(h, Ad, xd, yd, acc, coo_cfg, fused):
    os.environ["B200SP_HYB_FUSED"] = fused
    try:
        e, c = Ad.ell, Ad.coo
        h.spmv_hyb(Ad.num_rows, Ad.num_cols, e.num_cols_per_row, e.pitch, e.column_indices, e.values, c.num_entries,
                   c.row_indices, c.column_indices, c.values, xd, yd, accumulate=acc, coo_cfg=coo_cfg)
        torch.cuda.synchronize()
    finally:
        os.environ.pop("B200SP_HYB_FUSED", None)


@pytest.mark.parametrize("ndt,tdt", DTYPES)
def test_hyb_single_pass_equals_two_pass(ndt, tdt, dev, handle):
    """opt-in (B200SP_HYB_FUSED=1; measured slower than two launches, spmv_hyb_fused.cu) — the fused kernel (ELL rows
    of a tile's row range, then the tile's tail entries accumulating) adds, per row, the
    same two values in the same order as ELL launch + COO launch with the same tile shape: bit-identical y on any
    data, assign and accumulate, every instantiated shape; and within the regrouped-sum bar of the reference's
    sequential loops (sequential/multiply/hyb_spmv.h:35-57).  Matrices: a stencil split below its row length (one
    or more tail entries in every row), a random matrix, hub rows spanning many tiles, a tail concentrated in a few
    rows (tiles that own thousands of tail-free rows), tail-free rows at both ends"""
    rng = np.random.default_rng(33)
    cases = []
    p7 = O.poisson(7, (23, 19, 17), ndt, "coo")
    cases += [("p7_k6", p7, 6), ("p7_k3", p7, 3)]
    rnd = O.gallery_random(5000, 4000, 60000, ndt, "coo")
    cases += [("rand_k4", rnd, 4), ("rand_k1", rnd, 1)]
    # hubs: rows 10 and 4200 with 20 000 / 3 000 entries on top of a sparse background; rows >= 4500 empty
    bg = O.gallery_random(4500, 3000, 9000, ndt, "coo")
    rows_h = np.concatenate([bg["row_indices"], np.full(20000, 10, np.int32), np.full(3000, 4200, np.int32)])
    cols_h = np.concatenate([bg["column_indices"], rng.integers(0, 3000, 23000).astype(np.int32)])
    order = np.lexsort((cols_h, rows_h))
    hub = dict(format="coo", num_rows=6000, num_cols=3000, num_entries=len(rows_h), row_indices=rows_h[order],
               column_indices=cols_h[order], values=np.ones(len(rows_h), ndt))
    cases += [("hub_k2", hub, 2), ("hub_k8", hub, 8)]
    for name, coo, K in cases:
        coo = dict(coo)
        coo["values"] = (coo["values"] * rng.uniform(0.5, 1.5, coo["num_entries"])).astype(ndt)
        A = O.convert(coo, "hyb", num_entries_per_row=K)
        assert A["coo"]["num_entries"] > 0, name
        Ad = upload("hyb", A, dev)
        x = rng.uniform(-1, 1, A["num_cols"]).astype(ndt)
        y0 = rng.uniform(-1, 1, A["num_rows"]).astype(ndt)
        xd = tdev(x, dev)
        scale = O.spmv(abs_matrix(A), np.abs(x))
        for vw, u in ((4, 1), (4, 2), (8, 1), (8, 2)):
            cfg = capi.Cfg(kernel=capi.K_COO_WARP, block_size=256, vector_width=vw, unroll=u)
            for acc in (False, True):
                y2 = tdev(y0, dev)
                _hyb_direct(handle, Ad, xd, y2, acc, cfg, "0")
                y1 = tdev(y0, dev)
                n0 = handle.launch_count
                _hyb_direct(handle, Ad, xd, y1, acc, cfg, "1")
                assert handle.launch_count - n0 == 3, (name, vw, u)  # fused kernel, carry fix-up, gap work list
                assert torch.equal(y1, y2), (name, vw, u, acc)
                want = O.spmv(A, x, y0 if acc else None, accumulate=acc)
                assert scaled_err(y1.cpu().numpy(), want, scale + (np.abs(y0) if acc else 0)) <= TOL[np.dtype(ndt)], (name, vw, u, acc)
    # larger operators through the default dispatch (the tail picks warp tiles by itself from 2.1 M entries): a stencil
    # split at K = 6 — its first and last grid planes are 16 k rows without tail entries each, i.e. owner tiles and
    # the gap work list — and the hub matrix scaled up; integer data, exact against the oracle / the row degrees
    big = O.poisson(7, (128, 128, 160), ndt, "coo")
    Ab = upload("hyb", O.convert(big, "hyb", num_entries_per_row=6), dev)
    assert Ab.coo.num_entries > 148 * 8 * 256 * 7
    xh = ((np.arange(big["num_cols"]) % 21) - 10).astype(ndt)
    yb = torch.empty(Ab.num_rows, dtype=tdt, device=dev)
    os.environ["B200SP_HYB_FUSED"] = "1"
    try:
        cusp.multiply(Ab, tdev(xh, dev), yb)
        n0 = handle.launch_count
        cusp.multiply(Ab, tdev(xh, dev), yb)
        torch.cuda.synchronize()
    finally:
        os.environ.pop("B200SP_HYB_FUSED", None)
    assert handle.launch_count - n0 == 3  # fused kernel, carry fix-up, gap work list
    assert np.array_equal(yb.cpu().numpy(), O.spmv(O.convert(big, "csr"), xh))
    yb.zero_()
    cusp.multiply(Ab, tdev(xh, dev), yb)  # the default: ELL launch + tail launch
    assert np.array_equal(yb.cpu().numpy(), O.spmv(O.convert(big, "csr"), xh))


# ---------------------------------------------------------------------------
# generalized product (b200sp_spmv_generalized): functor triples by code
# ---------------------------------------------------------------------------
GEN_TRIPLES = [("constant", 1e30, "plus", "minimum"), ("identity", 0.0, "plus", "minimum"),
               ("constant", -1e30, "multiplies", "maximum"), ("constant", -1e30, "minimum", "maximum"),
               ("constant", 0.0, "project2nd", "plus"), ("constant", 3.0, "multiplies", "plus"),
               ("identity", 0.0, "maximum", "plus"), ("constant", 1e30, "maximum", "minimum"),
               ("identity", 0.0, "multiplies", "plus"), ("constant", 0.0, "multiplies", "plus"),
               ("constant", 2.0, "plus", "plus"), ("identity", 0.0, "project2nd", "maximum")]


@pytest.mark.parametrize("ndt,tdt", DTYPES)
@pytest.mark.parametrize("fmt", FORMATS)
def test_generalized_functor_triples(fmt, ndt, tdt, dev, handle):
    """every (initialize, combine, reduce) code x every format against the numpy restatement of the reference's host
    loops (oracle.spmv_generalized, pinned to the C oracle for the default triple): integer-valued data -> exact for
    every pair; rows with no entries get initialize(y) only; long rows (sub-warp CSR, COO tiles, hub rows)."""
    rng = np.random.default_rng(31)
    mats = [O.poisson(5, (37, 29), ndt, "coo")]
    if fmt != "dia":
        for m, n, s in ((355, 378, 2340), (64, 3000, 60000)):   # ragged rows; 64 long rows (~900 entries each)
            coo = O.gallery_random(m, n, s, ndt, "coo")
            coo["values"] = rng.integers(-4, 5, coo["num_entries"]).astype(ndt)
            mats.append(coo)
    for coo in mats:
        A = to_fmt(coo, fmt)
        Ad = upload_any(fmt, A, dev)
        x = rng.integers(-5, 6, coo["num_cols"]).astype(ndt)
        y0 = rng.integers(-8, 9, coo["num_rows"]).astype(ndt)
        xd = tdev(x, dev)
        for init, c0, comb, red in GEN_TRIPLES:
            want = O.spmv_generalized(A, x, y0, init, c0, comb, red)
            yd = tdev(y0, dev)
            handle.spmv_generalized(Ad.descriptor(), xd, yd, init, c0, comb, red)
            assert np.array_equal(yd.cpu().numpy(), want), (fmt, init, c0, comb, red, coo["num_rows"])


def test_generalized_min_plus_on_real_data(dev, handle):
    """(min, +) on non-integer data: min is exact, so any grouping gives the host loop's bits — one relaxation step of
    single-source shortest paths on a weighted R-MAT-like graph, all formats"""
    rng = np.random.default_rng(32)
    n, nnz = 20000, 400000
    rows = np.sort(rng.integers(0, n, nnz)).astype(np.int32)
    rows[3000:150000] = rows[3000]          # a hub row
    rows = np.sort(rows)
    cols = rng.integers(0, n, nnz).astype(np.int32)
    key = np.unique(rows.astype(np.int64) * n + cols)
    rows, cols = (key // n).astype(np.int32), (key % n).astype(np.int32)
    w = rng.uniform(0.1, 5.0, len(key)).astype(np.float32)
    coo = dict(format="coo", num_rows=n, num_cols=n, num_entries=len(key), row_indices=rows, column_indices=cols, values=w)
    dist = np.full(n, 1e30, np.float32)
    dist[rng.integers(0, n, 50)] = rng.uniform(0, 3, 50).astype(np.float32)
    for fmt in ("coo", "csr", "hyb"):
        A = to_fmt(coo, fmt)
        want = O.spmv_generalized(A, dist, dist, "identity", 0.0, "plus", "minimum")
        yd, xd = tdev(dist, dev), tdev(dist, dev)
        Ad = upload_any(fmt, A, dev)  # must outlive the call: the descriptor only borrows its arrays
        handle.spmv_generalized(Ad.descriptor(), xd, yd, "identity", 0.0, "plus", "minimum")
        assert np.array_equal(yd.cpu().numpy(), want), fmt


def test_csr_row_starts_bit_exact(dev, handle):
    """the balanced CSR kernel's preprocessing (gpu_compute_row_starts, cusp/system/cuda/ktt/csr_multiply.h:64-85)
    through b200sp_csr_row_starts against the oracle's restatement of the reference's host version: every entry of the
    int32 array, ragged and power-law rows, worker counts below / at / above the entry count"""
    rng = np.random.default_rng(33)
    for rows, maxlen in ((1, 5), (1000, 9), (40000, 40)):
        lens = rng.integers(0, maxlen, rows)
        lens[rng.integers(0, rows, max(1, rows // 7))] = 0
        if rows > 100:
            lens[rows // 3] = 200000  # a hub row spanning many chunks
        lens[0] += 1
        ro = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
        nnz = int(ro[-1])
        rod = tdev(ro, dev)
        for workers in (1, 3, 148 * 32, 148 * 32 * 16, nnz, nnz + 5):
            out = torch.full((workers,), -7, dtype=torch.int32, device=dev)
            handle.csr_row_starts(rows, nnz, rod, workers, out)
            assert np.array_equal(out.cpu().numpy(), O.compute_row_starts(ro, workers)), (rows, workers)
