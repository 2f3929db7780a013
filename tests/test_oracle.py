"""CPU suite, part 1: pins the oracle (oracle/oracle.cpp) against
  * every golden vector the reference's tests hold for this path
    (tests/golden/reference_fixtures.py, transcribed from testing/*.cu),
  * the reference's own host loops compiled unmodified (oracle/_ref) through the
    committed outputs tests/golden/ref_spmv.npz, and live when _ref is present,
  * scipy as an independent cross-check.
"""
import os

import numpy as np
import pytest

from golden import reference_fixtures as G
from helpers import rel_err
from oracle import oracle as O

FORMATS = ("csr", "coo", "dia", "ell", "hyb")


def _x(n, dt):
    return (np.arange(n) % 10).astype(dt)


@pytest.mark.parametrize("name", sorted(G.MULTIPLY_DENSE))
@pytest.mark.parametrize("fmt", FORMATS)
def test_multiply_fixtures_exact(name, fmt):
    """testing/multiply.cu:383-512: x = i%10, y0 = 10, expected = dense product, ASSERT_EQUAL"""
    D = G.MULTIPLY_DENSE[name].astype(np.float32)
    A = O.convert(O.dense_to_coo(D), fmt)
    x = _x(D.shape[1], np.float32)
    y = O.spmv(A, x, np.full(D.shape[0], 10, np.float32))
    assert np.array_equal(y, D @ x)
    # scaled variant, initialize = identity (multiply.cu:514-645): y = 10 + A x
    y = O.spmv(A, x, np.full(D.shape[0], 10, np.float32), accumulate=True)
    assert np.array_equal(y, 10 + D @ x)


@pytest.mark.parametrize("name", sorted(G.MULTIPLY_POISSON))
@pytest.mark.parametrize("fmt", FORMATS)
def test_multiply_poisson_fixtures_exact(name, fmt):
    A = O.poisson(5, G.MULTIPLY_POISSON[name], np.float32, fmt)
    D = O.to_dense(O.poisson(5, G.MULTIPLY_POISSON[name], np.float32, "dia"))
    x = _x(D.shape[1], np.float32)
    assert np.array_equal(O.spmv(A, x), D @ x)


def test_poisson_dense_images():
    """testing/poisson.cu:6-93"""
    assert np.array_equal(O.to_dense(O.poisson(5, (2, 3), np.float32)), G.POISSON5_2x3)
    assert np.array_equal(O.to_dense(O.poisson(9, (2, 3), np.float32)), G.POISSON9_2x3)
    assert np.array_equal(O.to_dense(O.poisson(7, (2, 2, 2), np.float32)), G.POISSON7_2x2x2)
    assert np.array_equal(O.to_dense(O.poisson(27, (2, 2, 2), np.float32)), G.POISSON27_2x2x2)


def _same_layout(a, b):
    for k, v in b.items():
        if isinstance(v, dict):
            _same_layout(a[k], v)
        elif isinstance(v, np.ndarray):
            assert np.array_equal(a[k], v), k
        else:
            assert a[k] == v, (k, a[k], v)


def test_convert_exact_layouts():
    """testing/convert.cu:63-200 and :405-497: exact array layouts at alignment 1"""
    _same_layout(O.convert(G.CONVERT_CSR, "dia", alignment=1), G.CONVERT_DIA)
    _same_layout(O.convert(G.CONVERT_CSR, "ell", alignment=1), G.CONVERT_ELL)
    _same_layout(O.convert(G.CONVERT_CSR, "coo"), G.CONVERT_COO)
    _same_layout(O.convert(G.CONVERT_COO, "csr"), G.CONVERT_CSR)
    _same_layout(O.convert(G.CONVERT_CSR, "hyb", alignment=1, num_entries_per_row=1), G.CONVERT_HYB)
    _same_layout(O.convert(G.CONVERT_DIA, "csr"), G.CONVERT_CSR)
    for src in (G.CONVERT_CSR, G.CONVERT_COO, G.CONVERT_DIA, G.CONVERT_ELL, G.CONVERT_HYB):
        assert np.array_equal(O.to_dense(src), G.CONVERT_DENSE)
        for fmt in FORMATS:
            assert np.array_equal(O.to_dense(O.convert(src, fmt)), G.CONVERT_DENSE), (src["format"], fmt)


def test_default_alignment_is_32():
    """cusp/detail/ell_matrix.inl:35-36: pitch = round_up(rows, 32)"""
    e = O.convert(G.CONVERT_CSR, "ell")
    assert e["pitch"] == 32 and len(e["values"]) == 96
    assert np.array_equal(O.to_dense(e), G.CONVERT_DENSE)


def test_format_utils():
    """testing/format_utils.cu:13-75"""
    coo = dict(format="coo", num_rows=7, num_cols=7, num_entries=10, row_indices=G.INDICES,
               column_indices=np.zeros(10, np.int32), values=np.ones(10, np.float32))
    csr = O.coo_to_csr(coo)
    assert np.array_equal(csr["row_offsets"], G.OFFSETS)
    assert np.array_equal(O.csr_to_coo(csr)["row_indices"], G.INDICES)


def test_hyb_split_rule():
    """compute_optimal_entries_per_row (generic/format_utils.inl:281-321) with
    relative_speed 3, breakeven 4096 (csr_to_other.h:250-253).  'parity unpinned' in
    the reference (no direct test); pinned here by hand evaluation of the rule."""
    # 10000 rows: 6000 of length 2, 3000 of length 5, 1000 of length 40
    lens = np.array([2] * 6000 + [5] * 3000 + [40] * 1000)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    # k=0: rows longer than 0 = 10000 -> 3*10000 < 10000 false, 10000 < 4096 false
    # k=2: longer than 2 = 4000 -> 12000 < 10000 false; 4000 < 4096 TRUE -> K = 2
    assert O.optimal_entries_per_row(offs) == 2
    # small matrices (< 4096 rows) always give K = 0: everything goes to COO
    assert O.optimal_entries_per_row(G.CONVERT_CSR["row_offsets"]) == 0
    lens = np.array([7] * 20000)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    assert O.optimal_entries_per_row(offs) == 7  # no k < max satisfies the rule -> max


def test_blas_known_answers():
    """testing/blas.cu:97-142, 287-352, 434-453"""
    f = np.float32
    a = G.BLAS_AXPBY
    assert np.array_equal(O.axpby(np.array(a["x"], f), np.array(a["y"], f), a["alpha"], a["beta"]), np.array(a["z"], f))
    a = G.BLAS_AXPY
    assert np.array_equal(O.axpy(np.array(a["x"], f), np.array(a["y"], f), a["alpha"]), np.array(a["out"], f))
    a = G.BLAS_DOT
    assert O.dot(np.array(a["x"], f), np.array(a["y"], f)) == a["result"]
    a = G.BLAS_NRM2
    assert O.nrm2(np.array(a["x"], f)) == a["result"]


def test_cg_reference_cases():
    """testing/cg.cu:46-99"""
    c = G.CG_CASE
    A = O.poisson(5, c["grid"], np.float32, "csr")
    b = np.ones(A["num_rows"], np.float32)
    x, it, conv, hist = O.cg(A, np.zeros_like(b), b, c["limit"], c["rel"])
    r = b - O.spmv(A, x)
    assert np.linalg.norm(r) < 1e-4 * np.linalg.norm(b)
    assert conv and it <= c["limit"] and len(hist) == it + 1
    # zero residual: diag(8,4), x = 1, b = A x -> 0 iterations, converged
    A = O.convert(O.dense_to_coo(np.array([[8, 0], [0, 4]], np.float32)), "csr")
    x0 = np.ones(2, np.float32)
    b = O.spmv(A, x0)
    x, it, conv, hist = O.cg(A, x0, b, 20, 0.0)
    assert it == 0 and conv and np.array_equal(x, x0)


def test_against_committed_reference_outputs(golden):
    """oracle == the reference's own templates (outputs generated by
    tests/golden/make_golden.py from oracle/_ref), bit for bit"""
    rng_cases = [("p5", 5, (13, 9)), ("p7", 7, (7, 6, 5)), ("p9", 9, (6, 7)), ("p27", 27, (4, 3, 5))]
    n = 0
    for name, st, grid in rng_cases:
        for dt in (np.float32, np.float64):
            dia = O.poisson(st, grid, dt, "dia")
            for fmt in FORMATS:
                key = f"{name}_{np.dtype(dt).name}_{fmt}"
                A = O.convert(dia, fmt)
                assert np.array_equal(O.spmv(A, golden[key + "_x"]), golden[key + "_y"]), key
                assert np.array_equal(O.spmv(A, golden[key + "_x"], golden[key + "_y0"], accumulate=True),
                                      golden[key + "_yacc"]), key
                n += 1
    for m, nn, s in ((24, 24, 150), (24, 12, 20), (300, 257, 4000)):
        for dt in (np.float32, np.float64):
            coo = O.gallery_random(m, nn, s, dt, "coo")
            for fmt in ("csr", "coo", "ell", "hyb"):
                key = f"rand{m}x{nn}_{np.dtype(dt).name}_{fmt}"
                coo["values"] = golden[key + "_vals"]
                A = O.convert(coo, fmt)
                assert np.array_equal(O.spmv(A, golden[key + "_x"]), golden[key + "_y"]), key
                n += 1
    assert n == 64


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built (reference absent)")
@pytest.mark.parametrize("fmt", FORMATS)
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_live_reference_equals_oracle(fmt, dt):
    rng = np.random.default_rng(5)
    A = O.convert(O.poisson(7, (9, 8, 7), dt, "dia"), fmt)
    if fmt != "dia":
        # perturb values so products are not exactly representable
        for tgt in ([A["ell"], A["coo"]] if fmt == "hyb" else [A]):
            tgt["values"] = (tgt["values"] * rng.uniform(0.5, 1.5, len(tgt["values"]))).astype(dt)
    x = rng.uniform(-1, 1, A["num_cols"]).astype(dt)
    y0 = rng.uniform(-1, 1, A["num_rows"]).astype(dt)
    for acc in (False, True):
        a = O.spmv(A, x, y0 if acc else None, accumulate=acc)
        b = O.spmv(A, x, y0 if acc else None, accumulate=acc, impl="ref")
        assert np.array_equal(a, b)
        if fmt in ("csr", "dia", "ell"):  # row-block threaded runs of the reference loop
            c = O.spmv(A, x, y0 if acc else None, accumulate=acc, impl="ref", nthreads=4)
            d = O.spmv(A, x, y0 if acc else None, accumulate=acc, impl="oracle_mt", nthreads=3)
            assert np.array_equal(a, c) and np.array_equal(a, d)


def test_scipy_cross_check():
    sp = pytest.importorskip("scipy.sparse")
    rng = np.random.default_rng(11)
    coo = O.gallery_random(200, 150, 3000, np.float64, "coo")
    coo["values"] = rng.uniform(0.5, 1.5, coo["num_entries"])
    S = sp.coo_matrix((coo["values"], (coo["row_indices"], coo["column_indices"])), shape=(200, 150)).tocsr()
    x = rng.uniform(0.5, 1.5, 150)
    want = S @ x
    for fmt in ("csr", "coo", "ell", "hyb"):
        assert rel_err(O.spmv(O.convert(coo, fmt), x), want) < 1e-13
    # structure: CSR from the oracle equals scipy's canonical CSR
    csr = O.convert(coo, "csr")
    assert np.array_equal(csr["row_offsets"], S.indptr) and np.array_equal(csr["column_indices"], S.indices)


def test_ellr_row_lengths_and_spmv():
    e = O.convert(G.CONVERT_CSR, "ell", alignment=1)
    r = O.to_ellr(e)
    assert np.array_equal(r["row_lengths"], [2, 1, 3, 1])
    x = np.arange(4, dtype=np.float32)
    assert np.array_equal(O.spmv(r, x), O.spmv(e, x))


def test_make_diagonal_symmetric():
    """cusp/ktt/matrix_generation.h:64-102 and testing/ktt.cu:274-281"""
    A = O.make_diagonal_symmetric(8, 8, 1, 3)
    assert list(A["diagonal_offsets"]) == [-1, 0, 1] and A["num_entries"] == 22
    assert np.array_equal(O.to_dense(A), np.eye(8) + np.eye(8, k=1) + np.eye(8, k=-1))
    with pytest.raises(RuntimeError):
        O.make_diagonal_symmetric(4, 4, 1, 64)
    for rows, cols, step, cnt in G.KTT_BANDED:
        B = O.make_diagonal_symmetric(rows, cols, step, cnt)
        assert B["diagonal_offsets"][0] == -512 and len(B["diagonal_offsets"]) == 1024


def test_reference_gpu_kernel_harness_enumerates_the_ktt_dia_space():
    """oracle/_ref/libcuspref_gpu.so (the reference's KTT DIA kernel compiled unmodified, measurement
    infrastructure for bench.py's reference_cuda_kernel leg): loads without a GPU, exports its C ABI and
    lists the 42 points of cusp/system/cuda/ktt/dia_multiply.h:24-55; only sm_100a code inside"""
    import ctypes as C
    import re
    import subprocess
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libcuspref_gpu.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libcuspref_gpu.so not built (needs /root/reference)")
    lib = C.CDLL(path)
    assert lib.cuspref_dia_num_configs() == 42
    p = (C.c_int * 4)()
    seen = set()
    for k in range(42):
        assert lib.cuspref_dia_config(k, p) == 0
        bs, pf, pt, sl = tuple(p)
        assert bs in (128, 256, 512) and pf in (0, 2, 3, 4) and pt in (0, 1) and sl in (0, 1)
        assert pf > 0 or pt == 0  # the KTT constraint
        seen.add((bs, pf, pt, sl))
    assert len(seen) == 42 and lib.cuspref_dia_config(42, p) != 0
    assert hasattr(lib, "cuspref_dia_spmv")
    if os.path.exists("/usr/local/cuda/bin/cuobjdump"):
        out = subprocess.check_output(["/usr/local/cuda/bin/cuobjdump", "-res-usage", path], text=True)
        assert len(re.findall(r"ktt_dia_vector_kernel", out)) == 84  # 42 points x {float, double}


def test_committed_cg_history_is_the_oracles():
    """tests/golden/cg_poisson7pt_f64.json (what bench.py and the full-size GPU test compare the solver's residual
    history with) is reproduced by the oracle CG on the grid that finishes in a blink; the 512^3 entry has the
    shape BASELINE configs[4] names"""
    import json
    import os
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cg_poisson7pt_f64.json")))
    g = gold["64x64x64"]
    A = O.poisson(7, (64, 64, 64), np.float64, "csr")
    n = A["num_rows"]
    x, it, conv, hist = O.cg(A, np.zeros(n), np.ones(n), g["iterations"], 0.0, 0.0)
    assert it == g["iterations"] and np.array_equal(hist, np.asarray(g["residuals"]))
    xc, itc, convc, histc = O.cg(A, np.zeros(n), np.ones(n), g["iterations"], 0.0, 0.0, compensated=True)
    assert np.array_equal(histc, np.asarray(g["residuals_compensated"]))
    assert np.max(np.abs(hist - histc) / histc) <= 1e-10   # 2.6e5 terms: the two summations still agree closely (7e-12)
    big = gold["512x512x512"]
    assert big["rows"] == 512 ** 3 and big["nnz"] == 937951232 and len(big["residuals"]) == big["iterations"] + 1 == 51
    assert len(big["residuals_compensated"]) == 51 and big["sequential_vs_compensated_max_rel_dev"] < 1e-7
    assert abs(big["residuals"][0] - np.sqrt(512 ** 3)) < 1e-6


def test_compute_row_starts_restates_the_balanced_csr_preprocessing():
    """cpu_compute_row_starts (cusp/system/cuda/ktt/csr_multiply.h:38-61): out[w] = the row containing entry
    w * ceil(nnz / workers), 0 beyond the matrix — against the definition by binary search, ragged rows incl. empty ones"""
    rng = np.random.default_rng(8)
    for rows in (1, 7, 300):
        lens = rng.integers(0, 9, rows)
        lens[rng.integers(0, rows, max(1, rows // 5))] = 0
        lens[0] += 1
        ro = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
        nnz = int(ro[-1])
        for workers in (1, 2, 5, 64, nnz, nnz + 9):
            got = O.compute_row_starts(ro, workers)
            chunk = -(-nnz // workers)
            e = np.arange(workers, dtype=np.int64) * chunk
            want = np.where(e < nnz, np.searchsorted(ro, e, side="right") - 1, 0).astype(np.int32)
            assert np.array_equal(got, want), (rows, workers)
