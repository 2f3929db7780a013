"""numpy front end of the CPU oracle (oracle/oracle.cpp) and of the reference's
own host loops (oracle/_ref/libcuspref.so, built from /root/reference).

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs — never by the product package.
Matrices are plain dicts of numpy arrays in the reference's layouts.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
_REF = os.path.join(_HERE, "_ref", "libcuspref.so")
_lib = None
_ref = None
I64 = C.c_int64


def build():
    """compile liboracle.so (always) and _ref/libcuspref.so (only where /root/reference exists)"""
    subprocess.check_call(["make", "-s", "-C", _HERE], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
        for s in ("f32", "f64"):
            ct = C.c_float if s == "f32" else C.c_double
            getattr(_lib, "oracle_dot_" + s).restype = ct
            getattr(_lib, "oracle_nrm2_" + s).restype = ct
            for n in ("cg_csr", "cg_csr_compensated", "krylov_csr", "stencil_dia", "dia_to_coo", "csr_to_hyb", "csr_to_dia", "ell_to_coo", "hyb_to_coo"):
                getattr(_lib, f"oracle_{n}_{s}").restype = I64
        for n in ("max_entries_per_row", "optimal_entries_per_row", "gallery_random", "make_diagonal"):
            getattr(_lib, "oracle_" + n).restype = I64
    return _lib


def ref_available() -> bool:
    return os.path.exists(_REF)


def ref():
    global _ref
    if _ref is None:
        _ref = C.CDLL(_REF)
    return _ref


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def _s(dt) -> str:
    dt = np.dtype(dt)
    if dt == np.float32:
        return "f32"
    if dt == np.float64:
        return "f64"
    raise TypeError(dt)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _ct(dt):
    return C.c_float if np.dtype(dt) == np.float32 else C.c_double


def _c(a, dt=None):
    return np.ascontiguousarray(a, dtype=dt)


# ---------------------------------------------------------------------------
# SpMV
# ---------------------------------------------------------------------------
def spmv(A: dict, x, y=None, accumulate=False, impl="oracle", nthreads=1):
    """y = A x or y += A x for a matrix dict with key 'format'.
    impl = 'oracle' (restatement), 'oracle_mt' (row-parallel restatement),
    'ref' (the reference's own templates, oracle/_ref)."""
    fmt = A["format"]
    vals = A["ell"]["values"] if fmt == "hyb" else A["values"]
    dt = vals.dtype
    s = _s(dt)
    x = _c(x, dt)
    rows, cols = int(A["num_rows"]), int(A["num_cols"])
    if y is None:
        y = np.zeros(rows, dtype=dt)
        assert not accumulate
    else:
        y = _c(y, dt).copy()
    acc = int(bool(accumulate))
    L = lib()
    if impl == "ref":
        R = ref()
        if fmt == "csr":
            getattr(R, "ref_spmv_csr_" + s)(I64(rows), I64(cols), I64(len(vals)), _p(A["row_offsets"]),
                                            _p(A["column_indices"]), _p(vals), _p(x), _p(y), acc, nthreads)
        elif fmt == "coo":
            getattr(R, "ref_spmv_coo_" + s)(I64(rows), I64(cols), I64(len(vals)), _p(A["row_indices"]),
                                            _p(A["column_indices"]), _p(vals), _p(x), _p(y), acc, nthreads)
        elif fmt == "dia":
            getattr(R, "ref_spmv_dia_" + s)(I64(rows), I64(cols), I64(len(A["diagonal_offsets"])),
                                            I64(A["pitch"]), _p(A["diagonal_offsets"]), _p(vals), _p(x), _p(y),
                                            acc, nthreads)
        elif fmt == "ell":
            getattr(R, "ref_spmv_ell_" + s)(I64(rows), I64(cols), I64(A["num_cols_per_row"]), I64(A["pitch"]),
                                            _p(A["column_indices"]), _p(vals), _p(x), _p(y), acc, nthreads)
        elif fmt == "hyb":
            e, c = A["ell"], A["coo"]
            getattr(R, "ref_spmv_hyb_" + s)(I64(rows), I64(cols), I64(e["num_cols_per_row"]), I64(e["pitch"]),
                                            _p(e["column_indices"]), _p(e["values"]), I64(len(c["values"])),
                                            _p(c["row_indices"]), _p(c["column_indices"]), _p(c["values"]),
                                            _p(x), _p(y), acc)
        else:
            raise ValueError(fmt)
        return y
    mt = impl == "oracle_mt"
    if fmt == "csr":
        if mt:
            getattr(L, "oracle_spmv_csr_mt_" + s)(I64(rows), _p(A["row_offsets"]), _p(A["column_indices"]),
                                                  _p(vals), _p(x), _p(y), acc, nthreads)
        else:
            getattr(L, "oracle_spmv_csr_" + s)(I64(rows), _p(A["row_offsets"]), _p(A["column_indices"]),
                                               _p(vals), _p(x), _p(y), acc)
    elif fmt == "coo":
        getattr(L, "oracle_spmv_coo_" + s)(I64(rows), I64(len(vals)), _p(A["row_indices"]),
                                           _p(A["column_indices"]), _p(vals), _p(x), _p(y), acc)
    elif fmt == "dia":
        args = (I64(rows), I64(cols), I64(len(A["diagonal_offsets"])), I64(A["pitch"]),
                _p(A["diagonal_offsets"]), _p(vals), _p(x), _p(y), acc)
        if mt:
            getattr(L, "oracle_spmv_dia_mt_" + s)(*args, nthreads)
        else:
            getattr(L, "oracle_spmv_dia_" + s)(*args)
    elif fmt == "ell":
        args = (I64(rows), I64(A["num_cols_per_row"]), I64(A["pitch"]), _p(A["column_indices"]), _p(vals),
                _p(x), _p(y), acc)
        if mt:
            getattr(L, "oracle_spmv_ell_mt_" + s)(*args, nthreads)
        else:
            getattr(L, "oracle_spmv_ell_" + s)(*args)
    elif fmt == "ellr":
        getattr(L, "oracle_spmv_ellr_" + s)(I64(rows), I64(A["num_cols_per_row"]), I64(A["pitch"]),
                                            _p(A["column_indices"]), _p(vals), _p(A["row_lengths"]), _p(x),
                                            _p(y), acc)
    elif fmt == "hyb":
        e, c = A["ell"], A["coo"]
        getattr(L, "oracle_spmv_hyb_" + s)(I64(rows), I64(e["num_cols_per_row"]), I64(e["pitch"]),
                                           _p(e["column_indices"]), _p(e["values"]), I64(len(c["values"])),
                                           _p(c["row_indices"]), _p(c["column_indices"]), _p(c["values"]),
                                           _p(x), _p(y), acc)
    else:
        raise ValueError(fmt)
    return y


# ---------------------------------------------------------------------------
# generalized product (numpy restatement of the reference's host loops with a functor triple)
# ---------------------------------------------------------------------------
_COMBINE = {"multiplies": lambda a, x: a * x, "plus": lambda a, x: a + x,
            "minimum": lambda a, x: np.where(x < a, x, a),      # thrust::minimum: rhs < lhs ? rhs : lhs
            "maximum": lambda a, x: np.where(a < x, x, a),      # thrust::maximum: lhs < rhs ? rhs : lhs
            "project2nd": lambda a, x: x + 0 * a}
_REDUCE = {"plus": lambda a, b: a + b, "minimum": lambda a, b: np.where(b < a, b, a), "maximum": lambda a, b: np.where(a < b, b, a)}


def stored_entries(A: dict):
    """(rows, cols, values) of the entries the reference's host loop of A's format visits, in that loop's order per
    row: CSR / COO storage order (sequential/multiply/csr_spmv.h:44-62, coo_spmv.h:45-62); ELL slot by slot without
    the padding (ell_spmv.h:52-72); DIA diagonal by diagonal, every slot whose column lies inside the matrix —
    explicit zeros included (dia_spmv.h:52-79); HYB = ELL part then COO part (hyb_spmv.h:45-56)."""
    fmt = A["format"]
    rows, cols = int(A["num_rows"]), int(A["num_cols"])
    if fmt == "csr":
        return csr_to_coo(A)["row_indices"].astype(np.int64), A["column_indices"].astype(np.int64), A["values"]
    if fmt == "coo":
        return A["row_indices"].astype(np.int64), A["column_indices"].astype(np.int64), A["values"]
    if fmt in ("ell", "ellr"):
        K, pitch = int(A["num_cols_per_row"]), int(A["pitch"])
        cj = A["column_indices"].reshape(K, pitch)[:, :rows]
        av = A["values"].reshape(K, pitch)[:, :rows]
        keep = cj >= 0
        if fmt == "ellr":
            keep &= np.arange(K)[:, None] < A["row_lengths"][None, :rows]
        ri = np.broadcast_to(np.arange(rows)[None, :], cj.shape)
        return ri[keep].astype(np.int64), cj[keep].astype(np.int64), av[keep]   # slot-major = per-row slot order
    if fmt == "dia":
        nd, pitch = len(A["diagonal_offsets"]), int(A["pitch"])
        av = A["values"].reshape(nd, pitch)[:, :rows]
        ri = np.broadcast_to(np.arange(rows, dtype=np.int64)[None, :], av.shape)
        cj = ri + A["diagonal_offsets"].astype(np.int64)[:, None]
        keep = (cj >= 0) & (cj < cols)
        return ri[keep], cj[keep], av[keep]
    if fmt == "hyb":
        r1, c1, v1 = stored_entries(A["ell"])
        r2, c2, v2 = stored_entries(A["coo"])
        return np.concatenate([r1, r2]), np.concatenate([c1, c2]), np.concatenate([v1, v2])
    raise ValueError(fmt)


def spmv_generalized(A: dict, x, y, initialize="constant", init_value=0.0, combine="multiplies", reduce="plus"):
    """y[i] = reduce(initialize(y[i]), combine(a_ij, x_j) ...) over the stored entries of row i, one entry at a time
    in the host loop's order (cusp/multiply.h:163-195; generic/multiply/generalized_spmv.h:61-303)."""
    vals = A["ell"]["values"] if A["format"] == "hyb" else A["values"]
    dt = vals.dtype
    x = _c(x, dt)
    acc = _c(y, dt).copy() if initialize == "identity" else np.full(int(A["num_rows"]), init_value, dtype=dt)
    ri, cj, av = stored_entries(A)
    if len(ri) == 0:
        return acc
    prod = _COMBINE[combine](av.astype(dt), x[cj]).astype(dt)
    # one entry per row per round keeps the per-row order of the reductions (and the rounding of `plus`)
    order = np.argsort(ri, kind="stable")
    ri, prod = ri[order], prod[order]
    first = np.r_[0, np.flatnonzero(np.diff(ri)) + 1]
    rank_in_row = np.arange(len(ri)) - np.repeat(first, np.diff(np.r_[first, len(ri)]))
    red = _REDUCE[reduce]
    for k in range(int(rank_in_row.max()) + 1):
        sel = rank_in_row == k
        acc[ri[sel]] = red(acc[ri[sel]], prod[sel]).astype(dt)
    return acc


# ---------------------------------------------------------------------------
# BLAS-1 / CG
# ---------------------------------------------------------------------------
def axpy(x, y, alpha):
    y = y.copy()
    getattr(lib(), "oracle_axpy_" + _s(y.dtype))(I64(len(y)), _ct(y.dtype)(alpha), _p(_c(x, y.dtype)), _p(y))
    return y


def axpby(x, y, alpha, beta):
    z = np.empty_like(x)
    ct = _ct(x.dtype)
    getattr(lib(), "oracle_axpby_" + _s(x.dtype))(I64(len(x)), ct(alpha), _p(_c(x)), ct(beta),
                                                  _p(_c(y, x.dtype)), _p(z))
    return z


def dot(x, y):
    return getattr(lib(), "oracle_dot_" + _s(x.dtype))(I64(len(x)), _p(_c(x)), _p(_c(y, x.dtype)))


def nrm2(x):
    return getattr(lib(), "oracle_nrm2_" + _s(x.dtype))(I64(len(x)), _p(_c(x)))


def cg(A: dict, x0, b, iteration_limit=500, relative_tolerance=1e-5, absolute_tolerance=0.0, compensated=False):
    """cusp::krylov::cg on a CSR dict.  Returns (x, iterations, converged, residuals).
    compensated=True: dot products and norms accumulated in long double and rounded once (the same iteration with
    the summation error taken out: what both the sequential sums and the engine's tree sums approximate)."""
    assert A["format"] == "csr"
    dt = A["values"].dtype
    x = _c(x0, dt).copy()
    b = _c(b, dt)
    hist = np.zeros(iteration_limit + 2, dtype=np.float64)
    nres = I64(0)
    conv = C.c_int(0)
    it = getattr(lib(), ("oracle_cg_csr_compensated_" if compensated else "oracle_cg_csr_") + _s(dt))(
        I64(A["num_rows"]), _p(A["row_offsets"]), _p(A["column_indices"]), _p(A["values"]), _p(x), _p(b),
        I64(iteration_limit), C.c_double(relative_tolerance), C.c_double(absolute_tolerance), _p(hist),
        C.byref(nres), C.byref(conv))
    return x, int(it), bool(conv.value), hist[: nres.value].copy()


SOLVERS = {"cg": 0, "bicgstab": 1, "cr": 2}


def krylov(solver: str, A: dict, x0, b, iteration_limit=500, relative_tolerance=1e-5, absolute_tolerance=0.0, dinv=None):
    """cusp::krylov::{cg,bicgstab,cr}(A, x, b, monitor, M) on a CSR dict with M = identity (dinv=None) or
    cusp::precond::diagonal (dinv = 1 / diag(A)).  Returns (x, iterations, converged, residuals): residuals holds one
    entry per monitor.finished() call, like monitor.residuals (BiCGStab: two per iteration)."""
    assert A["format"] == "csr"
    dt = A["values"].dtype
    x = _c(x0, dt).copy()
    b = _c(b, dt)
    dv = None if dinv is None else _c(dinv, dt)
    hist = np.zeros(2 * iteration_limit + 4, dtype=np.float64)
    nres = I64(0)
    conv = C.c_int(0)
    it = getattr(lib(), "oracle_krylov_csr_" + _s(dt))(
        C.c_int(SOLVERS[solver]), I64(A["num_rows"]), _p(A["row_offsets"]), _p(A["column_indices"]), _p(A["values"]),
        _p(dv) if dv is not None else None, _p(x), _p(b), I64(iteration_limit), C.c_double(relative_tolerance),
        C.c_double(absolute_tolerance), _p(hist), C.byref(nres), C.byref(conv))
    return x, int(it), bool(conv.value), hist[: nres.value].copy()


def extract_diagonal(A: dict):
    """cusp::extract_diagonal on a CSR dict (cusp/format_utils.h): the stored a_ii, 0 where the row has none"""
    assert A["format"] == "csr"
    n = A["num_rows"]
    d = np.zeros(n, dtype=A["values"].dtype)
    ri = csr_to_coo(A)["row_indices"]
    on = ri == A["column_indices"]
    d[ri[on]] = A["values"][on]
    return d


# ---------------------------------------------------------------------------
# gallery
# ---------------------------------------------------------------------------
_STENCILS = {
    # (points, centre value); dimension 0 fastest — cusp/gallery/detail/poisson.inl:29-96
    5: [((0, -1), -1), ((-1, 0), -1), ((0, 0), 4), ((1, 0), -1), ((0, 1), -1)],
    9: [((-1, -1), -1), ((0, -1), -1), ((1, -1), -1), ((-1, 0), -1), ((0, 0), 8), ((1, 0), -1),
        ((-1, 1), -1), ((0, 1), -1), ((1, 1), -1)],
    7: [((0, 0, -1), -1), ((0, -1, 0), -1), ((-1, 0, 0), -1), ((0, 0, 0), 6), ((1, 0, 0), -1),
        ((0, 1, 0), -1), ((0, 0, 1), -1)],
    27: [((i, j, k), 26 if (i, j, k) == (0, 0, 0) else -1)
         for k in (-1, 0, 1) for j in (-1, 0, 1) for i in (-1, 0, 1)],
}


def stencil_dia(points, grid, dtype=np.float64) -> dict:
    """generate_matrix_from_stencil -> dia_matrix (pitch = num_rows)"""
    ndim = len(grid)
    pts = _c([p for p, _ in points], np.int32).reshape(len(points), ndim)
    pv = _c([v for _, v in points], dtype)
    g = _c(grid, np.int64)
    rows = int(np.prod(g))
    offs = np.zeros(len(points), np.int32)
    vals = np.zeros(len(points) * rows, dtype)
    nnz = getattr(lib(), "oracle_stencil_dia_" + _s(dtype))(ndim, _p(g), len(points), _p(pts), _p(pv),
                                                             _p(offs), _p(vals))
    return dict(format="dia", num_rows=rows, num_cols=rows, num_entries=int(nnz), diagonal_offsets=offs,
                pitch=rows, values=vals)


def poisson(stencil: int, grid, dtype=np.float64, fmt="dia") -> dict:
    """cusp::gallery::poisson{5,9,7,27}pt into `fmt` through the reference's
    conversion rules (DIA first, then convert)."""
    return convert(stencil_dia(_STENCILS[stencil], grid, dtype), fmt)


def gallery_random(m, n, samples, dtype=np.float32, fmt="coo") -> dict:
    """cusp::gallery::random (glibc rand(), values 1)"""
    Ai = np.zeros(samples, np.int32)
    Aj = np.zeros(samples, np.int32)
    k = lib().oracle_gallery_random(I64(m), I64(n), I64(samples), _p(Ai), _p(Aj))
    coo = dict(format="coo", num_rows=m, num_cols=n, num_entries=int(k), row_indices=Ai[:k].copy(),
               column_indices=Aj[:k].copy(), values=np.ones(k, dtype))
    return convert(coo, fmt)


def make_diagonal_symmetric(rows, cols, offset_step, diagonal_count) -> dict:
    """cusp::ktt::make_diagonal_symmetric_matrix (cusp/ktt/matrix_generation.h:64-102)"""
    start = int(-offset_step * diagonal_count / 2)  # C++ int division truncates toward zero
    offs = _c([start + offset_step * i for i in range(diagonal_count)], np.int32)
    vals = np.zeros(diagonal_count * rows, np.float32)
    nnz = lib().oracle_make_diagonal(I64(rows), I64(cols), I64(diagonal_count), _p(offs), _p(vals))
    if nnz < 0:
        raise RuntimeError("make_diagonal_symmetric_matrix: Too many diagonals.")
    return dict(format="dia", num_rows=rows, num_cols=cols, num_entries=int(nnz), diagonal_offsets=offs,
                pitch=rows, values=vals)


# ---------------------------------------------------------------------------
# conversions (cusp::convert; via COO/CSR where the reference has no direct path,
# cusp/system/detail/generic/convert.inl:53-70)
# ---------------------------------------------------------------------------
def round_up(n, k):
    return k * ((n + k - 1) // k)


def dense_to_coo(M, dtype=None) -> dict:
    """array2d -> coo: row-major scan, non-zeros only (cusp array2d_to_other)"""
    M = np.asarray(M)
    dtype = dtype or M.dtype
    r, c = np.nonzero(M)
    return dict(format="coo", num_rows=M.shape[0], num_cols=M.shape[1], num_entries=len(r),
                row_indices=r.astype(np.int32), column_indices=c.astype(np.int32),
                values=M[r, c].astype(dtype))


def to_dense(A: dict):
    fmt = A["format"]
    if fmt == "hyb":
        return to_dense(A["ell"]) + to_dense(A["coo"])
    D = np.zeros((A["num_rows"], A["num_cols"]), A["values"].dtype)
    if fmt == "coo":
        np.add.at(D, (A["row_indices"], A["column_indices"]), A["values"])
    elif fmt == "csr":
        idx = np.zeros(len(A["values"]), np.int32)
        lib().oracle_offsets_to_indices(I64(A["num_rows"]), _p(A["row_offsets"]), _p(idx))
        np.add.at(D, (idx, A["column_indices"]), A["values"])
    elif fmt in ("ell", "ellr"):
        p, K = A["pitch"], A["num_cols_per_row"]
        for k in range(K):
            for i in range(A["num_rows"]):
                j = A["column_indices"][k * p + i]
                if j != -1:
                    D[i, j] += A["values"][k * p + i]
    elif fmt == "dia":
        p = A["pitch"]
        for d, off in enumerate(A["diagonal_offsets"]):
            for i in range(A["num_rows"]):
                j = i + int(off)
                if 0 <= j < A["num_cols"]:
                    D[i, j] += A["values"][d * p + i]
    return D


def coo_to_csr(A: dict) -> dict:
    offs = np.zeros(A["num_rows"] + 1, np.int32)
    lib().oracle_indices_to_offsets(I64(len(A["values"])), _p(_c(A["row_indices"], np.int32)),
                                    I64(A["num_rows"]), _p(offs))
    return dict(format="csr", num_rows=A["num_rows"], num_cols=A["num_cols"], num_entries=len(A["values"]),
                row_offsets=offs, column_indices=A["column_indices"].copy(), values=A["values"].copy())


def csr_to_coo(A: dict) -> dict:
    idx = np.zeros(len(A["values"]), np.int32)
    lib().oracle_offsets_to_indices(I64(A["num_rows"]), _p(A["row_offsets"]), _p(idx))
    return dict(format="coo", num_rows=A["num_rows"], num_cols=A["num_cols"], num_entries=len(A["values"]),
                row_indices=idx, column_indices=A["column_indices"].copy(), values=A["values"].copy())


def compute_row_starts(row_offsets, workers: int):
    """cpu_compute_row_starts (cusp/system/cuda/ktt/csr_multiply.h:38-61)"""
    ro = _c(row_offsets, np.int32)
    out = np.zeros(max(workers, 1), np.int32)[:workers]
    lib().oracle_compute_row_starts(I64(len(ro) - 1), I64(int(ro[-1])), _p(ro), I64(workers), _p(out))
    return out


def optimal_entries_per_row(row_offsets, relative_speed=3.0, breakeven_threshold=4096) -> int:
    return int(lib().oracle_optimal_entries_per_row(I64(len(row_offsets) - 1), _p(_c(row_offsets, np.int32)),
                                                    C.c_float(relative_speed), I64(breakeven_threshold)))


def max_entries_per_row(row_offsets) -> int:
    return int(lib().oracle_max_entries_per_row(I64(len(row_offsets) - 1), _p(_c(row_offsets, np.int32))))


def convert(A: dict, fmt: str, alignment=32, num_entries_per_row=0) -> dict:
    src = A["format"]
    if src == fmt:
        return A
    dt = (A["ell"]["values"] if src == "hyb" else A["values"]).dtype
    s = _s(dt)
    L = lib()
    if src == "dia":
        rows, cols, p = A["num_rows"], A["num_cols"], A["pitch"]
        nd = len(A["diagonal_offsets"])
        if fmt in ("coo", "csr"):
            n = getattr(L, "oracle_dia_to_coo_" + s)(I64(rows), I64(cols), I64(nd), I64(p),
                                                     _p(A["diagonal_offsets"]), _p(A["values"]), None, None,
                                                     None)
            Ai, Aj, Ax = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, dt)
            getattr(L, "oracle_dia_to_coo_" + s)(I64(rows), I64(cols), I64(nd), I64(p),
                                                 _p(A["diagonal_offsets"]), _p(A["values"]), _p(Ai), _p(Aj),
                                                 _p(Ax))
            coo = dict(format="coo", num_rows=rows, num_cols=cols, num_entries=int(n), row_indices=Ai,
                       column_indices=Aj, values=Ax)
            return coo if fmt == "coo" else coo_to_csr(coo)
        if fmt == "ell":
            cidx = np.zeros(nd * p, np.int32)
            ev = np.zeros(nd * p, dt)
            getattr(L, "oracle_dia_to_ell_" + s)(I64(rows), I64(nd), I64(p), _p(A["diagonal_offsets"]),
                                                 _p(A["values"]), _p(cidx), _p(ev))
            return dict(format="ell", num_rows=rows, num_cols=cols, num_entries=A["num_entries"],
                        num_cols_per_row=nd, pitch=p, column_indices=cidx, values=ev)
        return convert(convert(A, "csr"), fmt, alignment, num_entries_per_row)
    if src == "coo":
        if fmt == "csr":
            return coo_to_csr(A)
        return convert(coo_to_csr(A), fmt, alignment, num_entries_per_row)
    if src == "csr":
        rows, cols = A["num_rows"], A["num_cols"]
        if fmt == "coo":
            return csr_to_coo(A)
        if fmt == "ell":
            K = num_entries_per_row or max_entries_per_row(A["row_offsets"])
            p = round_up(rows, alignment)
            cidx, ev = np.zeros(K * p, np.int32), np.zeros(K * p, dt)
            getattr(L, "oracle_csr_to_ell_" + s)(I64(rows), _p(A["row_offsets"]), _p(A["column_indices"]),
                                                 _p(A["values"]), I64(K), I64(p), _p(cidx), _p(ev))
            # csr_to_other.h:186: num_entries = src.num_entries - count(values == 0)
            ne = len(A["values"]) - int(np.count_nonzero(A["values"] == 0))
            return dict(format="ell", num_rows=rows, num_cols=cols, num_entries=ne, num_cols_per_row=K,
                        pitch=p, column_indices=cidx, values=ev)
        if fmt == "hyb":
            K = num_entries_per_row or optimal_entries_per_row(A["row_offsets"])
            p = round_up(rows, alignment)
            ecidx, ev = np.zeros(K * p, np.int32), np.zeros(K * p, dt)
            n = getattr(L, "oracle_csr_to_hyb_" + s)(I64(rows), _p(A["row_offsets"]), _p(A["column_indices"]),
                                                     _p(A["values"]), I64(K), I64(p), _p(ecidx), _p(ev), None,
                                                     None, None)
            ci, cj, cv = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, dt)
            getattr(L, "oracle_csr_to_hyb_" + s)(I64(rows), _p(A["row_offsets"]), _p(A["column_indices"]),
                                                 _p(A["values"]), I64(K), I64(p), _p(ecidx), _p(ev), _p(ci),
                                                 _p(cj), _p(cv))
            ell = dict(format="ell", num_rows=rows, num_cols=cols, num_entries=len(A["values"]) - int(n),
                       num_cols_per_row=K, pitch=p, column_indices=ecidx, values=ev)
            coo = dict(format="coo", num_rows=rows, num_cols=cols, num_entries=int(n), row_indices=ci,
                       column_indices=cj, values=cv)
            return dict(format="hyb", num_rows=rows, num_cols=cols, num_entries=len(A["values"]), ell=ell,
                        coo=coo)
        if fmt == "dia":
            p = round_up(rows, alignment)
            nd = getattr(L, "oracle_csr_to_dia_" + s)(I64(rows), I64(cols), _p(A["row_offsets"]),
                                                      _p(A["column_indices"]), _p(A["values"]), I64(p), None,
                                                      None)
            offs, vals = np.zeros(nd, np.int32), np.zeros(nd * p, dt)
            getattr(L, "oracle_csr_to_dia_" + s)(I64(rows), I64(cols), _p(A["row_offsets"]),
                                                 _p(A["column_indices"]), _p(A["values"]), I64(p), _p(offs),
                                                 _p(vals))
            return dict(format="dia", num_rows=rows, num_cols=cols, num_entries=len(A["values"]),
                        diagonal_offsets=offs, pitch=p, values=vals)
    if src == "ell":   # ell_to_other.h:55-143 (keeps value != 0); to HYB: the ELL part is the matrix itself (:145-163)
        rows, cols, K, p = A["num_rows"], A["num_cols"], A["num_cols_per_row"], A["pitch"]
        if fmt == "hyb":
            coo = dict(format="coo", num_rows=rows, num_cols=cols, num_entries=0, row_indices=np.zeros(0, np.int32),
                       column_indices=np.zeros(0, np.int32), values=np.zeros(0, dt))
            return dict(format="hyb", num_rows=rows, num_cols=cols, num_entries=A["num_entries"], ell=dict(A), coo=coo)
        args = (I64(rows), I64(K), I64(p), _p(A["column_indices"]), _p(A["values"]))
        n = getattr(L, "oracle_ell_to_coo_" + s)(*args, None, None, None)
        Ai, Aj, Ax = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, dt)
        getattr(L, "oracle_ell_to_coo_" + s)(*args, _p(Ai), _p(Aj), _p(Ax))
        coo = dict(format="coo", num_rows=rows, num_cols=cols, num_entries=int(n), row_indices=Ai, column_indices=Aj,
                   values=Ax)
        return coo if fmt == "coo" else convert(coo, fmt, alignment, num_entries_per_row)
    if src == "hyb":   # hyb_to_other.h:45-56 + cusp/detail/coo_matrix.inl:269-341 (merge by (row, col), keeps valid columns)
        e, c = A["ell"], A["coo"]
        rows, cols = A["num_rows"], A["num_cols"]
        args = (I64(rows), I64(e["num_cols_per_row"]), I64(e["pitch"]), _p(e["column_indices"]), _p(e["values"]),
                I64(len(c["values"])), _p(c["row_indices"]), _p(c["column_indices"]), _p(c["values"]))
        n = getattr(L, "oracle_hyb_to_coo_" + s)(*args, None, None, None)
        Ai, Aj, Ax = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, dt)
        getattr(L, "oracle_hyb_to_coo_" + s)(*args, _p(Ai), _p(Aj), _p(Ax))
        coo = dict(format="coo", num_rows=rows, num_cols=cols, num_entries=int(n), row_indices=Ai, column_indices=Aj,
                   values=Ax)
        return coo if fmt == "coo" else convert(coo, fmt, alignment, num_entries_per_row)
    raise ValueError(f"{src} -> {fmt}")


def ell_row_lengths(A: dict):
    """cusp/ktt/detail/ellr_matrix.inl:16-52"""
    p, K, rows = A["pitch"], A["num_cols_per_row"], A["num_rows"]
    if K == 0:
        return np.zeros(rows, np.int32)
    c = A["column_indices"].reshape(K, p)[:, :rows]
    neg = c < 0
    first_neg = np.where(neg.any(axis=0), neg.argmax(axis=0), K)
    return first_neg.astype(np.int32)


def to_ellr(A: dict) -> dict:
    B = dict(A)
    B["format"] = "ellr"
    B["row_lengths"] = ell_row_lengths(A)
    return B
