"""GPU parity: cusp::blas level-1 and cusp::krylov::cg vs the oracle."""
import numpy as np
import pytest
import torch

import cusp_autotuned_b200 as cusp
from cusp_autotuned_b200 import blas, capi
from golden import reference_fixtures as G
from helpers import TOL, rel_err, tdev, upload
from oracle import oracle as O

pytestmark = pytest.mark.gpu
DTYPES = [(np.float32, torch.float32), (np.float64, torch.float64)]


def test_blas_known_answers(dev):
    """testing/blas.cu:60-142, 287-352, 434-453 (exact, fp32)"""
    t = lambda a: torch.tensor(a, dtype=torch.float32, device=dev)
    a = G.BLAS_AXPBY
    z = torch.zeros(4, dtype=torch.float32, device=dev)
    blas.axpby(t(a["x"]), t(a["y"]), z, a["alpha"], a["beta"])
    assert z.tolist() == a["z"]
    a = G.BLAS_AXPY
    y = t(a["y"])
    blas.axpy(t(a["x"]), y, a["alpha"])
    assert y.tolist() == a["out"]
    a = G.BLAS_DOT
    assert blas.dot(t(a["x"]), t(a["y"])) == a["result"]
    assert blas.dotc(t(a["x"]), t(a["y"])) == a["result"]
    assert blas.nrm2(t(G.BLAS_NRM2["x"])) == G.BLAS_NRM2["result"]
    # size checking -> invalid_input_exception
    w = torch.zeros(3, dtype=torch.float32, device=dev)
    with pytest.raises(cusp.InvalidInput):
        blas.axpy(t(a["x"]), w, 1.0)
    with pytest.raises(cusp.InvalidInput):
        blas.axpby(t(a["x"]), t(a["y"]), w, 2.0, 1.0)
    with pytest.raises(cusp.InvalidInput):
        blas.dot(t(a["x"]), w)
    with pytest.raises(cusp.InvalidInput):
        blas.copy(w, t(a["x"]))


@pytest.mark.parametrize("ndt,tdt", DTYPES)
@pytest.mark.parametrize("n", [0, 1, 31, 1024, 1025, 100003, 3_000_001])
def test_blas_elementwise_bit_exact(n, ndt, tdt, dev):
    """alpha*x + y written as the reference functors write it, no FMA: bit-identical"""
    rng = np.random.default_rng(n + 1)
    x = rng.uniform(-2, 2, n).astype(ndt)
    y = rng.uniform(-2, 2, n).astype(ndt)
    alpha, beta = ndt(0.7371), ndt(-1.3113)
    yd = tdev(y, dev)
    blas.axpy(tdev(x, dev), yd, float(alpha))
    assert np.array_equal(yd.cpu().numpy(), O.axpy(x, y, alpha))
    zd = torch.empty(n, dtype=tdt, device=dev)
    blas.axpby(tdev(x, dev), tdev(y, dev), zd, float(alpha), float(beta))
    assert np.array_equal(zd.cpu().numpy(), O.axpby(x, y, alpha, beta))
    blas.copy(tdev(x, dev), zd)
    assert np.array_equal(zd.cpu().numpy(), x)
    blas.scal(zd, float(alpha))
    assert np.array_equal(zd.cpu().numpy(), alpha * x)
    blas.fill(zd, 2.5)
    assert np.array_equal(zd.cpu().numpy(), np.full(n, 2.5, ndt))


@pytest.mark.parametrize("ndt,tdt", DTYPES)
@pytest.mark.parametrize("n", [0, 1, 255, 4096, 100003, 3_000_001])
def test_blas_reductions(n, ndt, tdt, dev):
    rng = np.random.default_rng(n + 7)
    # integer-valued data: every summation order is exact -> equality
    xi = rng.integers(-3, 4, n).astype(ndt)
    yi = rng.integers(-3, 4, n).astype(ndt)
    assert blas.dot(tdev(xi, dev), tdev(yi, dev)) == float(np.dot(xi.astype(np.float64), yi.astype(np.float64)))
    # positive data: per north star 1e-5 / 1e-12 relative to the sequential reference
    x = rng.uniform(0.5, 1.5, n).astype(ndt)
    y = rng.uniform(0.5, 1.5, n).astype(ndt)
    tol = TOL[np.dtype(ndt)]
    d = blas.dot(tdev(x, dev), tdev(y, dev))
    exact = float(np.dot(x.astype(np.float64), y.astype(np.float64)))
    if n:
        # the fp32 sequential reference itself drifts by ~n*eps on long sums: compare
        # both against the exactly accumulated value, the GPU tree must not be worse
        ref_err = abs(float(O.dot(x, y)) - exact) / exact
        assert abs(d - exact) / exact <= max(tol, ref_err)
        nr = blas.nrm2(tdev(x, dev))
        exact_n = float(np.sqrt(np.dot(x.astype(np.float64), x.astype(np.float64))))
        assert abs(nr - exact_n) / exact_n <= max(tol, abs(float(O.nrm2(x)) - exact_n) / exact_n)
    else:
        assert d == 0.0 and blas.nrm2(tdev(x, dev)) == 0.0
    # deterministic: bit-identical on repetition
    assert blas.dot(tdev(x, dev), tdev(y, dev)) == d


@pytest.mark.parametrize("ndt,tdt", DTYPES)
@pytest.mark.parametrize("fmt", ["csr", "dia", "ell", "coo", "hyb"])
def test_cg_reference_case(fmt, ndt, tdt, dev):
    """testing/cg.cu:46-72: poisson5pt(10,10), b = 1, monitor(b, 20, 1e-4)"""
    c = G.CG_CASE
    A = O.poisson(5, c["grid"], ndt, fmt)
    Ad = upload(fmt, A, dev)
    b = torch.ones(A["num_rows"], dtype=tdt, device=dev)
    x = torch.zeros_like(b)
    mon = cusp.monitor(b, c["limit"], c["rel"])
    cusp.krylov.cg(Ad, x, b, mon)
    r = torch.zeros_like(b)
    cusp.multiply(Ad, x, r, cfg=capi.Cfg())
    blas.axpby(r, b, r, -1.0, 1.0)
    assert blas.nrm2(r) < 1e-4 * blas.nrm2(b)
    assert mon.converged() and mon.iteration_count() <= c["limit"]
    # same iterate sequence as the reference algorithm (oracle on CSR)
    csr = O.poisson(5, c["grid"], ndt, "csr")
    xo, it, conv, hist = O.cg(csr, np.zeros(csr["num_rows"], ndt), np.ones(csr["num_rows"], ndt), c["limit"], c["rel"])
    assert mon.iteration_count() == it and len(mon.residuals) == len(hist)
    assert np.allclose(mon.residuals, hist, rtol=1e-4 if ndt == np.float32 else 1e-10)
    assert np.allclose(x.cpu().numpy(), xo, rtol=1e-4 if ndt == np.float32 else 1e-10, atol=0)


def test_cg_zero_residual(dev):
    """testing/cg.cu:75-99"""
    A = upload("csr", O.convert(O.dense_to_coo(np.array([[8, 0], [0, 4]], np.float32)), "csr"), dev)
    x = torch.ones(2, dtype=torch.float32, device=dev)
    b = torch.zeros(2, dtype=torch.float32, device=dev)
    cusp.multiply(A, x, b)
    mon = cusp.monitor(b, 20, 0.0)
    cusp.krylov.cg(A, x, b, mon)
    assert mon.converged() and mon.iteration_count() == 0
    assert x.tolist() == [1.0, 1.0] and mon.residual_norm() == 0.0


@pytest.mark.parametrize("check_interval", [1, 3, 16])
def test_cg_iteration_limit_and_history(check_interval, dev):
    """limit reached before convergence: count == limit, one residual per finished() call,
    independent of how often the host polls the device flag"""
    A = O.poisson(7, (12, 11, 10), np.float64, "csr")
    Ad = upload("dia", O.poisson(7, (12, 11, 10), np.float64, "dia"), dev)
    b = np.random.default_rng(3).uniform(-1, 1, A["num_rows"])
    xo, it, conv, hist = O.cg(A, np.zeros_like(b), b, 7, 1e-14)
    x = torch.zeros(A["num_rows"], dtype=torch.float64, device=dev)
    mon = cusp.monitor(None, 7, 1e-14)
    cusp.krylov.cg(Ad, x, tdev(b, dev), mon, check_interval=check_interval)
    assert it == 7 and not conv
    assert mon.iteration_count() == 7 and not mon.converged() and len(mon.residuals) == 8
    assert np.allclose(mon.residuals, hist, rtol=1e-10)
    assert np.allclose(x.cpu().numpy(), xo, rtol=1e-9, atol=1e-14)


def test_cg_size_mismatch(dev):
    A = upload("csr", O.poisson(5, (4, 4), np.float32, "csr"), dev)
    with pytest.raises(cusp.InvalidInput):
        cusp.krylov.cg(A, torch.zeros(15, device=dev), torch.zeros(16, device=dev))


def test_captured_products_equal_plain_calls(dev, handle):
    """b200sp_spmv_graph_create / b200sp_graph_launch: `count` products replayed from one CUDA graph give the bits of
    `count` plain calls — assign and accumulate forms, CSR / DIA / COO, torch's (legacy default) stream as caller"""
    A = O.poisson(5, (64, 48), np.float64, "coo")
    n = A["num_rows"]
    rng = np.random.default_rng(9)
    x = tdev(rng.uniform(-1, 1, n), dev)
    for fmt in ("csr", "dia", "coo", "ell", "hyb"):
        Ad = upload(fmt, O.convert(A, fmt), dev)
        d = Ad.descriptor()
        y_plain = torch.zeros(n, dtype=torch.float64, device=dev)
        handle.spmv(d, x, y_plain)
        y_g = torch.full((n,), 7.0, dtype=torch.float64, device=dev)
        g = handle.spmv_graph_create(d, x, y_g, 5)
        l0 = handle.launch_count
        handle.graph_launch(g)
        assert handle.launch_count > l0
        torch.cuda.synchronize()
        assert torch.equal(y_g, y_plain), fmt
        handle.graph_destroy(g)
        # accumulate: 3 replays of 4 captured products add 12 A x (plus the one plain product creation makes)
        y_acc = torch.zeros(n, dtype=torch.float64, device=dev)
        g = handle.spmv_graph_create(d, x, y_acc, 4, accumulate=True)
        for _ in range(3):
            handle.graph_launch(g)
        want = torch.zeros(n, dtype=torch.float64, device=dev)
        for _ in range(13):
            handle.spmv(d, x, want, accumulate=True)
        torch.cuda.synchronize()
        assert torch.equal(y_acc, want), fmt
        handle.graph_destroy(g)


@pytest.mark.parametrize("fmt", ["csr", "dia", "coo"])
def test_cg_graph_replay_is_the_same_solve(fmt, dev, monkeypatch):
    """small systems: b200sp_cg replays check_interval iterations from one CUDA graph (B200SP_CG_GRAPH) — iteration
    count, history and solution are those of the launch-by-launch solve, bit for bit"""
    A = O.poisson(5, (40, 33), np.float64, fmt)
    Ad = upload(fmt, A, dev)
    b = tdev(np.random.default_rng(2).uniform(-1, 1, A["num_rows"]), dev)
    out = {}
    monkeypatch.setenv("B200SP_CG_PERSISTENT", "0")  # (CSR would otherwise take the one-kernel solve in both modes)
    for mode in ("0", "1"):
        monkeypatch.setenv("B200SP_CG_GRAPH", mode)
        for ci in (1, 7, 16):
            x = torch.zeros_like(b)
            mon = cusp.monitor(b, 300, 1e-10)
            cusp.krylov.cg(Ad, x, b, mon, check_interval=ci)
            out[(mode, ci)] = (mon.iteration_count(), list(mon.residuals), x.clone())
    base = out[("0", 1)]
    assert base[0] > 20
    for k, v in out.items():
        assert v[0] == base[0] and v[1] == base[1] and torch.equal(v[2], base[2]), k


@pytest.mark.parametrize("ndt,tdt", [(np.float32, torch.float32), (np.float64, torch.float64)])
@pytest.mark.parametrize("grid", [(40, 33), (24, 20, 19), (64, 64, 16)])
def test_cg_two_kernel_iteration_is_the_same_solve(grid, ndt, tdt, dev, monkeypatch):
    """DIA through the bulk kernel: the direction update p = r + beta p is folded into the product (B200SP_CG_FUSE,
    csrc/cg.cu) — same expressions in the same order, so iteration count, history and solution equal the three-kernel
    solve bit for bit, launch by launch and from a replayed graph, whatever the poll interval (odd ones included: the
    graph holds an even number of iterations)"""
    A = O.poisson(5 if len(grid) == 2 else 7, grid, ndt, "dia")
    Ad = upload("dia", A, dev)
    b = tdev(np.random.default_rng(4).uniform(-1, 1, A["num_rows"]).astype(ndt), dev)
    out = {}
    for fuse in ("0", "1"):
        monkeypatch.setenv("B200SP_CG_FUSE", fuse)
        for graph in ("0", "1"):
            monkeypatch.setenv("B200SP_CG_GRAPH", graph)
            for ci in (1, 5, 16):
                x = torch.zeros_like(b)
                mon = cusp.monitor(b, 200, 1e-5 if ndt == np.float32 else 1e-10)
                cusp.krylov.cg(Ad, x, b, mon, check_interval=ci)
                out[(fuse, graph, ci)] = (mon.iteration_count(), list(mon.residuals), x.clone(), mon.converged())
    base = out[("0", "0", 1)]
    assert base[0] > 15 and base[3]
    for k, v in out.items():
        assert v[0] == base[0] and v[1] == base[1] and torch.equal(v[2], base[2]) and v[3] == base[3], k
    # and an iteration limit that stops the solve in mid-graph
    monkeypatch.setenv("B200SP_CG_GRAPH", "1")
    res = {}
    for fuse in ("0", "1"):
        monkeypatch.setenv("B200SP_CG_FUSE", fuse)
        x = torch.zeros_like(b)
        mon = cusp.monitor(b, 11, 0.0)
        cusp.krylov.cg(Ad, x, b, mon, check_interval=8)
        res[fuse] = (mon.iteration_count(), list(mon.residuals), x.clone())
    assert res["0"][0] == 11 and res["1"][0] == 11 and res["0"][1] == res["1"][1] and torch.equal(res["0"][2], res["1"][2])


@pytest.mark.parametrize("ndt,tdt", [(np.float32, torch.float32), (np.float64, torch.float64)])
@pytest.mark.parametrize("grid", [(10, 10), (37, 29), (512, 512), (20, 18, 16)])
def test_cg_persistent_kernel_is_the_same_iteration(grid, ndt, tdt, dev, monkeypatch):
    """small CSR systems on one GPU run as ONE persistent cooperative kernel (cg_small_csr_kernel: matrix slice and
    vectors in shared memory, two grid barriers per iteration).  Same expressions per element, dot products grouped by
    CTA instead of by 1024-element blocks: same iteration count, history and solution to rounding, the oracle's
    history to the usual bar; limit, poll interval and x0 != 0 behave like the ordinary path"""
    A = O.poisson(5 if len(grid) == 2 else 7, grid, ndt, "csr")
    n = A["num_rows"]
    Ad = upload("csr", A, dev)
    rng = np.random.default_rng(8)
    bn = rng.uniform(-1, 1, n).astype(ndt)
    x0 = rng.uniform(-1, 1, n).astype(ndt)
    b = tdev(bn, dev)
    rel = 1e-4 if ndt == np.float32 else 1e-9
    limit = 60 if n > 100000 else 400
    got = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("B200SP_CG_PERSISTENT", mode)
        x = tdev(x0, dev)
        mon = cusp.monitor(b, limit, rel)
        launches = cusp.default_handle().launch_count
        cusp.krylov.cg(Ad, x, b, mon, check_interval=7)
        got[mode] = (mon.iteration_count(), np.asarray(mon.residuals), x.cpu().numpy(), mon.converged(),
                     cusp.default_handle().launch_count - launches)
    a, p = got["0"], got["1"]
    assert p[4] <= 4, p[4]  # ||b||, the state set-up and ONE kernel for the whole solve
    assert a[4] >= 3 * a[0]
    assert p[0] == a[0] and p[3] == a[3] and len(p[1]) == len(a[1]), (p[0], a[0])
    tol = 1e-3 if ndt == np.float32 else 1e-9
    assert np.allclose(p[1], a[1], rtol=tol, atol=0), np.max(np.abs(p[1] - a[1]) / a[1])
    assert np.allclose(p[2], a[2], rtol=0, atol=(1e-3 if ndt == np.float32 else 1e-8) * np.abs(a[2]).max())
    xo, it, conv, hist = O.cg(A, x0, bn, limit, rel)
    if ndt == np.float64 and n < 100000:
        assert p[0] == it and np.allclose(p[1], hist, rtol=1e-9)


def test_cg_persistent_kernel_falls_back_when_a_slice_does_not_fit(dev, monkeypatch):
    """the host checks the mean slice, the kernel every slice: a block of heavy rows that overflows one CTA's shared
    memory raises the fallback flag and the solve takes the ordinary path — same bits as with the kernel turned off"""
    import scipy.sparse as sp
    n, heavy, width = 20000, 135, 2000
    rows = np.repeat(np.arange(heavy), width)
    cols = np.tile(np.arange(width), heavy)
    B = sp.coo_matrix((np.full(rows.size, 1e-3), (rows, cols)), shape=(n, n)).tocsr()
    M = (B + B.T + sp.identity(n) * 10.0).tocsr()
    M.sum_duplicates()
    M.sort_indices()
    A = {"format": "csr", "num_rows": n, "num_cols": n, "num_entries": int(M.nnz), "row_offsets": M.indptr.astype(np.int32),
         "column_indices": M.indices.astype(np.int32), "values": M.data.astype(np.float64)}
    Ad = upload("csr", A, dev)
    b = tdev(np.random.default_rng(1).uniform(-1, 1, n), dev)
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("B200SP_CG_PERSISTENT", mode)
        x = torch.zeros_like(b)
        mon = cusp.monitor(b, 50, 1e-10)
        cusp.krylov.cg(Ad, x, b, mon)
        out[mode] = (mon.iteration_count(), list(mon.residuals), x.clone(), mon.converged())
    assert out["0"][3] and out["0"][0] >= 2
    assert out["1"][0] == out["0"][0] and out["1"][1] == out["0"][1] and torch.equal(out["1"][2], out["0"][2])
