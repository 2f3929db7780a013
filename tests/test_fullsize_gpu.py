"""GPU: BASELINE.json's full sizes, checked through size-independent properties
(the oracle would need seconds-to-minutes per case at these sizes, so direct
comparison is kept to one sub-sampled case)."""
import numpy as np
import pytest
import torch

import cusp_autotuned_b200 as cusp
from cusp_autotuned_b200 import capi, convert, gallery
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _boundary_count_7pt(n, dev):
    """(A 1)_i for poisson7pt = 6 - #neighbours = number of faces of cell i on the boundary"""
    i = torch.arange(n, device=dev)
    f = ((i == 0) | (i == n - 1)).to(torch.float64) if n > 1 else torch.full((1,), 2.0, device=dev)
    if n > 1:
        f = (i == 0).to(torch.float64) + (i == n - 1).to(torch.float64)
    fx = f.view(1, 1, n).expand(n, n, n)
    fy = f.view(1, n, 1).expand(n, n, n)
    fz = f.view(n, 1, 1).expand(n, n, n)
    return (fx + fy + fz).reshape(-1)


@pytest.mark.parametrize("tdt", [torch.float32, torch.float64])
def test_poisson7pt_256_all_formats_agree_bitwise(tdt, dev, handle):
    """configs[1]: 16.7M rows.  DIA, ELL (both kernels), stream CSR and scalar CSR keep
    the reference's per-row order -> identical bits; A*1 has a closed form."""
    n = 256
    N = n ** 3
    ones = torch.ones(N, dtype=tdt, device=dev)
    x = ((torch.arange(N, device=dev) % 21) - 10).to(tdt)          # integer-valued: exact
    xr = (torch.rand(N, device=dev, dtype=torch.float64) + 0.5).to(tdt)  # generic data
    want_ones = _boundary_count_7pt(n, dev).to(tdt)
    outs = {}
    for fmt, cfgs in (("dia", [capi.Cfg(kernel=1), capi.Cfg(kernel=2)]),
                      ("ell", [capi.Cfg(kernel=1), capi.Cfg(kernel=2)]),
                      ("csr", [capi.Cfg(kernel=capi.K_CSR_STREAM), capi.Cfg(kernel=capi.K_CSR_RING),
                               capi.Cfg(kernel=capi.K_CSR_VECTOR, threads_per_row=8),
                               capi.Cfg(kernel=capi.K_CSR_VECTOR, threads_per_row=1)])):
        A = gallery.poisson(fmt, 7, (n, n, n), dtype=tdt)
        assert A.num_entries == 117047296
        for ci, cfg in enumerate(cfgs):
            y = torch.empty(N, dtype=tdt, device=dev)
            cusp.multiply(A, ones, y, cfg=cfg)
            assert torch.equal(y, want_ones), (fmt, ci)
            cusp.multiply(A, x, y, cfg=cfg)
            outs[(fmt, ci, "int")] = y.clone()
            cusp.multiply(A, xr, y, cfg=cfg)
            outs[(fmt, ci, "real")] = y.clone()
        del A
    base_i, base_r = outs[("dia", 0, "int")], outs[("dia", 0, "real")]
    for k, v in outs.items():
        if k[2] == "int":
            assert torch.equal(v, base_i), k
        elif k[:2] != ("csr", 2):  # every order-preserving kernel: same bits on real data too
            assert torch.equal(v, base_r), k
    # linearity on exactly representable data: A(2x + 1) == 2 A x + A 1
    A = gallery.poisson("dia", 7, (n, n, n), dtype=tdt)
    y = torch.empty(N, dtype=tdt, device=dev)
    cusp.multiply(A, 2 * x + ones, y, cfg=capi.Cfg())
    assert torch.equal(y, 2 * base_i + want_ones)
    # a leading slab against the oracle itself (rows [0, 2^20))
    m = 1 << 20
    Ah = O.poisson(7, (n, n, 17), np.float32 if tdt == torch.float32 else np.float64, "dia")
    xs = xr[: Ah["num_cols"]].cpu().numpy()
    want = O.spmv(Ah, xs)[:m]
    # rows < 2^20 of the 256^3 operator touch columns < 2^20 + 65536 only -> same rows as the 256x256x17 operator
    assert np.array_equal(base_r[:m].cpu().numpy(), want)


def test_poisson5pt_512_csr_fp64_vs_oracle(dev, handle):
    """configs[0]: the reference's own CPU-runnable case, compared entry by entry"""
    A = gallery.poisson5pt(512, 512, fmt="csr", dtype=torch.float64)
    assert A.num_rows == 262144 and A.num_entries == 1308672
    Ah = O.poisson(5, (512, 512), np.float64, "csr")
    assert np.array_equal(A.column_indices.cpu().numpy(), Ah["column_indices"])
    rng = np.random.default_rng(1234)
    x = rng.uniform(0.5, 1.5, 262144)
    xd = torch.from_numpy(x).to(dev)
    y = torch.empty(262144, dtype=torch.float64, device=dev)
    cusp.multiply(A, xd, y, cfg=capi.Cfg(kernel=capi.K_CSR_VECTOR, threads_per_row=1))
    assert np.array_equal(y.cpu().numpy(), O.spmv(Ah, x))           # scalar kernel: bit-exact
    cusp.multiply(A, xd, y)                                          # default (stream) kernel
    assert np.array_equal(y.cpu().numpy(), O.spmv(Ah, x))           # keeps the order too: bit-exact
    cusp.multiply(A, xd, y, cfg=capi.Cfg(kernel=capi.K_CSR_VECTOR, threads_per_row=8))  # sub-warp kernel
    scale = O.spmv(dict(Ah, values=np.abs(Ah["values"])), x)
    assert np.max(np.abs(y.cpu().numpy() - O.spmv(Ah, x)) / scale) <= 1e-12


def test_rmat_scale20_coo_hyb_vs_oracle(dev, handle):
    """configs[2] at scale 20 (the oracle finishes in seconds); all-ones values give
    exact integer row degrees, uniform values are checked at 1e-5 relative against an
    fp64-accumulated oracle (the fp32 sequential reference itself carries ~sqrt(n)*eps
    on hub rows)"""
    coo = convert.rmat(20, 16, seed=42, values="ones")
    n = coo.num_rows
    deg = torch.bincount(coo.row_indices.to(torch.int64), minlength=n).to(torch.float32)
    x1 = torch.ones(n, dtype=torch.float32, device=dev)
    y = torch.empty(n, dtype=torch.float32, device=dev)
    cusp.multiply(coo, x1, y)
    assert torch.equal(y, deg)
    hyb = convert.csr_to_hyb(convert.coo_to_csr(coo))
    cusp.multiply(hyb, x1, y)
    assert torch.equal(y, deg)
    assert hyb.ell.num_entries + hyb.coo.num_entries == coo.num_entries
    # uniform values
    coo = convert.rmat(20, 16, seed=42, values="uniform")
    x = torch.rand(n, device=dev) + 0.5
    cusp.multiply(coo, x, y)
    Ah = dict(format="coo", num_rows=n, num_cols=n, num_entries=coo.num_entries,
              row_indices=coo.row_indices.cpu().numpy(), column_indices=coo.column_indices.cpu().numpy(),
              values=coo.values.cpu().numpy().astype(np.float64))
    want64 = O.spmv(Ah, x.cpu().numpy().astype(np.float64))
    got = y.cpu().numpy().astype(np.float64)
    nzr = want64 > 0
    assert np.max(np.abs(got[nzr] - want64[nzr]) / want64[nzr]) <= 1e-5
    assert np.all(got[~nzr] == 0)
    hyb = convert.csr_to_hyb(convert.coo_to_csr(coo))
    cusp.multiply(hyb, x, y)
    got = y.cpu().numpy().astype(np.float64)
    assert np.max(np.abs(got[nzr] - want64[nzr]) / want64[nzr]) <= 1e-5


def test_cg_512cubed_residual_history_properties(dev, handle):
    """configs[4] on one GPU: residuals decrease monotonically for this SPD operator
    and the true residual ||b - A x|| matches the recurrence"""
    n = 512
    A = gallery.poisson("dia", 7, (n, n, n), dtype=torch.float64)
    assert A.num_entries == 937951232
    N = n ** 3
    b = torch.ones(N, dtype=torch.float64, device=dev)
    x = torch.zeros(N, dtype=torch.float64, device=dev)
    mon = cusp.monitor(None, 12, 0.0)
    cusp.krylov.cg(A, x, b, mon, check_interval=12)
    assert mon.iteration_count() == 12 and len(mon.residuals) == 13
    assert abs(mon.residuals[0] - np.sqrt(N)) <= 1e-9 * np.sqrt(N)
    r = torch.empty(N, dtype=torch.float64, device=dev)
    cusp.multiply(A, x, r, cfg=capi.Cfg())
    cusp.blas.axpby(b, r, r, 1.0, -1.0)
    true = cusp.blas.nrm2(r)
    assert abs(true - mon.residuals[-1]) <= 1e-8 * mon.residuals[0]


@pytest.mark.gpu
@pytest.mark.parametrize("tdt", (torch.float64, torch.float32))
def test_host_buffer_pipeline_is_bit_identical_to_the_device_product(tdt, dev, handle):
    """b200sp_spmv_host on a banded (DIA) operator runs as a chunked H2D / SpMV / D2H pipeline
    (csrc/api.cu); every row is still computed by the same kernel in the same order, so the
    host-side result equals the device product bit for bit — on generic data, for square and
    rectangular operators, and after the staging buffers have been reused"""
    import os
    n = 160  # 4 096 000 rows: 16 chunks of 256 000 rows, band 25 600
    A = gallery.poisson("dia", 7, (n, n, n), dtype=tdt)
    N = A.num_rows
    g = torch.Generator(device="cpu").manual_seed(5)
    for rep in range(2):
        xh = (torch.rand(N, generator=g, dtype=torch.float64) - 0.5).to(tdt).pin_memory()
        yh = torch.full((N,), 7.0, dtype=tdt).pin_memory()
        handle.spmv_host(A.descriptor(), xh, yh)
        y = torch.empty(N, dtype=tdt, device=dev)
        cusp.multiply(A, xh.to(dev), y)
        assert torch.equal(yh, y.cpu()), rep
    # the one-shot path (B200SP_HOST_ONE_SHOT) gives the same bits
    os.environ["B200SP_HOST_ONE_SHOT"] = "1"
    try:
        y2 = torch.empty(N, dtype=tdt).pin_memory()
        handle.spmv_host(A.descriptor(), xh, y2)
    finally:
        del os.environ["B200SP_HOST_ONE_SHOT"]
    assert torch.equal(y2, yh)
    # rectangular: more columns than rows (the last x piece runs to num_cols), offsets reaching past the rows
    rows, cols = 3_000_000, 3_200_000
    offs = torch.tensor([-70_000, -3, 0, 5, 150_000], dtype=torch.int32, device=dev)
    vals = (torch.rand(5 * rows, generator=torch.Generator(device=dev).manual_seed(3), device=dev,
                       dtype=torch.float64) + 0.5).to(tdt)
    from cusp_autotuned_b200.matrix import dia_matrix
    B = dia_matrix(rows, cols, 5 * rows, offs, rows, vals)
    xh = (torch.rand(cols, generator=g, dtype=torch.float64) - 0.5).to(tdt).pin_memory()
    yh = torch.empty(rows, dtype=tdt).pin_memory()
    handle.spmv_host(B.descriptor(), xh, yh)
    y = torch.empty(rows, dtype=tdt, device=dev)
    cusp.multiply(B, xh.to(dev), y)
    assert torch.equal(yh, y.cpu())
