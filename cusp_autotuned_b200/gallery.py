"""cusp::gallery::poisson5pt / poisson7pt (cusp/gallery/detail/poisson.inl:29-96)
built directly on the device in the target format (b200sp_poisson_*), bit-identical
to the reference's DIA-then-convert pipeline."""
from __future__ import annotations

import torch

from . import capi
from .matrix import coo_matrix, csr_matrix, default_handle, dia_matrix, ell_matrix


def _grid(stencil, dims):
    if stencil == 5:
        nx, ny = dims
        return int(nx), int(ny), 1
    nx, ny, nz = dims
    return int(nx), int(ny), int(nz)


def poisson(fmt: str, stencil: int, dims, dtype=torch.float64, device=None, row_begin=0, num_rows=None,
            halo_lo=0, halo_hi=0, handle=None):
    """Rows [row_begin, row_begin+num_rows) of the stencil operator in `fmt`
    ("dia" | "ell" | "csr" | "coo").  With halos, column indices are relative to the
    window [halo_lo | local | halo_hi] (row-partitioned operators)."""
    h = handle or default_handle()
    device = device or torch.device("cuda", torch.cuda.current_device())
    nx, ny, nz = _grid(stencil, dims)
    total = nx * ny * nz
    if num_rows is None:
        num_rows = total - row_begin
    partitioned = not (row_begin == 0 and num_rows == total and halo_lo == 0 and halo_hi == 0)
    num_cols = num_rows + halo_lo + halo_hi if partitioned else total
    col_shift = row_begin - halo_lo
    nnz = capi.poisson_num_entries(stencil, nx, ny, nz, row_begin, num_rows)
    if fmt == "dia":
        pitch = num_rows  # generate_matrix_from_stencil: 4-arg resize, pitch == num_rows
        offs = torch.empty(stencil, dtype=torch.int32, device=device)
        vals = torch.empty(stencil * pitch, dtype=dtype, device=device)
        h.poisson_dia(stencil, nx, ny, nz, row_begin, num_rows, col_shift, pitch, offs, vals)
        return dia_matrix(num_rows, num_cols, nnz, offs, pitch, vals)
    if fmt == "ell":
        pitch = num_rows  # DIA->ELL keeps the DIA pitch (dia_to_other.h:196)
        cidx = torch.empty(stencil * pitch, dtype=torch.int32, device=device)
        vals = torch.empty(stencil * pitch, dtype=dtype, device=device)
        h.poisson_ell(stencil, nx, ny, nz, row_begin, num_rows, col_shift, pitch, cidx, vals)
        return ell_matrix(num_rows, num_cols, nnz, stencil, pitch, cidx, vals)
    if fmt == "coo":  # DIA -> COO: same entries and order as CSR (dia_to_other.h:61-161), rows expanded
        A = poisson("csr", stencil, dims, dtype=dtype, device=device, row_begin=row_begin, num_rows=num_rows,
                    halo_lo=halo_lo, halo_hi=halo_hi, handle=h)
        ri = torch.empty(max(nnz, 1), dtype=torch.int32, device=device)[:nnz]
        h.offsets_to_indices(A.row_offsets, ri)
        return coo_matrix(A.num_rows, A.num_cols, ri, A.column_indices, A.values)
    if fmt == "csr":
        Ap = torch.empty(num_rows + 1, dtype=torch.int32, device=device)
        Aj = torch.empty(max(nnz, 1), dtype=torch.int32, device=device)[:nnz]
        Ax = torch.empty(max(nnz, 1), dtype=dtype, device=device)[:nnz]
        h.poisson_csr_offsets(stencil, nx, ny, nz, row_begin, num_rows, Ap)
        h.poisson_csr(stencil, nx, ny, nz, row_begin, num_rows, col_shift, Ap, Aj, Ax)
        return csr_matrix(num_rows, num_cols, Ap, Aj, Ax)
    raise capi.InvalidInput(capi.ST_INVALID_INPUT, f"gallery: unknown format {fmt!r}")


def poisson5pt(m, n, fmt="csr", dtype=torch.float64, **kw):
    return poisson(fmt, 5, (m, n), dtype=dtype, **kw)


def poisson7pt(m, n, k, fmt="csr", dtype=torch.float64, **kw):
    return poisson(fmt, 7, (m, n, k), dtype=dtype, **kw)


def from_host(fmt: str, arrays: dict, device=None):
    """upload numpy arrays laid out in the reference's format to a device container
    (cusp::<fmt>_matrix<int,V,device_memory> A_d = A_h)"""
    device = device or torch.device("cuda", torch.cuda.current_device())
    t = lambda a: torch.from_numpy(a).to(device)
    if fmt == "csr":
        return csr_matrix(arrays["num_rows"], arrays["num_cols"], t(arrays["row_offsets"]),
                          t(arrays["column_indices"]), t(arrays["values"]))
    if fmt == "coo":
        return coo_matrix(arrays["num_rows"], arrays["num_cols"], t(arrays["row_indices"]),
                          t(arrays["column_indices"]), t(arrays["values"]))
    if fmt == "ell":
        return ell_matrix(arrays["num_rows"], arrays["num_cols"], arrays["num_entries"],
                          arrays["num_cols_per_row"], arrays["pitch"], t(arrays["column_indices"]),
                          t(arrays["values"]))
    if fmt == "dia":
        return dia_matrix(arrays["num_rows"], arrays["num_cols"], arrays["num_entries"],
                          t(arrays["diagonal_offsets"]), arrays["pitch"], t(arrays["values"]))
    if fmt == "hyb":
        from .matrix import hyb_matrix
        return hyb_matrix(from_host("ell", arrays["ell"], device), from_host("coo", arrays["coo"], device))
    raise capi.InvalidInput(capi.ST_INVALID_INPUT, f"unknown format {fmt!r}")


def poisson7pt_benchmark_product(dims, row_begin, num_rows, dtype, device=None):
    """y = A x of the 7-point Poisson operator on grid `dims` for the benchmark vector x_j = (j mod 21) - 10
    (performance/spmv/benchmark.h:27-46), rows [row_begin, row_begin + num_rows), from index arithmetic alone: no
    matrix, no engine call.  Integer-valued, hence exact in fp32 / fp64 in any summation order — the independent
    check bench.py and tools/dist_check.py hold every timed product against.
    gallery/detail/poisson.inl:75-96: 6 on the diagonal, -1 per neighbour inside the grid, dimension 0 fastest."""
    device = device or torch.device("cuda", torch.cuda.current_device())
    nx, ny, nz = dims
    g = torch.arange(row_begin, row_begin + num_rows, device=device, dtype=torch.int64)
    xv = lambda j: (j % 21) - 10
    y = 6 * xv(g)
    ix = g % nx
    y -= torch.where(ix > 0, xv(g - 1), 0)
    y -= torch.where(ix < nx - 1, xv(g + 1), 0)
    del ix
    iy = (g // nx) % ny
    y -= torch.where(iy > 0, xv(g - nx), 0)
    y -= torch.where(iy < ny - 1, xv(g + nx), 0)
    del iy
    iz = g // (nx * ny)
    y -= torch.where(iz > 0, xv(g - nx * ny), 0)
    y -= torch.where(iz < nz - 1, xv(g + nx * ny), 0)
    return y.to(dtype)
