"""cusp::convert on the device, plus the R-MAT generator of the bench configurations.

  * rmat(): R-MAT power-law graph (SURVEY §8d cfg 3; not in the reference gallery), torch ops
  * coo_to_csr / csr_to_coo / csr_to_ell / csr_to_hyb / csr_to_dia / optimal_entries_per_row:
    the reference's conversion rules (cusp/system/detail/generic/conversions/csr_to_other.h:73-306,
    generic/format_utils.inl:36-110,281-321) through the engine's conversion kernels
    (csrc/convert.cu, b200sp_csr_to_* / b200sp_csr_convert_query).

Results are bit-identical to the oracle's restatement of the same rules
(tests/test_gallery_gpu.py, tests/test_convert_gpu.py)."""
from __future__ import annotations

import torch

from . import capi
from .matrix import coo_matrix, csr_matrix, default_handle, dia_matrix, ell_matrix, hyb_matrix


def round_up(n: int, k: int) -> int:
    return k * ((n + k - 1) // k)


def rmat(scale: int, edge_factor: int = 16, a=0.57, b=0.19, c=0.19, seed: int = 42, dtype=torch.float32,
         values: str = "uniform", device=None) -> coo_matrix:
    """2^scale vertices, edge_factor*2^scale sampled edges, self loops kept,
    duplicates removed, sorted by (row, col).  values: "uniform" U(0.5,1.5) or "ones"."""
    device = device or torch.device("cuda", torch.cuda.current_device())
    n = 1 << scale
    m = edge_factor * n
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    rows = torch.zeros(m, dtype=torch.int64, device=device)
    cols = torch.zeros(m, dtype=torch.int64, device=device)
    ab, abc = a + b, a + b + c
    for _ in range(scale):
        u = torch.rand(m, generator=g, device=device)
        rbit = (u >= ab).to(torch.int64)                     # quadrants c, d -> lower half
        cbit = (((u >= a) & (u < ab)) | (u >= abc)).to(torch.int64)  # quadrants b, d -> right half
        rows = (rows << 1) | rbit
        cols = (cols << 1) | cbit
        del u, rbit, cbit
    key = torch.unique(rows * n + cols)  # sorted, duplicates removed
    del rows, cols
    ri = (key // n).to(torch.int32)
    ci = (key % n).to(torch.int32)
    nnz = key.numel()
    del key
    if values == "ones":
        v = torch.ones(nnz, dtype=dtype, device=device)
    else:
        v = (torch.rand(nnz, generator=g, device=device, dtype=torch.float32) + 0.5).to(dtype)
    return coo_matrix(n, n, ri, ci, v)


def coo_to_csr(A: coo_matrix) -> csr_matrix:
    """indices_to_offsets (cusp/format_utils.h): offsets[i] = #entries with row < i (b200sp_indices_to_offsets)"""
    offs = torch.empty(A.num_rows + 1, dtype=torch.int32, device=A.values.device)
    default_handle().indices_to_offsets(A.row_indices, offs)
    return csr_matrix(A.num_rows, A.num_cols, offs, A.column_indices, A.values)


def csr_to_coo(A: csr_matrix) -> coo_matrix:
    """offsets_to_indices (b200sp_offsets_to_indices)"""
    ri = torch.empty(max(A.num_entries, 1), dtype=torch.int32, device=A.values.device)[:A.num_entries]
    default_handle().offsets_to_indices(A.row_offsets, ri)
    return coo_matrix(A.num_rows, A.num_cols, ri, A.column_indices, A.values)


def optimal_entries_per_row(row_offsets: torch.Tensor, relative_speed: float = 3.0,
                            breakeven_threshold: int = 4096) -> int:
    """compute_optimal_entries_per_row (generic/format_utils.inl:281-321) +
    speed_threshold_functor (cusp/detail/functional.inl:114-132), b200sp_csr_convert_query"""
    rows = row_offsets.numel() - 1
    if rows <= 0:
        return 0
    nnz = int(row_offsets[-1].item())
    info = default_handle().csr_convert_query(rows, rows, nnz, row_offsets, None, relative_speed, breakeven_threshold)
    return int(info.hyb_entries_per_row)


def csr_to_ell(A: csr_matrix, num_entries_per_row: int = 0, alignment: int = 32) -> ell_matrix:
    """csr_to_other.h:155-227 (b200sp_csr_to_ell)"""
    h = default_handle()
    dev = A.values.device
    K = num_entries_per_row
    if not K and A.num_entries:
        K = int(h.csr_convert_query(A.num_rows, A.num_cols, A.num_entries, A.row_offsets).max_entries_per_row)
    pitch = round_up(A.num_rows, alignment)
    cidx = torch.empty(max(K * pitch, 1), dtype=torch.int32, device=dev)[:K * pitch]
    vals = torch.empty(max(K * pitch, 1), dtype=A.values.dtype, device=dev)[:K * pitch]
    h.csr_to_ell(A.num_rows, K, pitch, A.row_offsets, A.column_indices, A.values, cidx, vals)
    ne = A.num_entries - h.count_zeros(A.values)
    return ell_matrix(A.num_rows, A.num_cols, ne, K, pitch, cidx, vals)


def csr_to_hyb(A: csr_matrix, num_entries_per_row: int | None = None, alignment: int = 32) -> hyb_matrix:
    """csr_to_other.h:229-306: first K entries of each row -> ELL, the rest -> COO in CSR order
    (b200sp_csr_convert_query + b200sp_csr_to_ell + b200sp_csr_to_coo_tail)"""
    h = default_handle()
    dev = A.values.device
    if num_entries_per_row is None:
        info = h.csr_convert_query(A.num_rows, A.num_cols, A.num_entries, A.row_offsets)
        K, tail = int(info.hyb_entries_per_row), int(info.hyb_coo_entries)
    else:
        K = int(num_entries_per_row)
        lens = (A.row_offsets[1:] - A.row_offsets[:-1]).to(torch.int64)
        tail = int(torch.clamp(lens - K, min=0).sum().item())
    pitch = round_up(A.num_rows, alignment)
    cidx = torch.empty(max(K * pitch, 1), dtype=torch.int32, device=dev)[:K * pitch]
    vals = torch.empty(max(K * pitch, 1), dtype=A.values.dtype, device=dev)[:K * pitch]
    h.csr_to_ell(A.num_rows, K, pitch, A.row_offsets, A.column_indices, A.values, cidx, vals)
    ri = torch.empty(max(tail, 1), dtype=torch.int32, device=dev)[:tail]
    ci = torch.empty(max(tail, 1), dtype=torch.int32, device=dev)[:tail]
    cv = torch.empty(max(tail, 1), dtype=A.values.dtype, device=dev)[:tail]
    h.csr_to_coo_tail(A.num_rows, K, A.row_offsets, A.column_indices, A.values, ri, ci, cv)
    coo = coo_matrix(A.num_rows, A.num_cols, ri, ci, cv)
    ell = ell_matrix(A.num_rows, A.num_cols, A.num_entries - tail, K, pitch, cidx, vals)
    return hyb_matrix(ell, coo)


def csr_to_dia(A: csr_matrix, alignment: int = 32) -> dia_matrix:
    """csr_to_other.h:73-153: occupied diagonals ascending, pitch = round_up(rows, alignment),
    refuses a fill-in > 3x on slabs > 1e6 slots (b200sp_csr_convert_query + b200sp_csr_to_dia)"""
    h = default_handle()
    dev = A.values.device
    info = h.csr_convert_query(A.num_rows, A.num_cols, A.num_entries, A.row_offsets, A.column_indices)
    nd = int(info.num_diagonals)
    size = float(nd) * float(A.num_rows)
    if 3.0 < size / max(1.0, float(A.num_entries)) and size > 1e6:
        raise capi.B200spError(capi.ST_INVALID_INPUT, "dia_matrix fill-in would exceed maximum tolerance")
    pitch = round_up(A.num_rows, alignment)
    offs = torch.empty(max(nd, 1), dtype=torch.int32, device=dev)[:nd]
    vals = torch.empty(max(nd * pitch, 1), dtype=A.values.dtype, device=dev)[:nd * pitch]
    h.csr_to_dia(A.num_rows, A.num_cols, nd, pitch, A.row_offsets, A.column_indices, A.values, offs, vals)
    return dia_matrix(A.num_rows, A.num_cols, A.num_entries, offs, pitch, vals)


def _empty(n, dtype, dev):
    return torch.empty(max(int(n), 1), dtype=dtype, device=dev)[:int(n)]


def dia_to_csr(A: dia_matrix) -> csr_matrix:
    """dia_to_other.h:110-161: row-major scan, value != 0 kept (b200sp_dia_to_csr_offsets / _fill)"""
    h, dev = default_handle(), A.values.device
    offs = torch.empty(A.num_rows + 1, dtype=torch.int32, device=dev)
    n = h.dia_to_csr_offsets(A.num_rows, A.num_diagonals, A.pitch, A.values, offs)
    cj, cv = _empty(n, torch.int32, dev), _empty(n, A.values.dtype, dev)
    h.dia_to_csr_fill(A.num_rows, A.num_diagonals, A.pitch, A.diagonal_offsets, A.values, offs, cj, cv)
    return csr_matrix(A.num_rows, A.num_cols, offs, cj, cv)


def ell_to_csr(A: ell_matrix) -> csr_matrix:
    """ell_to_other.h:100-143: row-major scan, value != 0 kept (b200sp_ell_to_csr_offsets / _fill)"""
    h, dev = default_handle(), A.values.device
    offs = torch.empty(A.num_rows + 1, dtype=torch.int32, device=dev)
    n = h.ell_to_csr_offsets(A.num_rows, A.num_cols_per_row, A.pitch, A.column_indices, A.values, offs)
    cj, cv = _empty(n, torch.int32, dev), _empty(n, A.values.dtype, dev)
    h.ell_to_csr_fill(A.num_rows, A.num_cols_per_row, A.pitch, A.column_indices, A.values, offs, cj, cv)
    return csr_matrix(A.num_rows, A.num_cols, offs, cj, cv)


def hyb_to_csr(A: hyb_matrix) -> csr_matrix:
    """hyb_to_other.h:45-56 + detail/coo_matrix.inl:269-341: ELL entries with a valid column merged with the COO
    entries by (row, column) (b200sp_hyb_to_csr_offsets / _fill)"""
    h, dev = default_handle(), A.ell.values.device
    e, c = A.ell, A.coo
    offs = torch.empty(A.num_rows + 1, dtype=torch.int32, device=dev)
    n = h.hyb_to_csr_offsets(A.num_rows, e.num_cols_per_row, e.pitch, e.column_indices, e.values, c.num_entries,
                             c.row_indices, offs)
    cj, cv = _empty(n, torch.int32, dev), _empty(n, e.values.dtype, dev)
    h.hyb_to_csr_fill(A.num_rows, e.num_cols_per_row, e.pitch, e.column_indices, e.values, c.num_entries, c.row_indices,
                      c.column_indices, c.values, offs, cj, cv)
    return csr_matrix(A.num_rows, A.num_cols, offs, cj, cv)


def dia_to_ell(A: dia_matrix) -> ell_matrix:
    """dia_to_other.h:163-251: K = #diagonals, pitch = the DIA pitch, non-zeros left-packed (b200sp_dia_to_ell)"""
    h, dev = default_handle(), A.values.device
    K, p = A.num_diagonals, A.pitch
    cidx, vals = _empty(K * p, torch.int32, dev), _empty(K * p, A.values.dtype, dev)
    h.dia_to_ell(A.num_rows, K, p, A.diagonal_offsets, A.values, cidx, vals)
    return ell_matrix(A.num_rows, A.num_cols, A.num_entries, K, p, cidx, vals)


def convert(A, fmt: str, **kw):
    """cusp::convert(src, dst) for device containers, every source x destination pair, on the device: the formats
    meet in CSR (the reference meets in COO/CSR too; the layouts are the rules of generic/conversions/*.h)."""
    src = {capi.FMT_CSR: "csr", capi.FMT_COO: "coo", capi.FMT_DIA: "dia", capi.FMT_ELL: "ell", capi.FMT_HYB: "hyb",
           capi.FMT_ELLR: "ell"}[A.format]
    if src == fmt:
        return A
    if src == "dia" and fmt == "ell":
        return dia_to_ell(A)
    if src == "ell" and fmt == "hyb":  # ell_to_other.h:145-163: the ELL part is the matrix, the COO part is empty
        dev = A.values.device
        z = lambda dt: torch.empty(1, dtype=dt, device=dev)[:0]
        return hyb_matrix(A, coo_matrix(A.num_rows, A.num_cols, z(torch.int32), z(torch.int32), z(A.values.dtype)))
    csr = {"csr": lambda M: M, "coo": coo_to_csr, "dia": dia_to_csr, "ell": ell_to_csr, "hyb": hyb_to_csr}[src](A)
    return {"csr": lambda M: M, "coo": csr_to_coo, "ell": csr_to_ell, "hyb": csr_to_hyb, "dia": csr_to_dia}[fmt](csr, **kw)
