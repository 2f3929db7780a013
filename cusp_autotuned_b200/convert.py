"""Device-side input construction for the bench configurations that the
reference builds with cusp::convert / has no generator for:

  * rmat(): R-MAT power-law graph (SURVEY §8d cfg 3; not in the reference gallery)
  * coo_to_csr / csr_to_hyb / csr_to_ell: the reference's conversion RULES
    (cusp/system/detail/generic/conversions/csr_to_other.h:155-306,
    generic/format_utils.inl:281-321) evaluated with torch tensor ops on the GPU.

This is input plumbing (sort / scan / scatter through PyTorch), not the SpMV hot
path; results are bit-identical to the oracle's restatement of the same rules
(tests/test_convert_gpu.py).  Hand-written conversion kernels are a "next" row
(SURVEY §8f-1)."""
from __future__ import annotations

import torch

from .matrix import coo_matrix, csr_matrix, ell_matrix, hyb_matrix


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
    """indices_to_offsets (cusp/format_utils.h): offsets[i] = #entries with row < i"""
    counts = torch.bincount(A.row_indices.to(torch.int64), minlength=A.num_rows)
    offs = torch.zeros(A.num_rows + 1, dtype=torch.int64, device=A.values.device)
    torch.cumsum(counts, 0, out=offs[1:])
    return csr_matrix(A.num_rows, A.num_cols, offs.to(torch.int32), A.column_indices, A.values)


def optimal_entries_per_row(row_offsets: torch.Tensor, relative_speed: float = 3.0,
                            breakeven_threshold: int = 4096) -> int:
    """compute_optimal_entries_per_row (generic/format_utils.inl:281-321) +
    speed_threshold_functor (cusp/detail/functional.inl:114-132)"""
    lens = (row_offsets[1:] - row_offsets[:-1]).to(torch.int64)
    num_rows = lens.numel()
    maxc = int(lens.max().item()) if num_rows else 0
    hist = torch.bincount(lens, minlength=maxc + 1)
    cum = torch.cumsum(hist, 0).cpu().tolist()  # cum[k] = #rows with length <= k
    import numpy as np
    for k in range(maxc):
        longer = num_rows - cum[k]
        # float32 arithmetic like the functor: relative_speed * (num_rows-rows) < num_rows
        if np.float32(relative_speed) * np.float32(longer) < np.float32(num_rows) or longer < breakeven_threshold:
            return k
    return maxc


def _slot_index(row_offsets: torch.Tensor, nnz: int):
    """position of every entry inside its row, and its row"""
    dev = row_offsets.device
    lens = (row_offsets[1:] - row_offsets[:-1]).to(torch.int64)
    rows = torch.repeat_interleave(torch.arange(lens.numel(), device=dev, dtype=torch.int64), lens)
    k = torch.arange(nnz, device=dev, dtype=torch.int64) - row_offsets.to(torch.int64)[rows]
    return rows, k


def csr_to_ell(A: csr_matrix, num_entries_per_row: int = 0, alignment: int = 32) -> ell_matrix:
    """csr_to_other.h:155-227"""
    dev = A.values.device
    rows, k = _slot_index(A.row_offsets, A.num_entries)
    K = num_entries_per_row or (int(k.max().item()) + 1 if A.num_entries else 0)
    pitch = round_up(A.num_rows, alignment)
    cidx = torch.full((K * pitch,), -1, dtype=torch.int32, device=dev)
    vals = torch.zeros(K * pitch, dtype=A.values.dtype, device=dev)
    keep = k < K
    pos = (k * pitch + rows)[keep]
    cidx[pos] = A.column_indices[keep]
    vals[pos] = A.values[keep]
    ne = A.num_entries - int((A.values == 0).sum().item())
    return ell_matrix(A.num_rows, A.num_cols, ne, K, pitch, cidx, vals)


def csr_to_hyb(A: csr_matrix, num_entries_per_row: int | None = None, alignment: int = 32) -> hyb_matrix:
    """csr_to_other.h:229-306: first K entries of each row -> ELL, the rest -> COO"""
    dev = A.values.device
    K = optimal_entries_per_row(A.row_offsets) if num_entries_per_row is None else num_entries_per_row
    rows, k = _slot_index(A.row_offsets, A.num_entries)
    pitch = round_up(A.num_rows, alignment)
    cidx = torch.full((K * pitch,), -1, dtype=torch.int32, device=dev)
    vals = torch.zeros(K * pitch, dtype=A.values.dtype, device=dev)
    in_ell = k < K
    pos = (k * pitch + rows)[in_ell]
    cidx[pos] = A.column_indices[in_ell]
    vals[pos] = A.values[in_ell]
    tail = ~in_ell
    coo = coo_matrix(A.num_rows, A.num_cols, rows[tail].to(torch.int32), A.column_indices[tail].contiguous(),
                     A.values[tail].contiguous())
    ell = ell_matrix(A.num_rows, A.num_cols, int(in_ell.sum().item()), K, pitch, cidx, vals)
    return hyb_matrix(ell, coo)
