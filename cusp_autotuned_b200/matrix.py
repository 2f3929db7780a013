"""Device-resident sparse containers with the reference's member names, and
`multiply(A, x, y)` (cusp/multiply.h:36-195).

Containers only hold torch tensors that already live on the GPU in the
reference's layouts (cusp/{csr,coo,dia,ell,hyb}_matrix.h; ELL/DIA column-major
with pitch).  They do no arithmetic; `multiply` hands raw pointers to the C ABI.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import capi

_default_handles = {}


def default_handle() -> capi.Handle:
    """one engine handle per (process, device) — like the reference's implicit
    per-context state, but explicit and per device"""
    if not torch.cuda.is_available():
        raise capi.B200spError(capi.ST_CUDA_ERROR, "no CUDA device: cusp_autotuned_b200 has no CPU fallback")
    dev = torch.cuda.current_device()
    h = _default_handles.get(dev)
    if h is None:
        h = _default_handles[dev] = capi.Handle()
    return h


def _dt(values: torch.Tensor) -> int:
    if values.dtype == torch.float32:
        return capi.F32
    if values.dtype == torch.float64:
        return capi.F64
    raise capi.InvalidInput(capi.ST_INVALID_INPUT, f"unsupported value type {values.dtype}")


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _check_dev(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise capi.InvalidInput(capi.ST_INVALID_INPUT, "device_memory container given a host tensor")
        if t is not None and not t.is_contiguous():
            raise capi.InvalidInput(capi.ST_INVALID_INPUT, "container arrays must be contiguous")


def _check_index(name, t, expected_len):
    """index arrays are int32 in the ABI: torch's default index dtype (int64 from arange / nonzero /
    crow_indices) would be reinterpreted by the kernels — refuse it here, and check the length the shape implies"""
    if t.dtype != torch.int32:
        raise capi.InvalidInput(capi.ST_INVALID_INPUT, f"{name} must be torch.int32, got {t.dtype}")
    if expected_len is not None and t.numel() != expected_len:
        raise capi.InvalidInput(capi.ST_INVALID_INPUT, f"{name} has {t.numel()} elements, the shape implies {expected_len}")


def _check_values(values, min_len=None):
    _dt(values)
    if min_len is not None and values.numel() < min_len:
        raise capi.InvalidInput(capi.ST_INVALID_INPUT, f"values has {values.numel()} elements, the shape needs {min_len}")


class _Base:
    format = -1

    def descriptor(self) -> capi.Matrix:
        raise NotImplementedError

    @property
    def shape(self):
        return (self.num_rows, self.num_cols)


class csr_matrix(_Base):
    """cusp::csr_matrix<int, V, device_memory> (cusp/csr_matrix.h:107-210)"""
    format = capi.FMT_CSR

    def __init__(self, num_rows, num_cols, row_offsets, column_indices, values):
        _check_values(values)
        _check_index("row_offsets", row_offsets, int(num_rows) + 1)
        _check_index("column_indices", column_indices, int(values.numel()))
        _check_dev(row_offsets, column_indices, values)
        self.num_rows, self.num_cols, self.num_entries = int(num_rows), int(num_cols), int(values.numel())
        self.row_offsets, self.column_indices, self.values = row_offsets, column_indices, values

    def descriptor(self):
        return capi.Matrix(format=self.format, dtype=_dt(self.values), num_rows=self.num_rows,
                           num_cols=self.num_cols, num_entries=self.num_entries,
                           row_offsets=_p(self.row_offsets), column_indices=_p(self.column_indices),
                           values=_p(self.values))


class coo_matrix(_Base):
    """cusp::coo_matrix (cusp/coo_matrix.h:116-225); rows sorted ascending"""
    format = capi.FMT_COO

    def __init__(self, num_rows, num_cols, row_indices, column_indices, values):
        _check_values(values)
        _check_index("row_indices", row_indices, int(values.numel()))
        _check_index("column_indices", column_indices, int(values.numel()))
        _check_dev(row_indices, column_indices, values)
        self.num_rows, self.num_cols, self.num_entries = int(num_rows), int(num_cols), int(values.numel())
        self.row_indices, self.column_indices, self.values = row_indices, column_indices, values

    def descriptor(self):
        return capi.Matrix(format=self.format, dtype=_dt(self.values), num_rows=self.num_rows,
                           num_cols=self.num_cols, num_entries=self.num_entries,
                           row_indices=_p(self.row_indices), column_indices=_p(self.column_indices),
                           values=_p(self.values))


class ell_matrix(_Base):
    """cusp::ell_matrix (cusp/ell_matrix.h:119-229): column_indices / values are
    column-major [num_cols_per_row][pitch] flat arrays, padding col = -1"""
    format = capi.FMT_ELL
    invalid_index = -1

    def __init__(self, num_rows, num_cols, num_entries, num_cols_per_row, pitch, column_indices, values):
        if int(pitch) < int(num_rows):
            raise capi.InvalidInput(capi.ST_INVALID_INPUT, "ell: pitch < num_rows")
        _check_values(values, int(num_cols_per_row) * int(pitch))
        _check_index("column_indices", column_indices, None)
        if column_indices.numel() < int(num_cols_per_row) * int(pitch):
            raise capi.InvalidInput(capi.ST_INVALID_INPUT, "ell: column_indices shorter than num_cols_per_row * pitch")
        _check_dev(column_indices, values)
        self.num_rows, self.num_cols, self.num_entries = int(num_rows), int(num_cols), int(num_entries)
        self.num_cols_per_row, self.pitch = int(num_cols_per_row), int(pitch)
        self.column_indices, self.values = column_indices, values

    def descriptor(self):
        return capi.Matrix(format=self.format, dtype=_dt(self.values), num_rows=self.num_rows,
                           num_cols=self.num_cols, num_entries=self.num_entries,
                           num_cols_per_row=self.num_cols_per_row, pitch=self.pitch,
                           column_indices=_p(self.column_indices), values=_p(self.values))


class ellr_matrix(ell_matrix):
    """cusp::ktt::ellr_matrix (cusp/ktt/ellr_matrix.h:18): ELL + row_lengths"""
    format = capi.FMT_ELLR

    def __init__(self, ell: ell_matrix, handle: Optional[capi.Handle] = None):
        super().__init__(ell.num_rows, ell.num_cols, ell.num_entries, ell.num_cols_per_row, ell.pitch,
                         ell.column_indices, ell.values)
        self.row_lengths = torch.empty(max(self.num_rows, 1), dtype=torch.int32, device=ell.values.device)
        (handle or default_handle()).ell_row_lengths(self.num_rows, self.num_cols_per_row, self.pitch,
                                                     self.column_indices, self.row_lengths)

    def descriptor(self):
        d = super().descriptor()
        d.format = capi.FMT_ELLR
        d.row_offsets = _p(self.row_lengths)
        return d


class dia_matrix(_Base):
    """cusp::dia_matrix (cusp/dia_matrix.h:120-227): values column-major [ndiag][pitch]"""
    format = capi.FMT_DIA

    def __init__(self, num_rows, num_cols, num_entries, diagonal_offsets, pitch, values):
        if int(pitch) < int(num_rows):
            raise capi.InvalidInput(capi.ST_INVALID_INPUT, "dia: pitch < num_rows")
        _check_index("diagonal_offsets", diagonal_offsets, None)
        _check_values(values, int(diagonal_offsets.numel()) * int(pitch))
        _check_dev(diagonal_offsets, values)
        self.num_rows, self.num_cols, self.num_entries = int(num_rows), int(num_cols), int(num_entries)
        self.diagonal_offsets, self.pitch, self.values = diagonal_offsets, int(pitch), values
        self.num_diagonals = int(diagonal_offsets.numel())

    def descriptor(self):
        return capi.Matrix(format=self.format, dtype=_dt(self.values), num_rows=self.num_rows,
                           num_cols=self.num_cols, num_entries=self.num_entries,
                           num_cols_per_row=self.num_diagonals, pitch=self.pitch,
                           diagonal_offsets=_p(self.diagonal_offsets), values=_p(self.values))


class hyb_matrix(_Base):
    """cusp::hyb_matrix (cusp/hyb_matrix.h:142-248): .ell + .coo"""
    format = capi.FMT_HYB

    def __init__(self, ell: ell_matrix, coo: coo_matrix):
        if ell.values.dtype != coo.values.dtype:
            raise capi.InvalidInput(capi.ST_INVALID_INPUT, "hyb: ell and coo parts must share one value type")
        if (ell.num_rows, ell.num_cols) != (coo.num_rows, coo.num_cols):
            raise capi.InvalidInput(capi.ST_INVALID_INPUT, "hyb: ell and coo parts must have the same shape")
        self.ell, self.coo = ell, coo
        self.num_rows, self.num_cols = ell.num_rows, ell.num_cols
        self.num_entries = ell.num_entries + coo.num_entries

    def descriptor(self):
        e, c = self.ell, self.coo
        return capi.Matrix(format=self.format, dtype=_dt(e.values), num_rows=self.num_rows,
                           num_cols=self.num_cols, num_entries=self.num_entries,
                           num_cols_per_row=e.num_cols_per_row, pitch=e.pitch,
                           column_indices=_p(e.column_indices), values=_p(e.values),
                           coo_num_entries=c.num_entries, coo_row_indices=_p(c.row_indices),
                           coo_column_indices=_p(c.column_indices), coo_values=_p(c.values))


def multiply(A: _Base, x: torch.Tensor, y: torch.Tensor, *, accumulate: bool = False,
             cfg: Optional[capi.Cfg] = None, handle: Optional[capi.Handle] = None) -> torch.Tensor:
    """cusp::multiply(A, x, y): y = A x  (accumulate=True: the 7-argument form with
    initialize = thrust::identity, y += A x).  Size mismatches raise InvalidInput
    like cusp::invalid_input_exception."""
    from . import ktt
    if x.numel() != A.num_cols or y.numel() != A.num_rows:
        raise capi.InvalidInput(capi.ST_INVALID_INPUT,
                                f"multiply: A is {A.num_rows}x{A.num_cols}, x has {x.numel()}, y has {y.numel()}")
    _check_dev(x, y)
    if x.dtype != y.dtype or x.dtype != A.descriptor_dtype():
        raise capi.InvalidInput(capi.ST_INVALID_INPUT, "multiply: A, x and y must share one value type")
    h = handle or default_handle()
    d = A.descriptor()
    # plain cusp::multiply on ELL/DIA does one step of dynamic tuning per call
    # while ktt is enabled (cusp/system/detail/generic/multiply.inl:141-154)
    if cfg is None and not accumulate and ktt.is_enabled() and A.format in (capi.FMT_ELL, capi.FMT_DIA,
                                                                            capi.FMT_ELLR):
        h.tune_step(d, x, y)
    else:
        h.spmv(d, x, y, accumulate=accumulate, cfg=cfg)
    return y


def multiply_block(A: "csr_matrix", X: torch.Tensor, Y: torch.Tensor, *, accumulate: bool = False,
                   handle: Optional[capi.Handle] = None) -> torch.Tensor:
    """cusp::multiply(A, X, Y) for a csr_matrix and row-major dense blocks X [num_cols x k],
    Y [num_rows x k] (cusp/system/cuda/detail/multiply/csr_block_spmv.h:181-222)."""
    if not isinstance(A, csr_matrix):
        raise capi.InvalidInput(capi.ST_NOT_IMPLEMENTED, "multiply_block: csr_matrix only (like the reference)")
    if X.dim() != 2 or Y.dim() != 2 or X.shape[0] != A.num_cols or Y.shape[0] != A.num_rows or X.shape[1] != Y.shape[1]:
        raise capi.InvalidInput(capi.ST_INVALID_INPUT, "multiply_block: shapes do not match")
    if not (X.is_cuda and Y.is_cuda):
        raise capi.InvalidInput(capi.ST_INVALID_INPUT, "device_memory container given a host tensor")
    if X.dtype != Y.dtype or X.dtype != A.values.dtype or X.stride(1) != 1 or Y.stride(1) != 1:
        raise capi.InvalidInput(capi.ST_INVALID_INPUT, "multiply_block: one value type, row-major blocks")
    h = handle or default_handle()
    h.spmm_csr(A.num_rows, A.num_cols, A.num_entries, A.row_offsets, A.column_indices, A.values, X.shape[1],
               X, X.stride(0), Y, Y.stride(0), accumulate=accumulate)
    return Y


def _descriptor_dtype(self):
    v = self.ell.values if isinstance(self, hyb_matrix) else self.values
    return v.dtype


_Base.descriptor_dtype = _descriptor_dtype
