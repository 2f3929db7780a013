"""cusp_autotuned_b200 — B200-native SpMV / BLAS-1 / CG engine behind the CUSP API.

Python-side mirror of the reference's operator interface for the hot path
(`cusp::multiply`, `cusp::blas`, `cusp::krylov::cg`, `cusp::ktt`), over the C ABI
of libb200sp.so (include/b200sp.h).  The C++ drop-in headers live in
include/cusp/.  Device memory and streams come from PyTorch; all arithmetic
runs in hand-written sm_100a kernels (csrc/).  No CPU fallback.
"""
from . import capi
from .capi import B200spError, InvalidInput, Cfg, Handle
from .matrix import (coo_matrix, csr_matrix, dia_matrix, ell_matrix, ellr_matrix, hyb_matrix,
                     default_handle, multiply, multiply_block)
from . import blas, gallery, krylov, ktt
from .krylov import monitor

__all__ = ["capi", "B200spError", "InvalidInput", "Cfg", "Handle", "coo_matrix", "csr_matrix",
           "dia_matrix", "ell_matrix", "ellr_matrix", "hyb_matrix", "default_handle", "multiply", "multiply_block",
           "blas", "gallery", "krylov", "ktt", "monitor"]
