"""cusp::krylov::{cg, bicgstab, cr} + cusp::monitor + cusp::precond::diagonal (cusp/krylov/detail/{cg,bicgstab,cr}.inl,
cusp/monitor.h:101-245, cusp/precond/diagonal.h) over b200sp_cg / b200sp_krylov."""
from __future__ import annotations

import sys

import numpy as np

from . import capi
from .matrix import default_handle


class monitor:
    """cusp::monitor<Real>(b, iteration_limit=500, relative_tolerance=1e-5,
    absolute_tolerance=0, verbose=False).  The stopping rule is evaluated on the
    device by b200sp_cg with exactly the reference's test
    ||r|| <= absolute + relative*||b||; this object carries the parameters in and
    the history out."""

    def __init__(self, b, iteration_limit=500, relative_tolerance=1e-5, absolute_tolerance=0.0, verbose=False):
        self._b = b
        self._limit = int(iteration_limit)
        self._rel = float(relative_tolerance)
        self._abs = float(absolute_tolerance)
        self.verbose = bool(verbose)
        self.reset(b)

    def reset(self, b):
        self._b = b
        self.b_norm = None
        self.r_norm = sys.float_info.max
        self._count = 0
        self.residuals = []

    def iteration_count(self):
        return self._count

    def iteration_limit(self):
        return self._limit

    def relative_tolerance(self):
        return self._rel

    def absolute_tolerance(self):
        return self._abs

    def residual_norm(self):
        return self.r_norm

    def tolerance(self):
        if self.b_norm is None:
            from . import blas
            self.b_norm = blas.nrm2(self._b)
        return self._abs + self._rel * self.b_norm

    def converged(self):
        return self.r_norm <= self.tolerance()

    def set_verbose(self, v=True):
        self.verbose = bool(v)

    def _absorb(self, res: capi.CgResult, hist):
        self._count = int(res.iteration_count)
        self.r_norm = float(res.residual_norm)
        self.b_norm = float(res.b_norm)
        self.residuals = [] if hist is None else [float(v) for v in hist]
        if self.verbose:
            for i, v in enumerate(self.residuals):
                print(f"       {i:10d}       {v:10.6e}")


def cg(A, x, b, mon: monitor | None = None, *, cfg=None, handle=None, check_interval=0, halo=None):
    """cusp::krylov::cg(A, x, b[, monitor]) with the identity preconditioner.
    x is updated in place; returns the monitor."""
    if A.num_rows != x.numel() or b.numel() != x.numel():
        raise capi.InvalidInput(capi.ST_INVALID_INPUT, "cg: dimension mismatch")
    if mon is None:
        mon = monitor(b)
    h = handle or default_handle()
    res, hist = h.cg(A.descriptor(), x, b, iteration_limit=mon.iteration_limit(),
                     relative_tolerance=mon.relative_tolerance(),
                     absolute_tolerance=mon.absolute_tolerance(), check_interval=check_interval, cfg=cfg,
                     halo=halo)
    mon._absorb(res, hist)
    return mon


class diagonal:
    """cusp::precond::diagonal<V, device_memory>(A): M = diag(A)^-1, held as `diagonal_reciprocals`
    (cusp/precond/detail/diagonal.inl:30-47: extract_diagonal, then reciprocal).  Set-up time: torch index ops."""

    def __init__(self, A):
        import torch
        from . import convert
        src = {capi.FMT_CSR: "csr", capi.FMT_COO: "coo", capi.FMT_DIA: "dia", capi.FMT_ELL: "ell", capi.FMT_HYB: "hyb"}[A.format]
        C = A if src == "coo" else convert.convert(A, "coo")
        d = torch.zeros(A.num_rows, dtype=C.values.dtype, device=C.values.device)
        on = C.row_indices == C.column_indices
        d[C.row_indices[on].to(torch.int64)] = C.values[on]
        self.diagonal_reciprocals = 1.0 / d


def _solve(solver, A, x, b, mon, M, cfg, handle, check_interval, halo):
    if A.num_rows != x.numel() or b.numel() != x.numel():
        raise capi.InvalidInput(capi.ST_INVALID_INPUT, f"{solver}: dimension mismatch")
    if mon is None:
        mon = monitor(b)
    h = handle or default_handle()
    dinv = None if M is None else M.diagonal_reciprocals
    res, hist = h.krylov(solver, A.descriptor(), x, b, diagonal_inverse=dinv, iteration_limit=mon.iteration_limit(),
                         relative_tolerance=mon.relative_tolerance(), absolute_tolerance=mon.absolute_tolerance(),
                         check_interval=check_interval, cfg=cfg, halo=halo)
    mon._absorb(res, hist)
    return mon


def pcg(A, x, b, mon: monitor | None = None, M: diagonal | None = None, *, cfg=None, handle=None, check_interval=0, halo=None):
    """cusp::krylov::cg(A, x, b, monitor, M) with M = None (identity) or a `diagonal` preconditioner"""
    return _solve("cg", A, x, b, mon, M, cfg, handle, check_interval, halo)


def bicgstab(A, x, b, mon: monitor | None = None, M: diagonal | None = None, *, cfg=None, handle=None, check_interval=0,
             halo=None):
    """cusp::krylov::bicgstab(A, x, b[, monitor[, M]]); monitor.residuals gets two entries per iteration (||r||, ||s||)"""
    return _solve("bicgstab", A, x, b, mon, M, cfg, handle, check_interval, halo)


def cr(A, x, b, mon: monitor | None = None, M: diagonal | None = None, *, cfg=None, handle=None, check_interval=0, halo=None):
    """cusp::krylov::cr(A, x, b[, monitor[, M]])"""
    return _solve("cr", A, x, b, mon, M, cfg, handle, check_interval, halo)
