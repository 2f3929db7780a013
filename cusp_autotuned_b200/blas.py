"""cusp::blas (cusp/detail/blas.inl:84-461) on device vectors, via the C ABI."""
from __future__ import annotations

import torch

from . import capi
from .matrix import default_handle


def _same(*ts):
    n = ts[0].numel()
    for t in ts:
        if t.numel() != n:  # cusp::assert_same_dimensions -> invalid_input_exception
            raise capi.InvalidInput(capi.ST_INVALID_INPUT, "array dimensions do not match")
        if not t.is_cuda:
            raise capi.InvalidInput(capi.ST_INVALID_INPUT, "device vectors required")


def axpy(x, y, alpha, handle=None):
    """y <- alpha*x + y"""
    _same(x, y)
    (handle or default_handle()).axpy(alpha, x, y)


def axpby(x, y, z, alpha, beta, handle=None):
    """z <- alpha*x + beta*y"""
    _same(x, y, z)
    (handle or default_handle()).axpby(alpha, x, beta, y, z)


def copy(x, y, handle=None):
    _same(x, y)
    (handle or default_handle()).copy(x, y)


def fill(x, alpha, handle=None):
    (handle or default_handle()).fill(alpha, x)


def scal(x, alpha, handle=None):
    (handle or default_handle()).scal(alpha, x)


def dot(x, y, handle=None) -> float:
    _same(x, y)
    return (handle or default_handle()).dot(x, y)


dotc = dot  # real value types: conj is the identity


def nrm2(x, handle=None) -> float:
    return (handle or default_handle()).nrm2(x)


def axpbypcz(x, y, z, output, alpha, beta, gamma, handle=None):
    """output <- alpha*x + beta*y + gamma*z"""
    _same(x, y, z, output)
    (handle or default_handle()).axpbypcz(alpha, x, beta, y, gamma, z, output)


def xmy(x, y, z, handle=None):
    """z <- x .* y"""
    _same(x, y, z)
    (handle or default_handle()).xmy(x, y, z)


def asum(x, handle=None) -> float:
    return (handle or default_handle()).asum(x)


nrm1 = asum


def nrmmax(x, handle=None) -> float:
    return (handle or default_handle()).nrmmax(x)


def amax(x, handle=None) -> int:
    """index of the first element of maximal magnitude"""
    return (handle or default_handle()).amax(x)
