"""cusp::ktt (cusp/ktt/ktt.h:14-127): enable / disable / multiply / tune /
reset_tuning over the engine's own tuner (b200sp_tune*)."""
from __future__ import annotations

from . import capi
from .matrix import default_handle

_enabled = True  # cusp/ktt/detail/ktt.inl:21 `is_enabled = true`


def enable():
    global _enabled
    _enabled = True


def disable():
    global _enabled
    _enabled = False


def is_enabled() -> bool:
    return _enabled


def get_tuner(handle=None):
    """the reference returns the global ::ktt::Tuner; here the engine handle owns the tuner state"""
    return handle or default_handle()


def multiply(A, x, y, configuration: capi.Cfg | None = None, handle=None):
    """cusp::ktt::multiply(A,x,y): one step of dynamic autotuning;
    cusp::ktt::multiply(A,x,y,conf): run exactly `conf`."""
    h = handle or default_handle()
    if configuration is not None:
        h.spmv(A.descriptor(), x, y, cfg=configuration)
        r = capi.TuneResult()
        r.cfg = configuration
        return r
    return h.tune_step(A.descriptor(), x, y)


def tune(A, x, y, reference=None, tol=0.0, repeats=5, handle=None):
    """cusp::ktt::tune(A,x,y[,reference_computation]): exhaustive offline tuning,
    every configuration validated; returns (best, results)."""
    h = handle or default_handle()
    return h.tune(A.descriptor(), x, y, y_reference=reference, tol=tol, repeats=repeats)


def reset_tuning(A=None, x=None, y=None, handle=None):
    (handle or default_handle()).tune_reset(A.descriptor() if A is not None else None)
