"""shared test helpers (parity metrics, uploads)"""
import numpy as np

# north_star tolerances: per-entry relative error vs the reference's result
TOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-12}


def rel_err(got, want):
    """max_i |got-want| / |want| (entries with want == 0 must match exactly)"""
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    d = np.abs(got - want)
    den = np.abs(want)
    e = np.where(den > 0, d / np.where(den > 0, den, 1), np.where(d == 0, 0.0, np.inf))
    return float(e.max()) if e.size else 0.0


def scaled_err(got, want, scale):
    """max_i |got-want| / scale_i with scale_i = sum_j |a_ij x_j| — the bound that
    stays meaningful when a row cancels (6x_i - sum of neighbours ~ 0)"""
    d = np.abs(np.asarray(got, np.float64) - np.asarray(want, np.float64))
    s = np.asarray(scale, np.float64)
    e = np.where(s > 0, d / np.where(s > 0, s, 1), np.where(d == 0, 0.0, np.inf))
    return float(e.max()) if e.size else 0.0


def abs_matrix(A):
    """|A| as a matrix dict (for the cancellation-safe error scale)"""
    B = dict(A)
    if A["format"] == "hyb":
        B["ell"] = abs_matrix(A["ell"])
        B["coo"] = abs_matrix(A["coo"])
    else:
        B["values"] = np.abs(A["values"])
    return B


def upload(fmt, A, dev):
    from cusp_autotuned_b200 import gallery
    return gallery.from_host(fmt, A, dev)


def tdev(a, dev):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)
