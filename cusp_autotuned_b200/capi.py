"""ctypes binding of libb200sp.so (the C ABI declared in include/b200sp.h).

This is the only way Python reaches the engine: every call below goes through
the C-ABI entry point of the same name.  PyTorch is used for device memory and
streams only.  There is no CPU fallback: if the shared library is missing, or no
B200 is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200sp.so")


class B200spError(RuntimeError):
    """cusp::runtime_exception analogue (cusp/exception.h:74-84)."""

    def __init__(self, status: int, message: str):
        super().__init__(f"[b200sp status {status}] {message}")
        self.status = status


class InvalidInput(B200spError, ValueError):
    """cusp::invalid_input_exception analogue (cusp/exception.h:53-58)."""


class Cfg(C.Structure):
    """b200sp_cfg — one point of the tuning space."""

    _fields_ = [
        ("kernel", C.c_int),
        ("block_size", C.c_int),
        ("threads_per_row", C.c_int),
        ("unroll", C.c_int),
        ("vector_width", C.c_int),
        ("tile_rows", C.c_int),
        ("stages", C.c_int),
        ("ctas_per_sm", C.c_int),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}

    def __repr__(self):
        return "Cfg(" + ", ".join(f"{k}={v}" for k, v in self.as_dict().items() if v) + ")"


class Functors(C.Structure):
    """b200sp_functors — (initialize, combine, reduce) of the generalized product, by code"""

    _fields_ = [("initialize", C.c_int), ("init_value", C.c_double), ("combine", C.c_int), ("reduce", C.c_int)]


INIT_CONSTANT, INIT_IDENTITY = 0, 1
COMBINE = {"multiplies": 0, "plus": 1, "minimum": 2, "maximum": 3, "project2nd": 4}
REDUCE = {"plus": 0, "minimum": 1, "maximum": 2}


class Matrix(C.Structure):
    """b200sp_matrix — non-owning device matrix descriptor."""

    _fields_ = [
        ("format", C.c_int),
        ("dtype", C.c_int),
        ("num_rows", C.c_int64),
        ("num_cols", C.c_int64),
        ("num_entries", C.c_int64),
        ("num_cols_per_row", C.c_int64),
        ("pitch", C.c_int64),
        ("row_offsets", C.c_void_p),
        ("row_indices", C.c_void_p),
        ("column_indices", C.c_void_p),
        ("diagonal_offsets", C.c_void_p),
        ("values", C.c_void_p),
        ("coo_num_entries", C.c_int64),
        ("coo_row_indices", C.c_void_p),
        ("coo_column_indices", C.c_void_p),
        ("coo_values", C.c_void_p),
    ]


class CgParams(C.Structure):
    _fields_ = [
        ("iteration_limit", C.c_int64),
        ("relative_tolerance", C.c_double),
        ("absolute_tolerance", C.c_double),
        ("check_interval", C.c_int),
    ]


class CgResult(C.Structure):
    _fields_ = [
        ("iteration_count", C.c_int64),
        ("converged", C.c_int),
        ("residual_norm", C.c_double),
        ("b_norm", C.c_double),
        ("num_residuals", C.c_int64),
    ]


class Halo(C.Structure):
    _fields_ = [("halo_lo", C.c_int64), ("halo_hi", C.c_int64)]


class ConvertInfo(C.Structure):
    """b200sp_convert_info"""
    _fields_ = [("max_entries_per_row", C.c_int64), ("hyb_entries_per_row", C.c_int64),
                ("hyb_coo_entries", C.c_int64), ("num_diagonals", C.c_int64)]


class TuneResult(C.Structure):
    _fields_ = [
        ("cfg", Cfg),
        ("status", C.c_int),
        ("milliseconds", C.c_float),
        ("max_rel_error", C.c_double),
    ]


FMT_CSR, FMT_ELL, FMT_DIA, FMT_COO, FMT_HYB, FMT_ELLR = range(6)
F32, F64 = 0, 1
K_CSR_VECTOR, K_CSR_STREAM, K_CSR_RING, K_CSR_BALANCED = 1, 2, 3, 4
K_ELL_LDG, K_ELL_BULK = 1, 2
K_DIA_LDG, K_DIA_BULK = 1, 2
K_COO_SEGSCAN, K_COO_RING, K_COO_WARP = 1, 2, 3
ST_OK, ST_INVALID_INPUT, ST_CUDA_ERROR, ST_NOT_IMPLEMENTED, ST_ALLOC_FAILED, ST_COMM_ERROR = range(6)

# every symbol include/b200sp.h declares (checked by tests/test_abi.py)
_SFX = ("f32", "f64")
EXPORTED_SYMBOLS = (
    ["b200sp_version", "b200sp_create", "b200sp_destroy", "b200sp_last_error_string",
     "b200sp_status_string", "b200sp_launch_count", "b200sp_set_l2_persist",
     "b200sp_ell_row_lengths", "b200sp_csr_row_starts", "b200sp_spmv", "b200sp_spmv_generalized", "b200sp_spmv_host", "b200sp_cg", "b200sp_krylov", "b200sp_spmv_graph_create", "b200sp_graph_launch",
     "b200sp_graph_destroy",
     "b200sp_comm_unique_id", "b200sp_comm_init", "b200sp_comm_destroy", "b200sp_cg_dist",
     "b200sp_spmv_dist", "b200sp_spmv_dist_host", "b200sp_spmv_dist_gather", "b200sp_comm_p2p_enabled", "b200sp_comm_timeouts", "b200sp_cfg_space", "b200sp_tune", "b200sp_tune_ex", "b200sp_tune_step",
     "b200sp_tune_reset", "b200sp_tune_lookup", "b200sp_tune_save", "b200sp_tune_load",
     "b200sp_poisson_num_entries", "b200sp_poisson_csr_offsets",
     "b200sp_offsets_to_indices", "b200sp_indices_to_offsets", "b200sp_csr_convert_query"]
    + [f"b200sp_{op}_{s}" for op in ("csr_to_ell", "csr_to_coo_tail", "csr_to_dia", "count_zeros", "dia_to_ell",
                                       "dia_to_csr_offsets", "dia_to_csr_fill", "ell_to_csr_offsets", "ell_to_csr_fill",
                                       "hyb_to_csr_offsets", "hyb_to_csr_fill") for s in _SFX]
    + [f"b200sp_spmv_{f}_{s}" for f in ("csr", "ell", "dia", "coo", "hyb", "ellr") for s in _SFX]
    + [f"b200sp_{op}_{s}" for op in ("axpy", "axpby", "axpbypcz", "xmy", "copy", "fill", "scal", "dot", "nrm2", "asum", "nrmmax",
                                       "amax") for s in _SFX]
    + [f"b200sp_poisson_{f}_{s}" for f in ("dia", "ell", "csr") for s in _SFX]
    + [f"b200sp_spmm_csr_{s}" for s in _SFX]
    + ["b200sp_coo_plan_create", "b200sp_coo_plan_destroy", "b200sp_coo_plan_info", "b200sp_coo_plan_attach",
       "b200sp_coo_plan_detach"]
    + [f"b200sp_spmv_coo_plan_{s}" for s in _SFX]
)

_lib: Optional[C.CDLL] = None


def load_library() -> C.CDLL:
    """dlopen libb200sp.so (built in-tree by __graft_entry__.build() / csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200spError(-1, f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    lib.b200sp_version.restype = C.c_int
    lib.b200sp_last_error_string.restype = C.c_char_p
    lib.b200sp_last_error_string.argtypes = [C.c_void_p]
    lib.b200sp_status_string.restype = C.c_char_p
    lib.b200sp_launch_count.restype = C.c_uint64
    lib.b200sp_launch_count.argtypes = [C.c_void_p]
    lib.b200sp_cfg_space.restype = C.c_int64
    lib.b200sp_cfg_space.argtypes = [C.c_int, C.c_int, C.POINTER(Cfg), C.c_int64]
    lib.b200sp_poisson_num_entries.restype = C.c_int64
    lib.b200sp_poisson_num_entries.argtypes = [C.c_int] + [C.c_int64] * 5
    lib.b200sp_tune_lookup.restype = C.c_int
    _lib = lib
    return lib


def _ptr(t):
    """device (or host) address of a torch tensor / numpy array / None."""
    if t is None:
        return C.c_void_p(0)
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)


def _sfx(dtype) -> str:
    import torch
    if dtype == torch.float32:
        return "f32"
    if dtype == torch.float64:
        return "f64"
    raise InvalidInput(ST_INVALID_INPUT, f"unsupported value type {dtype}: the engine computes in f32 or f64")


def _ctype(dtype):
    import torch
    return C.c_float if dtype == torch.float32 else C.c_double


def _stream() -> C.c_void_p:
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Handle:
    """b200sp_handle bound to the current CUDA device."""

    def __init__(self):
        self.lib = load_library()
        h = C.c_void_p()
        st = self.lib.b200sp_create(C.byref(h))
        if st != ST_OK:
            raise B200spError(st, self.lib.b200sp_last_error_string(None).decode())
        self._h = h

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.b200sp_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, st: int):
        if st == ST_OK:
            return
        msg = self.lib.b200sp_last_error_string(self._h).decode()
        raise (InvalidInput if st == ST_INVALID_INPUT else B200spError)(st, msg)

    @property
    def launch_count(self) -> int:
        return int(self.lib.b200sp_launch_count(self._h))

    # -- typed SpMV entry points ------------------------------------------------
    def spmv_csr(self, rows, cols, nnz, Ap, Aj, Ax, x, y, accumulate=False, cfg: Optional[Cfg] = None):
        f = getattr(self.lib, "b200sp_spmv_csr_" + _sfx(y.dtype))
        self.check(f(self._h, _stream(), C.c_int64(rows), C.c_int64(cols), C.c_int64(nnz), _ptr(Ap), _ptr(Aj),
                     _ptr(Ax), _ptr(x), _ptr(y), C.c_int(int(accumulate)), C.byref(cfg) if cfg else None))

    def spmm_csr(self, rows, cols, nnz, Ap, Aj, Ax, k, X, ldx, Y, ldy, accumulate=False):
        """Y[rows x k] = (accumulate ? Y : 0) + A X[cols x k], X / Y row-major (cusp::multiply(csr, array2d, array2d))"""
        f = getattr(self.lib, "b200sp_spmm_csr_" + _sfx(Y.dtype))
        self.check(f(self._h, _stream(), C.c_int64(rows), C.c_int64(cols), C.c_int64(nnz), _ptr(Ap), _ptr(Aj),
                     _ptr(Ax), C.c_int64(k), _ptr(X), C.c_int64(ldx), _ptr(Y), C.c_int64(ldy),
                     C.c_int(int(accumulate))))

    # -- inspector / executor COO product (hot columns of x in shared memory) -------------------------
    def coo_plan_create(self, rows, cols, nnz, Ai, Aj, dtype: int, table_bytes: int = 0):
        plan = C.c_void_p()
        self.check(self.lib.b200sp_coo_plan_create(self._h, _stream(), C.c_int64(rows), C.c_int64(cols), C.c_int64(nnz),
                                                   _ptr(Ai), _ptr(Aj), C.c_int(dtype), C.c_int64(table_bytes),
                                                   C.byref(plan)))
        return plan

    def coo_plan_destroy(self, plan):
        self.check(self.lib.b200sp_coo_plan_destroy(self._h, plan))

    def coo_plan_info(self, plan):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self.check(self.lib.b200sp_coo_plan_info(plan, C.byref(a), C.byref(b), C.byref(c)))
        return {"hot_columns": a.value, "hot_entries": b.value, "capacity": c.value}

    def coo_plan_attach(self, plan):
        self.check(self.lib.b200sp_coo_plan_attach(self._h, plan))

    def coo_plan_detach(self, plan):
        self.check(self.lib.b200sp_coo_plan_detach(self._h, plan))

    def spmv_coo_plan(self, plan, Ax, x, y, accumulate=False, cfg: Optional[Cfg] = None):
        f = getattr(self.lib, "b200sp_spmv_coo_plan_" + _sfx(y.dtype))
        self.check(f(self._h, _stream(), plan, _ptr(Ax), _ptr(x), _ptr(y), C.c_int(int(accumulate)),
                     C.byref(cfg) if cfg else None))

    def spmv_ell(self, rows, cols, K, pitch, cidx, vals, x, y, accumulate=False, cfg: Optional[Cfg] = None):
        f = getattr(self.lib, "b200sp_spmv_ell_" + _sfx(y.dtype))
        self.check(f(self._h, _stream(), C.c_int64(rows), C.c_int64(cols), C.c_int64(K), C.c_int64(pitch),
                     _ptr(cidx), _ptr(vals), _ptr(x), _ptr(y), C.c_int(int(accumulate)),
                     C.byref(cfg) if cfg else None))

    def spmv_ellr(self, rows, cols, K, pitch, cidx, vals, row_lengths, x, y, accumulate=False,
                  cfg: Optional[Cfg] = None):
        f = getattr(self.lib, "b200sp_spmv_ellr_" + _sfx(y.dtype))
        self.check(f(self._h, _stream(), C.c_int64(rows), C.c_int64(cols), C.c_int64(K), C.c_int64(pitch),
                     _ptr(cidx), _ptr(vals), _ptr(row_lengths), _ptr(x), _ptr(y), C.c_int(int(accumulate)),
                     C.byref(cfg) if cfg else None))

    def ell_row_lengths(self, rows, K, pitch, cidx, out):
        self.check(self.lib.b200sp_ell_row_lengths(self._h, _stream(), C.c_int64(rows), C.c_int64(K),
                                                   C.c_int64(pitch), _ptr(cidx), _ptr(out)))

    def spmv_dia(self, rows, cols, ndiag, pitch, offs, vals, x, y, accumulate=False, cfg: Optional[Cfg] = None):
        f = getattr(self.lib, "b200sp_spmv_dia_" + _sfx(y.dtype))
        self.check(f(self._h, _stream(), C.c_int64(rows), C.c_int64(cols), C.c_int64(ndiag), C.c_int64(pitch),
                     _ptr(offs), _ptr(vals), _ptr(x), _ptr(y), C.c_int(int(accumulate)),
                     C.byref(cfg) if cfg else None))

    def spmv_coo(self, rows, cols, nnz, Ai, Aj, Ax, x, y, accumulate=False, cfg: Optional[Cfg] = None):
        f = getattr(self.lib, "b200sp_spmv_coo_" + _sfx(y.dtype))
        self.check(f(self._h, _stream(), C.c_int64(rows), C.c_int64(cols), C.c_int64(nnz), _ptr(Ai), _ptr(Aj),
                     _ptr(Ax), _ptr(x), _ptr(y), C.c_int(int(accumulate)), C.byref(cfg) if cfg else None))

    def spmv_hyb(self, rows, cols, K, pitch, ecidx, evals, cnnz, ci, cj, cv, x, y, accumulate=False,
                 ell_cfg: Optional[Cfg] = None, coo_cfg: Optional[Cfg] = None):
        f = getattr(self.lib, "b200sp_spmv_hyb_" + _sfx(y.dtype))
        self.check(f(self._h, _stream(), C.c_int64(rows), C.c_int64(cols), C.c_int64(K), C.c_int64(pitch),
                     _ptr(ecidx), _ptr(evals), C.c_int64(cnnz), _ptr(ci), _ptr(cj), _ptr(cv), _ptr(x), _ptr(y),
                     C.c_int(int(accumulate)), C.byref(ell_cfg) if ell_cfg else None,
                     C.byref(coo_cfg) if coo_cfg else None))

    # -- descriptor based ---------------------------------------------------------
    def spmv(self, A: Matrix, x, y, accumulate=False, cfg: Optional[Cfg] = None):
        self.check(self.lib.b200sp_spmv(self._h, _stream(), C.byref(A), _ptr(x), _ptr(y),
                                        C.c_int(int(accumulate)), C.byref(cfg) if cfg else None))

    def spmv_generalized(self, A: Matrix, x, y, initialize="constant", init_value=0.0, combine="multiplies",
                         reduce="plus"):
        """y[i] = reduce(initialize(y[i]), combine(a_ij, x_j) ...): cusp::multiply's 7-argument form by functor code"""
        f = Functors(INIT_IDENTITY if initialize == "identity" else INIT_CONSTANT, float(init_value), COMBINE[combine],
                     REDUCE[reduce])
        self.check(self.lib.b200sp_spmv_generalized(self._h, _stream(), C.byref(A), _ptr(x), _ptr(y), C.byref(f)))

    def csr_row_starts(self, num_rows, num_entries, row_offsets, workers, out):
        """row_starts[w] = row containing entry w * ceil(nnz / workers) (balanced-CSR preprocessing)"""
        self.check(self.lib.b200sp_csr_row_starts(self._h, _stream(), C.c_int64(num_rows), C.c_int64(num_entries),
                                                  _ptr(row_offsets), C.c_int64(workers), _ptr(out)))

    def spmv_graph_create(self, A: Matrix, x, y, count: int, accumulate=False, cfg: Optional[Cfg] = None):
        """`count` back-to-back products captured in one CUDA graph (launch-bound sizes); replay with graph_launch"""
        g = C.c_void_p()
        self.check(self.lib.b200sp_spmv_graph_create(self._h, _stream(), C.byref(A), _ptr(x), _ptr(y), C.c_int(int(accumulate)),
                                                     C.byref(cfg) if cfg else None, C.c_int(count), C.byref(g)))
        return g

    def graph_launch(self, graph):
        self.check(self.lib.b200sp_graph_launch(self._h, _stream(), graph))

    def graph_destroy(self, graph):
        self.check(self.lib.b200sp_graph_destroy(self._h, graph))

    def spmv_host(self, A: Matrix, x_host, y_host, accumulate=False, cfg: Optional[Cfg] = None):
        self.check(self.lib.b200sp_spmv_host(self._h, _stream(), C.byref(A), _ptr(x_host), _ptr(y_host),
                                             C.c_int(int(accumulate)), C.byref(cfg) if cfg else None))

    def set_l2_persist(self, tensor=None):
        nbytes = 0 if tensor is None else tensor.numel() * tensor.element_size()
        self.check(self.lib.b200sp_set_l2_persist(self._h, _stream(), _ptr(tensor), C.c_size_t(nbytes)))

    # -- BLAS-1 -----------------------------------------------------------------------
    def axpy(self, alpha, x, y):
        f = getattr(self.lib, "b200sp_axpy_" + _sfx(y.dtype))
        self.check(f(self._h, _stream(), C.c_int64(y.numel()), _ctype(y.dtype)(alpha), _ptr(x), _ptr(y)))

    def axpby(self, alpha, x, beta, y, z):
        f = getattr(self.lib, "b200sp_axpby_" + _sfx(z.dtype))
        ct = _ctype(z.dtype)
        self.check(f(self._h, _stream(), C.c_int64(z.numel()), ct(alpha), _ptr(x), ct(beta), _ptr(y), _ptr(z)))

    def copy(self, x, y):
        f = getattr(self.lib, "b200sp_copy_" + _sfx(y.dtype))
        self.check(f(self._h, _stream(), C.c_int64(y.numel()), _ptr(x), _ptr(y)))

    def fill(self, alpha, x):
        f = getattr(self.lib, "b200sp_fill_" + _sfx(x.dtype))
        self.check(f(self._h, _stream(), C.c_int64(x.numel()), _ctype(x.dtype)(alpha), _ptr(x)))

    def scal(self, alpha, x):
        f = getattr(self.lib, "b200sp_scal_" + _sfx(x.dtype))
        self.check(f(self._h, _stream(), C.c_int64(x.numel()), _ctype(x.dtype)(alpha), _ptr(x)))

    def axpbypcz(self, alpha, x, beta, y, gamma, z, out):
        f = getattr(self.lib, "b200sp_axpbypcz_" + _sfx(out.dtype))
        ct = _ctype(out.dtype)
        self.check(f(self._h, _stream(), C.c_int64(out.numel()), ct(alpha), _ptr(x), ct(beta), _ptr(y), ct(gamma),
                     _ptr(z), _ptr(out)))

    def xmy(self, x, y, z):
        f = getattr(self.lib, "b200sp_xmy_" + _sfx(z.dtype))
        self.check(f(self._h, _stream(), C.c_int64(z.numel()), _ptr(x), _ptr(y), _ptr(z)))

    def _reduce1(self, name, x) -> float:
        f = getattr(self.lib, f"b200sp_{name}_" + _sfx(x.dtype))
        out = _ctype(x.dtype)()
        self.check(f(self._h, _stream(), C.c_int64(x.numel()), _ptr(x), None, C.byref(out)))
        return out.value

    def asum(self, x) -> float:
        return self._reduce1("asum", x)

    def nrmmax(self, x) -> float:
        return self._reduce1("nrmmax", x)

    def amax(self, x) -> int:
        f = getattr(self.lib, "b200sp_amax_" + _sfx(x.dtype))
        out = C.c_int()
        self.check(f(self._h, _stream(), C.c_int64(x.numel()), _ptr(x), C.byref(out)))
        return out.value

    def dot(self, x, y) -> float:
        f = getattr(self.lib, "b200sp_dot_" + _sfx(x.dtype))
        out = _ctype(x.dtype)()
        self.check(f(self._h, _stream(), C.c_int64(x.numel()), _ptr(x), _ptr(y), None, C.byref(out)))
        return out.value

    def nrm2(self, x) -> float:
        f = getattr(self.lib, "b200sp_nrm2_" + _sfx(x.dtype))
        out = _ctype(x.dtype)()
        self.check(f(self._h, _stream(), C.c_int64(x.numel()), _ptr(x), None, C.byref(out)))
        return out.value

    # -- CG ------------------------------------------------------------------------------
    def cg(self, A: Matrix, x, b, iteration_limit=500, relative_tolerance=1e-5, absolute_tolerance=0.0,
           check_interval=0, cfg: Optional[Cfg] = None, want_residuals=True, halo: Optional[Halo] = None):
        import numpy as np
        prm = CgParams(iteration_limit, relative_tolerance, absolute_tolerance, check_interval)
        res = CgResult()
        hist = np.zeros(iteration_limit + 2, dtype=np.float64) if want_residuals else None
        hp = hist.ctypes.data_as(C.c_void_p) if hist is not None else None
        if halo is None:
            st = self.lib.b200sp_cg(self._h, _stream(), C.byref(A), _ptr(x), _ptr(b), C.byref(prm),
                                    C.byref(cfg) if cfg else None, C.byref(res), hp)
        else:
            st = self.lib.b200sp_cg_dist(self._h, _stream(), C.byref(A), C.byref(halo), _ptr(x), _ptr(b),
                                         C.byref(prm), C.byref(cfg) if cfg else None, C.byref(res), hp)
        self.check(st)
        return res, (hist[: res.num_residuals] if hist is not None else None)

    SOLVERS = {"cg": 0, "bicgstab": 1, "cr": 2}

    def krylov(self, solver: str, A: Matrix, x, b, diagonal_inverse=None, iteration_limit=500, relative_tolerance=1e-5,
               absolute_tolerance=0.0, check_interval=0, cfg: Optional[Cfg] = None, halo: Optional[Halo] = None):
        """b200sp_krylov: fused cg / bicgstab / cr with the identity or a diagonal preconditioner; one GPU or the
        row-block partitioned form (halo).  Returns (result, monitor.residuals)."""
        import numpy as np
        prm = CgParams(iteration_limit, relative_tolerance, absolute_tolerance, check_interval)
        res = CgResult()
        hist = np.zeros(2 * iteration_limit + 4, dtype=np.float64)
        self.check(self.lib.b200sp_krylov(self._h, _stream(), C.c_int(self.SOLVERS[solver]), C.byref(A),
                                          C.byref(halo) if halo is not None else None, _ptr(x), _ptr(b),
                                          _ptr(diagonal_inverse), C.byref(prm), C.byref(cfg) if cfg else None,
                                          C.byref(res), hist.ctypes.data_as(C.c_void_p)))
        return res, hist[: res.num_residuals]

    # -- multi-GPU ---------------------------------------------------------------------------
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        self.check(self.lib.b200sp_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, unique_id: bytes, world_size: int, rank: int):
        buf = C.create_string_buffer(unique_id, 128)
        self.check(self.lib.b200sp_comm_init(self._h, buf, C.c_int(world_size), C.c_int(rank)))

    def comm_destroy(self):
        self.check(self.lib.b200sp_comm_destroy(self._h))

    def comm_timeouts(self) -> int:
        """cross-GPU waits that gave up since comm_init (must be 0 for valid results)"""
        self.lib.b200sp_comm_timeouts.restype = C.c_int64
        return int(self.lib.b200sp_comm_timeouts(self._h, _stream()))

    def comm_p2p_enabled(self) -> bool:
        """True when cg_dist runs its exchanges as stores into NVLink peer memory"""
        return bool(self.lib.b200sp_comm_p2p_enabled(self._h))

    def spmv_dist(self, A: Matrix, halo: Halo, x_window, y, cfg: Optional[Cfg] = None):
        self.check(self.lib.b200sp_spmv_dist(self._h, _stream(), C.byref(A), C.byref(halo), _ptr(x_window),
                                             _ptr(y), C.byref(cfg) if cfg else None))

    def spmv_dist_host(self, A: Matrix, halo: Halo, x_host_local, y_host_local, cfg: Optional[Cfg] = None):
        """this rank's slices of x / y in (pinned) host memory; halo planes exchanged between the GPUs"""
        self.check(self.lib.b200sp_spmv_dist_host(self._h, _stream(), C.byref(A), C.byref(halo), _ptr(x_host_local),
                                                  _ptr(y_host_local), C.byref(cfg) if cfg else None))

    def spmv_dist_gather(self, A: Matrix, slice_offsets, x_full, y, cfg: Optional[Cfg] = None):
        """row block with global column indices; gathers the other ranks' x slices, then multiplies"""
        offs = (C.c_int64 * len(slice_offsets))(*[int(o) for o in slice_offsets])
        self.check(self.lib.b200sp_spmv_dist_gather(self._h, _stream(), C.byref(A), offs, _ptr(x_full), _ptr(y),
                                                    C.byref(cfg) if cfg else None))

    # -- tuning ----------------------------------------------------------------------------------
    @staticmethod
    def cfg_space(fmt: int, dtype: int):
        lib = load_library()
        n = lib.b200sp_cfg_space(fmt, dtype, None, 0)
        arr = (Cfg * n)()
        lib.b200sp_cfg_space(fmt, dtype, arr, n)
        return list(arr)

    def tune(self, A: Matrix, x, y, y_reference=None, tol=0.0, repeats=5):
        n = self.lib.b200sp_cfg_space(A.format, A.dtype, None, 0)
        results = (TuneResult * n)()
        count = C.c_int64(0)
        best = Cfg()
        self.check(self.lib.b200sp_tune(self._h, _stream(), C.byref(A), _ptr(x), _ptr(y), _ptr(y_reference),
                                        C.c_double(tol), C.c_int(repeats), results, C.c_int64(n),
                                        C.byref(count), C.byref(best)))
        return best, list(results)[: count.value]

    def tune_ex(self, A: Matrix, x, y, order=None, stop=None, y_reference=None, tol=0.0, repeats=5):
        """b200sp_tune_ex: `order` = indices into cfg_space() (the searcher), `stop(result) -> bool` is consulted
        after every configuration (the stop condition).  Returns (best, results visited)."""
        n = self.lib.b200sp_cfg_space(A.format, A.dtype, None, 0)
        visits = len(order) if order is not None else n
        results = (TuneResult * max(visits, 1))()
        count = C.c_int64(0)
        best = Cfg()
        CB = C.CFUNCTYPE(C.c_int, C.POINTER(TuneResult), C.c_void_p)
        cb = CB(lambda r, _u: int(bool(stop(r.contents)))) if stop else C.cast(None, CB)
        arr = (C.c_int64 * max(visits, 1))(*(order if order is not None else []))
        self.check(self.lib.b200sp_tune_ex(self._h, _stream(), C.byref(A), _ptr(x), _ptr(y), _ptr(y_reference),
                                           C.c_double(tol), C.c_int(repeats), arr if order is not None else None,
                                           C.c_int64(len(order) if order is not None else 0), cb, None, results,
                                           C.c_int64(max(visits, 1)), C.byref(count), C.byref(best)))
        return best, list(results)[: count.value]

    def tune_step(self, A: Matrix, x, y) -> TuneResult:
        r = TuneResult()
        self.check(self.lib.b200sp_tune_step(self._h, _stream(), C.byref(A), _ptr(x), _ptr(y), C.byref(r)))
        return r

    def tune_reset(self, A: Optional[Matrix] = None):
        self.check(self.lib.b200sp_tune_reset(self._h, C.byref(A) if A is not None else None))

    def tune_lookup(self, A: Matrix) -> Optional[Cfg]:
        c = Cfg()
        return c if self.lib.b200sp_tune_lookup(self._h, C.byref(A), C.byref(c)) else None

    def tune_save(self, path: str):
        self.check(self.lib.b200sp_tune_save(self._h, path.encode()))

    def tune_load(self, path: str):
        self.check(self.lib.b200sp_tune_load(self._h, path.encode()))

    # -- gallery ------------------------------------------------------------------------------------
    def poisson_dia(self, stencil, nx, ny, nz, row_begin, num_rows, col_shift, pitch, offsets, values):
        f = getattr(self.lib, "b200sp_poisson_dia_" + _sfx(values.dtype))
        self.check(f(self._h, _stream(), C.c_int(stencil), C.c_int64(nx), C.c_int64(ny), C.c_int64(nz),
                     C.c_int64(row_begin), C.c_int64(num_rows), C.c_int64(col_shift), C.c_int64(pitch),
                     _ptr(offsets), _ptr(values)))

    def poisson_ell(self, stencil, nx, ny, nz, row_begin, num_rows, col_shift, pitch, cidx, values):
        f = getattr(self.lib, "b200sp_poisson_ell_" + _sfx(values.dtype))
        self.check(f(self._h, _stream(), C.c_int(stencil), C.c_int64(nx), C.c_int64(ny), C.c_int64(nz),
                     C.c_int64(row_begin), C.c_int64(num_rows), C.c_int64(col_shift), C.c_int64(pitch),
                     _ptr(cidx), _ptr(values)))

    # -- conversions on the device (csrc/convert.cu) ------------------------------------
    def offsets_to_indices(self, row_offsets, row_indices):
        self.check(self.lib.b200sp_offsets_to_indices(self._h, _stream(), C.c_int64(row_offsets.numel() - 1),
                                                      _ptr(row_offsets), _ptr(row_indices)))

    def indices_to_offsets(self, row_indices, row_offsets):
        self.check(self.lib.b200sp_indices_to_offsets(self._h, _stream(), C.c_int64(row_offsets.numel() - 1),
                                                      C.c_int64(row_indices.numel()), _ptr(row_indices),
                                                      _ptr(row_offsets)))

    def csr_convert_query(self, num_rows, num_cols, num_entries, row_offsets, column_indices=None,
                          relative_speed=3.0, breakeven_threshold=4096) -> "ConvertInfo":
        info = ConvertInfo()
        self.check(self.lib.b200sp_csr_convert_query(self._h, _stream(), C.c_int64(num_rows), C.c_int64(num_cols),
                                                     C.c_int64(num_entries), _ptr(row_offsets), _ptr(column_indices),
                                                     C.c_float(relative_speed), C.c_int64(breakeven_threshold),
                                                     C.byref(info)))
        return info

    def csr_to_ell(self, num_rows, K, pitch, row_offsets, column_indices, values, ell_cidx, ell_vals):
        f = getattr(self.lib, "b200sp_csr_to_ell_" + _sfx(values.dtype))
        self.check(f(self._h, _stream(), C.c_int64(num_rows), C.c_int64(K), C.c_int64(pitch), _ptr(row_offsets),
                     _ptr(column_indices), _ptr(values), _ptr(ell_cidx), _ptr(ell_vals)))

    # DIA / ELL / HYB -> CSR on the device: `_offsets` (row_offsets + the entry count), then `_fill`
    def dia_to_csr_offsets(self, num_rows, nd, pitch, values, row_offsets) -> int:
        n = C.c_int64(0)
        f = getattr(self.lib, "b200sp_dia_to_csr_offsets_" + _sfx(values.dtype))
        self.check(f(self._h, _stream(), C.c_int64(num_rows), C.c_int64(nd), C.c_int64(pitch), _ptr(values),
                     _ptr(row_offsets), C.byref(n)))
        return n.value

    def dia_to_csr_fill(self, num_rows, nd, pitch, offsets, values, row_offsets, column_indices, csr_values):
        f = getattr(self.lib, "b200sp_dia_to_csr_fill_" + _sfx(values.dtype))
        self.check(f(self._h, _stream(), C.c_int64(num_rows), C.c_int64(nd), C.c_int64(pitch), _ptr(offsets), _ptr(values),
                     _ptr(row_offsets), _ptr(column_indices), _ptr(csr_values)))

    def ell_to_csr_offsets(self, num_rows, K, pitch, cidx, values, row_offsets) -> int:
        n = C.c_int64(0)
        f = getattr(self.lib, "b200sp_ell_to_csr_offsets_" + _sfx(values.dtype))
        self.check(f(self._h, _stream(), C.c_int64(num_rows), C.c_int64(K), C.c_int64(pitch), _ptr(cidx), _ptr(values),
                     _ptr(row_offsets), C.byref(n)))
        return n.value

    def ell_to_csr_fill(self, num_rows, K, pitch, cidx, values, row_offsets, column_indices, csr_values):
        f = getattr(self.lib, "b200sp_ell_to_csr_fill_" + _sfx(values.dtype))
        self.check(f(self._h, _stream(), C.c_int64(num_rows), C.c_int64(K), C.c_int64(pitch), _ptr(cidx), _ptr(values),
                     _ptr(row_offsets), _ptr(column_indices), _ptr(csr_values)))

    def hyb_to_csr_offsets(self, num_rows, K, pitch, cidx, values, cnnz, coo_ri, row_offsets) -> int:
        n = C.c_int64(0)
        f = getattr(self.lib, "b200sp_hyb_to_csr_offsets_" + _sfx(values.dtype))
        self.check(f(self._h, _stream(), C.c_int64(num_rows), C.c_int64(K), C.c_int64(pitch), _ptr(cidx), _ptr(values),
                     C.c_int64(cnnz), _ptr(coo_ri), _ptr(row_offsets), C.byref(n)))
        return n.value

    def hyb_to_csr_fill(self, num_rows, K, pitch, cidx, values, cnnz, coo_ri, coo_ci, coo_v, row_offsets, column_indices,
                        csr_values):
        f = getattr(self.lib, "b200sp_hyb_to_csr_fill_" + _sfx(values.dtype))
        self.check(f(self._h, _stream(), C.c_int64(num_rows), C.c_int64(K), C.c_int64(pitch), _ptr(cidx), _ptr(values),
                     C.c_int64(cnnz), _ptr(coo_ri), _ptr(coo_ci), _ptr(coo_v), _ptr(row_offsets), _ptr(column_indices),
                     _ptr(csr_values)))

    def dia_to_ell(self, num_rows, nd, pitch, offsets, values, ell_cidx, ell_vals):
        f = getattr(self.lib, "b200sp_dia_to_ell_" + _sfx(values.dtype))
        self.check(f(self._h, _stream(), C.c_int64(num_rows), C.c_int64(nd), C.c_int64(pitch), _ptr(offsets), _ptr(values),
                     _ptr(ell_cidx), _ptr(ell_vals)))

    def csr_to_coo_tail(self, num_rows, K, row_offsets, column_indices, values, coo_ri, coo_ci, coo_v):
        f = getattr(self.lib, "b200sp_csr_to_coo_tail_" + _sfx(values.dtype))
        self.check(f(self._h, _stream(), C.c_int64(num_rows), C.c_int64(K), _ptr(row_offsets), _ptr(column_indices),
                     _ptr(values), _ptr(coo_ri), _ptr(coo_ci), _ptr(coo_v)))

    def csr_to_dia(self, num_rows, num_cols, num_diagonals, pitch, row_offsets, column_indices, values,
                   diagonal_offsets, dia_values):
        f = getattr(self.lib, "b200sp_csr_to_dia_" + _sfx(values.dtype))
        self.check(f(self._h, _stream(), C.c_int64(num_rows), C.c_int64(num_cols), C.c_int64(num_diagonals),
                     C.c_int64(pitch), _ptr(row_offsets), _ptr(column_indices), _ptr(values),
                     _ptr(diagonal_offsets), _ptr(dia_values)))

    def count_zeros(self, values) -> int:
        f = getattr(self.lib, "b200sp_count_zeros_" + _sfx(values.dtype))
        out = C.c_int64()
        self.check(f(self._h, _stream(), C.c_int64(values.numel()), _ptr(values), C.byref(out)))
        return out.value

    def poisson_csr_offsets(self, stencil, nx, ny, nz, row_begin, num_rows, row_offsets):
        self.check(self.lib.b200sp_poisson_csr_offsets(self._h, _stream(), C.c_int(stencil), C.c_int64(nx),
                                                       C.c_int64(ny), C.c_int64(nz), C.c_int64(row_begin),
                                                       C.c_int64(num_rows), _ptr(row_offsets)))

    def poisson_csr(self, stencil, nx, ny, nz, row_begin, num_rows, col_shift, row_offsets, cidx, values):
        f = getattr(self.lib, "b200sp_poisson_csr_" + _sfx(values.dtype))
        self.check(f(self._h, _stream(), C.c_int(stencil), C.c_int64(nx), C.c_int64(ny), C.c_int64(nz),
                     C.c_int64(row_begin), C.c_int64(num_rows), C.c_int64(col_shift), _ptr(row_offsets),
                     _ptr(cidx), _ptr(values)))


def poisson_num_entries(stencil, nx, ny, nz, row_begin, num_rows) -> int:
    return int(load_library().b200sp_poisson_num_entries(stencil, nx, ny, nz, row_begin, num_rows))
