"""Timing of the §8(f) rows on one B200 (CUDA events, inputs larger than L2): generalized products by functor code
(f2) beside the default triple on the same matrices, device conversions (f1) and the fused Krylov iterations (f3) on
poisson7pt 256^3.  Prints one JSON line; `profiles/r03_widen.json` is a copy of it."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import cusp_autotuned_b200 as cusp
from bench import compulsory_bytes
from cusp_autotuned_b200 import convert, gallery
from cusp_autotuned_b200 import krylov as K

PEAK = 6530.3
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
dev = torch.device("cuda", 0)
h = cusp.default_handle()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
out = {"grid": f"poisson7pt {n}^3", "peak_gbs": PEAK, "generalized": {}, "convert_ms": {}, "krylov": {}}


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for tdt, es in ((torch.float32, 4), (torch.float64, 8)):
    for fmt in ("dia", "ell", "csr", "coo", "hyb"):
        if fmt in ("coo", "hyb"):
            C0 = gallery.poisson("csr", 7, (n, n, n), dtype=tdt)
            A = convert.csr_to_coo(C0) if fmt == "coo" else convert.csr_to_hyb(C0)
            del C0
        else:
            A = gallery.poisson(fmt, 7, (n, n, n), dtype=tdt)
        x = ((torch.arange(A.num_cols, device=dev) % 21) - 10).to(tdt)
        y = torch.zeros(A.num_rows, dtype=tdt, device=dev)
        B = compulsory_bytes(A, es)
        d = A.descriptor()
        rec = {"default_ms": round(timed(lambda: h.spmv(d, x, y)), 5)}
        for name, kw in (("min_plus", dict(initialize="constant", init_value=float("inf"), combine="plus", reduce="minimum")),
                         ("max_times", dict(initialize="identity", combine="multiplies", reduce="maximum")),
                         ("plus_project2nd", dict(initialize="constant", init_value=0.0, combine="project2nd", reduce="plus"))):
            ms = timed(lambda: h.spmv_generalized(d, x, y, **kw))
            rec[name] = [round(ms, 5), round(B / ms / 1e6 / PEAK, 4)]
        rec["default_frac"] = round(B / rec["default_ms"] / 1e6 / PEAK, 4)
        out["generalized"][f"{fmt}_{'f32' if es == 4 else 'f64'}"] = rec
        del A, x, y
        torch.cuda.empty_cache()

# f1: device conversions at 256^3 fp64 (structure + values), wall clock around a synchronize
C = gallery.poisson("csr", 7, (n, n, n), dtype=torch.float64)
for name, fn in (("csr_to_coo", lambda: convert.csr_to_coo(C)), ("csr_to_ell", lambda: convert.csr_to_ell(C)),
                 ("csr_to_dia", lambda: convert.csr_to_dia(C)), ("csr_to_hyb", lambda: convert.csr_to_hyb(C))):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):  # the temporaries are cudaMalloc'd per call: take the fastest of five
        t = time.perf_counter()
        R = fn()
        torch.cuda.synchronize()
        best = min(best, (time.perf_counter() - t) * 1e3)
        del R
    out["convert_ms"][name] = round(best, 3)
D = convert.csr_to_dia(C)
E = convert.csr_to_ell(C)
H = convert.csr_to_hyb(C)
for name, fn in (("dia_to_csr", lambda: convert.dia_to_csr(D)), ("ell_to_csr", lambda: convert.ell_to_csr(E)),
                 ("hyb_to_csr", lambda: convert.hyb_to_csr(H)), ("dia_to_ell", lambda: convert.dia_to_ell(D))):
    fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    R = fn()
    torch.cuda.synchronize()
    out["convert_ms"][name] = round((time.perf_counter() - t) * 1e3, 3)
    del R
del C, E, H
torch.cuda.empty_cache()

# f3: fused Krylov iterations, DIA fp64, fixed 40 iterations (tolerance 0), Jacobi where the solver takes one
N = D.num_rows
b = torch.ones(N, dtype=torch.float64, device=dev)
M = K.diagonal(D)
Bspmv = compulsory_bytes(D, 8)
for solver, fn, vecs, prods in (("cg", K.cg, 9, 1), ("pcg_jacobi", K.pcg, 0, 1), ("bicgstab_jacobi", K.bicgstab, 0, 2),
                                ("cr_jacobi", K.cr, 0, 1)):
    its = 40
    for rep in range(2):
        xk = torch.zeros(N, dtype=torch.float64, device=dev)
        mon = cusp.monitor(b, its, 0.0, 0.0)
        torch.cuda.synchronize()
        t = time.perf_counter()
        if solver == "cg":
            fn(D, xk, b, mon, check_interval=its)
        else:
            fn(D, xk, b, mon, M, check_interval=its)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
    rec = {"ms_per_iter": round(dt / max(1, mon.iteration_count()) * 1e3, 4), "iterations": mon.iteration_count(),
           "products_per_iter": prods}
    if vecs:
        rec["frac"] = round((prods * Bspmv + vecs * N * 8) / (rec["ms_per_iter"] * 1e6) / PEAK, 4)
    out["krylov"][solver] = rec
print(json.dumps(out))
