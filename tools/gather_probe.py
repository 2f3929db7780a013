"""How fast can ANY kernel consume a scattered column stream?  (VERDICT r1 task 2: "a hand-written bare-gather
kernel within 5 % of the SpMV".)

Times tools/probe_kernels.cu's gather_sum kernels — the entry streams of a sparse product read exactly like
K_COO_WARP reads them (128/256-bit streaming loads), x gathered through ld.global.nc, products added in a register,
one store per lane; no segmented scan, no row logic — beside the engine's product on the same arrays:

  R-MAT scale S (default 24) COO fp32        engine default, engine planned, probe {cols | cols+vals | rows+cols+vals}
  random CSR 2^20 rows x k in {4, 32, 256}   engine default, probe cols+vals

  python tools/gather_probe.py [scale]   -> gpurun_out/gather_probe.json + a table on stdout
Median of REPS launches, CUDA events, 512 MiB L2 flush between launches.
"""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import cusp_autotuned_b200 as cusp
from cusp_autotuned_b200 import capi, convert
from cusp_autotuned_b200.matrix import csr_matrix

dev = torch.device("cuda", 0)
h = cusp.default_handle()
lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "build", "libprobe.so"))
lib.probe_gather_sum.restype = ctypes.c_int
lib.probe_gather_sum.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_longlong] + [ctypes.c_void_p] * 6
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
REPS = int(os.environ.get("PROBE_REPS", "9"))
SM_CLK = 148 * 1.965e9


WARM = os.environ.get("PROBE_WARM", "0") == "1"  # 1: no L2 flush between launches (bench.py's regime: x stays in L2)


def timeit(fn):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(REPS):
        if not WARM:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def burst(fn, n=20):
    """bench.py's regime: n back-to-back launches between one pair of events"""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def probe(variant, vpl, nnz, Ai, Aj, Ax, x, out):
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.probe_gather_sum(variant, vpl, nnz, Ai.data_ptr() if Ai is not None else None, Aj.data_ptr(), Ax.data_ptr(),
                              x.data_ptr(), out.data_ptr(), st)
    assert rc == 0, rc


def main():
    scale = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    res = []

    def rec(label, ms, nnz, extra=None):
        r = dict(label=label, ms=round(ms, 4), gathers_per_clk_per_sm=round(nnz / (ms * 1e-3) / SM_CLK, 3))
        r.update(extra or {})
        res.append(r)
        print(f"## {label}: {ms:.4f} ms  {r['gathers_per_clk_per_sm']:.2f} gathers/clk/SM" + (f"  burst {r['ms_burst']:.4f} ms" if "ms_burst" in r else ""), flush=True)

    C = convert.rmat(scale, 16, seed=42, dtype=torch.float32)
    nnz = C.num_entries
    x = torch.rand(C.num_cols, dtype=torch.float32, device=dev) + 0.5
    y = torch.empty(C.num_rows, dtype=torch.float32, device=dev)
    out = torch.empty(nnz // 4 + 1, dtype=torch.float32, device=dev)
    d = C.descriptor()
    ms_engine = timeit(lambda: h.spmv(d, x, y))
    rec(f"rmat s{scale} coo f32: engine default", ms_engine, nnz, dict(ms_burst=round(burst(lambda: h.spmv(d, x, y)), 4)))
    for vw, u, cps in ((8, 1, 0), (8, 2, 0), (8, 2, 8), (4, 1, 0), (4, 2, 0), (8, 4, 0)):
        cfg = capi.Cfg(kernel=capi.K_COO_WARP, vector_width=vw, unroll=u, ctas_per_sm=cps)
        rec(f"rmat s{scale} coo f32: engine K_COO_WARP v{vw} u{u} ctas/SM {cps or 'one-shot'}", timeit(lambda: h.spmv(d, x, y, cfg=cfg)), nnz,
            dict(ms_burst=round(burst(lambda: h.spmv(d, x, y, cfg=cfg)), 4)))
    plan = h.coo_plan_create(C.num_rows, C.num_cols, nnz, C.row_indices, C.column_indices, capi.F32, 0)
    ms_plan = timeit(lambda: h.spmv_coo_plan(plan, C.values, x, y))
    rec(f"rmat s{scale} coo f32: engine planned (hot-column table)", ms_plan, nnz)
    h.coo_plan_destroy(plan)
    for vpl in (8, 4):
        for variant, name in ((0, "columns only"), (1, "columns + values"), (2, "rows + columns + values")):
            ms = timeit(lambda: probe(variant, vpl, nnz, C.row_indices, C.column_indices, C.values, x, out))
            rec(f"rmat s{scale}: bare gather v{vpl}, {name}", ms, nnz, dict(vs_engine=round(ms_engine / ms, 3)))
    # the sum of a bare gather equals the sum of all products: sanity, not parity
    probe(2, 8, nnz, C.row_indices, C.column_indices, C.values, x, out)
    s_probe = out[: nnz // 8].double().sum().item()
    s_ref = (C.values[: nnz // 8 * 8].double() * x[C.column_indices[: nnz // 8 * 8].long()].double()).sum().item()
    assert abs(s_probe - s_ref) <= 1e-6 * abs(s_ref), (s_probe, s_ref)
    del C, x, y, out

    rows = 1 << 20
    for k in (4, 32, 256):
        g = torch.Generator(device=dev)
        g.manual_seed(k)
        cols = torch.randint(0, rows, (rows, k), generator=g, device=dev, dtype=torch.int32)
        cols, _ = torch.sort(cols, dim=1)
        Ap = (torch.arange(rows + 1, device=dev, dtype=torch.int64) * k).to(torch.int32)
        vals = torch.ones(rows * k, dtype=torch.float32, device=dev)
        A = csr_matrix(rows, rows, Ap, cols.reshape(-1).contiguous(), vals)
        x = torch.rand(rows, dtype=torch.float32, device=dev) + 0.5
        y = torch.empty(rows, dtype=torch.float32, device=dev)
        out = torch.empty(rows * k // 4 + 1, dtype=torch.float32, device=dev)
        dd = A.descriptor()
        ms_e = timeit(lambda: h.spmv(dd, x, y))
        rec(f"random csr 2^20 x {k} f32: engine default", ms_e, rows * k)
        for vpl in (8, 4):
            ms = timeit(lambda: probe(1, vpl, rows * k, None, A.column_indices, A.values, x, out))
            rec(f"random csr 2^20 x {k}: bare gather v{vpl}, columns + values", ms, rows * k, dict(vs_engine=round(ms_e / ms, 3)))
        del A, x, y, out, cols, vals
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/gather_probe%s.json" % ("_warm" if WARM else ""), "w"), indent=1)


if __name__ == "__main__":
    main()
