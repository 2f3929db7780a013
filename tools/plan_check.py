"""First hardware check of the EXPERIMENTAL inspector / executor COO product (b200sp_coo_plan_*,
csrc/spmv_coo_plan.cu): parity against the oracle on skewed matrices (integer data bit-exact, real data within the
north-star tolerance, assign and accumulate, several table sizes, fp32 / fp64), plan statistics against the column
histogram, and a timing of the plan path beside the default COO kernel on R-MAT scale 22.  Prints one JSON line;
exit status 0 iff every parity case holds.  Run in its own process by tests/test_zz_plan_gpu.py so that nothing this
not-yet-validated path does can disturb the rest of the GPU suite.

  python tools/plan_check.py [rmat_scale=22]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

import cusp_autotuned_b200 as cusp
from cusp_autotuned_b200 import capi, convert
from helpers import TOL, abs_matrix, scaled_err, tdev
from oracle import oracle as O


def skewed_coo(rng, n, nnz, ndt, integer):
    """R-MAT-like skew: rows and columns drawn with P(bit = 0) = 0.76 per level, sorted by (row, col)"""
    bits = int(np.ceil(np.log2(n)))

    def draw(m):
        v = np.zeros(m, np.int64)
        for _ in range(bits):
            v = (v << 1) | (rng.random(m) >= 0.76).astype(np.int64)
        return v % n
    rows, cols = draw(nnz), draw(nnz)
    order = np.lexsort((cols, rows))
    rows, cols = rows[order].astype(np.int32), cols[order].astype(np.int32)
    vals = (rng.integers(1, 4, nnz) if integer else rng.uniform(0.5, 1.5, nnz)).astype(ndt)
    return dict(format="coo", num_rows=n, num_cols=n, num_entries=nnz, row_indices=rows, column_indices=cols, values=vals)


def main():
    scale = int(sys.argv[1]) if len(sys.argv) > 1 else 22
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    h = cusp.default_handle()
    rng = np.random.default_rng(31)
    out = {"cases": [], "ok": True}
    for ndt, dt in ((np.float32, capi.F32), (np.float64, capi.F64)):
        for n, nnz, table in ((5000, 7168 * 3 + 5, 0), (70000, 7168 * 40 + 1, 4096), (300, 17, 0), (40000, 7168 * 200, 64 << 10)):
            for integer in (True, False):
                A = skewed_coo(rng, n, nnz, ndt, integer)
                x = (rng.integers(-3, 4, n) if integer else rng.uniform(0.5, 1.5, n)).astype(ndt)
                y0 = (rng.integers(-5, 5, n) if integer else rng.uniform(-1, 1, n)).astype(ndt)
                Ai, Aj, Ax = tdev(A["row_indices"], dev), tdev(A["column_indices"], dev), tdev(A["values"], dev)
                plan = h.coo_plan_create(n, n, nnz, Ai, Aj, dt, table)
                rec = {"dtype": ndt.__name__, "n": n, "nnz": nnz, "table_bytes": table, "integer": integer}
                try:
                    info = h.coo_plan_info(plan)
                    top = np.sort(np.bincount(A["column_indices"], minlength=n))[::-1]
                    rec["info"] = info
                    rec["stats_ok"] = bool(0 < info["hot_columns"] <= info["capacity"]
                                           and info["hot_entries"] == int(top[: info["hot_columns"]].sum()))
                    good = rec["stats_ok"]
                    for acc in (False, True):
                        y = tdev(y0.copy(), dev)
                        h.spmv_coo_plan(plan, Ax, tdev(x, dev), y, accumulate=acc)
                        torch.cuda.synchronize()
                        got = y.cpu().numpy()
                        want = O.spmv(A, x, y0 if acc else None, accumulate=acc)
                        if integer:
                            okc = bool(np.array_equal(got, want))
                        else:
                            sc = O.spmv(abs_matrix(A), np.abs(x)) + (np.abs(y0) if acc else 0)
                            okc = bool(scaled_err(got, want, sc) <= TOL[np.dtype(ndt)])
                        rec["accumulate" if acc else "assign"] = okc
                        good = good and okc
                    rec["ok"] = good
                finally:
                    h.coo_plan_destroy(plan)
                out["cases"].append(rec)
                out["ok"] = out["ok"] and rec.get("ok", False)
    # timing beside the default kernel (information only): table sizes at `scale`, then the bench's scale 24
    def timed(fn):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 10

    out["timing"] = []
    out["timing_agree"] = True
    for sc, tables in ((scale, (32 << 10, 64 << 10, 128 << 10, 160 << 10)), (24, (128 << 10,))):
        try:
            C = convert.rmat(sc, 16, seed=42, dtype=torch.float32)
            x = torch.rand(C.num_cols, dtype=torch.float32, device=dev) + 0.5
            y = torch.empty(C.num_rows, dtype=torch.float32, device=dev)
            yp = torch.empty_like(y)
            t_def = timed(lambda: cusp.multiply(C, x, y))
            for tb in tables:
                plan = h.coo_plan_create(C.num_rows, C.num_cols, C.num_entries, C.row_indices, C.column_indices, capi.F32, tb)
                try:
                    t_plan = timed(lambda: h.spmv_coo_plan(plan, C.values, x, yp))
                    info = h.coo_plan_info(plan)
                    agree = bool(torch.allclose(y, yp, rtol=1e-4, atol=1e-4))
                    out["timing_agree"] = out["timing_agree"] and agree
                    out["timing"].append({"matrix": f"R-MAT scale {sc} ef 16 fp32", "nnz": int(C.num_entries),
                                          "table_bytes": tb, "default_ms": t_def, "plan_ms": t_plan,
                                          "hot_columns": info["hot_columns"],
                                          "gathers_served_by_table": info["hot_entries"] / C.num_entries, "agree": agree})
                finally:
                    h.coo_plan_destroy(plan)
            del C, x, y, yp
            torch.cuda.empty_cache()
        except Exception as ex:
            out["timing"].append({"scale": sc, "error": repr(ex)})
            out["timing_agree"] = False
    print(json.dumps(out), flush=True)
    return 0 if out["ok"] else 1


if __name__ == "__main__":
    sys.exit(main())
