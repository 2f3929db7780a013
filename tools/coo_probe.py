"""COO on R-MAT: every K_COO_WARP shape and the hot-column plan executor (table sizes) beside the
default kernels, with an exactness check per variant.  Writes gpurun_out/coo_probe_s<scale>.json.

  python tools/coo_probe.py [scale ...]        (default: 22 24)

Per variant: median of REPS launches (CUDA events, 512 MiB L2 flush between launches), GB/s on compulsory bytes,
`exact` = all-ones matrix times all-ones x equals the row degrees bit for bit, `err` = max relative deviation from
the segmented-scan kernel on U(0.5, 1.5) data (positive: no cancellation)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import cusp_autotuned_b200 as cusp
from cusp_autotuned_b200 import capi, convert

dev = torch.device("cuda", 0)
h = cusp.default_handle()
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
REPS = int(os.environ.get("PROBE_REPS", "7"))


def timeit(fn):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(REPS):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    scales = [int(s) for s in sys.argv[1:]] or [22, 24]
    for scale in scales:
        C = convert.rmat(scale, 16, seed=42, dtype=torch.float32)
        n, nnz = C.num_rows, C.num_entries
        B = nnz * 12 + 2 * n * 4
        x = torch.rand(n, dtype=torch.float32, device=dev) + 0.5
        ones_v = torch.ones(nnz, dtype=torch.float32, device=dev)
        ones_x = torch.ones(n, dtype=torch.float32, device=dev)
        deg = torch.bincount(C.row_indices.to(torch.int64), minlength=n).to(torch.float32)
        y = torch.empty(n, dtype=torch.float32, device=dev)
        yref = torch.empty(n, dtype=torch.float32, device=dev)
        h.spmv_coo(n, n, nnz, C.row_indices, C.column_indices, C.values, x, yref,
                   cfg=capi.Cfg(kernel=capi.K_COO_SEGSCAN))
        scale_ref = yref.abs().clamp_min(1e-30)
        out = []

        def run(label, fn_vals, **extra):
            """fn_vals(values, x, y): launches the variant"""
            rec = dict(label=label, **extra)
            try:
                y.fill_(7.0)
                fn_vals(ones_v, ones_x, y)
                rec["exact"] = bool(torch.equal(y, deg))
                y.fill_(7.0)
                fn_vals(C.values, x, y)
                rec["err"] = float(((y - yref).abs() / scale_ref).max().item())
                # accumulate form on integer data
                y.fill_(3.0)
                fn_vals(ones_v, ones_x, y, True)
                rec["exact_acc"] = bool(torch.equal(y, deg + 3.0))
                ms = timeit(lambda: fn_vals(C.values, x, y))
                rec.update(ms=ms, gbs=B / ms / 1e6, frac=B / ms / 1e6 / 6530.3)
            except capi.B200spError as e:
                rec["error"] = str(e)[:200]
            out.append(rec)
            print(json.dumps(rec), flush=True)

        def coo(cfg):
            return lambda v, xx, yy, acc=False: h.spmv_coo(n, n, nnz, C.row_indices, C.column_indices, v, xx, yy,
                                                          accumulate=acc, cfg=cfg)

        run("default", coo(None))
        run("segscan 256x7", coo(capi.Cfg(kernel=capi.K_COO_SEGSCAN, block_size=256, unroll=7)))
        quick = os.environ.get("PROBE_QUICK") == "1"
        for vpl in (4, 8):
            for u in (1, 2, 4):
                for cps in ((0,) if quick else (0, 4, 8)):
                    cfg = capi.Cfg(kernel=capi.K_COO_WARP, vector_width=vpl, unroll=u, ctas_per_sm=cps)
                    run(f"warp v{vpl} u{u} cps{cps}", coo(cfg), vpl=vpl, u=u, cps=cps)
        # plan executor: table size x shape
        for tb in ((0,) if quick else (64 << 10, 96 << 10, 112 << 10, 128 << 10, 144 << 10, 160 << 10, 0)):
            try:
                plan = h.coo_plan_create(n, n, nnz, C.row_indices, C.column_indices, capi.F32, tb)
            except capi.B200spError as e:
                print("plan create failed:", e, flush=True)
                continue
            info = h.coo_plan_info(plan)
            info["hot_fraction"] = info["hot_entries"] / nnz
            for vpl, u in ((4, 1), (4, 2), (8, 1)):
                cfg = capi.Cfg(kernel=capi.K_COO_WARP, vector_width=vpl, unroll=u)
                run(f"plan tb{tb >> 10}K v{vpl} u{u}",
                    lambda v, xx, yy, acc=False, cfg=cfg: h.spmv_coo_plan(plan, v, xx, yy, accumulate=acc, cfg=cfg),
                    vpl=vpl, u=u, table_kib=tb >> 10, **info)
            # through the product entry point with the plan attached
            if tb == 0:
                h.coo_plan_attach(plan)
                run("attached plan via b200sp_spmv_coo", coo(None), **info)
                h.coo_plan_detach(plan)
            h.coo_plan_destroy(plan)
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(out, open(f"gpurun_out/coo_probe_s{scale}.json", "w"), indent=0)
        best = sorted((r for r in out if "ms" in r and r.get("exact") and r.get("exact_acc")), key=lambda r: r["ms"])[:8]
        print(f"## scale {scale}: nnz={nnz} bytes={B}")
        for r in best:
            print(f"   {r['label']:40s} {r['ms']:.4f} ms {r['gbs']:.0f} GB/s frac {r['frac']:.3f} err {r['err']:.1e}")
        del C, x, y, yref, ones_v, ones_x, deg


if __name__ == "__main__":
    main()
