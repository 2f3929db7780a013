"""Times the REFERENCE's own CUDA kernel on this GPU ("the existing kernel to beat"): every point of
the KTT DIA tuning space (cusp/system/cuda/ktt/dia_multiply.h:24-55), kernel source cusp/system/cuda/ktt/kernels/
dia_kernel.h compiled unmodified for sm_100a into oracle/_ref/libcuspref_gpu.so (oracle/Makefile), on the headline
workload (poisson7pt 256^3 DIA, fp64 and fp32, x_i = (i mod 21) - 10).  Every configuration's y is checked against
the engine's y (1e-12 / 1e-5 relative to sum_j |a_ij x_j|: the reference build contracts to FMA).  Prints one JSON
line.  Measurement infrastructure: bench.py runs it in a separate process so that nothing it does can touch the
engine's own numbers.

  python tools/ref_gpu_kernels.py [n=256] [reps=20]
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import cusp_autotuned_b200 as cusp
    from cusp_autotuned_b200 import gallery

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    path = os.path.join(ROOT, "oracle", "_ref", "libcuspref_gpu.so")
    if not os.path.exists(path):
        print(json.dumps({"unavailable": "oracle/_ref/libcuspref_gpu.so not built (needs /root/reference at build time)"}))
        return 0
    lib = C.CDLL(path)
    lib.cuspref_dia_spmv.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p]
    ncfg = lib.cuspref_dia_num_configs()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    h = cusp.default_handle()
    out = {"workload": f"poisson7pt {n}^3 DIA, y = A x", "kernel": "ktt_dia_vector_kernel (cusp/system/cuda/ktt/kernels/dia_kernel.h), "
           "unmodified, nvcc -O3 sm_100a", "configs": ncfg, "reps": reps, "dtypes": {}}
    for tdt, name, tol in ((torch.float64, "f64", 1e-12), (torch.float32, "f32", 1e-5)):
        A = gallery.poisson7pt(n, n, n, fmt="dia", dtype=tdt)
        es = 8 if tdt == torch.float64 else 4
        x = ((torch.arange(A.num_cols, device=dev) % 21) - 10).to(tdt)
        y_ours = torch.empty(A.num_rows, dtype=tdt, device=dev)
        y = torch.empty(A.num_rows, dtype=tdt, device=dev)
        d = A.descriptor()
        h.spmv(d, x, y_ours)
        scale = torch.empty_like(y_ours)
        Aabs = gallery.poisson7pt(n, n, n, fmt="dia", dtype=tdt)
        Aabs.values.abs_()
        h.spmv(Aabs.descriptor(), x.abs(), scale)
        del Aabs
        B = A.num_diagonals * A.pitch * es + A.num_diagonals * 4 + A.num_cols * es + A.num_rows * es
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

        def run(cfg):
            return lib.cuspref_dia_spmv(cfg, int(tdt == torch.float64), A.num_rows, A.num_cols, A.num_diagonals, A.pitch,
                                        A.diagonal_offsets.data_ptr(), A.values.data_ptr(), x.data_ptr(), y.data_ptr(), st)

        def timed(fn, k):
            fn()
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(k):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / k

        ours_ms = timed(lambda: h.spmv(d, x, y_ours), reps)
        recs = []
        for cfg in range(ncfg):
            p = (C.c_int * 4)()
            lib.cuspref_dia_config(cfg, p)
            y.fill_(float("nan"))
            rc = run(cfg)
            torch.cuda.synchronize()
            if rc != 0:
                recs.append({"cfg": list(p), "error": f"cudaError {rc}"})
                continue
            err = float(((y - y_ours).abs() / scale.clamp_min(1e-30)).max().item())
            ms = timed(lambda: run(cfg), reps)
            recs.append({"cfg": {"BLOCK_SIZE": p[0], "PREFETCH_FACTOR": p[1], "PREFETCH_TYPE": p[2], "SPECIAL_LOADS": p[3]},
                         "ms": ms, "gbs": B / ms / 1e6, "scaled_err_vs_engine": err, "ok": bool(err <= tol)})
        good = [r for r in recs if r.get("ok")]
        best = min(good, key=lambda r: r["ms"]) if good else None
        first = recs[0] if recs and recs[0].get("ok") else None  # KTT's first point: 128 / no prefetch / plain loads
        out["dtypes"][name] = {"bytes": B, "engine_ms": ours_ms, "engine_gbs": B / ours_ms / 1e6,
                               "reference_best": best, "reference_first_config": first,
                               "engine_speedup_over_reference_best": (best["ms"] / ours_ms) if best else None,
                               "valid_configs": len(good), "all": recs}
        del A, x, y, y_ours, scale
    coo_mode = os.environ.get("REF_COO", "top")  # full: every point; top: the fastest points of the committed full sweep; off
    if coo_mode != "off":
        out["coo"] = coo_leg(torch, cusp, h, dev, reps, coo_mode)
    print(json.dumps(out))
    return 0


def coo_leg(torch, cusp, h, dev, reps, mode):
    """configs[2]: the reference's KTT COO kernels (cusp/system/cuda/ktt/kernels/coo_kernel.h, unmodified, every
    compilable point of coo_multiply.h:22-54) on R-MAT scale 24 fp32 beside the engine's default and planned products.
    A reference product = its launcher's two kernels (zero_output + coo_spmv)."""
    from cusp_autotuned_b200 import capi, convert
    path = os.path.join(ROOT, "oracle", "_ref", "libcuspref_gpu_coo.so")
    if not os.path.exists(path):
        return {"unavailable": "oracle/_ref/libcuspref_gpu_coo.so not built"}
    lib = C.CDLL(path)
    lib.cuspref_coo_spmv_f32.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_void_p]
    scale_log2 = int(os.environ.get("REF_COO_RMAT_SCALE", "24"))
    A = convert.rmat(scale_log2, 16, seed=42, dtype=torch.float32)
    n, nnz = A.num_rows, A.num_entries
    x = torch.rand(n, dtype=torch.float32, device=dev) + 0.5
    y_ours = torch.empty(n, dtype=torch.float32, device=dev)
    y = torch.empty(n, dtype=torch.float32, device=dev)
    d = A.descriptor()
    h.spmv(d, x, y_ours)  # positive data: sum |a x| = y
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    B = nnz * 12 + 2 * n * 4

    def timed(fn, k):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / k

    k = max(3, reps // 4)
    ours_ms = timed(lambda: h.spmv(d, x, y_ours), k)
    plan = h.coo_plan_create(n, n, nnz, A.row_indices, A.column_indices, capi.F32, 0)
    h.coo_plan_attach(plan)
    planned_ms = timed(lambda: h.spmv(d, x, y), k)
    planned_ok = bool(torch.equal(y, y_ours) or float(((y - y_ours).abs() / y_ours.clamp_min(1e-30)).max().item()) <= 1e-5)
    h.coo_plan_detach(plan)
    h.coo_plan_destroy(plan)
    ncfg = lib.cuspref_coo_num_configs()
    names = ("BLOCK_SIZE", "VALUES_PER_THREAD", "IMPL", "USE_CARRY", "AVOID_ATOMIC", "SPECIAL_LOADS")
    # mode "top" (what bench.py runs, bounded time): only the 12 fastest points of the committed full sweep
    # (profiles/r03_ref_coo_kernels.json, made by REF_COO=full); without that file every 16th point
    selected = None
    if mode == "top":
        try:
            sweep = json.load(open(os.path.join(ROOT, "profiles", "r03_ref_coo_kernels.json")))
            selected = [tuple(r["cfg"][k] for k in names) for r in sweep["coo"]["reference_top"]][:12]
        except Exception:
            selected = None
    recs = []
    for cfg in range(ncfg):
        p = (C.c_int * 6)()
        lib.cuspref_coo_config(cfg, p)
        if mode == "top" and ((selected is not None and tuple(p) not in selected) or (selected is None and cfg % 16)):
            continue

        def run():
            return lib.cuspref_coo_spmv_f32(cfg, A.row_indices.data_ptr(), A.column_indices.data_ptr(), A.values.data_ptr(), nnz,
                                            x.data_ptr(), y.data_ptr(), n, st)
        y.fill_(float("nan"))
        rc = run()
        torch.cuda.synchronize()
        rec = {"cfg": dict(zip(names, list(p)))}
        if rc != 0:
            rec["error"] = f"cudaError {rc}"
            recs.append(rec)
            continue
        err = float(((y - y_ours).abs() / y_ours.clamp_min(1e-30)).max().item())
        rec["scaled_err_vs_engine"] = err
        rec["ok"] = bool(err <= 1e-4)  # atomics / regrouped fp32 sums over hub rows of 10^5 entries
        if rec["ok"]:
            rec["ms"] = timed(run, 3)
        recs.append(rec)
    good = [r for r in recs if r.get("ok")]
    best = min(good, key=lambda r: r["ms"]) if good else None
    good.sort(key=lambda r: r["ms"])
    return {"workload": f"R-MAT scale {scale_log2} ef 16 fp32 COO, y = A x", "bytes": B, "configs": ncfg, "valid_configs": len(good),
            "engine_default_ms": ours_ms, "engine_planned_ms": planned_ms, "engine_planned_matches": planned_ok,
            "mode": mode, "timed_configs": len(recs), "reference_best": best, "reference_top": good[:16],
            "engine_speedup_over_reference_best": (best["ms"] / ours_ms) if best else None,
            "planned_speedup_over_reference_best": (best["ms"] / planned_ms) if best else None,
            "invalid": [r for r in recs if not r.get("ok")][:8]}


if __name__ == "__main__":
    sys.exit(main())
