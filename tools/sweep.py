"""Tuning-space sweeps on one B200 (writes gpurun_out/sweep_<what>.json + a table on stdout).

  python tools/sweep.py csr256        CSR cfg space on poisson7pt 256^3, fp32+fp64
  python tools/sweep.py coo256        COO / HYB cfg spaces on poisson7pt 256^3 (coalesced gathers)
  python tools/sweep.py csr512sq      CSR cfg space on poisson5pt 512^2 (BASELINE configs[0], L2-resident)
  python tools/sweep.py rmat [scale]  CSR / COO / HYB on the R-MAT graph (power-law rows, random gathers)
  python tools/sweep.py random        BASELINE configs[3]: CSR space over random matrices, 2^20 rows,
                                      4..256 nnz/row (the csr_vector threads-per-row sweep of
                                      performance/csr_vector/csr_vector.cu:84-110 on the new space)

Median of `reps` launches, CUDA events, 512 MiB L2 flush between launches; GB/s on compulsory bytes.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import cusp_autotuned_b200 as cusp
from cusp_autotuned_b200 import capi, convert, gallery
from cusp_autotuned_b200.matrix import coo_matrix, csr_matrix

dev = torch.device("cuda", 0)
h = cusp.default_handle()
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
REPS = int(os.environ.get("SWEEP_REPS", "10"))
KNAME = {capi.FMT_CSR: {1: "vector", 2: "stream", 3: "ring", 4: "balanced"}, capi.FMT_COO: {1: "segscan", 2: "ring"},
         capi.FMT_HYB: {1: "ldg", 2: "bulk"}}


def timeit(fn):
    fn()
    torch.cuda.synchronize()
    ts = []
    inner = int(os.environ.get("SWEEP_WARM_LAUNCHES", "0"))  # > 0: L2-resident timing, mean of back-to-back launches
    for _ in range(REPS):
        if not inner:
            flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(max(1, inner)):
            fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) / max(1, inner))
    ts.sort()
    return ts[len(ts) // 2]


def comp_bytes(A, es):
    r, c, f = A.num_rows, A.num_cols, A.format
    if f == capi.FMT_CSR:
        return (r + 1) * 4 + A.num_entries * (4 + es) + c * es + r * es
    if f == capi.FMT_COO:
        return A.num_entries * (8 + es) + c * es + r * es
    if f == capi.FMT_HYB:
        e, co = A.ell, A.coo
        return e.num_cols_per_row * e.pitch * (4 + es) + co.num_entries * (8 + es) + c * es + r * es
    raise ValueError(f)


def csr_to_coo(A):
    lens = (A.row_offsets[1:] - A.row_offsets[:-1]).to(torch.int64)
    ri = torch.repeat_interleave(torch.arange(A.num_rows, device=dev, dtype=torch.int32), lens)
    return coo_matrix(A.num_rows, A.num_cols, ri, A.column_indices, A.values)


def sweep(label, A, x, space=None, out=None, check=True):
    es = x.element_size()
    y = torch.empty(A.num_rows, dtype=x.dtype, device=dev)
    d = A.descriptor()
    B = comp_bytes(A, es)
    # reference output: for COO the LDG kernel (the ring kernel must reproduce its bits)
    h.spmv(d, x, y, cfg=capi.Cfg(kernel=capi.K_COO_SEGSCAN) if A.format == capi.FMT_COO else None)
    yref = y.clone()
    scale = float(yref.abs().max().item()) or 1.0
    recs = []
    cfgs = [None] + list(space if space is not None else capi.Handle.cfg_space(A.format, 0))
    for cfg in cfgs:
        try:
            ms = timeit(lambda: h.spmv(d, x, y, cfg=cfg))
        except capi.B200spError as e:
            continue
        err = float((y - yref).abs().max().item()) / scale if check else 0.0
        rec = dict(label=label, cfg=(cfg.as_dict() if cfg else "default"), ms=ms, gbs=B / ms / 1e6, err=err)
        recs.append(rec)
    recs_named = [r for r in recs if r["cfg"] != "default"]
    dflt = [r for r in recs if r["cfg"] == "default"][0]
    print(f"## {label}: bytes={B}  default {dflt['ms']:.4f} ms {dflt['gbs']:.0f} GB/s", flush=True)
    fam = {}
    for r in recs_named:
        k = r["cfg"]["kernel"]
        if k not in fam or r["gbs"] > fam[k]["gbs"]:
            fam[k] = r
    for k, r in sorted(fam.items()):
        c = r["cfg"]
        print(f"   best {KNAME.get(A.format, {}).get(k, k)}: block={c['block_size']} tpr={c['threads_per_row']} "
              f"unroll={c['unroll']} stages={c['stages']} ctas/sm={c['ctas_per_sm']}  {r['ms']:.4f} ms "
              f"{r['gbs']:.0f} GB/s  maxerr={r['err']:.1e}", flush=True)
    if out is not None:
        out.extend(recs)
    return recs


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "csr256"
    out = []
    if what in ("csr256", "coo256"):
        n = int(os.environ.get("SWEEP_N", "256"))
        for dtype in (torch.float32, torch.float64):
            A = gallery.poisson7pt(n, n, n, fmt="csr", dtype=dtype)
            x = torch.rand(A.num_cols, dtype=dtype, device=dev) + 0.5
            if what == "csr256":
                sweep(f"csr poisson7pt {n}^3 {dtype}", A, x, out=out)
            else:
                C = csr_to_coo(A)
                sweep(f"coo poisson7pt {n}^3 {dtype}", C, x, out=out)
                H = convert.csr_to_hyb(A, num_entries_per_row=6)  # force a COO tail of one entry per interior row
                sweep(f"hyb(K=6) poisson7pt {n}^3 {dtype}", H, x, space=[], out=out)
            del A, x
    elif what == "csr512sq":  # BASELINE configs[0]: L2-resident, launch-latency territory
        for dtype in (torch.float32, torch.float64):
            A = gallery.poisson5pt(512, 512, fmt="csr", dtype=dtype)
            x = torch.rand(A.num_cols, dtype=dtype, device=dev) + 0.5
            sweep(f"csr poisson5pt 512^2 {dtype}", A, x, out=out)
    elif what == "gather":
        # how fast can this GPU gather 4-byte words at all?  (library kernel, not ours: torch.index_select)
        # R-MAT column stream vs uniformly random indices vs a sorted (coalescable) stream, 2^24-entry table
        C = convert.rmat(24, 16, seed=42, dtype=torch.float32)
        x = torch.rand(C.num_cols, dtype=torch.float32, device=dev)
        n = C.num_entries
        streams = {"rmat_columns": C.column_indices,
                   "uniform_random": torch.randint(0, C.num_cols, (n,), device=dev, dtype=torch.int32),
                   "sorted": torch.sort(torch.randint(0, C.num_cols, (n,), device=dev, dtype=torch.int32))[0]}
        for name, idx in streams.items():
            ms = timeit(lambda: torch.index_select(x, 0, idx))
            rec = dict(label=f"gather {name}", n=n, ms=ms, gathers_per_s=n / ms * 1e3,
                       gathers_per_clk_per_sm=n / (ms * 1e-3) / 148 / 1.965e9)
            out.append(rec)
            print(f"## gather {name}: {n} x 4 B from a 64 MiB table  {ms:.4f} ms  {rec['gathers_per_s'] / 1e9:.1f} G/s  "
                  f"{rec['gathers_per_clk_per_sm']:.2f} per clk per SM", flush=True)
    elif what == "rmat":
        scale = int(sys.argv[2]) if len(sys.argv) > 2 else 24
        C = convert.rmat(scale, 16, seed=42, dtype=torch.float32)
        x = torch.rand(C.num_cols, dtype=torch.float32, device=dev) + 0.5
        sweep(f"coo rmat s{scale}", C, x, out=out)
        if os.environ.get("SWEEP_ONLY") == "coo":
            json.dump(out, open(f"gpurun_out/sweep_{what}_coo.json", "w"))
            return
        A = convert.coo_to_csr(C)
        sweep(f"csr rmat s{scale}", A, x, out=out)
        H = convert.csr_to_hyb(A)
        sweep(f"hyb rmat s{scale} (K={H.ell.num_cols_per_row})", H, x, space=[], out=out)
    elif what == "spmm":  # CSR x dense block (cusp::multiply(csr, array2d, array2d)), poisson7pt 256^3
        n = int(os.environ.get("SWEEP_N", "256"))
        for dtype in (torch.float32, torch.float64):
            A = gallery.poisson7pt(n, n, n, fmt="csr", dtype=dtype)
            es = torch.empty(0, dtype=dtype).element_size()
            for k in (1, 2, 4, 8, 16, 32):
                X = torch.rand(A.num_cols, k, dtype=dtype, device=dev) + 0.5
                Y = torch.empty(A.num_rows, k, dtype=dtype, device=dev)
                ms = timeit(lambda: cusp.multiply_block(A, X, Y))
                B = (A.num_rows + 1) * 4 + A.num_entries * (4 + es) + (A.num_cols + A.num_rows) * k * es
                rec = dict(label=f"spmm csr poisson7pt {n}^3 {dtype} k={k}", ms=ms, gbs=B / ms / 1e6, bytes=B,
                           gflops=2.0 * A.num_entries * k / ms / 1e6)
                out.append(rec)
                print(f"## {rec['label']}: bytes={B}  {ms:.4f} ms  {rec['gbs']:.0f} GB/s  {rec['gflops']:.0f} GFLOP/s", flush=True)
                del X, Y
            del A
    elif what == "random":
        rows = 1 << 20
        for k in (4, 8, 16, 32, 64, 128, 256):
            g = torch.Generator(device=dev)
            g.manual_seed(k)
            cols = torch.randint(0, rows, (rows, k), generator=g, device=dev, dtype=torch.int32)
            cols, _ = torch.sort(cols, dim=1)
            Ap = (torch.arange(rows + 1, device=dev, dtype=torch.int64) * k).to(torch.int32)
            for dtype in (torch.float32, torch.float64):
                vals = torch.ones(rows * k, dtype=dtype, device=dev)
                A = csr_matrix(rows, rows, Ap, cols.reshape(-1).contiguous(), vals)
                x = torch.rand(rows, dtype=dtype, device=dev) + 0.5
                sweep(f"csr random 2^20 rows x {k} nnz/row {dtype}", A, x, out=out)
                del A, vals, x
            del cols
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/sweep_{what}.json", "w"))


if __name__ == "__main__":
    main()
