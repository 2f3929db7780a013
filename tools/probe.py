"""Scratch perf probe (not the bench): times every SpMV format/kernel variant on
poisson7pt n^3 and prints achieved GB/s on compulsory bytes."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cusp_autotuned_b200 as cusp
from cusp_autotuned_b200 import capi, gallery

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
fmts = sys.argv[2].split(",") if len(sys.argv) > 2 else ["dia", "ell", "csr"]
reps = 20
dev = torch.device("cuda", 0)
h = cusp.default_handle()
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

def timeit(fn):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts)//2], ts[0]

def comp_bytes(fmt, A, es):
    r, c = A.num_rows, A.num_cols
    if fmt == "dia": return A.num_diagonals*A.pitch*es + A.num_diagonals*4 + c*es + r*es
    if fmt == "ell": return A.num_cols_per_row*A.pitch*(4+es) + c*es + r*es
    if fmt == "csr": return (r+1)*4 + A.num_entries*(4+es) + c*es + r*es
    if fmt == "coo": return A.num_entries*(8+es) + c*es + r*es

out = []
for dtype in (torch.float32, torch.float64):
    es = 4 if dtype == torch.float32 else 8
    for fmt in fmts:
        A = gallery.poisson7pt(n, n, n, fmt=fmt, dtype=dtype)
        x = torch.rand(A.num_cols, dtype=dtype, device=dev) + 0.5
        y = torch.empty(A.num_rows, dtype=dtype, device=dev)
        d = A.descriptor()
        fid = {"dia": capi.FMT_DIA, "ell": capi.FMT_ELL, "csr": capi.FMT_CSR}[fmt]
        B = comp_bytes(fmt, A, es)
        yref = None
        for cfg in capi.Handle.cfg_space(fid, 0):
            try:
                med, best = timeit(lambda: h.spmv(d, x, y, cfg=cfg))
            except capi.B200spError as e:
                print(fmt, dtype, cfg, "ERR", e); continue
            if yref is None: yref = y.clone()
            ok = bool(torch.equal(y, yref)) if fmt != "csr" else bool(torch.allclose(y, yref, rtol=1e-5 if es == 4 else 1e-12))
            rec = dict(fmt=fmt, dtype=str(dtype), cfg=cfg.as_dict(), ms=med, ms_min=best, gbs=B/med/1e6, same=ok)
            out.append(rec)
            print(f"{fmt} {es*8} k={cfg.kernel} b={cfg.block_size} tpr={cfg.threads_per_row} u={cfg.unroll} st={cfg.stages} cps={cfg.ctas_per_sm}: {med:.4f} ms  {B/med/1e6:8.1f} GB/s same={ok}", flush=True)
        del A, x, y
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/probe_{n}_{'_'.join(fmts)}.json", "w"))
best = {}
for r in out:
    k = (r["fmt"], r["dtype"])
    if k not in best or r["gbs"] > best[k]["gbs"]: best[k] = r
for k, r in best.items(): print("BEST", k, r["cfg"], f'{r["gbs"]:.0f} GB/s')
