"""HYB with a one-entry-per-row COO tail (poisson7pt 256^3 forced to K = 6): which COO kernel family serves the tail
best.  Times the whole HYB product per tail configuration; prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import cusp_autotuned_b200 as cusp
from bench import compulsory_bytes
from cusp_autotuned_b200 import capi, convert, gallery

PEAK = 6530.3
dev = torch.device("cuda", 0)
h = cusp.default_handle()
n = 256
out = {}
for tdt, es in ((torch.float32, 4), (torch.float64, 8)):
    H = convert.csr_to_hyb(gallery.poisson("csr", 7, (n, n, n), dtype=tdt), num_entries_per_row=6)
    x = ((torch.arange(H.num_cols, device=dev) % 21) - 10).to(tdt)
    y = torch.zeros(H.num_rows, dtype=tdt, device=dev)
    want = torch.zeros_like(y)
    h.spmv(H.descriptor(), x, want)
    B = compulsory_bytes(H, es)
    e, c = H.ell, H.coo
    rec = {"tail_nnz": int(c.num_entries)}
    cfgs = {"default": None,
            "segscan": capi.Cfg(kernel=capi.K_COO_SEGSCAN),
            "ring": capi.Cfg(kernel=capi.K_COO_RING),
            "warp_v8_u1": capi.Cfg(kernel=capi.K_COO_WARP, vector_width=8, unroll=1),
            "warp_v4_u1": capi.Cfg(kernel=capi.K_COO_WARP, vector_width=4, unroll=1),
            "warp_v4_u2": capi.Cfg(kernel=capi.K_COO_WARP, vector_width=4, unroll=2)}
    for name, cc in cfgs.items():
        def run():
            h.spmv_hyb(H.num_rows, H.num_cols, e.num_cols_per_row, e.pitch, e.column_indices, e.values, c.num_entries,
                       c.row_indices, c.column_indices, c.values, x, y, coo_cfg=cc)
        try:
            for _ in range(3):
                run()
            ok = bool(torch.equal(y, want))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            rec[name] = [round(ms, 5), round(B / ms / 1e6 / PEAK, 4), ok]
        except Exception as ex:
            rec[name] = repr(ex)[:120]
    out["f32" if es == 4 else "f64"] = rec
    del H, x, y, want
    torch.cuda.empty_cache()
if len(sys.argv) > 1 and sys.argv[1].startswith("rmat"):  # python tools/hyb_probe.py rmat24: the graph operator's HYB
    sc = int(sys.argv[1][4:] or 24)
    C = convert.rmat(sc, 16, seed=42, dtype=torch.float32, values="ones")
    H = convert.csr_to_hyb(convert.coo_to_csr(C))
    deg = torch.bincount(C.row_indices.long(), minlength=C.num_rows).float()
    del C
    x = torch.ones(H.num_cols, dtype=torch.float32, device=dev)
    y = torch.zeros(H.num_rows, dtype=torch.float32, device=dev)
    B = compulsory_bytes(H, 4)
    d = H.descriptor()
    for _ in range(3):
        h.spmv(d, x, y)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        h.spmv(d, x, y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    out[f"rmat_s{sc}"] = {"K": int(H.ell.num_cols_per_row), "tail_nnz": int(H.coo.num_entries), "ms": round(ms, 5),
                          "frac": round(B / ms / 1e6 / PEAK, 4), "exact": bool(torch.equal(y, deg)),
                          "fused_env": os.environ.get("B200SP_HYB_FUSED", "1")}
print(json.dumps(out))
