"""Launches each hot kernel a few times on the BASELINE workloads (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cusp_autotuned_b200 as cusp
from cusp_autotuned_b200 import capi, gallery, convert

which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["dia", "ell", "csr", "coo"]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
h = cusp.default_handle()
dev = torch.device("cuda", 0)
n = 256
for fmt in which:
    if fmt == "cg":  # a few CG iterations on the headline DIA operator (fused-direction product + update kernel)
        from cusp_autotuned_b200 import krylov as K
        A = gallery.poisson("dia", 7, (n, n, n), dtype=torch.float64)
        b = torch.ones(A.num_rows, dtype=torch.float64, device=dev)
        xk = torch.zeros_like(b)
        K.cg(A, xk, b, cusp.monitor(b, 4, 0.0), check_interval=4)
        torch.cuda.synchronize()
        del A, b, xk
        continue
    if fmt.split(":")[0] == "spmm":  # "spmm:k[:f64]": CSR x dense block on the stencil operator
        parts = fmt.split(":")
        k = int(parts[1]) if len(parts) > 1 else 32
        dt = torch.float64 if (len(parts) > 2 and parts[2] == "f64") else torch.float32
        A = gallery.poisson("csr", 7, (n, n, n), dtype=dt)
        X = torch.rand(A.num_cols, k, dtype=dt, device=dev) + 0.5
        Y = torch.empty(A.num_rows, k, dtype=dt, device=dev)
        for _ in range(reps):
            cusp.multiply_block(A, X, Y)
        torch.cuda.synchronize()
        del A, X, Y
        continue
    if fmt.split(":")[0] in ("dia", "ell", "csr"):
        A = gallery.poisson(fmt, 7, (n, n, n), dtype=torch.float64)
        x = ((torch.arange(A.num_cols, device=dev) % 21) - 10).double()
    elif fmt == "coop":  # COO on the stencil operator: coalesced gathers, isolates the segmented scan
        from cusp_autotuned_b200.matrix import coo_matrix
        C = gallery.poisson("csr", 7, (n, n, n), dtype=torch.float64)
        lens = (C.row_offsets[1:] - C.row_offsets[:-1]).to(torch.int64)
        ri = torch.repeat_interleave(torch.arange(C.num_rows, device=dev, dtype=torch.int32), lens)
        A = coo_matrix(C.num_rows, C.num_cols, ri, C.column_indices, C.values)
        x = ((torch.arange(A.num_cols, device=dev) % 21) - 10).double()
    else:  # "coo": default kernel; "coow[:vw:u]": K_COO_WARP; "plan[:vw:u]": hot-column plan executor
        A = convert.rmat(int(os.environ.get("PROF_RMAT_SCALE", "22")), 16, seed=42, dtype=torch.float32)
        x = torch.rand(A.num_cols, device=dev) + 0.5
    y = torch.empty(A.num_rows, dtype=x.dtype, device=dev)
    d = A.descriptor()
    cfg = None
    parts = fmt.split(":")
    if parts[0] in ("coow", "plan"):
        vw, u = (int(v) for v in (parts[1:] + ["8", "1"][len(parts) - 1:])[:2])
        cfg = capi.Cfg(kernel=capi.K_COO_WARP, vector_width=vw, unroll=u)
    if parts[0] == "plan":
        plan = h.coo_plan_create(A.num_rows, A.num_cols, A.num_entries, A.row_indices, A.column_indices, capi.F32, 0)
        for _ in range(reps):
            h.spmv_coo_plan(plan, A.values, x, y, cfg=cfg)
        torch.cuda.synchronize()
        h.coo_plan_destroy(plan)
    else:
        for _ in range(reps):
            h.spmv(d, x, y, cfg=cfg)
    torch.cuda.synchronize()
    del A, x, y
print("done")
