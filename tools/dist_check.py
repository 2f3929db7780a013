#!/usr/bin/env python
"""Multi-GPU parity check of the row-partitioned operator, one process
per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29511 tools/dist_check.py

Every rank builds its row block of poisson7pt on the device, then
  * b200sp_spmv_dist: the gathered y must be BIT-IDENTICAL to the oracle's
    single-domain y (each y entry is computed by exactly one GPU, same order);
  * b200sp_cg_dist: iteration count equal and residual history within 1e-10
    relative of the oracle's sequential CG (only the dot-product grouping differs).
Prints one JSON line on rank 0 and exits non-zero on a mismatch.  The oracle is
the checker only.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch
import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import cusp_autotuned_b200 as cusp  # noqa: E402
from cusp_autotuned_b200 import capi, dist, gallery  # noqa: E402
from cusp_autotuned_b200.partition import plane_partition  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    rank, world, local = dist.init_process_group_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    h = cusp.default_handle()
    dist.init_engine_comm(h, rank, world)
    out = {"world": world, "comm": "nvlink-p2p" if (world > 1 and h.comm_p2p_enabled()) else "nccl", "cases": []}
    ok = True
    for grid in ((24, 20, 16), (32, 32, 8 * world), (17, 13, world)):
        for tdt, ndt in ((torch.float64, np.float64), (torch.float32, np.float32)):
            for fmt in ("dia", "ell", "csr"):
                blk = plane_partition(grid, world, rank)
                A = gallery.poisson(fmt, 7, grid, dtype=tdt, row_begin=blk.row_begin, num_rows=blk.num_rows,
                                    halo_lo=blk.halo_lo, halo_hi=blk.halo_hi)
                ref = O.poisson(7, grid, ndt, "dia")
                n = ref["num_rows"]
                rng = np.random.default_rng(11)
                xg = rng.uniform(0.5, 1.5, n).astype(ndt)
                xw = torch.zeros(blk.window, dtype=tdt, device=dev)
                xw[blk.halo_lo: blk.halo_lo + blk.num_rows] = torch.from_numpy(
                    xg[blk.row_begin: blk.row_begin + blk.num_rows]).to(dev)
                y = torch.empty(blk.num_rows, dtype=tdt, device=dev)
                halo = capi.Halo(blk.halo_lo, blk.halo_hi)
                if world > 1:
                    h.spmv_dist(A.descriptor(), halo, xw, y)
                else:
                    h.spmv(A.descriptor(), xw, y)
                want = O.spmv(O.convert(ref, fmt), xg)[blk.row_begin: blk.row_begin + blk.num_rows]
                got = y.cpu().numpy()
                exact = bool(np.array_equal(got, want)) if fmt != "csr" else bool(
                    np.allclose(got, want, rtol=1e-5 if ndt == np.float32 else 1e-12, atol=1e-5 if ndt == np.float32 else 1e-12))
                # CG
                b = np.ones(n, ndt)
                xo, it_o, conv_o, hist_o = O.cg(O.convert(ref, "csr"), np.zeros_like(b), b, 40, 1e-6)
                xl = torch.zeros(blk.num_rows, dtype=tdt, device=dev)
                bl = torch.ones(blk.num_rows, dtype=tdt, device=dev)
                res, hist = h.cg(A.descriptor(), xl, bl, iteration_limit=40, relative_tolerance=1e-6,
                                 check_interval=4, halo=halo if world > 1 else None)
                # fp64: 1e-10 relative over the whole history; fp32: the regrouped dot products
                # perturb the iterates at 1e-7 and CG amplifies that, so the history is compared against the
                # initial residual (absolute) instead
                hist_o = np.asarray(hist_o, dtype=np.float64)
                same_it = int(res.iteration_count) == int(it_o)
                m = min(len(hist), len(hist_o))
                hist_dev = float(np.max(np.abs(hist[:m] - hist_o[:m]) / hist_o[0])) if m else 0.0
                if ndt == np.float64:
                    hist_ok = bool(np.allclose(hist[:m], hist_o[:m], rtol=1e-10, atol=0))
                else:
                    # fp32: rounding differences of the regrouped dots grow with the iteration; the
                    # first 10 entries must agree to 1e-3 of ||r0||, the whole history to 1e-2
                    hist_ok = bool(np.all(np.abs(hist[:min(m, 10)] - hist_o[:min(m, 10)]) <= 1e-3 * hist_o[0])) \
                        and hist_dev <= 1e-2
                hist_ok = hist_ok and len(hist) == len(hist_o)
                x_ok = bool(np.allclose(xl.cpu().numpy(), xo[blk.row_begin: blk.row_begin + blk.num_rows],
                                        rtol=1e-3 if ndt == np.float32 else 1e-9))
                # the other fused solvers, partitioned (b200sp_krylov with halo): Jacobi-CG, BiCGStab, CR
                kr_ok = 1
                if fmt == "dia" and ndt == np.float64:
                    csr_ref = O.convert(ref, "csr")
                    dinv_g = 1.0 / O.extract_diagonal(csr_ref)
                    dinv_l = torch.from_numpy(dinv_g[blk.row_begin: blk.row_begin + blk.num_rows]).to(dev)
                    # a generic right-hand side: with b = 1 on this symmetric stencil BiCGStab's <r*, A M p> nearly vanishes
                    # and the iteration is erratic in any arithmetic
                    b2 = np.random.default_rng(23).uniform(-1.0, 1.0, n)
                    bl2 = torch.from_numpy(b2[blk.row_begin: blk.row_begin + blk.num_rows]).to(dev)
                    for solver, jac in (("cg", True), ("bicgstab", False), ("bicgstab", True), ("cr", False), ("cr", True)):
                        xo2, it2, conv2, hist2 = O.krylov(solver, csr_ref, np.zeros(n), b2, 300, 1e-6, dinv=dinv_g if jac else None)
                        xl2 = torch.zeros(blk.num_rows, dtype=tdt, device=dev)
                        res2, h2 = h.krylov(solver, A.descriptor(), xl2, bl2, diagonal_inverse=dinv_l if jac else None,
                                            iteration_limit=300, relative_tolerance=1e-6, check_interval=4,
                                            halo=halo if world > 1 else None)
                        # leading 8 iterations entry by entry (BiCGStab amplifies the regrouped sums' rounding quickly),
                        # the same verdict and count, the same solution to the solve's accuracy
                        per = 2 if solver == "bicgstab" else 1
                        m2 = min(len(h2), len(hist2), 8 * per)
                        bi = solver == "bicgstab"
                        good = (abs(int(res2.iteration_count) - it2) <= (2 + it2 // 20 if bi else 1) and bool(res2.converged) == bool(conv2) and
                                bool(np.allclose(h2[:m2], hist2[:m2], rtol=1e-8, atol=0)) and
                                bool(np.abs(xl2.cpu().numpy() - xo2[blk.row_begin: blk.row_begin + blk.num_rows]).max()
                                     <= 1e-4 * np.abs(xo2).max()))
                        if not good:
                            out.setdefault("krylov_failures", []).append(
                                [list(grid), solver, jac, int(res2.iteration_count), it2,
                                 float(np.max(np.abs(h2[:m2] - hist2[:m2]) / hist2[:m2])) if m2 else -1.0])
                        kr_ok &= int(good)
                flags = torch.tensor([int(exact), int(same_it), int(hist_ok), int(x_ok) & kr_ok], device=dev)
                if world > 1:
                    td.all_reduce(flags, op=td.ReduceOp.MIN)
                f = [int(v) for v in flags.cpu().tolist()]
                out["cases"].append({"grid": list(grid), "dtype": ndt.__name__, "fmt": fmt, "spmv_exact": f[0],
                                     "cg_same_iters": f[1], "cg_hist": f[2], "cg_x": f[3], "hist_dev": hist_dev,
                                     "iters": int(res.iteration_count)})
                ok = ok and all(f)
    # production-shaped case (--production / DIST_CHECK_PRODUCTION=1): 32 planes of 256^2 per rank, the shape of the
    # benchmarked jobs (bench.py: 256 planes per rank for the product, 64 for CG at 8 GPUs) rather than 2-8 planes.
    # Product: every bit against the closed form (index arithmetic, gallery.poisson7pt_benchmark_product), through the
    # device entry point and through host buffers (b200sp_spmv_dist_host).  CG: 40 iterations against the oracle CG of
    # the whole operator, history within 1e-10.
    if world > 1 and ("--production" in sys.argv or os.environ.get("DIST_CHECK_PRODUCTION") == "1"):
        grid = (256, 256, 32 * world)
        blk = plane_partition(grid, world, rank)
        prod = {"grid": list(grid), "planes_per_rank": 32}
        for fmt, tdt in (("dia", torch.float64), ("dia", torch.float32), ("ell", torch.float64), ("csr", torch.float64)):
            A = gallery.poisson(fmt, 7, grid, dtype=tdt, row_begin=blk.row_begin, num_rows=blk.num_rows,
                                halo_lo=blk.halo_lo, halo_hi=blk.halo_hi)
            halo = capi.Halo(blk.halo_lo, blk.halo_hi)
            xw = ((torch.arange(blk.window, device=dev) + blk.col_shift) % 21 - 10).to(tdt)
            xw[:blk.halo_lo] = float("nan")
            xw[blk.halo_lo + blk.num_rows:] = float("nan")
            y = torch.empty(blk.num_rows, dtype=tdt, device=dev)
            want = gallery.poisson7pt_benchmark_product(grid, blk.row_begin, blk.num_rows, tdt, dev)
            good = 1
            for rep in range(3):
                if rep % world == rank:
                    torch.cuda._sleep(2_000_000)
                y.fill_(float("nan"))
                h.spmv_dist(A.descriptor(), halo, xw, y)
                good &= int(torch.equal(y, want))
            xh = xw[blk.halo_lo:blk.halo_lo + blk.num_rows].cpu().pin_memory()
            yh = torch.empty(blk.num_rows, dtype=tdt).pin_memory()
            for rep in range(2):
                yh.fill_(float("nan"))
                h.spmv_dist_host(A.descriptor(), halo, xh, yh)
                good &= int(torch.equal(yh, want.cpu()))
            flags = torch.tensor([good], device=dev)
            td.all_reduce(flags, op=td.ReduceOp.MIN)
            prod[f"spmv_{fmt}_{'f64' if tdt == torch.float64 else 'f32'}_exact"] = int(flags.item())
            ok = ok and bool(flags.item())
            if fmt == "dia" and tdt == torch.float64:
                ref = O.poisson(7, grid, np.float64, "csr")
                nall = ref["num_rows"]
                # the yardstick for a long history is the oracle iteration with its dot products accumulated in long
                # double: on 10^7 unknowns the sequential fp64 sums of the plain oracle drift by ~1e-10 themselves
                xo, it_o, conv_o, hist_o = O.cg(ref, np.zeros(nall), np.ones(nall), 40, 0.0, compensated=True)
                del ref
                xl = torch.zeros(blk.num_rows, dtype=tdt, device=dev)
                bl = torch.ones(blk.num_rows, dtype=tdt, device=dev)
                res, hist = h.cg(A.descriptor(), xl, bl, iteration_limit=40, relative_tolerance=0.0, check_interval=8,
                                 halo=halo)
                m = min(len(hist), len(hist_o))
                devmax = float(np.max(np.abs(np.asarray(hist[:m]) - hist_o[:m]) / hist_o[:m]))
                x_ok = bool(np.abs(xl.cpu().numpy() - xo[blk.row_begin: blk.row_begin + blk.num_rows]).max() <= 1e-8 * np.abs(xo).max())
                flags = torch.tensor([int(devmax <= 1e-10 and len(hist) == len(hist_o) and int(res.iteration_count) == it_o), int(x_ok)],
                                     device=dev)
                td.all_reduce(flags, op=td.ReduceOp.MIN)
                prod["cg_hist_max_rel_dev"] = devmax
                prod["cg_hist_ok"], prod["cg_x_ok"] = int(flags[0].item()), int(flags[1].item())
                ok = ok and bool(flags[0].item()) and bool(flags[1].item())
                # the opt-in two-kernel form (r planes travel, p rebuilt inside the product) gives the same bits as
                # the default three-kernel form (direction update as its own kernel, p planes travel)
                os.environ["B200SP_CG_FUSE"] = "1"
                xl3 = torch.zeros(blk.num_rows, dtype=tdt, device=dev)
                res3, hist3 = h.cg(A.descriptor(), xl3, bl, iteration_limit=40, relative_tolerance=0.0, check_interval=8,
                                   halo=halo)
                del os.environ["B200SP_CG_FUSE"]
                same = int(list(hist3) == list(hist) and torch.equal(xl3, xl) and int(res3.iteration_count) == int(res.iteration_count))
                flags = torch.tensor([same], device=dev)
                td.all_reduce(flags, op=td.ReduceOp.MIN)
                prod["cg_two_vs_three_kernel_bits"] = int(flags.item())
                ok = ok and bool(flags.item())
            del A, xw, y, want
            torch.cuda.empty_cache()
        out["production"] = prod
    # skew stress: ranks take turns being late by ~1 ms before an exchange, so neighbours run one
    # exchange ahead of each other (epoch waits must be monotonic, staging double-buffered)
    if world > 1:
        grid = (64, 32, 4 * world)
        blk = plane_partition(grid, world, rank)
        A = gallery.poisson("dia", 7, grid, dtype=torch.float64, row_begin=blk.row_begin, num_rows=blk.num_rows,
                            halo_lo=blk.halo_lo, halo_hi=blk.halo_hi)
        halo = capi.Halo(blk.halo_lo, blk.halo_hi)
        ref = O.poisson(7, grid, np.float64, "dia")
        stress_ok = 1
        for it in range(48):
            xg = ((np.arange(ref["num_cols"]) * (it + 1)) % 23 - 11).astype(np.float64)
            xw = torch.full((blk.window,), float("nan"), dtype=torch.float64, device=dev)
            xw[blk.halo_lo: blk.halo_lo + blk.num_rows] = torch.from_numpy(
                xg[blk.row_begin: blk.row_begin + blk.num_rows]).to(dev)
            y = torch.empty(blk.num_rows, dtype=torch.float64, device=dev)
            if it % world == rank:
                torch.cuda._sleep(2_000_000)
            h.spmv_dist(A.descriptor(), halo, xw, y)
            want = O.spmv(ref, xg)[blk.row_begin: blk.row_begin + blk.num_rows]
            if not np.array_equal(y.cpu().numpy(), want):
                stress_ok = 0
        flags = torch.tensor([stress_ok], device=dev)
        td.all_reduce(flags, op=td.ReduceOp.MIN)
        out["skew_stress_ok"] = int(flags.item())
        ok = ok and bool(out["skew_stress_ok"])
    # graph operator (R-MAT): rows partitioned in contiguous blocks, global column indices, x gathered
    # from all ranks (b200sp_spmv_dist_gather).  Integer-valued data: y must equal the single-GPU
    # product of the whole matrix bit for bit; ranks take turns being late (staging parity).
    if world > 1:
        from cusp_autotuned_b200 import convert
        from cusp_autotuned_b200.matrix import coo_matrix
        from cusp_autotuned_b200.partition import row_block_offsets
        scale = 16
        C = convert.rmat(scale, 16, seed=7, dtype=torch.float32, values="ones")  # same stream on every GPU
        n = C.num_rows
        from cusp_autotuned_b200.partition import nnz_balanced_offsets
        graph_ok = 1
        # equal row counts, and nnz-balanced blocks cut at arbitrary rows (slices not 16-byte aligned)
        # (row blocks, x slices): equal row counts with x partitioned alike; nnz-balanced blocks cut at arbitrary
        # rows (slices not 16-byte aligned) with x alike; nnz-balanced rows with x in equal slices (decoupled)
        bal = nnz_balanced_offsets(C.row_indices, n, world, align=1)
        eq = row_block_offsets(n, world)
        for offs, xoffs, tdt in [(o, xo, t) for o, xo in ((eq, eq), (bal, bal), (bal, eq))
                                 for t in (torch.float32, torch.float64)]:
            bounds = torch.searchsorted(C.row_indices, torch.tensor(offs, dtype=torch.int32, device=dev))
            e0, e1 = int(bounds[rank]), int(bounds[rank + 1])
            nloc = offs[rank + 1] - offs[rank]
            vals = C.values.to(tdt)
            Cg = coo_matrix(n, n, C.row_indices, C.column_indices, vals)
            loc = coo_matrix(nloc, n, (C.row_indices[e0:e1] - offs[rank]).contiguous(),
                             C.column_indices[e0:e1].clone(), vals[e0:e1].clone())
            mats = {"coo": loc, "csr": convert.coo_to_csr(loc)}
            mats["hyb"] = convert.csr_to_hyb(mats["csr"])
            for it in range(6):
                xg = ((torch.arange(n, device=dev) * (it + 3)) % 17 - 8).to(tdt)
                yref = torch.empty(n, dtype=tdt, device=dev)
                h.spmv(Cg.descriptor(), xg, yref)
                for fmt, M in mats.items():
                    xf = torch.full((n,), float("nan"), dtype=tdt, device=dev)
                    xf[xoffs[rank]:xoffs[rank + 1]] = xg[xoffs[rank]:xoffs[rank + 1]]
                    y = torch.empty(nloc, dtype=tdt, device=dev)
                    if (it + len(fmt)) % world == rank:
                        torch.cuda._sleep(2_000_000)
                    h.spmv_dist_gather(M.descriptor(), xoffs, xf, y)
                    if not (torch.equal(xf, xg) and torch.equal(y, yref[offs[rank]:offs[rank + 1]])):
                        graph_ok = 0
        flags = torch.tensor([graph_ok], device=dev)
        td.all_reduce(flags, op=td.ReduceOp.MIN)
        out["graph_gather_ok"] = int(flags.item())
        out["graph"] = {"matrix": f"R-MAT scale {scale} ef 16 (ones)", "nnz": int(C.num_entries), "offsets": offs}
        ok = ok and bool(out["graph_gather_ok"])
    if world > 1:
        bad = torch.tensor([h.comm_timeouts()], dtype=torch.int64, device=dev)
        td.all_reduce(bad, op=td.ReduceOp.MAX)
        out["comm_timeouts"] = int(bad.item())
        ok = ok and out["comm_timeouts"] == 0
    out["ok"] = ok
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        h.comm_destroy()
        td.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
