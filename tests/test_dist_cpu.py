"""CPU suite: the N>1 path on world_size-2/3 `gloo` process groups.  Each rank
owns a row block of poisson7pt, exchanges halos following
cusp_autotuned_b200.partition.halo_plan (the plan libb200sp's comm_halo_exchange
executes with NCCL), runs the LOCAL product with the oracle and all-reduces the
dot products of a CG solve.  Results must equal the single-process operator."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

from cusp_autotuned_b200.partition import halo_plan, plane_partition
from oracle import oracle as O

DIMS = (6, 5, 9)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _local_block(blk):
    """rows [row_begin, +num_rows) of the global CSR with window-relative columns"""
    full = O.poisson(7, DIMS, np.float64, "csr")
    r0, r1 = blk.row_begin, blk.row_begin + blk.num_rows
    a, b = full["row_offsets"][r0], full["row_offsets"][r1]
    return dict(format="csr", num_rows=blk.num_rows, num_cols=blk.window, num_entries=int(b - a),
                row_offsets=(full["row_offsets"][r0:r1 + 1] - a).astype(np.int32),
                column_indices=(full["column_indices"][a:b] - blk.col_shift).astype(np.int32),
                values=full["values"][a:b].copy())


def _exchange(window, blk):
    reqs = []
    t = torch.from_numpy(window)
    recvs = []
    for peer, send, recv in halo_plan(blk):
        reqs.append(td.isend(t[send].clone(), dst=peer))
        buf = torch.empty(recv.stop - recv.start, dtype=t.dtype)
        reqs.append(td.irecv(buf, src=peer))
        recvs.append((recv, buf))
    for r in reqs:
        r.wait()
    for recv, buf in recvs:
        t[recv] = buf


def _allsum(v):
    t = torch.tensor([v], dtype=torch.float64)
    td.all_reduce(t)
    return float(t.item())


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cusp_autotuned_b200.dist import broadcast_bytes
        uid = broadcast_bytes(bytes(range(128)) if rank == 0 else None, 128)
        assert uid == bytes(range(128))  # unique-id hand-over works on gloo too

        blk = plane_partition(DIMS, world, rank)
        A = _local_block(blk)
        n = blk.num_rows
        rng = np.random.default_rng(0)
        xg = rng.uniform(-1, 1, int(np.prod(DIMS)))
        # partitioned SpMV
        win = np.zeros(blk.window)
        win[blk.halo_lo: blk.halo_lo + n] = xg[blk.row_begin: blk.row_begin + n]
        _exchange(win, blk)
        y = O.spmv(A, win)
        # partitioned CG, reference operation order (cg.inl:63-105), 12 iterations
        b = np.ones(n)
        x = np.zeros(n)
        pw = np.zeros(blk.window)
        r = b.copy()  # x0 = 0 -> r = b
        p = pw[blk.halo_lo: blk.halo_lo + n]
        p[:] = r
        rz = _allsum(float(np.dot(r, r)))
        hist = [np.sqrt(rz)]
        for _ in range(12):
            _exchange(pw, blk)
            yv = O.spmv(A, pw)
            alpha = rz / _allsum(float(np.dot(yv, p)))
            x += alpha * p
            r -= alpha * yv
            rz_old, rz = rz, _allsum(float(np.dot(r, r)))
            p[:] = r + (rz / rz_old) * p
            hist.append(np.sqrt(rz))
        out[rank] = (blk.row_begin, y, x, hist)
    finally:
        td.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_spmv_and_cg_match_single_process(world):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    full = O.poisson(7, DIMS, np.float64, "csr")
    xg = np.random.default_rng(0).uniform(-1, 1, full["num_rows"])
    parts = [out[r] for r in range(world)]
    y = np.concatenate([p[1] for p in parts])
    assert np.array_equal(y, O.spmv(full, xg))  # every y entry is computed by exactly one rank
    xs = np.concatenate([p[2] for p in parts])
    xo, it, conv, hist = O.cg(full, np.zeros(full["num_rows"]), np.ones(full["num_rows"]), 12, 0.0)
    assert np.allclose(parts[0][3], hist, rtol=1e-10)  # only the dot-product grouping differs
    assert np.allclose(xs, xo, rtol=1e-9, atol=1e-13)


# ---------------------------------------------------------------------------
# graph operators: contiguous row blocks, global column indices, x all-gathered
# (the plan b200sp_spmv_dist_gather executes over NVLink peer memory / NCCL)
# ---------------------------------------------------------------------------
def test_row_block_offsets_properties():
    from cusp_autotuned_b200.partition import row_block_offsets
    for n in (0, 1, 31, 32, 33, 1000, 65531, 1 << 24):
        for world in (1, 2, 3, 8):
            offs = row_block_offsets(n, world)
            assert len(offs) == world + 1 and offs[0] == 0 and offs[-1] == n
            sizes = np.diff(offs)
            assert (sizes >= 0).all()
            assert all(o % 32 == 0 for o in offs[:-1])  # 16-byte aligned fp32 slices
            if n >= 32 * world:
                assert sizes.max() - sizes.min() <= 32 + 31


def test_nnz_balanced_offsets_split_the_entries_evenly():
    from cusp_autotuned_b200.partition import nnz_balanced_offsets
    rng = np.random.default_rng(5)
    n = 5000
    deg = (rng.pareto(1.2, n) * 3).astype(np.int64) + (np.arange(n) % 400 == 0) * 3000 * (np.arange(n) < 2000)  # hubs early
    rows = np.repeat(np.arange(n), deg)
    for world in (1, 2, 4, 8):
        offs = nnz_balanced_offsets(rows, n, world)
        assert offs[0] == 0 and offs[-1] == n and all(a <= b for a, b in zip(offs, offs[1:]))
        assert all(o % 32 == 0 for o in offs[:-1])
        per = np.diff(np.searchsorted(rows, offs))
        assert per.sum() == len(rows)
        window = np.convolve(deg, np.ones(32, np.int64), mode="full").max()  # entries of any 32 consecutive rows
        assert per.max() <= len(rows) / world + 2 * window
    assert nnz_balanced_offsets([], 7, 3) == [0, 0, 0, 7]


def _graph_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cusp_autotuned_b200.partition import coo_row_block, row_block_offsets
        A = O.gallery_random(700, 700, 9000, np.float64, "coo")  # same matrix on every rank
        offs = row_block_offsets(A["num_rows"], world)
        e0, e1 = coo_row_block(A["row_indices"].tolist(), offs, rank)
        loc = dict(format="coo", num_rows=offs[rank + 1] - offs[rank], num_cols=A["num_cols"], num_entries=e1 - e0,
                   row_indices=(A["row_indices"][e0:e1] - offs[rank]).astype(np.int32),
                   column_indices=A["column_indices"][e0:e1].copy(), values=A["values"][e0:e1].copy())
        xg = np.random.default_rng(3).uniform(-1, 1, A["num_cols"])
        # all-gather of the x slices (slices are ragged: all_gather on padded tensors)
        m = max(np.diff(offs))
        mine = torch.zeros(m, dtype=torch.float64)
        mine[: offs[rank + 1] - offs[rank]] = torch.from_numpy(xg[offs[rank]:offs[rank + 1]])
        parts = [torch.zeros(m, dtype=torch.float64) for _ in range(world)]
        td.all_gather(parts, mine)
        xf = np.concatenate([parts[r][: offs[r + 1] - offs[r]].numpy() for r in range(world)])
        out[rank] = (xf, O.spmv(loc, xf))
    finally:
        td.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_graph_row_blocks_with_gathered_x_match_single_process(world):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_graph_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    A = O.gallery_random(700, 700, 9000, np.float64, "coo")
    xg = np.random.default_rng(3).uniform(-1, 1, A["num_cols"])
    for r in range(world):
        assert np.array_equal(out[r][0], xg)
    y = np.concatenate([out[r][1] for r in range(world)])
    assert np.array_equal(y, O.spmv(A, xg))  # rows are never split across ranks: same sums, same order
