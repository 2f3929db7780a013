"""CPU suite: host-side logic of the Python mirror (argument checking, partition,
monitor bookkeeping) — no device needed."""
import numpy as np
import pytest
import torch

import cusp_autotuned_b200 as cusp
from cusp_autotuned_b200 import capi
from cusp_autotuned_b200.partition import halo_plan, plane_partition


def test_containers_reject_host_tensors():
    i = torch.zeros(5, dtype=torch.int32)
    with pytest.raises(cusp.InvalidInput):
        cusp.csr_matrix(4, 4, i, i, torch.zeros(5))


def test_blas_size_checks_happen_before_any_device_work():
    x, w = torch.zeros(4), torch.zeros(3)
    for call in (lambda: cusp.blas.axpy(x, w, 1.0), lambda: cusp.blas.dot(x, w),
                 lambda: cusp.blas.axpby(x, x, w, 1.0, 1.0), lambda: cusp.blas.copy(w, x)):
        with pytest.raises(cusp.InvalidInput):
            call()


def test_plane_partition_covers_the_grid():
    for dims in ((4, 4, 8), (5, 3, 7), (512, 512, 512), (10, 9)):
        for world in (1, 2, 3, 4, 7):
            if dims[-1] < world:
                with pytest.raises(ValueError):
                    plane_partition(dims, world, 0)
                continue
            blocks = [plane_partition(dims, world, r) for r in range(world)]
            total = int(np.prod(dims))
            plane = total // dims[-1]
            assert blocks[0].row_begin == 0 and blocks[-1].row_begin + blocks[-1].num_rows == total
            for a, b in zip(blocks, blocks[1:]):
                assert a.row_begin + a.num_rows == b.row_begin
                assert a.halo_hi == b.halo_lo == plane  # symmetric halos
            assert blocks[0].halo_lo == 0 and blocks[-1].halo_hi == 0
            assert all(b.num_rows % plane == 0 and b.num_rows > 0 for b in blocks)
            assert max(b.num_rows for b in blocks) - min(b.num_rows for b in blocks) <= plane


def test_halo_plan_is_consistent_between_neighbours():
    world = 4
    blocks = [plane_partition((6, 5, 9), world, r) for r in range(world)]
    plans = [halo_plan(b) for b in blocks]
    for r, plan in enumerate(plans):
        for peer, send, recv in plan:
            back = [p for p in plans[peer] if p[0] == r]
            assert len(back) == 1
            assert (send.stop - send.start) == (back[0][2].stop - back[0][2].start)
            # what I send is the global range my neighbour expects in its halo
            mine = blocks[r].col_shift + np.arange(send.start, send.stop)
            theirs = blocks[peer].col_shift + np.arange(back[0][2].start, back[0][2].stop)
            assert np.array_equal(mine, theirs)


def test_monitor_bookkeeping():
    m = cusp.monitor(None, iteration_limit=7, relative_tolerance=1e-3, absolute_tolerance=0.5)
    assert m.iteration_limit() == 7 and m.iteration_count() == 0
    assert m.relative_tolerance() == 1e-3 and m.absolute_tolerance() == 0.5
    res = capi.CgResult(iteration_count=3, converged=1, residual_norm=0.25, b_norm=2.0, num_residuals=4)
    m._absorb(res, np.array([4.0, 2.0, 1.0, 0.25]))
    assert m.iteration_count() == 3 and m.residual_norm() == 0.25 and len(m.residuals) == 4
    assert m.tolerance() == 0.5 + 1e-3 * 2.0 and m.converged()


def test_ktt_switch():
    from cusp_autotuned_b200 import ktt
    assert ktt.is_enabled()  # cusp/ktt/detail/ktt.inl:21
    ktt.disable()
    assert not ktt.is_enabled()
    ktt.enable()
    assert ktt.is_enabled()
