"""Row-block partition of a stencil grid for the multi-GPU operator.

Pure host logic (no torch, no CUDA): the slowest grid axis is cut into
contiguous slabs of planes, one per rank; a rank's x window is
[halo_lo | local | halo_hi] with one plane of halo towards each neighbour."""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class RowBlock:
    rank: int
    world: int
    row_begin: int
    num_rows: int
    halo_lo: int
    halo_hi: int

    @property
    def col_shift(self) -> int:  # global column j lives at window index j - col_shift
        return self.row_begin - self.halo_lo

    @property
    def window(self) -> int:
        return self.halo_lo + self.num_rows + self.halo_hi


def plane_partition(dims, world: int, rank: int) -> RowBlock:
    """dims = (nx, ny) for the 5-point or (nx, ny, nz) for the 7-point stencil."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    dims = tuple(int(d) for d in dims)
    slow = dims[-1]
    plane = 1
    for d in dims[:-1]:
        plane *= d
    if slow < world:
        raise ValueError(f"cannot cut {slow} planes over {world} ranks")
    base, extra = divmod(slow, world)
    first = rank * base + min(rank, extra)
    count = base + (1 if rank < extra else 0)
    return RowBlock(rank, world, first * plane, count * plane,
                    plane if rank > 0 else 0, plane if rank < world - 1 else 0)


def halo_plan(blk: RowBlock):
    """the exchanges b200sp's comm_halo_exchange performs for one rank, as
    (peer, send_slice, recv_slice) over the window [halo_lo | local | halo_hi];
    halos are symmetric: a rank sends to a neighbour as much as it receives from it."""
    plan = []
    lo, n = blk.halo_lo, blk.num_rows
    if blk.rank > 0 and lo > 0:
        plan.append((blk.rank - 1, slice(lo, lo + lo), slice(0, lo)))
    if blk.rank < blk.world - 1 and blk.halo_hi > 0:
        hi = blk.halo_hi
        plan.append((blk.rank + 1, slice(lo + n - hi, lo + n), slice(lo + n, lo + n + hi)))
    return plan
