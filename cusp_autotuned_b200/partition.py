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


def row_block_offsets(num_rows: int, world: int):
    """contiguous row blocks for operators without grid structure (graphs): world+1 offsets,
    block r = [offsets[r], offsets[r+1]); sizes differ by at most one row and are multiples
    of 32 rows where possible (16-byte aligned fp32 slices for the peer-memory all-gather)"""
    if world < 1:
        raise ValueError("world must be >= 1")
    units, rem = divmod(int(num_rows), 32)
    base, extra = divmod(units, world)
    offs = [0]
    for r in range(world):
        offs.append(offs[-1] + 32 * (base + (1 if r < extra else 0)))
    offs[-1] += rem
    return offs


def coo_row_block(row_indices, offsets, rank: int):
    """entry range [e0, e1) of a row-sorted COO matrix whose rows fall into block `rank`
    (row_indices: any sorted sequence supporting searchsorted-style bisect)"""
    import bisect
    return bisect.bisect_left(row_indices, offsets[rank]), bisect.bisect_left(row_indices, offsets[rank + 1])


def nnz_balanced_offsets(row_indices, num_rows: int, world: int, align: int = 32):
    """row blocks with about nnz/world stored entries each (power-law graphs: equal row counts
    would give the block holding the hub rows most of the matrix).  `row_indices` is the
    row-sorted COO row array (anything indexable: list, numpy, torch); block boundaries are
    rounded to a multiple of `align` rows and kept ascending."""
    nnz = len(row_indices)
    offs = [0]
    for r in range(1, world):
        cut = int(row_indices[(nnz * r) // world]) if nnz else 0
        cut = ((cut + align // 2) // align) * align  # nearest multiple of `align`
        offs.append(min(max(cut, offs[-1]), int(num_rows)))
    offs.append(int(num_rows))
    return offs
