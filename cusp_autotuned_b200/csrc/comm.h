// comm.h — NCCL plumbing for the row-block partitioned operator.
// The reference has no communication layer at all; this is new functionality.
#pragma once
#include "common.cuh"

namespace b200sp {
// in-place sum all-reduce of `count` scalars living in device memory, on `st`
b200sp_status comm_allreduce_sum(b200sp_handle h, cudaStream_t st, void *dev, int count, bool is_double);
// dev[0] = sqrt(dev[0])
void comm_sqrt_inplace(b200sp_handle h, cudaStream_t st, void *dev, bool is_double);
// window = [halo_lo | n local | halo_hi]; fills both halos from the neighbours
// (rank-1 / rank+1) and sends them the matching edge slices of the local part.
// Halos are symmetric: what a rank receives from a neighbour has the size of
// what it sends to it (true for the stencil operators partitioned by planes).
b200sp_status comm_halo_exchange(b200sp_handle h, cudaStream_t st, void *window, i64 n, i64 halo_lo,
                                 i64 halo_hi, size_t elem);
}  // namespace b200sp
