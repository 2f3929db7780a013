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

// ---- NVLink peer-memory path -------------------------------------------------
// Mailbox layout (device memory, one per rank, mapped into every peer through CUDA IPC):
//   slot[ch][r]  ch in {0,1}: value + epoch written by rank r (scalar all-reduce by
//                "everyone writes to everyone, everyone sums in rank order")
//   halo_flag[2] epoch of the last halo push from rank-1 / rank+1
constexpr int P2P_MAX_WORLD = 16;
// A scalar travels as two 8-byte words, each = (32 data bits | 32-bit epoch tag).  An
// aligned 8-byte store is single-copy atomic, so data and tag arrive together and the
// reader needs no fence: it spins until both tags carry the expected epoch (the
// "LL" idea of NCCL's low-latency protocol).
struct MailSlot {
  unsigned long long w[2];
};
struct Mailbox {
  MailSlot slot[2][P2P_MAX_WORLD];
  unsigned long long halo_flag[2];  // CG: epoch of the last p-halo push from rank-1 / rank+1
  unsigned long long xchg_flag[2];  // spmv_dist: epoch of the last staged plane from rank-1 / rank+1
  unsigned long long xchg_go;       // spmv_dist: local "both planes have landed" signal
  unsigned long long timeouts;      // spin loops that gave up (a peer never arrived): results are invalid
  unsigned long long gather_flag[P2P_MAX_WORLD];  // spmv_dist_gather: epoch of rank r's last staged x slice
};
#ifdef __CUDACC__
// every cross-GPU spin is bounded (~4 s of SM clocks): a peer that died must not hang this GPU
struct SpinGuard {
  long long t0;
  __device__ SpinGuard() {
#ifdef __CUDA_ARCH__
    t0 = clock64();
#else
    t0 = 0;
#endif
  }
  __device__ bool expired(Mailbox *mine) {
#ifdef __CUDA_ARCH__
    // ~4 s of SM clocks (ranks of a real job arrive seconds apart after host-side work) for the first failure; once one wait has failed the job is broken
    // anyway (b200sp_comm_timeouts() != 0), so later waits give up at once instead of
    // stretching a dead run by half a second per kernel
    if (clock64() - t0 < (1ll << 33) && *reinterpret_cast<volatile unsigned long long *>(&mine->timeouts) == 0)
      return false;
    atomicAdd(&mine->timeouts, 1ull);
#endif
    return true;
  }
};
#endif
// staging buffer layout (bytes): [parity 0 | parity 1] x [from rank-1 | from rank+1] x P2P_STAGE_SIDE
constexpr size_t P2P_STAGE_SIDE = (size_t)16 << 20;
struct P2PView {  // passed by value to the CG kernels
  Mailbox *peer[P2P_MAX_WORLD];
  Mailbox *mine;
  int world, rank;
};
// all-gather `bytes` (<= 256) of host data from every rank into host_out[world*bytes]
b200sp_status comm_allgather_host(b200sp_handle h, cudaStream_t st, const void *host_in, size_t bytes,
                                  void *host_out);
// map the CG workspaces of rank-1 / rank+1 (collective; every rank passes its own cg_ws
// base, the byte offset of its p window inside it and its local sizes).  On return
// dst_lo / dst_hi are the addresses, in the neighbours' windows, where this rank's
// first halo_lo / last halo_hi elements of p belong (nullptr at the ends of the chain).
b200sp_status comm_p2p_map_windows(b200sp_handle h, cudaStream_t st, void *ws_base, size_t pwin_offset,
                                   i64 n, i64 halo_lo, i64 halo_hi, size_t elem, void **dst_lo,
                                   void **dst_hi);
// Halo exchange fused into the bulk kernel (b200sp_spmv_dist, NVLink peer memory).
// The 31 idle lanes of every CTA's producer warp push this rank's two edge planes into
// the neighbours' staging buffers, publish, wait for the neighbours' planes and copy
// them into the x window, all while the consumer warps stream interior tiles.  Tiles
// are visited in rotated order so that the tiles which read halo columns come last;
// their consumers wait on a local flag that is raised when the copy-out is complete.
struct FusedXchg {
  int enabled;  // 0: off   1: exchange inside the kernel   2: only wait for flags raised by someone else
  // consumers of tiles that read the lower / upper halo wait until *wait_lo / *wait_hi >= wait_epoch
  const unsigned long long *wait_lo, *wait_hi;
  unsigned long long wait_epoch;
  char *window;        // x window base [halo_lo | local | halo_hi]
  size_t local_off, n_bytes, lo_bytes, hi_bytes;
  char *stage_mine, *stage_lo_nbr, *stage_hi_nbr;  // nullptr where there is no neighbour
  Mailbox *mine, *mail_lo_nbr, *mail_hi_nbr;
  unsigned int *tickets;  // [2], zero on entry, reset by the kernel
  unsigned long long epoch;
  long long rot;            // tile rotation: sequence index s -> tile (s + rot) mod num_tiles
  long long lo_tiles;       // tiles [0, lo_tiles) read the lower halo
  long long hi_tile_begin;  // tiles [hi_tile_begin, num_tiles) read the upper halo
};

// fill `xc` for one exchange of the window [halo_lo | n | halo_hi] (advances the exchange
// epoch); returns false when the peer path cannot carry it (then use
// comm_halo_exchange_auto before the product instead)
bool comm_fused_xchg_prepare(b200sp_handle h, void *window, i64 n, i64 halo_lo, i64 halo_hi, size_t elem,
                             FusedXchg *xc);
P2PView comm_p2p_view(b200sp_handle h);
// halo exchange through NVLink peer memory (one kernel: store my edge planes into the
// neighbours' staging buffers, publish, wait for theirs, copy them into my window);
// falls back to NCCL send/recv when the peer path is not available or a plane exceeds
// P2P_STAGE_SIDE.  Same contract as comm_halo_exchange.
b200sp_status comm_halo_exchange_auto(b200sp_handle h, cudaStream_t st, void *window, i64 n, i64 halo_lo,
                                      i64 halo_hi, size_t elem);
// All-gather of a row-block partitioned vector (graph operators: every rank's rows read the
// whole of x).  x_full holds `slice_offsets[world]` elements; this rank's slice
// [slice_offsets[rank], slice_offsets[rank+1]) is valid on entry, all slices on return.
// Peer-memory path: one kernel copies the own slice into an IPC-shared staging buffer,
// publishes an epoch flag in every peer's mailbox, and pulls the other ranks' slices straight
// from their staging buffers over NVLink (rotated peer order, staging double-buffered by
// epoch parity).  Without CUDA IPC: grouped ncclSend / ncclRecv.  Collective.
b200sp_status comm_allgather_slices(b200sp_handle h, cudaStream_t st, void *x_full, const int64_t *slice_offsets,
                                    size_t elem);
// agree on a fresh solve id (max over ranks + 1): epochs never repeat across solves
b200sp_status comm_next_solve_id(b200sp_handle h, cudaStream_t st, unsigned long long *id);
}  // namespace b200sp
