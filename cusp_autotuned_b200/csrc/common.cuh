// common.cuh — shared device helpers and the handle of libb200sp.
//
// Everything in csrc/ is written for sm_100a only (B200): 148 SMs, 227 KB smem
// per CTA, cp.async.bulk + mbarrier staging, ld.global.cs / ld.global.nc hints.
// Compiled with -fmad=false so that kernels which keep the reference's
// per-row summation order (thread-per-row CSR/ELL/DIA, BLAS-1 updates) are
// bit-identical to the reference's host loops
// (cusp/system/detail/sequential/multiply/*.h), which x86-64 builds without FMA.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "b200sp.h"

#ifndef B200SP_NUM_SMS_FALLBACK
#define B200SP_NUM_SMS_FALLBACK 148
#endif

typedef long long i64;

// ---------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------
struct b200sp_tune_key {
  int format, dtype, rows_log2, nnz_per_row_log2;
  // what the defaults already look at when they pick a kernel family: CSR 1 banded / 2 scattered / 3 skewed row
  // lengths, COO and the COO part of HYB 1 banded / 2 one entry per row / 3 scattered columns, 0 where structure
  // does not enter (ELL, DIA).  A winner tuned on a banded CSR is not replayed on a random CSR of the same size.
  int structure;
  bool operator<(const b200sp_tune_key &o) const {
    if (format != o.format) return format < o.format;
    if (dtype != o.dtype) return dtype < o.dtype;
    if (rows_log2 != o.rows_log2) return rows_log2 < o.rows_log2;
    if (nnz_per_row_log2 != o.nnz_per_row_log2) return nnz_per_row_log2 < o.nnz_per_row_log2;
    return structure < o.structure;
  }
};

struct b200sp_tune_entry {
  b200sp_cfg best;
  bool has_best = false;
  // dynamic tuning progress (b200sp_tune_step)
  int64_t next_index = 0;
  float best_ms = 1e30f;
};

struct b200sp_context {
  int device = 0;
  int num_sms = B200SP_NUM_SMS_FALLBACK;
  int max_smem_optin = 0;
  size_t l2_bytes = 0;
  int max_persist_l2 = 0;
  char err[1024];
  uint64_t launches = 0;

  // device scratch, grown on demand, never shrunk (owned by the handle)
  void *scratch = nullptr;
  size_t scratch_bytes = 0;
  // reduction workspace: partial sums + ticket counters + result scalars
  double *red_partials = nullptr;  // RED_MAX_PARTIALS * 4 doubles
  unsigned int *red_counters = nullptr;  // 16 tickets
  double *dev_scalars = nullptr;   // 64 doubles (CG state lives here)
  double *pinned_scalars = nullptr;  // 64 doubles, pinned host
  // host-buffer entry points
  void *stage_x = nullptr, *stage_y = nullptr;
  size_t stage_x_bytes = 0, stage_y_bytes = 0;
  // pipelined host path (b200sp_spmv_host on banded matrices): copy streams + events
  void *copy_in_stream = nullptr, *copy_out_stream = nullptr;
  std::vector<void *> pipe_events;
  bool pipe_events_timed = false;  // B200SP_HOST_TRACE=1: the pipeline's events carry timestamps
  // CG workspace
  void *cg_ws = nullptr;
  size_t cg_ws_bytes = 0;
  double *cg_residuals = nullptr;  // device residual history
  size_t cg_residuals_cap = 0;

  std::map<b200sp_tune_key, b200sp_tune_entry> tune_cache;
  // CSR structure analysis (longest row) per (row_offsets pointer, rows, nnz): decides between the
  // row-split kernels and the nnz-balanced one when the caller gives no configuration.  Only a
  // performance hint — every kernel is correct for every matrix — so a stale entry is harmless.
  struct CsrKey {
    const void *ap;
    int64_t rows, nnz;
    bool operator<(const CsrKey &o) const {
      if (ap != o.ap) return ap < o.ap;
      if (rows != o.rows) return rows < o.rows;
      return nnz < o.nnz;
    }
  };
  std::map<CsrKey, int> csr_max_row;
  // COO gather-order probe per (column_indices pointer, element size, nnz): 1 = ring kernel
  std::map<CsrKey, int> coo_gather_order;
  std::vector<void *> tune_events;  // cudaEvent_t pair
  // K_COO_WARP on >= 2^27 scattered entries: persistent grid (1) or one tile per warp (0), whichever was faster on
  // the first product with these arrays (spmv_coo.cu: coo_warp_grid_choice); same tiles, same bits either way
  std::map<CsrKey, int> coo_grid_choice;
  std::vector<void *> coo_choice_events;  // three cudaEvent_t
  std::vector<void *> coo_plans;    // attached b200sp_coo_plan (spmv_coo_plan.cu)
  // set by the CG driver around its iteration: the DIA bulk kernel is launched with programmatic stream
  // serialization and waits (griddepcontrol.wait) before its first read of x / y (spmv_dia.cu, cg.cu)
  bool pdl_spmv = false;
  // CUDA-graph paths (small, launch-bound systems): a capturable stream and an ordering event owned by the handle
  void *graph_stream = nullptr, *graph_event = nullptr;

  // multi-GPU
  void *nccl_comm = nullptr;
  int world = 1, rank = 0;
  // NVLink peer-memory path (comm.cu): a 4 KiB mailbox per rank, IPC-mapped into every
  // peer, carries the CG scalars and the halo-arrival flags; the CG workspace of the two
  // neighbouring ranks is mapped on demand so halo planes are stored straight into it.
  bool p2p_ok = false;
  void *mail = nullptr;             // this rank's mailbox (device)
  void *peer_mail[16] = {nullptr};  // every rank's mailbox as seen from this device
  void *nbr_ws[2] = {nullptr, nullptr};        // mapped cg_ws of rank-1 / rank+1
  unsigned char nbr_ws_handle[2][64] = {{0}};  // IPC handles currently mapped in nbr_ws
  unsigned long long solve_id = 0;
  // staging for the peer-memory halo exchange of b200sp_spmv_dist: this rank's buffer
  // (IPC-exported) and the two neighbours' buffers as mapped here
  void *halo_stage = nullptr;
  void *nbr_stage[2] = {nullptr, nullptr};
  unsigned long long xchg_epoch = 0;
  // staging for the peer-memory all-gather of b200sp_spmv_dist_gather: 2 parities x gather_slice_cap
  // bytes, IPC-exported; every peer's buffer mapped here
  void *gather_stage = nullptr;
  void *peer_gather[16] = {nullptr};
  size_t gather_slice_cap = 0;
  unsigned long long gather_epoch = 0;
};

enum { RED_MAX_PARTIALS = 1 << 16 };

namespace b200sp {

b200sp_status set_error(b200sp_handle h, b200sp_status s, const char *fmt, ...);
b200sp_status ensure_scratch(b200sp_handle h, size_t bytes);

#define B200SP_CHECK_HANDLE(h)                                       \
  do {                                                               \
    if ((h) == nullptr)                                              \
      return b200sp::set_error(nullptr, B200SP_INVALID_INPUT, "null handle"); \
  } while (0)

#define B200SP_CUDA(h, expr)                                                       \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess)                                                         \
      return b200sp::set_error((h), B200SP_CUDA_ERROR, "%s failed: %s (%s:%d)", #expr, \
                               cudaGetErrorString(_e), __FILE__, __LINE__);        \
  } while (0)

#define B200SP_LAUNCH_CHECK(h, name)                                               \
  do {                                                                             \
    cudaError_t _e = cudaGetLastError();                                           \
    if (_e != cudaSuccess)                                                         \
      return b200sp::set_error((h), B200SP_CUDA_ERROR, "launch of %s failed: %s", name, \
                               cudaGetErrorString(_e));                            \
    (h)->launches++;                                                               \
  } while (0)

#define B200SP_REQUIRE(h, cond, msg)                                               \
  do {                                                                             \
    if (!(cond)) return b200sp::set_error((h), B200SP_INVALID_INPUT, "%s (%s)", msg, #cond); \
  } while (0)

static inline i64 ceil_div(i64 a, i64 b) { return (a + b - 1) / b; }

#ifdef __CUDACC__
// launch with (pdl) or without programmatic stream serialization; same argument conversion rules as <<< >>>
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_kernel_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                            bool pdl, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
#endif
static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
#ifdef __CUDACC__

// matrix streams: read once -> ld.global.cs (evict-first in L1 and L2) keeps
// the 126 MB L2 for x.   x gathers: ld.global.nc (read-only path, L1 cached).
// Written as volatile asm so that the compiler keeps a batch of loads together
// in program order (NVVM otherwise sinks each load next to its use to save
// registers, which serialises the memory pipeline of these HBM-bound kernels).
__device__ __forceinline__ float ld_stream(const float *p) {
  float v;
  asm volatile("ld.global.cs.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double ld_stream(const double *p) {
  double v;
  asm volatile("ld.global.cs.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_stream(const int *p) {
  int v;
  asm volatile("ld.global.cs.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_ro(const float *p) {
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double ld_ro(const double *p) {
  double v;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_ro(const int *p) {
  int v;
  asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
// Scheduling fence for a loaded register: an empty volatile asm that "modifies"
// the value.  Volatile asms keep program order among themselves, so
//   [all ld_* of a batch]  [pin() of every loaded register]  [arithmetic]
// forces every load of the batch to be issued before the first dependent
// instruction -> RPT*DU independent requests in flight per thread.
__device__ __forceinline__ void pin(float &v) { asm volatile("" : "+f"(v)); }
__device__ __forceinline__ void pin(double &v) { asm volatile("" : "+d"(v)); }
__device__ __forceinline__ void pin(int &v) { asm volatile("" : "+r"(v)); }

template <typename T>
__device__ __forceinline__ void st_stream(T *p, T v) {
  __stcs(p, v);
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// Sum over `WIDTH` adjacent lanes (WIDTH power of two <= 32); result valid in
// the first lane of each group.  Row-split CSR reduction.
template <int WIDTH, typename T>
__device__ __forceinline__ T subwarp_sum(T v) {
#pragma unroll
  for (int o = WIDTH / 2; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o, WIDTH);
  return v;
}

// Block-wide sum in a fixed order: warp shuffles, then warp 0 adds the
// per-warp values in warp order.  Result valid in thread 0.
template <int BLOCK, typename T>
__device__ __forceinline__ T block_sum(T v, T *smem /* >= BLOCK/32 */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) smem[w] = v;
  __syncthreads();
  T r = 0;
  if (w == 0) {
    r = (lane < BLOCK / 32) ? smem[lane] : T(0);
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}

// Deterministic grid-wide reduction tail ("last block done"): every CTA stores
// its partial, the CTA that takes the last ticket adds all partials in index
// order (strided over threads, then block_sum) and hands the total to `fin`.
// `partials` needs gridDim.x entries, `*ticket` must be 0 on entry and is reset.
template <int BLOCK, typename T, typename Fin>
__device__ __forceinline__ void grid_reduce_finish(T block_value /* thread 0 */, T *partials,
                                                   unsigned int *ticket, T *smem, Fin fin) {
  __shared__ bool is_last;
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = block_value;
    __threadfence();
    unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    T acc = 0;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += BLOCK)
      acc += *((volatile T *)(partials + i));
    acc = block_sum<BLOCK>(acc, smem);
    if (threadIdx.x == 0) {
      fin(acc);
      *ticket = 0;
      __threadfence();
    }
  }
}

// ---- mbarrier / bulk-async (TMA engine, 1-D) -------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// Releasing a ring stage: the consumer's `mbarrier.arrive` on the stage's `empty` barrier
// must not be issued while shared-memory loads of that stage are still outstanding.  On
// sm_100a nothing orders it behind them: ptxas puts no scoreboard wait on SYNCS.ARRIVE, the
// arrive overtakes LDS instructions still queued in the MIO pipe, the producer's next bulk
// copy lands, and the late LDS reads the NEW tile (measured with the COO ring kernel:
// arrive right after the loads -> a few warps per launch read a half-overwritten stage;
// membar.cta in between does not help; consuming every loaded register first does).
// consume_before_release(acc) is that consumption for a value every staged load flows
// into: a real compare-and-branch on it (taken only for one NaN bit pattern, and then it
// only executes `nanosleep 0`) that the scheduler cannot move behind the arrive.
__device__ __forceinline__ int hi_bits(float v) { return __float_as_int(v); }
__device__ __forceinline__ int hi_bits(double v) { return __double2hiint(v); }
template <typename T>
__device__ __forceinline__ void consume_before_release(T acc) {
  if (hi_bits(acc) == (int)0x7ff4dead) asm volatile("nanosleep.u32 0;");
}

// ---- programmatic dependent launch (PDL) ------------------------------------------------
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while its
// predecessor in the stream is still draining; pdl_wait() blocks until the predecessor has completed
// and its writes are visible (a no-op for a normal launch).  pdl_trigger() lets the NEXT dependent
// kernel's CTAs be scheduled once every CTA of this grid has called it (or exited).  Rule used in
// this library: a kernel triggers only AFTER its own pdl_wait(), so whatever ran before its
// predecessor is complete before any successor's pre-wait code runs — a pre-wait prologue may read
// anything its predecessor does not write.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// L2 eviction policy for slabs that are read exactly once
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// global -> shared bulk copy (bytes % 16 == 0, both addresses 16 B aligned);
// completion is signalled on `bar` as transaction bytes.  SASS: UBLKCP.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                         uint64_t *bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

#endif  // __CUDACC__

}  // namespace b200sp
