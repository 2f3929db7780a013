// spmv_dia.cu — DIA SpMV for sm_100a.
//
// Replaces spmv_dia_kernel (cusp/system/cuda/detail/multiply/dia_spmv.h:66-126)
// and the KTT kernel ktt_dia_vector_kernel (cusp/system/cuda/ktt/kernels/
// dia_kernel.h:129-252).  Semantics follow the host loop
// cusp/system/detail/sequential/multiply/dia_spmv.h:36-82:
//     y[i] = init(y[i]);  for d ascending:  y[i] += values(i,d) * x[i+off[d]]
// for every (i,d) with 0 <= i+off[d] < num_cols.  One thread owns a row and adds
// the diagonals in ascending d, so with -fmad=false the result is bit-identical
// to that loop.
//
// Data layout (HBM): values is column-major, entry (row,d) at values[d*pitch+row]
// -> for a fixed d a tile of rows is one contiguous slab.
//
//  K_DIA_LDG : thread handles RPT rows strided by BLOCK (perfectly coalesced
//              32-lane loads), diagonals unrolled 8-deep so RPT*8 slab loads and
//              RPT*8 x loads are in flight per thread; slabs via ld.global.cs,
//              x via ld.global.nc.
//  K_DIA_BULK: persistent CTAs; a producer thread stages [KC diagonals x R rows]
//              slabs into a ring of shared-memory stages with cp.async.bulk
//              (TMA engine, SASS UBLKCP) completing on mbarriers with an L2
//              evict-first policy; consumer warps read the slabs from smem and x
//              through ld.global.nc.  Needs pitch*sizeof(T) % 16 == 0.
//
// Algorithmic bytes / row: K*sizeof(T) slab + sizeof(T) x + sizeof(T) y.
#include "common.cuh"
#include "comm.h"

namespace b200sp {

template <typename T, int MODE>
b200sp_status reduce(b200sp_handle, cudaStream_t, i64, const T *, const T *, T *, T *);  // blas1.cu

typedef FusedXchg DiaXchg;  // comm.h

template <typename T>
struct DiaArgs {
  i64 rows, cols, pitch;
  int ndiag;
  const int *offs;
  const T *vals;
  const T *x;
  T *y;
  int accumulate;
  // optional fused dot product  sum_r y[r]*dotv[r]  (CG: <Ap,p>)
  const T *dotv;
  T *dot_partials;
  unsigned int *dot_ticket;
  T *dot_result;
  i64 row_begin;  // first row handled by this launch (remainder launches)
  DiaXchg xc;
  int pdl;  // launched with programmatic stream serialization: wait before the first read of x / y / dotv
  // CG with the direction update folded into the product (dia_bulk_kernel<..., FUSED = true>, cg.cu): the operand is
  // not read from x but rebuilt where it is gathered,  p[j] = fr[j] + beta * fp[j]  (p = z + beta p, z == r,
  // cusp/krylov/detail/cg.inl:97-99), and the owner of row i also stores p[i + fshift] into fpn.  All three are
  // windows [fhalo_lo | rows | fhalo_hi] in the operator's column coordinates; the halo rows of fpn are rebuilt
  // locally from the halo rows of fr / fp, so p itself never travels between GPUs.
  const T *fr, *fp;
  T *fpn;
  const T *fbeta;    // device scalar
  const int *fdone;  // device flag: the solve is over, do nothing
  int ffirst;        // first iteration: p = r
  i64 fshift, fhalo_lo, fhalo_hi;
  // tiles are handed to the persistent CTAs in runs of `run` consecutive tiles (CTA b: tiles [b*run, (b+1)*run), then
  // gridDim.x * run further on): the x lines a tile gathers at row +- R, 2R, ... are the ones its neighbours in the run
  // gather at offset 0, so they are L1 hits instead of another trip to L2
  int run;
};

constexpr int DIA_DU = 8;            // diagonals in flight per thread
constexpr int DIA_SMEM_OFFS = 2048;  // offsets cached in smem per chunk

// Row / column arithmetic is done in 32-bit unsigned: rows, cols < 2^31, so
// c = row + offset (mod 2^32) is a valid column  <=>  c < cols  (a true sum
// outside [0, 2^31) wraps to >= 2^31).  Loads are issued unconditionally on
// clamped addresses (predicated loads would need more predicate registers than
// exist and serialise the batch); the bounds decision is re-derived at the add.
template <typename T, int RPT>
__device__ __forceinline__ void dia_accumulate_group(const T *__restrict__ vals, i64 pitch,
                                                     const T *__restrict__ x, const int *s_off,
                                                     int d0, const unsigned (&rc)[RPT],
                                                     unsigned cols, T (&acc)[RPT]) {
  T v[DIA_DU][RPT], xv[DIA_DU][RPT];
  unsigned off[DIA_DU];
#pragma unroll
  for (int u = 0; u < DIA_DU; ++u) {
    off[u] = (unsigned)s_off[d0 + u];
    const T *slab = vals + (i64)(d0 + u) * pitch;
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      v[u][i] = ld_stream(slab + rc[i]);
      xv[u][i] = ld_ro(x + min(rc[i] + off[u], cols - 1));
    }
  }
#pragma unroll
  for (int u = 0; u < DIA_DU; ++u)
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      pin(v[u][i]);
      pin(xv[u][i]);
    }
#pragma unroll
  for (int u = 0; u < DIA_DU; ++u)
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const T t = acc[i] + v[u][i] * xv[u][i];
      acc[i] = (rc[i] + off[u] < cols) ? t : acc[i];
    }
}

template <typename T, int BLOCK, int RPT>
__global__ void __launch_bounds__(BLOCK) dia_ldg_kernel(DiaArgs<T> a) {
  __shared__ int s_off[DIA_SMEM_OFFS];
  __shared__ T s_red[32];

  const unsigned rows = (unsigned)a.rows, cols = (unsigned)a.cols;
  const unsigned base = (unsigned)a.row_begin + blockIdx.x * (unsigned)(BLOCK * RPT) + threadIdx.x;
  T acc[RPT];
  unsigned rc[RPT];
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    const unsigned r = base + i * BLOCK;
    rc[i] = min(r, rows - 1);
    acc[i] = (a.accumulate && r < rows) ? a.y[r] : T(0);
  }

  for (int dc = 0; dc < a.ndiag; dc += DIA_SMEM_OFFS) {
    const int nd = min(DIA_SMEM_OFFS, a.ndiag - dc);
    if (dc > 0) __syncthreads();
    for (int d = threadIdx.x; d < nd; d += BLOCK) s_off[d] = a.offs[dc + d];
    __syncthreads();
    const T *vals = a.vals + (i64)dc * a.pitch;
    int d0 = 0;
    for (; d0 + DIA_DU <= nd; d0 += DIA_DU)
      dia_accumulate_group<T, RPT>(vals, a.pitch, a.x, s_off, d0, rc, cols, acc);
    for (; d0 < nd; ++d0) {  // ragged tail of the diagonal list
      const unsigned off = (unsigned)s_off[d0];
      const T *slab = vals + (i64)d0 * a.pitch;
      T v[RPT], xv[RPT];
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        v[i] = ld_stream(slab + rc[i]);
        xv[i] = ld_ro(a.x + min(rc[i] + off, cols - 1));
      }
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const T t = acc[i] + v[i] * xv[i];
        acc[i] = (rc[i] + off < cols) ? t : acc[i];
      }
    }
  }

  T dsum = 0;
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    const unsigned r = base + i * BLOCK;
    if (r < rows) {
      a.y[r] = acc[i];
      if (a.dotv) dsum = dsum + acc[i] * ld_ro(a.dotv + r);
    }
  }
  if (a.dotv) {  // uniform branch
    T bs = block_sum<BLOCK>(dsum, s_red);
    grid_reduce_finish<BLOCK>(bs, a.dot_partials, a.dot_ticket, s_red,
                              [&](T total) { *a.dot_result = total; });
  }
}

// ---------------------------------------------------------------------------
// bulk-async (TMA) staged variant
// ---------------------------------------------------------------------------
constexpr int DIA_KC = 8;  // diagonals per stage

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void xchg_copy16(char *dst, const char *src, size_t bytes, size_t tid, size_t nthreads) {
  const size_t n16 = ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) ? 0 : bytes / 16;
  const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
  uint4 *d4 = reinterpret_cast<uint4 *>(dst);
  for (size_t i = tid; i < n16; i += nthreads) d4[i] = s4[i];
  for (size_t i = n16 * 16 + tid; i < bytes; i += nthreads) dst[i] = src[i];
}

// lanes 1..31 of the producer warp of every CTA (see DiaXchg)
__device__ __forceinline__ void dia_xchg_aux(const DiaXchg &xc, int lane) {
  constexpr unsigned AUX = 0xfffffffeu;
  const size_t aux = (size_t)blockIdx.x * 31 + (lane - 1), naux = (size_t)gridDim.x * 31;
  const size_t par = (size_t)(xc.epoch & 1) * 2 * P2P_STAGE_SIDE;
  char *local = xc.window + xc.local_off;
  // A: my low edge -> rank-1's "from rank+1" slot, my high edge -> rank+1's "from rank-1" slot
  if (xc.stage_lo_nbr) xchg_copy16(xc.stage_lo_nbr + par + P2P_STAGE_SIDE, local, xc.lo_bytes, aux, naux);
  if (xc.stage_hi_nbr) xchg_copy16(xc.stage_hi_nbr + par, local + xc.n_bytes - xc.hi_bytes, xc.hi_bytes, aux, naux);
  __threadfence_system();
  __syncwarp(AUX);
  if (lane == 1) {
    if (atomicAdd(&xc.tickets[0], 1u) == gridDim.x - 1) {  // every CTA has pushed: publish
      __threadfence_system();
      xc.tickets[0] = 0;
      if (xc.mail_lo_nbr)
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&xc.mail_lo_nbr->xchg_flag[1]), "l"(xc.epoch) : "memory");
      if (xc.mail_hi_nbr)
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&xc.mail_hi_nbr->xchg_flag[0]), "l"(xc.epoch) : "memory");
    }
    // B: the neighbours' planes have landed in my staging buffer
    for (int side = 0; side < 2; ++side) {
      if (!(side == 0 ? xc.mail_lo_nbr : xc.mail_hi_nbr)) continue;
      unsigned long long v;
      SpinGuard guard;
      do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(&xc.mine->xchg_flag[side]) : "memory");
      } while (v < xc.epoch && !guard.expired(xc.mine));  // monotonic epochs: a neighbour may be one exchange ahead
    }
  }
  __syncwarp(AUX);
  // C: staging -> halo regions of my window
  if (xc.mail_lo_nbr) xchg_copy16(xc.window, xc.stage_mine + par, xc.lo_bytes, aux, naux);
  if (xc.mail_hi_nbr) xchg_copy16(local + xc.n_bytes, xc.stage_mine + par + P2P_STAGE_SIDE, xc.hi_bytes, aux, naux);
  __threadfence();
  __syncwarp(AUX);
  if (lane == 1 && atomicAdd(&xc.tickets[1], 1u) == gridDim.x - 1) {
    __threadfence();
    xc.tickets[1] = 0;
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(&xc.mine->xchg_go), "l"(xc.epoch) : "memory");
  }
}

template <typename T, int BLOCK, int RPT, bool FUSED = false>
__global__ void __launch_bounds__(BLOCK + 32) dia_bulk_kernel(DiaArgs<T> a, int stages,
                                                              i64 num_tiles) {
  constexpr int R = BLOCK * RPT;
  if (FUSED && *a.fdone) return;  // uniform: the monitor has stopped the solve
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // layout: [stages][KC][R] T | full[stages] | empty[stages] | offs[ndiag]
  T *s_vals = reinterpret_cast<T *>(smem_raw);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)stages * DIA_KC * R * sizeof(T));
  uint64_t *empty = full + stages;
  int *s_off = reinterpret_cast<int *>(empty + stages);
  __shared__ T s_red[32];

  const int tid = threadIdx.x;
  for (int d = tid; d < a.ndiag; d += BLOCK + 32) s_off[d] = a.offs[d];
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], BLOCK / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();

  const int nchunks = (a.ndiag + DIA_KC - 1) / DIA_KC;
  const unsigned rows = (unsigned)a.rows, cols = (unsigned)a.cols;
  T dsum = 0;
  // FUSED: the operand element at window index j (first iteration: beta = 0 and fp = fr, so p = r without a branch)
  const T fbeta = FUSED ? *a.fbeta : T(0);
  const T *const opx = FUSED ? a.fr : a.x;
  const T *const opp = a.fp;
  auto operand = [opx, opp, fbeta](unsigned j) -> T {
    if constexpr (!FUSED) {
      return ld_ro(opx + j);
    } else {
      const T rv = ld_ro(opx + j);
      const T pv = ld_ro(opp + j);
      return T(1) * rv + fbeta * pv;
    }
  };
  // runs of consecutive tiles per CTA (DiaArgs::run): the plain product is indifferent to them, so it keeps the
  // one-tile stride at compile time
  const i64 run = FUSED ? (i64)a.run : 1;  // a power of two
  const i64 seq_first = FUSED ? (i64)blockIdx.x * run : (i64)blockIdx.x;
  auto seq_next = [run](i64 seq) -> i64 {
    if constexpr (!FUSED) {
      return seq + gridDim.x;
    } else {
      return ((seq + 1) & (run - 1)) ? seq + 1 : seq + 1 + (i64)(gridDim.x - 1) * run;
    }
  };

  if (tid >= BLOCK) {
    // ===== producer warp: one elected lane drives the TMA engine =====
    if (tid == BLOCK) {
      const uint64_t pol = l2_policy_evict_first();
      int s = 0;
      uint32_t ph = 0;
      for (i64 seq = seq_first; seq < num_tiles; seq = seq_next(seq)) {
        i64 tile = seq + a.xc.rot;
        if (tile >= num_tiles) tile -= num_tiles;
        const i64 r0 = a.row_begin + tile * R;
        if (r0 + R > a.rows) continue;  // ragged last tile: consumers load it directly
        for (int c = 0; c < nchunks; ++c) {
          const int d0 = c * DIA_KC;
          const int kc = min(DIA_KC, a.ndiag - d0);
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], (uint32_t)(kc * R * sizeof(T)));
          for (int u = 0; u < kc; ++u)
            bulk_g2s(s_vals + ((size_t)s * DIA_KC + u) * R, a.vals + (i64)(d0 + u) * a.pitch + r0,
                     (uint32_t)(R * sizeof(T)), &full[s], pol);
          if (++s == stages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    } else if (a.xc.enabled == 1) {
      if (a.pdl) pdl_wait();
      dia_xchg_aux(a.xc, tid - BLOCK);
    }
  } else {
    // ===== consumer warps =====
    // PDL: the producer lane above is already streaming matrix slabs (never written by a predecessor in a solver
    // loop) into the ring; x, y and dotv may only be read once the predecessor has completed
    if (a.pdl) {
      pdl_wait();
      pdl_trigger();
    }
    int s = 0;
    uint32_t ph = 0;
    const int lane = tid & 31;
    bool ready_lo = !a.xc.enabled, ready_hi = !a.xc.enabled;
    for (i64 seq = seq_first; seq < num_tiles; seq = seq_next(seq)) {
      i64 tile = seq + a.xc.rot;
      if (tile >= num_tiles) tile -= num_tiles;
      // first tile of this warp that reads halo columns: the planes must have landed
      if (!ready_lo && tile < a.xc.lo_tiles) {
        if (lane == 0) {
          SpinGuard guard;
          while (ld_acquire_sys_u64(a.xc.wait_lo) < a.xc.wait_epoch && !guard.expired(a.xc.mine)) {
          }
        }
        __syncwarp();
        ready_lo = true;
      }
      if (!ready_hi && tile >= a.xc.hi_tile_begin) {
        if (lane == 0) {
          SpinGuard guard;
          while (ld_acquire_sys_u64(a.xc.wait_hi) < a.xc.wait_epoch && !guard.expired(a.xc.mine)) {
          }
        }
        __syncwarp();
        ready_hi = true;
      }
      if (FUSED) {
        // the halo rows of the new direction, rebuilt from the neighbours' r planes (which the waits above have seen
        // arrive) by the tiles that sit next to them
        if (tile < a.xc.lo_tiles) {
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            const i64 q = tile * R + tid + i * BLOCK;
            if (q < a.fhalo_lo) a.fpn[q] = operand((unsigned)q);
          }
        }
        if (tile >= a.xc.hi_tile_begin) {
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            const i64 q = (tile - a.xc.hi_tile_begin) * R + tid + i * BLOCK;
            if (q < a.fhalo_hi) a.fpn[a.fshift + a.rows + q] = operand((unsigned)(a.fshift + a.rows + q));
          }
        }
      }
      const unsigned r0 = (unsigned)(a.row_begin + tile * R);
      T acc[RPT];
      if ((i64)r0 + R > a.rows) {
        // ragged last tile (< R rows): same arithmetic, slabs straight from global
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const unsigned r = r0 + tid + i * BLOCK;
          if (r >= rows) continue;
          T s_acc = a.accumulate ? a.y[r] : T(0);
          for (int d = 0; d < a.ndiag; ++d) {
            const unsigned cidx = r + (unsigned)s_off[d];
            if (cidx < cols)
              s_acc = s_acc + ld_stream(a.vals + (i64)d * a.pitch + r) * operand(cidx);
          }
          a.y[r] = s_acc;
          if (FUSED) {
            const T pn = operand(r + (unsigned)a.fshift);
            a.fpn[r + a.fshift] = pn;
            dsum = dsum + s_acc * pn;
          } else if (a.dotv) {
            dsum = dsum + s_acc * ld_ro(a.dotv + r);
          }
        }
        continue;
      }
#pragma unroll
      for (int i = 0; i < RPT; ++i) acc[i] = a.accumulate ? a.y[r0 + tid + i * BLOCK] : T(0);
      for (int c = 0; c < nchunks; ++c) {
        const int d0 = c * DIA_KC;
        const int kc = min(DIA_KC, a.ndiag - d0);
        // x for this chunk is requested before waiting on the slab stage
        T xv[DIA_KC][RPT];
        unsigned off[DIA_KC];
#pragma unroll
        for (int u = 0; u < DIA_KC; ++u) {
          off[u] = (unsigned)s_off[min(d0 + u, a.ndiag - 1)];
#pragma unroll
          for (int i = 0; i < RPT; ++i)
            {
              const unsigned j = min(r0 + tid + i * BLOCK + off[u], cols - 1);
              // (the plain product reads a.x straight from the parameter bank: routing it through the lambda's
              // captured pointer costs registers in the headline kernel)
              xv[u][i] = FUSED ? operand(j) : ld_ro(a.x + j);
            }
        }
        mbar_wait(&full[s], ph);
        const T *sv = s_vals + (size_t)s * DIA_KC * R;
#pragma unroll
        for (int u = 0; u < DIA_KC; ++u)
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            pin(xv[u][i]);
            const T t = acc[i] + sv[u * R + tid + i * BLOCK] * xv[u][i];
            acc[i] = (u < kc && r0 + tid + i * BLOCK + off[u] < cols) ? t : acc[i];
          }
#pragma unroll
        for (int i = 0; i < RPT; ++i) consume_before_release(acc[i]);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (++s == stages) {
          s = 0;
          ph ^= 1;
        }
      }
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const unsigned r = r0 + tid + i * BLOCK;
        a.y[r] = acc[i];
        if (FUSED) {
          const T pn = operand(r + (unsigned)a.fshift);
          a.fpn[r + a.fshift] = pn;
          dsum = dsum + acc[i] * pn;
        } else if (a.dotv) {
          dsum = dsum + acc[i] * ld_ro(a.dotv + r);
        }
      }
    }
  }
  if (FUSED || a.dotv) {
    if (tid >= BLOCK) dsum = 0;
    T bs = block_sum<BLOCK + 32>(dsum, s_red);
    grid_reduce_finish<BLOCK + 32>(bs, a.dot_partials, a.dot_ticket, s_red,
                                   [&](T total) { *a.dot_result = total; });
  }
}

// ---------------------------------------------------------------------------
// host dispatch
// ---------------------------------------------------------------------------
template <typename T, int BLOCK, int RPT>
static b200sp_status launch_ldg(b200sp_handle h, cudaStream_t st, DiaArgs<T> a, i64 nrows) {
  const i64 grid = ceil_div(nrows, (i64)BLOCK * RPT);
  if (grid == 0) return B200SP_OK;
  const T *late_dotv = nullptr;  // one partial per CTA: beyond the workspace the dot runs after the product
  if (a.dotv && grid > RED_MAX_PARTIALS) {
    late_dotv = a.dotv;
    a.dotv = nullptr;
  }
  dia_ldg_kernel<T, BLOCK, RPT><<<(unsigned)grid, BLOCK, 0, st>>>(a);
  B200SP_LAUNCH_CHECK(h, "dia_ldg_kernel");
  if (late_dotv) return reduce<T, 0>(h, st, nrows, a.y + a.row_begin, late_dotv + a.row_begin, a.dot_result, nullptr);
  return B200SP_OK;
}

template <typename T>
static b200sp_status dispatch_ldg(b200sp_handle h, cudaStream_t st, const DiaArgs<T> &a, i64 nrows,
                                  int block, int rpt) {
#define CASE(B, R) \
  if (block == B && rpt == R) return launch_ldg<T, B, R>(h, st, a, nrows);
  CASE(128, 1) CASE(128, 2) CASE(128, 4) CASE(256, 1) CASE(256, 2) CASE(256, 4) CASE(512, 1)
  CASE(512, 2) CASE(512, 4)
#undef CASE
  return set_error(h, B200SP_INVALID_INPUT, "dia ldg: unsupported block_size=%d unroll=%d", block, rpt);
}

template <typename T, int BLOCK, int RPT, bool FUSED = false>
static b200sp_status launch_bulk(b200sp_handle h, cudaStream_t st, DiaArgs<T> a, i64 num_tiles,
                                 int stages, int ctas_per_sm) {
  constexpr int R = BLOCK * RPT;
  auto kern = dia_bulk_kernel<T, BLOCK, RPT, FUSED>;
  size_t smem = (size_t)stages * DIA_KC * R * sizeof(T) + 2 * stages * sizeof(uint64_t) +
                (size_t)a.ndiag * sizeof(int) + 16;
  if (smem > (size_t)h->max_smem_optin)
    return set_error(h, B200SP_INVALID_INPUT, "dia bulk: %zu B smem exceeds %d", smem, h->max_smem_optin);
  B200SP_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // persistent grid: no more CTAs than are resident at once (a second wave would start
  // only after a first-wave CTA has finished all of its tiles)
  int resident = 0;
  B200SP_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, BLOCK + 32, smem));
  if (resident < 1) return set_error(h, B200SP_INVALID_INPUT, "dia bulk: configuration does not fit on an SM");
  i64 grid = (i64)h->num_sms * (ctas_per_sm < resident ? ctas_per_sm : resident);
  // runs of consecutive tiles per CTA, as long as every CTA still gets several runs (load balance).  Measured on
  // poisson7pt (B200SP_DIA_RUN): the plain product is indifferent up to 4 and loses 4 % at 16 (HBM-bound either way:
  // it keeps run = 1); the fused-direction product is best at 4 (3.08 ms per 512^3 iteration against 3.16 at 1, 3.25 at 16)
  int run = 1;  // (runs change which CTA sums which tiles of the fused dot product: 1 keeps the bits of the plain form)
  if (const char *e = getenv("B200SP_DIA_RUN")) run = (FUSED && atoi(e) > 0) ? atoi(e) : run;
  while (run & (run - 1)) run &= run - 1;  // a power of two
  while (run > 1 && num_tiles < grid * run * 4) run /= 2;
  a.run = run;
  if (grid > ceil_div(num_tiles, (i64)run)) grid = ceil_div(num_tiles, (i64)run);
  if ((FUSED || a.dotv) && grid > RED_MAX_PARTIALS) return set_error(h, B200SP_INVALID_INPUT, "dia: grid too large");
  a.pdl = (h->pdl_spmv && !FUSED) ? 1 : 0;
  B200SP_CUDA(h, launch_kernel_pdl(kern, dim3((unsigned)grid), dim3(BLOCK + 32), smem, st, a.pdl != 0, a, stages, num_tiles));
  h->launches++;
  return B200SP_OK;
}

template <typename T>
static b200sp_status dispatch_bulk(b200sp_handle h, cudaStream_t st, const DiaArgs<T> &a, i64 num_tiles,
                                   int block, int rpt, int stages, int cps) {
#define CASE(B, R) \
  if (block == B && rpt == R) return launch_bulk<T, B, R>(h, st, a, num_tiles, stages, cps);
  CASE(128, 2) CASE(128, 4) CASE(128, 8) CASE(256, 1) CASE(256, 2) CASE(256, 4)
#undef CASE
  return set_error(h, B200SP_INVALID_INPUT, "dia bulk: unsupported block_size=%d unroll=%d", block, rpt);
}

template <typename T>
static b200sp_status dispatch_bulk_fused(b200sp_handle h, cudaStream_t st, const DiaArgs<T> &a, i64 num_tiles,
                                         int block, int rpt, int stages, int cps) {
#define CASE(B, R) \
  if (block == B && rpt == R) return launch_bulk<T, B, R, true>(h, st, a, num_tiles, stages, cps);
  CASE(128, 2) CASE(128, 4) CASE(256, 1) CASE(256, 2)
#undef CASE
  return set_error(h, B200SP_INVALID_INPUT, "dia bulk (fused direction): unsupported block_size=%d unroll=%d", block, rpt);
}

// Engine defaults (overridden by cfg / tuning cache).  Chosen on B200 from the
// round-1 sweep in profiles/.
static void dia_defaults(b200sp_cfg &c, size_t elem) {
  // round-1 sweep on B200, poisson7pt 256^3 (profiles/r01_probe_256.md): the
  // bulk-async kernel wins for both value types (6.0 / 6.5 TB/s vs 4.8 / 6.3 LDG)
  if (c.kernel == 0) c.kernel = B200SP_K_DIA_BULK;
  if (c.block_size == 0) c.block_size = 128;
  if (c.unroll == 0) c.unroll = (elem == 4) ? 4 : 2;
  if (c.stages == 0) c.stages = 2;
  if (c.ctas_per_sm == 0) c.ctas_per_sm = 4;
}

template <typename T>
b200sp_status spmv_dia_xchg(b200sp_handle h, cudaStream_t st, i64 rows, i64 cols, i64 ndiag, i64 pitch,
                            const int *offs, const T *vals, const T *x, T *y, int accumulate,
                            const b200sp_cfg *cfg, const T *dotv, T *dot_result, const DiaXchg *xchg,
                            int *xchg_fused);

// would spmv_dia() with this configuration run the bulk kernel (the one that can carry
// a fused halo exchange)?
bool dia_can_fuse_xchg(i64 rows, i64 ndiag, i64 pitch, const void *vals, size_t elem, const b200sp_cfg *cfg) {
  b200sp_cfg c = cfg ? *cfg : b200sp_cfg{};
  dia_defaults(c, elem);
  if (c.kernel != B200SP_K_DIA_BULK || rows <= 0 || ndiag <= 0) return false;
  const int R = c.block_size * c.unroll;
  return (pitch * elem) % 16 == 0 && aligned16(vals) && ((size_t)R * elem) % 16 == 0;
}

// does spmv_dia_fused_direction() have an instantiation for this configuration?
bool dia_can_fuse_direction(i64 rows, i64 ndiag, i64 pitch, const void *vals, size_t elem, const b200sp_cfg *cfg) {
  if (!dia_can_fuse_xchg(rows, ndiag, pitch, vals, elem, cfg)) return false;
  b200sp_cfg c = cfg ? *cfg : b200sp_cfg{};
  dia_defaults(c, elem);
  return (c.block_size == 128 && (c.unroll == 2 || c.unroll == 4)) || (c.block_size == 256 && (c.unroll == 1 || c.unroll == 2));
}

// y = A p with p[j] = r[j] + beta p_old[j] rebuilt on the fly, p stored into p_new (window coordinates, own rows at
// [shift, shift + rows) plus both halos), <y, p> into *dot_result: the product and the direction update of one CG
// iteration in one pass (see DiaArgs).  `xchg` (enabled == 2) makes the tiles next to the halos wait for the
// neighbours' r planes.
template <typename T>
b200sp_status spmv_dia_fused_direction(b200sp_handle h, cudaStream_t st, i64 rows, i64 cols, i64 ndiag, i64 pitch,
                                       const int *offs, const T *vals, T *y, const b200sp_cfg *cfg, const T *r_win,
                                       const T *p_old_win, T *p_new_win, const T *beta, const int *done, int first,
                                       i64 shift, i64 halo_lo, i64 halo_hi, T *dot_result, const DiaXchg *xchg) {
  B200SP_CHECK_HANDLE(h);
  B200SP_REQUIRE(h, rows > 0 && ndiag > 0 && rows < (1ll << 31) && cols < (1ll << 31) && pitch >= rows,
                 "dia (fused direction): bad dimensions");
  B200SP_REQUIRE(h, y && offs && vals && r_win && p_old_win && p_new_win && beta && done && dot_result,
                 "dia (fused direction): null pointer");
  B200SP_REQUIRE(h, cols == shift + rows + halo_hi && shift == halo_lo, "dia (fused direction): window does not match the operator");
  b200sp_cfg c = cfg ? *cfg : b200sp_cfg{};
  dia_defaults(c, sizeof(T));
  if (const char *e = getenv("B200SP_DIA_FUSED_SHAPE")) {  // "block,unroll,stages,ctas" (experiments)
    int b = 0, u = 0, sg = 0, cp = 0;
    if (sscanf(e, "%d,%d,%d,%d", &b, &u, &sg, &cp) == 4) {
      c.block_size = b; c.unroll = u; c.stages = sg; c.ctas_per_sm = cp;
    }
  }
  B200SP_REQUIRE(h, dia_can_fuse_direction(rows, ndiag, pitch, vals, sizeof(T), &c), "dia (fused direction): configuration not supported");
  DiaArgs<T> a;
  memset(&a, 0, sizeof(a));
  a.rows = rows; a.cols = cols; a.pitch = pitch; a.ndiag = (int)ndiag;
  a.offs = offs; a.vals = vals; a.x = nullptr; a.y = y; a.accumulate = 0;
  a.dotv = nullptr; a.dot_result = dot_result;
  a.dot_partials = reinterpret_cast<T *>(h->red_partials);
  a.dot_ticket = h->red_counters;
  a.fr = r_win; a.fp = first ? r_win : p_old_win; a.fpn = p_new_win; a.fbeta = beta; a.fdone = done; a.ffirst = first;
  a.fshift = shift; a.fhalo_lo = halo_lo; a.fhalo_hi = halo_hi;
  const int R = c.block_size * c.unroll;
  const i64 tiles = ceil_div(rows, (i64)R);
  if (xchg && xchg->enabled) a.xc = *xchg;
  a.xc.lo_tiles = ceil_div(halo_lo, (i64)R);
  a.xc.hi_tile_begin = halo_hi > 0 ? (rows - halo_hi) / R : tiles;
  if (a.xc.hi_tile_begin < 0) a.xc.hi_tile_begin = 0;
  a.xc.rot = a.xc.lo_tiles;
  if (!a.xc.enabled || a.xc.lo_tiles + (tiles - a.xc.hi_tile_begin) >= tiles) a.xc.rot = 0;
  if (a.xc.rot >= tiles) a.xc.rot = 0;
  return dispatch_bulk_fused<T>(h, st, a, tiles, c.block_size, c.unroll, c.stages, c.ctas_per_sm);
}

template b200sp_status spmv_dia_fused_direction<float>(b200sp_handle, cudaStream_t, i64, i64, i64, i64, const int *,
                                                       const float *, float *, const b200sp_cfg *, const float *,
                                                       const float *, float *, const float *, const int *, int, i64, i64,
                                                       i64, float *, const DiaXchg *);
template b200sp_status spmv_dia_fused_direction<double>(b200sp_handle, cudaStream_t, i64, i64, i64, i64, const int *,
                                                        const double *, double *, const b200sp_cfg *, const double *,
                                                        const double *, double *, const double *, const int *, int, i64,
                                                        i64, i64, double *, const DiaXchg *);

template <typename T>
b200sp_status spmv_dia(b200sp_handle h, cudaStream_t st, i64 rows, i64 cols, i64 ndiag, i64 pitch,
                       const int *offs, const T *vals, const T *x, T *y, int accumulate,
                       const b200sp_cfg *cfg, const T *dotv, T *dot_result) {
  return spmv_dia_xchg<T>(h, st, rows, cols, ndiag, pitch, offs, vals, x, y, accumulate, cfg, dotv, dot_result,
                          nullptr, nullptr);
}

// same, with an optional halo exchange fused into the kernel.  *xchg_fused is set to 1
// when the launched kernel performs the exchange; otherwise the caller must have
// exchanged the halos itself before the call... so the caller asks first with
// dia_can_fuse_xchg() and only then passes `xchg`.
template <typename T>
b200sp_status spmv_dia_xchg(b200sp_handle h, cudaStream_t st, i64 rows, i64 cols, i64 ndiag, i64 pitch,
                            const int *offs, const T *vals, const T *x, T *y, int accumulate,
                            const b200sp_cfg *cfg, const T *dotv, T *dot_result, const DiaXchg *xchg,
                            int *xchg_fused) {
  B200SP_CHECK_HANDLE(h);
  B200SP_REQUIRE(h, rows >= 0 && cols >= 0 && ndiag >= 0, "dia: negative dimension");
  B200SP_REQUIRE(h, rows < (1ll << 31) && cols < (1ll << 31), "dia: int32 index range");
  B200SP_REQUIRE(h, pitch >= rows, "dia: pitch < num_rows");
  if (rows == 0) {
    if (dot_result) B200SP_CUDA(h, cudaMemsetAsync(dot_result, 0, sizeof(T), st));
    return B200SP_OK;
  }
  B200SP_REQUIRE(h, y != nullptr && (ndiag == 0 || (offs && vals && x)), "dia: null pointer");

  b200sp_cfg c = cfg ? *cfg : b200sp_cfg{};
  dia_defaults(c, sizeof(T));

  DiaArgs<T> a;
  a.rows = rows; a.cols = cols; a.pitch = pitch; a.ndiag = (int)ndiag;
  a.offs = offs; a.vals = vals; a.x = x; a.y = y; a.accumulate = accumulate;
  a.dotv = dotv; a.dot_result = dot_result;
  a.dot_partials = reinterpret_cast<T *>(h->red_partials);
  a.dot_ticket = h->red_counters;
  a.row_begin = 0;
  a.pdl = 0;
  memset(&a.xc, 0, sizeof(a.xc));

  if (c.kernel == B200SP_K_DIA_BULK) {
    const int R = c.block_size * c.unroll;
    const bool ok = (pitch * sizeof(T)) % 16 == 0 && aligned16(vals) && ((size_t)R * sizeof(T)) % 16 == 0 &&
                    ndiag > 0;
    const i64 tiles = ceil_div(rows, (i64)R);
    if (ok && xchg && xchg->enabled) {
      // fused halo exchange: the tiles that read halo columns are visited last
      a.xc = *xchg;
      const i64 lo_rows = (i64)(a.xc.lo_bytes / sizeof(T));  // 0 where there is no neighbour
      const i64 hi_rows = (i64)(a.xc.hi_bytes / sizeof(T));
      a.xc.lo_tiles = ceil_div(lo_rows, (i64)R);
      a.xc.hi_tile_begin = hi_rows > 0 ? (rows - hi_rows) / R : tiles;
      a.xc.rot = a.xc.lo_tiles;
      if (a.xc.lo_tiles + (tiles - a.xc.hi_tile_begin) >= tiles) a.xc.rot = 0;  // no interior to hide behind
      *xchg_fused = 1;
    }
    if (ok) return dispatch_bulk<T>(h, st, a, tiles, c.block_size, c.unroll, c.stages, c.ctas_per_sm);
    // layout not bulk-copyable -> LDG kernel (same results)
    c.kernel = B200SP_K_DIA_LDG;
    if (c.unroll > 4) c.unroll = 4;
  }
  if (c.kernel != B200SP_K_DIA_LDG)
    return set_error(h, B200SP_INVALID_INPUT, "dia: unknown kernel id %d", c.kernel);
  return dispatch_ldg<T>(h, st, a, rows, c.block_size, c.unroll);
}

template b200sp_status spmv_dia<float>(b200sp_handle, cudaStream_t, i64, i64, i64, i64, const int *,
                                       const float *, const float *, float *, int, const b200sp_cfg *,
                                       const float *, float *);
template b200sp_status spmv_dia<double>(b200sp_handle, cudaStream_t, i64, i64, i64, i64, const int *,
                                        const double *, const double *, double *, int,
                                        const b200sp_cfg *, const double *, double *);

template b200sp_status spmv_dia_xchg<float>(b200sp_handle, cudaStream_t, i64, i64, i64, i64, const int *,
                                            const float *, const float *, float *, int, const b200sp_cfg *,
                                            const float *, float *, const DiaXchg *, int *);
template b200sp_status spmv_dia_xchg<double>(b200sp_handle, cudaStream_t, i64, i64, i64, i64, const int *,
                                             const double *, const double *, double *, int, const b200sp_cfg *,
                                             const double *, double *, const DiaXchg *, int *);

}  // namespace b200sp

extern "C" {
b200sp_status b200sp_spmv_dia_f32(b200sp_handle h, b200sp_stream stream, int64_t num_rows,
                                  int64_t num_cols, int64_t num_diagonals, int64_t pitch,
                                  const int32_t *diagonal_offsets, const float *values, const float *x,
                                  float *y, int accumulate, const b200sp_cfg *cfg) {
  return b200sp::spmv_dia<float>(h, (cudaStream_t)stream, num_rows, num_cols, num_diagonals, pitch,
                                 diagonal_offsets, values, x, y, accumulate, cfg, nullptr, nullptr);
}
b200sp_status b200sp_spmv_dia_f64(b200sp_handle h, b200sp_stream stream, int64_t num_rows,
                                  int64_t num_cols, int64_t num_diagonals, int64_t pitch,
                                  const int32_t *diagonal_offsets, const double *values, const double *x,
                                  double *y, int accumulate, const b200sp_cfg *cfg) {
  return b200sp::spmv_dia<double>(h, (cudaStream_t)stream, num_rows, num_cols, num_diagonals, pitch,
                                  diagonal_offsets, values, x, y, accumulate, cfg, nullptr, nullptr);
}
}
