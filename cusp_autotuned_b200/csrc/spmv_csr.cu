// spmv_csr.cu — CSR SpMV for sm_100a.
//
// Replaces spmv_csr_vector_kernel (cusp/system/cuda/detail/multiply/
// csr_vector_spmv.h:66-161; volatile-smem warp-synchronous reduction, undefined
// on Volta+), spmv_csr_scalar_kernel (csr_scalar.h:48-73) and the KTT kernels
// csr_kernel_{naive,warp} (cusp/system/cuda/ktt/kernels/csr_kernel.h:160-235).
// Semantics: host loop cusp/system/detail/sequential/multiply/csr_spmv.h:35-74
//     acc = init(y[i]); for jj in row i: acc += Ax[jj]*x[Aj[jj]]; y[i] = acc
//
//  K_CSR_VECTOR : a sub-warp of TPR in {1,2,4,8,16,32} lanes owns a row; lanes
//     stride the row, partial sums are combined with __shfl_down_sync (width TPR);
//     each sub-warp keeps RPT independent rows in flight (adjacent sub-warps take
//     adjacent rows, so a warp's loads of Aj/Ax cover one contiguous nnz range).
//     TPR == 1 is the scalar kernel and keeps the reference's summation order
//     (bit-identical with -fmad=false); TPR > 1 changes only the grouping.
//     Long rows continue in a 4x-unrolled strided loop.
//
//  K_CSR_STREAM : a CTA owns R consecutive rows (R chosen on the host so that the
//     block's nnz fill one shared-memory chunk).  The block's contiguous nnz range
//     [Ap[r0], Ap[r0+R)) is streamed with perfectly coalesced ld.global.cs loads
//     [Ap[r0], Ap[r0+R)) of Aj and Ax is staged into shared memory by two
//     cp.async.bulk copies (TMA engine, UBLKCP, L2 evict-first) regardless of row
//     boundaries; then one thread per row walks its entries in order, gathering x
//     with 8 independent ld.global.nc in flight: acc = init(y); acc += a0*x0; ...
//     — exactly the reference's order, so this kernel is bit-identical to the host
//     loop on any data.  Because lanes map to ROWS in the gather phase, a warp's
//     gather touches neighbouring x entries for banded matrices (ELL-like L1
//     behaviour) instead of the 5-10 cache lines a lane-per-entry mapping touches.
//     Row pieces longer than 128 entries inside a chunk are reduced by a whole warp
//     instead (power-law hubs; only the grouping of that piece changes).
//
// Algorithmic bytes: (rows+1)*4 + nnz*(4+sizeof(T)) + cols*sizeof(T) + rows*sizeof(T).
#include "common.cuh"

namespace b200sp {

// spmv_coo.cu: nnz-balanced segmented scan over a CSR matrix (K_CSR_BALANCED)
template <typename T>
b200sp_status spmv_csr_balanced(b200sp_handle h, cudaStream_t st, i64 rows, i64 cols, i64 nnz, const int *Ap,
                                const int *Aj, const T *Ax, const T *x, T *y, int accumulate, int block, int vpt);

template <typename T, int MODE>
b200sp_status reduce(b200sp_handle, cudaStream_t, i64, const T *, const T *, T *, T *);  // blas1.cu

template <typename T>
struct CsrArgs {
  i64 rows, cols, nnz;
  const int *Ap;
  const int *Aj;
  const T *Ax;
  const T *x;
  T *y;
  int accumulate;
  const T *dotv;
  T *dot_partials;
  unsigned int *dot_ticket;
  T *dot_result;
};

template <typename T, int BLOCK, int TPR, int RPT>
__global__ void __launch_bounds__(BLOCK) csr_vector_kernel(CsrArgs<T> a) {
  constexpr int VPB = BLOCK / TPR;  // rows in flight per CTA per step
  __shared__ T s_red[32];
  const int lane = threadIdx.x & (TPR - 1);
  const unsigned vec = threadIdx.x / TPR;
  const unsigned rows = (unsigned)a.rows, cols = (unsigned)a.cols;
  const int last = (int)(a.nnz - 1);
  const unsigned base = blockIdx.x * (unsigned)(VPB * RPT) + vec;

  int s[RPT], e[RPT];
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    const unsigned r = base + i * VPB;
    const unsigned rc = min(r, rows - 1);
    s[i] = ld_ro(a.Ap + rc);
    e[i] = ld_ro(a.Ap + rc + 1);
  }
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    pin(s[i]);
    pin(e[i]);
    if (base + i * VPB >= rows) e[i] = s[i];
  }

  T sum[RPT];
  {
    int c[RPT];
    T v[RPT], xv[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const int jc = min(s[i] + lane, last);
      c[i] = ld_stream(a.Aj + jc);
      v[i] = ld_stream(a.Ax + jc);
    }
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      pin(c[i]);
      xv[i] = ld_ro(a.x + min((unsigned)c[i], cols - 1));
    }
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      pin(xv[i]);
      T init = T(0);
      if (TPR == 1 && a.accumulate && base + i * VPB < rows) init = a.y[base + i * VPB];
      const T t = init + v[i] * xv[i];
      sum[i] = (s[i] + lane < e[i]) ? t : init;
    }
  }
  // rows longer than TPR
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    for (int jj = s[i] + lane + TPR; jj < e[i]; jj += 4 * TPR) {
      int c[4];
      T v[4], xv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int jc = min(jj + q * TPR, last);
        c[q] = ld_stream(a.Aj + jc);
        v[q] = ld_stream(a.Ax + jc);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        pin(c[q]);
        xv[q] = ld_ro(a.x + min((unsigned)c[q], cols - 1));
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const T t = sum[i] + v[q] * xv[q];
        sum[i] = (jj + q * TPR < e[i]) ? t : sum[i];
      }
    }
  }

  T dsum = 0;
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    const unsigned r = base + i * VPB;
    T tot = subwarp_sum<TPR>(sum[i]);
    if (lane == 0 && r < rows) {
      if (TPR > 1 && a.accumulate) tot = a.y[r] + tot;
      a.y[r] = tot;
      if (a.dotv) dsum = dsum + tot * ld_ro(a.dotv + r);
    }
  }
  if (a.dotv) {
    T bs = block_sum<BLOCK>(dsum, s_red);
    grid_reduce_finish<BLOCK>(bs, a.dot_partials, a.dot_ticket, s_red,
                              [&](T total) { *a.dot_result = total; });
  }
}

// ---------------------------------------------------------------------------
// K_CSR_STREAM
// ---------------------------------------------------------------------------
constexpr int CSR_LONG = 128;   // row pieces longer than this are reduced by a warp
constexpr int CSR_MAXQ = 96;    // >= CAP/CSR_LONG + 2 for every instantiated CAP
constexpr int CSR_GU = 8;       // x gathers in flight per thread in the row phase

template <typename T, int BLOCK, int NPT>
__global__ void __launch_bounds__(BLOCK) csr_stream_kernel(CsrArgs<T> a, int R, int tma_aligned) {
  constexpr int CAP = BLOCK * NPT;
  constexpr int EPV = 16 / (int)sizeof(T);  // values per 16 bytes
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T *s_val = reinterpret_cast<T *>(smem_raw);              // CAP + 16   values of the chunk (+ slack)
  int *s_col = reinterpret_cast<int *>(s_val + CAP + 16);    // CAP + 16   column indices  (+ slack)
  T *s_acc = reinterpret_cast<T *>(s_col + CAP + 16);        // R          row accumulators
  int *s_off = reinterpret_cast<int *>(s_acc + R);           // R + 1      row offsets
  __shared__ uint64_t s_bar;
  __shared__ int s_long[CSR_MAXQ];
  __shared__ int s_nlong;
  __shared__ T s_red[32];

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const i64 r0 = (i64)blockIdx.x * R;
  const int nr = (int)min((i64)R, a.rows - r0);
  const unsigned cols = (unsigned)a.cols;
  const int nnz = (int)a.nnz;

  if (tid == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
    s_nlong = 0;
  }
  for (int i = tid; i <= nr; i += BLOCK) s_off[i] = ld_ro(a.Ap + r0 + i);
  for (int i = tid; i < nr; i += BLOCK) s_acc[i] = a.accumulate ? a.y[r0 + i] : T(0);
  __syncthreads();
  const int s0 = s_off[0], s1 = s_off[nr];
  uint32_t parity = 0;

  for (int lo = s0; lo < s1; lo += CAP) {
    const int hi = min(lo + CAP, s1);
    // ---- phase 1: the chunk's contiguous (Aj, Ax) range -> shared memory ---------------
    // bulk-async copies start at the enclosing 16-byte boundary; `sc`/`sv` are the
    // resulting shifts.  The last chunk of the matrix (rounded end past nnz) and
    // unaligned base pointers use plain coalesced loads.
    const int ga_c = lo & ~3, ga_v = lo & ~(EPV - 1);
    const int n_c = ((hi - ga_c) + 3) & ~3, n_v = ((hi - ga_v) + EPV - 1) & ~(EPV - 1);
    int sc = 0, sv = 0;
    if (tma_aligned && ga_c + n_c <= nnz && ga_v + n_v <= nnz) {
      if (tid == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const uint64_t pol = l2_policy_evict_first();
        mbar_expect_tx(&s_bar, (uint32_t)(n_c * sizeof(int) + n_v * sizeof(T)));
        bulk_g2s(s_col, a.Aj + ga_c, (uint32_t)(n_c * sizeof(int)), &s_bar, pol);
        bulk_g2s(s_val, a.Ax + ga_v, (uint32_t)(n_v * sizeof(T)), &s_bar, pol);
      }
      sc = lo - ga_c;
      sv = lo - ga_v;
      if (w == 0) mbar_wait(&s_bar, parity);  // one warp polls, the rest sleep at the barrier
      parity ^= 1;
      __syncthreads();
    } else {
      for (int idx = tid; idx < hi - lo; idx += BLOCK) {
        s_col[idx] = ld_stream(a.Aj + lo + idx);
        s_val[idx] = ld_stream(a.Ax + lo + idx);
      }
      __syncthreads();
    }
    const int *pc = s_col + sc - lo;  // pc[j], pv[j] for absolute entry index j
    const T *pv = s_val + sv - lo;

    // ---- phase 2: one thread per row, entries in order (the reference's order) ---------
    for (int i = tid; i < nr; i += BLOCK) {
      const int b = max(s_off[i], lo), e = min(s_off[i + 1], hi);
      if (b < e) {
        if (e - b <= CSR_LONG) {
          T t = s_acc[i];
          // slots past `e` read stale shared memory inside the buffers' 16-entry slack:
          // their column is clamped to a valid address and their product discarded
          for (int j = b; j < e; j += CSR_GU) {
            T xv[CSR_GU];
#pragma unroll
            for (int q = 0; q < CSR_GU; ++q) xv[q] = ld_ro(a.x + min((unsigned)pc[j + q], cols - 1));
#pragma unroll
            for (int q = 0; q < CSR_GU; ++q) {
              pin(xv[q]);
              const T u = t + pv[j + q] * xv[q];
              t = (j + q < e) ? u : t;
            }
          }
          s_acc[i] = t;
        } else {
          s_long[atomicAdd(&s_nlong, 1)] = i;
        }
      }
    }
    __syncthreads();
    const int nlong = s_nlong;
    if (nlong > 0) {  // uniform: power-law hub pieces, one warp each
      for (int q = w; q < nlong; q += BLOCK / 32) {
        const int i = s_long[q];
        const int b = max(s_off[i], lo), e = min(s_off[i + 1], hi);
        T part = T(0);
        for (int j = b + lane; j < e; j += 32) part = part + pv[j] * ld_ro(a.x + min((unsigned)pc[j], cols - 1));
        part = warp_sum(part);
        if (lane == 0) s_acc[i] = s_acc[i] + part;
      }
      __syncthreads();
      if (tid == 0) s_nlong = 0;
      __syncthreads();
    }
  }

  T dsum = 0;
  for (int i = tid; i < nr; i += BLOCK) {
    const T t = s_acc[i];
    a.y[r0 + i] = t;
    if (a.dotv) dsum = dsum + t * ld_ro(a.dotv + r0 + i);
  }
  if (a.dotv) {
    T bs = block_sum<BLOCK>(dsum, s_red);
    grid_reduce_finish<BLOCK>(bs, a.dot_partials, a.dot_ticket, s_red,
                              [&](T total) { *a.dot_result = total; });
  }
}

// ---------------------------------------------------------------------------
// K_CSR_RING — persistent, pipelined version of the stream kernel.
//
// grid = ctas_per_sm x 148 persistent CTAs of BLOCK consumer threads + one
// producer warp.  Tiles of R consecutive rows (R <= BLOCK, chosen on the host so
// that a tile's nnz fill ~one ring stage) are dealt round-robin.  For each tile the
// producer thread issues two cp.async.bulk copies (TMA engine, L2 evict-first) of
// the tile's contiguous [Ap[r0], Ap[r0+R)) range of Aj and Ax into the next free
// stage of an mbarrier ring, up to `stages` tiles ahead of the consumers; tiles
// with more entries than a stage holds (power-law hubs) become several chunks.
// Consumers: thread-per-row in the reference's entry order (bit-identical to
// csr_spmv.h:35-74 with -fmad=false), x through ld.global.nc, 8 gathers in
// flight; next tile's row offsets are prefetched while the current one is
// processed.  A row piece longer than CSR_LONG inside a chunk is reduced by its
// whole warp (only the grouping of that piece changes).  The column ring is
// zero-filled once, so slots past a row's end always hold a valid column and the
// gather needs no clamp; their products are discarded by a select.
// ---------------------------------------------------------------------------
template <typename T, int BLOCK, int NPT>
__global__ void __launch_bounds__(BLOCK + 32) csr_ring_kernel(CsrArgs<T> a, int R, int stages, i64 num_tiles) {
  constexpr int CAP = BLOCK * NPT;   // entries per stage
  constexpr int STR = CAP + 16;      // stage stride (alignment shift + read-ahead slack)
  constexpr int EPV = 16 / (int)sizeof(T);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T *s_val = reinterpret_cast<T *>(smem_raw);
  int *s_col = reinterpret_cast<int *>(smem_raw + (size_t)stages * STR * sizeof(T));
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)stages * STR * (sizeof(T) + sizeof(int)));
  uint64_t *empty = full + stages;
  __shared__ T s_red[32];

  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], BLOCK / 32);
    }
    mbar_fence_init();
  }
  for (int i = tid; i < stages * STR; i += BLOCK + 32) s_col[i] = 0;
  __syncthreads();

  const i64 rows = a.rows;
  const int nnz = (int)a.nnz;
  T dsum = 0;

  if (tid >= BLOCK) {
    // ------------------------------ producer --------------------------------
    if (tid == BLOCK) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      const uint64_t pol = l2_policy_evict_first();
      const int nnz_c = nnz & ~3, nnz_v = nnz & ~(EPV - 1);  // last 16-byte-complete entry
      int s = 0;
      uint32_t ph = 0;
      i64 tile = blockIdx.x;
      int s0 = 0, s1 = 0;
      if (tile < num_tiles) {
        s0 = ld_ro(a.Ap + tile * R);
        s1 = ld_ro(a.Ap + min(tile * R + R, rows));
      }
      while (tile < num_tiles) {
        const i64 next = tile + gridDim.x;
        int n0 = 0, n1 = 0;
        if (next < num_tiles) {  // bounds of the next tile: in flight while this one is issued
          n0 = ld_ro(a.Ap + next * R);
          n1 = ld_ro(a.Ap + min(next * R + R, rows));
        }
        for (int lo = s0; lo < s1; lo += CAP) {
          const int hi = min(lo + CAP, s1);
          mbar_wait(&empty[s], ph ^ 1);
          int *dc = s_col + (size_t)s * STR;
          T *dv = s_val + (size_t)s * STR;
          const int ga_c = lo & ~3, ga_v = lo & ~(EPV - 1);
          const int end_c = min((hi + 3) & ~3, nnz_c), end_v = min((hi + EPV - 1) & ~(EPV - 1), nnz_v);
          // the (at most 3) entries after the last complete 16 bytes of the arrays
          for (int j = max(end_c, ga_c); j < hi; ++j) dc[j - ga_c] = a.Aj[j];
          for (int j = max(end_v, ga_v); j < hi; ++j) dv[j - ga_v] = a.Ax[j];
          const int bc = max(end_c - ga_c, 0), bv = max(end_v - ga_v, 0);
          mbar_expect_tx(&full[s], (uint32_t)(bc * sizeof(int) + bv * sizeof(T)));
          if (bc > 0) bulk_g2s(dc, a.Aj + ga_c, (uint32_t)(bc * sizeof(int)), &full[s], pol);
          if (bv > 0) bulk_g2s(dv, a.Ax + ga_v, (uint32_t)(bv * sizeof(T)), &full[s], pol);
          if (++s == stages) {
            s = 0;
            ph ^= 1;
          }
        }
        tile = next;
        s0 = n0;
        s1 = n1;
      }
    }
  } else {
    // ------------------------------ consumers -------------------------------
    const int lane = tid & 31;
    int s = 0;
    uint32_t ph = 0;
    i64 tile = blockIdx.x;
    int b_i = 0, e_i = 0, s0 = 0, s1 = 0;
    if (tile < num_tiles) {
      const i64 r = tile * R + tid;
      if (tid < R && r < rows) {
        b_i = ld_ro(a.Ap + r);
        e_i = ld_ro(a.Ap + r + 1);
      }
      s0 = ld_ro(a.Ap + tile * R);
      s1 = ld_ro(a.Ap + min(tile * R + R, rows));
    }
    while (tile < num_tiles) {
      const i64 r = tile * R + tid;
      const bool valid = tid < R && r < rows;
      const i64 next = tile + gridDim.x;
      int nb = 0, ne = 0, n0 = 0, n1 = 0;
      if (next < num_tiles) {
        const i64 rn = next * R + tid;
        if (tid < R && rn < rows) {
          nb = ld_ro(a.Ap + rn);
          ne = ld_ro(a.Ap + rn + 1);
        }
        n0 = ld_ro(a.Ap + next * R);
        n1 = ld_ro(a.Ap + min(next * R + R, rows));
      }
      T acc = (valid && a.accumulate) ? a.y[r] : T(0);
      for (int lo = s0; lo < s1; lo += CAP) {
        const int hi = min(lo + CAP, s1);
        mbar_wait(&full[s], ph);
        const int *pc = s_col + (size_t)s * STR + (lo & 3) - lo;           // pc[j], absolute entry index j
        const T *pv = s_val + (size_t)s * STR + (lo & (EPV - 1)) - lo;
        const int b = max(b_i, lo), e = min(e_i, hi);
        const int len = e - b;
        if (len > 0 && len <= CSR_LONG) {
          T t = acc;
          for (int j = b; j < e; j += CSR_GU) {
            int c[CSR_GU];
            T xv[CSR_GU];
#pragma unroll
            for (int q = 0; q < CSR_GU; ++q) c[q] = pc[j + q];
#pragma unroll
            for (int q = 0; q < CSR_GU; ++q) xv[q] = ld_ro(a.x + (unsigned)c[q]);
#pragma unroll
            for (int q = 0; q < CSR_GU; ++q) {
              pin(xv[q]);
              const T u = t + pv[j + q] * xv[q];
              t = (j + q < e) ? u : t;
            }
          }
          acc = t;
        }
        unsigned long_mask = __ballot_sync(0xffffffffu, len > CSR_LONG);
        while (long_mask) {  // warp-uniform: hub pieces, the whole warp on one row piece
          const int src = __ffs(long_mask) - 1;
          long_mask &= long_mask - 1;
          const int lb = __shfl_sync(0xffffffffu, b, src), le = __shfl_sync(0xffffffffu, e, src);
          T part = T(0);
          for (int j = lb + lane; j < le; j += 32) part = part + pv[j] * ld_ro(a.x + (unsigned)pc[j]);
          part = warp_sum(part);
          part = __shfl_sync(0xffffffffu, part, 0);
          if (lane == src) acc = acc + part;
        }
        consume_before_release(acc);  // every staged value of this piece has flowed into acc
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (++s == stages) {
          s = 0;
          ph ^= 1;
        }
      }
      if (valid) {
        a.y[r] = acc;
        if (a.dotv) dsum = dsum + acc * ld_ro(a.dotv + r);
      }
      tile = next;
      b_i = nb;
      e_i = ne;
      s0 = n0;
      s1 = n1;
    }
  }
  if (a.dotv) {
    if (tid >= BLOCK) dsum = 0;
    T bs = block_sum<BLOCK + 32>(dsum, s_red);
    grid_reduce_finish<BLOCK + 32>(bs, a.dot_partials, a.dot_ticket, s_red,
                                   [&](T total) { *a.dot_result = total; });
  }
}

template <typename T, int BLOCK, int NPT>
static b200sp_status launch_ring(b200sp_handle h, cudaStream_t st, CsrArgs<T> a, int stages, int ctas_per_sm) {
  constexpr int CAP = BLOCK * NPT;
  const double mean = (double)a.nnz / (double)a.rows;
  // rows per tile: ~90 % of a stage at the mean row length, one row per consumer thread at most
  i64 R = (i64)(0.9 * (double)CAP / (mean > 1.0 ? mean : 1.0));
  R = (R / 32) * 32;
  if (R < 32) R = 32;
  if (R > BLOCK) R = BLOCK;
  const i64 num_tiles = ceil_div(a.rows, R);
  const size_t smem = (size_t)stages * (CAP + 16) * (sizeof(T) + sizeof(int)) + 2 * (size_t)stages * sizeof(uint64_t) + 16;
  if (smem > (size_t)h->max_smem_optin)
    return set_error(h, B200SP_INVALID_INPUT, "csr ring: %zu B smem exceeds %d", smem, h->max_smem_optin);
  auto kern = csr_ring_kernel<T, BLOCK, NPT>;
  B200SP_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // persistent grid: never more CTAs than are resident at once (a second wave would
  // start only when a first-wave CTA has finished ALL of its tiles)
  int resident = 0;
  B200SP_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, BLOCK + 32, smem));
  if (resident < 1) return set_error(h, B200SP_INVALID_INPUT, "csr ring: configuration does not fit on an SM");
  i64 grid = (i64)h->num_sms * (ctas_per_sm < resident ? ctas_per_sm : resident);
  if (grid > num_tiles) grid = num_tiles;
  if (a.dotv && grid > RED_MAX_PARTIALS)
    return set_error(h, B200SP_INVALID_INPUT, "csr: too many CTAs for fused dot");
  kern<<<(unsigned)grid, BLOCK + 32, smem, st>>>(a, (int)R, stages, num_tiles);
  B200SP_LAUNCH_CHECK(h, "csr_ring_kernel");
  return B200SP_OK;
}

template <typename T>
static b200sp_status dispatch_ring(b200sp_handle h, cudaStream_t st, const CsrArgs<T> &a, int block, int npt,
                                   int stages, int cps) {
  if (stages < 2 || stages > 8 || cps < 1 || cps > 16)
    return set_error(h, B200SP_INVALID_INPUT, "csr ring: unsupported stages=%d ctas_per_sm=%d", stages, cps);
#define CASE(B, N) \
  if (block == B && npt == N) return launch_ring<T, B, N>(h, st, a, stages, cps);
  CASE(128, 4) CASE(128, 8) CASE(128, 16)
  CASE(256, 4) CASE(256, 8)
#undef CASE
  return set_error(h, B200SP_INVALID_INPUT, "csr ring: unsupported block_size=%d unroll=%d", block, npt);
}

template <typename T, int BLOCK, int NPT>
static b200sp_status launch_stream(b200sp_handle h, cudaStream_t st, CsrArgs<T> a) {
  constexpr int CAP = BLOCK * NPT;
  static_assert(CAP / CSR_LONG + 2 <= CSR_MAXQ, "long-row queue too small");
  // rows per CTA: the block's nnz should fill about one chunk
  const double mean = (double)a.nnz / (double)a.rows;
  i64 R = (i64)((double)CAP / (mean > 1.0 ? mean : 1.0));
  R = (R / 32) * 32;
  if (R < 32) R = 32;
  if (R > 2048) R = 2048;
  const i64 grid = ceil_div(a.rows, R);
  const T *late_dotv = nullptr;  // see launch_vec
  if (a.dotv && grid > RED_MAX_PARTIALS) {
    late_dotv = a.dotv;
    a.dotv = nullptr;
  }
  const size_t smem = (size_t)(CAP + 16 + R) * sizeof(T) + (size_t)(CAP + 16 + R + 1) * sizeof(int);
  auto kern = csr_stream_kernel<T, BLOCK, NPT>;
  if (smem > 48 * 1024)
    B200SP_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tma_aligned = aligned16(a.Aj) && aligned16(a.Ax);
  kern<<<(unsigned)grid, BLOCK, smem, st>>>(a, (int)R, tma_aligned);
  B200SP_LAUNCH_CHECK(h, "csr_stream_kernel");
  if (late_dotv) return reduce<T, 0>(h, st, a.rows, a.y, late_dotv, a.dot_result, nullptr);
  return B200SP_OK;
}

template <typename T>
static b200sp_status dispatch_stream(b200sp_handle h, cudaStream_t st, const CsrArgs<T> &a, int block, int npt) {
#define CASE(B, N) \
  if (block == B && npt == N) return launch_stream<T, B, N>(h, st, a);
  CASE(128, 4) CASE(128, 8) CASE(128, 16)
  CASE(256, 4) CASE(256, 8) CASE(256, 16)
  CASE(512, 4) CASE(512, 8)
#undef CASE
  return set_error(h, B200SP_INVALID_INPUT, "csr stream: unsupported block_size=%d unroll=%d", block, npt);
}

template <typename T, int BLOCK, int TPR, int RPT>
static b200sp_status launch_vec(b200sp_handle h, cudaStream_t st, CsrArgs<T> a) {
  const i64 grid = ceil_div(a.rows, (i64)(BLOCK / TPR) * RPT);
  // the fused <y, dotv> epilogue keeps one partial per CTA: beyond the workspace the dot runs as its own
  // deterministic reduction after the product (as the COO / HYB / balanced paths always do)
  const T *late_dotv = nullptr;
  if (a.dotv && grid > RED_MAX_PARTIALS) {
    late_dotv = a.dotv;
    a.dotv = nullptr;
  }
  csr_vector_kernel<T, BLOCK, TPR, RPT><<<(unsigned)grid, BLOCK, 0, st>>>(a);
  B200SP_LAUNCH_CHECK(h, "csr_vector_kernel");
  if (late_dotv) return reduce<T, 0>(h, st, a.rows, a.y, late_dotv, a.dot_result, nullptr);
  return B200SP_OK;
}

template <typename T, int BLOCK>
static b200sp_status dispatch_tpr(b200sp_handle h, cudaStream_t st, const CsrArgs<T> &a, int tpr, int rpt) {
#define CASE(P, R) \
  if (tpr == P && rpt == R) return launch_vec<T, BLOCK, P, R>(h, st, a);
  CASE(1, 1) CASE(1, 2) CASE(1, 4)
  CASE(2, 1) CASE(2, 2) CASE(2, 4)
  CASE(4, 1) CASE(4, 2) CASE(4, 4)
  CASE(8, 1) CASE(8, 2) CASE(8, 4)
  CASE(16, 1) CASE(16, 2) CASE(16, 4)
  CASE(32, 1) CASE(32, 2) CASE(32, 4)
#undef CASE
  return set_error(h, B200SP_INVALID_INPUT, "csr: unsupported threads_per_row=%d unroll=%d", tpr, rpt);
}

// structure analysis, one pass over row_offsets (+ a 1/16 sample of column_indices), cached in
// the handle: out[0] = longest row, out[1] = sampled row pairs, out[2] = pairs whose first
// columns are within 32 of each other ("adjacent rows read adjacent x": stencils, banded
// matrices — the case where thread-per-row gathers coalesce)
__global__ void csr_analyze_kernel(i64 rows, const int *Ap, const int *Aj, int *out) {
  int m = 0, pairs = 0, close = 0;
  for (i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (i64)gridDim.x * blockDim.x) {
    const int lo = Ap[r], hi = Ap[r + 1];
    m = max(m, hi - lo);
    if ((r & 15) == 0 && r + 1 < rows && hi > lo) {
      const int hi2 = Ap[r + 2];
      if (hi2 > hi) {
        ++pairs;
        const int d = Aj[hi] - Aj[lo];
        close += (d >= -32 && d <= 32) ? 1 : 0;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m = max(m, __shfl_down_sync(0xffffffffu, m, o));
    pairs += __shfl_down_sync(0xffffffffu, pairs, o);
    close += __shfl_down_sync(0xffffffffu, close, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(out, m);
    if (pairs) atomicAdd(out + 1, pairs);
    if (close) atomicAdd(out + 2, close);
  }
}

struct CsrStructure {
  int longest_row;  // -1: unknown
  bool banded;      // most adjacent rows start at adjacent columns
};

static CsrStructure csr_structure(b200sp_handle h, cudaStream_t st, i64 rows, i64 nnz, const int *Ap, const int *Aj) {
  const b200sp_context::CsrKey key{Ap, rows, nnz};
  auto it = h->csr_max_row.find(key);
  if (it != h->csr_max_row.end()) return CsrStructure{it->second >> 1, (it->second & 1) != 0};
  int *d = reinterpret_cast<int *>(h->dev_scalars + 58);  // 3 ints
  int *p = reinterpret_cast<int *>(h->pinned_scalars + 58);
  CsrStructure r{-1, true};
  if (cudaMemsetAsync(d, 0, 3 * sizeof(int), st) == cudaSuccess) {
    i64 g = ceil_div(rows, 256);
    if (g > (i64)h->num_sms * 16) g = (i64)h->num_sms * 16;
    csr_analyze_kernel<<<(unsigned)g, 256, 0, st>>>(rows, Ap, Aj, d);
    h->launches++;
    if (cudaMemcpyAsync(p, d, 3 * sizeof(int), cudaMemcpyDeviceToHost, st) == cudaSuccess &&
        cudaStreamSynchronize(st) == cudaSuccess) {
      r.longest_row = p[0];
      r.banded = p[1] == 0 || 2 * (i64)p[2] >= (i64)p[1];
    }
  }
  cudaGetLastError();
  if (h->csr_max_row.size() > 256) h->csr_max_row.clear();
  if (r.longest_row >= 0) h->csr_max_row[key] = (r.longest_row << 1) | (r.banded ? 1 : 0);
  return r;
}

// structure class of a CSR matrix for the tuning-cache key (api.cu): 1 banded, 2 scattered, 3 skewed row lengths
int csr_structure_class(b200sp_handle h, cudaStream_t st, i64 rows, i64 nnz, const int *Ap, const int *Aj) {
  if (rows <= 0 || nnz <= 0 || !Ap || !Aj) return 0;
  const CsrStructure cs = csr_structure(h, st, rows, nnz, Ap, Aj);
  if (cs.longest_row < 0) return 0;
  const double mean = (double)nnz / (double)rows;
  if (cs.longest_row > 2048 && (double)cs.longest_row > 64.0 * (mean > 1.0 ? mean : 1.0)) return 3;
  return cs.banded ? 1 : 2;
}

static void csr_defaults(b200sp_cfg &c, i64 rows, i64 nnz, size_t elem, const CsrStructure &cs) {
  const int longest_row = cs.longest_row;
  const double mean = rows > 0 ? (double)nnz / (double)rows : 0.0;
  // skewed row lengths (power-law graphs): one hub row would serialise a row-split kernel
  if (c.kernel == 0 && longest_row > 2048 && (double)longest_row > 64.0 * (mean > 1.0 ? mean : 1.0))
    c.kernel = B200SP_K_CSR_BALANCED;
  if (c.kernel == 0) {
    // round-1 sweeps on B200 (profiles/r01_sweep_*.md): short rows -> thread-per-row from a
    // TMA-staged ring (RING; STREAM when the matrix is too small to fill persistent CTAs);
    // longer rows -> lane-per-entry sub-warps (VECTOR).  cusp::ktt::tune refines per matrix.
    if (mean <= 12.0 && cs.banded)
      c.kernel = (rows >= (i64)B200SP_NUM_SMS_FALLBACK * 4 * 256 * 2) ? B200SP_K_CSR_RING : B200SP_K_CSR_STREAM;
    else if (!cs.banded && mean >= 192.0 && c.block_size == 0 && c.unroll == 0) {
      // long rows, scattered columns (BASELINE configs[3] at 256 nnz/row): every row piece is longer than CSR_LONG, so
      // the stream kernel gives each to a warp that reads its entries from the bulk-copied chunk — 1.005 - 1.012 ms
      // against 1.087 - 1.094 for the sub-warp kernel on 2^20 x 256 fp32 (profiles/r01_sweep_random.json, the tuned
      // best of every bench line), which is the rate of a bare gather of that column stream (r04_gather_probe.md)
      c.kernel = B200SP_K_CSR_STREAM;
      c.block_size = (elem == 4) ? 128 : 256;
      c.unroll = 8;
    } else
      c.kernel = B200SP_K_CSR_VECTOR;  // scattered columns: lane-per-entry keeps the matrix stream coalesced
  }
  if (c.kernel == B200SP_K_CSR_RING) {
    if (c.block_size == 0) c.block_size = 256;
    if (c.unroll == 0) c.unroll = 8;
    if (c.stages == 0) c.stages = 2;
    if (c.ctas_per_sm == 0) c.ctas_per_sm = 4;
    return;
  }
  if (c.kernel == B200SP_K_CSR_BALANCED) {
    if (c.block_size == 0) c.block_size = 256;
    if (c.unroll == 0) c.unroll = 7;
    return;
  }
  if (c.kernel == B200SP_K_CSR_STREAM) {
    if (c.block_size == 0) c.block_size = 128;
    if (c.unroll == 0) c.unroll = (elem == 4) ? 16 : 8;
    return;
  }
  if (c.threads_per_row == 0) {
    // smallest power of two >= mean row length up to 32 nnz/row (the reference uses the
    // integer mean nnz/rows, csr_vector_spmv.h:236-257, which under-sizes e.g. 7 -> 4);
    // beyond that 16 lanes with deeper per-lane loops won the random-matrix sweep
    int t = 2;
    while (t < 32 && (double)t < mean) t *= 2;
    if (mean > 32.0) t = 16;
    c.threads_per_row = t;
  }
  if (c.block_size == 0) c.block_size = (mean > 32.0) ? 512 : ((elem == 4) ? 512 : 128);
  if (c.unroll == 0) c.unroll = (mean > 32.0) ? 1 : 4;
}

template <typename T>
b200sp_status spmv_csr(b200sp_handle h, cudaStream_t st, i64 rows, i64 cols, i64 nnz, const int *Ap,
                       const int *Aj, const T *Ax, const T *x, T *y, int accumulate,
                       const b200sp_cfg *cfg, const T *dotv, T *dot_result) {
  B200SP_CHECK_HANDLE(h);
  B200SP_REQUIRE(h, rows >= 0 && cols >= 0 && nnz >= 0, "csr: negative dimension");
  B200SP_REQUIRE(h, rows < (1ll << 31) && cols < (1ll << 31) && nnz < (1ll << 31), "csr: int32 index range");
  if (rows == 0) {
    if (dot_result) B200SP_CUDA(h, cudaMemsetAsync(dot_result, 0, sizeof(T), st));
    return B200SP_OK;
  }
  B200SP_REQUIRE(h, y != nullptr && Ap != nullptr, "csr: null pointer");
  if (nnz == 0) {  // every row is empty: y = init(y)
    if (!accumulate) B200SP_CUDA(h, cudaMemsetAsync(y, 0, (size_t)rows * sizeof(T), st));
    if (dot_result) B200SP_CUDA(h, cudaMemsetAsync(dot_result, 0, sizeof(T), st));
    return B200SP_OK;
  }
  B200SP_REQUIRE(h, Aj && Ax && x, "csr: null pointer");
  B200SP_REQUIRE(h, cols > 0, "csr: num_cols == 0 with stored entries");

  b200sp_cfg c = cfg ? *cfg : b200sp_cfg{};
  const CsrStructure cs = (c.kernel == 0) ? csr_structure(h, st, rows, nnz, Ap, Aj) : CsrStructure{-1, true};
  csr_defaults(c, rows, nnz, sizeof(T), cs);
  if (c.kernel == B200SP_K_CSR_BALANCED) {
    b200sp_status bs = spmv_csr_balanced<T>(h, st, rows, cols, nnz, Ap, Aj, Ax, x, y, accumulate, c.block_size, c.unroll);
    if (bs != B200SP_OK || !dotv) return bs;
    return reduce<T, 0>(h, st, rows, y, dotv, dot_result, nullptr);  // no fused epilogue: separate deterministic dot
  }
  if (c.kernel != B200SP_K_CSR_VECTOR && c.kernel != B200SP_K_CSR_STREAM && c.kernel != B200SP_K_CSR_RING)
    return set_error(h, B200SP_INVALID_INPUT, "csr: unknown kernel id %d", c.kernel);
  if (c.kernel == B200SP_K_CSR_RING && !(aligned16(Aj) && aligned16(Ax))) {
    // bulk copies need 16-byte aligned array bases: same arithmetic through the stream kernel's LDG path
    c.kernel = B200SP_K_CSR_STREAM;
    c.stages = c.ctas_per_sm = 0;
    if (c.block_size == 256 && c.unroll > 16) c.unroll = 16;
  }

  CsrArgs<T> a;
  a.rows = rows; a.cols = cols; a.nnz = nnz; a.Ap = Ap; a.Aj = Aj; a.Ax = Ax; a.x = x; a.y = y;
  a.accumulate = accumulate; a.dotv = dotv; a.dot_result = dot_result;
  a.dot_partials = reinterpret_cast<T *>(h->red_partials);
  a.dot_ticket = h->red_counters;

  if (c.kernel == B200SP_K_CSR_RING)
    return dispatch_ring<T>(h, st, a, c.block_size, c.unroll, c.stages, c.ctas_per_sm);
  if (c.kernel == B200SP_K_CSR_STREAM) return dispatch_stream<T>(h, st, a, c.block_size, c.unroll);
  switch (c.block_size) {
    case 128: return dispatch_tpr<T, 128>(h, st, a, c.threads_per_row, c.unroll);
    case 256: return dispatch_tpr<T, 256>(h, st, a, c.threads_per_row, c.unroll);
    case 512: return dispatch_tpr<T, 512>(h, st, a, c.threads_per_row, c.unroll);
  }
  return set_error(h, B200SP_INVALID_INPUT, "csr: unsupported block_size=%d", c.block_size);
}

template b200sp_status spmv_csr<float>(b200sp_handle, cudaStream_t, i64, i64, i64, const int *, const int *,
                                       const float *, const float *, float *, int, const b200sp_cfg *,
                                       const float *, float *);
template b200sp_status spmv_csr<double>(b200sp_handle, cudaStream_t, i64, i64, i64, const int *, const int *,
                                        const double *, const double *, double *, int, const b200sp_cfg *,
                                        const double *, double *);

}  // namespace b200sp

extern "C" {
#define DEF(T, sfx)                                                                             \
  b200sp_status b200sp_spmv_csr_##sfx(b200sp_handle h, b200sp_stream stream, int64_t num_rows,  \
                                      int64_t num_cols, int64_t num_entries,                    \
                                      const int32_t *row_offsets, const int32_t *column_indices, \
                                      const T *values, const T *x, T *y, int accumulate,        \
                                      const b200sp_cfg *cfg) {                                  \
    return b200sp::spmv_csr<T>(h, (cudaStream_t)stream, num_rows, num_cols, num_entries,        \
                               row_offsets, column_indices, values, x, y, accumulate, cfg,      \
                               nullptr, nullptr);                                               \
  }
DEF(float, f32)
DEF(double, f64)
#undef DEF
}
