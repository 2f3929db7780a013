// spmv_coo.cu — COO and HYB SpMV for sm_100a.
//
// Replaces, for coo_matrix on device_memory:
//   - the thrust::reduce_by_key path that actually runs today
//     (cusp/system/detail/generic/multiply/spmv.h:182-238: 3 kernels, 2 temporaries
//     of num_rows, allocations per call),
//   - the dormant spmv_coo_flat_kernel + reduce_update + serial trio
//     (cusp/system/cuda/detail/multiply/coo_flat_spmv.h:225-463, coo_serial.h:35-54),
//   - the KTT coo_spmv variants that all finish with global atomicAdd
//     (cusp/system/cuda/ktt/kernels/coo_kernel.h:25-392).
// and for hyb_matrix the two-pass generic/multiply/spmv.h:272-290.
// Semantics: host loop cusp/system/detail/sequential/multiply/coo_spmv.h:35-68
//     y[i] = init(y[i]);  for n ascending: y[Ai[n]] += Ax[n]*x[Aj[n]]
// with row indices sorted ascending (duplicates allowed), as cusp requires.
//
// K_COO_SEGSCAN — nnz-balanced, deterministic, no atomics:
//   * a CTA owns a tile of BLOCK*VPT consecutive entries (hub rows of any length
//     are split evenly: "merge-path" balance degenerates to an even nnz split
//     because COO stores the row of every entry);
//   * entries are loaded coalesced (ld.global.cs), x gathered (ld.global.nc),
//     products + rows parked in shared memory, then each thread reduces VPT
//     consecutive entries serially (VPT odd -> conflict-free) and a block-wide
//     segmented scan in shared memory/shuffles stitches rows across threads;
//   * rows that lie entirely inside a tile are written straight to y; the (at
//     most two) rows a tile shares with its neighbours go to a per-tile carry
//     record; a second tiny kernel walks each carry chain in tile order.
//   Summation order is fixed by (nnz, BLOCK, VPT) -> bit-reproducible.
//
// Algorithmic bytes: nnz*(8+sizeof(T)) + cols*sizeof(T) + rows*sizeof(T).
#include <stdlib.h>

#include "coo.cuh"

namespace b200sp {

// first row of every nnz tile of a CSR matrix: tile t starts at entry t*TILE, which lies in
// the row r with Ap[r] <= t*TILE < Ap[r+1].  Replaces gpu_compute_row_starts
// (cusp/system/cuda/ktt/csr_multiply.h:64-105, the preprocessing of the KTT kernel's
// DYNAMIC=2 "balanced" mode); like there it runs before every product (Ap may change).
__global__ void csr_tile_rows_kernel(i64 rows, const int *Ap, int tile, int *tile_first_row) {
  const i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int lo = Ap[r], hi = Ap[r + 1];
  for (int t = (lo + tile - 1) / tile; (i64)t * tile < hi; ++t) tile_first_row[t] = (int)r;
}

template <typename T, int BLOCK, int VPT, bool CSR>
__global__ void __launch_bounds__(BLOCK) coo_segscan_kernel(CooArgs<T> a) {
  constexpr int TILE = BLOCK * VPT;
  constexpr int NW = BLOCK / 32;
  __shared__ int s_row[TILE + 1];
  __shared__ T s_val[TILE];
  __shared__ T s_wv[NW];
  __shared__ int s_wf[NW];
  __shared__ int s_head_row;
  __shared__ T s_head_val;

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const i64 start = (i64)blockIdx.x * TILE;
  const int n = (int)min((i64)TILE, a.nnz - start);
  const unsigned cols = (unsigned)a.cols;

  // ---- coalesced load + gather + multiply -------------------------------
  {
    int r[VPT], c[VPT];
    T v[VPT], xv[VPT];
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const i64 g = min(start + i * BLOCK + tid, a.nnz - 1);
      r[i] = CSR ? -1 : ld_stream(a.Ai + g);
      c[i] = ld_stream(a.Aj + g);
      v[i] = ld_stream(a.Ax + g);
    }
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      pin(c[i]);
      xv[i] = ld_ro(a.x + min((unsigned)c[i], cols - 1));
    }
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      pin(xv[i]);
      const int idx = i * BLOCK + tid;
      const bool ok = idx < n;
      s_row[idx] = ok ? r[i] : -1;
      s_val[idx] = ok ? v[i] * xv[i] : T(0);
    }
  }
  int prev_row = -1;
  if (CSR) {
    // rebuild the row of every entry: each row that owns entries of this tile marks its
    // first slot, an inclusive max-scan spreads the marks (rows ascend)
    __shared__ int s_wmax[NW];
    const int r_first = a.tile_first_row[blockIdx.x];
    const int r_last = (start + TILE < a.nnz) ? a.tile_first_row[blockIdx.x + 1] : (int)a.rows - 1;
    if (tid == 0) {
      s_row[TILE] = (start + TILE < a.nnz) ? r_last : -1;
      s_head_row = -1;
      s_head_val = T(0);
    }
    prev_row = (ld_ro(a.Ap + r_first) < start) ? r_first : -2;  // does the first row continue from the previous tile?
    __syncthreads();  // slots hold -1 (or stale -1 beyond n) from the load phase
    for (int rr_ = r_first + tid; rr_ <= r_last; rr_ += BLOCK) {
      const i64 lo = ld_ro(a.Ap + rr_), hi = ld_ro(a.Ap + rr_ + 1);
      if (hi > lo && lo < start + n && hi > start) s_row[(int)(max(lo, start) - start)] = rr_;
    }
    __syncthreads();
    int loc[VPT], m = -1;
#pragma unroll
    for (int q = 0; q < VPT; ++q) {
      m = max(m, s_row[tid * VPT + q]);
      loc[q] = m;
    }
    int incl = m;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl = max(incl, u);
    }
    if (lane == 31) s_wmax[w] = incl;
    __syncthreads();
    int excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = -1;
    for (int k = 0; k < w; ++k) excl = max(excl, s_wmax[k]);
#pragma unroll
    for (int q = 0; q < VPT; ++q) {
      const int idx = tid * VPT + q;
      s_row[idx] = (idx < n) ? max(loc[q], excl) : -1;
    }
    __syncthreads();
  } else {
    if (tid == 0) {
      s_row[TILE] = (start + TILE < a.nnz) ? a.Ai[start + TILE] : -1;
      s_head_row = -1;
      s_head_val = T(0);
    }
    if (start > 0) prev_row = ld_ro(a.Ai + start - 1);
    __syncthreads();
  }

  // ---- per-thread serial segmented reduction over VPT consecutive entries ----
  int rr[VPT + 1];
  T pv[VPT];
#pragma unroll
  for (int q = 0; q < VPT; ++q) {
    rr[q] = s_row[tid * VPT + q];
    pv[q] = s_val[tid * VPT + q];
  }
  rr[VPT] = s_row[tid * VPT + VPT];

  T run = T(0), head = T(0);
  int head_row = -1;
  bool has_b = false;
#pragma unroll
  for (int q = 0; q < VPT; ++q) {
    run = run + pv[q];
    if (rr[q] != rr[q + 1]) {  // row rr[q] ends at this entry
      if (!has_b) {
        head = run;
        head_row = rr[q];
        has_b = true;
      } else if (rr[q] >= 0) {
        // began and ended inside this thread: sole owner of y[row]
        a.y[rr[q]] = a.accumulate ? a.y[rr[q]] + run : run;
      }
      run = T(0);
    }
  }

  // ---- block-wide segmented scan of (has_b, tail) ---------------------------
  T vi = run;
  int fi = has_b ? 1 : 0;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const T vu = __shfl_up_sync(0xffffffffu, vi, d);
    const int fu = __shfl_up_sync(0xffffffffu, fi, d);
    if (lane >= d) {
      if (!fi) vi = vu + vi;
      fi |= fu;
    }
  }
  if (lane == 31) {
    s_wv[w] = vi;
    s_wf[w] = fi;
  }
  __syncthreads();
  // exclusive prefix over warps (serial over <= 16 warps: fixed order)
  T VW = T(0);
  for (int k = 0; k < w; ++k) VW = s_wf[k] ? s_wv[k] : VW + s_wv[k];
  const T Vi = fi ? vi : VW + vi;  // block-inclusive tail sum
  T carry_in = __shfl_up_sync(0xffffffffu, Vi, 1);
  if (lane == 0) carry_in = VW;

  if (has_b && head_row >= 0) {
    const T total = carry_in + head;
    if (head_row == prev_row) {  // row continued from the previous tile
      s_head_row = head_row;
      s_head_val = total;
    } else {
      a.y[head_row] = a.accumulate ? a.y[head_row] + total : total;
    }
  }
  __syncthreads();
  if (tid == BLOCK - 1) {
    CooCarry<T> cr;
    cr.head_row = s_head_row;
    cr.head_val = s_head_val;
    cr.pad = 0;
    const int last_row = rr[VPT - 1];
    if (last_row >= 0 && last_row == rr[VPT]) {
      cr.tail_row = last_row;
      cr.tail_val = Vi;
      cr.leader = (last_row != prev_row) ? 1 : 0;
    } else {
      cr.tail_row = -1;
      cr.tail_val = T(0);
      cr.leader = 0;
    }
    a.carry[blockIdx.x] = cr;
  }
}

template <typename T>
b200sp_status launch_coo_fixup(b200sp_handle h, cudaStream_t st, i64 tiles, const CooCarry<T> *carry, T *y,
                               int accumulate) {
  coo_fixup_kernel<T, SpmvOps<T, 0, 0>><<<(unsigned)ceil_div(tiles, 256), 256, 0, st>>>(tiles, carry, y, accumulate);
  B200SP_LAUNCH_CHECK(h, "coo_fixup_kernel");
  return B200SP_OK;
}
template b200sp_status launch_coo_fixup<float>(b200sp_handle, cudaStream_t, i64, const CooCarry<float> *, float *, int);
template b200sp_status launch_coo_fixup<double>(b200sp_handle, cudaStream_t, i64, const CooCarry<double> *, double *,
                                                int);

// ---------------------------------------------------------------------------
// K_COO_RING — the same tiles and the same summation order as K_COO_SEGSCAN, but the three
// entry streams never touch the LSU/L1TEX path: persistent CTAs, one producer lane stages
// each tile's row / column / value ranges with three bulk copies (TMA engine, L2
// evict-first) into an mbarrier ring; the consumers read their VPT consecutive entries
// straight from shared memory (VPT odd -> conflict-free, no transposition buffer), gather
// x, run the serial + shuffle segmented scan and release the stage.  L1TEX then carries only
// the x gathers, which is what bounds COO on coalesced inputs (ncu: l1tex 80 % with the
// LDG kernel).  One named barrier per tile (the cross-warp carry arrays are double
// buffered by tile parity).  Needs 16-byte aligned array bases; spmv_coo falls back to
// K_COO_SEGSCAN otherwise.  Bit-identical to K_COO_SEGSCAN for equal (BLOCK, VPT).
// ---------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void consumer_bar() {
  asm volatile("bar.sync 1, %0;" ::"n"(N) : "memory");
}

template <typename T, int BLOCK, int VPT>
__global__ void __launch_bounds__(BLOCK + 32) coo_ring_kernel(CooArgs<T> a, int stages, i64 num_tiles) {
  constexpr int TILE = BLOCK * VPT;
  constexpr int NW = BLOCK / 32;
  constexpr int RSTR = TILE + 8;  // rows of entries [start-4, start+TILE+4): previous and next row ride along
  constexpr int EPV = 16 / (int)sizeof(T);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T *s_val = reinterpret_cast<T *>(smem_raw);
  int *s_col = reinterpret_cast<int *>(smem_raw + (size_t)stages * TILE * sizeof(T));
  int *s_row = s_col + (size_t)stages * TILE;
  uint64_t *full = reinterpret_cast<uint64_t *>(s_row + (size_t)stages * RSTR);
  uint64_t *empty = full + stages;
  __shared__ T s_wv[2][NW];
  __shared__ int s_wf[2][NW];

  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NW);
    }
    mbar_fence_init();
  }
  for (int i = tid; i < stages * TILE; i += BLOCK + 32) s_col[i] = 0;
  __syncthreads();

  const i64 nnz = a.nnz;
  if (tid >= BLOCK) {
    // ------------------------------ producer --------------------------------
    if (tid == BLOCK) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      const uint64_t pol = l2_policy_evict_first();
      int s = 0;
      uint32_t ph = 0;
      for (i64 tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const i64 start = tile * TILE;
        const int n = (int)min((i64)TILE, nnz - start);
        mbar_wait(&empty[s], ph ^ 1);
        T *dv = s_val + (size_t)s * TILE;
        int *dc = s_col + (size_t)s * TILE;
        int *dr = s_row + (size_t)s * RSTR + (start > 0 ? 0 : 4);
        const i64 g0 = start > 0 ? start - 4 : 0;
        const int nr = (int)(min(start + TILE + 4, nnz) - g0);
        const int bv = n & ~(EPV - 1), bc = n & ~3, br = nr & ~3;
        // the (at most 3) entries after the last complete 16 bytes of the arrays
        for (int j = bv; j < n; ++j) dv[j] = a.Ax[start + j];
        for (int j = bc; j < n; ++j) dc[j] = a.Aj[start + j];
        for (int j = br; j < nr; ++j) dr[j] = a.Ai[g0 + j];
        mbar_expect_tx(&full[s], (uint32_t)(bv * sizeof(T) + (bc + br) * sizeof(int)));
        if (br > 0) bulk_g2s(dr, a.Ai + g0, (uint32_t)(br * sizeof(int)), &full[s], pol);
        if (bc > 0) bulk_g2s(dc, a.Aj + start, (uint32_t)(bc * sizeof(int)), &full[s], pol);
        if (bv > 0) bulk_g2s(dv, a.Ax + start, (uint32_t)(bv * sizeof(T)), &full[s], pol);
        if (++s == stages) {
          s = 0;
          ph ^= 1;
        }
      }
    }
    return;
  }
  // ------------------------------ consumers -------------------------------
  const int lane = tid & 31, w = tid >> 5;
  const unsigned cols = (unsigned)a.cols;
  int s = 0, par = 0;
  uint32_t ph = 0;
  for (i64 tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, par ^= 1) {
    const i64 start = tile * TILE;
    const int n = (int)min((i64)TILE, nnz - start);
    const int n_rows = (start + TILE < nnz) ? TILE + 1 : n;  // entries whose row index was staged
    mbar_wait(&full[s], ph);
    const int *pr = s_row + (size_t)s * RSTR + 4;  // pr[j]: row of entry start + j
    const int *pc = s_col + (size_t)s * TILE;
    const T *pv_ = s_val + (size_t)s * TILE;
    int rr[VPT + 1], c[VPT];
    T v[VPT], xv[VPT];
    const int prev_row = (start > 0) ? pr[-1] : -1;
#pragma unroll
    for (int q = 0; q < VPT; ++q) {
      const int idx = tid * VPT + q;
      rr[q] = (idx < n) ? pr[idx] : -1;
      c[q] = pc[idx];
      v[q] = pv_[idx];
    }
    rr[VPT] = (tid * VPT + VPT < n_rows) ? pr[tid * VPT + VPT] : -1;
    const int s_used = s;  // released after the products are formed (see below)
    if (++s == stages) {
      s = 0;
      ph ^= 1;
    }
#pragma unroll
    for (int q = 0; q < VPT; ++q) xv[q] = ld_ro(a.x + min((unsigned)c[q], cols - 1));

    // ---- per-thread serial segmented reduction over VPT consecutive entries ----
    T run = T(0), head = T(0);
    int head_row = -1;
    bool has_b = false;
#pragma unroll
    for (int q = 0; q < VPT; ++q) {
      pin(xv[q]);
      const T p = (tid * VPT + q < n) ? v[q] * xv[q] : T(0);
      run = run + p;
      if (rr[q] != rr[q + 1]) {  // row rr[q] ends at this entry
        if (!has_b) {
          head = run;
          head_row = rr[q];
          has_b = true;
        } else if (rr[q] >= 0) {
          a.y[rr[q]] = a.accumulate ? a.y[rr[q]] + run : run;  // began and ended inside this thread
        }
        run = T(0);
      }
    }

    // ---- block-wide segmented scan of (has_b, tail) ---------------------------
    T vi = run;
    int fi = has_b ? 1 : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const T vu = __shfl_up_sync(0xffffffffu, vi, d);
      const int fu = __shfl_up_sync(0xffffffffu, fi, d);
      if (lane >= d) {
        if (!fi) vi = vu + vi;
        fi |= fu;
      }
    }
    if (lane == 31) {
      s_wv[par][w] = vi;
      s_wf[par][w] = fi;
    }
    consumer_bar<BLOCK>();
    T VW = T(0);
    int FW = 0;  // a row ended in an earlier warp of this tile
    for (int k = 0; k < w; ++k) {
      VW = s_wf[par][k] ? s_wv[par][k] : VW + s_wv[par][k];
      FW |= s_wf[par][k];
    }
    const T Vi = fi ? vi : VW + vi;  // block-inclusive tail sum
    T carry_in = __shfl_up_sync(0xffffffffu, Vi, 1);
    int f_before = __shfl_up_sync(0xffffffffu, fi, 1);
    if (lane == 0) {
      carry_in = VW;
      f_before = 0;
    }
    f_before |= FW;

    CooCarry<T> *cr = a.carry + tile;
    if (has_b && head_row >= 0) {
      const T total = carry_in + head;
      const bool continued = (head_row == prev_row);  // only the first row end of a tile can be
      if (continued) {
        cr->head_row = head_row;
        cr->head_val = total;
      } else {
        a.y[head_row] = a.accumulate ? a.y[head_row] + total : total;
        if (!f_before) {
          cr->head_row = -1;
          cr->head_val = T(0);
        }
      }
    }
    if (tid == BLOCK - 1) {
      if (!(fi | FW)) {  // no row ends inside this tile
        cr->head_row = -1;
        cr->head_val = T(0);
      }
      cr->pad = 0;
      const int last_row = rr[VPT - 1];
      if (last_row >= 0 && last_row == rr[VPT]) {
        cr->tail_row = last_row;
        cr->tail_val = Vi;
        cr->leader = (last_row != prev_row) ? 1 : 0;
      } else {
        cr->tail_row = -1;
        cr->tail_val = T(0);
        cr->leader = 0;
      }
    }
    // Release the stage only at the end of the tile, after every staged value has been consumed
    // (products -> shuffles / y stores, rows -> branches): see consume_before_release() in
    // common.cuh for what happens when the arrive is issued right after the loads.
    consume_before_release(Vi);
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s_used]);
  }
}

template <typename T, int BLOCK, int VPT>
static b200sp_status launch_coo_ring(b200sp_handle h, cudaStream_t st, CooArgs<T> a, int stages, int ctas_per_sm) {
  constexpr int TILE = BLOCK * VPT;
  const i64 tiles = ceil_div(a.nnz, (i64)TILE);
  b200sp_status s = ensure_scratch(h, (size_t)tiles * sizeof(CooCarry<T>));
  if (s != B200SP_OK) return s;
  a.carry = reinterpret_cast<CooCarry<T> *>(h->scratch);
  const size_t smem = (size_t)stages * ((size_t)TILE * (sizeof(T) + sizeof(int)) + (size_t)(TILE + 8) * sizeof(int)) +
                      2 * (size_t)stages * sizeof(uint64_t);
  if (smem > (size_t)h->max_smem_optin)
    return set_error(h, B200SP_INVALID_INPUT, "coo ring: %zu B smem exceeds %d", smem, h->max_smem_optin);
  auto kern = coo_ring_kernel<T, BLOCK, VPT>;
  B200SP_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int resident = 0;
  B200SP_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, BLOCK + 32, smem));
  if (resident < 1) return set_error(h, B200SP_INVALID_INPUT, "coo ring: configuration does not fit on an SM");
  i64 grid = (i64)h->num_sms * (ctas_per_sm < resident ? ctas_per_sm : resident);
  if (grid > tiles) grid = tiles;
  kern<<<(unsigned)grid, BLOCK + 32, smem, st>>>(a, stages, tiles);
  B200SP_LAUNCH_CHECK(h, "coo_ring_kernel");
  return launch_coo_fixup<T>(h, st, tiles, a.carry, a.y, a.accumulate);
}

template <typename T, int BLOCK, int VPT>
static b200sp_status launch_coo(b200sp_handle h, cudaStream_t st, CooArgs<T> a) {
  constexpr int TILE = BLOCK * VPT;
  const i64 tiles = ceil_div(a.nnz, (i64)TILE);
  b200sp_status s = ensure_scratch(h, (size_t)tiles * sizeof(CooCarry<T>));
  if (s != B200SP_OK) return s;
  a.carry = reinterpret_cast<CooCarry<T> *>(h->scratch);
  coo_segscan_kernel<T, BLOCK, VPT, false><<<(unsigned)tiles, BLOCK, 0, st>>>(a);
  B200SP_LAUNCH_CHECK(h, "coo_segscan_kernel");
  return launch_coo_fixup<T>(h, st, tiles, a.carry, a.y, a.accumulate);
}

// CSR through the nnz-balanced segmented scan (K_CSR_BALANCED)
template <typename T, int BLOCK, int VPT>
static b200sp_status launch_csr_balanced(b200sp_handle h, cudaStream_t st, CooArgs<T> a) {
  constexpr int TILE = BLOCK * VPT;
  const i64 tiles = ceil_div(a.nnz, (i64)TILE);
  const size_t carry_bytes = ((size_t)tiles * sizeof(CooCarry<T>) + 255) & ~(size_t)255;
  b200sp_status s = ensure_scratch(h, carry_bytes + (size_t)(tiles + 1) * sizeof(int));
  if (s != B200SP_OK) return s;
  a.carry = reinterpret_cast<CooCarry<T> *>(h->scratch);
  int *tfr = reinterpret_cast<int *>(reinterpret_cast<char *>(h->scratch) + carry_bytes);
  a.tile_first_row = tfr;
  csr_tile_rows_kernel<<<(unsigned)ceil_div(a.rows, 256), 256, 0, st>>>(a.rows, a.Ap, TILE, tfr);
  B200SP_LAUNCH_CHECK(h, "csr_tile_rows_kernel");
  coo_segscan_kernel<T, BLOCK, VPT, true><<<(unsigned)tiles, BLOCK, 0, st>>>(a);
  B200SP_LAUNCH_CHECK(h, "coo_segscan_kernel<csr>");
  return launch_coo_fixup<T>(h, st, tiles, a.carry, a.y, a.accumulate);
}

// entry used by spmv_csr (spmv_csr.cu) for cfg.kernel == B200SP_K_CSR_BALANCED
template <typename T>
b200sp_status spmv_csr_balanced(b200sp_handle h, cudaStream_t st, i64 rows, i64 cols, i64 nnz, const int *Ap,
                                const int *Aj, const T *Ax, const T *x, T *y, int accumulate, int block, int vpt) {
  if (!accumulate) B200SP_CUDA(h, cudaMemsetAsync(y, 0, (size_t)rows * sizeof(T), st));
  CooArgs<T> a;
  a.rows = rows; a.cols = cols; a.nnz = nnz; a.Ai = nullptr; a.Aj = Aj; a.Ax = Ax; a.x = x; a.y = y;
  a.accumulate = accumulate;
  a.carry = nullptr;
  a.Ap = Ap;
  a.tile_first_row = nullptr;
  a.scalar_loads = 0;
#define CASE(B, V) \
  if (block == B && vpt == V) return launch_csr_balanced<T, B, V>(h, st, a);
  CASE(128, 7) CASE(256, 5) CASE(256, 7) CASE(256, 9)
#undef CASE
  return set_error(h, B200SP_INVALID_INPUT, "csr balanced: unsupported block_size=%d unroll=%d", block, vpt);
}
template b200sp_status spmv_csr_balanced<float>(b200sp_handle, cudaStream_t, i64, i64, i64, const int *, const int *,
                                                const float *, const float *, float *, int, int, int);
template b200sp_status spmv_csr_balanced<double>(b200sp_handle, cudaStream_t, i64, i64, i64, const int *, const int *,
                                                 const double *, const double *, double *, int, int, int);

// Gather-locality probe for the default kernel choice.  The two COO kernels issue their x
// gathers in different lane orders: K_COO_SEGSCAN lane l of a warp reads entry 32k + l
// (consecutive entries), K_COO_RING lane l reads entry 7l + q (its own 7 consecutive entries).
// What a gather costs in L1TEX is the number of distinct 128-byte lines per instruction, so
// the probe counts exactly that for both orders on `samples` windows of 224 entries spread
// over the matrix (one warp per window, __match_any_sync on the line id).
// out[0] = lines in consecutive-entry order, out[1] = lines in 7-strided order.
__global__ void coo_analyze_kernel(i64 nnz, const int *Aj, int line_shift, int samples, int *out) {
  const int lane = threadIdx.x & 31;
  const int wid = (int)(((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (wid >= samples) return;
  const i64 windows = nnz / 224;
  const i64 base = (windows * wid / samples) * 224;
  int la = 0, lb = 0;
  for (int q = 0; q < 7; ++q) {
    const int ca = Aj[base + q * 32 + lane] >> line_shift;
    const int cb = Aj[base + lane * 7 + q] >> line_shift;
    const unsigned ma = __match_any_sync(0xffffffffu, ca), mb = __match_any_sync(0xffffffffu, cb);
    la += (__ffs(ma) - 1 == lane) ? 1 : 0;
    lb += (__ffs(mb) - 1 == lane) ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    la += __shfl_down_sync(0xffffffffu, la, o);
    lb += __shfl_down_sync(0xffffffffu, lb, o);
  }
  if (lane == 0) {
    atomicAdd(out, la);
    atomicAdd(out + 1, lb);
  }
}

// Which lane order suits this column stream?  1: the ring kernel's strided order coalesces (few lines per
// instruction, and no worse than the consecutive order: stencils, banded operators); 2: only the consecutive order
// coalesces (one entry per row, e.g. the COO tail of a stencil HYB) -> the LDG scan kernel; 0: neither does
// (scattered columns: graphs) -> the warp-autonomous kernel.  Cached per (column_indices, nnz, element size); a hint
// only — every kernel is correct on every matrix.
static int coo_gather_class(b200sp_handle h, cudaStream_t st, i64 nnz, const int *Aj, size_t elem) {
  const b200sp_context::CsrKey key{Aj, (int64_t)elem, nnz};
  auto it = h->coo_gather_order.find(key);
  if (it != h->coo_gather_order.end()) return it->second;
  const int samples = 2048;
  int *d = reinterpret_cast<int *>(h->dev_scalars + 54);  // 2 ints
  int *p = reinterpret_cast<int *>(h->pinned_scalars + 54);
  int cls = 2;
  if (cudaMemsetAsync(d, 0, 2 * sizeof(int), st) == cudaSuccess) {
    coo_analyze_kernel<<<samples / 8, 256, 0, st>>>(nnz, Aj, elem == 4 ? 5 : 4, samples, d);
    h->launches++;
    if (cudaMemcpyAsync(p, d, 2 * sizeof(int), cudaMemcpyDeviceToHost, st) == cudaSuccess &&
        cudaStreamSynchronize(st) == cudaSuccess) {
      const double per_instr_seq = (double)p[0] / (7.0 * samples), per_instr_str = (double)p[1] / (7.0 * samples);
      if (per_instr_str <= 4.0 && per_instr_str <= per_instr_seq)
        cls = 1;
      else if (per_instr_seq > 8.0 && per_instr_str > 8.0)
        cls = 0;
    }
  }
  cudaGetLastError();
  if (h->coo_gather_order.size() > 256) h->coo_gather_order.clear();
  h->coo_gather_order[key] = cls;
  return cls;
}

// structure class of a COO column stream for the tuning-cache key (api.cu): 1 ring order coalesces (banded),
// 2 consecutive order coalesces, 3 scattered; 0 for streams too short to probe
int coo_structure_class(b200sp_handle h, cudaStream_t st, i64 nnz, const int *Aj, size_t elem) {
  if (!Aj || nnz < (i64)h->num_sms * 8 * 256 * 7) return 0;
  const int cls = coo_gather_class(h, st, nnz, Aj, elem);
  return cls == 1 ? 1 : (cls == 2 ? 2 : 3);
}

static void coo_defaults(b200sp_cfg &c, b200sp_handle h, cudaStream_t st, i64 nnz, const int *Aj, size_t elem,
                         bool tma_ok, bool vec32_ok) {
  const bool no_shape = c.block_size == 0 && c.unroll == 0;
  if (c.block_size == 0) c.block_size = 256;
  if (c.unroll == 0) c.unroll = 7;
  // The persistent ring needs a few tiles per resident CTA to be worth its prologue, and its
  // gather order must coalesce (stencils, banded operators); scattered or consecutive-column
  // streams (graphs, one entry per row) run the LDG kernel with its 2048 threads per SM.
  if (c.kernel == 0) {
    const bool big = nnz >= (i64)h->num_sms * 8 * c.block_size * c.unroll;
    const int cls = big ? coo_gather_class(h, st, nnz, Aj, elem) : 2;
    if (tma_ok && big && cls == 1) {
      c.kernel = B200SP_K_COO_RING;
    } else if (big && cls == 0 && no_shape && vec32_ok) {
      // scattered columns (graphs): the warp-autonomous kernel, 256-bit loads (R-MAT scale 22 / 24: 0.47 of the copy
      // rate against 0.41 - 0.44 for the shared-memory scan, profiles/r03_coo_probe.md); a persistent grid once every
      // warp has tens of tiles
      c = b200sp_cfg{};
      c.kernel = B200SP_K_COO_WARP;
      c.block_size = 256;
      c.vector_width = 8;
      const bool many = nnz >= ((i64)1 << 27);
      c.unroll = many ? 2 : 1;
      c.ctas_per_sm = many ? 8 : 0;
      return;
    } else if (big && cls == 2 && no_shape && vec32_ok) {
      // consecutive entries gather neighbouring columns (a HYB tail with one entry per row): warp tiles with 128-bit
      // loads and two units per warp — poisson7pt 256^3 forced to K = 6: whole HYB product 0.253 / 0.361 ms (fp32 /
      // fp64) against 0.275 / 0.392 with the shared-memory scan (tools/hyb_probe.py)
      c = b200sp_cfg{};
      c.kernel = B200SP_K_COO_WARP;
      c.block_size = 256;
      c.vector_width = 4;
      c.unroll = 2;
      return;
    } else {
      c.kernel = B200SP_K_COO_SEGSCAN;
    }
  }
  if (c.kernel == B200SP_K_COO_RING && no_shape) {  // profiles/r02_results.md: 512x7 for fp64, 256x9 for fp32
    if (elem == 8)
      c.block_size = 512;
    else
      c.unroll = 9;
  }
  if (c.kernel == B200SP_K_COO_RING) {
    if (c.stages == 0) c.stages = 2;
    if (c.ctas_per_sm == 0) c.ctas_per_sm = 4;
  }
}

// spmv_coo_plan.cu
b200sp_coo_plan coo_attached_plan(b200sp_handle h, i64 rows, i64 cols, i64 nnz, const int *Ai, const int *Aj,
                                  size_t elem);
template <typename T>
b200sp_status spmv_coo_attached(b200sp_handle h, cudaStream_t st, b200sp_coo_plan p, const T *Ax, const T *x, T *y,
                                int accumulate, const b200sp_cfg *cfg);

// K_COO_WARP with the default shape for >= 2^27 scattered entries (v8, two units per tile) ran 4 % faster from a
// persistent grid than with one tile per warp on the boxes of round 2's first session (1.066 against 1.144 ms on R-MAT
// scale 24) and 10 - 15 % slower on the boxes of the second (1.23 - 1.28 against 1.12 ms; same binary, the one-shot
// time did not move: profiles/r04_gather_probe.md).  Both grids cut the entries into the same tiles and add them in
// the same order, so y is bit-identical either way — the first product with a given set of arrays times both
// (second of two launches each, CUDA events on the caller's stream, y = A x only: every launch rewrites y completely)
// and later products use the faster one.  A hint cached per (column_indices, nnz, element size) like the gather
// class; skipped (one tile per warp) while the stream is being captured.
template <typename T>
static b200sp_status spmv_coo_warp_auto_grid(b200sp_handle h, cudaStream_t st, const CooArgs<T> &a, b200sp_cfg c) {
  const b200sp_context::CsrKey key{a.Aj, (int64_t)sizeof(T), a.nnz};
  auto it = h->coo_grid_choice.find(key);
  if (it != h->coo_grid_choice.end()) {
    c.ctas_per_sm = it->second ? 8 : 0;
    return spmv_coo_warp<T>(h, st, a, c);
  }
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (a.accumulate || cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
    cudaGetLastError();  // y += A x cannot be repeated, a capture cannot be timed: one tile per warp, nothing cached
    c.ctas_per_sm = 0;
    return spmv_coo_warp<T>(h, st, a, c);
  }
  while (h->coo_choice_events.size() < 3) {
    cudaEvent_t e;
    B200SP_CUDA(h, cudaEventCreate(&e));
    h->coo_choice_events.push_back(e);
  }
  cudaEvent_t e0 = (cudaEvent_t)h->coo_choice_events[0], e1 = (cudaEvent_t)h->coo_choice_events[1],
              e2 = (cudaEvent_t)h->coo_choice_events[2];
  b200sp_cfg one = c, per = c;
  one.ctas_per_sm = 0;
  per.ctas_per_sm = 8;
  b200sp_status s;
  if ((s = spmv_coo_warp<T>(h, st, a, one)) != B200SP_OK) return s;
  if ((s = spmv_coo_warp<T>(h, st, a, per)) != B200SP_OK) return s;
  B200SP_CUDA(h, cudaEventRecord(e0, st));
  if ((s = spmv_coo_warp<T>(h, st, a, one)) != B200SP_OK) return s;
  B200SP_CUDA(h, cudaEventRecord(e1, st));
  if ((s = spmv_coo_warp<T>(h, st, a, per)) != B200SP_OK) return s;
  B200SP_CUDA(h, cudaEventRecord(e2, st));
  B200SP_CUDA(h, cudaEventSynchronize(e2));
  float ms_one = 0.f, ms_per = 0.f;
  B200SP_CUDA(h, cudaEventElapsedTime(&ms_one, e0, e1));
  B200SP_CUDA(h, cudaEventElapsedTime(&ms_per, e1, e2));
  if (h->coo_grid_choice.size() > 256) h->coo_grid_choice.clear();
  h->coo_grid_choice[key] = ms_per < 0.98f * ms_one ? 1 : 0;  // the persistent grid has to win by 2 %
  return B200SP_OK;
}

template <typename T>
b200sp_status spmv_coo(b200sp_handle h, cudaStream_t st, i64 rows, i64 cols, i64 nnz, const int *Ai,
                       const int *Aj, const T *Ax, const T *x, T *y, int accumulate,
                       const b200sp_cfg *cfg) {
  B200SP_CHECK_HANDLE(h);
  B200SP_REQUIRE(h, rows >= 0 && cols >= 0 && nnz >= 0, "coo: negative dimension");
  B200SP_REQUIRE(h, rows < (1ll << 31) && cols < (1ll << 31), "coo: int32 index range");
  if (rows == 0) return B200SP_OK;
  B200SP_REQUIRE(h, y != nullptr, "coo: null pointer");
  if (!accumulate) B200SP_CUDA(h, cudaMemsetAsync(y, 0, (size_t)rows * sizeof(T), st));
  if (nnz == 0) return B200SP_OK;
  B200SP_REQUIRE(h, Ai && Aj && Ax && x, "coo: null pointer");
  B200SP_REQUIRE(h, cols > 0, "coo: num_cols == 0 with stored entries");

  if (!h->coo_plans.empty() && (!cfg || cfg->kernel == 0 || cfg->kernel == B200SP_K_COO_WARP)) {
    // an attached plan for exactly these arrays (b200sp_coo_plan_attach): hot columns of x from shared memory
    if (b200sp_coo_plan p = coo_attached_plan(h, rows, cols, nnz, Ai, Aj, sizeof(T)))
      return spmv_coo_attached<T>(h, st, p, Ax, x, y, accumulate, cfg);
  }
  b200sp_cfg c = cfg ? *cfg : b200sp_cfg{};
  const bool tma_ok = aligned16(Ai) && aligned16(Aj) && aligned16(Ax);
  const bool vec32_ok = ((((uintptr_t)Ai | (uintptr_t)Aj | (uintptr_t)Ax) & 31) == 0);
  const bool no_cfg = c.kernel == 0 && c.block_size == 0 && c.unroll == 0;
  coo_defaults(c, h, st, nnz, Aj, sizeof(T), tma_ok, vec32_ok);
  // the default for >= 2^27 scattered entries: persistent or one-shot grid by measurement (below)
  const bool auto_grid = no_cfg && c.kernel == B200SP_K_COO_WARP && c.ctas_per_sm == 8;
  if (c.kernel == B200SP_K_COO_RING && !tma_ok) c.kernel = B200SP_K_COO_SEGSCAN;  // bulk copies need 16-byte bases
  if (c.kernel != B200SP_K_COO_SEGSCAN && c.kernel != B200SP_K_COO_RING && c.kernel != B200SP_K_COO_WARP)
    return set_error(h, B200SP_INVALID_INPUT, "coo: unknown kernel id %d", c.kernel);

  CooArgs<T> a;
  a.rows = rows; a.cols = cols; a.nnz = nnz; a.Ai = Ai; a.Aj = Aj; a.Ax = Ax; a.x = x; a.y = y;
  // y = A x: y was zeroed above, so complete rows are stored without reading y back
  a.accumulate = accumulate;
  a.carry = nullptr;
  a.Ap = nullptr;
  a.tile_first_row = nullptr;
  a.scalar_loads = 0;
  if (c.kernel == B200SP_K_COO_WARP) {
    // vector loads need 16- / 32-byte aligned bases; otherwise the scalar-load kernel gives the same sums
    const uintptr_t m = (uintptr_t)(c.vector_width == 8 ? 31 : 15);
    if ((((uintptr_t)Ai | (uintptr_t)Aj | (uintptr_t)Ax) & m) == 0) {
      if (auto_grid) return spmv_coo_warp_auto_grid<T>(h, st, a, c);
      return spmv_coo_warp<T>(h, st, a, c);
    }
    c = b200sp_cfg{};
    c.kernel = B200SP_K_COO_SEGSCAN;
    c.block_size = 256;
    c.unroll = 7;
  }
  if (c.kernel == B200SP_K_COO_RING) {
    if (c.stages < 2 || c.stages > 8 || c.ctas_per_sm < 1 || c.ctas_per_sm > 16)
      return set_error(h, B200SP_INVALID_INPUT, "coo ring: unsupported stages=%d ctas_per_sm=%d", c.stages,
                       c.ctas_per_sm);
#define CASE(B, V) \
  if (c.block_size == B && c.unroll == V) return launch_coo_ring<T, B, V>(h, st, a, c.stages, c.ctas_per_sm);
    CASE(128, 7) CASE(128, 9) CASE(128, 11)
    CASE(256, 5) CASE(256, 7) CASE(256, 9)
    CASE(512, 5) CASE(512, 7)
#undef CASE
    return set_error(h, B200SP_INVALID_INPUT, "coo ring: unsupported block_size=%d unroll=%d", c.block_size, c.unroll);
  }
#define CASE(B, V) \
  if (c.block_size == B && c.unroll == V) return launch_coo<T, B, V>(h, st, a);
  CASE(128, 5) CASE(128, 7) CASE(128, 9) CASE(128, 11)
  CASE(256, 5) CASE(256, 7) CASE(256, 9) CASE(256, 11)
  CASE(512, 5) CASE(512, 7)
#undef CASE
  return set_error(h, B200SP_INVALID_INPUT, "coo: unsupported block_size=%d unroll=%d", c.block_size, c.unroll);
}

template b200sp_status spmv_coo<float>(b200sp_handle, cudaStream_t, i64, i64, i64, const int *, const int *,
                                       const float *, const float *, float *, int, const b200sp_cfg *);
template b200sp_status spmv_coo<double>(b200sp_handle, cudaStream_t, i64, i64, i64, const int *, const int *,
                                        const double *, const double *, double *, int, const b200sp_cfg *);

// declared in spmv_ell.cu
template <typename T>
b200sp_status spmv_ell(b200sp_handle h, cudaStream_t st, i64 rows, i64 cols, i64 K, i64 pitch,
                       const int *cidx, const T *vals, const int *row_lengths, const T *x, T *y,
                       int accumulate, const b200sp_cfg *cfg, const T *dotv, T *dot_result);

// spmv_hyb_fused.cu
template <typename T>
b200sp_status spmv_hyb_fused(b200sp_handle h, cudaStream_t st, i64 rows, i64 cols, i64 K, i64 pitch, const int *ecidx,
                             const T *evals, i64 cnnz, const int *ci, const int *cj, const T *cv, const T *x, T *y,
                             int accumulate, const b200sp_cfg &c, int *done);

// HYB = ELL pass (init = caller's) then COO pass with identity
// (cusp/system/detail/sequential/multiply/hyb_spmv.h:35-57)
template <typename T>
b200sp_status spmv_hyb(b200sp_handle h, cudaStream_t st, i64 rows, i64 cols, i64 K, i64 pitch,
                       const int *ecidx, const T *evals, i64 cnnz, const int *ci, const int *cj,
                       const T *cv, const T *x, T *y, int accumulate, const b200sp_cfg *ecfg,
                       const b200sp_cfg *ccfg) {
  B200SP_CHECK_HANDLE(h);
  const char *fused_env = getenv("B200SP_HYB_FUSED");  // opt-in (measured slower, spmv_hyb_fused.cu)
  if (fused_env && fused_env[0] == '1' && cnnz > 0 && rows > 0 && rows < (1ll << 31) && cols > 0 && cols < (1ll << 31) && K >= 0 && ci && cj && cv && x && y &&
      (K == 0 || (ecidx && evals && pitch >= rows)) && (!ecfg || ecfg->kernel == 0) &&
      (!ccfg || ccfg->kernel == 0 || ccfg->kernel == B200SP_K_COO_WARP)) {
    // one pass over y (spmv_hyb_fused.cu) when the tail would run the warp-tile kernel anyway, no hot-column plan
    // is attached to it, and no tile would own a long stretch of tail-free rows
    const bool planned = !h->coo_plans.empty() && coo_attached_plan(h, rows, cols, cnnz, ci, cj, sizeof(T)) != nullptr;
    if (!planned) {
      b200sp_cfg c = ccfg ? *ccfg : b200sp_cfg{};
      const bool tma_ok = aligned16(ci) && aligned16(cj) && aligned16(cv);
      const bool vec32_ok = ((((uintptr_t)ci | (uintptr_t)cj | (uintptr_t)cv) & 31) == 0);
      coo_defaults(c, h, st, cnnz, cj, sizeof(T), tma_ok, vec32_ok);
      if (c.kernel == B200SP_K_COO_WARP) {
        int done = 0;
        b200sp_status s = spmv_hyb_fused<T>(h, st, rows, cols, K, pitch, ecidx, evals, cnnz, ci, cj, cv, x, y, accumulate,
                                            c, &done);
        if (s != B200SP_OK || done) return s;
      }
    }
  }
  if (!accumulate && cnnz > 0 && K == 1 && rows > 0) {
    // y = A x with a one-column ELL part (what the reference's split rule gives power-law graphs):
    // tail first (y zeroed, complete rows stored without reading y), then the ELL column
    // accumulating with coalesced y reads.  Per row the only add is tail + ell instead of
    // ell + tail (commutative), so the bits equal the ELL-then-COO order.
    b200sp_status s = spmv_coo<T>(h, st, rows, cols, cnnz, ci, cj, cv, x, y, 0, ccfg);
    if (s != B200SP_OK) return s;
    return spmv_ell<T>(h, st, rows, cols, K, pitch, ecidx, evals, nullptr, x, y, 1, ecfg, nullptr, nullptr);
  }
  b200sp_status s = spmv_ell<T>(h, st, rows, cols, K, pitch, ecidx, evals, nullptr, x, y, accumulate, ecfg,
                                nullptr, nullptr);
  if (s != B200SP_OK) return s;
  return spmv_coo<T>(h, st, rows, cols, cnnz, ci, cj, cv, x, y, 1, ccfg);
}

template b200sp_status spmv_hyb<float>(b200sp_handle, cudaStream_t, i64, i64, i64, i64, const int *,
                                       const float *, i64, const int *, const int *, const float *,
                                       const float *, float *, int, const b200sp_cfg *, const b200sp_cfg *);
template b200sp_status spmv_hyb<double>(b200sp_handle, cudaStream_t, i64, i64, i64, i64, const int *,
                                        const double *, i64, const int *, const int *, const double *,
                                        const double *, double *, int, const b200sp_cfg *, const b200sp_cfg *);

}  // namespace b200sp

extern "C" {
// row_starts[w] = the row that contains entry w * chunk, chunk = ceil(num_entries / workers); workers whose first
// entry lies beyond the matrix get 0 — cpu_compute_row_starts / gpu_compute_row_starts of the reference's balanced
// CSR kernel (cusp/system/cuda/ktt/csr_multiply.h:38-85).  K_CSR_BALANCED computes the same array with chunk = its
// tile size before every product (csr_tile_rows_kernel above).
b200sp_status b200sp_csr_row_starts(b200sp_handle h, b200sp_stream stream, int64_t num_rows, int64_t num_entries,
                                    const int32_t *row_offsets, int64_t workers, int32_t *row_starts) {
  B200SP_CHECK_HANDLE(h);
  B200SP_REQUIRE(h, num_rows >= 0 && num_entries >= 0 && workers >= 0 && workers < (1ll << 31), "csr_row_starts: bad sizes");
  if (workers == 0) return B200SP_OK;
  B200SP_REQUIRE(h, row_starts && (num_rows == 0 || row_offsets), "csr_row_starts: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  B200SP_CUDA(h, cudaMemsetAsync(row_starts, 0, (size_t)workers * sizeof(int32_t), st));
  if (num_rows == 0 || num_entries == 0) return B200SP_OK;
  const int64_t chunk = (num_entries + workers - 1) / workers;
  B200SP_REQUIRE(h, chunk < (1ll << 31), "csr_row_starts: chunk exceeds the int32 range");
  b200sp::csr_tile_rows_kernel<<<(unsigned)b200sp::ceil_div(num_rows, 256), 256, 0, st>>>(num_rows, row_offsets, (int)chunk,
                                                                                            row_starts);
  B200SP_LAUNCH_CHECK(h, "csr_tile_rows_kernel");
  return B200SP_OK;
}

#define DEF(T, sfx)                                                                              \
  b200sp_status b200sp_spmv_coo_##sfx(b200sp_handle h, b200sp_stream stream, int64_t num_rows,   \
                                      int64_t num_cols, int64_t num_entries,                     \
                                      const int32_t *row_indices, const int32_t *column_indices, \
                                      const T *values, const T *x, T *y, int accumulate,         \
                                      const b200sp_cfg *cfg) {                                   \
    return b200sp::spmv_coo<T>(h, (cudaStream_t)stream, num_rows, num_cols, num_entries,         \
                               row_indices, column_indices, values, x, y, accumulate, cfg);      \
  }                                                                                              \
  b200sp_status b200sp_spmv_hyb_##sfx(                                                           \
      b200sp_handle h, b200sp_stream stream, int64_t num_rows, int64_t num_cols,                 \
      int64_t ell_cols_per_row, int64_t ell_pitch, const int32_t *ell_column_indices,            \
      const T *ell_values, int64_t coo_num_entries, const int32_t *coo_row_indices,              \
      const int32_t *coo_column_indices, const T *coo_values, const T *x, T *y, int accumulate,  \
      const b200sp_cfg *ell_cfg, const b200sp_cfg *coo_cfg) {                                    \
    return b200sp::spmv_hyb<T>(h, (cudaStream_t)stream, num_rows, num_cols, ell_cols_per_row,    \
                               ell_pitch, ell_column_indices, ell_values, coo_num_entries,       \
                               coo_row_indices, coo_column_indices, coo_values, x, y,            \
                               accumulate, ell_cfg, coo_cfg);                                    \
  }
DEF(float, f32)
DEF(double, f64)
#undef DEF
}
