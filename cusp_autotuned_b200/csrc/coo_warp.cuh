// coo_warp.cuh — K_COO_WARP: warp-autonomous COO SpMV for gather-bound operators (power-law graphs).
//
// Same semantics as the other COO kernels (host loop cusp/system/detail/sequential/multiply/
// coo_spmv.h:35-68, rows sorted, duplicates allowed; replaces generic/multiply/spmv.h:182-238 and
// the atomicAdd kernels of cusp/system/cuda/ktt/kernels/coo_kernel.h:25-392), different mapping:
//
//   * a WARP owns a tile of U units x 32 lanes x VPL consecutive entries; every lane reads its VPL
//     consecutive rows / columns / values with one 128- or 256-bit load per array (cfg.vector_width
//     = VPL = 4 | 8; LDG.E.128 / LDG.E.ENL2.256 in SASS), so the three entry streams cost 3 LSU
//     instructions per 4 / 8 entries instead of 3 per entry, and nothing is transposed through
//     shared memory;
//   * products stay in registers: serial reduction over the lane's VPL entries, then one
//     ballot-bounded shuffle scan per unit (5 SHFL per 32 x VPL entries) stitches rows across lanes,
//     a register carry stitches units — no shared memory, no CTA barrier, no atomics;
//   * rows that end inside a tile are stored by the lane that sees the end; the row a tile shares
//     with its predecessor / successor goes to the tile's carry record, added up in tile order by
//     coo_fixup_kernel (deterministic for a given (nnz, VPL, U)).
//
// Why: ncu on K_COO_SEGSCAN over R-MAT scale 22 (profiles/r02_ncu_coo.md) showed the MIO pipe shared
// by the x gathers and the kernel's own shared-memory traffic (mio_throttle 8.5, barrier 4.4 cycles
// per issue); here the only MIO work besides the gathers is 7 shuffles per unit.
//
// With TABLE (the plan executor, b200sp_coo_plan_*): columns whose index has the sign bit set read
// x from a per-CTA shared-memory table of the matrix's most frequent columns instead of L1TEX.
#pragma once

#include "coo.cuh"

namespace b200sp {

// ---- 128 / 256-bit streaming loads --------------------------------------------------------------
// SPOL 0: ld.global.cs (evict-first in L1 and L2)   SPOL 1: L1::no_allocate + L2 evict-first policy
template <int SPOL>
__device__ __forceinline__ void ld_words4(const void *p, uint32_t *o, uint64_t pol) {
  if (SPOL == 0)
    asm volatile("ld.global.cs.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]) : "l"(p));
  else
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.b32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3])
                 : "l"(p), "l"(pol));
}
template <int SPOL>
__device__ __forceinline__ void ld_words8(const void *p, uint32_t *o, uint64_t pol) {
  if (SPOL == 0)
    asm volatile("ld.global.cs.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]), "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7])
                 : "l"(p));
  else
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]), "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7])
                 : "l"(p), "l"(pol));
}
template <int SPOL, int W>
__device__ __forceinline__ void ld_words(const void *p, uint32_t *o, uint64_t pol) {
  static_assert(W % 4 == 0, "whole 16-byte pieces");
  if (W % 8 == 0) {
#pragma unroll
    for (int i = 0; i < W; i += 8) ld_words8<SPOL>(reinterpret_cast<const char *>(p) + 4 * i, o + i, pol);
  } else {
#pragma unroll
    for (int i = 0; i < W; i += 4) ld_words4<SPOL>(reinterpret_cast<const char *>(p) + 4 * i, o + i, pol);
  }
}
__device__ __forceinline__ float from_words(const uint32_t *w, float) { return __uint_as_float(w[0]); }
__device__ __forceinline__ double from_words(const uint32_t *w, double) { return __hiloint2double((int)w[1], (int)w[0]); }

// ---- x gathers ----------------------------------------------------------------------------------
// XPOL 0: ld.global.nc   1: ld.global.nc.L1::evict_last   2: ld.global.nc.L1::no_allocate
template <int XPOL>
__device__ __forceinline__ float ld_x(const float *p) {
  float v;
  if (XPOL == 0) asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  else if (XPOL == 1) asm volatile("ld.global.nc.L1::evict_last.f32 %0, [%1];" : "=f"(v) : "l"(p));
  else asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
template <int XPOL>
__device__ __forceinline__ double ld_x(const double *p) {
  double v;
  if (XPOL == 0) asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
  else if (XPOL == 1) asm volatile("ld.global.nc.L1::evict_last.f64 %0, [%1];" : "=d"(v) : "l"(p));
  else asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

constexpr unsigned COO_HOT_FLAG = 0x80000000u;  // remapped column: sign bit | table slot

template <typename T, typename Ops>
__device__ __forceinline__ void coo_store_y(T *y, int row, T val, int accumulate) {
  y[row] = accumulate ? Ops::reduce(y[row], val) : val;
}

// ---- fused HYB (spmv_hyb_fused.cu): what a COO-tail tile needs to know about the ELL part -----------------------
template <typename T>
struct HybEll {
  const int *cidx;
  const T *vals;
  i64 pitch;
  int K;
  i64 rows;
  int accumulate;        // the caller's: y = y + A x
  int2 *work;            // row ranges without tail entries left to hyb_gap_kernel: [x, y] inclusive
  unsigned *work_count;  // ranges appended so far
  unsigned work_cap;
};
constexpr int HYB_GAP_INLINE = 8;    // shorter runs of tail-free rows are finished by the lane that meets them
constexpr int HYB_GAP_CHUNK = 1024;  // longer ones are cut into work items of at most this many rows

// init(y[r]) + the ELL slots of row r in ascending k: the value the ELL kernels leave in y[r]
template <typename T>
__device__ __noinline__ T hyb_row_init(const HybEll<T> &e, const T *x, const T *y, unsigned cols, i64 r) {
  T acc = e.accumulate ? y[r] : T(0);
  for (int k0 = 0; k0 < e.K; k0 += 4) {
    int c[4];
    T v[4], xv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const bool kin = k0 + k < e.K;
      const i64 so = (i64)(kin ? k0 + k : k0) * e.pitch + r;
      c[k] = ld_stream(e.cidx + so);
      v[k] = ld_stream(e.vals + so);
      if (!kin) c[k] = -1;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) xv[k] = ld_ro(x + min((unsigned)max(c[k], 0), cols - 1));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const T t = acc + v[k] * xv[k];
      acc = (c[k] != -1) ? t : acc;
    }
  }
  return acc;
}

// rows [lo, hi] hold no tail entry: finish them here when they are few, otherwise leave them to hyb_gap_kernel
template <typename T>
__device__ __noinline__ void hyb_gap(const HybEll<T> &e, const T *x, T *y, unsigned cols, i64 lo, i64 hi) {
  const i64 len = hi - lo + 1;
  if (len <= 0) return;
  if (len >= HYB_GAP_INLINE) {
    const unsigned n = (unsigned)((len + HYB_GAP_CHUNK - 1) / HYB_GAP_CHUNK);
    const unsigned base = atomicAdd(e.work_count, n);
    if (base + n <= e.work_cap) {  // always, by the capacity bound of launch_hyb_warp
      for (unsigned i = 0; i < n; ++i) {
        const i64 b = lo + (i64)i * HYB_GAP_CHUNK;
        e.work[base + i] = make_int2((int)b, (int)min(hi, b + HYB_GAP_CHUNK - 1));
      }
      return;
    }
    atomicSub(e.work_count, n);
  }
  for (i64 r = lo; r <= hi; ++r) y[r] = hyb_row_init(e, x, y, cols, r);
}

// One warp, one tile of U * 32 * VPL consecutive entries.  HYB (fused HYB kernel only): `owner` tiles finish their
// rows themselves — a row that ends here is stored as init + ELL slots + tail sum by the lane that sees its end, runs
// of rows without tail entries between two entries of the tile go to hyb_gap; other tiles accumulate into the y
// their warp has just initialised (a.accumulate = 1).
// PREY (y += A x on tails whose entries mostly end a row — one entry per row): the y values of the rows that end at an
// entry are fetched together with the x gathers instead of by a read-modify-write after the scan — two dependent
// round trips per tile instead of three; every row has one writer in this kernel, so the early read sees the same
// value, and the same two operands are added: bit-identical.
template <typename T, int VPL, int U, int XPOL, int SPOL, bool TABLE, typename Ops, bool HYB = false, bool PREY = false>
__device__ __forceinline__ void coo_warp_tile(const CooArgs<T> &a, const i64 tile, const int lane, const T *tab,
                                              const uint64_t pol, const HybEll<T> *hyb = nullptr,
                                              const bool owner = false) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int UNIT = 32 * VPL, WT = UNIT * U;
  constexpr int VW = VPL * (int)sizeof(T) / 4;  // 32-bit words of a lane's values
  const i64 start = tile * (i64)WT;
  const int n = (int)min((i64)WT, a.nnz - start);
  const unsigned cols = (unsigned)a.cols;

  int r[U][VPL], c[U][VPL];
  T v[U][VPL];
  bool vector_loads = false;
  if constexpr (VPL >= 4) {  // VPL = 1: a lane takes every 32nd entry of a unit — scalar, fully coalesced loads (below)
    vector_loads = n == WT && !a.scalar_loads;
    if (vector_loads) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const i64 e0 = start + u * UNIT + lane * VPL;
        uint32_t wr[VPL], wc[VPL], wv[VW > 0 ? VW : 1];
        ld_words<SPOL, VPL>(a.Ai + e0, wr, pol);
        ld_words<SPOL, VPL>(a.Aj + e0, wc, pol);
        ld_words<SPOL, VW>(a.Ax + e0, wv, pol);
#pragma unroll
        for (int q = 0; q < VPL; ++q) {
          r[u][q] = (int)wr[q];
          c[u][q] = (int)wc[q];
          v[u][q] = from_words(wv + q * ((int)sizeof(T) / 4), T());
        }
      }
    }
  }
  if (!vector_loads) {  // the last tile of the matrix (or unaligned arrays): guarded scalar loads, absent entries get row -1
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int q = 0; q < VPL; ++q) {
        const int idx = u * UNIT + lane * VPL + q;
        const bool ok = idx < n;
        const i64 g = ok ? start + idx : start;
        r[u][q] = ok ? ld_stream(a.Ai + g) : -1;
        c[u][q] = ok ? ld_stream(a.Aj + g) : 0;
        v[u][q] = ok ? ld_stream(a.Ax + g) : T(0);
      }
  }
  const int next_row = (start + WT < a.nnz) ? ld_ro(a.Ai + start + WT) : -1;
  const int prev_row = (start > 0) ? ld_ro(a.Ai + start - 1) : -1;

  // ---- gathers: all of the tile's loads are issued before the first product ----
  T xv[U][VPL];
#pragma unroll
  for (int u = 0; u < U; ++u)
#pragma unroll
    for (int q = 0; q < VPL; ++q) {
      const int cc = c[u][q];
      if (TABLE && cc < 0)
        xv[u][q] = tab[(unsigned)cc & ~COO_HOT_FLAG];
      else
        xv[u][q] = ld_x<XPOL>(a.x + min((unsigned)cc, cols - 1));
    }

  T yv[PREY ? U : 1][PREY ? VPL : 1];
  if (PREY) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int rn = __shfl_down_sync(FULL, r[u][0], 1);
      int r_after = next_row;
      if (u + 1 < U) r_after = __shfl_sync(FULL, r[(u + 1 < U) ? u + 1 : u][0], 0);
      if (lane == 31) rn = r_after;
#pragma unroll
      for (int q = 0; q < VPL; ++q) {
        const int nxt = (q + 1 < VPL) ? r[u][(q + 1 < VPL) ? q + 1 : q] : rn;
        T t = Ops::identity();
        if (a.accumulate && r[u][q] >= 0 && r[u][q] != nxt) t = a.y[r[u][q]];
        yv[PREY ? u : 0][PREY ? q : 0] = t;
      }
    }
  }

  T wcarry = Ops::identity();  // the open row's sum since its last row end in earlier units of this tile
  bool seen_end = false;  // a row ended earlier in this tile (warp-uniform)
  CooCarry<T> *rec = a.carry + tile;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    int rn = __shfl_down_sync(FULL, r[u][0], 1);
    int r_after = next_row;
    if (u + 1 < U) r_after = __shfl_sync(FULL, r[(u + 1 < U) ? u + 1 : u][0], 0);
    if (lane == 31) rn = r_after;

    // serial segmented reduction over the lane's VPL consecutive entries
    T run = Ops::identity(), head = Ops::identity(), head_y = Ops::identity();
    int head_row = -1;
    bool has_b = false;
#pragma unroll
    for (int q = 0; q < VPL; ++q) {
      const T p = (r[u][q] >= 0) ? Ops::combine(v[u][q], xv[u][q]) : Ops::identity();
      run = Ops::reduce(run, p);
      const int nxt = (q + 1 < VPL) ? r[u][(q + 1 < VPL) ? q + 1 : q] : rn;
      if (r[u][q] != nxt) {  // row r[u][q] ends at this entry
        if (!has_b) {
          head = run;
          head_row = r[u][q];
          if (PREY) head_y = yv[PREY ? u : 0][PREY ? q : 0];
          has_b = true;
        } else {
          // began and ended inside this lane
          if (HYB && owner) a.y[r[u][q]] = hyb_row_init(*hyb, a.x, a.y, cols, r[u][q]) + run;
          else if (PREY) a.y[r[u][q]] = a.accumulate ? Ops::reduce(yv[PREY ? u : 0][PREY ? q : 0], run) : run;
          else coo_store_y<T, Ops>(a.y, r[u][q], run, a.accumulate);
        }
        run = Ops::identity();
        if (HYB && owner && nxt > r[u][q] + 1 && !(lane == 31 && u == U - 1 && q == VPL - 1))
          hyb_gap(*hyb, a.x, a.y, cols, (i64)r[u][q] + 1, (i64)nxt - 1);  // tail-free rows up to the tile's next entry
      }
    }

    // inclusive scan of the lanes' open sums, restarted at every lane that saw a row end
    const unsigned m = __ballot_sync(FULL, has_b);
    const unsigned le = m & (FULL >> (31 - lane));  // row ends at lanes <= this one
    const int j = le ? 31 - __clz(le) : 0;          // the scan of this lane reaches back to lane j
    T vi = run;
    if (m != FULL) {  // every lane saw a row end (one entry per row): nothing crosses a lane, no scan
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const T up = __shfl_up_sync(FULL, vi, d);
        if (lane - d >= j) vi = Ops::reduce(up, vi);
      }
    }
    const T Vi = (le == 0) ? Ops::reduce(wcarry, vi) : vi;  // no row end so far in this unit: the earlier units' sum joins
    T cin = __shfl_up_sync(FULL, Vi, 1);
    if (lane == 0) cin = wcarry;
    if (has_b) {
      const T total = Ops::reduce(cin, head);
      const bool first_end = !seen_end && (m & ((1u << lane) - 1u)) == 0;  // first row end of the tile
      if (first_end) {
        const bool continued = (head_row == prev_row);  // the row came in from the previous tile
        rec->head_row = continued ? head_row : -1;
        rec->head_val = continued ? total : Ops::identity();
        if (!continued) {
          if (HYB && owner) a.y[head_row] = hyb_row_init(*hyb, a.x, a.y, cols, head_row) + total;
          else if (PREY) a.y[head_row] = a.accumulate ? Ops::reduce(head_y, total) : total;
          else coo_store_y<T, Ops>(a.y, head_row, total, a.accumulate);
        }
      } else {
        if (HYB && owner) a.y[head_row] = hyb_row_init(*hyb, a.x, a.y, cols, head_row) + total;
        else if (PREY) a.y[head_row] = a.accumulate ? Ops::reduce(head_y, total) : total;
        else coo_store_y<T, Ops>(a.y, head_row, total, a.accumulate);
      }
    }
    wcarry = __shfl_sync(FULL, Vi, 31);
    seen_end = seen_end || (m != 0);
  }
  if (lane == 31) {
    if (!seen_end) {
      rec->head_row = -1;
      rec->head_val = Ops::identity();
    }
    const int last_row = r[U - 1][VPL - 1];
    const bool open = last_row >= 0 && last_row == next_row;
    rec->tail_row = open ? last_row : -1;
    rec->tail_val = open ? wcarry : Ops::identity();
    rec->leader = (open && last_row != prev_row) ? 1 : 0;
    rec->pad = 0;
  }
}

template <typename T, int BLOCK, int MINB, int VPL, int U, int XPOL, int SPOL, bool TABLE, typename Ops = SpmvOps<T, 0, 0>,
          bool PREY = false>
__global__ void __launch_bounds__(BLOCK, MINB) coo_warp_kernel(CooArgs<T> a, i64 num_tiles, const int *hot_cols,
                                                               int hot) {
  extern __shared__ __align__(16) unsigned char coo_warp_smem[];
  T *tab = reinterpret_cast<T *>(coo_warp_smem);
  if (TABLE) {
    for (int i = threadIdx.x; i < hot; i += BLOCK) tab[i] = ld_ro(a.x + ld_ro(hot_cols + i));
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const i64 stride = (i64)gridDim.x * (BLOCK / 32);
  const uint64_t pol = SPOL ? l2_policy_evict_first() : 0;
  for (i64 tile = (i64)blockIdx.x * (BLOCK / 32) + (threadIdx.x >> 5); tile < num_tiles; tile += stride)
    coo_warp_tile<T, VPL, U, XPOL, SPOL, TABLE, Ops, false, PREY>(a, tile, lane, tab, pol);
}

template <typename T, int BLOCK, int MINB, int VPL, int U, int XPOL, int SPOL, bool TABLE, typename Ops = SpmvOps<T, 0, 0>,
          bool PREY = false>
static b200sp_status launch_coo_warp(b200sp_handle h, cudaStream_t st, CooArgs<T> a, int ctas_per_sm,
                                     const int *hot_cols, int hot, int capacity) {
  constexpr int WT = 32 * VPL * U;
  const i64 tiles = ceil_div(a.nnz, (i64)WT);
  b200sp_status s = ensure_scratch(h, (size_t)tiles * sizeof(CooCarry<T>));
  if (s != B200SP_OK) return s;
  a.carry = reinterpret_cast<CooCarry<T> *>(h->scratch);
  auto kern = coo_warp_kernel<T, BLOCK, MINB, VPL, U, XPOL, SPOL, TABLE, Ops, PREY>;
  const size_t smem = TABLE ? (size_t)capacity * sizeof(T) : 0;
  if constexpr (TABLE) {
    if (smem > (size_t)h->max_smem_optin)
      return set_error(h, B200SP_INVALID_INPUT, "coo plan: table of %zu B exceeds %d", smem, h->max_smem_optin);
    B200SP_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  } else {
    // no shared memory: give the whole unified array to L1 (x sectors of hot columns stay on the SM)
    B200SP_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 0));
  }
  i64 grid = ceil_div(tiles, (i64)(BLOCK / 32));
  if (TABLE && ctas_per_sm <= 0) ctas_per_sm = 1;
  if (ctas_per_sm > 0) {
    int resident = 0;
    B200SP_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, BLOCK, smem));
    if (resident < 1) return set_error(h, B200SP_INVALID_INPUT, "coo warp: configuration does not fit on an SM");
    const i64 persistent = (i64)h->num_sms * (ctas_per_sm < resident ? ctas_per_sm : resident);
    if (persistent < grid) grid = persistent;
  }
  kern<<<(unsigned)grid, BLOCK, smem, st>>>(a, tiles, hot_cols, hot);
  B200SP_LAUNCH_CHECK(h, "coo_warp_kernel");
  coo_fixup_kernel<T, Ops><<<(unsigned)ceil_div(tiles, 256), 256, 0, st>>>(tiles, a.carry, a.y, a.accumulate);
  B200SP_LAUNCH_CHECK(h, "coo_fixup_kernel");
  return B200SP_OK;
}

}  // namespace b200sp
