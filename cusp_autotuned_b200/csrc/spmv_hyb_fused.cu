// spmv_hyb_fused.cu — HYB product in ONE pass over y (VERDICT r1 "single-pass HYB").
//
// The reference runs the ELL part and then the COO part accumulating into y
// (cusp/system/detail/generic/multiply/spmv.h:272-290, sequential/multiply/hyb_spmv.h:35-57); the two-launch form
// here (spmv_hyb in spmv_coo.cu) costs a second launch and a read-modify-write of all of y between the parts
// (+ a memset when the tail runs first).  hyb_warp_kernel does both parts in the warp tiles of K_COO_WARP:
//
//   * the COO tail is cut into warp tiles of U*32*VPL consecutive entries exactly as in coo_warp.cuh;
//   * tile t also owns the ELL rows (p_t, L_t], p_t = row of the entry in front of the tile (-1 for tile 0),
//     L_t = row of the tile's last entry (num_rows-1 for the last tile): these ranges partition [0, num_rows);
//     the warp first computes  y[r] = init(y[r]) + sum_k ell(r, k) * x[col]  for its range — 32 consecutive rows
//     per step, coalesced column-major ELL slabs, slots in ascending k like the ELL kernels (bit-identical to them) —
//   * then runs its COO tile in accumulate mode: a row that ends inside the tile began after p_t, so its y was
//     initialised by this very warp (ordered by __syncwarp); rows that cross tile boundaries go through the carry
//     records and are added by coo_fixup_kernel after the grid, as in every COO kernel.
// Per row: y = (init + ELL slots in order) + tail sum — the same two values added in the same order as the
// two-launch form with the same tile shape, so the results are bit-identical to it.
//
// A tile's ELL range is as long as the gap in front of its rows (a tail concentrated in a few rows, the boundary
// planes of a stencil whose interior rows alone reach the tail).  Tiles whose range exceeds HYB_COOP_CAP rows do not
// walk it: they become "owner" tiles (coo_warp.cuh) — a row that ends in the tile is stored as init + ELL + tail by the
// lane that sees the end (no read-modify-write), the open last row is initialised by one lane, and every run of
// tail-free rows (in front of the tile, between two of its entries, behind the last entry of the matrix) is either
// finished on the spot (< 8 rows) or appended to a work list in pieces of at most 1024 rows; tail-free rows have no
// ordering constraint, so hyb_gap_kernel works the list off after the grid with a warp per piece.  Work per tile is
// bounded for every input, and the list holds at most rows/8 + rows/1024 pieces.
//
// MEASURED (B200, tools/hyb_probe.py, profiles/r04_hyb_single_pass.md): bit-identical to the two-launch form on every
// test matrix, and SLOWER — poisson7pt 256^3 split at K = 6 (one tail entry per row): 0.535 against 0.253 ms (fp32),
// 0.674 against 0.359 ms (fp64); R-MAT scale 24 (K = 1, 97 % of the entries in the tail): 1.965 against 1.143 ms.
// Both products are bound by the latency of dependent round trips per warp, not by bytes: a warp tile of the
// two-launch tail does two (entries, then x gathers); here the tile's row range has to be known first (one), then the
// ELL slots are fetched (two), gathered (three) and stored before the tail entries may accumulate into them (a fourth
// for the read-modify-write) — and an ELL part as long as the stencil's runs at the LDG rate inside a warp instead of
// the bulk-copy ring's.  What the second launch and the y round trip cost (5 - 12 % of the bytes) is less than what
// the serialised phases add.  Kept behind B200SP_HYB_FUSED=1 with its parity test; the default stays ELL launch +
// COO launch.
#include <stdlib.h>

#include "coo_warp.cuh"

namespace b200sp {

constexpr int HYB_COOP_CAP = 1024;  // rows a tile's warp initialises itself before its tail entries accumulate

// ELL rows [lo, hi] by one warp: 32 consecutive rows per step, RU steps in flight, slots in chunks of KC.
template <typename T>
__device__ __forceinline__ void hyb_ell_rows(const HybEll<T> &e, const T *x, T *y, unsigned cols, i64 lo, i64 hi,
                                             int lane) {
  constexpr int RU = 4, KC = 4;
  for (i64 r0 = lo + lane; r0 <= hi; r0 += 32 * RU) {
    T acc[RU];
    i64 rc[RU];
    bool ok[RU];
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      const i64 r = r0 + 32 * u;
      ok[u] = r <= hi;
      rc[u] = ok[u] ? r : hi;
      acc[u] = (ok[u] && e.accumulate) ? y[r] : T(0);
    }
    for (int k0 = 0; k0 < e.K; k0 += KC) {
      int c[KC][RU];
      T v[KC][RU], xv[KC][RU];
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        const bool kin = k0 + k < e.K;
        const i64 so = (i64)(kin ? k0 + k : k0) * e.pitch;
#pragma unroll
        for (int u = 0; u < RU; ++u) {
          c[k][u] = ld_stream(e.cidx + so + rc[u]);
          v[k][u] = ld_stream(e.vals + so + rc[u]);
          if (!kin) c[k][u] = -1;
        }
      }
#pragma unroll
      for (int k = 0; k < KC; ++k)
#pragma unroll
        for (int u = 0; u < RU; ++u) {
          pin(c[k][u]);
          xv[k][u] = ld_ro(x + min((unsigned)max(c[k][u], 0), cols - 1));
        }
#pragma unroll
      for (int k = 0; k < KC; ++k)
#pragma unroll
        for (int u = 0; u < RU; ++u) {
          pin(xv[k][u]);
          const T t = acc[u] + v[k][u] * xv[k][u];
          acc[u] = (c[k][u] != -1) ? t : acc[u];
        }
    }
#pragma unroll
    for (int u = 0; u < RU; ++u)
      if (ok[u]) y[rc[u]] = acc[u];
  }
}

template <typename T, int BLOCK, int MINB, int VPL, int U>
__global__ void __launch_bounds__(BLOCK, MINB) hyb_warp_kernel(CooArgs<T> a, HybEll<T> e, i64 num_tiles) {
  constexpr int WT = 32 * VPL * U;
  const int lane = threadIdx.x & 31;
  const unsigned cols = (unsigned)a.cols;
  const i64 stride = (i64)gridDim.x * (BLOCK / 32);
  for (i64 tile = (i64)blockIdx.x * (BLOCK / 32) + (threadIdx.x >> 5); tile < num_tiles; tile += stride) {
    const i64 start = tile * (i64)WT;
    const bool last_tile = tile == num_tiles - 1;
    const i64 p = (start > 0) ? (i64)ld_ro(a.Ai + start - 1) : -1;                     // row in front of the tile
    const i64 le = (i64)ld_ro(a.Ai + (last_tile ? a.nnz - 1 : start + WT - 1));        // row of the tile's last entry
    const i64 L = last_tile ? e.rows - 1 : le;
    const bool owner = L - p > HYB_COOP_CAP;
    if (!owner) {
      hyb_ell_rows<T>(e, a.x, a.y, cols, p + 1, L, lane);
    } else {
      if (lane == 0) hyb_gap(e, a.x, a.y, cols, p + 1, (i64)ld_ro(a.Ai + start) - 1);  // rows in front of the first entry
      if (lane == 1 && last_tile) hyb_gap(e, a.x, a.y, cols, le + 1, e.rows - 1);      // rows behind the last entry
      // the row left open at the end of the tile is completed by the carry fix-up: its ELL part goes in now
      if (lane == 2 && !last_tile && le > p && (i64)ld_ro(a.Ai + start + WT) == le)
        a.y[le] = hyb_row_init(e, a.x, a.y, cols, le);
    }
    __syncwarp();  // the range's y values are visible to every lane of this warp before the tail accumulates
    coo_warp_tile<T, VPL, U, 0, 0, false, SpmvOps<T, 0, 0>, true>(a, tile, lane, nullptr, 0, &e, owner);
  }
}

// the work list: a warp per piece of at most HYB_GAP_CHUNK tail-free rows
template <typename T>
__global__ void __launch_bounds__(256) hyb_gap_kernel(HybEll<T> e, const T *x, T *y, unsigned cols) {
  const unsigned n = min(*e.work_count, e.work_cap);
  const int lane = threadIdx.x & 31;
  for (unsigned i = blockIdx.x * 8 + (threadIdx.x >> 5); i < n; i += gridDim.x * 8) {
    const int2 w = e.work[i];
    hyb_ell_rows<T>(e, x, y, cols, w.x, w.y, lane);
  }
}

template <typename T, int MINB, int VPL, int U>
static b200sp_status launch_hyb_warp(b200sp_handle h, cudaStream_t st, CooArgs<T> a, HybEll<T> e, int ctas_per_sm) {
  constexpr int WT = 32 * VPL * U;
  const i64 tiles = ceil_div(a.nnz, (i64)WT);
  const size_t carry_bytes = ((size_t)tiles * sizeof(CooCarry<T>) + 15) & ~(size_t)15;
  const i64 work_cap = a.rows / HYB_GAP_INLINE + a.rows / HYB_GAP_CHUNK + 64;
  b200sp_status s = ensure_scratch(h, carry_bytes + (size_t)work_cap * sizeof(int2));
  if (s != B200SP_OK) return s;
  a.carry = reinterpret_cast<CooCarry<T> *>(h->scratch);
  a.accumulate = 1;  // tiles that are not owners accumulate into the y their warp has just initialised
  e.work = reinterpret_cast<int2 *>(reinterpret_cast<char *>(h->scratch) + carry_bytes);
  e.work_cap = (unsigned)work_cap;
  e.work_count = reinterpret_cast<unsigned *>(h->dev_scalars + 55);
  B200SP_CUDA(h, cudaMemsetAsync(e.work_count, 0, sizeof(unsigned), st));
  auto kern = hyb_warp_kernel<T, 256, MINB, VPL, U>;
  B200SP_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 0));
  i64 grid = ceil_div(tiles, (i64)8);
  if (ctas_per_sm > 0) {
    int resident = 0;
    B200SP_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, 256, 0));
    if (resident < 1) return set_error(h, B200SP_INVALID_INPUT, "hyb fused: configuration does not fit on an SM");
    const i64 persistent = (i64)h->num_sms * (ctas_per_sm < resident ? ctas_per_sm : resident);
    if (persistent < grid) grid = persistent;
  }
  kern<<<(unsigned)grid, 256, 0, st>>>(a, e, tiles);
  B200SP_LAUNCH_CHECK(h, "hyb_warp_kernel");
  coo_fixup_kernel<T, SpmvOps<T, 0, 0>><<<(unsigned)ceil_div(tiles, 256), 256, 0, st>>>(tiles, a.carry, a.y, 1);
  B200SP_LAUNCH_CHECK(h, "coo_fixup_kernel");
  const i64 gap_grid = min((i64)h->num_sms * 8, ceil_div(work_cap, (i64)8));
  hyb_gap_kernel<T><<<(unsigned)gap_grid, 256, 0, st>>>(e, a.x, a.y, (unsigned)a.cols);
  B200SP_LAUNCH_CHECK(h, "hyb_gap_kernel");
  return B200SP_OK;
}

// Returns B200SP_OK and sets *done = 1 when the fused kernel ran; *done = 0 (and OK) when the caller should take the
// two-launch form: B200SP_HYB_FUSED is not 1 (the default, see MEASURED in the header), arrays not aligned for the
// tile's vector loads, or a shape outside the instantiated set.  `c` is the COO configuration the tail would run with (kernel ==
// K_COO_WARP).
template <typename T>
b200sp_status spmv_hyb_fused(b200sp_handle h, cudaStream_t st, i64 rows, i64 cols, i64 K, i64 pitch, const int *ecidx,
                             const T *evals, i64 cnnz, const int *ci, const int *cj, const T *cv, const T *x, T *y,
                             int accumulate, const b200sp_cfg &c, int *done) {
  *done = 0;
  const char *env = getenv("B200SP_HYB_FUSED");
  const int mode = env ? atoi(env) : 0;  // opt-in: measured slower than the two-launch form (header)
  if (mode == 0 || c.kernel != B200SP_K_COO_WARP || cnnz <= 0 || rows <= 0 || K < 0 || K > (1 << 20)) return B200SP_OK;
  const int vpl = c.vector_width ? c.vector_width : 8, u = c.unroll ? c.unroll : 1;
  const uintptr_t m = (uintptr_t)(vpl == 8 ? 31 : 15);
  if ((((uintptr_t)ci | (uintptr_t)cj | (uintptr_t)cv) & m) != 0) return B200SP_OK;
  if (!((vpl == 4 || vpl == 8) && (u == 1 || u == 2))) return B200SP_OK;
  if (K > 0 && (!ecidx || !evals)) return set_error(h, B200SP_INVALID_INPUT, "hyb: null ELL arrays");
  CooArgs<T> a;
  a.rows = rows; a.cols = cols; a.nnz = cnnz; a.Ai = ci; a.Aj = cj; a.Ax = cv; a.x = x; a.y = y;
  a.accumulate = 1;
  a.carry = nullptr;
  a.Ap = nullptr;
  a.tile_first_row = nullptr;
  a.scalar_loads = 0;
  HybEll<T> e;
  e.cidx = ecidx; e.vals = evals; e.pitch = pitch; e.K = (int)K; e.rows = rows; e.accumulate = accumulate;
  e.work = nullptr; e.work_count = nullptr; e.work_cap = 0;
  b200sp_status s;
  // one resident CTA fewer than the plain tile kernels: the ELL range loop keeps 32 loads in flight per lane
  if constexpr (sizeof(T) == 4) {
    if (vpl == 4 && u == 1) s = launch_hyb_warp<T, 4, 4, 1>(h, st, a, e, c.ctas_per_sm);
    else if (vpl == 4) s = launch_hyb_warp<T, 3, 4, 2>(h, st, a, e, c.ctas_per_sm);
    else if (u == 1) s = launch_hyb_warp<T, 3, 8, 1>(h, st, a, e, c.ctas_per_sm);
    else s = launch_hyb_warp<T, 2, 8, 2>(h, st, a, e, c.ctas_per_sm);
  } else {
    if (vpl == 4 && u == 1) s = launch_hyb_warp<T, 3, 4, 1>(h, st, a, e, c.ctas_per_sm);
    else if (vpl == 4) s = launch_hyb_warp<T, 2, 4, 2>(h, st, a, e, c.ctas_per_sm);
    else if (u == 1) s = launch_hyb_warp<T, 2, 8, 1>(h, st, a, e, c.ctas_per_sm);
    else s = launch_hyb_warp<T, 2, 8, 2>(h, st, a, e, c.ctas_per_sm);
  }
  if (s == B200SP_OK) *done = 1;
  return s;
}

template b200sp_status spmv_hyb_fused<float>(b200sp_handle, cudaStream_t, i64, i64, i64, i64, const int *, const float *,
                                             i64, const int *, const int *, const float *, const float *, float *, int,
                                             const b200sp_cfg &, int *);
template b200sp_status spmv_hyb_fused<double>(b200sp_handle, cudaStream_t, i64, i64, i64, i64, const int *,
                                              const double *, i64, const int *, const int *, const double *,
                                              const double *, double *, int, const b200sp_cfg &, int *);

}  // namespace b200sp
