// spmv_coo_warp.cu — dispatch of K_COO_WARP (coo_warp.cuh): configuration -> template instance.
//
//   cfg.vector_width = entries per lane per load (4: 128-bit, 8: 256-bit loads)
//   cfg.unroll       = units per warp tile (1, 2, 4): a lane keeps vector_width * unroll gathers in flight
//   cfg.block_size   = 256
//   cfg.ctas_per_sm  = 0: one tile per warp; n: persistent grid of n CTAs per SM
//
// Cache policies are fixed by measurement (R-MAT scale 22 / 24 sweep, profiles/r03_coo_probe.md): x gathers
// ld.global.nc with the default L1 policy (L1::evict_last -3 %, L1::no_allocate -40 %: the L1 hits on hot x sectors
// matter); entry streams ld.global.cs here, and L1::no_allocate + L2 evict-first in the plan executor, where the
// table leaves little L1 (0.89 ms against 0.99 ms).
#include <stdlib.h>

#include "coo_warp.cuh"

namespace b200sp {

template <typename T>
b200sp_status spmv_coo_warp(b200sp_handle h, cudaStream_t st, CooArgs<T> a, const b200sp_cfg &c) {
  const int vpl = c.vector_width ? c.vector_width : 8, u = c.unroll ? c.unroll : 1;
  if (c.block_size != 0 && c.block_size != 256)
    return set_error(h, B200SP_INVALID_INPUT, "coo warp: unsupported block_size=%d", c.block_size);
  // a lane's vector loads need 16- / 32-byte aligned array bases
  const uintptr_t need_idx = (uintptr_t)(4 * vpl) - 1, need_val = (uintptr_t)(sizeof(T) * vpl > 32 ? 32 : sizeof(T) * vpl) - 1;
  if (((uintptr_t)a.Ai & need_idx) || ((uintptr_t)a.Aj & need_idx) || ((uintptr_t)a.Ax & need_val))
    return set_error(h, B200SP_INVALID_INPUT, "coo warp: arrays not aligned for %d-entry vector loads", vpl);
  // y += A x with 128-bit tiles (the shape of one-entry-per-row HYB tails): y of the row ends is fetched with the x
  // gathers (PREY, coo_warp.cuh) — B200SP_COO_PREY=0 keeps the read-modify-write after the scan
  const char *pe = getenv("B200SP_COO_PREY");
  const bool prey = a.accumulate && vpl == 4 && !(pe && pe[0] == '0');
#define CASEP(V, UU, MINB)    \
  if (prey && vpl == V && u == UU) \
    return launch_coo_warp<T, 256, MINB, V, UU, 0, 0, false, SpmvOps<T, 0, 0>, true>(h, st, a, c.ctas_per_sm, nullptr, 0, 0);
  if constexpr (sizeof(T) == 4) {
    CASEP(4, 1, 6) CASEP(4, 2, 4)
  } else {
    CASEP(4, 1, 4) CASEP(4, 2, 3)
  }
#undef CASEP
#define CASE(V, UU, MINB) \
  if (vpl == V && u == UU) return launch_coo_warp<T, 256, MINB, V, UU, 0, 0, false>(h, st, a, c.ctas_per_sm, nullptr, 0, 0);
  // (vector_width 1 — a lane takes every 32nd entry of a unit, scalar coalesced loads, 4 or 8 units per tile, so that
  // the x and y accesses of one-entry-per-row tails coalesce — was measured on the stencil's K = 6 HYB: 0.2465 against
  // 0.2426 ms for the 128-bit shape in fp32, 0.370 against 0.351 in fp64; not instantiated)
  if constexpr (sizeof(T) == 4) {
    CASE(4, 1, 6) CASE(4, 2, 4) CASE(4, 4, 3) CASE(8, 1, 4) CASE(8, 2, 3) CASE(8, 4, 2)
  } else {
    CASE(4, 1, 4) CASE(4, 2, 3) CASE(4, 4, 2) CASE(8, 1, 3) CASE(8, 2, 2)
  }
#undef CASE
  return set_error(h, B200SP_INVALID_INPUT, "coo warp: unsupported vector_width=%d unroll=%d", vpl, u);
}

template b200sp_status spmv_coo_warp<float>(b200sp_handle, cudaStream_t, CooArgs<float>, const b200sp_cfg &);
template b200sp_status spmv_coo_warp<double>(b200sp_handle, cudaStream_t, CooArgs<double>, const b200sp_cfg &);

}  // namespace b200sp
