// spmv_coo_warp.cu — dispatch of K_COO_WARP (coo_warp.cuh): configuration -> template instance.
//
//   cfg.vector_width = entries per lane per load (4: 128-bit, 8: 256-bit loads)
//   cfg.unroll       = units per warp tile (1, 2, 4): a lane keeps vector_width * unroll gathers in flight
//   cfg.block_size   = 256
//   cfg.ctas_per_sm  = 0: one tile per warp; n: persistent grid of n CTAs per SM
//   cfg.stages       = cache-policy variant (x gathers: 0 nc, 1 nc + L1 evict_last, 2 nc + L1 no_allocate;
//                      +4: entry streams with L1::no_allocate + L2 evict-first instead of ld.global.cs)
#include "coo_warp.cuh"

namespace b200sp {

template <typename T>
b200sp_status spmv_coo_warp(b200sp_handle h, cudaStream_t st, CooArgs<T> a, const b200sp_cfg &c) {
  const int vpl = c.vector_width ? c.vector_width : 4, u = c.unroll ? c.unroll : 2;
  const int xpol = c.stages & 3, spol = (c.stages >> 2) & 1;
  if (c.block_size != 0 && c.block_size != 256)
    return set_error(h, B200SP_INVALID_INPUT, "coo warp: unsupported block_size=%d", c.block_size);
  // a lane's vector loads need 16- / 32-byte aligned array bases
  const uintptr_t need_idx = (uintptr_t)(4 * vpl) - 1, need_val = (uintptr_t)(sizeof(T) * vpl > 32 ? 32 : sizeof(T) * vpl) - 1;
  if (((uintptr_t)a.Ai & need_idx) || ((uintptr_t)a.Aj & need_idx) || ((uintptr_t)a.Ax & need_val))
    return set_error(h, B200SP_INVALID_INPUT, "coo warp: arrays not aligned for %d-entry vector loads", vpl);
#define CASE(V, UU, MINB, X, S)                                \
  if (vpl == V && u == UU && xpol == X && spol == S)           \
    return launch_coo_warp<T, 256, MINB, V, UU, X, S, false>(h, st, a, c.ctas_per_sm, nullptr, 0, 0);
#define CASES(V, UU, MINB) \
  CASE(V, UU, MINB, 0, 0) CASE(V, UU, MINB, 1, 0) CASE(V, UU, MINB, 2, 0) CASE(V, UU, MINB, 0, 1) CASE(V, UU, MINB, 1, 1) CASE(V, UU, MINB, 2, 1)
  if constexpr (sizeof(T) == 4) {
    CASES(4, 1, 6) CASES(4, 2, 4) CASES(4, 4, 3) CASES(8, 1, 4) CASES(8, 2, 3) CASES(8, 4, 2)
  } else {
    CASES(4, 1, 4) CASES(4, 2, 3) CASES(4, 4, 2) CASES(8, 1, 3) CASES(8, 2, 2)
  }
#undef CASES
#undef CASE
  return set_error(h, B200SP_INVALID_INPUT, "coo warp: unsupported vector_width=%d unroll=%d stages=%d", vpl, u,
                   c.stages);
}

template b200sp_status spmv_coo_warp<float>(b200sp_handle, cudaStream_t, CooArgs<float>, const b200sp_cfg &);
template b200sp_status spmv_coo_warp<double>(b200sp_handle, cudaStream_t, CooArgs<double>, const b200sp_cfg &);

}  // namespace b200sp
