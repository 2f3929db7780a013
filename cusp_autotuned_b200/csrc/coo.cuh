// coo.cuh — types shared by the COO kernels (spmv_coo.cu: K_COO_SEGSCAN / K_COO_RING,
// spmv_coo_warp.cu: K_COO_WARP and the hot-column plan executor) and the carry fix-up.
#pragma once

#include "common.cuh"

namespace b200sp {

// What a tile of consecutive entries leaves for its neighbours: the row that continues from
// the previous tile and ends here (head), and the row left open at the end of the tile (tail).
template <typename T>
struct CooCarry {
  int head_row;  // row continued from the previous tile that ends here, or -1
  int tail_row;  // row left open at the end of this tile, or -1
  int leader;    // tail_row began inside this tile
  int pad;
  T head_val;
  T tail_val;
};

template <typename T>
struct CooArgs {
  i64 rows, cols, nnz;
  const int *Ai;
  const int *Aj;
  const T *Ax;
  const T *x;
  T *y;
  int accumulate;
  CooCarry<T> *carry;
  // CSR source (K_CSR_BALANCED): row indices are rebuilt per tile from row_offsets
  const int *Ap;
  const int *tile_first_row;  // row that contains entry t*TILE, for every tile t
  int scalar_loads;           // K_COO_WARP: array bases not aligned for vector loads -> guarded scalar loads
};

// ---- the (combine, reduce) pair of the reference's generalized product ---------------------------
// y[i] = reduce(init(y[i]), combine(a_ij, x_j) ...)   cusp/system/detail/generic/multiply/generalized_spmv.h:61-303,
// cusp/multiply.h:163-195.  The C ABI names the functors by code (b200sp_functors); every kernel that is templated
// on Ops runs the default product with SpmvOps<T, 0, 0> = (multiplies, plus) and the same instructions as before.
template <typename T>
__device__ __forceinline__ T ops_inf();
template <>
__device__ __forceinline__ float ops_inf<float>() { return __int_as_float(0x7f800000); }
template <>
__device__ __forceinline__ double ops_inf<double>() { return __longlong_as_double(0x7ff0000000000000ll); }

template <typename T, int COMBINE, int REDUCE>
struct SpmvOps {
  // COMBINE: 0 multiplies, 1 plus, 2 minimum, 3 maximum, 4 project2nd (the x operand)
  static __device__ __forceinline__ T combine(T a, T x) {
    if (COMBINE == 0) return a * x;
    if (COMBINE == 1) return a + x;
    if (COMBINE == 2) return x < a ? x : a;  // thrust::minimum: rhs < lhs ? rhs : lhs
    if (COMBINE == 3) return a < x ? x : a;  // thrust::maximum: lhs < rhs ? rhs : lhs
    return x;
  }
  // REDUCE: 0 plus, 1 minimum, 2 maximum
  static __device__ __forceinline__ T reduce(T a, T b) {
    if (REDUCE == 0) return a + b;
    if (REDUCE == 1) return b < a ? b : a;
    return a < b ? b : a;
  }
  // neutral element of reduce (what an empty partial contributes)
  static __device__ __forceinline__ T identity() {
    if (REDUCE == 0) return T(0);
    if (REDUCE == 1) return ops_inf<T>();
    return -ops_inf<T>();
  }
};

// Second pass of every COO kernel: one thread per tile.  A tile whose open tail row began inside it ("leader")
// owns that row's total: its tail, then the tails of the following tiles that lie entirely inside the row, then the
// head of the tile in which the row ends.  Most chains are one record long and are finished by the leader's own
// lane; a chain that goes on (a hub row spanning many tiles) is walked by the whole warp, 32 records per step, with
// a fixed shuffle tree per step — the order of the reductions depends only on the tile shape, never on timing.
template <typename T, typename Ops>
__global__ void coo_fixup_kernel(i64 num_tiles, const CooCarry<T> *carry, T *y, int accumulate) {
  constexpr unsigned FULL = 0xffffffffu;
  const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  bool leader = false, walk = false;
  int row = -1;
  T total = Ops::identity();
  if (t < num_tiles) {
    const CooCarry<T> me = carry[t];
    if (me.tail_row >= 0 && me.leader) {
      leader = true;
      row = me.tail_row;
      total = me.tail_val;
      if (t + 1 < num_tiles) {
        const CooCarry<T> nx = carry[t + 1];
        if (nx.head_row == row) {
          total = Ops::reduce(total, nx.head_val);
        } else if (nx.tail_row == row && !nx.leader) {
          total = Ops::reduce(total, nx.tail_val);
          walk = true;
        }
      }
    }
  }
  unsigned walkers = __ballot_sync(FULL, walk);
  while (walkers) {
    const int src = __ffs(walkers) - 1;
    walkers &= walkers - 1;
    const i64 t0 = __shfl_sync(FULL, t, src);
    const int wrow = __shfl_sync(FULL, row, src);
    T wtot = __shfl_sync(FULL, total, src);
    for (i64 base = t0 + 2;; base += 32) {
      const i64 u = base + lane;
      int cls = 2;  // 0: tile inside the row, 1: the row ends in this tile, 2: not part of the chain
      T val = Ops::identity();
      if (u < num_tiles) {
        const CooCarry<T> nx = carry[u];
        if (nx.head_row == wrow) {
          cls = 1;
          val = nx.head_val;
        } else if (nx.tail_row == wrow && !nx.leader) {
          cls = 0;
          val = nx.tail_val;
        }
      }
      const unsigned stop = __ballot_sync(FULL, cls != 0);
      const int first = stop ? __ffs(stop) - 1 : 32;
      T contrib = (lane < first || (lane == first && cls == 1)) ? val : Ops::identity();
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) contrib = Ops::reduce(contrib, __shfl_xor_sync(FULL, contrib, o));
      wtot = Ops::reduce(wtot, contrib);
      if (stop) break;
    }
    if (lane == src) total = wtot;
  }
  if (leader) y[row] = accumulate ? Ops::reduce(y[row], total) : total;  // !accumulate: y[row] holds the memset zero
}

// the default (multiplies, plus) instance, launched by every COO path (spmv_coo.cu)
template <typename T>
b200sp_status launch_coo_fixup(b200sp_handle h, cudaStream_t st, i64 tiles, const CooCarry<T> *carry, T *y,
                               int accumulate);

// K_COO_WARP (spmv_coo_warp.cu).  hot_cols / hot: the plan executor's table (nullptr / 0 otherwise);
// then a.Aj is the plan's remapped column array.
template <typename T>
b200sp_status spmv_coo_warp(b200sp_handle h, cudaStream_t st, CooArgs<T> a, const b200sp_cfg &c);
template <typename T>
b200sp_status spmv_coo_hot(b200sp_handle h, cudaStream_t st, CooArgs<T> a, const b200sp_cfg &c,
                           const int *hot_cols, int hot, int capacity);

}  // namespace b200sp
