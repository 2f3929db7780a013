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
};

// second pass of every COO kernel: leaders add up their carry chain in tile order (spmv_coo.cu)
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
