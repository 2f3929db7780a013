// spmm_csr.cu — CSR x dense block:  Y[rows x k] = (init Y) + A * X[cols x k], X and Y row-major
// ("block SpMV" for block Krylov methods).  Replaces, for csr_matrix x array2d on device_memory,
// BlockSpmvKernel / __spmv_csr_block (cusp/system/cuda/detail/multiply/csr_block_spmv.h:36-222).
// Semantics: host loop cusp/system/detail/sequential/multiply/csr_block_spmv.h:52-77
//     acc[j] = init(Y(i,j));  for jj ascending: acc[j] += Ax[jj] * X(Aj[jj], j);  Y(i,j) = acc[j]
// — per (row, column) the same summation order, so results are bit-identical to it (-fmad=false).
//
// Kernel: a sub-warp of K lanes (K = 8, 16 or 32 >= min(k,32)) owns a row; lane j < k owns column j
// of the block.  The K lanes load K consecutive entries of the row coalesced (ld.global.cs) and
// hand them round with shuffles, so every entry costs one coalesced K*sizeof(T) gather of X
// (ld.global.nc) with 8 gathers in flight per lane.  Blocks wider than 32 columns are processed
// in column chunks of 32.  No shared memory (the reference stages the entries there).
// Algorithmic bytes: (rows+1)*4 + nnz*(4+V) + cols*k*V + rows*k*V.
#include "common.cuh"

namespace b200sp {

// spmv_csr.cu
template <typename T>
b200sp_status spmv_csr(b200sp_handle h, cudaStream_t st, i64 rows, i64 cols, i64 nnz, const int *Ap, const int *Aj,
                       const T *Ax, const T *x, T *y, int accumulate, const b200sp_cfg *cfg, const T *dotv,
                       T *dot_result);

template <typename T>
struct SpmmArgs {
  i64 rows, cols;
  const int *Ap, *Aj;
  const T *Ax;
  const T *X;
  T *Y;
  i64 ldx, ldy;
  int k;  // columns handled by this launch (<= 32), starting at X / Y
  int accumulate;
};

// R rows per sub-warp, interleaved over the CTA (row = first + i*(256/K) + sub-warp): their
// offsets, entries and gathers are issued together, so a thread has R*G independent gathers in
// flight instead of a chain of dependent loads per row (one row per sub-warp ran at 1-1.6 TB/s,
// latency-bound).
template <typename T, int K, int R>
__global__ void __launch_bounds__(256) csr_spmm_kernel(SpmmArgs<T> a) {
  constexpr int G = 4;          // gathers in flight per lane and row
  constexpr int SW = 256 / K;   // sub-warps per CTA
  const int lane = threadIdx.x & (K - 1);
  const i64 first = (i64)blockIdx.x * (SW * R) + threadIdx.x / K;
  const bool col_ok = lane < a.k;
  const int xl = col_ok ? lane : 0;
  int lo[R], hi[R];
  T acc[R];
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const i64 row = first + (i64)i * SW;
    lo[i] = hi[i] = 0;
    if (row < a.rows) {
      lo[i] = ld_ro(a.Ap + row);
      hi[i] = ld_ro(a.Ap + row + 1);
    }
  }
  int span = 0;
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const i64 row = first + (i64)i * SW;
    acc[i] = (row < a.rows && col_ok && a.accumulate) ? a.Y[row * a.ldy + lane] : T(0);
    span = max(span, hi[i] - lo[i]);
  }
  // the longest row among the warp's rows bounds the trip count (shuffles need uniform control)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) span = max(span, __shfl_xor_sync(0xffffffffu, span, o));
  for (int base = 0; base < span; base += K) {
    int c[R];
    T v[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const int jj = lo[i] + base + lane;
      c[i] = 0;
      v[i] = T(0);
      if (jj < hi[i]) {
        c[i] = ld_stream(a.Aj + jj);
        v[i] = ld_stream(a.Ax + jj);
      }
    }
    const int nmax = min(K, span - base);  // warp-uniform
#pragma unroll
    for (int g0 = 0; g0 < K; g0 += G) {
      if (g0 >= nmax) break;
      constexpr int GQ = G < K ? G : K;  // sub-warps narrower than G hold fewer entries per chunk
      T vv[R][GQ], xv[R][GQ];
#pragma unroll
      for (int i = 0; i < R; ++i)
#pragma unroll
        for (int q = 0; q < GQ; ++q) {
          const int cc = __shfl_sync(0xffffffffu, c[i], g0 + q, K);
          vv[i][q] = __shfl_sync(0xffffffffu, v[i], g0 + q, K);
          xv[i][q] = ld_ro(a.X + (i64)cc * a.ldx + xl);  // slots past the row's end gather row 0 of X (valid)
        }
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const int n = min(K, hi[i] - lo[i] - base);  // entries of row i in this chunk (<= 0: none)
#pragma unroll
        for (int q = 0; q < GQ; ++q) {
          pin(xv[i][q]);
          const T t = acc[i] + vv[i][q] * xv[i][q];
          acc[i] = (g0 + q < n) ? t : acc[i];
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const i64 row = first + (i64)i * SW;
    if (row < a.rows && col_ok) a.Y[row * a.ldy + lane] = acc[i];
  }
}

template <typename T, int K, int R>
static b200sp_status launch_spmm(b200sp_handle h, cudaStream_t st, const SpmmArgs<T> &a) {
  const i64 grid = ceil_div(a.rows, (i64)(256 / K) * R);
  csr_spmm_kernel<T, K, R><<<(unsigned)grid, 256, 0, st>>>(a);
  B200SP_LAUNCH_CHECK(h, "csr_spmm_kernel");
  return B200SP_OK;
}

template <typename T>
b200sp_status spmm_csr(b200sp_handle h, cudaStream_t st, i64 rows, i64 cols, i64 nnz, const int *Ap, const int *Aj,
                       const T *Ax, i64 k, const T *X, i64 ldx, T *Y, i64 ldy, int accumulate) {
  B200SP_CHECK_HANDLE(h);
  B200SP_REQUIRE(h, rows >= 0 && cols >= 0 && nnz >= 0 && k >= 0, "spmm csr: negative dimension");
  B200SP_REQUIRE(h, rows < (1ll << 31) && cols < (1ll << 31) && nnz < (1ll << 31), "spmm csr: int32 index range");
  B200SP_REQUIRE(h, ldx >= k && ldy >= k, "spmm csr: leading dimension smaller than the block width");
  if (rows == 0 || k == 0) return B200SP_OK;
  B200SP_REQUIRE(h, Ap && Y, "spmm csr: null pointer");
  B200SP_REQUIRE(h, nnz == 0 || (Aj && Ax && X && cols > 0), "spmm csr: null pointer");
  B200SP_REQUIRE(h, rows * 32 / 256 < (1ll << 31), "spmm csr: too many rows for one launch");
  SpmmArgs<T> a;
  a.rows = rows; a.cols = cols; a.Ap = Ap; a.Aj = Aj; a.Ax = Ax; a.ldx = ldx; a.ldy = ldy; a.accumulate = accumulate;
  if (k == 1 && ldx == 1 && ldy == 1)  // a single contiguous column is a vector (csr_block_spmv.h:198-201)
    return spmv_csr<T>(h, st, rows, cols, nnz, Ap, Aj, Ax, X, Y, accumulate, nullptr, nullptr, nullptr);
  const T dummy = T(0);
  for (i64 c0 = 0; c0 < k; c0 += 32) {  // column chunks of 32
    a.k = (int)((k - c0 < 32) ? k - c0 : 32);
    a.X = X ? X + c0 : &dummy;
    a.Y = Y + c0;
    b200sp_status s;
    // sub-warp width = block width rounded up to a power of two, 4 rows (2 for 32 lanes) per sub-warp
    if (a.k <= 2) s = launch_spmm<T, 2, 4>(h, st, a);
    else if (a.k <= 4) s = launch_spmm<T, 4, 4>(h, st, a);
    else if (a.k <= 8) s = launch_spmm<T, 8, 4>(h, st, a);
    else if (a.k <= 16) s = launch_spmm<T, 16, 4>(h, st, a);
    else s = launch_spmm<T, 32, 2>(h, st, a);
    if (s != B200SP_OK) return s;
  }
  return B200SP_OK;
}

}  // namespace b200sp

extern "C" {
#define DEF(T, sfx)                                                                                         \
  b200sp_status b200sp_spmm_csr_##sfx(b200sp_handle h, b200sp_stream stream, int64_t num_rows, int64_t num_cols, \
                                      int64_t num_entries, const int32_t *row_offsets,                      \
                                      const int32_t *column_indices, const T *values, int64_t block_cols,   \
                                      const T *X, int64_t ldx, T *Y, int64_t ldy, int accumulate) {         \
    return b200sp::spmm_csr<T>(h, (cudaStream_t)stream, num_rows, num_cols, num_entries, row_offsets,       \
                               column_indices, values, block_cols, X, ldx, Y, ldy, accumulate);             \
  }
DEF(float, f32)
DEF(double, f64)
#undef DEF
}
