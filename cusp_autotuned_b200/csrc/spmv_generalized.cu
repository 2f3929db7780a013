// spmv_generalized.cu — the generalized product on the device:
//     y[i] = reduce(initialize(y[i]), combine(a_ij, x_j) ...)       for the stored entries of row i
// cusp::multiply(A, x, y, initialize, combine, reduce) and cusp::generalized_spmv
// (cusp/multiply.h:163-280; cusp/system/detail/generic/multiply/generalized_spmv.h:61-303; device kernels templated
// on the functor triple: cusp/system/cuda/detail/multiply/csr_vector_spmv.h:66-161, ell_spmv.h:47-93,
// dia_spmv.h:66-126; tested by testing/generalized_spmv.cu).  A C ABI cannot take C++ functor objects, so the
// triple is named by code (b200sp_functors): initialize in {constant(c), identity}, combine in {multiplies, plus,
// minimum, maximum, project2nd}, reduce in {plus, minimum, maximum} — the arithmetic and the (min,+) / (max,x) /
// (min,max) ... semirings.  (multiplies, plus) with constant(0) | identity is routed to the tuned kernels of
// b200sp_spmv; everything else runs here:
//   * initialize is applied to all of y first (memset / fill / nothing), like the COO and ELL host loops do
//     (sequential/multiply/coo_spmv.h:45-49, ell_spmv.h:52-56);
//   * ELL / ELL-R / DIA: one thread per row, slots / diagonals in storage order — the reference's order of
//     operations per row, bit-identical to the host loop for every functor pair;
//   * CSR: one thread per row (mean row length <= 8: the host loop's order, bit-identical) or a sub-warp per row
//     with partial results combined by a shuffle tree (regrouped: exact for min / max, tolerance-level for plus);
//   * COO (and the tail of HYB): K_COO_WARP's warp tiles with the segmented scan and the carry fix-up running on
//     `reduce` instead of `+` (coo_warp.cuh is templated on the pair).
#include "coo_warp.cuh"

namespace b200sp {

// Every kernel below keeps 4 - 8 independent (entry, x) load pairs in flight per thread — all entry loads of a batch,
// then all gathers, then the reductions in storage order (the first version walked one entry at a time through three
// dependent loads: 0.41 - 0.53 of the copy rate on poisson7pt 256^3) — and applies `initialize` itself (init_const:
// y[r] starts from init_value, no fill pass and no read of y; otherwise from the caller's y[r]).
constexpr int GEN_G = 4;   // CSR: per lane of a row
// ELL / DIA: per row — 8 in fp32 (a 7-point stencil row is one batch), 4 in fp64 (measured: ELL 0.426 against 0.487 ms,
// DIA 0.207 against 0.244 ms with 8)
template <typename T>
struct GenGr {
  static constexpr int value = sizeof(T) == 4 ? 8 : 4;
};

template <typename T, typename Ops>
__global__ void __launch_bounds__(256) gen_ell_kernel(i64 rows, i64 cols, i64 pitch, int K, const int *cidx,
                                                      const T *vals, const int *row_lengths, const T *x, T *y,
                                                      int init_const, T init_value) {
  constexpr int GR = GenGr<T>::value;
  const i64 r = (i64)blockIdx.x * 256 + threadIdx.x;
  if (r >= rows) return;
  T acc = init_const ? init_value : y[r];
  const int kmax = row_lengths ? min(K, row_lengths[r]) : K;
  for (int n0 = 0; n0 < kmax; n0 += GR) {
    int c[GR];
    T v[GR], xv[GR];
#pragma unroll
    for (int q = 0; q < GR; ++q) {
      const bool in = n0 + q < kmax;
      const i64 so = (i64)(in ? n0 + q : n0) * pitch + r;
      c[q] = ld_stream(cidx + so);
      v[q] = ld_stream(vals + so);
      if (!in || c[q] >= cols) c[q] = -1;
    }
#pragma unroll
    for (int q = 0; q < GR; ++q) {
      pin(c[q]);
      xv[q] = ld_ro(x + max(c[q], 0));
    }
#pragma unroll
    for (int q = 0; q < GR; ++q) {
      pin(xv[q]);
      if (c[q] >= 0) acc = Ops::reduce(acc, Ops::combine(v[q], xv[q]));
    }
  }
  y[r] = acc;
}

template <typename T, typename Ops>
__global__ void __launch_bounds__(256) gen_dia_kernel(i64 rows, i64 cols, i64 pitch, int ndiag, const int *offs,
                                                      const T *vals, const T *x, T *y, int init_const, T init_value) {
  // DIA: x is read at consecutive addresses, the loop is short and the plain one-diagonal-at-a-time form is the
  // fastest in fp32 (0.187 against 0.201 - 0.262 ms for batches of 4 / 8 or offsets staged in shared memory); fp64
  // gains from batches of 4 (0.207 against 0.260 ms)
  constexpr int GR = sizeof(T) == 4 ? 1 : 4;
  const i64 r = (i64)blockIdx.x * 256 + threadIdx.x;
  if (r >= rows) return;
  T acc = init_const ? init_value : y[r];
  for (int d0 = 0; d0 < ndiag; d0 += GR) {
    i64 c[GR];
    T v[GR], xv[GR];
#pragma unroll
    for (int q = 0; q < GR; ++q) {
      const bool in = d0 + q < ndiag;
      const int d = in ? d0 + q : d0;
      c[q] = r + (i64)ld_ro(offs + d);
      if (!in || (unsigned long long)c[q] >= (unsigned long long)cols) c[q] = -1;
      v[q] = ld_stream(vals + (i64)d * pitch + r);
    }
#pragma unroll
    for (int q = 0; q < GR; ++q) xv[q] = ld_ro(x + (c[q] < 0 ? 0 : c[q]));
#pragma unroll
    for (int q = 0; q < GR; ++q) {
      pin(xv[q]);
      if (c[q] >= 0) acc = Ops::reduce(acc, Ops::combine(v[q], xv[q]));
    }
  }
  y[r] = acc;
}

// TPR lanes per row (power of two); TPR == 1 keeps the host loop's order
template <typename T, typename Ops, int TPR>
__global__ void __launch_bounds__(256) gen_csr_kernel(i64 rows, i64 cols, const int *Ap, const int *Aj, const T *Ax,
                                                      const T *x, T *y, int init_const, T init_value) {
  const i64 r = ((i64)blockIdx.x * 256 + threadIdx.x) / TPR;
  const int sub = threadIdx.x % TPR;
  const bool live = r < rows;
  const int lo = live ? ld_ro(Ap + r) : 0, hi = live ? ld_ro(Ap + r + 1) : 0;
  const T start = (live && !init_const) ? y[r] : init_value;  // what initialize leaves in y[r]
  T acc = (TPR == 1) ? start : Ops::identity();
  for (int k0 = lo + sub; k0 < hi; k0 += GEN_G * TPR) {
    unsigned c[GEN_G];
    T v[GEN_G], xv[GEN_G];
#pragma unroll
    for (int q = 0; q < GEN_G; ++q) {
      const int k = min(k0 + q * TPR, hi - 1);
      c[q] = (unsigned)ld_stream(Aj + k);
      v[q] = ld_stream(Ax + k);
    }
#pragma unroll
    for (int q = 0; q < GEN_G; ++q) xv[q] = ld_ro(x + min(c[q], (unsigned)cols - 1));
#pragma unroll
    for (int q = 0; q < GEN_G; ++q) {
      pin(xv[q]);
      if (k0 + q * TPR < hi) acc = Ops::reduce(acc, Ops::combine(v[q], xv[q]));
    }
  }
  if (TPR > 1) {
#pragma unroll
    for (int o = TPR / 2; o > 0; o >>= 1) acc = Ops::reduce(acc, __shfl_down_sync(0xffffffffu, acc, o, TPR));
    if (live && sub == 0) y[r] = Ops::reduce(start, acc);
  } else if (live) {
    y[r] = acc;
  }
}

template <typename T>
__global__ void gen_fill_kernel(i64 n, T v, T *y) {
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) y[i] = v;
}

template <typename T, typename Ops>
static b200sp_status gen_coo(b200sp_handle h, cudaStream_t st, i64 rows, i64 cols, i64 nnz, const int *Ai, const int *Aj,
                             const T *Ax, const T *x, T *y) {
  if (nnz == 0) return B200SP_OK;
  B200SP_REQUIRE(h, Ai && Aj && Ax && cols > 0, "generalized coo: null pointer");
  CooArgs<T> a;
  a.rows = rows; a.cols = cols; a.nnz = nnz; a.Ai = Ai; a.Aj = Aj; a.Ax = Ax; a.x = x; a.y = y;
  a.accumulate = 1;  // initialize was applied to every row already: y[row] = reduce(y[row], row total)
  a.carry = nullptr; a.Ap = nullptr; a.tile_first_row = nullptr;
  a.scalar_loads = ((((uintptr_t)Ai | (uintptr_t)Aj) & 15) || ((uintptr_t)Ax & (4 * sizeof(T) > 32 ? 31 : 4 * sizeof(T) - 1))) ? 1 : 0;
  return launch_coo_warp<T, 256, 3, 4, 2, 0, 0, false, Ops>(h, st, a, 0, nullptr, 0, 0);
}

template <typename T, typename Ops>
static b200sp_status gen_run(b200sp_handle h, cudaStream_t st, const b200sp_matrix *A, const T *x, T *y, int init_const,
                             T init_value) {
  const i64 rows = A->num_rows, cols = A->num_cols;
  const T *vals = reinterpret_cast<const T *>(A->values);
  const unsigned grid_rows = (unsigned)ceil_div(rows, 256);
  switch (A->format) {
    case B200SP_FMT_CSR: {
      if (A->num_entries == 0) return B200SP_OK;
      B200SP_REQUIRE(h, A->row_offsets && A->column_indices && vals && cols > 0, "generalized csr: null pointer");
      const double mean = (double)A->num_entries / (double)rows;
      if (mean <= 8.0) {
        gen_csr_kernel<T, Ops, 1><<<grid_rows, 256, 0, st>>>(rows, cols, A->row_offsets, A->column_indices, vals, x, y,
                                                             init_const, init_value);
      } else if (mean <= 48.0) {
        gen_csr_kernel<T, Ops, 8><<<(unsigned)ceil_div(rows * 8, 256), 256, 0, st>>>(
            rows, cols, A->row_offsets, A->column_indices, vals, x, y, init_const, init_value);
      } else {
        gen_csr_kernel<T, Ops, 32><<<(unsigned)ceil_div(rows * 32, 256), 256, 0, st>>>(
            rows, cols, A->row_offsets, A->column_indices, vals, x, y, init_const, init_value);
      }
      B200SP_LAUNCH_CHECK(h, "gen_csr_kernel");
      return B200SP_OK;
    }
    case B200SP_FMT_ELL:
    case B200SP_FMT_ELLR:
    case B200SP_FMT_HYB: {
      if (A->num_cols_per_row > 0) {
        B200SP_REQUIRE(h, A->column_indices && vals && A->pitch >= rows, "generalized ell: bad arrays");
        gen_ell_kernel<T, Ops><<<grid_rows, 256, 0, st>>>(rows, cols, A->pitch, (int)A->num_cols_per_row, A->column_indices,
                                                          vals, A->format == B200SP_FMT_ELLR ? A->row_offsets : nullptr, x, y,
                                                          init_const, init_value);
        B200SP_LAUNCH_CHECK(h, "gen_ell_kernel");
      }
      if (A->format != B200SP_FMT_HYB) return B200SP_OK;
      // the COO part continues from what the ELL part left in y (hyb_spmv.h:52-56: initialize = identity)
      return gen_coo<T, Ops>(h, st, rows, cols, A->coo_num_entries, A->coo_row_indices, A->coo_column_indices,
                             reinterpret_cast<const T *>(A->coo_values), x, y);
    }
    case B200SP_FMT_DIA:
      if (A->num_cols_per_row == 0) return B200SP_OK;
      B200SP_REQUIRE(h, A->diagonal_offsets && vals && A->pitch >= rows, "generalized dia: bad arrays");
      gen_dia_kernel<T, Ops><<<grid_rows, 256, 0, st>>>(rows, cols, A->pitch, (int)A->num_cols_per_row, A->diagonal_offsets,
                                                        vals, x, y, init_const, init_value);
      B200SP_LAUNCH_CHECK(h, "gen_dia_kernel");
      return B200SP_OK;
    case B200SP_FMT_COO:
      return gen_coo<T, Ops>(h, st, rows, cols, A->num_entries, A->row_indices, A->column_indices, vals, x, y);
  }
  return set_error(h, B200SP_INVALID_INPUT, "generalized spmv: unknown format %d", (int)A->format);
}

template <typename T>
static b200sp_status gen_dispatch(b200sp_handle h, cudaStream_t st, const b200sp_matrix *A, const T *x, T *y,
                                  const b200sp_functors *f) {
  const i64 rows = A->num_rows;
  if (rows == 0) return B200SP_OK;
  B200SP_REQUIRE(h, y != nullptr && (x != nullptr || A->num_cols == 0), "generalized spmv: null vector");
  // initialize(y[i]) for every row, stored or not.  A row kernel that visits every row applies a constant itself
  // (no fill pass, no read of y); COO — whose tiles only see rows with entries — and matrices without a row kernel
  // to run get the fill.
  if (f->initialize != B200SP_INIT_CONSTANT && f->initialize != B200SP_INIT_IDENTITY)
    return set_error(h, B200SP_INVALID_INPUT, "generalized spmv: unknown initialize code %d", f->initialize);
  const bool constant = f->initialize == B200SP_INIT_CONSTANT;
  const bool row_kernel = (A->format == B200SP_FMT_CSR && A->num_entries > 0) ||
                          ((A->format == B200SP_FMT_ELL || A->format == B200SP_FMT_ELLR || A->format == B200SP_FMT_HYB ||
                            A->format == B200SP_FMT_DIA) && A->num_cols_per_row > 0);
  const int init_const = (constant && row_kernel) ? 1 : 0;
  const T init_value = (T)f->init_value;
  if (constant && !row_kernel) {
    if (f->init_value == 0.0) {
      B200SP_CUDA(h, cudaMemsetAsync(y, 0, (size_t)rows * sizeof(T), st));
    } else {
      const i64 g = ceil_div(rows, 256) < (i64)h->num_sms * 8 ? ceil_div(rows, 256) : (i64)h->num_sms * 8;
      gen_fill_kernel<T><<<(unsigned)g, 256, 0, st>>>(rows, init_value, y);
      B200SP_LAUNCH_CHECK(h, "gen_fill_kernel");
    }
  }
#define CASE(C, R) \
  if (f->combine == C && f->reduce == R) return gen_run<T, SpmvOps<T, C, R>>(h, st, A, x, y, init_const, init_value);
  CASE(0, 0) CASE(1, 0) CASE(2, 0) CASE(3, 0) CASE(4, 0)
  CASE(0, 1) CASE(1, 1) CASE(2, 1) CASE(3, 1) CASE(4, 1)
  CASE(0, 2) CASE(1, 2) CASE(2, 2) CASE(3, 2) CASE(4, 2)
#undef CASE
  return set_error(h, B200SP_NOT_IMPLEMENTED, "generalized spmv: unsupported functor codes combine=%d reduce=%d", f->combine,
                   f->reduce);
}

}  // namespace b200sp

extern "C" b200sp_status b200sp_spmv_generalized(b200sp_handle h, b200sp_stream stream, const b200sp_matrix *A,
                                                 const void *x, void *y, const b200sp_functors *functors) {
  B200SP_CHECK_HANDLE(h);
  B200SP_REQUIRE(h, A != nullptr && functors != nullptr, "generalized spmv: null argument");
  B200SP_REQUIRE(h, A->num_rows >= 0 && A->num_cols >= 0 && A->num_rows < (1ll << 31) && A->num_cols < (1ll << 31),
                 "generalized spmv: dimensions outside the int32 index range");
  // the default pair with the two initializers the tuned kernels know: the fast path
  if (functors->combine == B200SP_COMBINE_MULTIPLIES && functors->reduce == B200SP_REDUCE_PLUS &&
      (functors->initialize == B200SP_INIT_IDENTITY ||
       (functors->initialize == B200SP_INIT_CONSTANT && functors->init_value == 0.0)))
    return b200sp_spmv(h, stream, A, x, y, functors->initialize == B200SP_INIT_IDENTITY ? 1 : 0, nullptr);
  if (A->dtype == B200SP_F32)
    return b200sp::gen_dispatch<float>(h, (cudaStream_t)stream, A, (const float *)x, (float *)y, functors);
  if (A->dtype == B200SP_F64)
    return b200sp::gen_dispatch<double>(h, (cudaStream_t)stream, A, (const double *)x, (double *)y, functors);
  return b200sp::set_error(h, B200SP_INVALID_INPUT, "generalized spmv: unknown dtype %d", (int)A->dtype);
}
