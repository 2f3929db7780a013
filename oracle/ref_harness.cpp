// ref_harness.cpp — thin C ABI around the REFERENCE's own host SpMV loops.
// TEST INFRASTRUCTURE ONLY (checker + CPU baseline), never linked by the product.
//
// The five headers cusp/system/detail/sequential/multiply/{csr,coo,dia,ell,hyb}_spmv.h
// are included UNMODIFIED from /root/reference at build time (oracle/Makefile,
// output oracle/_ref/libcuspref.so); nothing from the reference is copied into
// this repository.  The rest of the reference does not build here (needs KTT
// v2.1 and Thrust-1.x internals), so the templates are instantiated on
// duck-typed views exposing exactly the members they touch.
//
// Multi-threading: the reference's host path is single-threaded for every format
// but CSR-with-OpenMP.  For "all host threads" numbers the harness splits the
// rows into contiguous blocks and runs the reference loop on a view of each
// block in its own std::thread; per-row arithmetic is the reference's own.
#include <cuda_runtime.h>  // only so that thrust's host headers find their config
#include <thrust/functional.h>
#include <thrust/system/cpp/execution_policy.h>

#include <cusp/system/detail/sequential/multiply/coo_spmv.h>
#include <cusp/system/detail/sequential/multiply/csr_spmv.h>
#include <cusp/system/detail/sequential/multiply/dia_spmv.h>
#include <cusp/system/detail/sequential/multiply/ell_spmv.h>
#include <cusp/system/detail/sequential/multiply/hyb_spmv.h>

#include <cstdint>
#include <thread>
#include <vector>

typedef int64_t i64;

template <typename T>
struct Vec {
  typedef T value_type;
  T *p;
  T &operator[](size_t i) { return p[i]; }
  const T &operator[](size_t i) const { return p[i]; }
};
template <typename T>
struct CVec {
  typedef T value_type;
  const T *p;
  const T &operator[](size_t i) const { return p[i]; }
};
// column-major 2-D view with pitch and a row shift (row-block views)
template <typename T>
struct Arr2 {
  const T *p;
  size_t pitch, num_cols, row0;
  const T &operator()(size_t i, size_t j) const { return p[j * pitch + row0 + i]; }
};

template <typename T>
struct Csr {
  typedef int index_type;
  typedef T value_type;
  size_t num_rows, num_cols, num_entries;
  CVec<int> row_offsets, column_indices;
  CVec<T> values;
};
template <typename T>
struct Coo {
  typedef int index_type;
  typedef T value_type;
  size_t num_rows, num_cols, num_entries;
  CVec<int> row_indices, column_indices;
  CVec<T> values;
};
template <typename T>
struct Dia {
  typedef int index_type;
  typedef T value_type;
  size_t num_rows, num_cols, num_entries;
  CVec<int> diagonal_offsets;
  Arr2<T> values;
};
template <typename T>
struct Ell {
  typedef int index_type;
  typedef T value_type;
  static const int invalid_index = -1;
  size_t num_rows, num_cols, num_entries;
  Arr2<int> column_indices;
  Arr2<T> values;
};
template <typename T>
struct Hyb {
  typedef int index_type;
  typedef T value_type;
  size_t num_rows, num_cols, num_entries;
  Ell<T> ell;
  Coo<T> coo;
};

template <typename T>
struct Zero {
  T operator()(const T &) const { return T(0); }
};
template <typename T>
struct Ident {
  T operator()(const T &v) const { return v; }
};

namespace seq = cusp::system::detail::sequential;

template <typename T, typename M, typename Fmt>
static void run(const M &A, const T *x, T *y, int acc, Fmt fmt) {
  thrust::cpp::tag exec;
  CVec<T> xv{x};
  Vec<T> yv{y};
  if (acc)
    seq::multiply(exec, A, xv, yv, Ident<T>(), thrust::multiplies<T>(), thrust::plus<T>(), fmt,
                  cusp::array1d_format(), cusp::array1d_format());
  else
    seq::multiply(exec, A, xv, yv, Zero<T>(), thrust::multiplies<T>(), thrust::plus<T>(), fmt,
                  cusp::array1d_format(), cusp::array1d_format());
}

template <typename F>
static void parallel_blocks(i64 rows, int nthreads, F f) {
  if (nthreads <= 1 || rows < 2 * nthreads) {
    f(0, rows);
    return;
  }
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++) {
    const i64 r0 = rows * t / nthreads, r1 = rows * (t + 1) / nthreads;
    th.emplace_back([=]() { f(r0, r1); });
  }
  for (auto &t : th) t.join();
}

extern "C" {
#define DEF(T, sfx)                                                                                           \
  void ref_spmv_csr_##sfx(i64 rows, i64 cols, i64 nnz, const int *Ap, const int *Aj, const T *Ax, const T *x, \
                          T *y, int acc, int nthreads) {                                                      \
    parallel_blocks(rows, nthreads, [=](i64 r0, i64 r1) {                                                     \
      Csr<T> A{(size_t)(r1 - r0), (size_t)cols, (size_t)nnz, {Ap + r0}, {Aj}, {Ax}};                          \
      run<T>(A, x, y + r0, acc, cusp::csr_format());                                                          \
    });                                                                                                       \
  }                                                                                                           \
  void ref_spmv_coo_##sfx(i64 rows, i64 cols, i64 nnz, const int *Ai, const int *Aj, const T *Ax, const T *x, \
                          T *y, int acc, int nthreads) {                                                      \
    (void)nthreads; /* entry-ordered loop: run single-threaded as the reference does */                       \
    Coo<T> A{(size_t)rows, (size_t)cols, (size_t)nnz, {Ai}, {Aj}, {Ax}};                                      \
    run<T>(A, x, y, acc, cusp::coo_format());                                                                 \
  }                                                                                                           \
  void ref_spmv_dia_##sfx(i64 rows, i64 cols, i64 nd, i64 pitch, const int *offs, const T *vals, const T *x,  \
                          T *y, int acc, int nthreads) {                                                      \
    parallel_blocks(rows, nthreads, [=](i64 r0, i64 r1) {                                                     \
      /* rows [r0,r1) as a matrix of its own: offsets shift by r0 */                                          \
      std::vector<int> o((size_t)nd);                                                                         \
      for (i64 d = 0; d < nd; d++) o[(size_t)d] = (int)(offs[d] + r0);                                        \
      Dia<T> A{(size_t)(r1 - r0), (size_t)cols, 0, {o.data()}, {vals, (size_t)pitch, (size_t)nd, (size_t)r0}}; \
      run<T>(A, x, y + r0, acc, cusp::dia_format());                                                          \
    });                                                                                                       \
  }                                                                                                           \
  void ref_spmv_ell_##sfx(i64 rows, i64 cols, i64 K, i64 pitch, const int *cidx, const T *vals, const T *x,   \
                          T *y, int acc, int nthreads) {                                                      \
    parallel_blocks(rows, nthreads, [=](i64 r0, i64 r1) {                                                     \
      Ell<T> A{(size_t)(r1 - r0), (size_t)cols, 0, {cidx, (size_t)pitch, (size_t)K, (size_t)r0},              \
               {vals, (size_t)pitch, (size_t)K, (size_t)r0}};                                                 \
      run<T>(A, x, y + r0, acc, cusp::ell_format());                                                          \
    });                                                                                                       \
  }                                                                                                           \
  void ref_spmv_hyb_##sfx(i64 rows, i64 cols, i64 K, i64 pitch, const int *ecidx, const T *evals, i64 cnnz,   \
                          const int *ci, const int *cj, const T *cv, const T *x, T *y, int acc) {             \
    Hyb<T> A{(size_t)rows, (size_t)cols, 0,                                                                   \
             Ell<T>{(size_t)rows, (size_t)cols, 0, {ecidx, (size_t)pitch, (size_t)K, 0},                      \
                    {evals, (size_t)pitch, (size_t)K, 0}},                                                    \
             Coo<T>{(size_t)rows, (size_t)cols, (size_t)cnnz, {ci}, {cj}, {cv}}};                             \
    run<T>(A, x, y, acc, cusp::hyb_format());                                                                 \
  }
DEF(float, f32)
DEF(double, f64)
#undef DEF
int ref_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }
}
