// cusp/multiply.h — cusp::multiply(A, x, y) and the generalized forms
// (reference: cusp/multiply.h:36-195, cusp/detail/multiply.inl:27-105,
// cusp/system/detail/generic/multiply.inl:94-191).
//
// Dispatch (what ADL on the execution policy does in the reference):
//   * device_memory, sparse format, int32 indices, float/double values and the
//     functor triple (constant_functor(0) | identity, multiplies, plus):
//       -> the C ABI, b200sp_spmv(handle, stream, descriptor, x, y, accumulate, cfg)
//          — replaces cusp::system::cuda::detail::multiply(exec, A, x, y, init,
//          combine, reduce, <fmt>_format, array1d_format, array1d_format)
//          (csr_vector_spmv.h:218-258, ell_spmv.h:96-155, dia_spmv.h:129-188,
//          coo_flat_spmv.h:486-502, generic/multiply/spmv.h:272-290 for HYB).
//       For ELL / DIA (and ELL-R) with cusp::ktt enabled the 3-argument form runs
//       one step of dynamic autotuning per call, as generic/multiply.inl:141-154
//       does through KTT (b200sp_tune_step instead of an NVRTC compile).
//   * host_memory: the sequential loops, functor-generic — this is the
//     reference's host path (system/detail/sequential/multiply/*_spmv.h) and is
//     host code by definition; it is never used for device containers.
//   * anything else on the device (other functors, other value types):
//     cusp::not_implemented_exception — there is no silent fallback.
#pragma once
#include <algorithm>
#include <vector>

#include "array1d.h"
#include "array2d.h"
#include "blas/blas.h"
#include "detail/descriptor.h"
#include "functional.h"
#include "ktt/state.h"

namespace cusp {
namespace detail {

// ---------------------------------------------------------------------------
// host loops (reference order of operations per entry)
// ---------------------------------------------------------------------------
template <typename M, typename V1, typename V2, typename F0, typename F1, typename F2>
void host_multiply(const M &A, const V1 &x, V2 &y, F0 initialize, F1 combine, F2 reduce, array2d_format) {
  typedef typename V2::value_type T;
  for (size_t i = 0; i < A.num_rows; ++i) {
    T acc = initialize(y[i]);
    for (size_t j = 0; j < A.num_cols; ++j) acc = reduce(acc, combine((T)A(i, j), (T)x[j]));
    y[i] = acc;
  }
}
template <typename M, typename V1, typename V2, typename F0, typename F1, typename F2>
void host_multiply(const M &A, const V1 &x, V2 &y, F0 initialize, F1 combine, F2 reduce, csr_format) {
  typedef typename V2::value_type T;
  for (size_t i = 0; i < A.num_rows; ++i) {
    T acc = initialize(y[i]);
    const size_t lo = (size_t)A.row_offsets[i], hi = (size_t)A.row_offsets[i + 1];
    for (size_t k = lo; k < hi; ++k) acc = reduce(acc, combine((T)A.values[k], (T)x[(size_t)A.column_indices[k]]));
    y[i] = acc;
  }
}
template <typename M, typename V1, typename V2, typename F0, typename F1, typename F2>
void host_multiply(const M &A, const V1 &x, V2 &y, F0 initialize, F1 combine, F2 reduce, coo_format) {
  typedef typename V2::value_type T;
  for (size_t i = 0; i < A.num_rows; ++i) y[i] = initialize(y[i]);
  for (size_t k = 0; k < A.num_entries; ++k) {
    const size_t i = (size_t)A.row_indices[k];
    y[i] = reduce((T)y[i], combine((T)A.values[k], (T)x[(size_t)A.column_indices[k]]));
  }
}
template <typename M, typename V1, typename V2, typename F0, typename F1, typename F2>
void host_multiply(const M &A, const V1 &x, V2 &y, F0 initialize, F1 combine, F2 reduce, dia_format) {
  typedef typename V2::value_type T;
  typedef long long ll;
  for (size_t i = 0; i < A.num_rows; ++i) y[i] = initialize(y[i]);
  const size_t nd = A.values.num_cols;
  for (size_t d = 0; d < nd; ++d) {  // diagonal by diagonal, like dia_spmv.h:58-79
    const ll off = (ll)A.diagonal_offsets[d];
    const size_t r0 = (size_t)std::max<ll>(0, -off), c0 = (size_t)std::max<ll>(0, off);
    if (r0 >= A.num_rows || c0 >= A.num_cols) continue;
    const size_t len = std::min(A.num_rows - r0, A.num_cols - c0);
    for (size_t n = 0; n < len; ++n) y[r0 + n] = reduce((T)y[r0 + n], combine((T)A.values(r0 + n, d), (T)x[c0 + n]));
  }
}
template <typename M, typename V1, typename V2, typename F1, typename F2>
void host_ell_accumulate(const M &A, const V1 &x, V2 &y, F1 combine, F2 reduce) {
  typedef typename V2::value_type T;
  typedef typename M::index_type I;
  const I invalid = static_cast<I>(-1);
  const size_t K = A.column_indices.num_cols;
  for (size_t n = 0; n < K; ++n)  // slot-major, like ell_spmv.h:58-72
    for (size_t i = 0; i < A.num_rows; ++i) {
      const I j = A.column_indices(i, n);
      if (j != invalid) y[i] = reduce((T)y[i], combine((T)A.values(i, n), (T)x[(size_t)j]));
    }
}
template <typename M, typename V1, typename V2, typename F0, typename F1, typename F2>
void host_multiply(const M &A, const V1 &x, V2 &y, F0 initialize, F1 combine, F2 reduce, ell_format) {
  for (size_t i = 0; i < A.num_rows; ++i) y[i] = initialize(y[i]);
  host_ell_accumulate(A, x, y, combine, reduce);
}
template <typename M, typename V1, typename V2, typename F0, typename F1, typename F2>
void host_multiply(const M &A, const V1 &x, V2 &y, F0 initialize, F1 combine, F2 reduce, hyb_format) {
  typedef typename V2::value_type T;
  host_multiply(A.ell, x, y, initialize, combine, reduce, ell_format());
  host_multiply(A.coo, x, y, identity_function<T>(), combine, reduce, coo_format());  // hyb_spmv.h:52-56
}

// ---------------------------------------------------------------------------
// device: C ABI
// ---------------------------------------------------------------------------
template <typename M, typename V1, typename V2>
void device_spmv(const M &A, const V1 &x, V2 &y, int accumulate, bool allow_dynamic_tuning, std::true_type) {
  static_assert(std::is_same<typename M::value_type, typename V1::value_type>::value &&
                    std::is_same<typename M::value_type, typename V2::value_type>::value,
                "cusp::multiply on device_memory: A, x and y must share one value type (float or double)");
  b200sp_matrix d = describe(A);
  typedef typename M::format F;
  const bool tunable = std::is_same<F, ell_format>::value || std::is_same<F, dia_format>::value;
  if (tunable && allow_dynamic_tuning && !accumulate && cusp::ktt::detail::is_enabled())
    check(b200sp_tune_step(engine(), current_stream(), &d, raw_ptr(x), raw_ptr(y), nullptr));
  else
    check(b200sp_spmv(engine(), current_stream(), &d, raw_ptr(x), raw_ptr(y), accumulate, nullptr));
}
template <typename M, typename V1, typename V2>
void device_spmv(const M &, const V1 &, V2 &, int, bool, std::false_type) {
  throw cusp::not_implemented_exception(
      "cusp::multiply on device_memory: the B200 engine takes sparse matrices with 32-bit indices and float/double "
      "values");
}

template <typename M, typename V1, typename V2>
void check_shapes(const M &A, const V1 &x, const V2 &y) {
  if (A.num_cols != x.size() || A.num_rows != y.size())
    throw cusp::invalid_input_exception("cusp::multiply: matrix and vector dimensions do not match");
}

// host memory
template <typename M, typename V1, typename V2, typename F0, typename F1, typename F2>
void multiply_in(host_memory, const M &A, const V1 &x, V2 &y, F0 initialize, F1 combine, F2 reduce, bool) {
  host_multiply(A, x, y, initialize, combine, reduce, typename M::format());
}
// device memory
template <typename M, typename V1, typename V2>
void device_generalized(const M &A, const V1 &x, V2 &y, const b200sp_functors &f, std::true_type) {
  static_assert(std::is_same<typename M::value_type, typename V1::value_type>::value &&
                    std::is_same<typename M::value_type, typename V2::value_type>::value,
                "cusp::multiply on device_memory: A, x and y must share one value type (float or double)");
  b200sp_matrix d = describe(A);
  check(b200sp_spmv_generalized(engine(), current_stream(), &d, raw_ptr(x), raw_ptr(y), &f));
}
template <typename M, typename V1, typename V2>
void device_generalized(const M &, const V1 &, V2 &, const b200sp_functors &, std::false_type) {
  throw cusp::not_implemented_exception(
      "cusp::multiply on device_memory: the B200 engine takes sparse matrices with 32-bit indices and float/double "
      "values");
}
template <typename M, typename V1, typename V2, typename F0, typename F1, typename F2>
void multiply_in(device_memory, const M &A, const V1 &x, V2 &y, F0 initialize, F1, F2, bool allow_dynamic_tuning) {
  const int acc = init_kind<F0>::of(initialize);
  if (acc >= 0 && is_multiplies<F1>::value && is_plus<F2>::value) {  // the tuned kernels
    device_spmv(A, x, y, acc, allow_dynamic_tuning, std::integral_constant<bool, abi_matrix<M>::value>());
    return;
  }
  // any other triple the C ABI can name (b200sp_spmv_generalized): functor types -> codes
  b200sp_functors f;
  f.init_value = 0.0;
  f.initialize = init_code<F0>::of(initialize, f.init_value);
  f.combine = combine_code<F1>::value;
  f.reduce = reduce_code<F2>::value;
  if (f.initialize < 0 || f.combine < 0 || f.reduce < 0)
    throw cusp::not_implemented_exception(
        "cusp::multiply on device_memory: initialize in {constant_functor, identity}, combine in {multiplies, plus, "
        "minimum, maximum, project2nd}, reduce in {plus, minimum, maximum} (a C ABI cannot run user functor code on the "
        "device)");
  device_generalized(A, x, y, f, std::integral_constant<bool, abi_matrix<M>::value>());
}

// ---------------------------------------------------------------------------
// CSR x dense block: C = (init C) + A B for array2d B, C
// (host: sequential/multiply/csr_block_spmv.h:52-77; device: cuda/detail/multiply/csr_block_spmv.h:181-222
//  -> b200sp_spmm_csr_<t>).  Like the reference, only csr_matrix takes dense right-hand sides.
// ---------------------------------------------------------------------------
template <typename V>
struct is_array2d : std::is_same<typename V::format, array2d_format> {};

inline b200sp_status spmm_(int64_t r, int64_t c, int64_t n, const int *Ap, const int *Aj, const float *Ax, int64_t k,
                           const float *X, int64_t ldx, float *Y, int64_t ldy, int acc) {
  return b200sp_spmm_csr_f32(engine(), current_stream(), r, c, n, Ap, Aj, Ax, k, X, ldx, Y, ldy, acc);
}
inline b200sp_status spmm_(int64_t r, int64_t c, int64_t n, const int *Ap, const int *Aj, const double *Ax, int64_t k,
                           const double *X, int64_t ldx, double *Y, int64_t ldy, int acc) {
  return b200sp_spmm_csr_f64(engine(), current_stream(), r, c, n, Ap, Aj, Ax, k, X, ldx, Y, ldy, acc);
}

template <typename M, typename B, typename C, typename F0, typename F1, typename F2>
void block_multiply_in(host_memory, const M &A, const B &X, C &Y, F0 initialize, F1 combine, F2 reduce) {
  typedef typename C::value_type T;
  std::vector<T> acc(X.num_cols);
  for (size_t i = 0; i < A.num_rows; ++i) {
    for (size_t k = 0; k < X.num_cols; ++k) acc[k] = initialize((T)Y(i, k));
    const size_t lo = (size_t)A.row_offsets[i], hi = (size_t)A.row_offsets[i + 1];
    for (size_t jj = lo; jj < hi; ++jj) {
      const size_t j = (size_t)A.column_indices[jj];
      const T a = (T)A.values[jj];
      for (size_t k = 0; k < X.num_cols; ++k) acc[k] = reduce(acc[k], combine(a, (T)X(j, k)));
    }
    for (size_t k = 0; k < X.num_cols; ++k) Y(i, k) = acc[k];
  }
}
template <typename M, typename B, typename C, typename F0, typename F1, typename F2>
void block_multiply_in(device_memory, const M &A, const B &X, C &Y, F0 initialize, F1, F2) {
  typedef typename M::value_type T;
  const int acc = init_kind<F0>::of(initialize);
  constexpr bool abi = std::is_same<typename M::index_type, int>::value &&
                       (std::is_same<T, float>::value || std::is_same<T, double>::value) &&
                       std::is_same<typename B::value_type, T>::value && std::is_same<typename C::value_type, T>::value &&
                       std::is_same<typename B::orientation, row_major>::value &&
                       std::is_same<typename C::orientation, row_major>::value;
  if (acc < 0 || !is_multiplies<F1>::value || !is_plus<F2>::value || !abi)
    throw cusp::not_implemented_exception(
        "cusp::multiply(csr, array2d, array2d) on device_memory: int32 indices, one float/double value type, row-major "
        "blocks and (constant_functor(0) | identity, multiplies, plus)");
  if constexpr (abi)
    check(spmm_((int64_t)A.num_rows, (int64_t)A.num_cols, (int64_t)A.num_entries, raw_ptr(A.row_offsets),
                raw_ptr(A.column_indices), raw_ptr(A.values), (int64_t)X.num_cols, raw_ptr(X.values), (int64_t)X.pitch,
                raw_ptr(Y.values), (int64_t)Y.pitch, acc));
}
template <typename M, typename B, typename C, typename F0, typename F1, typename F2>
void block_multiply(const M &A, const B &X, C &Y, F0 initialize, F1 combine, F2 reduce) {
  static_assert(std::is_same<typename M::memory_space, typename B::memory_space>::value &&
                    std::is_same<typename M::memory_space, typename C::memory_space>::value,
                "cusp::multiply: A, B and C must live in the same memory space");
  if (!std::is_same<typename M::format, csr_format>::value)
    throw cusp::not_implemented_exception("cusp::multiply(sparse, array2d, array2d): csr_matrix only");
  if (A.num_cols != X.num_rows || A.num_rows != Y.num_rows || X.num_cols != Y.num_cols)
    throw cusp::invalid_input_exception("cusp::multiply: matrix and block dimensions do not match");
  if constexpr (std::is_same<typename M::format, csr_format>::value)
    block_multiply_in(typename M::memory_space(), A, X, Y, initialize, combine, reduce);
}

}  // namespace detail

// ---- 7-argument form (cusp/multiply.h:163-195) ------------------------------
namespace detail {
namespace adl_default {
template <typename P, typename LinearOperator, typename MatrixOrVector1, typename V2, typename UnaryFunction,
          typename BinaryFunction1, typename BinaryFunction2>
void multiply(cusp::execution_policy<P> &, const LinearOperator &A, const MatrixOrVector1 &B, V2 &C,
              UnaryFunction initialize, BinaryFunction1 combine, BinaryFunction2 reduce) {
  if constexpr (std::is_base_of<sparse_format, typename MatrixOrVector1::format>::value) {
    throw cusp::not_implemented_exception("cusp::multiply(A, B, C) with a sparse B (SpGEMM) is not part of the B200 engine");
  } else if constexpr (detail::is_array2d<MatrixOrVector1>::value && std::is_base_of<sparse_format, typename LinearOperator::format>::value) {
    detail::block_multiply(A, B, C, initialize, combine, reduce);
  } else {
    static_assert(std::is_same<typename LinearOperator::memory_space, typename MatrixOrVector1::memory_space>::value &&
                      std::is_same<typename LinearOperator::memory_space, typename V2::memory_space>::value,
                  "cusp::multiply: A, x and y must live in the same memory space");
    detail::check_shapes(A, B, C);
    detail::multiply_in(typename LinearOperator::memory_space(), A, B, C, initialize, combine, reduce, false);
  }
}
}  // namespace adl_default
}  // namespace detail
template <typename P, typename LinearOperator, typename MatrixOrVector1, typename MatrixOrVector2, typename UnaryFunction,
          typename BinaryFunction1, typename BinaryFunction2>
void multiply(const execution_policy<P> &exec, const LinearOperator &A, const MatrixOrVector1 &B, MatrixOrVector2 &&C,
              UnaryFunction initialize, BinaryFunction1 combine, BinaryFunction2 reduce) {
  using detail::adl_default::multiply;
  detail::stream_scope<P> on_stream(detail::derived_cast(exec));
  multiply(detail::derived_cast(exec), A, B, C, initialize, combine, reduce);
}
template <typename LinearOperator, typename MatrixOrVector1, typename MatrixOrVector2, typename UnaryFunction,
          typename BinaryFunction1, typename BinaryFunction2>
void multiply(const LinearOperator &A, const MatrixOrVector1 &B, MatrixOrVector2 &&C, UnaryFunction initialize,
              BinaryFunction1 combine, BinaryFunction2 reduce) {
  multiply(typename LinearOperator::memory_space(), A, B, C, initialize, combine, reduce);
}

namespace detail {
// operators that are not matrices (cusp::linear_operator subclasses, user
// functors with unknown_format) are applied through operator()
// (generic/multiply.inl:59-73)
template <typename LinearOperator, typename V1, typename V2>
void multiply3(const LinearOperator &A, const V1 &x, V2 &y, unknown_format) {
  const_cast<LinearOperator &>(A)(x, y);
}
template <typename LinearOperator, typename V1, typename V2>
void multiply3(const LinearOperator &A, const V1 &x, V2 &y, known_format) {
  typedef typename V2::value_type T;
  if constexpr (std::is_base_of<sparse_format, typename V1::format>::value) {
    // sparse x sparse (cusp/system/detail/generic/multiply/spgemm.h): outside the SpMV hot path (out of scope)
    throw cusp::not_implemented_exception("cusp::multiply(A, B, C) with a sparse B (SpGEMM) is not part of the B200 engine");
  } else if constexpr (is_array2d<V1>::value && std::is_base_of<sparse_format, typename LinearOperator::format>::value) {
    block_multiply(A, x, y, constant_functor<T>(T(0)), multiplies_function<T>(), plus_function<T>());
    return;
  } else {
  static_assert(std::is_same<typename LinearOperator::memory_space, typename V1::memory_space>::value &&
                    std::is_same<typename LinearOperator::memory_space, typename V2::memory_space>::value,
                "cusp::multiply: A, x and y must live in the same memory space");
  check_shapes(A, x, y);
  multiply_in(typename LinearOperator::memory_space(), A, x, y, constant_functor<T>(T(0)), multiplies_function<T>(),
              plus_function<T>(), true);
  }
}
}  // namespace detail

// ---- 3 / 4-argument forms (cusp/multiply.h:36-99) ---------------------------
template <typename LinearOperator, typename MatrixOrVector1, typename MatrixOrVector2>
void multiply(const LinearOperator &A, const MatrixOrVector1 &B, MatrixOrVector2 &&C) {
  detail::multiply3(A, B, C, typename LinearOperator::format());
}
namespace detail {
namespace adl_default {
template <typename P, typename LinearOperator, typename MatrixOrVector1, typename V2>
void multiply(cusp::execution_policy<P> &, const LinearOperator &A, const MatrixOrVector1 &B, V2 &C) {
  detail::multiply3(A, B, C, typename LinearOperator::format());
}
}  // namespace adl_default
}  // namespace detail
template <typename P, typename LinearOperator, typename MatrixOrVector1, typename MatrixOrVector2>
void multiply(const execution_policy<P> &exec, const LinearOperator &A, const MatrixOrVector1 &B, MatrixOrVector2 &&C) {
  using detail::adl_default::multiply;
  detail::stream_scope<P> on_stream(detail::derived_cast(exec));
  multiply(detail::derived_cast(exec), A, B, C);
}

// ---- generalized_spmv (cusp/multiply.h:197-280): z = reduce(y, A (combine) x)
template <typename LinearOperator, typename V1, typename V2, typename V3, typename BinaryFunction1,
          typename BinaryFunction2>
void generalized_spmv(const LinearOperator &A, const V1 &x, const V2 &y, V3 &&z, BinaryFunction1 combine,
                      BinaryFunction2 reduce) {
  typedef typename std::decay<V3>::type Z;
  typedef typename Z::value_type T;
  cusp::blas::copy(y, z);
  multiply(A, x, z, identity_function<T>(), combine, reduce);
}

}  // namespace cusp
