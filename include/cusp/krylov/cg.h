// cusp/krylov/cg.h — cusp::krylov::cg(A, x, b[, monitor[, M]])
// (reference: cusp/krylov/cg.h:43-150, cusp/krylov/detail/cg.inl:35-180).
//
// Two routes with one iterate sequence:
//   * fused: device matrix the C ABI takes + cusp::monitor<V> + the identity or the
//     diagonal (cusp::precond::diagonal) preconditioner  ->  b200sp_krylov / b200sp_cg
//     (3 kernels per iteration, scalars stay on the device, the monitor rule is
//     evaluated there and the residual history comes back in one piece);
//   * generic: any linear operator / monitor / preconditioner  ->  the textbook
//     loop over cusp::multiply and cusp::blas, operation by operation in the
//     order of cg.inl:63-105 (each of those calls is again a C-ABI call on
//     device containers).
#pragma once
#include <type_traits>
#include <vector>

#include "../array1d.h"
#include "../blas/blas.h"
#include "../detail/descriptor.h"
#include "../linear_operator.h"
#include "../monitor.h"
#include "../multiply.h"
#include "detail/fused.h"

namespace cusp {
namespace krylov {
namespace detail {

template <typename LinearOperator, typename V1, typename V2, typename Monitor, typename Preconditioner>
void cg_generic(const LinearOperator &A, V1 &x, const V2 &b, Monitor &monitor, Preconditioner &M) {
  typedef typename LinearOperator::value_type ValueType;
  typedef typename LinearOperator::memory_space Space;
  const size_t N = A.num_rows;
  cusp::array1d<ValueType, Space> y(N), z(N), r(N), p(N);

  cusp::multiply(A, x, y);                                      // y <- A x
  cusp::blas::axpby(b, y, r, ValueType(1), ValueType(-1));      // r <- b - y
  cusp::multiply(M, r, z);                                      // z <- M r
  cusp::blas::copy(z, p);                                       // p <- z
  ValueType rz = cusp::blas::dotc(r, z);
  while (!monitor.finished(r)) {
    cusp::multiply(A, p, y);                                    // y <- A p
    const ValueType alpha = rz / cusp::blas::dotc(y, p);
    cusp::blas::axpy(p, x, alpha);                              // x += alpha p
    cusp::blas::axpy(y, r, -alpha);                             // r -= alpha y
    cusp::multiply(M, r, z);                                    // z <- M r
    const ValueType rz_old = rz;
    rz = cusp::blas::dotc(r, z);
    const ValueType beta = rz / rz_old;
    cusp::blas::axpby(z, p, p, ValueType(1), beta);             // p <- z + beta p
    ++monitor;
  }
}

template <typename LinearOperator, typename V1, typename V2, typename Monitor, typename Preconditioner>
void cg_dispatch(const LinearOperator &A, V1 &x, const V2 &b, Monitor &monitor, Preconditioner &M, std::true_type) {
  if (monitor.iteration_count() != 0) return cg_generic(A, x, b, monitor, M);  // a monitor in mid-count: step by step
  solve_fused(B200SP_SOLVER_CG, A, x, b, monitor, M, 1);
}
template <typename LinearOperator, typename V1, typename V2, typename Monitor, typename Preconditioner>
void cg_dispatch(const LinearOperator &A, V1 &x, const V2 &b, Monitor &monitor, Preconditioner &M, std::false_type) {
  cg_generic(A, x, b, monitor, M);
}

}  // namespace detail

template <typename LinearOperator, typename VectorType1, typename VectorType2, typename Monitor,
          typename Preconditioner>
void cg(const LinearOperator &A, VectorType1 &x, const VectorType2 &b, Monitor &monitor, Preconditioner &M) {
  if (A.num_rows != A.num_cols || x.size() != A.num_rows || b.size() != A.num_rows)
    throw cusp::invalid_input_exception("cusp::krylov::cg: A must be square and match x, b");
  typedef detail::can_fuse<LinearOperator, VectorType1, VectorType2, Monitor, Preconditioner> fused;
  detail::cg_dispatch(A, x, b, monitor, M, fused());
}

template <typename LinearOperator, typename VectorType1, typename VectorType2, typename Monitor>
void cg(const LinearOperator &A, VectorType1 &x, const VectorType2 &b, Monitor &monitor) {
  typedef typename LinearOperator::value_type ValueType;
  typedef typename LinearOperator::memory_space MemorySpace;
  cusp::identity_operator<ValueType, MemorySpace> M(A.num_rows, A.num_cols);
  cg(A, x, b, monitor, M);
}

template <typename LinearOperator, typename VectorType1, typename VectorType2>
void cg(const LinearOperator &A, VectorType1 &x, const VectorType2 &b) {
  typedef typename LinearOperator::value_type ValueType;
  cusp::monitor<ValueType> monitor(b);
  cg(A, x, b, monitor);
}

// leading execution policy (cg.h:43-70)
namespace detail {
namespace adl_default {
template <typename P, typename LinearOperator, typename VectorType1, typename VectorType2, typename Monitor,
          typename Preconditioner>
void cg(cusp::execution_policy<P> &, const LinearOperator &A, VectorType1 &x, const VectorType2 &b,
        Monitor &monitor, Preconditioner &M) {
  cusp::krylov::cg(A, x, b, monitor, M);
}
}  // namespace adl_default
}  // namespace detail
// leading execution policy: dispatched on the derived policy (cusp/memory.h: derived_cast)
template <typename P, typename LinearOperator, typename VectorType1, typename VectorType2, typename Monitor,
          typename Preconditioner>
void cg(const cusp::execution_policy<P> &exec, const LinearOperator &A, VectorType1 &x, const VectorType2 &b,
        Monitor &monitor, Preconditioner &M) {
  using detail::adl_default::cg;
  cusp::detail::stream_scope<P> on_stream(cusp::detail::derived_cast(exec));
  cg(cusp::detail::derived_cast(exec), A, x, b, monitor, M);
}

}  // namespace krylov
}  // namespace cusp
