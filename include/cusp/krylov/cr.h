// cusp/krylov/cr.h — cusp::krylov::cr(A, x, b[, monitor[, M]]): conjugate residuals
// (reference: cusp/krylov/cr.h, cusp/krylov/detail/cr.inl:35-128), for symmetric
// (possibly indefinite) A.  SURVEY §8(f) row 3: same hot-path kernels, another
// fusion pattern; operations are issued in the reference's order, including its
// periodic recomputation of the true residual every 8 iterations.
#pragma once
#include "../array1d.h"
#include "../blas/blas.h"
#include "../linear_operator.h"
#include "../monitor.h"
#include "../multiply.h"
#include "detail/fused.h"

namespace cusp {
namespace krylov {

namespace detail {
template <typename LinearOperator, typename VectorType1, typename VectorType2, typename Monitor,
          typename Preconditioner>
void cr_generic(const LinearOperator &A, VectorType1 &x, const VectorType2 &b, Monitor &monitor, Preconditioner &M) {
  typedef typename LinearOperator::value_type ValueType;
  typedef typename LinearOperator::memory_space Space;
  if (A.num_rows != A.num_cols || x.size() != A.num_rows || b.size() != A.num_rows)
    throw cusp::invalid_input_exception("cusp::krylov::cr: A must be square and match x, b");
  const size_t N = A.num_rows;
  const size_t recompute_r = 8;  // interval at which r is recomputed from b - A x
  cusp::array1d<ValueType, Space> y(N), z(N), r(N), p(N), Az(N), Ax(N);

  cusp::multiply(A, x, Ax);
  cusp::blas::axpby(b, Ax, r, ValueType(1), ValueType(-1));  // r <- b - A x
  cusp::multiply(M, r, z);                                   // z <- M r
  cusp::blas::copy(z, p);
  cusp::multiply(A, p, y);                                   // y <- A p
  cusp::multiply(A, z, Az);
  ValueType rz = cusp::blas::dotc(r, Az);

  while (!monitor.finished(r)) {
    const ValueType alpha = rz / cusp::blas::dotc(y, y);
    cusp::blas::axpy(p, x, alpha);                           // x += alpha p
    const size_t iter = monitor.iteration_count();
    if ((iter % recompute_r) && (iter > 0)) {
      cusp::blas::axpy(y, r, -alpha);                        // r -= alpha A p
    } else {
      cusp::multiply(A, x, Ax);
      cusp::blas::axpby(b, Ax, r, ValueType(1), ValueType(-1));
    }
    cusp::multiply(M, r, z);
    cusp::multiply(A, z, Az);
    const ValueType rz_old = rz;
    rz = cusp::blas::dotc(r, Az);
    const ValueType beta = rz / rz_old;
    cusp::blas::axpby(z, p, p, ValueType(1), beta);          // p <- z + beta p
    cusp::blas::axpby(Az, y, y, ValueType(1), beta);         // y <- A z + beta y  (= A p)
    ++monitor;
  }
}
template <typename LinearOperator, typename V1, typename V2, typename Monitor, typename Preconditioner>
void cr_dispatch(const LinearOperator &A, V1 &x, const V2 &b, Monitor &monitor, Preconditioner &M, std::true_type) {
  if (monitor.iteration_count() != 0) return cr_generic(A, x, b, monitor, M);
  solve_fused(B200SP_SOLVER_CR, A, x, b, monitor, M, 1);
}
template <typename LinearOperator, typename V1, typename V2, typename Monitor, typename Preconditioner>
void cr_dispatch(const LinearOperator &A, V1 &x, const V2 &b, Monitor &monitor, Preconditioner &M, std::false_type) {
  cr_generic(A, x, b, monitor, M);
}
}  // namespace detail

// fused device solve (b200sp_krylov: scalars and the monitor on the device, 2-3 fused vector kernels per iteration)
// where the operands allow it (detail/fused.h), otherwise operation by operation; one iterate sequence either way
template <typename LinearOperator, typename VectorType1, typename VectorType2, typename Monitor,
          typename Preconditioner>
void cr(const LinearOperator &A, VectorType1 &x, const VectorType2 &b, Monitor &monitor, Preconditioner &M) {
  if (A.num_rows != A.num_cols || x.size() != A.num_rows || b.size() != A.num_rows)
    throw cusp::invalid_input_exception("cusp::krylov::cr: A must be square and match x, b");
  detail::cr_dispatch(A, x, b, monitor, M, detail::can_fuse<LinearOperator, VectorType1, VectorType2, Monitor, Preconditioner>());
}

template <typename LinearOperator, typename VectorType1, typename VectorType2, typename Monitor>
void cr(const LinearOperator &A, VectorType1 &x, const VectorType2 &b, Monitor &monitor) {
  cusp::identity_operator<typename LinearOperator::value_type, typename LinearOperator::memory_space> M(A.num_rows,
                                                                                                        A.num_cols);
  cr(A, x, b, monitor, M);
}

template <typename LinearOperator, typename VectorType1, typename VectorType2>
void cr(const LinearOperator &A, VectorType1 &x, const VectorType2 &b) {
  cusp::monitor<typename LinearOperator::value_type> monitor(b);
  cr(A, x, b, monitor);
}

namespace detail {
namespace adl_default {
template <typename P, typename LinearOperator, typename VectorType1, typename VectorType2, typename Monitor,
          typename Preconditioner>
void cr(cusp::execution_policy<P> &, const LinearOperator &A, VectorType1 &x, const VectorType2 &b,
        Monitor &monitor, Preconditioner &M) {
  cusp::krylov::cr(A, x, b, monitor, M);
}
}  // namespace adl_default
}  // namespace detail
// leading execution policy: dispatched on the derived policy (cusp/memory.h: derived_cast)
template <typename P, typename LinearOperator, typename VectorType1, typename VectorType2, typename Monitor,
          typename Preconditioner>
void cr(const cusp::execution_policy<P> &exec, const LinearOperator &A, VectorType1 &x, const VectorType2 &b,
        Monitor &monitor, Preconditioner &M) {
  using detail::adl_default::cr;
  cusp::detail::stream_scope<P> on_stream(cusp::detail::derived_cast(exec));
  cr(cusp::detail::derived_cast(exec), A, x, b, monitor, M);
}

}  // namespace krylov
}  // namespace cusp
