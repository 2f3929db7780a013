// cusp/krylov/bicgstab.h — cusp::krylov::bicgstab(A, x, b[, monitor[, M]])
// (reference: cusp/krylov/bicgstab.h, cusp/krylov/detail/bicgstab.inl:35-123).
// the solver is a different fusion pattern over the same
// hot-path kernels — every cusp::multiply / cusp::blas call below is a C-ABI call
// on device containers (b200sp_spmv, b200sp_axpby, b200sp_axpbypcz, b200sp_dot),
// issued in the reference's order so the iterate sequence is the same.
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
void bicgstab_generic(const LinearOperator &A, VectorType1 &x, const VectorType2 &b, Monitor &monitor, Preconditioner &M) {
  typedef typename LinearOperator::value_type ValueType;
  typedef typename LinearOperator::memory_space Space;
  if (A.num_rows != A.num_cols || x.size() != A.num_rows || b.size() != A.num_rows)
    throw cusp::invalid_input_exception("cusp::krylov::bicgstab: A must be square and match x, b");
  const size_t N = A.num_rows;
  cusp::array1d<ValueType, Space> p(N), r(N), r_star(N), s(N), Mp(N), AMp(N), Ms(N), AMs(N);

  cusp::multiply(A, x, r);                                       // r <- A x
  cusp::blas::axpby(b, r, r, ValueType(1), ValueType(-1));       // r <- b - r
  cusp::blas::copy(r, p);
  cusp::blas::copy(r, r_star);
  ValueType rho_old = cusp::blas::dotc(r_star, r);

  while (!monitor.finished(r)) {
    cusp::multiply(M, p, Mp);
    cusp::multiply(A, Mp, AMp);
    const ValueType alpha = rho_old / cusp::blas::dotc(r_star, AMp);
    cusp::blas::axpby(r, AMp, s, ValueType(1), ValueType(-alpha));   // s <- r - alpha A M p
    if (monitor.finished(s)) {
      cusp::blas::axpby(x, Mp, x, ValueType(1), ValueType(alpha));   // x += alpha M p
      break;
    }
    cusp::multiply(M, s, Ms);
    cusp::multiply(A, Ms, AMs);
    const ValueType omega = cusp::blas::dotc(AMs, s) / cusp::blas::dotc(AMs, AMs);
    cusp::blas::axpbypcz(x, Mp, Ms, x, ValueType(1), alpha, omega);  // x += alpha M p + omega M s
    cusp::blas::axpby(s, AMs, r, ValueType(1), -omega);              // r <- s - omega A M s
    const ValueType rho_new = cusp::blas::dotc(r_star, r);
    const ValueType beta = (rho_new / rho_old) * (alpha / omega);
    rho_old = rho_new;
    cusp::blas::axpbypcz(r, p, AMp, p, ValueType(1), beta, -beta * omega);  // p <- r + beta (p - omega A M p)
    ++monitor;
  }
}
template <typename LinearOperator, typename V1, typename V2, typename Monitor, typename Preconditioner>
void bicgstab_dispatch(const LinearOperator &A, V1 &x, const V2 &b, Monitor &monitor, Preconditioner &M, std::true_type) {
  if (monitor.iteration_count() != 0) return bicgstab_generic(A, x, b, monitor, M);
  solve_fused(B200SP_SOLVER_BICGSTAB, A, x, b, monitor, M, 2);
}
template <typename LinearOperator, typename V1, typename V2, typename Monitor, typename Preconditioner>
void bicgstab_dispatch(const LinearOperator &A, V1 &x, const V2 &b, Monitor &monitor, Preconditioner &M, std::false_type) {
  bicgstab_generic(A, x, b, monitor, M);
}
}  // namespace detail

// fused device solve (b200sp_krylov: scalars and the monitor on the device, 2-3 fused vector kernels per iteration)
// where the operands allow it (detail/fused.h), otherwise operation by operation; one iterate sequence either way
template <typename LinearOperator, typename VectorType1, typename VectorType2, typename Monitor,
          typename Preconditioner>
void bicgstab(const LinearOperator &A, VectorType1 &x, const VectorType2 &b, Monitor &monitor, Preconditioner &M) {
  if (A.num_rows != A.num_cols || x.size() != A.num_rows || b.size() != A.num_rows)
    throw cusp::invalid_input_exception("cusp::krylov::bicgstab: A must be square and match x, b");
  detail::bicgstab_dispatch(A, x, b, monitor, M, detail::can_fuse<LinearOperator, VectorType1, VectorType2, Monitor, Preconditioner>());
}

template <typename LinearOperator, typename VectorType1, typename VectorType2, typename Monitor>
void bicgstab(const LinearOperator &A, VectorType1 &x, const VectorType2 &b, Monitor &monitor) {
  cusp::identity_operator<typename LinearOperator::value_type, typename LinearOperator::memory_space> M(A.num_rows,
                                                                                                        A.num_cols);
  bicgstab(A, x, b, monitor, M);
}

template <typename LinearOperator, typename VectorType1, typename VectorType2>
void bicgstab(const LinearOperator &A, VectorType1 &x, const VectorType2 &b) {
  cusp::monitor<typename LinearOperator::value_type> monitor(b);
  bicgstab(A, x, b, monitor);
}

namespace detail {
namespace adl_default {
template <typename P, typename LinearOperator, typename VectorType1, typename VectorType2, typename Monitor,
          typename Preconditioner>
void bicgstab(cusp::execution_policy<P> &, const LinearOperator &A, VectorType1 &x, const VectorType2 &b,
              Monitor &monitor, Preconditioner &M) {
  cusp::krylov::bicgstab(A, x, b, monitor, M);
}
}  // namespace adl_default
}  // namespace detail
// leading execution policy: dispatched on the derived policy (cusp/memory.h: derived_cast)
template <typename P, typename LinearOperator, typename VectorType1, typename VectorType2, typename Monitor,
          typename Preconditioner>
void bicgstab(const cusp::execution_policy<P> &exec, const LinearOperator &A, VectorType1 &x, const VectorType2 &b,
              Monitor &monitor, Preconditioner &M) {
  using detail::adl_default::bicgstab;
  cusp::detail::stream_scope<P> on_stream(cusp::detail::derived_cast(exec));
  bicgstab(cusp::detail::derived_cast(exec), A, x, b, monitor, M);
}

}  // namespace krylov
}  // namespace cusp
