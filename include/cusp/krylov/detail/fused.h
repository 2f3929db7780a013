// cusp/krylov/detail/fused.h — which solver calls can run as ONE fused device solve (b200sp_krylov):
// a device matrix the C ABI takes, a cusp::monitor<V> that has not counted iterations yet, vectors of the matrix's
// value type, and a preconditioner the engine knows — cusp::identity_operator or cusp::precond::diagonal on the
// device (its diagonal_reciprocals array is handed over as `diagonal_inverse`).  Everything else takes the
// operation-by-operation route of the solver's own header (same iterate sequence, one C-ABI call per operation).
#pragma once
#include <type_traits>
#include <vector>

#include "../../detail/descriptor.h"
#include "../../linear_operator.h"
#include "../../monitor.h"
#include "../../precond/diagonal.h"

namespace cusp {
namespace krylov {
namespace detail {

template <typename T>
struct is_identity_operator : std::false_type {};
template <typename V, typename S, typename I>
struct is_identity_operator<cusp::identity_operator<V, S, I>> : std::true_type {};
template <typename T>
struct is_cusp_monitor : std::false_type {};
template <typename V>
struct is_cusp_monitor<cusp::monitor<V>> : std::true_type {};

// 0: identity, 1: diagonal on the device, -1: anything else
template <typename P>
struct precond_kind : std::integral_constant<int, is_identity_operator<P>::value ? 0 : -1> {};
template <typename V>
struct precond_kind<cusp::precond::diagonal<V, cusp::device_memory>> : std::integral_constant<int, 1> {};

template <typename V, typename S, typename I>
const V *diagonal_inverse_of(const cusp::identity_operator<V, S, I> &) {
  return nullptr;
}
template <typename V>
const V *diagonal_inverse_of(const cusp::precond::diagonal<V, cusp::device_memory> &M) {
  return cusp::detail::raw_ptr(M.diagonal_reciprocals);
}
template <typename P>
const void *diagonal_inverse_of(const P &) {
  return nullptr;
}

template <typename LinearOperator, typename V1, typename V2, typename Monitor, typename Preconditioner>
struct can_fuse
    : std::integral_constant<
          bool, cusp::detail::abi_matrix<LinearOperator>::value && is_cusp_monitor<Monitor>::value &&
                    (precond_kind<typename std::remove_const<Preconditioner>::type>::value >= 0) &&
                    std::is_same<typename V1::value_type, typename LinearOperator::value_type>::value &&
                    std::is_same<typename V2::value_type, typename LinearOperator::value_type>::value> {};

// residual_slots: monitor.finished() calls per iteration (BiCGStab: 2)
template <typename LinearOperator, typename V1, typename V2, typename Monitor, typename Preconditioner>
void solve_fused(b200sp_solver solver, const LinearOperator &A, V1 &x, const V2 &b, Monitor &monitor, const Preconditioner &M,
                 size_t residual_slots) {
  using namespace cusp::detail;
  b200sp_matrix d = describe(A);
  b200sp_cg_params prm;
  prm.iteration_limit = (int64_t)monitor.iteration_limit();
  prm.relative_tolerance = (double)monitor.relative_tolerance();
  prm.absolute_tolerance = (double)monitor.absolute_tolerance();
  prm.check_interval = 0;
  b200sp_cg_result res;
  std::vector<double> history(residual_slots * monitor.iteration_limit() + 4, 0.0);
  check(b200sp_krylov(engine(), current_stream(), solver, &d, nullptr, raw_ptr(x), raw_ptr(b),
                      (const void *)diagonal_inverse_of(M), &prm, nullptr, &res, history.data()));
  monitor.absorb((size_t)res.iteration_count, history.data(), (size_t)res.num_residuals, res.b_norm);
}

}  // namespace detail
}  // namespace krylov
}  // namespace cusp
