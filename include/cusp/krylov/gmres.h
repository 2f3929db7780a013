// cusp/krylov/gmres.h — cusp::krylov::gmres(A, x, b, restart[, monitor[, M]])
// (reference: cusp/krylov/gmres.h, cusp/krylov/detail/gmres.inl:36-227).
// SURVEY §8(f) row 3.  Restarted GMRES with modified Gram-Schmidt: the Arnoldi basis is a
// column-major array2d in the operator's memory space, every inner product / update /
// product below is a C-ABI call on device containers (b200sp_spmv, b200sp_dot, b200sp_axpy,
// b200sp_scal, b200sp_nrm2) in the reference's order; the (restart+1) x restart Hessenberg
// matrix, the Givens rotations and the triangular solve stay on the host like there.
#pragma once
#include <cmath>

#include "../array1d.h"
#include "../array2d.h"
#include "../blas/blas.h"
#include "../linear_operator.h"
#include "../monitor.h"
#include "../multiply.h"

namespace cusp {
namespace krylov {
namespace gmres_detail {

// gmres.inl:36-71 (real value types)
template <typename T>
void ApplyPlaneRotation(T &dx, T &dy, const T &cs, const T &sn) {
  const T temp = cs * dx + sn * dy;
  dy = -sn * dx + cs * dy;
  dx = temp;
}
template <typename T>
void GeneratePlaneRotation(const T &dx, const T &dy, T &cs, T &sn) {
  if (dx == T(0)) {
    cs = T(0);
    sn = T(1);
  } else {
    const T scale = std::abs(dx) + std::abs(dy);
    const T norm = scale * std::sqrt(std::abs(dx / scale) * std::abs(dx / scale) +
                                     std::abs(dy / scale) * std::abs(dy / scale));
    const T alpha = dx / std::abs(dx);
    cs = std::abs(dx) / norm;
    sn = alpha * dy / norm;
  }
}
// gmres.inl:73-91
template <typename Hessenberg, typename V>
void PlaneRotation(Hessenberg &H, V &cs, V &sn, V &s, const int i) {
  typedef typename V::value_type T;
  for (int k = 0; k < i; k++) {
    T a = H(k, i), b = H(k + 1, i);
    ApplyPlaneRotation(a, b, (T)cs[k], (T)sn[k]);
    H(k, i) = a;
    H(k + 1, i) = b;
  }
  T c, sN;
  GeneratePlaneRotation((T)H(i, i), (T)H(i + 1, i), c, sN);
  cs[i] = c;
  sn[i] = sN;
  T a = H(i, i), b = H(i + 1, i);
  ApplyPlaneRotation(a, b, c, sN);
  H(i, i) = a;
  H(i + 1, i) = b;
  T s0 = s[i], s1 = s[i + 1];
  ApplyPlaneRotation(s0, s1, c, sN);
  s[i] = s0;
  s[i + 1] = s1;
}

}  // namespace gmres_detail

template <typename LinearOperator, typename VectorType1, typename VectorType2, typename Monitor,
          typename Preconditioner>
void gmres(const LinearOperator &A, VectorType1 &x, const VectorType2 &b, const size_t restart, Monitor &monitor,
           Preconditioner &M) {
  typedef typename LinearOperator::value_type ValueType;
  typedef typename LinearOperator::memory_space Space;
  if (A.num_rows != A.num_cols || x.size() != A.num_rows || b.size() != A.num_rows)
    throw cusp::invalid_input_exception("cusp::krylov::gmres: A must be square and match x, b");
  if (restart == 0) throw cusp::invalid_input_exception("cusp::krylov::gmres: restart must be positive");
  const size_t N = A.num_rows;
  const int R = (int)restart;
  int i, j, k;
  ValueType beta = 0;

  cusp::array1d<ValueType, Space> w(N), V0(N);
  cusp::array2d<ValueType, Space, cusp::column_major> V(N, R + 1, ValueType(0));  // Arnoldi basis
  // host workspace (gmres.inl:127-133)
  cusp::array2d<ValueType, cusp::host_memory, cusp::column_major> H(R + 1, R, ValueType(0));
  cusp::array1d<ValueType, cusp::host_memory> s(R + 1), cs(R), sn(R), resid(1);

  do {
    cusp::multiply(A, x, w);                          // w = A x
    cusp::blas::axpy(b, w, ValueType(-1));            // w = A x - b
    cusp::multiply(M, w, w);                          // w = M w
    beta = cusp::blas::nrm2(w);
    cusp::blas::scal(w, ValueType(-1.0 / beta));      // w = -w / beta
    {
      auto v0 = V.column(0);
      cusp::blas::copy(w, v0);
    }
    cusp::blas::fill(s, ValueType(0));
    s[0] = beta;
    i = -1;
    resid[0] = std::abs((ValueType)s[0]);
    if (monitor.finished(resid)) break;

    do {
      ++i;
      ++monitor;
      cusp::multiply(A, w, V0);
      cusp::multiply(M, V0, w);                       // w = M A V(i)
      for (k = 0; k <= i; k++) {
        auto vk = V.column(k);
        H(k, i) = cusp::blas::dotc(vk, w);            // H(k,i) = <V(k), w>
        cusp::blas::axpy(vk, w, -(ValueType)H(k, i)); // w -= H(k,i) V(k)
      }
      H(i + 1, i) = cusp::blas::nrm2(w);
      cusp::blas::scal(w, ValueType(1) / (ValueType)H(i + 1, i));
      {
        auto vn = V.column(i + 1);
        cusp::blas::copy(w, vn);
      }
      gmres_detail::PlaneRotation(H, cs, sn, s, i);
      resid[0] = std::abs((ValueType)s[i + 1]);
      if (monitor.finished(resid)) break;
    } while (i + 1 < R && monitor.iteration_count() + 1 <= monitor.iteration_limit());

    // back substitution on the host (gmres.inl:193-200)
    for (j = i; j >= 0; j--) {
      s[j] = (ValueType)s[j] / (ValueType)H(j, j);
      for (k = j - 1; k >= 0; k--) s[k] = (ValueType)s[k] - (ValueType)H(k, j) * (ValueType)s[j];
    }
    for (j = 0; j <= i; j++) {
      auto vj = V.column(j);
      cusp::blas::axpy(vj, x, (ValueType)s[j]);       // x += s[j] V(j)
    }
  } while (!monitor.finished(resid));
}

template <typename LinearOperator, typename VectorType1, typename VectorType2, typename Monitor>
void gmres(const LinearOperator &A, VectorType1 &x, const VectorType2 &b, const size_t restart, Monitor &monitor) {
  cusp::identity_operator<typename LinearOperator::value_type, typename LinearOperator::memory_space> M(A.num_rows,
                                                                                                        A.num_cols);
  gmres(A, x, b, restart, monitor, M);
}

template <typename LinearOperator, typename VectorType1, typename VectorType2>
void gmres(const LinearOperator &A, VectorType1 &x, const VectorType2 &b, const size_t restart) {
  cusp::monitor<typename LinearOperator::value_type> monitor(b);
  gmres(A, x, b, restart, monitor);
}

namespace detail {
namespace adl_default {
template <typename P, typename LinearOperator, typename VectorType1, typename VectorType2, typename Monitor,
          typename Preconditioner>
void gmres(cusp::execution_policy<P> &, const LinearOperator &A, VectorType1 &x, const VectorType2 &b,
           const size_t restart, Monitor &monitor, Preconditioner &M) {
  cusp::krylov::gmres(A, x, b, restart, monitor, M);
}
}  // namespace adl_default
}  // namespace detail
// leading execution policy: dispatched on the derived policy (cusp/memory.h: derived_cast)
template <typename P, typename LinearOperator, typename VectorType1, typename VectorType2, typename Monitor,
          typename Preconditioner>
void gmres(const cusp::execution_policy<P> &exec, const LinearOperator &A, VectorType1 &x, const VectorType2 &b,
           const size_t restart, Monitor &monitor, Preconditioner &M) {
  using detail::adl_default::gmres;
  cusp::detail::stream_scope<P> on_stream(cusp::detail::derived_cast(exec));
  gmres(cusp::detail::derived_cast(exec), A, x, b, restart, monitor, M);
}

}  // namespace krylov
}  // namespace cusp
