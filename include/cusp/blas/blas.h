// cusp/blas/blas.h — cusp::blas level-1 front end
// (reference: cusp/blas/blas.h -> cusp/blas.h, cusp/detail/blas.inl:84-461,
// cusp/system/detail/generic/blas.h:99-340).
//
// Device arrays (float / double) go straight to the C ABI (b200sp_<op>_<t>,
// include/b200sp.h); host arrays run the same formulas as plain loops, like the
// reference's host_memory path.  Every function exists with and without a
// leading execution policy.  Size mismatches throw
// cusp::invalid_input_exception (testing/blas.cu:141).
// BLAS-2/3 entry points throw cusp::not_implemented_exception like the
// reference's generic back end.
#pragma once
#include <cmath>
#include <type_traits>

#include "../array1d.h"
#include "../array2d.h"

namespace cusp {
namespace blas {
namespace detail {

using cusp::detail::check;
using cusp::detail::current_stream;
using cusp::detail::engine;
using cusp::detail::raw_ptr;

template <typename A, typename B>
inline void same_size(const A &a, const B &b, const char *what) {
  if (a.size() != b.size()) throw cusp::invalid_input_exception(std::string(what) + ": array dimensions do not match");
}

template <typename Array>
struct is_device : std::is_same<typename Array::memory_space, cusp::device_memory> {};

// the device engine computes in float or double only
template <typename T>
struct abi_type : std::false_type {};
template <>
struct abi_type<float> : std::true_type {};
template <>
struct abi_type<double> : std::true_type {};

template <typename T>
inline void require_abi(const char *what) {
  if (!abi_type<T>::value)
    throw cusp::not_implemented_exception(std::string(what) + ": device BLAS supports float and double");
}

// typed trampolines: pick the _f32 / _f64 symbol without reinterpret_casts at the call sites
inline b200sp_status axpy_(int64_t n, float a, const float *x, float *y) {
  return b200sp_axpy_f32(engine(), current_stream(), n, a, x, y);
}
inline b200sp_status axpy_(int64_t n, double a, const double *x, double *y) {
  return b200sp_axpy_f64(engine(), current_stream(), n, a, x, y);
}
inline b200sp_status axpby_(int64_t n, float a, const float *x, float b, const float *y, float *z) {
  return b200sp_axpby_f32(engine(), current_stream(), n, a, x, b, y, z);
}
inline b200sp_status axpby_(int64_t n, double a, const double *x, double b, const double *y, double *z) {
  return b200sp_axpby_f64(engine(), current_stream(), n, a, x, b, y, z);
}
inline b200sp_status axpbypcz_(int64_t n, float a, const float *x, float b, const float *y, float c, const float *z,
                               float *o) {
  return b200sp_axpbypcz_f32(engine(), current_stream(), n, a, x, b, y, c, z, o);
}
inline b200sp_status axpbypcz_(int64_t n, double a, const double *x, double b, const double *y, double c,
                               const double *z, double *o) {
  return b200sp_axpbypcz_f64(engine(), current_stream(), n, a, x, b, y, c, z, o);
}
inline b200sp_status xmy_(int64_t n, const float *x, const float *y, float *z) {
  return b200sp_xmy_f32(engine(), current_stream(), n, x, y, z);
}
inline b200sp_status xmy_(int64_t n, const double *x, const double *y, double *z) {
  return b200sp_xmy_f64(engine(), current_stream(), n, x, y, z);
}
inline b200sp_status copy_(int64_t n, const float *x, float *y) {
  return b200sp_copy_f32(engine(), current_stream(), n, x, y);
}
inline b200sp_status copy_(int64_t n, const double *x, double *y) {
  return b200sp_copy_f64(engine(), current_stream(), n, x, y);
}
inline b200sp_status fill_(int64_t n, float a, float *x) { return b200sp_fill_f32(engine(), current_stream(), n, a, x); }
inline b200sp_status fill_(int64_t n, double a, double *x) {
  return b200sp_fill_f64(engine(), current_stream(), n, a, x);
}
inline b200sp_status scal_(int64_t n, float a, float *x) { return b200sp_scal_f32(engine(), current_stream(), n, a, x); }
inline b200sp_status scal_(int64_t n, double a, double *x) {
  return b200sp_scal_f64(engine(), current_stream(), n, a, x);
}
inline b200sp_status dot_(int64_t n, const float *x, const float *y, float *r) {
  return b200sp_dot_f32(engine(), current_stream(), n, x, y, nullptr, r);
}
inline b200sp_status dot_(int64_t n, const double *x, const double *y, double *r) {
  return b200sp_dot_f64(engine(), current_stream(), n, x, y, nullptr, r);
}
inline b200sp_status nrm2_(int64_t n, const float *x, float *r) {
  return b200sp_nrm2_f32(engine(), current_stream(), n, x, nullptr, r);
}
inline b200sp_status nrm2_(int64_t n, const double *x, double *r) {
  return b200sp_nrm2_f64(engine(), current_stream(), n, x, nullptr, r);
}
inline b200sp_status asum_(int64_t n, const float *x, float *r) {
  return b200sp_asum_f32(engine(), current_stream(), n, x, nullptr, r);
}
inline b200sp_status asum_(int64_t n, const double *x, double *r) {
  return b200sp_asum_f64(engine(), current_stream(), n, x, nullptr, r);
}
inline b200sp_status nrmmax_(int64_t n, const float *x, float *r) {
  return b200sp_nrmmax_f32(engine(), current_stream(), n, x, nullptr, r);
}
inline b200sp_status nrmmax_(int64_t n, const double *x, double *r) {
  return b200sp_nrmmax_f64(engine(), current_stream(), n, x, nullptr, r);
}
inline b200sp_status amax_(int64_t n, const float *x, int *r) {
  return b200sp_amax_f32(engine(), current_stream(), n, x, r);
}
inline b200sp_status amax_(int64_t n, const double *x, int *r) {
  return b200sp_amax_f64(engine(), current_stream(), n, x, r);
}
// unsupported value types on the device: never called (require_abi throws first)
template <typename... Args>
inline b200sp_status axpy_(Args...) { return B200SP_NOT_IMPLEMENTED; }
template <typename... Args>
inline b200sp_status axpby_(Args...) { return B200SP_NOT_IMPLEMENTED; }
template <typename... Args>
inline b200sp_status axpbypcz_(Args...) { return B200SP_NOT_IMPLEMENTED; }
template <typename... Args>
inline b200sp_status xmy_(Args...) { return B200SP_NOT_IMPLEMENTED; }
template <typename... Args>
inline b200sp_status copy_(Args...) { return B200SP_NOT_IMPLEMENTED; }
template <typename... Args>
inline b200sp_status fill_(Args...) { return B200SP_NOT_IMPLEMENTED; }
template <typename... Args>
inline b200sp_status scal_(Args...) { return B200SP_NOT_IMPLEMENTED; }
template <typename... Args>
inline b200sp_status dot_(Args...) { return B200SP_NOT_IMPLEMENTED; }
template <typename... Args>
inline b200sp_status nrm2_(Args...) { return B200SP_NOT_IMPLEMENTED; }
template <typename... Args>
inline b200sp_status asum_(Args...) { return B200SP_NOT_IMPLEMENTED; }
template <typename... Args>
inline b200sp_status nrmmax_(Args...) { return B200SP_NOT_IMPLEMENTED; }
template <typename... Args>
inline b200sp_status amax_(Args...) { return B200SP_NOT_IMPLEMENTED; }

// const / mutable raw pointers of an array or view (exact types, so the typed
// trampolines above win overload resolution over the variadic fallbacks)
template <typename A>
inline const typename A::value_type *cptr(const A &a) {
  return raw_ptr(a);
}
template <typename A>
inline typename std::remove_const<A>::type::value_type *mptr(A &a) {
  return raw_ptr(a);
}

template <typename T>
inline T abs_of(const T &v) {
  return v < T(0) ? -v : v;
}

// all operands of one call must live in the same memory space
template <typename A, typename B>
inline void same_space() {
  static_assert(std::is_same<typename A::memory_space, typename B::memory_space>::value,
                "cusp::blas: operands must share a memory space");
}

}  // namespace detail

// ---- y <- alpha*x + y ------------------------------------------------------
template <typename Array1, typename Array2, typename Scalar>
void axpy(const Array1 &x, Array2 &&y, const Scalar alpha) {
  typedef typename std::decay<Array2>::type A2;
  typedef typename A2::value_type T;
  detail::same_space<Array1, A2>();
  detail::same_size(x, y, "axpy");
  if (detail::is_device<A2>::value) {
    detail::require_abi<T>("axpy");
    detail::check(detail::axpy_((int64_t)y.size(), (T)alpha, detail::cptr(x), detail::mptr(y)));
  } else {
    auto px = detail::cptr(x);
    auto py = detail::mptr(y);
    for (size_t i = 0; i < y.size(); ++i) py[i] = (T)alpha * px[i] + py[i];
  }
}

// ---- z <- alpha*x + beta*y -------------------------------------------------
template <typename Array1, typename Array2, typename Array3, typename Scalar1, typename Scalar2>
void axpby(const Array1 &x, const Array2 &y, Array3 &&z, const Scalar1 alpha, const Scalar2 beta) {
  typedef typename std::decay<Array3>::type A3;
  typedef typename A3::value_type T;
  detail::same_space<Array1, A3>();
  detail::same_space<Array2, A3>();
  detail::same_size(x, y, "axpby");
  detail::same_size(x, z, "axpby");
  if (detail::is_device<A3>::value) {
    detail::require_abi<T>("axpby");
    detail::check(detail::axpby_((int64_t)z.size(), (T)alpha, detail::cptr(x), (T)beta, detail::cptr(y),
                                 detail::mptr(z)));
  } else {
    auto px = detail::cptr(x);
    auto py = detail::cptr(y);
    auto pz = detail::mptr(z);
    for (size_t i = 0; i < z.size(); ++i) pz[i] = (T)alpha * px[i] + (T)beta * py[i];
  }
}

// ---- output <- alpha*x + beta*y + gamma*z ----------------------------------
template <typename Array1, typename Array2, typename Array3, typename Array4, typename S1, typename S2, typename S3>
void axpbypcz(const Array1 &x, const Array2 &y, const Array3 &z, Array4 &&output, const S1 alpha, const S2 beta,
              const S3 gamma) {
  typedef typename std::decay<Array4>::type A4;
  typedef typename A4::value_type T;
  detail::same_space<Array1, A4>();
  detail::same_size(x, y, "axpbypcz");
  detail::same_size(x, z, "axpbypcz");
  detail::same_size(x, output, "axpbypcz");
  if (detail::is_device<A4>::value) {
    detail::require_abi<T>("axpbypcz");
    detail::check(detail::axpbypcz_((int64_t)output.size(), (T)alpha, detail::cptr(x), (T)beta, detail::cptr(y),
                                    (T)gamma, detail::cptr(z), detail::mptr(output)));
  } else {
    auto px = detail::cptr(x);
    auto py = detail::cptr(y);
    auto pz = detail::cptr(z);
    auto po = detail::mptr(output);
    for (size_t i = 0; i < output.size(); ++i) po[i] = (T)alpha * px[i] + (T)beta * py[i] + (T)gamma * pz[i];
  }
}

// ---- z <- x .* y -----------------------------------------------------------
template <typename Array1, typename Array2, typename Array3>
void xmy(const Array1 &x, const Array2 &y, Array3 &&z) {
  typedef typename std::decay<Array3>::type A3;
  typedef typename A3::value_type T;
  detail::same_space<Array1, A3>();
  detail::same_size(x, y, "xmy");
  detail::same_size(x, z, "xmy");
  if (detail::is_device<A3>::value) {
    detail::require_abi<T>("xmy");
    detail::check(detail::xmy_((int64_t)z.size(), detail::cptr(x), detail::cptr(y), detail::mptr(z)));
  } else {
    auto px = detail::cptr(x);
    auto py = detail::cptr(y);
    auto pz = detail::mptr(z);
    for (size_t i = 0; i < z.size(); ++i) pz[i] = px[i] * py[i];
  }
}

// ---- y <- x ----------------------------------------------------------------
template <typename Array1, typename Array2>
void copy(const Array1 &x, Array2 &&y) {
  typedef typename std::decay<Array2>::type A2;
  typedef typename A2::value_type T;
  detail::same_size(x, y, "copy");
  if (detail::is_device<A2>::value && detail::is_device<Array1>::value &&
      std::is_same<T, typename Array1::value_type>::value && detail::abi_type<T>::value) {
    detail::check(detail::copy_((int64_t)y.size(), detail::cptr(x), detail::mptr(y)));
  } else {
    cusp::detail::any_copy<T, typename Array1::value_type, typename Array1::memory_space, typename A2::memory_space>(
        detail::cptr(x), detail::mptr(y), y.size());
  }
}

// ---- x <- alpha ------------------------------------------------------------
template <typename Array, typename Scalar>
void fill(Array &&x, const Scalar alpha) {
  typedef typename std::decay<Array>::type A;
  typedef typename A::value_type T;
  if (detail::is_device<A>::value && detail::abi_type<T>::value) {
    detail::check(detail::fill_((int64_t)x.size(), (T)alpha, detail::mptr(x)));
  } else if (detail::is_device<A>::value) {
    std::vector<T> h(x.size(), (T)alpha);
    cusp::detail::raw_copy<T, cusp::host_memory, cusp::device_memory>(h.data(), detail::mptr(x), h.size());
  } else {
    auto px = detail::mptr(x);
    for (size_t i = 0; i < x.size(); ++i) px[i] = (T)alpha;
  }
}

// ---- x <- alpha*x ----------------------------------------------------------
template <typename Array, typename Scalar>
void scal(Array &&x, const Scalar alpha) {
  typedef typename std::decay<Array>::type A;
  typedef typename A::value_type T;
  if (detail::is_device<A>::value) {
    detail::require_abi<T>("scal");
    detail::check(detail::scal_((int64_t)x.size(), (T)alpha, detail::mptr(x)));
  } else {
    auto px = detail::mptr(x);
    for (size_t i = 0; i < x.size(); ++i) px[i] = (T)alpha * px[i];
  }
}

// ---- reductions (return the scalar to the host, like the reference) --------
template <typename Array1, typename Array2>
typename Array1::value_type dot(const Array1 &x, const Array2 &y) {
  typedef typename Array1::value_type T;
  detail::same_space<Array1, Array2>();
  detail::same_size(x, y, "dot");
  T r = T(0);
  if (detail::is_device<Array1>::value) {
    detail::require_abi<T>("dot");
    detail::check(detail::dot_((int64_t)x.size(), detail::cptr(x), detail::cptr(y), &r));
  } else {
    auto px = detail::cptr(x);
    auto py = detail::cptr(y);
    for (size_t i = 0; i < x.size(); ++i) r = r + px[i] * py[i];
  }
  return r;
}
// real value types: conj is the identity
template <typename Array1, typename Array2>
typename Array1::value_type dotc(const Array1 &x, const Array2 &y) {
  return dot(x, y);
}

template <typename Array>
typename Array::value_type nrm2(const Array &x) {
  typedef typename Array::value_type T;
  T r = T(0);
  if (detail::is_device<Array>::value) {
    detail::require_abi<T>("nrm2");
    detail::check(detail::nrm2_((int64_t)x.size(), detail::cptr(x), &r));
  } else {
    auto px = detail::cptr(x);
    for (size_t i = 0; i < x.size(); ++i) r = r + px[i] * px[i];
    r = (T)std::sqrt(r);
  }
  return r;
}

template <typename Array>
typename Array::value_type asum(const Array &x) {
  typedef typename Array::value_type T;
  T r = T(0);
  if (detail::is_device<Array>::value) {
    detail::require_abi<T>("asum");
    detail::check(detail::asum_((int64_t)x.size(), detail::cptr(x), &r));
  } else {
    auto px = detail::cptr(x);
    for (size_t i = 0; i < x.size(); ++i) r = r + detail::abs_of(px[i]);
  }
  return r;
}
template <typename Array>
typename Array::value_type nrm1(const Array &x) {
  return asum(x);
}

template <typename Array>
typename Array::value_type nrmmax(const Array &x) {
  typedef typename Array::value_type T;
  T r = T(0);
  if (detail::is_device<Array>::value) {
    detail::require_abi<T>("nrmmax");
    detail::check(detail::nrmmax_((int64_t)x.size(), detail::cptr(x), &r));
  } else {
    auto px = detail::cptr(x);
    for (size_t i = 0; i < x.size(); ++i) r = std::max(r, detail::abs_of(px[i]));
  }
  return r;
}

// index of the first element of maximal magnitude
template <typename Array>
int amax(const Array &x) {
  typedef typename Array::value_type T;
  int r = 0;
  if (detail::is_device<Array>::value) {
    detail::require_abi<T>("amax");
    detail::check(detail::amax_((int64_t)x.size(), detail::cptr(x), &r));
  } else {
    auto px = detail::cptr(x);
    T best = T(0);
    for (size_t i = 0; i < x.size(); ++i)
      if (i == 0 || detail::abs_of(px[i]) > best) {
        r = (int)i;
        best = detail::abs_of(px[i]);
      }
  }
  return r;
}

// ---- overloads with a leading execution policy ------------------------------
// Dispatched on the derived policy (cusp/memory.h: derived_cast): a user policy's own overload
// `void axpy(my_system&, ...)` is found by argument-dependent lookup (testing/blas.cu:700-1208,
// TestBlasDispatch); otherwise the call lands on adl_default::<name>, which forwards to the
// policy-free function above.
namespace detail {
namespace adl_default {
#define CUSP_B200_POLICY_DEFAULT_VOID(name)                                   \
  template <typename P, typename... Args>                                     \
  void name(cusp::execution_policy<P> &, Args &&... args) {                   \
    cusp::blas::name(std::forward<Args>(args)...);                            \
  }
CUSP_B200_POLICY_DEFAULT_VOID(axpy)
CUSP_B200_POLICY_DEFAULT_VOID(axpby)
CUSP_B200_POLICY_DEFAULT_VOID(axpbypcz)
CUSP_B200_POLICY_DEFAULT_VOID(xmy)
CUSP_B200_POLICY_DEFAULT_VOID(copy)
CUSP_B200_POLICY_DEFAULT_VOID(fill)
CUSP_B200_POLICY_DEFAULT_VOID(scal)
#undef CUSP_B200_POLICY_DEFAULT_VOID
template <typename P, typename Array1, typename Array2>
typename Array1::value_type dot(cusp::execution_policy<P> &, const Array1 &x, const Array2 &y) {
  return cusp::blas::dot(x, y);
}
template <typename P, typename Array1, typename Array2>
typename Array1::value_type dotc(cusp::execution_policy<P> &, const Array1 &x, const Array2 &y) {
  return cusp::blas::dot(x, y);
}
template <typename P, typename Array>
typename Array::value_type nrm2(cusp::execution_policy<P> &, const Array &x) {
  return cusp::blas::nrm2(x);
}
template <typename P, typename Array>
typename Array::value_type nrm1(cusp::execution_policy<P> &, const Array &x) {
  return cusp::blas::asum(x);
}
template <typename P, typename Array>
typename Array::value_type asum(cusp::execution_policy<P> &, const Array &x) {
  return cusp::blas::asum(x);
}
template <typename P, typename Array>
typename Array::value_type nrmmax(cusp::execution_policy<P> &, const Array &x) {
  return cusp::blas::nrmmax(x);
}
template <typename P, typename Array>
int amax(cusp::execution_policy<P> &, const Array &x) {
  return cusp::blas::amax(x);
}
}  // namespace adl_default
}  // namespace detail

#define CUSP_B200_POLICY_VOID(name)                                           \
  template <typename P, typename... Args>                                     \
  void name(const cusp::execution_policy<P> &exec, Args &&... args) {         \
    using detail::adl_default::name;                                          \
    cusp::detail::stream_scope<P> on_stream(cusp::detail::derived_cast(exec)); \
    name(cusp::detail::derived_cast(exec), std::forward<Args>(args)...);      \
  }
CUSP_B200_POLICY_VOID(axpy)
CUSP_B200_POLICY_VOID(axpby)
CUSP_B200_POLICY_VOID(axpbypcz)
CUSP_B200_POLICY_VOID(xmy)
CUSP_B200_POLICY_VOID(copy)
CUSP_B200_POLICY_VOID(fill)
CUSP_B200_POLICY_VOID(scal)
#undef CUSP_B200_POLICY_VOID

#define CUSP_B200_POLICY_VALUE(name)                                                          \
  template <typename P, typename... Args>                                                     \
  auto name(const cusp::execution_policy<P> &exec, Args &&... args) {                         \
    using detail::adl_default::name;                                                          \
    cusp::detail::stream_scope<P> on_stream(cusp::detail::derived_cast(exec));                \
    return name(cusp::detail::derived_cast(exec), std::forward<Args>(args)...);               \
  }
CUSP_B200_POLICY_VALUE(dot)
CUSP_B200_POLICY_VALUE(dotc)
CUSP_B200_POLICY_VALUE(nrm2)
CUSP_B200_POLICY_VALUE(nrm1)
CUSP_B200_POLICY_VALUE(asum)
CUSP_B200_POLICY_VALUE(nrmmax)
CUSP_B200_POLICY_VALUE(amax)
#undef CUSP_B200_POLICY_VALUE

// ---- BLAS-2/3: not provided by the generic back end either ------------------
template <typename... Args>
void gemv(Args &&...) {
  throw cusp::not_implemented_exception("CUSP GEMV not implemented");
}
template <typename... Args>
void gemm(Args &&...) {
  throw cusp::not_implemented_exception("CUSP GEMM not implemented");
}

}  // namespace blas
}  // namespace cusp
