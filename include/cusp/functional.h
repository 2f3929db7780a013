// cusp/functional.h — the functors the multiply front end recognises
// (reference: cusp/functional.h:426-439 constant_functor; thrust::identity /
// multiplies / plus are accepted when Thrust is available, std:: ones always).
#pragma once
#include <functional>
#include <type_traits>

#if !defined(CUSP_B200_NO_THRUST) && defined(__has_include)
#if __has_include(<thrust/functional.h>) && defined(__CUDACC__)
#include <thrust/functional.h>
#define CUSP_B200_HAVE_THRUST_FUNCTIONAL 1
#endif
#endif

namespace cusp {

template <typename T>
struct constant_functor {
  typedef T result_type;
  T val;
  constant_functor(const T &v = T()) : val(v) {}
  template <typename U>
  T operator()(const U &) const {
    return val;
  }
};
template <typename T>
struct identity_function {
  typedef T result_type;
  const T &operator()(const T &x) const { return x; }
};
template <typename T>
struct multiplies_function {
  T operator()(const T &a, const T &b) const { return a * b; }
};
template <typename T>
struct plus_function {
  T operator()(const T &a, const T &b) const { return a + b; }
};

namespace detail {
// which C-ABI `accumulate` value an initialize functor means: 0, 1 or -1 (none)
template <typename F>
struct init_kind {
  static int of(const F &) { return -1; }
};
template <typename T>
struct init_kind<constant_functor<T>> {
  static int of(const constant_functor<T> &f) { return f.val == T(0) ? 0 : -1; }
};
template <typename T>
struct init_kind<identity_function<T>> {
  static int of(const identity_function<T> &) { return 1; }
};
template <typename F>
struct is_multiplies : std::false_type {};
template <typename T>
struct is_multiplies<multiplies_function<T>> : std::true_type {};
template <typename T>
struct is_multiplies<std::multiplies<T>> : std::true_type {};
template <typename F>
struct is_plus : std::false_type {};
template <typename T>
struct is_plus<plus_function<T>> : std::true_type {};
template <typename T>
struct is_plus<std::plus<T>> : std::true_type {};
#ifdef CUSP_B200_HAVE_THRUST_FUNCTIONAL
template <typename T>
struct init_kind<thrust::identity<T>> {
  static int of(const thrust::identity<T> &) { return 1; }
};
template <typename T>
struct is_multiplies<thrust::multiplies<T>> : std::true_type {};
template <typename T>
struct is_plus<thrust::plus<T>> : std::true_type {};
#endif
}  // namespace detail
}  // namespace cusp
