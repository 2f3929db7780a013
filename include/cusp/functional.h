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

template <typename T>
struct minimum_function {
  T operator()(const T &a, const T &b) const { return b < a ? b : a; }
};
template <typename T>
struct maximum_function {
  T operator()(const T &a, const T &b) const { return a < b ? b : a; }
};
template <typename T>
struct project2nd_function {
  const T &operator()(const T &, const T &b) const { return b; }
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
// functor type -> code of b200sp_combine_op / b200sp_reduce_op (b200sp_spmv_generalized); -1: not expressible
template <typename F>
struct combine_code : std::integral_constant<int, -1> {};
template <typename F>
struct reduce_code : std::integral_constant<int, -1> {};
#define CUSP_B200_CODE(trait, F, code) \
  template <typename T>                \
  struct trait<F<T>> : std::integral_constant<int, code> {};
CUSP_B200_CODE(combine_code, multiplies_function, 0)
CUSP_B200_CODE(combine_code, std::multiplies, 0)
CUSP_B200_CODE(combine_code, plus_function, 1)
CUSP_B200_CODE(combine_code, std::plus, 1)
CUSP_B200_CODE(combine_code, minimum_function, 2)
CUSP_B200_CODE(combine_code, maximum_function, 3)
CUSP_B200_CODE(combine_code, project2nd_function, 4)
CUSP_B200_CODE(reduce_code, plus_function, 0)
CUSP_B200_CODE(reduce_code, std::plus, 0)
CUSP_B200_CODE(reduce_code, minimum_function, 1)
CUSP_B200_CODE(reduce_code, maximum_function, 2)
// initialize: constant_functor(c) -> (0, c), identity -> (1, 0); anything else -> (-1, 0)
template <typename F>
struct init_code {
  static int of(const F &, double &) { return -1; }
};
template <typename T>
struct init_code<constant_functor<T>> {
  static int of(const constant_functor<T> &f, double &v) {
    v = (double)f.val;
    return 0;
  }
};
template <typename T>
struct init_code<identity_function<T>> {
  static int of(const identity_function<T> &, double &) { return 1; }
};
#ifdef CUSP_B200_HAVE_THRUST_FUNCTIONAL
CUSP_B200_CODE(combine_code, thrust::multiplies, 0)
CUSP_B200_CODE(combine_code, thrust::plus, 1)
CUSP_B200_CODE(combine_code, thrust::minimum, 2)
CUSP_B200_CODE(combine_code, thrust::maximum, 3)
CUSP_B200_CODE(combine_code, thrust::project2nd, 4)
CUSP_B200_CODE(reduce_code, thrust::plus, 0)
CUSP_B200_CODE(reduce_code, thrust::minimum, 1)
CUSP_B200_CODE(reduce_code, thrust::maximum, 2)
template <typename T>
struct init_code<thrust::identity<T>> {
  static int of(const thrust::identity<T> &, double &) { return 1; }
};
#endif
#undef CUSP_B200_CODE
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
