// cusp/complex.h — the type-level part of the reference's cusp/complex.h (norm_type, abs / conj / norm for real
// scalars; cusp::complex = thrust::complex where Thrust is at hand, i.e. under nvcc).  The B200 engine computes in
// float / double only (complex is out of scope): device containers of complex values exist as types, every
// engine entry point refuses them with cusp::not_implemented_exception.  This header is what lets the reference's own
// test framework (testing/unittest/*.h) and user code that merely mentions cusp::complex compile against include/cusp.
#pragma once
#include <cmath>

#if !defined(CUSP_B200_NO_THRUST) && defined(__has_include)
#if __has_include(<thrust/complex.h>) && defined(__CUDACC__)
#include <thrust/complex.h>
#define CUSP_B200_HAVE_THRUST_COMPLEX 1
#endif
#endif

#ifdef __CUDACC__
#define CUSP_B200_HD __host__ __device__
#else
#define CUSP_B200_HD
#endif

namespace cusp {

template <typename T>
struct norm_type {
  typedef T type;
};

#ifdef CUSP_B200_HAVE_THRUST_COMPLEX
using thrust::complex;
template <typename T>
struct norm_type<thrust::complex<T>> {
  typedef T type;
};
using thrust::abs;
using thrust::conj;
using thrust::norm;
using thrust::sqrt;
#endif

template <typename T>
CUSP_B200_HD inline T conj(const T &z) {
  return z;
}
template <typename T>
CUSP_B200_HD inline typename norm_type<T>::type abs(const T &z) {
  return z > 0 ? z : -z;
}
template <typename T>
CUSP_B200_HD inline typename norm_type<T>::type norm(const T &z) {
  return cusp::abs(z);
}

}  // namespace cusp
