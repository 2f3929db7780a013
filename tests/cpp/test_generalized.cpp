// cusp::generalized_spmv — testing/generalized_spmv.cu:13-180 restated: z = y + A x with
// (multiplies, plus) for {coo,csr,dia,ell,hyb} x {host_memory,device_memory}: the known-answer
// 5x4 matrix, then poisson5pt / gallery::random matrices against cusp::multiply + y (exact:
// integer-valued data).  On device_memory the call is cusp::blas::copy + b200sp_spmv(accumulate); other functor
// triples (semirings) go to b200sp_spmv_generalized (TestGeneralizedFunctors).
#include <cusp/array2d.h>
#include <cusp/coo_matrix.h>
#include <cusp/csr_matrix.h>
#include <cusp/dia_matrix.h>
#include <cusp/ell_matrix.h>
#include <cusp/gallery/poisson.h>
#include <cusp/gallery/random.h>
#include <cusp/hyb_matrix.h>
#include <cusp/multiply.h>

#include <vector>

#include "check.h"

template <typename TestMatrix>
void GeneralizedSpMV() {
  typedef typename TestMatrix::value_type ValueType;
  typedef typename TestMatrix::memory_space MemorySpace;
  const bool is_dia = std::is_same<typename TestMatrix::format, cusp::dia_format>::value;
  {
    cusp::array2d<ValueType, cusp::host_memory> A(5, 4, ValueType(0));
    A(0, 0) = 13; A(0, 1) = 80; A(1, 1) = 27; A(2, 0) = 55; A(2, 2) = 24; A(2, 3) = 42;
    A(3, 1) = 69; A(3, 3) = 83; A(4, 2) = 27;
    TestMatrix test_matrix(A);
    cusp::array1d<ValueType, MemorySpace> x(4), y(5), z(5, -1);
    x[0] = 1; x[1] = 2; x[2] = 3; x[3] = 4;
    y[0] = 10; y[1] = 20; y[2] = 30; y[3] = 40; y[4] = 50;
    cusp::generalized_spmv(test_matrix, x, y, z, cusp::multiplies_function<ValueType>(),
                           cusp::plus_function<ValueType>());
    ASSERT_EQUAL((ValueType)z[0], (ValueType)183);
    ASSERT_EQUAL((ValueType)z[1], (ValueType)74);
    ASSERT_EQUAL((ValueType)z[2], (ValueType)325);
    ASSERT_EQUAL((ValueType)z[3], (ValueType)510);
    ASSERT_EQUAL((ValueType)z[4], (ValueType)131);
  }
  typedef cusp::coo_matrix<int, ValueType, cusp::host_memory> HostMatrix;
  std::vector<HostMatrix> matrices;
  const int grids[5][2] = {{5, 5}, {10, 10}, {117, 113}, {313, 444}, {876, 321}};
  for (auto &g : grids) {
    HostMatrix M;
    cusp::gallery::poisson5pt(M, g[0], g[1]);
    matrices.push_back(M);
  }
  if (!is_dia) {  // random patterns have no diagonal structure (the reference's DIA conversion refuses them)
    const int rnd[5][3] = {{21, 23, 5}, {45, 37, 15}, {129, 127, 40}, {355, 378, 234}, {512, 512, 276}};
    for (auto &r : rnd) {
      HostMatrix M;
      cusp::gallery::random(M, r[0], r[1], r[2]);
      matrices.push_back(M);
    }
  }
  for (size_t i = 0; i < matrices.size(); i++) {
    TestMatrix M(matrices[i]);
    cusp::array1d<ValueType, cusp::host_memory> xh(M.num_cols), yh(M.num_rows);
    for (size_t k = 0; k < xh.size(); ++k) xh[k] = (ValueType)((k * 7 + i) % 2);          // random_integers<bool>
    for (size_t k = 0; k < yh.size(); ++k) yh[k] = (ValueType)((int)((k * 13 + i) % 256) - 128);  // <char>
    cusp::array1d<ValueType, MemorySpace> x(xh), y(yh), z(M.num_rows, ValueType(-7));
    cusp::generalized_spmv(M, x, y, z, cusp::multiplies_function<ValueType>(), cusp::plus_function<ValueType>());
    cusp::array1d<ValueType, MemorySpace> reference(M.num_rows, ValueType(0));
    cusp::multiply(M, x, reference);
    cusp::blas::axpy(y, reference, ValueType(1));
    ASSERT_EQUAL(z, reference);
  }
}

// Functor triples other than (multiplies, plus): the reference's kernels are templated on them
// (generic/multiply/generalized_spmv.h:61-303); on device_memory the shim maps the functor TYPES to the codes of
// b200sp_spmv_generalized.  Device result == host loop result for every format (integer data: exact for plus,
// min / max are exact on any data); a functor the ABI cannot name throws not_implemented on the device only.
template <typename T>
struct my_combine {
  T operator()(const T &a, const T &b) const { return a * b + T(1); }
};
template <typename Matrix>
void GeneralizedFunctors() {
  typedef typename Matrix::value_type T;
  typedef typename Matrix::memory_space MemorySpace;
  const bool is_dia = std::is_same<typename Matrix::format, cusp::dia_format>::value;
  typedef cusp::coo_matrix<int, T, cusp::host_memory> HostMatrix;
  std::vector<HostMatrix> matrices;
  {
    HostMatrix M;
    cusp::gallery::poisson5pt(M, 37, 29);
    matrices.push_back(M);
  }
  if (!is_dia) {
    HostMatrix M;
    cusp::gallery::random(M, 355, 378, 2340);
    for (size_t k = 0; k < M.num_entries; ++k) M.values[k] = (T)((int)(k % 9) - 4);
    matrices.push_back(M);
  }
  for (size_t i = 0; i < matrices.size(); ++i) {
    typedef typename Matrix::template rebind<cusp::host_memory>::type HostSame;
    HostSame Mh(matrices[i]);
    Matrix M(matrices[i]);
    cusp::array1d<T, cusp::host_memory> xh(M.num_cols), yh(M.num_rows);
    for (size_t k = 0; k < xh.size(); ++k) xh[k] = (T)((int)((k * 7 + i) % 11) - 5);
    for (size_t k = 0; k < yh.size(); ++k) yh[k] = (T)((int)((k * 13 + i) % 17) - 8);
    cusp::array1d<T, MemorySpace> x(xh);
#define CHECK_TRIPLE(INIT, COMBINE, REDUCE)                \
  {                                                        \
    cusp::array1d<T, cusp::host_memory> want(yh);          \
    cusp::multiply(Mh, xh, want, INIT, COMBINE, REDUCE);   \
    cusp::array1d<T, MemorySpace> got(yh);                 \
    cusp::multiply(M, x, got, INIT, COMBINE, REDUCE);      \
    ASSERT_EQUAL(got, want);                               \
  }
    CHECK_TRIPLE(cusp::constant_functor<T>(T(1e30)), cusp::plus_function<T>(), cusp::minimum_function<T>())  // (min,+)
    CHECK_TRIPLE(cusp::identity_function<T>(), cusp::plus_function<T>(), cusp::minimum_function<T>())
    CHECK_TRIPLE(cusp::constant_functor<T>(T(-1e30)), cusp::multiplies_function<T>(), cusp::maximum_function<T>())  // (max,x)
    CHECK_TRIPLE(cusp::constant_functor<T>(T(-1e30)), cusp::minimum_function<T>(), cusp::maximum_function<T>())  // (max,min)
    CHECK_TRIPLE(cusp::constant_functor<T>(T(0)), cusp::project2nd_function<T>(), cusp::plus_function<T>())   // sum of x over the pattern
    CHECK_TRIPLE(cusp::constant_functor<T>(T(3)), cusp::multiplies_function<T>(), cusp::plus_function<T>())   // y = 3 + A x
    CHECK_TRIPLE(cusp::identity_function<T>(), cusp::maximum_function<T>(), cusp::plus_function<T>())
#undef CHECK_TRIPLE
    cusp::array1d<T, MemorySpace> y(yh);
    if (std::is_same<MemorySpace, cusp::device_memory>::value) {
      ASSERT_THROWS(cusp::multiply(M, x, y, cusp::identity_function<T>(), my_combine<T>(), cusp::plus_function<T>()),
                    cusp::not_implemented_exception);
    } else {
      cusp::multiply(M, x, y, cusp::identity_function<T>(), my_combine<T>(), cusp::plus_function<T>());  // host: any functor
    }
  }
}
template <class MemorySpace>
void TestGeneralizedFunctors() {
  GeneralizedFunctors<cusp::coo_matrix<int, float, MemorySpace>>();
  GeneralizedFunctors<cusp::csr_matrix<int, float, MemorySpace>>();
  GeneralizedFunctors<cusp::dia_matrix<int, float, MemorySpace>>();
  GeneralizedFunctors<cusp::ell_matrix<int, float, MemorySpace>>();
  GeneralizedFunctors<cusp::hyb_matrix<int, float, MemorySpace>>();
  GeneralizedFunctors<cusp::csr_matrix<int, double, MemorySpace>>();
  GeneralizedFunctors<cusp::coo_matrix<int, double, MemorySpace>>();
}
TEST_HOST_DEVICE(TestGeneralizedFunctors)

template <class MemorySpace>
void TestGeneralizedSpMV() {
  GeneralizedSpMV<cusp::coo_matrix<int, float, MemorySpace>>();
  GeneralizedSpMV<cusp::csr_matrix<int, float, MemorySpace>>();
  GeneralizedSpMV<cusp::dia_matrix<int, float, MemorySpace>>();
  GeneralizedSpMV<cusp::ell_matrix<int, float, MemorySpace>>();
  GeneralizedSpMV<cusp::hyb_matrix<int, float, MemorySpace>>();
  GeneralizedSpMV<cusp::csr_matrix<int, double, MemorySpace>>();
  GeneralizedSpMV<cusp::hyb_matrix<int, double, MemorySpace>>();
}
TEST_HOST_DEVICE(TestGeneralizedSpMV)

// cusp::multiply(csr_matrix, array2d, array2d): CSR x dense block
// (cuda/detail/multiply/csr_block_spmv.h:181-222, sequential/multiply/csr_block_spmv.h:52-77).
// Column j of the result must equal cusp::multiply(A, column j) exactly (same order of additions).
template <class MemorySpace>
void TestCsrBlockMultiply() {
  cusp::csr_matrix<int, double, cusp::host_memory> Ah;
  cusp::gallery::poisson5pt(Ah, 23, 17);
  for (size_t n = 0; n < Ah.num_entries; ++n) Ah.values[n] = (double)Ah.values[n] * (0.5 + 0.001 * (double)(n % 97));
  cusp::csr_matrix<int, double, MemorySpace> A(Ah);
  for (size_t k : {1u, 2u, 3u, 5u, 8u, 16u, 20u, 32u, 40u}) {
    cusp::array2d<double, cusp::host_memory> Xh(A.num_cols, k), Y0(A.num_rows, k, 3.0);
    for (size_t i = 0; i < A.num_cols; ++i)
      for (size_t j = 0; j < k; ++j) Xh(i, j) = 0.125 * (double)((i * 7 + j * 3) % 19) - 1.0;
    cusp::array2d<double, MemorySpace> X(Xh), Y(Y0);
    cusp::multiply(A, X, Y);
    cusp::array2d<double, cusp::host_memory> Yh(Y);
    for (size_t j = 0; j < k; ++j) {
      cusp::array1d<double, cusp::host_memory> xj(A.num_cols), yj(A.num_rows);
      for (size_t i = 0; i < A.num_cols; ++i) xj[i] = Xh(i, j);
      cusp::multiply(Ah, xj, yj);
      for (size_t i = 0; i < A.num_rows; ++i) ASSERT_EQUAL((double)Yh(i, j), (double)yj[i]);
    }
    // y += A x  (initialize = identity)
    cusp::array2d<double, MemorySpace> Z(Y0);
    cusp::multiply(A, X, Z, cusp::identity_function<double>(), cusp::multiplies_function<double>(),
                   cusp::plus_function<double>());
    cusp::array2d<double, cusp::host_memory> Zh(Z);
    cusp::array2d<double, cusp::host_memory> Wh(Y0);
    cusp::multiply(Ah, Xh, Wh, cusp::identity_function<double>(), cusp::multiplies_function<double>(),
                   cusp::plus_function<double>());
    ASSERT_EQUAL(Zh == Wh, true);
  }
  cusp::array2d<double, MemorySpace> Xbad(A.num_cols + 1, 2), Ybad(A.num_rows, 2);
  ASSERT_THROWS(cusp::multiply(A, Xbad, Ybad), cusp::invalid_input_exception);
  cusp::coo_matrix<int, double, MemorySpace> C(Ah);
  cusp::array2d<double, MemorySpace> X2(A.num_cols, 2), Y2(A.num_rows, 2);
  ASSERT_THROWS(cusp::multiply(C, X2, Y2), cusp::not_implemented_exception);
}
TEST_HOST_DEVICE(TestCsrBlockMultiply)
