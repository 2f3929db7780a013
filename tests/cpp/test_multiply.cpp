// cusp::multiply — the reference's SpMV unit tests, same inputs and the same
// exact-equality assertion:
//   testing/multiply.cu:383-512  y = A x   for container / matrix view / array view
//   testing/multiply.cu:514-645  y += A x  (initialize = identity)
// for {coo,csr,dia,ell,hyb}_matrix<int,float|double> x {host_memory,device_memory},
// plus ELL-R (cusp::ktt::ellr_matrix) and error behaviour.
#include <cusp/array2d.h>
#include <cusp/coo_matrix.h>
#include <cusp/csr_matrix.h>
#include <cusp/dia_matrix.h>
#include <cusp/ell_matrix.h>
#include <cusp/gallery/poisson.h>
#include <cusp/hyb_matrix.h>
#include <cusp/ktt/ellr_matrix.h>
#include <cusp/ktt/ktt.h>
#include <cusp/multiply.h>

#include "check.h"

template <typename SparseMatrixType, typename DenseMatrixType>
void CompareSparseMatrixVectorMultiply(const DenseMatrixType &A) {
  typedef typename SparseMatrixType::value_type ValueType;
  typedef typename SparseMatrixType::memory_space MemorySpace;

  cusp::array1d<ValueType, cusp::host_memory> x(A.num_cols);
  cusp::array1d<ValueType, cusp::host_memory> y(A.num_rows, 10);
  for (size_t i = 0; i < x.size(); i++) x[i] = i % 10;
  cusp::multiply(A, x, y);  // dense host product = expected

  {  // container
    SparseMatrixType _A(A);
    cusp::array1d<ValueType, MemorySpace> _x(x);
    cusp::array1d<ValueType, MemorySpace> _y(A.num_rows, 10);
    cusp::multiply(_A, _x, _y);
    ASSERT_EQUAL(_y, y);
  }
  {  // matrix view
    SparseMatrixType _A(A);
    cusp::array1d<ValueType, MemorySpace> _x(x);
    cusp::array1d<ValueType, MemorySpace> _y(A.num_rows, 10);
    typename SparseMatrixType::view _V(_A);
    cusp::multiply(_V, _x, _y);
    ASSERT_EQUAL(_y, y);
  }
  {  // array views
    SparseMatrixType _A(A);
    cusp::array1d<ValueType, MemorySpace> _x(x);
    cusp::array1d<ValueType, MemorySpace> _y(A.num_rows, 10);
    typename cusp::array1d<ValueType, MemorySpace>::view _Vx(_x), _Vy(_y);
    cusp::multiply(_A, _Vx, _Vy);
    ASSERT_EQUAL(_Vy, y);
  }
  {  // explicit policy
    SparseMatrixType _A(A);
    cusp::array1d<ValueType, MemorySpace> _x(x);
    cusp::array1d<ValueType, MemorySpace> _y(A.num_rows, 10);
    MemorySpace exec;
    cusp::multiply(exec, _A, _x, _y);
    ASSERT_EQUAL(_y, y);
  }
}

template <typename SparseMatrixType, typename DenseMatrixType>
void CompareScaledSparseMatrixVectorMultiply(const DenseMatrixType &A) {
  typedef typename SparseMatrixType::value_type ValueType;
  typedef typename SparseMatrixType::memory_space MemorySpace;

  cusp::array1d<ValueType, cusp::host_memory> x(A.num_cols);
  cusp::array1d<ValueType, cusp::host_memory> y(A.num_rows, 10);
  for (size_t i = 0; i < x.size(); i++) x[i] = i % 10;
  cusp::identity_function<ValueType> initialize;
  cusp::multiplies_function<ValueType> combine;
  cusp::plus_function<ValueType> reduce;
  cusp::multiply(A, x, y, initialize, combine, reduce);  // y = 10 + A x

  SparseMatrixType _A(A);
  cusp::array1d<ValueType, MemorySpace> _x(x);
  cusp::array1d<ValueType, MemorySpace> _y(A.num_rows, 10);
  cusp::multiply(_A, _x, _y, initialize, combine, reduce);
  ASSERT_EQUAL(_y, y);

  // std:: functors and the explicit zero-initialiser are recognised too
  cusp::array1d<ValueType, MemorySpace> _z(A.num_rows, 10);
  cusp::multiply(_A, _x, _z, cusp::constant_functor<ValueType>(0), std::multiplies<ValueType>(),
                 std::plus<ValueType>());
  cusp::array1d<ValueType, cusp::host_memory> z(A.num_rows, 10);
  cusp::multiply(A, x, z);
  ASSERT_EQUAL(_z, z);
}

template <typename ValueType>
struct TestMatrices {
  typedef cusp::array2d<ValueType, cusp::host_memory> Dense;
  Dense A, B, C, D, E, F, G, H;
  TestMatrices() : A(5, 4), B(2, 4), C(2, 2), D(2, 1), E(2, 2), F(2, 3) {
    const ValueType a[5][4] = {{13, 80, 0, 0}, {0, 27, 0, 0}, {55, 0, 24, 42}, {0, 69, 0, 83}, {0, 0, 27, 0}};
    for (int i = 0; i < 5; ++i)
      for (int j = 0; j < 4; ++j) A(i, j) = a[i][j];
    const ValueType b[2][4] = {{0, 2, 3, 4}, {5, 0, 0, 8}};
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 4; ++j) B(i, j) = b[i][j];
    C(0, 0) = 0; C(0, 1) = 0; C(1, 0) = 3; C(1, 1) = 5;
    D(0, 0) = 2; D(1, 0) = 3;
    E(0, 0) = 0; E(0, 1) = 0; E(1, 0) = 0; E(1, 1) = 0;
    F(0, 0) = 0; F(0, 1) = 1.5; F(0, 2) = 3.0; F(1, 0) = 0.5; F(1, 1) = 0; F(1, 2) = 0;
    cusp::gallery::poisson5pt(G, 4, 6);
    cusp::gallery::poisson5pt(H, 8, 3);
  }
};

template <class TestMatrix>
void TestSparseMatrixVectorMultiply() {
  TestMatrices<typename TestMatrix::value_type> m;
  CompareSparseMatrixVectorMultiply<TestMatrix>(m.A);
  CompareSparseMatrixVectorMultiply<TestMatrix>(m.B);
  CompareSparseMatrixVectorMultiply<TestMatrix>(m.C);
  CompareSparseMatrixVectorMultiply<TestMatrix>(m.D);
  CompareSparseMatrixVectorMultiply<TestMatrix>(m.E);
  CompareSparseMatrixVectorMultiply<TestMatrix>(m.F);
  CompareSparseMatrixVectorMultiply<TestMatrix>(m.G);
  CompareSparseMatrixVectorMultiply<TestMatrix>(m.H);
}
template <class TestMatrix>
void TestScaledSparseMatrixVectorMultiply() {
  TestMatrices<typename TestMatrix::value_type> m;
  CompareScaledSparseMatrixVectorMultiply<TestMatrix>(m.A);
  CompareScaledSparseMatrixVectorMultiply<TestMatrix>(m.B);
  CompareScaledSparseMatrixVectorMultiply<TestMatrix>(m.C);
  CompareScaledSparseMatrixVectorMultiply<TestMatrix>(m.D);
  CompareScaledSparseMatrixVectorMultiply<TestMatrix>(m.E);
  CompareScaledSparseMatrixVectorMultiply<TestMatrix>(m.F);
  CompareScaledSparseMatrixVectorMultiply<TestMatrix>(m.G);
  CompareScaledSparseMatrixVectorMultiply<TestMatrix>(m.H);
}

// DECLARE_SPARSE_MATRIX_UNITTEST (testing/unittest/matrix.h:27-61) + fp64
#define SPARSE_CASES(VTEST, V, vname)                                                                       \
  static check::registrar r_##VTEST##vname##CooH(#VTEST "<coo," #V ",host>", false,                         \
                                                 VTEST<cusp::coo_matrix<int, V, cusp::host_memory>>);       \
  static check::registrar r_##VTEST##vname##CsrH(#VTEST "<csr," #V ",host>", false,                         \
                                                 VTEST<cusp::csr_matrix<int, V, cusp::host_memory>>);       \
  static check::registrar r_##VTEST##vname##DiaH(#VTEST "<dia," #V ",host>", false,                         \
                                                 VTEST<cusp::dia_matrix<int, V, cusp::host_memory>>);       \
  static check::registrar r_##VTEST##vname##EllH(#VTEST "<ell," #V ",host>", false,                         \
                                                 VTEST<cusp::ell_matrix<int, V, cusp::host_memory>>);       \
  static check::registrar r_##VTEST##vname##HybH(#VTEST "<hyb," #V ",host>", false,                         \
                                                 VTEST<cusp::hyb_matrix<int, V, cusp::host_memory>>);       \
  static check::registrar r_##VTEST##vname##CooD(#VTEST "<coo," #V ",device>", true,                        \
                                                 VTEST<cusp::coo_matrix<int, V, cusp::device_memory>>);     \
  static check::registrar r_##VTEST##vname##CsrD(#VTEST "<csr," #V ",device>", true,                        \
                                                 VTEST<cusp::csr_matrix<int, V, cusp::device_memory>>);     \
  static check::registrar r_##VTEST##vname##DiaD(#VTEST "<dia," #V ",device>", true,                        \
                                                 VTEST<cusp::dia_matrix<int, V, cusp::device_memory>>);     \
  static check::registrar r_##VTEST##vname##EllD(#VTEST "<ell," #V ",device>", true,                        \
                                                 VTEST<cusp::ell_matrix<int, V, cusp::device_memory>>);     \
  static check::registrar r_##VTEST##vname##HybD(#VTEST "<hyb," #V ",device>", true,                        \
                                                 VTEST<cusp::hyb_matrix<int, V, cusp::device_memory>>);
SPARSE_CASES(TestSparseMatrixVectorMultiply, float, f32)
SPARSE_CASES(TestSparseMatrixVectorMultiply, double, f64)
SPARSE_CASES(TestScaledSparseMatrixVectorMultiply, float, f32)
SPARSE_CASES(TestScaledSparseMatrixVectorMultiply, double, f64)

// ELL-R takes the ELL route of cusp::multiply and the ELL-R kernel underneath
template <typename MemorySpace>
void TestEllrMultiply() {
  TestMatrices<float> m;
  typedef cusp::ktt::ellr_matrix<int, float, MemorySpace> Ellr;
  Ellr A(m.A);
  cusp::array1d<int, cusp::host_memory> lengths(A.row_lengths);
  ASSERT_EQUAL(lengths.size(), (size_t)5);
  ASSERT_EQUAL(lengths[0], 2); ASSERT_EQUAL(lengths[1], 1); ASSERT_EQUAL(lengths[2], 3);
  ASSERT_EQUAL(lengths[3], 2); ASSERT_EQUAL(lengths[4], 1);
  CompareSparseMatrixVectorMultiply<Ellr>(m.A);
  CompareSparseMatrixVectorMultiply<Ellr>(m.G);
}
TEST_HOST_DEVICE(TestEllrMultiply)

// with ktt disabled, ELL / DIA take the plain default kernels (ktt.cu:173-181)
void TestMultiplyKttDisabled() {
  TestMatrices<float> m;
  cusp::ktt::disable();
  CompareSparseMatrixVectorMultiply<cusp::dia_matrix<int, float, cusp::device_memory>>(m.G);
  CompareSparseMatrixVectorMultiply<cusp::ell_matrix<int, float, cusp::device_memory>>(m.G);
  cusp::ktt::enable();
  // enabled: repeated calls walk the tuning space (one configuration per call), result unchanged
  for (int rep = 0; rep < 40; ++rep)
    CompareSparseMatrixVectorMultiply<cusp::dia_matrix<int, float, cusp::device_memory>>(m.H);
}
TEST_DEVICE(TestMultiplyKttDisabled)

template <typename MemorySpace>
void TestMultiplyErrors() {
  cusp::csr_matrix<int, float, MemorySpace> A(3, 4, 0);
  cusp::array1d<float, MemorySpace> x(3), y(3);
  ASSERT_THROWS(cusp::multiply(A, x, y), cusp::invalid_input_exception);
}
TEST_HOST_DEVICE(TestMultiplyErrors)

// a functor the C ABI cannot name (user code) must fail loudly on the device; the named ones run there
// (b200sp_spmv_generalized) and give the host loop's result
template <typename T>
struct user_combine {
  T operator()(const T &a, const T &b) const { return a * b + T(1); }
};
void TestDeviceUnsupportedFunctors() {
  TestMatrices<float> m;
  cusp::csr_matrix<int, float, cusp::device_memory> A(m.A);
  cusp::array1d<float, cusp::device_memory> x(4, 1.0f), y(5, 0.0f);
  ASSERT_THROWS(cusp::multiply(A, x, y, cusp::constant_functor<float>(1.0f), user_combine<float>(),
                               cusp::plus_function<float>()),
                cusp::not_implemented_exception);
  // the host path is functor-generic (generalized SpMV, testing/generalized_spmv.cu)
  cusp::csr_matrix<int, float, cusp::host_memory> Ah(m.A);
  cusp::array1d<float, cusp::host_memory> xh(4, 1.0f), yh(5, 0.0f);
  cusp::multiply(Ah, xh, yh, cusp::constant_functor<float>(1.0f), cusp::multiplies_function<float>(),
                 cusp::plus_function<float>());
  ASSERT_EQUAL(yh[0], 94.0f);
  // constant_functor(1) is expressible by code: same result on the device
  cusp::multiply(A, x, y, cusp::constant_functor<float>(1.0f), cusp::multiplies_function<float>(),
                 cusp::plus_function<float>());
  cusp::array1d<float, cusp::host_memory> yd(y);
  ASSERT_EQUAL(yd[0], 94.0f);
  ASSERT_EQUAL(yd == yh, true);
}
TEST_DEVICE(TestDeviceUnsupportedFunctors)

// sparse x sparse is outside the hot path: a clear exception, not a compile error inside the shape check
template <class MemorySpace>
void TestSparseTimesSparseIsRefused() {
  TestMatrices<float> m;
  cusp::csr_matrix<int, float, MemorySpace> A(m.A), B(m.A), C;
  ASSERT_THROWS(cusp::multiply(A, B, C), cusp::not_implemented_exception);
}
TEST_HOST_DEVICE(TestSparseTimesSparseIsRefused)
