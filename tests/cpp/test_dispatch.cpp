// Execution-policy dispatch ("user policies ... must still reach user-provided overloads
// untouched", unittest/special_types.h:107-139).  The reference's dispatch tests restated:
//   testing/multiply.cu:792-858  TestMatrixMatrixMultiplyDispatch / TestMatrixVectorMultiplyDispatch
//   testing/cg.cu:14-44          TestConjugateGradientDispatch
//   testing/gmres.cu:9-37        TestGeneralizedMinResDispatch
//   testing/convert.cu:662-690   TestConvertDispatch
//   testing/blas.cu:700-1208     TestBlasDispatch (the BLAS-1 functions)
// A user policy derives from cusp::execution_policy<Derived> (the reference: thrust::device_execution_policy);
// its overloads live in ITS namespace and are found by argument-dependent lookup from inside
// cusp::multiply(policy, ...) etc.  Without an overload the library's implementation runs.
#include <cusp/array1d.h>
#include <cusp/blas/blas.h>
#include <cusp/convert.h>
#include <cusp/csr_matrix.h>
#include <cusp/gallery/poisson.h>
#include <cusp/hyb_matrix.h>
#include <cusp/krylov/cg.h>
#include <cusp/krylov/gmres.h>
#include <cusp/monitor.h>
#include <cusp/multiply.h>

#include "check.h"

class my_system : public cusp::execution_policy<my_system> {
 public:
  my_system(int) : correctly_dispatched(false), num_copies(0) {}
  my_system(const my_system &other) : correctly_dispatched(false), num_copies(other.num_copies + 1) {}
  void validate_dispatch() { correctly_dispatched = (num_copies == 0); }  // reached, and not through a copy
  bool is_valid() { return correctly_dispatched; }

 private:
  bool correctly_dispatched;
  unsigned int num_copies;
  my_system();
};

// a policy WITHOUT overloads: everything falls through to the library
struct plain_system : cusp::execution_policy<plain_system> {};

template <typename MatrixType1, typename MatrixType2, typename MatrixType3>
void multiply(my_system &system, const MatrixType1 &, const MatrixType2 &, MatrixType3 &) {
  system.validate_dispatch();
}
template <typename LinearOperator, typename MatrixOrVector1, typename MatrixOrVector2, typename UnaryFunction,
          typename BinaryFunction1, typename BinaryFunction2>
void multiply(my_system &system, const LinearOperator &, const MatrixOrVector1 &, MatrixOrVector2 &, UnaryFunction,
              BinaryFunction1, BinaryFunction2) {
  system.validate_dispatch();
}
template <class LinearOperator, class VectorType1, class VectorType2, class Monitor, class Preconditioner>
void cg(my_system &system, const LinearOperator &, VectorType1 &, const VectorType2 &, Monitor &, Preconditioner &) {
  system.validate_dispatch();
}
template <class LinearOperator, class VectorType1, class VectorType2, class Monitor, class Preconditioner>
void gmres(my_system &system, const LinearOperator &, VectorType1 &, const VectorType2 &, const size_t, Monitor &,
           Preconditioner &) {
  system.validate_dispatch();
}
template <typename MatrixType1, typename MatrixType2>
void convert(my_system &system, const MatrixType1 &, MatrixType2 &) {
  system.validate_dispatch();
}
template <typename Array>
int amax(my_system &system, const Array &) {
  system.validate_dispatch();
  return 0;
}
template <typename Array>
typename Array::value_type asum(my_system &system, const Array &) {
  system.validate_dispatch();
  return 0;
}
template <typename Array1, typename Array2, typename ScalarType>
void axpy(my_system &system, const Array1 &, Array2 &, const ScalarType) {
  system.validate_dispatch();
}
template <typename Array1, typename Array2, typename Array3, typename ScalarType1, typename ScalarType2>
void axpby(my_system &system, const Array1 &, const Array2 &, Array3 &, ScalarType1, ScalarType2) {
  system.validate_dispatch();
}
template <typename Array1, typename Array2>
void copy(my_system &system, const Array1 &, Array2 &) {
  system.validate_dispatch();
}
template <typename Array1, typename Array2>
typename Array1::value_type dot(my_system &system, const Array1 &, const Array2 &) {
  system.validate_dispatch();
  return 0;
}
template <typename Array, typename ScalarType>
void fill(my_system &system, Array &, const ScalarType) {
  system.validate_dispatch();
}
template <typename Array>
typename Array::value_type nrm2(my_system &system, const Array &) {
  system.validate_dispatch();
  return 0;
}
template <typename Array, typename ScalarType>
void scal(my_system &system, Array &, const ScalarType) {
  system.validate_dispatch();
}

void TestMatrixMultiplyDispatch() {  // multiply.cu:792-858
  cusp::csr_matrix<int, float, cusp::host_memory> A, B, C;
  cusp::array1d<float, cusp::host_memory> x;
  {
    my_system sys(0);
    cusp::multiply(sys, A, B, C);
    ASSERT_EQUAL(true, sys.is_valid());
  }
  {
    my_system sys(0);
    cusp::multiply(sys, A, x, x);
    ASSERT_EQUAL(true, sys.is_valid());
  }
  {
    my_system sys(0);
    cusp::multiply(sys, A, x, x, cusp::constant_functor<float>(), cusp::multiplies_function<float>(),
                   cusp::plus_function<float>());
    ASSERT_EQUAL(true, sys.is_valid());
  }
}
TEST_HOST(TestMatrixMultiplyDispatch)

void TestKrylovDispatch() {  // cg.cu:14-44, gmres.cu:9-37
  cusp::csr_matrix<int, float, cusp::host_memory> A;
  cusp::gallery::poisson5pt(A, 10, 10);
  cusp::array1d<float, cusp::host_memory> x(A.num_rows, 0.0f);
  cusp::monitor<float> monitor(x, 20, 1e-4);
  cusp::identity_operator<float, cusp::host_memory> M(A.num_rows, A.num_cols);
  {
    my_system sys(0);
    cusp::krylov::cg(sys, A, x, x, monitor, M);
    ASSERT_EQUAL(true, sys.is_valid());
  }
  {
    my_system sys(0);
    cusp::krylov::gmres(sys, A, x, x, 20, monitor, M);
    ASSERT_EQUAL(true, sys.is_valid());
  }
}
TEST_HOST(TestKrylovDispatch)

void TestConvertDispatch() {  // convert.cu:662-690
  cusp::csr_matrix<int, float, cusp::host_memory> A;
  cusp::hyb_matrix<int, float, cusp::host_memory> B;
  my_system sys(0);
  cusp::convert(sys, A, B);
  ASSERT_EQUAL(true, sys.is_valid());
}
TEST_HOST(TestConvertDispatch)

void TestBlasDispatch() {  // blas.cu:942-1208 (BLAS-1)
  cusp::array1d<float, cusp::host_memory> x;
#define CHECK_DISPATCH(call)             \
  {                                      \
    my_system sys(0);                    \
    call;                                \
    ASSERT_EQUAL(true, sys.is_valid());  \
  }
  CHECK_DISPATCH(cusp::blas::amax(sys, x))
  CHECK_DISPATCH(cusp::blas::asum(sys, x))
  CHECK_DISPATCH(cusp::blas::axpy(sys, x, x, 1.0f))
  CHECK_DISPATCH(cusp::blas::axpby(sys, x, x, x, 1.0f, 1.0f))
  CHECK_DISPATCH(cusp::blas::copy(sys, x, x))
  CHECK_DISPATCH(cusp::blas::dot(sys, x, x))
  CHECK_DISPATCH(cusp::blas::fill(sys, x, 1.0f))
  CHECK_DISPATCH(cusp::blas::nrm2(sys, x))
  CHECK_DISPATCH(cusp::blas::scal(sys, x, 1.0f))
#undef CHECK_DISPATCH
}
TEST_HOST(TestBlasDispatch)

// a policy without overloads, and the built-in tags used as policies, run the library's implementation
template <class MemorySpace>
void TestPolicyFallsThroughToTheLibrary() {
  cusp::csr_matrix<int, double, MemorySpace> A;
  cusp::gallery::poisson5pt(A, 9, 8);
  cusp::array1d<double, MemorySpace> x(A.num_cols, 1.0), y0(A.num_rows, -1.0), y1(A.num_rows, -2.0), y2(A.num_rows, -3.0);
  cusp::multiply(A, x, y0);
  plain_system mine;
  cusp::multiply(mine, A, x, y1);
  cusp::multiply(MemorySpace(), A, x, y2);
  ASSERT_EQUAL(y0 == y1, true);
  ASSERT_EQUAL(y0 == y2, true);
  ASSERT_EQUAL(cusp::blas::dot(mine, x, y0), cusp::blas::dot(x, y0));
  ASSERT_EQUAL(cusp::blas::nrm2(MemorySpace(), y0), cusp::blas::nrm2(y0));
  cusp::blas::axpy(mine, x, y1, 2.0);
  cusp::blas::axpy(x, y2, 2.0);
  ASSERT_EQUAL(y1 == y2, true);
  cusp::array1d<double, MemorySpace> b(A.num_rows, 1.0), s0(A.num_rows, 0.0), s1(A.num_rows, 0.0);
  cusp::identity_operator<double, MemorySpace> M(A.num_rows, A.num_cols);
  cusp::monitor<double> m0(b, 50, 1e-10), m1(b, 50, 1e-10);
  cusp::krylov::cg(A, s0, b, m0, M);
  cusp::krylov::cg(mine, A, s1, b, m1, M);
  ASSERT_EQUAL(m0.iteration_count(), m1.iteration_count());
  ASSERT_EQUAL(s0 == s1, true);
  cusp::hyb_matrix<int, double, MemorySpace> H0, H1;
  cusp::convert(A, H0);
  cusp::convert(mine, A, H1);
  ASSERT_EQUAL(H0.ell.values.values == H1.ell.values.values, true);
}
static void TestPolicyFallsThroughToTheLibraryHost() { TestPolicyFallsThroughToTheLibrary<cusp::host_memory>(); }
TEST_HOST(TestPolicyFallsThroughToTheLibraryHost)
