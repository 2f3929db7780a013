// cusp::blas known answers — the reference's testing/blas.cu:
//   amax :8-27, axpy :60-92, axpby :97-142, axpbypcz :147-200, xmy :205-250,
//   copy :255-282, dot/dotc :287-352, fill :357-395, nrm1 :400-429, nrm2 :434-453,
//   nrmmax :458-483, scal :488-520; size mismatches throw invalid_input_exception.
#include <cusp/array1d.h>
#include <cusp/blas/blas.h>

#include "check.h"

template <class MemorySpace>
void TestAmax() {
  typedef cusp::array1d<float, MemorySpace> Array;
  typedef typename Array::view View;
  Array x(6);
  View view_x(x);
  x[0] = 0.0f; x[1] = -5.0f; x[2] = 4.0f; x[3] = -3.0f; x[4] = 7.0f; x[5] = 1.0f;
  ASSERT_EQUAL(cusp::blas::amax(x), 4);
  ASSERT_EQUAL(cusp::blas::amax(view_x), 4);
}
TEST_HOST_DEVICE(TestAmax)

template <class MemorySpace>
void TestAxpy() {
  typedef cusp::array1d<float, MemorySpace> Array;
  typedef typename Array::view View;
  Array x(4), y(4);
  x[0] = 7.0f; y[0] = 0.0f; x[1] = 5.0f; y[1] = -2.0f; x[2] = 4.0f; y[2] = 0.0f; x[3] = -3.0f; y[3] = 5.0f;
  cusp::blas::axpy(x, y, 2.0f);
  ASSERT_EQUAL(y[0], 14.0f); ASSERT_EQUAL(y[1], 8.0f); ASSERT_EQUAL(y[2], 8.0f); ASSERT_EQUAL(y[3], -1.0f);
  y[0] = 0.0f; y[1] = -2.0f; y[2] = 0.0f; y[3] = 5.0f;
  View vx(x), vy(y);
  cusp::blas::axpy(vx, vy, 2.0f);
  ASSERT_EQUAL(y[0], 14.0f); ASSERT_EQUAL(y[1], 8.0f); ASSERT_EQUAL(y[2], 8.0f); ASSERT_EQUAL(y[3], -1.0f);
  Array w(3);
  ASSERT_THROWS(cusp::blas::axpy(x, w, 1.0f), cusp::invalid_input_exception);
}
TEST_HOST_DEVICE(TestAxpy)

template <class MemorySpace>
void TestAxpby() {
  typedef cusp::array1d<float, MemorySpace> Array;
  typedef typename Array::view View;
  Array x(4), y(4), z(4, 0);
  x[0] = 7.0f; y[0] = 0.0f; x[1] = 5.0f; y[1] = -2.0f; x[2] = 4.0f; y[2] = 0.0f; x[3] = -3.0f; y[3] = 5.0f;
  cusp::blas::axpby(x, y, z, 2.0f, 1.0f);
  ASSERT_EQUAL(z[0], 14.0f); ASSERT_EQUAL(z[1], 8.0f); ASSERT_EQUAL(z[2], 8.0f); ASSERT_EQUAL(z[3], -1.0f);
  z[0] = z[1] = z[2] = z[3] = 0.0f;
  View vx(x), vy(y), vz(z);
  cusp::blas::axpby(vx, vy, vz, 2.0f, 1.0f);
  ASSERT_EQUAL(z[0], 14.0f); ASSERT_EQUAL(z[1], 8.0f); ASSERT_EQUAL(z[2], 8.0f); ASSERT_EQUAL(z[3], -1.0f);
  Array w(3);
  ASSERT_THROWS(cusp::blas::axpby(x, y, w, 2.0f, 1.0f), cusp::invalid_input_exception);
}
TEST_HOST_DEVICE(TestAxpby)

template <class MemorySpace>
void TestAxpbypcz() {
  typedef cusp::array1d<float, MemorySpace> Array;
  Array x(4), y(4), z(4), w(4, 0);
  x[0] = 7.0f; y[0] = 0.0f; z[0] = 1.0f;
  x[1] = 5.0f; y[1] = -2.0f; z[1] = 0.0f;
  x[2] = 4.0f; y[2] = 0.0f; z[2] = 3.0f;
  x[3] = -3.0f; y[3] = 5.0f; z[3] = -2.0f;
  cusp::blas::axpbypcz(x, y, z, w, 2.0f, 1.0f, 3.0f);
  ASSERT_EQUAL(w[0], 17.0f); ASSERT_EQUAL(w[1], 8.0f); ASSERT_EQUAL(w[2], 17.0f); ASSERT_EQUAL(w[3], -7.0f);
  Array output(3);
  ASSERT_THROWS(cusp::blas::axpbypcz(x, y, z, output, 2.0f, 1.0f, 3.0f), cusp::invalid_input_exception);
}
TEST_HOST_DEVICE(TestAxpbypcz)

template <class MemorySpace>
void TestXmy() {
  typedef cusp::array1d<float, MemorySpace> Array;
  Array x(4), y(4), z(4, 0);
  x[0] = 7.0f; y[0] = 0.0f; x[1] = 5.0f; y[1] = -2.0f; x[2] = 4.0f; y[2] = 0.0f; x[3] = -3.0f; y[3] = 5.0f;
  cusp::blas::xmy(x, y, z);
  ASSERT_EQUAL(z[0], 0.0f); ASSERT_EQUAL(z[1], -10.0f); ASSERT_EQUAL(z[2], 0.0f); ASSERT_EQUAL(z[3], -15.0f);
  Array output(3);
  ASSERT_THROWS(cusp::blas::xmy(x, y, output), cusp::invalid_input_exception);
}
TEST_HOST_DEVICE(TestXmy)

template <class MemorySpace>
void TestCopy() {
  typedef cusp::array1d<float, MemorySpace> Array;
  typedef typename Array::view View;
  Array x(4);
  x[0] = 7.0f; x[1] = 5.0f; x[2] = 4.0f; x[3] = -3.0f;
  {
    Array y(4, -1);
    cusp::blas::copy(x, y);
    ASSERT_EQUAL(x == y, true);
  }
  {
    Array y(4, -1);
    View vx(x), vy(y);
    cusp::blas::copy(vx, vy);
    ASSERT_EQUAL(x == y, true);
  }
  Array w(3);
  ASSERT_THROWS(cusp::blas::copy(w, x), cusp::invalid_input_exception);
}
TEST_HOST_DEVICE(TestCopy)

template <class MemorySpace>
void TestDot() {
  typedef cusp::array1d<float, MemorySpace> Array;
  typedef typename Array::view View;
  Array x(6), y(6);
  x[0] = 7.0f; y[0] = 0.0f; x[1] = 5.0f; y[1] = -2.0f; x[2] = 4.0f; y[2] = 0.0f;
  x[3] = -3.0f; y[3] = 5.0f; x[4] = 0.0f; y[4] = 6.0f; x[5] = 4.0f; y[5] = 1.0f;
  ASSERT_EQUAL(cusp::blas::dot(x, y), -21.0f);
  ASSERT_EQUAL(cusp::blas::dotc(x, y), -21.0f);
  ASSERT_EQUAL(cusp::blas::dot(View(x), View(y)), -21.0f);
  Array w(3);
  ASSERT_THROWS(cusp::blas::dot(x, w), cusp::invalid_input_exception);
}
TEST_HOST_DEVICE(TestDot)

template <class MemorySpace>
void TestFill() {
  typedef cusp::array1d<float, MemorySpace> Array;
  Array x(4);
  x[0] = 7.0f; x[1] = 5.0f; x[2] = 4.0f; x[3] = -3.0f;
  cusp::blas::fill(x, 2.0f);
  for (int i = 0; i < 4; ++i) ASSERT_EQUAL(x[i], 2.0f);
  typename Array::view v(x);
  cusp::blas::fill(v, 1.0f);
  for (int i = 0; i < 4; ++i) ASSERT_EQUAL(x[i], 1.0f);
}
TEST_HOST_DEVICE(TestFill)

template <class MemorySpace>
void TestNorms() {
  typedef cusp::array1d<float, MemorySpace> Array;
  Array x(6);
  x[0] = 7.0f; x[1] = 5.0f; x[2] = 4.0f; x[3] = -3.0f; x[4] = 0.0f; x[5] = 1.0f;
  ASSERT_EQUAL(cusp::blas::nrm1(x), 20.0f);
  ASSERT_EQUAL(cusp::blas::asum(x), 20.0f);
  ASSERT_EQUAL(cusp::blas::nrm2(x), 10.0f);
  ASSERT_EQUAL(cusp::blas::nrmmax(x), 7.0f);
  ASSERT_EQUAL(cusp::blas::nrm2(typename Array::view(x)), 10.0f);
}
TEST_HOST_DEVICE(TestNorms)

template <class MemorySpace>
void TestScal() {
  typedef cusp::array1d<float, MemorySpace> Array;
  Array x(6);
  x[0] = 7.0f; x[1] = 5.0f; x[2] = 4.0f; x[3] = -3.0f; x[4] = 0.0f; x[5] = 4.0f;
  cusp::blas::scal(x, 4.0f);
  ASSERT_EQUAL(x[0], 28.0f); ASSERT_EQUAL(x[1], 20.0f); ASSERT_EQUAL(x[2], 16.0f);
  ASSERT_EQUAL(x[3], -12.0f); ASSERT_EQUAL(x[4], 0.0f); ASSERT_EQUAL(x[5], 16.0f);
}
TEST_HOST_DEVICE(TestScal)

// the policy overloads reach the same code (testing/blas.cu:942-975 dispatch tests)
template <class MemorySpace>
void TestBlasWithPolicy() {
  typedef cusp::array1d<double, MemorySpace> Array;
  Array x(3, 2.0), y(3, 1.0);
  MemorySpace exec;
  cusp::blas::axpy(exec, x, y, 3.0);
  ASSERT_EQUAL(y[2], 7.0);
  ASSERT_EQUAL(cusp::blas::dot(exec, x, y), 42.0);
  ASSERT_EQUAL(cusp::blas::nrm2(exec, Array(4, 2.0)), 4.0);
}
TEST_HOST_DEVICE(TestBlasWithPolicy)
