// cusp::krylov::cg + cusp::monitor — the reference's testing/cg.cu:46-99 on host
// and device, the fused device route against the generic operation-by-operation
// route (same iterate sequence, cg.inl:63-105), and monitor bookkeeping
// (monitor.inl:107-111,178-208).
#include <cusp/array2d.h>
#include <cusp/csr_matrix.h>
#include <cusp/dia_matrix.h>
#include <cusp/ell_matrix.h>
#include <cusp/gallery/poisson.h>
#include <cusp/hyb_matrix.h>
#include <cusp/krylov/cg.h>
#include <cusp/monitor.h>
#include <cusp/multiply.h>

#include "check.h"

template <class MemorySpace>
void TestConjugateGradient() {
  cusp::csr_matrix<int, float, MemorySpace> A;
  cusp::gallery::poisson5pt(A, 10, 10);
  cusp::array1d<float, MemorySpace> x(A.num_rows, 0.0f);
  cusp::array1d<float, MemorySpace> b(A.num_rows, 1.0f);
  cusp::monitor<float> monitor(b, 20, 1e-4);
  cusp::krylov::cg(A, x, b, monitor);
  cusp::array1d<float, MemorySpace> residual(A.num_rows, 0.0f);
  cusp::multiply(A, x, residual);
  cusp::blas::axpby(residual, b, residual, -1.0f, 1.0f);
  ASSERT_EQUAL(cusp::blas::nrm2(residual) < 1e-4 * cusp::blas::nrm2(b), true);
  ASSERT_EQUAL(monitor.residuals.size(), monitor.iteration_count() + 1);
}
TEST_HOST_DEVICE(TestConjugateGradient)

template <class MemorySpace>
void TestConjugateGradientZeroResidual() {
  cusp::array2d<float, MemorySpace> M(2, 2);
  M(0, 0) = 8; M(0, 1) = 0; M(1, 0) = 0; M(1, 1) = 4;
  cusp::csr_matrix<int, float, MemorySpace> A(M);
  cusp::array1d<float, MemorySpace> x(A.num_rows, 1.0f);
  cusp::array1d<float, MemorySpace> b(A.num_rows);
  cusp::multiply(A, x, b);
  cusp::monitor<float> monitor(b, 20, 0.0f);
  cusp::krylov::cg(A, x, b, monitor);
  cusp::array1d<float, MemorySpace> residual(A.num_rows, 0.0f);
  cusp::multiply(A, x, residual);
  cusp::blas::axpby(residual, b, residual, -1.0f, 1.0f);
  ASSERT_EQUAL(monitor.converged(), true);
  ASSERT_EQUAL(monitor.iteration_count(), (size_t)0);
  ASSERT_EQUAL(cusp::blas::nrm2(residual), 0.0f);
}
TEST_HOST_DEVICE(TestConjugateGradientZeroResidual)

// a monitor type the fused route does not know: forces the generic loop
template <typename Real>
struct counting_monitor : cusp::monitor<Real> {
  template <typename V>
  counting_monitor(const V &b, size_t limit, Real rel) : cusp::monitor<Real>(b, limit, rel) {}
};

// fused b200sp_cg == the generic loop over cusp::multiply / cusp::blas, per format
template <typename Matrix>
void CompareFusedAndGeneric(double hist_tol) {
  typedef typename Matrix::value_type V;
  Matrix A;
  cusp::gallery::poisson7pt(A, 12, 10, 8);
  cusp::array1d<V, cusp::device_memory> b(A.num_rows, V(1));
  cusp::array1d<V, cusp::device_memory> x1(A.num_rows, V(0)), x2(A.num_rows, V(0));
  cusp::monitor<V> m1(b, 60, V(1e-6));
  counting_monitor<V> m2(b, 60, V(1e-6));
  cusp::krylov::cg(A, x1, b, m1);  // fused
  cusp::krylov::cg(A, x2, b, m2);  // generic
  ASSERT_EQUAL(m1.iteration_count(), m2.iteration_count());
  ASSERT_EQUAL(m1.converged(), m2.converged());
  ASSERT_EQUAL(m1.residuals.size(), m2.residuals.size());
  for (size_t i = 0; i < m1.residuals.size(); ++i)
    ASSERT_TRUE(std::fabs((double)m1.residuals[i] - (double)m2.residuals[i]) <= hist_tol * (double)m2.residuals[0]);
  cusp::array1d<V, cusp::host_memory> h1(x1), h2(x2);
  for (size_t i = 0; i < h1.size(); ++i) ASSERT_NEAR(h1[i], h2[i], 100 * hist_tol * std::fabs((double)h2[i]) + 1e-30);
}
void TestCgFusedVsGeneric() {
  CompareFusedAndGeneric<cusp::csr_matrix<int, double, cusp::device_memory>>(1e-10);
  CompareFusedAndGeneric<cusp::dia_matrix<int, double, cusp::device_memory>>(1e-10);
  CompareFusedAndGeneric<cusp::ell_matrix<int, double, cusp::device_memory>>(1e-10);
  CompareFusedAndGeneric<cusp::hyb_matrix<int, double, cusp::device_memory>>(1e-10);
  CompareFusedAndGeneric<cusp::coo_matrix<int, double, cusp::device_memory>>(1e-10);
  CompareFusedAndGeneric<cusp::dia_matrix<int, float, cusp::device_memory>>(1e-4);
}
TEST_DEVICE(TestCgFusedVsGeneric)

// device CG reproduces the host CG history (same operation order per entry)
void TestCgDeviceVsHost() {
  cusp::csr_matrix<int, double, cusp::host_memory> Ah;
  cusp::gallery::poisson5pt(Ah, 24, 17);
  cusp::csr_matrix<int, double, cusp::device_memory> Ad(Ah);
  cusp::array1d<double, cusp::host_memory> bh(Ah.num_rows, 1.0), xh(Ah.num_rows, 0.0);
  cusp::array1d<double, cusp::device_memory> bd(bh), xd(xh);
  cusp::monitor<double> mh(bh, 200, 1e-9), md(bd, 200, 1e-9);
  cusp::krylov::cg(Ah, xh, bh, mh);
  cusp::krylov::cg(Ad, xd, bd, md);
  ASSERT_EQUAL(mh.iteration_count(), md.iteration_count());
  ASSERT_TRUE(md.converged());
  for (size_t i = 0; i < mh.residuals.size(); ++i)
    ASSERT_TRUE(std::fabs(mh.residuals[i] - md.residuals[i]) <= 1e-10 * mh.residuals[0]);
}
TEST_DEVICE(TestCgDeviceVsHost)

template <class MemorySpace>
void TestCgDefaultMonitorAndLimit() {
  cusp::csr_matrix<int, double, MemorySpace> A;
  cusp::gallery::poisson5pt(A, 16, 16);
  cusp::array1d<double, MemorySpace> x(A.num_rows, 0.0), b(A.num_rows, 1.0);
  cusp::krylov::cg(A, x, b);  // default monitor: 500 iterations, 1e-5 relative
  cusp::array1d<double, MemorySpace> r(A.num_rows);
  cusp::multiply(A, x, r);
  cusp::blas::axpby(r, b, r, -1.0, 1.0);
  ASSERT_TRUE(cusp::blas::nrm2(r) <= 1e-5 * cusp::blas::nrm2(b));
  // iteration limit reached: not converged, count == limit
  cusp::array1d<double, MemorySpace> x2(A.num_rows, 0.0);
  cusp::monitor<double> mon(b, 3, 1e-12);
  cusp::krylov::cg(A, x2, b, mon);
  ASSERT_EQUAL(mon.iteration_count(), (size_t)3);
  ASSERT_EQUAL(mon.converged(), false);
  ASSERT_EQUAL(mon.residuals.size(), (size_t)4);
  // monitor bookkeeping
  ASSERT_EQUAL(mon.iteration_limit(), (size_t)3);
  ASSERT_EQUAL(mon.tolerance(), 1e-12 * cusp::blas::nrm2(b));
  mon.reset(b);
  ASSERT_EQUAL(mon.iteration_count(), (size_t)0);
  ASSERT_EQUAL(mon.residuals.size(), (size_t)0);
}
TEST_HOST_DEVICE(TestCgDefaultMonitorAndLimit)

template <class MemorySpace>
void TestCgShapeErrors() {
  cusp::csr_matrix<int, float, MemorySpace> A(3, 4, 0);
  cusp::array1d<float, MemorySpace> x(3), b(3);
  ASSERT_THROWS(cusp::krylov::cg(A, x, b), cusp::invalid_input_exception);
}
TEST_HOST_DEVICE(TestCgShapeErrors)

// testing/monitor.cu:5-69 — the monitor on its own: tolerance = abs + rel * ||b||, finished() records the
// residual norm, the iteration limit ends the run, reset() re-arms with a new right-hand side
template <typename MemorySpace>
void TestMonitorSimple() {
  cusp::array1d<float, MemorySpace> b(2);
  b[0] = 10;
  b[1] = 0;
  cusp::array1d<float, MemorySpace> r(2);
  r[0] = 10;
  r[1] = 0;
  cusp::monitor<float> monitor(b, 5, 0.5, 1.0);
  ASSERT_EQUAL(monitor.finished(r), false);
  ASSERT_EQUAL(monitor.iteration_count(), (size_t)0);
  ASSERT_EQUAL(monitor.iteration_limit(), (size_t)5);
  ASSERT_EQUAL(monitor.relative_tolerance(), 0.5f);
  ASSERT_EQUAL(monitor.absolute_tolerance(), 1.0f);
  ASSERT_EQUAL(monitor.tolerance(), 6.0f);
  ++monitor;
  ASSERT_EQUAL(monitor.finished(r), false);
  ASSERT_EQUAL(monitor.iteration_count(), (size_t)1);
  ASSERT_EQUAL(monitor.residual_norm(), 10.0f);
  r[0] = 2;
  ASSERT_EQUAL(monitor.finished(r), true);
  ASSERT_EQUAL(monitor.iteration_count(), (size_t)1);
  ASSERT_EQUAL(monitor.residual_norm(), 2.0f);
  ASSERT_EQUAL(monitor.converged(), true);
  r[0] = 7;
  ASSERT_EQUAL(monitor.finished(r), false);
  ASSERT_EQUAL(monitor.residual_norm(), 7.0f);
  ++monitor;
  ASSERT_EQUAL(monitor.finished(r), false);
  ASSERT_EQUAL(monitor.iteration_count(), (size_t)2);
  ++monitor;
  ++monitor;
  ASSERT_EQUAL(monitor.finished(r), false);
  ASSERT_EQUAL(monitor.iteration_count(), (size_t)4);
  ++monitor;
  ASSERT_EQUAL(monitor.finished(r), true);  // iteration limit
  ASSERT_EQUAL(monitor.iteration_count(), (size_t)5);
  ASSERT_EQUAL(monitor.residual_norm(), 7.0f);
  ASSERT_EQUAL(monitor.converged(), false);
  monitor.reset(r);
  ASSERT_EQUAL(monitor.finished(r), false);
  ASSERT_EQUAL(monitor.iteration_count(), (size_t)0);
  ASSERT_EQUAL(monitor.residual_norm(), 7.0f);
}
static void TestMonitorSimpleHost() { TestMonitorSimple<cusp::host_memory>(); }
TEST_HOST(TestMonitorSimpleHost)
