// SURVEY §8(f) row 3 — further Krylov solvers and the Jacobi preconditioner over the same
// hot-path kernels: the reference's testing/cr.cu:35-87 (and the same protocol for
// bicgstab, which the reference tests inside testing/krylov programs), host and device,
// plus device == host iterate histories (same operation order per entry).
#include <cusp/array2d.h>
#include <cusp/csr_matrix.h>
#include <cusp/dia_matrix.h>
#include <cusp/gallery/poisson.h>
#include <cusp/krylov/bicgstab.h>
#include <cusp/krylov/cg.h>
#include <cusp/krylov/cr.h>
#include <cusp/krylov/gmres.h>
#include <cusp/monitor.h>
#include <cusp/multiply.h>
#include <cusp/precond/diagonal.h>

#include "check.h"

template <typename MemorySpace, typename Solver>
void solve_poisson(Solver solver) {
  cusp::csr_matrix<int, float, MemorySpace> A;
  cusp::gallery::poisson5pt(A, 10, 10);
  cusp::array1d<float, MemorySpace> x(A.num_rows, 0.0f), b(A.num_rows, 1.0f);
  cusp::monitor<float> monitor(b, 40, 1e-4);
  solver(A, x, b, monitor);
  cusp::array1d<float, MemorySpace> residual(A.num_rows, 0.0f);
  cusp::multiply(A, x, residual);
  cusp::blas::axpby(residual, b, residual, -1.0f, 1.0f);
  ASSERT_EQUAL(cusp::blas::nrm2(residual) < 1e-4 * cusp::blas::nrm2(b), true);
  ASSERT_TRUE(monitor.converged());
}

template <class MemorySpace>
void TestConjugateResidual() {
  solve_poisson<MemorySpace>([](auto &A, auto &x, auto &b, auto &m) { cusp::krylov::cr(A, x, b, m); });
}
TEST_HOST_DEVICE(TestConjugateResidual)

template <class MemorySpace>
void TestBiCGStab() {
  solve_poisson<MemorySpace>([](auto &A, auto &x, auto &b, auto &m) { cusp::krylov::bicgstab(A, x, b, m); });
}
TEST_HOST_DEVICE(TestBiCGStab)

template <class MemorySpace>
void TestConjugateResidualZeroResidual() {  // testing/cr.cu:60-87
  cusp::array2d<float, MemorySpace> M(2, 2);
  M(0, 0) = 8; M(0, 1) = 0; M(1, 0) = 0; M(1, 1) = 4;
  cusp::csr_matrix<int, float, MemorySpace> A(M);
  cusp::array1d<float, MemorySpace> x(A.num_rows, 1.0f), b(A.num_rows);
  cusp::multiply(A, x, b);
  cusp::monitor<float> monitor(b, 20, 0.0f);
  cusp::krylov::cr(A, x, b, monitor);
  cusp::array1d<float, MemorySpace> residual(A.num_rows, 0.0f);
  cusp::multiply(A, x, residual);
  cusp::blas::axpby(residual, b, residual, -1.0f, 1.0f);
  ASSERT_EQUAL(monitor.converged(), true);
  ASSERT_EQUAL(monitor.iteration_count(), (size_t)0);
  ASSERT_EQUAL(cusp::blas::nrm2(residual), 0.0f);
}
TEST_HOST_DEVICE(TestConjugateResidualZeroResidual)

// Jacobi-preconditioned CG on a badly scaled SPD system: D A D with D = diag(1 .. 1000)
template <class MemorySpace>
void TestDiagonalPreconditionedCg() {
  cusp::csr_matrix<int, double, cusp::host_memory> Ah;
  cusp::gallery::poisson5pt(Ah, 12, 9);
  const size_t n = Ah.num_rows;
  for (size_t i = 0; i < n; ++i)
    for (int k = Ah.row_offsets[i]; k < Ah.row_offsets[i + 1]; ++k) {
      const double di = 1.0 + 999.0 * (double)i / (double)(n - 1);
      const double dj = 1.0 + 999.0 * (double)Ah.column_indices[k] / (double)(n - 1);
      Ah.values[k] *= di * dj;
    }
  cusp::csr_matrix<int, double, MemorySpace> A(Ah);
  cusp::array1d<double, MemorySpace> b(n, 1.0), x0(n, 0.0), x1(n, 0.0);
  cusp::monitor<double> plain(b, 2000, 1e-10), jacobi(b, 2000, 1e-10);
  cusp::krylov::cg(A, x0, b, plain);
  cusp::precond::diagonal<double, MemorySpace> M(A);
  cusp::krylov::cg(A, x1, b, jacobi, M);
  ASSERT_TRUE(jacobi.converged());
  ASSERT_TRUE(jacobi.iteration_count() * 3 < plain.iteration_count());  // the scaling is what hurt plain CG
  cusp::array1d<double, MemorySpace> r(n);
  cusp::multiply(A, x1, r);
  cusp::blas::axpby(r, b, r, -1.0, 1.0);
  ASSERT_TRUE(cusp::blas::nrm2(r) <= 1e-9 * cusp::blas::nrm2(b));
  // the preconditioner itself: M r = r ./ diag(A)
  cusp::array1d<double, MemorySpace> ones(n, 1.0), z(n);
  M(ones, z);
  cusp::array1d<double, cusp::host_memory> zh(z);
  ASSERT_EQUAL(zh[0], 1.0 / 4.0);
}
TEST_HOST_DEVICE(TestDiagonalPreconditionedCg)

// device and host run the same operation sequence: equal iteration counts, histories within 1e-10
template <typename Solver>
void compare_device_and_host(Solver solver) {
  cusp::csr_matrix<int, double, cusp::host_memory> Ah;
  cusp::gallery::poisson5pt(Ah, 20, 15);
  cusp::csr_matrix<int, double, cusp::device_memory> Ad(Ah);
  cusp::array1d<double, cusp::host_memory> bh(Ah.num_rows, 1.0), xh(Ah.num_rows, 0.0);
  cusp::array1d<double, cusp::device_memory> bd(bh), xd(xh);
  cusp::monitor<double> mh(bh, 300, 1e-9), md(bd, 300, 1e-9);
  solver(Ah, xh, bh, mh);
  solver(Ad, xd, bd, md);
  ASSERT_EQUAL(mh.iteration_count(), md.iteration_count());
  ASSERT_TRUE(md.converged());
  ASSERT_EQUAL(mh.residuals.size(), md.residuals.size());
  for (size_t i = 0; i < mh.residuals.size(); ++i)
    ASSERT_TRUE(std::fabs(mh.residuals[i] - md.residuals[i]) <= 1e-9 * mh.residuals[0]);
}
void TestKrylovDeviceVsHost() {
  compare_device_and_host([](auto &A, auto &x, auto &b, auto &m) { cusp::krylov::cr(A, x, b, m); });
  compare_device_and_host([](auto &A, auto &x, auto &b, auto &m) { cusp::krylov::bicgstab(A, x, b, m); });
}
TEST_DEVICE(TestKrylovDeviceVsHost)


// testing/gmres.cu:39-62
template <class MemorySpace>
void TestGeneralizedMinRes() {
  const size_t restart = 20;
  cusp::csr_matrix<int, float, MemorySpace> A;
  cusp::gallery::poisson5pt(A, 10, 10);
  cusp::array1d<float, MemorySpace> x(A.num_rows, 0.0f), b(A.num_rows, 1.0f);
  cusp::monitor<float> monitor(b, 20, 1e-4);
  cusp::krylov::gmres(A, x, b, restart, monitor);
  cusp::array1d<float, MemorySpace> residual(A.num_rows, 0.0f);
  cusp::multiply(A, x, residual);
  cusp::blas::axpby(residual, b, residual, -1.0f, 1.0f);
  ASSERT_EQUAL(cusp::blas::nrm2(residual) < 1e-4 * cusp::blas::nrm2(b), true);
}
TEST_HOST_DEVICE(TestGeneralizedMinRes)

// restarts (restart < iterations needed), fp64, non-symmetric operator, Jacobi preconditioner
template <class MemorySpace>
void TestGeneralizedMinResRestartedNonSymmetric() {
  cusp::csr_matrix<int, double, cusp::host_memory> P;
  cusp::gallery::poisson5pt(P, 12, 9);
  for (size_t r = 0; r < P.num_rows; ++r)
    for (int e = P.row_offsets[r]; e < P.row_offsets[r + 1]; ++e) {
      if ((size_t)P.column_indices[e] == r + 1) P.values[e] = -1.5;  // convection-like skew, rows stay dominant
      if ((size_t)P.column_indices[e] + 1 == r) P.values[e] = -0.5;
      if ((size_t)P.column_indices[e] == r) P.values[e] = 4.0 + 0.01 * (double)(r % 7);
    }
  cusp::csr_matrix<int, double, MemorySpace> A(P);
  cusp::array1d<double, MemorySpace> x(A.num_rows, 0.0), b(A.num_rows, 1.0);
  cusp::monitor<double> monitor(b, 400, 1e-10);
  cusp::precond::diagonal<double, MemorySpace> M(A);
  cusp::krylov::gmres(A, x, b, 7, monitor, M);
  ASSERT_TRUE(monitor.converged());
  ASSERT_TRUE(monitor.iteration_count() > 7);  // restarted at least once
  cusp::array1d<double, MemorySpace> residual(A.num_rows, 0.0);
  cusp::multiply(A, x, residual);
  cusp::blas::axpby(residual, b, residual, -1.0, 1.0);
  ASSERT_TRUE(cusp::blas::nrm2(residual) < 1e-8 * cusp::blas::nrm2(b));
}
TEST_HOST_DEVICE(TestGeneralizedMinResRestartedNonSymmetric)

// array2d::column / row views alias the storage
void TestArray2dLines() {
  cusp::array2d<float, cusp::host_memory, cusp::column_major> C(3, 2, 0.0f);
  auto c1 = C.column(1);
  c1[2] = 5.0f;
  ASSERT_EQUAL(C(2, 1), 5.0f);
  ASSERT_EQUAL(c1.size(), (size_t)3);
  cusp::array2d<float, cusp::host_memory> Rm(2, 4, 0.0f);
  auto r1 = Rm.row(1);
  r1[3] = 7.0f;
  ASSERT_EQUAL(Rm(1, 3), 7.0f);
}
TEST_HOST(TestArray2dLines)
