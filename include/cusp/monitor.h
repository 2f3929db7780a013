// cusp/monitor.h — cusp::monitor<ValueType>: convergence test + residual history
// (reference: cusp/monitor.h:101-245, cusp/detail/monitor.inl:26-253).
//   finished(r): r_norm = nrm2(r); residuals.push_back(r_norm);
//                true when r_norm <= absolute + relative*||b|| or the iteration
//                limit is reached                      (monitor.inl:178-208, 107-111)
// The fused device CG (b200sp_cg) evaluates the same rule on the device and
// hands the history back through absorb().
#pragma once
#include <cmath>
#include <cstddef>
#include <iomanip>
#include <iostream>
#include <limits>
#include <vector>

#include "array1d.h"
#include "blas/blas.h"
#include "exception.h"

namespace cusp {

template <typename ValueType>
class monitor {
 public:
  typedef ValueType Real;  // real value types only (norm_type<ValueType>)

  template <typename VectorType>
  monitor(const VectorType &b, size_t iteration_limit = 500, Real relative_tolerance = 1e-5,
          Real absolute_tolerance = 0, bool verbose = false)
      : verbose(verbose),
        b_norm(cusp::blas::nrm2(b)),
        r_norm(std::numeric_limits<Real>::max()),
        iteration_limit_(iteration_limit),
        iteration_count_(0),
        relative_tolerance_(relative_tolerance),
        absolute_tolerance_(absolute_tolerance) {
    if (verbose) {
      std::cout << "Solver will continue until ";
      std::cout << "residual norm " << relative_tolerance << " or reaching ";
      std::cout << iteration_limit << " iterations " << std::endl;
      std::cout << "  Iteration Number  | Residual Norm" << std::endl;
    }
    residuals.reserve(iteration_limit);
  }

  void operator++() { ++iteration_count_; }
  bool converged() const { return residual_norm() <= tolerance(); }
  Real residual_norm() const { return r_norm; }
  size_t iteration_count() const { return iteration_count_; }
  size_t iteration_limit() const { return iteration_limit_; }
  Real relative_tolerance() const { return relative_tolerance_; }
  Real absolute_tolerance() const { return absolute_tolerance_; }
  Real tolerance() const { return absolute_tolerance() + relative_tolerance() * b_norm; }
  void set_verbose(bool v = true) { verbose = v; }
  bool is_verbose() { return verbose; }

  template <typename Vector>
  void reset(const Vector &b) {
    b_norm = cusp::blas::nrm2(b);
    r_norm = std::numeric_limits<Real>::max();
    iteration_count_ = 0;
    residuals.resize(0);
  }

  template <typename Vector>
  bool finished(const Vector &r) {
    r_norm = cusp::blas::nrm2(r);
    return record_and_test();
  }
  template <typename P, typename Vector>
  bool finished(const execution_policy<P> &, const Vector &r) {
    return finished(r);
  }

  void print() {
    if (iteration_count() == 0) {
      std::cout << "Monitor configured with " << tolerance() << " tolerance ";
      std::cout << "and iteration limit " << iteration_limit() << std::endl;
      return;
    }
    if (converged())
      std::cout << "Solver converged to " << tolerance() << " tolerance";
    else if (iteration_count() >= iteration_limit())
      std::cout << "Solver reached iteration limit " << iteration_limit() << " before converging";
    else
      throw cusp::runtime_exception("Monitor is in inconsistent state.");
    std::cout << " to (" << residual_norm() << " final residual)" << std::endl;
    std::cout << "Ran " << iteration_count();
    std::cout << " iterations with a final residual of ";
    std::cout << r_norm << std::endl;
    std::cout << "geometric convergence factor : " << geometric_rate() << std::endl;
    std::cout << "immediate convergence factor : " << immediate_rate() << std::endl;
    std::cout << "average convergence factor   : " << average_rate() << std::endl;
  }

  // convergence-rate summaries (monitor.inl:210-253)
  Real immediate_rate() {
    const size_t n = residuals.size();
    return n < 2 ? Real(0) : residuals[n - 1] / residuals[n - 2];
  }
  Real geometric_rate() {
    const size_t n = residuals.size();
    return n < 2 ? Real(0) : (Real)std::pow((double)(residuals[n - 1] / residuals[0]), 1.0 / (double)(n - 1));
  }
  Real average_rate() {
    const size_t n = residuals.size();
    if (n < 2) return Real(0);
    double s = 0;
    for (size_t i = 1; i < n; ++i) s += (double)(residuals[i] / residuals[i - 1]);
    return (Real)(s / (double)(n - 1));
  }

  // take over the outcome of a fused device solve: the history is exactly what
  // finished() would have recorded call by call
  void absorb(size_t iterations, const double *history, size_t num_residuals, double b_norm_from_solver) {
    (void)b_norm_from_solver;
    iteration_count_ = iterations;
    residuals.resize(0);
    for (size_t i = 0; i < num_residuals; ++i) {
      residuals.push_back((Real)history[i]);
      if (verbose) print_line(i, (Real)history[i]);
    }
    if (num_residuals) r_norm = (Real)history[num_residuals - 1];
    if (verbose) {
      if (converged())
        std::cout << "Successfully converged after " << iteration_count() << " iterations." << std::endl;
      else
        std::cout << "Failed to converge after " << iteration_count() << " iterations." << std::endl;
    }
  }

  bool verbose;
  std::vector<Real> residuals;  // cusp::array1d<Real, host_memory> in the reference; same operator[] / size()

 protected:
  void print_line(size_t it, Real r) {
    std::cout << "       " << std::setw(10) << it;
    std::cout << "       " << std::setw(10) << std::scientific << r << std::endl;
  }
  bool record_and_test() {
    residuals.push_back(r_norm);
    if (verbose) print_line(iteration_count(), residual_norm());
    if (converged()) {
      if (verbose) std::cout << "Successfully converged after " << iteration_count() << " iterations." << std::endl;
      return true;
    } else if (iteration_count() >= iteration_limit()) {
      if (verbose) std::cout << "Failed to converge after " << iteration_count() << " iterations." << std::endl;
      return true;
    }
    return false;
  }

  Real b_norm;
  Real r_norm;
  size_t iteration_limit_;
  size_t iteration_count_;
  Real relative_tolerance_;
  Real absolute_tolerance_;
};

}  // namespace cusp
