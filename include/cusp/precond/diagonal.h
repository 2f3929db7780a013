// cusp/precond/diagonal.h — cusp::precond::diagonal<ValueType, MemorySpace>: Jacobi
// preconditioner M = diag(A)^-1 (reference: cusp/precond/diagonal.h,
// cusp/precond/detail/diagonal.inl:30-68).  Applying it is one cusp::blas::xmy
// (b200sp_xmy on the device); with it cusp::krylov::cg takes the generic
// operation-by-operation route over the same kernels.
#pragma once
#include <vector>

#include "../array1d.h"
#include "../blas/blas.h"
#include "../format_utils.h"
#include "../linear_operator.h"

namespace cusp {
namespace precond {

template <typename ValueType, typename MemorySpace>
class diagonal : public cusp::linear_operator<ValueType, MemorySpace> {
  typedef cusp::linear_operator<ValueType, MemorySpace> Parent;

 public:
  cusp::array1d<ValueType, MemorySpace> diagonal_reciprocals;

  diagonal() {}
  template <typename MatrixType>
  diagonal(const MatrixType &A) : Parent((int)A.num_rows, (int)A.num_cols, (int)A.num_rows) {
    cusp::array1d<ValueType, cusp::host_memory> d;
    cusp::extract_diagonal(A, d);  // setup time: through the host
    for (size_t i = 0; i < d.size(); ++i) d[i] = ValueType(1) / d[i];
    diagonal_reciprocals = d;
  }
  template <typename VectorType1, typename VectorType2>
  void operator()(const VectorType1 &x, VectorType2 &y) const {
    cusp::blas::xmy(diagonal_reciprocals, x, y);
  }
};

}  // namespace precond
}  // namespace cusp
