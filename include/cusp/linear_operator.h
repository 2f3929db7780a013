// cusp/linear_operator.h — cusp::linear_operator / identity_operator
// (reference: cusp/linear_operator.h:94-223).
#pragma once
#include "blas/blas.h"
#include "detail/matrix_base.h"

namespace cusp {

template <typename ValueType, typename MemorySpace, typename IndexType = int>
class linear_operator : public detail::matrix_base<IndexType, ValueType, MemorySpace, unknown_format> {
  typedef detail::matrix_base<IndexType, ValueType, MemorySpace, unknown_format> Parent;

 public:
  linear_operator() {}
  linear_operator(IndexType r, IndexType c) : Parent(r, c) {}
  linear_operator(IndexType r, IndexType c, IndexType n) : Parent(r, c, n) {}
};

template <typename ValueType, typename MemorySpace, typename IndexType = int>
class identity_operator : public linear_operator<ValueType, MemorySpace, IndexType> {
  typedef linear_operator<ValueType, MemorySpace, IndexType> Parent;

 public:
  identity_operator() {}
  identity_operator(IndexType r, IndexType c) : Parent(r, c) {}
  template <typename VectorType1, typename VectorType2>
  void operator()(const VectorType1 &x, VectorType2 &y) const {
    cusp::blas::copy(x, y);
  }
};

}  // namespace cusp
