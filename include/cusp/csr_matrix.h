// cusp/csr_matrix.h — cusp::csr_matrix / csr_matrix_view
// (reference: cusp/csr_matrix.h:107-210 container, :243-400 view).
#pragma once
#include "detail/matrix_base.h"

namespace cusp {

template <typename Array1, typename Array2, typename Array3, typename IndexType = typename Array1::value_type,
          typename ValueType = typename Array3::value_type, typename MemorySpace = typename Array1::memory_space>
class csr_matrix_view;

template <typename IndexType, typename ValueType, typename MemorySpace>
class csr_matrix : public detail::matrix_base<IndexType, ValueType, MemorySpace, csr_format> {
  typedef detail::matrix_base<IndexType, ValueType, MemorySpace, csr_format> Parent;

 public:
  typedef array1d<IndexType, MemorySpace> row_offsets_array_type;
  typedef array1d<IndexType, MemorySpace> column_indices_array_type;
  typedef array1d<ValueType, MemorySpace> values_array_type;
  typedef csr_matrix container;
  typedef csr_matrix_view<typename row_offsets_array_type::view, typename column_indices_array_type::view,
                          typename values_array_type::view, IndexType, ValueType, MemorySpace>
      view;
  typedef csr_matrix_view<typename row_offsets_array_type::const_view,
                          typename column_indices_array_type::const_view, typename values_array_type::const_view,
                          IndexType, ValueType, MemorySpace>
      const_view;
  template <typename Space>
  struct rebind {
    typedef csr_matrix<IndexType, ValueType, Space> type;
  };

  row_offsets_array_type row_offsets;
  column_indices_array_type column_indices;
  values_array_type values;

  csr_matrix() {}
  csr_matrix(size_t r, size_t c, size_t n) : Parent(r, c, n), row_offsets(r + 1), column_indices(n), values(n) {}
  template <typename MatrixType, typename = typename std::enable_if<detail::has_format<MatrixType>::value>::type>
  csr_matrix(const MatrixType &m) {
    cusp::convert(m, *this);
  }
  template <typename MatrixType, typename = typename std::enable_if<detail::has_format<MatrixType>::value>::type>
  csr_matrix &operator=(const MatrixType &m) {
    cusp::convert(m, *this);
    return *this;
  }
  void resize(size_t r, size_t c, size_t n) {
    Parent::resize(r, c, n);
    row_offsets.resize(r + 1);
    column_indices.resize(n);
    values.resize(n);
  }
  void swap(csr_matrix &o) {
    Parent::swap(o);
    row_offsets.swap(o.row_offsets);
    column_indices.swap(o.column_indices);
    values.swap(o.values);
  }
};

template <typename Array1, typename Array2, typename Array3, typename IndexType, typename ValueType,
          typename MemorySpace>
class csr_matrix_view : public detail::matrix_base<IndexType, ValueType, MemorySpace, csr_format> {
  typedef detail::matrix_base<IndexType, ValueType, MemorySpace, csr_format> Parent;

 public:
  typedef Array1 row_offsets_array_type;
  typedef Array2 column_indices_array_type;
  typedef Array3 values_array_type;
  typedef csr_matrix<IndexType, ValueType, MemorySpace> container;
  typedef csr_matrix_view view;

  Array1 row_offsets;
  Array2 column_indices;
  Array3 values;

  csr_matrix_view() {}
  csr_matrix_view(size_t r, size_t c, size_t n, const Array1 &ro, const Array2 &ci, const Array3 &v)
      : Parent(r, c, n), row_offsets(ro), column_indices(ci), values(v) {}
  template <typename Matrix, typename = typename std::enable_if<detail::has_format<Matrix>::value>::type>
  csr_matrix_view(Matrix &m)
      : Parent(m), row_offsets(m.row_offsets), column_indices(m.column_indices), values(m.values) {}

  void resize(size_t r, size_t c, size_t n) {
    Parent::resize(r, c, n);
    row_offsets.resize(r + 1);
    column_indices.resize(n);
    values.resize(n);
  }
};

template <typename Array1, typename Array2, typename Array3>
csr_matrix_view<Array1, Array2, Array3> make_csr_matrix_view(size_t r, size_t c, size_t n, const Array1 &ro,
                                                             const Array2 &ci, const Array3 &v) {
  return csr_matrix_view<Array1, Array2, Array3>(r, c, n, ro, ci, v);
}
template <typename I, typename V, typename S>
typename csr_matrix<I, V, S>::view make_csr_matrix_view(csr_matrix<I, V, S> &m) {
  return typename csr_matrix<I, V, S>::view(m);
}
template <typename I, typename V, typename S>
typename csr_matrix<I, V, S>::const_view make_csr_matrix_view(const csr_matrix<I, V, S> &m) {
  return typename csr_matrix<I, V, S>::const_view(m);
}
// a view of a view is the same view (csr_matrix_view.cu: "construct view from view")
template <typename A1, typename A2, typename A3, typename I, typename V, typename S>
csr_matrix_view<A1, A2, A3, I, V, S> make_csr_matrix_view(const csr_matrix_view<A1, A2, A3, I, V, S> &v) {
  return v;
}

}  // namespace cusp
#include "convert.h"
