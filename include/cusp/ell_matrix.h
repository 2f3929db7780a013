// cusp/ell_matrix.h — cusp::ell_matrix / ell_matrix_view
// (reference: cusp/ell_matrix.h:119-229, cusp/detail/ell_matrix.inl:28-85).
// column_indices / values are column-major array2d with a shared pitch: slot
// (row, n) at [n*pitch + row]; padding: column invalid_index (-1), value 0.
#pragma once
#include "detail/matrix_base.h"

namespace cusp {

template <typename Array1, typename Array2, typename IndexType = typename Array1::value_type,
          typename ValueType = typename Array2::value_type, typename MemorySpace = typename Array1::memory_space>
class ell_matrix_view;

template <typename IndexType, typename ValueType, typename MemorySpace>
class ell_matrix : public detail::matrix_base<IndexType, ValueType, MemorySpace, ell_format> {
  typedef detail::matrix_base<IndexType, ValueType, MemorySpace, ell_format> Parent;

 public:
  typedef array2d<IndexType, MemorySpace, column_major> column_indices_array_type;
  typedef array2d<ValueType, MemorySpace, column_major> values_array_type;
  typedef ell_matrix container;
  typedef ell_matrix_view<typename column_indices_array_type::view, typename values_array_type::view, IndexType,
                          ValueType, MemorySpace>
      view;
  typedef ell_matrix_view<typename column_indices_array_type::const_view, typename values_array_type::const_view,
                          IndexType, ValueType, MemorySpace>
      const_view;
  template <typename Space>
  struct rebind {
    typedef ell_matrix<IndexType, ValueType, Space> type;
  };
  static const IndexType invalid_index = static_cast<IndexType>(-1);

  column_indices_array_type column_indices;
  values_array_type values;

  ell_matrix() {}
  ell_matrix(size_t r, size_t c, size_t n, size_t num_entries_per_row, size_t alignment = 32) : Parent(r, c, n) {
    column_indices.resize(r, num_entries_per_row, detail::round_up(r, alignment));
    values.resize(r, num_entries_per_row, detail::round_up(r, alignment));
  }
  template <typename MatrixType, typename = typename std::enable_if<detail::has_format<MatrixType>::value>::type>
  ell_matrix(const MatrixType &m) {
    cusp::convert(m, *this);
  }
  template <typename MatrixType, typename = typename std::enable_if<detail::has_format<MatrixType>::value>::type>
  ell_matrix &operator=(const MatrixType &m) {
    cusp::convert(m, *this);
    return *this;
  }
  void resize(size_t r, size_t c, size_t n, size_t num_entries_per_row) {
    Parent::resize(r, c, n);
    column_indices.resize(r, num_entries_per_row);
    values.resize(r, num_entries_per_row);
  }
  void resize(size_t r, size_t c, size_t n, size_t num_entries_per_row, size_t alignment) {
    Parent::resize(r, c, n);
    column_indices.resize(r, num_entries_per_row, detail::round_up(r, alignment));
    values.resize(r, num_entries_per_row, detail::round_up(r, alignment));
  }
  void swap(ell_matrix &o) {
    Parent::swap(o);
    column_indices.swap(o.column_indices);
    values.swap(o.values);
  }
};
template <typename I, typename V, typename S>
const I ell_matrix<I, V, S>::invalid_index;

template <typename Array1, typename Array2, typename IndexType, typename ValueType, typename MemorySpace>
class ell_matrix_view : public detail::matrix_base<IndexType, ValueType, MemorySpace, ell_format> {
  typedef detail::matrix_base<IndexType, ValueType, MemorySpace, ell_format> Parent;

 public:
  typedef Array1 column_indices_array_type;
  typedef Array2 values_array_type;
  typedef ell_matrix<IndexType, ValueType, MemorySpace> container;
  typedef ell_matrix_view view;
  static const IndexType invalid_index = static_cast<IndexType>(-1);

  Array1 column_indices;
  Array2 values;

  ell_matrix_view() {}
  ell_matrix_view(size_t r, size_t c, size_t n, const Array1 &ci, const Array2 &v)
      : Parent(r, c, n), column_indices(ci), values(v) {}
  template <typename Matrix, typename = typename std::enable_if<detail::has_format<Matrix>::value>::type>
  ell_matrix_view(Matrix &m) : Parent(m), column_indices(m.column_indices), values(m.values) {}

  void resize(size_t r, size_t c, size_t n, size_t num_entries_per_row) {
    Parent::resize(r, c, n);
    column_indices.resize(r, num_entries_per_row);
    values.resize(r, num_entries_per_row);
  }
};
template <typename A1, typename A2, typename I, typename V, typename S>
const I ell_matrix_view<A1, A2, I, V, S>::invalid_index;

template <typename Array1, typename Array2>
ell_matrix_view<Array1, Array2> make_ell_matrix_view(size_t r, size_t c, size_t n, const Array1 &ci,
                                                     const Array2 &v) {
  return ell_matrix_view<Array1, Array2>(r, c, n, ci, v);
}
template <typename I, typename V, typename S>
typename ell_matrix<I, V, S>::view make_ell_matrix_view(ell_matrix<I, V, S> &m) {
  return typename ell_matrix<I, V, S>::view(m);
}
template <typename I, typename V, typename S>
typename ell_matrix<I, V, S>::const_view make_ell_matrix_view(const ell_matrix<I, V, S> &m) {
  return typename ell_matrix<I, V, S>::const_view(m);
}
// a view of a view is the same view (ell_matrix_view.cu: "construct view from view")
template <typename A1, typename A2, typename I, typename V, typename S>
ell_matrix_view<A1, A2, I, V, S> make_ell_matrix_view(const ell_matrix_view<A1, A2, I, V, S> &v) {
  return v;
}

}  // namespace cusp
#include "convert.h"
