// cusp/dia_matrix.h — cusp::dia_matrix / dia_matrix_view
// (reference: cusp/dia_matrix.h:120-227, cusp/detail/dia_matrix.inl:26-90).
// values is a column-major array2d: entry (row, d) at values.values[d*pitch + row].
#pragma once
#include "detail/matrix_base.h"

namespace cusp {

template <typename Array1, typename Array2, typename IndexType = typename Array1::value_type,
          typename ValueType = typename Array2::value_type, typename MemorySpace = typename Array1::memory_space>
class dia_matrix_view;

template <typename IndexType, typename ValueType, typename MemorySpace>
class dia_matrix : public detail::matrix_base<IndexType, ValueType, MemorySpace, dia_format> {
  typedef detail::matrix_base<IndexType, ValueType, MemorySpace, dia_format> Parent;

 public:
  typedef array1d<IndexType, MemorySpace> diagonal_offsets_array_type;
  typedef array2d<ValueType, MemorySpace, column_major> values_array_type;
  typedef dia_matrix container;
  typedef dia_matrix_view<typename diagonal_offsets_array_type::view, typename values_array_type::view, IndexType,
                          ValueType, MemorySpace>
      view;
  typedef dia_matrix_view<typename diagonal_offsets_array_type::const_view,
                          typename values_array_type::const_view, IndexType, ValueType, MemorySpace>
      const_view;
  template <typename Space>
  struct rebind {
    typedef dia_matrix<IndexType, ValueType, Space> type;
  };

  diagonal_offsets_array_type diagonal_offsets;
  values_array_type values;

  dia_matrix() {}
  dia_matrix(size_t r, size_t c, size_t n, size_t num_diagonals, size_t alignment = 32)
      : Parent(r, c, n), diagonal_offsets(num_diagonals) {
    values.resize(r, num_diagonals, detail::round_up(r, alignment));
  }
  template <typename MatrixType, typename = typename std::enable_if<detail::has_format<MatrixType>::value>::type>
  dia_matrix(const MatrixType &m) {
    cusp::convert(m, *this);
  }
  template <typename MatrixType, typename = typename std::enable_if<detail::has_format<MatrixType>::value>::type>
  dia_matrix &operator=(const MatrixType &m) {
    cusp::convert(m, *this);
    return *this;
  }
  void resize(size_t r, size_t c, size_t n, size_t num_diagonals) {
    Parent::resize(r, c, n);
    diagonal_offsets.resize(num_diagonals);
    values.resize(r, num_diagonals);
  }
  void resize(size_t r, size_t c, size_t n, size_t num_diagonals, size_t alignment) {
    Parent::resize(r, c, n);
    diagonal_offsets.resize(num_diagonals);
    values.resize(r, num_diagonals, detail::round_up(r, alignment));
  }
  void swap(dia_matrix &o) {
    Parent::swap(o);
    diagonal_offsets.swap(o.diagonal_offsets);
    values.swap(o.values);
  }
};

template <typename Array1, typename Array2, typename IndexType, typename ValueType, typename MemorySpace>
class dia_matrix_view : public detail::matrix_base<IndexType, ValueType, MemorySpace, dia_format> {
  typedef detail::matrix_base<IndexType, ValueType, MemorySpace, dia_format> Parent;

 public:
  typedef Array1 diagonal_offsets_array_type;
  typedef Array2 values_array_type;
  typedef dia_matrix<IndexType, ValueType, MemorySpace> container;
  typedef dia_matrix_view view;

  Array1 diagonal_offsets;
  Array2 values;

  dia_matrix_view() {}
  dia_matrix_view(size_t r, size_t c, size_t n, const Array1 &offs, const Array2 &v)
      : Parent(r, c, n), diagonal_offsets(offs), values(v) {}
  template <typename Matrix, typename = typename std::enable_if<detail::has_format<Matrix>::value>::type>
  dia_matrix_view(Matrix &m) : Parent(m), diagonal_offsets(m.diagonal_offsets), values(m.values) {}

  void resize(size_t r, size_t c, size_t n, size_t num_diagonals) {
    Parent::resize(r, c, n);
    diagonal_offsets.resize(num_diagonals);
    values.resize(r, num_diagonals);
  }
};

template <typename Array1, typename Array2>
dia_matrix_view<Array1, Array2> make_dia_matrix_view(size_t r, size_t c, size_t n, const Array1 &offs,
                                                     const Array2 &v) {
  return dia_matrix_view<Array1, Array2>(r, c, n, offs, v);
}
template <typename I, typename V, typename S>
typename dia_matrix<I, V, S>::view make_dia_matrix_view(dia_matrix<I, V, S> &m) {
  return typename dia_matrix<I, V, S>::view(m);
}
template <typename I, typename V, typename S>
typename dia_matrix<I, V, S>::const_view make_dia_matrix_view(const dia_matrix<I, V, S> &m) {
  return typename dia_matrix<I, V, S>::const_view(m);
}
// a view of a view is the same view (dia_matrix_view.cu: "construct view from view")
template <typename A1, typename A2, typename I, typename V, typename S>
dia_matrix_view<A1, A2, I, V, S> make_dia_matrix_view(const dia_matrix_view<A1, A2, I, V, S> &v) {
  return v;
}

}  // namespace cusp
#include "convert.h"
