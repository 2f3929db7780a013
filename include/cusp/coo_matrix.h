// cusp/coo_matrix.h — cusp::coo_matrix / coo_matrix_view
// (reference: cusp/coo_matrix.h:116-225 container, :258-420 view,
// cusp/detail/coo_matrix.inl:95-127 sorting helpers).
#pragma once
#include <numeric>

#include "detail/matrix_base.h"

namespace cusp {

template <typename Array1, typename Array2, typename Array3, typename IndexType = typename Array1::value_type,
          typename ValueType = typename Array3::value_type, typename MemorySpace = typename Array1::memory_space>
class coo_matrix_view;

template <typename IndexType, typename ValueType, typename MemorySpace>
class coo_matrix : public detail::matrix_base<IndexType, ValueType, MemorySpace, coo_format> {
  typedef detail::matrix_base<IndexType, ValueType, MemorySpace, coo_format> Parent;

 public:
  typedef array1d<IndexType, MemorySpace> row_indices_array_type;
  typedef array1d<IndexType, MemorySpace> column_indices_array_type;
  typedef array1d<ValueType, MemorySpace> values_array_type;
  typedef coo_matrix container;
  typedef coo_matrix_view<typename row_indices_array_type::view, typename column_indices_array_type::view,
                          typename values_array_type::view, IndexType, ValueType, MemorySpace>
      view;
  typedef coo_matrix_view<typename row_indices_array_type::const_view,
                          typename column_indices_array_type::const_view, typename values_array_type::const_view,
                          IndexType, ValueType, MemorySpace>
      const_view;
  template <typename Space>
  struct rebind {
    typedef coo_matrix<IndexType, ValueType, Space> type;
  };

  row_indices_array_type row_indices;
  column_indices_array_type column_indices;
  values_array_type values;

  coo_matrix() {}
  coo_matrix(size_t r, size_t c, size_t n) : Parent(r, c, n), row_indices(n), column_indices(n), values(n) {}
  template <typename MatrixType, typename = typename std::enable_if<detail::has_format<MatrixType>::value>::type>
  coo_matrix(const MatrixType &m) {
    cusp::convert(m, *this);
  }
  template <typename MatrixType, typename = typename std::enable_if<detail::has_format<MatrixType>::value>::type>
  coo_matrix &operator=(const MatrixType &m) {
    cusp::convert(m, *this);
    return *this;
  }

  void resize(size_t r, size_t c, size_t n) {
    Parent::resize(r, c, n);
    row_indices.resize(n);
    column_indices.resize(n);
    values.resize(n);
  }
  void swap(coo_matrix &o) {
    Parent::swap(o);
    row_indices.swap(o.row_indices);
    column_indices.swap(o.column_indices);
    values.swap(o.values);
  }

  // cusp/detail/coo_matrix.inl:95-127 (stable, like cusp::sort_by_row)
  void sort_by_row() { sort_impl(false); }
  void sort_by_row_and_column() { sort_impl(true); }
  bool is_sorted_by_row() const {
    auto r = detail::to_host_vector(row_indices);
    return std::is_sorted(r.begin(), r.end());
  }
  bool is_sorted_by_row_and_column() const {
    auto r = detail::to_host_vector(row_indices);
    auto c = detail::to_host_vector(column_indices);
    for (size_t k = 1; k < r.size(); ++k)
      if (r[k - 1] > r[k] || (r[k - 1] == r[k] && c[k - 1] > c[k])) return false;
    return true;
  }

 private:
  void sort_impl(bool by_col) {
    auto r = detail::to_host_vector(row_indices);
    auto c = detail::to_host_vector(column_indices);
    auto v = detail::to_host_vector(values);
    std::vector<size_t> perm(r.size());
    std::iota(perm.begin(), perm.end(), (size_t)0);
    std::stable_sort(perm.begin(), perm.end(), [&](size_t a, size_t b) {
      if (r[a] != r[b]) return r[a] < r[b];
      return by_col && c[a] < c[b];
    });
    std::vector<IndexType> r2(r.size()), c2(r.size());
    std::vector<ValueType> v2(r.size());
    for (size_t k = 0; k < perm.size(); ++k) {
      r2[k] = r[perm[k]];
      c2[k] = c[perm[k]];
      v2[k] = v[perm[k]];
    }
    detail::raw_copy<IndexType, host_memory, MemorySpace>(r2.data(), detail::raw_ptr(row_indices), r2.size());
    detail::raw_copy<IndexType, host_memory, MemorySpace>(c2.data(), detail::raw_ptr(column_indices), c2.size());
    detail::raw_copy<ValueType, host_memory, MemorySpace>(v2.data(), detail::raw_ptr(values), v2.size());
  }
};

template <typename Array1, typename Array2, typename Array3, typename IndexType, typename ValueType,
          typename MemorySpace>
class coo_matrix_view : public detail::matrix_base<IndexType, ValueType, MemorySpace, coo_format> {
  typedef detail::matrix_base<IndexType, ValueType, MemorySpace, coo_format> Parent;

 public:
  typedef Array1 row_indices_array_type;
  typedef Array2 column_indices_array_type;
  typedef Array3 values_array_type;
  typedef coo_matrix<IndexType, ValueType, MemorySpace> container;
  typedef coo_matrix_view view;

  Array1 row_indices;
  Array2 column_indices;
  Array3 values;

  coo_matrix_view() {}
  coo_matrix_view(size_t r, size_t c, size_t n, const Array1 &ri, const Array2 &ci, const Array3 &v)
      : Parent(r, c, n), row_indices(ri), column_indices(ci), values(v) {}
  template <typename Matrix, typename = typename std::enable_if<detail::has_format<Matrix>::value>::type>
  coo_matrix_view(Matrix &m)
      : Parent(m), row_indices(m.row_indices), column_indices(m.column_indices), values(m.values) {}

  void resize(size_t r, size_t c, size_t n) {
    Parent::resize(r, c, n);
    row_indices.resize(n);
    column_indices.resize(n);
    values.resize(n);
  }
};

template <typename Array1, typename Array2, typename Array3>
coo_matrix_view<Array1, Array2, Array3> make_coo_matrix_view(size_t r, size_t c, size_t n, const Array1 &ri,
                                                             const Array2 &ci, const Array3 &v) {
  return coo_matrix_view<Array1, Array2, Array3>(r, c, n, ri, ci, v);
}
template <typename I, typename V, typename S>
typename coo_matrix<I, V, S>::view make_coo_matrix_view(coo_matrix<I, V, S> &m) {
  return typename coo_matrix<I, V, S>::view(m);
}
template <typename I, typename V, typename S>
typename coo_matrix<I, V, S>::const_view make_coo_matrix_view(const coo_matrix<I, V, S> &m) {
  return typename coo_matrix<I, V, S>::const_view(m);
}
// a view of a view is the same view (coo_matrix_view.cu: "construct view from view")
template <typename A1, typename A2, typename A3, typename I, typename V, typename S>
coo_matrix_view<A1, A2, A3, I, V, S> make_coo_matrix_view(const coo_matrix_view<A1, A2, A3, I, V, S> &v) {
  return v;
}

}  // namespace cusp
#include "convert.h"
