// cusp/hyb_matrix.h — cusp::hyb_matrix / hyb_matrix_view
// (reference: cusp/hyb_matrix.h:142-248, cusp/detail/hyb_matrix.inl): an ELL part
// for the first K entries of every row plus a COO part for the overflow.
#pragma once
#include "coo_matrix.h"
#include "ell_matrix.h"

namespace cusp {

template <typename EllView, typename CooView, typename IndexType = typename EllView::index_type,
          typename ValueType = typename EllView::value_type, typename MemorySpace = typename EllView::memory_space>
class hyb_matrix_view;

template <typename IndexType, typename ValueType, typename MemorySpace>
class hyb_matrix : public detail::matrix_base<IndexType, ValueType, MemorySpace, hyb_format> {
  typedef detail::matrix_base<IndexType, ValueType, MemorySpace, hyb_format> Parent;

 public:
  typedef cusp::ell_matrix<IndexType, ValueType, MemorySpace> ell_matrix_type;
  typedef cusp::coo_matrix<IndexType, ValueType, MemorySpace> coo_matrix_type;
  typedef hyb_matrix container;
  typedef hyb_matrix_view<typename ell_matrix_type::view, typename coo_matrix_type::view, IndexType, ValueType,
                          MemorySpace>
      view;
  typedef hyb_matrix_view<typename ell_matrix_type::const_view, typename coo_matrix_type::const_view, IndexType,
                          ValueType, MemorySpace>
      const_view;
  template <typename Space>
  struct rebind {
    typedef hyb_matrix<IndexType, ValueType, Space> type;
  };

  ell_matrix_type ell;
  coo_matrix_type coo;

  hyb_matrix() {}
  hyb_matrix(size_t r, size_t c, size_t num_ell_entries, size_t num_coo_entries, size_t num_entries_per_row,
             size_t alignment = 32)
      : Parent(r, c, num_ell_entries + num_coo_entries),
        ell(r, c, num_ell_entries, num_entries_per_row, alignment),
        coo(r, c, num_coo_entries) {}
  template <typename MatrixType, typename = typename std::enable_if<detail::has_format<MatrixType>::value>::type>
  hyb_matrix(const MatrixType &m) {
    cusp::convert(m, *this);
  }
  template <typename MatrixType, typename = typename std::enable_if<detail::has_format<MatrixType>::value>::type>
  hyb_matrix &operator=(const MatrixType &m) {
    cusp::convert(m, *this);
    return *this;
  }
  void resize(size_t r, size_t c, size_t num_ell_entries, size_t num_coo_entries, size_t num_entries_per_row,
              size_t alignment = 32) {
    Parent::resize(r, c, num_ell_entries + num_coo_entries);
    ell.resize(r, c, num_ell_entries, num_entries_per_row, alignment);
    coo.resize(r, c, num_coo_entries);
  }
  void swap(hyb_matrix &o) {
    Parent::swap(o);
    ell.swap(o.ell);
    coo.swap(o.coo);
  }
};

template <typename EllView, typename CooView, typename IndexType, typename ValueType, typename MemorySpace>
class hyb_matrix_view : public detail::matrix_base<IndexType, ValueType, MemorySpace, hyb_format> {
  typedef detail::matrix_base<IndexType, ValueType, MemorySpace, hyb_format> Parent;

 public:
  typedef EllView ell_matrix_type;
  typedef CooView coo_matrix_type;
  typedef hyb_matrix<IndexType, ValueType, MemorySpace> container;
  typedef hyb_matrix_view view;

  EllView ell;
  CooView coo;

  hyb_matrix_view() {}
  hyb_matrix_view(const EllView &e, const CooView &c)
      : Parent(e.num_rows, e.num_cols, e.num_entries + c.num_entries), ell(e), coo(c) {}
  template <typename Matrix, typename = typename std::enable_if<detail::has_format<Matrix>::value>::type>
  hyb_matrix_view(Matrix &m) : Parent(m), ell(m.ell), coo(m.coo) {}
};

template <typename EllView, typename CooView>
hyb_matrix_view<EllView, CooView> make_hyb_matrix_view(const EllView &e, const CooView &c) {
  return hyb_matrix_view<EllView, CooView>(e, c);
}
template <typename I, typename V, typename S>
typename hyb_matrix<I, V, S>::view make_hyb_matrix_view(hyb_matrix<I, V, S> &m) {
  return typename hyb_matrix<I, V, S>::view(m);
}
template <typename I, typename V, typename S>
typename hyb_matrix<I, V, S>::const_view make_hyb_matrix_view(const hyb_matrix<I, V, S> &m) {
  return typename hyb_matrix<I, V, S>::const_view(m);
}
// a view of a view is the same view (hyb_matrix_view.cu: "construct view from view")
template <typename E, typename C, typename I, typename V, typename S>
hyb_matrix_view<E, C, I, V, S> make_hyb_matrix_view(const hyb_matrix_view<E, C, I, V, S> &v) {
  return v;
}

}  // namespace cusp
#include "convert.h"
