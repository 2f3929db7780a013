// cusp/array2d.h — dense 2-D array with pitch (reference: cusp/array2d.h:152-231,
// cusp/detail/array2d.inl:100-126).  Only what the ELL / DIA containers and the
// hot-path tests use: storage, pitch, (i,j) access, resize, cross-space copies,
// equality.  Entry (i,j) lives at  row_major: i*pitch + j,  column_major: j*pitch + i.
#pragma once
#include "array1d.h"

namespace cusp {

namespace detail {
template <typename Orientation>
struct orient;
template <>
struct orient<row_major> {
  static size_t minor(size_t, size_t c) { return c; }
  static size_t major(size_t r, size_t) { return r; }
  static size_t index(size_t i, size_t j, size_t pitch) { return i * pitch + j; }
};
template <>
struct orient<column_major> {
  static size_t minor(size_t r, size_t) { return r; }
  static size_t major(size_t, size_t c) { return c; }
  static size_t index(size_t i, size_t j, size_t pitch) { return j * pitch + i; }
};
inline size_t round_up(size_t n, size_t k) { return k ? k * ((n + k - 1) / k) : n; }
}  // namespace detail

template <typename ArrayView, typename Orientation>
class array2d_view;

// defined in cusp/convert.h
template <typename SourceType, typename DestinationType>
void convert(const SourceType &src, DestinationType &dst);

template <typename T, typename MemorySpace, typename Orientation = row_major>
class array2d {
 public:
  typedef T value_type;
  typedef MemorySpace memory_space;
  typedef array2d_format format;
  typedef Orientation orientation;
  typedef array1d<T, MemorySpace> values_array_type;
  typedef array2d container;
  typedef array2d_view<typename values_array_type::view, Orientation> view;
  typedef array2d_view<typename values_array_type::const_view, Orientation> const_view;
  template <typename Space>
  struct rebind {
    typedef array2d<T, Space, Orientation> type;
  };

  size_t num_rows = 0, num_cols = 0, num_entries = 0, pitch = 0;
  values_array_type values;

  array2d() {}
  array2d(size_t r, size_t c) { resize(r, c); }
  array2d(size_t r, size_t c, const T &value) {
    resize(r, c);
    values.assign(values.size(), value);
  }
  array2d(size_t r, size_t c, const T &value, size_t p) {
    resize(r, c, p);
    values.assign(values.size(), value);
  }
  template <typename U, typename Space>
  array2d(const array2d<U, Space, Orientation> &o)
      : num_rows(o.num_rows), num_cols(o.num_cols), num_entries(o.num_entries), pitch(o.pitch), values(o.values) {}
  // dense image of any sparse matrix (cusp/detail/array2d.inl:100-126 -> cusp::convert)
  template <typename MatrixType,
            typename = typename std::enable_if<std::is_base_of<sparse_format, typename MatrixType::format>::value>::type>
  array2d(const MatrixType &m) {
    cusp::convert(m, *this);
  }
  template <typename MatrixType,
            typename = typename std::enable_if<std::is_base_of<sparse_format, typename MatrixType::format>::value>::type>
  array2d &operator=(const MatrixType &m) {
    cusp::convert(m, *this);
    return *this;
  }
  template <typename U, typename Space>
  array2d &operator=(const array2d<U, Space, Orientation> &o) {
    num_rows = o.num_rows;
    num_cols = o.num_cols;
    num_entries = o.num_entries;
    pitch = o.pitch;
    values = o.values;
    return *this;
  }

  // the other storage order: entries are re-laid out (cusp/detail/array2d.inl, testing/array2d.cu:200-229)
  template <typename U, typename Space, typename O2,
            typename = typename std::enable_if<!std::is_same<O2, Orientation>::value>::type>
  array2d(const array2d<U, Space, O2> &o) {
    assign_reoriented(o);
  }
  template <typename U, typename Space, typename O2,
            typename = typename std::enable_if<!std::is_same<O2, Orientation>::value>::type>
  array2d &operator=(const array2d<U, Space, O2> &o) {
    assign_reoriented(o);
    return *this;
  }

  void resize(size_t r, size_t c) { resize(r, c, detail::orient<Orientation>::minor(r, c)); }
  void resize(size_t r, size_t c, size_t p) {
    if (p < detail::orient<Orientation>::minor(r, c))
      throw cusp::invalid_input_exception("array2d: pitch smaller than the minor dimension");
    num_rows = r;
    num_cols = c;
    num_entries = r * c;
    pitch = p;
    values.resize(p * detail::orient<Orientation>::major(r, c));
  }
  void swap(array2d &o) {
    std::swap(num_rows, o.num_rows);
    std::swap(num_cols, o.num_cols);
    std::swap(num_entries, o.num_entries);
    std::swap(pitch, o.pitch);
    values.swap(o.values);
  }
  typename values_array_type::reference operator()(size_t i, size_t j) {
    return values[detail::orient<Orientation>::index(i, j, pitch)];
  }
  T operator()(size_t i, size_t j) const { return values[detail::orient<Orientation>::index(i, j, pitch)]; }

  // contiguous lines of the storage order as array1d views (cusp/array2d.h: row(i) of a row_major,
  // column(j) of a column_major array; the strided direction is not provided here)
  typedef typename values_array_type::view row_view;

  typedef typename values_array_type::view column_view;
  row_view row(size_t i) {
    static_assert(std::is_same<Orientation, row_major>::value, "array2d::row(): contiguous only for row_major");
    return values.subarray(i * pitch, num_cols);
  }
  column_view column(size_t j) {
    static_assert(std::is_same<Orientation, column_major>::value,
                  "array2d::column(): contiguous only for column_major");
    return values.subarray(j * pitch, num_rows);
  }

 private:
  template <typename U, typename Space, typename O2>
  void assign_reoriented(const array2d<U, Space, O2> &o) {
    auto h = detail::to_host_vector(o.values);
    resize(o.num_rows, o.num_cols);
    std::vector<T> mine(values.size(), T(0));
    for (size_t i = 0; i < o.num_rows; ++i)
      for (size_t j = 0; j < o.num_cols; ++j)
        mine[detail::orient<Orientation>::index(i, j, pitch)] = (T)h[detail::orient<O2>::index(i, j, o.pitch)];
    detail::raw_copy<T, host_memory, MemorySpace>(mine.data(), detail::raw_ptr(values), mine.size());
  }
};

template <typename ArrayView, typename Orientation = row_major>
class array2d_view {
 public:
  typedef typename ArrayView::value_type value_type;
  typedef typename ArrayView::memory_space memory_space;
  typedef array2d_format format;
  typedef Orientation orientation;
  typedef ArrayView values_array_type;
  typedef array2d<value_type, memory_space, Orientation> container;
  typedef array2d_view view;

  size_t num_rows = 0, num_cols = 0, num_entries = 0, pitch = 0;
  ArrayView values;

  array2d_view() {}
  array2d_view(size_t r, size_t c, size_t p, const ArrayView &v)
      : num_rows(r), num_cols(c), num_entries(r * c), pitch(p), values(v) {}
  template <typename T, typename Space>
  array2d_view(array2d<T, Space, Orientation> &a)
      : num_rows(a.num_rows), num_cols(a.num_cols), num_entries(a.num_entries), pitch(a.pitch), values(a.values) {}
  template <typename T, typename Space>
  array2d_view(const array2d<T, Space, Orientation> &a)
      : num_rows(a.num_rows), num_cols(a.num_cols), num_entries(a.num_entries), pitch(a.pitch), values(a.values) {}
  template <typename OtherView>
  array2d_view(const array2d_view<OtherView, Orientation> &a)
      : num_rows(a.num_rows), num_cols(a.num_cols), num_entries(a.num_entries), pitch(a.pitch), values(a.values) {}

  void resize(size_t r, size_t c) { resize(r, c, detail::orient<Orientation>::minor(r, c)); }
  void resize(size_t r, size_t c, size_t p) {
    values.resize(p * detail::orient<Orientation>::major(r, c));
    num_rows = r;
    num_cols = c;
    num_entries = r * c;
    pitch = p;
  }
  typename ArrayView::reference operator()(size_t i, size_t j) const {
    return values[detail::orient<Orientation>::index(i, j, pitch)];
  }
};

template <typename ArrayView, typename Orientation>
array2d_view<ArrayView, Orientation> make_array2d_view(size_t r, size_t c, size_t p, const ArrayView &v,
                                                       Orientation) {
  return array2d_view<ArrayView, Orientation>(r, c, p, v);
}
template <typename T, typename Space, typename O>
typename array2d<T, Space, O>::view make_array2d_view(array2d<T, Space, O> &a) {
  return typename array2d<T, Space, O>::view(a);
}
template <typename T, typename Space, typename O>
typename array2d<T, Space, O>::const_view make_array2d_view(const array2d<T, Space, O> &a) {
  return typename array2d<T, Space, O>::const_view(a);
}

// logical equality (pitch padding ignored), across memory spaces
template <typename T1, typename S1, typename O1, typename T2, typename S2, typename O2>
bool operator==(const array2d<T1, S1, O1> &a, const array2d<T2, S2, O2> &b) {
  if (a.num_rows != b.num_rows || a.num_cols != b.num_cols) return false;
  auto ha = detail::to_host_vector(a.values);
  auto hb = detail::to_host_vector(b.values);
  for (size_t i = 0; i < a.num_rows; ++i)
    for (size_t j = 0; j < a.num_cols; ++j)
      if (!(ha[detail::orient<O1>::index(i, j, a.pitch)] == hb[detail::orient<O2>::index(i, j, b.pitch)]))
        return false;
  return true;
}
template <typename T1, typename S1, typename O1, typename T2, typename S2, typename O2>
bool operator!=(const array2d<T1, S1, O1> &a, const array2d<T2, S2, O2> &b) {
  return !(a == b);
}

}  // namespace cusp
