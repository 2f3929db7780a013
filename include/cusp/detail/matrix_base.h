// cusp/detail/matrix_base.h — shape + typedefs shared by every matrix container
// and view (reference: cusp/detail/matrix_base.h:30-75).
#pragma once
#include <cstddef>
#include <type_traits>
#include <utility>

#include "../array1d.h"
#include "../array2d.h"

namespace cusp {

// defined in cusp/convert.h (included at the end of every matrix header)
template <typename SourceType, typename DestinationType>
void convert(const SourceType &src, DestinationType &dst);

namespace detail {

template <typename IndexType, typename ValueType, typename MemorySpace, typename Format>
class matrix_base {
 public:
  typedef IndexType index_type;
  typedef ValueType value_type;
  typedef MemorySpace memory_space;
  typedef Format format;

  size_t num_rows = 0, num_cols = 0, num_entries = 0;

  matrix_base() {}
  matrix_base(size_t r, size_t c) : num_rows(r), num_cols(c) {}
  matrix_base(size_t r, size_t c, size_t n) : num_rows(r), num_cols(c), num_entries(n) {}
  template <typename Matrix>
  explicit matrix_base(const Matrix &m) : num_rows(m.num_rows), num_cols(m.num_cols), num_entries(m.num_entries) {}

  void resize(size_t r, size_t c, size_t n) {
    num_rows = r;
    num_cols = c;
    num_entries = n;
  }
  void swap(matrix_base &o) {
    std::swap(num_rows, o.num_rows);
    std::swap(num_cols, o.num_cols);
    std::swap(num_entries, o.num_entries);
  }
};

// SFINAE: T is a cusp matrix/array (has a nested `format`)
template <typename T, typename = void>
struct has_format : std::false_type {};
template <typename T>
struct has_format<T, typename std::conditional<false, typename T::format, void>::type> : std::true_type {};

}  // namespace detail
}  // namespace cusp
