// cusp/ktt/ellr_matrix.h — cusp::ktt::ellr_matrix: ELL plus per-row lengths
// (reference: cusp/ktt/ellr_matrix.h:18-93, cusp/ktt/detail/ellr_matrix.inl:16-137).
// row_lengths[i] = number of leading slots of row i whose column index is not
// the padding marker.  On the device they are computed by
// b200sp_ell_row_lengths (the reference uses a thrust::transform,
// ellr_matrix.inl:16-52); cusp::multiply then takes the ELL-R kernel
// (b200sp_spmv_ellr_*), which stops each row at its length.
#pragma once
#include <type_traits>

#include "../detail/descriptor.h"
#include "../ell_matrix.h"

namespace cusp {
namespace ktt {

template <typename IndexType, typename ValueType, typename MemorySpace>
class ellr_matrix : public cusp::ell_matrix<IndexType, ValueType, MemorySpace> {
  typedef cusp::ell_matrix<IndexType, ValueType, MemorySpace> Parent;

 public:
  typedef cusp::array1d<IndexType, MemorySpace> row_lengths_array_type;
  typedef ellr_matrix container;
  template <typename Space>
  struct rebind {
    typedef ellr_matrix<IndexType, ValueType, Space> type;
  };

  row_lengths_array_type row_lengths;

  ellr_matrix() {}
  ellr_matrix(const size_t num_rows, const size_t num_cols, const size_t num_entries, const size_t num_entries_per_row,
              const size_t alignment = 32)
      : Parent(num_rows, num_cols, num_entries, num_entries_per_row, alignment), row_lengths(num_rows) {}
  ellr_matrix(const ellr_matrix &o) : Parent(static_cast<const Parent &>(o)), row_lengths(o.row_lengths) {}
  template <typename MatrixType, typename = typename std::enable_if<cusp::detail::has_format<MatrixType>::value>::type>
  ellr_matrix(const MatrixType &matrix) : Parent(matrix) {
    update_row_lengths();
  }

  void resize(const size_t num_rows, const size_t num_cols, const size_t num_entries,
              const size_t num_entries_per_row) {
    Parent::resize(num_rows, num_cols, num_entries, num_entries_per_row);
    row_lengths.resize(num_rows);
  }
  void resize(const size_t num_rows, const size_t num_cols, const size_t num_entries, const size_t num_entries_per_row,
              const size_t alignment) {
    Parent::resize(num_rows, num_cols, num_entries, num_entries_per_row, alignment);
    row_lengths.resize(num_rows);
  }
  void swap(ellr_matrix &matrix) {
    Parent::swap(matrix);
    row_lengths.swap(matrix.row_lengths);
  }
  ellr_matrix &operator=(const ellr_matrix &o) {
    Parent::operator=(static_cast<const Parent &>(o));
    row_lengths = o.row_lengths;
    return *this;
  }
  template <typename MatrixType, typename = typename std::enable_if<cusp::detail::has_format<MatrixType>::value>::type>
  ellr_matrix &operator=(const MatrixType &matrix) {
    Parent::operator=(matrix);
    update_row_lengths();
    return *this;
  }

  void update_row_lengths() {
    row_lengths.resize(this->num_rows);
    compute(MemorySpace());
  }

 private:
  void compute(cusp::host_memory) {
    const size_t K = this->column_indices.num_cols;
    for (size_t i = 0; i < this->num_rows; ++i) {
      size_t len = 0;
      while (len < K && this->column_indices(i, len) != Parent::invalid_index) ++len;
      row_lengths[i] = (IndexType)len;
    }
  }
  void compute(cusp::device_memory) {
    static_assert(sizeof(IndexType) == 4, "device ellr_matrix takes 32-bit indices");
    if (this->num_rows == 0) return;
    cusp::detail::check(b200sp_ell_row_lengths(
        cusp::detail::engine(), cusp::detail::current_stream(), (int64_t)this->num_rows,
        (int64_t)this->column_indices.num_cols, (int64_t)this->column_indices.pitch,
        reinterpret_cast<const int32_t *>(cusp::detail::raw_ptr(this->column_indices.values)),
        reinterpret_cast<int32_t *>(cusp::detail::raw_ptr(row_lengths))));
  }
};

}  // namespace ktt
}  // namespace cusp
