// cusp/detail/descriptor.h — unpack a device-resident cusp matrix (container or
// view) into the non-owning b200sp_matrix descriptor of the C ABI.  This is the
// job the reference does inline at every launch site with
// thrust::raw_pointer_cast(&A.values[0]), A.values.pitch,
// A.column_indices.num_cols ... (cuda/detail/multiply/ell_spmv.h:128-135,
// dia_spmv.h:162-180, csr_vector_spmv.h:190-214, cuda/ktt/utils.h:57-60).
#pragma once
#include <cstring>
#include <type_traits>

#include "../array1d.h"
#include "../memory.h"

namespace cusp {
namespace detail {

template <typename T>
struct is_abi_value : std::integral_constant<bool, std::is_same<T, float>::value || std::is_same<T, double>::value> {};

// ELL-R: an ELL matrix that also carries row_lengths (cusp/ktt/ellr_matrix.h:18-25)
template <typename M, typename = void>
struct has_row_lengths : std::false_type {};
template <typename M>
struct has_row_lengths<M, typename std::conditional<false, decltype(std::declval<const M &>().row_lengths), void>::type>
    : std::true_type {};

template <typename M>
inline void describe_common(const M &A, b200sp_matrix &d, b200sp_format f) {
  std::memset(&d, 0, sizeof(d));
  d.format = f;
  d.dtype = dtype_of<typename M::value_type>::value;
  d.num_rows = (int64_t)A.num_rows;
  d.num_cols = (int64_t)A.num_cols;
  d.num_entries = (int64_t)A.num_entries;
}

template <typename M>
inline void describe(const M &A, b200sp_matrix &d, csr_format) {
  describe_common(A, d, B200SP_FMT_CSR);
  d.row_offsets = raw_ptr(A.row_offsets);
  d.column_indices = raw_ptr(A.column_indices);
  d.values = raw_ptr(A.values);
}
template <typename M>
inline void describe(const M &A, b200sp_matrix &d, coo_format) {
  describe_common(A, d, B200SP_FMT_COO);
  d.row_indices = raw_ptr(A.row_indices);
  d.column_indices = raw_ptr(A.column_indices);
  d.values = raw_ptr(A.values);
}
template <typename M>
inline void describe(const M &A, b200sp_matrix &d, dia_format) {
  describe_common(A, d, B200SP_FMT_DIA);
  d.num_cols_per_row = (int64_t)A.diagonal_offsets.size();
  d.pitch = (int64_t)A.values.pitch;
  d.diagonal_offsets = raw_ptr(A.diagonal_offsets);
  d.values = raw_ptr(A.values.values);
}
template <typename M>
inline void describe_ellr(const M &, b200sp_matrix &, std::false_type) {}
template <typename M>
inline void describe_ellr(const M &A, b200sp_matrix &d, std::true_type) {
  d.format = B200SP_FMT_ELLR;
  d.row_offsets = raw_ptr(A.row_lengths);
}
template <typename M>
inline void describe(const M &A, b200sp_matrix &d, ell_format) {
  describe_common(A, d, B200SP_FMT_ELL);
  d.num_cols_per_row = (int64_t)A.column_indices.num_cols;
  d.pitch = (int64_t)A.column_indices.pitch;
  d.column_indices = raw_ptr(A.column_indices.values);
  d.values = raw_ptr(A.values.values);
  describe_ellr(A, d, has_row_lengths<M>());
}
template <typename M>
inline void describe(const M &A, b200sp_matrix &d, hyb_format) {
  describe_common(A, d, B200SP_FMT_HYB);
  d.num_cols_per_row = (int64_t)A.ell.column_indices.num_cols;
  d.pitch = (int64_t)A.ell.column_indices.pitch;
  d.column_indices = raw_ptr(A.ell.column_indices.values);
  d.values = raw_ptr(A.ell.values.values);
  d.coo_num_entries = (int64_t)A.coo.num_entries;
  d.coo_row_indices = raw_ptr(A.coo.row_indices);
  d.coo_column_indices = raw_ptr(A.coo.column_indices);
  d.coo_values = raw_ptr(A.coo.values);
}

// can this matrix type go through the C ABI?  (device memory, int32 indices, float/double values)
template <typename M, typename = void>
struct abi_matrix : std::false_type {};
template <typename M>
struct abi_matrix<M, typename std::conditional<false, typename M::index_type, void>::type>
    : std::integral_constant<bool, std::is_same<typename M::memory_space, device_memory>::value &&
                                       sizeof(typename M::index_type) == 4 &&
                                       std::is_integral<typename M::index_type>::value &&
                                       is_abi_value<typename M::value_type>::value &&
                                       std::is_base_of<sparse_format, typename M::format>::value> {};

template <typename M>
inline b200sp_matrix describe(const M &A) {
  b200sp_matrix d;
  describe(A, d, typename M::format());
  return d;
}

}  // namespace detail
}  // namespace cusp
