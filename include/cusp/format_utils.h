// cusp/format_utils.h — offsets_to_indices / indices_to_offsets / extract_diagonal
// (reference: cusp/format_utils.h, generic/format_utils.inl:36-279;
// pinned by testing/format_utils.cu:13-75).  Setup-time helpers: they run on the
// host and upload (compute_{max,optimal}_entries_per_row live in convert.h).
#pragma once
#include <vector>

#include "convert.h"

namespace cusp {

// offsets[i] .. offsets[i+1] -> indices[k] = i
template <typename OffsetArray, typename IndexArray>
void offsets_to_indices(const OffsetArray &offsets, IndexArray &indices) {
  typedef typename IndexArray::value_type I;
  auto off = detail::to_host_vector(offsets);
  std::vector<I> idx(indices.size());
  for (size_t i = 0; i + 1 < off.size(); ++i)
    for (size_t k = (size_t)off[i]; k < (size_t)off[i + 1] && k < idx.size(); ++k) idx[k] = (I)i;
  detail::raw_copy<I, host_memory, typename IndexArray::memory_space>(idx.data(), detail::raw_ptr(indices), idx.size());
}

// sorted indices -> offsets[i] = number of indices < i
template <typename IndexArray, typename OffsetArray>
void indices_to_offsets(const IndexArray &indices, OffsetArray &offsets) {
  typedef typename OffsetArray::value_type I;
  auto idx = detail::to_host_vector(indices);
  std::vector<I> off(offsets.size());
  size_t k = 0;
  for (size_t i = 0; i < off.size(); ++i) {
    while (k < idx.size() && (size_t)idx[k] < i) ++k;
    off[i] = (I)k;
  }
  detail::raw_copy<I, host_memory, typename OffsetArray::memory_space>(off.data(), detail::raw_ptr(offsets), off.size());
}

// main diagonal of any sparse matrix (missing entries are 0; duplicates are summed like the COO path)
template <typename Matrix, typename Array>
void extract_diagonal(const Matrix &A, Array &output) {
  typedef typename Array::value_type V;
  detail::host_csr<typename Matrix::index_type, V> H;
  detail::gather(A, H, typename Matrix::format());
  const size_t n = std::min(A.num_rows, A.num_cols);
  std::vector<V> d(n, V(0));
  for (size_t i = 0; i < n; ++i)
    for (auto k = H.offsets[i]; k < H.offsets[i + 1]; ++k)
      if ((size_t)H.columns[k] == i) d[i] += H.values[k];
  output.resize(n);
  detail::raw_copy<V, host_memory, typename Array::memory_space>(d.data(), detail::raw_ptr(output), n);
}

}  // namespace cusp
