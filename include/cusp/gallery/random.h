// cusp/gallery/random.h — cusp::gallery::random(matrix, m, n, num_samples)
// (reference: cusp/gallery/random.h, gallery/detail/random.inl:33-63):
// srand(m ^ n ^ num_samples); num_samples draws (rand() % m, rand() % n) with
// value 1; sorted by (row, column); duplicates removed; converted to `matrix`.
// The C library generator is part of the definition: the same libc gives the
// same matrix as the reference (BASELINE configs[3] inputs).
#pragma once
#include <algorithm>
#include <cstdlib>
#include <utility>
#include <vector>

#include "../convert.h"

namespace cusp {
namespace gallery {

template <typename MatrixType>
void random(MatrixType &matrix, const size_t m, const size_t n, const size_t num_samples) {
  typedef typename MatrixType::index_type IndexType;
  typedef typename MatrixType::value_type ValueType;

  std::vector<std::pair<IndexType, IndexType>> entries(num_samples);
  srand((unsigned)(m ^ n ^ num_samples));
  for (size_t k = 0; k < num_samples; ++k) {
    const IndexType r = (IndexType)(rand() % m);  // row first, then column: the draw order matters
    const IndexType c = (IndexType)(rand() % n);
    entries[k] = std::make_pair(r, c);
  }
  std::sort(entries.begin(), entries.end());
  entries.erase(std::unique(entries.begin(), entries.end()), entries.end());

  cusp::coo_matrix<IndexType, ValueType, cusp::host_memory> coo(m, n, entries.size());
  for (size_t k = 0; k < entries.size(); ++k) {
    coo.row_indices[k] = entries[k].first;
    coo.column_indices[k] = entries[k].second;
    coo.values[k] = ValueType(1);
  }
  matrix = coo;
}

}  // namespace gallery
}  // namespace cusp
