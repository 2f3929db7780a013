// cusp/ktt/matrix_generation.h — banded all-ones DIA test matrices
// (reference: cusp/ktt/matrix_generation.h:14-102; used by testing/ktt.cu:274-281
// and main.cu's DRAM-traffic sweeps).  pitch == rows, one slab column per offset,
// ones exactly on the in-range part of every diagonal.
#pragma once
#include <algorithm>
#include <stdexcept>
#include <vector>

#include "../dia_matrix.h"

namespace cusp {
namespace ktt {

inline cusp::dia_matrix<int, float, cusp::host_memory> make_diagonal_matrix(int rows, int cols,
                                                                             const std::vector<int> &diag_offsets) {
  typedef cusp::dia_matrix<int, float, cusp::host_memory> Dia;
  const size_t nd = diag_offsets.size();
  Dia A;
  A.resize((size_t)rows, (size_t)cols, 0, nd);  // pitch == rows, slots zero-initialised below
  std::fill(A.values.values.begin(), A.values.values.end(), 0.0f);
  size_t filled = 0;
  for (size_t d = 0; d < nd; ++d) {
    const int off = diag_offsets[d];
    const int first_row = off < 0 ? -off : 0;
    const int first_col = off < 0 ? 0 : off;
    if (first_row >= rows || first_col >= cols) throw std::runtime_error("make_diagonal_matrix: Diagonal out of bounds.");
    A.diagonal_offsets[d] = off;
    const int len = std::min(rows - first_row, cols - first_col);
    for (int r = first_row; r < first_row + len; ++r) A.values((size_t)r, d) = 1.0f;
    filled += (size_t)len;
  }
  A.num_entries = filled;
  return A;
}

// main diagonal plus the same number of diagonals on either side, `offset_step` apart
inline cusp::dia_matrix<int, float, cusp::host_memory> make_diagonal_symmetric_matrix(int rows, int cols,
                                                                                       int offset_step,
                                                                                       int diagonal_count) {
  std::vector<int> offsets;
  offsets.reserve((size_t)diagonal_count);
  const int first = -offset_step * diagonal_count / 2;
  for (int i = 0; i < diagonal_count; ++i) {
    const int off = first + offset_step * i;
    const int first_row = off < 0 ? -off : 0;
    const int first_col = off < 0 ? 0 : off;
    if (first_row >= rows || first_col >= cols)
      throw std::runtime_error("make_diagonal_symmetric_matrix: Too many diagonals.");
    offsets.push_back(off);
  }
  return make_diagonal_matrix(rows, cols, offsets);
}

}  // namespace ktt
}  // namespace cusp
