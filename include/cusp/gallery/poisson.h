// cusp/gallery/poisson.h — cusp::gallery::poisson5pt / 9pt / 7pt / 27pt
// (reference: cusp/gallery/poisson.h, gallery/detail/poisson.inl:29-116,
// gallery/detail/stencil.inl:33-206).
//
// The reference builds a DIA matrix from the stencil (pitch = num_rows, one
// diagonal per stencil point in stencil order, out-of-grid slots = 0,
// num_entries = number of non-zero slots) and cusp::convert()s it to the
// requested type.  Host containers follow exactly that route here.  For device
// DIA / ELL / CSR matrices of float/double and the 5- and 7-point stencils the
// arrays are produced directly on the GPU by the engine's builders
// (b200sp_poisson_{dia,ell,csr}_*), bit-identical to the reference pipeline and
// without the O(rows) launches of conversions/dia_to_other.h:227-251.
#pragma once
#include <array>
#include <vector>

#include "../convert.h"
#include "../detail/descriptor.h"

namespace cusp {
namespace gallery {
namespace detail {

struct stencil_point {
  std::array<long long, 3> d;  // offsets along x (fastest), y, z
  double value;
};

// host: stencil -> DIA, the reference's generate_matrix_from_stencil
template <typename I, typename V>
void stencil_to_dia(cusp::dia_matrix<I, V, cusp::host_memory> &A, const std::vector<stencil_point> &stencil,
                    const std::array<size_t, 3> &grid, int dims) {
  const size_t rows = grid[0] * grid[1] * grid[2];
  const size_t strides[3] = {1, grid[0], grid[0] * grid[1]};
  A.resize(rows, rows, 0, stencil.size());  // pitch == rows (stencil.inl:174)
  size_t nnz = 0;
  for (size_t s = 0; s < stencil.size(); ++s) {
    long long off = 0;
    for (int k = 0; k < dims; ++k) off += (long long)strides[k] * stencil[s].d[k];
    A.diagonal_offsets[s] = (I)off;
    for (size_t row = 0; row < rows; ++row) {
      size_t rem = row;
      bool inside = true;
      for (int k = 0; k < dims; ++k) {  // inside_grid: every shifted coordinate stays in [0, n_k)
        const long long c = (long long)(rem % grid[k]) + stencil[s].d[k];
        rem /= grid[k];
        if (c < 0 || c >= (long long)grid[k]) inside = false;
      }
      const V v = inside ? (V)stencil[s].value : V(0);
      A.values(row, s) = v;
      if (v != V(0)) ++nnz;
    }
  }
  A.num_entries = nnz;
}

template <typename M>
void from_stencil_host_route(M &matrix, const std::vector<stencil_point> &stencil, const std::array<size_t, 3> &grid,
                             int dims) {
  cusp::dia_matrix<typename M::index_type, typename M::value_type, cusp::host_memory> dia;
  stencil_to_dia(dia, stencil, grid, dims);
  cusp::convert(dia, matrix);
}
// dense destinations have no index_type
template <typename T, typename S, typename O>
void from_stencil_host_route(cusp::array2d<T, S, O> &matrix, const std::vector<stencil_point> &stencil,
                             const std::array<size_t, 3> &grid, int dims) {
  cusp::dia_matrix<int, T, cusp::host_memory> dia;
  stencil_to_dia(dia, stencil, grid, dims);
  cusp::convert(dia, matrix);
}

using cusp::detail::check;
using cusp::detail::current_stream;
using cusp::detail::engine;
using cusp::detail::raw_ptr;

inline b200sp_status pdia(int st, int64_t nx, int64_t ny, int64_t nz, int64_t rows, int64_t pitch, int *off, float *v) {
  return b200sp_poisson_dia_f32(engine(), current_stream(), st, nx, ny, nz, 0, rows, 0, pitch, off, v);
}
inline b200sp_status pdia(int st, int64_t nx, int64_t ny, int64_t nz, int64_t rows, int64_t pitch, int *off,
                          double *v) {
  return b200sp_poisson_dia_f64(engine(), current_stream(), st, nx, ny, nz, 0, rows, 0, pitch, off, v);
}
inline b200sp_status pell(int st, int64_t nx, int64_t ny, int64_t nz, int64_t rows, int64_t pitch, int *ci, float *v) {
  return b200sp_poisson_ell_f32(engine(), current_stream(), st, nx, ny, nz, 0, rows, 0, pitch, ci, v);
}
inline b200sp_status pell(int st, int64_t nx, int64_t ny, int64_t nz, int64_t rows, int64_t pitch, int *ci,
                          double *v) {
  return b200sp_poisson_ell_f64(engine(), current_stream(), st, nx, ny, nz, 0, rows, 0, pitch, ci, v);
}
inline b200sp_status pcsr(int st, int64_t nx, int64_t ny, int64_t nz, int64_t rows, const int *ro, int *ci, float *v) {
  return b200sp_poisson_csr_f32(engine(), current_stream(), st, nx, ny, nz, 0, rows, 0, ro, ci, v);
}
inline b200sp_status pcsr(int st, int64_t nx, int64_t ny, int64_t nz, int64_t rows, const int *ro, int *ci,
                          double *v) {
  return b200sp_poisson_csr_f64(engine(), current_stream(), st, nx, ny, nz, 0, rows, 0, ro, ci, v);
}

// device builders for the 5-/7-point operators
template <typename V>
bool device_build(cusp::dia_matrix<int, V, cusp::device_memory> &A, int st, size_t nx, size_t ny, size_t nz) {
  const size_t rows = nx * ny * nz;
  const size_t nnz = (size_t)b200sp_poisson_num_entries(st, (int64_t)nx, (int64_t)ny, (int64_t)nz, 0, (int64_t)rows);
  A.resize(rows, rows, nnz, (size_t)st);
  check(pdia(st, (int64_t)nx, (int64_t)ny, (int64_t)nz, (int64_t)rows, (int64_t)A.values.pitch,
             raw_ptr(A.diagonal_offsets), raw_ptr(A.values.values)));
  return true;
}
template <typename V>
bool device_build(cusp::ell_matrix<int, V, cusp::device_memory> &A, int st, size_t nx, size_t ny, size_t nz) {
  const size_t rows = nx * ny * nz;
  const size_t nnz = (size_t)b200sp_poisson_num_entries(st, (int64_t)nx, (int64_t)ny, (int64_t)nz, 0, (int64_t)rows);
  A.resize(rows, rows, nnz, (size_t)st);  // DIA -> ELL keeps K = #diagonals and the DIA pitch (= rows)
  check(pell(st, (int64_t)nx, (int64_t)ny, (int64_t)nz, (int64_t)rows, (int64_t)A.values.pitch,
             raw_ptr(A.column_indices.values), raw_ptr(A.values.values)));
  return true;
}
template <typename V>
bool device_build(cusp::csr_matrix<int, V, cusp::device_memory> &A, int st, size_t nx, size_t ny, size_t nz) {
  const size_t rows = nx * ny * nz;
  const size_t nnz = (size_t)b200sp_poisson_num_entries(st, (int64_t)nx, (int64_t)ny, (int64_t)nz, 0, (int64_t)rows);
  A.resize(rows, rows, nnz);
  check(b200sp_poisson_csr_offsets(engine(), current_stream(), st, (int64_t)nx, (int64_t)ny, (int64_t)nz, 0,
                                   (int64_t)rows, raw_ptr(A.row_offsets)));
  check(pcsr(st, (int64_t)nx, (int64_t)ny, (int64_t)nz, (int64_t)rows, raw_ptr(A.row_offsets),
             raw_ptr(A.column_indices), raw_ptr(A.values)));
  return true;
}
template <typename M>
bool device_build(M &, int, size_t, size_t, size_t) {
  return false;
}
template <typename M>
struct device_buildable : std::false_type {};
template <typename V>
struct device_buildable<cusp::dia_matrix<int, V, cusp::device_memory>> : cusp::detail::is_abi_value<V> {};
template <typename V>
struct device_buildable<cusp::ell_matrix<int, V, cusp::device_memory>> : cusp::detail::is_abi_value<V> {};
template <typename V>
struct device_buildable<cusp::csr_matrix<int, V, cusp::device_memory>> : cusp::detail::is_abi_value<V> {};

template <typename M>
void build(M &matrix, const std::vector<stencil_point> &stencil, const std::array<size_t, 3> &grid, int dims,
           int abi_stencil, std::true_type) {
  if (abi_stencil && grid[0] * grid[1] * grid[2] > 0) {
    device_build(matrix, abi_stencil, grid[0], grid[1], grid[2]);
    return;
  }
  from_stencil_host_route(matrix, stencil, grid, dims);
}
template <typename M>
void build(M &matrix, const std::vector<stencil_point> &stencil, const std::array<size_t, 3> &grid, int dims, int,
           std::false_type) {
  from_stencil_host_route(matrix, stencil, grid, dims);
}

}  // namespace detail

template <typename MatrixType>
void poisson5pt(MatrixType &matrix, const size_t m, const size_t n) {
  const std::vector<detail::stencil_point> st = {
      {{0, -1, 0}, -1}, {{-1, 0, 0}, -1}, {{0, 0, 0}, 4}, {{1, 0, 0}, -1}, {{0, 1, 0}, -1}};
  detail::build(matrix, st, {m, n, 1}, 2, 5, detail::device_buildable<MatrixType>());
}

template <typename MatrixType>
void poisson9pt(MatrixType &matrix, const size_t m, const size_t n) {
  std::vector<detail::stencil_point> st;
  for (long long j = -1; j <= 1; ++j)
    for (long long i = -1; i <= 1; ++i) st.push_back({{i, j, 0}, (i == 0 && j == 0) ? 8.0 : -1.0});
  detail::build(matrix, st, {m, n, 1}, 2, 0, std::false_type());
}

template <typename MatrixType>
void poisson7pt(MatrixType &matrix, const size_t m, const size_t n, const size_t k) {
  const std::vector<detail::stencil_point> st = {{{0, 0, -1}, -1}, {{0, -1, 0}, -1}, {{-1, 0, 0}, -1}, {{0, 0, 0}, 6},
                                                 {{1, 0, 0}, -1},  {{0, 1, 0}, -1},  {{0, 0, 1}, -1}};
  detail::build(matrix, st, {m, n, k}, 3, 7, detail::device_buildable<MatrixType>());
}

template <typename MatrixType>
void poisson27pt(MatrixType &matrix, const size_t m, const size_t n, const size_t l) {
  std::vector<detail::stencil_point> st;
  for (long long k = -1; k <= 1; ++k)
    for (long long j = -1; j <= 1; ++j)
      for (long long i = -1; i <= 1; ++i) st.push_back({{i, j, k}, (i == 0 && j == 0 && k == 0) ? 26.0 : -1.0});
  detail::build(matrix, st, {m, n, l}, 3, 0, std::false_type());
}

}  // namespace gallery
}  // namespace cusp
